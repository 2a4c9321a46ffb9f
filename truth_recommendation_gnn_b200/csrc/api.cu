// Library-level entry points of include/trg_b200.h: error string, launch counter, and the
// dispatcher for the projection kernels.
#include <cstdarg>

#include "common.cuh"

namespace trg {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int proj_simt_launch(const trg_proj_term* terms, int n_terms, const float* bias, int64_t n_rows,
                     int hidden, int dtype, int relu, void* out, cudaStream_t st);

}  // namespace trg

using namespace trg;

extern "C" int trg_abi_version(void) { return TRG_ABI_VERSION; }
extern "C" const char* trg_last_error(void) { return g_err; }
extern "C" int64_t trg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int trg_sage_proj_fwd(const trg_proj_term* terms, int32_t n_terms, const float* bias,
                                 int64_t n_rows, int32_t hidden, int dtype, int relu, void* out,
                                 void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TRG_CHECK_ARG(terms && n_terms >= 1 && n_terms <= 4, "trg_sage_proj_fwd: n_terms=%d not in 1..4", n_terms);
  TRG_CHECK_ARG(dtype == TRG_F32 || dtype == TRG_BF16, "trg_sage_proj_fwd: unknown dtype %d", dtype);
  TRG_CHECK_ARG(n_rows >= 0 && hidden > 0 && hidden % 4 == 0, "trg_sage_proj_fwd: bad n_rows/hidden");
  if (n_rows == 0) return TRG_OK;
  const int es = dtype == TRG_BF16 ? 2 : 4;
  TRG_CHECK_ARG(out && (uintptr_t)out % 16 == 0 && (hidden * es) % 16 == 0,
                "trg_sage_proj_fwd: out must be 16-byte aligned with 16-byte-multiple rows");
  for (int i = 0; i < n_terms; ++i) {
    TRG_CHECK_ARG(terms[i].a && terms[i].w && terms[i].k > 0 && (terms[i].k * es) % 16 == 0 &&
                      (uintptr_t)terms[i].a % 16 == 0 && (uintptr_t)terms[i].w % 16 == 0,
                  "trg_sage_proj_fwd: term %d needs 16-byte aligned A/W and 16-byte-multiple rows", i);
  }
  return proj_simt_launch(terms, n_terms, bias, n_rows, hidden, dtype, relu, out, st);
}
