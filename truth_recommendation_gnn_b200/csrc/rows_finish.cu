// trg_rows_finish -- one pass over the owned rows of a table that was reduced across GPUs:
//     out[r, :] = gate( row_scale[r] * in[r, :] + add[r, :] )
// (1/deg of a source-partitioned mean; the local gradient term; the ReLU backward of train_gnn.py:187-198).
// Replaces the torch element-wise kernels the multi-GPU step launched after each reduce-scatter
// (`rs * inv_deg[:, None]`, `g_loc.add_(g)`, `threshold_backward`): HBM-bound, every byte touched once,
// 16-byte vectors, fp32 arithmetic, one rounding to the storage dtype.
#include "common.cuh"

namespace trg {
namespace {

template <typename TI, typename T>
__global__ void __launch_bounds__(256) rows_finish_kernel(const TI* __restrict__ in, const float* __restrict__ row_scale,
                                                          const T* add, const T* __restrict__ relu_of,
                                                          long long n_rows, int feat, T* out) {
  constexpr int kVec = Elem<T>::kVec;                 // elements per 16-byte vector of the storage dtype
  const int vecs_per_row = feat / kVec;
  const long long n_vec = n_rows * vecs_per_row;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n_vec;
       v += (long long)gridDim.x * blockDim.x) {
    const long long r = v / vecs_per_row;
    const size_t e0 = (size_t)v * kVec;
    float f[kVec];
    if (sizeof(TI) == 4) {
#pragma unroll
      for (int k = 0; k < kVec; k += 4) {
        const uint4 x = ldg_row(reinterpret_cast<const float*>(in) + e0 + k);
        f[k] = __uint_as_float(x.x); f[k + 1] = __uint_as_float(x.y);
        f[k + 2] = __uint_as_float(x.z); f[k + 3] = __uint_as_float(x.w);
      }
    } else {
      Elem<T>::unpack(ldg_row(reinterpret_cast<const T*>(in) + e0), f);
    }
    if (row_scale) {
      const float s = __ldg(row_scale + r);
#pragma unroll
      for (int k = 0; k < kVec; ++k) f[k] *= s;
    }
    if (add) {
      float g[kVec];
      Elem<T>::unpack(*reinterpret_cast<const uint4*>(add + e0), g);
#pragma unroll
      for (int k = 0; k < kVec; ++k) f[k] += g[k];
    }
    if (relu_of) {
      float g[kVec];
      Elem<T>::unpack(ldg_row(relu_of + e0), g);
#pragma unroll
      for (int k = 0; k < kVec; ++k) f[k] = g[k] > 0.f ? f[k] : 0.f;
    }
    *reinterpret_cast<uint4*>(out + e0) = Elem<T>::pack(f);
  }
}

}  // namespace
}  // namespace trg

using namespace trg;

extern "C" int trg_rows_finish(const void* in, int in_dtype, const float* row_scale, const void* add,
                               const void* relu_of, int64_t n_rows, int32_t feat, int dtype, void* out,
                               void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TRG_CHECK_ARG(n_rows >= 0 && feat > 0, "trg_rows_finish: bad n_rows/feat");
  TRG_CHECK_ARG(dtype == TRG_F32 || dtype == TRG_BF16, "trg_rows_finish: unknown dtype %d", dtype);
  TRG_CHECK_ARG(in_dtype == dtype || in_dtype == TRG_F32, "trg_rows_finish: in_dtype must be dtype or TRG_F32");
  if (n_rows == 0) return TRG_OK;
  const int es = dtype == TRG_BF16 ? 2 : 4;
  TRG_CHECK_ARG(in && out && (feat * es) % 16 == 0, "trg_rows_finish: NULL table or rows not a multiple of 16 bytes");
  TRG_CHECK_ARG(((uintptr_t)in | (uintptr_t)out | (uintptr_t)add | (uintptr_t)relu_of) % 16 == 0,
                "trg_rows_finish: tables must be 16-byte aligned");
  const long long n_vec = (long long)n_rows * (feat * es / 16);
  const int grid = (int)std::min<long long>((n_vec + 255) / 256, (long long)grid_sms() * 8);
  if (dtype == TRG_F32)
    rows_finish_kernel<float, float><<<grid, 256, 0, st>>>((const float*)in, row_scale, (const float*)add,
                                                           (const float*)relu_of, n_rows, feat, (float*)out);
  else if (in_dtype == TRG_F32)
    rows_finish_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(
        (const float*)in, row_scale, (const __nv_bfloat16*)add, (const __nv_bfloat16*)relu_of, n_rows, feat,
        (__nv_bfloat16*)out);
  else
    rows_finish_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(
        (const __nv_bfloat16*)in, row_scale, (const __nv_bfloat16*)add, (const __nv_bfloat16*)relu_of, n_rows,
        feat, (__nv_bfloat16*)out);
  count_launch();
  TRG_LAUNCH_OK();
  return TRG_OK;
}
