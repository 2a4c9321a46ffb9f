// K4 -- fused positive/negative edge scoring + BCE-with-logits, forward and (user side) backward.
//
// Replaces train_gnn.py:259-281 (four [E,H] gathers, two mul+rowsum, two BCEWithLogitsLoss) and
// the index_put_(accumulate=True) half of loss.backward() (train_gnn.py:283) that lands on
// user_emb.  Positive edges arrive grouped by user (CSR keyed on pos_u), so the user row is
// read ONCE per user instead of twice per edge, dloss/du is produced in the same pass without
// atomics, and per-edge coefficients c_pos/c_neg are left for the post-side gather passes
// (trg_gather_wsum).  HBM-bound: 2 random post rows per edge.
#include "common.cuh"

namespace trg {
namespace {

constexpr int kThreads = 256;

struct BceArgs {
  const int* rowptr;
  const int* col_p;
  const int* eid;
  const long long* neg_p;
  const void* u;
  const void* p;
  const float* wbar;
  float* c_pos;
  float* c_neg;
  void* g_u;
  double* partials;  // [gridDim.x][2]
  int64_t n_users;
  float inv_e;
  int row_vecs;
};

__device__ __forceinline__ float softplus(float x) {  // log(1 + e^x), stable
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoid(float x) {
  const float e = expf(-fabsf(x));
  return x >= 0.f ? __fdiv_rn(1.f, 1.f + e) : __fdiv_rn(e, 1.f + e);
}

template <typename T, int LPR, int VPL>
__global__ void __launch_bounds__(kThreads) edge_bce(const BceArgs a) {
  constexpr int kVec = Elem<T>::kVec;
  constexpr int kUnroll = VPL == 1 ? 4 : 2;
  __shared__ double red[2][kThreads / 32];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPR - 1);
  const unsigned gmask = LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << (lane & ~(LPR - 1)));
  const int64_t row = (int64_t)blockIdx.x * (kThreads / LPR) + threadIdx.x / LPR;
  const bool want_grad = a.g_u != nullptr;
  float sp_sum = 0.f, sn_sum = 0.f;

  if (row < a.n_users) {
    const int beg = ldg_stream(a.rowptr + row);
    const int end = ldg_stream(a.rowptr + row + 1);
    const size_t row_bytes = (size_t)a.row_vecs * 16;
    const char* pb = reinterpret_cast<const char*>(a.p);
    const float wbar = __ldg(a.wbar);

    bool act[VPL];
    float uf[VPL][kVec], ga[VPL][kVec];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      act[i] = gl + i * LPR < a.row_vecs;
#pragma unroll
      for (int k = 0; k < kVec; ++k) uf[i][k] = ga[i][k] = 0.f;
      if (act[i] && end > beg)
        Elem<T>::unpack(ldg_row(reinterpret_cast<const char*>(a.u) + (size_t)row * row_bytes +
                                (size_t)(gl + i * LPR) * 16),
                        uf[i]);
    }

    for (int j = beg; j < end; j += LPR) {
      const int my = j + gl;
      int cp = 0, cn = 0, e = 0;
      if (my < end) {
        cp = ldg_stream(a.col_p + my);
        e = ldg_stream(a.eid + my);
        cn = (int)ldg_stream(a.neg_p + e);
      }
      float my_cpos = 0.f, my_cneg = 0.f;
      const int cnt = min(LPR, end - j);
      for (int t = 0; t < cnt; t += kUnroll) {
        uint4 vp[kUnroll][VPL], vn[kUnroll][VPL];
#pragma unroll
        for (int q = 0; q < kUnroll; ++q) {
          const int cpq = __shfl_sync(gmask, cp, t + q, LPR);
          const int cnq = __shfl_sync(gmask, cn, t + q, LPR);
          if (t + q < cnt) {
#pragma unroll
            for (int i = 0; i < VPL; ++i)
              if (act[i]) {
                vp[q][i] = ldg_row(pb + (size_t)cpq * row_bytes + (size_t)(gl + i * LPR) * 16);
                vn[q][i] = ldg_row(pb + (size_t)cnq * row_bytes + (size_t)(gl + i * LPR) * 16);
              }
          }
        }
#pragma unroll
        for (int q = 0; q < kUnroll; ++q) {
          if (t + q < cnt) {  // group-uniform
            float fp[VPL][kVec], fn[VPL][kVec];
            float dp = 0.f, dn = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
              if (act[i]) {
                Elem<T>::unpack(vp[q][i], fp[i]);
                Elem<T>::unpack(vn[q][i], fn[i]);
#pragma unroll
                for (int k = 0; k < kVec; ++k) {
                  dp = fmaf(uf[i][k], fp[i][k], dp);
                  dn = fmaf(uf[i][k], fn[i][k], dn);
                }
              }
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) {
              dp += __shfl_xor_sync(gmask, dp, o, LPR);
              dn += __shfl_xor_sync(gmask, dn, o, LPR);
            }
            sp_sum += softplus(-dp);  // BCEWithLogits(x, 1) = softplus(-x)
            sn_sum += softplus(dn);   // BCEWithLogits(x, 0) = softplus(x)
            if (want_grad) {
              const float cpos = -wbar * sigmoid(-dp) * a.inv_e;
              const float cneg = sigmoid(dn) * a.inv_e;
              if (gl == t + q) {
                my_cpos = cpos;
                my_cneg = cneg;
              }
#pragma unroll
              for (int i = 0; i < VPL; ++i)
                if (act[i])
#pragma unroll
                  for (int k = 0; k < kVec; ++k)
                    ga[i][k] = fmaf(cpos, fp[i][k], fmaf(cneg, fn[i][k], ga[i][k]));
            }
          }
        }
      }
      if (want_grad && my < end) {
        a.c_pos[e] = my_cpos;
        a.c_neg[e] = my_cneg;
      }
    }
    if (want_grad) {
      char* ob = reinterpret_cast<char*>(a.g_u) + (size_t)row * row_bytes;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (act[i]) stg_stream(ob + (size_t)(gl + i * LPR) * 16, Elem<T>::pack(ga[i]));
    }
    if (gl != 0) sp_sum = sn_sum = 0.f;  // every lane of a group holds the same sums
  }

  // deterministic CTA reduction -> one (sp, sn) pair of doubles per CTA
  double dsp = sp_sum, dsn = sn_sum;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dsp += __shfl_xor_sync(0xffffffffu, dsp, o);
    dsn += __shfl_xor_sync(0xffffffffu, dsn, o);
  }
  if (lane == 0) {
    red[0][threadIdx.x >> 5] = dsp;
    red[1][threadIdx.x >> 5] = dsn;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0, s1 = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      s0 += red[0][w];
      s1 += red[1][w];
    }
    a.partials[2 * (size_t)blockIdx.x] = s0;
    a.partials[2 * (size_t)blockIdx.x + 1] = s1;
  }
}

// Fixed-order final reduction: loss = wbar * mean(softplus(-pos)) + mean(softplus(neg)).
__global__ void __launch_bounds__(1024) edge_bce_finish(const double* __restrict__ partials,
                                                        int64_t n_parts, const float* wbar,
                                                        double n_edges, float* loss_out) {
  __shared__ double red[2][32];
  double s0 = 0, s1 = 0;
  for (int64_t i = threadIdx.x; i < n_parts; i += 1024) {
    s0 += partials[2 * i];
    s1 += partials[2 * i + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s0;
    red[1][threadIdx.x >> 5] = s1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s0 = s1 = 0;
    for (int w = 0; w < 32; ++w) {
      s0 += red[0][w];
      s1 += red[1][w];
    }
    const float pos_loss = (float)(s0 / n_edges);  // 0/0 = NaN for E = 0, like torch's mean
    const float neg_loss = (float)(s1 / n_edges);
    loss_out[0] = __ldg(wbar) * pos_loss + neg_loss;
  }
}

template <typename T>
int launch_bce(BceArgs& a, int64_t* n_blocks_out, cudaStream_t st, bool dry) {
  const int rv = a.row_vecs;
#define TRG_BCE_CASE(LPR, VPL)                                                   \
  {                                                                              \
    const int64_t grid = ceil_div<int64_t>(a.n_users, kThreads / LPR);           \
    *n_blocks_out = grid;                                                        \
    if (!dry) edge_bce<T, LPR, VPL><<<(unsigned)grid, kThreads, 0, st>>>(a);     \
  }
  if (rv <= 1) TRG_BCE_CASE(1, 1)
  else if (rv <= 2) TRG_BCE_CASE(2, 1)
  else if (rv <= 4) TRG_BCE_CASE(4, 1)
  else if (rv <= 8) TRG_BCE_CASE(8, 1)
  else if (rv <= 16) TRG_BCE_CASE(16, 1)
  else if (rv <= 32) TRG_BCE_CASE(32, 1)
  else if (rv <= 64) TRG_BCE_CASE(32, 2)
  else {
    set_error("trg_edge_bce_fwd: rows wider than 1024 bytes are not supported (row_vecs=%d)", rv);
    return TRG_E_UNSUPPORTED;
  }
#undef TRG_BCE_CASE
  return TRG_OK;
}

}  // namespace
}  // namespace trg

using namespace trg;

extern "C" size_t trg_edge_bce_workspace_bytes(int64_t n_users) {
  if (n_users < 0) return 0;
  // worst case: one warp per user -> 8 users per CTA, 2 doubles per CTA
  return align_up((size_t)(ceil_div<int64_t>(n_users, 8) + 1) * 16, 256);
}

extern "C" int trg_edge_bce_fwd(const int32_t* rowptr_u, const int32_t* col_p, const int32_t* eid,
                                const int64_t* neg_p, const void* u, const void* p, int64_t n_users,
                                int64_t n_edges, int32_t hidden, int dtype, const float* wbar,
                                float* loss_out, float* c_pos, float* c_neg, void* g_u,
                                void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TRG_CHECK_ARG(n_users >= 0 && n_edges >= 0, "trg_edge_bce_fwd: negative size");
  TRG_CHECK_ARG(loss_out && wbar, "trg_edge_bce_fwd: NULL loss_out/wbar");
  TRG_CHECK_ARG((c_pos == nullptr) == (g_u == nullptr) && (c_neg == nullptr) == (g_u == nullptr),
                "trg_edge_bce_fwd: c_pos, c_neg and g_u must be all NULL or all non-NULL");
  TRG_CHECK_ARG(n_users == 0 || (rowptr_u && u && p), "trg_edge_bce_fwd: NULL rowptr / tables");
  TRG_CHECK_ARG(((uintptr_t)u | (uintptr_t)p | (uintptr_t)g_u) % 16 == 0,
                "trg_edge_bce_fwd: tables must be 16-byte aligned");
  const int es = dtype == TRG_BF16 ? 2 : 4;
  TRG_CHECK_ARG(dtype == TRG_F32 || dtype == TRG_BF16, "trg_edge_bce_fwd: unknown dtype %d", dtype);
  TRG_CHECK_ARG(hidden > 0 && (hidden * es) % 16 == 0,
                "trg_edge_bce_fwd: row width %d x %d bytes is not a multiple of 16 bytes", hidden, es);
  if (workspace == nullptr || workspace_bytes < trg_edge_bce_workspace_bytes(n_users)) {
    set_error("trg_edge_bce_fwd: workspace %zu < required %zu", workspace_bytes,
              trg_edge_bce_workspace_bytes(n_users));
    return TRG_E_WORKSPACE;
  }
  BceArgs a{};
  a.rowptr = rowptr_u; a.col_p = col_p; a.eid = eid; a.neg_p = (const long long*)neg_p;
  a.u = u; a.p = p; a.wbar = wbar; a.c_pos = c_pos; a.c_neg = c_neg; a.g_u = g_u;
  a.partials = reinterpret_cast<double*>(workspace);
  a.n_users = n_users;
  a.inv_e = n_edges > 0 ? (float)(1.0 / (double)n_edges) : 0.f;
  a.row_vecs = hidden * es / 16;
  int64_t n_blocks = 0;
  if (n_users > 0) {
    int rc = dtype == TRG_F32 ? launch_bce<float>(a, &n_blocks, st, false)
                              : launch_bce<__nv_bfloat16>(a, &n_blocks, st, false);
    if (rc) return rc;
    count_launch();
    TRG_LAUNCH_OK();
  }
  edge_bce_finish<<<1, 1024, 0, st>>>(a.partials, n_blocks, wbar, (double)n_edges, loss_out);
  count_launch();
  TRG_LAUNCH_OK();
  return TRG_OK;
}
