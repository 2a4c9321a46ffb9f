// K1 / K2 -- fused neighbour gather + segmented reduction over a destination-sorted CSR.
//
//   trg_sage_agg_fwd : mean[r] = (sum_{j in row r} x[col[j]]) / max(deg r, 1)
//                      replaces x_src.index_select(0, src) + scatter(reduce='mean') inside the
//                      SAGEConv calls at train_gnn.py:177-184,194-197 (never materialises [E,F]).
//   trg_sage_agg_bwd : g_src[s] = sum_{j in row s of the transposed CSR} g_mean[c] * inv_deg[c]
//                      replaces autograd's index_add/gather; atomic-free.
//   trg_gather_wsum  : out[r] (+)= scale * sum_j coef[eid[j]] * x[col[j]]  (loss backward).
//
// HBM-bound: one group of LPR lanes owns one destination row; each lane moves 16-byte vectors
// of the gathered source rows (coalesced across the group: a 512 B fp32 H=128 row is one
// LDG.128 per lane of a full warp) with up to 8 row loads in flight per lane; neighbour ids are
// read coalesced (one per lane) and broadcast by shuffle.  Neighbours are added in CSR order,
// which is edge order inside a row (stable sort), so fp32 sums equal the CPU scatter_add_ bit
// for bit.
#include <type_traits>

#include "common.cuh"

namespace trg {
namespace {

constexpr int kThreads = 256;

enum Mode { kMean = 0, kNbrScale = 1, kEdgeCoef = 2 };

struct GatherArgs {
  const int* rowptr;
  const int* col;
  const int* eid;          // kEdgeCoef
  const float* coef;       // kEdgeCoef: indexed by eid
  const float* nbr_scale;  // kNbrScale: indexed by col (nullable)
  const float* scale;      // device scalar, nullable
  const void* x;
  void* out;
  float* inv_deg_out;  // kMean, nullable
  int64_t n_rows;
  int row_vecs;  // 16-byte vectors per row
  int accumulate;
  const void* relu_of;  // nullable, non-mean modes: [n_rows, feat] forward activation; out = relu_of > 0 ? out : 0
  // long-row splitting (nullable vinfo): rowptr/n_rows then describe VIRTUAL rows; a virtual row is
  // either a whole short row (vinfo >= 0: its row id) or one <= T-edge slice of a long row
  // (vinfo < 0: partial-sum slot -(vinfo+1)); slices are summed in order by gather_combine_long
  const int* vinfo;
  float* partial;        // [n_slots][feat] fp32
  const int* long_rows;  // [n_long] original row ids
  const int* long_ptr;   // [n_long + 1] slot ranges
  const int* rowptr_orig;
  int64_t n_long;
};

// fp32 output rows for a table stored in T (OUTF32; non-mean modes): partial sums that are reduced across
// GPUs travel in fp32 so that a bf16 model rounds once, after the cross-rank add.  Lane-local layout: the
// kVec elements a lane holds land at byte offset (lane offset in the T row) * 4 / sizeof(T).
template <int N>
__device__ __forceinline__ void load_f32_plain(const char* p, float* f) {
#pragma unroll
  for (int k = 0; k < N; k += (N >= 4 ? 4 : 2)) {
    if (N >= 4) {
      const float4 v = *reinterpret_cast<const float4*>(p + k * 4);
      f[k] = v.x; f[k + 1] = v.y; f[k + 2] = v.z; f[k + 3] = v.w;
    } else {
      const float2 v = *reinterpret_cast<const float2*>(p + k * 4);
      f[k] = v.x; f[k + 1] = v.y;
    }
  }
}
template <int N>
__device__ __forceinline__ void store_f32_stream(char* p, const float* f) {
#pragma unroll
  for (int k = 0; k < N; k += (N >= 4 ? 4 : 2)) {
    if (N >= 4)
      stg_stream(p + k * 4, make_uint4(__float_as_uint(f[k]), __float_as_uint(f[k + 1]), __float_as_uint(f[k + 2]),
                                       __float_as_uint(f[k + 3])));
    else
      asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p + k * 4), "r"(__float_as_uint(f[k])),
                   "r"(__float_as_uint(f[k + 1])) : "memory");
  }
}

template <typename T, int LPR, int VPL, int MODE, bool OUTF32 = false>
__global__ void __launch_bounds__(kThreads) gather_reduce(const GatherArgs a) {
  constexpr int kVec = Elem<T>::kVec;
  constexpr int kOutMul = OUTF32 ? 4 / (int)sizeof(T) : 1;   // output bytes per input byte
  constexpr int kUnroll = VPL == 1 ? 8 : (VPL == 2 ? 4 : 2);
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPR - 1);
  const unsigned gmask = LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << (lane & ~(LPR - 1)));
  const int64_t row = (int64_t)blockIdx.x * (kThreads / LPR) + threadIdx.x / LPR;
  if (row >= a.n_rows) return;

  const int beg = ldg_stream(a.rowptr + row);
  const int end = ldg_stream(a.rowptr + row + 1);
  const size_t row_bytes = (size_t)a.row_vecs * 16;
  const char* xb = reinterpret_cast<const char*>(a.x);

  bool act[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) act[i] = gl + i * LPR < a.row_vecs;

  float acc[VPL][kVec];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int k = 0; k < kVec; ++k) acc[i][k] = 0.f;

  for (int j = beg; j < end; j += LPR) {
    const int my = j + gl;
    int c = 0;
    float wgt = 1.f;
    if (my < end) {
      c = ldg_stream(a.col + my);
      if (MODE == kNbrScale) {
        if (a.nbr_scale) wgt = __ldg(a.nbr_scale + c);
      } else if (MODE == kEdgeCoef) {
        wgt = __ldg(a.coef + ldg_stream(a.eid + my));
      }
    }
    const int cnt = min(LPR, end - j);
    for (int t = 0; t < cnt; t += kUnroll) {
      uint4 v[kUnroll][VPL];
      float wv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        // shuffles are executed by the whole group (cnt is group-uniform)
        const int cu = __shfl_sync(gmask, c, t + u, LPR);
        if (MODE != kMean) wv[u] = __shfl_sync(gmask, wgt, t + u, LPR);
        if (t + u < cnt) {
          const char* rp = xb + (size_t)cu * row_bytes + (size_t)gl * 16;
#pragma unroll
          for (int i = 0; i < VPL; ++i)
            if (act[i]) v[u][i] = ldg_row(rp + (size_t)i * LPR * 16);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (t + u < cnt) {
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            if (act[i]) {
              float f[kVec];
              Elem<T>::unpack(v[u][i], f);
#pragma unroll
              for (int k = 0; k < kVec; ++k) {
                if (MODE == kMean)
                  acc[i][k] = __fadd_rn(acc[i][k], f[k]);  // plain adds, CSR order: bit-exact
                else
                  acc[i][k] = fmaf(wv[u], f[k], acc[i][k]);
              }
            }
          }
        }
      }
    }
  }

  float post = 1.f;
  if (MODE == kMean) {
    const float cntf = (float)max(end - beg, 1);
    if (gl == 0 && a.inv_deg_out) a.inv_deg_out[row] = __fdiv_rn(1.f, cntf);
#pragma unroll
    for (int i = 0; i < VPL; ++i)
#pragma unroll
      for (int k = 0; k < kVec; ++k) acc[i][k] = __fdiv_rn(acc[i][k], cntf);  // sum / count (IEEE)
  } else if (a.scale) {
    post = __ldg(a.scale);
  }
  char* ob = reinterpret_cast<char*>(a.out) + ((size_t)row * row_bytes + (size_t)gl * 16) * kOutMul;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (!act[i]) continue;
    if (MODE != kMean) {
#pragma unroll
      for (int k = 0; k < kVec; ++k) acc[i][k] *= post;
      if (a.accumulate) {
        float f[kVec];
        if (OUTF32) load_f32_plain<kVec>(ob + (size_t)i * LPR * 16 * kOutMul, f);
        else Elem<T>::unpack(*reinterpret_cast<const uint4*>(ob + (size_t)i * LPR * 16), f);
#pragma unroll
        for (int k = 0; k < kVec; ++k) acc[i][k] += f[k];
      }
      if (a.relu_of) {
        float f[kVec];
        Elem<T>::unpack(ldg_row(reinterpret_cast<const char*>(a.relu_of) + (size_t)row * row_bytes +
                                (size_t)(gl + i * LPR) * 16), f);
#pragma unroll
        for (int k = 0; k < kVec; ++k) acc[i][k] = f[k] > 0.f ? acc[i][k] : 0.f;
      }
    }
    if (OUTF32) store_f32_stream<kVec>(ob + (size_t)i * LPR * 16 * kOutMul, acc[i]);
    else stg_stream(ob + (size_t)i * LPR * 16, Elem<T>::pack(acc[i]));
  }
}

// v2: each group of LPR lanes owns R CONSECUTIVE destination rows and walks their edges as one
// continuous stream (the CSR stores them back to back): neighbour ids for the next 32-edge chunk
// are prefetched while the current chunk's rows are in flight, and the running sum is flushed at
// row boundaries.  Removes the per-row rowptr -> col -> row dependency chain that made short rows
// (in-degree ~8) latency-bound (profiles/README.md, r1_v1).  Same CSR-order adds: still bit-exact.
// UN (0 = default): row loads in flight per lane.  Measured on 256-byte rows (bf16 H = 128, 8-byte lanes): 16 in
// flight at 3 CTAs / SM is SLOWER than 8 at 4 CTAs / SM (agg_fwd 15.5 vs 12.0 ms at config 3): that shape is not
// short of loads in flight (profiles/README.md, r2).
template <typename T, int LPR, int VPL, int MODE, int R, bool VIRT, int VB = 16, bool OUTF32 = false, int UN = 0>
__global__ void __launch_bounds__(kThreads, UN > 8 ? 3 : (VPL == 1 ? 4 : 2)) gather_reduce_seg(const GatherArgs a) {
  using V = Vec<T, VB>;
  constexpr int kVec = V::kVec;
  constexpr int kOutMul = OUTF32 ? 4 / (int)sizeof(T) : 1;
  constexpr int kUnroll = UN > 0 ? UN : (VPL == 1 ? 8 : (VPL == 2 ? 4 : 2));
  static_assert(R < LPR, "row boundaries are held one per lane");
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LPR - 1);
  const unsigned gmask = LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << (lane & ~(LPR - 1)));
  const int64_t r0 = ((int64_t)blockIdx.x * (kThreads / LPR) + threadIdx.x / LPR) * R;
  if (r0 >= a.n_rows) return;
  const int nr = (int)((a.n_rows - r0) < (int64_t)R ? (a.n_rows - r0) : (int64_t)R);
  const int my_ptr = ldg_stream(a.rowptr + r0 + min(gl, nr));   // lane l: start of row l (l <= nr)
  const int e_end = __shfl_sync(gmask, my_ptr, nr, LPR);
  int e0 = __shfl_sync(gmask, my_ptr, 0, LPR);
  const size_t row_bytes = (size_t)a.row_vecs * VB;
  const char* xb = reinterpret_cast<const char*>(a.x);
  const float post = (MODE != kMean && a.scale) ? __ldg(a.scale) : 1.f;

  bool act[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) act[i] = gl + i * LPR < a.row_vecs;
  float acc[VPL][kVec];
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int k = 0; k < kVec; ++k) acc[i][k] = 0.f;

  int cur = 0;                                            // row being accumulated (group-uniform)
  int cur_beg = e0;
  int cur_end = __shfl_sync(gmask, my_ptr, 1, LPR);

  auto flush = [&]() {
    int64_t orow = r0 + cur;
    if (VIRT) {
      const int info = __ldg(a.vinfo + r0 + cur);
      if (info < 0) {   // slice of a long row: raw fp32 partial sum, combined later in slice order
        float* pb = a.partial + (size_t)(-(info + 1)) * ((size_t)a.row_vecs * kVec);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
          if (act[i]) {
#pragma unroll
            for (int k = 0; k < kVec; ++k) pb[(size_t)(gl + i * LPR) * kVec + k] = acc[i][k];
          }
#pragma unroll
          for (int k = 0; k < kVec; ++k) acc[i][k] = 0.f;
        }
        ++cur;
        cur_beg = cur_end;
        cur_end = __shfl_sync(gmask, my_ptr, min(cur + 1, nr), LPR);
        return;
      }
      orow = info;
    }
    char* ob = reinterpret_cast<char*>(a.out) + ((size_t)orow * row_bytes + (size_t)gl * VB) * kOutMul;
    if (MODE == kMean) {
      const float cntf = (float)max(cur_end - cur_beg, 1);
      if (gl == 0 && a.inv_deg_out) a.inv_deg_out[orow] = __fdiv_rn(1.f, cntf);
#pragma unroll
      for (int i = 0; i < VPL; ++i)
#pragma unroll
        for (int k = 0; k < kVec; ++k) acc[i][k] = __fdiv_rn(acc[i][k], cntf);   // sum / count (IEEE)
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (act[i]) {
        if (MODE != kMean) {
#pragma unroll
          for (int k = 0; k < kVec; ++k) acc[i][k] *= post;
          if (a.accumulate) {
            float f[kVec];
            if (OUTF32) load_f32_plain<kVec>(ob + (size_t)i * LPR * VB * kOutMul, f);
            else V::unpack(V::load_plain(ob + (size_t)i * LPR * VB), f);
#pragma unroll
            for (int k = 0; k < kVec; ++k) acc[i][k] += f[k];
          }
          if (a.relu_of) {   // ReLU derivative of the layer that produced the rows this gradient is for
            float f[kVec];
            V::unpack(V::load(reinterpret_cast<const char*>(a.relu_of) + (size_t)orow * row_bytes +
                              (size_t)(gl + i * LPR) * VB), f);
#pragma unroll
            for (int k = 0; k < kVec; ++k) acc[i][k] = f[k] > 0.f ? acc[i][k] : 0.f;
          }
        }
        if (OUTF32) store_f32_stream<kVec>(ob + (size_t)i * LPR * VB * kOutMul, acc[i]);
        else V::store(ob + (size_t)i * LPR * VB, acc[i]);
      }
#pragma unroll
      for (int k = 0; k < kVec; ++k) acc[i][k] = 0.f;
    }
    ++cur;
    cur_beg = cur_end;
    cur_end = __shfl_sync(gmask, my_ptr, min(cur + 1, nr), LPR);
  };

  // index pipeline: c / wgt of the current chunk are ready; the next chunk's are in flight
  auto load_idx = [&](int e, int& c, int& aux) {
    c = 0; aux = 0;
    if (e < e_end) {
      c = ldg_stream(a.col + e);
      if (MODE == kEdgeCoef) aux = ldg_stream(a.eid + e);
    }
  };
  auto load_wgt = [&](int e, int c, int aux) -> float {
    if (e >= e_end) return 0.f;
    if (MODE == kNbrScale) return a.nbr_scale ? __ldg(a.nbr_scale + c) : 1.f;
    if (MODE == kEdgeCoef) return __ldg(a.coef + aux);
    return 1.f;
  };
  int c_cur, aux_cur, c_nxt, aux_nxt;
  load_idx(e0 + gl, c_cur, aux_cur);
  float w_cur = load_wgt(e0 + gl, c_cur, aux_cur);
  load_idx(e0 + LPR + gl, c_nxt, aux_nxt);

  while (e0 < e_end) {
    const int cnt = min(LPR, e_end - e0);
    // issue the dependent weight lookup of the next chunk and the ids of the one after it
    const float w_nxt = (MODE == kMean) ? 1.f : load_wgt(e0 + LPR + gl, c_nxt, aux_nxt);
    int c_nn, aux_nn;
    load_idx(e0 + 2 * LPR + gl, c_nn, aux_nn);

    for (int t = 0; t < cnt; t += kUnroll) {
      typename V::Raw v[kUnroll][VPL];
      float wv[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int cu = __shfl_sync(gmask, c_cur, t + u, LPR);
        if (MODE != kMean) wv[u] = __shfl_sync(gmask, w_cur, t + u, LPR);
        if (t + u < cnt) {
          const char* rp = xb + (size_t)cu * row_bytes + (size_t)gl * VB;
#pragma unroll
          for (int i = 0; i < VPL; ++i)
            if (act[i]) v[u][i] = V::load(rp + (size_t)i * LPR * VB);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        if (t + u < cnt) {
          const int e = e0 + t + u;
          while (e >= cur_end) flush();                     // group-uniform: close finished rows
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            if (act[i]) {
              float f[kVec];
              V::unpack(v[u][i], f);
#pragma unroll
              for (int k = 0; k < kVec; ++k) {
                if (MODE == kMean)
                  acc[i][k] = __fadd_rn(acc[i][k], f[k]);   // plain adds, CSR order: bit-exact
                else
                  acc[i][k] = fmaf(wv[u], f[k], acc[i][k]);
              }
            }
          }
        }
      }
    }
    e0 += LPR;
    c_cur = c_nxt; aux_cur = aux_nxt; w_cur = w_nxt;
    c_nxt = c_nn; aux_nxt = aux_nn;
  }
  while (cur < nr) flush();                                 // last row and trailing empty rows
}

// Second stage for long rows: sum the slices' fp32 partials in slice order (deterministic), then the
// mode's epilogue (mean / scale / accumulate).  One CTA per long row.
template <typename T, int MODE, bool OUTF32 = false>
__global__ void __launch_bounds__(128) gather_combine_long(const GatherArgs a) {
  using TO = typename std::conditional<OUTF32, float, T>::type;
  const int i = blockIdx.x;
  const int64_t orow = a.long_rows[i];
  const int s0 = a.long_ptr[i], s1 = a.long_ptr[i + 1];
  const int feat = a.row_vecs * Elem<T>::kVec;
  const float post = (MODE != kMean && a.scale) ? __ldg(a.scale) : 1.f;
  const float cntf = (float)max(a.rowptr_orig[orow + 1] - a.rowptr_orig[orow], 1);
  if (MODE == kMean && threadIdx.x == 0 && a.inv_deg_out) a.inv_deg_out[orow] = __fdiv_rn(1.f, cntf);
  TO* out = reinterpret_cast<TO*>(a.out) + (size_t)orow * feat;
  for (int f = threadIdx.x; f < feat; f += blockDim.x) {
    float s = 0.f;
    for (int sl = s0; sl < s1; ++sl) s += a.partial[(size_t)sl * feat + f];
    if (MODE == kMean) {
      s = __fdiv_rn(s, cntf);
    } else {
      s *= post;
      if (a.accumulate) s += (float)out[f];
      if (a.relu_of && !((float)reinterpret_cast<const T*>(a.relu_of)[(size_t)orow * feat + f] > 0.f)) s = 0.f;
    }
    out[f] = (TO)s;
  }
}

template <typename T, int MODE, bool OUTF32 = false>
int launch_gather(const GatherArgs& a, cudaStream_t st) {
  if (a.n_rows == 0) return TRG_OK;
  const int rv = a.row_vecs;
#define TRG_GATHER_CASE(LPR, VPL)                                                         \
  {                                                                                       \
    const int64_t grid = ceil_div<int64_t>(a.n_rows, kThreads / LPR);                     \
    gather_reduce<T, LPR, VPL, MODE, OUTF32><<<(unsigned)grid, kThreads, 0, st>>>(a);             \
  }
#define TRG_GATHER_SEG(LPR, VPL, R)                                                       \
  {                                                                                       \
    const int64_t grid = ceil_div<int64_t>(a.n_rows, (int64_t)(kThreads / LPR) * R);      \
    if (a.vinfo)                                                                          \
      gather_reduce_seg<T, LPR, VPL, MODE, R, true, 16, OUTF32><<<(unsigned)grid, kThreads, 0, st>>>(a);  \
    else                                                                                  \
      gather_reduce_seg<T, LPR, VPL, MODE, R, false, 16, OUTF32><<<(unsigned)grid, kThreads, 0, st>>>(a); \
  }
#define TRG_GATHER_SEG8()                                                                 \
  {                                                                                       \
    GatherArgs b = a;                                                                     \
    b.row_vecs = a.row_vecs * 2;                                                          \
    const int64_t grid = ceil_div<int64_t>(b.n_rows, (int64_t)(kThreads / 32) * 8);       \
    if (b.vinfo)                                                                          \
      gather_reduce_seg<T, 32, 1, MODE, 8, true, 8, OUTF32><<<(unsigned)grid, kThreads, 0, st>>>(b);  \
    else                                                                                  \
      gather_reduce_seg<T, 32, 1, MODE, 8, false, 8, OUTF32><<<(unsigned)grid, kThreads, 0, st>>>(b); \
  }
  if (rv <= 1) TRG_GATHER_CASE(1, 1)
  else if (rv <= 2) TRG_GATHER_CASE(2, 1)
  else if (rv <= 4) TRG_GATHER_CASE(4, 1)
  else if (rv <= 8) TRG_GATHER_SEG(8, 1, 4)
  else if (rv == 16) TRG_GATHER_SEG8()          // 256-byte rows: one full warp per row, 8-byte lanes
  else if (rv <= 16) TRG_GATHER_SEG(16, 1, 8)
  else if (rv <= 32) TRG_GATHER_SEG(32, 1, 8)
  else if (rv <= 64) TRG_GATHER_SEG(32, 2, 8)
  else if (rv <= 128) TRG_GATHER_SEG(32, 4, 8)
  else {
    set_error("gather: rows wider than 2048 bytes are not supported (row_vecs=%d)", rv);
    return TRG_E_UNSUPPORTED;
  }
#undef TRG_GATHER_CASE
#undef TRG_GATHER_SEG
#undef TRG_GATHER_SEG8
  count_launch();
  TRG_LAUNCH_OK();
  if (a.vinfo && a.n_long > 0) {
    gather_combine_long<T, MODE, OUTF32><<<(unsigned)a.n_long, 128, 0, st>>>(a);
    count_launch();
    TRG_LAUNCH_OK();
  }
  return TRG_OK;
}

template <int MODE>
int dispatch(const GatherArgs& a, int dtype, int out_dtype, cudaStream_t st) {
  if (out_dtype != dtype && !(MODE != kMean && dtype == TRG_BF16 && out_dtype == TRG_F32)) {
    set_error("gather: out_dtype %d with dtype %d is not supported (fp32 output rows exist for bf16 sums only)",
              out_dtype, dtype);
    return TRG_E_UNSUPPORTED;
  }
  if (dtype == TRG_F32) return launch_gather<float, MODE>(a, st);
  if (dtype == TRG_BF16) {
    if constexpr (MODE != kMean) {
      if (out_dtype == TRG_F32) return launch_gather<__nv_bfloat16, MODE, true>(a, st);
    }
    return launch_gather<__nv_bfloat16, MODE>(a, st);
  }
  set_error("gather: unknown dtype %d", dtype);
  return TRG_E_ARG;
}

int apply_long(GatherArgs& a, const trg_long_rows* lr, const int32_t* rowptr, const char* who) {
  if (!lr || lr->n_long <= 0) return TRG_OK;
  if (a.row_vecs < 5) {
    set_error("%s: long-row splitting needs rows of at least 80 bytes", who);
    return TRG_E_UNSUPPORTED;
  }
  if (!lr->vrowptr || !lr->vinfo || !lr->long_rows || !lr->long_ptr || !lr->partial) {
    set_error("%s: incomplete trg_long_rows", who);
    return TRG_E_ARG;
  }
  a.rowptr_orig = rowptr;
  a.rowptr = lr->vrowptr; a.vinfo = lr->vinfo; a.n_rows = lr->n_vrows;
  a.partial = lr->partial; a.long_rows = lr->long_rows; a.long_ptr = lr->long_ptr; a.n_long = lr->n_long;
  return TRG_OK;
}

int row_vecs_of(int feat, int dtype, const char* who, int* out) {
  const int es = dtype == TRG_BF16 ? 2 : 4;
  if (feat <= 0 || (feat * es) % 16 != 0) {
    set_error("%s: row width %d x %d bytes is not a multiple of 16 bytes", who, feat, es);
    return TRG_E_ARG;
  }
  *out = feat * es / 16;
  return TRG_OK;
}

}  // namespace
}  // namespace trg

using namespace trg;

extern "C" int trg_sage_agg_fwd(const int32_t* rowptr, const int32_t* col, const void* x_src,
                                int64_t n_dst, int32_t feat, int dtype, void* mean_out,
                                float* inv_deg_out, const trg_long_rows* lr, void* stream) {
  TRG_CHECK_ARG(n_dst >= 0, "trg_sage_agg_fwd: n_dst < 0");
  if (n_dst == 0) return TRG_OK;
  TRG_CHECK_ARG(rowptr && mean_out, "trg_sage_agg_fwd: NULL rowptr/out");
  TRG_CHECK_ARG(((uintptr_t)x_src | (uintptr_t)mean_out) % 16 == 0, "trg_sage_agg_fwd: tables must be 16-byte aligned");
  GatherArgs a{};
  int rc = row_vecs_of(feat, dtype, "trg_sage_agg_fwd", &a.row_vecs);
  if (rc) return rc;
  a.rowptr = rowptr; a.col = col; a.x = x_src; a.out = mean_out; a.inv_deg_out = inv_deg_out;
  a.n_rows = n_dst;
  rc = apply_long(a, lr, rowptr, "trg_sage_agg_fwd");
  if (rc) return rc;
  return dispatch<kMean>(a, dtype, dtype, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int trg_sage_agg_bwd(const int32_t* rowptr_t, const int32_t* col_t, const float* inv_deg,
                                const void* g_mean, int64_t n_src, int32_t feat, int dtype,
                                void* g_src_out, int out_dtype, int accumulate, const void* relu_of,
                                const trg_long_rows* lr, void* stream) {
  TRG_CHECK_ARG(n_src >= 0, "trg_sage_agg_bwd: n_src < 0");
  if (n_src == 0) return TRG_OK;
  TRG_CHECK_ARG(rowptr_t && g_src_out, "trg_sage_agg_bwd: NULL rowptr/out");
  TRG_CHECK_ARG(((uintptr_t)g_mean | (uintptr_t)g_src_out | (uintptr_t)relu_of) % 16 == 0,
                "trg_sage_agg_bwd: tables must be 16-byte aligned");
  GatherArgs a{};
  int rc = row_vecs_of(feat, dtype, "trg_sage_agg_bwd", &a.row_vecs);
  if (rc) return rc;
  a.rowptr = rowptr_t; a.col = col_t; a.nbr_scale = inv_deg; a.x = g_mean; a.out = g_src_out;
  a.n_rows = n_src; a.accumulate = accumulate; a.relu_of = relu_of;
  rc = apply_long(a, lr, rowptr_t, "trg_sage_agg_bwd");
  if (rc) return rc;
  return dispatch<kNbrScale>(a, dtype, out_dtype, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int trg_gather_wsum(const int32_t* rowptr, const int32_t* col, const int32_t* eid,
                               const float* coef, const float* scale, const void* x, int64_t n_rows,
                               int32_t feat, int dtype, void* out, int out_dtype, int accumulate, const void* relu_of,
                               const trg_long_rows* lr, void* stream) {
  TRG_CHECK_ARG(n_rows >= 0, "trg_gather_wsum: n_rows < 0");
  if (n_rows == 0) return TRG_OK;
  TRG_CHECK_ARG(rowptr && out, "trg_gather_wsum: NULL rowptr/out");
  TRG_CHECK_ARG(((uintptr_t)x | (uintptr_t)out | (uintptr_t)relu_of) % 16 == 0,
                "trg_gather_wsum: tables must be 16-byte aligned");
  GatherArgs a{};
  int rc = row_vecs_of(feat, dtype, "trg_gather_wsum", &a.row_vecs);
  if (rc) return rc;
  a.rowptr = rowptr; a.col = col; a.eid = eid; a.coef = coef; a.scale = scale; a.x = x; a.out = out;
  a.n_rows = n_rows; a.accumulate = accumulate; a.relu_of = relu_of;
  rc = apply_long(a, lr, rowptr, "trg_gather_wsum");
  if (rc) return rc;
  return dispatch<kEdgeCoef>(a, dtype, out_dtype, reinterpret_cast<cudaStream_t>(stream));
}
