// trg_peer_reduce_rows -- reduce-scatter of a partial-sum table FUSED with the row finish, over peer memory:
//     out[r, :] = gate( row_scale[r] * sum_{g = 0 .. G-1} parts[g][row0 + r, :] + add[r, :] )
// Every rank holds a full-height [G * n_rows, feat] table of partial sums (the source-partitioned post -> user
// aggregation, the dL/du partials of the loss and of the layers' backward: SURVEY.md §8e) in memory its peers
// can address (NVLink / NVSwitch P2P mappings).  The owner of rows [row0, row0 + n_rows) reads those rows from
// all G tables with plain 16-byte loads -- its own from HBM, the other G - 1 straight over NVLink -- adds them
// in rank order (fixed, so the result is deterministic), applies 1/deg, the local gradient term and the ReLU
// backward of train_gnn.py:187-198, and writes the finished rows once in the storage dtype.  It replaces an
// NCCL reduce-scatter (a separate output table written and read again) plus trg_rows_finish: the reduced table
// never exists in memory, and the transfer needs no SM-resident library kernel of its own.
//
// NVLink-bound: (G - 1) / G of the bytes come over the links (~2.5 us away), so what matters is loads in
// flight: every thread issues the G loads of a vector (and of the next UN - 1 vectors) before it adds
// anything.  The grid is a multiple of the SM count with small CTAs, so that the kernel spreads thinly over
// all SMs next to the compute kernel it overlaps with.
#include "common.cuh"

namespace trg {
namespace {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
  const void* p[kMaxPeers];
};

__device__ __forceinline__ uint4 ld_peer(const void* p) {
  uint4 r;
  // not volatile: the compiler is free to issue the G loads of a vector back to back, ahead of the adds
  asm("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(p));
  return r;
}

// TI: element type of the partial tables (float, or the storage type); T: storage type of add / relu_of / out.
// G > 0: compile-time rank count (2, 4, 8); G = 0: run-time n_peers <= kMaxPeers.
// UN units (16 output bytes each) per thread and iteration, CTA-strided: G * UN * kIn independent 16-byte loads
// are in flight per thread before the first add (UN = 8 / G: eight rows of loads whatever the rank count).
template <typename TI, typename T, int G, int UN>
__global__ void __launch_bounds__(256) peer_reduce_rows_kernel(PeerPtrs parts, int n_peers, long long row0,
                                                               const float* __restrict__ row_scale, const T* add,
                                                               const T* __restrict__ relu_of, long long n_rows,
                                                               int feat, T* out) {
  constexpr int kVec = Elem<T>::kVec;
  constexpr int kIn = (int)(kVec * sizeof(TI) / 16);
  constexpr int kG = G > 0 ? G : kMaxPeers;
  const int units_per_row = feat / kVec;
  const long long n_units = n_rows * units_per_row;
  const size_t in_off = (size_t)row0 * feat;
  const int ng = G > 0 ? G : n_peers;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long u0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; u0 < n_units; u0 += stride * UN) {
    uint4 raw[UN][kG][kIn];
#pragma unroll
    for (int j = 0; j < UN; ++j) {
      const long long u = u0 + j * stride;
      if (u < n_units) {
#pragma unroll
        for (int g = 0; g < kG; ++g) {
          if (g < ng) {
#pragma unroll
            for (int k = 0; k < kIn; ++k)
              raw[j][g][k] = ld_peer(reinterpret_cast<const TI*>(parts.p[g]) + in_off + (size_t)u * kVec + k * (16 / sizeof(TI)));
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < UN; ++j) {
      const long long u = u0 + j * stride;
      if (u >= n_units) break;
      const long long r = u / units_per_row;
      const size_t e0 = (size_t)u * kVec;
      float f[kVec];
#pragma unroll
      for (int k = 0; k < kVec; ++k) f[k] = 0.f;
#pragma unroll
      for (int g = 0; g < kG; ++g) {
        if (g < ng) {
          float x[kVec];
          if (sizeof(TI) == 4) {
#pragma unroll
            for (int k = 0; k < kIn; ++k) {
              x[4 * k] = __uint_as_float(raw[j][g][k].x); x[4 * k + 1] = __uint_as_float(raw[j][g][k].y);
              x[4 * k + 2] = __uint_as_float(raw[j][g][k].z); x[4 * k + 3] = __uint_as_float(raw[j][g][k].w);
            }
          } else {
            Elem<T>::unpack(raw[j][g][0], x);
          }
#pragma unroll
          for (int k = 0; k < kVec; ++k) f[k] = g == 0 ? x[k] : __fadd_rn(f[k], x[k]);   // rank order, no fma contraction
        }
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
}

template <typename TI, typename T>
void launch_peer_reduce(const PeerPtrs& pp, int n_peers, long long row0, const float* row_scale, const void* add,
                        const void* relu_of, long long n_rows, int feat, void* out, int grid, cudaStream_t st) {
  auto a = (const T*)add;
  auto ro = (const T*)relu_of;
  auto o = (T*)out;
  switch (n_peers) {
    case 2: peer_reduce_rows_kernel<TI, T, 2, 4><<<grid, 256, 0, st>>>(pp, n_peers, row0, row_scale, a, ro, n_rows, feat, o); break;
    case 4: peer_reduce_rows_kernel<TI, T, 4, 2><<<grid, 256, 0, st>>>(pp, n_peers, row0, row_scale, a, ro, n_rows, feat, o); break;
    case 8: peer_reduce_rows_kernel<TI, T, 8, 1><<<grid, 256, 0, st>>>(pp, n_peers, row0, row_scale, a, ro, n_rows, feat, o); break;
    default: peer_reduce_rows_kernel<TI, T, 0, 1><<<grid, 256, 0, st>>>(pp, n_peers, row0, row_scale, a, ro, n_rows, feat, o); break;
  }
}

}  // namespace
}  // namespace trg

using namespace trg;

extern "C" int trg_peer_reduce_rows(const void* const* parts, int32_t n_peers, int64_t row0, int in_dtype,
                                    const float* row_scale, const void* add, const void* relu_of, int64_t n_rows,
                                    int32_t feat, int dtype, void* out, int32_t ctas_per_sm, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TRG_CHECK_ARG(parts && n_peers >= 1 && n_peers <= kMaxPeers, "trg_peer_reduce_rows: 1 <= n_peers <= %d", kMaxPeers);
  TRG_CHECK_ARG(n_rows >= 0 && row0 >= 0 && feat > 0, "trg_peer_reduce_rows: bad n_rows/row0/feat");
  TRG_CHECK_ARG(dtype == TRG_F32 || dtype == TRG_BF16, "trg_peer_reduce_rows: unknown dtype %d", dtype);
  TRG_CHECK_ARG(in_dtype == dtype || in_dtype == TRG_F32, "trg_peer_reduce_rows: in_dtype must be dtype or TRG_F32");
  if (n_rows == 0) return TRG_OK;
  const int es = dtype == TRG_BF16 ? 2 : 4;
  TRG_CHECK_ARG(out && (feat * es) % 16 == 0, "trg_peer_reduce_rows: NULL out or rows not a multiple of 16 bytes");
  PeerPtrs pp{};
  uintptr_t bits = (uintptr_t)out | (uintptr_t)add | (uintptr_t)relu_of;
  for (int g = 0; g < n_peers; ++g) {
    TRG_CHECK_ARG(parts[g], "trg_peer_reduce_rows: parts[%d] is NULL", g);
    pp.p[g] = parts[g];
    bits |= (uintptr_t)parts[g];
  }
  TRG_CHECK_ARG(bits % 16 == 0, "trg_peer_reduce_rows: tables must be 16-byte aligned");
  const long long n_units = (long long)n_rows * (feat * es / 16);
  const int per_sm = ctas_per_sm > 0 ? ctas_per_sm : 2;
  const int grid = (int)std::min<long long>((n_units + 255) / 256, (long long)grid_sms() * per_sm);
  if (dtype == TRG_F32)
    launch_peer_reduce<float, float>(pp, n_peers, row0, row_scale, add, relu_of, n_rows, feat, out, grid, st);
  else if (in_dtype == TRG_F32)
    launch_peer_reduce<float, __nv_bfloat16>(pp, n_peers, row0, row_scale, add, relu_of, n_rows, feat, out, grid, st);
  else
    launch_peer_reduce<__nv_bfloat16, __nv_bfloat16>(pp, n_peers, row0, row_scale, add, relu_of, n_rows, feat, out, grid, st);
  count_launch();
  TRG_LAUNCH_OK();
  return TRG_OK;
}
