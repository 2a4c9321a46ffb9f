// K0 -- destination-sorted CSR builder (trg_csr_build): stable LSD radix sort of the COO edge
// list by key + degree histogram + exclusive scan.  HBM-bound integer work; bit-exact with
// argsort(key, stable) / bincount / cumsum (oracle/csr.py).
//
// Replaces: the COO edge_index consumption of PyG's propagate for the tensors built at
// build_graph.py:387,394,402 and train_gnn.py:128-142 (reference keeps COO and scatters with
// atomics; a stable sort once per static graph makes every later pass atomic-free).
#include <algorithm>
#include <cstdio>

#include "common.cuh"

namespace trg {
namespace {

constexpr int kThreads = 256;
constexpr int kItems = 16;
constexpr int kTile = kThreads * kItems;  // 4096 keys per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kRadix = 256;

// ---------------------------------------------------------------------------------------------
// exclusive scan of int32 arrays (3 launches: tile sums, scan of tile sums, apply)
// ---------------------------------------------------------------------------------------------
template <int NWARPS>
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  int ws = (lane < NWARPS) ? warp_sums[lane] : 0;
  int winc = ws;
#pragma unroll
  for (int o = 1; o < NWARPS; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += t;
  }
  const int wbase = __shfl_sync(0xffffffffu, winc - ws, w);
  total = __shfl_sync(0xffffffffu, winc, NWARPS - 1);
  __syncthreads();
  return wbase + inc - v;
}

__global__ void __launch_bounds__(kThreads) scan_tile_sums(const int* __restrict__ in, int64_t n,
                                                           int* __restrict__ tile_sums) {
  __shared__ int warp_sums[kWarps];
  const int64_t base = (int64_t)blockIdx.x * kTile;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    int64_t idx = base + (int64_t)i * kThreads + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  int total;
  block_exclusive_scan<kWarps>(s, warp_sums, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_tile_offsets(int* __restrict__ tile_sums, int n_tiles) {
  __shared__ int warp_sums[32];
  int carry = 0;
  for (int base = 0; base < n_tiles; base += 1024) {
    int idx = base + threadIdx.x;
    int v = idx < n_tiles ? tile_sums[idx] : 0;
    int total;
    int ex = block_exclusive_scan<32>(v, warp_sums, total);
    if (idx < n_tiles) tile_sums[idx] = carry + ex;
    carry += total;
  }
}

// in-place capable: every thread reads its items before any thread of the CTA writes them.
__global__ void __launch_bounds__(kThreads) scan_apply(const int* in, int* out, int64_t n,
                                                       const int* __restrict__ tile_offsets) {
  __shared__ int warp_sums[kWarps];
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  int v[kItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int total;
  int run = block_exclusive_scan<kWarps>(s, warp_sums, total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

int exclusive_scan(const int* in, int* out, int64_t n, int* tile_sums, cudaStream_t st) {
  if (n <= 0) return TRG_OK;
  const int n_tiles = (int)ceil_div<int64_t>(n, kTile);
  scan_tile_sums<<<n_tiles, kThreads, 0, st>>>(in, n, tile_sums);
  scan_tile_offsets<<<1, 1024, 0, st>>>(tile_sums, n_tiles);
  scan_apply<<<n_tiles, kThreads, 0, st>>>(in, out, n, tile_sums);
  count_launch(3);
  TRG_LAUNCH_OK();
  return TRG_OK;
}

// ---------------------------------------------------------------------------------------------
// degree histogram
// ---------------------------------------------------------------------------------------------
// An id outside [0, n_key) would be an out-of-bounds atomic here and an out-of-bounds row read in every
// kernel that later follows the CSR: abort the kernel instead (the reference raises IndexError on the CPU and
// torch's CUDA index kernels raise a device-side assert; this is the same loud failure, not silent
// corruption).  Static graphs are validated on the host at cache-fill time; this guard covers the per-step
// structures built from caller-supplied negatives (train_gnn.py:272).
__global__ void __launch_bounds__(kThreads) count_keys(const long long* __restrict__ key, int64_t n,
                                                       long long n_key, int* __restrict__ deg) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const long long k = ldg_stream(key + i);
    if ((unsigned long long)k >= (unsigned long long)n_key) {
      printf("trg_csr_build: key %lld at edge %lld is outside [0, %lld)\n", k, (long long)i, n_key);
      __trap();
    }
    atomicAdd(&deg[(int)k], 1);
  }
}


// ---------------------------------------------------------------------------------------------
// stable range selection (multi-GPU: this rank's share of the step's negatives, train_gnn.py:272)
// ---------------------------------------------------------------------------------------------
// out = the (key - lo, other) pairs with lo <= key < hi, in input order, padded to `capacity` entries with
// (pad_key, pad_other).  No host synchronisation: the consumers take `capacity` as their edge count and the
// padding lands in a sentinel row past the last real one.  Three launches: per-tile counts, scan of the
// tile counts (also publishes the total and aborts on overflow), stable scatter + padding.
__global__ void __launch_bounds__(kThreads) select_count(const long long* __restrict__ key, int64_t n,
                                                         long long lo, long long hi, int* __restrict__ tile_counts) {
  __shared__ int warp_sums[kWarps];
  const int64_t base = (int64_t)blockIdx.x * kTile;
  int c = 0;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int64_t idx = base + (int64_t)i * kThreads + threadIdx.x;
    if (idx < n) {
      const long long k = ldg_stream(key + idx);
      c += (k >= lo && k < hi) ? 1 : 0;
    }
  }
  int total;
  block_exclusive_scan<kWarps>(c, warp_sums, total);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) select_offsets(int* __restrict__ tile_counts, int n_tiles,
                                                       long long capacity, int* __restrict__ count_out) {
  __shared__ int warp_sums[32];
  int carry = 0;
  for (int base = 0; base < n_tiles; base += 1024) {
    const int idx = base + threadIdx.x;
    const int v = idx < n_tiles ? tile_counts[idx] : 0;
    int total;
    const int ex = block_exclusive_scan<32>(v, warp_sums, total);
    if (idx < n_tiles) tile_counts[idx] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) {
    *count_out = carry;
    if ((long long)carry > capacity) {
      printf("trg_select_range: %d selected entries exceed the capacity %lld (negatives far from uniform? pass a larger "
             "capacity)\n", carry, capacity);
      __trap();
    }
  }
}

// warp w owns the contiguous span [w * 512, w * 512 + 512) of the tile, walked 32 keys at a time: the output
// position of a selected key = tile offset + earlier warps + earlier rounds + lower lanes (ballot) -> stable.
__global__ void __launch_bounds__(kThreads)
    select_scatter(const long long* __restrict__ key, const long long* __restrict__ other, int64_t n, long long lo,
                   long long hi, const int* __restrict__ tile_offsets, const int* __restrict__ count,
                   long long capacity, long long pad_key, long long pad_other, long long* __restrict__ key_out,
                   long long* __restrict__ other_out) {
  __shared__ int wtot[kWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t wbase = (int64_t)blockIdx.x * kTile + (int64_t)w * (kItems * 32);
  long long kv[kItems];
  unsigned sel[kItems];
  int mine = 0;
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int64_t idx = wbase + r * 32 + lane;
    kv[r] = idx < n ? ldg_stream(key + idx) : lo - 1;
    const bool in = kv[r] >= lo && kv[r] < hi;
    sel[r] = __ballot_sync(0xffffffffu, in);
    mine += __popc(sel[r]);
  }
  if (lane == 0) wtot[w] = mine;      // every lane holds the warp total (ballots are warp-wide)
  __syncthreads();
  int run = tile_offsets[blockIdx.x];
  for (int i = 0; i < w; ++i) run += wtot[i];
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    if ((sel[r] >> lane) & 1u) {
      const int dest = run + __popc(sel[r] & ((1u << lane) - 1u));
      if (dest < capacity) {
        key_out[dest] = kv[r] - lo;
        other_out[dest] = ldg_stream(other + wbase + r * 32 + lane);
      }
    }
    run += __popc(sel[r]);
  }
  // padding [count, capacity): spread over the whole grid
  const long long cnt = *count;
  for (long long i = cnt + (long long)blockIdx.x * kThreads + threadIdx.x; i < capacity;
       i += (long long)gridDim.x * kThreads) {
    key_out[i] = pad_key;
    other_out[i] = pad_other;
  }
}

// ---------------------------------------------------------------------------------------------
// one LSD pass: per-CTA digit histogram -> global scan -> stable scatter
// ---------------------------------------------------------------------------------------------
// Passes after the first read and write (key, edge id) as ONE interleaved 8-byte pair: the scatter is
// bound by the number of scattered store instructions, not by bytes (ncu: a pass that stores keys and
// ids separately took 0.80 ms at 40 M edges, the last pass -- ids only -- 0.34 ms).
__device__ __forceinline__ uint2 ldg_pair(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
template <bool FIRST>
__device__ __forceinline__ uint32_t load_key(const void* keys, int64_t idx) {
  if (FIRST) return (uint32_t)ldg_stream(reinterpret_cast<const long long*>(keys) + idx);
  return ldg_pair(reinterpret_cast<const uint2*>(keys) + idx).x;
}

template <bool FIRST>
__global__ void __launch_bounds__(kThreads) radix_hist(const void* __restrict__ keys, int64_t n,
                                                       int shift, int* __restrict__ hist,
                                                       int n_ctas) {
  __shared__ int h[kRadix];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    int64_t idx = base + (int64_t)i * kThreads + threadIdx.x;
    if (idx < n) atomicAdd(&h[(load_key<FIRST>(keys, idx) >> shift) & (kRadix - 1)], 1);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * n_ctas + blockIdx.x] = h[threadIdx.x];
}

// Stability: warp w owns the contiguous span [w*512, w*512+512) of the tile and walks it 32 keys
// at a time; a key's rank among equal digits = (earlier CTAs) + (earlier warps) + (earlier
// rounds of this warp) + (lower lanes of this round, via match_any).
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kThreads)
    radix_scatter(const void* __restrict__ keys_in, int64_t n, int shift, const int* __restrict__ gbase,
                  int n_ctas, uint2* __restrict__ pairs_out, uint32_t* __restrict__ vals_out) {
  __shared__ int wcnt[kWarps][kRadix];
  __shared__ int sbase[kRadix];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kWarps; ++i) wcnt[i][threadIdx.x] = 0;
  const int64_t wbase = (int64_t)blockIdx.x * kTile + (int64_t)w * (kItems * 32);

  uint32_t key[kItems], val[kItems];
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    int64_t idx = wbase + r * 32 + lane;
    bool valid = idx < n;
    if (FIRST) {
      key[r] = valid ? load_key<true>(keys_in, idx) : 0u;
      val[r] = (uint32_t)idx;
    } else {
      const uint2 kv = valid ? ldg_pair(reinterpret_cast<const uint2*>(keys_in) + idx) : make_uint2(0u, 0u);
      key[r] = kv.x;
      val[r] = kv.y;
    }
  }
  __syncthreads();

  int rank[kItems];
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const bool valid = wbase + r * 32 + lane < n;
    const int d = valid ? (int)((key[r] >> shift) & (kRadix - 1)) : kRadix;  // sentinel never matches
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int prev = valid ? wcnt[w][d] : 0;
    __syncwarp();
    if (valid && lane == __ffs(peers) - 1) wcnt[w][d] = prev + __popc(peers);
    __syncwarp();
    rank[r] = prev + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  {
    int run = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) {
      int t = wcnt[i][threadIdx.x];
      wcnt[i][threadIdx.x] = run;
      run += t;
    }
    sbase[threadIdx.x] = gbase[(int64_t)threadIdx.x * n_ctas + blockIdx.x];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    if (wbase + r * 32 + lane < n) {
      const int d = (int)((key[r] >> shift) & (kRadix - 1));
      const int dest = sbase[d] + wcnt[w][d] + rank[r];
      if (LAST)
        vals_out[dest] = val[r];
      else
        pairs_out[dest] = make_uint2(key[r], val[r]);
    }
  }
}

__global__ void __launch_bounds__(kThreads) gather_col(const long long* __restrict__ other,
                                                       const int* __restrict__ eid, int64_t n,
                                                       int* __restrict__ col) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) col[i] = (int)other[ldg_stream(eid + i)];
}

int num_passes(int64_t n_key) {
  int bits = 1;
  while (bits < 32 && ((int64_t)1 << bits) < n_key) ++bits;
  return (bits + 7) / 8;
}

struct Workspace {
  uint2* pairs[2];       // (key, edge id) ping-pong buffers
  uint32_t* vals;        // edge ids of the last pass when the caller does not want them
  int *hist, *tile_sums;
  size_t bytes;
};

Workspace carve(void* ws, int64_t e, int64_t n_key) {
  Workspace w;
  const size_t ebytes = align_up((size_t)(e > 0 ? e : 1) * 4, 256);
  const int64_t n_ctas = ceil_div<int64_t>(e > 0 ? e : 1, kTile);
  const size_t hist_bytes = align_up((size_t)kRadix * n_ctas * 4, 256);
  const int64_t scan_len = (int64_t)kRadix * n_ctas > n_key + 1 ? (int64_t)kRadix * n_ctas : n_key + 1;
  const size_t sums_bytes = align_up((size_t)ceil_div<int64_t>(scan_len, kTile) * 4 + 4, 256);
  char* p = reinterpret_cast<char*>(ws);
  w.pairs[0] = (uint2*)p; p += 2 * ebytes;
  w.pairs[1] = (uint2*)p; p += 2 * ebytes;
  w.vals = (uint32_t*)p; p += ebytes;
  w.hist = (int*)p; p += hist_bytes;
  w.tile_sums = (int*)p; p += sums_bytes;
  w.bytes = (size_t)(p - reinterpret_cast<char*>(ws));
  return w;
}

}  // namespace
}  // namespace trg

using namespace trg;

extern "C" size_t trg_csr_workspace_bytes(int64_t n_edges, int64_t n_key) {
  if (n_edges < 0 || n_key < 0) return 0;
  return carve(nullptr, n_edges, n_key).bytes;
}

extern "C" int trg_csr_build(const int64_t* other, const int64_t* key, int64_t e, int64_t n_key,
                             int32_t* rowptr, int32_t* col, int32_t* eid, void* workspace,
                             size_t workspace_bytes, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TRG_CHECK_ARG(e >= 0 && n_key >= 0, "trg_csr_build: negative size");
  TRG_CHECK_ARG(e < ((int64_t)1 << 31) && n_key < ((int64_t)1 << 31) - 1,
                "trg_csr_build: n_edges=%lld / n_key=%lld exceed int32 CSR range", (long long)e,
                (long long)n_key);
  TRG_CHECK_ARG(rowptr != nullptr, "trg_csr_build: rowptr is NULL");
  TRG_CHECK_ARG(e == 0 || (key != nullptr && (col == nullptr || other != nullptr)),
                "trg_csr_build: NULL edge arrays with n_edges > 0");
  TRG_CHECK_ARG(e == 0 || n_key > 0, "trg_csr_build: edges present but n_key == 0");
  Workspace w = carve(workspace, e, n_key);
  if (workspace == nullptr || workspace_bytes < w.bytes) {
    set_error("trg_csr_build: workspace %zu < required %zu", workspace_bytes, w.bytes);
    return TRG_E_WORKSPACE;
  }
  TRG_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(n_key + 1) * sizeof(int32_t), st));
  if (e == 0) return TRG_OK;

  // degrees -> rowptr (exclusive scan over n_key + 1 entries; entry n_key is 0 -> rowptr[n_key] = E)
  {
    int grid = (int)std::min<int64_t>(ceil_div<int64_t>(e, kThreads), (int64_t)kNumSMs * 16);
    count_keys<<<grid, kThreads, 0, st>>>((const long long*)key, e, (long long)n_key, rowptr);
    count_launch();
    TRG_LAUNCH_OK();
    int rc = exclusive_scan(rowptr, rowptr, n_key + 1, w.tile_sums, st);
    if (rc) return rc;
  }
  if (col == nullptr && eid == nullptr) return TRG_OK;

  const int passes = num_passes(n_key);
  const int n_ctas = (int)ceil_div<int64_t>(e, kTile);
  const void* kin = key;
  const uint32_t* vin = nullptr;
  for (int p = 0; p < passes; ++p) {
    const bool first = p == 0, last = p == passes - 1;
    const int shift = 8 * p;
    uint2* kout = w.pairs[p & 1];
    uint32_t* vout = eid ? (uint32_t*)eid : w.vals;     // written by the last pass only
    if (first)
      radix_hist<true><<<n_ctas, kThreads, 0, st>>>(kin, e, shift, w.hist, n_ctas);
    else
      radix_hist<false><<<n_ctas, kThreads, 0, st>>>(kin, e, shift, w.hist, n_ctas);
    count_launch();
    TRG_LAUNCH_OK();
    int rc = exclusive_scan(w.hist, w.hist, (int64_t)kRadix * n_ctas, w.tile_sums, st);
    if (rc) return rc;
    if (first && last)
      radix_scatter<true, true><<<n_ctas, kThreads, 0, st>>>(kin, e, shift, w.hist, n_ctas, kout, vout);
    else if (first)
      radix_scatter<true, false><<<n_ctas, kThreads, 0, st>>>(kin, e, shift, w.hist, n_ctas, kout, vout);
    else if (last)
      radix_scatter<false, true><<<n_ctas, kThreads, 0, st>>>(kin, e, shift, w.hist, n_ctas, kout, vout);
    else
      radix_scatter<false, false><<<n_ctas, kThreads, 0, st>>>(kin, e, shift, w.hist, n_ctas, kout, vout);
    count_launch();
    TRG_LAUNCH_OK();
    kin = kout;
    vin = vout;
  }
  if (col != nullptr) {
    int grid = (int)std::min<int64_t>(ceil_div<int64_t>(e, kThreads), (int64_t)kNumSMs * 16);
    gather_col<<<grid, kThreads, 0, st>>>((const long long*)other, (const int*)vin, e, col);
    count_launch();
    TRG_LAUNCH_OK();
  }
  return TRG_OK;
}

extern "C" size_t trg_select_range_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  return align_up((size_t)(ceil_div<int64_t>(n > 0 ? n : 1, kTile) + 1) * 4, 256);
}

extern "C" int trg_select_range(const int64_t* key, const int64_t* other, int64_t n, int64_t lo, int64_t hi,
                                int64_t capacity, int64_t pad_key, int64_t pad_other, int64_t* key_out,
                                int64_t* other_out, int32_t* count_out, void* workspace, size_t workspace_bytes,
                                void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  TRG_CHECK_ARG(n >= 0 && capacity >= 0 && lo <= hi, "trg_select_range: bad sizes");
  TRG_CHECK_ARG(n < ((int64_t)1 << 31) && capacity < ((int64_t)1 << 31), "trg_select_range: sizes exceed int32 range");
  TRG_CHECK_ARG(count_out && (capacity == 0 || (key_out && other_out)), "trg_select_range: NULL outputs");
  TRG_CHECK_ARG(n == 0 || (key && other), "trg_select_range: NULL inputs with n > 0");
  if (workspace == nullptr || workspace_bytes < trg_select_range_workspace_bytes(n)) {
    set_error("trg_select_range: workspace %zu < required %zu", workspace_bytes, trg_select_range_workspace_bytes(n));
    return TRG_E_WORKSPACE;
  }
  int* tile_counts = reinterpret_cast<int*>(workspace);
  const int n_tiles = (int)ceil_div<int64_t>(n > 0 ? n : 1, kTile);
  select_count<<<n_tiles, kThreads, 0, st>>>((const long long*)key, n, lo, hi, tile_counts);
  select_offsets<<<1, 1024, 0, st>>>(tile_counts, n_tiles, capacity, count_out);
  select_scatter<<<n_tiles, kThreads, 0, st>>>((const long long*)key, (const long long*)other, n, lo, hi, tile_counts,
                                               count_out, capacity, pad_key, pad_other, (long long*)key_out,
                                               (long long*)other_out);
  count_launch(3);
  TRG_LAUNCH_OK();
  return TRG_OK;
}
