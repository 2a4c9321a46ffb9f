// K5 (tcgen05 form) -- bf16 score contraction + streaming top-k on the tensor cores.
//
// Replaces  scores = torch.mm(user_emb, known_post_emb.T); torch.topk(scores, min(K, n))
// (inference.py:427-428) for batched queries against a large catalogue (BASELINE config 5:
// 4096 x 50M, K = 100).  The [B, P] score matrix is never written: it only ever exists as
// [128 x N] fp32 accumulator tiles in TMEM.
//
// CTA = 128 queries (one TMEM lane each) x one contiguous catalogue split.
//   warp 0   TMA: the query block once (all k-blocks resident), then catalogue tiles [N posts x H]
//            through an mbarrier ring (SWIZZLE_128B, K-major)
//   warp 1   one thread issues tcgen05.mma kind::f16 (bf16 in, fp32 accumulate), accumulators
//            double-buffered in TMEM columns [0, 256)
//   warps 4+ selection: thread r owns query row r (= TMEM lane r); it reads its row of the accumulator
//            tile with tcgen05.ld, rejects the whole tile with one max tree + one warp vote against the
//            row's K-th best, pushes the few survivors to a per-thread queue in shared memory, and the
//            warp merges full queues into the row's sorted top-K list (score desc, id asc; the stream is
//            in ascending id order, so a tie with the K-th best never enters).  The lists live in TMEM
//            columns [256, 512) of the row's own lane, which leaves shared memory to the catalogue ring.
// Per-split lists go to the workspace and are merged by trg_topk_merge (also the multi-GPU merge).
// Tensor-bound: 2*B*P*H flops against 2*P*H bytes of catalogue (AI = B = 4096 flop/B).
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"

namespace trg {
namespace tc {

constexpr int kQRows = 128;
// TMEM columns: [0,256) two accumulator tiles | [256,384) the row's top-K scores | [384,512) their ids.
// Each query row is one TMEM lane, so a row's sorted list lives in its own lane (K <= 128): the 103 KB
// the lists took in shared memory now hold catalogue ring stages (ablation in profiles/README.md: with
// a 3 x 16 KB ring the TMA + MMA pipeline alone ran at 0.97 us per 128-post tile, 3.4x the MMA time).
constexpr int kTmemColsTopk = 512;
constexpr uint32_t kListScoreCol = 256, kListIdCol = 384;
constexpr long long kPadIdTc = 0x7fffffffffffffffLL;

struct ScoreTcParams {
  CUtensorMap q_map;
  CUtensorMap c_map;
  long long n_query, n_cat, id_offset;
  long long tiles_per_split;
  int k, n_splits, kblocks;   // kblocks = H / 64
  int dbg;                    // ablation switches (TRG_TOPK_DBG): 1 = no selection, 2 = no tcgen05.ld either, 4 = no MMA
  int* thr_shared;            // [B] ordered-int keys of the best published K-th score per query row
  float* part_vals;           // [B][n_splits][k]
  long long* part_ids;        // [B][n_splits][k]
};

// float <-> int key whose signed order equals the float order (for atomicMax on scores of any sign)
__device__ __forceinline__ int float_key(float f) {
  int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
__global__ void fill_int(int* p, long long n, int v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// a ranks before b under (score desc, id asc)
__device__ __forceinline__ bool ranks_before(float sa, uint32_t ia, float sb, uint32_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

// Warp-cooperative merge of one row's candidate queue (<= 32 unsorted entries) into that row's
// sorted top-K list (best first, K <= 128): bitonic sort of the candidates across the lanes, then a
// rank merge -- every element's final slot = its own index + the number of elements of the OTHER
// sequence ranking before it (binary searches), so all writes are conflict-free and in place.
// ~200 warp instructions per merge, i.e. a few per candidate, instead of a per-candidate heap walk
// executed by a single lane.  Returns the new list length.
__device__ __forceinline__ int warp_merge_row(uint32_t lv_a, uint32_t li_a, uint32_t qs_a, uint32_t qi_a,
                                              int n_c, int m_l, int K, int lane) {
  float cs = -INFINITY;
  uint32_t ci = 0xffffffffu;
  if (lane < n_c) {
    cs = __uint_as_float(lds32(qs_a + 4u * lane));
    ci = lds32(qi_a + 4u * lane);
  }
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float ps = __shfl_xor_sync(0xffffffffu, cs, j);
      const uint32_t pi = __shfl_xor_sync(0xffffffffu, ci, j);
      const bool up = (lane & k) == 0, lower = (lane & j) == 0;
      const bool mine_first = ranks_before(cs, ci, ps, pi);
      if ((lower == up) != mine_first) { cs = ps; ci = pi; }
    }
  }
  // list entries of this lane: i = lane + 32 t
  float ls[4];
  uint32_t lid[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int i = lane + 32 * t;
    ls[t] = -INFINITY;
    lid[t] = 0xffffffffu;
    if (i < m_l) {
      ls[t] = __uint_as_float(lds32(lv_a + 4u * i));
      lid[t] = lds32(li_a + 4u * i);
    }
  }
  // candidate (sorted position = lane): number of list entries ranking before it
  int lo = 0, hi = m_l;
  while (lo < hi) {   // warp-uniform trip count bound: log2(128) + 1; lanes may finish early
    const int mid = (lo + hi) >> 1;
    const float s = __uint_as_float(lds32(lv_a + 4u * mid));
    const uint32_t id = lds32(li_a + 4u * mid);
    if (ranks_before(s, id, cs, ci)) lo = mid + 1; else hi = mid;
  }
  const int cpos = lane + lo;
  // list entries: number of (sorted) candidates ranking before each
  int lpos[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    int a = 0, b = n_c;
#pragma unroll
    for (int it = 0; it < 6; ++it) {   // 2^5 = 32 candidates (+1 step to converge)
      const int mid = (a + b) >> 1;
      const float s = __shfl_sync(0xffffffffu, cs, mid & 31);
      const uint32_t id = __shfl_sync(0xffffffffu, ci, mid & 31);
      if (a < b) {
        if (ranks_before(s, id, ls[t], lid[t])) a = mid + 1; else b = mid;
      }
    }
    lpos[t] = lane + 32 * t + a;
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (lane + 32 * t < m_l && lpos[t] < K) {
      sts32(lv_a + 4u * lpos[t], __float_as_uint(ls[t]));
      sts32(li_a + 4u * lpos[t], lid[t]);
    }
  }
  if (lane < n_c && cpos < K) {
    sts32(lv_a + 4u * cpos, __float_as_uint(cs));
    sts32(li_a + 4u * cpos, ci);
  }
  __syncwarp();
  return min(K, m_l + n_c);
}

// NS = catalogue rows per ring stage (128 / 64 / 32: what fits beside Q, the lists and the queues);
// the accumulator tile is always 128 posts wide (128 / NS stages per tile).
//
// Selection runs on 4 * NPART warps: warp (wq, part) owns the TMEM lane quarter wq (query rows
// 32 wq .. 32 wq + 31, one per lane) and the column range [part * 128/NPART, ...) of every tile, so a
// row is scanned by NPART threads of different warps.  ncu on the one-warp-per-quarter form
// (profiles/README.md): the tensor pipe was 16 % busy and the single selection warp of each SM
// sub-partition issued in 15 % of its cycles -- every tcgen05.ld / vote / branch latency was exposed
// because nothing else could issue.  With NPART warps per sub-partition the latencies overlap and a
// tile is released after 128/NPART scores per thread instead of 128.
//
// Shared per-row state (the row's sorted top-K list, its length, and a 64-bit snapshot (K-th score,
// K-th id) read with one LDS.64 per tile) is only written under a per-row lock by the warp that merges
// a candidate queue into the list.  The snapshot a scanning thread holds may be stale, which is
// conservative: the K-th best only improves, so nothing that belongs in the answer is ever filtered
// out, and the merge ranks candidates exactly under (score desc, id asc).
template <int NS, int NPART>
__global__ void __launch_bounds__(128 + 128 * NPART, 1)
    score_topk_tc_kernel(const __grid_constant__ ScoreTcParams p, int n_stages, int list_stride) {
  constexpr int N = 128;                 // posts per accumulator tile
  constexpr int kSub = N / NS;           // ring stages per accumulator tile
  constexpr int CW = N / NPART;          // columns scanned per thread
  constexpr int QCAP = 32;               // per-thread candidate queue
  constexpr int kChains = 8;             // interleaved max chains (CW / 8 sites each)
  constexpr int kCheckEvery = N / CW;    // chains between queue-overflow checks: <= 16 pushes in between
  constexpr int kSel = 128 * NPART;      // selection threads
  constexpr int kQStride = QCAP + 1;     // odd: lanes own consecutive rows -> conflict-free
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int q_bytes = p.kblocks * kQRows * 128;
  const int stage_bytes = p.kblocks * NS * 128;
  unsigned char* q_smem = smem;
  unsigned char* ring = smem + q_bytes;
  float* sv = reinterpret_cast<float*>(ring + (size_t)n_stages * stage_bytes);   // merge scratch [4*NPART warps][128]
  uint32_t* si = reinterpret_cast<uint32_t*>(sv + 4 * NPART * 128);
  float* cq_s = reinterpret_cast<float*>(si + 4 * NPART * 128);                 // [kSel][kQStride]
  uint32_t* cq_i = reinterpret_cast<uint32_t*>(cq_s + kSel * kQStride);
  uint2* row_thr = reinterpret_cast<uint2*>(
      (reinterpret_cast<uintptr_t>(cq_i + kSel * kQStride) + 7) & ~static_cast<uintptr_t>(7));   // [128] (K-th score, K-th id)
  int* row_m = reinterpret_cast<int*>(row_thr + kQRows);                       // [128] list length
  int* q_lock = row_m + kQRows;                                                // [4] one per TMEM lane quarter
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_lock + 4);
  uint64_t* full = bars;          // [8]
  uint64_t* empty = full + 8;     // [8]
  uint64_t* q_full = empty + 8;   // [1]
  uint64_t* tmem_full = q_full + 1;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q0 = (long long)blockIdx.x * kQRows;
  const int split = blockIdx.y;
  const long long n_tiles_total = (p.n_cat + N - 1) / N;
  const long long tile0 = (long long)split * p.tiles_per_split;
  const long long tile1 = min(tile0 + p.tiles_per_split, n_tiles_total);
  const int n_tiles = tile1 > tile0 ? (int)(tile1 - tile0) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q_map);
    tma_prefetch_desc(&p.c_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    mbar_init(smem_u32(q_full), 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), kSel);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), kTmemColsTopk);
  if (threadIdx.x >= 128 && threadIdx.x < 256) {      // part 0 threads: one per row
    const int r = threadIdx.x - 128;
    row_thr[r] = make_uint2(0xff800000u, 0xffffffffu);   // (-inf, max id): "list not full"
    row_m[r] = 0;
    if (r < 4) q_lock[r] = 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // query block: every k-block, resident for the CTA's lifetime (rows past B are zero-filled)
      mbar_arrive_expect_tx(smem_u32(q_full), (uint32_t)q_bytes);
      for (int kb = 0; kb < p.kblocks; ++kb)
        tma_load_2d(smem_u32(q_smem + kb * kQRows * 128), &p.q_map, smem_u32(q_full), kb * 64, (int)q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < n_tiles * kSub; ++t) {
        mbar_wait_backoff(smem_u32(&empty[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&full[stage]);
        mbar_arrive_expect_tx(fb, (uint32_t)stage_bytes);
        const int row = (int)(tile0 * N + (long long)t * NS);
        for (int kb = 0; kb < p.kblocks; ++kb)
          tma_load_2d(smem_u32(ring + (size_t)stage * stage_bytes + kb * NS * 128), &p.c_map, fb, kb * 64, row);
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtBF16, 0, 0, kQRows, NS);
      mbar_wait(smem_u32(q_full), 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait_backoff(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        for (int sub = 0; sub < kSub; ++sub) {
          mbar_wait_backoff(smem_u32(&full[stage]), phase);
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(acc * N + sub * NS);
          for (int kb = 0; kb < p.kblocks; ++kb) {
            const uint64_t qd = make_smem_desc_sw128(smem_u32(q_smem + kb * kQRows * 128), 0, 1024);
            const uint64_t cd = make_smem_desc_sw128(smem_u32(ring + (size_t)stage * stage_bytes + kb * NS * 128), 0, 1024);
            if (!(p.dbg & 4)) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_ss<false>(d, qd + (uint64_t)(2 * k), cd + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&empty[stage]));
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&tmem_full[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== selection: thread = (query row, column part) =====================
    const int wq = warp & 3;
    const int part = (warp - 4) >> 2;
    const int r = wq * 32 + lane;
    const bool row_ok = q0 + r < p.n_query;
    const int K = p.k;
    const uint32_t qs_base = smem_u32(cq_s), qi_base = smem_u32(cq_i);
    const uint32_t thr_addr = smem_u32(row_thr + r);
    const int qslot = part * kQRows + r;
    const uint32_t my_qs = qs_base + 4u * (uint32_t)(qslot * kQStride), my_qi = qi_base + 4u * (uint32_t)(qslot * kQStride);
    const uint32_t sv_a = smem_u32(sv + (warp - 4) * 128), si_a = smem_u32(si + (warp - 4) * 128);   // this warp's scratch
    const uint32_t tl_s = tmem_base + ((uint32_t)(wq * 32) << 16) + kListScoreCol;   // this thread's list lane
    const uint32_t tl_i = tmem_base + ((uint32_t)(wq * 32) << 16) + kListIdCol;
    int cnt = 0;               // entries in this thread's candidate queue
    // best K-th score published by any catalogue split for this row; query rows past B (zero-filled by
    // TMA) get +inf, so nothing of theirs ever passes the filter
    float thr_g = row_ok ? -INFINITY : INFINITY;
    // Merge the queue of every lane holding at least `min_fill` candidates into that row's list.  The
    // list of row L sits in TMEM lane L, reachable only by thread L of the warps of this quarter, and
    // tcgen05.ld/st move all 32 lanes at once: the list is staged through the warp's scratch (thread L
    // copies its 16-column chunks out), merged there by the whole warp (warp_merge_row), and written
    // back chunk by chunk -- the other 31 threads store back the values they just loaded.  NPART > 1:
    // the warps sharing a quarter take the quarter's lock, since a write-back rewrites every lane.
    auto drain = [&](int min_fill) {
      unsigned need = __ballot_sync(0xffffffffu, cnt >= min_fill && cnt > 0);
      if (!need) return;
      if (NPART > 1) {
        if (lane == 0) {
          while (atomicCAS(q_lock + wq, 0, 1) != 0) {
          }
        }
        __syncwarp();
        tc_fence_after();
      }
      while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        const int row = wq * 32 + L;
        const int n_c = __shfl_sync(0xffffffffu, cnt, L);
        asm volatile("" ::: "memory");
        const int m_l = *reinterpret_cast<volatile int*>(row_m + row);
#pragma unroll 1
        for (int c = 0; c * 16 < m_l; ++c) {                 // warp-uniform trip count
          uint32_t rs[16], ri[16];
          tmem_ld_32x16(tl_s + (uint32_t)(c * 16), rs);
          tmem_ld_32x16(tl_i + (uint32_t)(c * 16), ri);
          tmem_ld_wait();
          if (lane == L) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              sts128(sv_a + 4u * (uint32_t)(c * 16 + j), make_uint4(rs[j], rs[j + 1], rs[j + 2], rs[j + 3]));
              sts128(si_a + 4u * (uint32_t)(c * 16 + j), make_uint4(ri[j], ri[j + 1], ri[j + 2], ri[j + 3]));
            }
          }
        }
        __syncwarp();
        const uint32_t qrow = (uint32_t)((part * kQRows + row) * kQStride);
        const int nm = warp_merge_row(sv_a, si_a, qs_base + 4u * qrow, qi_base + 4u * qrow, n_c, m_l, K, lane);
#pragma unroll 1
        for (int c = 0; c * 16 < nm; ++c) {
          uint32_t rs[16], ri[16];
          tmem_ld_32x16(tl_s + (uint32_t)(c * 16), rs);
          tmem_ld_32x16(tl_i + (uint32_t)(c * 16), ri);
          tmem_ld_wait();
          if (lane == L) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const uint4 a4 = lds128(sv_a + 4u * (uint32_t)(c * 16 + j));
              const uint4 b4 = lds128(si_a + 4u * (uint32_t)(c * 16 + j));
              rs[j] = a4.x; rs[j + 1] = a4.y; rs[j + 2] = a4.z; rs[j + 3] = a4.w;
              ri[j] = b4.x; ri[j + 1] = b4.y; ri[j + 2] = b4.z; ri[j + 3] = b4.w;
            }
          }
          tmem_st_32x16(tl_s + (uint32_t)(c * 16), rs);
          tmem_st_32x16(tl_i + (uint32_t)(c * 16), ri);
        }
        tmem_st_wait();
        if (lane == L) {
          *reinterpret_cast<volatile int*>(row_m + row) = nm;
          if (nm == K) {
            const uint32_t ts = lds32(sv_a + 4u * (uint32_t)(K - 1));
            const uint32_t ti = lds32(si_a + 4u * (uint32_t)(K - 1));
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(smem_u32(row_thr + row)), "r"(ts), "r"(ti) : "memory");
            // publish the row's K-th best to the other catalogue splits (monotone, so stale reads are safe)
            if (q0 + row < p.n_query) atomicMax(p.thr_shared + q0 + row, float_key(__uint_as_float(ts)));
          }
          cnt = 0;
        }
        __syncwarp();
      }
      if (NPART > 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          atomicExch(q_lock + wq, 0);
        }
      }
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    // TRG_TOPK_DBG & 8: cycle accounting of the selection loop (printed by CTA (0,0))
    const bool prof = (p.dbg & 8) != 0;
    long long c_wait = 0, c_ld = 0, c_fast = 0, c_slow = 0, n_slow = 0, n_drain = 0, c_drain = 0, tk = 0;
    for (int t = 0; t < n_tiles; ++t) {
      if (prof) tk = clock64();
      // Threshold published by the other catalogue splits for this query row: the K-th best of ANY
      // subset is a lower bound of the global K-th best, so scores strictly below it can never be in
      // the answer (ties with it are decided locally).  Refreshed every 16 tiles, off the critical path.
      int tg_key = 0;
      const bool refresh = (t & 15) == 0 && row_ok;
      if (refresh) tg_key = __ldcg(p.thr_shared + q0 + r);
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tc_fence_after();
      if (prof) { const long long n = clock64(); c_wait += n - tk; tk = n; }
      const long long p0 = (tile0 + t) * N;                 // first post of the tile (local id)
      const long long left = p.n_cat - p0;
      // valid columns of this part (catalogue tail: TMA zero-fills the rest, filtered at push time)
      const int nvalid = (left < (long long)N ? (int)left : N) - part * CW;
      const uint32_t base_idx = (uint32_t)(p0 - tile0 * N) + (uint32_t)(part * CW);  // relative to the split start
      // accumulator row -> registers, 32 columns at a time, with the running maxima of chunk c-1 taken
      // while chunk c is in flight (cycle accounting, TRG_TOPK_DBG=8: the four loads issued back to back
      // cost ~420 clocks per tile on their own, as much as the rest of the fast path and the MMA)
      uint32_t v[CW];
      float mx[kChains];
      const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * N + part * CW);
      if (p.dbg & 2) {
#pragma unroll
        for (int j = 0; j < CW; ++j) v[j] = 0xff800000u;
#pragma unroll
        for (int c = 0; c < kChains; ++c) mx[c] = -INFINITY;
      } else {
        tmem_ld_32x32(t_row, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < kChains; ++c) mx[c] = __uint_as_float(v[c]);
#define TRG_TOPK_MAX_CHUNK(CH, J0)                                                              \
  _Pragma("unroll") for (int j = (J0); j < 32; ++j)                                             \
      mx[j & (kChains - 1)] = fmaxf(mx[j & (kChains - 1)], __uint_as_float(v[(CH) * 32 + j]));
        if constexpr (CW > 32) tmem_ld_32x32(t_row + 32u, v + 32);
        TRG_TOPK_MAX_CHUNK(0, kChains)
        if constexpr (CW > 32) tmem_ld_wait();
        if constexpr (CW > 64) tmem_ld_32x32(t_row + 64u, v + 64);
        if constexpr (CW > 32) { TRG_TOPK_MAX_CHUNK(1, 0) }
        if constexpr (CW > 64) tmem_ld_wait();
        if constexpr (CW > 96) tmem_ld_32x32(t_row + 96u, v + 96);
        if constexpr (CW > 64) { TRG_TOPK_MAX_CHUNK(2, 0) }
        if constexpr (CW > 96) {
          tmem_ld_wait();
          TRG_TOPK_MAX_CHUNK(3, 0)
        }
#undef TRG_TOPK_MAX_CHUNK
      }
      // the accumulator buffer can be handed back to the MMA warp already: scores are in registers
      tc_fence_before();
      mbar_arrive(smem_u32(&tmem_empty[acc]));
      if (prof) { const long long n = clock64(); c_ld += n - tk; tk = n; }
      if (refresh) thr_g = fmaxf(thr_g, key_float(tg_key));
      if (p.dbg & 1) {
        if (v[0] == 0x12345678u) cnt = 1;   // keep the loads alive
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      // this row's (K-th score, K-th id), one consistent 64-bit read
      uint32_t snap_s, snap_i;
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(snap_s), "=r"(snap_i) : "r"(thr_addr) : "memory");
      const float thr = __uint_as_float(snap_s);
      const uint32_t kth_id = snap_i;
      // Common case (no score of the tile can enter any list of this warp): kChains interleaved max
      // chains, ONE warp vote, done.  Otherwise only chains holding a candidate are looked at score by
      // score; survivors go to the thread's queue, and queues are merged by the whole warp (checked
      // every kCheckEvery chains = at most 16 pushes, so a queue of 32 entries cannot overflow).
      float mall = mx[0];
#pragma unroll
      for (int c = 1; c < kChains; ++c) mall = fmaxf(mall, mx[c]);
      const bool any_cand = __any_sync(0xffffffffu, mall >= thr && mall >= thr_g);
      if (prof) { const long long n = clock64(); c_fast += n - tk; tk = n; }
      if (any_cand) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
          if (mx[c] >= thr && mx[c] >= thr_g) {
#pragma unroll
            for (int j = c; j < CW; j += kChains) {
              const float s = __uint_as_float(v[j]);
              const uint32_t id = base_idx + (uint32_t)j;
              // survivor iff it can rank before the K-th best (equal score: only with a lower id)
              if (j < nvalid && s >= thr_g && (s > thr || (s == thr && id < kth_id))) {
                sts32(my_qs + 4u * cnt, __float_as_uint(s));
                sts32(my_qi + 4u * cnt, id);
                ++cnt;
              }
            }
          }
          if ((c + 1) % kCheckEvery == 0 && __any_sync(0xffffffffu, cnt > QCAP - 16)) {
            long long td = 0;
            if (prof) td = clock64();
            drain(QCAP - 16 + 1);
            if (prof) { c_drain += clock64() - td; ++n_drain; }
          }
        }
        if (prof) { const long long n = clock64(); c_slow += n - tk; tk = n; ++n_slow; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (prof && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0)
      printf("sel warp %d: tiles %d | cycles/tile: wait %.0f ld %.0f fast %.0f slow %.0f (of which drain %.0f) | slow tiles %.3f/tile drains %.4f/tile, %.0f cyc/slow %.0f cyc/drain\n",
             warp, n_tiles, (double)c_wait / n_tiles, (double)c_ld / n_tiles, (double)c_fast / n_tiles,
             (double)c_slow / n_tiles, (double)c_drain / n_tiles, (double)n_slow / n_tiles, (double)n_drain / n_tiles,
             n_slow ? (double)c_slow / n_slow : 0.0, n_drain ? (double)c_drain / n_drain : 0.0);
    drain(1);   // merge what is left in the queues
    asm volatile("bar.sync 1, %0;" ::"n"(kSel) : "memory");   // every part of every row has merged
    if (part == 0) {     // every thread reads its own row's list back from its TMEM lane
      tc_fence_after();
      const int m = row_m[r];
      float* ov = p.part_vals + ((q0 + r) * p.n_splits + split) * K;
      long long* oi = p.part_ids + ((q0 + r) * p.n_splits + split) * K;
      const long long idbase = p.id_offset + tile0 * N;
#pragma unroll 1
      for (int c = 0; c * 16 < K; ++c) {
        uint32_t rs[16], ri[16];
        tmem_ld_32x16(tl_s + (uint32_t)(c * 16), rs);
        tmem_ld_32x16(tl_i + (uint32_t)(c * 16), ri);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int i = c * 16 + j;
            if (i < K) {
              ov[i] = i < m ? __uint_as_float(rs[j]) : -INFINITY;
              oi[i] = i < m ? idbase + (long long)ri[j] : kPadIdTc;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemColsTopk);
}

// ---- scan / helper split (v2) ------------------------------------------------------------------------
// Cycle accounting of the kernel above (profiles/README.md, r1_v15): a selection warp spends 1 780 cycles
// in every tile that holds a candidate (12 % of the tiles) and 9 400 in every queue merge, and because
// the accumulator ring is two tiles deep the other three warps wait for it -- ~40 % of the tile time.
// Here the warps that read the accumulators never leave the fast path:
//   warps 4-7  SCAN: tcgen05.ld of the row, max tree, one vote.  A lane whose row has a candidate copies
//              its 128 scores (32 x STS.128) into a slot of its quarter's ring and moves on.
//   warps 8-11 HELPER (same TMEM lane quarter as warp - 4): take slots in order, test the 128 scores
//              cooperatively (4 per lane) against the row's exact K-th best, append survivors to the row's
//              queue and merge full queues into the row's list in TMEM.  Only the helper touches queues
//              and lists, so there are no locks; the scan warps filter with a snapshot that may be stale
//              (conservative), the helper ranks exactly, in stream order (= ascending id).
//
// Tried and measured (config 5): the query block as a TMEM-resident A operand (QT = true: copied once with
// tcgen05.st, MMAs issued as umma_ts) with 96-post tiles so that accumulators (2 x 96), lists (2 x 128) and Q
// (64 columns) fit the 512 TMEM columns.  Exact (all top-k tests pass) but no faster than Q in shared memory
// at the same tile width (89.6 vs 89.7 ms), and 96-post tiles cost 15 % against 128-post ones (74 - 78 ms):
// the slow shared-memory stores of the hand-off (~30 cycles per STS.128) are NOT caused by the MMA's operand
// reads.  The product therefore runs 128-post tiles with Q in shared memory; the QT path stays compiled for
// kTileN2 = 96 builds.
constexpr int kSlotRing = 8;                 // slots per quarter
constexpr int kTileN2 = 128;                 // posts per accumulator tile
constexpr int kSlotWords = kTileN2 + 4;      // scores + header (lane, base index, valid columns), 16 B aligned
constexpr int kQCap2 = 32;
constexpr uint32_t kListScoreCol2 = 2 * kTileN2, kListIdCol2 = kListScoreCol2 + 128, kQCol2 = kListIdCol2 + 128;

template <int NS, bool QT>
__global__ void __launch_bounds__(384, 1)
    score_topk_tc2_kernel(const __grid_constant__ ScoreTcParams p, int n_stages) {
  constexpr int N = kTileN2;
  constexpr int kSub = N / NS;
  constexpr int kChains = 8;
  constexpr int kQStride = kQCap2 + 1;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int q_bytes = p.kblocks * kQRows * 128;
  const int stage_bytes = p.kblocks * NS * 128;
  unsigned char* q_smem = smem;
  unsigned char* ring = smem + q_bytes;
  uint32_t* slots = reinterpret_cast<uint32_t*>(ring + (size_t)n_stages * stage_bytes);   // [4][kSlotRing][kSlotWords]
  float* sv = reinterpret_cast<float*>(slots + 4 * kSlotRing * kSlotWords);              // merge scratch [4][128]
  uint32_t* si = reinterpret_cast<uint32_t*>(sv + 4 * 128);
  float* cq_s = reinterpret_cast<float*>(si + 4 * 128);                                  // [128][kQStride]
  uint32_t* cq_i = reinterpret_cast<uint32_t*>(cq_s + kQRows * kQStride);
  uint2* row_thr = reinterpret_cast<uint2*>(
      (reinterpret_cast<uintptr_t>(cq_i + kQRows * kQStride) + 7) & ~static_cast<uintptr_t>(7));   // (K-th score, K-th id)
  float* row_tg = reinterpret_cast<float*>(row_thr + kQRows);                            // threshold of the other splits
  volatile int* ctl = reinterpret_cast<volatile int*>(row_tg + kQRows);                  // [4][4]: head, tail, done
  uint64_t* bars = reinterpret_cast<uint64_t*>(const_cast<int*>(ctl) + 16);
  uint64_t* full = bars;          // [8]
  uint64_t* empty = full + 8;     // [8]
  uint64_t* q_full = empty + 8;   // [1]
  uint64_t* tmem_full = q_full + 1;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint64_t* q_tmem = tmem_empty + 2;     // [1] Q copied into tensor memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_tmem + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q0 = (long long)blockIdx.x * kQRows;
  const int split = blockIdx.y;
  const long long n_tiles_total = (p.n_cat + N - 1) / N;
  const long long tile0 = (long long)split * p.tiles_per_split;
  const long long tile1 = min(tile0 + p.tiles_per_split, n_tiles_total);
  const int n_tiles = tile1 > tile0 ? (int)(tile1 - tile0) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.q_map);
    tma_prefetch_desc(&p.c_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    mbar_init(smem_u32(q_full), 1);
    mbar_init(smem_u32(q_tmem), 128);
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), kTmemColsTopk);
  if (threadIdx.x >= 128 && threadIdx.x < 256) {
    const int r = threadIdx.x - 128;
    row_thr[r] = make_uint2(0xff800000u, 0xffffffffu);   // (-inf, max id): "list not full"
    row_tg[r] = -INFINITY;
    if (r < 16) ctl[r] = 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(smem_u32(q_full), (uint32_t)q_bytes);
      for (int kb = 0; kb < p.kblocks; ++kb)
        tma_load_2d(smem_u32(q_smem + kb * kQRows * 128), &p.q_map, smem_u32(q_full), kb * 64, (int)q0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < n_tiles * kSub; ++t) {
        mbar_wait_backoff(smem_u32(&empty[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&full[stage]);
        mbar_arrive_expect_tx(fb, (uint32_t)stage_bytes);
        const int row = (int)(tile0 * N + (long long)t * NS);
        for (int kb = 0; kb < p.kblocks; ++kb)
          tma_load_2d(smem_u32(ring + (size_t)stage * stage_bytes + kb * NS * 128), &p.c_map, fb, kb * 64, row);
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtBF16, 0, 0, kQRows, NS);
      if (QT) {
        mbar_wait(smem_u32(q_tmem), 0);
        tc_fence_after();
      } else {
        mbar_wait(smem_u32(q_full), 0);
      }
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait_backoff(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        for (int sub = 0; sub < kSub; ++sub) {
          mbar_wait_backoff(smem_u32(&full[stage]), phase);
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(acc * N + sub * NS);
          for (int kb = 0; kb < p.kblocks; ++kb) {
            const uint64_t qd = make_smem_desc_sw128(smem_u32(q_smem + kb * kQRows * 128), 0, 1024);
            const uint64_t cd = make_smem_desc_sw128(smem_u32(ring + (size_t)stage * stage_bytes + kb * NS * 128), 0, 1024);
            if (!(p.dbg & 4)) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (QT)     // Q from tensor memory: 64 bf16 of a k-block = 32 columns, 16 per MMA = 8 columns
                  umma_ts<false>(d, tmem_base + kQCol2 + (uint32_t)(kb * 32 + k * 8), cd + (uint64_t)(2 * k), idesc,
                                 (kb | k) ? 1u : 0u);
                else
                  umma_ss<false>(d, qd + (uint64_t)(2 * k), cd + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
              }
            }
          }
          umma_commit(smem_u32(&empty[stage]));
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&tmem_full[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== SCAN: one query row per thread, fast path only =====================
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const bool row_ok = q0 + r < p.n_query;
    const uint32_t thr_addr = smem_u32(row_thr + r);
    const uint32_t slot_base = smem_u32(slots + wq * kSlotRing * kSlotWords);
    volatile int* c_head = ctl + wq * 4 + 0;
    volatile int* c_tail = ctl + wq * 4 + 1;
    volatile int* c_done = ctl + wq * 4 + 2;
    float thr_g = row_ok ? -INFINITY : INFINITY;   // rows past B (zero-filled by TMA) never produce candidates
    if (QT) {
      // query block: shared memory (SWIZZLE_128B, K-major) -> this thread's TMEM lane; the 32-bit words go over
      // as they are (bf16 pair k, k+1 of the row = one A-operand column)
      mbar_wait(smem_u32(q_full), 0);
      const uint32_t qrow = smem_u32(q_smem) + (uint32_t)r * 128u;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        uint32_t w[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 x = lds128(qrow + (uint32_t)(kb * kQRows * 128) + ((((uint32_t)c) ^ ((uint32_t)r & 7u)) << 4));
          w[4 * c] = x.x; w[4 * c + 1] = x.y; w[4 * c + 2] = x.z; w[4 * c + 3] = x.w;
        }
        tmem_st_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + kQCol2 + (uint32_t)(kb * 32), w);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(q_tmem));
    }
    int head = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool prof = (p.dbg & 8) != 0;      // cycle accounting, printed by CTA (0,0)
    long long c_wait = 0, c_ld = 0, c_fast = 0, c_ev = 0, n_ev = 0, tk = 0, c_e1 = 0, c_e2 = 0, c_e3 = 0;
    for (int t = 0; t < n_tiles; ++t) {
      if (prof) tk = clock64();
      int tg_key = 0;
      const bool refresh = (t & 15) == 0 && row_ok;
      if (refresh) tg_key = __ldcg(p.thr_shared + q0 + r);
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tc_fence_after();
      if (prof) { const long long n = clock64(); c_wait += n - tk; tk = n; }
      const long long p0 = (tile0 + t) * N;
      const long long left = p.n_cat - p0;
      const int nvalid = left < (long long)N ? (int)left : N;
      const uint32_t base_idx = (uint32_t)(p0 - tile0 * N);
      uint32_t v[N];
      const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * N);
      if (p.dbg & 2) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = 0xff800000u;
      } else {
#pragma unroll
        for (int c = 0; c < N / 32; ++c) tmem_ld_32x32(t_row + (uint32_t)(c * 32), v + c * 32);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tmem_empty[acc]));      // scores are in registers: hand the buffer back
      if (prof) { const long long n = clock64(); c_ld += n - tk; tk = n; }
      if (refresh) {
        thr_g = fmaxf(thr_g, key_float(tg_key));
        row_tg[r] = thr_g;                           // the helper filters with it too
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (p.dbg & 1) {
        if (v[0] == 0x12345678u) head = 1;
        continue;
      }
      uint32_t snap_s, snap_i;
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(snap_s), "=r"(snap_i) : "r"(thr_addr) : "memory");
      const float thr = __uint_as_float(snap_s);
      float mx[kChains];
#pragma unroll
      for (int c = 0; c < kChains; ++c) mx[c] = __uint_as_float(v[c]);
#pragma unroll
      for (int j = kChains; j < N; ++j) mx[j & (kChains - 1)] = fmaxf(mx[j & (kChains - 1)], __uint_as_float(v[j]));
      float mall = mx[0];
#pragma unroll
      for (int c = 1; c < kChains; ++c) mall = fmaxf(mall, mx[c]);
      const bool cand = mall >= thr && mall >= thr_g;
      const unsigned mask = __ballot_sync(0xffffffffu, cand);
      if (prof) { const long long n = clock64(); c_fast += n - tk; tk = n; }
      // every lane with a candidate hands its whole row of the tile to the helper; when more lanes have
      // one than the ring has free slots (the first tiles: every list is still empty) they go in rounds
      unsigned rem = mask;
      while (rem) {
        int space = 0;
        if (lane == 0) {
          while ((space = kSlotRing - (head - *c_tail)) <= 0) {
          }
        }
        space = __shfl_sync(0xffffffffu, space, 0);
        long long te = 0;
        if (prof) { te = clock64(); c_e1 += te - tk; }
        const bool pending = (rem >> lane) & 1u;
        const int k = __popc(rem & ((1u << lane) - 1u));
        const bool go = pending && k < space;
        if (go) {
          const uint32_t sa = slot_base + (uint32_t)(((head + k) & (kSlotRing - 1)) * kSlotWords * 4);
#pragma unroll
          for (int i = 0; i < N / 4; ++i) sts128(sa + (uint32_t)(i * 16), make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
          sts128(sa + (uint32_t)(N * 4), make_uint4((uint32_t)lane, base_idx, (uint32_t)nvalid, 0u));
        }
        __syncwarp();
        if (prof) { const long long n = clock64(); c_e2 += n - te; te = n; }
        head += min(__popc(rem), space);
        if (lane == 0) {
          if (!(p.dbg & 16)) __threadfence_block();
          *c_head = head;
        }
        if (prof) { const long long n = clock64(); c_e3 += n - te; }
        rem = __ballot_sync(0xffffffffu, pending && !go);
      }
      if (prof && mask) { c_ev += clock64() - tk; ++n_ev; }
    }
    if (prof && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0)
      printf("scan warp %d: tiles %d | cycles/tile: wait %.0f ld %.0f fast %.0f hand-off %.0f | hot tiles %.3f/tile, %.0f cyc each "
             "(ring space %.0f, copy %.0f, publish %.0f)\n",
             warp, n_tiles, (double)c_wait / n_tiles, (double)c_ld / n_tiles, (double)c_fast / n_tiles,
             (double)c_ev / n_tiles, (double)n_ev / n_tiles, n_ev ? (double)c_ev / n_ev : 0.0,
             n_ev ? (double)c_e1 / n_ev : 0.0, n_ev ? (double)c_e2 / n_ev : 0.0, n_ev ? (double)c_e3 / n_ev : 0.0);
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      *c_done = 1;
    }
  } else if (warp >= 8) {
    // ===================== HELPER: candidates -> queues -> lists (TMEM) =====================
    const int wq = warp & 3;
    const int K = p.k;
    const uint32_t qs_base = smem_u32(cq_s), qi_base = smem_u32(cq_i);
    const uint32_t sv_a = smem_u32(sv + wq * 128), si_a = smem_u32(si + wq * 128);
    const uint32_t tl_s = tmem_base + ((uint32_t)(wq * 32) << 16) + kListScoreCol2;
    const uint32_t tl_i = tmem_base + ((uint32_t)(wq * 32) << 16) + kListIdCol2;
    const uint32_t slot_base = smem_u32(slots + wq * kSlotRing * kSlotWords);
    volatile int* c_head = ctl + wq * 4 + 0;
    volatile int* c_tail = ctl + wq * 4 + 1;
    volatile int* c_done = ctl + wq * 4 + 2;
    int cnt = 0;     // lane l: entries in the queue of row 32 wq + l
    int m = 0;       // lane l: entries in the list of row 32 wq + l
    // merge the last `n_c` (<= 32) queue entries of row L into its list (TMEM lane L, staged through scratch)
    auto merge_row = [&](int L, int n_c, int q_off) {
      const int row = wq * 32 + L;
      const int m_l = __shfl_sync(0xffffffffu, m, L);
#pragma unroll 1
      for (int c = 0; c * 16 < m_l; ++c) {
        uint32_t rs[16], ri[16];
        tmem_ld_32x16(tl_s + (uint32_t)(c * 16), rs);
        tmem_ld_32x16(tl_i + (uint32_t)(c * 16), ri);
        tmem_ld_wait();
        if (lane == L) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            sts128(sv_a + 4u * (uint32_t)(c * 16 + j), make_uint4(rs[j], rs[j + 1], rs[j + 2], rs[j + 3]));
            sts128(si_a + 4u * (uint32_t)(c * 16 + j), make_uint4(ri[j], ri[j + 1], ri[j + 2], ri[j + 3]));
          }
        }
      }
      __syncwarp();
      const uint32_t qrow = (uint32_t)(row * kQStride + q_off);
      const int nm = warp_merge_row(sv_a, si_a, qs_base + 4u * qrow, qi_base + 4u * qrow, n_c, m_l, K, lane);
#pragma unroll 1
      for (int c = 0; c * 16 < nm; ++c) {
        uint32_t rs[16], ri[16];
        tmem_ld_32x16(tl_s + (uint32_t)(c * 16), rs);
        tmem_ld_32x16(tl_i + (uint32_t)(c * 16), ri);
        tmem_ld_wait();
        if (lane == L) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const uint4 a4 = lds128(sv_a + 4u * (uint32_t)(c * 16 + j));
            const uint4 b4 = lds128(si_a + 4u * (uint32_t)(c * 16 + j));
            rs[j] = a4.x; rs[j + 1] = a4.y; rs[j + 2] = a4.z; rs[j + 3] = a4.w;
            ri[j] = b4.x; ri[j + 1] = b4.y; ri[j + 2] = b4.z; ri[j + 3] = b4.w;
          }
        }
        tmem_st_32x16(tl_s + (uint32_t)(c * 16), rs);
        tmem_st_32x16(tl_i + (uint32_t)(c * 16), ri);
      }
      tmem_st_wait();
      if (lane == L) {
        m = nm;
        if (nm == K) {
          const uint32_t ts = lds32(sv_a + 4u * (uint32_t)(K - 1));
          const uint32_t ti = lds32(si_a + 4u * (uint32_t)(K - 1));
          asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(smem_u32(row_thr + row)), "r"(ts), "r"(ti) : "memory");
          if (q0 + row < p.n_query) atomicMax(p.thr_shared + q0 + row, float_key(__uint_as_float(ts)));
        }
      }
      __syncwarp();
    };
    auto drain_row = [&](int L) {      // empty row L's queue, 32 candidates at a time
      int c = __shfl_sync(0xffffffffu, cnt, L);
      while (c > 0) {
        const int n_c = min(c, 32);
        merge_row(L, n_c, c - n_c);
        c -= n_c;
      }
      if (lane == L) cnt = 0;
    };
    int tail = 0;
    for (;;) {
      int h = 0, dn = 0;
      if (lane == 0) {
        h = *c_head;
        if (h == tail) {
          dn = *c_done;
          if (dn) h = *c_head;      // done is published after the last head
        }
      }
      h = __shfl_sync(0xffffffffu, h, 0);
      dn = __shfl_sync(0xffffffffu, dn, 0);
      if (h == tail) {
        if (dn) break;
        __nanosleep(64);
        continue;
      }
      __threadfence_block();
      for (; tail != h; ++tail) {
        const uint32_t sa = slot_base + (uint32_t)((tail & (kSlotRing - 1)) * kSlotWords * 4);
        const uint4 hdr = lds128(sa + (uint32_t)(N * 4));
        const int L = (int)hdr.x;
        const uint32_t base_idx = hdr.y;
        const int nvalid = (int)hdr.z;
        const int row = wq * 32 + L;
        const float tg = *reinterpret_cast<volatile float*>(row_tg + row);
#pragma unroll 1
        for (int i = 0; i < N / 32; ++i) {
          // the row's exact K-th best (it may have moved in the previous sub-step's merge)
          uint32_t ts, ti;
          asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(ts), "=r"(ti) : "r"(smem_u32(row_thr + row)) : "memory");
          const float thr = __uint_as_float(ts);
          const int j = i * 32 + lane;
          const float s = __uint_as_float(lds32(sa + (uint32_t)(j * 4)));
          const uint32_t id = base_idx + (uint32_t)j;
          const bool pass = j < nvalid && s >= tg && (s > thr || (s == thr && id < ti));
          const unsigned pm = __ballot_sync(0xffffffffu, pass);
          if (pm) {
            const int n = __popc(pm);
            int c = __shfl_sync(0xffffffffu, cnt, L);
            if (c + n > kQCap2) {            // make room: merge what is queued (the threshold only gets tighter;
              drain_row(L);                  // survivors of the stale test are re-ranked exactly by the merge)
              c = 0;
            }
            if (pass) {
              const int k = c + __popc(pm & ((1u << lane) - 1u));
              sts32(qs_base + 4u * (uint32_t)(row * kQStride + k), __float_as_uint(s));
              sts32(qi_base + 4u * (uint32_t)(row * kQStride + k), id);
            }
            __syncwarp();
            if (lane == L) cnt = c + n;
          }
        }
        if (__shfl_sync(0xffffffffu, cnt, L) > 16) drain_row(L);
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          *c_tail = tail + 1;
        }
      }
    }
    // merge what is left in the queues, then every thread writes its own row's list
    {
      unsigned need = __ballot_sync(0xffffffffu, cnt > 0);
      while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        drain_row(L);
      }
    }
    {
      const int r = wq * 32 + lane;
      const bool row_ok = q0 + r < p.n_query;
      float* ov = p.part_vals + ((q0 + r) * p.n_splits + split) * K;
      long long* oi = p.part_ids + ((q0 + r) * p.n_splits + split) * K;
      const long long idbase = p.id_offset + tile0 * N;
#pragma unroll 1
      for (int c = 0; c * 16 < K; ++c) {
        uint32_t rs[16], ri[16];
        tmem_ld_32x16(tl_s + (uint32_t)(c * 16), rs);
        tmem_ld_32x16(tl_i + (uint32_t)(c * 16), ri);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int i = c * 16 + j;
            if (i < K) {
              ov[i] = i < m ? __uint_as_float(rs[j]) : -INFINITY;
              oi[i] = i < m ? idbase + (long long)ri[j] : kPadIdTc;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemColsTopk);
}

int make_tmap_2d(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols,
                 uint64_t ld_elems, uint32_t box_rows);

struct ScoreCfg;
static ScoreCfg pick_cfg(int hidden, int k);
bool score_tc_fits(int hidden, int k);
bool score_tc_eligible(int hidden, int dtype, int k) {
  return dtype == TRG_BF16 && hidden % 64 == 0 && hidden >= 64 && hidden <= 256 && k >= 1 && k <= 128 &&
         score_tc_fits(hidden, k) && get_encode_tiled() != nullptr;
}

constexpr int kSmemLimit = 227 * 1024;
struct ScoreCfg { int ns, npart, stages, smem; };
// Selection warps per TMEM lane quarter (TRG_TOPK_NPART=1|2).  Measured at config 5 (4096 x 50M, K=100):
// 1 -> 96.8 ms, 2 -> 105-117 ms, 4 -> 137 ms.  More warps do not help because the step is not bound by
// one warp's latency: every candidate event (about K ln(n/K) per row, ~1.5 per 128-post tile per CTA)
// sends one warp down the slow path, and with a two-deep accumulator ring every other warp then waits
// for it -- the tile time is the MAXIMUM over the selection warps, and more warps means more maxima.
static int score_npart() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("TRG_TOPK_NPART");
    v = (e && e[0] == '2') ? 2 : 1;
  }
  return v;
}
static int fixed_smem(int hidden, int k, int npart) {
  const int qcap = 32;
  (void)k;   // the lists live in TMEM
  return (hidden / 64) * kQRows * 128 + 4 * npart * 128 * 8 /*merge scratch*/ + 128 * npart * (qcap + 1) * 8 +
         kQRows * 12 + 16 /*row state*/ + 1024 + 512;
}
// largest ring stage (catalogue rows) for which >= 2 stages fit beside Q, the lists and the queues
static ScoreCfg pick_cfg(int hidden, int k) {
  const int npart = score_npart();
  const int fixed = fixed_smem(hidden, k, npart);
  for (int min_stages : {4, 2})      // deep ring first (TMA latency x bandwidth), widest stage that allows it
    for (int ns : {128, 64, 32}) {
      const int stage = (hidden / 64) * ns * 128;
      const int stages = std::min(8, (kSmemLimit - fixed) / stage);
      if (stages >= min_stages) return {ns, npart, stages, fixed + stages * stage};
    }
  return {0, 0, 0, 0};
}

static bool score_v1();
int score_tc_splits(int64_t n_query, int64_t n_cat, int hidden, int k, long long* tiles_per_split) {
  (void)hidden; (void)k;
  const int64_t tile = score_v1() ? 128 : kTileN2;
  const int64_t n_tiles = (n_cat + tile - 1) / tile;
  const int64_t q_blocks = (n_query + kQRows - 1) / kQRows;
  int64_t s = std::max<int64_t>(1, kNumSMs / q_blocks);
  s = std::min<int64_t>(s, std::max<int64_t>(1, n_tiles / 8));   // at least ~8 tiles per split
  s = std::min<int64_t>(s, 65535);
  *tiles_per_split = (n_tiles + s - 1) / s;
  return (int)((n_tiles + *tiles_per_split - 1) / *tiles_per_split);
}

static bool score_v1() {          // A/B: TRG_TOPK_V=1 keeps the one-role selection kernel
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TRG_TOPK_V");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
static int fixed_smem2(int hidden) {
  return (hidden / 64) * kQRows * 128 + 4 * kSlotRing * kSlotWords * 4 + 4 * 128 * 8 + kQRows * (kQCap2 + 1) * 8 +
         kQRows * 12 + 64 + 8 + 1024 + 512;
}
static ScoreCfg pick_cfg2(int hidden) {
  const int fixed = fixed_smem2(hidden);
  for (int min_stages : {4, 2})
    for (int ns : {kTileN2, kTileN2 / 2}) {
      const int stage = (hidden / 64) * ns * 128;
      const int stages = std::min(8, (kSmemLimit - fixed) / stage);
      if (stages >= min_stages) return {ns, 1, stages, fixed + stages * stage};
    }
  return {0, 0, 0, 0};
}
bool score_tc_fits(int hidden, int k) { return pick_cfg(hidden, k).ns > 0 && pick_cfg2(hidden).ns > 0; }

size_t score_tc_workspace_bytes(int64_t n_query, int64_t n_cat, int hidden, int k) {
  long long tps;
  const int splits = score_tc_splits(n_query, n_cat, hidden, k, &tps);
  return align_up((size_t)n_query * splits * k * 4, 256) + align_up((size_t)n_query * splits * k * 8, 256) +
         align_up((size_t)n_query * 4, 256);
}

int score_topk_tc(const void* q, const void* cat, int64_t n_query, int64_t n_cat, int hidden, int k,
                  int64_t id_offset, float* vals_out, int64_t* ids_out, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
  const size_t need = score_tc_workspace_bytes(n_query, n_cat, hidden, k);
  if (!ws || ws_bytes < need) {
    set_error("trg_score_topk: workspace %zu < required %zu", ws_bytes, need);
    return TRG_E_WORKSPACE;
  }
  const ScoreCfg cfg = score_v1() ? pick_cfg(hidden, k) : pick_cfg2(hidden);
  if (cfg.ns == 0) {
    set_error("trg_score_topk(bf16): hidden=%d k=%d does not fit in shared memory", hidden, k);
    return TRG_E_UNSUPPORTED;
  }
  ScoreTcParams p{};
  int rc = make_tmap_2d(&p.q_map, q, TRG_BF16, (uint64_t)n_query, hidden, hidden, kQRows);
  if (rc) return rc;
  rc = make_tmap_2d(&p.c_map, cat, TRG_BF16, (uint64_t)n_cat, hidden, hidden, cfg.ns);
  if (rc) return rc;
  p.n_query = n_query; p.n_cat = n_cat; p.id_offset = id_offset; p.k = k; p.kblocks = hidden / 64;
  { const char* e = getenv("TRG_TOPK_DBG"); p.dbg = e ? atoi(e) : 0; }
  p.n_splits = score_tc_splits(n_query, n_cat, hidden, k, &p.tiles_per_split);
  p.part_vals = reinterpret_cast<float*>(ws);
  p.part_ids = reinterpret_cast<long long*>(reinterpret_cast<char*>(ws) +
                                            align_up((size_t)n_query * p.n_splits * k * 4, 256));
  p.thr_shared = reinterpret_cast<int*>(reinterpret_cast<char*>(p.part_ids) +
                                        align_up((size_t)n_query * p.n_splits * k * 8, 256));
  fill_int<<<(unsigned)std::min<int64_t>(1024, (n_query + 255) / 256), 256, 0, st>>>(p.thr_shared, n_query, (int)0x807fffff);  // key of -inf
  count_launch();
  const int list_stride = k | 1;                          // odd stride: conflict-free row access
  const int smem = cfg.smem;
  const int n_stages = cfg.stages;
  dim3 grid((unsigned)((n_query + kQRows - 1) / kQRows), (unsigned)p.n_splits);
#define TRG_SCORE_LAUNCH(NS, NP)                                                                      \
  {                                                                                                   \
    static int set_smem = 0;                                                                          \
    if (smem > set_smem) {                                                                            \
      TRG_CUDA(cudaFuncSetAttribute(score_topk_tc_kernel<NS, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      set_smem = smem;                                                                                \
    }                                                                                                 \
    score_topk_tc_kernel<NS, NP><<<grid, 128 + 128 * NP, smem, st>>>(p, n_stages, list_stride);       \
  }
#define TRG_SCORE_LAUNCH2(NS, QT)                                                                     \
  {                                                                                                   \
    static int set_smem = 0;                                                                          \
    if (smem > set_smem) {                                                                            \
      TRG_CUDA(cudaFuncSetAttribute(score_topk_tc2_kernel<NS, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      set_smem = smem;                                                                                \
    }                                                                                                 \
    score_topk_tc2_kernel<NS, QT><<<grid, 384, smem, st>>>(p, n_stages);                              \
  }
  if (!score_v1()) {
    // Q operand in tensor memory: only when it fits beside accumulators and lists (kTileN2 = 96 builds)
    const bool qt = (int)kQCol2 + 64 <= kTmemColsTopk && hidden <= 128 && !getenv("TRG_TOPK_QSMEM");
    if (cfg.ns == kTileN2) { if (qt) TRG_SCORE_LAUNCH2(kTileN2, true) else TRG_SCORE_LAUNCH2(kTileN2, false) }
    else { if (qt) TRG_SCORE_LAUNCH2(kTileN2 / 2, true) else TRG_SCORE_LAUNCH2(kTileN2 / 2, false) }
  } else if (cfg.npart == 2) {
    if (cfg.ns == 128) TRG_SCORE_LAUNCH(128, 2)
    else if (cfg.ns == 64) TRG_SCORE_LAUNCH(64, 2)
    else TRG_SCORE_LAUNCH(32, 2)
  } else {
    if (cfg.ns == 128) TRG_SCORE_LAUNCH(128, 1)
    else if (cfg.ns == 64) TRG_SCORE_LAUNCH(64, 1)
    else TRG_SCORE_LAUNCH(32, 1)
  }
#undef TRG_SCORE_LAUNCH
#undef TRG_SCORE_LAUNCH2
  count_launch();
  TRG_LAUNCH_OK();
  return trg_topk_merge(p.part_vals, (const int64_t*)p.part_ids, n_query, p.n_splits, k, k, vals_out,
                        ids_out, st);
}

}  // namespace tc
}  // namespace trg
