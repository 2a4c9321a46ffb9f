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
//            double-buffered in TMEM
//   warps 4-7 selection: thread r owns query row r; it reads its row of the accumulator tile with
//            tcgen05.ld, compares against the running K-th best score held in a register and inserts
//            the few survivors into its K-entry min-heap in shared memory (ordered by score desc, id
//            asc: the stream is in ascending id order, so a tie with the K-th best never enters)
// Per-split lists go to the workspace and are merged by trg_topk_merge (also the multi-GPU merge).
// Tensor-bound: 2*B*P*H flops against 2*P*H bytes of catalogue (AI = B = 4096 flop/B).
#include <algorithm>
#include <cfloat>

#include "tc_common.cuh"

namespace trg {
namespace tc {

constexpr int kQRows = 128;
constexpr long long kPadIdTc = 0x7fffffffffffffffLL;

struct ScoreTcParams {
  CUtensorMap q_map;
  CUtensorMap c_map;
  long long n_query, n_cat, id_offset;
  long long tiles_per_split;
  int k, n_splits, kblocks;   // kblocks = H / 64
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
// the accumulator tile is always 128 posts wide (128 / NS stages per tile), so the selection warps
// pay their per-tile costs (barriers, TMEM load latency) once per 128 scores.
template <int NS, int QCAP>
__global__ void __launch_bounds__(256, 1)
    score_topk_tc_kernel(const __grid_constant__ ScoreTcParams p, int n_stages, int list_stride) {
  constexpr int N = 128;                 // posts per accumulator tile
  constexpr int kSub = N / NS;           // ring stages per accumulator tile
  constexpr int kQStride = QCAP + 1;     // odd: lanes own consecutive rows -> conflict-free
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int q_bytes = p.kblocks * kQRows * 128;
  const int stage_bytes = p.kblocks * NS * 128;
  unsigned char* q_smem = smem;
  unsigned char* ring = smem + q_bytes;
  float* lv = reinterpret_cast<float*>(ring + (size_t)n_stages * stage_bytes);
  uint32_t* li = reinterpret_cast<uint32_t*>(lv + kQRows * list_stride);
  float* cq_s = reinterpret_cast<float*>(li + kQRows * list_stride);          // [kQRows][kQStride]
  uint32_t* cq_i = reinterpret_cast<uint32_t*>(cq_s + kQRows * kQStride);
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(cq_i + kQRows * kQStride) + 7) & ~static_cast<uintptr_t>(7));
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
      mbar_init(smem_u32(&tmem_empty[a]), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 2 * N);
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
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss<false>(d, qd + (uint64_t)(2 * k), cd + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty[stage]));
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&tmem_full[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== selection: one query row per thread =====================
    const int wq = warp & 3;
    const int r = wq * 32 + lane;
    const bool row_ok = q0 + r < p.n_query;
    const int K = p.k;
    const uint32_t lv_base = smem_u32(lv), li_base = smem_u32(li);
    const uint32_t qs_base = smem_u32(cq_s), qi_base = smem_u32(cq_i);
    const uint32_t my_qs = qs_base + 4u * (uint32_t)(r * kQStride), my_qi = qi_base + 4u * (uint32_t)(r * kQStride);
    int m = 0;                 // entries in this row's sorted list
    int cnt = 0;               // entries in this row's candidate queue
    float thr = -INFINITY;     // K-th best (score, id) once the list is full
    uint32_t kth_id = 0xffffffffu;
    float thr_g = -INFINITY;   // best K-th score published by any catalogue split for this row
    // merge the queues of every row of this warp holding at least `min_fill` candidates
    auto drain = [&](int min_fill) {
      unsigned need = __ballot_sync(0xffffffffu, cnt >= min_fill && cnt > 0);
      while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        const int row = wq * 32 + L;
        const int n_c = __shfl_sync(0xffffffffu, cnt, L), m_l = __shfl_sync(0xffffffffu, m, L);
        const int nm = warp_merge_row(lv_base + 4u * (uint32_t)(row * list_stride), li_base + 4u * (uint32_t)(row * list_stride),
                                      qs_base + 4u * (uint32_t)(row * kQStride), qi_base + 4u * (uint32_t)(row * kQStride),
                                      n_c, m_l, K, lane);
        if (lane == L) {
          m = nm;
          cnt = 0;
          if (m == K) {
            thr = __uint_as_float(lds32(lv_base + 4u * (uint32_t)(row * list_stride + K - 1)));
            kth_id = lds32(li_base + 4u * (uint32_t)(row * list_stride + K - 1));
          }
        }
      }
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = 0; t < n_tiles; ++t) {
      // Threshold published by the other catalogue splits for this query row: the K-th best of ANY
      // subset is a lower bound of the global K-th best, so scores strictly below it can never be in
      // the answer (ties with it are decided locally).  Refreshed every 16 tiles, off the critical path.
      int tg_key = 0;
      const bool refresh = (t & 15) == 0 && row_ok;
      if (refresh) tg_key = __ldcg(p.thr_shared + q0 + r);
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tc_fence_after();
      const long long p0 = (tile0 + t) * N;                 // first post of the tile (local id)
      const long long left = p.n_cat - p0;
      const int nvalid = left < (long long)N ? (int)left : N;
      const uint32_t base_idx = (uint32_t)(p0 - tile0 * N);  // index relative to the split start
      // whole accumulator row -> registers: all tcgen05.ld issued back to back, one wait
      uint32_t v[N];
#pragma unroll
      for (int c = 0; c < N / 32; ++c)
        tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * N + c * 32), v + c * 32);
      tmem_ld_wait();
      // the accumulator buffer can be handed back to the MMA warp already: scores are in registers
      tc_fence_before();
      mbar_arrive(smem_u32(&tmem_empty[acc]));
      if (refresh) thr_g = fmaxf(thr_g, key_float(tg_key));
      if (nvalid < N || !row_ok) {              // catalogue tail / query rows past B: never candidates
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (j >= nvalid || !row_ok) v[j] = 0xff800000u;   // -inf
      }
      // Common case (no score of the tile can enter any list of this warp): 8 interleaved max chains,
      // ONE warp vote, done.  Otherwise only chains holding a candidate are looked at score by score;
      // survivors go to the row's queue, and queues are merged by the whole warp (checked after every
      // chain = N/8 sites, so a queue of QCAP >= 2*N/8 entries cannot overflow).
      constexpr int kChains = 2 * N / QCAP;                 // 8 (QCAP 32) or 16 (QCAP 16)
      constexpr int kPer = N / kChains;                     // sites per chain = QCAP / 2
      float mx[kChains];
#pragma unroll
      for (int c = 0; c < kChains; ++c) mx[c] = __uint_as_float(v[c]);
#pragma unroll
      for (int j = kChains; j < N; ++j) mx[j & (kChains - 1)] = fmaxf(mx[j & (kChains - 1)], __uint_as_float(v[j]));
      float mall = mx[0];
#pragma unroll
      for (int c = 1; c < kChains; ++c) mall = fmaxf(mall, mx[c]);
      const float before = thr;
      if (__any_sync(0xffffffffu, mall > -INFINITY && mall >= thr && mall >= thr_g)) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
          if (mx[c] > -INFINITY && mx[c] >= thr && mx[c] >= thr_g) {
#pragma unroll
            for (int j = c; j < N; j += kChains) {
              const float s = __uint_as_float(v[j]);
              const uint32_t id = base_idx + (uint32_t)j;
              // survivor iff it can rank before the K-th best (equal score: only with a lower id)
              if (s > -INFINITY && s >= thr_g && (m < K || s > thr || (s == thr && id < kth_id))) {
                sts32(my_qs + 4u * cnt, __float_as_uint(s));
                sts32(my_qi + 4u * cnt, id);
                ++cnt;
              }
            }
          }
          if (__any_sync(0xffffffffu, cnt > QCAP - kPer)) drain(QCAP - kPer + 1);
        }
      }
      if (row_ok && thr > before && thr > thr_g) atomicMax(p.thr_shared + q0 + r, float_key(thr));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    drain(1);   // merge what is left in the queues
    if (row_ok) {
      const float* mylv = lv + r * list_stride;
      const uint32_t* myli = li + r * list_stride;
      float* ov = p.part_vals + ((q0 + r) * p.n_splits + split) * K;
      long long* oi = p.part_ids + ((q0 + r) * p.n_splits + split) * K;
      const long long idbase = p.id_offset + tile0 * N;
      for (int i = 0; i < K; ++i) {
        ov[i] = i < m ? mylv[i] : -INFINITY;
        oi[i] = i < m ? idbase + (long long)myli[i] : kPadIdTc;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * N);
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
struct ScoreCfg { int ns, qcap, stages, smem; };
static int fixed_smem(int hidden, int k, int qcap) {
  return (hidden / 64) * kQRows * 128 + kQRows * (k | 1) * 8 + kQRows * (qcap + 1) * 8 + 1024 + 512;
}
// largest ring stage (catalogue rows) and candidate queue for which >= 2 stages fit beside Q and lists
static ScoreCfg pick_cfg(int hidden, int k) {
  for (int qcap : {32, 16})
    for (int ns : {128, 64, 32}) {
      const int stage = (hidden / 64) * ns * 128;
      const int fixed = fixed_smem(hidden, k, qcap);
      if (fixed + 2 * stage <= kSmemLimit) {
        const int stages = std::min(8, (kSmemLimit - fixed) / stage);
        return {ns, qcap, stages, fixed + stages * stage};
      }
    }
  return {0, 0, 0, 0};
}

int score_tc_splits(int64_t n_query, int64_t n_cat, int hidden, int k, long long* tiles_per_split) {
  (void)hidden; (void)k;
  const int64_t n_tiles = (n_cat + 127) / 128;
  const int64_t q_blocks = (n_query + kQRows - 1) / kQRows;
  int64_t s = std::max<int64_t>(1, kNumSMs / q_blocks);
  s = std::min<int64_t>(s, std::max<int64_t>(1, n_tiles / 8));   // at least ~8 tiles per split
  s = std::min<int64_t>(s, 65535);
  *tiles_per_split = (n_tiles + s - 1) / s;
  return (int)((n_tiles + *tiles_per_split - 1) / *tiles_per_split);
}

bool score_tc_fits(int hidden, int k) { return pick_cfg(hidden, k).ns > 0; }

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
  const ScoreCfg cfg = pick_cfg(hidden, k);
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
#define TRG_SCORE_LAUNCH(NS, QC)                                                                      \
  {                                                                                                   \
    static int set_smem = 0;                                                                          \
    if (smem > set_smem) {                                                                            \
      TRG_CUDA(cudaFuncSetAttribute(score_topk_tc_kernel<NS, QC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      set_smem = smem;                                                                                \
    }                                                                                                 \
    score_topk_tc_kernel<NS, QC><<<grid, 256, smem, st>>>(p, n_stages, list_stride);                  \
  }
  if (cfg.qcap == 32) {
    if (cfg.ns == 128) TRG_SCORE_LAUNCH(128, 32)
    else if (cfg.ns == 64) TRG_SCORE_LAUNCH(64, 32)
    else TRG_SCORE_LAUNCH(32, 32)
  } else {
    if (cfg.ns == 128) TRG_SCORE_LAUNCH(128, 16)
    else if (cfg.ns == 64) TRG_SCORE_LAUNCH(64, 16)
    else TRG_SCORE_LAUNCH(32, 16)
  }
#undef TRG_SCORE_LAUNCH
  count_launch();
  TRG_LAUNCH_OK();
  return trg_topk_merge(p.part_vals, (const int64_t*)p.part_ids, n_query, p.n_splits, k, k, vals_out,
                        ids_out, st);
}

}  // namespace tc
}  // namespace trg
