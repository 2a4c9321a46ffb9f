/* Plain-C restatement of the two integer algorithms on the hot path.  TEST INFRASTRUCTURE:
 * only tests/, smoke() and bench.py's cpu_baseline leg may load this.
 *
 *  - oracle_csr_by_dst: stable counting sort of the COO edge list by destination; the contract
 *    the CUDA builder must match bit for bit (reference input format: build_graph.py:387,394,402;
 *    train_gnn.py:128-142).  Equivalent to argsort(dst, stable) + bincount + cumsum.
 *  - oracle_topk_row: top-k of one score row under (score desc, id asc), the canonical form of
 *    torch.topk at inference.py:428.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int oracle_csr_by_dst(const int64_t* src, const int64_t* dst, int64_t e, int64_t n_dst,
                      int64_t* rowptr /* n_dst+1 */, int64_t* col /* e */, int64_t* eid /* e */) {
    memset(rowptr, 0, (size_t)(n_dst + 1) * sizeof(int64_t));
    for (int64_t i = 0; i < e; ++i) {
        if (dst[i] < 0 || dst[i] >= n_dst) return 1;
        rowptr[dst[i] + 1]++;
    }
    for (int64_t i = 0; i < n_dst; ++i) rowptr[i + 1] += rowptr[i];
    int64_t* cur = (int64_t*)malloc((size_t)(n_dst > 0 ? n_dst : 1) * sizeof(int64_t));
    if (!cur) return 2;
    memcpy(cur, rowptr, (size_t)n_dst * sizeof(int64_t));
    for (int64_t i = 0; i < e; ++i) {            /* ascending edge id => stable */
        int64_t p = cur[dst[i]]++;
        col[p] = src[i];
        eid[p] = i;
    }
    free(cur);
    return 0;
}

/* a ranks before b iff score larger, or equal score and smaller id */
static int before(float sa, int64_t ia, float sb, int64_t ib) {
    return sa > sb || (sa == sb && ia < ib);
}

int oracle_topk_row(const float* scores, int64_t n, int64_t k, int64_t id_offset,
                    float* vals /* k */, int64_t* ids /* k */) {
    if (k > n) k = n;
    int64_t m = 0;                               /* sorted insertion list, best first */
    for (int64_t i = 0; i < n; ++i) {
        float s = scores[i];
        int64_t id = i + id_offset;
        if (m == k && !before(s, id, vals[m - 1], ids[m - 1])) continue;
        int64_t p = (m < k) ? m++ : k - 1;
        while (p > 0 && before(s, id, vals[p - 1], ids[p - 1])) {
            vals[p] = vals[p - 1]; ids[p] = ids[p - 1]; --p;
        }
        vals[p] = s; ids[p] = id;
    }
    return 0;
}
