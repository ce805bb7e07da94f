/* Plain-C restatement of the SVS retrieve hot path.  TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the
 * product (svs_b200/); only tests/ and bench.py's CPU-baseline legs may use it.  It exists as a second, independent
 * checker next to oracle/svs_oracle.py (which calls NumPy, the very dependency the reference calls): the two are
 * compared with each other and with the golden vectors the real reference produced (tests/test_oracle_c.py).
 *
 * Reference lines restated (paths relative to /root/reference):
 *   svs_scores       src/svs/kb.py:1185, 1623      x = np.dot(embeddings_matrix, query_vec)   (float32 result)
 *   svs_get_top_k    src/svs/util.py:190-203       clip k to n; k <= 0 -> nothing; the k largest scores
 *                                                  (np.argpartition treats NaN as largest), returned sorted by
 *                                                  (score, index) descending -- `sorted(..., reverse=True)` on tuples
 *   svs_superheavy   src/svs/kb.py:1622-1627       dot, top-k, row index -> embeddings.id
 *   svs_blob_to_row  src/svs/embeddings/util.py:19-23  little-endian float32 blob -> floats (length % 4 == 0)
 *
 * The arithmetic of np.dot lives in OpenBLAS (sgemv, float32, blocked, FMA): its summation ORDER is not part of the
 * reference, so scores are reproduced here to rounding (<= 1e-6 absolute on unit vectors), not bit for bit; the
 * float32 dot product below accumulates in float like sgemv does.  Which of several elements tied exactly at the
 * k-th score np.argpartition keeps is introselect-dependent; here the (score, index) order
 * decides, i.e. the larger index wins -- identical whenever the boundary is not tied.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_ref/libsvs_oracle_c.so oracle/svs_oracle_c.c   (oracle/build_oracle_c.py)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* x[i] = sum_j M[i*d + j] * q[j], float32 accumulation, four partial sums (the order is unspecified in the reference). */
void svs_scores(const float* M, int64_t n, int32_t d, const float* q, float* x)
{
    for (int64_t i = 0; i < n; ++i) {
        const float* row = M + i * (int64_t)d;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int32_t j = 0;
        for (; j + 3 < d; j += 4) {
            a0 += row[j] * q[j]; a1 += row[j + 1] * q[j + 1]; a2 += row[j + 2] * q[j + 2]; a3 += row[j + 3] * q[j + 3];
        }
        for (; j < d; ++j) a0 += row[j] * q[j];
        x[i] = (a0 + a1) + (a2 + a3);
    }
}

typedef struct { float score; int64_t index; } pair_t;

/* Python's tuple order on (float(score), int(index)) with reverse=True; NaN first (argpartition's "largest"). */
static int cmp_desc(const void* pa, const void* pb)
{
    const pair_t* a = (const pair_t*)pa; const pair_t* b = (const pair_t*)pb;
    const int an = isnan(a->score), bn = isnan(b->score);
    if (an != bn) return an ? -1 : 1;
    if (!an) {
        if (a->score > b->score) return -1;
        if (a->score < b->score) return 1;
    }
    if (a->index > b->index) return -1;
    if (a->index < b->index) return 1;
    return 0;
}

/* util.py:190-203.  Returns the number of results (min(k, n), 0 for k <= 0). */
int64_t svs_get_top_k(const float* scores, int64_t n, int64_t k, float* out_scores, int64_t* out_index)
{
    if (k > n) k = n;                                   /* util.py:198-199 */
    if (k <= 0) return 0;                               /* util.py:200-201 */
    pair_t* all = (pair_t*)malloc((size_t)n * sizeof(pair_t));
    if (!all) return -1;
    for (int64_t i = 0; i < n; ++i) { all[i].score = scores[i]; all[i].index = i; }
    qsort(all, (size_t)n, sizeof(pair_t), cmp_desc);    /* selection + final order in one total order */
    for (int64_t i = 0; i < k; ++i) { out_scores[i] = all[i].score; out_index[i] = all[i].index; }
    free(all);
    return k;
}

/* kb.py:1622-1627: x = dot(M, q); top-k; index -> embeddings.id. */
int64_t svs_superheavy(const float* M, const int64_t* emb_id_lookup, int64_t n, int32_t d, const float* q, int64_t k,
                       float* out_scores, int64_t* out_emb_ids)
{
    float* x = (float*)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    if (!x) return -1;
    svs_scores(M, n, d, q, x);
    const int64_t c = svs_get_top_k(x, n, k, out_scores, out_emb_ids);
    for (int64_t i = 0; i < c; ++i) out_emb_ids[i] = emb_id_lookup[out_emb_ids[i]];
    free(x);
    return c;
}

/* embeddings/util.py:19-23: '<{n}f' unpack.  Returns the number of floats, or -1 if nbytes % 4 != 0 (the assert). */
int64_t svs_blob_to_row(const unsigned char* blob, int64_t nbytes, float* out)
{
    if (nbytes % 4) return -1;
    for (int64_t i = 0; i < nbytes / 4; ++i) {
        const uint32_t u = (uint32_t)blob[4 * i] | ((uint32_t)blob[4 * i + 1] << 8) | ((uint32_t)blob[4 * i + 2] << 16)
                         | ((uint32_t)blob[4 * i + 3] << 24);
        memcpy(out + i, &u, 4);
    }
    return nbytes / 4;
}
