/* ORACLE (test infrastructure only -- never linked into the product library).
 *
 * Plain-C restatement of the reference's only native component,
 *   /root/reference/torchreid/metrics/rank_cylib/rank_cy.pyx:154-249  (eval_market1501_cy + function_cumsum)
 * with the numpy call it makes (np.argsort, rank_cy.pyx:165) replaced by a stable merge argsort
 * (ties -> ascending gallery index; SURVEY.md section 7 hard part 5).
 * Accumulators are `float` exactly like the Cython `float` cdefs (rank_cy.pyx:169-183,230-239).
 * Single-threaded by construction, like the reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void merge_argsort(const float *key, int64_t *idx, int64_t *tmp, long n) {
    for (long i = 0; i < n; ++i) idx[i] = i;
    for (long width = 1; width < n; width *= 2) {
        for (long lo = 0; lo < n; lo += 2 * width) {
            long mid = lo + width < n ? lo + width : n;
            long hi = lo + 2 * width < n ? lo + 2 * width : n;
            long a = lo, b = mid, o = lo;
            while (a < mid && b < hi) {
                /* stable: take from the left run unless right is strictly smaller */
                if (key[idx[b]] < key[idx[a]]) tmp[o++] = idx[b++];
                else tmp[o++] = idx[a++];
            }
            while (a < mid) tmp[o++] = idx[a++];
            while (b < hi) tmp[o++] = idx[b++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int64_t));
    }
}

/* returns 0 on success, 1 when no query is valid (the reference asserts, rank_cy.pyx:227) */
int oracle_eval_market1501(const float *distmat, const int64_t *q_pids, const int64_t *g_pids,
                           const int64_t *q_camids, const int64_t *g_camids,
                           long num_q, long num_g, long max_rank,
                           float *avg_cmc /* [max_rank] */, float *mAP_out, int64_t *num_valid_out) {
    if (num_g < max_rank) max_rank = num_g;
    int64_t *order = (int64_t *)malloc((size_t)num_g * sizeof(int64_t));
    int64_t *tmp = (int64_t *)malloc((size_t)num_g * sizeof(int64_t));
    float *raw_cmc = (float *)malloc((size_t)num_g * sizeof(float));
    float *cmc = (float *)malloc((size_t)num_g * sizeof(float));
    float *all_cmc = (float *)calloc((size_t)num_q * (size_t)max_rank, sizeof(float));
    float *all_AP = (float *)calloc((size_t)num_q, sizeof(float));
    float num_valid_q = 0.f;

    for (long q = 0; q < num_q; ++q) {
        const int64_t q_pid = q_pids[q], q_camid = q_camids[q];
        merge_argsort(distmat + (size_t)q * (size_t)num_g, order, tmp, num_g);
        long num_g_real = 0;
        int meet = 0;
        for (long g = 0; g < num_g; ++g) {
            const int64_t gi = order[g];
            if (g_pids[gi] != q_pid || g_camids[gi] != q_camid) {
                const float match = (g_pids[gi] == q_pid) ? 1.f : 0.f;
                raw_cmc[num_g_real++] = match;
                if (match > 1e-31f) meet = 1;
            }
        }
        if (!meet) continue;
        /* function_cumsum + clip (rank_cy.pyx:208-211) */
        cmc[0] = raw_cmc[0];
        for (long g = 1; g < num_g_real; ++g) cmc[g] = raw_cmc[g] + cmc[g - 1];
        /* AP from the unclipped cumsum (rank_cy.pyx:219-225) */
        float num_rel = 0.f, tmp_sum = 0.f;
        for (long g = 0; g < num_g_real; ++g) {
            /* Cython: float / (long + 1.) promotes to double; the float accumulator is rounded once per term */
            tmp_sum = (float)((double)tmp_sum + ((double)cmc[g] / ((double)g + 1.)) * (double)raw_cmc[g]);
            num_rel += raw_cmc[g];
        }
        /* NB: entries beyond num_g_real keep stale values in the reference's reused buffer; with
         * max_rank <= num_g_real (always true for market-scale galleries) they are never read. */
        for (long r = 0; r < max_rank; ++r) {
            float v = r < num_g_real ? cmc[r] : (num_g_real > 0 ? cmc[num_g_real - 1] : 0.f);
            all_cmc[(size_t)q * (size_t)max_rank + r] = v > 1.f ? 1.f : v;
        }
        all_AP[q] = tmp_sum / num_rel;
        num_valid_q += 1.f;
    }
    int rc = 0;
    if (num_valid_q > 0.f) {
        for (long r = 0; r < max_rank; ++r) {
            float s = 0.f;
            for (long q = 0; q < num_q; ++q) s += all_cmc[(size_t)q * (size_t)max_rank + r];
            avg_cmc[r] = s / num_valid_q;
        }
        float m = 0.f;
        for (long q = 0; q < num_q; ++q) m += all_AP[q];
        *mAP_out = m / num_valid_q;
    } else {
        rc = 1;
    }
    *num_valid_out = (int64_t)num_valid_q;
    free(order); free(tmp); free(raw_cmc); free(cmc); free(all_cmc); free(all_AP);
    return rc;
}
