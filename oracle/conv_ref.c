/*
 * oracle/conv_ref.c -- plain-C fp64 restatement of the conv core.  TEST INFRASTRUCTURE ONLY.
 *
 * Independent (non-torch) restatement of the arithmetic of
 * torch-geometric 2.7.0 `TransformerConv(H, H/h, heads=h, edge_dim=H, beta=True)` as the
 * reference constructs and calls it (/root/reference/scripts/train.py:308,315,326,334),
 * used to cross-check the pure-torch shim under oracle/pyg_shim (neither is pinned by the
 * reference's own tests: PARITY UNPINNED, see oracle/__init__.py).
 *
 * Also restates the graph plan: a stable sort of edges by target (CSR) and by source (CSC),
 * i.e. what `torch.sort(edge_index[1], stable=True)` yields.
 *
 * Only tests/ may load the resulting oracle/_ref/libconv_ref.so.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Stable counting sort of edge ids by key[e] in [0, n_nodes).  rowptr has n_nodes+1 entries. */
int ref_stable_sort_by_key(const int64_t *key, int64_t n_edges, int64_t n_nodes, int32_t *rowptr,
                           int32_t *eid)
{
    memset(rowptr, 0, (size_t)(n_nodes + 1) * sizeof(int32_t));
    for (int64_t e = 0; e < n_edges; ++e) {
        if (key[e] < 0 || key[e] >= n_nodes) return -1;
        rowptr[key[e] + 1] += 1;
    }
    for (int64_t i = 0; i < n_nodes; ++i) rowptr[i + 1] += rowptr[i];
    int32_t *cursor = (int32_t *)malloc((size_t)(n_nodes > 0 ? n_nodes : 1) * sizeof(int32_t));
    if (!cursor) return -2;
    memcpy(cursor, rowptr, (size_t)n_nodes * sizeof(int32_t));
    for (int64_t e = 0; e < n_edges; ++e) eid[cursor[key[e]]++] = (int32_t)e;
    free(cursor);
    return 0;
}

/*
 * Message + segment softmax + aggregation.
 *   q,k,v : [n_nodes, heads*C]   e : [n_edges, heads*C]   src,dst : [n_edges]
 *   agg   : [n_nodes, heads*C] (output)
 * s = <q_i, k_j + e_ij> / sqrt(C) per head; alpha = exp(s - amax_i) / (sum_i exp(s - amax_i) + 1e-16);
 * agg_i = sum_j alpha * (v_j + e_ij).  Rows without in-edges stay 0.
 */
int ref_conv_core_fwd(const double *q, const double *k, const double *v, const double *e,
                      const int64_t *src, const int64_t *dst, int64_t n_edges, int64_t n_nodes,
                      int heads, int C, double *agg)
{
    const int64_t H = (int64_t)heads * C;
    const int64_t nh = n_nodes * heads;
    double *amax = (double *)malloc((size_t)(nh > 0 ? nh : 1) * sizeof(double));
    double *zsum = (double *)calloc((size_t)(nh > 0 ? nh : 1), sizeof(double));
    char *seen = (char *)calloc((size_t)(nh > 0 ? nh : 1), 1);
    double *logit = (double *)malloc((size_t)(n_edges * heads > 0 ? n_edges * heads : 1) * sizeof(double));
    if (!amax || !zsum || !seen || !logit) return -2;
    const double inv = 1.0 / sqrt((double)C);
    for (int64_t ed = 0; ed < n_edges; ++ed) {
        const int64_t i = dst[ed], j = src[ed];
        for (int t = 0; t < heads; ++t) {
            double s = 0.0;
            for (int c = 0; c < C; ++c) {
                const int64_t o = (int64_t)t * C + c;
                s += q[i * H + o] * (k[j * H + o] + e[ed * H + o]);
            }
            s *= inv;
            logit[ed * heads + t] = s;
            if (!seen[i * heads + t] || s > amax[i * heads + t]) {
                amax[i * heads + t] = s;
                seen[i * heads + t] = 1;
            }
        }
    }
    for (int64_t ed = 0; ed < n_edges; ++ed)
        for (int t = 0; t < heads; ++t)
            zsum[dst[ed] * heads + t] += exp(logit[ed * heads + t] - amax[dst[ed] * heads + t]);
    memset(agg, 0, (size_t)(n_nodes * H) * sizeof(double));
    for (int64_t ed = 0; ed < n_edges; ++ed) {
        const int64_t i = dst[ed], j = src[ed];
        for (int t = 0; t < heads; ++t) {
            const double a = exp(logit[ed * heads + t] - amax[i * heads + t]) / (zsum[i * heads + t] + 1e-16);
            for (int c = 0; c < C; ++c) {
                const int64_t o = (int64_t)t * C + c;
                agg[i * H + o] += a * (v[j * H + o] + e[ed * H + o]);
            }
        }
    }
    free(amax); free(zsum); free(seen); free(logit);
    return 0;
}

/* beta gate + LayerNorm + ReLU + residual (train.py:316-317 around the conv's beta-gated skip):
 *   beta = sigmoid(w[0:H].agg + w[H:2H].xr + w[2H:3H].(agg - xr));  o = beta*xr + (1-beta)*agg
 *   out  = x + relu(LN(o) * gamma + bias),  LN eps = 1e-5, biased variance. */
int ref_gate_ln_relu_res(const double *agg, const double *xr, const double *x, const double *wbeta,
                         const double *gamma, const double *bias, int64_t n_rows, int H, double *out)
{
    double *o = (double *)malloc((size_t)H * sizeof(double));
    if (!o) return -2;
    for (int64_t r = 0; r < n_rows; ++r) {
        double z = 0.0;
        for (int c = 0; c < H; ++c) {
            const double a = agg[r * H + c], s = xr[r * H + c];
            z += wbeta[c] * a + wbeta[H + c] * s + wbeta[2 * H + c] * (a - s);
        }
        const double beta = 1.0 / (1.0 + exp(-z));
        double mean = 0.0;
        for (int c = 0; c < H; ++c) {
            o[c] = beta * xr[r * H + c] + (1.0 - beta) * agg[r * H + c];
            mean += o[c];
        }
        mean /= H;
        double var = 0.0;
        for (int c = 0; c < H; ++c) var += (o[c] - mean) * (o[c] - mean);
        var /= H;
        const double rstd = 1.0 / sqrt(var + 1e-5);
        for (int c = 0; c < H; ++c) {
            const double y = (o[c] - mean) * rstd * gamma[c] + bias[c];
            out[r * H + c] = x[r * H + c] + (y > 0.0 ? y : 0.0);
        }
    }
    free(o);
    return 0;
}
