/*
 * alignn_b200.h -- C ABI of the B200-native ALIGNN message-passing hot path.
 *
 * This is the drop-in boundary: a plain C shared library (libalignn_b200.so, sm_100a) whose
 * entry points are what a binding for the reference's hot path would call.  The reference
 * (conorjmoran/gnn-elasticity-predictor) has no native code of its own: it reaches its
 * arithmetic through torch-geometric 2.7.0 (`TransformerConv`, `global_mean_pool`) from
 * `scripts/train.py`.  Each entry point below cites the reference interface whose arithmetic it
 * replaces.  INTEGRATION.md shows the ctypes stub and the module-level swap a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device, 16-byte aligned, contiguous
 *     row-major; the caller (PyTorch) owns every buffer: nothing is allocated or freed in here;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     the host, there is no global mutable state, and every call is CUDA-graph capturable;
 *   - `dtype` selects the storage type of the projected operands (q,k,v,e and their gradients):
 *     ALIGNN_F32 or ALIGNN_BF16.  Accumulation, softmax statistics, the aggregate, LayerNorm and
 *     the residual stream are always fp32;
 *   - return value: ALIGNN_OK (0), an ALIGNN_ERR_* code, or 1000 + cudaError_t.  Nothing throws.
 *
 * Graph plan (built once per batch, shared by all layers and by forward+backward)
 *   CSR ("by target"): rowptr[n_nodes+1], col[n_edges] = source of the p-th edge in stable
 *   target-sorted order, eid[n_edges] = its position in the caller's edge list.
 *   CSC ("by source"): rowptr_t, col_t (= target), eid_t -- stable source-sorted order.
 *   Indices are int32; duplicates, self loops, empty rows and arbitrary degree skew are legal.
 */
#ifndef ALIGNN_B200_H
#define ALIGNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALIGNN_ABI_VERSION 3

#define ALIGNN_F32 0
#define ALIGNN_BF16 1

#define ALIGNN_OK 0
#define ALIGNN_ERR_BAD_ARG 1      /* null pointer, negative size, misaligned buffer              */
#define ALIGNN_ERR_BAD_SHAPE 2    /* hidden % heads != 0, sizes beyond int32 indexing             */
#define ALIGNN_ERR_WORKSPACE 3    /* workspace smaller than alignn_plan_workspace_bytes()         */
#define ALIGNN_ERR_BAD_DTYPE 4
#define ALIGNN_ERR_CUDA_BASE 1000 /* 1000 + cudaError_t                                           */

int alignn_abi_version(void);
const char *alignn_error_string(int code);

/* ---- graph plan ------------------------------------------------------------------------------
 * Replaces the index handling inside PyG `MessagePassing.propagate` / `utils.softmax` /
 * `utils.scatter` as driven by `conv(x, edge_index, edge_attr)` at reference
 * scripts/train.py:315,334 (edge_index[0] = source j, edge_index[1] = target i, aggregation on
 * edge_index[1]).  Result is bit-identical to torch.sort(edge_index[k], stable=True).
 *
 * edge_index : int64 [2, n_edges] (row 0 = source, row 1 = target), as PyG stores it.
 * status     : int32[1]; set to 1 if any index lies outside [0, n_nodes) (such edges are dropped
 *              from the plan -- the caller decides whether to read the flag).
 */
size_t alignn_plan_workspace_bytes(int64_t n_edges, int64_t n_nodes);
int alignn_build_plan(const int64_t *edge_index, int64_t n_edges, int64_t n_nodes,
                      int32_t *rowptr, int32_t *col, int32_t *eid,
                      int32_t *rowptr_t, int32_t *col_t, int32_t *eid_t,
                      int32_t *status, void *workspace, size_t workspace_bytes, void *stream);

/* ---- fused edge-attention conv core, forward ----------------------------------------------------
 * Replaces `TransformerConv.message` + `utils.softmax` + 'add' aggregation (PyG 2.7.0; reference
 * call sites scripts/train.py:315,334):
 *   s_ij,t = <q_i,t , k_j,t + e_ij,t> / sqrt(C);  a = exp(s - max_i) / (sum_i exp(s - max_i) + 1e-16)
 *   a~ = dropout(a, p);  agg_i = sum_j a~_ij,t (v_j,t + e_ij,t)           (0 for rows w/o in-edges)
 * One pass over the target-sorted edge list, online softmax, no atomics, deterministic.
 *
 * q,k,v : [n_nodes, hidden] dtype;  e : [n_edges, hidden] dtype, rows in the CALLER's edge order
 * agg   : [n_nodes, hidden] f32 (out);  stat_m, stat_z : [n_nodes, heads] f32 (out; running max in
 *         log2 units and the softmax denominator, saved for backward)
 * p_drop in [0,1): attention dropout probability (0 = eval); (seed, offset) key the Philox stream.
 */
int alignn_conv_fwd(const void *q, const void *k, const void *v, const void *e,
                    const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                    float *agg, float *stat_m, float *stat_z,
                    int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                    float p_drop, uint64_t seed, uint64_t offset, void *stream);

/* ---- conv core, backward ------------------------------------------------------------------------
 * Autograd of the above (the reference gets it from torch autograd through the PyG ops;
 * scripts/train.py:691,697 `.backward()`).  Two passes, no atomics:
 *   target-sorted pass: recompute a from (stat_m, stat_z); dq, de, per-edge (a~, ds/sqrt(C)) -> coef
 *   source-sorted pass: dk_j = sum_i ds_ij q_i / sqrt(C);  dv_j = sum_i a~_ij dagg_i
 * dagg, agg : [n_nodes, hidden] f32;  dq,dk,dv : [n_nodes, hidden] dtype;  de : [n_edges, hidden] dtype
 * coef : [n_edges, 2*heads] f32 scratch owned by the caller.
 */
int alignn_conv_bwd(const float *dagg, const float *agg,
                    const void *q, const void *k, const void *v, const void *e,
                    const float *stat_m, const float *stat_z,
                    const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                    const int32_t *rowptr_t, const int32_t *col_t, const int32_t *eid_t,
                    void *dq, void *dk, void *dv, void *de, float *coef,
                    int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                    float p_drop, uint64_t seed, uint64_t offset, void *stream);

/* ---- beta-gated skip + LayerNorm + ReLU + dropout + residual --------------------------------------
 * Replaces the tail of `TransformerConv.forward` (beta = sigmoid(lin_beta([agg, x_r, agg - x_r]));
 * out = beta x_r + (1-beta) agg) and `x + Dropout(ReLU(LayerNorm(out)))` of
 * EdgeUpdateBlock / NodeUpdateBlock (reference scripts/train.py:316-317, 335-336).
 * agg : f32;  xr : dtype (skip projection);  x, y : f32 [n_rows, hidden] (residual in / block out);
 * y_lp : optional dtype copy of y (may be NULL);  wbeta : f32 [3*hidden];  gamma, bias : f32 [hidden]
 * beta, mean, rstd : f32 [n_rows] saved for backward.
 */
int alignn_gate_ln_fwd(const float *agg, const void *xr, const float *x,
                       const float *wbeta, const float *gamma, const float *bias,
                       float *y, void *y_lp, float *beta, float *mean, float *rstd,
                       int64_t n_rows, int hidden, int dtype, float eps,
                       float p_drop, uint64_t seed, uint64_t offset, void *stream);

/* Backward of the above w.r.t. agg, xr (dx = dy is the caller's residual pass-through) and the
 * parameters.  partials : f32 [alignn_gate_ln_bwd_partial_rows() * 5 * hidden] scratch;
 * dparams : f32 [5 * hidden] = (d wbeta[0:H], d wbeta[H:2H], d wbeta[2H:3H], d gamma, d bias),
 * reduced in a fixed order (deterministic). */
int64_t alignn_gate_ln_bwd_partial_rows(void);
int alignn_gate_ln_bwd(const float *dy, const float *agg, const void *xr,
                       const float *wbeta, const float *gamma, const float *bias,
                       const float *beta, const float *mean, const float *rstd,
                       float *dagg, void *dxr, float *partials, float *dparams,
                       int64_t n_rows, int hidden, int dtype,
                       float p_drop, uint64_t seed, uint64_t offset, void *stream);

/* ---- per-graph mean pooling -------------------------------------------------------------------------
 * Replaces `global_mean_pool(node_state, data.batch)` (reference scripts/train.py:388,562):
 * pooled_g = sum_{n in graph g} x_n / max(count_g, 1).  (rowptr, eid) is a plan over `batch`
 * (key = graph id), so unsorted `batch` vectors are handled too.
 */
int alignn_segment_mean_fwd(const float *x, const int32_t *rowptr, const int32_t *eid, float *pooled,
                            int64_t n_graphs, int hidden, void *stream);
int alignn_segment_mean_bwd(const float *dpooled, const int32_t *rowptr, const int32_t *eid, float *dx,
                            int64_t n_graphs, int hidden, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ALIGNN_B200_H */
