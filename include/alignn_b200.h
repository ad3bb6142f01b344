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

#define ALIGNN_ABI_VERSION 22

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
 * status     : int32[1] bit flags; bit 0 (value 1) is set if any index lies outside [0, n_nodes) (such edges are dropped
 *              from the plan -- the caller decides whether to read the flag).  Dropped edges sort after
 *              rowptr[n_nodes] in input order; eid / eid_t stay permutations of [0, n_edges) and col / col_t
 *              hold 0 there.  Shape-bucket padding (batching.py) relies on this: padded edges carry the
 *              index -1 and are never visited by any kernel.
 */
size_t alignn_plan_workspace_bytes(int64_t n_edges, int64_t n_nodes);
int alignn_build_plan(const int64_t *edge_index, int64_t n_edges, int64_t n_nodes,
                      int32_t *rowptr, int32_t *col, int32_t *eid,
                      int32_t *rowptr_t, int32_t *col_t, int32_t *eid_t,
                      int32_t *status, void *workspace, size_t workspace_bytes, void *stream);

/* Same plan with hints.  ALIGNN_PLAN_SOURCE_SORTED: the caller states that edge_index[0] is already
 * non-decreasing -- true for every graph `fetch.py:389-396,421-444` emits (bonds and angles are generated
 * source-major) and for any PyG collate of such graphs -- so the stable source sort is the identity and the
 * CSC half of the plan needs no radix passes.  The hint is VERIFIED on the device: if the sort keys are not
 * non-decreasing, bit 1 (value 2) of `status` is set and the CSC half is invalid. */
#define ALIGNN_PLAN_SOURCE_SORTED 1
int alignn_build_plan_ex(const int64_t *edge_index, int64_t n_edges, int64_t n_nodes,
                         int32_t *rowptr, int32_t *col, int32_t *eid,
                         int32_t *rowptr_t, int32_t *col_t, int32_t *eid_t,
                         int32_t *status, void *workspace, size_t workspace_bytes, int flags, void *stream);

/* The same with a KEY BOUND: the caller states that every index is < key_bound (<= n_nodes) -- e.g. the batch's
 * `lg_active_rows`: with PyG's default collate every line-graph index is below N_atoms + E_last -- so the
 * radix sorts run over log2(key_bound) bits instead of log2(n_nodes) (2 passes instead of 3 at BASELINE config 2).  Rows at
 * or beyond the bound come out empty.  Verified on the device: a legal index >= key_bound sets bit 2 (value 4) of `status`
 * and the edge is dropped like an out-of-range one.  key_bound < 0 or > n_nodes means n_nodes. */
int alignn_build_plan_bounded(const int64_t *edge_index, int64_t n_edges, int64_t n_nodes, int64_t key_bound,
                              int32_t *rowptr, int32_t *col, int32_t *eid,
                              int32_t *rowptr_t, int32_t *col_t, int32_t *eid_t,
                              int32_t *status, void *workspace, size_t workspace_bytes, int flags, void *stream);

/* pos_t[p] = position in the target-sorted (CSR) order of the p-th edge of the source-sorted (CSC) order: lets the
 * source-sorted backward pass read per-edge records the target-sorted pass wrote in CSR order (`coef`) without an edge-id
 * indirection.  eid / eid_t: the plan's permutations; scratch: int32 [n_edges]. */
int alignn_plan_csc_positions(const int32_t *eid, const int32_t *eid_t, int64_t n_edges, int32_t *scratch,
                              int32_t *pos_t, void *stream);

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

/* ==== streaming path: edge attention with LINEAR edge features (hidden = 256, heads in {1,2,4}) ============
 * Same reference arithmetic as alignn_conv_fwd/bwd, plus the per-edge dense projections that feed it
 * (`lin_edge` of PyG TransformerConv, reference scripts/train.py:308,326; the second Linear of `angle_encoder`,
 * :360-364; `edge_proj`, :324,333).  With per-edge features f and e = Wc f + c the [E,H] tensor e is never
 * materialised: logits use qt_i,t = Wc[t]^T q_i,t (a per-NODE projection), the aggregate is returned as
 *   aggv_i,t = sum_j a~ v_j,t ,  abar_i,t = sum_j a~ f_ij  (then agg = aggv + Wc[t] abar_t + c_t S_t per node),
 * and backward returns bbar_i,t = sum_j ds f_ij / sqrt(C) and the feature gradient
 *   df_ij = sum_t ( ds_ij,t qt_i,t / sqrt(C) + a~_ij,t gt_i,t ),  gt_i,t = Wc[t]^T dagg_i,t,
 * optionally accumulated onto df_in (sum over layers sharing f) and ReLU-masked by (f > 0).
 * q,k,v are rows of strided buffers (ld* in elements, multiples of 8); qt, gt, abar, bbar are [heads, Nn, 256].
 */
int alignn_edgeattn_supported(int hidden, int heads);
int alignn_edgeattn_fwd(const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                        const void *qt, const void *feat,
                        const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                        float *aggv, void *abar, float *stat_m, float *stat_z, float *stat_s,
                        int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                        float p_drop, uint64_t seed, uint64_t offset, void *stream);
int alignn_edgeattn_bwd_dst(const float *dagg, const float *agg,
                            const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                            const void *qt, const void *gt, const float *cvec,
                            const void *feat, const float *stat_m, const float *stat_z,
                            const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                            void *dq, int64_t lddq, void *bbar, float *coef,
                            const void *df_in, void *df_out, int relu_mask,
                            int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                            float p_drop, uint64_t seed, uint64_t offset, void *stream);
int alignn_edgeattn_bwd_src(const float *dagg, const void *q, int64_t ldq, const float *coef,
                            const int32_t *rowptr_t, const int32_t *col_t, const int32_t *eid_t,
                            void *dk, void *dv, int64_t ldd, int64_t n_nodes, int64_t n_edges,
                            int hidden, int heads, int dtype, void *stream);

/* alignn_edgeattn_bwd_src with the upstream gradient in storage dtype (bf16 only): one third less gather traffic. */
int alignn_edgeattn_bwd_src_lp(const void *dagg_lp, const void *q, int64_t ldq, const float *coef,
                               const int32_t *rowptr_t, const int32_t *col_t, const int32_t *eid_t,
                               void *dk, void *dv, int64_t ldd, int64_t n_nodes, int64_t n_edges,
                               int hidden, int heads, int dtype, void *stream);

/* Tensor-core variants of the streaming kernels (hidden = 256, heads = 4, bf16 storage): identical contract to
 * alignn_edgeattn_fwd / _bwd_dst, with the per-edge contractions on mma.sync (chunks of 16 in-edges of one target
 * row) and attention dropout keyed by the edge's CSR position (so a forward/backward pair must use the same
 * family of kernels). */
int alignn_edgeattn_mma_supported(int hidden, int heads, int dtype);
int alignn_edgeattn_mma_fwd(const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                            const void *qt, const void *feat,
                            const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                            float *aggv, void *abar, float *stat_m, float *stat_z, float *stat_s,
                            int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                            float p_drop, uint64_t seed, uint64_t offset, void *stream);

int alignn_edgeattn_mma_bwd_dst(const float *dagg, const void *dagg_lp, const float *agg,
                                const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                                const void *qt, const void *gt, const float *cvec,
                                const void *feat, const float *stat_m, const float *stat_z,
                                const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                                void *dq, int64_t lddq, void *bbar, float *coef,
                                const void *df_in, void *df_out, int relu_mask,
                                int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                                float p_drop, uint64_t seed, uint64_t offset, void *stream);

/* Strided variants: qt / gt / abar / bbar addressed as base[row * ld + head * hs + channel] (e.g. column blocks of a
 * fused projection), optional device dropout counter `rng_step` added to `offset`. */
int alignn_edgeattn_mma_fwd_s(const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                              const void *qt, int64_t ldqt, int64_t hsqt, const void *feat,
                              const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                              float *aggv, void *abar, int64_t ldab, int64_t hsab,
                              float *stat_m, float *stat_z, float *stat_s,
                              int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                              float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream);
int alignn_edgeattn_mma_bwd_dst_s(const float *dagg, const void *dagg_lp, const float *agg,
                                  const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                                  const void *qt, int64_t ldqt, int64_t hsqt, const void *gt, int64_t ldgt, int64_t hsgt,
                                  const float *cvec, const void *feat, const float *stat_m, const float *stat_z,
                                  const int32_t *rowptr, const int32_t *col, const int32_t *eid,
                                  void *dq, int64_t lddq, void *bbar, int64_t ldbb, int64_t hsbb, float *coef,
                                  const void *df_in, void *df_out, int64_t lddf, int relu_mask,
                                  int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                                  float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream);

/* ---- line-graph attention with the angle embedding recomputed in-kernel (hidden 256, 4 heads, bf16) ---------------
 * Replaces `angle_encoder(lg_edge_attr)` (reference scripts/train.py:360-364, 554) + the line-graph TransformerConv
 * of EdgeUpdateBlock (train.py:308, 315): f_ij = h1 = relu(W1 a_ij + b1) is rebuilt from the packed angle features
 * inside the kernel instead of being read as a [L, 256] tensor (see csrc/lgattn.cu).  Same outputs as
 * alignn_edgeattn_fwd; qt / abar are addressed as base[row * ld + head * hs + channel]; `rng_step` (optional device
 * counter) is added to `offset` so that a captured CUDA graph draws fresh dropout masks on every replay.
 *   alignn_lg_pack_angles : a_csr[p] = (bf16(a[eid[p], :]), 1, 0...)  -- [L, 16] bf16 in target-sorted order.
 * `work`: 8 bytes of device memory, ZERO on entry; the kernels leave it zero on exit (dynamic row-range scheduler: a
 * persistent grid of one CTA per SM whose warps draw units of ~128 edges from this counter, so rows of very different
 * in-degree balance; every row is still reduced by one warp in CSR order, results do not depend on the schedule).
 * One buffer per stream: launches that may run concurrently must not share it. */
int alignn_lgattn_supported(int hidden, int heads, int in_dim, int dtype);
int alignn_lg_pack_angles(const float *a, const int32_t *eid, void *a_csr, int64_t n_edges, int in_dim, void *stream);
int alignn_lgattn_fwd(const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                      const void *qt, int64_t ldqt, int64_t hsqt,
                      const void *a_csr, const float *w1, const float *b1, int in_dim,
                      const int32_t *rowptr, const int32_t *col,
                      float *aggv, void *abar, int64_t ldab, int64_t hsab,
                      float *stat_m, float *stat_z, float *stat_s,
                      int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                      float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *work, void *stream);

/* The same forward on the 5th-generation tensor cores (csrc/lgattn_tc.cu): 128-angle tiles, every contraction a
 * tcgen05.mma with M = 128 (H1 = A W1^T; logits = [H1 | K] [QT ; Qbd]^T; [abar | agg]^T = [H1 | V]^T P), accumulators
 * in TMEM, operands in 128-byte-swizzled shared memory, K / V rows gathered with cp.async.  Same arguments, outputs,
 * dropout masks and statistics as alignn_lgattn_fwd (no `work` counter: static row-range partition). */
int alignn_lgattn_fwd_tc(const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                         const void *qt, int64_t ldqt, int64_t hsqt,
                         const void *a_csr, const float *w1, const float *b1, int in_dim,
                         const int32_t *rowptr, const int32_t *col,
                         float *aggv, void *abar, int64_t ldab, int64_t hsab,
                         float *stat_m, float *stat_z, float *stat_s,
                         int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                         float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream);

/* Backward of alignn_lgattn_fwd w.r.t. q (dq), qt (bbar) and -- through coef, consumed by alignn_edgeattn_bwd_src with
 * the plan's CSC->CSR position map in place of eid_t -- k, v.  coef: f32 [L, 8] in target-sorted order, (a~_0..3,
 * ds_0..3) per angle.  No [L, 256] feature gradient is produced: */
int alignn_lgattn_bwd_dst(const float *dagg, const void *dagg_lp, const float *agg,
                          const void *q, const void *k, const void *v, int64_t ldq, int64_t ldk, int64_t ldv,
                          const void *qt, int64_t ldqt, int64_t hsqt, const void *gt, int64_t ldgt, int64_t hsgt,
                          const float *cvec, const void *a_csr, const float *w1, const float *b1, int in_dim,
                          const float *stat_m, const float *stat_z, const int32_t *rowptr, const int32_t *col,
                          void *dq, int64_t lddq, void *bbar, int64_t ldbb, int64_t hsbb, float *coef,
                          int64_t n_nodes, int64_t n_edges, int hidden, int heads, int dtype,
                          float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *work, void *stream);
/* ... the gradient of the first angle-encoder Linear (reference scripts/train.py:360) is formed once per step from the
 * coefficients of all (<= 4 per call) line-graph layers: out[r*256 + c] = dW1[c, r] (r < in_dim), out[in_dim*256 + c] =
 * db1[c].  coef / qt / gt are HOST arrays of n_layers device pointers; partials: f32
 * [alignn_lg_angle_grad_partial_floats()] scratch.  Deterministic (fixed-order reductions). */
int64_t alignn_lg_angle_grad_partial_floats(int64_t n_nodes, int64_t n_edges);
int alignn_lg_angle_grad(const void *a_csr, const float *w1, const float *b1, int in_dim,
                         const int32_t *rowptr, int n_layers, const float *const *coef,
                         const void *const *qt, const void *const *gt,
                         int64_t ldqt, int64_t hsqt, int64_t ldgt, int64_t hsgt,
                         float *partials, float *out, int64_t n_nodes, int64_t n_edges, void *stream);
/* the same with a cap on the grid (max_blocks > 0: at most that many CTAs, one per SM, so that concurrent streams keep the
 * remaining SMs -- the engine runs the tail of the backward beside this kernel); same partials buffer. */
int alignn_lg_angle_grad2(const void *a_csr, const float *w1, const float *b1, int in_dim,
                         const int32_t *rowptr, int n_layers, const float *const *coef,
                         const void *const *qt, const void *const *gt,
                         int64_t ldqt, int64_t hsqt, int64_t ldgt, int64_t hsgt,
                         float *partials, float *out, int64_t n_nodes, int64_t n_edges, int max_blocks, void *stream);

/* ---- fused global-norm clip + AdamW over one flat parameter bucket (reference scripts/train.py:690-699, 1516-1540) --
 * params/grads/exp_avg/exp_avg_sq: f32 [n]; shadow_bf16: optional bf16 copy of the updated parameters; partials:
 * f32 [alignn_adamw_partial_floats()]; step: f32[1] device step count (advanced here); lr2: f32[2] device learning
 * rates of [0, split) and [split, n); norm_out: optional f32[1] total gradient norm; grad_scale multiplies the raw
 * gradient (loss-scale / world-size factors).  max_norm <= 0 disables clipping. */
int64_t alignn_adamw_partial_floats(void);
int alignn_clip_adamw_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                           void *shadow_bf16, float *partials, float *step, const float *lr2,
                           float *norm_out, int64_t n, int64_t split,
                           float beta1, float beta2, float eps, float weight_decay, float max_norm,
                           float grad_scale, void *stream);

/* Epilogue variants for the streaming path: the aggregate arrives in parts (aggv f32 [rows,H]; agge storage dtype
 * [heads, rows, C]; c_t * S_t), xr / dxr are strided column slices; agg_out receives the assembled aggregate (saved
 * for backward), dagg_lp a storage-dtype copy of dagg for the gt projection; rng_step: optional device counter added
 * to the dropout offset (hidden in {8..256} dividing 256 only). */
int alignn_gate_ln_fwd2(const float *aggv, const void *agge, const float *cvec, const float *stat_s,
                        int heads, const void *xr, int64_t ldxr, const float *x,
                        const float *wbeta, const float *gamma, const float *bias,
                        float *agg_out, float *y, void *y_lp, float *beta, float *mean, float *rstd,
                        int64_t n_rows, int hidden, int dtype, float eps,
                        float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream);
int alignn_gate_ln_bwd2(const float *dy, const float *agg, const void *xr, int64_t ldxr,
                        const float *wbeta, const float *gamma, const float *bias,
                        const float *beta, const float *mean, const float *rstd,
                        float *dagg, void *dagg_lp, void *dxr, int64_t lddxr, float *partials, float *dparams,
                        int64_t n_rows, int hidden, int dtype,
                        float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream);

/* alignn_gate_ln_fwd2 for batches whose trailing rows are ISOLATED (no in-edges): rows >= agg_rows take agg = 0 without
 * reading aggv / agge / stat_s (agge is [heads, agg_rows, C]) and agg_out is written for rows < agg_rows only.  With
 * PyG's default collate the reference offsets lg_edge_index by atoms, so for B > 1 most bond rows of the line graph
 * are isolated; agg_rows < 0 means all rows.  alignn_gate_ln_bwd3 takes the same bound (dagg / dagg_lp
 * are written for rows < agg_rows only). */
int alignn_gate_ln_fwd3(const float *aggv, const void *agge, const float *cvec, const float *stat_s,
                        int heads, int64_t agg_rows, const void *xr, int64_t ldxr, const float *x,
                        const float *wbeta, const float *gamma, const float *bias,
                        float *agg_out, float *y, void *y_lp, float *beta, float *mean, float *rstd,
                        int64_t n_rows, int hidden, int dtype, float eps,
                        float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream);
/* alignn_gate_ln_fwd3 with the residual input optionally in the STORAGE dtype: when `x` is NULL the kernel reads `x_lp`
 * ([n_rows, hidden], same dtype as xr) instead -- the first block of each chain, whose input is the bf16 encoder output
 * (reference: `edge_state + Dropout(...)` promotes bf16 + fp32 -> fp32, train.py:317), needs no fp32 copy of it. */
int alignn_gate_ln_fwd4(const float *aggv, const void *agge, const float *cvec, const float *stat_s,
                        int heads, int64_t agg_rows, const void *xr, int64_t ldxr, const float *x, const void *x_lp,
                        const float *wbeta, const float *gamma, const float *bias,
                        float *agg_out, float *y, void *y_lp, float *beta, float *mean, float *rstd,
                        int64_t n_rows, int hidden, int dtype, float eps,
                        float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream);

/* As alignn_gate_ln_bwd2, plus the gradient of the folded edge-projection bias c: dparams[5*hidden + ch] =
 * sum_rows dagg[row, ch] * stat_s[row, head(ch)] (the `c_t * S_t` term of the aggregate; c = W_e b of the Linear folded
 * into lin_edge, reference train.py:324,333 / :360-364).  partials: [partial_rows * 6 * hidden], dparams: [6 * hidden];
 * hidden in {8..256} dividing 256 only.  The upstream gradient is dy (f32, may be null = 0) + dy2 (storage dtype, row
 * stride lddy2, may be null): the second addend is the feature gradient the atom-graph conv sends to the bond states
 * (autograd's sum over the two consumers of `edge_state`, train.py:559-560), added on load instead of in its own pass. */
int alignn_gate_ln_bwd3(const float *dy, const void *dy2, int64_t lddy2, const float *agg, const void *xr, int64_t ldxr,
                        const float *wbeta, const float *gamma, const float *bias,
                        const float *beta, const float *mean, const float *rstd,
                        const float *stat_s, int heads, int64_t agg_rows,
                        float *dagg, void *dagg_lp, void *dxr, int64_t lddxr, float *partials, float *dparams,
                        int64_t n_rows, int hidden, int dtype,
                        float p_drop, uint64_t seed, uint64_t offset, const uint64_t *rng_step, void *stream);

/* Column sums out[c] = sum_r x[r * ld + c], c < width: the bias gradients of the stacked node projections
 * (lin_query/key/value/skip of PyG TransformerConv, constructed at reference train.py:308,326; torch autograd's
 * `grad.sum(0)`).  Deterministic two-stage reduction; partials: f32 [alignn_colsum_partial_floats(width)]. */
int alignn_colsum_supported(int width);
int64_t alignn_colsum_partial_floats(int width);
int alignn_colsum(const void *x, int64_t ld, int64_t n_rows, int width, int dtype, float *partials, float *out,
                  void *stream);

/* First angle-encoder layer: h1 = relu(W1 a + b1) (reference scripts/train.py:360-362 applied at :554) and its
 * parameter gradients from the ReLU-masked feature gradient dpre: out[f*256 + c] = dW1[c,f] (f < in_dim),
 * out[in_dim*256 + c] = db1[c].  in_dim <= 16, hidden = 256. */
int alignn_angle_supported(int in_dim, int hidden);
int alignn_angle_h1_fwd(const float *a, const float *w1, const float *b1, void *h1, int64_t n_edges,
                        int in_dim, int hidden, int dtype, void *stream);
int64_t alignn_angle_partial_floats(int in_dim);
int alignn_angle_h1_bwd(const void *dpre, const float *a, float *partials, float *out, int64_t n_edges,
                        int in_dim, int hidden, int dtype, void *stream);

/* ---- device-resident dataset: collate ---------------------------------------------------
 * Replaces `PtGraphDataset.__getitem__` (reference scripts/train.py:130-172: one `torch.load` per sample per
 * epoch) + the host collate of `torch_geometric.loader.DataLoader` (train.py:2037) for batches drawn from a
 * store of graphs that already lives in HBM: all graphs concatenated field by field (the `Data` fields of
 * scripts/fetch.py:614-651), indices LOCAL to each graph, `*_ptr` the per-graph row offsets.
 *
 * alignn_collate writes the batch PyG's default `Batch.from_data_list` would produce for the graphs
 * `sel[0..n_sel)`, in that order: feature rows concatenated, `edge_index` shifted by the running ATOM count,
 * `lg_edge_index` shifted by the running atom count too (`lg_inc_bonds = 0`: PyG's default `__inc__`, what the
 * reference trains on -- SURVEY.md A9) or by the running BOND count (`lg_inc_bonds = 1`), `batch[n]` = position
 * of the atom's graph in `sel`, `train_idx[g] = sel[g]`, per-graph rows gathered.  Output sizes in `out` may
 * exceed the selection's totals (shape buckets): the tail is filled exactly like batching.pad_batch -- zero
 * feature rows, index -1, padded atoms in graph `n_sel`, `y = 1`, `train_idx = -1`.  Any `out` pointer may be
 * NULL (field skipped).  Bit-exact (copies and int64 adds only).
 *
 * seg_ptr : int64 [3, n_sel + 1] device scratch; on return the running atom / bond / angle counts.
 * totals  : HOST int64[3]: the selection's atom / bond / angle totals (the caller knows them from its host copy
 *           of the offsets -- it sized `out` with them); they size the launches and are re-derived on the device.
 * maxima  : HOST int64[3]: upper bounds of the per-graph atom / bond / angle counts in the selection (grid sizing).
 * status  : int32[1] bit flags: 1 = an entry of `sel` outside [0, n_graphs), 2 = the device-side totals differ from
 *           `totals`, 4 = a graph larger than `maxima` (results still correct).  With 1 or 2 nothing is written. */
typedef struct alignn_graph_store {
    const float *x, *edge_attr, *lg_edge_attr, *global_x, *sg_one_hot, *y;
    const int64_t *edge_index;    /* [2, n_bonds]  (row 1 at + n_bonds)  */
    const int64_t *lg_edge_index; /* [2, n_angles] (row 1 at + n_angles) */
    const int64_t *node_ptr, *bond_ptr, *angle_ptr; /* [n_graphs + 1] */
    int64_t n_graphs, n_nodes, n_bonds, n_angles;
    int32_t node_dim, edge_dim, angle_dim, global_dim, sg_dim, target_dim;
} alignn_graph_store;

typedef struct alignn_batch_out {
    float *x, *edge_attr, *lg_edge_attr, *global_x, *sg_one_hot, *y;
    int64_t *edge_index, *lg_edge_index, *batch, *train_idx;
    int64_t n_graphs, n_nodes, n_bonds, n_angles; /* allocated sizes (>= the selection's totals) */
} alignn_batch_out;

int alignn_collate(const alignn_graph_store *store, const int64_t *sel, int64_t n_sel, int lg_inc_bonds,
                   const alignn_batch_out *out, int64_t *seg_ptr, const int64_t *totals, const int64_t *maxima,
                   int32_t *status, void *stream);

/* ---- bond features and line graph on the device ---------------------------------------
 * Replaces the two loops of `build_graph_from_structure` (reference scripts/fetch.py:385-396 bonds,
 * :417-447 line graph) with their helpers `_edge_geom` (:250-263), `_angle_between_vectors` (:266-273) and
 * `_rbf_expand` (:311-316).  Arithmetic in float64 in the reference's operation order, stored as float32 /
 * int64 like `to_pyg_data` (:629-633).  Bonds are the reference's directed `(i, j, jimage)` list, i-major
 * (:189-207), for one structure or for many concatenated ones (global atom ids; `atom_graph[a]` = structure of
 * atom a selects its lattice; NULL = one structure).
 *
 * alignn_bond_features  : dirv [E,3] f64 (unit i->j, 0 for zero-length bonds), edge_attr [E, n_rbf + 4] f32 =
 *                         exp(-gamma (dist - c_k)^2) | |EN_i - EN_j| | dirv.
 * alignn_linegraph_count: counts[e1] = number of bonds (j -> k) leaving j = dst(e1) other than e1's exact reverse
 *                         image; `out_ptr [A+1]` = first bond leaving each atom.
 * alignn_linegraph_fill : with angle_ptr = exclusive scan of counts ([E+1]): lg_edge_index [2, L] (e1-major, then
 *                         bond order -- the reference's emission order; ids minus graph_bond_ptr[structure] when
 *                         that pointer is given, i.e. LOCAL ids) and lg_edge_attr [L, n_ang + 3] =
 *                         exp(-gamma (angle - c_k)^2) | angle | cos | sin, angle at j between j->i and j->k. */
int alignn_bond_features(const double *frac, const double *lattice, const int64_t *atom_graph, const double *en,
                         const int64_t *bond_src, const int64_t *bond_dst, const int32_t *bond_image, int64_t n_bonds,
                         const double *rbf_centers, int n_rbf, double rbf_gamma, double *dirv, float *edge_attr,
                         void *stream);
int alignn_linegraph_count(const int64_t *bond_src, const int64_t *bond_dst, const int32_t *bond_image,
                           const int64_t *out_ptr, int64_t n_bonds, int64_t *counts, void *stream);
int alignn_linegraph_fill(const double *frac, const double *lattice, const int64_t *atom_graph,
                          const int64_t *graph_bond_ptr, const int64_t *bond_src, const int64_t *bond_dst,
                          const int32_t *bond_image, const int64_t *out_ptr, const double *dirv,
                          const int64_t *angle_ptr, int64_t n_bonds, const double *angle_centers, int n_ang,
                          double angle_gamma, int64_t *lg_edge_index, int64_t n_angles, float *lg_edge_attr,
                          void *stream);

/* ---- ensemble post-processing ------------------------------------------------------------------
 * Replaces the tail of the member loop of `ensemble_collect` (reference scripts/train.py:876-894, 903; same arithmetic at
 * scripts/predict.py:604-623 and scripts/evaluate.py:244-261), `apply_conformal_intervals` (train.py:1053-1076) and
 * `LogTransformer.inverse_transform_tensor` (train.py:281-296) by one kernel over the stacked member outputs.
 * member_means / member_logvars : f32 [n_members, n_graphs, n_targets] (logvars NULL = homoscedastic members, var_z =
 *     mean mu^2 - mean_z^2).  q : f32 [n_targets] conformal quantile (NULL = no interval), scaled = 1 for the "scaled"
 *     method (interval = q * std_z), 0 for "absolute".  log_means / log_stds : f32 [n_targets] of the fitted LogTransformer
 *     (both NULL = outputs stay in z-space).  var_z, std_z, mean_orig, lower_orig, upper_orig may be NULL. */
int alignn_ensemble_post(const float *member_means, const float *member_logvars, int n_members, int64_t n_graphs,
                         int n_targets, float min_logvar_floor, const float *q, int scaled,
                         const float *log_means, const float *log_stds, float *mean_z, float *var_z, float *std_z,
                         float *mean_orig, float *lower_orig, float *upper_orig, void *stream);

/* ---- training loss ----------------------------------------------------------------------------
 * Replaces the loss arithmetic of `train_epoch_hetero` (reference scripts/train.py:655-681) and its autograd backward:
 *   lv = max(logvar, floor);  loss = mean_b mean_t [w_b] 0.5 (lv + (mean - target)^2 / exp(lv)) + l2 * mean_{b,t} (0.5 lv)^2
 * mean / logvar / target : f32 [n_graphs, n_targets] (target already z-scored, train.py:650).  mask : optional f32
 * [n_graphs], 1 = real graph (means run over the real graphs; shape-bucket padding).  weight : optional f32 [n_graphs]
 * per-sample weights (train.py:661-675).  Outputs: loss f32[1]; dmean / dlogvar (optional) = gradient of loss. */
int alignn_gaussian_nll(const float *mean, const float *logvar, const float *target, const float *mask,
                        const float *weight, int64_t n_graphs, int n_targets, float min_logvar_floor,
                        float log_sigma_l2, float *loss, float *dmean, float *dlogvar, void *stream);

/* ---- weight + bias gradient of a node projection in one pass (tcgen05 / TMEM) ---------------------------------------
 * Replaces autograd's AddmmBackward of PyG TransformerConv's lin_query / lin_key / lin_value / lin_skip (reference
 * scripts/train.py:691): C[M, N] = A^T B and colsum[M] = sum_k A[k, :], for A [K, M] (row stride lda) = gradient of the
 * projection output and B [K, N] (row stride ldb) = the block input, both bf16, N = 256 (alignn_wgrad_supported).  One
 * streaming pass over A and B (csrc/wgrad_tc.cu: split over K and over 256-channel groups of M, UMMA 128x256x16 with
 * both operands MN-major, accumulators in TMEM), per-CTA partials reduced in a fixed order.
 * partials : f32 [alignn_wgrad_partial_floats(K, M)] scratch.  colsum may be NULL. */
int alignn_wgrad_supported(int n, int dtype);
int64_t alignn_wgrad_partial_floats(int64_t K, int M);
int alignn_wgrad(const void *a, int64_t lda, const void *b, int64_t ldb, int64_t K, int M, int N, int dtype,
                 float *partials, float *c, float *colsum, void *stream);

/* ---- per-node projection GEMM (tcgen05 / TMEM, TMA-fed) -------------------------------------------------------------
 * Replaces the `nn.Linear`s PyG TransformerConv applies to the node state (reference scripts/train.py:308, 326:
 * lin_query / lin_key / lin_value / lin_skip, called from TransformerConv.forward at train.py:315, 334), stacked and
 * folded as gnn_elasticity_predictor_b200/fused.py describes:
 *     C[M, N] = A[M, K] . W[N, K]^T + bias[N]      A, W, bias, C bf16 (row strides lda / ldw / ldc elements), fp32 accumulate
 * K = 256 and N a multiple of 256 (alignn_proj_tc_supported).  csrc/proj_tc.cu: persistent CTAs, the [256, 256] weight block
 * resident in shared memory, [128 x 64] boxes of A through a four-stage TMA ring (128-byte swizzle), 16 tcgen05.mma
 * 128x256x16 per tile, two TMEM accumulators so the epilogue (bias, bf16, swizzled staging, TMA store) overlaps the next
 * tile's MMAs.
 * bias may be NULL.  Never allocates, never synchronises. */
int alignn_proj_tc_supported(int K, int N, int dtype);
int alignn_proj_tc(const void *a, int64_t lda, const void *w, int64_t ldw, const void *bias, void *c, int64_t ldc,
                   int64_t M, int N, int K, int dtype, void *stream);
/* The same pass over A for TWO column groups of one stacked weight W [w_rows, K]: C0[M, n_full] from W rows
 * [w_row_full, +n_full) over all M rows, C1[M_pre, n_pre] from W rows [w_row_pre, +n_pre) over the first M_pre rows; bias is
 * indexed like the rows of W.  What one conv block needs: x_r for every row + q | k | v | qt_0..3 for the rows that have
 * line-graph neighbours (trunk.py). */
int alignn_proj_tc2(const void *a, int64_t lda, const void *w, int64_t ldw, int64_t w_rows, const void *bias,
                    void *c0, int64_t ldc0, int n_full, int w_row_full, void *c1, int64_t ldc1, int n_pre, int w_row_pre,
                    int64_t M, int64_t M_pre, int K, int dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ALIGNN_B200_H */
