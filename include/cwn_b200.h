/* cwn_b200 — C ABI of the B200-native cochain message-passing hot path.
 *
 * The reference (twitter-research/cwn) is pure Python: its hot path bottoms out in three third-party calls,
 *   Tensor.index_select(dim, idx)                         mp/cell_mp.py:195-198, data/complex.py:579-580,587-588
 *   torch_scatter.scatter(src, index, dim=-2, dim_size=N, reduce)   mp/cell_mp.py:437-440,456-459,476-479; mp/layers.py:487
 *   torch_geometric global_add_pool / global_mean_pool     mp/nn.py:50-60 (= scatter over the `batch` vector)
 * plus the per-message Linear(2F->F)+activation of SparseCINConv(use_coboundaries=True), mp/layers.py:210-211,290-293.
 * This library is what replaces that boundary. There is no native ABI in the reference to copy; the entry points
 * below are what a binding for this path binds (see INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch caching allocator); the library never
 *     allocates, frees or keeps pointers, and is re-entrant / thread-safe (autograd calls it from worker threads);
 *   - matrices are row-major fp32 with an explicit leading dimension `ld` (in elements);
 *   - adjacency columns arrive as int64 (the reference API: `assert index.dtype == torch.long`,
 *     mp/cell_mp.py:158) and are narrowed ONCE per batch to int32 CSR plans that every layer, forward and
 *     backward, reuses;
 *   - all work is stream-ordered on `stream` (a cudaStream_t); no call synchronises the device;
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = argument error (CWN_E_*); cwn_last_error_string() gives the
 *     calling thread's last message. No exception crosses the boundary.
 */
#ifndef CWN_B200_H
#define CWN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cwn_stream_t; /* cudaStream_t */

enum {
  CWN_OK = 0,
  CWN_E_NULL = -1,      /* required pointer is NULL */
  CWN_E_SHAPE = -2,     /* negative size, ld < F, F <= 0, sizes beyond int32 range */
  CWN_E_ENUM = -3,      /* unknown reduce / activation code */
  CWN_E_WORKSPACE = -4, /* workspace too small */
  CWN_E_ALIGN = -5      /* pointer not 4-byte aligned */
};

/* aggregation of the messages that reach a destination (reference: aggr_up/aggr_down/aggr_boundary,
 * mp/cell_mp.py:84-86,104-105). Rows without messages are 0 for every mode (mp/test_cell_mp.py:114-134). */
enum { CWN_REDUCE_ADD = 0, CWN_REDUCE_MEAN = 1, CWN_REDUCE_MAX = 2 };

/* activation of the coboundary message MLP (reference mp/nn.py:7-27) */
enum { CWN_ACT_ID = 0, CWN_ACT_RELU = 1, CWN_ACT_ELU = 2, CWN_ACT_SIGMOID = 3, CWN_ACT_TANH = 4 };

const char* cwn_version(void);
const char* cwn_last_error_string(void);
/* number of kernels this library has launched in this process (for bench.py's gpu_launches) */
unsigned long long cwn_launch_count(void);
/* test hook: route the grouped dense entry points through their generic (any shape / alignment) kernels instead of
 * the fast path for 16-byte aligned operands, so the two can be compared on the same inputs. Process-wide.
 * mask bit 0: cwn_linear_fwd_grouped, bit 1: cwn_unit_bwd_grouped; 0 restores the default. */
int cwn_debug_force_generic_dense(int32_t mask);

/* ---------------------------------------------------------------------------------------------------------
 * CSR plan: group the E messages of one adjacency by one of its columns.
 *   key[e] in [0, n_rows)  : the grouping column (destination = index[1] for the forward pass; source = index[0]
 *                            for the gradient w.r.t. the gathered operand; shared_coboundaries for the gradient
 *                            w.r.t. the coboundary operand)
 *   pay0, pay1 (nullable)  : other per-message columns to carry along (e.g. source and coboundary ids)
 * Outputs (all int32, caller-allocated):
 *   rowptr[n_rows+1]       : messages of row r are positions [rowptr[r], rowptr[r+1])
 *   perm[E]                : original message id at each position; STABLE (ascending e inside a row), so a
 *                            sequential in-row accumulation reproduces the order of CPU scatter_add_ exactly
 *   pay0_sorted[E], pay1_sorted[E] : pay*[perm[i]] narrowed to int32 (only written if the input is non-NULL)
 *   flags[1] (nullable)    : bit 0 set if some key was outside [0, n_rows) (such messages are dropped)
 * Replaces: the unsorted atomics scatter of torch_scatter (reference mp/cell_mp.py:439-440).
 */
size_t cwn_csr_plan_workspace_bytes(int64_t E, int64_t n_rows);
int cwn_csr_plan_build(const int64_t* key, const int64_t* pay0, const int64_t* pay1, int64_t E, int64_t n_rows,
                       int32_t* rowptr, int32_t* perm, int32_t* pay0_sorted, int32_t* pay1_sorted,
                       int32_t* flags, void* workspace, size_t workspace_bytes, cwn_stream_t stream);

/* All the plans of one batch in ONE kernel launch (one CTA per plan, block-wide stable radix sort in shared memory).
 * Every plan must have E <= cwn_csr_plan_small_capacity() messages; `descs` is a HOST array read during the call
 * (its device pointers follow the conventions of cwn_csr_plan_build). No workspace. */
typedef struct {
  const int64_t* key;
  const int64_t* pay0; /* nullable */
  const int64_t* pay1; /* nullable */
  int64_t E;
  int64_t n_rows;
  int32_t* rowptr;
  int32_t* perm;
  int32_t* pay0_sorted;
  int32_t* pay1_sorted;
} cwn_plan_desc;
int64_t cwn_csr_plan_small_capacity(void);
int cwn_csr_plan_build_small(const cwn_plan_desc* descs, int32_t n_plans, int32_t* flags, cwn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused gather -> (identity message) -> reduce, one destination row per thread group, no atomics:
 *   out[r,:] = (x_res ? (1 + *eps) * x_res[r,:] : 0) + REDUCE_{i in row r} x_src[ idx ? idx[i] : i , :]
 * `idx` = the plan's source column (fused identity pass: boundary messages, upper messages without coboundaries,
 * InitReduceConv, and — with the transposed plan — their gradients), or the plan's `perm` (aggregation of
 * messages a user hook materialised), or NULL (sorted segments: the per-complex readout, rowptr = ptr).
 * `eps` (device scalar, nullable => 0) is the GIN epsilon of mp/layers.py:191-192; x_res requires CWN_REDUCE_ADD.
 * Replaces: index_select + scatter (mp/cell_mp.py:198 + :439-440 / :478-479), mp/layers.py:485-487, mp/nn.py:59.
 */
int cwn_csr_gather_reduce_f32(const float* x_src, int64_t ld_src, const int32_t* rowptr, const int32_t* idx,
                              int64_t n_rows, int32_t F, const float* x_res, int64_t ld_res, const float* eps,
                              float* out, int64_t ld_out, int32_t reduce, cwn_stream_t stream);

/* Max aggregation with argument tracking, for the backward pass of `aggr='max'` (reference mp/cell_mp.py:437-440 ->
 * torch_scatter.scatter(reduce='max'), whose CPU kernel keeps the FIRST maximum in message order):
 *   out[r,f] = max_{i in row r} x_src[idx[i], f]  (0 for rows without messages);  arg[r,f] = perm[i*] (message id) or -1
 *   gX[s,f]  = SUM_{i in row s of the by-source plan} (arg[dst[i], f] == perm[i]) ? G[dst[i], f] : 0
 * perm = the plan's stable permutation (original message id of every plan position); deterministic, no atomics. */
int cwn_csr_gather_max_arg_f32(const float* x_src, int64_t ld_src, const int32_t* rowptr, const int32_t* idx,
                               const int32_t* perm, int64_t n_rows, int32_t F, float* out, int64_t ld_out,
                               int32_t* arg /* [n_rows, F] */, cwn_stream_t stream);
int cwn_csr_max_bwd_f32(const float* G, int64_t ld_g, const int32_t* arg, const int32_t* rowptr, const int32_t* dst,
                        const int32_t* perm, int64_t n_rows, int32_t F, float* gX, int64_t ld_gx, cwn_stream_t stream);

/* The pass above with a SECOND residual operand and an optional plan:
 *   out[r,:] = (1+eps) * x_res[r,:] + (1+eps2) * x_res2[r,:] + SUM_{i in row r} x_src[idx[i],:]
 * rowptr == NULL means "no messages" (out = the two residual terms). This is the gradient fan-in of a cochain's features
 * inside one SparseCINConv layer: (1+eps1) gU_d + (1+eps2) gB_d + the transposed boundary pass of dimension d+1 —
 * one launch instead of a gather, two scaled copies and two additions. x_res / x_res2 nullable. */
int cwn_csr_gather_reduce2_f32(const float* x_src, int64_t ld_src, const int32_t* rowptr, const int32_t* idx,
                               int64_t n_rows, int32_t F, const float* x_res, int64_t ld_res, const float* eps,
                               const float* x_res2, int64_t ld_res2, const float* eps2, float* out, int64_t ld_out,
                               cwn_stream_t stream);

/* Row gather out[e,:] = scale * x[idx[e],:] (int64 idx straight from the API). Used for operands of user-defined
 * message hooks (reference __lift__, mp/cell_mp.py:195-198), lazily requested `up_attr`/`down_attr`
 * (data/complex.py:579-580,587-588) and the gradient of the readout. */
int cwn_gather_rows_f32(const float* x, int64_t ld_x, const int64_t* idx, int64_t E, int32_t F, float scale,
                        float* out, int64_t ld_out, cwn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Upper-adjacency pass with coboundary features (SparseCINConv(use_coboundaries=True), mp/layers.py:210-211,
 * 290-293): message_e = act(W [x[s_e] ; y[c_e]] + b). With W = [W1 | W2] the per-message Linear splits into two
 * per-CELL products P = x W1^T, Q = y W2^T + b (dense, done by the caller), and the pass becomes memory-bound:
 *   out[r,:] = (x_res ? (1 + *eps) * x_res[r,:] : 0) + SUM_{i in row r} act( P[src[i],:] + Q[cob[i],:] )
 * (plan grouped by destination; src/cob = the plan's two payload columns).
 */
int cwn_csr_cob_fwd_f32(const float* P, int64_t ld_p, const float* Q, int64_t ld_q, const int32_t* rowptr,
                        const int32_t* src, const int32_t* cob, int64_t n_rows, int32_t F, int32_t act,
                        const float* x_res, int64_t ld_res, const float* eps, float* out, int64_t ld_out,
                        cwn_stream_t stream);

/* Gradient of the pass above w.r.t. ONE of its two gathered operands A (the other is B), on the plan grouped by
 * A's index column, so A[r,:] is read once per row and nothing is atomically accumulated:
 *   gA[r,:] = SUM_{i in row r} G[dst[i],:] * act'( A[r,:] + B[oth[i],:] )
 * Call once with (A,B) = (P,Q) on the by-source plan and once with (A,B) = (Q,P) on the by-coboundary plan. */
int cwn_csr_cob_bwd_f32(const float* G, int64_t ld_g, const float* A, int64_t ld_a, const float* B, int64_t ld_b,
                        const int32_t* rowptr, const int32_t* dst, const int32_t* oth, int64_t n_rows, int32_t F,
                        int32_t act, float* gA, int64_t ld_ga, cwn_stream_t stream);

/* fp64 instantiations of cwn_csr_gather_reduce_f32, cwn_gather_rows_f32, cwn_csr_cob_fwd_f32 and cwn_csr_cob_bwd_f32
 * (same contracts; `eps` and `scale` are double): the reference runs its strongly-regular-graph isomorphism experiments
 * in float64 (exp/run_exp.py:41-43). Same in-row order => bit-identical to a sequential CPU scatter_add_ in double.
 * Plain kernels (those datasets are small); matrices need 8-byte alignment only. */
int cwn_csr_gather_reduce_f64(const double* x_src, int64_t ld_src, const int32_t* rowptr, const int32_t* idx,
                              int64_t n_rows, int32_t F, const double* x_res, int64_t ld_res, const double* eps,
                              double* out, int64_t ld_out, int32_t reduce, cwn_stream_t stream);
int cwn_gather_rows_f64(const double* x, int64_t ld_x, const int64_t* idx, int64_t E, int32_t F, double scale,
                        double* out, int64_t ld_out, cwn_stream_t stream);
int cwn_csr_cob_fwd_f64(const double* P, int64_t ld_p, const double* Q, int64_t ld_q, const int32_t* rowptr,
                        const int32_t* src, const int32_t* cob, int64_t n_rows, int32_t F, int32_t act,
                        const double* x_res, int64_t ld_res, const double* eps, double* out, int64_t ld_out,
                        cwn_stream_t stream);
int cwn_csr_cob_bwd_f64(const double* G, int64_t ld_g, const double* A, int64_t ld_a, const double* B, int64_t ld_b,
                        const int32_t* rowptr, const int32_t* dst, const int32_t* oth, int64_t n_rows, int32_t F,
                        int32_t act, double* gA, int64_t ld_ga, cwn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * The three passes above for the HBM-bound regime (tens of thousands of rows and more per launch), warp-specialised:
 * one producer warp per persistent CTA streams, several tiles ahead, the plan slices, the operand-row WINDOWS and the
 * tile's own residual rows into shared memory with TMA bulk copies; eight consumer warps run the rows out of shared
 * memory (csrc/gsa_ws.cu). Same results, bit for bit, as cwn_csr_gather_reduce_f32 / cwn_csr_cob_fwd_f32 /
 * cwn_csr_cob_bwd_f32 (same in-row order, same roundings). Requirements: F % 4 == 0, F <= 128, every matrix, rowptr,
 * payload column and `windows` 16-byte aligned, every ld % 4 == 0.
 *
 * cwn_csr_tile_windows — once per plan and tile height: for tile t (rows [t*tile_rows, (t+1)*tile_rows))
 *   windows[8t..8t+7] = { m0, m1 (message range), lo0, cnt0 (rows lo0 .. lo0+cnt0-1 of the operand behind pay0 are the
 *   only ones its messages read), lo1, cnt1 (same for pay1; 0 if pay1 == NULL), 0, 0 }
 * followed by 4 statistics words the CALLER ZEROES before the call: { max cnt0, max cnt1, max 4-aligned message hull, 0 }.
 * `windows` therefore holds 8 * ceil(n_rows / tile_rows) + 4 int32. The caller sizes the kernels' buffers from the
 * statistics (cap_rows*, cap_msgs); a tile that exceeds them is still computed correctly, from global memory.
 * cwn_csr_ws_stages — pipeline depth the configuration gets (0 or 1: do not use these entry points for it).
 * E = number of messages of the plan (length of the payload columns). */
int cwn_csr_tile_windows(const int32_t* rowptr, const int32_t* pay0, const int32_t* pay1 /* nullable */, int64_t n_rows,
                         int32_t tile_rows, int32_t* windows, cwn_stream_t stream);
int cwn_csr_ws_consumer_threads(void); /* rows of a tile are shared by consumer_threads / lanes_per_row(F) groups */
int cwn_csr_ws_lanes_per_row(int32_t F);
int cwn_csr_ws_stages(int32_t F, int32_t tile_rows, int32_t cap_rows0, int32_t cap_rows1, int32_t cap_msgs,
                      int32_t n_arrays, int32_t has_row_operand);
int cwn_csr_gather_reduce_ws_f32(const float* x_src, int64_t ld_src, const int32_t* rowptr, const int32_t* idx,
                                 int64_t E, const int32_t* windows, int32_t tile_rows, int32_t cap_rows,
                                 int32_t cap_msgs, int64_t n_rows, int32_t F, const float* x_res, int64_t ld_res,
                                 const float* eps, float* out, int64_t ld_out, int32_t reduce /* add | mean */,
                                 cwn_stream_t stream);
int cwn_csr_cob_fwd_ws_f32(const float* P, int64_t ld_p, const float* Q, int64_t ld_q, const int32_t* rowptr,
                           const int32_t* src, const int32_t* cob, int64_t E, const int32_t* windows,
                           int32_t tile_rows, int32_t cap_rows0, int32_t cap_rows1, int32_t cap_msgs, int64_t n_rows,
                           int32_t F, int32_t act, const float* x_res, int64_t ld_res, const float* eps, float* out,
                           int64_t ld_out, cwn_stream_t stream);
int cwn_csr_cob_bwd_ws_f32(const float* G, int64_t ld_g, const float* A, int64_t ld_a, const float* B, int64_t ld_b,
                           const int32_t* rowptr, const int32_t* dst, const int32_t* oth, int64_t E,
                           const int32_t* windows, int32_t tile_rows, int32_t cap_rows0, int32_t cap_rows1,
                           int32_t cap_msgs, int64_t n_rows, int32_t F, int32_t act, float* gA, int64_t ld_ga,
                           cwn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * K5 — message MLP of the dense CIN layers with BatchNorm over the MESSAGE population (reference mp/layers.py:94-103,
 * nets mp/models.py:40-47: Linear(2F -> F), act, BatchNorm1d). With the Linear in split-weight form (P, Q as in
 * cwn_csr_cob_fwd_f32) a message is a_e = act(P[src_e] + Q[att_e]) and its BatchNorm is affine, so
 *   out[t] = scale * S[t] + deg(t) * (beta - scale * mean),  S = cwn_csr_cob_fwd_f32 without residual,
 * and only the statistics and the backward need message-level passes of their own:
 *   cwn_cin_msg_sq_f32 : out[t,:] = SUM_{i in row t} (act(P[src[i]] + Q[att[i]]) - mu)^2      (var = colsum / E)
 *   cwn_cin_msg_bwd_f32: gA[r,:]  = SUM_{i in row r} scale (G[dst[i]] - c1 - ahat_i c2) act'(A[r] + B[oth[i]]),
 *                        ahat_i = (act(A[r] + B[oth[i]]) - mu) rstd — the BatchNorm backward over the messages folded
 *                        into the gradient of one gathered operand (by-source plan for P, by-attribute plan for Q).
 * mu, scale (= gamma * rstd), rstd, c1, c2: DEVICE vectors [F]. No [E, F] tensor is ever materialised. */
int cwn_cin_msg_sq_f32(const float* P, int64_t ld_p, const float* Q, int64_t ld_q, const int32_t* rowptr,
                       const int32_t* src, const int32_t* att, int64_t n_rows, int32_t F, int32_t act, const float* mu,
                       float* out, int64_t ld_out, cwn_stream_t stream);
int cwn_cin_msg_bwd_f32(const float* G, int64_t ld_g, const float* A, int64_t ld_a, const float* B, int64_t ld_b,
                        const int32_t* rowptr, const int32_t* dst, const int32_t* oth, int64_t n_rows, int32_t F,
                        int32_t act, const float* scale, const float* mu, const float* rstd, const float* c1,
                        const float* c2, float* gA, int64_t ld_ga, cwn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Dense update / combine nets of SparseCINConv (reference mp/layers.py:191-199, 303-325):
 *   Linear -> BatchNorm -> act -> Linear -> BatchNorm -> act  (x2 branches),  Linear(2H->H) -> BatchNorm -> act.
 * One "unit" is  z = f_in(X) W^T + b  followed by BatchNorm statistics over the rows, where the input transform
 * f_in(x)[k] = act_in((x[k] - in_mean[k]) * in_scale[k] + in_beta[k]) is the PREVIOUS unit's BatchNorm + activation
 * applied on the fly while the tile is loaded (so normalised activations are never written to HBM), and X may be the
 * virtual concatenation [X0 | X1] (combine_nn's torch.cat). All entry points are GROUPED: `descs` is a host array of
 * up to CWN_MAX_GROUP problems (the two branches of the three cochain dimensions) served by ONE launch.
 * Products: tcgen05 tensor cores (kind::tf32 on hi/lo split operands, accumulation spread over several TMEM
 * accumulators — as accurate as an fp32 FMA chain, see csrc/dense_tc5.cuh) when h and K are powers of two <= 128 and
 * the operands are 16-byte aligned; fp32 FFMA kernels otherwise (plain TF32 cannot meet the 1e-5 rtol parity gate).
 * CWN_B200_DENSE_TC5=0 forces the FFMA kernels.
 */
#define CWN_MAX_GROUP 8

typedef struct {
  const float* x0; int64_t ld_x0; int32_t k0;              /* first input block  [n_rows, k0] */
  const float* x1; int64_t ld_x1; int32_t k1;              /* optional second block (k1 = 0: absent) */
  const float* in_mean0; const float* in_scale0; const float* in_beta0; /* per-column input transform, NULL = none */
  const float* in_mean1; const float* in_scale1; const float* in_beta1;
  int32_t in_act;                                          /* CWN_ACT_* applied after the affine */
  const float* w; int64_t ld_w;                            /* [h, k0 + k1] row-major (torch Linear.weight) */
  const float* bias;                                       /* [h], nullable */
  float* z; int64_t ld_z;                                  /* output [n_rows, h] */
  float* stats;                                            /* nullable: per row tile (64 rows) [n_tiles, 2, h] = (mean, M2) */
  int64_t n_rows; int32_t h;
  /* optional fused BatchNorm finalisation (same outputs as cwn_bn_finalize_grouped): the LAST CTA of the problem to
   * finish merges the per-tile partials, so no second launch is needed. `counter` must point to a zero int32; it is
   * reset to zero before the kernel ends. bn_mean == NULL disables it. */
  const float* bn_gamma; float bn_eps; float bn_momentum; int32_t bn_training;
  float* bn_running_mean; float* bn_running_var; int64_t* bn_num_batches_tracked;
  float* bn_mean; float* bn_scale; float* bn_rstd;
  int32_t* counter;
  int32_t tile_rows;   /* 64 or 32 (0 = 64): rows per CTA tile = granularity of `stats`; the same for a whole group.
                        * 32 doubles the CTA count of small problems (latency-bound launches) */
  const int32_t* n_rows_live; /* nullable DEVICE scalar: only the first *n_rows_live (<= n_rows) rows are real, the rest is
                        * padding of a fixed-capacity batch (ragged batches replayed through one CUDA graph,
                        * cwn_b200/bucketed.py). All n_rows rows of z are still written, but the BatchNorm statistics
                        * (and running statistics) see the live rows only. Tensor-core path only (CWN_E_SHAPE otherwise). */
} cwn_linear_desc;
int cwn_linear_fwd_grouped(const cwn_linear_desc* descs, int32_t n, cwn_stream_t stream);

/* BatchNorm statistics from the per-tile partials of cwn_linear_fwd_grouped (training), or from the running
 * statistics (training = 0). Writes mean[h], scale[h] = gamma * rstd, rstd[h]; in training mode also updates
 * running_mean / running_var (unbiased) with `momentum` and increments num_batches_tracked (all nullable). */
typedef struct {
  const float* stats; int32_t n_tiles; int64_t n_rows; int32_t h;
  const float* gamma; const float* beta; float eps; float momentum; int32_t training;
  float* running_mean; float* running_var; int64_t* num_batches_tracked;
  float* mean; float* scale; float* rstd;
  int32_t tile_rows;   /* rows per tile of `stats` (0 = 64) */
} cwn_bn_desc;
int cwn_bn_finalize_grouped(const cwn_bn_desc* descs, int32_t n, cwn_stream_t stream);

/* out = act((z - mean) * scale + beta): the layer output that neighbouring cells gather from. */
typedef struct {
  const float* z; int64_t ld_z; const float* mean; const float* scale; const float* beta; int32_t act;
  float* out; int64_t ld_out; int64_t n_rows; int32_t h;
} cwn_bn_act_desc;
int cwn_bn_act_grouped(const cwn_bn_act_desc* descs, int32_t n, cwn_stream_t stream);

/* Backward of one unit. With y = (z - mean) * scale + beta, out = act(y), zhat = (z - mean) * rstd:
 *   step 1 (cwn_unit_bwd_reduce_grouped): per-tile partials of s1 = SUM_rows g_out * act'(y), s2 = SUM_rows g_out * act'(y) * zhat
 *   step 2 (cwn_unit_bwd_finalize_grouped): c1 = s1/N, c2 = s2/N, g_gamma = s2, g_beta = s1
 *   step 3 (cwn_unit_bwd_grouped): g_z = scale * (g_out*act'(y) - c1 - zhat*c2)   [has_bn = 0: g_z = g_out * act'(z)]
 *            g_in = g_z W (gradient w.r.t. f_in(X), split into g_in0 | g_in1), per-CTA partials of
 *            g_W = g_z^T f_in(X) and g_b = SUM_rows g_z
 *   step 4 (cwn_wgrad_finalize_grouped): ordered sum of the partials -> g_W, g_b (accumulated INTO the outputs if accumulate != 0)
 * Everything is deterministic (no atomics). */
typedef struct {
  /* forward operands (saved) */
  const float* x0; int64_t ld_x0; int32_t k0; const float* x1; int64_t ld_x1; int32_t k1;
  const float* in_mean0; const float* in_scale0; const float* in_beta0;
  const float* in_mean1; const float* in_scale1; const float* in_beta1; int32_t in_act;
  const float* w; int64_t ld_w;
  const float* z; int64_t ld_z; int32_t has_bn; int32_t act;   /* act of THIS unit's output (CWN_ACT_ID if none) */
  const float* mean; const float* scale; const float* rstd; const float* beta;
  /* incoming gradient w.r.t. the unit's output */
  const float* g_out; int64_t ld_g;
  /* BN-backward reductions */
  float* red_partials;   /* [n_tiles, 2, h] */
  float* c1; float* c2;  /* [h] each */
  float* g_gamma; float* g_beta; int32_t accumulate_affine;
  /* outputs */
  float* g_in0; int64_t ld_gi0; float* g_in1; int64_t ld_gi1;   /* nullable: gradient not needed */
  float* w_partials;     /* [n_ctas, h, k0+k1] */
  float* b_partials;     /* [n_ctas, h] */
  int32_t n_ctas;        /* CTAs assigned to this problem in step 3 (each strides over the row tiles) */
  float* g_w; int64_t ld_gw; float* g_b; int32_t accumulate_w;
  int64_t n_rows; int32_t h;
  int32_t* counter;      /* nullable: zero int32; if set, the last CTA of cwn_unit_bwd_reduce_grouped performs step 2 */
  int32_t tile_rows;     /* 64 or 32 (0 = 64), same for a whole group; sizes red_partials and bounds n_ctas */
  int32_t accumulate_in; /* step 3 adds the input gradient INTO g_in0 / g_in1 instead of overwriting them (every element is
                          * owned by one thread: deterministic). Lets a caller sum the gradient contributions of several
                          * consumers of one tensor without extra elementwise launches; problems of ONE launch must not
                          * share an output buffer */
  const int32_t* n_rows_live; /* nullable DEVICE scalar, as in cwn_linear_desc: rows >= *n_rows_live are padding. Their g_z
                          * is zero (so they add nothing to g_W / g_b and their g_in rows are written as zeros) and the
                          * BatchNorm-backward means c1, c2 divide by the live count. cwn_unit_bwd_grouped: tensor-core
                          * path only (CWN_E_SHAPE otherwise) */
  /* FUSED REDUCTION OF THE UPSTREAM UNITS (optional; tensor-core path only, CWN_E_SHAPE otherwise). Input block i of
   * this unit is the output z of an upstream unit U_i whose BatchNorm + activation are this unit's in_mean_i / in_scale_i /
   * in_beta_i / in_act; g_in_i is therefore U_i's g_out, and steps 1-2 of U_i (its BatchNorm-backward column sums and
   * their finalisation) can be taken from the g_in tile while it is still in shared memory instead of by a launch of
   * their own that re-reads g_in and z from memory. Set next_red_i to U_i's red_partials ([n_tiles, 2, k_i], tiles of 64
   * rows) to enable it for block i; then next_rstd_i, next_c1_i, next_c2_i are required, next_g_gamma_i / next_g_beta_i /
   * next_accumulate_affine_i have the meaning of g_gamma / g_beta / accumulate_affine of U_i, and next_counter (zero
   * int32, reset by the kernel) elects the CTA that finalises. Requires g_in_i != NULL and accumulate_in == 0 (g_in_i must
   * be U_i's complete output gradient). The caller then skips cwn_unit_bwd_reduce_grouped for U_i.
   * cwn_unit_bwd_fuses_reduce(descs, n) tells whether a group will take the tensor-core path (1) or not (0). */
  const float* next_rstd0; float* next_red0; float* next_c1_0; float* next_c2_0; float* next_g_gamma0; float* next_g_beta0;
  const float* next_rstd1; float* next_red1; float* next_c1_1; float* next_c2_1; float* next_g_gamma1; float* next_g_beta1;
  int32_t next_accumulate_affine0; int32_t next_accumulate_affine1;
  int32_t* next_counter;
} cwn_unit_bwd_desc;
int cwn_unit_bwd_reduce_grouped(const cwn_unit_bwd_desc* descs, int32_t n, cwn_stream_t stream);
int cwn_unit_bwd_finalize_grouped(const cwn_unit_bwd_desc* descs, int32_t n, cwn_stream_t stream);
int cwn_unit_bwd_grouped(const cwn_unit_bwd_desc* descs, int32_t n, cwn_stream_t stream);
int cwn_unit_bwd_fuses_reduce(const cwn_unit_bwd_desc* descs, int32_t n);
int cwn_wgrad_finalize_grouped(const cwn_unit_bwd_desc* descs, int32_t n, cwn_stream_t stream);

/* Adam (torch.optim.Adam semantics: L2 weight decay, no amsgrad; reference exp/run_exp.py:343) over flat parameter /
 * gradient / moment buffers in ONE launch. `step` (device int32, number of updates so far) is advanced by the kernel;
 * `counter` must point to a zero int32 and is reset by the kernel. zero_grad != 0 clears `grad` after use. */
int cwn_adam_step_f32(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int32_t* step, int32_t* counter, int32_t zero_grad,
                      cwn_stream_t stream);

/* Data parallel: the gradient all-reduce (average over `world` ranks of one node) fused with the Adam step above, in ONE
 * launch over NVLink peer memory — no NCCL call on the step. `grad_ptrs` / `signal_pads` are DEVICE arrays [world] of
 * every rank's flat gradient bucket / signal pad as mapped into THIS process (symmetric memory: cudaIpc / fabric
 * handles; torch.distributed._symmetric_memory's buffer_ptrs_dev). Pads: n_ctas * world zero-initialised uint32 per
 * rank, private to this entry point. n % 4 == 0, buffers 16-byte aligned, n_ctas <= 148 and identical on every rank.
 * Two-shot: rank r averages slice r of all buckets in rank order (bit-identical results on every rank) and publishes it
 * to every bucket; then each rank runs Adam locally. `*error` (device, zeroed by the caller) becomes non-zero if a
 * peer did not show up within ~2 s (the kernel gives up rather than hang). Replaces: nothing in the reference (single
 * device, exp/run_exp.py:22-23); torch DDP's NCCL all-reduce + torch.optim.Adam in a multi-GPU port of it. */
int cwn_allreduce_adam_step_f32(float* param, float* const* grad_ptrs, uint32_t* const* signal_pads, int32_t rank,
                                int32_t world, int32_t n_ctas, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int32_t* step, int32_t* counter,
                                int32_t zero_grad, int32_t* error, cwn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * GPU-side collation (replaces the CPU loops of ComplexBatch.from_complex_list / CochainBatch.from_cochain_list,
 * data/complex.py:323-458, 690-728). Every tensor of a batch is a concatenation of per-complex segments of a
 * device-resident dataset, each shifted by a per-segment constant (the running cell-count offsets of `__inc__`,
 * data/complex.py:148-169). One job = one output tensor (or one row of a [2, E] index):
 *   dst[dst_start[j] + t] = src[src_start[j] + t] (+ add[j])      for t in [0, dst_start[j+1] - dst_start[j])
 * CWN_COLLATE_FILL writes the segment number j instead (the `batch` vector). All jobs of a batch go in ONE launch.
 * src_start[n_segments], dst_start[n_segments + 1], add[n_segments] (nullable) are DEVICE int64 arrays. */
enum { CWN_COLLATE_I64 = 0, CWN_COLLATE_F32_ROWS = 1, CWN_COLLATE_FILL = 2 };
typedef struct {
  const void* src; void* dst;
  const int64_t* src_start; const int64_t* dst_start; const int64_t* add;
  int32_t n_segments; int32_t kind; int32_t row_elems; /* floats per row for CWN_COLLATE_F32_ROWS */
  int64_t n_out;                                        /* = dst_start[n_segments], known on the host */
} cwn_collate_job;
int cwn_collate(const cwn_collate_job* jobs, int32_t n_jobs, cwn_stream_t stream);

/* Debug aid: sets bit 1 of flags[0] if any idx[e] is outside [0, n). (The reference relies on torch's device
 * assert for out-of-range indices.) */
int cwn_check_index_range(const int64_t* idx, int64_t E, int64_t n, int32_t* flags, cwn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Readout head of the SparseCIN-family models (reference mp/nn.py:50-60 pool_complex; mp/models.py:230-254 and
 * mp/molec_models.py:137-161: per-dimension lin1 + act, sum/mean over dimensions, lin2), one launch forward and
 * two backward instead of ~45 library launches:
 *   pooled_d[b] = SUM|MEAN_{i in complex b} x_d[i] ;  z_d[b] = pooled_d[b] W1_d^T + b1_d ;
 *   h[b] = SUM|MEAN_d act(z_d[b]) ;  out[b] = h[b] W2^T + b2
 * The cells of complex b in dimension d are rows perm[rowptr[b] .. rowptr[b+1]) of x_d (row plan of `batch_d`;
 * perm NULL = identity; rowptr NULL = the dimension is absent from the batch: pooled_d = 0, as pool_complex does).
 * Dropout must be inactive (the Python wrapper falls back to torch otherwise).
 */
#define CWN_MAX_HEAD_DIMS 4
typedef struct {
  const float* x; int64_t ld_x;              /* [n_d, K] */
  const int32_t* rowptr; const int32_t* perm;
  const float* w1; const float* b1;          /* [H2, K] row-major, [H2] (nullable) */
  float* pooled;                             /* [B, K]  written by fwd, read by bwd */
  float* z;                                  /* [B, H2] pre-activations, written by fwd, read by bwd */
  /* backward only */
  float* g_z;                                /* [B, H2] scratch */
  float* g_x; int64_t ld_gx;                 /* [n_d, K], nullable */
  float* g_w1; float* g_b1;                  /* nullable */
  int32_t accumulate;                        /* add into g_w1 / g_b1 instead of overwriting */
} cwn_head_dim;
int cwn_readout_head_fwd(const cwn_head_dim* dims, int32_t n_dims, int64_t B, int32_t K, int32_t H2, int32_t out_size,
                         int32_t act, int32_t pool_mean, int32_t final_mean, const float* w2, const float* b2,
                         float* h /* [B, H2] */, float* out /* [B, out_size] */, cwn_stream_t stream);
int cwn_readout_head_bwd(const cwn_head_dim* dims, int32_t n_dims, int64_t B, int32_t K, int32_t H2, int32_t out_size,
                         int32_t act, int32_t pool_mean, int32_t final_mean, const float* w2, const float* h,
                         const float* g_out, float* g_w2, float* g_b2, int32_t accumulate_out, cwn_stream_t stream);
/* The same in two parts (`parts`: bit 0 = input gradients g_z / g_x, bit 1 = parameter gradients). Nothing in the rest
 * of a backward pass reads the parameter gradients, so a caller can launch part 2 on another stream after part 1 and
 * keep its ~40 us of ordered sums off the critical path. cwn_readout_head_bwd == parts 3. */
int cwn_readout_head_bwd_parts(const cwn_head_dim* dims, int32_t n_dims, int64_t B, int32_t K, int32_t H2,
                               int32_t out_size, int32_t act, int32_t pool_mean, int32_t final_mean, const float* w2,
                               const float* h, const float* g_out, float* g_w2, float* g_b2, int32_t accumulate_out,
                               int32_t parts, cwn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CWN_B200_H */
