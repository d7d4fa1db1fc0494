/* uavnet -- C-ABI of the kernels under the A3C actor-critic MLP (the caller of the env step path, SURVEY.md 8(f1/f2)).
 *
 * The reference builds two 3-layer MLPs on the flattened (nBS+1)*G*G observation in TensorFlow 1.x
 * (main.py:143-156: actor 50000->200->200->625 softmax, critic 50000->200->200->1, relu6, N(0,0.1) kernels) and
 * trains them with two RMSProp optimisers (main.py:300-301) from 10-step rollouts (main.py:212-238).  The observation
 * has ~44 non-zero cells of 50 000, so the first layer is a gather-sum of weight rows: the env kernel emits the
 * non-zero cells as flat indices (uavenv_out.obs_idx) and these entry points consume them.  The small dense layers
 * (200x200, 200x625, 200x1) and their gradients run on the tcgen05 tensor cores through uavnet_gemm.
 *
 * Plain C types, raw device pointers, the stream is a cudaStream_t passed as void*.  Every entry point returns 0 or a
 * negative UAVNET_E* code and only enqueues work on the stream.  One process per GPU (the launch model of this
 * library), but a process may drive several: every entry point makes the device that owns its first pointer argument
 * current, and the host-side state (SM count, function attributes, the gemm error word) is kept per device ordinal.  Not
 * thread-safe.
 */
#ifndef UAVNET_H
#define UAVNET_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { UAVNET_OK = 0, UAVNET_EINVAL = -1, UAVNET_ECUDA = -2 };
#define UAVNET_MAX_PEERS 8

/* First dense layer on a sparse count vector (tf.layers.dense(self.s, 200, relu6), main.py:147,151):
 *   out[m, :] = act( b + sum_k W[idx[m,k], :] ),  act = relu6 (relu6 != 0) or identity.
 * idx int32 [M,K] (duplicates count twice), W float32 [n_rows, H] row-major, H a multiple of 4 (actor and critic first
 * layers are stored side by side: H = 400), out float32 [M,H]. */
int uavnet_sparse_fwd(const int32_t *idx, int64_t M, int32_t K, int64_t n_rows, const float *W, const float *b,
                      int32_t H, float *out, int32_t relu6, void *stream);

/* Its weight gradient: dW[idx[m,k], :] += dpre[m, :] (float atomics; dW float32 [n_rows,H], zeroed by the caller or by
 * uavnet_rmsprop), and db[:] += sum_m dpre[m, :] is left to the caller (a column sum). */
int uavnet_sparse_bwd(const int32_t *idx, int64_t M, int32_t K, int64_t n_rows, const float *dpre, int32_t H, float *dW,
                      void *stream);

/* The same weight gradient, gather side: the (sample, slot) pairs are bucketed by row with a counting sort (histogram,
 * exclusive scan, fill of (sample, row) into `workspace`), then threads walk fixed-size slices of the sorted list, add up the
 * gradient vectors with plain loads and flush one 16-byte RED per row change -- ~60x fewer float atomics than the scatter
 * version, and balanced whatever the row histogram looks like.  col_passes: the sum runs separately over H / col_passes
 * columns at a time so that the gathered operand stays L2-resident (2 for the rollout batch: 65 MB per half).  workspace:
 * device memory of uavnet_sparse_bwd_gather_workspace(M, K, n_rows) bytes, 16-byte aligned.  Indices outside [0, n_rows)
 * are ignored.  Results are deterministic up to fp32 summation order (like the scatter version's). */
int64_t uavnet_sparse_bwd_gather_workspace(int64_t M, int32_t K, int64_t n_rows);
int uavnet_sparse_bwd_gather(const int32_t *idx, int64_t M, int32_t K, int64_t n_rows, const float *dpre, int32_t H, float *dW,
                             void *workspace, int32_t col_passes, void *stream);
/* The two halves of it: _prepare buckets the pairs (needs the indices only -- a caller can run it on a side stream as soon
 * as the rollout's indices exist, under the dense layers' backward products), _apply does the sums. */
int uavnet_sparse_bwd_gather_prepare(const int32_t *idx, int64_t M, int32_t K, int64_t n_rows, void *workspace, void *stream);
int uavnet_sparse_bwd_gather_apply(int64_t M, int32_t K, int64_t n_rows, const float *dpre, int32_t H, float *dW, void *workspace,
                                   int32_t col_passes, void *stream);
int uavnet_sparse_bwd_gather_apply_cols(int64_t M, int32_t K, int64_t n_rows, const float *dpre, int32_t H, int32_t col0, int32_t n_cols,
                                        float *dW, void *workspace, void *stream);

/* Actor head of the rollout (main.py:149,165-169): prob = softmax(logits) and action ~ np.random.choice(A, p=prob) by
 * inverse CDF with one Philox4x32-10 uniform per sample, keyed by (seed, row_offset + row, counter) -- the first action
 * whose cumulative probability exceeds u.  logits float32 [M,A]; prob float32 [M,A] out (may be NULL); action int64 [M]
 * out (may be NULL).  counter = *counter_dev (a uint32 in device memory, NULL = 0) + counter_add: keeping the running
 * count on the device lets a CUDA graph that contains the call draw fresh numbers on every replay. */
int uavnet_softmax_sample(const float *logits, int64_t M, int32_t A, uint64_t seed, uint32_t row_offset,
                          const uint32_t *counter_dev, uint32_t counter_add, float *prob, int64_t *action, void *stream);

/* d(a_loss)/d(logits) of the actor loss (main.py:68-76), fused over the softmax output:
 *   a_loss = mean_i -( log(prob[i,a_i] + 1e-5) * td_i + beta * H_i ),  H_i = -sum_j prob_ij log(prob_ij + 1e-5)
 * prob float32 [M,A] (softmax output), a_his int64 [M], td float32 [M] (v_target - v, treated as constant),
 * dz float32 [M,A] out with leading dimension ldz >= A (rows padded to 16 bytes feed uavnet_gemm's vector path),
 * loss_row float32 [M] out (the per-sample loss term; its mean is a_loss; may be NULL). */
int uavnet_actor_head_bwd(const float *prob, const int64_t *a_his, const float *td, int64_t M, int32_t A, float beta,
                          float *dz, int64_t ldz, float *loss_row, void *stream);

/* out[m,j] = dv[m] * w[j] * (0 < h[m,j] < 6): the data gradient of the critic's value head (main.py:153, 200 -> 1)
 * through the relu6 of its second layer -- a rank-1 product, memory-bound.  dv float32 [M], w float32 [H], h and out
 * float32 [M,H] contiguous, H a multiple of 4. */
int uavnet_rank1_mask(const float *dv, const float *w, const float *h, int64_t M, int32_t H, float *out, void *stream);

/* The critic's side of the losses in one launch (main.py:64-66): td[i] = v_target[i] - v[i] (the TD error both losses use),
 * dv[i] = d(mean td^2)/dv[i] = -2 td[i] / M, loss2[1] = mean(td^2) = c_loss; loss2[0] is zeroed for uavnet_mean_rows.  loss2: 2 floats. */
int uavnet_critic_td(const float *v_target, const float *v, int64_t M, float *td, float *dv, float *loss2, void *stream);
/* *out_accum += mean(rows[0..M)): a_loss from the per-sample terms uavnet_actor_head_bwd writes (main.py:76) */
int uavnet_mean_rows(const float *rows, int64_t M, float *out_accum, void *stream);

/* Bookkeeping of one rollout step (main.py:199-211: ep_r += r, buffer_r.append(r)) for E envs in one launch:
 * reward_out[e] = (float) reward[e] (the env kernel's float64 reward), done_out[e] = done[e], ep_return[e] += reward[e];
 * where done[e], the finished episode's return goes to ep_finished[e] and ep_return[e] restarts at 0 (ep_r of
 * main.py:188,246-252).  ep_return / ep_finished may be NULL. */
int uavnet_rollout_record(const double *reward, const uint8_t *done, int64_t E, float *reward_out, uint8_t *done_out,
                          double *ep_return, double *ep_finished, void *stream);

/* Discounted n-step value targets of the worker loop (main.py:217-227), batched over envs: walking the rollout
 * backwards, v = r[t] + gamma * (done[t] ? 0 : v), starting from the bootstrap value v_boot of the state after the last
 * step.  rewards float32 [T,E], dones uint8 [T,E], v_boot float32 [E], out float32 [T,E]. */
int uavnet_nstep_targets(const float *rewards, const uint8_t *dones, const float *v_boot, int32_t T, int64_t E, float gamma,
                         float *out, void *stream);

/* TensorFlow-1 RMSPropOptimizer step (main.py:300-301; decay 0.9, momentum 0, epsilon 1e-10, slot `ms` starts at 1):
 *   g = grad * grad_scale;  ms = decay*ms + (1-decay)*g*g;  param -= lr * g / sqrt(ms + eps);  grad = 0 (if zero_grad)
 * over n float32 elements (any n; 16-byte aligned pointers). */
int uavnet_rmsprop(float *param, float *grad, float *ms, int64_t n, float lr, float decay, float eps, float grad_scale,
                   int32_t zero_grad, void *stream);

/* ---- the gradient push fused with the optimiser over NVLink peer memory (one process per GPU) ----
 * uavnet_p2p_alloc: cudaMalloc (zeroed) + its 64-byte CUDA IPC handle, to be exchanged between the ranks' processes;
 * uavnet_p2p_open: map a peer's buffer into this process (peer access over NVLink / NVSwitch is enabled lazily);
 * uavnet_p2p_close / uavnet_p2p_free: unmap a peer's buffer / free an own one. */
int uavnet_p2p_alloc(int64_t bytes, void **dev_ptr, uint8_t *handle64);
int uavnet_p2p_open(const uint8_t *handle64, void **dev_ptr);
int uavnet_p2p_close(void *dev_ptr);
int uavnet_p2p_free(void *dev_ptr);

/* One kernel per rank instead of all-reduce + optimiser (main.py:85-86,159-163): rank `rank` of `world` owns the slice
 * [rank*ceil(n/4/world)*4, ...) of the flat buffers; for each owned element it sums grads[r][i] over all ranks (peer
 * loads, fixed order), scales by 1/world, applies the TF1 RMSProp step with its local slot ms_local[i], and stores the
 * new parameter into params[r][i] of EVERY rank (peer stores).  Only the owned slice of the own gradient buffer is
 * zeroed; the caller clears the rest after the second collective (the peers are still reading it).  grads / params: HOST arrays of
 * `world` device pointers (own buffer at index `rank`, peers' mapped with uavnet_p2p_open); n a multiple of 4.
 * The caller must order the launch after all ranks finished writing their gradients and order the next use of the
 * parameters after all ranks' launches completed (two stream-ordered collectives, e.g. 4-byte all-reduces). */
int uavnet_p2p_rmsprop(float *const *grads, float *const *params, float *ms_local, int64_t n, int32_t rank, int32_t world,
                       float lr, float decay, float eps, void *stream);

/* The same push with the ranks ordered by flag words in peer memory instead of by the caller's collectives: no NCCL
 * call, no host round trip, replayable inside a CUDA graph.  flags: HOST array of `world` device pointers to 256-byte
 * zero-initialised IPC buffers (uavnet_p2p_alloc / uavnet_p2p_open), own buffer at index `rank`.  Two launches: the push
 * kernel announces "my gradients are complete" to every rank, waits for everyone's announcement, does reduce-scatter +
 * RMSProp + all-gather and announces "my slice is everywhere"; the finish kernel waits for everyone's second
 * announcement and zeroes the rest of the own gradient buffer.  After it (stream order) the parameters of every rank
 * are identical and the gradient buffer is zero.  Every rank must call it the same number of times.  Waits are bounded
 * (about 2 s): uavnet_p2p_push_status reports the number of completed pushes and whether a wait ever gave up. */
int uavnet_p2p_push(float *const *grads, float *const *params, uint32_t *const *flags, float *ms_local, int64_t n, int32_t rank,
                    int32_t world, float lr, float decay, float eps, void *stream);
int uavnet_p2p_push_status(const uint32_t *flags_own, uint32_t *epoch_out, uint32_t *timeout_out);
/* The same push for PART of the flat buffers: up to two parts, each a column range of a row-major matrix inside the buffers
 * (offset / row_width / col0 / n_cols in float32 elements, multiples of 4; a plain range is one row).  b may be NULL.  The
 * elements of a, then of b, are numbered consecutively and rank r owns the r-th of `world` equal slices of that numbering.
 * What the learner uses it for: the actor half of the first layer's gradient (columns 0..199 of [50000, 400]) is pushed as
 * soon as it is complete, on a side stream under the critic half's gather pass; the rest follows at the end.  Every element
 * must be covered by exactly one push per update, the same way on every rank. */
typedef struct uavnet_push_part { int64_t offset; int64_t rows; int32_t row_width; int32_t col0; int32_t n_cols; } uavnet_push_part;
int uavnet_p2p_push_part(float *const *grads, float *const *params, uint32_t *const *flags, float *ms_local, const uavnet_push_part *a,
                         const uavnet_push_part *b, int32_t rank, int32_t world, float lr, float decay, float eps, void *stream);

/* ---- the dense layers (main.py:148-149,152-153: 200->200 relu6, 200->625 softmax logits, 200->1) and their gradients
 * on the 5th-generation tensor cores: tcgen05.mma kind::tf32, fp32 accumulation in tensor memory, fused epilogue ----
 *   D[M,N] (+)= op(A)[M,K] . op(B)[K,N]
 *   a_trans = 0: A is [M,K] row-major (leading dimension lda);  1: A is stored [K,M] row-major (weight gradients X^T . dY)
 *   b_trans = 0: B is [K,N] row-major (a weight matrix as tf.layers.dense stores it);  1: B is stored [N,K] row-major
 *                (data gradients dY . W^T read the same weight matrix)
 * epilogue, in this order: + bias[N]; relu6; dot_out[m] = sum_n D[m,n] * dot_w[n] + *dot_b (the critic's value head on top
 * of its second layer; N <= 256); D *= (0 < mask_src[m,n] < 6) (relu6 backward, mask_src = the layer's output); then either
 * a plain store or, with accumulate != 0, float REDs (D += ...) from split_k slices of K (0 = chosen to fill the GPU).
 * colsum (accumulate only): colsum[n] += sum_k op(B)[k,n] -- a row of ones appended to op(A), i.e. the bias gradient of
 * the layer comes out of the same pass; M may be 0 (column sums only).  D may be NULL when only dot_out / colsum is wanted.
 * out_colsum (plain stores only): out_colsum[n] += sum_m D[m,n] of the values stored -- the bias gradient of the layer
 * BELOW a data-gradient product comes out of its epilogue.
 * precision: UAVNET_GEMM_TF32 (operands rounded to 10 mantissa bits) or UAVNET_GEMM_3XTF32 (hi/lo split, three MMAs per
 * k-step: fp32-class accuracy).  Any M, N <= 65536, K, leading dimensions and alignments (16-byte aligned operands with
 * leading dimensions that are multiples of 4 take vector loads).  Only enqueues work; uavnet_gemm_check() synchronises
 * the device and returns non-zero if any launch since the start gave up on a barrier (a bug, never a data condition). */
enum { UAVNET_GEMM_TF32 = 0, UAVNET_GEMM_3XTF32 = 1 };
typedef struct uavnet_gemm_desc {
    const float *A; int64_t lda; int32_t a_trans;
    const float *B; int64_t ldb; int32_t b_trans;
    float *D; int64_t ldd;
    int64_t M; int32_t N; int64_t K;
    const float *bias; int32_t relu6;
    const float *mask_src; int64_t ld_mask;
    int32_t accumulate; int32_t split_k;
    float *colsum;
    float *out_colsum;
    const float *dot_w; const float *dot_b; float *dot_out;
    int32_t precision;
} uavnet_gemm_desc;
int uavnet_gemm(const uavnet_gemm_desc *desc, void *stream);
int uavnet_gemm_check(void);

#ifdef __cplusplus
}
#endif
#endif /* UAVNET_H */
