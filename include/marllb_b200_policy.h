/*
 * marllb_b200_policy.h -- C ABI of the batched policy kernels (QMIX / SAC-GRU rows of
 * a13-a19).  All pointers are DEVICE pointers to float32 unless noted; every
 * function is asynchronous on `stream` (cudaStream_t as void*) and returns 0 or a negative
 * MLB_E* code (include/marllb_b200.h).
 *
 * Reference modules these kernels stand in for (paths relative to simulation-mode/):
 *   problem-05-qmix/src/agent_network.py:63-87     AgentQNetwork.forward  (GRU cell + 3 linears)
 *   problem-05-qmix/src/mixing_network.py:78-117   QMixingNetwork.forward (hypernets, |.|, bmm, ELU)
 *   problem-05-qmix/src/qmix_agent.py:126-170,192-307  select_actions / update
 *   problem-04-sac-gru/src/networks.py:82-147,209-237  PolicyNetwork.forward/sample, QNetwork.forward
 *   problem-04-sac-gru/src/sac_agent.py:124-255    select_action / update_parameters
 * Parameter tensors keep torch's layouts (nn.Linear weight [out][in]; nn.GRU weight_ih [3H][in],
 * gate order r,z,n) so reference state_dicts load unchanged.
 */
#ifndef MARLLB_B200_POLICY_H
#define MARLLB_B200_POLICY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { MLB_ACT_NONE = 0, MLB_ACT_RELU = 1, MLB_ACT_ABS = 2 };

/*
 * Batched strided GEMM with fused bias / activation epilogue (fp32, FFMA):
 *   C[b][m][n] = act( beta*C[b][m][n] + sum_k A[b](m,k) * B[b](k,n) + bias[b][n] )
 *   A(m,k) = A[b*a_bs + m*a_rs + k*a_cs],  B(k,n) = B[b*b_bs + k*b_rs + n*b_cs],
 *   C row-major with leading dimension ldc and batch stride c_bs; bias nullable, stride bias_bs.
 * Covers nn.Linear forward (B = W^T: b_rs=1, b_cs=K), input gradients (B = W) and weight
 * gradients (A = dY^T) with accumulation (beta = 1).
 */
int mlb_gemm(const float *A, int64_t a_bs, int64_t a_rs, int64_t a_cs,
             const float *B, int64_t b_bs, int64_t b_rs, int64_t b_cs,
             float *C, int64_t c_bs, int64_t ldc, const float *bias, int64_t bias_bs,
             int32_t M, int32_t N, int32_t K, int32_t batch, float beta, int32_t act, void *stream);

/*
 * nn.Linear forward on the tensor cores (tcgen05, 3xTF32 with fp32 accumulation in tensor memory):
 *   C[m][n] = act( sum_k X[m*lda + k] * W[n*ldw + k] + bias[n] ),  X [M][K], W [N][K] (nn.Linear layout).
 * lda and ldw must be multiples of 4 and X, W 16-byte aligned (TMA); any M, N, K >= 1.
 * Same results as mlb_gemm within fp32 rounding (error ~1e-6 relative); used for the large-M
 * (rollout batch) calls of AgentQNetwork / PolicyNetwork / QNetwork forward.
 */
int mlb_linear_tc_supported(int32_t M, int32_t N, int32_t K, int64_t lda, int64_t ldw, int64_t ldc);
int mlb_linear_tc(const float *X, int64_t lda, const float *W, int64_t ldw, const float *bias,
                  float *C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t act, void *stream);

/*
 * The general product of the UPDATE path on the tensor cores (tcgen05, 3xTF32; csrc/mlb_gemm_tc.cu):
 *   C[m*ldc + n] = act( beta * C + sum_k A[m*a_rs + k*a_cs] * B[k*b_rs + n*b_cs] + bias[n] )
 * with each operand either contraction-contiguous (a_cs == 1 / b_rs == 1) or contiguous along its other index
 * (a_rs == 1 / b_cs == 1): nn.Linear forward, its input gradient dx = dy W and its weight gradient dW = dy^T x
 * (autograd of sac_agent.py:197-231, qmix_agent.py:275-285).  Few-tile / long-K shapes are split over K (partial tiles
 * to a workspace, summed in split order: deterministic).  The non-unit strides and ldc must be multiples of 4, N a
 * multiple of 4 and >= 32, K >= 32, pointers 16-byte aligned: mlb_gemm_tc_supported tells; mlb_gemm routes supported
 * single-matrix calls above a size threshold here by itself.  Returns MLB_ESTATE / MLB_ENOMEM (nothing launched) when
 * a tensor map or the workspace cannot be had (e.g. workspace growth during stream capture).
 */
int mlb_gemm_tc_supported(const float *A, int64_t a_rs, int64_t a_cs, const float *B, int64_t b_rs, int64_t b_cs,
                          const float *C, int64_t ldc, int32_t M, int32_t N, int32_t K);
int mlb_gemm_tc(const float *A, int64_t a_rs, int64_t a_cs, const float *B, int64_t b_rs, int64_t b_cs, float *C,
                int64_t ldc, const float *bias, int32_t M, int32_t N, int32_t K, float beta, int32_t act,
                void *stream);

/* Split-K products (mlb_gemm, mlb_gemm_tc) keep their partial tiles in a per-device workspace.  Work issued on
 * CONCURRENT streams must not share it: the calling host thread selects one of MLB_WS_SLOTS workspaces (default 0)
 * for the products it issues from then on -- SAC_GRU_Agent.update_parameters runs the twin critics Q1 / Q2
 * (sac_agent.py:193-207, 213-215: independent until the loss) on two streams with slots 0 / 1. */
#define MLB_WS_SLOTS 8
int mlb_set_workspace_slot(int32_t slot);
int mlb_get_workspace_slot(void);
/* Give slots 1 .. n_slots-1 the capacity slot 0 has reached on the current device (nothing can be allocated inside a
 * stream capture): call after an eager warm-up of the same code on one stream. */
int mlb_reserve_workspace_slots(int32_t n_slots);

/* nn.GRU single step, gate part (torch gate order r,z,n):
 *   gi = x W_ih^T + b_ih, gh = h W_hh^T + b_hh (computed by mlb_gemm), both [M][3H];
 *   r = sigmoid(gi_r+gh_r), z = sigmoid(gi_z+gh_z), n = tanh(gi_n + r*gh_n), h' = (1-z)*n + z*h.
 * gates (nullable) receives [M][3H] = (r, z, n) for the backward pass. */
int mlb_gru_gates_forward(const float *gi, const float *gh, const float *h, float *h_new,
                          float *gates, int32_t M, int32_t H, void *stream);
/* backward of the above: given dh_new, saved gates, h and gh_n (= gh[:, 2H:3H]) produce
 * dgi [M][3H], dgh [M][3H] and dh_direct [M][H] (= dh_new * z; caller adds dgh W_hh). */
int mlb_gru_gates_backward(const float *dh_new, const float *gates, const float *h, const float *gh,
                           float *dgi, float *dgh, float *dh_direct, int32_t M, int32_t H, void *stream);

/* elementwise / reductions */
int mlb_relu_backward(const float *y, const float *dy, float *dx, int64_t n, void *stream); /* dx = dy*(y>0) */
int mlb_colsum(const float *dy, float *db, int32_t M, int32_t N, int64_t ld, float beta, void *stream); /* db = beta*db + sum_m dy */
int mlb_axpby(float a, const float *x, float b, float *y, int64_t n, void *stream);       /* y = a*x + b*y (soft update: networks.py:248-260) */
int mlb_sumsq(const float *x, int64_t n, double *out_accum, void *stream);                 /* *out += sum x^2 (clip_grad_norm_) */
int mlb_scale(float *x, int64_t n, const double *norm_sq, float max_norm, void *stream);   /* x *= min(1, max_norm/(sqrt(norm_sq)+1e-6)) */
/* torch.optim.Adam step (no weight decay, no amsgrad): p, g, m, v of n elements; step = t >= 1 */
int mlb_adam(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1,
             float beta2, float eps, int32_t step, void *stream);

/* The same with the step counter in device memory (step_dev[0] = steps taken so far; advanced by
 * the call) and a 2-float scratch coef_dev: no host value changes from step to step, so an
 * optimiser step can be captured in a CUDA graph. */
int mlb_adam_dev(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1,
                 float beta2, float eps, int32_t *step_dev, float *coef_dev, void *stream);

/* QMIX epsilon-greedy selection (qmix_agent.py:159-164): q [M][K]; u [M] uniform draws,
 * rnd [M] int32 pre-drawn random actions; action = u < epsilon ? rnd : argmax (first max). */
int mlb_egreedy_select(const float *q, const float *u, const int32_t *rnd, float epsilon,
                       int32_t *action, float *q_sel, int32_t M, int32_t K, void *stream);
/* QMIX actions -> per-server env action (rl_controller.py:314-321 puts the weight on the chosen
 * server of each agent): out[m][j] = (action[m*stride] == j) ? hot : cold, m < M agents, j < Sa. */
int mlb_onehot_action(const int32_t *action, int32_t M, int32_t Sa, int32_t stride, uint8_t hot,
                      uint8_t cold, uint8_t *out, void *stream);
/* row max (target Q: qmix_agent.py:253) and gather q[m][idx[m]] */
int mlb_row_max(const float *q, float *out, int32_t *argmax, int32_t M, int32_t K, void *stream);

/* QMIX mixer core (mixing_network.py:104-116) for M samples, A agents, E embed:
 *   hidden = elu(q[M][A] . w1[M][A][E] + b1[M][E]);  q_tot = hidden . w2[M][E] + b2[M]
 * (w1, w2 already |.|'d by the hypernet epilogue).  hidden_out nullable [M][E]. */
int mlb_mixer_forward(const float *q, const float *w1, const float *b1, const float *w2,
                      const float *b2, float *q_tot, float *hidden_out, int32_t M, int32_t A,
                      int32_t E, void *stream);
/* backward: dq_tot [M] -> dq [M][A], dw1 [M][A][E], db1 [M][E], dw2 [M][E], db2 [M]
 * (gradients w.r.t. the |.|'d tensors; the abs backward is sign(pre) applied by the caller op). */
int mlb_mixer_backward(const float *dq_tot, const float *q, const float *w1, const float *w2,
                       const float *hidden, float *dq, float *dw1, float *db1, float *dw2,
                       float *db2, int32_t M, int32_t A, int32_t E, void *stream);
int mlb_abs_backward(const float *pre, const float *dy, float *dx, int64_t n, void *stream); /* dx = dy*sign(pre) */

/* SAC tanh-Gaussian head (networks.py:112-147) on [M][A] tensors:
 *   log_std clamped to [lo,hi]; x = mean + exp(log_std)*eps; y = tanh(x); action = y*scale+bias;
 *   logp[m] = sum_a( -0.5*eps^2 - log_std - 0.5*log(2pi) - log(scale*(1-y^2)+1e-6) );
 *   mean_action = tanh(mean)*scale + bias. */
int mlb_tanh_gaussian_forward(const float *mean, const float *log_std_raw, const float *eps,
                              float lo, float hi, float scale, float bias, float *action,
                              float *logp, float *mean_action, int32_t M, int32_t A, void *stream);
/* backward: d_action [M][A] and d_logp [M] -> d_mean, d_log_std_raw (zero where clamped) */
int mlb_tanh_gaussian_backward(const float *mean, const float *log_std_raw, const float *eps,
                               float lo, float hi, float scale, const float *d_action,
                               const float *d_logp, float *d_mean, float *d_log_std_raw,
                               int32_t M, int32_t A, void *stream);

int mlb_abs_forward(const float *x, float *y, int64_t n, void *stream);                     /* y = |x| (mixing_network.py:92,99) */

/* QMIX TD loss (qmix_agent.py:263-278) over [B][T] row-major tensors:
 *   targets = reward_sum + gamma*(1-done)*shift(target_q_tot)   (shift: t <- t+1, last column 0)
 *   mask[b][t] = t < seq_len[b];  loss = sum((q_tot-targets)^2*mask)/sum(mask)
 * writes targets, dq_tot = 2*(q_tot-targets)*mask/sum(mask) and stats[0..2] (double) =
 * {loss, mean(q_tot), mean(targets)} (means over ALL B*T entries like the reference's .mean()). */
int mlb_qmix_td_loss(const float *q_tot, const float *target_q_tot, const float *reward_sum,
                     const float *done, const int32_t *seq_len, float gamma, float *targets,
                     float *dq_tot, double *stats, int32_t B, int32_t T, void *stream);

/* SAC critic target (sac_agent.py:180-190): y = r + (1-d)*gamma*(min(q1n,q2n) - alpha*logp_next) */
int mlb_sac_q_target(const float *reward, const float *done, const float *q1n, const float *q2n,
                     const float *logp_next, const float *alpha, float gamma, float *y, int32_t M,
                     void *stream);
/* F.mse_loss(q, y) (mean) and its gradient dq = 2*(q-y)/M; loss (double) accumulated into *loss */
int mlb_mse_loss(const float *q, const float *y, float *dq, double *loss, int32_t M, void *stream);
/* SAC actor loss (sac_agent.py:210-216): L = mean(alpha*logp - min(q1,q2));
 * d_logp = alpha/M, dq1/dq2 = -1/M routed to the smaller critic (ties: q1, like torch.min). */
int mlb_sac_policy_loss(const float *logp, const float *q1, const float *q2, const float *alpha,
                        float *d_logp, float *dq1, float *dq2, double *loss, int32_t M, void *stream);
/* temperature loss (sac_agent.py:223-231): L = -mean(log_alpha*(logp+target_entropy));
 * writes d_log_alpha[0] = -mean(logp+target_entropy) and *loss. */
int mlb_sac_alpha_loss(const float *logp, const float *log_alpha, float target_entropy,
                       float *d_log_alpha, double *loss, int32_t M, void *stream);
int mlb_exp_scalar(const float *x, float *y, void *stream);                                  /* y[0] = exp(x[0]) */

/* ---- original-paper agents (src/lb/sac_qmix.py:195-460, src/lb/sac_gru_discrete.py:128-359):
 * multi-head categorical outputs, rows of n <= 64 classes ---- */
/* F.softmax(x, dim=-1) over rows [rows][n] (sac_qmix.py:250, sac_gru_discrete.py:189) and its backward
 * dx = y * (dy - sum_k dy_k y_k) */
int mlb_softmax_forward(const float *x, float *y, int64_t rows, int32_t n, void *stream);
int mlb_softmax_backward(const float *y, const float *dy, float *dx, int64_t rows, int32_t n, void *stream);
/* torch.cat([state, one_hot(last_action)], -1) (sac_qmix.py:231-236): out [rows][F + heads*n] */
int mlb_concat_onehot(const float *x, const int32_t *action, float *out, int64_t rows, int32_t F,
                      int32_t heads, int32_t n, void *stream);
/* Categorical(p) over rows: class = given[r] if given, else inverse-CDF of u[r] if u, else argmax (first
 * maximum, np.argmax).  Nullable outputs: action [rows], logp [rows] = log p[class], psel [rows] = p[class]
 * (sac_qmix.py:268-274, 431-432; sac_gru_discrete.py:201-209). */
int mlb_categorical(const float *p, const float *u, const int32_t *given, int32_t *action, float *logp,
                    float *psel, int64_t rows, int32_t n, void *stream);
/* d/dlogits of sum_r g[r/group] * log p[r][action[r]]: dlogits[r][k] = g * ((k==a_r) - p[r][k]) */
int mlb_logprob_backward(const float *p, const int32_t *action, const float *g, float *dlogits,
                         int64_t rows, int32_t n, int32_t group, void *stream);
/* backward of the gather of the chosen class: d[r][k] = (k == action[r]) ? g[r] : 0 */
int mlb_scatter_class(const float *g, const int32_t *action, float *d, int64_t rows, int32_t n, void *stream);
/* WeightedQMixingNetwork.forward (problem-05-qmix/src/mixing_network.py:231-246): out[m] = sum_a q[m][a]*w[m][a],
 * and its backward dq = g*w, dw = g*q (g [M]) */
int mlb_weighted_sum_forward(const float *q, const float *w, float *out, int32_t M, int32_t A, void *stream);
int mlb_weighted_sum_backward(const float *g, const float *q, const float *w, float *dq, float *dw,
                              int32_t M, int32_t A, void *stream);
/* QMix_Trainer._build_td_lambda_targets (sac_qmix.py:449-460) over [B][T]:
 * ret[:,T-1] = tq[:,T-1]; ret[:,t] = lambda*gamma*ret[:,t+1] + (r[:,t] + (1-lambda)*gamma*tq[:,t+1]) */
int mlb_td_lambda_targets(const float *reward, const float *target_q, float *ret, float gamma,
                          float td_lambda, int32_t B, int32_t T, void *stream);
/* reward normalisation of SAC_Trainer.update (sac_gru_discrete.py:299-300) over [B][T]:
 * out = scale * (r - mean_b r) / (std_b r + 1e-6), unbiased std over the batch dimension (B >= 2) */
int mlb_reward_normalize(const float *reward, float *out, float scale, int32_t B, int32_t T, void *stream);
/* discrete SAC critic target (sac_gru_discrete.py:316-319): y = r + gamma*(min(q1n,q2n) - alpha*logp_next) */
int mlb_dsac_q_target(const float *reward, const float *q1n, const float *q2n, const float *logp_next,
                      const float *alpha, float gamma, float *y, int32_t M, void *stream);

/* nn.GRU over a whole sequence in one launch (the unrolled time loops of QMIXAgent.update, qmix_agent.py:217-224,
 * 246-253): gi_all [T][B][3H] = x_t W_ih^T + b_ih for every step (one GEMM), h0 [B][H] -> hs [T][B][H]; hprev, ghs,
 * gates (nullable) receive what the backward pass needs.  mlb_gru_seq_backward: dhs [T][B][H] gradient w.r.t. every
 * output -> dgi_all, dgh_all [T][B][3H] (the weight gradients are then two GEMMs over T*B rows) and dh0 (nullable).
 * 3H <= 1024 threads and W_hh resident in shared memory: H <= 128.  Same arithmetic as the per-step kernels. */
int mlb_gru_seq_forward(const float *gi_all, const float *W_hh, const float *b_hh, const float *h0, float *hs,
                        float *hprev, float *ghs, float *gates, int32_t T, int32_t B, int32_t H, void *stream);
int mlb_gru_seq_backward(const float *dhs, const float *gates, const float *hprev, const float *ghs,
                         const float *W_hh, float *dgi_all, float *dgh_all, float *dh0, int32_t T, int32_t B,
                         int32_t H, void *stream);

/* Device-resident replay ring for the batched rollout (ReplayBuffer.push / sample, problem-04-sac-gru/src/
 * replay_buffer.py:35-94, keeps one transition per Python call in a host deque).  Ring tensors: r_state /
 * r_next_state [capacity][state_dim], r_action [capacity][action_dim], r_hidden [capacity][hidden_dim], r_reward /
 * r_done [capacity] float.  mlb_replay_push writes the n transitions of one env step to rows (pos + e) % capacity and
 * advances the device-held write position (so a CUDA graph replay lands in the right place); mlb_replay_gather copies
 * the rows idx[0..batch) into contiguous batch tensors (uniform indices drawn by the caller). */
int mlb_replay_push(const float *state, const float *action, const double *reward, const float *next_state,
                    const uint8_t *done, const float *hidden, float *r_state, float *r_action, float *r_reward,
                    float *r_next_state, float *r_done, float *r_hidden, int64_t *pos_dev, int32_t n,
                    int32_t capacity, int32_t state_dim, int32_t action_dim, int32_t hidden_dim, void *stream);
int mlb_replay_gather(const float *r_state, const float *r_action, const float *r_reward, const float *r_next_state,
                      const float *r_done, const float *r_hidden, const int64_t *idx, float *state, float *action,
                      float *reward, float *next_state, float *done, float *hidden, int32_t batch,
                      int32_t state_dim, int32_t action_dim, int32_t hidden_dim, void *stream);

#ifdef __cplusplus
}
#endif
#endif
