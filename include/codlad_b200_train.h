/* codlad_b200 -- C ABI of the train_latent step (BASELINE.json configs[4]).
 *
 * The reference trains the denoiser with torch autograd (train_latent.py:184-261: diffusion.training_losses ->
 * accelerator.backward -> clip_grad_norm_ -> AdamW.step -> update_ema).  This header is the operator set the B200 training step is
 * composed of: device pointers + sizes + a cudaStream_t (as void*), fp32 row-major tensors, hidden width 128, `int` return codes
 * (0 = ok; cb2_last_error() of codlad_b200.h holds the message).  Every reduction is deterministic (fixed order, no atomics).
 * The host side that strings them into the forward / backward of ProteinMPNN_diffusion_new is codlad_b200/train.py; the gradient
 * all-reduce of the data-parallel step is torch.distributed (NCCL) on the flat gradient buffer.
 */
#ifndef CODLAD_B200_TRAIN_H
#define CODLAD_B200_TRAIN_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef CB2_API
#define CB2_API __attribute__((visibility("default")))
#endif

/* Replaces: nn.Linear forward / its two autograd gradients (every W*, dense.W_in / W_out, adaLN_modulation.1, x_in, W_e, edge_embedding,
 * W_out.linear of models/latent_model.py:78-165 and models/protein_mpnn_utils.py:208-330).  C[M,N] (+)= op(A)[M,K] op(B)[K,N]:
 * a_kc != 0: A stored [M][lda] (k contiguous), else [K][lda]; b_kc != 0: B stored [N][ldb] (k contiguous), else [K][ldb].
 *   forward Y = X W^T: (X, W, a_kc=1, b_kc=1);  dgrad dX = dY W: (dY, W, 1, 0);  wgrad dW = dY^T X: (dY, X, 0, 0). */
CB2_API int cb2t_gemm(const float* A, const float* B, float* C, int M, int N, int K, long long lda, long long ldb, long long ldc,
                      int a_kc, int b_kc, int accumulate, void* stream);

/* Arithmetic of cb2t_gemm: 0 = fp32 SIMT everywhere (the gradient-parity mode), 1 = TF32 tensor cores (tcgen05.mma kind::tf32, fp32
 * accumulation) wherever the shape allows -- what the reference itself trains with (train_latent.py:24-25, allow_tf32 = True): forward
 * and data gradients when both operands are K-contiguous (pass W^T for dgrad), weight gradients when both are MN-contiguous and the
 * output is made of full 128 x 128 blocks; everything else stays on the SIMT path. */
CB2_API int cb2t_set_gemm_mode(int mode);

/* Replaces: `+ bias` and torch.nn.GELU() (exact erf form; protein_mpnn_utils.py:227,243,325).  Z [rows, cols] <- Z + bias (kept as the
 * pre-activation), Y <- GELU(Z) when Y != NULL.  bias may be NULL. */
CB2_API int cb2t_bias_gelu_fwd(float* Z, const float* bias, long long rows, int cols, float* Y, void* stream);
/* Replaces: torch.nn.Linear followed by torch.nn.GELU() in one call (the W2 / W3 / W12 / W13 / dense layers, protein_mpnn_utils.py:240-247,
 * 261-270, 300-307, 319-330).  Z [M, N] = X [M, K] W [N, K]^T + bias (kept: the backward needs the pre-activation), Y = GELU(Z) when Y != NULL
 * (same pitch ldz as Z).  In TF32 mode the bias and the GELU run in the GEMM's epilogue; otherwise it is cb2t_gemm + cb2t_bias_gelu_fwd. */
CB2_API int cb2t_linear_bias_gelu_fwd(const float* X, const float* W, const float* bias, float* Z, float* Y, int M, int N, int K, long long ldx,
                              long long ldw, long long ldz, void* stream);
/* dpre = dY * GELU'(pre) (autograd of the above; in place allowed). */
CB2_API int cb2t_gelu_bwd(const float* pre, const float* dY, long long n, float* dpre, void* stream);
/* The same with the bias gradient of the layer that produced `pre` in the same pass: colsum_out [cols] (+)= column sums of dpre
 * (autograd's `.sum(0)` for a Linear bias; two passes, fixed order).  pre, dY, dpre [rows, cols] contiguous, cols % 4 == 0. */
CB2_API int cb2t_gelu_bwd_colsum(const float* pre, const float* dY, long long rows, int cols, float* dpre, float* colsum_out, int accumulate,
                         void* stream);

/* mode 0: out = SiLU(a) (adaLN_modulation.0, t_embedder.mlp.1); 1: out = b * SiLU'(a); 2: out = a + b; 3: out = a * scale; 4: out = a * b (dropout masks). */
CB2_API int cb2t_elementwise(int mode, const float* a, const float* b, float scale, long long n, float* out, void* stream);

/* Replaces: cat_neighbors_nodes + cat([h_V_i, .]) + W1 / W11 (protein_mpnn_utils.py:240-243, 261-264; decoder :300-303 with
 * latent_model.py:258-262) with W1 factored into its blocks: Z [E,128] holds the h_E block's product on entry and
 * z = Z + Pa[e / K] + Pc[nbr_node[e]] (+ bias) on exit; Y = GELU(z).  nbr_node [E] = batch-global node index of every edge's neighbour. */
CB2_API int cb2t_edge_combine_gelu_fwd(float* Z, const float* Pa, const float* Pc, const float* bias, const int* nbr_node, int K, long long E,
                                       float* Y, void* stream);
/* Its backward onto the per-node products: dPa[i] = sum_k dZ[i,k];  dPc[j] = sum over the edges e with nbr_node[e] = j, listed by the
 * reverse CSR (rev_ptr [N+1], rev_edge [E]: edge ids sorted by neighbour, ascending within a neighbour). */
CB2_API int cb2t_edge_gather_bwd(const float* dZ, int K, int N, const int* rev_ptr, const int* rev_edge, float* dPa, float* dPc, void* stream);

/* Replaces: mask_attend * h_message, torch.sum(h_message, -2) / scale (protein_mpnn_utils.py:244-246, 304-306) and its autograd.
 * mask_e [E] may be NULL (the decoder passes no mask). */
CB2_API int cb2t_masked_sum_fwd(const float* M, const float* mask_e, int K, int N, float scale, float* S, void* stream);
CB2_API int cb2t_masked_sum_bwd(const float* dS, const float* mask_e, int K, long long E, float scale, float* dM, void* stream);

/* Replaces: norm(h + dropout(dh)) -> gate * modulate(., shift, scale) [* mask_V] (protein_mpnn_utils.py:247-259, 266-270, 307-316;
 * FinalLayer, latent_model.py:31-34 with gate = NULL; the affine LayerNorm of the featuriser, :521, with rows_per_member = rows,
 * scale = weight - 1, shift = bias, gate = NULL).  x = A + drop * B (B, drop may be NULL; drop holds mask / (1 - p));
 * y = row_mask * gate[m] * ((x - mean) rstd (1 + scale[m]) + shift[m]) with m = row / rows_per_member and the modulation vectors at
 * shift + m * mod_stride etc.  X (needed when B != NULL) receives x, stats [rows, 2] = (mean, rstd). */
CB2_API int cb2t_ln_mod_fwd(const float* A, const float* B, const float* drop, long long rows, long long rows_per_member, const float* shift,
                            const float* scale, const float* gate, long long mod_stride, const float* row_mask, float eps, float* X, float* stats,
                            float* Y, void* stream);
/* Its autograd: dX [rows,128] and the per-member gradients of the modulation vectors, d_shift / d_scale / d_gate at the same stride
 * (written, or added to when accumulate != 0; d_gate ignored when gate == NULL). */
CB2_API int cb2t_ln_mod_bwd(const float* dY, const float* X, const float* stats, long long rows, long long rows_per_member, const float* shift,
                            const float* scale, const float* gate, long long mod_stride, const float* row_mask, float* dX, float* d_shift,
                            float* d_scale, float* d_gate, int accumulate, void* stream);

/* Replaces: the input side of CA_ProteinFeatures.forward (protein_mpnn_utils.py:478-516): for every (i, k) neighbour pair the 144 RBF
 * values of the nine C-alpha distances and the 7 orientation features, raw [F, L, K, 152] (column 151 = 0).  X [F,L,3], idx / D [F,L,K]
 * from cb2_knn_topk.  The trainable projections (positional embedding, edge_embedding, norm_edges, W_e) follow as cb2t_gemm / cb2t_ln_mod. */
CB2_API int cb2t_edge_raw_features(const float* X, const int* idx, const float* D, int F, int L, int K, float* raw, void* stream);
/* Z[r, :] += T[idx[r], :] (128 columns): the positional embedding folded through edge_embedding, looked up per edge. */
CB2_API int cb2t_row_gather_add(float* Z, const float* T, const int* idx, long long rows, void* stream);

/* Replaces: the autograd of an embedding lookup (W_s, latent_model.py:225; the one-hot of PositionalEncodings, protein_mpnn_utils.py:340-343):
 * out[c, :] (+)= sum of the rows of X [n, cols] with idx == c.  cols <= 128, classes <= 72. */
CB2_API int cb2t_index_sum(const float* X, const int* idx, long long n, int cols, int classes, float* out, int accumulate, void* stream);
/* Bias gradients: out[c] (+)= sum_r X[r * ld + c]. */
CB2_API int cb2t_colsum(const float* X, long long rows, int cols, long long ld, float* out, int accumulate, void* stream);
/* out[0] = sum x^2 (the global gradient norm of clip_grad_norm_, train_latent.py:252). */
CB2_API int cb2t_sumsq(const float* x, long long n, float* out, void* stream);

/* Replaces: clip_grad_norm_ (scaling) + torch.optim.AdamW.step + update_ema (train_latent.py:252-261, utils/train_module.py:101-111) on
 * flat buffers of n parameters.  grad_sumsq (device, may be NULL) with max_norm > 0 applies coef = min(1, max_norm / (sqrt(sumsq) + 1e-6))
 * to the gradients without a host round trip.  step = 1, 2, ... (bias correction); ema may be NULL. */
CB2_API int cb2t_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, long long n, float lr, float beta1, float beta2, float eps,
                           float weight_decay, int step, float ema_decay, const float* grad_sumsq, float max_norm, void* stream);

#ifdef __cplusplus
}
#endif
#endif
