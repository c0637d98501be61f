/*
 * s2t_b200.h -- C ABI of libs2t_b200.so: the B200 (sm_100a) implementation of
 * speech2text's transducer-loss hot path.
 *
 * Every entry point replaces one call the reference makes into its un-vendored
 * native dependencies (k2 / torchaudio); the reference-side binding is the
 * ctypes stub in speech2text_b200/_lib.py (shown in INTEGRATION.md).
 *
 * Conventions
 *   - all pointers are DEVICE pointers into caller-owned buffers (the Python
 *     host allocates them with torch); the library neither frees nor retains
 *     them; outputs are fully written by the call (no pre-zeroing needed unless
 *     a parameter says "accumulated");
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     performs no host synchronisation and no allocation;
 *   - tensors are contiguous row-major with the shapes given; int64 index
 *     tensors exactly as the reference produces them;
 *   - return value 0 = success; non-zero = error, message via s2t_last_error()
 *     (thread-local).  No C++ exceptions cross the boundary.
 *   - rnnt_type is always "regular"; boundary rows are [0, 0, S_b, T_b].
 */
#ifndef S2T_B200_H_
#define S2T_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2T_ABI_VERSION 3

/* dtype codes for logits tensors */
#define S2T_F32 0
#define S2T_BF16 1
#define S2T_F16 2
/* activation codes (reference: model/joiner/joiner.py:44-49) */
#define S2T_ACT_RELU 0
#define S2T_ACT_TANH 1
/* joiner arithmetic modes */
#define S2T_MODE_FP32_SIMT 0 /* strict fp32 FMA contractions (parity mode)        */
#define S2T_MODE_BF16_TC 1   /* bf16 operands, fp32 accumulate on tcgen05 / TMEM  */

int s2t_abi_version(void);
const char* s2t_last_error(void);

/* Measurement hooks used by bench.py (no effect on results).
 * s2t_launch_count: kernels launched by this library since load.
 * s2t_profile_enable(1): bracket every launch with CUDA events on its stream;
 * s2t_profile_report: wait for them, write "name\tcount\ttotal_ms\n" lines, clear. */
long long s2t_launch_count(void);
void s2t_profile_enable(int on);
int s2t_profile_report(char* buf, size_t n);

/* Bytes of lattice-DP scratch ("alpha_ws" below) for a lattice whose columns hold `slots`
 * symbol positions: slots = S+1 for the k2 layout, R for (B,T,R) band / full layouts. */
size_t s2t_lattice_workspace_bytes(int B, int S, int T, int slots);

/* ---------------------------------------------------------------------------
 * k2.mutual_information_recursion(px, py, boundary, return_grad)
 *   reference call sites: model/joiner/joiner.py:100-110 (inside rnnt_loss_smoothed),
 *   model/loss/pruned_rnnt_loss.py:39-48 (inside rnnt_loss_pruned).
 * px (B,S,T+1), py (B,S+1,T) fp32; boundary (B,4) int64 or NULL.
 * alpha_ws: scratch of s2t_lattice_workspace_bytes(B,S,T,S+1) bytes.  scores (B).  px_grad/py_grad: occupation
 * probabilities, same shapes as px/py, or both NULL for scores only.
 */
int s2t_mutual_information(const float* px, const float* py, const int64_t* boundary, int B, int S, int T,
                           void* alpha_ws, float* scores, float* px_grad, float* py_grad, void* stream);

/* ---------------------------------------------------------------------------
 * k2.rnnt_loss_smoothed(lm, am, symbols, termination_symbol, lm_only_scale,
 *                       am_only_scale, boundary, reduction, return_grad=True)
 *   reference call site: model/joiner/joiner.py:100-110.
 * am (B,T,V), lm (B,S+1,V) fp32; symbols (B,S) int64.
 * Outputs: am_max (B,T), lm_max (B,S+1), px (B,S,T+1), py (B,S+1,T),
 * nrm (B,S+1,T) [log-normalisers, kept for the backward], alpha_ws (scratch, as above),
 * scores (B) = log P(y|x) per utterance (reduction is the caller's),
 * px_grad (B,S,T+1), py_grad (B,S+1,T).
 * lm_only_scale / am_only_scale: k2's lm-only / am-only interpolation (JoinerConfig.lm_scale / am_scale);
 * both >= 0, sum < 1; 0 drops the term.
 * mode: S2T_MODE_FP32_SIMT = fp32 FMA contraction; S2T_MODE_BF16_TC = 3xF16 tensor-core
 * normaliser (fp32-level accuracy) and bf16 tensor-core backward contractions.
 * workspace: s2t_simple_workspace_bytes(mode,B,T,S,V) bytes.
 * row_max_ready != 0: am_max / lm_max already hold the row maxima of am / lm (by-products of s2t_linear_fwd);
 * the row-max pass over am and lm is skipped (tensor-core mode).  row_max_ready == 2: in addition
 * s2t_simple_loss_prep_lm has run on this workspace (the lm side of the normaliser: per-position records and the
 * split operand of exp(lm - max)) -- it depends on lm only, so a caller that computes lm on a second stream issues it
 * there, next to whatever produces am.
 */
size_t s2t_simple_workspace_bytes(int mode, int B, int T, int S, int V);
int s2t_simple_loss_fwd(int mode, const float* am, const float* lm, const int64_t* symbols, const int64_t* boundary,
                        int B, int T, int S, int V, int blank, float lm_only_scale, float am_only_scale,
                        float* am_max, float* lm_max, float* px, float* py, float* nrm, void* alpha_ws,
                        float* scores, float* px_grad, float* py_grad, void* workspace, int row_max_ready, void* stream);

/* Backward of the above: grad_scores (B) = d loss / d scores[b].
 * workspace: THE buffer the forward call wrote (in S2T_MODE_BF16_TC it holds the bf16 exp(am - max) /
 * exp(lm - max) operands that the forward pass produced as a by-product), unchanged.
 * d_am (B,T,V), d_lm (B,S+1,V) are overwritten. */
int s2t_simple_loss_prep_lm(int mode, const float* lm, const float* lm_max, const int64_t* symbols, int B, int T, int S,
                            int V, int blank, void* workspace, void* stream);
int s2t_simple_loss_bwd(int mode, const float* am, const float* lm, const int64_t* symbols, const float* am_max,
                        const float* lm_max, const float* nrm, const float* px_grad, const float* py_grad,
                        const float* grad_scores, int B, int T, int S, int V, int blank, float lm_only_scale,
                        float am_only_scale, void* workspace, float* d_am, float* d_lm, void* stream);

/* ---------------------------------------------------------------------------
 * k2.get_rnnt_prune_ranges(px_grad, py_grad, boundary, s_range)
 *   reference call site: model/joiner/joiner.py:112-117.
 * s_range must already be clamped by the caller (s_range > S -> S+1).
 * variant 0 = "A" (k2 v1.24.3 get_rnnt_prune_ranges), 1 = "B" (SURVEY.md A.4).
 * ranges (B,T,s_range) int64, bit-exact.
 */
int s2t_prune_ranges(const float* px_grad, const float* py_grad, const int64_t* boundary, int B, int S, int T,
                     int s_range, int variant, int64_t* ranges, void* stream);

/* ---------------------------------------------------------------------------
 * Loss on materialised logits (B,T,R,V) of dtype `dtype`:
 *   ranges != NULL: k2.rnnt_loss_pruned          (model/loss/pruned_rnnt_loss.py:39-48)
 *   ranges == NULL: torchaudio rnnt_loss, R = S+1 (model/loss/rnnt_loss.py:42-44)
 * Outputs: lse, px, py, occ_px, occ_py (B,T,R) fp32; alpha_ws scratch of
 * s2t_lattice_workspace_bytes(B,S,T,R) bytes;
 * scores (B) = log P(y|x).
 */
int s2t_logits_loss_fwd(const void* logits, int dtype, const int64_t* symbols, const int64_t* ranges,
                        const int64_t* boundary, int B, int T, int S, int R, int V, int blank,
                        float delay_penalty, float* lse, float* px, float* py, void* alpha_ws, float* scores,
                        float* occ_px, float* occ_py, void* stream);

/* grad (B,T,R,V), same dtype as logits, overwritten with d loss / d logits given
 * grad_scores[b] = d loss / d scores[b]:
 *   grad = grad_scores[b] * clip(occ_px [c==sym] + occ_py [c==blank] - (occ_px+occ_py) softmax_c)
 * where clip() is torchaudio's `clamp` (model/loss/rnnt_loss.py:27-29), applied when clamp > 0. */
int s2t_logits_loss_bwd(const void* logits, int dtype, const int64_t* symbols, const int64_t* ranges,
                        const float* lse, const float* occ_px, const float* occ_py, const float* grad_scores,
                        int B, int T, int S, int R, int V, int blank, float clamp, void* grad, void* stream);

/* ---------------------------------------------------------------------------
 * Fused pruned joiner + loss: never materialises (B,T,R,V).
 *   replaces k2.do_rnnt_pruning (joiner.py:121-123), add/act/out-projection
 *   (joiner.py:176-178) and k2.rnnt_loss_pruned (pruned_rnnt_loss.py:39-48);
 *   with ranges == NULL and R = S+1 it is the unpruned joiner (joiner.py:166-178)
 *   + torchaudio rnnt_loss (rnnt_loss.py:42-44).
 * am (B,T,V), lm (B,S+1,V) fp32 are the projected encoder / predictor outputs.
 * W1 (I,V), b1 (I), W2 (V,I), b2 (V): out-projection; I == 0 and NULL weights
 * when use_out_project=False.
 * workspace: s2t_joiner_workspace_bytes() bytes; the SAME buffer, untouched,
 * must be passed to the backward call (it carries the hidden activations).
 */
size_t s2t_joiner_workspace_bytes(int mode, int B, int T, int R, int V, int I);

int s2t_joiner_loss_fwd(int mode, const float* am, const float* lm, const int64_t* symbols,
                        const int64_t* ranges, const int64_t* boundary, const float* W1, const float* b1,
                        const float* W2, const float* b2, int B, int T, int S, int R, int V, int I, int act,
                        int blank, float delay_penalty, void* workspace, float* lse, float* px, float* py,
                        void* alpha_ws, float* scores, float* occ_px, float* occ_py, void* stream);

/* The two halves of s2t_joiner_loss_fwd as separate calls (same arguments, same results): the joiner's log-probs
 * lse / px / py (k2.get_rnnt_logprobs_pruned's role in k2.rnnt_loss_pruned) and the band lattice over them
 * (k2.mutual_information_recursion's role).  The lattice keeps one CTA per utterance busy and leaves the other SMs
 * idle, so a caller with independent work at hand -- the gradient contractions of the simple loss, which only need the
 * occupation probabilities of ITS lattice -- issues that work on a second stream between the two calls. */
int s2t_joiner_logprobs_fwd(int mode, const float* am, const float* lm, const int64_t* symbols,
                            const int64_t* ranges, const int64_t* boundary, const float* W1, const float* b1,
                            const float* W2, const float* b2, int B, int T, int S, int R, int V, int I, int act,
                            int blank, float delay_penalty, void* workspace, float* lse, float* px, float* py,
                            void* stream);
int s2t_band_lattice_fwd(const float* px, const float* py, const int64_t* ranges, const int64_t* boundary, int B,
                         int S, int T, int R, void* alpha_ws, float* scores, float* occ_px, float* occ_py,
                         void* stream);

/* out[0] = sum_i x[i] * w[i], n <= 1024: the loss reduction (-mean / -sum of the per-utterance scores, with a constant
 * weight vector) as one launch. */
int s2t_weighted_sum(const float* x, const float* w, int n, float* out, void* stream);

/* x0[g, 0..n0) and x1[g, 0..n1) (either may be NULL) are multiplied by num[g] / den[g] for every group g whose two
 * values differ -- groups with num == den cost no memory traffic -- and den[g] becomes (num[g] != 0 ? num[g] : 1).
 * Used for gradients that were computed ahead of the backward pass with a predicted upstream scale `den`
 * (the scale of the previous step): the backward pass corrects them with the actual scale `num`. */
int s2t_rescale_groups(float* x0, int64_t n0, float* x1, int64_t n1, int groups, const float* num, float* den,
                       void* stream);

/* d_am (B,T,V), d_lm (B,S+1,V), dW1, db1, dW2, db2 are overwritten. */
int s2t_joiner_loss_bwd(int mode, const float* am, const float* lm, const int64_t* symbols,
                        const int64_t* ranges, const int64_t* boundary, const float* W1, const float* b1,
                        const float* W2, const float* b2, int B, int T, int S, int R, int V, int I, int act,
                        int blank, float clamp, void* workspace, const float* lse, const float* occ_px,
                        const float* occ_py, const float* grad_scores, float* d_am, float* d_lm, float* dW1,
                        float* db1, float* dW2, float* db2, void* stream);

/* ---------------------------------------------------------------------------
 * nn.Linear for the joiner's projections in bf16 tensor-core mode
 *   reference: self._enc_proj / self._pre_proj, model/joiner/joiner.py:41-42, 148-149.
 * x (M,K), W (N,K), b (N) fp32 -> y (M,N) fp32 = x W^T + b: 3xF16 forward (fp32-level accuracy), bf16 operands in
 * the backward contractions, fp32 accumulation.
 * workspace: s2t_linear_workspace_bytes(M,N,K) bytes, the same buffer for fwd and bwd
 * (it carries the packed bf16 x and W^T).  bwd overwrites dx (may be NULL), dW, db.  The upstream gradient is
 * dy + dy2 (dy2 may be NULL): y feeds two consumers in this path (simple loss and joiner) and their two gradient
 * terms are added while dy is packed instead of by a separate pass over (M,N).
 */
size_t s2t_linear_workspace_bytes(int64_t M, int N, int K);
/* row_max (M) or NULL: by-product max_n y[m, n] from the epilogue: what the simple-loss normaliser
 * needs as am_max / lm_max, without a second pass over y (pass it on with row_max_ready = 1). */
/* x_dtype / dx_dtype: S2T_F32 or S2T_BF16 -- bf16 activations are read (and their gradient written) as such. */
int s2t_linear_fwd(const void* x, int x_dtype, const float* W, const float* b, int64_t M, int N, int K, void* workspace,
                   float* y, float* row_max, void* stream);
int s2t_linear_bwd(const float* dy, const float* dy2, const float* W, int64_t M, int N, int K, void* workspace,
                   void* dx, int dx_dtype, float* dW, float* db, void* stream);

/* Debug / materialised mode: write the logits (B,T,R,V) fp32 the fused path never stores. */
int s2t_joiner_materialize(int mode, const float* am, const float* lm, const int64_t* ranges, const float* W1,
                           const float* b1, const float* W2, const float* b2, int B, int T, int S, int R, int V,
                           int I, int act, void* workspace, float* logits, void* stream);

/* ---------------------------------------------------------------------------
 * CTC loss with the log-softmax fused in.
 * Replaces F.log_softmax + transpose + nn.CTCLoss of /root/reference/model/loss/ctc_loss.py:35-41 (the CTC branch of
 * PrunedRnntTask / CtcHybridRnnt, task_factory/rnnt_task.py:341-349, 485-496); the (T,B,V) log-probabilities are
 * never written.
 *   logits (B,T,V) fp32; targets (B,S) int64, padded; logit_lengths / target_lengths (B) int64.
 *   fwd: lse (B,T) and nll (B) = -log P(targets | logits) per utterance (+inf without a valid alignment);
 *        the workspace (s2t_ctc_workspace_bytes) holds the lattice and must reach bwd unchanged.
 *   bwd: grad_logits (B,T,V) = grad_nll[b] * d nll[b] / d logits (fully written; zero for padding frames and,
 *        with zero_infinity != 0, for utterances whose nll is infinite).
 * Reductions ("mean" divides by the target lengths as torch does) stay with the caller.
 * ------------------------------------------------------------------------- */
size_t s2t_ctc_workspace_bytes(int B, int T, int S, int V);
int s2t_ctc_loss_fwd(const float* logits, const int64_t* targets, const int64_t* logit_lengths,
                     const int64_t* target_lengths, int B, int T, int S, int V, int blank, void* workspace, float* lse,
                     float* nll, void* stream);
int s2t_ctc_loss_bwd(const float* logits, const int64_t* targets, const int64_t* logit_lengths,
                     const int64_t* target_lengths, int B, int T, int S, int V, int blank, const void* workspace,
                     const float* lse, const float* nll, const float* grad_nll, int zero_infinity, float* grad_logits,
                     void* stream);

/* ---------------------------------------------------------------------------
 * Stateless predictor front end: embedding lookup + depthwise Conv1d over the token context.
 * Replaces self._embedding + self._conv of /root/reference/model/predictor/stateless_predictor.py:90-97
 * (StatelessPredictor.forward); the Linear that follows goes through s2t_linear_fwd/bwd.
 *   emb (N,E) fp32 = _embedding.weight; conv_w (E,C) fp32 = _conv.weight viewed (E,1,C) -> (E,C);
 *   ctx (B,L) int64 = [state | blank | tokens], L = U + C; h (B, L-C+1, E):
 *   h[b,u,e] = sum_k conv_w[e,k] * emb[ctx[b,u+k], e].  Context sizes 1..8.
 *   bwd: d_emb (N,E) and d_conv_w (E,C) are fully written.
 * ------------------------------------------------------------------------- */
int s2t_predictor_embed_conv_fwd(const float* emb, const float* conv_w, const int64_t* ctx, int B, int L, int C, int E,
                                 int N, float* h, void* stream);
int s2t_predictor_embed_conv_bwd(const float* emb, const float* conv_w, const int64_t* ctx, const float* d_h, int B,
                                 int L, int C, int E, int N, float* d_emb, float* d_conv_w, void* stream);

/* ---------------------------------------------------------------------------
 * Batched greedy RNN-T decoding, device resident: one CTA per utterance walks its lattice.
 * Replaces the Python loop of RnntGreedyDecoding.decode (/root/reference/model/decoding.py:225-271), i.e. the
 * per-step calls of Joiner.streaming_step (model/joiner/joiner.py:184-207) and StatelessPredictor.streaming_step
 * (model/predictor/stateless_predictor.py:107-124), for a whole batch (batch_search, decoding.py:32-48).
 *   am (B,T,V) fp32 = enc_proj(encoder_out), computed by the caller; lengths (B) int64;
 *   predictor: emb (N,E), conv_w (E,C), Wo (D,E), bo (D); joiner: Wp (V,D), bp (V) [_pre_proj] and the
 *   out-projection W1 (I,V), b1, W2 (V,I), b2 (NULL / I = 0 when use_out_project = False).
 *   A frame emits at most max_token_step + 1 tokens (decoding.py:252).  tokens (B,max_out) int64, n_tokens (B) int32.
 * ------------------------------------------------------------------------- */
int s2t_rnnt_greedy_decode(const float* am, const int64_t* lengths, const float* emb, const float* conv_w, const float* Wo,
                           const float* bo, const float* Wp, const float* bp, const float* W1, const float* b1,
                           const float* W2, const float* b2, int B, int T, int V, int N, int E, int C, int D, int I, int act,
                           int blank, int max_token_step, int max_out, int64_t* tokens, int* n_tokens, void* stream);

/* Beam search of the same models (RnntBeamDecoding, model/decoding.py:295-425): at most one token per frame, every
 * beam proposes its top_k classes, the beam best candidates by accumulated log-probability survive (stable order,
 * hypotheses are not merged).  tokens (B, T) / n_tokens (B): the best hypothesis; best_score (B): its log-probability.
 * workspace: s2t_rnnt_beam_workspace_bytes(B, T, V, beam).  beam <= 8, top_k <= 8. */
size_t s2t_rnnt_beam_workspace_bytes(int B, int T, int V, int beam);
int s2t_rnnt_beam_decode(const float* am, const int64_t* lengths, const float* emb, const float* conv_w, const float* Wo,
                         const float* bo, const float* Wp, const float* bp, const float* W1, const float* b1,
                         const float* W2, const float* b2, int B, int T, int V, int N, int E, int C, int D, int I, int act,
                         int blank, int beam, int top_k, void* workspace, int64_t* tokens, int* n_tokens, float* best_score,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2T_B200_H_ */
