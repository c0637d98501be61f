// Problem descriptor shared by the fused joiner + loss implementations.
#pragma once
#include "common.cuh"

namespace s2t {

// Rows are m = (b, t, r), r < R.  Slot r of frame t is symbol position
// ranges[b, t, r] (pruned, /root/reference/model/joiner/joiner.py:112-123) or r itself when
// ranges == nullptr (unpruned joiner, joiner.py:166-176, R = S + 1).
// logits[m, :] = W2 (W1 act(am[b,t,:] + lm[b,s,:]) + b1) + b2     (I > 0, joiner.py:51-57)
//             = act(am[b,t,:] + lm[b,s,:])                         (I == 0, use_out_project=False)
struct JoinerProblem {
  const float* am;          // (B, T, V)
  const float* lm;          // (B, S+1, V)
  const int64_t* sym;       // (B, S)
  const int64_t* ranges;    // (B, T, R) or nullptr
  const int64_t* boundary;  // (B, 4)
  const float* W1;          // (I, V)
  const float* b1;          // (I)
  const float* W2;          // (V, I)
  const float* b2;          // (V)
  int B, T, S, R, V, I;
  int act;    // Activation
  int blank;  // termination symbol
  float delay_penalty;
};

size_t joiner_simt_workspace_bytes(int64_t M, int V, int I, int64_t* chunk_rows_out);
int joiner_simt_forward(const JoinerProblem& p, void* workspace, float* lse, float* px, float* py,
                        cudaStream_t stream);
int joiner_simt_materialize(const JoinerProblem& p, void* workspace, float* logits_out, cudaStream_t stream);
int joiner_simt_backward(const JoinerProblem& p, void* workspace, const float* lse, const float* occ_px,
                         const float* occ_py, const float* coef, float clamp, float* d_am, float* d_lm,
                         float* dW1, float* db1, float* dW2, float* db2, cudaStream_t stream);

size_t joiner_tc_workspace_bytes(int64_t M, int V, int I);
int joiner_tc_forward(const JoinerProblem& p, void* workspace, float* lse, float* px, float* py,
                      cudaStream_t stream);
int joiner_tc_backward(const JoinerProblem& p, void* workspace, const float* lse, const float* occ_px,
                       const float* occ_py, const float* coef, float clamp, float* d_am, float* d_lm,
                       float* dW1, float* db1, float* dW2, float* db2, cudaStream_t stream);

}  // namespace s2t
