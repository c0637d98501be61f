// Lattice views shared by the DP kernels and their callers.
//
// A "lattice view" describes where the log-probs of the two RNN-T transitions
// live in HBM.  Cell (s, t) emits symbol s (-> (s+1, t), log-prob px) or blank
// (-> (s, t+1), log-prob py).  Frame t only holds `slots` consecutive symbol
// positions starting at sb(t) = ranges[b, t, 0] (0 when ranges == nullptr):
//   px(s, t) = px[b*px_bs + t*px_ts + (s - sb(t))*px_rs]   if 0 <= s-sb(t) < rx
//   py(s, t) = py[b*py_bs + t*py_ts + (s - sb(t))*py_rs]   if 0 <= s-sb(t) < ry
// and -inf otherwise.  The same strides address the occupation outputs; alpha
// uses (a_bs, a_ts, a_rs) with ry slots and T+1 columns.
//
//   simple / smoothed lattice (k2 layout):
//       px (B,S,T+1): px_ts=1, px_rs=T+1, rx=S ; py (B,S+1,T): py_ts=1, py_rs=T, ry=S+1
//   pruned band: px,py (B,T,R): ts=R, rs=1, rx=ry=R, ranges (B,T,R)
//   vanilla full lattice (Appendix B): px,py (B,T,U+1): ts=U+1, rs=1, rx=ry=U+1
#pragma once
#include "common.cuh"

namespace s2t {

struct LatticeView {
  const float* px;
  const float* py;
  int64_t px_bs, px_ts, px_rs;
  int64_t py_bs, py_ts, py_rs;
  int rx, ry;
  bool px_at_tb;  // k2 layout: px has T+1 columns and the column t == T_b is read (k2 stores -inf there);
                  // band layouts: false, no symbol may be emitted after the last frame
  const int64_t* ranges;  // (B, T, R) or nullptr
  int64_t rg_bs, rg_ts;
  const int64_t* boundary;  // (B, 4) [0, 0, S_b, T_b] or nullptr
  int B, S, T;
  // scratch, carved from one workspace by lattice_carve_workspace()
  float* alpha;  // alpha relative to aoff[diagonal]
  int64_t a_bs, a_ts, a_rs;
  double* aoff;    // (B, S+T+2) fp64 offset of every alpha diagonal
  double* logp_d;  // (B) log P(y|x) in fp64
};

// bytes of DP scratch for alpha with `slots` symbol positions per column (T+1 columns)
size_t lattice_workspace_bytes(int B, int S, int T, int slots);
void lattice_carve_workspace(LatticeView& v, void* ws, int slots);

// Runs alpha, then beta + occupation.  occ_px / occ_py use the px / py strides
// and must be zero-filled by the caller (cells outside the live region are not
// visited).  logp[b] = log P(y|x) (k2: mutual_information_recursion scores).
int launch_lattice_fwd_bwd(const LatticeView& v, float* logp, float* occ_px, float* occ_py,
                           cudaStream_t stream);
// alpha only (scores without gradients)
int launch_lattice_fwd(const LatticeView& v, float* logp, cudaStream_t stream);

// warp-synchronous fast paths (lattice_fast.cu)
bool simple_lattice_fast_ok(int S, int T);
// unpruned lattice handed over in band layout (B, T, S+1): transposed onto the systolic full-lattice kernel
bool full_lattice_fast_ok(int S, int T);
size_t full_lattice_fast_workspace_bytes(int B, int S, int T);
int launch_full_lattice_fast(const float* px, const float* py, const int64_t* boundary, int B, int S, int T, void* ws,
                             float* logp, float* occ_px, float* occ_py, cudaStream_t stream);
size_t simple_lattice_fast_workspace_bytes(int B, int S, int T);
// occ_px / occ_py may both be nullptr (scores only); otherwise every element of them is written
int launch_simple_lattice_fast(const float* px, const float* py, const int64_t* boundary, int B, int S, int T,
                               void* ws, float* logp, float* occ_px, float* occ_py, cudaStream_t stream);
bool band_lattice_fast_ok(int S, int T, int R);
int launch_band_lattice_fast(const float* px, const float* py, const int64_t* ranges, const int64_t* boundary,
                             int B, int S, int T, int R, float* logp, float* occ_px, float* occ_py,
                             cudaStream_t stream);

}  // namespace s2t
