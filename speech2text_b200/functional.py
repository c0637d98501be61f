"""torch.autograd wrappers around the C ABI (include/s2t_b200.h).

PyTorch is plumbing here: it owns device memory, the current stream and the
autograd tape; all arithmetic on the hot path is in libs2t_b200.so.
Function names mirror the k2 / torchaudio calls of the reference they replace.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import check, lib, ptr, stream

PRUNE_VARIANTS = {"A": 0, "B": 1}


def _on_tensor_device(fn):
    """Run an autograd.Function's forward / backward with the device of its first CUDA tensor current: the C ABI
    takes raw pointers plus ``torch.cuda.current_stream()``, and the library's side streams are keyed by
    ``cudaGetDevice`` -- with tensors on a non-current device (one process driving several GPUs, or before
    ``torch.cuda.set_device``) the kernels would otherwise be launched on the wrong device and stream."""
    import functools

    @functools.wraps(fn)
    def wrapped(ctx, *args):
        dev = None
        for a in args + tuple(getattr(ctx, "saved_tensors", ()) or ()):
            if isinstance(a, Tensor) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise _lib.S2TError(f"tensors on different devices handed to one call: {dev} and {a.device}")
        if dev is None:
            return fn(ctx, *args)
        with torch.cuda.device(dev):
            return fn(ctx, *args)

    return wrapped


def _f32c(t: Tensor) -> Tensor:
    return t.to(torch.float32).contiguous()


def _i64c(t: Tensor) -> Tensor:
    return t.to(torch.int64).contiguous()


def _dp_scratch(B: int, S: int, T: int, slots: int, device) -> Tensor:
    return torch.empty((lib().s2t_lattice_workspace_bytes(B, S, T, slots),), dtype=torch.uint8, device=device)


def make_boundary(target_lengths: Tensor, encoder_out_lengths: Tensor, device) -> Tensor:
    """/root/reference/model/joiner/joiner.py:89-93 -- rows [0, 0, S_b, T_b], int64.
    Lengths may arrive as float tensors (joiner_test.py:56-58)."""
    B = target_lengths.shape[0]
    s = target_lengths.to(device=device, dtype=torch.int64)
    t = encoder_out_lengths.to(device=device, dtype=torch.int64)
    if not s.is_cuda:
        boundary = torch.zeros((B, 4), dtype=torch.int64, device=device)
        boundary[:, 2] = s
        boundary[:, 3] = t
        return boundary
    z = _zeros_i64(B, device)
    return torch.stack((z, z, s, t), dim=1)  # one kernel instead of a fill and two strided copies


# ---------------------------------------------------------------------------
# k2.mutual_information_recursion
# ---------------------------------------------------------------------------
def mutual_information_recursion(px: Tensor, py: Tensor, boundary: Optional[Tensor] = None,
                                 return_grad: bool = False):
    """Scores (B,) and, optionally, the occupation probabilities.  Not
    differentiable by itself (the fused losses below own the autograd edges)."""
    px, py = _f32c(px), _f32c(py)
    B, S, T1 = px.shape
    T = py.shape[-1]
    assert T1 == T + 1 and py.shape == (B, S + 1, T), (px.shape, py.shape)
    if boundary is not None:
        boundary = _i64c(boundary)
    alpha = _dp_scratch(B, S, T, S + 1, px.device)
    scores = torch.empty((B,), dtype=torch.float32, device=px.device)
    px_grad = torch.empty_like(px) if return_grad else None
    py_grad = torch.empty_like(py) if return_grad else None
    check(lib().s2t_mutual_information(ptr(px), ptr(py), ptr(boundary), B, S, T, ptr(alpha), ptr(scores),
                                       ptr(px_grad), ptr(py_grad), stream()))
    return (scores, (px_grad, py_grad)) if return_grad else scores


# ---------------------------------------------------------------------------
# k2.rnnt_loss_smoothed(..., return_grad=True)      joiner.py:100-110
# ---------------------------------------------------------------------------
_SIDE_STREAMS: dict = {}
_ONES: dict = {}
_ZEROS_I64: dict = {}


def _side_stream(device) -> "torch.cuda.Stream":
    s = _SIDE_STREAMS.get(device)
    if s is None:
        s = _SIDE_STREAMS[device] = torch.cuda.Stream(device=device)
    return s


def _zeros_i64(n: int, device) -> Tensor:
    key = (n, device)
    t = _ZEROS_I64.get(key)
    if t is None:
        t = _ZEROS_I64[key] = torch.zeros((n,), dtype=torch.int64, device=device)
    return t


def _ones(n: int, device) -> Tensor:
    key = (n, device)
    t = _ONES.get(key)
    if t is None:
        t = _ONES[key] = torch.ones((n,), dtype=torch.float32, device=device)
    return t


def _early_simple_backward(mode: int) -> bool:
    """S2T_B200_EARLY_SIMPLE_BWD: "0" off, "1" on; default: on in tensor-core mode.  Off while the library's
    per-launch event timer runs (side-stream launches would be timed against the wrong neighbours)."""
    v = os.environ.get("S2T_B200_EARLY_SIMPLE_BWD")
    on = (mode == _lib.MODE_BF16_TC) if v is None else v != "0"
    return on and not _lib.profiling()


_PENDING = []  # at most one _EarlySimpleBackward per device, waiting for its slot next to the band lattice
_PRED = {}     # (device, B) -> upstream scale of the simple loss's scores seen by the previous backward pass (never 0)


def _predicted_scale(B: int, device) -> Tensor:
    key = (device, B)
    t = _PRED.get(key)
    if t is None:
        t = _PRED[key] = torch.ones((B,), dtype=torch.float32, device=device)
    return t


class _EarlySimpleBackward:
    """The gradient of the simple loss needs nothing that the backward pass brings besides one scale per utterance:
    the occupation probabilities exist at the end of its forward call (k2 computes them there too, for the prune
    ranges).  Its contractions are therefore issued ahead of time on a side stream, in the one place of the step
    where most of the GPU idles -- next to the joiner's band lattice, which keeps one CTA per utterance busy --
    with the scale the previous step's backward pass saw; backward waits for them and corrects the scale where it
    differs (`s2t_rescale_groups`: no memory traffic for utterances whose scale was predicted right)."""

    def __init__(self, mode, tensors, dims, blank, scales):
        self.mode, self.tensors, self.dims, self.blank, self.scales = mode, tensors, dims, blank, scales
        am, lm = tensors[0], tensors[1]
        self.d_am, self.d_lm = torch.empty_like(am), torch.empty_like(lm)
        self.pred = _predicted_scale(dims[0], am.device)
        self.side = None

    def launch(self):
        am, lm, symbols, am_max, lm_max, nrm, px_grad, py_grad, ws = self.tensors
        B, T, S, V = self.dims
        dev = am.device
        cur, side = torch.cuda.current_stream(dev), _side_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            check(lib().s2t_simple_loss_bwd(self.mode, ptr(am), ptr(lm), ptr(symbols), ptr(am_max), ptr(lm_max),
                                            ptr(nrm), ptr(px_grad), ptr(py_grad), ptr(self.pred), B, T, S, V,
                                            self.blank, self.scales[0], self.scales[1], ptr(ws), ptr(self.d_am),
                                            ptr(self.d_lm), stream()))
        for t in self.tensors + (self.d_am, self.d_lm, self.pred):
            t.record_stream(side)
        self.side = side
        self.tensors = None


def _launch_pending(device) -> None:
    """Called where a latency-bound kernel is about to own the GPU (between the joiner's log-prob kernels and its
    band lattice)."""
    for e in [e for e in _PENDING if e.d_am.device == device]:
        _PENDING.remove(e)
        e.launch()


class _SimpleLoss(torch.autograd.Function):
    """k2.rnnt_loss_smoothed(return_grad=True); the backward contractions may run ahead of time
    (_EarlySimpleBackward)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, am: Tensor, lm: Tensor, symbols: Tensor, boundary: Tensor, blank: int,
                lm_only_scale: float, am_only_scale: float, mode: int, row_max=None, prepared_ws=None):
        am, lm = _f32c(am), _f32c(lm)
        symbols, boundary = _i64c(symbols), _i64c(boundary)
        B, T, V = am.shape
        S = lm.shape[1] - 1
        assert lm.shape == (B, S + 1, V) and symbols.shape == (B, S), (am.shape, lm.shape, symbols.shape)
        dev = am.device
        f32 = dict(dtype=torch.float32, device=dev)
        # the row maxima may arrive with am / lm (by-products of the projection epilogue)
        ready = (row_max is not None and mode == _lib.MODE_BF16_TC and row_max[0].shape == (B, T)
                 and row_max[1].shape == (B, S + 1))
        am_max = _f32c(row_max[0]) if ready else torch.empty((B, T), **f32)
        lm_max = _f32c(row_max[1]) if ready else torch.empty((B, S + 1), **f32)
        px = torch.empty((B, S, T + 1), **f32)
        py = torch.empty((B, S + 1, T), **f32)
        nrm = torch.empty((B, S + 1, T), **f32)
        alpha = _dp_scratch(B, S, T, S + 1, dev)
        scores = torch.empty((B,), **f32)
        px_grad = torch.empty((B, S, T + 1), **f32)
        py_grad = torch.empty((B, S + 1, T), **f32)
        nbytes = lib().s2t_simple_workspace_bytes(mode, B, T, S, V)
        # the lm side of the normaliser may already sit in the workspace (simple_loss_prepare_lm)
        lm_ready = ready and prepared_ws is not None and prepared_ws.numel() == nbytes and prepared_ws.device == dev
        ws = prepared_ws if lm_ready else torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        check(lib().s2t_simple_loss_fwd(mode, ptr(am), ptr(lm), ptr(symbols), ptr(boundary), B, T, S, V, blank,
                                        float(lm_only_scale), float(am_only_scale), ptr(am_max), ptr(lm_max),
                                        ptr(px), ptr(py), ptr(nrm), ptr(alpha), ptr(scores), ptr(px_grad),
                                        ptr(py_grad), ptr(ws), 2 if lm_ready else (1 if ready else 0), stream()))
        ctx.early = None
        if (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) and _early_simple_backward(mode):
            _PENDING[:] = [e for e in _PENDING if e.d_am.device != dev]  # an older one falls back to the plain path
            ctx.early = _EarlySimpleBackward(mode, (am, lm, symbols, am_max, lm_max, nrm, px_grad, py_grad, ws),
                                             (B, T, S, V), blank, (float(lm_only_scale), float(am_only_scale)))
            _PENDING.append(ctx.early)
        # the workspace travels to backward: in tensor-core mode it holds the bf16 exp(am - max) / exp(lm - max)
        # operands the forward pass produced on the side
        ctx.save_for_backward(am, lm, symbols, am_max, lm_max, nrm, px_grad, py_grad, ws)
        ctx.blank = blank
        ctx.mode = mode
        ctx.scales = (float(lm_only_scale), float(am_only_scale))
        ctx.mark_non_differentiable(px_grad, py_grad)
        ctx.set_materialize_grads(False)  # no zero-filled (B,S,T+1) / (B,S+1,T) gradients for the two lattices
        return scores, px_grad, py_grad

    @staticmethod
    @_on_tensor_device
    def backward(ctx, grad_scores, _gx, _gy):
        am, lm, symbols, am_max, lm_max, nrm, px_grad, py_grad, ws = ctx.saved_tensors
        B, T, V = am.shape
        S = lm.shape[1] - 1
        early, ctx.early = ctx.early, None
        if early is not None and early in _PENDING:
            _PENDING.remove(early)  # no lattice kernel came by: the plain path below
        if early is not None and early.side is not None:
            cur = torch.cuda.current_stream(am.device)
            cur.wait_stream(early.side)  # also what joins the side stream back into a CUDA-graph capture of the step
            if grad_scores is None:
                return (None,) * 10
            check(lib().s2t_rescale_groups(ptr(early.d_am), T * V, ptr(early.d_lm), (S + 1) * V, B,
                                           ptr(_f32c(grad_scores)), ptr(early.pred), stream()))
            return early.d_am, early.d_lm, None, None, None, None, None, None, None, None
        if grad_scores is None:
            return (None,) * 10
        grad_scores = _f32c(grad_scores)
        d_am = torch.empty_like(am)
        d_lm = torch.empty_like(lm)
        check(lib().s2t_simple_loss_bwd(ctx.mode, ptr(am), ptr(lm), ptr(symbols), ptr(am_max), ptr(lm_max), ptr(nrm),
                                        ptr(px_grad), ptr(py_grad), ptr(grad_scores), B, T, S, V, ctx.blank,
                                        ctx.scales[0], ctx.scales[1], ptr(ws), ptr(d_am), ptr(d_lm), stream()))
        if _early_simple_backward(ctx.mode):  # what the next step's early launch will assume
            check(lib().s2t_rescale_groups(None, 0, None, 0, B, ptr(grad_scores), ptr(_predicted_scale(B, am.device)),
                                           stream()))
        return d_am, d_lm, None, None, None, None, None, None, None, None


def simple_loss_prepare_lm(lm: Tensor, lm_max: Tensor, symbols: Tensor, T: int, blank: int, mode: int,
                           on_side_stream: bool = False) -> Optional[Tensor]:
    """The lm side of ``rnnt_loss_smoothed``'s normaliser (per-position records and the split operand of
    exp(lm - max)) ahead of the call: it needs lm and its row maxima only.  Returns the workspace to hand to
    ``rnnt_loss_smoothed(prepared_ws=...)``, or None when this mode has nothing to prepare.  ``on_side_stream``: launch
    behind a projection that was launched with ``on_side_stream=True`` (the caller joins with ``join_side_stream``)."""
    if mode != _lib.MODE_BF16_TC or not lm.is_cuda or lm.dtype != torch.float32 or not lm.is_contiguous():
        return None
    dev = lm.device
    with torch.cuda.device(dev):
        B, S1, V = lm.shape
        S = S1 - 1
        symbols = _i64c(symbols.to(dev))
        if symbols.shape != (B, S) or lm_max.shape != (B, S1):
            return None
        ws = torch.empty((lib().s2t_simple_workspace_bytes(mode, B, T, S, V),), dtype=torch.uint8, device=dev)
        if on_side_stream:
            side = _side_stream(dev)  # ordered behind the projection that produced lm on this stream
            for t in (lm, lm_max, symbols, ws):
                t.record_stream(side)
            st = ctypes.c_void_p(side.cuda_stream)
        else:
            st = stream()
        check(lib().s2t_simple_loss_prep_lm(mode, ptr(lm), ptr(_f32c(lm_max)), ptr(symbols), B, T, S, V, blank, ptr(ws), st))
    return ws


def rnnt_loss_smoothed(lm: Tensor, am: Tensor, symbols: Tensor, termination_symbol: int,
                       lm_only_scale: float = 0.0, am_only_scale: float = 0.0,
                       boundary: Optional[Tensor] = None, reduction: str = "mean",
                       return_grad: bool = False, mode: int = _lib.MODE_FP32_SIMT, row_max=None, prepared_ws=None):
    """Drop-in for ``k2.rnnt_loss_smoothed`` (rnnt_type='regular').  ``mode`` picks the arithmetic of
    the normaliser contraction: fp32 FMA, or tensor cores (3xTF32 forward, bf16 backward)."""
    if boundary is None:
        B, T = am.shape[0], am.shape[1]
        boundary = torch.zeros((B, 4), dtype=torch.int64, device=am.device)
        boundary[:, 2] = lm.shape[1] - 1
        boundary[:, 3] = T
    scores, px_grad, py_grad = _SimpleLoss.apply(am, lm, symbols, boundary, termination_symbol,
                                                 lm_only_scale, am_only_scale, mode, row_max, prepared_ws)
    loss = _reduce(scores, reduction)
    return (loss, (px_grad, py_grad)) if return_grad else loss


_REDUCE_WEIGHTS: dict = {}


class _ScaledSum(torch.autograd.Function):
    """sum_b scores[b] * w[b] for a constant weight vector w, in one launch."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, scores: Tensor, w: Tensor):
        scores = _f32c(scores)
        out = torch.empty((), dtype=torch.float32, device=scores.device)
        check(lib().s2t_weighted_sum(ptr(scores), ptr(w), scores.numel(), ptr(out), stream()))
        ctx.save_for_backward(w)
        return out

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        return g * w, None


def _reduce(scores: Tensor, reduction: str) -> Tensor:
    """-scores / -mean / -sum.  mean and sum are ONE dot product with a cached constant vector (-1/B or -1): the
    step is a chain of short kernels, and ``-torch.mean(x)`` is two of them forward and two more backward."""
    if reduction == "none":
        return -scores
    if reduction not in ("mean", "sum"):
        raise ValueError(f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")
    n = scores.numel()
    key = (reduction, n, scores.device, scores.dtype)
    w = _REDUCE_WEIGHTS.get(key)
    if w is None:
        w = torch.full((n,), -1.0 / n if reduction == "mean" else -1.0, dtype=scores.dtype, device=scores.device)
        _REDUCE_WEIGHTS[key] = w
    if scores.is_cuda and scores.dtype == torch.float32 and n <= 1024:
        return _ScaledSum.apply(scores.reshape(-1), w)  # one launch (cuBLAS' dot is two; a (1 x n) mv dispatches to it too)
    return torch.dot(scores.reshape(-1), w)


# ---------------------------------------------------------------------------
# k2.get_rnnt_prune_ranges                              joiner.py:112-117
# ---------------------------------------------------------------------------
def get_rnnt_prune_ranges(px_grad: Tensor, py_grad: Tensor, boundary: Tensor, s_range: int,
                          variant: Optional[str] = None) -> Tensor:
    variant = variant or os.environ.get("S2T_B200_PRUNE_VARIANT", "B")
    if variant not in PRUNE_VARIANTS:
        raise ValueError(f"S2T_B200_PRUNE_VARIANT must be A or B, got {variant}")
    px_grad, py_grad, boundary = _f32c(px_grad), _f32c(py_grad), _i64c(boundary)
    B, S, T1 = px_grad.shape
    T = py_grad.shape[-1]
    assert T1 == T + 1, "only rnnt_type='regular' lattices are supported"
    assert py_grad.shape == (B, S + 1, T), py_grad.shape
    assert boundary.shape == (B, 4), boundary.shape
    assert S >= 1, S
    assert T >= S, (T, S)
    if s_range > S:  # no symbol is pruned; keep indexing with ranges valid
        s_range = S + 1
    assert s_range >= 2, ("Pruning range for standard RNN-T should be equal to or greater than 2, "
                          "or no valid paths could survive pruning.")
    ranges = torch.empty((B, T, s_range), dtype=torch.int64, device=px_grad.device)
    check(lib().s2t_prune_ranges(ptr(px_grad), ptr(py_grad), ptr(boundary), B, S, T, s_range,
                                 PRUNE_VARIANTS[variant], ptr(ranges), stream()))
    return ranges


# ---------------------------------------------------------------------------
# loss on materialised logits: k2.rnnt_loss_pruned / torchaudio rnnt_loss
# ---------------------------------------------------------------------------
class _LogitsLoss(torch.autograd.Function):

    @staticmethod
    @_on_tensor_device
    def forward(ctx, logits: Tensor, symbols: Tensor, ranges: Optional[Tensor], boundary: Tensor,
                blank: int, delay_penalty: float, clamp: float):
        logits = logits.contiguous()
        symbols, boundary = _i64c(symbols), _i64c(boundary)
        B, T, R, V = logits.shape
        S = symbols.shape[1]
        if ranges is not None:
            ranges = _i64c(ranges)
            assert ranges.shape == (B, T, R), (ranges.shape, logits.shape)
        dev = logits.device
        f32 = dict(dtype=torch.float32, device=dev)
        lse, px, py = torch.empty((3, B, T, R), **f32).unbind(0)  # one buffer: the library zero-fills padding rows in one go
        occ_px = torch.empty((B, T, R), **f32)
        occ_py = torch.empty((B, T, R), **f32)
        alpha = _dp_scratch(B, S, T, R, dev)
        scores = torch.empty((B,), **f32)
        check(lib().s2t_logits_loss_fwd(ptr(logits), _lib.dtype_code(logits.dtype), ptr(symbols), ptr(ranges),
                                        ptr(boundary), B, T, S, R, V, blank, float(delay_penalty), ptr(lse),
                                        ptr(px), ptr(py), ptr(alpha), ptr(scores), ptr(occ_px), ptr(occ_py),
                                        stream()))
        ctx.save_for_backward(logits, symbols, ranges if ranges is not None else torch.empty(0, device=dev),
                              lse, occ_px, occ_py)
        ctx.has_ranges = ranges is not None
        ctx.blank, ctx.clamp, ctx.S = blank, clamp, S
        return scores

    @staticmethod
    @_on_tensor_device
    def backward(ctx, grad_scores):
        logits, symbols, ranges, lse, occ_px, occ_py = ctx.saved_tensors
        B, T, R, V = logits.shape
        grad = torch.empty_like(logits)
        check(lib().s2t_logits_loss_bwd(ptr(logits), _lib.dtype_code(logits.dtype), ptr(symbols),
                                        ptr(ranges if ctx.has_ranges else None), ptr(lse), ptr(occ_px),
                                        ptr(occ_py), ptr(_f32c(grad_scores)), B, T, ctx.S, R, V, ctx.blank,
                                        float(ctx.clamp), ptr(grad), stream()))
        return grad, None, None, None, None, None, None


def logits_scores(logits: Tensor, symbols: Tensor, ranges: Optional[Tensor], boundary: Tensor,
                  blank: int = 0, delay_penalty: float = 0.0, clamp: float = -1.0) -> Tensor:
    """log P(y|x) per utterance from a materialised (B,T,R,V) logits tensor."""
    return _LogitsLoss.apply(logits, symbols, ranges, boundary, blank, delay_penalty, clamp)


# ---------------------------------------------------------------------------
# fused joiner + loss (logits never materialised)
# ---------------------------------------------------------------------------
class _JoinerLoss(torch.autograd.Function):

    @staticmethod
    @_on_tensor_device
    def forward(ctx, am: Tensor, lm: Tensor, W1, b1, W2, b2, symbols: Tensor, ranges: Optional[Tensor],
                boundary: Tensor, act: int, blank: int, delay_penalty: float, clamp: float, mode: int,
                sinks=None):
        ctx.sink_params = sinks  # the four out-projection parameters when they may carry a bound gradient sink
        am, lm = _f32c(am), _f32c(lm)
        symbols, boundary = _i64c(symbols), _i64c(boundary)
        B, T, V = am.shape
        S = lm.shape[1] - 1
        dev = am.device
        if ranges is not None:
            ranges = _i64c(ranges)
            R = ranges.shape[2]
            assert ranges.shape == (B, T, R)
        else:
            R = S + 1
        has_proj = W1 is not None
        if has_proj:
            W1, b1, W2, b2 = _f32c(W1), _f32c(b1), _f32c(W2), _f32c(b2)
            I = W1.shape[0]
            assert W1.shape == (I, V) and W2.shape == (V, I) and b1.shape == (I,) and b2.shape == (V,)
        else:
            I = 0
        f32 = dict(dtype=torch.float32, device=dev)
        nbytes = lib().s2t_joiner_workspace_bytes(mode, B, T, R, V, I)
        workspace = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        lse = torch.empty((B, T, R), **f32)
        px = torch.empty((B, T, R), **f32)
        py = torch.empty((B, T, R), **f32)
        occ_px = torch.empty((B, T, R), **f32)
        occ_py = torch.empty((B, T, R), **f32)
        alpha = _dp_scratch(B, S, T, R, dev)
        scores = torch.empty((B,), **f32)
        check(lib().s2t_joiner_logprobs_fwd(mode, ptr(am), ptr(lm), ptr(symbols), ptr(ranges), ptr(boundary),
                                            ptr(W1), ptr(b1), ptr(W2), ptr(b2), B, T, S, R, V, I, act, blank,
                                            float(delay_penalty), ptr(workspace), ptr(lse), ptr(px), ptr(py),
                                            stream()))
        _launch_pending(dev)  # the simple loss's gradient contractions run next to the lattice kernel
        check(lib().s2t_band_lattice_fwd(ptr(px), ptr(py), ptr(ranges), ptr(boundary), B, S, T, R, ptr(alpha),
                                         ptr(scores), ptr(occ_px), ptr(occ_py), stream()))
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(am, lm, W1 if has_proj else empty, b1 if has_proj else empty,
                              W2 if has_proj else empty, b2 if has_proj else empty, symbols,
                              ranges if ranges is not None else empty, boundary, workspace, lse, occ_px, occ_py)
        ctx.meta = (has_proj, ranges is not None, S, R, I, act, blank, clamp, mode)
        return scores

    @staticmethod
    @_on_tensor_device
    def backward(ctx, grad_scores):
        (am, lm, W1, b1, W2, b2, symbols, ranges, boundary, workspace, lse, occ_px, occ_py) = ctx.saved_tensors
        has_proj, has_ranges, S, R, I, act, blank, clamp, mode = ctx.meta
        B, T, V = am.shape
        d_am = torch.empty_like(am)
        d_lm = torch.empty_like(lm)
        sinks = claim_grad_sinks(ctx.sink_params) if has_proj else None
        if sinks is not None:
            dW1, db1, dW2, db2 = sinks
        elif has_proj:
            dW1, db1, dW2, db2 = (torch.empty_like(W1), torch.empty_like(b1), torch.empty_like(W2),
                                  torch.empty_like(b2))
        else:
            W1 = b1 = W2 = b2 = dW1 = db1 = dW2 = db2 = None
        check(lib().s2t_joiner_loss_bwd(mode, ptr(am), ptr(lm), ptr(symbols), ptr(ranges if has_ranges else None),
                                        ptr(boundary), ptr(W1), ptr(b1), ptr(W2), ptr(b2), B, T, S, R, V, I, act,
                                        blank, float(clamp), ptr(workspace), ptr(lse), ptr(occ_px), ptr(occ_py),
                                        ptr(_f32c(grad_scores)), ptr(d_am), ptr(d_lm), ptr(dW1), ptr(db1),
                                        ptr(dW2), ptr(db2), stream()))
        if sinks is not None:
            dW1 = db1 = dW2 = db2 = None  # already where the optimizer / all-reduce reads them
        return (d_am, d_lm, dW1, db1, dW2, db2, None, None, None, None, None, None, None, None, None)


def joiner_scores(am: Tensor, lm: Tensor, W1, b1, W2, b2, symbols: Tensor, ranges: Optional[Tensor],
                  boundary: Tensor, act: int, blank: int = 0, delay_penalty: float = 0.0,
                  clamp: float = -1.0, mode: int = _lib.MODE_FP32_SIMT) -> Tensor:
    """log P(y|x) per utterance of the (pruned or full) joiner lattice, fused."""
    params = (W1, b1, W2, b2)
    return _JoinerLoss.apply(am, lm, W1, b1, W2, b2, symbols, ranges, boundary, act, blank, delay_penalty,
                             clamp, mode, params if all(grad_sink(p) is not None for p in params) else None)


def joiner_materialize(am: Tensor, lm: Tensor, W1, b1, W2, b2, ranges: Optional[Tensor], act: int,
                       mode: int = _lib.MODE_FP32_SIMT) -> Tensor:
    """The (B,T,R,V) logits the fused path never stores (debug / materialised mode, no autograd)."""
    am, lm = _f32c(am.detach()), _f32c(lm.detach())
    B, T, V = am.shape
    S = lm.shape[1] - 1
    if ranges is not None:
        ranges = _i64c(ranges)
        R = ranges.shape[2]
    else:
        R = S + 1
    I = 0
    if W1 is not None:
        W1, b1, W2, b2 = (_f32c(x.detach()) for x in (W1, b1, W2, b2))
        I = W1.shape[0]
    nbytes = lib().s2t_joiner_workspace_bytes(mode, B, T, R, V, I)
    workspace = torch.empty((nbytes,), dtype=torch.uint8, device=am.device)
    logits = torch.empty((B, T, R, V), dtype=torch.float32, device=am.device)
    check(lib().s2t_joiner_materialize(mode, ptr(am), ptr(lm), ptr(ranges), ptr(W1), ptr(b1), ptr(W2), ptr(b2),
                                       B, T, S, R, V, I, act, ptr(workspace), ptr(logits), stream()))
    return logits


# ---------------------------------------------------------------------------
# CTC loss with the log-softmax fused in               ctc_loss.py:35-41
# ---------------------------------------------------------------------------
class _CtcLoss(torch.autograd.Function):

    @staticmethod
    @_on_tensor_device
    def forward(ctx, logits: Tensor, targets: Tensor, logits_length: Tensor, targets_length: Tensor, blank: int,
                zero_infinity: bool):
        x = _f32c(logits)
        B, T, V = x.shape
        dev = x.device
        targets = _i64c(targets.to(dev))
        if targets.dim() != 2 or targets.shape[0] != B:
            raise ValueError("ctc_loss: targets must be a padded (B, S) tensor (as the reference's dataset produces)")
        S = targets.shape[1]
        in_len, tgt_len = _i64c(logits_length.to(dev)), _i64c(targets_length.to(dev))
        ws = torch.empty((lib().s2t_ctc_workspace_bytes(B, T, S, V),), dtype=torch.uint8, device=dev)
        lse = torch.empty((B, T), dtype=torch.float32, device=dev)
        nll = torch.empty((B,), dtype=torch.float32, device=dev)
        check(lib().s2t_ctc_loss_fwd(ptr(x), ptr(targets) if S > 0 else ptr(None), ptr(in_len), ptr(tgt_len), B, T, S, V,
                                     blank, ptr(ws), ptr(lse), ptr(nll), stream()))
        ctx.save_for_backward(x, targets, in_len, tgt_len, ws, lse, nll)
        ctx.meta = (S, blank, bool(zero_infinity), logits.dtype)
        return torch.where(torch.isinf(nll), torch.zeros_like(nll), nll) if zero_infinity else nll.clone()

    @staticmethod
    @_on_tensor_device
    def backward(ctx, grad_nll):
        x, targets, in_len, tgt_len, ws, lse, nll = ctx.saved_tensors
        S, blank, zero_infinity, in_dtype = ctx.meta
        B, T, V = x.shape
        grad = torch.empty_like(x)
        check(lib().s2t_ctc_loss_bwd(ptr(x), ptr(targets) if S > 0 else ptr(None), ptr(in_len), ptr(tgt_len), B, T, S, V,
                                     blank, ptr(ws), ptr(lse), ptr(nll), ptr(_f32c(grad_nll)), 1 if zero_infinity else 0,
                                     ptr(grad), stream()))
        return (grad if in_dtype == torch.float32 else grad.to(in_dtype)), None, None, None, None, None


def ctc_loss(logits: Tensor, targets: Tensor, logits_length: Tensor, targets_length: Tensor, blank: int = 0,
             reduction: str = "mean", zero_infinity: bool = True) -> Tensor:
    """``nn.CTCLoss(blank, reduction, zero_infinity)(F.log_softmax(logits, -1).transpose(0, 1), ...)`` on (B, T, V)
    logits, with the log-softmax fused in and its gradient written straight into d logits."""
    nll = _CtcLoss.apply(logits, targets, logits_length, targets_length, blank, zero_infinity)
    if reduction == "none":
        return nll
    if reduction == "sum":
        return nll.sum()
    if reduction == "mean":  # torch: mean over the batch of nll / target length
        return (nll / targets_length.to(nll.device).clamp_min(1).to(nll.dtype)).mean()
    raise ValueError(f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")


# ---------------------------------------------------------------------------
# stateless predictor front end: embedding + depthwise conv      stateless_predictor.py:90-97
# ---------------------------------------------------------------------------
class _PredictorEmbedConv(torch.autograd.Function):

    @staticmethod
    @_on_tensor_device
    def forward(ctx, tokens: Tensor, emb: Tensor, conv_w: Tensor):
        tokens = _i64c(tokens)
        emb = _f32c(emb)
        E = emb.shape[1]
        C = conv_w.shape[-1]
        w = _f32c(conv_w).reshape(E, C)
        B, L = tokens.shape
        N = emb.shape[0]
        h = torch.empty((B, L - C + 1, E), dtype=torch.float32, device=emb.device)
        check(lib().s2t_predictor_embed_conv_fwd(ptr(emb), ptr(w), ptr(tokens), B, L, C, E, N, ptr(h), stream()))
        ctx.save_for_backward(tokens, emb, w)
        ctx.w_shape = conv_w.shape
        return h

    @staticmethod
    @_on_tensor_device
    def backward(ctx, d_h):
        tokens, emb, w = ctx.saved_tensors
        B, L = tokens.shape
        N, E = emb.shape
        C = w.shape[1]
        d_emb = torch.empty_like(emb)
        d_w = torch.empty_like(w)
        check(lib().s2t_predictor_embed_conv_bwd(ptr(emb), ptr(w), ptr(tokens), ptr(_f32c(d_h)), B, L, C, E, N, ptr(d_emb),
                                                 ptr(d_w), stream()))
        return None, d_emb, d_w.reshape(ctx.w_shape)


def predictor_embed_conv(tokens: Tensor, emb: Tensor, conv_w: Tensor) -> Tensor:
    """``conv1d(embedding(tokens).transpose(1, 2), conv_w, groups=E).transpose(1, 2)`` for a depthwise ``conv_w`` of
    shape (E, 1, C): (B, L) tokens -> (B, L - C + 1, E), as one fused kernel."""
    return _PredictorEmbedConv.apply(tokens, emb, conv_w)


# ---------------------------------------------------------------------------
# nn.Linear on the tensor cores (joiner projections, bf16 mode)
# ---------------------------------------------------------------------------
class _LinearTC(torch.autograd.Function):
    """y = x W^T + b, returned twice (the second output is an alias of the first): the path feeds y to two
    consumers -- the simple loss and the joiner -- and with one output each their gradients arrive here
    separately, so that the sum is formed while dy is packed for the backward contractions instead of by a
    separate pass of autograd over two (M, N) tensors."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x: Tensor, W: Tensor, b: Tensor, sink_params=None, want_row_max: bool = False,
                on_side_stream: bool = False):
        ctx.sink_params = sink_params  # (W, b) parameters when they may carry a bound gradient sink
        lead = x.shape[:-1]
        K = x.shape[-1]
        # fp32 or bf16 activations are read as they are (anything else goes through fp32)
        x2 = (x if x.dtype in (torch.float32, torch.bfloat16) else x.float()).contiguous().reshape(-1, K)
        W, b = _f32c(W), _f32c(b)
        M, N = x2.shape[0], W.shape[0]
        ws = torch.empty((lib().s2t_linear_workspace_bytes(M, N, K),), dtype=torch.uint8, device=x.device)
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        # by-product of the epilogue: max over the output features of every row
        row_max = torch.empty((M,), dtype=torch.float32, device=x.device) if want_row_max else None
        if on_side_stream:
            # The caller has a second, independent projection to launch (the predictor side is 0.7 of a wave of
            # tiles, the encoder side 2.7: together 3.4 instead of 1 + 3 rounds) and joins with join_side_stream().
            # Everything is allocated on the current stream; the side stream only runs the kernels.
            cur, side = torch.cuda.current_stream(x.device), _side_stream(x.device)
            side.wait_stream(cur)
            for t in (x2, W, b, ws, y, row_max):
                if t is not None:
                    t.record_stream(side)
            st = ctypes.c_void_p(side.cuda_stream)
        else:
            st = stream()
        check(lib().s2t_linear_fwd(ptr(x2), _lib.dtype_code(x2.dtype), ptr(W), ptr(b), M, N, K, ptr(ws), ptr(y),
                                   ptr(row_max), st))
        ctx.save_for_backward(W, ws)
        ctx.set_materialize_grads(False)  # no zero-filled gradient for the row-max by-product (or an unused alias)
        ctx.dims = (M, N, K, lead, x.requires_grad, x.dtype)
        y = y.reshape(*lead, N)
        if want_row_max:
            row_max = row_max.reshape(*lead)
            ctx.mark_non_differentiable(row_max)
            return y, y.view_as(y), row_max
        return y, y.view_as(y)

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dy, dy_alias, _row_max_grad=None):
        W, ws = ctx.saved_tensors
        M, N, K, lead, need_dx, x_dtype = ctx.dims
        if dy is None:
            dy, dy_alias = dy_alias, None
        if dy is None:
            return None, None, None, None, None, None
        sinks = claim_grad_sinks(ctx.sink_params)
        sink_W, sink_b = sinks if sinks is not None else (None, None)
        dy2 = _f32c(dy).reshape(M, N)
        dy3 = _f32c(dy_alias).reshape(M, N) if dy_alias is not None else None
        dx_dtype = x_dtype if x_dtype == torch.bfloat16 else torch.float32  # bf16 activations: bf16 gradient, written as such
        dx = torch.empty((M, K), dtype=dx_dtype, device=dy.device) if need_dx else None
        # a bound gradient buffer (FlatGradBucket.bind) is written in place and autograd gets no gradient to add
        dW = sink_W if sink_W is not None else torch.empty_like(W)
        db = sink_b if sink_b is not None else torch.empty((N,), dtype=torch.float32, device=dy.device)
        check(lib().s2t_linear_bwd(ptr(dy2), ptr(dy3), ptr(W), M, N, K, ptr(ws), ptr(dx), _lib.dtype_code(dx_dtype), ptr(dW),
                                   ptr(db), stream()))
        if need_dx and x_dtype != dx_dtype:
            dx = dx.to(x_dtype)
        return ((dx.reshape(*lead, K) if need_dx else None), None if sink_W is not None else dW,
                None if sink_b is not None else db, None, None, None)


def grad_sink(p: Optional[Tensor]) -> Optional[Tensor]:
    """The buffer ``FlatGradBucket.bind`` attached to a parameter (a view of the flat all-reduce buffer), if any:
    the backward kernels then write the gradient there instead of handing a temporary to autograd's accumulate."""
    rec = getattr(p, "_s2t_grad_sink", None) if p is not None else None
    if rec is None:
        return None
    sink = rec[0]
    if sink.shape != p.shape or sink.dtype != torch.float32 or not sink.is_contiguous() or sink.device != p.device:
        raise ValueError("gradient sink must be a contiguous fp32 tensor of the parameter's shape on its device")
    return sink


def claim_grad_sinks(params):
    """Called by a backward pass that wants to OVERWRITE the bound gradient buffers of ``params``.  Returns the
    sinks only if that is what autograd's accumulate would have produced: every ``p.grad`` still aliases its
    sink (``optimizer.zero_grad(set_to_none=True)`` drops the alias: the kernels would then write into an orphaned
    buffer and the optimizer would skip the parameter) and no earlier backward pass has written the sink since the
    last ``FlatGradBucket.zero()`` (gradient accumulation, a second joiner call in one step).  Otherwise None: the
    caller returns its gradients to autograd, which accumulates them as usual."""
    if params is None:
        return None
    recs = []
    for p in params:
        rec = getattr(p, "_s2t_grad_sink", None) if p is not None else None
        if rec is None:
            return None
        sink, bucket = rec
        if p.grad is None or p.grad.data_ptr() != sink.data_ptr() or p.grad.shape != sink.shape:
            return None
        if bucket is not None and id(p) in bucket.written:
            return None
        recs.append((p, sink, bucket))
    for p, _, bucket in recs:
        if bucket is not None:
            bucket.written.add(id(p))
    return tuple(sink for _, sink, _ in recs)


def linear_tc(x: Tensor, W: Tensor, b: Tensor) -> Tensor:
    """``F.linear(x, W, b)`` on tcgen05 (3xF16 split forward, bf16 backward); fp32 in, fp32 out."""
    return linear_tc_pair(x, W, b)[0]


def linear_tc_pair(x: Tensor, W: Tensor, b: Tensor, row_max: bool = False, on_side_stream: bool = False):
    """Same, as two aliases of the result for two consumers (see _LinearTC).  ``row_max=True`` appends the per-row
    maximum of the result (a by-product of the epilogue) as a third output.  ``on_side_stream=True`` launches the
    kernels on the library's side stream: the caller issues its other independent work and then calls
    ``join_side_stream`` before anything reads the results."""
    params = (W, b)
    return _LinearTC.apply(x, W, b, params if all(grad_sink(p) is not None for p in params) else None, row_max,
                           on_side_stream)


def join_side_stream(device) -> None:
    """The current stream waits for what was launched with ``on_side_stream=True``."""
    torch.cuda.current_stream(device).wait_stream(_side_stream(device))
