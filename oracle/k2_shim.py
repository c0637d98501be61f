"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the k2
functions the reference's transducer-loss path calls.

Reference call sites (all under /root/reference):
  * model/joiner/joiner.py:100-110   k2.rnnt_loss_smoothed(..., return_grad=True)
  * model/joiner/joiner.py:112-117   k2.get_rnnt_prune_ranges
  * model/joiner/joiner.py:121-123   k2.do_rnnt_pruning
  * model/loss/pruned_rnnt_loss.py:39-48   k2.rnnt_loss_pruned

k2 is an un-vendored third-party dependency (requirements.txt:4 pins
``k2==1.24.3.dev20240615+cuda11.6.torch1.13.1``; Dockerfile.build:28-35 builds
tag ``v1.24.3``) and cannot be imported or built in this image, and the
reference's own tests assert no numeric result for this path
(pruned_rnnt_loss_test.py:45-46 only logs).  Hence: **PARITY UNPINNED at the
k2 boundary**.  What pins this restatement instead:
  * torchaudio's compiled ``rnnt_loss`` (the reference's own vanilla back end,
    model/loss/rnnt_loss.py:27-29) agrees with ``rnnt_loss_smoothed`` on the
    trivial joiner ``am + lm`` and with ``rnnt_loss_pruned`` whenever
    ``s_range >= S + 1``  (tests/test_oracle.py);
  * fp64 autograd through an independent loop DP agrees with the occupation
    probabilities returned here (tests/test_oracle.py).

The module mirrors k2's public names so that it can be installed as
``sys.modules["k2"]`` and the reference's ``joiner.py`` /
``pruned_rnnt_loss.py`` run verbatim on top of it (oracle/make_golden.py).
The arithmetic follows k2/python/k2/rnnt_loss.py and mutual_information.py as
published at tag v1.24.3.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline /
reference arm may import this module.  The product package
(``speech2text_b200``) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libs2t_oracle.so")
_lib = None


def build_native(force: bool = False) -> str:
    """Compile oracle/mutual_information.c with gcc (a few hundred ms)."""
    src = os.path.join(_HERE, "mutual_information.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call([
            "gcc", "-O2", "-fPIC", "-shared", "-fno-fast-math", "-o", _LIB_PATH,
            src, "-lm"
        ])
    return _LIB_PATH


def _native():
    global _lib
    if _lib is None:
        build_native()
        _lib = ctypes.CDLL(_LIB_PATH)
        for sfx in ("f32", "f64"):
            getattr(_lib, f"s2t_oracle_mi_forward_{sfx}").restype = None
            getattr(_lib, f"s2t_oracle_mi_backward_{sfx}").restype = None
    return _lib


def _ptr(t: Optional[Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


# --------------------------------------------------------------------------
# mutual_information.py
# --------------------------------------------------------------------------
def _over_utterances(call, B: int) -> None:
    """k2's CPU kernel is one serial loop over the batch; utterances are independent, so the all-thread figure of
    the CPU baseline splits [0, B) over ``torch.get_num_threads()`` host threads (ctypes releases
    the GIL).  ``torch.set_num_threads(1)`` gives k2's own serial schedule.  Results do not depend on the split."""
    n = max(1, min(B, torch.get_num_threads()))
    if n == 1:
        call(0, B)
        return
    import concurrent.futures
    per = -(-B // n)
    with concurrent.futures.ThreadPoolExecutor(max_workers=n) as pool:
        list(pool.map(lambda lo: call(lo, min(B, lo + per)), range(0, B, per)))


def _mi_forward(px: Tensor, py: Tensor, boundary: Optional[Tensor],
                p: Tensor) -> Tensor:
    B, S, T1 = px.shape
    T = py.shape[-1]
    assert T1 == T + 1, "oracle restates the 'regular' topology only"
    ans = torch.empty(B, dtype=px.dtype)
    sfx = {torch.float32: "f32", torch.float64: "f64"}[px.dtype]
    fn = getattr(_native(), f"s2t_oracle_mi_forward_{sfx}")
    _over_utterances(lambda lo, hi: fn(_ptr(px), _ptr(py), _ptr(boundary), _ptr(p), _ptr(ans), ctypes.c_int(B),
                                       ctypes.c_int(S), ctypes.c_int(T), ctypes.c_int(lo), ctypes.c_int(hi)), B)
    return ans


def _mi_backward(px: Tensor, py: Tensor, boundary: Optional[Tensor], p: Tensor,
                 ans_grad: Tensor) -> Tuple[Tensor, Tensor]:
    B, S, T1 = px.shape
    T = py.shape[-1]
    p_grad = torch.zeros(B, S + 1, T + 1, dtype=px.dtype)
    px_grad = torch.zeros_like(px)
    py_grad = torch.zeros_like(py)
    sfx = {torch.float32: "f32", torch.float64: "f64"}[px.dtype]
    fn = getattr(_native(), f"s2t_oracle_mi_backward_{sfx}")
    _over_utterances(lambda lo, hi: fn(_ptr(px), _ptr(py), _ptr(boundary), _ptr(p), _ptr(ans_grad), _ptr(p_grad),
                                       _ptr(px_grad), _ptr(py_grad), ctypes.c_int(B), ctypes.c_int(S),
                                       ctypes.c_int(T), ctypes.c_int(lo), ctypes.c_int(hi)), B)
    return px_grad, py_grad


class MutualInformationRecursionFunction(torch.autograd.Function):
    """k2/python/k2/mutual_information.py: forward computes p and, when any
    gradient is needed, the occupation probabilities at once; backward scales
    them by the incoming per-utterance gradient (in place, as k2 does)."""

    @staticmethod
    def forward(ctx, px, py, pxy_grads, boundary=None, return_grad=False):
        B, S, T1 = px.shape
        T = py.shape[-1]
        assert T1 in (T, T + 1)
        assert py.shape == (B, S + 1, T)
        if boundary is not None:
            assert boundary.shape == (B, 4)
        # k2 leaves p uninitialised outside the boundary rectangle (torch.empty);
        # use zeros so the oracle is deterministic.
        p = torch.zeros(B, S + 1, T + 1, dtype=px.dtype)
        ans = _mi_forward(px, py, boundary, p)
        px_grad, py_grad = None, None
        if return_grad or px.requires_grad or py.requires_grad:
            ans_grad = torch.ones(B, dtype=px.dtype)
            px_grad, py_grad = _mi_backward(px, py, boundary, p, ans_grad)
            ctx.save_for_backward(px_grad, py_grad)
        assert len(pxy_grads) == 2
        pxy_grads[0] = px_grad
        pxy_grads[1] = py_grad
        return ans

    @staticmethod
    def backward(ctx, ans_grad):
        px_grad, py_grad = ctx.saved_tensors
        (B,) = ans_grad.shape
        ans_grad = ans_grad.reshape(B, 1, 1)
        px_grad = px_grad * ans_grad
        py_grad = py_grad * ans_grad
        return px_grad, py_grad, None, None, None


def mutual_information_recursion(
    px: Tensor,
    py: Tensor,
    boundary: Optional[Tensor] = None,
    return_grad: bool = False,
) -> Union[Tuple[Tensor, Tuple[Tensor, Tensor]], Tensor]:
    assert px.ndim == 3
    B, S, T1 = px.shape
    T = py.shape[-1]
    assert px.shape[-1] in (T, T + 1)
    assert py.shape == (B, S + 1, T)
    assert px.dtype == py.dtype
    if boundary is not None:
        assert boundary.dtype == torch.int64
        assert boundary.shape == (B, 4)
        for s_begin, t_begin, s_end, t_end in boundary.tolist():
            assert 0 <= s_begin <= s_end <= S
            assert 0 <= t_begin <= t_end <= T
    px, py = px.contiguous(), py.contiguous()
    pxy_grads: List[Optional[Tensor]] = [None, None]
    scores = MutualInformationRecursionFunction.apply(px, py, pxy_grads,
                                                      boundary, return_grad)
    px_grad, py_grad = pxy_grads
    return (scores, (px_grad, py_grad)) if return_grad else scores


# --------------------------------------------------------------------------
# ops.py: monotonic_lower_bound
# --------------------------------------------------------------------------
def monotonic_lower_bound(src: Tensor, inplace: bool = False) -> Tensor:
    """y[..., i] = min(x[..., i], x[..., i+1], ...): the largest non-decreasing
    sequence that is <= src element-wise (k2.monotonic_lower_bound)."""
    assert src.ndim in (1, 2)
    flipped = torch.flip(src, dims=[-1])
    out = torch.flip(torch.cummin(flipped, dim=-1).values, dims=[-1])
    if inplace:
        src.copy_(out)
        return src
    return out.contiguous()


# --------------------------------------------------------------------------
# rnnt_loss.py
# --------------------------------------------------------------------------
def fix_for_boundary(px: Tensor, boundary: Optional[Tensor] = None) -> Tensor:
    """px[b, :, boundary[b, 3]] = -inf (no symbol may be emitted after the
    last frame); only for px of shape (B, S, T+1)."""
    if boundary is None:
        return px
    B, S, T1 = px.shape
    boundary = boundary[:, 3].reshape(B, 1, 1).expand(B, S, T1)
    return px.scatter_(dim=2, index=boundary, value=float("-inf"))


def get_rnnt_logprobs_smoothed(
    lm: Tensor,
    am: Tensor,
    symbols: Tensor,
    termination_symbol: int,
    lm_only_scale: float = 0.1,
    am_only_scale: float = 0.1,
    boundary: Optional[Tensor] = None,
    rnnt_type: str = "regular",
) -> Tuple[Tensor, Tensor]:
    assert lm.ndim == 3, lm.ndim
    assert am.ndim == 3, am.ndim
    assert lm.shape[0] == am.shape[0], (lm.shape[0], am.shape[0])
    assert lm.shape[2] == am.shape[2], (lm.shape[2], am.shape[2])
    (B, T, C) = am.shape
    S = lm.shape[1] - 1
    assert symbols.shape == (B, S), symbols.shape
    assert S >= 0, S
    assert rnnt_type == "regular", "oracle restates rnnt_type='regular' only"

    # Caution: some parts of this code are a little less clear than they could
    # be due to optimizations (upstream comment): exp of stabilised inputs and
    # one batched matmul give the log-normalisers of the trivial joiner.
    am_max, _ = torch.max(am, dim=2, keepdim=True)  # (B, T, 1)
    lm_max, _ = torch.max(lm, dim=2, keepdim=True)  # (B, S+1, 1)
    am_probs = (am - am_max).exp()  # (B, T, C)
    lm_probs = (lm - lm_max).exp()  # (B, S+1, C)
    tiny = torch.finfo(lm_probs.dtype).tiny
    normalizers = (torch.matmul(lm_probs, am_probs.transpose(1, 2)) +
                   tiny).log()  # (B, S+1, T)

    # normalisers for the am-only and lm-only interpolation terms
    lmonly_normalizers = lm_probs.sum(dim=2, keepdim=True)  # (B, S+1, 1)
    unigram_lm = (torch.mean(lm_probs / lmonly_normalizers, dim=(0, 1),
                             keepdim=True) + tiny)  # (1, 1, C)
    amonly_normalizers = (torch.mv(am_probs.reshape(-1, C),
                                   unigram_lm.reshape(C)).reshape(
                                       B, T, 1).log() + am_max)  # (B, T, 1)
    amonly_normalizers = amonly_normalizers.transpose(1, 2)  # (B, 1, T)
    unigram_lm = unigram_lm.log()
    lmonly_normalizers = (lmonly_normalizers.log() + lm_max)  # (B, S+1, 1)

    # add lm_max and am_max to normalizers, to make it as if we had not
    # subtracted am_max and lm_max above.
    normalizers = normalizers + lm_max + am_max.transpose(1, 2)  # (B, S+1, T)

    # px is the probs of the actual symbols (not yet normalized)..
    px_am = torch.gather(
        am.unsqueeze(1).expand(B, S, T, C),
        dim=3,
        index=symbols.reshape(B, S, 1, 1).expand(B, S, T, 1),
    ).squeeze(-1)  # (B, S, T)
    px_am = torch.cat(
        (px_am,
         torch.full((B, S, 1), float("-inf"), device=px_am.device,
                    dtype=px_am.dtype)),
        dim=2,
    )  # now (B, S, T+1), index [:, :, T] has -inf
    px_lm = torch.gather(lm[:, :S], dim=2,
                         index=symbols.unsqueeze(-1))  # (B, S, 1)
    px_lm_unigram = torch.gather(unigram_lm.expand(B, S, C), dim=2,
                                 index=symbols.unsqueeze(-1))  # (B, S, 1)

    px = px_am + px_lm  # (B, S, T+1), last one is -infinity
    px_amonly = px_am + px_lm_unigram  # (B, S, T+1)
    px_lmonly = px_lm - lmonly_normalizers[:, :S, :]  # (B, S, 1)

    px[:, :, :T] -= normalizers[:, :S, :]
    px_amonly[:, :, :T] -= amonly_normalizers

    # py is the probs of termination symbols
    py_am = am[:, :, termination_symbol].unsqueeze(1)  # (B, 1, T)
    py_lm = lm[:, :, termination_symbol].unsqueeze(2)  # (B, S+1, 1)
    py = py_am + py_lm - normalizers

    py_lm_unigram = unigram_lm[0][0][termination_symbol]  # scalar
    py_amonly = py_am + py_lm_unigram - amonly_normalizers  # (B, 1, T)
    py_lmonly = py_lm - lmonly_normalizers  # (B, S+1, 1)

    combined_scale = 1.0 - lm_only_scale - am_only_scale

    # We need to avoid exact zeros in the scales because otherwise multiplying
    # -inf by zero generates nan.
    if lm_only_scale == 0.0:
        lm_only_scale = 1.0e-20
    if am_only_scale == 0.0:
        am_only_scale = 1.0e-20

    px_interp = (px * combined_scale + px_lmonly * lm_only_scale +
                 px_amonly * am_only_scale)
    py_interp = (py * combined_scale + py_lmonly * lm_only_scale +
                 py_amonly * am_only_scale)

    px_interp = fix_for_boundary(px_interp, boundary)
    return (px_interp, py_interp)


def _delay_penalty(px: Tensor, boundary: Optional[Tensor],
                   delay_penalty: float) -> Tensor:
    B, S, T0 = px.shape
    T = T0 - 1  # regular
    if boundary is None:
        offset = torch.tensor((T - 1) / 2, dtype=px.dtype,
                              device=px.device).expand(B, 1, 1)
    else:
        offset = (boundary[:, 3] - 1) / 2
    penalty = offset.reshape(B, 1, 1) - torch.arange(
        T0, device=px.device).reshape(1, 1, T0)
    penalty = penalty * delay_penalty
    return px + penalty.to(px.dtype)


def _reduce(negated_loss: Tensor, reduction: str) -> Tensor:
    if reduction == "none":
        return -negated_loss
    elif reduction == "mean":
        return -torch.mean(negated_loss)
    elif reduction == "sum":
        return -torch.sum(negated_loss)
    raise ValueError(
        f"reduction should be ('none' | 'mean' | 'sum'), given {reduction}")


def rnnt_loss_smoothed(
    lm: Tensor,
    am: Tensor,
    symbols: Tensor,
    termination_symbol: int,
    lm_only_scale: float = 0.1,
    am_only_scale: float = 0.1,
    boundary: Optional[Tensor] = None,
    rnnt_type: str = "regular",
    delay_penalty: float = 0.0,
    reduction: Optional[str] = "mean",
    return_grad: bool = False,
):
    px, py = get_rnnt_logprobs_smoothed(
        lm=lm,
        am=am,
        symbols=symbols,
        termination_symbol=termination_symbol,
        lm_only_scale=lm_only_scale,
        am_only_scale=am_only_scale,
        boundary=boundary,
        rnnt_type=rnnt_type,
    )
    if delay_penalty > 0.0:
        px = _delay_penalty(px, boundary, delay_penalty)
    scores_and_grads = mutual_information_recursion(px=px,
                                                    py=py,
                                                    boundary=boundary,
                                                    return_grad=return_grad)
    negated_loss = scores_and_grads[0] if return_grad else scores_and_grads
    loss = _reduce(negated_loss, reduction)
    return (loss, scores_and_grads[1]) if return_grad else loss


def _adjust_pruning_lower_bound(s_begin: Tensor, s_range: int) -> Tensor:
    (B, T) = s_begin.shape
    s_begin = monotonic_lower_bound(s_begin)
    # do the magic transformation
    s_begin = -(s_begin -
                (s_range - 1) * torch.arange(0, T, device=s_begin.device))
    # make the transformed tensor non-decreasing
    s_begin = monotonic_lower_bound(s_begin)
    # make start symbol zero
    s_begin = torch.clamp(s_begin, min=0)
    # do the magic transformation again to recover s_begin
    s_begin = -(s_begin -
                (s_range - 1) * torch.arange(0, T, device=s_begin.device))
    return s_begin


#: which published variant of get_rnnt_prune_ranges to use:
#: "B" = cumulative symmetric px+py window: upstream's ``get_rnnt_prune_ranges`` since early 2023, i.e. what
#:       both of the reference's pins resolve to (requirements.txt:4 ``k2==1.24.3.dev20240615``, a 2024 build of
#:       the 1.24.3 sources; Dockerfile.build:28-35 tag v1.24.3, mid 2023) -- the default;
#: "A" = sliding-window sum of py_grad minus padded px_grad (upstream's older function, kept there as
#:       ``get_rnnt_prune_ranges_deprecated``).
#: k2 is not installable offline, so the choice cannot be verified here: both are built,
#: both have golden vectors (tests/golden/<case>.{A,B}.*.npz) and every parity test runs under both.
PRUNE_RANGES_VARIANT = "B"


def get_rnnt_prune_ranges(
    px_grad: Tensor,
    py_grad: Tensor,
    boundary: Tensor,
    s_range: int,
    variant: Optional[str] = None,
) -> Tensor:
    variant = variant or PRUNE_RANGES_VARIANT
    (B, S, T1) = px_grad.shape
    T = py_grad.shape[-1]
    assert T1 in [T, T + 1], T1
    S1 = S + 1
    assert py_grad.shape == (B, S + 1, T), py_grad.shape
    assert boundary.shape == (B, 4), boundary.shape
    assert S >= 1, S
    assert T >= S, (T, S)

    # s_range > S means we won't prune out any symbols.
    if s_range > S:
        s_range = S + 1
    if T1 == T:
        assert s_range >= 1
    else:
        assert s_range >= 2, (
            "Pruning range for standard RNN-T should be equal to or greater "
            "than 2, or no valid paths could survive pruning.")

    if variant == "A":
        # k2 forms the window with as_strided + torch.sum(axis=2); the summation
        # order of that reduction is a torch implementation detail, so the oracle
        # fixes it: left to right over the window (the bit-exactness contract of
        # SURVEY.md A.4).  tests/test_oracle.py checks this choice reproduces the
        # ranges of the reference run (golden vectors) exactly.
        n_cand = S1 - s_range + 1
        blk_sum_grad = py_grad[:, 0:n_cand, :].clone()
        for k in range(1, s_range):
            blk_sum_grad = blk_sum_grad + py_grad[:, k:k + n_cand, :]
        px_pad = torch.zeros((B, 1, T1), dtype=px_grad.dtype,
                             device=px_grad.device)
        px_grad_pad = torch.cat((px_pad, px_grad), dim=1)  # (B, S1, T1)
        final_grad = blk_sum_grad - px_grad_pad[:, :S1 - s_range + 1, :T]
        s_begin = torch.argmax(final_grad, dim=1)  # (B, T)
    elif variant == "B":
        px_pad = torch.zeros((B, 1, T1), dtype=px_grad.dtype,
                             device=px_grad.device)
        py_pad = torch.zeros((B, S1, T1 - T), dtype=py_grad.dtype,
                             device=py_grad.device)
        tot = torch.cat((px_grad, px_pad), dim=1) + torch.cat(
            (py_grad, py_pad), dim=2)  # (B, S1, T1)
        # sequential cumsum over s in the working dtype (torch.cumsum accumulates
        # fp32 in double on CPU and as a tree on GPU; the oracle fixes the order)
        cs_rows = [torch.zeros((B, T1), dtype=tot.dtype, device=tot.device)]
        for j in range(S1):
            cs_rows.append(cs_rows[-1] + tot[:, j, :])
        cs = torch.stack(cs_rows, dim=1)  # (B,S1+1,T1)
        diff = cs[:, s_range:, :] - cs[:, :S1 + 1 - s_range, :]
        s_begin = torch.argmax(diff[:, :, :T], dim=1)  # (B, T)
    else:
        raise ValueError(f"unknown prune-range variant {variant}")

    # Handle the values of s_begin in padding positions: the last real frame
    # and all padding frames are pinned to len(symbols) - s_range + 1.
    mask = torch.arange(0, T, device=px_grad.device).reshape(1, T).expand(B, T)
    mask = mask < boundary[:, 3].reshape(B, 1) - 1
    s_begin_padding = boundary[:, 2].reshape(B, 1) - s_range + 1
    s_begin_padding = torch.clamp(s_begin_padding, min=0)
    s_begin = torch.where(mask, s_begin, s_begin_padding)

    s_begin = _adjust_pruning_lower_bound(s_begin, 2 if T1 == T else s_range)

    ranges = s_begin.reshape((B, T, 1)).expand(
        (B, T, s_range)) + torch.arange(s_range, device=px_grad.device)
    return ranges


def do_rnnt_pruning(am: Tensor, lm: Tensor,
                    ranges: Tensor) -> Tuple[Tensor, Tensor]:
    assert ranges.shape[0] == am.shape[0]
    assert ranges.shape[0] == lm.shape[0]
    assert am.shape[1] == ranges.shape[1]
    (B, T, s_range) = ranges.shape
    (B, S1, C) = lm.shape
    am_pruning = am.unsqueeze(2).expand((B, T, s_range, am.shape[-1]))
    lm_pruning = torch.gather(
        lm.unsqueeze(1).expand((B, T, S1, C)),
        dim=2,
        index=ranges.reshape((B, T, s_range, 1)).expand(
            (B, T, s_range, C)),
    )
    return am_pruning, lm_pruning


def _roll_by_shifts(src: Tensor, shifts: Tensor) -> Tensor:
    """Roll the last axis of ``src`` (B, T, S) right by shifts[b, t]."""
    assert src.dim() == 3
    (B, T, S) = src.shape
    assert shifts.shape == (B, T)
    index = (torch.arange(S, device=src.device).view(
        (1, S)).repeat((T, 1)).repeat((B, 1, 1)))
    index = (index - shifts.reshape(B, T, 1)) % S
    return torch.gather(src, 2, index)


def get_rnnt_logprobs_pruned(
    logits: Tensor,
    symbols: Tensor,
    ranges: Tensor,
    termination_symbol: int,
    boundary: Tensor,
    rnnt_type: str = "regular",
) -> Tuple[Tensor, Tensor]:
    assert logits.ndim == 4, logits.ndim
    (B, T, s_range, C) = logits.shape
    assert ranges.shape == (B, T, s_range), ranges.shape
    (B, S) = symbols.shape
    assert S >= 0, S
    assert rnnt_type == "regular", "oracle restates rnnt_type='regular' only"

    normalizers = torch.logsumexp(logits, dim=3)

    symbols_with_terminal = torch.cat(
        (symbols,
         torch.tensor([termination_symbol] * B, dtype=torch.int64,
                      device=symbols.device).reshape((B, 1))),
        dim=1,
    )
    pruned_symbols = torch.gather(
        symbols_with_terminal.unsqueeze(1).expand((B, T, S + 1)),
        dim=2,
        index=ranges,
    )  # (B, T, s_range)

    px = torch.gather(logits, dim=3,
                      index=pruned_symbols.reshape(B, T, s_range,
                                                   1)).squeeze(-1)
    px = px - normalizers
    px = torch.cat(
        (px,
         torch.full((B, T, S + 1 - s_range), float("-inf"), device=px.device,
                    dtype=px.dtype)),
        dim=2,
    )  # (B, T, S+1)
    px = _roll_by_shifts(px, ranges[:, :, 0])[:, :, :S]
    px = px.permute((0, 2, 1))
    px = torch.cat(
        (px,
         torch.full((B, S, 1), float("-inf"), device=px.device,
                    dtype=px.dtype)),
        dim=2,
    )  # (B, S, T+1)

    py = logits[:, :, :, termination_symbol].clone()  # (B, T, s_range)
    py = py - normalizers
    py = torch.cat(
        (py,
         torch.full((B, T, S + 1 - s_range), float("-inf"), device=py.device,
                    dtype=py.dtype)),
        dim=2,
    )
    py = _roll_by_shifts(py, ranges[:, :, 0])
    py = py.permute((0, 2, 1))  # (B, S+1, T)

    px = fix_for_boundary(px, boundary)
    return (px, py)


def rnnt_loss_pruned(
    logits: Tensor,
    symbols: Tensor,
    ranges: Tensor,
    termination_symbol: int,
    boundary: Tensor = None,
    rnnt_type: str = "regular",
    delay_penalty: float = 0.0,
    reduction: Optional[str] = "mean",
) -> Tensor:
    px, py = get_rnnt_logprobs_pruned(
        logits=logits,
        symbols=symbols,
        ranges=ranges,
        termination_symbol=termination_symbol,
        boundary=boundary,
        rnnt_type=rnnt_type,
    )
    if delay_penalty > 0.0:
        px = _delay_penalty(px, boundary, delay_penalty)
    scores = mutual_information_recursion(px=px, py=py, boundary=boundary)
    return _reduce(scores, reduction)


def get_rnnt_logprobs(
    lm: Tensor,
    am: Tensor,
    symbols: Tensor,
    termination_symbol: int,
    boundary: Optional[Tensor] = None,
    rnnt_type: str = "regular",
) -> Tuple[Tensor, Tensor]:
    """Unsmoothed "simple" log-probs (k2.get_rnnt_logprobs); equals the
    smoothed version with both scales exactly 0 and no 1e-20 substitution."""
    (B, T, C) = am.shape
    S = lm.shape[1] - 1
    am_max, _ = torch.max(am, dim=2, keepdim=True)
    lm_max, _ = torch.max(lm, dim=2, keepdim=True)
    am_probs = (am - am_max).exp()
    lm_probs = (lm - lm_max).exp()
    normalizers = (torch.matmul(lm_probs, am_probs.transpose(1, 2)) +
                   torch.finfo(am_probs.dtype).tiny).log()
    normalizers = normalizers + lm_max + am_max.transpose(1, 2)
    px_am = torch.gather(
        am.unsqueeze(1).expand(B, S, T, C), dim=3,
        index=symbols.reshape(B, S, 1, 1).expand(B, S, T, 1)).squeeze(-1)
    px_am = torch.cat((px_am, torch.full((B, S, 1), float("-inf"),
                                         dtype=px_am.dtype)), dim=2)
    px_lm = torch.gather(lm[:, :S], dim=2, index=symbols.unsqueeze(-1))
    px = px_am + px_lm
    px[:, :, :T] -= normalizers[:, :S, :]
    py_am = am[:, :, termination_symbol].unsqueeze(1)
    py_lm = lm[:, :, termination_symbol].unsqueeze(2)
    py = py_am + py_lm - normalizers
    px = fix_for_boundary(px, boundary)
    return (px, py)
