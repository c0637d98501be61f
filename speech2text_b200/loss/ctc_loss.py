"""B200 drop-in for ``model/loss/ctc_loss.py`` (reference: ``F.log_softmax`` + ``nn.CTCLoss(zero_infinity=True)``,
/root/reference/model/loss/ctc_loss.py:13-41) -- SURVEY.md section 8 row f-2, the CTC half of BASELINE config 4
(``PrunedRnntTask`` with ``enable_ctc``, rnnt_task.py:485-496, and ``CtcHybridRnnt``, :341-349).

Same config dataclass and ``forward`` kwargs; on a CUDA device the log-softmax, the alpha/beta lattice and the
gradient w.r.t. the logits run as fused kernels (``functional.ctc_loss`` -> ``s2t_ctc_loss_fwd/bwd``): the
(T, B, V) log-probabilities the reference materialises are never written.  No CPU path (S2TError)."""
from __future__ import annotations

import dataclasses

import torch
import torch.nn as nn

from .. import _lib
from .. import functional as F2


@dataclasses.dataclass
class CtcLossConfig:
    """ Config of CTCLoss (ctc_loss.py:13-18) """
    blank_label: int = 0
    reduction: str = "mean"
    zero_infinity: bool = True


class CtcLoss(nn.Module):
    """ Ctc Loss on sm_100a """

    def __init__(self, config: CtcLossConfig):
        super(CtcLoss, self).__init__()
        self._blank_label = config.blank_label
        self._reduction = config.reduction
        self._zero_infinity = config.zero_infinity
        if self._reduction not in ("mean", "sum", "none"):
            raise ValueError(f"reduction should be ('none' | 'mean' | 'sum'), given {self._reduction}")

    def forward(self, logits, targets, logits_length, targets_length):
        """ logits (B, T, N) raw scores (the log-softmax is applied inside, as the reference does before
            nn.CTCLoss); targets (B, S) padded; logits_length, targets_length (B). """
        if not logits.is_cuda:
            raise _lib.S2TError("speech2text_b200.CtcLoss needs CUDA tensors: this build has no CPU path")
        return F2.ctc_loss(logits, targets, logits_length, targets_length, blank=self._blank_label,
                           reduction=self._reduction, zero_infinity=self._zero_infinity)
