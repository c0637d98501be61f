"""B200 drop-in for ``model/loss/pruned_rnnt_loss.py`` (reference wraps
``k2.rnnt_loss_pruned``, /root/reference/model/loss/pruned_rnnt_loss.py:34-50).

``forward`` keeps the kwarg names the task passes (rnnt_task.py:474-481).  When
``logits`` is the ``LazyJoinerLogits`` handle returned by the fused Joiner, the
joiner contraction, log-sum-exp, gather and banded lattice DP run as fused
kernels; a real (B,T,R,V) tensor takes the streaming materialised path.
"""
from __future__ import annotations

import dataclasses

import torch
import torch.nn as nn

from .. import functional as F2
from ..joiner import LazyJoinerLogits


@dataclasses.dataclass
class PrunedRnntLossConfig:
    """ Pruned Rnnt Loss Config (pruned_rnnt_loss.py:14-20) """
    termination_symbol: int = 0  # <blank id> = 0
    rnnt_type: str = "regular"
    delay_penalty: float = 0.0
    reduction: str = "mean"


class PrunedRnntLoss(nn.Module):
    """ Pruned Rnnt Loss on sm_100a """

    def __init__(self, config: PrunedRnntLossConfig) -> None:
        super(PrunedRnntLoss, self).__init__()
        self._termination_symbol = config.termination_symbol
        self._rnnt_type = config.rnnt_type
        self._delay_penalty = config.delay_penalty
        self._reduction = config.reduction
        if self._rnnt_type != "regular":
            raise NotImplementedError(
                f"rnnt_type={self._rnnt_type!r}: only the 'regular' transducer topology is built "
                "(the reference's joiner always uses it for the simple loss, joiner.py:100-110)")
        if self._reduction not in ("mean", "sum", "none"):
            raise ValueError(f"reduction should be ('none' | 'mean' | 'sum'), given {self._reduction}")

    def forward(self, logits, targets: torch.Tensor, logits_length: torch.Tensor,
                targets_length: torch.Tensor, boundary: torch.Tensor, ranges: torch.Tensor):
        # logits_length / targets_length are unused, as in the reference: lengths come from boundary.
        if isinstance(logits, LazyJoinerLogits):
            targets = targets.to(logits.device)
            scores = F2.joiner_scores(logits.am, logits.lm, logits.W1, logits.b1, logits.W2, logits.b2,
                                      targets, ranges, boundary, logits.act,
                                      blank=self._termination_symbol, delay_penalty=self._delay_penalty,
                                      mode=logits.mode)
        else:
            # NOTE: the reference casts logits to fp32 first; the kernel reads fp32/bf16/fp16
            # logits directly and accumulates in fp32, which is the same arithmetic.
            scores = F2.logits_scores(logits, targets.to(logits.device), ranges, boundary,
                                      blank=self._termination_symbol, delay_penalty=self._delay_penalty)
        return F2._reduce(scores, self._reduction)
