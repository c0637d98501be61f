"""B200 drop-in for ``model/loss/rnnt_loss.py`` (reference wraps
``torchaudio.transforms.RNNTLoss``, /root/reference/model/loss/rnnt_loss.py:27-45):
full-lattice RNN-T loss with the log-softmax fused in.
"""
from __future__ import annotations

import dataclasses

import torch
import torch.nn as nn

from .. import functional as F2
from ..joiner import LazyJoinerLogits


@dataclasses.dataclass
class RnntLossConfig:
    """ Config of RnntLoss (rnnt_loss.py:13-18) """
    blank_label: int = 0  # Blank label index
    clamp: float = -1  # Clamp for gradients
    reduction: str = "mean"  # Specifies the reduction to apply to the output


class RnntLoss(nn.Module):
    """ Full-lattice Rnnt Loss on sm_100a """

    def __init__(self, config: RnntLossConfig) -> None:
        super(RnntLoss, self).__init__()
        self._blank = config.blank_label
        self._clamp = config.clamp
        self._reduction = config.reduction
        if self._reduction not in ("mean", "sum", "none"):
            raise ValueError(f"reduction should be ('none' | 'mean' | 'sum'), given {self._reduction}")

    def forward(self, logits, targets: torch.Tensor, logits_length: torch.Tensor,
                targets_length: torch.Tensor):
        """ logits: (B, max_T, 1 + max_U, V) tensor or LazyJoinerLogits; targets: (B, max_U);
            logits_length, targets_length: (B).  As with torchaudio, max(logits_length) must equal
            max_T and max(targets_length) max_U (not re-checked here: that would cost a host sync). """
        dev = logits.device
        boundary = F2.make_boundary(targets_length, logits_length, dev)
        targets = targets.to(dev)
        if isinstance(logits, LazyJoinerLogits):
            assert logits.ranges is None, "RnntLoss expects the unpruned joiner (prune_range=-1)"
            scores = F2.joiner_scores(logits.am, logits.lm, logits.W1, logits.b1, logits.W2, logits.b2,
                                      targets, None, boundary, logits.act, blank=self._blank,
                                      clamp=self._clamp, mode=logits.mode)
        else:
            assert logits.dim() == 4 and logits.shape[2] == targets.shape[1] + 1, (
                "logits must be (B, T, 1 + U, V) with targets (B, U)")
            scores = F2.logits_scores(logits, targets, None, boundary, blank=self._blank, clamp=self._clamp)
        return F2._reduce(scores, self._reduction)
