"""CTC loss wrapper with the reference's interface (/root/reference/model/loss/ctc_loss.py:13-41).
Not on the hot path this round: torch's native CTC, kept so that
``PrunedRnntTask(enable_ctc=True)`` / ``CtcHybridRnnt`` can build their losses."""
from __future__ import annotations

import dataclasses

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclasses.dataclass
class CtcLossConfig:
    blank_label: int = 0
    reduction: str = "mean"
    zero_infinity: bool = True


class CtcLoss(nn.Module):

    def __init__(self, config: CtcLossConfig):
        super().__init__()
        self._loss = nn.CTCLoss(blank=config.blank_label, reduction=config.reduction,
                                zero_infinity=config.zero_infinity)

    def forward(self, logits, targets, logits_length, targets_length):
        # (B, T, N) logits -> (T, B, N) fp32 log-probs
        log_probs = F.log_softmax(logits, dim=-1).transpose(0, 1).to(dtype=torch.float32)
        return self._loss(log_probs, targets, logits_length, targets_length)
