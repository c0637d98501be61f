"""Loss factory with the reference's interface (/root/reference/model/loss/loss.py:19-53):
``Loss(config)(batch_dict)`` -> ``self.loss(**batch)``.  The transducer losses
are the sm_100a implementations; losses outside the hot path (CTC, masked CE /
KL / MAE) are resolved lazily from the reference package when it is importable.
"""
from __future__ import annotations

import importlib
from typing import Dict

import torch
import torch.nn as nn

from .pruned_rnnt_loss import PrunedRnntLoss, PrunedRnntLossConfig
from .rnnt_loss import RnntLoss, RnntLossConfig

# name -> (module, class, config class) of losses that are NOT on the hot path
_OFF_PATH = {
    "CTC": ("ctc_loss", "CtcLoss", "CtcLossConfig"),
    "MaskedCELoss": ("cross_entropy", "MaskedCELoss", "MaskedCELossConfig"),
    "MaskedKLDiv": ("kl_divergence", "MaskedKLDivergence", "MaskedKLDivergenceConfig"),
    "MaeLoss": ("mae_loss", "MaeLoss", "MaeLossConfig"),
}


class Loss(nn.Module):
    """ Loss interface for all designed losses """

    def __init__(self, config) -> None:
        super(Loss, self).__init__()
        name = config["model"]
        if name == "Rnnt":
            self.loss = RnntLoss(config=RnntLossConfig(**config["config"]))
        elif name == "Pruned_Rnnt":
            self.loss = PrunedRnntLoss(config=PrunedRnntLossConfig(**config["config"]))
        elif name in _OFF_PATH:
            mod_name, cls_name, cfg_name = _OFF_PATH[name]
            mod = None
            for pkg in ("speech2text_b200.loss", "model.loss"):
                try:
                    mod = importlib.import_module(f"{pkg}.{mod_name}")
                    break
                except ImportError:
                    continue
            if mod is None:
                raise ValueError(f"Not support {name} loss (module {mod_name} is outside the hot path "
                                 "and the reference package is not importable)")
            self.loss = getattr(mod, cls_name)(config=getattr(mod, cfg_name)(**config["config"]))
        else:
            raise ValueError("Not support {} loss".format(name))

    def forward(self, batch: Dict[str, torch.Tensor]):
        """ Loss training graph: kwarg names are the contract (rnnt_task.py:227-232, 474-481). """
        return self.loss(**batch)

    def predict(self, logits: torch.Tensor):
        """ Predict step for metric compute """
        if hasattr(self.loss, "predict"):
            return self.loss.predict(logits)
