"""ORACLE tooling (test infrastructure, NOT product code): the seeded parity
cases shared by oracle/make_golden.py (which runs the reference on them) and
tests/ (which run the CUDA path and the oracle port on them).

Everything is drawn from numpy ``RandomState`` streams so the same arrays are
regenerated bit-exactly on any box, independent of the torch version.

Shapes follow the reference's own hot-path tests and BASELINE.json configs:
  * joiner_test        /root/reference/model/joiner/joiner_test.py:25-32, 53-71
  * pruned_loss_test   /root/reference/model/loss/pruned_rnnt_loss_test.py:21-46
  * c1_zipformer       config/training/zipformer_stateless_pruned_rnnt.yaml:82-95
                       with sample_data-shaped lengths
  * rnnt_loss_test     /root/reference/model/loss/rnnt_loss_test.py:19-29 shapes
                       via the unpruned joiner (joiner_test.py:17-23, 34-51)
"""
from __future__ import annotations

import zlib

import numpy as np

CASES = {
    # reference joiner_test.py pruned fixture: no out-projection, ReLU
    "joiner_test": dict(
        joiner=dict(input_dim=256, output_dim=128, activation="relu",
                    prune_range=5, use_out_project=False),
        B=4, T=200, U=15, t_lens=[197, 200, 65, 80], u_lens=[8, 10, 9, 2],
        loss=dict(termination_symbol=0, reduction="mean"),
        simple_loss_scale=0.5, pruned_loss_scale=0.5, scale=1.0, uniform=True),
    # reference pruned_rnnt_loss_test.py fixture: 512 -> 1000, out-proj 256
    "pruned_loss_test": dict(
        joiner=dict(input_dim=512, output_dim=1000, activation="relu",
                    prune_range=5),
        B=4, T=200, U=15, t_lens=[197, 200, 65, 80], u_lens=[8, 10, 9, 2],
        loss=dict(termination_symbol=0, reduction="mean"),
        simple_loss_scale=0.5, pruned_loss_scale=0.5, scale=1.0, uniform=True),
    # BASELINE config 1 shape (zipformer yaml joiner block, sample_data lengths)
    "c1_zipformer": dict(
        joiner=dict(input_dim=256, output_dim=128, prune_range=5,
                    use_out_project=False),
        B=4, T=327, U=123, t_lens=[327, 285, 245, 178], u_lens=[123, 83, 97, 59],
        loss=dict(termination_symbol=0, reduction="mean"),
        simple_loss_scale=0.5, pruned_loss_scale=0.5, scale=0.5, uniform=False),
    # tanh + out-projection + non-zero smoothing scales, sum reduction
    "tanh_smoothed": dict(
        joiner=dict(input_dim=96, output_dim=72, inner_dim=40, activation="tanh",
                    prune_range=4, lm_scale=0.25, am_scale=0.1),
        B=5, T=61, U=17, t_lens=[61, 50, 33, 61, 20], u_lens=[17, 12, 1, 9, 17],
        loss=dict(termination_symbol=0, reduction="sum"),
        simple_loss_scale=0.3, pruned_loss_scale=0.7, scale=0.7, uniform=False),
    # edge cases: S_b < prune_range, T_b == S_b, single-symbol targets, 'none'
    "edge_short": dict(
        joiner=dict(input_dim=48, output_dim=33, inner_dim=16, activation="relu",
                    prune_range=5),
        B=6, T=24, U=7, t_lens=[24, 7, 3, 12, 24, 5], u_lens=[7, 7, 2, 1, 3, 4],
        loss=dict(termination_symbol=0, reduction="none"),
        simple_loss_scale=0.5, pruned_loss_scale=0.5, scale=0.8, uniform=False),
    # prune_range larger than S: k2 clamps s_range to S+1 (full lattice)
    "range_clamped": dict(
        joiner=dict(input_dim=40, output_dim=29, inner_dim=24, activation="tanh",
                    prune_range=9),
        B=3, T=31, U=6, t_lens=[31, 20, 8], u_lens=[6, 4, 6],
        loss=dict(termination_symbol=0, reduction="mean", delay_penalty=0.05),
        simple_loss_scale=0.5, pruned_loss_scale=0.5, scale=0.8, uniform=False),
    # vanilla path: unpruned joiner + torchaudio RNNTLoss
    "vanilla_rnnt": dict(
        joiner=dict(input_dim=64, output_dim=40, inner_dim=32, activation="relu",
                    prune_range=-1),
        B=3, T=50, U=12, t_lens=[50, 37, 22], u_lens=[12, 7, 9],
        loss=dict(blank_label=0, clamp=-1, reduction="mean"),
        scale=0.8, uniform=False),
}


def _seed(name: str) -> int:
    return zlib.crc32(name.encode()) & 0x7FFFFFFF


def make_weights(joiner_cfg: dict, rs: np.random.RandomState) -> dict:
    """nn.Linear-style U(-1/sqrt(fan_in), 1/sqrt(fan_in)) init, state_dict keys
    as in /root/reference/model/joiner/joiner.py:41-55."""
    D, V = joiner_cfg["input_dim"], joiner_cfg["output_dim"]
    inner = joiner_cfg.get("inner_dim", 256)

    def lin(prefix, fan_out, fan_in):
        k = 1.0 / np.sqrt(fan_in)
        return {
            f"{prefix}.weight": rs.uniform(-k, k, (fan_out, fan_in)).astype(np.float32),
            f"{prefix}.bias": rs.uniform(-k, k, (fan_out,)).astype(np.float32),
        }

    w = {}
    w.update(lin("_enc_proj", V, D))
    w.update(lin("_pre_proj", V, D))
    if joiner_cfg.get("use_out_project", True):
        w.update(lin("_out_projection.0", inner, V))
        w.update(lin("_out_projection.1", V, inner))
    return w


def make_case(name: str) -> dict:
    spec = CASES[name]
    rs = np.random.RandomState(_seed(name))
    B, T, U = spec["B"], spec["T"], spec["U"]
    D, V = spec["joiner"]["input_dim"], spec["joiner"]["output_dim"]
    if spec["uniform"]:  # the reference tests use torch.rand
        enc = rs.uniform(0, 1, (B, T, D))
        pred = rs.uniform(0, 1, (B, U + 1, D))
    else:
        enc = rs.standard_normal((B, T, D))
        pred = rs.standard_normal((B, U + 1, D))
    enc = (enc * spec["scale"]).astype(np.float32)
    pred = (pred * spec["scale"]).astype(np.float32)
    u_lens = np.asarray(spec["u_lens"], dtype=np.int64)
    t_lens = np.asarray(spec["t_lens"], dtype=np.int64)
    target = rs.randint(1, V, (B, U)).astype(np.int64)
    if name != "joiner_test" and name != "pruned_loss_test":
        # dataset/utils.py:189-191 pads labels with 0 past the true length; the
        # two reference tests above use un-padded randint targets.
        for b in range(B):
            target[b, u_lens[b]:] = 0
    return dict(
        encoder_out=enc,
        predict_out=pred,
        encoder_out_lengths=t_lens,
        target_lengths=u_lens,
        target=target,
        weights=make_weights(spec["joiner"], rs),
    )
