#!/usr/bin/env python
"""bench.py -- pruned RNN-T loss forward+backward throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--mode fp32|bf16]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores

One "step" = one pass of the hot path over one synthetic batch:
(encoder_out, predict_out, lengths, labels) -> joiner projections -> simple loss +
occupation probs -> prune ranges -> fused pruned joiner + loss -> backward to
d_encoder_out, d_predict_out and all joiner weight gradients, loss = 0.5 simple + 0.5 pruned
(/root/reference/task_factory/rnnt_task.py:469-514).  Under torchrun each rank owns its own
utterances (weak scaling) and the step ends with ONE all-reduce of the flat joiner weight
gradient and one of the scalar losses.

Prints ONE JSON line (rank 0).  `value` = utterances/s with inputs resident in HBM; `e2e` =
the same metric through the public module API with inputs in pinned HOST memory (H2D of the
step's inputs and D2H of the losses inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "pruned RNN-T loss fwd+bwd utterances/sec"

# BASELINE.json configs: the metric is quoted at 1/2/4/8 B200 on config 3.
WORKLOADS = {
    "c1": dict(B=4, T=327, U=123, V=128, D=256, R=5, I=0, act="relu",
               desc="zipformer_stateless_pruned_rnnt.yaml joiner, sample_data-shaped lengths"),
    "c2": dict(B=32, T=250, U=50, V=500, D=512, R=-1, I=256, act="tanh",
               desc="vanilla Rnnt full-lattice loss, synthetic B=32 T=250 U=50 V=500 joiner D=512"),
    # "bf16 joiner" (BASELINE config 3, SURVEY 8(d) e = 2): the activations arrive as bf16 in tensor-core mode
    "c3": dict(B=64, T=400, U=100, V=500, D=512, R=5, I=256, act="tanh", in_dtype="bf16",
               desc="pruned RNN-T prune_range=5, synthetic B=64 T=400 U=100 V=500 D=512, bf16 joiner"),
    "c4": dict(B=128, T=500, U=125, V=2000, D=512, R=5, I=256, act="tanh", ctc=True,
               desc="CTC + pruned RNN-T loss on shared encoder output (PrunedRnntTask enable_ctc, rnnt_task.py:485-496): "
                    "B=128 T=500 U=125 V=2000 D=512 prune_range=5"),
    "c5": dict(B=256, T=1000, U=250, V=5000, D=1024, R=5, I=256, act="tanh",
               desc="large-vocab pruned RNN-T V=5000 D=1024 prune_range=5, B=256/GPU"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm=p["hbm_gbs"], tc_burst=p["bf16_tflops"], tc_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------
def make_batch(cfg, seed: int, in_dtype: str = "f32"):
    """Host tensors.  Longest item full length, others T_b ~ U[0.6T, T], S_b ~ T_b*U/T*U[0.8,1].
    in_dtype "bf16": encoder_out / predict_out are rounded to bf16 (both arms start from the same rounded values)."""
    g = torch.Generator().manual_seed(seed)
    B, T, U, V, D = cfg["B"], cfg["T"], cfg["U"], cfg["V"], cfg["D"]
    enc = torch.randn(B, T, D, generator=g) * 0.5
    pred = torch.randn(B, U + 1, D, generator=g) * 0.5
    t_len = torch.randint(int(0.6 * T), T + 1, (B,), generator=g)
    frac = 0.8 + 0.2 * torch.rand(B, generator=g)
    s_len = torch.clamp(torch.round(t_len.float() * U / T * frac).long(), 1, U)
    t_len[0], s_len[0] = T, U
    s_len = torch.minimum(s_len, t_len)
    labels = torch.randint(1, V - 1, (B, U), generator=g)
    for b in range(B):
        labels[b, s_len[b]:] = 0  # pad_sequence(padding_value=0), dataset/utils.py:189-191
    if in_dtype == "bf16":
        enc, pred = enc.bfloat16(), pred.bfloat16()
    return dict(enc=enc, pred=pred, t_len=t_len, s_len=s_len, labels=labels)


def input_dtype(cfg, mode: str) -> str:
    return cfg.get("in_dtype", "f32") if mode == "bf16" else "f32"


def build_modules(cfg, device, mode: str):
    from speech2text_b200 import Joiner, JoinerConfig, Loss
    os.environ["S2T_B200_FUSED"] = "1"
    os.environ["S2T_B200_JOINER_MODE"] = mode
    torch.manual_seed(1234)  # build_task.py:49
    joiner = Joiner(JoinerConfig(input_dim=cfg["D"], output_dim=cfg["V"], inner_dim=max(cfg["I"], 1),
                                 activation=cfg["act"], prune_range=cfg["R"],
                                 use_out_project=cfg["I"] > 0)).to(device)
    if cfg["R"] > 0:
        loss = Loss({"model": "Pruned_Rnnt", "config": {"termination_symbol": 0, "reduction": "mean"}})
    else:
        loss = Loss({"model": "Rnnt", "config": {"blank_label": 0, "clamp": -1, "reduction": "mean"}})
    if cfg.get("ctc"):
        # PrunedRnntTask with loss.enable_ctc (rnnt_task.py:434-445): a Projector head (model/decoder/projector.py:
        # Linear D -> V; its dropout is switched off here so that both arms see the same numbers) and the CTC loss
        joiner.ctc_proj = torch.nn.Linear(cfg["D"], cfg["V"]).to(device)
        joiner.ctc_loss = Loss({"model": "CTC", "config": {"blank_label": 0, "reduction": "mean", "zero_infinity": True}})
    return joiner, loss


def hot_path_step(joiner, loss_mod, enc, t_len, pred, s_len, labels):
    """rnnt_task.py:469-499 through the reference-facing module API."""
    if joiner.prune_range <= 0:  # vanilla full-lattice loss (rnnt_task.py:326-339)
        logits, _, _, _ = joiner(enc, t_len, pred, s_len)
        total = loss_mod({"logits": logits, "logits_length": t_len, "targets": labels, "targets_length": s_len})
        total.backward()
        return total, total.detach(), total.detach()
    logits, boundary, ranges, simple = joiner(enc, t_len, pred, s_len, labels)
    pruned = loss_mod({"logits": logits, "logits_length": t_len, "targets": labels, "targets_length": s_len,
                       "boundary": boundary, "ranges": ranges})
    total = 0.5 * simple + 0.5 * pruned
    if hasattr(joiner, "ctc_proj"):  # rnnt_task.py:485-496: loss = simple_scale * simple + pruned_scale * pruned + ctc
        total = total + ctc_branch(joiner, enc, t_len, labels, s_len)
    total.backward()
    return total, simple, pruned


def ctc_branch(joiner, enc, t_len, labels, s_len):
    from speech2text_b200 import functional as F2
    proj = joiner.ctc_proj
    if os.environ.get("S2T_B200_JOINER_MODE", "fp32") == "bf16":
        ctc_logits = F2.linear_tc(enc, proj.weight, proj.bias)  # the Projector's Linear on tcgen05 as well
    else:
        ctc_logits = proj(enc)
    return joiner.ctc_loss({"logits": ctc_logits, "logits_length": t_len, "targets": labels, "targets_length": s_len})


# ---------------------------------------------------------------------------------------------
# algorithmic work of each kernel (per launch group), for the roofline of the dominant one
# ---------------------------------------------------------------------------------------------
def kernel_work(cfg, live_frames: float = 1.0):
    """live_frames = real frames / padded frames of the batch: the joiner kernels skip the row tiles of padding
    frames, so their work is counted on the real frames only (the padded count would flatter them)."""
    B, T, U, V, D, R, I = (cfg[k] for k in ("B", "T", "U", "V", "D", "R", "I"))
    M = B * T * (R if R > 0 else U + 1) * live_frames  # joiner rows: pruned band, or the vanilla full lattice
    S1 = U + 1
    gemm = 2.0 * M * V * max(I, 1)
    simple = 2.0 * B * S1 * T * V
    lat_cells = B * (S1 * (T + 1))
    band_cells = B * T * R
    proj_enc = 2.0 * B * T * D * V
    proj_pred = 2.0 * B * S1 * D * V
    rows = float(B * T + B * S1)  # projections: encoder-side and predictor-side launch together
    Ii = max(I, 1)
    # name: (algorithmic FLOPs, algorithmic HBM bytes) per STEP, summed over the kernel's launches in it (two for the
    # projections, one per row chunk when the joiner backward is chunked).  Bytes = every operand the kernel must read once
    # plus every result it must write once, in the dtype it is stored in (packed bf16 operands 2 B, fp32 4 B); weights
    # and per-row scalars are left out.  The roofline that bounds a kernel is whichever of FLOPs / tensor peak and
    # bytes / HBM peak takes longer: every contraction of this path is short and wide (K <= 512 against 10^5 rows),
    # which puts most of them on the HBM side even at the bf16 tensor rate.
    w = {
        # strict-fp32 SIMT mode
        "joiner_hidden_gemm": (gemm, 4.0 * M * (V + Ii)), "joiner_logits_gemm": (gemm, 4.0 * M * (V + Ii)),
        "joiner_dhidden_gemm": (gemm, 4.0 * M * (V + Ii)), "joiner_dW2_gemm": (gemm, 4.0 * M * (V + Ii)),
        "joiner_dW1_gemm": (gemm, 4.0 * M * (V + Ii)), "joiner_djoint_gemm": (gemm, 4.0 * M * (V + Ii)),
        "simple_normaliser_gemm": (simple, 4.0 * (B * T * V + B * S1 * V) + 8.0 * B * S1 * T),
        "simple_d_am_gemm": (simple, 4.0 * (B * S1 * T + B * S1 * V + B * T * V)),
        "simple_d_lm_gemm": (simple, 4.0 * (B * S1 * T + B * S1 * V + B * T * V)),
        # bf16 tensor-core mode (tcgen05): every joiner contraction is 2*M*V*I
        # fused forward (both contractions in one kernel, the hidden tile never leaves the SM on its way to GEMM2):
        # am and lm rows read once (fp32), hidden rows written once in bf16 for the backward, lse / px / py out
        "tc_joiner_fwd_fused": (2.0 * gemm, 4.0 * (B * T * V * live_frames + B * S1 * V) + 2.0 * M * Ii + 12.0 * M),
        "tc_joiner_hidden_gemm": (gemm, 2.0 * M * V + 2.0 * M * Ii),            # act(am+lm) rows in, hidden rows out
        "tc_joiner_logits_lse_gemm": (gemm, 2.0 * M * Ii + 12.0 * M),            # hidden in, lse / px / py out
        "tc_joiner_grad_logits_gemm": (gemm, 2.0 * M * Ii + 2.0 * M * V),        # hidden in, d logits (bf16) out
        "tc_joiner_dhidden_gemm": (gemm, 2.0 * M * V + 2.0 * M * Ii),            # d logits in, d hidden out
        "tc_joiner_dW2_gemm": (gemm, 2.0 * M * V + 2.0 * M * Ii),                # d logits and hidden in
        "tc_joiner_dW1_gemm": (gemm, 2.0 * M * V + 2.0 * M * Ii),                # d hidden and act rows in
        "tc_joiner_djoint_gemm": (gemm, 2.0 * M * V + 2.0 * M * Ii),
        "tc_joiner_dh_gemm": (gemm, 2.0 * M * Ii + 2.0 * M * V),                 # d hidden in, d act rows (bf16) out
        "joint_pack_kernel": (0.0, 4.0 * (B * T * V + B * S1 * V) + 2.0 * M * V),
        # one pass over the d act rows (bf16): those, am and lm read once, d_am written once, d_lm read-modify-write
        "djoint_reduce_kernel": (0.0, M * V * 2.0 + 2 * B * T * V * 4.0 + 3 * B * S1 * V * 4.0),
        "simple_px_kernel": (0.0, 4.0 * (2 * B * U * (T + 1) + B * U * T)),
        # simple (smoothed) loss on tensor cores: exp(am - max) built on the fly, contraction with exp(lm - max)
        # over V, nrm/py emitted from the epilogue: am and lm read once, nrm/py written once
        "tc_simple_normaliser_gemm_3xtf32": (simple, 4.0 * (B * T * V + B * S1 * V) + 8.0 * B * S1 * T),
        "tc_simple_normaliser_gemm_3xf16": (simple, 4.0 * (B * T * V + B * S1 * V) + 8.0 * B * S1 * T),
        "tc_simple_d_am_gemm": (simple, 2.0 * (B * S1 * T + B * S1 * V) + 8.0 * B * T * V),
        "tc_simple_d_lm_gemm": (simple, 2.0 * (B * S1 * T + B * T * V) + 8.0 * B * S1 * V),
        "simple_w_kernel": (0.0, 12.0 * B * S1 * T), "row_max_kernel": (0.0, 4.0 * (B * T * V + B * S1 * V)),
        "tc_linear_fwd_gemm_3xtf32": (2.0 * rows * D * V, 4.0 * rows * (D + V)),
        "tc_linear_fwd_gemm_3xf16": (2.0 * rows * D * V, 4.0 * rows * (D + V)),
        "tc_linear_dx_gemm": (2.0 * rows * D * V, rows * (2.0 * V + 4.0 * D)),
        "tc_linear_dW_gemm": (2.0 * rows * D * V, 2.0 * rows * (V + D)),
        "lse_gather_kernel": (0.0, M * V * 4.0), "logits_grad_kernel": (0.0, 2.0 * M * V * 4),
        "joint_act_kernel": (0.0, M * V * 4.0 * 3), "joint_grad_kernel": (0.0, M * V * 4.0 * 5),
        # lattices: alpha reads px,py and writes alpha; beta the same; occupation reads 4, writes 2
        "lattice_alpha_kernel": (0.0, 12.0 * (lat_cells + band_cells) / 2),
        "lattice_beta_kernel": (0.0, 20.0 * (lat_cells + band_cells) / 2),
        "simple_lattice_kernel": (0.0, 2 * 12.0 * lat_cells),
        "simple_occupation_kernel": (0.0, 24.0 * lat_cells),
        "band_lattice_kernel": (0.0, 20.0 * band_cells),
        "prune_ranges_kernel": (0.0, B * ((U * (T + 1) + S1 * T) * 4.0 + T * R * 8.0)),
        # CTC branch (config 4): logits read once for lse + emissions, once more with the gradient write; the lattice
        # reads the emissions and writes alpha / beta~ (2U+1 states each)
        "ctc_emit_kernel": (0.0, 4.0 * B * T * (V + S1)), "ctc_grad_kernel": (0.0, 4.0 * B * T * (2 * V + 2 * (2 * U + 1))),
        "ctc_lattice_kernel": (0.0, 4.0 * B * T * (2 * S1 + 2 * (2 * U + 1))),
    }
    return w


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # pragma: no cover
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # pragma: no cover
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU path (oracle port over the k2 restatement + torch CPU ops)
# ---------------------------------------------------------------------------------------------
def port_once(cfg, batch, n, joiner_state, ranges_override=None):
    """One forward + backward of the reference's CPU path (oracle port) on the first n utterances."""
    from oracle import reference_port as port
    jc = dict(input_dim=cfg["D"], output_dim=cfg["V"], inner_dim=max(cfg["I"], 1), activation=cfg["act"],
              prune_range=cfg["R"], use_out_project=cfg["I"] > 0)
    w = {k: v.clone().float().requires_grad_(True) for k, v in joiner_state.items()}
    e = batch["enc"][:n].float().clone().requires_grad_(True)
    p = batch["pred"][:n].float().clone().requires_grad_(True)
    t_len, s_len, labels = batch["t_len"][:n], batch["s_len"][:n], batch["labels"][:n]
    if cfg["R"] <= 0:
        logits, _, _, _ = port.joiner_forward(w, jc, e, t_len, p, s_len, None)
        total = port.rnnt_loss(logits, labels, t_len, s_len)
        total.backward()
        return dict(total=total.detach(), d_enc=e.grad, d_pred=p.grad)
    logits, boundary, ranges, simple = port.joiner_forward(w, jc, e, t_len, p, s_len, labels,
                                                           ranges_override=ranges_override)
    pruned = port.pruned_rnnt_loss(logits, labels, boundary, ranges)
    total = 0.5 * simple + 0.5 * pruned
    if "ctc_proj.weight" in w:
        ctc_logits = torch.nn.functional.linear(e, w["ctc_proj.weight"], w["ctc_proj.bias"])
        total = total + port.ctc_loss(ctc_logits, labels, t_len, s_len)
    total.backward()
    return dict(total=total.detach(), simple=simple.detach(), pruned=pruned.detach(), ranges=ranges, d_enc=e.grad,
                d_pred=p.grad)


def cpu_arm(cfg, batch, n_utts: int, steps: int, warmup: int, joiner_state=None, threads: int = 0):
    """Times the reference's CPU path (oracle port) on the first n utterances of the batch.  Returns (baseline dict,
    seconds per step, results of the last step) -- the results feed the bench line's ``parity`` block."""
    from oracle import reference_port as port
    n = min(n_utts, cfg["B"])
    torch.set_num_threads(threads if threads > 0 else (os.cpu_count() or 1))
    if joiner_state is None:
        jc = dict(input_dim=cfg["D"], output_dim=cfg["V"], inner_dim=max(cfg["I"], 1), activation=cfg["act"],
                  prune_range=cfg["R"], use_out_project=cfg["I"] > 0)
        from oracle.cases import make_weights
        import numpy as np
        joiner_state = {k: torch.from_numpy(v) for k, v in make_weights(jc, np.random.RandomState(1234)).items()}
        if cfg.get("ctc"):
            lin = torch.nn.Linear(cfg["D"], cfg["V"])
            joiner_state.update({"ctc_proj.weight": lin.weight.detach(), "ctc_proj.bias": lin.bias.detach()})
    last = {}

    def one():
        last.clear()
        last.update(port_once(cfg, batch, n, joiner_state))

    for _ in range(warmup):
        one()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return dict(value=n / sec, unit="utt/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{n} of {cfg['B']} utterances per step (same shapes/lengths as the GPU batch), "
                       f"{steps} steps after {warmup} warm-up, {sec:.2f} s/step"), sec, last


def parity_block(cfg, batch, ref, n, joiner, loss_mod, dev, mode, joiner_state):
    """The timed configuration checked against the CPU baseline's results on the SAME batch and weights (first n
    utterances, one eager step): relative loss errors, the rate of prune-range entries that differ, and the
    gradient of encoder_out over the utterances whose windows all agree (max |diff| / max |ref|)."""
    sub = {k: v[:n].to(dev) for k, v in batch.items()}
    enc = sub["enc"].detach().requires_grad_(True)
    pred = sub["pred"].detach().requires_grad_(True)
    joiner.zero_grad(set_to_none=True)  # drops the bucket aliases: this step goes through plain autograd
    ranges = None
    if joiner.prune_range > 0:
        logits, boundary, ranges, simple = joiner(enc, sub["t_len"], pred, sub["s_len"], sub["labels"])
        pruned = loss_mod({"logits": logits, "logits_length": sub["t_len"], "targets": sub["labels"],
                           "targets_length": sub["s_len"], "boundary": boundary, "ranges": ranges})
        total = 0.5 * simple + 0.5 * pruned
        if hasattr(joiner, "ctc_proj"):
            total = total + ctc_branch(joiner, enc, sub["t_len"], sub["labels"], sub["s_len"])
    else:
        logits, _, _, _ = joiner(enc, sub["t_len"], pred, sub["s_len"])
        total = loss_mod({"logits": logits, "logits_length": sub["t_len"], "targets": sub["labels"],
                          "targets_length": sub["s_len"]})
    total.backward()
    torch.cuda.synchronize()

    def rel(a, b):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

    out = {"utterances": n, "against": "cpu_baseline (oracle/reference_port.py, fp32) on the same batch and weights",
           "loss_rel": rel(total, ref["total"]), "tolerance": "fp32 1e-5 loss / 1e-4 gradients, bf16 joiner 1e-2 (north_star)",
           "mode": mode}
    g, gr = enc.grad.float().cpu(), ref["d_enc"]
    if "ranges" in ref:
        same = (ranges.cpu() == ref["ranges"]).all(dim=2).all(dim=1)
        out["ranges_mismatch"] = float((ranges.cpu() != ref["ranges"]).float().mean())
        out["utterances_with_identical_ranges"] = float(same.float().mean())
        out["simple_loss_rel"] = rel(simple, ref["simple"])
        out["pruned_loss_rel"] = rel(pruned, ref["pruned"])
        out["grad_rel"] = rel(g[same], gr[same]) if bool(same.any()) else None
        out["grad_rel_all_utterances"] = rel(g, gr)
        if out["ranges_mismatch"] > 0.0:
            # the window is an argmax over fp32 sums that tie up to rounding whenever several windows hold all of a
            # frame's occupation mass: the CPU path once more ON THE RANGES SELECTED HERE compares everything
            # downstream of the selection, over all utterances
            forced = port_once(cfg, batch, n, joiner_state, ranges_override=ranges.cpu())
            out["forced_ranges"] = {"pruned_loss_rel": rel(pruned, forced["pruned"]), "loss_rel": rel(total, forced["total"]),
                                    "grad_rel": rel(g, forced["d_enc"]),
                                    "grad_pred_rel": rel(pred.grad.float().cpu(), forced["d_pred"])}
    else:
        out["grad_rel"] = rel(g, gr)
    return out


def emit(line: dict) -> None:
    """The one JSON line goes to the real stdout; everything else this process (or NCCL's version banner,
    which is written to fd 1 from C) prints is diverted to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--mode", default=os.environ.get("S2T_BENCH_MODE", "bf16"), choices=["fp32", "bf16"],
                    help="joiner arithmetic: bf16 tensor cores (BASELINE config 3: 'bf16 joiner') or strict fp32")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-utts", type=int, default=64, help="utterances per CPU-baseline step")
    ap.add_argument("--cpu-steps", type=int, default=8, help="timed CPU-baseline steps (about 10 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="issue every step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()

    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": f"{args.workload}: {cfg['desc']}", "global_batch": cfg["B"] * world,
              "per_gpu_batch": cfg["B"], "T": cfg["T"], "U": cfg["U"], "V": cfg["V"], "D": cfg["D"],
              "prune_range": cfg["R"], "inner_dim": cfg["I"], "activation": cfg["act"],
              "parallelism": f"utterance-sharded x{world}"}

    # ----------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        batch = make_batch(cfg, 1234, input_dtype(cfg, args.mode))
        steps = max(1, min(args.steps, args.cpu_steps))
        base, sec, _ = cpu_arm(cfg, batch, args.cpu_utts, steps, min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "utt/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "utt/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        emit(line)
        return

    # ----------------------------------------------------------------------------- our arm
    import torch.distributed as dist
    from speech2text_b200 import _lib
    from speech2text_b200.distributed import FlatGradBucket

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback in the product path)"
    torch.cuda.set_device(local_rank)
    cpus_before = os.sched_getaffinity(0)
    try:  # run (and first-touch the pinned staging buffers) on the CPUs next to this rank's GPU: a host buffer on the
        # far socket costs a third of the host -> device bandwidth the end-to-end number is bound by
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:  # pragma: no cover - affinity is an optimisation, never a requirement
        pass
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()  # fail loudly if the CUDA library is missing

    in_dtype = input_dtype(cfg, args.mode)
    config["input_dtype"] = in_dtype
    batch = make_batch(cfg, 1234 + rank, in_dtype)
    joiner, loss_mod = build_modules(cfg, dev, args.mode)
    # dW kernels write straight into the all-reduce buffer; its three tail slots carry the logged losses
    bucket = FlatGradBucket(joiner.parameters(), extra_scalars=3).bind()
    d_in = {k: v.to(dev) for k, v in batch.items()}
    enc = d_in["enc"].requires_grad_(True)
    pred = d_in["pred"].requires_grad_(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_core():
        bucket.zero()
        enc.grad = None
        pred.grad = None
        return hot_path_step(joiner, loss_mod, enc, d_in["t_len"], pred, d_in["s_len"], d_in["labels"])

    def finish(losses):
        # ONE collective per step (gradients + logged scalars, averaged inside NCCL), part of the captured graph
        if world > 1:
            bucket.put_scalars(list(losses))
            bucket.all_reduce(average=True)
        return losses

    def step():  # eager: every kernel issued from Python
        return finish(step_core())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    def timed(n_steps, fn):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        for s, e in evs:
            flush.zero_()  # L2 flush, outside the timed interval
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in evs)

    # the step is a fixed sequence of ~350 short kernels: capture it once, replay it (speech2text_b200.graph)
    execution = "eager"
    run_step = step
    launches0 = _lib.launch_count()
    step()
    launches_per_step = _lib.launch_count() - launches0
    graphed = None
    graph_priority = int(os.environ.get("S2T_BENCH_GRAPH_PRIORITY", "0"))
    if not args.eager:
        from speech2text_b200.graph import GraphedStep
        if world > 1 and os.environ.get("S2T_BENCH_GRAPH_COLLECTIVE", "1") != "0":
            try:  # the all-reduce inside the graph: no launch gap between the last dW kernel and NCCL
                graphed = GraphedStep(step, priority=graph_priority)
                run_step = graphed
                execution = "cuda graph replay of the step (forward + backward + the gradient/scalar all-reduce)"
            except Exception as exc:
                graphed = None
                torch.cuda.synchronize()
                sys.stderr.write(f"graph capture with the collective failed ({type(exc).__name__}: {exc}); "
                                 "capturing the compute part only\n")
        if graphed is None:
            try:
                graphed = GraphedStep(step_core, priority=graph_priority)
                run_step = lambda: finish(graphed())
                execution = "cuda graph replay of the eager step (forward + backward)" + (
                    ", all-reduce issued after it" if world > 1 else "")
            except Exception as exc:  # keep measuring, but say what happened
                graphed = None
                execution = f"eager (graph capture failed: {type(exc).__name__}: {exc})"
                torch.cuda.synchronize()
    for _ in range(3):
        run_step()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    total_ms = timed(args.steps, run_step)
    barrier()
    launches = launches_per_step * args.steps
    clocks = sampler.stop()

    # per-kernel device time, same steps again with the library's event timer on
    _lib.profile_enable(True)
    prof_ms = timed(args.steps, step)
    _lib.profile_enable(False)
    report = _lib.profile_report()

    # end to end: pinned host inputs -> H2D -> public module API -> D2H of the losses
    h_in = {k: v.pin_memory() for k, v in batch.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in h_in.values())
    loss_host = torch.empty(3, dtype=torch.float32).pin_memory()

    # the copy of step i+1 travels on the prefetcher's stream while step i computes (every step still copies
    # its own inputs from pinned memory inside the timed region and reads its losses back)
    from speech2text_b200.prefetch import HostBatchPrefetcher
    pf = HostBatchPrefetcher(dev)

    def e2e_core(d):
        e = d["enc"].detach().requires_grad_(True)  # fresh autograd leaves over the slot's storage
        p = d["pred"].detach().requires_grad_(True)
        bucket.zero()
        return hot_path_step(joiner, loss_mod, e, d["t_len"], p, d["s_len"], d["labels"])

    slot_graphs = {}

    def e2e_step():
        d = pf.get()
        pf.put(h_in)
        key = d["enc"].data_ptr()
        if execution.startswith("cuda graph") and key not in slot_graphs:
            # one graph per device slot of the prefetcher (a graph reads fixed addresses)
            from speech2text_b200.graph import GraphedStep
            slot_graphs[key] = GraphedStep(lambda: e2e_core(d), warmup=1, pool=graphed.pool() if graphed else None)
        losses = slot_graphs[key]() if key in slot_graphs else e2e_core(d)
        vec = bucket.put_scalars(list(losses))
        if world > 1:
            bucket.all_reduce(average=True)
        loss_host.copy_(vec, non_blocking=True)

    pf.put(h_in)
    for _ in range(4):
        e2e_step()
    barrier()
    e2e_ms = timed(args.steps, e2e_step)
    barrier()

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = t.tolist()
    utts = cfg["B"] * world * args.steps
    value = utts / (total_ms / 1e3)
    e2e_value = utts / (e2e_ms / 1e3)

    if rank == 0:
        pk = peaks()
        work = kernel_work(cfg, float(batch["t_len"].sum()) / (cfg["B"] * cfg["T"]))
        roof = None
        kernel_roofs = {}

        def kernel_roof(name, cnt, ms):
            # work of all launches of this kernel in one step / their time in one step  (= per-launch work / average
            # launch duration, the launches of a step being equal shares of it)
            flops, nbytes = work.get(name, (0.0, 0.0))
            step_s = ms / args.steps / 1e3
            t_tc, t_hbm = flops / (pk["tc_sustained"] * 1e12), nbytes / (pk["hbm"] * 1e9)
            if t_tc > t_hbm:
                return {"bound": "tensor", "achieved": flops / step_s / 1e12, "peak": pk["tc_sustained"], "unit": "TFLOP/s",
                        "frac": t_tc / step_s}
            return {"bound": "hbm", "achieved": nbytes / step_s / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": t_hbm / step_s}

        if report:
            for kname, (kcnt, kms) in report.items():
                if kname not in work:
                    continue  # small helper kernels (operand packs, row metadata, scatters) carry no work model
                kr = kernel_roof(kname, kcnt, kms)
                kernel_roofs[kname] = {"bound": kr["bound"], "frac": round(kr["frac"], 4)}
            top = max(report.items(), key=lambda kv: kv[1][1])
            name, (cnt, ms) = top
            kr = kernel_roof(name, cnt, ms)
            traffic = None
            try:  # DRAM bytes per launch of this kernel from the committed ncu --set full capture
                rec = None
                for fname in ("r2_traffic.json", "r1_traffic.json"):
                    fpath = os.path.join(ROOT, "profiles", fname)
                    if rec is None and os.path.exists(fpath):
                        with open(fpath) as f:
                            rec = json.load(f).get(name)
                if rec:
                    traffic = rec["dram_bytes_read"] + rec["dram_bytes_write"]
            except OSError:
                pass
            roof = {"bound": kr["bound"], "kernel": name, "achieved": kr["achieved"], "peak": kr["peak"], "unit": kr["unit"],
                    "frac": kr["frac"], "traffic": traffic, "peak_source": pk["source"],
                    "avg_launch_ms": ms / cnt, "launches_timed": cnt,
                    "share_of_step": ms / max(sum(v[1] for v in report.values()), 1e-9),
                    "how": "library CUDA-event timer around every launch, eager second pass of the same steps; "
                           "share_of_step = this kernel's time / the sum over all of the library's kernels; bound = the "
                           "slower of algorithmic FLOPs / tensor peak and algorithmic bytes / HBM peak for this kernel"}
            if name.endswith("3xf16"):
                fl3 = 3.0 * work[name][0]
                roof["issued_tensor_frac"] = fl3 / (pk["tc_sustained"] * 1e12) / (ms / args.steps / 1e3)
                roof["note"] = ("3xF16 (hi/lo half split for fp32-level accuracy): the kernel issues 3x the algorithmic "
                                "MMA work; issued_tensor_frac = that work / tensor peak / time")
            # the whole step against SURVEY.md section 8(d)'s algorithmic work per utterance
            Bc, Tc, Sc, Vc, Dc, Rc, Ic = (cfg[k] for k in ("B", "T", "U", "V", "D", "R", "I"))
            if Rc > 0:
                fl = 6.0 * (Tc + Sc + 1) * Dc * Vc + 6.0 * (Sc + 1) * Tc * Vc + (12.0 * Tc * Rc * Vc * Ic if Ic > 0 else 0.0)
                fl += 6.0 * Tc * Dc * Vc if cfg.get("ctc") else 0.0
                if cfg.get("ctc"):  # Projector D -> V (fwd, dx, dW) and the CTC loss on its logits (two reads, one gradient write)
                    fl_ctc, by_ctc = 6.0 * Tc * Dc * Vc, 3.0 * Tc * Vc * 4 + 5.0 * Tc * (Sc + 1) * 4
                else:
                    fl_ctc = by_ctc = 0.0
                by = (by_ctc + 3.0 * (Tc + Sc + 1) * Dc * (2 if in_dtype == "bf16" else 4) + 6.0 * (Tc + Sc + 1) * Vc * 4 + 4.0 * (Sc * (Tc + 1) + (Sc + 1) * Tc) * 4
                      + 2.0 * Tc * Rc * 8 + 4.0 * Tc * Rc * 2 * 4)
            else:
                fl = 6.0 * (Tc + Sc + 1) * Dc * Vc + 12.0 * Tc * (Sc + 1) * Vc * max(Ic, 1)
                by = 3.0 * (Tc + Sc + 1) * Dc * 4 + 6.0 * (Tc + Sc + 1) * Vc * 4
            t_tc, t_hbm = Bc * fl / (pk["tc_sustained"] * 1e12), Bc * by / (pk["hbm"] * 1e9)
            step_s = total_ms / args.steps / 1e3
            roof["step"] = {"algorithmic_gflop_per_utt": fl / 1e9, "algorithmic_mb_per_utt": by / 1e6,
                            "bound": "tensor" if t_tc > t_hbm else "hbm", "roofline_ms": max(t_tc, t_hbm) * 1e3,
                            "frac": max(t_tc, t_hbm) / step_s,
                            "note": "SURVEY.md 8(d): padded T and S, logits never counted; the step is spread over "
                                    "~26 kernels (kernel_rooflines), none above a tenth of it"}
        kernels = {k: {"launches": c, "ms_per_step": ms / args.steps} for k, (c, ms) in
                   sorted(report.items(), key=lambda kv: -kv[1][1])}
        line = {"metric": METRIC, "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.mode == "fp32" else "bf16",
                "data": "synthetic", "config": dict(config, l2="flushed between steps (256 MiB memset, untimed)", execution=execution,
                                                    joiner_mode=args.mode),
                "e2e": {"value": e2e_value, "unit": "utt/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": 12, "ms_per_step": e2e_ms / args.steps,
                        "h2d": "pinned host -> device on a copy stream, one step ahead of the compute stream "
                               "(speech2text_b200.prefetch.HostBatchPrefetcher)"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernels_ms_per_step": kernels,
                "kernel_rooflines": kernel_roofs}
        if not args.no_cpu_baseline and world == 1:
            for tid in os.listdir("/proc/self/task"):  # the CPU arm gets every host core back
                try:
                    os.sched_setaffinity(int(tid), cpus_before)
                except OSError:
                    pass
            state = {k: v.detach().cpu() for k, v in joiner.state_dict().items()}
            base, _, ref = cpu_arm(cfg, batch, args.cpu_utts, args.cpu_steps, 1, joiner_state=state)
            # k2's own CPU schedule is serial over the batch: the 1-thread figure beside the all-thread one
            base1, sec1, _ = cpu_arm(cfg, batch, args.cpu_utts, 1, 0, joiner_state=state, threads=1)
            base["value_1thread"] = base1["value"]
            base["sample_1thread"] = f"1 step, {sec1:.2f} s"
            line["cpu_baseline"] = base
            line["parity"] = parity_block(cfg, batch, ref, min(args.cpu_utts, cfg["B"]), joiner, loss_mod, dev, args.mode, state)
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        # The captured graphs hold NCCL kernels; tearing the communicator down under them can block at exit.  Every
        # rank has finished its work: meet once more, then leave without the teardown.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
