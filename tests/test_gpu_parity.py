"""GPU parity tests: every call goes through the C ABI of libs2t_b200.so (via
speech2text_b200.functional / the Joiner + Loss modules) and is compared with
the CPU oracle on the same seeded inputs and with the golden vectors minted
from the reference.

Tolerances (north_star): fp32 relative 1e-5 on the loss, 1e-4 on gradients;
prune ranges bit-exact given identical occupation inputs.
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import VARIANTS, check_summary, load_golden, rel_err
from oracle import k2_shim as k2
from oracle import reference_port as port
from oracle.cases import CASES, make_case

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4


def _dev():
    return torch.device("cuda:0")


def _rand_lattice(B, S, T, seed, lens=None):
    g = torch.Generator().manual_seed(seed)
    px = torch.randn(B, S, T + 1, generator=g) - 1.0
    py = torch.randn(B, S + 1, T, generator=g) - 1.0
    boundary = torch.zeros(B, 4, dtype=torch.int64)
    if lens is None:
        boundary[:, 2] = torch.randint(1, S + 1, (B,), generator=g)
        boundary[:, 3] = torch.randint(1, T + 1, (B,), generator=g)
        boundary[0, 2], boundary[0, 3] = S, T
    else:
        boundary[:, 2] = torch.tensor(lens[0])
        boundary[:, 3] = torch.tensor(lens[1])
    return px, py, boundary


@pytest.mark.parametrize("B,S,T", [(1, 1, 1), (3, 4, 6), (5, 31, 33), (4, 40, 150), (2, 123, 327), (2, 300, 64)])
def test_mutual_information_matches_oracle(B, S, T):
    from speech2text_b200 import functional as F2
    px, py, boundary = _rand_lattice(B, S, T, seed=B * 1000 + S * 10 + T)
    # k2 semantics: no symbol after the last frame
    px = k2.fix_for_boundary(px, boundary)
    # truth: the oracle recursion in fp64 on the same fp32 inputs (the fp32 oracle, like k2's fp32
    # CPU kernel, is itself only good to a few ulp(|logP|) here)
    s_ref, (gx_ref, gy_ref) = k2.mutual_information_recursion(px.double(), py.double(), boundary,
                                                              return_grad=True)
    s32, (gx32, gy32) = k2.mutual_information_recursion(px, py, boundary, return_grad=True)
    s, (gx, gy) = F2.mutual_information_recursion(px.to(_dev()), py.to(_dev()), boundary.to(_dev()),
                                                  return_grad=True)
    torch.cuda.synchronize()
    assert rel_err(s, s_ref) < 5e-6
    assert rel_err(s, s32) < LOSS_RTOL
    assert (gx.cpu() - gx_ref).abs().max() < 2e-5  # occupation probabilities live in [0, 1]
    assert (gy.cpu() - gy_ref).abs().max() < 2e-5
    assert (gx.cpu() - gx32).abs().max() < 2e-5 + 2 * (gx32 - gx_ref).abs().max()
    assert (gy.cpu() - gy32).abs().max() < 2e-5 + 2 * (gy32 - gy_ref).abs().max()
    # without a boundary tensor
    s2 = F2.mutual_information_recursion(px.to(_dev()), py.to(_dev()), None)
    s2_ref = k2.mutual_information_recursion(px, py, None)
    assert rel_err(s2, s2_ref) < LOSS_RTOL


def _toy(B, T, S, V, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    am = torch.randn(B, T, V, generator=g) * scale
    lm = torch.randn(B, S + 1, V, generator=g) * scale
    sym = torch.randint(1, V, (B, S), generator=g)
    boundary = torch.zeros(B, 4, dtype=torch.int64)
    Tl = torch.randint(max(S, T // 2), T + 1, (B,), generator=g)
    Sl = torch.randint(1, S + 1, (B,), generator=g)
    Tl[0], Sl[0] = T, S
    Sl = torch.minimum(Sl, Tl)
    for b in range(B):
        sym[b, Sl[b]:] = 0
    boundary[:, 2], boundary[:, 3] = Sl, Tl
    return am, lm, sym, boundary


@pytest.mark.parametrize("B,T,S,V", [(2, 9, 3, 5), (3, 50, 17, 33), (4, 200, 15, 128), (2, 130, 129, 257)])
def test_simple_loss_and_grads_match_oracle(B, T, S, V):
    from speech2text_b200 import functional as F2
    am, lm, sym, boundary = _toy(B, T, S, V, seed=T + S + V)
    # truth = the oracle in fp64 on the same fp32 inputs
    am_r, lm_r = am.double().requires_grad_(True), lm.double().requires_grad_(True)
    loss_r, (gx_r, gy_r) = k2.rnnt_loss_smoothed(lm=lm_r, am=am_r, symbols=sym, termination_symbol=0,
                                                 lm_only_scale=0.0, am_only_scale=0.0, boundary=boundary,
                                                 reduction="none", return_grad=True)
    w = torch.linspace(0.5, 1.5, B)
    (loss_r * w.double()).sum().backward()

    am_g, lm_g = am.to(_dev()).requires_grad_(True), lm.to(_dev()).requires_grad_(True)
    loss_g, (gx, gy) = F2.rnnt_loss_smoothed(lm=lm_g, am=am_g, symbols=sym.to(_dev()), termination_symbol=0,
                                             boundary=boundary.to(_dev()), reduction="none", return_grad=True)
    (loss_g * w.to(_dev())).sum().backward()
    torch.cuda.synchronize()
    assert rel_err(loss_g, loss_r) < LOSS_RTOL
    # px / py themselves are fp32 here (as in k2), so the occupations carry their rounding
    assert (gx.cpu() - gx_r).abs().max() < 1e-4
    assert (gy.cpu() - gy_r).abs().max() < 1e-4
    assert rel_err(am_g.grad, am_r.grad) < GRAD_RTOL
    assert rel_err(lm_g.grad, lm_r.grad) < GRAD_RTOL


@pytest.mark.parametrize("variant", ["A", "B"])
@pytest.mark.parametrize("B,T,S,V,R", [(3, 40, 12, 9, 2), (3, 40, 12, 9, 5), (4, 200, 15, 64, 5),
                                       (2, 327, 123, 32, 5), (2, 30, 4, 8, 9), (2, 1100, 40, 8, 3)])
def test_prune_ranges_bit_exact(variant, B, T, S, V, R):
    """Identical occupation inputs (computed once by the oracle) -> identical int64 ranges."""
    from speech2text_b200 import functional as F2
    am, lm, sym, boundary = _toy(B, T, S, V, seed=R * 7 + T)
    _, (gx, gy) = k2.rnnt_loss_smoothed(lm=lm, am=am, symbols=sym, termination_symbol=0, lm_only_scale=0.0,
                                        am_only_scale=0.0, boundary=boundary, reduction="none",
                                        return_grad=True)
    ref = k2.get_rnnt_prune_ranges(gx, gy, boundary, R, variant=variant)
    got = F2.get_rnnt_prune_ranges(gx.to(_dev()), gy.to(_dev()), boundary.to(_dev()), R, variant=variant)
    torch.cuda.synchronize()
    assert got.dtype == torch.int64 and got.shape == ref.shape
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_pruned_loss_on_materialised_logits(dtype):
    from speech2text_b200 import functional as F2
    B, T, S, V, R = 3, 60, 14, 37, 5
    am, lm, sym, boundary = _toy(B, T, S, V, seed=11)
    _, (gx, gy) = k2.rnnt_loss_smoothed(lm=lm, am=am, symbols=sym, termination_symbol=0, lm_only_scale=0.0,
                                        am_only_scale=0.0, boundary=boundary, reduction="none",
                                        return_grad=True)
    ranges = k2.get_rnnt_prune_ranges(gx, gy, boundary, R)
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(B, T, R, V, generator=g) * 2).to(dtype)
    lr = logits.float().clone().requires_grad_(True)
    loss_r = k2.rnnt_loss_pruned(lr, sym, ranges, 0, boundary, delay_penalty=0.02, reduction="none")
    w = torch.linspace(0.5, 1.5, B)
    (loss_r * w).sum().backward()
    lg = logits.to(_dev()).requires_grad_(True)
    scores = F2.logits_scores(lg, sym.to(_dev()), ranges.to(_dev()), boundary.to(_dev()), blank=0,
                              delay_penalty=0.02)
    ((-scores) * w.to(_dev())).sum().backward()
    torch.cuda.synchronize()
    assert rel_err(-scores, loss_r) < LOSS_RTOL
    gtol = GRAD_RTOL if dtype == torch.float32 else 1e-2  # grad is rounded to the logits dtype
    assert rel_err(lg.grad.float(), lr.grad) < gtol
    assert lg.grad.dtype == dtype


@pytest.mark.parametrize("clamp", [-1.0, 0.05])
def test_vanilla_loss_on_materialised_logits_matches_torchaudio(clamp):
    import torchaudio
    from speech2text_b200.loss.rnnt_loss import RnntLoss, RnntLossConfig
    B, T, U, V = 3, 40, 9, 21
    g = torch.Generator().manual_seed(2)
    logits = torch.randn(B, T, U + 1, V, generator=g)
    tgt = torch.randint(1, V, (B, U), generator=g)
    Tl = torch.tensor([40, 33, 12])
    Ul = torch.tensor([9, 4, 7])
    lr = logits.clone().requires_grad_(True)
    ref = torchaudio.functional.rnnt_loss(lr, tgt.int(), Tl.int(), Ul.int(), blank=0, clamp=clamp,
                                          reduction="mean")
    ref.backward()
    mod = RnntLoss(RnntLossConfig(blank_label=0, clamp=clamp, reduction="mean"))
    lg = logits.to(_dev()).requires_grad_(True)
    out = mod(logits=lg, targets=tgt.to(_dev()), logits_length=Tl.to(_dev()), targets_length=Ul.to(_dev()))
    out.backward()
    torch.cuda.synchronize()
    assert rel_err(out, ref) < LOSS_RTOL
    assert rel_err(lg.grad, lr.grad) < GRAD_RTOL


def _run_modules(name, fused: bool, monkeypatch, mode: str = "fp32", variant: str = "B"):
    """One fwd+bwd through the drop-in modules exactly as rnnt_task.py:469-514 strings them."""
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.loss.loss import Loss
    monkeypatch.setenv("S2T_B200_FUSED", "1" if fused else "0")
    monkeypatch.setenv("S2T_B200_JOINER_MODE", mode)
    monkeypatch.setenv("S2T_B200_PRUNE_VARIANT", variant)
    spec, case = CASES[name], make_case(name)
    dev = _dev()
    joiner = Joiner(JoinerConfig(**spec["joiner"]))
    joiner.load_state_dict({k: torch.from_numpy(v) for k, v in case["weights"].items()})
    joiner = joiner.to(dev)
    enc = torch.from_numpy(case["encoder_out"]).to(dev).requires_grad_(True)
    pred = torch.from_numpy(case["predict_out"]).to(dev).requires_grad_(True)
    # lengths arrive as float tensors in the reference's tests (joiner_test.py:56-58)
    enc_len = torch.from_numpy(case["encoder_out_lengths"]).float().to(dev)
    tgt_len = torch.from_numpy(case["target_lengths"]).float().to(dev)
    tgt = torch.from_numpy(case["target"]).to(dev)
    out = {}
    if spec["joiner"].get("prune_range", 5) > 0:
        loss_mod = Loss({"model": "Pruned_Rnnt", "config": spec["loss"]})
        logits, boundary, ranges, simple = joiner(enc, enc_len, pred, tgt_len, tgt)
        pruned = loss_mod({"logits": logits, "logits_length": enc_len, "targets": tgt,
                           "targets_length": tgt_len, "boundary": boundary, "ranges": ranges})
        total = (spec["simple_loss_scale"] * simple + spec["pruned_loss_scale"] * pruned).mean()
        out.update(simple_loss=simple, pruned_loss=pruned, ranges=ranges, boundary=boundary, logits=logits)
    else:
        loss_mod = Loss({"model": "Rnnt", "config": spec["loss"]})
        logits, boundary, ranges, simple = joiner(enc, enc_len, pred, tgt_len)
        assert boundary is None and ranges is None and simple is None
        loss = loss_mod({"logits": logits, "logits_length": enc_len.long(), "targets": tgt,
                         "targets_length": tgt_len.long()})
        total = loss.mean()
        out.update(rnnt_loss=loss, logits=logits)
    total.backward()
    torch.cuda.synchronize()
    out["total_loss"] = total
    out["d_encoder_out"], out["d_predict_out"] = enc.grad, pred.grad
    for k, p in joiner.named_parameters():
        out["d" + k] = p.grad
    return out


SUPPORTED = list(CASES)  # includes tanh_smoothed: non-zero lm_scale / am_scale (k2's lm-only / am-only interpolation)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", SUPPORTED)
def test_training_step_matches_reference_goldens(name, fused, variant, monkeypatch):
    """Full hot path through the reference-facing modules vs the golden vectors minted by
    running the reference verbatim (oracle/make_golden.py), under both published variants of the
    prune-range selection."""
    if CASES[name]["joiner"].get("prune_range", 5) <= 0 and variant != VARIANTS[-1]:
        pytest.skip("the vanilla path has no prune ranges")
    gold = load_golden(name, "f32", variant)
    out = _run_modules(name, fused, monkeypatch, variant=variant)
    pruned = CASES[name]["joiner"].get("prune_range", 5) > 0
    assert tuple(out["logits"].shape) == tuple(gold["logits_shape"])
    if pruned:
        assert np.array_equal(out["boundary"].cpu().numpy(), gold["boundary"])
        np.testing.assert_allclose(out["simple_loss"].item(), gold["simple_loss"], rtol=LOSS_RTOL)
        mism = (out["ranges"].cpu().numpy() != gold["ranges"]).mean()
        _record("fp32_goldens", f"{name}.{variant}.{'fused' if fused else 'materialised'}", dict(ranges_mismatch=float(mism)))
        if variant == "A":
            assert mism == 0.0, f"{name}: {mism:.4%} of range entries differ from the reference run"
        elif mism > 0.0:
            # Variant B takes the argmax of cumulative-sum differences: every window that holds (nearly) all of a
            # frame's occupation mass gives the same sum up to fp32 rounding, so the reference's choice among them
            # is decided by the last bits of ITS occupation probabilities (test_prune_ranges_bit_exact pins the
            # selection itself, given identical inputs).  Everything downstream of the selection is then compared
            # at full tolerance against the oracle port run on the ranges chosen here.
            assert mism <= 0.02, f"{name}: {mism:.4%} of range entries differ from the reference run"
            _check_against_port_with_ranges(name, out, variant)
            return
        np.testing.assert_allclose(out["pruned_loss"].detach().cpu().double().numpy(), gold["pruned_loss"],
                                   rtol=LOSS_RTOL)
    else:
        np.testing.assert_allclose(out["rnnt_loss"].item(), gold["rnnt_loss"], rtol=LOSS_RTOL)
    np.testing.assert_allclose(out["total_loss"].item(), gold["total_loss"], rtol=LOSS_RTOL)
    for key in [k[:-len(".stride")] for k in gold if k.endswith(".stride") and k.startswith("d")]:
        check_summary(out[key], gold, key, rtol=GRAD_RTOL, what=name)


def _check_against_port_with_ranges(name, out, variant, ltol=LOSS_RTOL, gtol=GRAD_RTOL):
    spec, case = CASES[name], make_case(name)
    ref = port.training_step_loss(case["weights"], spec, case, prune_variant=variant,
                                  ranges_override=out["ranges"].cpu())
    np.testing.assert_allclose(out["pruned_loss"].detach().cpu().numpy(), ref["pruned_loss"].numpy(), rtol=ltol)
    np.testing.assert_allclose(out["total_loss"].item(), ref["total_loss"].item(), rtol=ltol)
    for key in [k for k in ref if k.startswith("d")]:
        err = rel_err(out[key], ref[key])
        assert err <= gtol, f"{name}[{variant}, forced ranges] {key}: {err:.3e}"


BF16_RTOL = 1e-2  # north_star: bf16-joiner relative 1e-2


# Tensor-core mode keeps everything that feeds the prune-range argmax at fp32 accuracy (3xF16 projections and
# normaliser, fp32 lattice): on the golden cases the ranges are identical to the reference run's; the bound below
# leaves room for one near-tie frame per case (north_star: ranges bit-exact GIVEN identical occupation inputs --
# test_prune_ranges_bit_exact; end to end the inputs differ in the last bits).
BF16_RANGE_MISMATCH = 0.01


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", SUPPORTED)
def test_bf16_tensor_core_joiner_matches_reference_goldens(name, variant, monkeypatch):
    """Tensor-core mode: bf16 operands / fp32 accumulation on tcgen05 for every joiner contraction (fwd + bwd), 3xF16
    projections and simple-loss normaliser, fp32 lattices.  Cases without the out-projection (the reference's
    zipformer yaml, joiner_test.py) have no joiner contraction: their act + log-sum-exp pass is the fp32 one."""
    if CASES[name]["joiner"].get("prune_range", 5) <= 0 and variant != VARIANTS[-1]:
        pytest.skip("the vanilla path has no prune ranges")
    gold = load_golden(name, "f32", variant)
    out = _run_modules(name, True, monkeypatch, mode="bf16", variant=variant)
    if "ranges" in out:
        mism = (out["ranges"].cpu().numpy() != gold["ranges"]).mean()
        _record("bf16_goldens", f"{name}.{variant}", dict(ranges_mismatch=float(mism)))
        assert mism <= (BF16_RANGE_MISMATCH if variant == "A" else 0.02), f"{name}: {mism:.2%} of range entries differ"
        np.testing.assert_allclose(out["simple_loss"].item(), gold["simple_loss"], rtol=BF16_RTOL)
        if mism > 0.0:  # near-tie windows (see the fp32 test): compare downstream of the selection
            _check_against_port_with_ranges(name, out, variant, BF16_RTOL, BF16_RTOL)
            return
        np.testing.assert_allclose(out["pruned_loss"].detach().cpu().double().numpy(), gold["pruned_loss"],
                                   rtol=BF16_RTOL)
    np.testing.assert_allclose(out["total_loss"].item(), gold["total_loss"], rtol=BF16_RTOL)
    for key in [k[:-len(".stride")] for k in gold if k.endswith(".stride") and k.startswith("d")]:
        check_summary(out[key], gold, key, rtol=BF16_RTOL, what=name + "[bf16]")


@pytest.mark.parametrize("keep_joint", [True, False])
@pytest.mark.parametrize("name", ["pruned_loss_test", "range_clamped"])
def test_bf16_chunked_backward_matches_single_chunk(name, keep_joint, monkeypatch):
    """The backward pass works on row chunks (bounded scratch); forcing tiny chunks, and the variant that
    rebuilds act(am + lm) on the fly instead of keeping it from forward, must give the same gradients."""
    ref = _run_modules(name, True, monkeypatch, mode="bf16")
    monkeypatch.setenv("S2T_B200_CHUNK_ROWS", "256")
    if not keep_joint:
        monkeypatch.setenv("S2T_B200_NO_KEEP_JOINT", "1")
    out = _run_modules(name, True, monkeypatch, mode="bf16")
    assert rel_err(out["total_loss"], ref["total_loss"]) < 1e-6
    for k in ref:
        if k.startswith("d"):
            assert rel_err(out[k], ref[k]) < 2e-3, k  # bf16 operands, different summation order


def test_bf16_full_size_c3_agrees_with_fp32_path(monkeypatch):
    """BASELINE config 3 (B=64, T=400, U=100, V=500, D=512, R=5, I=256): the tensor-core path against the
    strict-fp32 SIMT path of this library (itself pinned to the oracle at small sizes)."""
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.loss.loss import Loss
    B, T, U, V, D, R, I = 64, 400, 100, 500, 512, 5, 256
    g = torch.Generator().manual_seed(7)
    enc0 = torch.randn(B, T, D, generator=g)
    pred0 = torch.randn(B, U + 1, D, generator=g)
    tgt = torch.randint(1, V, (B, U), generator=g)
    t_len = torch.randint(int(0.7 * T), T + 1, (B,), generator=g)
    t_len[0] = T
    s_len = torch.clamp((t_len.float() * U / T * 0.9).long(), 1, U)
    s_len[0] = U
    cfg = JoinerConfig(input_dim=D, output_dim=V, inner_dim=I, activation="tanh", prune_range=R, use_out_project=True)
    torch.manual_seed(11)
    joiner = Joiner(cfg).to(_dev())
    results = {}
    for mode in ("fp32", "bf16"):
        monkeypatch.setenv("S2T_B200_FUSED", "1")
        monkeypatch.setenv("S2T_B200_JOINER_MODE", mode)
        joiner.zero_grad(set_to_none=True)
        enc = enc0.to(_dev()).requires_grad_(True)
        pred = pred0.to(_dev()).requires_grad_(True)
        loss_mod = Loss({"model": "Pruned_Rnnt", "config": {"termination_symbol": 0, "reduction": "mean"}})
        logits, boundary, ranges, simple = joiner(enc, t_len.to(_dev()), pred, s_len.to(_dev()), tgt.to(_dev()))
        pruned = loss_mod({"logits": logits, "logits_length": t_len.to(_dev()), "targets": tgt.to(_dev()),
                           "targets_length": s_len.to(_dev()), "boundary": boundary, "ranges": ranges})
        (0.5 * simple + pruned).backward()
        torch.cuda.synchronize()
        results[mode] = dict(simple=simple.detach(), pruned=pruned.detach(), ranges=ranges, d_enc=enc.grad, d_pred=pred.grad,
                             **{"d" + k: p.grad.clone() for k, p in joiner.named_parameters()})
    a, b = results["fp32"], results["bf16"]
    assert (a["ranges"] != b["ranges"]).float().mean() < 0.01
    assert rel_err(b["simple"], a["simple"]) < 1e-4
    assert rel_err(b["pruned"], a["pruned"]) < BF16_RTOL
    # a near-tie frame that picks the neighbouring window moves that utterance's gradient visibly: compare the
    # per-utterance gradients where both modes chose the same ranges, the weight gradients (sums over all) looser
    same = (a["ranges"] == b["ranges"]).all(dim=2).all(dim=1)
    assert same.float().mean() > 0.5, same.float().mean()
    for k in ("d_enc", "d_pred"):
        assert rel_err(b[k][same], a[k][same]) < 2 * BF16_RTOL, k
    for k in a:
        if k.startswith("d") and k not in ("d_enc", "d_pred"):
            assert rel_err(b[k], a[k]) < 5 * BF16_RTOL, k


# ---------------------------------------------------------------------------------------------
# BASELINE.json configurations at size, against the CPU port of the reference
# ---------------------------------------------------------------------------------------------
def _record(section: str, key: str, values: dict):
    """Measured parity numbers of this run -> gpurun_out/parity_r2.json (evidence, not an assertion)."""
    import json
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_r2.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = {}
        if os.path.exists(path):
            with open(path) as fh:
                data = json.load(fh)
        data.setdefault(section, {})[key] = values
        with open(path, "w") as fh:
            json.dump(data, fh, indent=1, sort_keys=True)
    except OSError:
        pass


# (B, T, U, V, D, R, I): c3 is the benchmarked configuration at full size; c4 (CTC hybrid: V=2000, U=125 per
# BASELINE.md section 3) and c5 (V=5000, D=1024, T=1000: 20 LSE column tiles, chunked backward) keep their
# per-utterance shape with the batch cut to what the CPU port finishes in seconds and in ~15 GB of host memory.
SIZE_CASES = {
    "c3": (64, 400, 100, 500, 512, 5, 256),
    "c4_b16": (16, 500, 125, 2000, 512, 5, 256),
    "c5_b3": (3, 1000, 250, 5000, 1024, 5, 256),
}


def _size_case(name, seed=1234):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    B, T, U, V, D, R, I = SIZE_CASES[name]
    cfg = dict(B=B, T=T, U=U, V=V, D=D, R=R, I=I, act="tanh")
    batch = bench.make_batch(cfg, seed)
    jc = dict(input_dim=D, output_dim=V, inner_dim=I, activation="tanh", prune_range=R, use_out_project=True)
    return cfg, batch, jc


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(SIZE_CASES))
def test_baseline_configs_at_size_match_the_reference_port(name, mode, monkeypatch):
    """One training step (rnnt_task.py:469-514) at BASELINE size through the drop-in modules against
    oracle.reference_port.training_step_loss on the same batch and weights: losses, prune ranges and every gradient.
    Tolerances are north_star's: fp32 1e-5 loss / 1e-4 gradients, bf16 joiner 1e-2."""
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.loss.loss import Loss
    from oracle.cases import make_weights
    cfg, batch, jc = _size_case(name)
    weights = make_weights(jc, np.random.RandomState(99))
    if mode == "bf16":
        # the bf16-joiner configuration feeds bf16 activations: both arms from the same rounded values
        batch["enc"] = batch["enc"].bfloat16().float()
        batch["pred"] = batch["pred"].bfloat16().float()
    spec = dict(joiner=jc, loss=dict(termination_symbol=0, reduction="mean"), simple_loss_scale=0.5,
                pruned_loss_scale=0.5)
    case = dict(encoder_out=batch["enc"], predict_out=batch["pred"], encoder_out_lengths=batch["t_len"],
                target_lengths=batch["s_len"], target=batch["labels"])
    torch.set_num_threads(os.cpu_count() or 1)
    ref = port.training_step_loss(weights, spec, case)

    monkeypatch.setenv("S2T_B200_FUSED", "1")
    monkeypatch.setenv("S2T_B200_JOINER_MODE", mode)
    monkeypatch.delenv("S2T_B200_PRUNE_VARIANT", raising=False)
    dev = _dev()
    joiner = Joiner(JoinerConfig(**jc))
    joiner.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    joiner = joiner.to(dev)
    loss_mod = Loss({"model": "Pruned_Rnnt", "config": spec["loss"]})
    enc = batch["enc"].detach().to(dev).requires_grad_(True)
    pred = batch["pred"].detach().to(dev).requires_grad_(True)
    t_len, s_len, labels = (batch[k].to(dev) for k in ("t_len", "s_len", "labels"))
    logits, boundary, ranges, simple = joiner(enc, t_len, pred, s_len, labels)
    pruned = loss_mod({"logits": logits, "logits_length": t_len, "targets": labels, "targets_length": s_len,
                       "boundary": boundary, "ranges": ranges})
    total = 0.5 * simple + 0.5 * pruned
    total.backward()
    torch.cuda.synchronize()

    ltol, gtol = (LOSS_RTOL, GRAD_RTOL) if mode == "fp32" else (BF16_RTOL, BF16_RTOL)
    mism = (ranges.cpu() != ref["ranges"]).float().mean().item()
    same = (ranges.cpu() == ref["ranges"]).all(dim=2).all(dim=1)  # utterances whose every window agrees
    if mism > 0.0:
        # near-tie windows (see test_training_step_matches_reference_goldens): the port once more on the ranges
        # selected here, so that everything downstream of the selection is compared at full tolerance
        ref = port.training_step_loss(weights, spec, case, ranges_override=ranges.cpu())
    got = {"d_encoder_out": enc.grad, "d_predict_out": pred.grad}
    got.update({"d" + k: p.grad for k, p in joiner.named_parameters()})
    errs = dict(ranges_mismatch=mism, utterances_with_identical_ranges=float(same.float().mean()),
                simple_loss_rel=abs(simple.item() - ref["simple_loss"].item()) / abs(ref["simple_loss"].item()),
                pruned_loss_rel=abs(pruned.item() - ref["pruned_loss"].item()) / abs(ref["pruned_loss"].item()),
                total_loss_rel=abs(total.item() - ref["total_loss"].item()) / abs(ref["total_loss"].item()))
    for k in got:
        errs[k] = rel_err(got[k], ref[k])
    _record("size_cases", f"{name}.{mode}", errs)
    assert mism <= 0.02, errs
    assert errs["simple_loss_rel"] <= (LOSS_RTOL if mode == "fp32" else 1e-4), errs
    assert errs["pruned_loss_rel"] <= ltol and errs["total_loss_rel"] <= ltol, errs
    exact = None
    for k in got:
        # weight gradients are sums over every lattice row of the batch (10^5..10^6 of them): fp32 summation order
        # (split-K partial sums met by atomics) and, in tensor-core mode, the bf16 operand rounding accumulate in them,
        # hence their wider band
        acts = k in ("d_encoder_out", "d_predict_out")
        lim = (gtol if acts else 2 * gtol) if mode == "fp32" else (2 * gtol if acts else 5 * gtol)
        if errs[k] <= lim:
            continue
        # At these lengths the reference's own fp32 lattice (k2's recursion, restated in oracle/mutual_information.c)
        # is only good to a few 1e-4: exp(alpha + p + beta - log P) with |log P| in the thousands.  The yardstick is
        # then the exact result of the reference's formulas (the port in fp64 without its float32 casts, on the same
        # ranges): the kernels must be at least as close to it as the reference's fp32 run is.
        assert mode == "fp32", (k, errs)
        if exact is None:
            exact = port.training_step_loss(weights, spec, case, dtype=torch.float64, ranges_override=ranges.cpu(),
                                            exact=True)
        e_ref, e_got = rel_err(ref[k], exact[k]), rel_err(got[k], exact[k])
        errs[k + ".vs_exact"] = dict(reference_fp32=e_ref, kernels=e_got)
        _record("size_cases", f"{name}.{mode}", errs)
        assert e_got <= max(lim, 1.25 * e_ref), (k, errs)


@pytest.mark.parametrize("name", ["joiner_test", "tanh_smoothed", "range_clamped"])
def test_lazy_logits_materialize_matches_port(name, monkeypatch):
    """The logits the fused path never stores, written out by the debug entry point."""
    spec, case = CASES[name], make_case(name)
    if spec["joiner"].get("lm_scale", 0.0) != 0.0:
        spec = dict(spec, joiner=dict(spec["joiner"], lm_scale=0.0, am_scale=0.0))
    from speech2text_b200.joiner import Joiner, JoinerConfig
    monkeypatch.setenv("S2T_B200_FUSED", "1")
    joiner = Joiner(JoinerConfig(**spec["joiner"]))
    joiner.load_state_dict({k: torch.from_numpy(v) for k, v in case["weights"].items()})
    joiner = joiner.to(_dev())
    args = [torch.from_numpy(case[k]).to(_dev()) for k in
            ("encoder_out", "encoder_out_lengths", "predict_out", "target_lengths", "target")]
    handle, boundary, ranges, simple = joiner(*args)
    w = {k: torch.from_numpy(v) for k, v in case["weights"].items()}
    cpu_args = [torch.from_numpy(case[k]) for k in
                ("encoder_out", "encoder_out_lengths", "predict_out", "target_lengths", "target")]
    _, _, ref_ranges, _ = port.joiner_forward(w, spec["joiner"], *cpu_args)
    assert (ranges.cpu() != ref_ranges).float().mean() <= 0.02  # near-tie windows of the cumulative variant
    # the logits are a function of the ranges: compare them on the ranges selected here
    ref_logits, _, _, ref_simple = port.joiner_forward(w, spec["joiner"], *cpu_args, ranges_override=ranges.cpu())
    assert tuple(handle.shape) == tuple(ref_logits.shape)
    got = handle.materialize()
    torch.cuda.synchronize()
    assert rel_err(got, ref_logits) < 1e-5
    assert rel_err(simple, ref_simple) < LOSS_RTOL


@pytest.mark.parametrize("scales", [(0.25, 0.1), (0.3, 0.0), (0.0, 0.2)])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_smoothed_simple_loss_matches_oracle(scales, mode):
    """k2.rnnt_loss_smoothed with non-zero lm_only_scale / am_only_scale: loss, occupation probabilities and the
    gradients w.r.t. am / lm (including the terms through the batch-wide unigram) against the fp64 oracle."""
    from speech2text_b200 import _lib
    from speech2text_b200 import functional as F2
    B, T, S, V = 3, 37, 11, 29
    g = torch.Generator().manual_seed(99)
    am = torch.randn(B, T, V, generator=g)
    lm = torch.randn(B, S + 1, V, generator=g)
    sym = torch.randint(1, V, (B, S), generator=g)
    boundary = torch.tensor([[0, 0, S, T], [0, 0, 7, 30], [0, 0, 3, 12]], dtype=torch.int64)
    am_r = am.double().requires_grad_(True)
    lm_r = lm.double().requires_grad_(True)
    loss_r, (gx_r, gy_r) = k2.rnnt_loss_smoothed(lm=lm_r, am=am_r, symbols=sym, termination_symbol=0,
                                                 lm_only_scale=scales[0], am_only_scale=scales[1], boundary=boundary,
                                                 reduction="mean", return_grad=True)
    loss_r.backward()
    am_g = am.to(_dev()).requires_grad_(True)
    lm_g = lm.to(_dev()).requires_grad_(True)
    m = _lib.MODE_BF16_TC if mode == "bf16" else _lib.MODE_FP32_SIMT
    loss_g, (gx, gy) = F2.rnnt_loss_smoothed(lm=lm_g, am=am_g, symbols=sym.to(_dev()), termination_symbol=0,
                                             lm_only_scale=scales[0], am_only_scale=scales[1],
                                             boundary=boundary.to(_dev()), reduction="mean", return_grad=True, mode=m)
    loss_g.backward()
    torch.cuda.synchronize()
    gtol = GRAD_RTOL if mode == "fp32" else BF16_RTOL
    assert rel_err(loss_g, loss_r) < LOSS_RTOL
    assert rel_err(gx, gx_r) < GRAD_RTOL and rel_err(gy, gy_r) < GRAD_RTOL
    assert rel_err(am_g.grad, am_r.grad) < gtol
    assert rel_err(lm_g.grad, lm_r.grad) < gtol


def test_full_size_properties_c3():
    """BASELINE config 3 size (B=64, T=400, U=100, V=500): size-independent properties of the
    occupation probabilities and ranges (the oracle would take minutes here)."""
    from speech2text_b200 import functional as F2
    B, T, S, V, R = 64, 400, 100, 500, 5
    g = torch.Generator(device="cuda").manual_seed(1234)
    am = torch.randn(B, T, V, generator=g, device="cuda") * 0.5
    lm = torch.randn(B, S + 1, V, generator=g, device="cuda") * 0.5
    sym = torch.randint(1, V - 1, (B, S), generator=g, device="cuda")
    Tl = torch.randint(int(0.6 * T), T + 1, (B,), generator=g, device="cuda")
    Sl = torch.clamp((Tl.float() * S / T * 0.9).long(), 1, S)
    Tl[0], Sl[0] = T, S
    boundary = torch.zeros(B, 4, dtype=torch.int64, device="cuda")
    boundary[:, 2], boundary[:, 3] = Sl, Tl
    loss, (gx, gy) = F2.rnnt_loss_smoothed(lm=lm, am=am, symbols=sym, termination_symbol=0, boundary=boundary,
                                           reduction="none", return_grad=True)
    ranges = F2.get_rnnt_prune_ranges(gx, gy, boundary, R)
    torch.cuda.synchronize()
    assert torch.isfinite(loss).all() and (loss > 0).all()
    t_idx = torch.arange(T, device="cuda")[None, :]
    live = (t_idx < Tl[:, None]).float()
    # exactly one blank per live frame, S_b symbols per utterance
    assert ((gy.sum(1) - live).abs().max() < 1e-3)
    assert ((gx.sum((1, 2)) - Sl.float()).abs().max() < 1e-2)
    assert (gx >= 0).all() and (gx <= 1 + 1e-4).all() and (gy >= 0).all() and (gy <= 1 + 1e-4).all()
    s0 = ranges[:, :, 0]
    assert (s0[:, 0] == 0).all()
    d = s0[:, 1:] - s0[:, :-1]
    assert (d >= 0).all() and (d <= R - 1).all()
    last = torch.gather(s0, 1, (Tl - 1)[:, None]).squeeze(1)
    assert torch.equal(last, torch.clamp(Sl - R + 1, min=0))
    assert torch.equal(ranges[:, :, 1:] - ranges[:, :, :-1], torch.ones_like(ranges[:, :, 1:]))


def test_host_batch_prefetcher_delivers_batches_in_order():
    from speech2text_b200.prefetch import HostBatchPrefetcher
    pf = HostBatchPrefetcher(_dev())
    host = [{"x": torch.full((1 << 20,), float(i)).pin_memory(), "n": torch.tensor([i]).pin_memory()} for i in range(4)]
    pf.put(host[0])
    for i in range(4):
        d = pf.get()
        if i + 1 < 4:
            pf.put(host[i + 1])
        assert d["x"].is_cuda and float(d["x"].sum().item()) == float(i) * (1 << 20)
        assert int(d["n"].item()) == i
    assert len(pf) == 0
    with pytest.raises(ValueError):
        pf.put({"x": torch.zeros(4)})
    pf.put(host[0])
    with pytest.raises(RuntimeError):  # depth 2: one batch in use, one in flight
        pf.put(host[1])


def test_graphed_step_replays_the_eager_step(monkeypatch):
    """speech2text_b200.graph.GraphedStep: the captured forward + backward gives the eager results."""
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.loss.loss import Loss
    from speech2text_b200.graph import GraphedStep
    monkeypatch.setenv("S2T_B200_FUSED", "1")
    monkeypatch.setenv("S2T_B200_JOINER_MODE", "bf16")
    name = "pruned_loss_test"
    spec, case = CASES[name], make_case(name)
    dev = _dev()
    joiner = Joiner(JoinerConfig(**spec["joiner"]))
    joiner.load_state_dict({k: torch.from_numpy(v) for k, v in case["weights"].items()})
    joiner = joiner.to(dev)
    loss_mod = Loss({"model": "Pruned_Rnnt", "config": spec["loss"]})
    enc = torch.from_numpy(case["encoder_out"]).to(dev).requires_grad_(True)
    pred = torch.from_numpy(case["predict_out"]).to(dev).requires_grad_(True)
    enc_len = torch.from_numpy(case["encoder_out_lengths"]).float().to(dev)
    tgt_len = torch.from_numpy(case["target_lengths"]).float().to(dev)
    tgt = torch.from_numpy(case["target"]).to(dev)

    def step():
        joiner.zero_grad(set_to_none=False)
        enc.grad = None
        pred.grad = None
        logits, boundary, ranges, simple = joiner(enc, enc_len, pred, tgt_len, tgt)
        pruned = loss_mod({"logits": logits, "logits_length": enc_len, "targets": tgt, "targets_length": tgt_len,
                           "boundary": boundary, "ranges": ranges})
        total = (0.5 * simple + pruned).mean()
        total.backward()
        return total

    for p in joiner.parameters():
        p.grad = torch.zeros_like(p)
    ref = step().detach().clone()
    torch.cuda.synchronize()
    ref_enc = enc.grad.clone()
    ref_w = {k: p.grad.clone() for k, p in joiner.named_parameters()}
    g = GraphedStep(step)
    enc.data.mul_(1.0)  # same data: replay must reproduce the eager numbers
    out = g()
    torch.cuda.synchronize()
    assert rel_err(out, ref) < 1e-6
    assert rel_err(enc.grad, ref_enc) < 2e-3  # atomics order + bf16 operands
    for k, p in joiner.named_parameters():
        assert rel_err(p.grad, ref_w[k]) < 2e-3, k
    # new data through the same graph
    with torch.no_grad():
        enc.mul_(0.5)
    out2 = g()
    torch.cuda.synchronize()
    eager2 = step().detach()
    torch.cuda.synchronize()
    assert rel_err(out2, eager2) < 1e-6


# fp32 gradients at this size: the reference's own fp32 recursion (torchaudio / k2) is ~1e-3 away from the fp64
# truth used here (|log P| ~ 1500); this library's fp32 lattice (fp64 offsets, fp32 values) lands at ~1e-4, so the
# bound against the fp64 truth is 2e-4 (the 1e-4 bar of the north star is against the fp32 reference itself).
@pytest.mark.parametrize("mode,tol", [("fp32", (LOSS_RTOL, 2 * GRAD_RTOL)), ("bf16", (BF16_RTOL, 2 * BF16_RTOL))])
def test_vanilla_full_size_c2_matches_torchaudio(mode, tol, monkeypatch):
    """BASELINE config 2 (vanilla full-lattice RNN-T, B=32 T=250 U=50 V=500 D=512): the fused joiner + loss against
    torchaudio's compiled rnnt_loss (the reference's vanilla back end, rnnt_loss.py:42-44) on the logits the
    reference's joiner ops produce from the same weights."""
    torchaudio = pytest.importorskip("torchaudio")
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.loss.loss import Loss
    B, T, U, V, D, I = 32, 250, 50, 500, 512, 256
    g = torch.Generator().manual_seed(21)
    enc0 = torch.randn(B, T, D, generator=g) * 0.5
    pred0 = torch.randn(B, U + 1, D, generator=g) * 0.5
    tgt = torch.randint(1, V, (B, U), generator=g)
    t_len = torch.randint(int(0.6 * T), T + 1, (B,), generator=g)
    s_len = torch.clamp((t_len.float() * U / T * 0.9).long(), 1, U)
    t_len[0], s_len[0] = T, U
    torch.manual_seed(5)
    joiner = Joiner(JoinerConfig(input_dim=D, output_dim=V, inner_dim=I, activation="tanh", prune_range=-1,
                                 use_out_project=True)).to(_dev())
    # reference chain in plain torch on the GPU (projections, add, tanh, out-projection), torchaudio loss on the CPU
    enc_r = enc0.to(_dev()).requires_grad_(True)
    pred_r = pred0.to(_dev()).requires_grad_(True)
    am = joiner._enc_proj(enc_r).unsqueeze(2)
    lm = joiner._pre_proj(pred_r).unsqueeze(1)
    logits = joiner._out_projection(torch.tanh(am + lm))
    # torchaudio's fp32 recursion carries ~1e-3 relative noise in the gradients at |log P| ~ 1500: it pins the loss;
    # the gradient reference is the oracle's k2 restatement run in fp64 with a range that covers the lattice
    logits_cpu = logits.detach().cpu().double().requires_grad_(True)
    ta = torchaudio.functional.rnnt_loss(logits_cpu.detach().float(), tgt.int(), t_len.int(), s_len.int(), blank=0,
                                         clamp=-1, reduction="mean")
    from oracle import k2_shim
    bnd = torch.zeros(B, 4, dtype=torch.int64)
    bnd[:, 2], bnd[:, 3] = s_len, t_len
    full = torch.arange(U + 1).expand(B, T, U + 1).contiguous()
    ref = k2_shim.rnnt_loss_pruned(logits_cpu, tgt, full, 0, bnd, reduction="mean")
    assert rel_err(ref, ta) < LOSS_RTOL
    ref.backward()
    logits.backward(logits_cpu.grad.float().to(_dev()))
    ref_grads = {"d_enc": enc_r.grad.clone(), "d_pred": pred_r.grad.clone(),
                 **{"d" + k: p.grad.clone() for k, p in joiner.named_parameters()}}
    joiner.zero_grad(set_to_none=True)
    del logits, am, lm
    # the drop-in path
    monkeypatch.setenv("S2T_B200_FUSED", "1")
    monkeypatch.setenv("S2T_B200_JOINER_MODE", mode)
    enc = enc0.to(_dev()).requires_grad_(True)
    pred = pred0.to(_dev()).requires_grad_(True)
    loss_mod = Loss({"model": "Rnnt", "config": {"blank_label": 0, "clamp": -1, "reduction": "mean"}})
    out, boundary, ranges, simple = joiner(enc, t_len.to(_dev()), pred, s_len.to(_dev()))
    assert boundary is None and ranges is None and simple is None and tuple(out.shape) == (B, T, U + 1, V)
    loss = loss_mod({"logits": out, "logits_length": t_len.to(_dev()), "targets": tgt.to(_dev()),
                     "targets_length": s_len.to(_dev())})
    loss.backward()
    torch.cuda.synchronize()
    assert rel_err(loss, ref) < tol[0]
    got = {"d_enc": enc.grad, "d_pred": pred.grad, **{"d" + k: p.grad for k, p in joiner.named_parameters()}}
    for k, v in ref_grads.items():
        assert rel_err(got[k], v) < tol[1], k


ODD_SHAPES = [
    # B, T, U, V, D, I, R, act
    (2, 7, 3, 33, 16, 24, 2, "tanh"),        # tiny, V % 4 != 0
    (3, 50, 20, 130, 64, 96, 5, "relu"),     # V % 4 != 0, I not a multiple of 64
    (2, 300, 260, 64, 32, 64, 5, "tanh"),    # S + 1 > 256: generic simple lattice
    (2, 64, 40, 72, 32, 40, 33, "relu"),     # R > 32: generic band lattice, several symbol-position windows
    (4, 128, 30, 257, 48, 300, 4, "tanh"),   # I > 256 (two 256-column tiles of the hidden layer), odd V
    (1, 700, 20, 40, 16, 32, 8, "relu"),     # one long utterance
    (5, 96, 95, 48, 24, 16, 3, "tanh"),      # U close to T: the band moves almost every frame
]


@pytest.mark.parametrize("shape", ODD_SHAPES, ids=lambda s: "x".join(str(v) for v in s))
def test_bf16_path_agrees_with_fp32_path_on_odd_shapes(shape, monkeypatch):
    """Shapes that leave the fast paths (odd vocabulary sizes, wide ranges, long transcripts, inner dimensions
    that need padding): the tensor-core path against the strict-fp32 path of the same library."""
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.loss.loss import Loss
    B, T, U, V, D, I, R, act = shape
    g = torch.Generator().manual_seed(sum(shape[:7]))
    enc0 = torch.randn(B, T, D, generator=g)
    pred0 = torch.randn(B, U + 1, D, generator=g)
    tgt = torch.randint(1, V, (B, U), generator=g)
    t_len = torch.randint(max(U, int(0.7 * T)), T + 1, (B,), generator=g)
    t_len[0] = T
    s_len = torch.minimum(torch.randint(max(1, U // 2), U + 1, (B,), generator=g), t_len)
    s_len[0] = U
    torch.manual_seed(3)
    joiner = Joiner(JoinerConfig(input_dim=D, output_dim=V, inner_dim=I, activation=act, prune_range=R,
                                 use_out_project=True)).to(_dev())
    res = {}
    for mode in ("fp32", "bf16"):
        monkeypatch.setenv("S2T_B200_FUSED", "1")
        monkeypatch.setenv("S2T_B200_JOINER_MODE", mode)
        joiner.zero_grad(set_to_none=True)
        enc = enc0.to(_dev()).requires_grad_(True)
        pred = pred0.to(_dev()).requires_grad_(True)
        loss_mod = Loss({"model": "Pruned_Rnnt", "config": {"termination_symbol": 0, "reduction": "mean"}})
        logits, boundary, ranges, simple = joiner(enc, t_len.to(_dev()), pred, s_len.to(_dev()), tgt.to(_dev()))
        pruned = loss_mod({"logits": logits, "logits_length": t_len.to(_dev()), "targets": tgt.to(_dev()),
                           "targets_length": s_len.to(_dev()), "boundary": boundary, "ranges": ranges})
        (0.5 * simple + pruned).backward()
        torch.cuda.synchronize()
        res[mode] = dict(simple=simple.detach(), pruned=pruned.detach(), ranges=ranges, d_enc=enc.grad, d_pred=pred.grad,
                         **{"d" + k: p.grad.clone() for k, p in joiner.named_parameters()})
    a, b = res["fp32"], res["bf16"]
    assert torch.isfinite(b["pruned"]).all() and torch.isfinite(b["simple"]).all()
    assert rel_err(b["simple"], a["simple"]) < 1e-4
    same = (a["ranges"] == b["ranges"]).all(dim=2).all(dim=1)
    if bool(same.all()):
        assert rel_err(b["pruned"], a["pruned"]) < BF16_RTOL
        for k in a:
            if k.startswith("d"):
                assert rel_err(b[k], a[k]) < 3 * BF16_RTOL, k
    else:  # a near-tie frame picked the neighbouring window: compare what is comparable
        assert rel_err(b["pruned"], a["pruned"]) < 5 * BF16_RTOL
        if bool(same.any()):
            assert rel_err(b["d_enc"][same], a["d_enc"][same]) < 3 * BF16_RTOL


@pytest.mark.parametrize("pair", [False, True], ids=["single_cta", "cta_pair"])
@pytest.mark.parametrize("shape", [(128, 128, 64, 128, 1), (300, 200, 100, 128, 1), (1000, 500, 300, 256, 1),
                                   (130, 256, 8192, 256, 7), (4096, 512, 512, 256, 1)],
                         ids=lambda s: "x".join(str(v) for v in s))
def test_streaming_contraction_kernel_single_cta_and_cta_pair(shape, pair, monkeypatch):
    """The tcgen05 contraction kernel behind every projection of the path, through its debug entry
    (s2t_tc_gemm: C = A B^T with bf16-rounded operands, fp32 accumulation): one CTA per tile and the
    cta_group::2 pair mode (M = 256 MMA across two SMs, half of B per CTA) against torch on the same
    bf16-rounded operands.  Odd sizes exercise the row / column / k tails and the odd last row tile of a pair."""
    import ctypes
    from speech2text_b200 import _lib
    L = _lib.lib()
    L.s2t_tc_gemm_workspace_bytes.restype = ctypes.c_size_t
    L.s2t_tc_gemm_workspace_bytes.argtypes = [ctypes.c_int] * 3
    L.s2t_tc_gemm.restype = ctypes.c_int
    L.s2t_tc_gemm.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 5 + [ctypes.c_void_p] * 2
    monkeypatch.setenv("S2T_GEMM_CLUSTER", "2" if pair else "1")
    M, N, K, bn, splits = shape
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g).to(_dev())
    B = torch.randn(N, K, generator=g).to(_dev())
    C = torch.full((M, N), float("nan"), device=_dev())
    ws = torch.empty(L.s2t_tc_gemm_workspace_bytes(M, N, K), dtype=torch.uint8, device=_dev())
    _lib.check(L.s2t_tc_gemm(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, bn, splits, _lib.ptr(ws), _lib.stream()))
    torch.cuda.synchronize()
    ref = A.bfloat16().double() @ B.bfloat16().double().t()
    assert not torch.isnan(C).any()
    assert ((C.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5


@pytest.mark.parametrize("shape", [(3, 37, 9, 70, 40, 4, 24), (2, 150, 40, 500, 64, 5, 256), (2, 61, 12, 1000, 48, 5, 200)],
                         ids=lambda s: "x".join(str(v) for v in s))
def test_fused_joiner_forward_matches_the_three_kernel_path(shape, monkeypatch):
    """The single-kernel joiner forward (gather -> hidden -> logits -> lse / px / py inside one CTA) against the
    joint_pack + hidden + logits + combine kernels it replaces: same operands, same accumulation order, so the
    log-probabilities agree to fp32 rounding of the log-sum-exp; gradients follow (the backward pass reads the hidden
    rows either path stored)."""
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.loss.loss import Loss
    B, T, U, V, D, R, I = shape
    g = torch.Generator().manual_seed(5)
    enc0 = torch.randn(B, T, D, generator=g) * 0.7
    pred0 = torch.randn(B, U + 1, D, generator=g) * 0.7
    tgt = torch.randint(1, V, (B, U), generator=g)
    t_len = torch.randint(max(U, int(0.6 * T)), T + 1, (B,), generator=g)
    t_len[0] = T
    s_len = torch.clamp((t_len.float() * U / T * 0.9).long(), 1, U)
    s_len[0] = U
    torch.manual_seed(3)
    joiner = Joiner(JoinerConfig(input_dim=D, output_dim=V, inner_dim=I, activation="tanh", prune_range=R)).to(_dev())
    monkeypatch.setenv("S2T_B200_FUSED", "1")
    monkeypatch.setenv("S2T_B200_JOINER_MODE", "bf16")
    res = {}
    for name, env in (("fused", None), ("split", "1")):
        if env is None:
            monkeypatch.delenv("S2T_B200_NO_FUSED_FWD", raising=False)
        else:
            monkeypatch.setenv("S2T_B200_NO_FUSED_FWD", env)
        joiner.zero_grad(set_to_none=True)
        enc = enc0.to(_dev()).requires_grad_(True)
        pred = pred0.to(_dev()).requires_grad_(True)
        loss_mod = Loss({"model": "Pruned_Rnnt", "config": {"termination_symbol": 0, "reduction": "none"}})
        logits, boundary, ranges, simple = joiner(enc, t_len.to(_dev()), pred, s_len.to(_dev()), tgt.to(_dev()))
        pruned = loss_mod({"logits": logits, "logits_length": t_len.to(_dev()), "targets": tgt.to(_dev()),
                           "targets_length": s_len.to(_dev()), "boundary": boundary, "ranges": ranges})
        (0.5 * simple + pruned.sum()).backward()
        torch.cuda.synchronize()
        res[name] = dict(pruned=pruned.detach().clone(), ranges=ranges.clone(), d_enc=enc.grad.clone(),
                         d_pred=pred.grad.clone(), **{"d" + k: p.grad.clone() for k, p in joiner.named_parameters()})
    a, b = res["split"], res["fused"]
    assert torch.equal(a["ranges"], b["ranges"])
    assert torch.isfinite(b["pruned"]).all()
    assert rel_err(b["pruned"], a["pruned"]) < 1e-5
    for k in a:
        if k.startswith("d"):
            assert rel_err(b[k], a[k]) < 5e-3, k  # 1-ulp differences of lse move the occupation probabilities by ~3e-5


def test_bound_gradient_bucket_receives_the_same_gradients(monkeypatch):
    """FlatGradBucket.bind(): the weight-gradient kernels write straight into the flat all-reduce buffer
     and autograd's accumulate is skipped; the buffer must hold what autograd would have
    produced, also when the step is repeated (overwrite, not accumulate)."""
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.loss.loss import Loss
    from speech2text_b200.distributed import FlatGradBucket
    monkeypatch.setenv("S2T_B200_FUSED", "1")
    monkeypatch.setenv("S2T_B200_JOINER_MODE", "bf16")
    name = "pruned_loss_test"
    spec, case = CASES[name], make_case(name)
    dev = _dev()
    joiner = Joiner(JoinerConfig(**spec["joiner"]))
    joiner.load_state_dict({k: torch.from_numpy(v) for k, v in case["weights"].items()})
    joiner = joiner.to(dev)
    loss_mod = Loss({"model": "Pruned_Rnnt", "config": spec["loss"]})
    enc = torch.from_numpy(case["encoder_out"]).to(dev).requires_grad_(True)
    pred = torch.from_numpy(case["predict_out"]).to(dev).requires_grad_(True)
    enc_len = torch.from_numpy(case["encoder_out_lengths"]).float().to(dev)
    tgt_len = torch.from_numpy(case["target_lengths"]).float().to(dev)
    tgt = torch.from_numpy(case["target"]).to(dev)

    def step():
        enc.grad = None
        pred.grad = None
        logits, boundary, ranges, simple = joiner(enc, enc_len, pred, tgt_len, tgt)
        pruned = loss_mod({"logits": logits, "logits_length": enc_len, "targets": tgt, "targets_length": tgt_len,
                           "boundary": boundary, "ranges": ranges})
        total = 0.5 * simple + 0.5 * pruned
        total.backward()
        return total.detach()

    joiner.zero_grad(set_to_none=True)
    ref_loss = step()
    ref = {k: p.grad.clone() for k, p in joiner.named_parameters()}
    ref_enc = enc.grad.clone()
    bucket = FlatGradBucket(joiner.parameters()).bind()
    for _ in range(2):  # second pass: overwritten, not accumulated
        bucket.zero()
        out = step()
    torch.cuda.synchronize()
    assert rel_err(out, ref_loss) < 1e-6
    assert rel_err(enc.grad, ref_enc) < 2e-3  # atomics order
    off = 0
    for k, p in joiner.named_parameters():
        assert p.grad.data_ptr() == bucket.flat.data_ptr() + 4 * off, k  # still the bucket view
        assert rel_err(p.grad, ref[k]) < 2e-3, k
        off += p.numel()
    bucket.unbind()


def test_padding_tiles_are_skipped_without_changing_results(monkeypatch):
    """Row tiles of the joiner lattice that hold padding frames only are never computed (live-tile list).
    Same step with the skip disabled (S2T_B200_NO_DEAD_SKIP) must give the same losses and gradients; the
    lengths leave most of some utterances' frames as padding and the backward runs in several row chunks."""
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.loss.loss import Loss
    monkeypatch.setenv("S2T_B200_FUSED", "1")
    monkeypatch.setenv("S2T_B200_JOINER_MODE", "bf16")
    monkeypatch.setenv("S2T_B200_CHUNK_ROWS", "1024")
    dev = _dev()
    B, T, U, V, D = 6, 230, 40, 200, 128
    g = torch.Generator().manual_seed(77)
    torch.manual_seed(5)
    joiner = Joiner(JoinerConfig(input_dim=D, output_dim=V, inner_dim=64, activation="tanh", prune_range=5)).to(dev)
    loss_mod = Loss({"model": "Pruned_Rnnt", "config": {"termination_symbol": 0, "reduction": "mean"}})
    enc = (torch.randn(B, T, D, generator=g) * 0.5).to(dev).requires_grad_(True)
    pred = (torch.randn(B, U + 1, D, generator=g) * 0.5).to(dev).requires_grad_(True)
    t_len = torch.tensor([230, 60, 200, 33, 128, 129], device=dev)
    s_len = torch.tensor([40, 12, 31, 7, 40, 25], device=dev)
    labels = torch.randint(1, V - 1, (B, U), generator=g)
    for b in range(B):
        labels[b, s_len[b]:] = 0
    labels = labels.to(dev)

    def step():
        joiner.zero_grad(set_to_none=True)
        enc.grad = None
        pred.grad = None
        logits, boundary, ranges, simple = joiner(enc, t_len, pred, s_len, labels)
        pruned = loss_mod({"logits": logits, "logits_length": t_len, "targets": labels, "targets_length": s_len,
                           "boundary": boundary, "ranges": ranges})
        (0.5 * simple + 0.5 * pruned).backward()
        torch.cuda.synchronize()
        return (simple.detach().clone(), pruned.detach().clone(), enc.grad.clone(), pred.grad.clone(),
                {k: p.grad.clone() for k, p in joiner.named_parameters()})

    skip = step()
    monkeypatch.setenv("S2T_B200_NO_DEAD_SKIP", "1")
    full = step()
    assert torch.isfinite(skip[2]).all() and torch.isfinite(skip[3]).all()
    assert rel_err(skip[0], full[0]) < 1e-6 and rel_err(skip[1], full[1]) < 1e-6
    assert rel_err(skip[2], full[2]) < 1e-4 and rel_err(skip[3], full[3]) < 1e-4  # summation order of the reductions
    for k in skip[4]:
        assert torch.isfinite(skip[4][k]).all(), k
        assert rel_err(skip[4][k], full[4][k]) < 1e-4, k
    # padding frames receive exactly zero gradient
    for b in range(B):
        assert float(skip[2][b, int(t_len[b]):].abs().max()) == 0.0 if int(t_len[b]) < T else True


def test_linear_tc_reads_bf16_activations_natively():
    """bf16 activations (BASELINE config 3's "bf16 joiner") go through the projection as they are: the forward must
    equal the fp32 path on the same (bf16-representable) values bit for bit, the row-max by-product must be the max of
    the stored result, and the gradient comes back as bf16 (the fp32 gradient rounded once)."""
    from speech2text_b200 import functional as F2
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    for M, K, N in ((300, 96, 70), (1000, 512, 500)):
        x = (torch.randn(2, M // 2, K, generator=g) * 0.5).bfloat16().to(dev)
        W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
        b = torch.randn(N, generator=g).to(dev)
        dy = torch.randn(2, M // 2, N, generator=g).to(dev)
        xb = x.clone().requires_grad_(True)
        xf = x.float().requires_grad_(True)
        yb, _, rm = F2.linear_tc_pair(xb, W, b, row_max=True)
        yf = F2.linear_tc(xf, W, b)
        assert torch.equal(yb, yf)
        assert torch.equal(rm, yb.max(dim=-1).values)
        ref = torch.nn.functional.linear(x.double(), W.double(), b.double())
        assert rel_err(yb, ref) < 1e-5
        (yb * dy).sum().backward()
        (yf * dy).sum().backward()
        torch.cuda.synchronize()
        assert xb.grad.dtype == torch.bfloat16 and xf.grad.dtype == torch.float32
        assert torch.equal(xb.grad, xf.grad.bfloat16())


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_simple_loss_gradient_issued_next_to_the_band_lattice(mode, monkeypatch):
    """The simple loss's gradient contractions may be issued ahead of the backward pass, next to the joiner's band
    lattice, with the upstream scale of the previous step (functional._EarlySimpleBackward); backward corrects the scale.
    Same kernels, same operands: the gradients equal the plain path's (up to the order of fp32 atomics) when the prediction was right, and up
    to the rounding of one extra multiply (fp32 mode) or of the bf16 operand coef * W (tensor-core mode) when it was not.  Scales 0.5 -> 0.5 -> 0.125 -> 0 -> 2 cover: first use
    (prediction = 1), a right prediction, a wrong one, a zero scale (the prediction must not become 0), recovery."""
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.loss.loss import Loss
    from speech2text_b200 import functional as F2
    B, T, U, V, D, R, I = 3, 90, 21, 120, 48, 5, 256
    g = torch.Generator().manual_seed(11)
    enc0 = torch.randn(B, T, D, generator=g) * 0.7
    pred0 = torch.randn(B, U + 1, D, generator=g) * 0.7
    tgt = torch.randint(1, V, (B, U), generator=g).to(_dev())
    t_len = torch.tensor([T, T - 17, T - 40]).to(_dev())
    s_len = torch.tensor([U, U - 3, U - 9]).to(_dev())
    torch.manual_seed(3)
    joiner = Joiner(JoinerConfig(input_dim=D, output_dim=V, inner_dim=I, activation="tanh", prune_range=R)).to(_dev())
    loss_mod = Loss({"model": "Pruned_Rnnt", "config": {"termination_symbol": 0, "reduction": "mean"}})
    monkeypatch.setenv("S2T_B200_FUSED", "1")
    monkeypatch.setenv("S2T_B200_JOINER_MODE", mode)

    def step(scale):
        joiner.zero_grad(set_to_none=True)
        enc = enc0.to(_dev()).requires_grad_(True)
        pred = pred0.to(_dev()).requires_grad_(True)
        logits, boundary, ranges, simple = joiner(enc, t_len, pred, s_len, tgt)
        pruned = loss_mod({"logits": logits, "logits_length": t_len, "targets": tgt, "targets_length": s_len,
                           "boundary": boundary, "ranges": ranges})
        (scale * simple + 0.5 * pruned).backward()
        torch.cuda.synchronize()
        return dict(d_enc=enc.grad.clone(), d_pred=pred.grad.clone(),
                    **{"d" + k: p.grad.clone() for k, p in joiner.named_parameters()})

    scales = [0.5, 0.5, 0.125, 0.0, 2.0]
    monkeypatch.setenv("S2T_B200_EARLY_SIMPLE_BWD", "0")
    plain = [step(s) for s in scales]
    monkeypatch.setenv("S2T_B200_EARLY_SIMPLE_BWD", "1")
    F2._PRED.clear()
    launched = []
    real_launch = F2._EarlySimpleBackward.launch
    monkeypatch.setattr(F2._EarlySimpleBackward, "launch", lambda self: (launched.append(1), real_launch(self))[1])
    early = [step(s) for s in scales]
    assert len(launched) == len(scales) and not F2._PENDING  # every step took the early path
    for i, (a, b) in enumerate(zip(plain, early)):
        for k in a:
            if i == 1:  # prediction right: nothing is rescaled; what is left is the order noise of fp32 atomics
                assert rel_err(b[k], a[k]) < (1e-5 if k not in ("d_enc", "d_pred") else 1e-6), (i, k, rel_err(b[k], a[k]))
            else:
                # tensor-core mode rounds coef * W to bf16: a different coef is a different (equally good) rounding
                assert rel_err(b[k], a[k]) < (1e-2 if mode == "bf16" else 1e-6), (i, k, rel_err(b[k], a[k]))
    assert all(float(v.abs().min()) > 0 for v in F2._PRED.values())


def test_backward_uses_the_layout_its_forward_carved(monkeypatch):
    """The workspace layout (backward row chunks, kept J, skipped padding tiles) depends on test hooks read from the
    environment; the backward call takes what ITS forward call used (remembered by workspace address) even if the
    environment changed in between."""
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.loss.loss import Loss
    B, T, U, V, D, R, I = 3, 120, 30, 300, 64, 5, 256
    g = torch.Generator().manual_seed(21)
    enc0 = torch.randn(B, T, D, generator=g) * 0.7
    pred0 = torch.randn(B, U + 1, D, generator=g) * 0.7
    tgt = torch.randint(1, V, (B, U), generator=g).to(_dev())
    t_len = torch.tensor([T, T - 31, T - 50]).to(_dev())
    s_len = torch.tensor([U, U - 5, U - 12]).to(_dev())
    torch.manual_seed(3)
    joiner = Joiner(JoinerConfig(input_dim=D, output_dim=V, inner_dim=I, activation="tanh", prune_range=R)).to(_dev())
    loss_mod = Loss({"model": "Pruned_Rnnt", "config": {"termination_symbol": 0, "reduction": "mean"}})
    monkeypatch.setenv("S2T_B200_FUSED", "1")
    monkeypatch.setenv("S2T_B200_JOINER_MODE", "bf16")

    def run(switch_env):
        joiner.zero_grad(set_to_none=True)
        monkeypatch.setenv("S2T_B200_CHUNK_ROWS", "256")
        monkeypatch.setenv("S2T_B200_NO_KEEP_JOINT", "1")
        enc = enc0.to(_dev()).requires_grad_(True)
        pred = pred0.to(_dev()).requires_grad_(True)
        logits, boundary, ranges, simple = joiner(enc, t_len, pred, s_len, tgt)
        pruned = loss_mod({"logits": logits, "logits_length": t_len, "targets": tgt, "targets_length": s_len,
                           "boundary": boundary, "ranges": ranges})
        if switch_env:
            monkeypatch.delenv("S2T_B200_CHUNK_ROWS")
            monkeypatch.delenv("S2T_B200_NO_KEEP_JOINT")
        (0.5 * simple + 0.5 * pruned).backward()
        torch.cuda.synchronize()
        return dict(d_enc=enc.grad.clone(), d_pred=pred.grad.clone(),
                    **{"d" + k: p.grad.clone() for k, p in joiner.named_parameters()})

    a, b = run(False), run(True)
    for k in a:
        assert torch.isfinite(b[k]).all(), k
        assert rel_err(b[k], a[k]) < 1e-5, (k, rel_err(b[k], a[k]))
