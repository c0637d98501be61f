"""CPU tests of the oracle itself: the restatement must reproduce the golden
vectors minted from the reference (oracle/make_golden.py) and agree with the
independent anchors (torchaudio, fp64 autograd through a loop DP)."""
import numpy as np
import pytest
import torch

from conftest import VARIANTS, check_summary, load_golden
from oracle import k2_shim as k2
from oracle import reference_port as port
from oracle.cases import CASES, make_case


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_port_matches_reference_goldens(name, tag, variant):
    spec = CASES[name]
    pruned = spec["joiner"].get("prune_range", 5) > 0
    if tag == "f64" and not pruned:
        pytest.skip("torchaudio has no fp64 kernel")
    if not pruned and variant != VARIANTS[-1]:
        pytest.skip("the vanilla path has no prune ranges")
    gold = load_golden(name, tag, variant)
    case = make_case(name)
    dtype = torch.float32 if tag == "f32" else torch.float64
    out = port.training_step_loss(case["weights"], spec, case, dtype=dtype, prune_variant=variant)
    tol = 2e-6 if tag == "f32" else 1e-12
    if pruned:
        assert np.array_equal(out["ranges"].numpy(), gold["ranges"])
        assert np.array_equal(out["boundary"].numpy(), gold["boundary"])
        np.testing.assert_allclose(out["simple_loss"].double().numpy(), gold["simple_loss"], rtol=tol)
        np.testing.assert_allclose(out["pruned_loss"].double().numpy(), gold["pruned_loss"], rtol=tol)
    else:
        np.testing.assert_allclose(out["rnnt_loss"].double().numpy(), gold["rnnt_loss"], rtol=tol)
    for key in [k[:-len(".stride")] for k in gold if k.endswith(".stride") and k.startswith("d")]:
        check_summary(out[key], gold, key, rtol=50 * tol, what=f"{name}.{tag}")


def test_f32_and_f64_goldens_agree():
    """fp32 reference run vs fp64 reference run: the tolerance budget of north_star
    (1e-5 loss, 1e-4 grads) has to be achievable by fp32 arithmetic at all."""
    for variant, (name, spec) in [(v, c) for v in VARIANTS for c in CASES.items()]:
        if spec["joiner"].get("prune_range", 5) <= 0:
            continue
        g32, g64 = load_golden(name, "f32", variant), load_golden(name, "f64", variant)
        np.testing.assert_allclose(g32["pruned_loss"], g64["pruned_loss"], rtol=1e-5)
        np.testing.assert_allclose(g32["simple_loss"], g64["simple_loss"], rtol=1e-5)
        mism = (g32["ranges"] != g64["ranges"]).mean()
        assert mism < 0.01, (name, mism)


def _toy(B=3, T=20, S=7, V=11, seed=0):
    g = torch.Generator().manual_seed(seed)
    am = torch.randn(B, T, V, generator=g)
    lm = torch.randn(B, S + 1, V, generator=g)
    sym = torch.randint(1, V, (B, S), generator=g)
    Tl = torch.tensor([T, T - 5, T - 11])[:B]
    Sl = torch.tensor([S, S - 2, S - 4])[:B]
    boundary = torch.zeros(B, 4, dtype=torch.int64)
    boundary[:, 2] = Sl
    boundary[:, 3] = Tl
    return am, lm, sym, Tl, Sl, boundary


def test_simple_equals_torchaudio_on_trivial_joiner():
    import torchaudio
    am, lm, sym, Tl, Sl, boundary = _toy()
    am.requires_grad_(True)
    loss, _ = k2.rnnt_loss_smoothed(lm=lm, am=am, symbols=sym, termination_symbol=0, lm_only_scale=0.0,
                                    am_only_scale=0.0, boundary=boundary, reduction="none", return_grad=True)
    logits = am[:, :, None, :] + lm[:, None, :, :]
    ta = torchaudio.functional.rnnt_loss(logits, sym.int(), Tl.int(), Sl.int(), blank=0, reduction="none")
    torch.testing.assert_close(loss, ta, rtol=1e-5, atol=1e-5)
    g1, = torch.autograd.grad(loss.sum(), am, retain_graph=True)
    g2, = torch.autograd.grad(ta.sum(), am)
    torch.testing.assert_close(g1, g2, rtol=1e-4, atol=2e-5)


def test_pruned_equals_torchaudio_when_range_covers_lattice():
    import torchaudio
    am, lm, sym, Tl, Sl, boundary = _toy(seed=1)
    S, V = sym.shape[1], am.shape[2]
    _, (pxg, pyg) = k2.rnnt_loss_smoothed(lm=lm, am=am, symbols=sym, termination_symbol=0, lm_only_scale=0.0,
                                          am_only_scale=0.0, boundary=boundary, reduction="none",
                                          return_grad=True)
    ranges = k2.get_rnnt_prune_ranges(pxg, pyg, boundary, S + 3)
    assert ranges.shape[2] == S + 1
    W = torch.randn(V, V, generator=torch.Generator().manual_seed(5))
    am_p, lm_p = k2.do_rnnt_pruning(am, lm, ranges)
    pl = k2.rnnt_loss_pruned(torch.tanh(am_p + lm_p) @ W, sym, ranges, 0, boundary, reduction="none")
    full = torch.tanh(am[:, :, None, :] + lm[:, None, :, :]) @ W
    ta = torchaudio.functional.rnnt_loss(full, sym.int(), Tl.int(), Sl.int(), blank=0, reduction="none")
    torch.testing.assert_close(pl, ta, rtol=1e-5, atol=1e-5)


def test_occupation_probs_match_fp64_autograd_of_loop_dp():
    """Independent check of mutual_information backward: differentiate a plain
    python loop DP in fp64 with autograd."""
    g = torch.Generator().manual_seed(3)
    B, S, T = 2, 4, 6
    px = torch.randn(B, S, T + 1, generator=g, dtype=torch.float64)
    py = torch.randn(B, S + 1, T, generator=g, dtype=torch.float64)
    boundary = torch.tensor([[0, 0, 4, 6], [0, 0, 2, 5]])
    px.requires_grad_(True)
    py.requires_grad_(True)
    tot = []
    for b in range(B):
        sb, tb, se, te = boundary[b].tolist()
        p = {(sb, tb): torch.zeros((), dtype=torch.float64)}
        for s in range(sb, se + 1):
            for t in range(tb, te + 1):
                if (s, t) == (sb, tb):
                    continue
                terms = []
                if s > sb:
                    terms.append(p[(s - 1, t)] + px[b, s - 1, t])
                if t > tb:
                    terms.append(p[(s, t - 1)] + py[b, s, t - 1])
                p[(s, t)] = torch.logsumexp(torch.stack(terms), 0)
        tot.append(p[(se, te)])
    tot = torch.stack(tot)
    gx, gy = torch.autograd.grad(tot.sum(), [px, py])
    scores, (pxg, pyg) = k2.mutual_information_recursion(px.detach(), py.detach(), boundary, return_grad=True)
    torch.testing.assert_close(scores, tot.detach(), rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(pxg, gx, rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(pyg, gy, rtol=1e-9, atol=1e-12)


def test_prune_range_invariants():
    am, lm, sym, Tl, Sl, boundary = _toy(B=3, T=40, S=12, V=9, seed=7)
    _, (pxg, pyg) = k2.rnnt_loss_smoothed(lm=lm, am=am, symbols=sym, termination_symbol=0, lm_only_scale=0.0,
                                          am_only_scale=0.0, boundary=boundary, reduction="none",
                                          return_grad=True)
    # occupation mass: exactly one blank per real frame
    for b in range(3):
        torch.testing.assert_close(pyg[b, :, :Tl[b]].sum(0), torch.ones(int(Tl[b])), rtol=1e-4, atol=1e-4)
    for variant in ("A", "B"):
        for R in (2, 3, 5):
            r = k2.get_rnnt_prune_ranges(pxg, pyg, boundary, R, variant=variant)
            s0 = r[:, :, 0]
            assert (s0[:, 0] == 0).all()
            d = s0[:, 1:] - s0[:, :-1]
            assert (d >= 0).all() and (d <= R - 1).all()
            for b in range(3):
                assert s0[b, int(Tl[b]) - 1] == max(int(Sl[b]) - R + 1, 0)


def test_monotonic_lower_bound():
    x = torch.tensor([[3, 1, 4, 1, 5, 9, 2, 6]])
    assert k2.monotonic_lower_bound(x).tolist() == [[1, 1, 1, 1, 2, 2, 2, 6]]


@pytest.mark.parametrize("name", ["stateless_predictor", "stateless_predictor_ctx2"])
def test_predictor_port_matches_reference_goldens(name):
    """oracle.reference_port.stateless_predictor_forward against the outputs of the reference's StatelessPredictor
    run verbatim (oracle/make_golden.py::run_predictor_case)."""
    from conftest import GOLDEN
    import os
    from oracle.make_golden import make_predictor_case
    gold = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    cfg, w, tokens, grad = make_predictor_case(name)
    wt = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in w.items()}
    state = torch.zeros(1, cfg["context_size"] - 1, dtype=torch.int32)
    out, out_state = port.stateless_predictor_forward(wt, torch.from_numpy(tokens), state, cfg["context_size"])
    (out * torch.from_numpy(grad)).sum().backward()
    np.testing.assert_allclose(out.detach().numpy(), gold["output"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(out_state.numpy(), gold["out_state"])
    for k, v in wt.items():
        np.testing.assert_allclose(v.grad.numpy(), gold["d" + k], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["greedy_outproj", "greedy_plain"])
def test_greedy_decode_port_matches_reference_goldens(name):
    """oracle.reference_port.rnnt_greedy_decode against the token sequences the reference's RnntGreedyDecoding
    produced (verbatim run, oracle/make_golden.py::run_greedy_case)."""
    import os
    from conftest import GOLDEN
    from oracle.make_golden import make_greedy_case
    gold = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    g, pcfg, pw, jw, enc = make_greedy_case(name)
    pw = {k: torch.from_numpy(v) for k, v in pw.items()}
    jw = {k: torch.from_numpy(v) for k, v in jw.items()}
    for i, e in enumerate(enc):
        got = port.rnnt_greedy_decode(pw, jw, g["joiner"], torch.from_numpy(e), pcfg["context_size"])
        assert got == gold[f"tokens_{i}"].tolist(), (name, i)


@pytest.mark.parametrize("name", ["greedy_outproj", "greedy_plain"])
def test_beam_decode_port_matches_reference_goldens(name):
    """oracle.reference_port.rnnt_beam_decode against the hypotheses (and their log-probabilities) the reference's
    RnntBeamDecoding produced (verbatim run, beam 4, top-k 4, oracle/make_golden.py::run_greedy_case)."""
    import os
    from conftest import GOLDEN
    from oracle.make_golden import BEAM_SIZE, BEAM_TOP_K, make_greedy_case
    gold = dict(np.load(os.path.join(GOLDEN, name.replace("greedy", "beam") + ".npz")))
    g, pcfg, pw, jw, enc = make_greedy_case(name)
    pw = {k: torch.from_numpy(v) for k, v in pw.items()}
    jw = {k: torch.from_numpy(v) for k, v in jw.items()}
    for i, e in enumerate(enc):
        toks, score = port.rnnt_beam_decode(pw, jw, g["joiner"], torch.from_numpy(e), pcfg["context_size"], BEAM_SIZE,
                                            BEAM_TOP_K)
        assert toks == gold[f"tokens_{i}"].tolist(), (name, i)
        assert abs(score - float(gold[f"score_{i}"])) < 1e-4 * abs(float(gold[f"score_{i}"]))
