"""ORACLE tooling (test infrastructure, NOT product code).

Mints the golden vectors under tests/golden/ by running the REFERENCE's own
modules, verbatim, in this build container:

    /root/reference/model/joiner/joiner.py        (Joiner, JoinerConfig)
    /root/reference/model/loss/pruned_rnnt_loss.py (PrunedRnntLoss)
    /root/reference/model/loss/rnnt_loss.py        (RnntLoss -> torchaudio)

``k2`` (un-vendored, not installable here) is satisfied by oracle/k2_shim.py
installed as ``sys.modules["k2"]``; ``onnx`` (only used by the export methods)
by an empty stub module.  /root/reference does not exist on the GPU box, so the
outputs are committed as small .npz fixtures and this script is the recipe.

Inputs and weights are drawn from numpy RandomState streams (stable across
torch versions) by ``oracle.cases.make_case`` so tests can regenerate them
bit-exactly without the reference.

Usage (build container only):  python -m oracle.make_golden
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    """The three reference modules, loaded from their files under /root/reference (the reference's ``model`` is a
    namespace package, so with this repo's import-path mirror on sys.path ``import model.joiner.joiner`` would
    resolve to the mirror; none of the three imports anything from ``model.*``)."""
    import importlib.util
    from oracle import k2_shim
    sys.modules["k2"] = k2_shim
    sys.modules.setdefault("onnx", types.ModuleType("onnx"))

    def load(rel):
        path = os.path.join(REFERENCE, rel)
        spec = importlib.util.spec_from_file_location("_reference_" + os.path.basename(rel)[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert mod.__file__.startswith(REFERENCE), mod.__file__
        return mod

    return (load("model/joiner/joiner.py"), load("model/loss/pruned_rnnt_loss.py"), load("model/loss/rnnt_loss.py"))


def summarize(t: torch.Tensor, max_elems: int = 16384):
    """Full tensor if small, else a strided sub-sample + sum + L2 norm."""
    flat = t.detach().reshape(-1).to(torch.float64 if t.is_floating_point() else t.dtype)
    n = flat.numel()
    stride = max(1, -(-n // max_elems))
    sub = flat[::stride]
    out = {
        "stride": np.int64(stride),
        "numel": np.int64(n),
        "sub": sub.to(t.dtype).numpy(),
    }
    if t.is_floating_point():
        out["sum"] = np.float64(flat.sum().item())
        out["norm"] = np.float64(flat.norm().item())
    return out


def run_case(name, joiner_mod, pruned_mod, rnnt_mod, dtype=torch.float32, variant="B"):
    from oracle.cases import CASES, make_case
    from oracle import k2_shim
    k2_shim.PRUNE_RANGES_VARIANT = variant  # the reference calls k2.get_rnnt_prune_ranges without a variant
    spec = CASES[name]
    case = make_case(name)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    joiner = joiner_mod.Joiner(joiner_mod.JoinerConfig(**spec["joiner"]))
    joiner.load_state_dict({k: torch.from_numpy(v) for k, v in case["weights"].items()})
    joiner = joiner.to(dtype)
    enc = torch.from_numpy(case["encoder_out"]).to(dtype).requires_grad_(True)
    pred = torch.from_numpy(case["predict_out"]).to(dtype).requires_grad_(True)
    enc_len = torch.from_numpy(case["encoder_out_lengths"])
    tgt_len = torch.from_numpy(case["target_lengths"])
    tgt = torch.from_numpy(case["target"])

    res = {}
    if spec["joiner"].get("prune_range", 5) > 0:
        loss_cfg = dict(spec.get("loss", {}))
        loss_mod = pruned_mod.PrunedRnntLoss(pruned_mod.PrunedRnntLossConfig(**loss_cfg))
        logits, boundary, ranges, simple = joiner(enc, enc_len, pred, tgt_len, tgt)
        pruned = loss_mod(logits=logits, targets=tgt, logits_length=enc_len,
                          targets_length=tgt_len, boundary=boundary, ranges=ranges)
        # rnnt_task.py:496-499 / 514
        total = (spec["simple_loss_scale"] * simple + spec["pruned_loss_scale"] * pruned).mean()
        total.backward()
        res["simple_loss"] = simple.detach().to(torch.float64).numpy()
        res["pruned_loss"] = pruned.detach().to(torch.float64).numpy()
        res["boundary"] = boundary.numpy()
        res["ranges"] = ranges.numpy()
        res["logits_shape"] = np.array(logits.shape, dtype=np.int64)
        for k, v in summarize(logits).items():
            res[f"logits.{k}"] = v
    else:
        loss_mod = rnnt_mod.RnntLoss(rnnt_mod.RnntLossConfig(**spec.get("loss", {})))
        logits, boundary, ranges, simple = joiner(enc, enc_len, pred, tgt_len)
        assert boundary is None and ranges is None and simple is None
        loss = loss_mod(logits=logits.float(), targets=tgt, logits_length=enc_len,
                        targets_length=tgt_len)
        total = loss.mean()
        total.backward()
        res["rnnt_loss"] = loss.detach().to(torch.float64).numpy()
        res["logits_shape"] = np.array(logits.shape, dtype=np.int64)
    res["total_loss"] = total.detach().to(torch.float64).numpy()
    grads = {"d_encoder_out": enc.grad, "d_predict_out": pred.grad}
    for pname, p in joiner.named_parameters():
        grads["d" + pname] = p.grad if p.grad is not None else torch.zeros_like(p)
    for gname, g in grads.items():
        for k, v in summarize(g).items():
            res[f"{gname}.{k}"] = v
    return res


PREDICTOR_CASES = {
    # zipformer_stateless_pruned_rnnt.yaml:74-80 shape (context 5) at a small size, and an odd one
    "stateless_predictor": dict(num_symbols=128, output_dim=96, symbol_embedding_dim=64, context_size=5, B=4, U=23),
    "stateless_predictor_ctx2": dict(num_symbols=37, output_dim=20, symbol_embedding_dim=18, context_size=2, B=3, U=9),
}


def make_predictor_case(name):
    """Seeded weights (reference state_dict keys), tokens and an upstream gradient for a predictor case."""
    cfg = PREDICTOR_CASES[name]
    rs = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    N, D, E, C = cfg["num_symbols"], cfg["output_dim"], cfg["symbol_embedding_dim"], cfg["context_size"]
    w = {
        "_embedding.weight": rs.standard_normal((N, E)).astype(np.float32),
        "_conv.weight": (rs.uniform(-1, 1, (E, 1, C)) / np.sqrt(C)).astype(np.float32),
        "_output_linear.weight": (rs.uniform(-1, 1, (D, E)) / np.sqrt(E)).astype(np.float32),
        "_output_linear.bias": (rs.uniform(-1, 1, (D,)) / np.sqrt(E)).astype(np.float32),
    }
    tokens = rs.randint(1, N, (cfg["B"], cfg["U"])).astype(np.int64)
    grad = rs.standard_normal((cfg["B"], cfg["U"] + 1, D)).astype(np.float32)
    return cfg, w, tokens, grad


def run_predictor_case(name):
    """The reference's StatelessPredictor, verbatim, on a seeded case: output, out_state and all gradients."""
    import importlib.util
    sys.modules.setdefault("onnx", types.ModuleType("onnx"))
    spec = importlib.util.spec_from_file_location("_reference_stateless_predictor",
                                                  os.path.join(REFERENCE, "model/predictor/stateless_predictor.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg, w, tokens, grad = make_predictor_case(name)
    pred = mod.StatelessPredictor(mod.StatelessPredictorConfig(
        num_symbols=cfg["num_symbols"], output_dim=cfg["output_dim"],
        symbol_embedding_dim=cfg["symbol_embedding_dim"], context_size=cfg["context_size"]))
    pred.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    lengths = torch.full((cfg["B"],), cfg["U"], dtype=torch.int64)
    out, out_len, out_state = pred(torch.from_numpy(tokens), lengths, pred.init_state())
    (out * torch.from_numpy(grad)).sum().backward()
    res = {"output": out.detach().numpy(), "out_state": out_state.numpy(), "lengths": out_len.numpy()}
    for k, p in pred.named_parameters():
        res["d" + k] = p.grad.numpy()
    return res


BEAM_SIZE, BEAM_TOP_K = 4, 4  # RnntBeamDecoding's defaults (model/decoding.py:310-311)

GREEDY_CASES = {
    # (predictor case, joiner config, T per utterance, blank bias): small models with a blank-leaning joiner bias so
    # that the search emits a realistic mix of blanks and symbols
    "greedy_outproj": dict(predictor="stateless_predictor", T=[40, 27, 33], blank_bias=12.0,
                           joiner=dict(input_dim=96, output_dim=128, inner_dim=32, activation="tanh", prune_range=5)),
    "greedy_plain": dict(predictor="stateless_predictor_ctx2", T=[25, 31], blank_bias=7.0,
                         joiner=dict(input_dim=20, output_dim=37, activation="relu", prune_range=5, use_out_project=False)),
}


def make_greedy_case(name):
    from oracle.cases import make_weights
    g = GREEDY_CASES[name]
    pcfg, pw, _, _ = make_predictor_case(g["predictor"])
    rs = np.random.RandomState(sum(map(ord, name)))
    jw = make_weights(g["joiner"], rs)
    for k in jw:  # larger weights: peaked distributions, no near-ties between the two arithmetic orders
        jw[k] = (jw[k] * 4.0).astype(np.float32)
    last = "_out_projection.1.bias" if g["joiner"].get("use_out_project", True) else "_pre_proj.bias"
    jw[last][0] += g["blank_bias"]
    enc = [rs.standard_normal((1, t, g["joiner"]["input_dim"])).astype(np.float32) for t in g["T"]]
    return g, pcfg, pw, jw, enc


def run_greedy_case(name):
    """The reference's RnntGreedyDecoding.decode (model/decoding.py:225-271) and RnntBeamDecoding.decode (:350-425),
    verbatim, with the reference's Joiner and
    StatelessPredictor, on seeded weights and encoder outputs.  Its imports of the tokenizer / factory modules are
    satisfied by stubs (they are type annotations there); the tokenizer stub returns the token ids."""
    import importlib.util
    from oracle import k2_shim
    sys.modules["k2"] = k2_shim
    for stub in ("onnx", "glog"):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    sys.modules["glog"].info = lambda *a, **k: None
    saved = {k: sys.modules.get(k) for k in ("dataset", "dataset.utils", "model.predictor.predictor", "model.joiner.joiner",
                                             "torchaudio.models.decoder")}

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    joiner_mod = load("_reference_joiner", "model/joiner/joiner.py")
    pred_mod = load("_reference_stateless_predictor", "model/predictor/stateless_predictor.py")
    ds, dsu = types.ModuleType("dataset"), types.ModuleType("dataset.utils")
    dsu.Tokenizer = object
    pp = types.ModuleType("model.predictor.predictor")
    pp.Predictor = object
    jj = types.ModuleType("model.joiner.joiner")
    jj.Joiner = joiner_mod.Joiner
    tad = types.ModuleType("torchaudio.models.decoder")  # the CTC lexicon decoder needs flashlight-text (not installed)
    tad.ctc_decoder = None
    sys.modules.update({"dataset": ds, "dataset.utils": dsu, "model.predictor.predictor": pp, "model.joiner.joiner": jj,
                        "torchaudio.models.decoder": tad})
    try:
        dec_mod = load("_reference_decoding", "model/decoding.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    g, pcfg, pw, jw, enc = make_greedy_case(name)
    pred = pred_mod.StatelessPredictor(pred_mod.StatelessPredictorConfig(
        num_symbols=pcfg["num_symbols"], output_dim=pcfg["output_dim"],
        symbol_embedding_dim=pcfg["symbol_embedding_dim"], context_size=pcfg["context_size"]))
    pred.load_state_dict({k: torch.from_numpy(v) for k, v in pw.items()})
    joiner = joiner_mod.Joiner(joiner_mod.JoinerConfig(**g["joiner"]))
    joiner.load_state_dict({k: torch.from_numpy(v) for k, v in jw.items()})

    class Tok:
        def decode(self, t):
            return [int(x) for x in t.tolist()]

    session = dec_mod.RnntGreedyDecoding(Tok(), pred.eval(), joiner.eval())
    beam_session = dec_mod.RnntBeamDecoding(Tok(), pred.eval(), joiner.eval(), beam_size=BEAM_SIZE, cutoff_top_k=BEAM_TOP_K)
    res, beam = {}, {}
    for i, e in enumerate(enc):
        res[f"tokens_{i}"] = np.asarray(session.decode(torch.from_numpy(e)), dtype=np.int64)
        beam[f"tokens_{i}"] = np.asarray(beam_session.decode(torch.from_numpy(e)), dtype=np.int64)
        beam[f"score_{i}"] = np.asarray(float(beam_session._decoding_state.best_beam.score), dtype=np.float64)
    return res, beam


def main():
    from oracle.cases import CASES
    os.makedirs(OUT, exist_ok=True)
    for name in PREDICTOR_CASES:
        res = run_predictor_case(name)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **res)
        print(f"{name}: output {res['output'].shape} -> {os.path.getsize(path) / 1024:.0f} KiB")
    for name in GREEDY_CASES:
        res, beam = run_greedy_case(name)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **res)
        np.savez_compressed(os.path.join(OUT, name.replace("greedy", "beam") + ".npz"), **beam)
        print(f"{name}: " + ", ".join(f"{k}: {len(v)} tokens" for k, v in res.items()))
        print(f"  beam: " + ", ".join(f"{k}: {v.tolist() if v.ndim == 0 else len(v)}" for k, v in beam.items()))
    joiner_mod, pruned_mod, rnnt_mod = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    for name, spec in CASES.items():
        pruned = spec["joiner"].get("prune_range", 5) > 0
        # both published variants of k2.get_rnnt_prune_ranges for every pruned case
        for variant in (("A", "B") if pruned else (None,)):
            for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
                if dt == torch.float64 and not pruned:
                    continue  # torchaudio's rnnt_loss has no fp64 kernel
                res = run_case(name, joiner_mod, pruned_mod, rnnt_mod, dt, variant or "B")
                stem = f"{name}.{variant}.{tag}" if variant else f"{name}.{tag}"
                path = os.path.join(OUT, stem + ".npz")
                np.savez_compressed(path, **res)
                keys = [k for k in ("simple_loss", "pruned_loss", "rnnt_loss") if k in res]
                print(f"{stem}: " + ", ".join(f"{k}={res[k]}" for k in keys),
                      f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
