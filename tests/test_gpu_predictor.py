"""GPU parity tests of the stateless predictor front end through the C ABI against the golden
vectors of the reference's StatelessPredictor (/root/reference/model/predictor/stateless_predictor.py, run verbatim by
oracle/make_golden.py) and the oracle port.  fp32 SIMT gather + FMA: 1e-5 on the output, 1e-4 on gradients."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import reference_port as port
from oracle.make_golden import PREDICTOR_CASES, make_predictor_case

pytestmark = pytest.mark.gpu


def _module(name, dev):
    from model.predictor.stateless_predictor import StatelessPredictor, StatelessPredictorConfig
    cfg, w, tokens, grad = make_predictor_case(name)
    pred = StatelessPredictor(StatelessPredictorConfig(num_symbols=cfg["num_symbols"], output_dim=cfg["output_dim"],
                                                       symbol_embedding_dim=cfg["symbol_embedding_dim"],
                                                       context_size=cfg["context_size"]))
    pred.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    return cfg, pred.to(dev), torch.from_numpy(tokens).to(dev), torch.from_numpy(grad).to(dev)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(PREDICTOR_CASES))
def test_stateless_predictor_matches_reference_goldens(name, mode, monkeypatch):
    monkeypatch.setenv("S2T_B200_JOINER_MODE", mode)
    dev = torch.device("cuda:0")
    gold = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    cfg, pred, tokens, grad = _module(name, dev)
    lengths = torch.full((cfg["B"],), cfg["U"], dtype=torch.int64, device=dev)
    out, out_len, out_state = pred(tokens, lengths, pred.init_state())
    (out * grad).sum().backward()
    torch.cuda.synchronize()
    assert tuple(out.shape) == gold["output"].shape and torch.equal(out_len, lengths)
    assert np.array_equal(out_state.cpu().numpy(), gold["out_state"])
    # fp32: FMA gather kernel + nn.Linear; bf16 mode: the Linear runs as the 3xF16 tensor-core GEMM forward
    # (fp32-level accuracy) with bf16 operands in its backward contractions
    otol, gtol = (1e-5, 1e-4) if mode == "fp32" else (1e-5, 1e-2)
    assert rel_err(out, torch.from_numpy(gold["output"])) <= otol
    for k, p in pred.named_parameters():
        assert rel_err(p.grad, torch.from_numpy(gold["d" + k])) <= gtol, k


def test_embed_conv_kernel_matches_torch_at_size():
    """The fused gather + depthwise conv at the predictor size of the zipformer yaml (E=512, C=5) on a c3-sized label
    batch, odd E (scalar path) and repeated tokens (the embedding gradient meets in one row)."""
    from speech2text_b200 import functional as F2
    dev = torch.device("cuda:0")
    for B, U, N, E, C in ((64, 100, 500, 512, 5), (3, 11, 7, 30, 2), (5, 40, 50, 66, 8)):
        g = torch.Generator().manual_seed(B + U)
        emb = torch.randn(N, E, generator=g)
        w = torch.randn(E, 1, C, generator=g) * 0.3
        tok = torch.randint(0, N, (B, U + C), generator=g)
        tok[0, :] = 1
        grad = torch.randn(B, U + 1, E, generator=g)
        e64, w64 = emb.double().requires_grad_(True), w.double().requires_grad_(True)
        ref = torch.nn.functional.conv1d(torch.nn.functional.embedding(tok, e64).transpose(1, 2), w64, groups=E).transpose(1, 2)
        (ref * grad.double()).sum().backward()
        ed, wd = emb.to(dev).requires_grad_(True), w.to(dev).requires_grad_(True)
        got = F2.predictor_embed_conv(tok.to(dev), ed, wd)
        (got * grad.to(dev)).sum().backward()
        torch.cuda.synchronize()
        assert rel_err(got, ref) <= 1e-5
        assert rel_err(ed.grad, e64.grad) <= 1e-4 and rel_err(wd.grad, w64.grad) <= 1e-4


def test_predictor_feeds_the_joiner_like_the_reference_task():
    """rnnt_task.py:464-471: predictor -> joiner on the same device; cpu tensors fail loudly."""
    from speech2text_b200._lib import S2TError
    cfg, pred, tokens, _ = _module("stateless_predictor", torch.device("cuda:0"))
    with pytest.raises(S2TError):
        pred.cpu()(tokens.cpu(), torch.full((cfg["B"],), cfg["U"]), pred.init_state())
    pred = pred.cuda()
    out_step, state = pred.streaming_step(tokens[:1, :1], pred.init_state().cuda())  # decoding path: plain torch
    w = {k: v.detach().cpu() for k, v in pred.state_dict().items()}
    ref, _ = port.stateless_predictor_forward(w, tokens[:1, :0].cpu(), pred.init_state(), cfg["context_size"])
    assert out_step.shape == (1, 1, cfg["output_dim"]) and state.shape == (1, cfg["context_size"] - 1)
