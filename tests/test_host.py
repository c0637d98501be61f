"""CPU tests of the host-side mirror of the reference interface (no GPU, no kernels)."""
import dataclasses

import pytest
import torch

from speech2text_b200 import Joiner, JoinerConfig, LazyJoinerLogits, Loss
from speech2text_b200._lib import S2TError
from speech2text_b200.distributed import FlatGradBucket, shard_bounds
from speech2text_b200.loss.pruned_rnnt_loss import PrunedRnntLoss, PrunedRnntLossConfig
from speech2text_b200.loss.rnnt_loss import RnntLossConfig


def test_joiner_config_fields_and_defaults_match_reference():
    # /root/reference/model/joiner/joiner.py:16-26
    fields = [(f.name, f.default) for f in dataclasses.fields(JoinerConfig)]
    assert fields == [("input_dim", dataclasses.MISSING), ("output_dim", dataclasses.MISSING), ("inner_dim", 256),
                      ("activation", "relu"), ("prune_range", 5), ("lm_scale", 0.0), ("am_scale", 0.0),
                      ("use_out_project", True)]
    # yaml blocks are splatted into the dataclass: unknown keys must raise, as in the reference
    with pytest.raises(TypeError):
        JoinerConfig(input_dim=4, output_dim=4, fused=True)
    assert [(f.name, f.default) for f in dataclasses.fields(PrunedRnntLossConfig)] == [
        ("termination_symbol", 0), ("rnnt_type", "regular"), ("delay_penalty", 0.0), ("reduction", "mean")]
    assert [(f.name, f.default) for f in dataclasses.fields(RnntLossConfig)] == [
        ("blank_label", 0), ("clamp", -1), ("reduction", "mean")]


def test_state_dict_keys_match_reference_checkpoints():
    # checkpoints are loaded by key: joiner.py:41-55
    j = Joiner(JoinerConfig(input_dim=8, output_dim=6, inner_dim=4))
    assert sorted(j.state_dict()) == sorted([
        "_enc_proj.weight", "_enc_proj.bias", "_pre_proj.weight", "_pre_proj.bias", "_out_projection.0.weight",
        "_out_projection.0.bias", "_out_projection.1.weight", "_out_projection.1.bias"])
    j2 = Joiner(JoinerConfig(input_dim=8, output_dim=6, use_out_project=False))
    assert sorted(j2.state_dict()) == sorted(["_enc_proj.weight", "_enc_proj.bias", "_pre_proj.weight",
                                               "_pre_proj.bias"])
    assert j.prune_range == 5 and j.blank_token == 0


def test_error_conventions():
    with pytest.raises(ValueError):
        Joiner(JoinerConfig(input_dim=4, output_dim=4, activation="gelu"))  # joiner.py:49
    with pytest.raises(ValueError):
        Loss({"model": "NoSuchLoss", "config": {}})  # loss.py:41
    with pytest.raises(NotImplementedError):
        PrunedRnntLoss(PrunedRnntLossConfig(rnnt_type="modified"))
    with pytest.raises(ValueError):
        PrunedRnntLoss(PrunedRnntLossConfig(reduction="median"))


def test_forward_on_cpu_fails_loudly_no_fallback():
    j = Joiner(JoinerConfig(input_dim=8, output_dim=6, use_out_project=False))
    with pytest.raises(S2TError):
        j(torch.rand(2, 5, 8), torch.tensor([5, 4]), torch.rand(2, 3, 8), torch.tensor([2, 1]),
          torch.randint(1, 6, (2, 2)))


def test_streaming_step_is_scriptable_and_matches_plain_torch():
    # joiner_test.py:73-86 scripts the joiner and calls streaming_step with a beam of 4
    torch.manual_seed(0)
    for cfg in (dict(input_dim=16, output_dim=12, prune_range=5, use_out_project=False),
                dict(input_dim=16, output_dim=12, inner_dim=8, activation="tanh")):
        j = Joiner(JoinerConfig(**cfg)).eval()
        ts = torch.jit.script(j)
        enc, pred = torch.rand(1, 1, 16), torch.rand(4, 1, 16)
        out = ts.streaming_step(enc, pred)
        assert out.shape == (4, 12)
        am = torch.nn.functional.linear(enc, j._enc_proj.weight, j._enc_proj.bias)
        lm = torch.nn.functional.linear(pred, j._pre_proj.weight, j._pre_proj.bias)
        ref = j._out_projection(j._activation(am.unsqueeze(2) + lm.unsqueeze(1)))
        ref = torch.log_softmax(ref, dim=-1).squeeze(1).squeeze(1)
        torch.testing.assert_close(out, ref)
        out2 = ts.sherpa_onnx_streaming_step(torch.rand(3, 16), torch.rand(3, 16))
        assert out2.shape == (3, 12)


def test_lazy_logits_handle_quacks_like_the_tensor():
    am, lm = torch.zeros(2, 7, 5), torch.zeros(2, 4, 5)
    ranges = torch.zeros(2, 7, 3, dtype=torch.int64)
    h = LazyJoinerLogits(am, lm, None, None, None, None, ranges, 0, 0)
    assert tuple(h.shape) == (2, 7, 3, 5) and h.size(2) == 3 and h.dim() == 4
    assert h.to(torch.float32) is h and h.float() is h and h.dtype == torch.float32
    h2 = LazyJoinerLogits(am, lm, None, None, None, None, None, 0, 0)
    assert tuple(h2.shape) == (2, 7, 4, 5)  # unpruned: R = U + 1


def test_loss_factory_routes_like_the_reference():
    assert type(Loss({"model": "Pruned_Rnnt", "config": {"termination_symbol": 0, "reduction": "mean"}}).loss
                ).__name__ == "PrunedRnntLoss"
    assert type(Loss({"model": "Rnnt", "config": {"blank_label": 0, "clamp": -1, "reduction": "mean"}}).loss
                ).__name__ == "RnntLoss"
    assert type(Loss({"model": "CTC", "config": {}}).loss).__name__ == "CtcLoss"


def test_model_package_mirror_resolves_to_this_repo():
    import model.joiner.joiner as mj
    import model.loss.loss as ml
    import model.loss.pruned_rnnt_loss as mp
    import model.loss.rnnt_loss as mr
    assert mj.Joiner is Joiner and ml.Loss is Loss
    assert mp.PrunedRnntLoss is PrunedRnntLoss and mr.RnntLossConfig is RnntLossConfig


def test_shard_bounds_partition_utterances():
    for n, w in ((64, 8), (10, 4), (3, 8), (0, 2)):
        got = [i for r in range(w) for i in shard_bounds(n, r, w)]
        assert got == list(range(n))
        sizes = [len(shard_bounds(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_flat_grad_bucket_views():
    lin = torch.nn.Linear(3, 2)
    bucket = FlatGradBucket(lin.parameters())
    assert bucket.flat.numel() == 8
    lin(torch.ones(4, 3)).sum().backward()
    # autograd accumulated into the views, i.e. into the flat buffer
    torch.testing.assert_close(bucket.flat[:6].view(2, 3), torch.full((2, 3), 4.0))
    torch.testing.assert_close(bucket.flat[6:], torch.full((2,), 4.0))
    assert lin.weight.grad.data_ptr() == bucket.flat.data_ptr()
    bucket.zero()
    assert float(lin.weight.grad.abs().sum()) == 0.0


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/model"), reason="reference checkout not present")
def test_off_path_losses_resolve_to_the_reference_through_the_mirror():
    # zero-edit drop-in: PYTHONPATH=<this repo>:<reference>; ssl/nnlm/cif tasks import
    # model.loss.loss.Loss and ask for MaskedCELoss / MaskedKLDiv / MaeLoss, which live only in the reference
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import model.loss.cross_entropy as m, model.loss.loss as l, model.joiner.joiner as j\n"
            "assert m.__file__.startswith('/root/reference'), m.__file__\n"
            "assert l.__file__.startswith(%r) and j.__file__.startswith(%r)\n"
            "from model.loss.loss import Loss\n"
            "for name in ('MaskedCELoss', 'MaskedKLDiv', 'MaeLoss'):\n"
            "    assert type(Loss({'model': name, 'config': {}}).loss).__module__.startswith('model.loss.'), name\n"
            "print('ok')\n" % (root, root))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + "/root/reference")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_bound_sinks_are_claimed_only_when_overwriting_equals_accumulating():
    # zero_grad(set_to_none=True) and a second backward before zero() must fall back to autograd
    from speech2text_b200.functional import claim_grad_sinks, grad_sink
    lin = torch.nn.Linear(3, 2)
    params = (lin.weight, lin.bias)
    bucket = FlatGradBucket(lin.parameters())
    assert grad_sink(lin.weight) is None and claim_grad_sinks(params) is None  # not bound
    bucket.bind()
    assert grad_sink(lin.weight).data_ptr() == bucket.flat.data_ptr()
    first = claim_grad_sinks(params)
    assert first is not None and first[0].data_ptr() == lin.weight.grad.data_ptr()
    assert claim_grad_sinks(params) is None          # second backward before zero(): accumulate through autograd
    bucket.zero()
    assert claim_grad_sinks(params) is not None
    bucket.zero()
    lin.zero_grad(set_to_none=True)                    # p.grad no longer aliases the flat buffer
    assert claim_grad_sinks(params) is None
    bucket.attach()                                    # re-alias (and re-bind)
    assert claim_grad_sinks(params) is not None
    bucket.unbind()
    assert claim_grad_sinks(params) is None
