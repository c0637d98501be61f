"""ORACLE (test infrastructure, NOT product code): plain-torch CPU port of the
reference's host-side plumbing for the transducer-loss path, on top of
oracle/k2_shim.py.  It exists because /root/reference cannot travel to the GPU
box; tests/test_oracle.py checks it against the golden vectors that
oracle/make_golden.py minted by running the reference verbatim.

Follows, line by line in behaviour (not in text):
  * Joiner.forward / _do_rnnt_prune  /root/reference/model/joiner/joiner.py:74-182
  * PrunedRnntLoss.forward           /root/reference/model/loss/pruned_rnnt_loss.py:34-50
  * RnntLoss.forward                 /root/reference/model/loss/rnnt_loss.py:31-45
  * PrunedRnntTask.training_step mix /root/reference/task_factory/rnnt_task.py:469-499

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from oracle import k2_shim as k2


def _act(name: str):
    if name == "relu":
        return torch.relu
    if name == "tanh":
        return torch.tanh
    raise ValueError(f"Unsupported activation {name}")


def joiner_forward(w: Dict[str, torch.Tensor], cfg: dict, encoder_out, encoder_out_lengths,
                   predict_out, target_lengths, target: Optional[torch.Tensor] = None,
                   prune_variant: Optional[str] = None, ranges_override: Optional[torch.Tensor] = None,
                   exact: bool = False):
    """joiner.py:126-182.  ``w`` uses the reference's state_dict keys.

    ``ranges_override`` (test hook, not in the reference): use these prune ranges instead of the ones selected
    here.  The selection is an argmax over fp32 window sums; with the cumulative variant every window that holds
    (nearly) all of a frame's occupation mass ties up to rounding, so two correct implementations pick different --
    equivalent -- windows on some frames.  Forcing the ranges under test lets everything downstream of the
    (non-differentiable) selection be compared at full tolerance.

    ``exact`` (test hook): skip the reference's casts to float32 (joiner.py:99-102, pruned_rnnt_loss.py:40), so that
    fp64 inputs give the exact result of the reference's formulas -- the yardstick for "which fp32 run is closer"."""
    prune_range = cfg.get("prune_range", 5)
    act = _act(cfg.get("activation", "relu"))
    am = F.linear(encoder_out, w["_enc_proj.weight"], w["_enc_proj.bias"])
    lm = F.linear(predict_out, w["_pre_proj.weight"], w["_pre_proj.bias"])
    boundary = ranges = simple_loss = None
    if prune_range > 0:
        assert target.shape[0] == target_lengths.shape[0]
        boundary = torch.zeros((am.size(0), 4), dtype=torch.int64)
        boundary[:, 2] = target_lengths
        boundary[:, 3] = encoder_out_lengths
        assert target.dim() == 2
        simple_loss, (px_grad, py_grad) = k2.rnnt_loss_smoothed(
            lm=lm if exact else lm.to(dtype=torch.float32),  # "Pruned rnnt loss strictly required fp32" (joiner.py:99-102)
            am=am if exact else am.to(dtype=torch.float32),
            symbols=target,
            termination_symbol=0,
            lm_only_scale=cfg.get("lm_scale", 0.0),
            am_only_scale=cfg.get("am_scale", 0.0),
            boundary=boundary,
            reduction="mean",
            return_grad=True,
        )
        ranges = k2.get_rnnt_prune_ranges(px_grad=px_grad, py_grad=py_grad, boundary=boundary,
                                          s_range=prune_range, variant=prune_variant)
        if ranges_override is not None:
            assert ranges_override.shape == ranges.shape and ranges_override.dtype == ranges.dtype
            ranges = ranges_override
        am, lm = k2.do_rnnt_pruning(am=am, lm=lm, ranges=ranges)
    else:
        am = am.unsqueeze(2)
        lm = lm.unsqueeze(1)
    h = act(am + lm)
    if cfg.get("use_out_project", True):
        h = F.linear(h, w["_out_projection.0.weight"], w["_out_projection.0.bias"])
        h = F.linear(h, w["_out_projection.1.weight"], w["_out_projection.1.bias"])
    return h, boundary, ranges, simple_loss


def pruned_rnnt_loss(logits, targets, boundary, ranges, termination_symbol=0,
                     rnnt_type="regular", delay_penalty=0.0, reduction="mean", exact=False):
    """pruned_rnnt_loss.py:34-50."""
    return k2.rnnt_loss_pruned(
        logits=logits if exact else logits.to(torch.float32),  # pruned_rnnt_loss.py:40
        symbols=targets, ranges=ranges, termination_symbol=termination_symbol,
        boundary=boundary, rnnt_type=rnnt_type, delay_penalty=delay_penalty,
        reduction=reduction)


def rnnt_loss(logits, targets, logits_length, targets_length, blank_label=0, clamp=-1,
              reduction="mean"):
    """rnnt_loss.py:31-45 (torchaudio's compiled CPU kernel)."""
    import torchaudio
    return torchaudio.functional.rnnt_loss(
        logits, targets.to(torch.int32), logits_length.to(torch.int32),
        targets_length.to(torch.int32), blank=blank_label, clamp=clamp, reduction=reduction)


def ctc_loss(logits, targets, logits_length, targets_length, blank_label=0, reduction="mean", zero_infinity=True):
    """ctc_loss.py:35-41 (torch's CPU kernel, the reference's own back end; fp64 logits give the fp64 oracle)."""
    log_probs = F.log_softmax(logits, dim=-1).transpose(0, 1)
    if log_probs.dtype != torch.float64:
        log_probs = log_probs.to(dtype=torch.float32)
    return F.ctc_loss(log_probs, targets, logits_length, targets_length, blank=blank_label, reduction=reduction,
                      zero_infinity=zero_infinity)


def stateless_predictor_forward(w: Dict[str, torch.Tensor], tokens, state, context_size: int):
    """StatelessPredictor.forward, /root/reference/model/predictor/stateless_predictor.py:74-99: left-pad with
    <blank> = 0, prepend the state, embedding -> depthwise Conv1d(kernel = context_size) -> Linear.
    ``w`` uses the reference's state_dict keys.  Returns (output (B, 1 + U, D), out_state)."""
    bs = tokens.shape[0]
    state = state.repeat(bs, 1)
    padded = F.pad(tokens.float(), (1, 0, 0, 0), value=0.0).to(torch.int32)
    ctxed = torch.concat([state, padded], dim=1)
    out_state = ctxed[:, ctxed.shape[1] - context_size:]
    embs = F.embedding(ctxed, w["_embedding.weight"]).transpose(1, 2)
    conv = F.conv1d(embs, w["_conv.weight"], groups=w["_conv.weight"].shape[0]).transpose(1, 2)
    return F.linear(conv, w["_output_linear.weight"], w["_output_linear.bias"]), out_state


def rnnt_greedy_decode(pw: Dict[str, torch.Tensor], jw: Dict[str, torch.Tensor], jcfg: dict, hidden_states,
                       context_size: int, max_token_step: int = 10):
    """RnntGreedyDecoding.decode, /root/reference/model/decoding.py:225-271, for ONE utterance (1, T, D) with a
    stateless predictor (streaming_step, stateless_predictor.py:107-124) and the joiner's streaming_step
    (joiner.py:184-207).  ``pw`` / ``jw``: predictor / joiner weights under the reference's state_dict keys."""
    act = _act(jcfg.get("activation", "relu"))

    def pred_step(token, state):  # token (1,1), state (1, C-1)
        ctxed = torch.concat([state, token], dim=1)
        out_state = ctxed[:, ctxed.shape[1] - context_size + 1:]
        embs = F.embedding(ctxed, pw["_embedding.weight"]).transpose(1, 2)
        conv = F.conv1d(embs, pw["_conv.weight"], groups=pw["_conv.weight"].shape[0]).transpose(1, 2)
        return F.linear(conv, pw["_output_linear.weight"], pw["_output_linear.bias"]), out_state

    def joiner_step(enc, pred):  # (1,1,D), (1,1,D) -> (1, V) log-probs
        a = F.linear(enc, jw["_enc_proj.weight"], jw["_enc_proj.bias"]).unsqueeze(2)
        l = F.linear(pred, jw["_pre_proj.weight"], jw["_pre_proj.bias"]).unsqueeze(1)
        h = act(a + l)
        if jcfg.get("use_out_project", True):
            h = F.linear(h, jw["_out_projection.0.weight"], jw["_out_projection.0.bias"])
            h = F.linear(h, jw["_out_projection.1.weight"], jw["_out_projection.1.bias"])
        return torch.log_softmax(h, dim=-1).squeeze(1).squeeze(1)

    state = torch.zeros(1, context_size - 1, dtype=torch.int32)
    token = torch.zeros(1, 1, dtype=torch.long)
    pred_out, state = pred_step(token, state)
    t, n_step, out = 0, 0, []
    while t < hidden_states.shape[1]:
        tok = joiner_step(hidden_states[:, t:t + 1, :], pred_out).argmax(dim=-1).unsqueeze(0)
        if torch.allclose(tok, torch.zeros(1, 1).long()) or n_step > max_token_step:
            t, n_step = t + 1, 0
            continue
        n_step += 1
        token = tok
        pred_out, state = pred_step(token, state)
        out.append(tok.item())
    return out


def rnnt_beam_decode(pw: Dict[str, torch.Tensor], jw: Dict[str, torch.Tensor], jcfg: dict, hidden_states,
                     context_size: int, beam_size: int = 4, cutoff_top_k: int = 4):
    """RnntBeamDecoding.decode, /root/reference/model/decoding.py:295-425, for ONE utterance (1, T, D) with a stateless
    predictor: at most one token per frame; every beam proposes its ``cutoff_top_k`` best classes; a blank (class 0)
    keeps the hypothesis, any other class extends it; the ``beam_size`` best candidates by accumulated log-probability
    survive (Python's stable sort: ties in order of creation); hypotheses are not merged.  Returns (tokens, score)."""
    act = _act(jcfg.get("activation", "relu"))

    def pred_step(token, state):
        ctxed = torch.concat([state, token], dim=1)
        out_state = ctxed[:, ctxed.shape[1] - context_size + 1:]
        embs = F.embedding(ctxed, pw["_embedding.weight"]).transpose(1, 2)
        conv = F.conv1d(embs, pw["_conv.weight"], groups=pw["_conv.weight"].shape[0]).transpose(1, 2)
        return F.linear(conv, pw["_output_linear.weight"], pw["_output_linear.bias"]), out_state

    def joiner_step(enc, pred):  # (1,1,D), (n,1,D) -> (n, V) log-probs
        a = F.linear(enc, jw["_enc_proj.weight"], jw["_enc_proj.bias"]).unsqueeze(2)
        l = F.linear(pred, jw["_pre_proj.weight"], jw["_pre_proj.bias"]).unsqueeze(1)
        h = act(a + l)
        if jcfg.get("use_out_project", True):
            h = F.linear(h, jw["_out_projection.0.weight"], jw["_out_projection.0.bias"])
            h = F.linear(h, jw["_out_projection.1.weight"], jw["_out_projection.1.bias"])
        return torch.log_softmax(h, dim=-1).squeeze(1).squeeze(1)

    state = torch.zeros(1, context_size - 1, dtype=torch.int32)
    pred_out, state = pred_step(torch.zeros(1, 1, dtype=torch.long), state)
    beams = [dict(tokens=[], blank=True, score=0.0, state=state, pred=pred_out)]
    for t in range(hidden_states.shape[1]):
        logp = joiner_step(hidden_states[:, t:t + 1, :], torch.cat([b["pred"] for b in beams], dim=0))
        new = []
        for i, b in enumerate(beams):
            for tok in torch.argsort(logp[i], descending=True).tolist()[:cutoff_top_k]:
                sc = b["score"] + logp[i][tok]
                if tok == 0:
                    new.append(dict(tokens=b["tokens"], blank=True, score=sc, state=b["state"], pred=b["pred"]))
                else:
                    new.append(dict(tokens=b["tokens"] + [tok], blank=False, score=sc, state=b["state"], pred=None))
        beams = sorted(new, key=lambda x: x["score"], reverse=True)[:beam_size]
        for b in beams:
            if not b["blank"]:
                b["pred"], b["state"] = pred_step(torch.tensor([[b["tokens"][-1]]]).long(), b["state"])
                b["blank"] = True
    return beams[0]["tokens"], float(beams[0]["score"])


def training_step_loss(w, spec: dict, case: dict, dtype=torch.float32, prune_variant=None, ranges_override=None,
                       exact=False):
    """One fwd+bwd of the hot path exactly as rnnt_task.py:469-514 strings it
    together.  Returns a dict of losses, ranges and gradients."""
    cfg = spec["joiner"]
    w = {k: torch.as_tensor(v).detach().to(dtype).clone().requires_grad_(True) for k, v in w.items()}
    # detached copies: the caller's tensors must not become autograd leaves of this run
    enc = torch.as_tensor(case["encoder_out"]).detach().to(dtype).clone().requires_grad_(True)
    pred = torch.as_tensor(case["predict_out"]).detach().to(dtype).clone().requires_grad_(True)
    enc_len = torch.as_tensor(case["encoder_out_lengths"])
    tgt_len = torch.as_tensor(case["target_lengths"])
    tgt = torch.as_tensor(case["target"])
    out = {}
    if cfg.get("prune_range", 5) > 0:
        logits, boundary, ranges, simple = joiner_forward(w, cfg, enc, enc_len, pred, tgt_len,
                                                          tgt, prune_variant, ranges_override, exact)
        pruned = pruned_rnnt_loss(logits, tgt, boundary, ranges, exact=exact, **spec.get("loss", {}))
        total = (spec["simple_loss_scale"] * simple + spec["pruned_loss_scale"] * pruned).mean()
        out.update(simple_loss=simple.detach(), pruned_loss=pruned.detach(),
                   boundary=boundary, ranges=ranges, logits=logits.detach())
    else:
        logits, _, _, _ = joiner_forward(w, cfg, enc, enc_len, pred, tgt_len)
        loss = rnnt_loss(logits.float(), tgt, enc_len, tgt_len, **spec.get("loss", {}))
        total = loss.mean()
        out.update(rnnt_loss=loss.detach(), logits=logits.detach())
    total.backward()
    out["total_loss"] = total.detach()
    out["d_encoder_out"] = enc.grad
    out["d_predict_out"] = pred.grad
    for k, v in w.items():
        out["d" + k] = v.grad if v.grad is not None else torch.zeros_like(v)
    return out
