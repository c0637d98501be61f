"""B200 drop-in for the RNN-T greedy and beam searches of ``model/decoding.py``.

``RnntGreedyDecoding`` keeps the reference's constructor and ``decode(hidden_states) -> str`` contract
(/root/reference/model/decoding.py:196-271) and ``batch_search`` its signature (:32-48), so ``AsrMetric``
(model/utils.py:99-136) and the inference tasks run unchanged.  What changes: for a stateless predictor on a
CUDA device the whole batch is decoded by ONE kernel launch (``s2t_rnnt_greedy_decode``: one CTA per utterance walks
its lattice on the device) instead of a Python loop with ~10 launches and a ``.item()`` synchronisation per
lattice step and utterance.  Other predictors (LSTM) take the reference's step-by-step loop through
``streaming_step``.
"""
from __future__ import annotations

from typing import List

import torch

from . import _lib
from . import functional as F2
from ._lib import check, lib, ptr, stream


class RnntGreedyDecoding:
    """ Rnnt greedy decoding: tokenizer, predictor and joiner as in the reference. """

    def __init__(self, tokenizer, predictor, joiner, max_token_step=10):
        self._tokenizer = tokenizer
        self._predictor = predictor
        self._joiner = joiner
        # limit of token steps (lattice moves upward) within one time step (decoding.py:213-215)
        self._max_token_step = max_token_step
        assert hasattr(self._predictor, "streaming_step") and hasattr(self._joiner, "streaming_step"), (
            "Predictor and Joiner should impl streaming_step for decoding.")

    # ---- device-resident path ------------------------------------------------------------------
    def _stateless(self):
        p = getattr(self._predictor, "predictor", self._predictor)  # model.predictor.predictor.Predictor wraps it
        ok = all(hasattr(p, n) for n in ("_embedding", "_conv", "_output_linear", "_context_size"))
        return p if ok and 1 <= p._context_size <= 8 else None

    def _fast_path(self, hidden_states: torch.Tensor) -> bool:
        j = self._joiner
        return (hidden_states.is_cuda and self._stateless() is not None and hasattr(j, "_enc_proj")
                and hasattr(j, "_pre_proj") and hasattr(j, "_act_code"))

    @torch.no_grad()
    def batch_decode_tokens(self, hidden_states: torch.Tensor, lengths: torch.Tensor) -> List[List[int]]:
        """Token ids of every utterance of (B, T, D) encoder outputs with true lengths (B)."""
        if not self._fast_path(hidden_states):
            return [self._decode_tokens_stepwise(hidden_states[i:i + 1, :int(lengths[i])])
                    for i in range(hidden_states.shape[0])]
        dev = hidden_states.device
        with torch.cuda.device(dev):
            p, j = self._stateless(), self._joiner
            B, T, _ = hidden_states.shape
            f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
            # every frame is visited at least once: the encoder-side projection is one batched tensor-core GEMM
            am = F2.linear_tc(f32(hidden_states), f32(j._enc_proj.weight), f32(j._enc_proj.bias))
            V = am.shape[-1]
            emb, Wo, bo = f32(p._embedding.weight), f32(p._output_linear.weight), f32(p._output_linear.bias)
            E, C = emb.shape[1], p._context_size
            conv_w = f32(p._conv.weight).reshape(E, C)
            Wp, bp = f32(j._pre_proj.weight), f32(j._pre_proj.bias)
            W1, b1, W2, b2 = j._out_proj_params()
            I = 0
            if W1 is not None:
                W1, b1, W2, b2 = f32(W1), f32(b1), f32(W2), f32(b2)
                I = W1.shape[0]
            lens = lengths.to(device=dev, dtype=torch.int64).contiguous()
            max_out = T * (self._max_token_step + 1)
            tokens = torch.zeros((B, max_out), dtype=torch.int64, device=dev)
            n_tok = torch.zeros((B,), dtype=torch.int32, device=dev)
            check(lib().s2t_rnnt_greedy_decode(ptr(am), ptr(lens), ptr(emb), ptr(conv_w), ptr(Wo), ptr(bo), ptr(Wp), ptr(bp),
                                               ptr(W1), ptr(b1), ptr(W2), ptr(b2), B, T, V, emb.shape[0], E, C, Wo.shape[0], I,
                                               j._act_code, j.blank_token, self._max_token_step, max_out, ptr(tokens),
                                               ptr(n_tok), stream()))
            n_host = n_tok.cpu().tolist()  # the one host synchronisation of the whole batch
            tok_host = tokens[:, :max(n_host + [1])].cpu()
        return [tok_host[i, :n_host[i]].tolist() for i in range(B)]

    # ---- the reference's loop, for predictors the kernel does not cover ---------------------------
    @torch.no_grad()
    def _decode_tokens_stepwise(self, hidden_states: torch.Tensor) -> List[int]:
        assert hidden_states.shape[0] == 1, "Support BatchSize = 1 only."
        pred_state = self._predictor.init_state()
        curr_token = torch.zeros((1, 1), dtype=torch.long, device=hidden_states.device)  # <blank_id>
        pred_out, pred_state = self._predictor.streaming_step(curr_token, pred_state)
        t, n_step, out = 0, 0, []
        while t < hidden_states.shape[1]:
            logp = self._joiner.streaming_step(hidden_states[:, t:t + 1, :], pred_out)  # (1, V)
            tok = int(logp.argmax(dim=-1))
            if tok == 0 or n_step > self._max_token_step:
                t, n_step = t + 1, 0
                continue
            n_step += 1
            curr_token = torch.full((1, 1), tok, dtype=torch.long, device=hidden_states.device)
            pred_out, pred_state = self._predictor.streaming_step(curr_token, pred_state)
            out.append(tok)
        return out

    def decode(self, hidden_states: torch.Tensor) -> str:
        """ hidden_states (1, T, D) encoder output -> decoded text (decoding.py:225-271). """
        assert hidden_states.shape[0] == 1, "Support BatchSize = 1 only."
        lengths = torch.tensor([hidden_states.shape[1]])
        tokens = self.batch_decode_tokens(hidden_states, lengths)[0]
        return self._tokenizer.decode(torch.Tensor(tokens).long())


class RnntBeamDecoding(RnntGreedyDecoding):
    """ Beam search of Rnnt Asr system, token step restricted to 1 per frame (decoding.py:295-425): same constructor
        and ``decode(hidden_states) -> str``.  With a stateless predictor on a CUDA device the whole batch is searched
        by one launch of ``s2t_rnnt_beam_decode`` (one CTA per utterance); other predictors take the reference's loop. """

    def __init__(self, tokenizer, predictor, joiner, beam_size=4, cutoff_top_k=4):
        super().__init__(tokenizer, predictor, joiner, max_token_step=1)
        self._beam_size = beam_size
        self._cutoff_top_k = cutoff_top_k
        self.best_scores = None  # log-probabilities of the best hypotheses of the last batch

    def _fast_path(self, hidden_states: torch.Tensor) -> bool:
        return (super()._fast_path(hidden_states) and 1 <= self._beam_size <= 8 and 1 <= self._cutoff_top_k <= 8
                and self._joiner.blank_token == 0)  # the reference's beam search hard-codes <blank_id> = 0

    @torch.no_grad()
    def batch_decode_tokens(self, hidden_states: torch.Tensor, lengths: torch.Tensor) -> List[List[int]]:
        if not self._fast_path(hidden_states):
            out = [self._beam_stepwise(hidden_states[i:i + 1, :int(lengths[i])]) for i in range(hidden_states.shape[0])]
            self.best_scores = [s for _, s in out]
            return [t for t, _ in out]
        dev = hidden_states.device
        with torch.cuda.device(dev):
            p, j = self._stateless(), self._joiner
            B, T, _ = hidden_states.shape
            f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
            am = F2.linear_tc(f32(hidden_states), f32(j._enc_proj.weight), f32(j._enc_proj.bias))
            V = am.shape[-1]
            emb, Wo, bo = f32(p._embedding.weight), f32(p._output_linear.weight), f32(p._output_linear.bias)
            E, C = emb.shape[1], p._context_size
            conv_w = f32(p._conv.weight).reshape(E, C)
            Wp, bp = f32(j._pre_proj.weight), f32(j._pre_proj.bias)
            W1, b1, W2, b2 = j._out_proj_params()
            I = 0
            if W1 is not None:
                W1, b1, W2, b2 = f32(W1), f32(b1), f32(W2), f32(b2)
                I = W1.shape[0]
            lens = lengths.to(device=dev, dtype=torch.int64).contiguous()
            ws = torch.empty((lib().s2t_rnnt_beam_workspace_bytes(B, T, V, self._beam_size),), dtype=torch.uint8, device=dev)
            tokens = torch.zeros((B, T), dtype=torch.int64, device=dev)
            n_tok = torch.zeros((B,), dtype=torch.int32, device=dev)
            best = torch.zeros((B,), dtype=torch.float32, device=dev)
            check(lib().s2t_rnnt_beam_decode(ptr(am), ptr(lens), ptr(emb), ptr(conv_w), ptr(Wo), ptr(bo), ptr(Wp), ptr(bp),
                                             ptr(W1), ptr(b1), ptr(W2), ptr(b2), B, T, V, emb.shape[0], E, C, Wo.shape[0], I,
                                             j._act_code, 0, self._beam_size, self._cutoff_top_k, ptr(ws), ptr(tokens),
                                             ptr(n_tok), ptr(best), stream()))
            n_host = n_tok.cpu().tolist()  # the one host synchronisation of the whole batch
            tok_host = tokens[:, :max(n_host + [1])].cpu()
            self.best_scores = best.cpu().tolist()
        return [tok_host[i, :n_host[i]].tolist() for i in range(B)]

    @torch.no_grad()
    def _beam_stepwise(self, hidden_states: torch.Tensor):
        """The reference's loop (decoding.py:350-425) over this library's ``streaming_step``s."""
        dev = hidden_states.device
        pred_state = self._predictor.init_state()
        pred_out, pred_state = self._predictor.streaming_step(torch.zeros((1, 1), dtype=torch.long, device=dev), pred_state)
        beams = [dict(tokens=[], blank=True, score=0.0, state=pred_state, pred=pred_out)]
        for t in range(hidden_states.shape[1]):
            logp = self._joiner.streaming_step(hidden_states[:, t:t + 1, :], torch.cat([b["pred"] for b in beams], dim=0))
            new = []
            for i, b in enumerate(beams):
                for tok in torch.argsort(logp[i], descending=True).tolist()[:self._cutoff_top_k]:
                    sc = b["score"] + logp[i][tok]
                    if tok == 0:
                        new.append(dict(tokens=b["tokens"], blank=True, score=sc, state=b["state"], pred=b["pred"]))
                    else:
                        new.append(dict(tokens=b["tokens"] + [tok], blank=False, score=sc, state=b["state"], pred=None))
            beams = sorted(new, key=lambda x: x["score"], reverse=True)[:self._beam_size]
            for b in beams:
                if not b["blank"]:
                    b["pred"], b["state"] = self._predictor.streaming_step(
                        torch.tensor([[b["tokens"][-1]]], dtype=torch.long, device=dev), b["state"])
                    b["blank"] = True
        return beams[0]["tokens"], float(beams[0]["score"])


def batch_search(hidden_states: torch.Tensor, inputs_length: torch.Tensor, decode_session):
    """ Batch decoding (decoding.py:32-48): a list of decoded texts.  An ``RnntGreedyDecoding`` session decodes
        the whole batch at once; any other session is called utterance by utterance as in the reference. """
    if isinstance(decode_session, RnntGreedyDecoding):
        all_tokens = decode_session.batch_decode_tokens(hidden_states, inputs_length)
        return [decode_session._tokenizer.decode(torch.Tensor(t).long()) for t in all_tokens]
    results = []
    for entry_id in range(hidden_states.shape[0]):
        n = int(inputs_length[entry_id])
        results.append(decode_session.decode(hidden_states[entry_id:entry_id + 1, :n, :]))
    return results
