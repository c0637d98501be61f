"""GPU parity tests of the device-resident greedy RNN-T search through the C ABI against the
token sequences of the reference's RnntGreedyDecoding (/root/reference/model/decoding.py:225-271, run verbatim by
oracle/make_golden.py) and the oracle port.  Index output: the token sequences must be identical."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import reference_port as port
from oracle.make_golden import GREEDY_CASES, make_greedy_case

pytestmark = pytest.mark.gpu


class _Tok:
    def decode(self, t):
        return [int(x) for x in t.tolist()]


def _session(name, dev, wrap_predictor=False):
    from model.decoding import RnntGreedyDecoding
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.predictor.stateless_predictor import StatelessPredictor, StatelessPredictorConfig
    g, pcfg, pw, jw, enc = make_greedy_case(name)
    pred = StatelessPredictor(StatelessPredictorConfig(num_symbols=pcfg["num_symbols"], output_dim=pcfg["output_dim"],
                                                       symbol_embedding_dim=pcfg["symbol_embedding_dim"],
                                                       context_size=pcfg["context_size"]))
    pred.load_state_dict({k: torch.from_numpy(v) for k, v in pw.items()})
    joiner = Joiner(JoinerConfig(**g["joiner"]))
    joiner.load_state_dict({k: torch.from_numpy(v) for k, v in jw.items()})
    pred, joiner = pred.to(dev).eval(), joiner.to(dev).eval()
    if wrap_predictor:  # model.predictor.predictor.Predictor keeps the module under .predictor
        holder = torch.nn.Module()
        holder.predictor = pred
        holder.streaming_step = pred.streaming_step
        holder.init_state = pred.init_state
        pred = holder
    return RnntGreedyDecoding(_Tok(), pred, joiner), g, pcfg, pw, jw, enc


@pytest.mark.parametrize("name", list(GREEDY_CASES))
def test_batched_greedy_decode_matches_reference_goldens(name):
    from model.decoding import batch_search
    dev = torch.device("cuda:0")
    session, g, pcfg, pw, jw, enc = _session(name, dev, wrap_predictor=(name == "greedy_outproj"))
    gold = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    T = max(e.shape[1] for e in enc)
    batch = torch.zeros(len(enc), T, enc[0].shape[2])
    for i, e in enumerate(enc):
        batch[i, :e.shape[1]] = torch.from_numpy(e[0])
    lengths = torch.tensor([e.shape[1] for e in enc])
    got = batch_search(batch.to(dev), lengths.to(dev), session)  # ONE launch for the whole batch
    for i in range(len(enc)):
        assert got[i] == gold[f"tokens_{i}"].tolist(), (name, i)
    # decode(): the reference's single-utterance entry point
    assert session.decode(torch.from_numpy(enc[0]).to(dev)) == gold["tokens_0"].tolist()


def test_greedy_decode_at_size_matches_the_port():
    """A c3-like evaluation shape (V=500, D=512, E=512, I=256, context 5), 8 utterances of up to 120 frames."""
    from model.decoding import RnntGreedyDecoding
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.predictor.stateless_predictor import StatelessPredictor, StatelessPredictorConfig
    dev = torch.device("cuda:0")
    torch.manual_seed(21)
    V, D, E, I, C = 500, 512, 512, 256, 5
    pred = StatelessPredictor(StatelessPredictorConfig(num_symbols=V, output_dim=D, symbol_embedding_dim=E, context_size=C))
    jc = dict(input_dim=D, output_dim=V, inner_dim=I, activation="tanh", prune_range=5)
    joiner = Joiner(JoinerConfig(**jc))
    with torch.no_grad():
        for p in joiner.parameters():
            p.mul_(6.0)
        joiner._out_projection[1].bias[0] += 34.0
    lens = torch.tensor([120, 97, 64, 120, 33, 80, 111, 5])
    hidden = torch.randn(len(lens), 120, D)
    pw = {k: v.detach() for k, v in pred.state_dict().items()}
    jw = {k: v.detach() for k, v in joiner.state_dict().items()}
    session = RnntGreedyDecoding(_Tok(), pred.to(dev).eval(), joiner.to(dev).eval())
    got = session.batch_decode_tokens(hidden.to(dev), lens.to(dev))
    same = 0
    for i in range(len(lens)):
        ref = port.rnnt_greedy_decode(pw, jw, jc, hidden[i:i + 1, :int(lens[i])], C)
        assert len(ref) > 0
        same += got[i] == ref
    # different summation orders of the two fp32 matrix-vector products can flip an argmax between two classes whose
    # logits tie to the last bits; after such a flip the context differs and the tails diverge.  With peaked
    # distributions that is rare: all utterances must agree here.
    assert same == len(lens), f"{same} of {len(lens)} utterances identical"


def test_other_predictors_take_the_stepwise_path():
    """A predictor the kernel does not cover (no embedding / conv attributes) is decoded through streaming_step, like
    the reference does."""
    from model.decoding import RnntGreedyDecoding
    dev = torch.device("cuda:0")
    session, g, pcfg, pw, jw, enc = _session("greedy_plain", dev)

    class Opaque(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def init_state(self):
            return self.inner.init_state().to(dev)

        def streaming_step(self, tok, state):
            return self.inner.streaming_step(tok, state)

    slow = RnntGreedyDecoding(_Tok(), Opaque(session._predictor), session._joiner)
    gold = dict(np.load(os.path.join(GOLDEN, "greedy_plain.npz")))
    assert slow.decode(torch.from_numpy(enc[1]).to(dev)) == gold["tokens_1"].tolist()


def _beam_session(name, dev):
    from model.decoding import RnntBeamDecoding
    from oracle.make_golden import BEAM_SIZE, BEAM_TOP_K
    greedy, g, pcfg, pw, jw, enc = _session(name, dev)
    return RnntBeamDecoding(_Tok(), greedy._predictor, greedy._joiner, beam_size=BEAM_SIZE, cutoff_top_k=BEAM_TOP_K), enc


@pytest.mark.parametrize("name", list(GREEDY_CASES))
def test_batched_beam_search_matches_reference_goldens(name):
    """The device-resident beam search (one CTA per utterance, one launch per batch) against the hypotheses of the
    reference's RnntBeamDecoding (/root/reference/model/decoding.py:295-425, run verbatim by oracle/make_golden.py):
    identical token sequences, log-probabilities of the best hypothesis to 1e-4."""
    from model.decoding import DecodingFactory, RnntBeamDecoding, batch_search
    assert DecodingFactory.rnnt_beam_decoding.value is RnntBeamDecoding  # the factory hands out the replacement
    dev = torch.device("cuda:0")
    session, enc = _beam_session(name, dev)
    gold = dict(np.load(os.path.join(GOLDEN, name.replace("greedy", "beam") + ".npz")))
    T = max(e.shape[1] for e in enc)
    batch = torch.zeros(len(enc), T, enc[0].shape[2])
    for i, e in enumerate(enc):
        batch[i, :e.shape[1]] = torch.from_numpy(e[0])
    lengths = torch.tensor([e.shape[1] for e in enc])
    got = batch_search(batch.to(dev), lengths.to(dev), session)
    for i in range(len(enc)):
        assert got[i] == gold[f"tokens_{i}"].tolist(), (name, i)
        ref = float(gold[f"score_{i}"])
        assert abs(session.best_scores[i] - ref) < 1e-4 * abs(ref), (session.best_scores[i], ref)
    assert session.decode(torch.from_numpy(enc[0]).to(dev)) == gold["tokens_0"].tolist()
    # the stepwise path (what an LSTM predictor takes) gives the same hypotheses
    toks, score = session._beam_stepwise(torch.from_numpy(enc[1]).to(dev))
    assert toks == gold["tokens_1"].tolist()


def test_beam_search_at_size_matches_the_port():
    """c3-like evaluation shape (V=500, D=512, E=512, I=256, context 5), 6 utterances of up to 100 frames, beam 4."""
    from model.decoding import RnntBeamDecoding
    from model.joiner.joiner import Joiner, JoinerConfig
    from model.predictor.stateless_predictor import StatelessPredictor, StatelessPredictorConfig
    dev = torch.device("cuda:0")
    torch.manual_seed(23)
    V, D, E, I, C = 500, 512, 512, 256, 5
    pred = StatelessPredictor(StatelessPredictorConfig(num_symbols=V, output_dim=D, symbol_embedding_dim=E, context_size=C))
    jc = dict(input_dim=D, output_dim=V, inner_dim=I, activation="tanh", prune_range=5)
    joiner = Joiner(JoinerConfig(**jc))
    with torch.no_grad():
        for p in joiner.parameters():
            p.mul_(6.0)
        joiner._out_projection[1].bias[0] += 30.0
    lens = torch.tensor([100, 77, 64, 100, 33, 5])
    hidden = torch.randn(len(lens), 100, D)
    pw = {k: v.detach() for k, v in pred.state_dict().items()}
    jw = {k: v.detach() for k, v in joiner.state_dict().items()}
    session = RnntBeamDecoding(_Tok(), pred.to(dev).eval(), joiner.to(dev).eval(), beam_size=4, cutoff_top_k=4)
    got = session.batch_decode_tokens(hidden.to(dev), lens.to(dev))
    same = 0
    for i in range(len(lens)):
        ref, score = port.rnnt_beam_decode(pw, jw, jc, hidden[i:i + 1, :int(lens[i])], C, 4, 4)
        # the best hypothesis' log-probability agrees even where two hypotheses tie to fp32 rounding and swap places
        assert abs(session.best_scores[i] - score) < 2e-4 * max(1.0, abs(score)), (i, session.best_scores[i], score)
        same += got[i] == ref
    assert same >= len(lens) - 1, f"{same} of {len(lens)} utterances identical"
