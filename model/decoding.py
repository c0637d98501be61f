"""Drop-in for /root/reference/model/decoding.py: everything the reference defines there (CTC decoders, the RNN-T
lexicon search, ...) is re-exported from the reference checkout when it is importable; ``RnntGreedyDecoding``,
``RnntBeamDecoding`` and ``batch_search`` are the device-resident versions (speech2text_b200.decoding, SURVEY.md 8
row f-4), also inside ``DecodingFactory``."""
import importlib.util as _ilu
import os as _os
import sys as _sys

from model import __path__ as _roots

_here = _os.path.dirname(_os.path.abspath(__file__))
for _root in list(_roots):
    _cand = _os.path.join(_root, "decoding.py")
    if _os.path.abspath(_root) != _here and _os.path.exists(_cand):
        _spec = _ilu.spec_from_file_location("model._reference_decoding", _cand)
        _mod = _ilu.module_from_spec(_spec)
        _sys.modules["model._reference_decoding"] = _mod
        try:
            _spec.loader.exec_module(_mod)
        except ImportError:  # optional dependencies of the other decoders (flashlight, torchaudio's ctc_decoder, ...)
            del _sys.modules["model._reference_decoding"]
        else:
            globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("_")})
        break

from speech2text_b200.decoding import RnntBeamDecoding, RnntGreedyDecoding, batch_search  # noqa: E402,F401

# the reference's factory enum (decoding.py:427-435) holds the classes themselves: rebuild it around the replacements
import enum as _enum  # noqa: E402

_members = {m.name: m.value for m in globals()["DecodingFactory"]} if "DecodingFactory" in globals() else {}
_members.update(rnnt_greedy_decoding=RnntGreedyDecoding, rnnt_beam_decoding=RnntBeamDecoding)
DecodingFactory = _enum.unique(_enum.Enum("DecodingFactory", _members))  # without the reference: the two RNN-T searches
