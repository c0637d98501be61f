"""The C-ABI library must load on a box without a GPU and export every symbol that
include/s2t_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "s2t_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(s2t_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path_entry_points():
    syms = _declared_symbols()
    for must in ("s2t_simple_loss_fwd", "s2t_simple_loss_bwd", "s2t_prune_ranges", "s2t_joiner_loss_fwd",
                 "s2t_joiner_loss_bwd", "s2t_logits_loss_fwd", "s2t_logits_loss_bwd", "s2t_mutual_information",
                 "s2t_linear_fwd", "s2t_linear_bwd", "s2t_last_error", "s2t_abi_version"):
        assert must in syms, must


def test_library_loads_and_exports_every_declared_symbol():
    from speech2text_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with `python -m speech2text_b200.build`"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(handle, s)]
    assert not missing, f"declared in include/s2t_b200.h but not exported: {missing}"
    handle.s2t_abi_version.restype = ctypes.c_int
    assert handle.s2t_abi_version() == 3


def test_ctypes_signatures_cover_the_header():
    from speech2text_b200 import _lib
    declared = set(_declared_symbols())
    bound = set(_lib.exported_symbols())
    assert declared <= bound, f"no ctypes signature for: {sorted(declared - bound)}"
    assert bound <= declared, f"ctypes binds undeclared symbols: {sorted(bound - declared)}"


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure; the product path must not route through it."""
    pkg = os.path.join(ROOT, "speech2text_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M):
                    offenders.append(os.path.join(dirpath, f))
    for f in ("model/joiner/joiner.py", "model/loss/loss.py", "model/loss/pruned_rnnt_loss.py",
              "model/loss/rnnt_loss.py"):
        text = open(os.path.join(ROOT, f)).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M):
            offenders.append(f)
    assert not offenders, offenders
