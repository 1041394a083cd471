"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to compute without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import tnqs_b200 as tq
from tnqs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.exists(_lib.LIB_PATH),
                                reason="libtnqs_b200.so not built (run __graft_entry__.build())")


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tnqs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tnqs_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 20
    assert sorted(_lib.SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None
    assert b"sm_100a" in lib.tnqs_version()


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.ApplyOpts) == 48
    assert C.sizeof(_lib.BpOpts) == 40
    assert C.sizeof(_lib.BpReport) == 16
    assert C.sizeof(_lib.Stats) == 152


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = tq.named_grid((2, 2))
    with pytest.raises(tq.TnqsError) as ei:
        tq.BeliefPropagationCache(tq.zerostate(np.complex64, g))
    assert ei.value.code == 7  # TNQS_ENOGPU


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tensornetworkquantumsimulator.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower().replace("# oracle", ""), f


def test_host_circuit_marshalling():
    g = tq.named_grid((2, 2))
    layer = [("Rx", [(1, 1)], 0.3), ("Rzz", [(1, 1), (2, 1)], 0.2), ("Z", (2, 2))]
    nverts, verts, mats = tq.circuit_arrays(layer, g)
    assert list(nverts) == [1, 2, 1]
    assert verts.tolist() == [[0, -1], [0, 1], [3, -1]]
    assert mats.shape == (2 * (4 + 16 + 4),)
    with pytest.raises(tq.ArgumentError):
        tq.circuit_arrays([("Rq", [(1, 1)], 0.3)], g)
    with pytest.raises(tq.ArgumentError):
        tq.circuit_arrays([("Rx", [(9, 9)], 0.3)], g)
