"""The C-ABI library loads on a CPU-only box and exports every symbol include/cnvetti_b200.h
declares; without a GPU compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cnvetti_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cnvb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from cnvetti_b200 import _native
    L = C.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export: %s" % n
    assert sorted(_native.SYMBOLS) == names, "python binding and header disagree"
    assert b"sm_100a" in _native.lib().cnvb_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cnvetti_b200 as cb
    with pytest.raises(cb.CnvbError) as e:
        cb.Context(0)
    assert "no CUDA device" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cnvetti_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "cnv_oracle" not in txt, f


def test_pack_reads_host_only():
    import cnvetti_b200 as cb
    pos = np.array([100, 200], np.int32)
    isize = np.array([300, -300], np.int32)
    flag = np.array([99, 147], np.uint16)
    cigar_off = np.array([0, 1, 3], np.uint64)
    cigar = np.array([(100 << 4) | 0, (70 << 4) | 4, (30 << 4) | 0], np.uint32)
    end, mid, frag_end, flags = cb.pack_reads(pos, isize, flag, cigar_off, cigar, 0.6)
    assert end.tolist() == [200, 230] and mid.tolist() == [250, 50] and frag_end.tolist() == [400, -100]
    assert flags.tolist() == [99, 147 | 0x1000 | 0x2000]


def test_pile_threshold_host_function(oracle):
    """cnvb_pile_threshold is pure host arithmetic (no GPU): the reference's own known-answer test
    (lib-coverage/src/lib.rs:797-807) and random histograms against the oracle."""
    import json
    import os
    from cnvetti_b200 import _native as N
    from cnvetti_b200 import piles
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "stats_kat.json")))["pile_threshold"]
    assert piles.compute_threshold(kat["y"], kat["fdr"]) == kat["expected"]
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(300):
        n = int(rng.integers(3, 40))
        y = np.sort(rng.integers(0, 5000, n))[::-1].astype(np.uint64)
        y[0] = 0
        fdr = float(rng.choice([1e-4, 1e-3, 1e-2, 0.05]))
        try:
            exp = oracle.pile_threshold(y, fdr)
        except oracle.OraclePanic:
            with pytest.raises(N.CnvbError):
                piles.compute_threshold(y, fdr)
            continue
        assert piles.compute_threshold(y, fdr) == exp
        checked += 1
    assert checked > 50
    with pytest.raises(N.CnvbError, match="No coverage above 1"):
        piles.compute_threshold([0, 3], 0.01)
    with pytest.raises(N.CnvbError, match="y\\[0\\] must be == 0"):
        piles.compute_threshold([1, 3, 2], 0.01)
