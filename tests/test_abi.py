"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header
declares, validates arguments, refuses to compute without a GPU (no CPU fallback), and
its host-side set-up routines agree bit for bit with the oracle's restatement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import unconfined_b200 as ub
from oracle import oracle
from helpers import load_deck, ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "unconfined_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(unc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 17
    lib = ub.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/unconfined_b200.h but not exported"
    assert b"sm_100a" in lib.unc_version()


def test_struct_layout_matches_header():
    # 2 int32, 3 double, 2 int32, ptr, 6 int32, 2 ptr, 12 double, 2 int32, 5 double
    assert C.sizeof(ub.UncParams) == 8 + 24 + 8 + 8 + 24 + 16 + 96 + 8 + 40
    assert C.sizeof(ub.UncParams) == C.sizeof(oracle.OrcParams)


def test_host_tables_bitwise_equal_oracle():
    assert np.array_equal(ub.j0_zeros(40), oracle.j0_zeros(40))
    tD = 10.0 ** np.linspace(-3, 8, 57)
    for j0s in ((1, 1), (2, 6), (7, 3)):
        assert np.array_equal(ub.split_index(tD, j0s), oracle.split_index(tD, j0s))
    z = np.linspace(0, 1, 101)
    assert np.array_equal(ub.zlay(z, 0.7, 0.2), oracle.zlay(z, 0.7, 0.2))


def test_no_cpu_fallback_and_argument_validation():
    d, pd = load_deck("hantush-input.dat")
    prm = ub.Params(pd)
    if ub.device_count() == 0:
        with pytest.raises(ub.UncError) as e:
            ub.eval_grid(prm, d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"])
        assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    # parameter validation happens before any device work
    # model 6: only MNtype 1 (closed-form Malama variant) exists; 0 (ARB) and 2 (FD) are refused
    for mn in (0, 2):
        bad = dict(pd, model=6, mn_type=mn, mn_ak=0.5, mn_b=20.0, mn_psia=0.02, mn_psik=0.02, mn_Sy=0.3)
        with pytest.raises(ub.UncError) as e:
            ub.eval_grid(ub.Params(bad), d["tD"], d["sv"], d["rD"], d["zD"], d["zLay"])
        assert e.value.code == -2 and "MNtype" in str(e.value)
    # empty inputs succeed trivially
    s, ds = ub.eval_points(prm, [], [], [], [], [])
    assert s.size == 0
    # sv out of range of the supplied zeros
    if ub.device_count() > 0:
        with pytest.raises(ub.UncError):
            ub.eval_grid(prm, d["tD"][:1], np.array([5], np.int32), d["rD"], d["zD"], d["zLay"])


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import unconfined_b200.api as api
    monkeypatch.setattr(api, "_LIB", None)
    monkeypatch.setattr(api, "_SO", str(tmp_path / "nope.so"))
    with pytest.raises(ub.UncError):
        api.lib()
