"""The CPU oracle (oracle/gk_oracle.c) against the reference's golden vectors and against
fixtures produced by the unmodified reference (oracle/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle.api import phi_mean


def _ped(oracle, g):
    r = g["rows"]
    return oracle.genealogy(r[:, 0], r[:, 1], r[:, 2], r[:, 3], sorted=bool(g["sorted"]))


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_fixture(oracle, name):
    g = load_golden(name)
    ped = _ped(oracle, g)
    # loader rank order: create.cpp:121-199
    assert np.array_equal(ped.ids, g["rank_ids"])
    assert oracle.pro(ped) == [int(x) for x in g["default_pro"]]
    assert oracle.founder(ped) == [int(x) for x in g["default_founders"]]
    pro = [int(x) for x in g["pro"]]
    assert oracle.cut_sizes(ped, pro) == [int(x) for x in g["cut_sizes"]]
    assert oracle.required_gb(ped, pro) == float(g["required_gb"])
    phi = oracle.phi(ped, pro)
    assert phi.tobytes() == g["phi"].tobytes()          # bit-exact, including deep pedigrees
    assert hashlib.sha256(phi.tobytes()).hexdigest().encode() == bytes(g["phi_sha256"])
    if "gc" in g:
        anc = [int(x) for x in g["anc"]]
        gc = oracle.gc(ped, pro, anc)
        assert gc.tobytes() == g["gc"].tobytes()
        assert hashlib.sha256(gc.tobytes()).hexdigest().encode() == bytes(g["gc_sha256"])
        if ped.n <= 500:
            assert oracle.gc(ped, pro, anc, dfs=True).tobytes() == g["gc"].tobytes()


def test_reference_known_answers(oracle):
    """The literal assertions of the reference's own tests."""
    g = load_golden("geneaJi")
    ped = _ped(oracle, g)
    assert oracle.pro(ped) == [1, 2, 29]                                   # test_geneaJi.py:8-9
    assert oracle.founder(ped) == [17, 19, 20, 23, 25, 26]                 # :12-13
    assert oracle.phi(ped, [1, 2])[0, 1] == 0.37109375                     # :26-27
    want = np.array([[0.591796875, 0.37109375, 0.072265625],
                     [0.37109375, 0.591796875, 0.072265625],
                     [0.072265625, 0.072265625, 0.53515625]])
    assert np.array_equal(oracle.phi(ped), want)                           # :30-36
    assert phi_mean(oracle.phi(ped)) == 0.171875                           # :39-41
    gc = oracle.gc(ped)                                                    # compute.py:360-363
    assert np.array_equal(gc[0], [0.3125, 0.21875, 0.0625, 0.0, 0.109375, 0.109375])
    assert np.array_equal(gc[2], [0.125, 0.0625, 0.25, 0.25, 0.15625, 0.15625])
    c = load_golden("custom10")
    assert phi_mean(oracle.phi(_ped(oracle, c))) == 0.125                  # test_custom.py:11-15


def test_genea140_known_answers(oracle):
    g = load_golden("genea140")
    ped = _ped(oracle, g)
    assert ped.n == 41523
    phi = oracle.phi(ped)
    assert phi_mean(phi) == 0.0011437357709631094                          # test_genea140.py:52-54
    assert hashlib.sha256(phi.tobytes()).hexdigest() == \
        "74133d7870442ae93cbc6aba364859dc34c6a250b18868bd61a97180cb8fae26"
    gc = oracle.gc(ped)
    assert gc.sum(axis=0).sum() == 140                                     # test_genea140.py:29-30
    assert hashlib.sha256(gc.tobytes()).hexdigest() == \
        "c532170fff1c61f9dbcedf22ff4b47312bea697036d641cfc23962f9e409a693"


def test_unknown_id_raises(oracle):
    g = load_golden("geneaJi")
    ped = _ped(oracle, g)
    with pytest.raises(IndexError):
        oracle.phi(ped, [1, 999])
