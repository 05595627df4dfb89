"""Host logic on the CPU: the C-ABI library loads and exports every declared symbol, the
planner reproduces the reference's cuts, and the tile plan -- interpreted by
tests/plan_emulator.py -- reproduces the reference's kinship matrices bit for bit.
No kernel is launched here."""
import ctypes as C
import re

import numpy as np
import pytest

import geneakit_b200 as gk
from geneakit_b200 import _lib
from conftest import ROOT, golden_names, load_golden
from plan_emulator import emulate


def _ped(g):
    r = g["rows"]
    return gk.cgeneakit.load_pedigree_from_vectors(r[:, 0], r[:, 1], r[:, 2], r[:, 3], bool(g["sorted"]))


def test_library_exports_every_declared_symbol():
    header = open(ROOT + "/include/geneakit_b200.h").read()
    declared = set(re.findall(r"\b(gk_[a-z0-9_]+)\s*\(", header))
    declared -= {"gk_opts", "gk_phi_stats", "gk_step_view", "gk_gc_stats"}
    assert declared, "no declarations found"
    raw = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(raw, s)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= declared
    assert _lib.lib.gk_abi_version() == 2


@pytest.mark.parametrize("name", golden_names())
def test_loader_and_cuts_match_reference(name):
    g = load_golden(name)
    ped = _ped(g)
    assert np.array_equal(ped.ids, g["rank_ids"])                       # create.cpp:121-199
    assert gk.pro(ped) == [int(x) for x in g["default_pro"]]
    assert gk.founder(ped) == [int(x) for x in g["default_founders"]]
    pro = [int(x) for x in g["pro"]]
    r = ped.ranks(pro)
    sizes = np.zeros(256, dtype=np.int32)
    G = _lib.lib.gk_cut_sizes(len(ped), ped.sire, ped.dam, len(r), r, sizes, 256)
    assert [int(x) for x in sizes[:G]] == [int(x) for x in g["cut_sizes"]]
    assert gk.cgeneakit.get_required_memory_for_kinships(ped, pro) == float(g["required_gb"])


@pytest.mark.parametrize("compress", [0, 1])
@pytest.mark.parametrize("name", golden_names(small_only=True))
def test_tile_plan_reproduces_reference_phi(name, compress):
    g = load_golden(name)
    ped = _ped(g)
    job = gk.cgeneakit.DevicePhi(ped, [int(x) for x in g["pro"]], compress=compress)
    st = job.stats()
    assert job.cut_sizes() == [int(x) for x in g["cut_sizes"]]
    if st["n_cuts"] - 1 > 26:
        assert st["ranked"] == 1 and st["compressed"] == 0
    got = emulate(job)
    assert got.tobytes() == g["phi"].tobytes()
    job.close()


def test_unknown_proband_raises_index_error():
    g = load_golden("geneaJi")
    ped = _ped(g)
    with pytest.raises(IndexError):
        gk.cgeneakit.DevicePhi(ped, [1, 999])
    with pytest.raises(IndexError):
        gk.cgeneakit.get_required_memory_for_kinships(ped, [12345])


def test_compression_shrinks_the_matrices():
    ped, pro = gk.synthetic_pedigree(6000, 6, 400, seed=3)
    a = gk.cgeneakit.DevicePhi(ped, pro, compress=0)
    b = gk.cgeneakit.DevicePhi(ped, pro, compress=1)
    assert a.cut_sizes() == b.cut_sizes()
    assert a.class_sizes()[:-1] == a.cut_sizes()[:-1]
    assert sum(b.class_sizes()) < sum(a.class_sizes())
    assert b.stats()["cells_device"] < a.stats()["cells_device"]
    a.close(); b.close()


def test_correction_overflow_falls_back_to_one_row_per_individual():
    """An individual of a multi-member sibship with tens of thousands of mates would need a quadratic correction list:
    the planner reports it and gk_phi_plan plans again uncompressed on its own (a drop-in must not ask the caller)."""
    mates = 17000
    # founders 0, 1 -> full sibs 2, 3 (a two-member class: both are parents); 2 mates with `mates` founders, one child each
    sire = [-1, -1, 0, 0] + [-1] * mates + [2] * (mates - 1) + [3]
    dam = [-1, -1, 1, 1] + [-1] * mates + list(range(4, 4 + mates))
    n = len(sire)
    ids = np.arange(1, n + 1)
    ped = gk.cgeneakit.Pedigree(ids, sire, dam, np.ones(n))
    pro = [int(x) for x in ids[4 + mates:]]                   # every child
    job = gk.cgeneakit.DevicePhi(ped, pro)                    # planning is host-only
    st = job.stats()
    assert st["compressed"] == 0 and st["n_cuts"] == 3
    job.close()
    few = 50
    sire2 = [-1, -1, 0, 0] + [-1] * few + [2] * (few - 1) + [3]
    dam2 = [-1, -1, 1, 1] + [-1] * few + list(range(4, 4 + few))
    ids2 = np.arange(1, len(sire2) + 1)
    small = gk.cgeneakit.DevicePhi(gk.cgeneakit.Pedigree(ids2, sire2, dam2, np.ones(len(sire2))), [int(x) for x in ids2[4 + few:]])
    assert small.stats()["compressed"] == 1 and small.stats()["corrections"] > 0
    small.close()


@pytest.mark.parametrize("world", [1, 3])
def test_planner_does_not_depend_on_the_thread_count(world, monkeypatch):
    """The cuts are planned in order on a pool of host threads (a cut waits for the order of the previous one): the plan
    must be the same arrays whatever the number of threads, for one-GPU and for sharded (lineage-grouped) jobs."""
    from plan_emulator import step_view
    ped, pro = gk.synthetic_pedigree(30000, 10, 1500, seed=7)
    plans = []
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("GK_PLAN_THREADS", threads)
        job = gk.cgeneakit.DevicePhi(ped, pro, rank=0, world=world)
        G = job.stats()["n_cuts"]
        plans.append([step_view(job, g) for g in range(G)])
        job.close()
    for other in plans[1:]:
        assert len(other) == len(plans[0])
        for a, b in zip(plans[0], other):
            assert a["K"] == b["K"] and a["panel_rows"] == b["panel_rows"]
            for key in ("pF", "pM", "flags", "cls_ind", "cls_size", "pg_i", "pg_j", "pg_off", "pt_cls", "pt_mult"):
                assert np.array_equal(a[key], b[key]), key
