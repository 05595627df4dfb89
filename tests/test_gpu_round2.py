"""Round-2 parity tests (GPU, through the C ABI): both step kernels on every fixture, the benchmark configurations against
the oracle at sizes it finishes in seconds (C3 in full, 1/10-width replicas of C4 and C5, phi and gc), the one-process
multi-GPU driver behind gk_phi_f64 (ranks emulated on one GPU), and the consumers of SURVEY 8f (gen.f, gen.phiOver,
gen.phiCI, gen.phi(sparse=True)) against outputs of the unmodified reference (tests/golden, oracle/make_golden.py)."""
import numpy as np
import pytest

import geneakit_b200 as gk
from conftest import golden_names, load_golden
from oracle.api import Oracle, Reference, phi_mean

pytestmark = pytest.mark.gpu


def _ped(g):
    r = g["rows"]
    return gk.cgeneakit.load_pedigree_from_vectors(r[:, 0], r[:, 1], r[:, 2], r[:, 3], bool(g["sorted"]))


def _checker():
    """The unmodified reference when it was built in this tree (oracle/_ref), else our C restatement of it."""
    return Reference() if Reference.available() else Oracle()


def _checker_ped(eng, ped):
    f = np.where(ped.sire >= 0, ped.ids[np.maximum(ped.sire, 0)], 0)
    d = np.where(ped.dam >= 0, ped.ids[np.maximum(ped.dam, 0)], 0)
    return eng.genealogy(ped.ids, f, d, ped.sex, sorted=True)


# ---- every fixture through each step kernel (the default picks per cut size: small cuts would never reach the TMA kernel) ----
@pytest.mark.parametrize("kernel", ["g4", "ldgsts"])
@pytest.mark.parametrize("name", golden_names())
def test_fixture_through_each_step_kernel(name, kernel, monkeypatch):
    monkeypatch.setenv("GK_STEP_KERNEL", kernel)
    monkeypatch.setenv("GK_G4_MIN_K", "32")
    g = load_golden(name)
    ped = _ped(g)
    pro = [int(x) for x in g["pro"]]
    for compress in (0, 1):
        got = gk.cgeneakit.compute_kinships(ped, pro, False, compress=compress)
        assert got.tobytes() == g["phi"].tobytes()


@pytest.mark.parametrize("kernel", ["g4", "ldgsts"])
def test_synthetic_through_each_step_kernel(oracle, kernel, monkeypatch):
    monkeypatch.setenv("GK_STEP_KERNEL", kernel)
    monkeypatch.setenv("GK_G4_MIN_K", "32")
    ped, pro = gk.synthetic_pedigree(60000, 12, 3000, seed=5)
    want = oracle.phi(_checker_ped(oracle, ped), pro)
    for compress in (0, 1):
        assert gk.cgeneakit.compute_kinships(ped, pro, False, compress=compress).tobytes() == want.tobytes()
    ped, pro = gk.synthetic_pedigree(12000, 30, 400, seed=5, founders=20)          # ranked mode (30 generations)
    want = oracle.phi(_checker_ped(oracle, ped), pro)
    assert gk.cgeneakit.compute_kinships(ped, pro).tobytes() == want.tobytes()


# ---- the benchmark's configurations against the oracle -------------------------------------------------------------
def test_c3_full_size_matches_the_oracle():
    """BASELINE configs[2] at FULL size (200 k individuals, 12 generations, 10 k probands, seed 3): bit-exact."""
    eng = _checker()
    ped, pro = gk.synthetic_pedigree(200_000, 12, 10_000, seed=3)
    want = eng.phi(_checker_ped(eng, ped), pro)
    job = gk.cgeneakit.DevicePhi(ped, pro).upload().run()
    got = job.to_host()
    assert got.tobytes() == want.tobytes()
    ref_mean = phi_mean(want)
    assert abs(job.mean() - ref_mean) <= 1e-12 * ref_mean
    job.close()


def test_c4_tenth_width_replica_matches_the_oracle():
    """BASELINE configs[3] at 1/10 width (200 k individuals, 20 generations, 10 k probands, seed 4): phi and gc bit-exact."""
    eng = _checker()
    ped, pro = gk.synthetic_pedigree(200_000, 20, 10_000, seed=4)
    h = _checker_ped(eng, ped)
    assert gk.cgeneakit.compute_kinships(ped, pro).tobytes() == eng.phi(h, pro).tobytes()
    founders = gk.founder(ped)[:300]
    lin = Oracle()                                       # the reference's gc enumerates paths: exponential at this depth
    want = lin.gc(_checker_ped(lin, ped), pro, founders)
    assert gk.cgeneakit.compute_genetic_contributions(ped, pro, founders).tobytes() == want.tobytes()


def test_c5_tenth_width_replica_matches_the_oracle():
    """BASELINE configs[4] at 1/10 width (500 k individuals, 30 generations, 200 founders, 5 k probands, seed 5): RANKED mode
    at a real size (FP64 rounds, the reference's rank-driven grouping decides the bits); phi bit-exact, gc bit-exact."""
    eng = _checker()
    ped, pro = gk.synthetic_pedigree(500_000, 30, 5_000, seed=5, founders=200)
    h = _checker_ped(eng, ped)
    job = gk.cgeneakit.DevicePhi(ped, pro).upload().run()
    assert job.stats()["ranked"] == 1
    want = eng.phi(h, pro)
    assert job.to_host().tobytes() == want.tobytes()
    job.close()
    founders = gk.founder(ped)
    lin = Oracle()
    want = lin.gc(_checker_ped(lin, ped), pro, founders)
    got = gk.cgeneakit.compute_genetic_contributions(ped, pro, founders)
    assert got.tobytes() == want.tobytes()
    assert abs(got.sum() - len(pro)) < 1e-9              # every proband's founders contribute 1 in total


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name", ["overlap400_s1", "genea140", "overlap400_s3"])
def test_sharded_ranks_through_the_default_kernel(name, world, monkeypatch):
    """The default step kernel in sharded jobs (every tile J of another rank's panels, J <= I inside the own block)."""
    from geneakit_b200.distributed import VirtualShardedPhi
    monkeypatch.setenv("GK_STEP_KERNEL", "g4")
    monkeypatch.setenv("GK_G4_MIN_K", "32")
    g = load_golden(name)
    for compress in (0, 1):
        v = VirtualShardedPhi(_ped(g), [int(x) for x in g["pro"]], world=world, compress=compress).upload().run()
        assert v.to_host().tobytes() == g["phi"].tobytes()
        v.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["deep45_shuffled", "deep30_w40"])
def test_ranked_sharded_ranks_through_the_default_kernel(name, world, monkeypatch):
    """Deep (ranked) sharded jobs run the default step kernel too: tiles J <= I, own rows only, mirror tiles pulled."""
    from geneakit_b200.distributed import VirtualShardedPhi
    monkeypatch.setenv("GK_STEP_KERNEL", "g4")
    monkeypatch.setenv("GK_G4_MIN_K", "32")
    g = load_golden(name)
    v = VirtualShardedPhi(_ped(g), [int(x) for x in g["pro"]], world=world).upload().run()
    assert v.jobs[0].stats()["ranked"] == 1
    assert v.to_host().tobytes() == g["phi"].tobytes()
    v.close()


# ---- one process, several GPUs, behind the one-shot entry point --------------------------------------------------------
@pytest.mark.parametrize("gpus", [2, 3, 8])
@pytest.mark.parametrize("name", ["overlap400_s1", "geneaJi_dup_anc", "founders_only", "deep45_shuffled", "genea140"])
def test_multi_gpu_driver_with_virtual_devices(name, gpus, monkeypatch):
    """gk_phi_f64 with opts.world = N, opts.rank = -1 drives N ranks from this process (one plan, peer pointers, the
    diagonal kernel writes every replica of the self-kinship vector).  The test box has one GPU: the ranks share it on one
    stream (GENEAKIT_VIRTUAL_DEVICES=1); the same code drives N devices with events between the cuts."""
    monkeypatch.setenv("GENEAKIT_VIRTUAL_DEVICES", "1")
    g = load_golden(name)
    ped = _ped(g)
    pro = [int(x) for x in g["pro"]]
    got = gk.cgeneakit.compute_kinships(ped, pro, False, gpus=gpus)
    assert got.tobytes() == g["phi"].tobytes()


def test_gen_phi_on_all_the_gpus_of_the_box():
    """gen.phi(ped, pro=..., gpus=N) on real devices (skipped below two GPUs)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    ped, pro = gk.synthetic_pedigree(60000, 12, 3000, seed=5)
    want = gk.phi(ped, pro=pro)
    got = gk.phi(ped, pro=pro, gpus=n)
    assert got.values.tobytes() == want.values.tobytes()
    g = load_golden("deep45_shuffled")
    got = gk.cgeneakit.compute_kinships(_ped(g), [int(x) for x in g["pro"]], False, gpus=min(n, 4))
    assert got.tobytes() == g["phi"].tobytes()


# ---- SURVEY 8f rows -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_f_matches_the_reference(name):
    g = load_golden(name)
    ped = _ped(g)
    pro = [int(x) for x in g["pro"]]
    got = gk.f(ped, pro=pro)
    assert list(got.index) == pro and list(got.columns) == ["F"]
    assert np.max(np.abs(got["F"].values - g["f"])) <= 1e-15       # exact up to 26 generations; the reference's own
    if not name.startswith("deep"):                                  # algorithm rounds differently beyond
        assert got["F"].values.tobytes() == g["f"].tobytes()
    job = gk.phi_device(ped, pro=pro)
    assert job.f().tobytes() == got["F"].values.tobytes()
    job.close()


def test_f_known_answer():
    ji = _ped(load_golden("geneaJi"))
    assert gk.f(ji).loc[1, "F"] == 0.18359375                        # reference test/test_geneaJi.py:22-23
    assert list(gk.f(ji)["F"]) == [0.18359375, 0.18359375, 0.0703125]  # geneakit/compute.py:268-272


@pytest.mark.parametrize("name", [n for n in golden_names() if "sp_data" in load_golden(n)])
def test_sparse_matches_the_reference(name):
    g = load_golden(name)
    ped = _ped(g)
    pro = [int(x) for x in g["pro"]]
    m = gk.phi(ped, pro=pro, sparse=True)
    assert m.shape == (len(pro), len(pro)) and m.dtype == np.float32
    assert np.array_equal(m.indptr, g["sp_indptr"]) and np.array_equal(m.indices, g["sp_indices"])
    assert m.data.tobytes() == g["sp_data"].tobytes()               # float32, bit for bit
    dense = gk.phi(ped, pro=pro).values
    assert np.max(np.abs(m.toarray() - np.tril(dense))) <= 1e-7     # reference test/test_genea140.py:57-59 (1e-8 on genea140)


def test_phi_over_matches_the_reference_loop():
    ji = _ped(load_golden("geneaJi"))
    job = gk.phi_device(ji)
    over = gk.phiOver(job, 0.1)                                      # geneakit/compute.py:150-155: one pair, (2, 1, 0.371094)
    assert over.shape == (1, 3) and int(over.pro1[0]) == 2 and int(over.pro2[0]) == 1 and over.kinship[0] == 0.37109375
    job.close()
    g = load_golden("overlap400_s1")
    ped, pro = _ped(g), [int(x) for x in g["pro"]]
    job = gk.phi_device(ped, pro=pro)
    for thr in (0.0, 0.03, 0.2, 0.6):
        got = gk.phiOver(job, thr)
        phi = g["phi"]
        want = [(pro[i], pro[j], phi[i, j]) for i in range(len(pro)) for j in range(i) if phi[i, j] > thr]   # the reference's loop
        assert [(int(a), int(b), c) for a, b, c in zip(got.pro1, got.pro2, got.kinship)] == want
    job.close()


def test_phi_ci_statistic_and_bootstrap():
    g = load_golden("genea140")
    ped, pro = _ped(g), [int(x) for x in g["pro"]]
    phi = g["phi"]
    job = gk.phi_device(ped, pro=pro)
    rng = np.random.default_rng(7)
    idx = rng.integers(0, len(pro), (40, len(pro)))
    got = job.ci_stats(idx)

    def statistic(ix):                                               # geneakit/compute.py:226-229
        sub = phi[ix][:, ix]
        sel = sub[sub < 0.5]
        return np.mean(sel) if sel.size else np.nan
    want = np.array([statistic(ix) for ix in idx])
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    # the whole gen.phiCI: same resamples as scipy's bootstrap draws for this seed -> same percentile interval
    a = gk.phiCI(job, b=300, random_state=11)
    b = gk.phiCI(gk.phi(ped, pro=pro), b=300, random_state=11)
    assert list(a.columns) == ["2.5%", "5.0%", "95.0%", "97.5%"]
    assert np.allclose(a.values, b.values, rtol=1e-9, atol=0)
    job.close()


# ---- job plumbing added this round ---------------------------------------------------------------------------------------
def test_replan_and_cache_reuse():
    g = load_golden("overlap400_s2")
    ped, pro = _ped(g), [int(x) for x in g["pro"]]
    job = gk.cgeneakit.DevicePhi(ped, pro).upload().run()
    assert job.to_host().tobytes() == g["phi"].tobytes()
    job.replan().run()                                               # plan + H2D again on the same buffers
    assert job.to_host().tobytes() == g["phi"].tobytes()
    job.close()
    for _ in range(3):                                               # buffers come back from the library's cache
        assert gk.cgeneakit.compute_kinships(ped, pro).tobytes() == g["phi"].tobytes()
    from geneakit_b200._lib import lib
    lib.gk_cache_release()
    assert gk.cgeneakit.compute_kinships(ped, pro).tobytes() == g["phi"].tobytes()


@pytest.mark.parametrize("name", ["genea140", "deep30_w40", "overlap400_s1"])
def test_symmetric_copy_out(name, monkeypatch):
    """The one-shot call copies the block rows of the lower triangle and mirrors them on the host: same bits."""
    g = load_golden(name)
    ped, pro = _ped(g), [int(x) for x in g["pro"]]
    monkeypatch.setenv("GK_D2H_SYMMETRIC", "1")
    assert gk.cgeneakit.compute_kinships(ped, pro).tobytes() == g["phi"].tobytes()
    monkeypatch.setenv("GK_D2H_SYMMETRIC", "0")
    assert gk.cgeneakit.compute_kinships(ped, pro).tobytes() == g["phi"].tobytes()


def test_gc_in_blocks_of_ancestor_columns(monkeypatch):
    """A wide ancestor list is propagated block of columns by block of columns (HBM budget): same bits."""
    from geneakit_b200 import _lib
    import ctypes as C
    g = load_golden("genea140")
    ped = _ped(g)
    pro, anc = [int(x) for x in g["pro"]], [int(x) for x in g["anc"]]
    monkeypatch.setenv("GK_GC_BUDGET_MB", "64")                       # forces several blocks of the 7399 founders
    got = gk.cgeneakit.compute_genetic_contributions(ped, pro, anc)
    assert got.tobytes() == g["gc"].tobytes()
    st = _lib.gk_gc_stats()
    assert _lib.lib.gk_gc_last_stats(C.byref(st)) == 0 and st.column_blocks > 1


def _random_frame(size, seed):
    """Overlapping generations with 25 % unknown fathers / mothers, as the reference's own property test builds them
    (test/test_edit.py:332-343)."""
    rng = np.random.default_rng(seed)
    rows = []
    for ident in range(1, size + 1):
        sex = int(rng.integers(1, 3))
        f = int(rng.integers(1, ident)) if ident > 1 and rng.random() < 0.75 else 0
        m = int(rng.integers(1, ident)) if ident > 1 and rng.random() < 0.75 else 0
        if f == m:
            m = 0
        rows.append((ident, f, m, sex))
    return np.asarray(rows, dtype=np.int64)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_phi_gc_f_do_not_depend_on_the_rank_order(seed):
    """The reference's own invariant (test/test_edit.py:363-397, analytics()): the same pedigree loaded in another order
    gets other ranks, and phi / gc / f must agree to 12 decimals."""
    rows = _random_frame(300, seed)
    rng = np.random.default_rng(100 + seed)
    a = gk.genealogy(__import__("pandas").DataFrame(rows, columns=["ind", "father", "mother", "sex"]))
    shuffled = rows[rng.permutation(len(rows))]
    b = gk.genealogy(__import__("pandas").DataFrame(shuffled, columns=["ind", "father", "mother", "sex"]))
    assert sorted(gk.pro(a)) == sorted(gk.pro(b)) and not np.array_equal(a.ids, b.ids)
    pro = gk.pro(a)
    assert np.array_equal(np.round(gk.phi(a, pro=pro).values, 12), np.round(gk.phi(b, pro=pro).values, 12))
    assert np.array_equal(np.round(gk.gc(a, pro=pro).values, 12), np.round(gk.gc(b, pro=pro).values, 12))
    assert np.array_equal(np.round(gk.f(a, pro=pro).values, 12), np.round(gk.f(b, pro=pro).values, 12))


def test_one_shot_call_falls_back_to_one_row_per_individual_mid_pass():
    """An individual of a two-member sibship with 17 000 mates: the sibship compression would need a quadratic correction
    list.  The one-shot call finds out while it already runs the first cuts (the cuts are planned while the device works),
    drains, and runs again with one row per individual -- the caller sees nothing but the right matrix: half sibs 1/8,
    the child of the other sib 1/16 with all of them, 1/2 on the diagonal."""
    mates = 17000
    sire = [-1, -1, 0, 0] + [-1] * mates + [2] * (mates - 1) + [3]
    dam = [-1, -1, 1, 1] + [-1] * mates + list(range(4, 4 + mates))
    n = len(sire)
    ids = np.arange(1, n + 1)
    ped = gk.cgeneakit.Pedigree(ids, sire, dam, np.ones(n))
    pro = [int(x) for x in ids[4 + mates:]]
    got = gk.cgeneakit.compute_kinships(ped, pro)
    m = len(pro)
    assert got.shape == (m, m)
    assert np.all(np.diag(got) == 0.5)
    assert np.all(got[-1, :-1] == 0.0625) and np.all(got[:-1, -1] == 0.0625)
    block = got[:2000, :2000]
    assert np.all(block[~np.eye(2000, dtype=bool)] == 0.125)
    assert np.array_equal(got[5000:5100], got[:, 5000:5100].T)
