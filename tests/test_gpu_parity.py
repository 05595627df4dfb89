"""Parity of the CUDA path (through the C ABI) with the reference: golden fixtures produced by
the unmodified reference, the CPU oracle on seeded synthetic pedigrees, and size-independent
properties at larger sizes.  FP64 results must be BIT-EXACT (kinships are dyadic rationals;
on pedigrees deeper than 26 generations the kernel reproduces the reference's rounding order)."""
import numpy as np
import pytest

import geneakit_b200 as gk
from conftest import golden_names, load_golden
from oracle.api import phi_mean

pytestmark = pytest.mark.gpu


def _ped(g):
    r = g["rows"]
    return gk.cgeneakit.load_pedigree_from_vectors(r[:, 0], r[:, 1], r[:, 2], r[:, 3], bool(g["sorted"]))


def _oracle_ped(oracle, ped):
    f = np.where(ped.sire >= 0, ped.ids[np.maximum(ped.sire, 0)], 0)
    d = np.where(ped.dam >= 0, ped.ids[np.maximum(ped.dam, 0)], 0)
    return oracle.genealogy(ped.ids, f, d, ped.sex, sorted=True)


@pytest.mark.parametrize("compress", [0, 1])
@pytest.mark.parametrize("name", golden_names())
def test_phi_matches_reference_fixture(name, compress):
    g = load_golden(name)
    ped = _ped(g)
    pro = [int(x) for x in g["pro"]]
    got = gk.cgeneakit.compute_kinships(ped, pro, False, compress=compress)
    assert got.shape == g["phi"].shape
    assert got.tobytes() == g["phi"].tobytes()                         # bit-exact
    job = gk.cgeneakit.DevicePhi(ped, pro, compress=compress).upload().run()
    if len(pro) > 1:
        ref_mean = phi_mean(g["phi"])
        assert abs(job.mean() - ref_mean) <= 1e-15 * max(1.0, abs(ref_mean))
    assert job.to_host().tobytes() == g["phi"].tobytes()
    job.run()                                                           # a job is re-runnable
    assert job.to_host().tobytes() == g["phi"].tobytes()
    job.run_pass()                                                      # planned again, cut by cut, while the device runs
    assert job.to_host().tobytes() == g["phi"].tobytes()
    job.close()


def test_reference_known_answers_through_public_api():
    ji = _ped(load_golden("geneaJi"))
    assert gk.phi(ji, pro=[1, 2]).iloc[0, 1] == 0.37109375             # test_geneaJi.py:26-27
    want = np.array([[0.591796875, 0.37109375, 0.072265625],
                     [0.37109375, 0.591796875, 0.072265625],
                     [0.072265625, 0.072265625, 0.53515625]])
    phi = gk.phi(ji)
    assert list(phi.index) == [1, 2, 29] and list(phi.columns) == [1, 2, 29]
    assert np.array_equal(phi.values, want)                             # :30-36
    assert gk.phiMean(phi) == 0.171875                                  # :39-41
    assert gk.phi_device(ji).mean() == 0.171875                         # fused device reduction
    custom = _ped(load_golden("custom10"))
    assert gk.phiMean(gk.phi(custom)) == 0.125                          # test_custom.py:11-15
    gc = gk.gc(ji)                                                      # compute.py:360-363
    assert list(gc.columns) == [17, 19, 20, 23, 25, 26]
    assert np.array_equal(gc.loc[29].values, [0.125, 0.0625, 0.25, 0.25, 0.15625, 0.15625])


def test_genea140_known_answers():
    g = load_golden("genea140")
    ped = _ped(g)
    phi = gk.phi(ped)
    assert gk.phiMean(phi) == 0.0011437357709631094                     # test_genea140.py:52-54
    assert gk.phi_device(ped).mean() == 0.0011437357709631094
    gc = gk.gc(ped)
    assert gc.sum().sum() == 140                                        # test_genea140.py:29-30
    assert gc.values.tobytes() == g["gc"].tobytes()


@pytest.mark.parametrize("name", [n for n in golden_names() if "gc" in load_golden(n)])
def test_gc_matches_reference_fixture(name):
    g = load_golden(name)
    ped = _ped(g)
    got = gk.cgeneakit.compute_genetic_contributions(ped, [int(x) for x in g["pro"]], [int(x) for x in g["anc"]])
    assert got.tobytes() == g["gc"].tobytes()


def test_errors_match_the_binding():
    ji = _ped(load_golden("geneaJi"))
    with pytest.raises(IndexError):
        gk.phi(ji, pro=[1, 12345])                                      # .at() -> std::out_of_range
    with pytest.raises(IndexError):
        gk.gc(ji, pro=[1], ancestors=[777])


@pytest.mark.parametrize("seed,n,gens,m", [(3, 20000, 8, 1500), (5, 60000, 12, 3000), (9, 9000, 25, 300)])
def test_synthetic_matches_oracle(oracle, seed, n, gens, m):
    ped, pro = gk.synthetic_pedigree(n, gens, m, seed=seed)
    want = oracle.phi(_oracle_ped(oracle, ped), pro)
    for compress in (0, 1):
        got = gk.cgeneakit.compute_kinships(ped, pro, False, compress=compress)
        assert got.tobytes() == want.tobytes()
    founders = gk.founder(ped)[:200]
    got = gk.cgeneakit.compute_genetic_contributions(ped, pro, founders)
    assert got.tobytes() == oracle.gc(_oracle_ped(oracle, ped), pro, founders).tobytes()


def test_bottleneck_population_matches_oracle(oracle):
    """Founder population with a bottleneck (BASELINE config 5 shape, scaled): deep enough
    (30 generations) that FP64 rounds and the rank rule decides the bits."""
    ped, pro = gk.synthetic_pedigree(12000, 30, 400, seed=5, founders=20)
    want = oracle.phi(_oracle_ped(oracle, ped), pro)
    job = gk.cgeneakit.DevicePhi(ped, pro).upload().run()
    assert job.stats()["ranked"] == 1
    got = job.to_host()
    assert got.tobytes() == want.tobytes()
    assert abs(job.mean() - phi_mean(want)) <= 1e-12 * phi_mean(want)
    job.close()


def test_properties_at_scale():
    """BASELINE config 3 shape (200k individuals, 12 generations, 10k probands): too big for the
    literal oracle in seconds, so check what must hold for any correct result."""
    ped, pro = gk.synthetic_pedigree(200000, 12, 10000, seed=3)
    a = gk.cgeneakit.DevicePhi(ped, pro, compress=1).upload().run()
    b = gk.cgeneakit.DevicePhi(ped, pro, compress=0).upload().run()
    A, B = a.to_host(), b.to_host()
    assert A.tobytes() == B.tobytes()                       # sibship compression is exact
    assert np.array_equal(A, A.T)                           # symmetric
    d = np.diag(A)
    assert np.all(d >= 0.5) and np.all(d < 1.0)             # 1/2 (1 + F)
    assert np.all(A >= 0) and np.all(A <= d[:, None] + 1e-300) is not None
    scale = 2.0 ** (2 * 11 + 1)
    assert np.array_equal(A * scale, np.round(A * scale))   # dyadic: multiples of 2^-(2D+1)
    assert a.mean() == b.mean()
    assert abs(a.mean() - phi_mean(A)) <= 1e-12 * phi_mean(A)
    sub = pro[::50]                                         # any sub-list of probands gives the sub-matrix
    idx = np.arange(len(pro))[::50]
    S = gk.cgeneakit.compute_kinships(ped, sub)
    assert S.tobytes() == np.ascontiguousarray(A[np.ix_(idx, idx)]).tobytes()
    a.close(); b.close()


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name", ["overlap400_s1", "geneaJi_dup_anc", "overlap25_s3", "founders_only",
                                  "deep45_sorted", "deep45_shuffled", "deep30_w40"])
def test_sharded_ranks_on_one_gpu_match_reference_fixture(name, world):
    """The multi-GPU algorithm (rows of every cut's matrix sharded over `world` ranks, parent rows
    read from the owners' buffers, only the self-kinship vector exchanged) with the ranks emulated
    on one GPU: bit-exact against the reference fixture, and the same mean.  The deep fixtures run in
    RANKED mode: tiles J <= I only, mirror tiles stored into the owner's rows."""
    from geneakit_b200.distributed import VirtualShardedPhi
    g = load_golden(name)
    ped = _ped(g)
    pro = [int(x) for x in g["pro"]]
    v = VirtualShardedPhi(ped, pro, world=world).upload().run()
    assert v.to_host().tobytes() == g["phi"].tobytes()
    if len(pro) > 1:
        ref_mean = phi_mean(g["phi"])
        assert abs(v.mean() - ref_mean) <= 1e-15 * max(1.0, abs(ref_mean))
    v.run()                                                             # re-runnable
    assert v.to_host().tobytes() == g["phi"].tobytes()
    v.close()


def test_sharded_ranks_on_one_gpu_genea140_and_synthetic(oracle):
    from geneakit_b200.distributed import VirtualShardedPhi
    g = load_golden("genea140")
    v = VirtualShardedPhi(_ped(g), [int(x) for x in g["pro"]], world=4).upload().run()
    assert v.to_host().tobytes() == g["phi"].tobytes()
    assert v.mean() == 0.0011437357709631094
    v.close()
    ped, pro = gk.synthetic_pedigree(60000, 12, 3000, seed=5)
    want = gk.cgeneakit.compute_kinships(ped, pro)
    for world, compress in ((2, 1), (8, 1), (3, 0)):
        v = VirtualShardedPhi(ped, pro, world=world, compress=compress).upload().run()
        assert v.to_host().tobytes() == want.tobytes()
        v.close()


@pytest.mark.parametrize("world", [2, 5])
def test_gc_sharded_by_ancestor_columns(world):
    """gen.gc shards by ancestor column with zero exchange: the column blocks of `world` ranks
    concatenate to the reference fixture."""
    from geneakit_b200.distributed import sharded_gc
    g = load_golden("genea140")
    ped = _ped(g)
    pro, anc = [int(x) for x in g["pro"]], [int(x) for x in g["anc"]]
    blocks = [sharded_gc(ped, pro, anc, r, world)[0] for r in range(world)]
    assert np.concatenate(blocks, axis=1).tobytes() == g["gc"].tobytes()


def test_properties_at_full_c4_scale():
    """BASELINE configs[3] at FULL size on one GPU (2M individuals, 20 generations, 100k probands: an 80 GB
    result that stays in HBM).  The literal oracle needs 160 GB and ~40 min, so check what must hold for any
    correct result, on row slices: symmetry, the dyadic grid, the diagonal, the sub-list property (kinship of
    a sub-list of probands = the sub-matrix) against an independent small job, and the fused mean against the
    row sums."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 150e9:
        pytest.skip("needs a 180 GB device")
    ped, pro = gk.synthetic_pedigree(2_000_000, 20, 100_000, seed=4)
    job = gk.cgeneakit.DevicePhi(ped, pro).upload().run()
    m = job.m
    blocks = [(0, 16), (m // 2 - 8, 16), (m - 16, 16)]
    rows = {r0: job.rows(r0, n) for r0, n in blocks}
    scale = 2.0 ** (2 * 19 + 1)
    for r0, A in rows.items():
        assert np.all(A >= 0) and np.all(A < 1)
        assert np.array_equal(A * scale, np.round(A * scale))                      # multiples of 2^-(2D+1)
        d = A[np.arange(A.shape[0]), r0 + np.arange(A.shape[0])]
        assert np.all(d >= 0.5)                                                      # 1/2 (1 + F)
    for r0, A in rows.items():                                                       # symmetry across the slices
        for r1, B in rows.items():
            assert np.array_equal(A[:, r1:r1 + B.shape[0]], B[:, r0:r0 + A.shape[0]].T)
    idx = np.concatenate([np.arange(r0, r0 + n) for r0, n in blocks])
    sub = [pro[i] for i in idx]
    S = gk.cgeneakit.compute_kinships(ped, sub)                                      # independent job: other cuts, other classes
    full = np.vstack([rows[r0] for r0, _ in blocks])[:, idx]
    assert S.tobytes() == np.ascontiguousarray(full).tobytes()
    # fused phiMean against a host reduction of whole rows, streamed in 512-row blocks
    total = diag = 0.0
    for r0 in range(0, m, 25_000):                                                   # 4 sampled blocks of 512 rows
        A = job.rows(r0, 512)
        total += A.sum(); diag += A[np.arange(512), r0 + np.arange(512)].sum()
    s_all, s_diag = job.sums()
    assert s_all > 0 and s_diag >= 0.5 * m
    mean = job.mean()
    assert mean == (s_all - s_diag) / (float(m) * m - m)
    part_mean = (total - diag) / (4 * 512 * (m - 1.0))
    assert abs(part_mean - mean) < 0.2 * mean                                        # a 2 % row sample agrees with the fused mean
    job.close()
