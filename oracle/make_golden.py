"""TEST INFRASTRUCTURE: regenerate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/libgkref.so):
    python oracle/make_golden.py
Each fixture holds the INPUT in file order (ids, fathers, mothers, sexes, 0 = unknown parent),
the proband / ancestor id lists that were passed, and what the reference returned:
rank order after loading, cut sizes, phi matrix, gc matrix, required GB.
The fixtures travel to the GPU box; the reference does not.
"""
import hashlib
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.api import Reference  # noqa: E402

REF_DATA = "/root/reference/geneakit/datasets/"
OUT = os.path.join(ROOT, "tests", "golden")


def read_csv4(path):
    rows = []
    with open(path) as fh:
        next(fh)
        for line in fh:
            t = line.split()
            if len(t) >= 4:
                rows.append([int(x) for x in t[:4]])
    return np.asarray(rows, dtype=np.int32)


def overlapping_pedigree(size, seed, p_known=0.75):
    """Overlapping generations with half-founders: every individual picks each parent, with
    probability p_known, among the earlier individuals of the right sex (the shape the
    reference's own property tests use, test/test_edit.py:332-343)."""
    g = random.Random(seed)
    rows, males, females = [], [], []
    for ident in range(1, size + 1):
        sex = g.choice([1, 2])
        f = g.choice(males) if males and g.random() < p_known else 0
        m = g.choice(females) if females and g.random() < p_known else 0
        rows.append((ident, f, m, sex))
        (males if sex == 1 else females).append(ident)
    return np.asarray(rows, dtype=np.int32)


def narrow_deep_pedigree(width, gens, seed):
    """Discrete generations of `width` individuals, random mating inside the previous
    generation: deep and heavily inbred (kinships need > 53 bits beyond 26 generations, so
    the reference's rank-driven rounding becomes observable)."""
    rng = np.random.default_rng(seed)
    rows, prev = [], []
    ident = 0
    for g in range(gens):
        cur = []
        for k in range(width):
            ident += 1
            sex = 1 + (k % 2)
            if g == 0:
                f = m = 0
            else:
                f = int(rng.choice([x for x, s in prev if s == 1]))
                m = int(rng.choice([x for x, s in prev if s == 2]))
            rows.append((ident, f, m, sex))
            cur.append((ident, sex))
        prev = cur
    return np.asarray(rows, dtype=np.int32)


def relabel_and_shuffle(rows, seed):
    """Random id relabelling + random file order, so the loader's DFS sort produces ranks that
    differ from both id order and generation order."""
    rng = np.random.default_rng(seed)
    n = len(rows)
    new_id = rng.permutation(n) + 1
    m = np.zeros(n + 1, dtype=np.int64)
    m[rows[:, 0]] = new_id
    out = rows.copy()
    out[:, 0] = m[rows[:, 0]]
    out[:, 1] = np.where(rows[:, 1] > 0, m[rows[:, 1]], 0)
    out[:, 2] = np.where(rows[:, 2] > 0, m[rows[:, 2]], 0)
    return out[rng.permutation(n)]


def emit(R, name, rows, pro=None, anc=None, sorted_flag=False, with_gc=True, gc_sparse=False):
    h = R.genealogy(rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3], sorted=sorted_flag)
    flat = R.flatten(h)
    pro_l = R.pro(h) if pro is None else list(pro)
    anc_l = R.founder(h) if anc is None else list(anc)
    phi = R.phi(h, pro_l)
    d = dict(rows=rows.astype(np.int32), sorted=np.int32(sorted_flag),
             pro=np.asarray(pro_l, dtype=np.int32), pro_default=np.int32(pro is None),
             anc=np.asarray(anc_l, dtype=np.int32), anc_default=np.int32(anc is None),
             rank_ids=flat.ids, default_pro=np.asarray(R.pro(h), dtype=np.int32),
             default_founders=np.asarray(R.founder(h), dtype=np.int32),
             cut_sizes=np.asarray(R.cut_sizes(h, pro_l), dtype=np.int32),
             required_gb=np.float64(R.required_gb(h, pro_l)),
             phi=phi, phi_sha256=np.bytes_(hashlib.sha256(phi.tobytes()).hexdigest()))
    d["f"] = R.f(h, pro_l)                                   # gen.f (compute_inbreedings)
    if len(set(pro_l)) == len(pro_l):                        # gen.phi(sparse=True): CSR of the lower triangle, float32
        d["sp_data"], d["sp_indices"], d["sp_indptr"] = R.sparse(h, pro_l)
    if with_gc:
        gc = R.gc(h, pro_l, anc_l)
        d["gc_sha256"] = np.bytes_(hashlib.sha256(gc.tobytes()).hexdigest())
        d["gc_shape"] = np.asarray(gc.shape, dtype=np.int64)
        if gc_sparse:
            r, c = np.nonzero(gc)
            d["gc_r"], d["gc_c"], d["gc_v"] = r.astype(np.int32), c.astype(np.int32), gc[r, c]
        else:
            d["gc"] = gc
    R.free(h)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("%-28s n=%6d m=%5d cuts=%s" % (name, len(rows), len(pro_l), list(d["cut_sizes"])[:24]))


def main():
    os.makedirs(OUT, exist_ok=True)
    R = Reference()
    ji = read_csv4(REF_DATA + "geneaJi.csv")
    g140 = read_csv4(REF_DATA + "genea140.csv")
    pop = [int(l.split()[0]) for l in open(REF_DATA + "pop140.csv").read().splitlines()[1:] if l.strip()]

    emit(R, "geneaJi", ji)
    emit(R, "geneaJi_pair", ji, pro=[1, 2])
    emit(R, "geneaJi_dup_anc", ji, pro=[29, 1, 28, 1, 13, 2, 17], anc=[17, 13, 25, 1, 17])
    custom = np.asarray([[1, 0, 0, 1], [2, 0, 0, 2], [3, 0, 0, 1], [4, 1, 2, 2], [5, 1, 2, 2],
                         [6, 0, 0, 1], [7, 3, 4, 2], [8, 3, 4, 1], [9, 6, 5, 1], [10, 6, 5, 2]],
                        dtype=np.int32)                      # README / test/test_custom.py
    emit(R, "custom10", custom)
    emit(R, "founders_only", custom, pro=[1, 2, 3, 3, 6])     # single cut: 0.5*I even for duplicates
    emit(R, "genea140", g140, gc_sparse=True)
    emit(R, "genea140_pop140", g140, pro=pop, gc_sparse=True)   # BASELINE config 2
    for seed in range(6):
        emit(R, "overlap25_s%d" % seed, overlapping_pedigree(25, seed))
    for seed in (1, 2, 3):
        rows = relabel_and_shuffle(overlapping_pedigree(400, seed), 100 + seed)
        emit(R, "overlap400_s%d" % seed, rows)
    rows = relabel_and_shuffle(overlapping_pedigree(400, 7), 107)
    rng = np.random.default_rng(7)
    pick = [int(x) for x in rng.choice(rows[:, 0], size=60, replace=True)]   # ancestors + duplicates
    emit(R, "overlap400_explicit", rows, pro=pick, anc=[int(x) for x in rng.choice(rows[:, 0], 40)])
    deep = narrow_deep_pedigree(14, 45, 11)
    last = [int(x) for x in deep[-14:, 0]]
    emit(R, "deep45_sorted", deep, pro=last, sorted_flag=True, with_gc=False)
    deep_s = relabel_and_shuffle(narrow_deep_pedigree(14, 45, 12), 212)
    emit(R, "deep45_shuffled", deep_s, with_gc=False)
    deep30 = relabel_and_shuffle(narrow_deep_pedigree(40, 30, 13), 213)
    emit(R, "deep30_w40", deep30, with_gc=False)
    gc_deep = narrow_deep_pedigree(14, 18, 14)
    emit(R, "gcdeep18", relabel_and_shuffle(gc_deep, 214))


if __name__ == "__main__":
    main()
