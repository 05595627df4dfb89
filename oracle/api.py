"""TEST INFRASTRUCTURE ONLY: ctypes front-ends for the CPU oracle.

* ``Oracle``  -> oracle/libgkoracle.so, our plain-C restatement (oracle/gk_oracle.c).
* ``Reference`` -> oracle/_ref/libgkref.so, the UNMODIFIED reference C++ behind
  oracle/ref_harness.cpp (only present where it was built from /root/reference).

Nothing in geneakit_b200/ imports this module; it is the checker used by tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_I = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_D = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(verbose=False):
    """Compile libgkoracle.so (always) and _ref/libgkref.so (when the reference tree exists)."""
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64).astype(np.int32))


class FlatPedigree:
    """Rank-space pedigree: ids[k], sire[k], dam[k] (parent RANK or -1), sex[k]."""

    def __init__(self, ids, sire, dam, sex):
        self.ids = _i32(ids)
        self.sire = _i32(sire)
        self.dam = _i32(dam)
        self.sex = _i32(sex)
        self.n = len(self.ids)
        order = np.argsort(self.ids, kind="stable")
        self._sorted_ids = self.ids[order]
        self._sorted_rank = order.astype(np.int32)

    def ranks(self, id_list):
        """ids -> ranks; IndexError on an unknown id (reference: std::out_of_range from .at(),
        compute.cpp:649)."""
        q = np.asarray(id_list, dtype=np.int64)
        pos = np.searchsorted(self._sorted_ids, q)
        pos = np.clip(pos, 0, max(self.n - 1, 0))
        if self.n == 0 or np.any(self._sorted_ids[pos] != q):
            raise IndexError("unknown individual id")
        return np.ascontiguousarray(self._sorted_rank[pos])


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "libgkoracle.so")
        if not os.path.exists(path):
            build()
        L = self.L = C.CDLL(path)
        L.gko_dfs_order.argtypes = [C.c_int, _I, _I, _I]
        L.gko_childless.argtypes = [C.c_int, _I, _I, _I]
        L.gko_parentless.argtypes = [C.c_int, _I, _I, _I]
        L.gko_cut_sizes.argtypes = [C.c_int, _I, _I, C.c_int, _I, _I, C.c_int]
        L.gko_required_gb.argtypes = [C.c_int, _I, _I, C.c_int, _I]
        L.gko_required_gb.restype = C.c_double
        L.gko_phi.argtypes = [C.c_int, _I, _I, C.c_int, _I, _D]
        L.gko_gc.argtypes = [C.c_int, _I, _I, C.c_int, _I, C.c_int, _I, _D]
        L.gko_gc_dfs.argtypes = [C.c_int, _I, _I, C.c_int, _I, C.c_int, _I, _D]
        L.gko_set_threads.argtypes = [C.c_int]

    # --- loader (create.cpp:158-218) -------------------------------------------------------
    def genealogy(self, ids, fathers, mothers, sexes, sorted=False):
        ids = np.asarray(ids, dtype=np.int64)
        fathers = np.asarray(fathers, dtype=np.int64)
        mothers = np.asarray(mothers, dtype=np.int64)
        sexes = np.asarray(sexes, dtype=np.int64)
        n = len(ids)
        o = np.argsort(ids, kind="stable")
        sid = ids[o]

        def to_pos(par):
            pos = np.searchsorted(sid, par)
            pos = np.clip(pos, 0, max(n - 1, 0))
            known = par != 0
            if n and np.any(known & (sid[pos] != par)):
                raise IndexError("parent id not in pedigree")
            return np.where(known, o[pos], -1).astype(np.int32)

        fpos, mpos = to_pos(fathers), to_pos(mothers)
        if sorted:
            order = np.arange(n, dtype=np.int32)
            if np.any(fpos >= order) or np.any(mpos >= order):
                raise IndexError("sorted=True but a child precedes a parent")
        else:
            order = np.empty(n, dtype=np.int32)
            rc = self.L.gko_dfs_order(n, np.ascontiguousarray(fpos), np.ascontiguousarray(mpos), order)
            if rc != 0:
                raise ValueError("pedigree is not a DAG (rc=%d)" % rc)
        rank_of = np.empty(n, dtype=np.int32)
        rank_of[order] = np.arange(n, dtype=np.int32)
        sire = np.where(fpos[order] >= 0, rank_of[np.maximum(fpos[order], 0)], -1)
        dam = np.where(mpos[order] >= 0, rank_of[np.maximum(mpos[order], 0)], -1)
        return FlatPedigree(ids[order], sire, dam, sexes[order])

    def genealogy_file(self, path, sorted=False):
        rows = []
        with open(path) as fh:
            next(fh)
            for line in fh:
                # create.cpp:30-46: any non-digit run separates the four integers
                cur, vals = "", []
                for ch in line:
                    if ch.isdigit():
                        cur += ch
                    elif cur:
                        vals.append(int(cur)); cur = ""
                if cur:
                    vals.append(int(cur))
                if len(vals) < 4:
                    if not line.strip():
                        continue
                    raise ValueError("Invalid line format in pedigree file: " + line)
                rows.append(vals[:4])
        a = np.asarray(rows, dtype=np.int64).reshape(-1, 4)
        return self.genealogy(a[:, 0], a[:, 1], a[:, 2], a[:, 3], sorted=sorted)

    # --- identify.cpp:28-38,55-65 ------------------------------------------------------------
    def pro(self, ped):
        out = np.empty(ped.n, dtype=np.int32)
        c = self.L.gko_childless(ped.n, ped.sire, ped.dam, out)
        return sorted(int(x) for x in ped.ids[out[:c]])

    def founder(self, ped):
        out = np.empty(ped.n, dtype=np.int32)
        c = self.L.gko_parentless(ped.n, ped.sire, ped.dam, out)
        return sorted(int(x) for x in ped.ids[out[:c]])

    # --- compute.cpp ---------------------------------------------------------------------------
    def phi(self, ped, pro=None):
        pro = self.pro(ped) if pro is None or len(pro) == 0 else list(pro)
        r = ped.ranks(pro)
        out = np.empty((len(r), len(r)), dtype=np.float64)
        rc = self.L.gko_phi(ped.n, ped.sire, ped.dam, len(r), r, out)
        if rc != 0:
            raise RuntimeError("gko_phi rc=%d" % rc)
        return out

    def gc(self, ped, pro=None, ancestors=None, dfs=False):
        pro = self.pro(ped) if pro is None or len(pro) == 0 else list(pro)
        anc = self.founder(ped) if ancestors is None or len(ancestors) == 0 else list(ancestors)
        r, a = ped.ranks(pro), ped.ranks(anc)
        out = np.empty((len(r), len(a)), dtype=np.float64)
        fn = self.L.gko_gc_dfs if dfs else self.L.gko_gc
        rc = fn(ped.n, ped.sire, ped.dam, len(r), r, len(a), a, out)
        if rc != 0:
            raise RuntimeError("gko_gc rc=%d" % rc)
        return out

    def cut_sizes(self, ped, pro=None):
        pro = self.pro(ped) if pro is None or len(pro) == 0 else list(pro)
        r = ped.ranks(pro)
        sizes = np.zeros(4096, dtype=np.int32)
        g = self.L.gko_cut_sizes(ped.n, ped.sire, ped.dam, len(r), r, sizes, len(sizes))
        if g < 0:
            raise RuntimeError("gko_cut_sizes rc=%d" % g)
        return [int(x) for x in sizes[:g]]

    def required_gb(self, ped, pro=None):
        pro = self.pro(ped) if pro is None or len(pro) == 0 else list(pro)
        r = ped.ranks(pro)
        return float(self.L.gko_required_gb(ped.n, ped.sire, ped.dam, len(r), r))

    def set_threads(self, t):
        self.L.gko_set_threads(int(t))


def phi_mean(mat):
    """compute.py:124-132 dense branch on a plain ndarray: column sums, then their sum (what
    DataFrame.sum().sum() does), minus the diagonal, over n^2 - n."""
    mat = np.asarray(mat, dtype=np.float64)
    total = mat.sum(axis=0).sum()
    diag = np.diag(mat).sum()
    n = mat.shape[0]
    return (total - diag) / (n ** 2 - n)


class Reference:
    """The unmodified reference core (oracle/_ref/libgkref.so)."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libgkref.so"))

    def __init__(self):
        L = self.L = C.CDLL(os.path.join(HERE, "_ref", "libgkref.so"))
        L.gkref_load_vectors.argtypes = [C.c_int, _I, _I, _I, _I, C.c_int]
        L.gkref_load_vectors.restype = C.c_void_p
        L.gkref_load_file.argtypes = [C.c_char_p, C.c_int]
        L.gkref_load_file.restype = C.c_void_p
        L.gkref_free.argtypes = [C.c_void_p]
        L.gkref_size.argtypes = [C.c_void_p]
        L.gkref_flatten.argtypes = [C.c_void_p, _I, _I, _I, _I]
        L.gkref_proband_ids.argtypes = [C.c_void_p, C.c_void_p]
        L.gkref_founder_ids.argtypes = [C.c_void_p, C.c_void_p]
        L.gkref_phi.argtypes = [C.c_void_p, C.c_int, _I, _D, C.c_int]
        L.gkref_gc.argtypes = [C.c_void_p, C.c_int, _I, C.c_int, _I, _D]
        L.gkref_required_gb.argtypes = [C.c_void_p, C.c_int, _I]
        L.gkref_required_gb.restype = C.c_double
        L.gkref_cut_sizes.argtypes = [C.c_void_p, C.c_int, _I, _I, C.c_int]
        L.gkref_set_threads.argtypes = [C.c_int]
        L.gkref_last_error.restype = C.c_char_p
        if hasattr(L, "gkref_f"):
            L.gkref_f.argtypes = [C.c_void_p, C.c_int, _I, _D]
            L.gkref_sparse.argtypes = [C.c_void_p, C.c_int, _I, C.c_void_p, C.c_void_p, C.c_void_p]
            L.gkref_sparse.restype = C.c_longlong

    def genealogy(self, ids, fathers, mothers, sexes, sorted=False):
        h = self.L.gkref_load_vectors(len(ids), _i32(ids), _i32(fathers), _i32(mothers), _i32(sexes),
                                      int(sorted))
        if not h:
            raise IndexError(self.L.gkref_last_error().decode())
        return h

    def genealogy_file(self, path, sorted=False):
        h = self.L.gkref_load_file(path.encode(), int(sorted))
        if not h:
            raise IndexError(self.L.gkref_last_error().decode())
        return h

    def free(self, h):
        self.L.gkref_free(h)

    def flatten(self, h):
        n = self.L.gkref_size(h)
        a, b, c, d = (np.empty(n, dtype=np.int32) for _ in range(4))
        self.L.gkref_flatten(h, a, b, c, d)
        return FlatPedigree(a, b, c, d)

    def pro(self, h):
        n = self.L.gkref_size(h)
        out = np.empty(n, dtype=np.int32)
        c = self.L.gkref_proband_ids(h, out.ctypes.data)
        return [int(x) for x in out[:c]]

    def founder(self, h):
        n = self.L.gkref_size(h)
        out = np.empty(n, dtype=np.int32)
        c = self.L.gkref_founder_ids(h, out.ctypes.data)
        return [int(x) for x in out[:c]]

    def phi(self, h, pro=None, verbose=False):
        pro = self.pro(h) if pro is None or len(pro) == 0 else list(pro)
        p = _i32(pro)
        out = np.empty((len(p), len(p)), dtype=np.float64)
        rc = self.L.gkref_phi(h, len(p), p, out, int(verbose))
        if rc == 1:
            raise IndexError(self.L.gkref_last_error().decode())
        if rc:
            raise RuntimeError(self.L.gkref_last_error().decode())
        return out

    def gc(self, h, pro=None, ancestors=None):
        pro = self.pro(h) if pro is None or len(pro) == 0 else list(pro)
        anc = self.founder(h) if ancestors is None or len(ancestors) == 0 else list(ancestors)
        p, a = _i32(pro), _i32(anc)
        out = np.empty((len(p), len(a)), dtype=np.float64)
        rc = self.L.gkref_gc(h, len(p), p, len(a), a, out)
        if rc == 1:
            raise IndexError(self.L.gkref_last_error().decode())
        if rc:
            raise RuntimeError(self.L.gkref_last_error().decode())
        return out

    def f(self, h, pro=None):
        """gen.f: the reference's compute_inbreedings."""
        pro = self.pro(h) if pro is None or len(pro) == 0 else list(pro)
        p = _i32(pro)
        out = np.empty(len(p), dtype=np.float64)
        if self.L.gkref_f(h, len(p), p, out):
            raise RuntimeError(self.L.gkref_last_error().decode())
        return out

    def sparse(self, h, pro=None):
        """gen.phi(sparse=True): (data float32, indices int32, indptr int64) of the reference's CSR."""
        pro = self.pro(h) if pro is None or len(pro) == 0 else list(pro)
        p = _i32(pro)
        nnz = self.L.gkref_sparse(h, len(p), p, None, None, None)
        if nnz < 0:
            raise RuntimeError(self.L.gkref_last_error().decode())
        data = np.empty(max(nnz, 1), dtype=np.float32); idx = np.empty(max(nnz, 1), dtype=np.int32)
        ptr = np.empty(len(p) + 1, dtype=np.int64)
        self.L.gkref_sparse(h, len(p), p, data.ctypes.data, idx.ctypes.data, ptr.ctypes.data)
        return data[:nnz], idx[:nnz], ptr

    def cut_sizes(self, h, pro=None):
        p = _i32([] if pro is None else pro)
        sizes = np.zeros(4096, dtype=np.int32)
        g = self.L.gkref_cut_sizes(h, len(p), p, sizes, len(sizes))
        if g < 0:
            raise RuntimeError(self.L.gkref_last_error().decode())
        return [int(x) for x in sizes[:g]]

    def required_gb(self, h, pro=None):
        p = _i32([] if pro is None else pro)
        return float(self.L.gkref_required_gb(h, len(p), p))

    def set_threads(self, t):
        self.L.gkref_set_threads(int(t))
