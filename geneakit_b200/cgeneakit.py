"""Host-side mirror of the reference's `cgeneakit` binding for the kinship path.

Same function names, argument order and error behaviour as the nanobind module
(reference geneakit/lib/cgeneakit.cpp) for the entry points that sit on the hot path:

    load_pedigree_from_file / load_pedigree_from_vectors   (:201-205)
    get_proband_ids / get_founder_ids                      (:233-236)
    get_required_memory_for_kinships                       (:408-409)
    compute_kinships                                       (:411-431)
    compute_genetic_contributions                          (:491-511)

The pedigree is held flat, in RANK order (the order of Pedigree::ids, pedigree.hpp:34-44),
which is exactly what the C ABI consumes; every compute call goes to libgeneakit_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check, make_opts


class Pedigree:
    """Flat rank-space pedigree: ids[k], sire[k] / dam[k] (parent rank, -1 unknown), sex[k]."""

    def __init__(self, ids, sire, dam, sex):
        self.ids = np.ascontiguousarray(ids, dtype=np.int32)
        self.sire = np.ascontiguousarray(sire, dtype=np.int32)
        self.dam = np.ascontiguousarray(dam, dtype=np.int32)
        self.sex = np.ascontiguousarray(sex, dtype=np.int32)
        order = np.argsort(self.ids, kind="stable")
        self._sorted_ids = self.ids[order]
        self._sorted_rank = order.astype(np.int32)

    def __len__(self):
        return len(self.ids)

    def ranks(self, id_list):
        """ids -> ranks.  Unknown id -> IndexError, as `individuals.at(id)` does upstream
        (lib/compute.cpp:646-650 through nanobind's std::out_of_range mapping)."""
        q = np.asarray(list(id_list), dtype=np.int64)
        n = len(self.ids)
        pos = np.clip(np.searchsorted(self._sorted_ids, q), 0, max(n - 1, 0))
        if n == 0 or np.any(self._sorted_ids[pos] != q):
            raise IndexError("unordered_map::at")
        return np.ascontiguousarray(self._sorted_rank[pos])

    def __repr__(self):
        n = len(self.ids)
        links = int(np.count_nonzero(self.sire >= 0) + np.count_nonzero(self.dam >= 0))
        return ("A pedigree with:\n - %d individuals;\n - %d parent-child relations;\n - %d men;\n"
                " - %d women;\n - %d probands." % (n, links, int(np.count_nonzero(self.sex == 1)),
                                                 int(np.count_nonzero(self.sex == 2)),
                                                 len(get_proband_ids(self))))


def load_pedigree_from_vectors(ids, father_ids, mother_ids, sexes, sorted=False):
    """lib/create.cpp:211-218 (+ convert_pedigree :158-199).  Parent id 0 = unknown."""
    ids = np.asarray(ids, dtype=np.int64)
    fathers = np.asarray(father_ids, dtype=np.int64)
    mothers = np.asarray(mother_ids, dtype=np.int64)
    sexes = np.asarray(sexes, dtype=np.int64)
    n = len(ids)
    o = np.argsort(ids, kind="stable")
    sid = ids[o]

    def positions(par):
        pos = np.clip(np.searchsorted(sid, par), 0, max(n - 1, 0))
        known = par != 0
        if n and np.any(known & (sid[pos] != par)):
            raise IndexError("unordered_map::at")          # parent id absent from the pedigree
        return np.ascontiguousarray(np.where(known, o[pos], -1).astype(np.int32))

    fpos, mpos = positions(fathers), positions(mothers)
    if sorted:
        order = np.arange(n, dtype=np.int32)
        if np.any(fpos >= order) or np.any(mpos >= order):
            raise IndexError("unordered_map::at")          # child listed before a parent
    else:
        order = np.empty(n, dtype=np.int32)
        check(lib.gk_dfs_order(n, fpos, mpos, order))
    rank_of = np.empty(n, dtype=np.int32)
    rank_of[order] = np.arange(n, dtype=np.int32)
    fo, mo = fpos[order], mpos[order]
    sire = np.where(fo >= 0, rank_of[np.maximum(fo, 0)], -1)
    dam = np.where(mo >= 0, rank_of[np.maximum(mo, 0)], -1)
    return Pedigree(ids[order], sire, dam, sexes[order])


def load_pedigree_from_file(pedigree_file, sorted=False):
    """lib/create.cpp:202-208: header skipped, four unsigned integers per line separated by
    any run of non-digits (:30-46)."""
    with open(pedigree_file, "rb") as fh:
        fh.readline()
        body = fh.read()
    table = bytes(c if 48 <= c <= 57 or c == 10 else 32 for c in range(256))
    rows = []
    for line in body.translate(table).split(b"\n"):
        t = line.split()
        if not t:
            continue
        if len(t) < 4:
            raise ValueError("Invalid line format in pedigree file: " + line.decode(errors="replace"))
        rows.append((int(t[0]), int(t[1]), int(t[2]), int(t[3])))
    a = np.asarray(rows, dtype=np.int64).reshape(-1, 4)
    return load_pedigree_from_vectors(a[:, 0], a[:, 1], a[:, 2], a[:, 3], sorted)


def get_proband_ids(pedigree):
    """lib/identify.cpp:55-65: childless individuals, sorted by id."""
    n = len(pedigree)
    out = np.empty(max(n, 1), dtype=np.int32)
    c = lib.gk_childless(n, pedigree.sire, pedigree.dam, out.ctypes.data)
    return [int(x) for x in np.sort(pedigree.ids[out[:c]])]


def get_founder_ids(pedigree):
    """lib/identify.cpp:28-38: individuals with no known parent, sorted by id."""
    n = len(pedigree)
    out = np.empty(max(n, 1), dtype=np.int32)
    c = lib.gk_parentless(n, pedigree.sire, pedigree.dam, out.ctypes.data)
    return [int(x) for x in np.sort(pedigree.ids[out[:c]])]


def _pro_ranks(pedigree, proband_ids):
    if proband_ids is None or len(proband_ids) == 0:
        proband_ids = get_proband_ids(pedigree)            # lib/compute.cpp:641-643
    return pedigree.ranks(proband_ids)


def get_required_memory_for_kinships(pedigree, proband_ids=()):
    r = _pro_ranks(pedigree, proband_ids)
    gb = lib.gk_phi_required_gb(len(pedigree), pedigree.sire, pedigree.dam, len(r), r)
    if gb < 0:
        raise ValueError(lib.gk_last_error().decode())
    return float(gb)


def _pinned_matrix(rows, cols):
    """Result storage owned by the array, like the nb::capsule of lib/cgeneakit.cpp:416-429,
    but page-locked so the device->host copy runs at PCIe rate."""
    nbytes = rows * cols * 8
    ptr = lib.gk_host_alloc(nbytes)
    if not ptr:
        raise _lib.GkError("cannot allocate %d bytes of pinned host memory (%s); geneakit_b200 needs a "
                           "CUDA device, there is no CPU fallback" % (nbytes, lib.gk_last_error().decode()))
    buf = (C.c_double * (rows * cols)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.float64).reshape(rows, cols)
    owner = _PinnedOwner(ptr)
    _OWNERS[id(buf)] = owner
    import weakref
    weakref.finalize(buf, _release, id(buf))
    return arr, ptr


class _PinnedOwner:
    def __init__(self, ptr):
        self.ptr = ptr


_OWNERS = {}


def _release(key):
    o = _OWNERS.pop(key, None)
    if o is not None:
        lib.gk_host_free(o.ptr)


def compute_kinships(pedigree, proband_ids=(), verbose=False, **gk):
    """lib/cgeneakit.cpp:411-431 -> gk_phi_f64.  Returns an (m, m) float64 ndarray."""
    r = _pro_ranks(pedigree, proband_ids)
    m = len(r)
    out, ptr = _pinned_matrix(m, m)
    opts = make_opts(verbose=verbose, **gk)
    check(lib.gk_phi_f64(len(pedigree), pedigree.sire, pedigree.dam, m, r, ptr, None, C.byref(opts)))
    return out


def compute_inbreedings(pedigree, proband_ids=(), **gk):
    """lib/cgeneakit.cpp (compute_inbreedings binding) -> gk_f_f64: F of every proband, caller order."""
    r = _pro_ranks(pedigree, proband_ids)
    out = np.empty(len(r), dtype=np.float64)
    opts = make_opts(**gk)
    check(lib.gk_f_f64(len(pedigree), pedigree.sire, pedigree.dam, len(r), r, out.ctypes.data, C.byref(opts)))
    return out


def compute_kinships_sparse(pedigree, proband_ids=(), verbose=False, **gk):
    """lib/cgeneakit.cpp:433-486 (CSR branch): (data float32, indices int32, indptr) of the lower-triangular kinship matrix,
    compacted on the device from the dense result."""
    job = DevicePhi(pedigree, proband_ids, verbose, **gk).upload().run()
    try:
        return job.sparse_csr()
    finally:
        job.close()


def compute_genetic_contributions(pedigree, proband_ids=(), ancestor_ids=(), **gk):
    """lib/cgeneakit.cpp:491-511 -> gk_gc_f64.  Rows = probands, columns = ancestors."""
    r = _pro_ranks(pedigree, proband_ids)
    if ancestor_ids is None or len(ancestor_ids) == 0:
        ancestor_ids = get_founder_ids(pedigree)            # lib/compute.cpp:837-839
    a = pedigree.ranks(ancestor_ids)
    out, ptr = _pinned_matrix(len(r), len(a))
    opts = make_opts(**gk)
    check(lib.gk_gc_f64(len(pedigree), pedigree.sire, pedigree.dam, len(r), r, len(a), a, ptr, C.byref(opts)))
    return out


class DevicePhi:
    """Device-resident kinship job (gk_phi_plan/upload/run): the matrix stays in HBM; the mean
    comes from the fused reduction; `to_host()` is the only PCIe transfer."""

    def __init__(self, pedigree, proband_ids=(), verbose=False, **gk):
        self._h = C.c_void_p()
        r = _pro_ranks(pedigree, proband_ids)
        self.m = len(r)
        self._opts = make_opts(verbose=verbose, **gk)
        r0, r1 = C.c_int64(), C.c_int64()
        lib.gk_shard_rows(self.m, self._opts.rank, max(self._opts.world, 1), C.byref(r0), C.byref(r1))
        self.row0, self.row1 = r0.value, r1.value          # result rows held by this job (multi-GPU shard)
        check(lib.gk_phi_plan(len(pedigree), pedigree.sire, pedigree.dam, self.m, r,
                              C.byref(self._opts), C.byref(self._h)))

    def upload(self):
        check(lib.gk_phi_upload(self._h)); return self

    def run(self):
        check(lib.gk_phi_run(self._h)); return self

    def sync(self):
        check(lib.gk_phi_sync(self._h)); return self

    def mean(self):
        s, d, mu = C.c_double(), C.c_double(), C.c_double()
        check(lib.gk_phi_mean(self._h, C.byref(s), C.byref(d), C.byref(mu)))
        return mu.value

    def sums(self):
        s, d, mu = C.c_double(), C.c_double(), C.c_double()
        check(lib.gk_phi_mean(self._h, C.byref(s), C.byref(d), C.byref(mu)))
        return s.value, d.value

    def to_host(self, out=None):
        if out is None:
            out, ptr = _pinned_matrix(self.row1 - self.row0, self.m)
        else:
            ptr = out.ctypes.data
        check(lib.gk_phi_copy_out(self._h, ptr))
        return out

    def replan(self):
        """Plan again on the host and copy the plan to the device again (buffers stay allocated)."""
        check(lib.gk_phi_replan(self._h)); return self

    def run_pass(self):
        """One whole pass from the job's inputs -- plan, copy, every cut, expansion -- with the host planning the later
        cuts while the device runs the earlier ones (gk_phi_pass; asynchronous like run())."""
        check(lib.gk_phi_pass(self._h)); return self

    def diag(self):
        out = np.empty(self.m, dtype=np.float64)
        check(lib.gk_phi_diag(self._h, out.ctypes.data))
        return out

    def f(self):
        """Inbreeding coefficients 2 phi(i, i) - 1 of the probands (gen.f), from the device's self kinships."""
        out = np.empty(self.m, dtype=np.float64)
        check(lib.gk_phi_f(self._h, out.ctypes.data))
        return out

    def over(self, threshold):
        """gen.phiOver on the device: (row index, column index, kinship) of the pairs i > j above the threshold."""
        n = C.c_int64()
        check(lib.gk_phi_over(self._h, float(threshold), C.byref(n)))
        i = np.empty(max(n.value, 1), dtype=np.int32); j = np.empty(max(n.value, 1), dtype=np.int32)
        v = np.empty(max(n.value, 1), dtype=np.float64)
        check(lib.gk_phi_over_fetch(self._h, i.ctypes.data, j.ctypes.data, v.ctypes.data))
        return i[:n.value], j[:n.value], v[:n.value]

    def sparse_csr(self):
        """(data float32, indices int32, indptr int64): the CSR of gen.phi(sparse=True), lower triangle with the diagonal."""
        n = C.c_int64()
        check(lib.gk_phi_sparse(self._h, C.byref(n)))
        indptr = np.empty(self.m + 1, dtype=np.int64)
        indices = np.empty(max(n.value, 1), dtype=np.int32); data = np.empty(max(n.value, 1), dtype=np.float32)
        check(lib.gk_phi_sparse_fetch(self._h, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data))
        return data[:n.value], indices[:n.value], indptr

    def ci_stats(self, idx):
        """gen.phiCI's statistic for every row of `idx` (b x n resample indices): mean of the entries < 0.5 of phi[idx][:, idx]."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        b, n = idx.shape
        out = np.empty(b, dtype=np.float64)
        check(lib.gk_phi_ci_stats(self._h, idx.reshape(-1), b, n, out.ctypes.data))
        return out

    def rows(self, row_begin, nrows):
        """Rows [row_begin, row_begin + nrows) of the result, copied to the host (a slice, not the M x M matrix)."""
        out = np.empty((nrows, self.m), dtype=np.float64)
        check(lib.gk_phi_copy_rows(self._h, int(row_begin), int(nrows), out.ctypes.data))
        return out

    def device_pointer(self):
        p, ld = C.c_void_p(), C.c_int64()
        check(lib.gk_phi_device_result(self._h, C.byref(p), C.byref(ld)))
        return p.value, ld.value

    def stats(self):
        s = _lib.gk_phi_stats()
        check(lib.gk_phi_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def cut_sizes(self):
        g = self.stats()["n_cuts"]
        a = np.zeros(g, dtype=np.int32)
        check(lib.gk_phi_cut_sizes(self._h, a, g))
        return [int(x) for x in a]

    def class_sizes(self):
        g = self.stats()["n_cuts"]
        a = np.zeros(g, dtype=np.int32)
        check(lib.gk_phi_class_sizes(self._h, a, g))
        return [int(x) for x in a]

    def step_times_ms(self):
        g = self.stats()["n_cuts"]
        a = np.zeros(max(g, 1), dtype=np.float32)
        check(lib.gk_phi_step_times(self._h, a, g))
        return [float(x) for x in a[:g]]

    def step_bytes(self):
        g = self.stats()["n_cuts"]
        a = np.zeros(max(g, 1), dtype=np.int64)
        check(lib.gk_phi_step_bytes(self._h, a, g))
        return [int(x) for x in a[:g]]

    def close(self):
        if self._h:
            lib.gk_phi_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
