"""Test helper: a numpy interpreter of the STEP PLAN (DESIGN.md "Plan layout").

It executes on the CPU exactly what phi_step_kernel + phi_diag_kernel do with the plan arrays
exposed by gk_phi_step_view: phase 1 (N = 1/2 (A[F] + A[M]) per new class), phase 2
(B[j][i] = 1/2 (N[i][F_j] + N[i][M_j]), evaluated only where position j <= position i and mirrored),
then the self-kinship / "same individual" corrections -- so the host planner can be checked against
the reference fixtures without a GPU.  It is NOT a product path (it lives under tests/)."""
import ctypes as C

import numpy as np

from geneakit_b200 import _lib
from geneakit_b200._lib import lib, check

SINGLETON, KEPT, BOTH = 1, 2, 4


def _arr(ptr, n):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=np.int32)
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


def step_view(job, g):
    v = _lib.gk_step_view()
    check(lib.gk_phi_step_view(job._h, g, C.byref(v)))
    K, Kpad, ng = v.K, v.Kpad, v.ngroups
    d = dict(K=K, Kpad=Kpad, K_prev=v.K_prev, Kpad_prev=v.Kpad_prev, np=v.np, panel=v.panel,
             ranked=bool(v.ranked), panel_rows=v.panel_rows, panel_begin=v.panel_begin, panel_end=v.panel_end)
    d["pF"], d["pM"], d["flags"] = _arr(v.pF, Kpad), _arr(v.pM, Kpad), _arr(v.flags, Kpad)
    d["cls_ind"], d["cls_size"] = _arr(v.cls_ind, K), _arr(v.cls_size, K)
    d["pg_i"], d["pg_j"], d["pg_off"] = _arr(v.pg_i, ng), _arr(v.pg_j, ng), _arr(v.pg_off, ng + 1 if ng else 0)
    nt = int(d["pg_off"][-1]) if ng else 0
    d["pt_cls"], d["pt_mult"] = _arr(v.pt_cls, nt), _arr(v.pt_mult, nt)
    return d


def emulate_step(v, A, dprev):
    """One cut: (A, dprev) of the previous cut -> (B, dnew); A is Kpad_prev x Kpad_prev."""
    K, Kpad = v["K"], v["Kpad"]
    assert A.shape == (v["Kpad_prev"], v["Kpad_prev"]) and v["np"] * v["panel"] == Kpad
    F, M, fl = v["pF"], v["pM"], v["flags"]
    assert np.all(F[K:] == -1) and np.all(M[K:] == -1) and F.max(initial=-1) < v["K_prev"] and M.max(initial=-1) < v["K_prev"]
    Aext = np.vstack([A, np.zeros((1, A.shape[1]))])                     # row -1 = zeros (unknown parent)
    N = 0.5 * (Aext[F] + Aext[M])                                         # phase 1: Kpad x Kpad_prev
    Next = np.hstack([N, np.zeros((Kpad, 1))])
    V = 0.5 * (Next[:, F].T + Next[:, M].T)                               # V[j][i] = 1/2 (N[i][F_j] + N[i][M_j])
    B = np.triu(V) + np.triu(V, 1).T                                      # valid where j <= i; mirrored
    dnew = np.full(Kpad, 0.5)
    for c in range(K):
        if fl[c] & KEPT:
            assert F[c] == M[c] and F[c] >= 0
            dnew[c] = dprev[F[c]]
        elif fl[c] & BOTH:
            dnew[c] = 0.5 + 0.5 * A[F[c], M[c]]
        if fl[c] & SINGLETON:
            B[c, c] = dnew[c]
    for k in range(len(v["pg_i"])):
        i, j = int(v["pg_i"][k]), int(v["pg_j"][k])
        assert i >= j and not (i == j and (fl[i] & SINGLETON))
        add = 0.0
        for t in range(v["pg_off"][k], v["pg_off"][k + 1]):
            c = int(v["pt_cls"][t])
            add += 0.25 * float(v["pt_mult"][t]) * (dprev[c] - A[c, c])
        B[i, j] += add
        if i != j:
            B[j, i] += add
    return B, dnew


def emulate(job):
    st = job.stats()
    G, m = st["n_cuts"], st["m"]
    cls_p, ind_p, klast = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.c_int64()
    check(lib.gk_phi_final_view(job._h, C.byref(cls_p), C.byref(ind_p), C.byref(klast)))
    fcls, find = _arr(cls_p, m), _arr(ind_p, m)
    if G == 1:
        return 0.5 * np.eye(m)
    v0 = step_view(job, 0)
    A = np.zeros((v0["Kpad"], v0["Kpad"]))
    for c in range(v0["K"]):
        if v0["flags"][c] & SINGLETON:
            A[c, c] = 0.5
    d = np.full(v0["Kpad"], 0.5)
    for g in range(1, G):
        A, d = emulate_step(step_view(job, g), A, d)
    out = A[np.ix_(fcls, fcls)]
    same = find[:, None] == find[None, :]
    return np.where(same, d[fcls][:, None] * np.ones((1, m)), out)


def emulate_step_rows(v, A, dprev):
    """Sharded mode (a rank of a multi-GPU job): only the rows of this job's own panels, every tile J
    evaluated, nothing mirrored: returns (c0, c1, B[c0:c1, :], dnew[c0:c1])."""
    K, Kpad, T = v["K"], v["Kpad"], v["panel"]
    c0, c1 = v["panel_begin"] * T, v["panel_end"] * T
    F, M, fl = v["pF"], v["pM"], v["flags"]
    Aext = np.vstack([A, np.zeros((1, A.shape[1]))])
    N = 0.5 * (Aext[F[c0:c1]] + Aext[M[c0:c1]])                          # phase 1 of the own panels (rows may be remote)
    Next = np.hstack([N, np.zeros((c1 - c0, 1))])
    B = 0.5 * (Next[:, F] + Next[:, M])                                   # B[i][j] = 1/2 (N[i][F_j] + N[i][M_j]), every j
    dnew = np.full(c1 - c0, 0.5)
    for c in range(c0, min(c1, K)):
        if fl[c] & KEPT:
            dnew[c - c0] = dprev[F[c]]
        elif fl[c] & BOTH:
            dnew[c - c0] = 0.5 + 0.5 * A[F[c], M[c]]
        if fl[c] & SINGLETON:
            B[c - c0, c] = dnew[c - c0]
    for k in range(len(v["pg_i"])):
        i, j = int(v["pg_i"][k]), int(v["pg_j"][k])
        add = 0.0
        for t in range(v["pg_off"][k], v["pg_off"][k + 1]):
            c = int(v["pt_cls"][t])
            add += 0.25 * float(v["pt_mult"][t]) * (dprev[c] - A[c, c])
        if c0 <= i < c1:
            B[i - c0, j] += add
        if i != j and c0 <= j < c1:
            B[j - c0, i] += add
    return c0, c1, B, dnew
