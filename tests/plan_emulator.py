"""Test helper: a numpy interpreter of the TILE PLAN (DESIGN.md "Plan layout").

It executes, on the CPU and tile pair by tile pair, exactly what phi_step_kernel does with the
plan arrays exposed by gk_phi_step_view -- including the symmetric fetch of the window x list
block, the conflict-list gated "same individual" rule and the mirror store -- so the host
planner can be checked against the oracle without a GPU.  It is NOT a product path (it lives
under tests/)."""
import ctypes as C

import numpy as np

from geneakit_b200 import _lib
from geneakit_b200._lib import lib, check

MISS = 0xFFFF


def _arr(ptr, n):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=np.int32)
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


def step_view(job, g):
    v = _lib.gk_step_view()
    check(lib.gk_phi_step_view(job._h, g, C.byref(v)))
    nt, W, LMAX, DW, T = v.nt, v.win, v.lmax, v.desc_words, v.tile
    d = dict(K=v.K, ld=v.ld, K_prev=v.K_prev, nt=nt, T=T, W=W, LMAX=LMAX, nr_max=v.nr_max,
             conf_max=v.conf_max, ranked=bool(v.ranked))
    desc = _arr(v.desc, nt * DW).reshape(nt, DW)
    d["c0"], d["nl"], d["cnt"] = desc[:, 0], desc[:, 1] & 0xFFFF, desc[:, 1] >> 16
    d["nconf"], d["start"] = desc[:, 2], desc[:, 3]
    d["list"] = desc[:, v.off_list:v.off_list + LMAX]
    d["conf"] = desc[:, v.off_conf:v.off_conf + v.conf_max]
    d["la"] = desc[:, v.off_la:v.off_la + T]
    d["pind"] = desc[:, v.off_pind:v.off_pind + 2 * T].reshape(nt, T, 2)
    d["rank"] = desc[:, v.off_rank:v.off_rank + T]
    return d


def emulate(job):
    st = job.stats()
    G, m = st["n_cuts"], st["m"]
    cls_p, ind_p, klast = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.c_int64()
    check(lib.gk_phi_final_view(job._h, C.byref(cls_p), C.byref(ind_p), C.byref(klast)))
    fcls, find = _arr(cls_p, m), _arr(ind_p, m)
    if G == 1:
        return 0.5 * np.eye(m)
    k0 = C.c_int64()
    check(lib.gk_phi_initial_classes(job._h, C.byref(k0)))
    Gp = np.zeros((k0.value, k0.value))
    dp = np.full(k0.value, 0.5)
    for g in range(1, G):
        v = step_view(job, g)
        K, T, W, nt = v["K"], v["T"], v["W"], v["nt"]
        assert v["K_prev"] == Gp.shape[0]
        assert int(v["cnt"].sum()) == K and np.array_equal(v["start"], np.concatenate([[0], np.cumsum(v["cnt"])[:-1]]))
        assert v["nl"].max(initial=0) <= v["LMAX"] and v["cnt"].max(initial=0) <= T
        Gn = np.full((K, K), np.nan)
        dn = np.full(K, np.nan)
        last = Gp.shape[0] - 1

        def local_classes(t):
            nl = v["nl"][t]
            win = np.minimum(v["c0"][t] + np.arange(W), last)
            return np.concatenate([win, v["list"][t, :nl]]).astype(np.int64), nl

        for I in range(nt):
            rI, nlI = local_classes(I)
            nI, sI = v["cnt"][I], v["start"][I]
            for J in range(I + 1):
                cJ, nlJ = local_classes(J)
                nJ, sJ = v["cnt"][J], v["start"][J]
                S = np.empty((W + nlI, W + nlJ))
                S[:, :W] = Gp[np.ix_(rI, cJ[:W])]                      # rows x window columns
                S[:W, W:] = Gp[np.ix_(cJ[W:], rI[:W])].T               # window rows x list cols via symmetry
                S[W:, W:] = Gp[np.ix_(rI[W:], cJ[W:])]                 # list x list
                slow = I == J or v["nconf"][I] < 0 or v["nconf"][J] < 0 or \
                    J in [int(x) for x in v["conf"][I, :max(v["nconf"][I], 0)]]
                laI, laJ = v["la"][I, :nI], v["la"][J, :nJ]
                A = [laI & 0xFFFF, (laI >> 16) & 0xFFFF]
                B = [laJ & 0xFFFF, (laJ >> 16) & 0xFFFF]
                IA = [v["pind"][I, :nI, 0], v["pind"][I, :nI, 1]]
                IB = [v["pind"][J, :nJ, 0], v["pind"][J, :nJ, 1]]

                def take(qa, qb):
                    ra, cb = A[qa], B[qb]
                    ok = (ra[:, None] != MISS) & (cb[None, :] != MISS)
                    rr = np.where(ra == MISS, 0, ra)
                    cc = np.where(cb == MISS, 0, cb)
                    assert rr.max(initial=0) < S.shape[0] and cc.max(initial=0) < S.shape[1]
                    val = S[np.ix_(rr, cc)]
                    same = IA[qa][:, None] == IB[qb][None, :]
                    if not slow:
                        assert not (ok & same).any(), "identity needed on a fast-path tile pair"
                    else:
                        val = np.where(same, dp[rI[rr]][:, None], val)
                    return np.where(ok, val, 0.0)

                p, q, r, s = take(0, 0), take(1, 0), take(0, 1), take(1, 1)
                if v["ranked"]:
                    lower = v["rank"][I, :nI][:, None] < v["rank"][J, :nJ][None, :]
                    out = np.where(lower, 0.25 * ((p + q) + (r + s)), 0.25 * ((p + r) + (q + s)))
                else:
                    out = 0.25 * ((p + q) + (r + s))
                Gn[sI:sI + nI, sJ:sJ + nJ] = out
                if I != J:
                    Gn[sJ:sJ + nJ, sI:sI + nI] = out.T
                else:
                    for k in range(nI):
                        x0, x1 = A[0][k], A[1][k]
                        if x0 != MISS and IA[0][k] == IA[1][k]:
                            dn[sI + k] = dp[rI[x0]]
                        elif x0 != MISS and x1 != MISS:
                            dn[sI + k] = 0.5 + 0.5 * S[x0, x1]
                        else:
                            dn[sI + k] = 0.5
        assert not np.isnan(Gn).any() and not np.isnan(dn).any()
        Gp, dp = Gn, dn
    out = Gp[np.ix_(fcls, fcls)]
    same = find[:, None] == find[None, :]
    return np.where(same, dp[fcls][:, None] * np.ones((1, m)), out)
