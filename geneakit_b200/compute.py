"""gen.phi / gen.phiMean / gen.gc with the reference's signatures and return types
(reference geneakit/compute.py:9-92, :95-132, :340-373), computed on the GPU."""
import time

import numpy as np
import pandas as pd

from . import cgeneakit


def phi(gen, **kwargs):
    """Compute kinship coefficients between probands (reference compute.py:9-92).

    Args:
        gen (cgeneakit.Pedigree): genealogy object.
        pro (list, optional): proband ids, default all probands.
        verbose (bool): print the cut sizes and the elapsed time.
        compute (bool): if False only print the memory estimate.
        sparse (bool): the experimental sparse algorithm is outside this package's scope.
    Returns:
        pd.DataFrame indexed by the proband ids on both axes.
    """
    pro = kwargs.get('pro', None)
    if pro is None:
        pro = cgeneakit.get_proband_ids(gen)
    verbose = kwargs.get('verbose', False)
    compute = kwargs.get('compute', True)
    sparse = kwargs.get('sparse', False)
    if not compute:
        required_memory = cgeneakit.get_required_memory_for_kinships(gen, pro)
        print(f'You will require at least {round(required_memory, 2)} GB of RAM.')
        return
    if sparse:
        # reference compute.py:43-83: CSR by default, dense-layout sparse DataFrame otherwise (raw=False)
        from scipy.sparse import csr_matrix
        if verbose:
            begin = time.time()
        data, indices, indptr = cgeneakit.compute_kinships_sparse(gen, pro, verbose)
        matrix = csr_matrix((data, indices, indptr), shape=(len(pro), len(pro)))
        if verbose:
            print(f'Elapsed time: {round(time.time() - begin, 2)} seconds')
        if kwargs.get('raw', True):
            return matrix
        return pd.DataFrame.sparse.from_spmatrix(matrix, index=pro, columns=pro)
    if verbose:
        begin = time.time()
    cmatrix = cgeneakit.compute_kinships(gen, pro, verbose, **({'gpus': kwargs['gpus']} if 'gpus' in kwargs else {}))
    kinship_matrix = pd.DataFrame(cmatrix, index=pro, columns=pro, copy=False)
    if verbose:
        print(f'Elapsed time: {round(time.time() - begin, 2)} seconds')
    return kinship_matrix


def phiMean(kinship_matrix):
    """Mean kinship excluding self pairs (reference compute.py:95-132, dense branch):
    (sum of all entries - trace) / (n^2 - n)."""
    values = kinship_matrix.values if isinstance(kinship_matrix, pd.DataFrame) else np.asarray(kinship_matrix)
    total = values.sum(axis=0).sum()
    diag_sum = np.diag(values).sum()
    n = values.shape[0]
    return (total - diag_sum) / (n**2 - n)


def phi_device(gen, **kwargs):
    """Device-resident variant: returns a cgeneakit.DevicePhi whose matrix stays in HBM;
    `.mean()` is the fused on-device phiMean, `.to_host()` copies the matrix out."""
    pro = kwargs.get('pro', None)
    if pro is None:
        pro = cgeneakit.get_proband_ids(gen)
    job = cgeneakit.DevicePhi(gen, pro, kwargs.get('verbose', False),
                              **{k: v for k, v in kwargs.items() if k in ('compress', 'order', 'device')})
    job.ids = list(pro)
    return job.upload().run()


def phiOver(phiMatrix, threshold):
    """Pairs of probands whose kinship exceeds a threshold (reference compute.py:135-190).  A DevicePhi job is scanned
    on the GPU (the matrix never leaves HBM); a DataFrame / array is scanned like the reference does."""
    if isinstance(phiMatrix, cgeneakit.DevicePhi):
        i, j, v = phiMatrix.over(threshold)
        ids = getattr(phiMatrix, 'ids', None)
        p1 = [ids[k] for k in i] if ids is not None else i
        p2 = [ids[k] for k in j] if ids is not None else j
        return pd.DataFrame({'pro1': p1, 'pro2': p2, 'kinship': v})
    values = phiMatrix.values if isinstance(phiMatrix, pd.DataFrame) else np.asarray(phiMatrix)
    i, j = np.nonzero(np.tril(values > threshold, -1))
    if isinstance(phiMatrix, pd.DataFrame):
        return pd.DataFrame({'pro1': phiMatrix.index[i], 'pro2': phiMatrix.columns[j], 'kinship': values[i, j]})
    return pd.DataFrame({'pro1': i, 'pro2': j, 'kinship': values[i, j]})


def phiCI(phiMatrix, prob=[0.025, 0.05, 0.95, 0.975], b=5000, random_state=None):
    """Bootstrap confidence intervals of the mean kinship (reference compute.py:193-252: scipy's percentile bootstrap of
    the mean of the sub-matrix entries < 0.5).  A DevicePhi job evaluates the b statistics on the GPU; the resamples are
    drawn exactly as scipy.stats.bootstrap draws them (rng.integers(0, n, (b, n))), so a seed gives scipy's result."""
    if not isinstance(phiMatrix, cgeneakit.DevicePhi):
        from scipy.stats import bootstrap
        phi_array = phiMatrix.to_numpy() if isinstance(phiMatrix, pd.DataFrame) else np.asarray(phiMatrix)
        n = phi_array.shape[0]

        def statistic(indices):
            sub = phi_array[indices][:, indices]
            sel = sub[sub < 0.5]
            return np.mean(sel) if sel.size > 0 else np.nan
    else:
        n = phiMatrix.m
    sorted_prob = sorted(prob)
    quantiles = {}
    i, j = 0, len(sorted_prob) - 1
    while i < j:
        lower, upper = sorted_prob[i], sorted_prob[j]
        if isinstance(phiMatrix, cgeneakit.DevicePhi):
            rng = np.random.default_rng(random_state)
            stats = phiMatrix.ci_stats(rng.integers(0, n, (b, n)))
            alpha = (1 - (upper - lower)) / 2
            lo, hi = np.percentile(stats, [alpha * 100, (1 - alpha) * 100])
        else:
            res = bootstrap((np.arange(n),), statistic, method='percentile', confidence_level=upper - lower,
                            n_resamples=b, vectorized=False,
                            random_state=None if random_state is None else np.random.default_rng(random_state))
            lo, hi = res.confidence_interval.low, res.confidence_interval.high
        quantiles[lower], quantiles[upper] = lo, hi
        i += 1
        j -= 1
    results = [quantiles.get(p, np.nan) for p in prob]
    return pd.DataFrame([results], columns=[f"{p*100}%" for p in prob], index=["Mean"])


def f(gen, **kwargs):
    """Inbreeding coefficients of the probands (reference compute.py:255-282): DataFrame indexed by proband id, column 'F'."""
    pro = kwargs.get('pro', None)
    if pro is None:
        pro = cgeneakit.get_proband_ids(gen)
    cmatrix = cgeneakit.compute_inbreedings(gen, pro)
    return pd.DataFrame(cmatrix, index=pro, columns=['F'], copy=False)


def gc(pedigree, **kwargs):
    """Genetic contribution of ancestors to probands (reference compute.py:340-373):
    rows = proband ids, columns = ancestor ids (default: all founders)."""
    pro = kwargs.get('pro', None)
    ancestors = kwargs.get('ancestors', None)
    if pro is None:
        pro = cgeneakit.get_proband_ids(pedigree)
    if ancestors is None:
        ancestors = cgeneakit.get_founder_ids(pedigree)
    cmatrix = cgeneakit.compute_genetic_contributions(pedigree, pro, ancestors)
    return pd.DataFrame(cmatrix, index=pro, columns=ancestors, copy=False)
