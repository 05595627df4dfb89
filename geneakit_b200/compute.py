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
        raise NotImplementedError(
            "geneakit_b200 accelerates the dense path only; phi(sparse=True) stays with the reference")
    if verbose:
        begin = time.time()
    cmatrix = cgeneakit.compute_kinships(gen, pro, verbose)
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
    return job.upload().run()


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
