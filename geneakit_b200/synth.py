"""Synthetic pedigrees of BASELINE.json's shapes: discrete generations,
monogamous couples drawn at random inside the previous generation, each child draws a couple
uniformly; ids 1..N in generation order (already topologically sorted: load with sorted=True)."""
import numpy as np


def generation_widths(n_individuals, n_generations, founders=None):
    """Equal widths, or (founders given) w_g = min(W, founders*2^g) with W solved so that the
    widths sum to n_individuals (BASELINE config 5: bottlenecked founder population)."""
    if founders is None:
        w = n_individuals // n_generations
        widths = [w] * n_generations
        widths[-1] += n_individuals - w * n_generations
        return widths
    lo, hi = founders, n_individuals
    while lo < hi:
        mid = (lo + hi + 1) // 2
        if sum(min(mid, founders * 2 ** g) for g in range(n_generations)) <= n_individuals:
            lo = mid
        else:
            hi = mid - 1
    widths = [min(lo, founders * 2 ** g) for g in range(n_generations)]
    widths[-1] += n_individuals - sum(widths)
    return widths


def synthetic_arrays(n_individuals, n_generations, seed, founders=None):
    """Returns (ids, fathers, mothers, sexes, generation_offsets); parent id 0 = unknown."""
    rng = np.random.default_rng(seed)
    widths = generation_widths(n_individuals, n_generations, founders)
    offs = np.concatenate([[0], np.cumsum(widths)]).astype(np.int64)
    n = int(offs[-1])
    ids = np.arange(1, n + 1, dtype=np.int32)
    sexes = (1 + (np.arange(n) % 2)).astype(np.int32)      # alternate male / female
    fathers = np.zeros(n, dtype=np.int32)
    mothers = np.zeros(n, dtype=np.int32)
    for g in range(1, n_generations):
        lo, hi = int(offs[g - 1]), int(offs[g])
        prev = ids[lo:hi]
        males = prev[sexes[lo:hi] == 1].copy()
        females = prev[sexes[lo:hi] == 2].copy()
        rng.shuffle(males)
        rng.shuffle(females)
        c = min(len(males), len(females))
        pick = rng.integers(0, c, size=int(offs[g + 1] - offs[g]))
        fathers[offs[g]:offs[g + 1]] = males[pick]
        mothers[offs[g]:offs[g + 1]] = females[pick]
    return ids, fathers, mothers, sexes, offs


def synthetic_probands(ids, offs, n_probands, seed):
    """Probands sampled without replacement from the last generation, sorted by id (numpy only)."""
    rng = np.random.default_rng(seed + 1000003)
    last = ids[offs[-2]:offs[-1]]
    k = min(n_probands, len(last))
    pro = np.sort(rng.choice(last, size=k, replace=False)) if k < len(last) else last.copy()
    return [int(x) for x in pro]


def synthetic_pedigree(n_individuals, n_generations, n_probands, seed, founders=None):
    """(Pedigree, proband ids): probands sampled without replacement from the last generation,
    sorted by id."""
    from . import cgeneakit          # (this module stays importable without the CUDA library: the reference arm of bench.py)
    ids, fathers, mothers, sexes, offs = synthetic_arrays(n_individuals, n_generations, seed, founders)
    ped = cgeneakit.load_pedigree_from_vectors(ids, fathers, mothers, sexes, True)
    return ped, synthetic_probands(ids, offs, n_probands, seed)
