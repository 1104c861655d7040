"""Bootstrap multiplicities and splitter seeds of a forest's trees, computed by the library's host
threads (csrc/bootstrap.cu; no GPU involved).

What one task of the reference derives from a tree seed (ref ensemble.py:51-55, 97-99:
``bincount(RandomState(seed).randint(0, n, n))``) and scikit-learn's splitter seed
(SK/tree/_splitter.pyx:155), for many trees at once and bit-identical to numpy's legacy generator."""
import numpy as np

from . import _lib


def bootstrap_counts(seeds, n_samples, bootstrap=True, n_threads=0):
    """(counts uint8 [n_trees, n_samples] or None when bootstrap is False, rand_r uint32 [n_trees])."""
    seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.int64))
    if seeds.ndim != 1 or len(seeds) == 0 or seeds.min() < 0 or seeds.max() > 0xFFFFFFFF:
        raise ValueError("tree seeds must be integers in [0, 2**32 - 1]")
    s32 = seeds.astype(np.uint32)
    t = len(s32)
    counts = np.empty((t, int(n_samples)), dtype=np.uint8) if bootstrap else None
    rand_r = np.empty(t, dtype=np.uint32)
    rc = _lib.load().skd_bootstrap_counts(t, _lib.ptr(s32), int(n_samples), int(bool(bootstrap)),
                                          _lib.ptr(counts) if bootstrap else None, _lib.ptr(rand_r), int(n_threads))
    if rc == 1:
        raise NotImplementedError("a bootstrap multiplicity above 255 does not fit the device format")
    if rc != 0:
        raise ValueError("skd_bootstrap_counts: bad arguments")
    return counts, rand_r
