"""Validation helpers (host side); mirrors /root/reference/skdist/distribute/validation.py
for the functions the hot path calls: ``_check_estimator`` (validation.py:14-20) and
``_check_is_fitted`` (validation.py:23-29)."""
from sklearn.utils.validation import check_is_fitted


def _check_estimator(estimator, verbose=False):
    """Print backend awareness (ref validation.py:14-20: spark vs local).  Here the
    backend is always the B200 engine; ``sc`` is accepted and ignored."""
    if verbose:
        from .. import parallel
        rank, world, _ = parallel.dist_info()
        if rank == 0:
            print("skdist_b200: running on %d B200 process(es); sc=%s is ignored"
                  % (world, "None" if getattr(estimator, "sc", None) is None else "given"))


def _check_is_fitted(estimator, attributes=None):
    return check_is_fitted(estimator)
