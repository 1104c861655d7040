"""oracle/ -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``skdist_b200/`` (the product) may import, call or link this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker /
the CPU arm that is timed beside the GPU arm.

Pinning status (see DESIGN.md "Oracle"):

* The reference (Ibotta/sk-dist @ v0.1.9) is pure Python; every FLOP of the hot
  path is executed by its third-party dependency scikit-learn (reference pins
  ``>=0.20,<0.23.2`` in setup.py:30-31; this image and the GPU box carry 1.9.0)
  and scipy 1.18.1 (L-BFGS-B, LAPACK).  The oracle therefore restates the
  reference's *driver loops* (search.py:315-571, multiclass.py:288-335,
  ensemble.py:177-336) and per-task functions (search.py:180-288) in numpy and
  calls the same installed scikit-learn estimators the reference calls.
* The restatement is pinned against the reference's own per-task functions,
  imported unmodified from /root/reference under a 3-line in-memory shim
  (``oracle/refshim.py``), by ``tests/golden/make_golden.py``; its outputs are
  committed as ``tests/golden/*.npz`` and re-checked by ``tests/test_oracle.py``.
* The reference's own tests pin only toy predictions
  (skdist/distribute/tests/test_search.py:37-56 etc.); those are reproduced in
  ``tests/test_reference_cases.py``.
"""
