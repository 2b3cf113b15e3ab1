"""CPU oracle for the BOCF EI-CF hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A line-faithful numpy/scipy restatement of the reference's algorithm for the
path named in BASELINE.json (multi-output GP posterior -> MC composite EI and
its pathwise gradient).  Every function cites the reference file:line it
follows (paths relative to the reference checkout).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package, and only as the checker or
the timed CPU baseline.  ``bocf_b200`` (the product) never imports it.

Parity pin: the reference ships no golden vectors for this path (SURVEY.md
section 8c).  The oracle is instead pinned against outputs of the reference's
OWN source files executed in the build container by ``tests/golden/make_golden.py``
(real reference modules loaded with stubs for the absent third-party packages
paramz/pathos); the resulting fixtures live in ``tests/golden/*.npz``.  The C
routine ``_grad_X`` is additionally compiled from the reference source into
``oracle/_ref/`` (see ``oracle/Makefile``) and checked against the restatement.
"""
from . import kern, linalg, gp, models, utility, acquisitions  # noqa: F401
