"""CPU oracle for the soundgen source-filter synthesis path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy (float64 / x87 long double)
restatement of the reference's R algorithm for the hot path; it exists to check
the CUDA path and to serve as the timed CPU baseline in ``bench.py``.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it.  The product package
(``soundgen_beta_b200``) never imports anything from here.

PARITY UNPINNED: the reference (a pure-R package) ships no tests, no golden
vectors and no known-answer values, and neither R nor the R sources of
``stats::spline/approx/fft`` are available in the build container, so this
restatement cannot be checked against outputs of the reference itself.  It is
pinned only by the hand-derived known answers of SURVEY.md Appendix B (kept in
``tests/golden/appendix_b.json``) and by invariants (see ``tests/``).
"""
