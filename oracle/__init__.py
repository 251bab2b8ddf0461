"""CPU oracle for the DLPM sampling hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``dlpm_b200/`` (the product) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / CPU baseline.

It is a from-scratch restatement (numpy float64 for the stable-law arithmetic,
torch CPU fp32 for the tensor arithmetic) of the reference algorithm; every
function cites the reference ``file:line`` (relative to ``/root/reference``) it
follows.  The restatement is PINNED: ``tests/golden/make_golden.py`` imports the
real reference in the build container (``oracle/ref_import.py``), runs it with
injected noise and stores input/output vectors under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this oracle against those vectors on
every CPU test run (and, when ``/root/reference`` is present, against the live
reference as well).  The reference itself ships no tests or golden vectors
(SURVEY.md section 4), so these generated vectors are the pin.
"""
