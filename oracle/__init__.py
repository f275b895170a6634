"""CPU oracle for the RLVI hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package restates, in plain NumPy / CPU torch, the algorithm of the reference
(akarakulev/rlvi) for the E-step + weighted M-step path.  Every function cites
the reference file:line it follows.

Rules (task statement, item 3):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import anything from here -- as the checker or the CPU baseline,
    never as the thing that is shipped or measured as the product;
  * nothing under ``rlvi_b200/`` imports this package (tests/test_boundary.py enforces that).

Pinning status: the reference ships NO golden vectors / known-answer tests for this path
(SURVEY.md section 8c).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in
the build container by ``oracle/make_golden.py`` (which imports ``/root/reference`` through
``oracle/ref_shim.py``); the resulting vectors are committed under ``tests/golden/`` and
``tests/test_oracle_golden.py`` checks the restatement against them on every run.
"""
