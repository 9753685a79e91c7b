"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's StackGAN training path (torch CPU fp32/fp64),
the harness that runs the *real* reference from /root/reference (only available
in the build container) and the script that turns its outputs into the golden
fixtures under tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import anything from this package.  The product package
(imagegenerator_b200/) never imports it and has no CPU fallback.

Parity status: the reference has no tests or golden vectors of its own
(SURVEY.md section 4).  The oracle is pinned against outputs of the reference
itself, run in the build container through oracle/ref_harness.py, and committed
as fixtures by oracle/make_golden.py (tests/golden/*.pt).
"""
