"""CPU oracle for the BOPL uEI hot path — TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy restatement of the reference's algorithm (RaulAstudillo06/BOPL) for the path
SURVEY.md §8 names: multi-output GP posterior (mean / variance / gradients) and the uEI /
uEI_affine acquisition functions.  Every function cites the reference file:line it follows and
calls the same LAPACK routines through SciPy that the reference's `GPy/util/linalg.py` calls.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and only as the checker / the CPU baseline.  The product
(`bopl_b200/`) never imports it; there is no CPU fallback.

PARITY PINNING: the reference holds no golden vectors or tests for this path,
and it cannot be imported in the build container (`paramz`, `pathos` absent), so the oracle is
pinned by (i) the nearest known-answer tests the vendored suites hold — closed-form EI KATs
(`GPyOpt/testing/acquisitions_tests/test_ei_acquisition.py:21-37`), predict ≡ textbook formula
(`GPy/testing/model_tests.py:63-82`), finite-difference kernel-gradient checks
(`GPy/testing/kernel_tests.py:352-434`) — see `tests/test_oracle_*.py`; and (ii) a loop-faithful
form that follows the reference's Python loops statement by statement, against which the
vectorised form is checked.  For uEI itself this is "parity unpinned by reference fixtures".
"""
