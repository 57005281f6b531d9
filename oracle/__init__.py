"""CPU oracle for the BOPL uEI hot path — TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy restatement of the reference's algorithm (RaulAstudillo06/BOPL) for the path
SURVEY.md §8 names: multi-output GP posterior (mean / variance / gradients) and the uEI /
uEI_affine acquisition functions.  Every function cites the reference file:line it follows and
calls the same LAPACK routines through SciPy that the reference's `GPy/util/linalg.py` calls.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and only as the checker / the CPU baseline.  The product
(`bopl_b200/`) never imports it; there is no CPU fallback.

PARITY PINNING: the reference's own test suites hold no golden vectors for this path and the
reference cannot be imported as a package in the build container (`paramz`, `pathos` absent).  The oracle is pinned
against OUTPUTS OF THE REFERENCE'S OWN CODE all the same: `tests/golden/make_golden.py` executes the reference's source
for the whole path straight from /root/reference (modules imported with the plumbing imports stubbed; classes whose
bases need paramz rebuilt from their unchanged method source) and commits the results as `tests/golden/*.npz`;
`tests/test_oracle_pins.py` checks every oracle function against them, plus the nearest known answers of the vendored
suites — closed-form EI KATs (`GPyOpt/testing/acquisitions_tests/test_ei_acquisition.py:21-37`), predict == textbook
formula (`GPy/testing/model_tests.py:63-82`), finite-difference gradient checks (`GPy/testing/kernel_tests.py:352-434`) —
and a loop-faithful form that follows the reference's Python loops statement by statement.
"""
