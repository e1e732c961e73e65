"""The flux kernels evaluate fp64 division / reciprocal / square root as
straight-line code (csrc/vlct_fpops.cuh) so that independent chains overlap.
This pins them, on the device, to the built-in IEEE operators: wherever the
range guard passes the bits must be identical; everything else is flagged for
re-evaluation with the built-in operator (and solver-like operands must
essentially never be flagged)."""
import ctypes as C

import pytest

from enzo_e_b200 import lib as _lib

pytestmark = pytest.mark.gpu

OPS = ("div", "rcp", "sqrt", "pair", "divz")


def run(n, seed, mode):
    lib = _lib.load()
    out = (C.c_longlong * 10)()
    rc = lib.vlct_selftest_fpops(n, seed, mode, out)
    assert rc == 0
    return {op: (out[2 * i], out[2 * i + 1]) for i, op in enumerate(OPS)}


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_solver_like_operands_are_exact_and_stay_on_the_fast_path(seed):
    n = 1 << 26
    res = run(n, seed, 0)
    for op, (wrong, slow) in res.items():
        assert wrong == 0, (op, wrong)
    # exponents in [-40, 40]: nothing leaves the fast paths' range, except that
    # sqrt of a negative operand (half of the samples) is re-evaluated
    assert all(res[k][1] == 0 for k in ("div", "rcp", "pair", "divz"))
    assert abs(res["sqrt"][1] / n - 0.5) < 0.01


@pytest.mark.parametrize("seed", [11, 12])
def test_arbitrary_bit_patterns(seed):
    res = run(1 << 26, seed, 1)
    for op, (wrong, slow) in res.items():
        assert wrong == 0, (op, wrong)
        assert slow > 0      # extreme exponents must be caught by the guard


def test_special_operands():
    n = 1 << 22
    res = run(n, 5, 2)
    for op, (wrong, slow) in res.items():
        assert wrong == 0, (op, wrong)
    # zero numerators (2 of 16 specials, half of them bit-perturbed) go down
    # the slow path in div but are resolved in place by divz
    assert res["divz"][1] < res["div"][1]
