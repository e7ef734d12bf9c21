"""CPU (not gpu): the C restatement (oracle/mf_oracle.c) reproduces every golden vector generated from the unmodified
reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import pytest

from helpers import KERNEL_FIXTURES, check_diffusion_against_golden, check_kernels_against_golden, check_psolve52, load_golden


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", KERNEL_FIXTURES)
def test_port_reproduces_reference_golden(name, prec, port32, port64):
    O = port32 if prec == 4 else port64
    # the float build of the reference is deterministic here (double accumulators of float products): bit-exact CG too
    check_kernels_against_golden(O, load_golden(name, prec), prec, exact_reductions=(prec == 4))
    check_diffusion_against_golden(O, load_golden(name, prec), prec)


def test_port_reproduces_test_0100_and_0110(port32):
    check_psolve52(port32)
