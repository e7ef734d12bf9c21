"""GPU: the CUDA path against the golden vectors generated from the unmodified reference (tests/golden/), including the
scenario of the reference's own tools/tests/test_0100_psolve.py and test_0110_mgsolve.py."""
import pytest

pytestmark = pytest.mark.gpu

from helpers import KERNEL_FIXTURES, check_diffusion_against_golden, check_kernels_against_golden, check_psolve52, load_golden  # noqa: E402


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", KERNEL_FIXTURES)
def test_cuda_reproduces_reference_golden(name, prec):
    from cuda_impl import CudaImpl
    check_kernels_against_golden(CudaImpl(prec), load_golden(name, prec), prec, exact_reductions=False)
    check_diffusion_against_golden(CudaImpl(prec), load_golden(name, prec), prec)


def test_cuda_reproduces_test_0100_and_0110():
    from cuda_impl import CudaImpl
    check_psolve52(CudaImpl(4))
