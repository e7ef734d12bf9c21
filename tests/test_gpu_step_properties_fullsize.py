"""GPU: size-independent properties of the projection at BASELINE.json's full single-GPU sizes (configs 3 and 4: synthetic smoke plume
with obstacle at 256^3 and 512^3) -- linearity and symmetry of ApplyMatrix, compatibility of the right-hand side, divergence below the
solver tolerance after solvePressure for every preconditioner, idempotence of the projection.  The reference cannot run these sizes
inside a test; its results at 256^3 / 512^3 are compared by tools/run_configs.py (profiles/r1_configs.md)."""
import pytest

pytestmark = pytest.mark.gpu

from helpers import check_projection_properties  # noqa: E402


def test_projection_properties_256_all_preconditioners():
    from cuda_impl import CudaImpl
    its = check_projection_properties(CudaImpl(4), 256, 4, [0, 1, 2, 3], random_vel=True)
    assert its[2] <= 12 and its[3] <= 12 and its[1] < its[0]


def test_projection_properties_256_double():
    from cuda_impl import CudaImpl
    its = check_projection_properties(CudaImpl(8), 256, 8, [0, 3], accuracy=1e-8, random_vel=False)
    assert its[3] <= 20


def test_projection_properties_512():
    """BASELINE configs[3]: 512^3 single-GPU projection (PcNone, the bench workload, and PcMGStatic)"""
    from cuda_impl import CudaImpl
    its = check_projection_properties(CudaImpl(4), 512, 4, [0, 3], random_vel=False)
    assert 1000 < its[0] < 2500 and its[3] <= 12
