"""CPU (not gpu): the size-independent properties of the projection (tests/helpers.py::check_projection_properties) hold for the CPU
restatement on small grids -- the same checker runs on the GPU at BASELINE's 256^3 and 512^3 (tests/test_gpu_step_properties_fullsize.py),
where no second implementation can be run beside the CUDA path."""
import pytest

from helpers import check_projection_properties


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("res,pcs", [(24, [0, 1, 2, 3]), ((40, 36, 28), [0, 1, 2]), ((48, 40, 1), [0, 2])])
def test_projection_properties_port(res, pcs, prec, port32, port64):
    its = check_projection_properties(port32 if prec == 4 else port64, res, prec, pcs)
    assert its[0] > its[2] > 0          # multigrid needs far fewer iterations than plain CG
