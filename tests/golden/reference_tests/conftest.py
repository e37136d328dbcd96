"""Runs the reference's own operator tests against the B200 kernels.

`test_operator.py` in this directory is /root/reference/tests/operator/test_operator.py, byte for byte
(sha256 7df2879fbf3f720c389385a8915a20102c7c996545a562accfe66eb1746de193; tests/test_abi_cpu.py checks the hash).
It is vendored because the round-1 review asked for the reference's tests to run *unchanged* in the driver's GPU
suite and /root/reference does not exist on the GPU box.  It is test input, not product source.

The file imports `from optical_flow import normalize, resize, scale, warp` and builds CPU tensors; tests/conftest.py
puts `torch-optical-flow_b200/` on sys.path, so those names are the drop-in package's, which stages host tensors
through the GPU (there is no CPU compute path).  Every test here therefore needs the device: marked `gpu`.

`test_scale` fails against the reference itself: `scaled[:, 0]` is (1, 2, 2) while `expected_x`
is (1, 1, 2, 2) and `torch.equal` requires equal shapes (test_operator.py:53).  The values are right on both sides
(tests/test_gpu_parity.py::test_resize_reference_known_answers restates the check with matching shapes), so it is an
expected failure here too -- strict, so a change of behaviour on either side is noticed."""
import pytest


def pytest_collection_modifyitems(config, items):
    for item in items:
        if "reference_tests" not in str(item.fspath):
            continue
        item.add_marker(pytest.mark.gpu)
        if item.name == "test_scale":
            item.add_marker(pytest.mark.xfail(
                strict=True, reason="reference test bug: torch.equal on shapes (1,2,2) vs (1,1,2,2), test_operator.py:53"))


@pytest.fixture(autouse=True)
def _need_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
