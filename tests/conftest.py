import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


# The fp32 path has two GEMM modes (ops.set_fp32_mode): "x3" (tcgen05, six bf16 products of hi / mid / lo planes; the
# default) and "ffma" (true-fp32 FFMA kernels).  Every GPU test of the fp32 path - by name: f32 / fp32 - runs in both.
@pytest.fixture
def _fp32_mode(request):
    from sl_hwgat_b200 import ops
    prev = ops.set_fp32_mode(request.param)
    yield request.param
    ops.set_fp32_mode(prev)


def pytest_generate_tests(metafunc):
    name = metafunc.function.__name__
    mod = metafunc.module.__name__.rsplit(".", 1)[-1]
    if mod.startswith("test_gpu") and mod != "test_gpu_x3" and ("f32" in name or "fp32" in name):
        if "_fp32_mode" not in metafunc.fixturenames:
            metafunc.fixturenames.append("_fp32_mode")
        metafunc.parametrize("_fp32_mode", ["x3", "ffma"], indirect=True)
