"""The x3 mode of the fp32 path (csrc/gemm_x3.cu): fp32 Linear layers as six bf16 tcgen05 products of hi / mid / lo
planes.  Checked against fp64 matmuls of the same fp32 inputs; the bound is the north_star's fp32 tolerance (1e-5),
the observed error is printed (-s) and is at the level of an fp32 GEMM."""
import pytest
import torch

from tests._util import rel_inf, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture
def x3_mode():
    from sl_hwgat_b200 import ops
    prev = ops.set_fp32_mode("x3")
    yield
    ops.set_fp32_mode(prev)


@pytest.mark.parametrize("n,d_in,d_out", [(128, 128, 128), (384, 256, 768), (4096, 512, 1536), (128 * 37, 1024, 512),
                                          (65536, 128, 384), (128 * 149 * 3, 128, 256), (262144, 256, 128)])
def test_linear_x3_against_fp64(n, d_in, d_out, x3_mode):
    """forward, input gradient, weight gradient (token split in bounded chains) and bias gradient; inputs with a
    non-zero mean so that the sums grow (the case in which a truncating accumulator drifts)."""
    from sl_hwgat_b200 import ops
    assert ops.linear_x3_active(n, d_in, d_out)
    g = torch.Generator().manual_seed(n + d_in + d_out)
    x = (torch.randn(n, d_in, generator=g) + 0.3).cuda().requires_grad_(True)
    w = (torch.randn(d_out, d_in, generator=g) / d_in ** 0.5 + 0.01).cuda().requires_grad_(True)
    b = torch.randn(d_out, generator=g).cuda().requires_grad_(True)
    dy = (torch.randn(n, d_out, generator=g) + 0.2).cuda()
    before = ops._lib.launch_count()
    y = ops.linear_f32(x, w, b)
    y.backward(dy)
    torch.cuda.synchronize()
    assert ops._lib.launch_count() - before >= 3 + 6            # splits + GEMMs + column sums: the x3 path ran (FFMA: 4)
    xd, wd, bd, dyd = x.detach().double(), w.detach().double(), b.detach().double(), dy.double()
    errs = {"y": rel_l2(y, xd @ wd.t() + bd), "dx": rel_l2(x.grad, dyd @ wd), "dw": rel_l2(w.grad, dyd.t() @ xd),
            "db": rel_l2(b.grad, dyd.sum(0)), "y_inf": rel_inf(y, xd @ wd.t() + bd)}
    print(f"x3 n={n} d_in={d_in} d_out={d_out}: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert max(errs.values()) < 5e-6, errs


def test_x3_off_in_deterministic_mode_and_for_other_shapes(x3_mode):
    from sl_hwgat_b200 import ops
    assert not ops.linear_x3_active(100, 128, 128)          # n % 128
    assert not ops.linear_x3_active(128, 128, 262)          # the classifier head
    prev = ops.set_deterministic(True)
    try:
        assert not ops.linear_x3_active(128, 128, 128)
        x = torch.randn(256, 128, device="cuda", requires_grad=True)
        w = torch.randn(128, 128, device="cuda", requires_grad=True)
        y = ops.linear_f32(x, w, None)
        y.sum().backward()
        assert rel_l2(y, x.detach().double() @ w.detach().double().t()) < 1e-6
    finally:
        ops.set_deterministic(prev)


def test_linear_x3_writes_nothing_outside_its_outputs(x3_mode):
    """straight through the C ABI with NaN guard rows behind every output (the pool has no compute-sanitizer)"""
    from sl_hwgat_b200 import _lib, ops
    lib = _lib.load()
    n, d_in, d_out = 128 * 5, 256, 384
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, d_in, generator=g).cuda()
    w = (torch.randn(d_out, d_in, generator=g) / 16).cuda()
    b = torch.randn(d_out, generator=g).cuda()
    dy = torch.randn(n, d_out, generator=g).cuda()
    nan = float("nan")
    y = torch.full((n + 128, d_out), nan, device="cuda")
    dx = torch.full((n + 128, d_in), nan, device="cuda")
    dw = torch.full((d_out + 8, d_in), nan, device="cuda")
    db = torch.full((d_out + 8,), nan, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hwgat_linear_f32_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), n, d_in, d_out, st), "fwd")
    _lib.check(lib.hwgat_linear_f32_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), dx.data_ptr(), dw.data_ptr(),
                                        db.data_ptr(), n, d_in, d_out, st), "bwd")
    torch.cuda.synchronize()
    assert torch.isnan(y[n:]).all() and torch.isnan(dx[n:]).all() and torch.isnan(dw[d_out:]).all() and torch.isnan(db[d_out:]).all()
    assert rel_l2(y[:n], x.double() @ w.double().t() + b.double()) < 2e-6
    assert rel_l2(dx[:n], dy.double() @ w.double()) < 2e-6
    assert rel_l2(dw[:d_out], dy.double().t() @ x.double()) < 2e-6
    assert rel_l2(db[:d_out], dy.double().sum(0)) < 2e-6
