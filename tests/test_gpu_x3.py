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
    assert ops._lib.launch_count() - before >= 3 + 5            # splits + GEMMs + column sums: the x3 path ran (FFMA: 4)
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


# ------------------------------------------------------------------ K5 / K6 / K7 with fp32 activations (the fp32 chain)
def _ln64(x, g, b, eps=1e-5):
    return torch.nn.functional.layer_norm(x, x.shape[-1:], g, b, eps)


@pytest.mark.parametrize("d", [128, 256, 512])
def test_layer_norm_residual_fp32_io(d):
    """K5 / K5' with float32 y / dy against fp64 LayerNorm autograd; the residual gradient is added inside K5'."""
    from sl_hwgat_b200 import ops
    g = torch.Generator().manual_seed(d)
    n = 1000                                             # not a multiple of the rows per warp
    x = (torch.randn(n, d, generator=g) * 2 + 0.5).cuda().requires_grad_(True)
    gm = (1 + 0.1 * torch.randn(d, generator=g)).cuda().requires_grad_(True)
    bt = (0.1 * torch.randn(d, generator=g)).cuda().requires_grad_(True)
    g_res, g_y = torch.randn(n, d, generator=g).cuda(), torch.randn(n, d, generator=g).cuda()
    xr, y = ops.layer_norm_residual(x, gm, bt, 1e-5, io=torch.float32)
    assert y.dtype == torch.float32 and torch.equal(xr, x.detach())
    torch.autograd.backward([xr, y], [g_res, g_y])
    x64, g64, b64 = (t.detach().double().requires_grad_(True) for t in (x, gm, bt))
    y64 = _ln64(x64, g64, b64)
    torch.autograd.backward([x64 * 1.0, y64], [g_res.double(), g_y.double()])
    assert rel_l2(y, y64) < 1e-6
    assert rel_l2(x.grad, x64.grad) < 1e-6 and rel_l2(gm.grad, g64.grad) < 2e-6 and rel_l2(bt.grad, b64.grad) < 2e-6


@pytest.mark.parametrize("with_ln", [True, False])
@pytest.mark.parametrize("d,p", [(128, 0.0), (256, 0.1), (512, 0.3)])
def test_bias_dropout_add_ln_fp32_io(d, p, with_ln):
    """K6 / K6' with float32 a0 / y / dy / d_a0.  The dropout mask is read off the output (kept entries carry
    (a0 + bias) / (1 - p)), then everything is compared with fp64 autograd under that mask."""
    from sl_hwgat_b200 import ops
    g = torch.Generator().manual_seed(d + int(p * 100))
    n = 128 * 9
    res = torch.randn(n, d, generator=g).cuda().requires_grad_(True)
    a0 = (torch.rand(n, d, generator=g) * 2 + 2.0).cuda().requires_grad_(True)      # away from zero: the mask is readable
    bias = (0.1 * torch.randn(d, generator=g)).cuda().requires_grad_(True)
    norm = torch.nn.LayerNorm(d).cuda() if with_ln else None
    if norm is not None:
        with torch.no_grad():
            norm.weight.add_(0.1 * torch.randn(d, generator=g).cuda())
            norm.bias.add_(0.1 * torch.randn(d, generator=g).cuda())
    torch.manual_seed(3)
    x1, y = ops.bias_dropout_add_ln(res, a0, bias, norm, p, True, io=torch.float32)
    kept = (x1.detach() - res.detach()).abs() > 1e-3 if p else torch.ones_like(res, dtype=torch.bool)
    assert abs(1 - kept.float().mean().item() - p) < 1e-2
    g_x1 = torch.randn(n, d, generator=g).cuda()
    g_y = torch.randn(n, d, generator=g).cuda()
    if with_ln:
        assert y.dtype == torch.float32
        torch.autograd.backward([x1, y], [g_x1, g_y])
    else:
        assert y is None
        x1.backward(g_x1)
    r64, a64, b64 = (t.detach().double().requires_grad_(True) for t in (res, a0, bias))
    pq = round(p * 65536) / 65536          # the kernels hold the drop probability as a 16-bit threshold
    x64 = r64 + torch.where(kept, (a64 + b64) / (1 - pq), torch.zeros_like(a64))
    if with_ln:
        w64, nb64 = norm.weight.detach().double().requires_grad_(True), norm.bias.detach().double().requires_grad_(True)
        y64 = _ln64(x64, w64, nb64)
        torch.autograd.backward([x64, y64], [g_x1.double(), g_y.double()])
        assert rel_l2(y, y64) < 3e-6              # fp32 LayerNorm of rows with mean ~3
        assert rel_l2(norm.weight.grad, w64.grad) < 5e-6 and rel_l2(norm.bias.grad, nb64.grad) < 5e-6
    else:
        x64.backward(g_x1.double())
    assert rel_l2(x1, x64) < 1e-6
    assert rel_l2(res.grad, r64.grad) < 3e-6 and rel_l2(a0.grad, a64.grad) < 3e-6 and rel_l2(bias.grad, b64.grad) < 5e-6


def test_merge_fold_fp32_io_equals_k6_k4_k5():
    """the level boundary with fp32 activations: K6 storing the merged layout + K5 over 2d-wide rows == K6 -> K4 -> K5,
    forward and backward, bit for bit (outputs, input gradients; parameter gradients up to the order of the atomics)"""
    from sl_hwgat_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, F, K, d = 3, 8, 64, 128
    res0 = torch.randn(B, F, K, d, generator=g).cuda()
    a00 = torch.randn(B, F, K, d, generator=g).cuda()
    bias0 = (0.1 * torch.randn(d, generator=g)).cuda()
    norm = torch.nn.LayerNorm(2 * d).cuda()
    gxm = torch.randn(B, F // 2, K, 2 * d, generator=g).cuda()
    gy = torch.randn(B, F // 2, K, 2 * d, generator=g).cuda()

    def run(folded):
        res, a0, bias = (t.clone().requires_grad_(True) for t in (res0, a00, bias0))
        norm.zero_grad(set_to_none=True)
        if folded:
            xm, y = ops.bias_dropout_add_merge_ln(res, a0, bias, norm, 0.0, True, io=torch.float32)
        else:
            x1, _ = ops.bias_dropout_add_ln(res, a0, bias, None, 0.0, True, io=torch.float32)
            xm = ops.temporal_merge(x1)
            xm, y = ops.layer_norm_residual(xm, norm.weight, norm.bias, norm.eps, io=torch.float32)
        torch.autograd.backward([xm, y], [gxm, gy])
        return xm.detach(), y.detach(), res.grad, a0.grad, bias.grad, norm.weight.grad.clone(), norm.bias.grad.clone()

    a, b = run(True), run(False)
    for i in range(4):
        assert torch.equal(a[i], b[i]), i
    for i in range(4, 7):
        assert rel_l2(a[i], b[i]) < 1e-6, i


@pytest.mark.parametrize("cols,p", [(256, 0.0), (512, 0.1), (1024, 0.5)])
def test_bias_gelu_dropout_fp32_io(cols, p):
    """K7 / K7' with float32 tensors: exact-erf GELU (A&S 7.1.26, 1.5e-7 absolute) against fp64, the mask read off the
    output and reused for the gradient check."""
    from sl_hwgat_b200 import ops
    g = torch.Generator().manual_seed(cols)
    n = 128 * 5
    if p == 0:      # both signs of the pre-activation (the two branches of the erfc form)
        u = (torch.randn(n, cols, generator=g) * 1.5 + 0.5).cuda().requires_grad_(True)
    else:           # pre-activations >= 0.5: gelu is never near zero, so "output == 0" identifies the dropped entries
        u = (torch.rand(n, cols, generator=g) * 3 + 1.0).cuda().requires_grad_(True)
    bias = (0.4 * torch.rand(cols, generator=g) - 0.2).cuda().requires_grad_(True)
    torch.manual_seed(4)
    out = ops.bias_gelu_dropout(u, bias, p, True, io=torch.float32)
    assert out.dtype == torch.float32
    pre = (u.detach().double() + bias.detach().double())
    ref = torch.nn.functional.gelu(pre)
    big = ref.abs() > 1e-2 if p else torch.ones_like(ref, dtype=torch.bool)
    kept = (out.detach() != 0) if p else torch.ones_like(ref, dtype=torch.bool)
    assert p == 0 or bool(big.all())
    assert abs(1 - kept[big].float().mean().item() - p) < 1.5e-2
    pq = round(p * 65536) / 65536          # the kernels hold the drop probability as a 16-bit threshold
    ref_m = torch.where(kept, ref / (1 - pq), torch.zeros_like(ref))
    assert float((out.detach().double() - ref_m)[big].abs().max()) < 2e-6
    dg = torch.randn(n, cols, generator=g).cuda()
    out.backward(dg)
    pre64 = pre.clone().requires_grad_(True)
    (torch.where(kept, torch.nn.functional.gelu(pre64) / (1 - pq), torch.zeros_like(pre64))).backward(dg.double())
    assert rel_l2(u.grad, pre64.grad) < 2e-6 and rel_l2(bias.grad, pre64.grad.sum(0)) < 5e-6
