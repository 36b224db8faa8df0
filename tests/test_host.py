"""CPU-side checks: the C-ABI library loads and exports every symbol the header
declares, the drop-in module keeps the reference's state_dict, the product path
fails loudly without CUDA, and the data-parallel host logic (gloo, world 2)."""
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from oracle import hwgate_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from sl_hwgat_b200 import _lib
    header = open(os.path.join(ROOT, "include", "hwgat_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(hwgat_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hwgat_version() == _lib.ABI_VERSION
    assert b"no fallback" in lib.hwgat_error_string(1002)
    # size query is pure host arithmetic: qkv + dqkv in fp32, dqkv only in bf16
    assert lib.hwgat_attn_workspace_bytes(2, 4, 64, 128, 2, _lib.F32, 1) == 2 * 4 * 64 * 384 * 4 * 2
    assert lib.hwgat_attn_workspace_bytes(2, 4, 64, 128, 2, _lib.BF16, 0) == (384 * 128 + 2 * 192 * 16) * 2
    assert lib.hwgat_attn_workspace_bytes(2, 4, 64, 128, 2, _lib.BF16, 1) == (2 * 4 * 64 * 384 + 2 * 384 * 128 + 2 * 192 * 16) * 2


def test_argument_errors_without_gpu():
    """status codes of the boundary that need no device work"""
    from sl_hwgat_b200 import _lib
    lib = _lib.load()
    # NULL output
    assert lib.hwgat_mask_build(None, 4, 16, 2, 8, 0, None, None) == 1000
    # W != 16 is unsupported (no fallback)
    st = lib.hwgat_attn_fwd(None, None, None, None, -1.0, None, None, 0, 1, 4, 64, 128, 2, 8, 2, 0, 0, 0, None)
    assert st == 1002
    # d != heads * 64
    st = lib.hwgat_attn_fwd(None, None, None, None, -1.0, None, None, 0, 1, 4, 64, 128, 4, 16, 2, 0, 0, 0, None)
    assert st == 1002
    # NULL tensors with a supported geometry
    st = lib.hwgat_attn_fwd(None, None, None, None, -1.0, None, None, 0, 1, 4, 64, 128, 2, 16, 2, 0, 0, 0, None)
    assert st == 1000
    # shift with the pre-partitioned layout is inconsistent
    st = lib.hwgat_attn_fwd(None, None, None, None, -1.0, None, None, 0, 1, 4, 64, 128, 2, 16, 2, 1, 1, 0, None)
    assert st == 1001
    with pytest.raises(_lib.HwgatError):
        _lib.check(1002, "x")


def _cpu_model(T=64, classes=262):
    from sl_hwgat_b200.models import HWGATE
    adj = torch.from_numpy(O.window_adjacency(O.HWGATEConfig().edges, 16, 2).astype(np.float32))
    return HWGATE.Model(2, 64, T, classes, 128, 2, True, [2, 2, 4], [2, 4, 8], 16, adj, 0.1, 0.0, 2.,
                        torch.nn.LayerNorm, None)


def test_dropin_state_dict_matches_reference(golden_dir):
    G = np.load(os.path.join(golden_dir, "full_model.npz"))
    m = _cpu_model()
    assert list(m.state_dict().keys()) == list(G["state_dict_names"])
    assert [str(tuple(v.shape)) for v in m.state_dict().values()] == list(G["state_dict_shapes"])
    sd = O.make_state_dict(O.HWGATEConfig(temporal_dim=64, num_classes=262), seed=1001)
    m.load_state_dict(sd, strict=True)
    # trainable parameter count of the INCLUDE model as main.py:56 prints it
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 9865734
    # the float shift-mask buffers equal what the oracle derives (pinned to the reference)
    for i, layer in enumerate(m.layers):
        F = 64 // 2 ** i
        for j, blk in enumerate(layer.blocks):
            if j % 2:
                assert np.array_equal(blk.attn_mask.numpy() != 0, O.shift_window_mask(F, 4, 16, 2, 1))
            else:
                assert blk.attn_mask is None
        assert tuple(layer.adj_mat.shape) == (F // 2 * 4, 32, 32)


def test_product_path_fails_loudly_without_cuda():
    from sl_hwgat_b200 import _lib
    m = _cpu_model().eval()
    with pytest.raises(_lib.HwgatError):
        m(torch.zeros(1, 64, 64, 2))
    from sl_hwgat_b200 import ops
    with pytest.raises(_lib.HwgatError):
        ops.temporal_merge(torch.zeros(1, 4, 64, 128))
    from sl_hwgat_b200.losses import SmoothedCrossEntropyLoss
    with pytest.raises(_lib.HwgatError):
        SmoothedCrossEntropyLoss()(torch.zeros(2, 5), torch.zeros(2, dtype=torch.long))
    if not torch.cuda.is_available():
        from sl_hwgat_b200.models import model_params
        with pytest.raises(RuntimeError):
            model_params.HWGATEParams({"num_class": 262, "src_len": 64}, 2, "cpu")


def test_no_product_import_of_oracle():
    """the oracle is test infrastructure: nothing under sl_hwgat_b200/ may import it"""
    pkg = os.path.join(ROOT, "sl_hwgat_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src, os.path.join(dp, f)


def test_shard_batch():
    from sl_hwgat_b200.parallel import shard_batch
    for n in (0, 1, 7, 8, 512, 1023):
        for world in (1, 2, 3, 8):
            parts = [shard_batch(n, r, world) for r in range(world)]
            assert parts[0].start == 0 and parts[-1].stop == n
            assert all(parts[i].stop == parts[i + 1].start for i in range(world - 1))
            sizes = [p.stop - p.start for p in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    from sl_hwgat_b200.parallel import GradientAllReduce, broadcast_parameters, shard_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)            # different init per rank: broadcast must fix it
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 32),
                                torch.nn.LayerNorm(32), torch.nn.Linear(32, 5))
    for p in model[2].parameters():
        p.requires_grad_(False)              # a frozen parameter, like Model.B
    broadcast_parameters(model, 0)
    sync = GradientAllReduce(model, bucket_bytes=1024)   # several buckets
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(12, 16, generator=g), torch.randint(0, 5, (12,), generator=g)
    sl = shard_batch(12, rank, world)
    views = [p.grad.data_ptr() for p in model.parameters() if p.requires_grad]
    for step in range(3):                    # three steps: state must reset between them
        if step == 2:
            sync.zero_grad()                 # the intended call: one memset per bucket, .grad stays a bucket view
        else:
            model.zero_grad(set_to_none=(step == 1))   # also tolerated: in-place zero, and .grad replaced (copied in)
        loss = torch.nn.functional.cross_entropy(model(x[sl]), y[sl])
        loss.backward()
        sync.finish()
        # every .grad is (again) the view into its flat bucket: no unpack copy happened
        assert views == [p.grad.data_ptr() for p in model.parameters() if p.requires_grad]
    if rank == 0:
        torch.save({"grads": [p.grad for p in model.parameters() if p.requires_grad],
                    "state": model.state_dict(), "x": x, "y": y}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2_equals_full_batch(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "dp.pt")
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 32),
                                torch.nn.LayerNorm(32), torch.nn.Linear(32, 5))
    model.load_state_dict(r["state"])
    for p in model[2].parameters():
        p.requires_grad_(False)
    torch.nn.functional.cross_entropy(model(r["x"]), r["y"]).backward()   # one process, whole batch
    full = [p.grad for p in model.parameters() if p.requires_grad]
    assert len(full) == len(r["grads"])
    for a, b in zip(r["grads"], full):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)


def test_fused_adamw_has_no_cpu_path():
    """The optimizer mirrors torch.optim.AdamW's interface but only updates CUDA tensors."""
    import torch
    from sl_hwgat_b200 import _lib
    from sl_hwgat_b200.optim import AdamW
    p = torch.ones(4, requires_grad=True)
    opt = AdamW([p], lr=1e-3)
    assert opt.defaults["betas"] == (0.9, 0.999) and opt.defaults["weight_decay"] == 1e-2   # torch's defaults
    p.grad = torch.ones(4)
    with pytest.raises(_lib.HwgatError):
        opt.step()
    with pytest.raises(NotImplementedError):
        AdamW([p], amsgrad=True)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys,
    labelled with the batch it actually ran; non-zero ranks print nothing."""
    import json
    import subprocess
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["value"] > 0 and j["higher_is_better"] is True
    assert j["config"]["per_gpu_batch"] == 8 and "workload" in j["config"]
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
    env["RANK"] = "1"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_header_is_plain_c_and_links_from_a_c_host(tmp_path):
    """include/hwgat_b200.h compiles as C99 and a C program linked against the library can call it (no GPU needed for
    the housekeeping entry points): the boundary carries no C++ or torch types."""
    import shutil
    import subprocess
    from sl_hwgat_b200 import _lib
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    _lib.load()
    exe = str(tmp_path / "host_check")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "host_check.c"), "-o", exe, "-L", libdir, "-lhwgat_b200",
           "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "c abi ok" in r.stdout, r.stdout + r.stderr


def test_scratch_failure_is_retried_once_after_releasing_the_torch_cache(monkeypatch):
    """entry points that take stream-ordered scratch (the x3 planes) retry a HWGAT_ERR_WORKSPACE once, after
    torch.cuda.empty_cache(): the pool cannot grow while the caching allocator holds the free memory"""
    import torch
    from sl_hwgat_b200 import _lib
    calls, released = [], []
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: released.append(1))

    def fake(*args):
        calls.append(args)
        return _lib.ERR_WORKSPACE if len(calls) == 1 else 0

    wrapped = _lib._with_scratch_retry(fake)
    assert wrapped(1, 2) == 0 and len(calls) == 2 and released == [1]
    calls.clear(), released.clear()
    always = _lib._with_scratch_retry(lambda *a: _lib.ERR_WORKSPACE)
    assert always() == _lib.ERR_WORKSPACE and released == [1]            # one retry, then the error is reported
    ok = _lib._with_scratch_retry(lambda *a: 0)
    assert ok() == 0 and released == [1]
    lib = _lib.load()
    assert all(callable(getattr(lib, n)) for n in _lib._SCRATCH_CALLS)
