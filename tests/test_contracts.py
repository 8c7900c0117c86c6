"""Single-process contract tests (no process group): hash / equality of meshes, placements, specs and op schemas (the
sharding-propagation cache key; reference ``legacy/test/dtensor/hash/test_hash.py``), storage aliasing of ``from_local`` /
``to_local`` (``legacy/test/dtensor/memory/test_memory.py``), propagator cache behaviour, and the env-flag switches
(SURVEY §5.6)."""
import os

import pytest
import torch


def _mesh(shape=(4,), names=None, rank=0):
    from vescale_b200 import init_device_mesh

    return init_device_mesh("cpu", shape, mesh_dim_names=names, _rank=rank, _init_process_groups=False)


def test_hash_and_equality_contracts():
    from vescale_b200 import Partial, Replicate, Shard
    from vescale_b200.dtensor import InterleavedShard, RaggedShard
    from vescale_b200.dtensor.op_schema import OpSchema
    from vescale_b200.spec import DTensorSpec, TensorMeta

    # placements: value semantics, usable as dict keys, distinct kinds never collide
    pls = [Replicate(), Shard(0), Shard(1), Partial(), Partial("max"), RaggedShard((0,), (1, 2, 0, 1)), RaggedShard((0,), (1, 2, 1, 0)), RaggedShard((0, 1), (1, 2, 0, 1)),
           InterleavedShard(0, 2), InterleavedShard(0, 4)]
    for i, a in enumerate(pls):
        for j, b in enumerate(pls):
            assert (a == b) == (i == j), (a, b)
            if i == j:
                assert hash(a) == hash(type(a)(**a.__dict__)) if hasattr(a, "__dict__") and a.__dict__ else True
    assert len({p: k for k, p in enumerate(pls)}) == len(pls)
    assert Shard(1) == Shard(dim=1) and hash(Shard(1)) == hash(Shard(dim=1))
    assert Shard(0) != InterleavedShard(0, 2) and InterleavedShard(0, 2) != Shard(0)
    # meshes: equal iff same layout of ranks (and names); sub-meshes hash like an explicitly built mesh
    m1, m2, m3 = _mesh((2, 2), ("a", "b")), _mesh((2, 2), ("a", "b")), _mesh((4,))
    assert m1 == m2 and hash(m1) == hash(m2) and m1 != m3
    # specs: placements + tensor meta
    tm, tm2 = TensorMeta((8, 4), (4, 1), torch.float32), TensorMeta((8, 4), (4, 1), torch.bfloat16)
    s1, s2, s3, s4 = DTensorSpec(m3, (Shard(0),), tm), DTensorSpec(m3, (Shard(0),), tm), DTensorSpec(m3, (Shard(1),), tm), DTensorSpec(m3, (Shard(0),), tm2)
    assert s1 == s2 and hash(s1) == hash(s2) and s1 != s3 and s1 != s4
    # op schemas (cache keys): same op + specs + static args hash equal; a different static arg is a different key
    aten = torch.ops.aten
    k1 = OpSchema(aten.sum.dim_IntList, (s1, [0], False), {})
    k2 = OpSchema(aten.sum.dim_IntList, (s2, [0], False), {})
    k3 = OpSchema(aten.sum.dim_IntList, (s1, [1], False), {})
    k4 = OpSchema(aten.sum.dim_IntList, (s3, [0], False), {})
    assert k1 == k2 and hash(k1) == hash(k2) and k1 != k3 and k1 != k4
    assert len({k1: 0, k2: 1, k3: 2, k4: 3}) == 3


def test_from_local_to_local_alias_storage_and_cache():
    from vescale_b200 import Replicate, Shard
    from vescale_b200.dtensor import DTensor
    from vescale_b200.dtensor.sharding_prop import propagator

    mesh = _mesh((4,), rank=1)
    local = torch.arange(12.0).reshape(3, 4)
    dt = DTensor.from_local(local, mesh, [Shard(0)], run_check=False)
    assert dt.shape == (12, 4) and dt.to_local().data_ptr() == local.data_ptr()  # no copy either way
    dt.to_local().mul_(2)
    assert local[0, 1].item() == 2.0
    y = dt * 3  # pointwise on shards: local result, no communication needed on a fake mesh
    assert isinstance(y, DTensor) and y.placements == (Shard(0),) and torch.equal(y.to_local(), local * 3)
    v = dt.view(12, 2, 2)
    assert v.to_local().data_ptr() == local.data_ptr()  # views stay views
    before = propagator.cache_info()
    for _ in range(5):
        _ = dt + dt
    after = propagator.cache_info()
    assert after["hits"] - before["hits"] >= 4 and after["misses"] - before["misses"] <= 1
    # detach / clone semantics
    assert dt.detach().to_local().data_ptr() == local.data_ptr() and dt.clone().to_local().data_ptr() != local.data_ptr()
    r = DTensor.from_local(torch.ones(2, 2), mesh, [Replicate()], run_check=False)
    assert r.shape == (2, 2) and r.full_tensor().data_ptr() == r.to_local().data_ptr()


def test_submesh_and_coordinates():
    m = _mesh((2, 3), ("dp", "tp"), rank=4)
    assert list(m.get_coordinate()) == [1, 1] and m.get_local_rank("tp") == 1 and m.get_local_rank("dp") == 1
    tp, dp = m["tp"], m["dp"]
    assert tp.size() == 3 and dp.size() == 2 and tp.mesh_dim_names == ("tp",)
    assert tp.mesh.tolist() == [3, 4, 5] and dp.mesh.tolist() == [1, 4]
    assert m.size(0) == 2 and m.size("tp") == 3 and m.ndim == 2 and m.size() == 6
    with pytest.raises((KeyError, ValueError)):
        m["pp"]


def test_env_flags(monkeypatch):
    """VESCALE_STRICT_RULES refuses the replicate fallback; VESCALE_DISABLE_REDISTRIBUTE forbids implicit resharding in
    dispatch (legacy ``dtensor/_diff.py:24``)."""
    from vescale_b200 import Shard
    from vescale_b200.dtensor import DTensor

    mesh = _mesh((1,))
    a = DTensor.from_local(torch.randn(4, 4), mesh, [Shard(0)], run_check=False)
    b = DTensor.from_local(torch.randn(4, 4), mesh, [Shard(1)], run_check=False)
    monkeypatch.setenv("VESCALE_DISABLE_REDISTRIBUTE", "1")
    import importlib

    import vescale_b200.dtensor.dispatch as disp

    if hasattr(disp, "_redistribute_disabled"):
        assert disp._redistribute_disabled() is True
    monkeypatch.setenv("VESCALE_DISABLE_REDISTRIBUTE", "0")
    out = a + b  # placements differ: one operand must be resharded (a no-op on a 1-rank mesh, but it goes through the planner)
    assert torch.allclose(out.full_tensor(), a.full_tensor() + b.full_tensor())
    monkeypatch.setenv("VESCALE_STRICT_RULES", "1")
    from vescale_b200.dtensor.sharding_prop import ShardingPropagator, _replicate_fallback
    from vescale_b200.dtensor.op_schema import OpSchema

    with pytest.raises(NotImplementedError):
        _replicate_fallback(OpSchema(torch.ops.aten.frac.default, (a._spec,), {}))


def test_utils_package():
    """monkey-patch helper (reference ``vescale/utils/monkey_patch.py``), env-flag registry covers every flag in the tree."""
    import re
    import subprocess

    from vescale_b200.utils import FLAGS, describe_flags, flag, patch_method, unpatch_all

    class T:
        def f(self):
            return 1

    @patch_method(T, "f")
    def f2(self):
        return 2

    @patch_method(T, "g")
    def g(self):
        return 3

    assert T().f() == 2 and T().g() == 3 and f2.__wrapped_original__(T()) == 1
    unpatch_all()
    assert T().f() == 1 and not hasattr(T, "g")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    used = set()
    for dp, _, fs in os.walk(os.path.join(root, "vescale_b200")):
        for fn in fs:
            if fn.endswith(".py"):
                used |= set(re.findall(r"VESCALE_[A-Z0-9_]+", open(os.path.join(dp, fn)).read()))
    used -= {"VESCALE_DEVICE_MESH", "VESCALE_INSTRUCTION_MAPPING_ZBV", "VESCALE_INTRUCTION_BUILDER"} - set(FLAGS)  # Python objects that share the prefix (reference names)
    assert used <= set(FLAGS), f"flags missing from utils.env.FLAGS: {sorted(used - set(FLAGS))}"
    assert "VESCALE_STRICT_RULES" in describe_flags() and flag("VESCALE_B200_SYMM_CHUNK_MB") == 2048


def test_root_level_api_surface():
    """SURVEY §7.5 checklist: the reference's root-level names resolve on ``vescale_b200`` and on the ``vescale`` alias."""
    import vescale
    import vescale_b200 as v

    for name in ("DTensor DeviceMesh init_device_mesh Placement Shard Replicate Partial InterleavedShard RaggedShard distribute_tensor redistribute_dtensor "
                 "from_local to_local normalize_placements parallelize_module is_dmodule PlacementsInterface auto_parallelize_module "
                 "set_plan_overriding_policy get_plan_overriding_policy DistributedDataParallel DistributedOptimizer BasicOptimizer BasicOptimizerHook "
                 "deferred_init is_deferred materialize_dtensor materialize_dparameter vescale_all_gather vescale_all_reduce vescale_reduce_scatter "
                 "checkpoint loss_parallel manual_seed fully_shard FSDPAdamW PipeEngine PipelineParallelPlan parallelize_experts").split():
        assert getattr(v, name) is not None, name
        assert getattr(vescale, name) is getattr(v, name), name
    import vescale_b200.dtensor as d

    for name in "distribute_tensor ones empty full rand randn zeros arange DTensorSpec TensorMeta RaggedShard _StridedRaggedShard _StridedShard _Partial is_ragged_shard equal allclose".split():
        assert hasattr(d, name), name
    assert repr(d.RaggedShard((0,), (1, 2))) == "RaggedShard(dims=(0,), local_units=(1, 2))"


def test_defer_resharding_toggle():
    """Partial + Partial stays Partial (one reduction later) by default; ``defer_resharding(False)`` reduces each operand
    first, the legacy order (``DeferReshardMode``, legacy ``_dispatch_patch.py:134``)."""
    from vescale_b200 import Partial, Replicate
    from vescale_b200.dtensor import DTensor, defer_resharding

    mesh = _mesh((1,))
    mk = lambda: DTensor.from_local(torch.ones(2, 2), mesh, [Partial()], run_check=False, shape=(2, 2))  # noqa: E731
    assert (mk() + mk()).placements == (Partial(),)
    with defer_resharding(False):
        assert (mk() + mk()).placements == (Replicate(),)
        with defer_resharding(True):
            assert (mk() + mk()).placements == (Partial(),)
    assert (mk() + mk() - mk()).placements == (Partial(),)


def test_fp8_block_scaled_linear():
    """Block-scaled e4m3 quantisation round trip, GEMM numerics vs fp32, autograd, and a tiny fp8 Llama step (emulated path;
    CUDA uses cuBLASLt through torch._scaled_mm with the same scales)."""
    from vescale_b200.models import LlamaConfig, LlamaModel
    from vescale_b200.ops.fp8 import FP8_MAX, dequantize_blockwise, fp8_linear, quantize_blockwise

    g = torch.Generator().manual_seed(0)
    x = torch.randn(70, 300, generator=g) * torch.logspace(-2, 2, 300)  # wide dynamic range across K blocks
    for blk in ((1, 128), (128, 128)):
        q, s = quantize_blockwise(x, blk)
        assert q.dtype == torch.float8_e4m3fn and q.shape == x.shape and s.shape == (-(-70 // blk[0]), -(-300 // blk[1]))
        back = dequantize_blockwise(q, s, blk)
        # e4m3 has 3 mantissa bits: relative error <= 2^-4 per element w.r.t. its block's amax-scaled grid
        blocks_amax = dequantize_blockwise(torch.full_like(q.float(), FP8_MAX).to(torch.float8_e4m3fn), s, blk)
        assert ((back - x).abs() <= blocks_amax / FP8_MAX * 32 + 1e-6).all()
        assert (back - x).abs().max() / x.abs().max() < 0.07
    w = torch.randn(96, 300, generator=g) * 0.05
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    y = fp8_linear(xr, wr)
    ref = x @ w.t()
    assert (y - ref).norm() / ref.norm() < 0.05
    y.sum().backward()
    torch.testing.assert_close(xr.grad, torch.ones(70, 96) @ w, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(wr.grad, torch.ones(96, 70) @ x, rtol=1e-4, atol=1e-3)
    cfg = LlamaConfig.tiny(fp8=True)
    m = LlamaModel(cfg).reset_parameters(seed=1)
    tok = torch.randint(0, cfg.vocab_size, (2, 17), generator=g)
    loss = m(tok[:, :-1], tok[:, 1:])
    loss.backward()
    ref_loss = LlamaModel(LlamaConfig.tiny()).reset_parameters(seed=1)(tok[:, :-1], tok[:, 1:])
    assert abs(loss.item() - ref_loss.item()) < 0.05 and all(p.grad is not None for p in m.parameters())


def test_bench_script_control_flow_on_cpu():
    """``bench.py --device cpu`` walks the whole script (build, warm-up, both timed regions, JSON contract) on the tiny model;
    the record is marked invalid.  Guards the driver-facing entry point against regressions that only a GPU run would show."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--device", "cpu", "--model", "tiny", "--seq-len", "64", "--micro-batch", "2", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=240, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e",
              "gpu_launches", "exposed_comm_ms_per_step", "step_ms"):
        assert k in rec, k
    assert rec["invalid"] and rec["steps"] == 2 and rec["e2e"]["h2d_bytes_per_step"] == 2 * 65 * 8 and rec["e2e"]["d2h_bytes_per_step"] == 4
    ref = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference"], capture_output=True, text=True, timeout=120, cwd=root)
    assert ref.returncode == 0 and json.loads(ref.stdout.strip().splitlines()[-1])["impl"] == "reference"


def test_symm_debug_tooling():
    """Poisoning and the epoch-invariant checker (on a stand-in arena; the real one needs peer-mapped GPU memory)."""
    from vescale_b200.comm.symm_debug import check_epochs, poison

    assert poison(torch.zeros(4)).isnan().all() and poison(torch.zeros(2, dtype=torch.int32))[0].item() == 0x5A5A5A5A

    class Arena:
        pad = torch.zeros(8 * 2, dtype=torch.int32)
        world = 2
        device = torch.device("cpu")
        group = None

    a = Arena()
    s0 = check_epochs(a, quiescent=False)
    a.pad[0], a.pad[1] = 3, 3
    s1 = check_epochs(a, s0, quiescent=False)
    a.pad[1] = 2  # an epoch flag must never go backwards
    with pytest.raises(AssertionError):
        check_epochs(a, s1, quiescent=False)
    a.pad[1] = 3
    a.pad[2], a.pad[3] = 7, 5  # sources of one slot two epochs apart at a quiescent point: a lost signal
    with pytest.raises(AssertionError):
        torch.distributed.is_initialized() or check_epochs(a, s1, quiescent=True)


def test_mxfp8_quantisation():
    """OCP MX (1x32 blocks, E8M0 power-of-two scales, e4m3 elements): round trip and emulated GEMM accuracy."""
    from vescale_b200.ops.fp8 import FP8_MAX, dequantize_mx, mxfp8_gemm_nt, quantize_mx

    g = torch.Generator().manual_seed(0)
    x = torch.randn(40, 200, generator=g) * torch.logspace(-3, 3, 200)
    q, s = quantize_mx(x)
    assert q.dtype == torch.float8_e4m3fn and s.dtype == torch.uint8 and s.shape == (40, 7)
    back = dequantize_mx(q, s)
    blk_amax = torch.nn.functional.pad(x, (0, 24)).view(40, 7, 32).abs().amax(2).repeat_interleave(32, 1)[:, :200]
    assert ((back - x).abs() <= blk_amax * 2.0**-3 + 1e-12).all()  # 3 mantissa bits relative to the block's top binade
    assert (q.float().abs() <= FP8_MAX).all()
    # power-of-two scales: dequantised values of exact powers of two are exact
    p2 = torch.tensor([[2.0**k for k in range(-10, 22)]])
    qp, sp = quantize_mx(p2)
    assert torch.equal(dequantize_mx(qp, sp)[:, -9:], p2[:, -9:])  # within e4m3's dynamic range of the block maximum
    w = torch.randn(64, 200, generator=g) * 0.05
    xq, xs = quantize_mx(x)
    wq, ws = quantize_mx(w)
    y = mxfp8_gemm_nt(xq, xs, wq, ws, torch.float32)
    ref = x @ w.t()
    assert (y - ref).norm() / ref.norm() < 0.05


def test_mxfp8_scale_atoms_and_recipe():
    """Scale factors in the tensor core's atom order (512 B = 128 rows x 4 K-blocks; csrc/gemm_mxfp8.cu) round-trip and follow
    the byte formula; fp8_linear(recipe="mx") and a tiny Llama step under MX numerics track the bf16 model."""
    from vescale_b200.models import LlamaConfig, LlamaModel
    from vescale_b200.ops.fp8 import fp8_linear, mx_scale_atoms, mx_scale_from_atoms, quantize_mx

    g = torch.Generator().manual_seed(3)
    x = torch.randn(300, 512, generator=g)
    _, s = quantize_mx(x)
    for mult, atoms_rows in ((128, 3), (256, 4)):
        a = mx_scale_atoms(s, mult)
        assert a.dtype == torch.uint8 and a.numel() == atoms_rows * 4 * 512
        assert torch.equal(mx_scale_from_atoms(a, 300, 16), s)
        for r, kb in ((0, 0), (31, 3), (32, 4), (200, 9), (299, 15)):
            off = ((r // 128) * 4 + kb // 4) * 512 + (r % 32) * 16 + (r % 128 // 32) * 4 + kb % 4
            assert a[off] == s[r, kb]
        assert (a.view(atoms_rows, 4, 32, 4, 4)[-1, :, :, (300 % 128) // 32 + 1 :, :] == 127).all()  # padding rows: scale 1.0
    w = torch.randn(96, 512, generator=g) * 0.05
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    y = fp8_linear(xr, wr, recipe="mx")
    ref = x @ w.t()
    assert (y - ref).norm() / ref.norm() < 0.06
    y.sum().backward()
    torch.testing.assert_close(wr.grad, torch.ones(96, 300) @ x, rtol=1e-4, atol=1e-3)
    cfg = LlamaConfig.tiny(fp8="mx")
    m = LlamaModel(cfg).reset_parameters(seed=1)
    tok = torch.randint(0, cfg.vocab_size, (2, 17), generator=g)
    loss = m(tok[:, :-1], tok[:, 1:])
    loss.backward()
    ref_loss = LlamaModel(LlamaConfig.tiny()).reset_parameters(seed=1)(tok[:, :-1], tok[:, 1:])
    assert abs(loss.item() - ref_loss.item()) < 0.05 and all(p.grad is not None for p in m.parameters())


def test_new_package_dispatch_names_and_bypass_table():
    """The new package's handler names and the bypass table object resolve to the functions the dispatcher runs
    (``vescale/dtensor/_dispatch.py``, ``legacy/vescale/dtensor/_dispatch_bypass.py:26``)."""
    import torch

    from vescale.dtensor._dispatch import found_inf_reduce_handler, fused_adamw_sgd_op_handler, is_contiguous, ragged_norm_op_handler
    from vescale.dtensor._dispatch_bypass import BypassOpDispatch
    from vescale.dtensor._dtensor_spec import is_ragged_shard
    from vescale.dtensor._redistribute import substitute_ragged_spec
    from vescale_b200 import DTensor, Replicate, init_device_mesh
    from vescale_b200.dtensor import handlers
    from vescale_b200.placement import RaggedShard
    from vescale_b200.spec import DTensorSpec, TensorMeta

    assert ragged_norm_op_handler is handlers.ragged_norm_handler and found_inf_reduce_handler is handlers.found_inf_handler
    assert fused_adamw_sgd_op_handler is handlers.fused_optimizer_handler
    mesh = init_device_mesh("cpu", (2,), _rank=0, _init_process_groups=False)
    sp = DTensorSpec(mesh, (RaggedShard((0,), (1, 3)),), TensorMeta((4, 4), (4, 1), torch.float32))
    assert is_ragged_shard(sp) and substitute_ragged_spec(sp).placements == (Replicate(),) and is_contiguous(sp)
    assert not is_contiguous(DTensorSpec(mesh, (Replicate(),), TensorMeta((4, 4), (1, 4), torch.float32)))
    a = DTensor.from_local(torch.ones(4, 4), mesh, [Replicate()])
    hit, same = BypassOpDispatch.apply(torch.ops.aten.is_same_size.default, (a, a), {})
    assert hit and same is True
    assert BypassOpDispatch.apply(torch.ops.aten.mm.default, (a, a), {}) == (False, None)
