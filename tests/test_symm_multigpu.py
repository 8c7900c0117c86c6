"""Symmetric-memory kernels on >= 2 GPUs of one box: collectives vs NCCL, FSDP(symm) vs FSDP(nccl)."""

import pytest
import torch
import torch.distributed as dist

from common import run_distributed

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _collectives(rank, world):
    from vescale_b200 import init_device_mesh
    from vescale_b200.comm.symm import SymmUnitComm
    from vescale_b200.ops import _ext

    _ext.load(required=True)
    dev = torch.device("cuda", rank)
    mesh = init_device_mesh("cuda", (world,))
    comm = SymmUnitComm(mesh, 0, dev, chunk_bytes=256 << 20)

    class U:  # minimal stand-in for FSDPUnit
        sumsq = None

    u = U()
    S = 64 * 4096 * 5
    shard = comm.alloc(S, torch.bfloat16)
    g = torch.Generator(device=dev).manual_seed(rank)
    shard.copy_(torch.randn(S, device=dev, generator=g).bfloat16())
    full = torch.empty(S * world, dtype=torch.bfloat16, device=dev)
    ref = torch.empty_like(full)
    dist.all_gather_into_tensor(ref, shard)
    for it in range(3):
        full.zero_()
        comm.all_gather(shard, full, u)
        torch.cuda.synchronize()
        assert torch.equal(full, ref), f"all_gather mismatch at iter {it}"
    # reduce-scatter: fp32 accumulate of bf16 peers, scaled, + sum of squares
    fg = comm.alloc(S * world, torch.bfloat16)
    for use_mm in (False, True):
        comm.use_multimem = use_mm
        for it in range(2):
            fg.copy_(torch.randn(S * world, device=dev, generator=g).bfloat16())
            torch.cuda.synchronize()
            dist.barrier()
            gathered = [torch.empty_like(fg) for _ in range(world)]
            dist.all_gather(gathered, fg)
            want = sum(t[rank * S : (rank + 1) * S].float() for t in gathered) * 0.5
            out = torch.empty(S, dtype=torch.float32, device=dev)
            comm.wait_buffer_free(fg)
            comm.reduce_scatter(fg, out, 0.5, u)
            torch.cuda.synchronize()
            tol = dict(rtol=1e-5, atol=1e-5) if not use_mm or comm.arena.multicast_ptr(fg) == 0 else dict(rtol=2e-2, atol=2e-2)
            torch.testing.assert_close(out, want, **tol)
            torch.testing.assert_close(u.sumsq[0], out.pow(2).sum(), rtol=1e-3, atol=1e-3)
            dist.barrier()
    if rank == 0:
        print(f"multicast available: {comm.arena.multicast_ptr(fg) != 0}", flush=True)


def _fsdp_backends(rank, world):
    from vescale_b200 import init_device_mesh
    from vescale_b200.models import LlamaConfig, LlamaModel
    from vescale_b200.optim import FSDPAdamW
    from vescale_b200.parallel.fsdp import fully_shard

    dev = torch.device("cuda", rank)
    cfg = LlamaConfig(vocab_size=2048, hidden_size=512, intermediate_size=1024, num_layers=3, num_heads=8, num_kv_heads=2, head_dim=64, max_seq_len=256)
    results = {}
    from vescale_b200.ops import _ext

    for backend, clip, fused in (("nccl", 1.0, False), ("symm", 1.0, False), ("symm", None, True), ("nccl", None, False), ("symm-wag", 1.0, False)):
        mesh = init_device_mesh("cuda", (world,))
        model = LlamaModel(cfg, device=dev).reset_parameters(seed=3)
        wag = backend == "symm-wag"  # exposed all-gathers (prefetch=0) with the unit's first GEMM gathering its own weight
        if wag:
            backend = "symm"
            _ext.LAUNCH_COUNTER.update(n=0, enabled=True, by_op={})
        for blk in model.layers:
            if wag:
                fully_shard(blk, mesh, comm_backend="symm", reshard_after_forward=True, prefetch=0, fuse_first_gemm=True)
                continue
            fully_shard(blk, mesh, comm_backend=backend, reshard_after_forward=(backend == "nccl"))
        fully_shard(model.embed, mesh, comm_backend=backend)
        fully_shard(model.head, mesh, comm_backend=backend)
        fully_shard(model, mesh, comm_backend=backend, **({"prefetch": 0} if wag else {}))
        opt = FSDPAdamW(model, lr=1e-3, max_grad_norm=clip, fused_reduce=fused)
        losses = []
        if rank == 0:
            print(f"[fsdp-backends] {backend} clip={clip} fused={fused} wag={wag}", flush=True)
        for s in range(4):
            g = torch.Generator().manual_seed(100 * s + rank)
            tok = torch.randint(0, cfg.vocab_size, (2, 257), generator=g).to(dev)
            loss = model(tok[:, :-1], tok[:, 1:])
            loss.backward()
            opt.step()
            opt.zero_grad()
            losses.append(loss.item())
            if rank == 0:
                print(f"   step {s} loss {losses[-1]:.4f}", flush=True)
        params = torch.cat([p.full_tensor().reshape(-1).float() for p in model.parameters()])
        results[("symm-wag" if wag else backend, clip, fused)] = (losses, params)
        if wag:
            n_wag = _ext.LAUNCH_COUNTER["by_op"].get("wag_gemm", 0)
            _ext.LAUNCH_COUNTER["enabled"] = False
            assert n_wag == 4 * cfg.num_layers, f"fused all-gather⊕GEMM ran {n_wag} times"
        del model, opt
        torch.cuda.synchronize()
        dist.barrier()
    for a, b in ((("nccl", 1.0, False), ("symm", 1.0, False)), (("nccl", None, False), ("symm", None, True)), (("nccl", 1.0, False), ("symm-wag", 1.0, False))):
        la, pa = results[a]
        lb, pb = results[b]
        assert all(abs(x - y) < 3e-2 for x, y in zip(la, lb)), (a, b, la, lb)
        rel = (pa - pb).norm() / pa.norm()
        assert rel < 2e-3, (a, b, rel.item())


@pytest.mark.timeout(240)
def test_symm_collectives_match_nccl():
    run_distributed(_collectives, min(torch.cuda.device_count(), 8), backend="nccl")


@pytest.mark.timeout(300)
def test_fsdp_symm_matches_nccl():
    run_distributed(_fsdp_backends, min(torch.cuda.device_count(), 8), backend="nccl")


def _fused_tp_setup(rank, world):
    from vescale_b200 import init_device_mesh
    from vescale_b200.comm.fused_tp import FusedTP
    from vescale_b200.ops import _ext

    _ext.load(required=True)
    dev = torch.device("cuda", rank)
    mesh = init_device_mesh("cuda", (world,), mesh_dim_names=("TP",))
    return FusedTP(mesh, "TP", dev), dev, torch.Generator(device=dev)


def _fused_tp_ag(rank, world):
    tp, dev, g = _fused_tp_setup(rank, world)
    for (Ml, K, Nr) in ((256, 512, 256), (1024, 4096, 3072 // world * 2), (512, 1024, 264)):
        M = Ml * world
        for it in range(3):
            g.manual_seed(1000 * it + rank)
            x_local = (torch.randn(Ml, K, device=dev, generator=g) * 0.5).bfloat16()
            w = (torch.randn(Nr, K, device=dev, generator=g) * 0.05).bfloat16()
            xs = [torch.empty_like(x_local) for _ in range(world)]
            dist.all_gather(xs, x_local)
            ref = torch.cat(xs).float() @ w.float().t()
            y, x_full = tp.ag_gemm(x_local, w)
            torch.cuda.synchronize()
            assert torch.equal(x_full, torch.cat(xs)), f"ag_gemm gathered buffer mismatch {Ml,K,Nr} it{it}"
            err = (y.float() - ref).abs().max().item()
            assert err < 0.02 * ref.abs().max().item() + 0.05, ("ag_gemm", Ml, K, Nr, it, err)
        if rank == 0:
            print(f"[fused_tp] ag_gemm {Ml, K, Nr} ok", flush=True)


def _fused_tp_rs(rank, world):
    tp, dev, g = _fused_tp_setup(rank, world)
    for (M, Kr, N) in ((256 * world, 512, 256), (2048 * world // 2, 2048, 4096), (512 * world, 1024, 264)):
        for it in range(3):
            g.manual_seed(77 * it + rank)
            x = (torch.randn(M, Kr, device=dev, generator=g) * 0.5).bfloat16()
            w = (torch.randn(N, Kr, device=dev, generator=g) * 0.05).bfloat16()
            part = (x.float() @ w.float().t())
            tot = part.clone()
            dist.all_reduce(tot)
            ref = tot[rank * (M // world) : (rank + 1) * (M // world)]
            y = tp.gemm_rs(x, w)
            torch.cuda.synchronize()
            err = (y.float() - ref).abs().max().item()
            assert err < 0.03 * ref.abs().max().item() + 0.05, ("gemm_rs", M, Kr, N, it, err)
            dist.barrier()
        if rank == 0:
            print(f"[fused_tp] gemm_rs {M, Kr, N} ok", flush=True)


def _fused_tp_autograd(rank, world):
    # column-parallel then row-parallel MLP under SP equals the single-device MLP
    tp, dev, g = _fused_tp_setup(rank, world)
    H, F, Ml = 512, 1024, 256
    g.manual_seed(5)
    w1 = (torch.randn(F, H, device=dev, generator=g) * 0.05).bfloat16()
    w2 = (torch.randn(H, F, device=dev, generator=g) * 0.05).bfloat16()
    g.manual_seed(100 + rank)
    x_local = torch.randn(Ml, H, device=dev, generator=g).bfloat16().requires_grad_()
    w1s = w1[rank * F // world : (rank + 1) * F // world].clone().requires_grad_()
    w2s = w2[:, rank * F // world : (rank + 1) * F // world].clone().requires_grad_()
    out = tp.linear_rs(torch.relu(tp.ag_linear(x_local, w1s)), w2s)
    torch.cuda.synchronize()
    if rank == 0:
        print("[fused_tp] forward done", flush=True)
    out.float().pow(2).sum().backward()
    torch.cuda.synchronize()
    if rank == 0:
        print("[fused_tp] backward done", flush=True)
    xs = [torch.empty_like(x_local) for _ in range(world)]
    dist.all_gather(xs, x_local.detach())
    xf = torch.cat(xs).float().requires_grad_()
    w1f, w2f = w1.float().requires_grad_(), w2.float().requires_grad_()
    ref = torch.relu(xf @ w1f.t()) @ w2f.t()
    ref.pow(2).sum().backward()
    sl = slice(rank * Ml, (rank + 1) * Ml)
    assert (out.float() - ref[sl]).abs().max().item() < 0.05 * ref.abs().max().item() + 0.05
    assert (x_local.grad.float() - xf.grad[sl]).abs().max().item() < 0.08 * xf.grad.abs().max().item() + 0.05
    assert (w1s.grad.float() - w1f.grad[rank * F // world : (rank + 1) * F // world]).abs().max().item() < 0.08 * w1f.grad.abs().max().item() + 0.05


@pytest.mark.timeout(240)
def test_fused_tp_ag_gemm():
    run_distributed(_fused_tp_ag, min(torch.cuda.device_count(), 8), backend="nccl")


@pytest.mark.timeout(240)
def test_fused_tp_gemm_rs():
    run_distributed(_fused_tp_rs, min(torch.cuda.device_count(), 8), backend="nccl")


@pytest.mark.timeout(240)
def test_fused_tp_autograd():
    run_distributed(_fused_tp_autograd, min(torch.cuda.device_count(), 8), backend="nccl")


def _symm_moe(rank, world):
    from vescale_b200 import init_device_mesh
    from vescale_b200.ops import _ext
    from vescale_b200.parallel.moe import MoEConfig, MoELayer
    from vescale_b200.parallel.moe.symm_dispatch import SymmMoEDispatcher

    _ext.load(required=True)
    dev = torch.device("cuda", rank)
    mesh = init_device_mesh("cuda", (world,), mesh_dim_names=("EP",))
    H, F, E, k, T = 512, 1024, 2 * world, 2, 384
    cfg = MoEConfig(H, F, E, k, dtype=torch.bfloat16)
    g = torch.Generator(device=dev).manual_seed(11)
    layer = MoELayer(cfg, mesh.get_group(0), device=dev)
    layer.reset_parameters(g)
    disp = SymmMoEDispatcher(mesh, E, H, F, max_tokens=T, top_k=k, capacity_factor=float(world), device=dev)
    for it in range(3):
        gx = torch.Generator(device=dev).manual_seed(100 * it + rank)
        x = torch.randn(T, H, device=dev, generator=gx).bfloat16()
        xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
        layer.symm_dispatcher = None
        ya = layer(xa)  # NCCL all_to_all_single path (baseline)
        ya.float().pow(2).sum().backward()
        ga = {n: p.grad.clone() for n, p in layer.named_parameters()}
        layer.zero_grad()
        layer.use_symmetric_dispatch(disp)
        yb = layer(xb)
        yb.float().pow(2).sum().backward()
        torch.cuda.synchronize()
        scale = ya.float().abs().max().item()
        assert (ya.float() - yb.float()).abs().max().item() < 0.03 * scale + 1e-3, (it, "fwd")
        assert (xa.grad.float() - xb.grad.float()).abs().max().item() < 0.05 * xa.grad.float().abs().max().item() + 1e-3, (it, "dx")
        for n, p in layer.named_parameters():
            ref = ga[n].float()
            assert (p.grad.float() - ref).abs().max().item() < 0.06 * ref.abs().max().item() + 1e-3, (it, n)
        layer.zero_grad()
        dist.barrier()


@pytest.mark.timeout(240)
def test_symm_moe_dispatch_matches_nccl():
    run_distributed(_symm_moe, min(torch.cuda.device_count(), 8), backend="nccl")


def _tensor_collectives(rank, world):
    """csrc/symm_collectives.cu: all-reduce (one-shot / two-shot / NVLS), a2a with folded permutes, ragged puts, vocab CE —
    each against NCCL / a single-device fp32 reference — and DTensor.redistribute routed through them."""
    from vescale_b200 import DTensor, Replicate, Shard, init_device_mesh
    from vescale_b200.comm import collectives as C
    from vescale_b200.comm.symm_collectives import disable_symmetric_collectives, enable_symmetric_collectives
    from vescale_b200.dtensor import RaggedShard
    from vescale_b200.ops import _ext

    _ext.load(required=True)
    dev = torch.device("cuda", rank)
    mesh = init_device_mesh("cuda", (world,))
    (sc,) = enable_symmetric_collectives(mesh, reserve_bytes=64 << 20)
    g = torch.Generator(device=dev).manual_seed(100 + rank)

    # ---- all-reduce
    for dtype in (torch.float32, torch.bfloat16):
        for n in (7, 4096, 1 << 20, (1 << 22) + 24):
            for mm in (False, True):
                sc.use_multimem = mm
                x = torch.randn(n, device=dev, generator=g).to(dtype)
                want = x.clone().float()
                dist.all_reduce(want)
                got = sc.all_reduce(x.clone())
                torch.cuda.synchronize()
                tol = dict(rtol=2e-2, atol=3e-2) if dtype == torch.bfloat16 else dict(rtol=1e-5, atol=1e-5)
                torch.testing.assert_close(got.float(), want, **tol)
                if n % 8 == 0 and mm:  # operand resident in symmetric memory: zero-copy path
                    xs = sc.empty(n, dtype)
                    xs.copy_(x)
                    got = sc.all_reduce(xs)
                    torch.cuda.synchronize()
                    torch.testing.assert_close(got.float(), want, **tol)
    x = torch.randn(3, 1000, device=dev, generator=g)
    want = x.clone()
    dist.all_reduce(want)
    torch.testing.assert_close(C.mesh_all_reduce(x, mesh, "avg", 0), want / world, rtol=1e-5, atol=1e-5)

    # ---- Shard(i) -> Shard(j) with the permutes folded into the put
    for shape, i, j, dtype in (((4 * world, 6, 8 * world), 0, 2, torch.bfloat16), ((3, 2 * world, 5, 4 * world, 8), 3, 1, torch.float32), ((2 * world, 3 * world), 1, 0, torch.bfloat16)):
        full = torch.arange(int(torch.tensor(shape).prod()), device=dev, dtype=torch.float32).reshape(shape).to(dtype)
        mine = full.chunk(world, dim=i)[rank].contiguous()
        got = sc.all_to_all_permute(mine, i, j)
        torch.cuda.synchronize()
        assert torch.equal(got, full.chunk(world, dim=j)[rank]), (shape, i, j)
        dt = DTensor.from_local(mine, mesh, [Shard(i)])
        assert torch.equal(dt.redistribute(mesh, [Shard(j)]).to_local(), full.chunk(world, dim=j)[rank])

    # ---- ragged interval exchange (RaggedShard -> RaggedShard, gather to root)
    numel = 512 * world * (world + 1)
    full = torch.randn(numel, device=dev, generator=torch.Generator(device=dev).manual_seed(5)).bfloat16()
    src_units = tuple(range(1, world + 1))
    dst_units = tuple(reversed(src_units))
    root_units = tuple(1 if r == world - 1 else 0 for r in range(world))
    a = DTensor.from_local(full[slice(*RaggedShard((0,), src_units).flat_range(numel, rank))].clone(), mesh, [RaggedShard((0,), src_units)], shape=(numel,), stride=(1,))
    for units in (dst_units, root_units, src_units):
        b = a.redistribute(mesh, [RaggedShard((0,), units)])
        lo, hi = RaggedShard((0,), units).flat_range(numel, rank)
        torch.cuda.synchronize()
        assert torch.equal(b.to_local().reshape(-1), full[lo:hi]), units

    # ---- vocab-parallel cross entropy: one launch vs the fp32 single-device reference
    T, V = 1000, 1024 * world
    logits = torch.randn(T, V, device=dev, generator=torch.Generator(device=dev).manual_seed(9)).mul(3).bfloat16()
    target = torch.randint(0, V, (T,), device=dev, generator=torch.Generator(device=dev).manual_seed(10))
    target[::17] = -100
    ref_logits = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_logits, target, ignore_index=-100)
    ref.backward()
    for it in range(3):
        shard = logits[:, rank * (V // world) : (rank + 1) * (V // world)].contiguous().requires_grad_(True)
        buf = shard * 1.0  # non-leaf buffer the kernel may consume
        loss = sc.vocab_parallel_cross_entropy(buf, target)
        loss.backward()
        torch.cuda.synchronize()
        assert abs(loss.item() - ref.item()) < 2e-3, (it, loss.item(), ref.item())
        want = ref_logits.grad[:, rank * (V // world) : (rank + 1) * (V // world)]
        assert (shard.grad.float() - want).abs().max().item() < 2e-5 + 0.01 * want.abs().max().item()
    disable_symmetric_collectives()
    dist.barrier()


@pytest.mark.timeout(300)
def test_symm_tensor_collectives():
    run_distributed(_tensor_collectives, min(torch.cuda.device_count(), 8), backend="nccl")


def _llama_tp_fused(rank, world):
    """2-D Llama (FSDP x TP) with every TP collective fused into its GEMM (fwd and bwd) vs the same model on NCCL + cuBLAS."""
    from vescale_b200 import init_device_mesh
    from vescale_b200.comm.fused_tp import FusedTP, PlainTP
    from vescale_b200.models import LlamaConfig
    from vescale_b200.models.llama_tp import LlamaTPModel
    from vescale_b200.ops import _ext
    from vescale_b200.optim import FSDPAdamW
    from vescale_b200.parallel.fsdp import fully_shard

    _ext.load(required=True)
    dev = torch.device("cuda", rank)
    tp_size = 2
    dp = world // tp_size
    cfg = LlamaConfig(vocab_size=2048, hidden_size=512, intermediate_size=1024, num_layers=2, num_heads=8, num_kv_heads=2, head_dim=64, max_seq_len=512)
    mesh = init_device_mesh("cuda", (dp, tp_size), mesh_dim_names=("dp", "tp"))
    res = {}
    for impl in ("plain", "fused"):
        tp = FusedTP(mesh, "tp", dev) if impl == "fused" else PlainTP(mesh, "tp")
        model = LlamaTPModel(cfg, tp, device=dev).reset_parameters(seed=5)
        for blk in model.layers:
            fully_shard(blk, mesh, mesh_dim="dp")
        fully_shard(model.embed, mesh, mesh_dim="dp")
        fully_shard(model.head, mesh, mesh_dim="dp")
        fully_shard(model, mesh, mesh_dim="dp")
        opt = FSDPAdamW(model, lr=1e-3, max_grad_norm=1.0, tp_group=mesh.get_group("tp"))
        _ext.LAUNCH_COUNTER.update(n=0, enabled=True, by_op={})
        losses = []
        for s in range(4):
            g = torch.Generator().manual_seed(100 * s + mesh.get_local_rank("dp"))
            tok = torch.randint(0, cfg.vocab_size, (2, 513), generator=g).to(dev)
            share = model(tok[:, :-1], tok[:, 1:])
            share.backward()
            opt.step()
            opt.zero_grad()
            losses.append(model.loss_for_logging(share).item())
        by_op = dict(_ext.LAUNCH_COUNTER["by_op"])
        _ext.LAUNCH_COUNTER["enabled"] = False
        if impl == "fused":  # 2 ag + 2 rs per block forward, and the duals in backward
            assert by_op.get("ag_gemm", 0) == 4 * cfg.num_layers * 4 and by_op.get("gemm_rs", 0) == 4 * cfg.num_layers * 4, by_op
        res[impl] = losses
        if rank == 0:
            print(f"[llama-tp] {impl}: {losses}", flush=True)
        del model, opt
        torch.cuda.synchronize()
        dist.barrier()
    assert all(abs(a - b) < 3e-2 for a, b in zip(res["plain"], res["fused"])), res


@pytest.mark.timeout(300)
def test_llama_tp_fused_matches_plain():
    run_distributed(_llama_tp_fused, min(torch.cuda.device_count(), 8) // 2 * 2, backend="nccl")


def _reduce_scatter_nvls(rank, world):
    """Tensor-level reduce-scatter kernel (P2P pull and NVLS) against NCCL, strided output, and the gemm_rs built on it."""
    from vescale_b200 import init_device_mesh
    from vescale_b200.comm.fused_tp import FusedTP
    from vescale_b200.comm.symm_collectives import SymmCollectives
    from vescale_b200.ops import _ext

    _ext.load(required=True)
    dev = torch.device("cuda", rank)
    mesh = init_device_mesh("cuda", (world,))
    sc = SymmCollectives(mesh, 0, dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for dtype, tol in ((torch.bfloat16, 2e-2), (torch.float32, 1e-5)):
        for rows, row in ((3, 8), (64, 1024), (257, 4096)):
            x = torch.randn(world * rows, row, device=dev, generator=g).to(dtype)
            want = torch.empty(rows, row, device=dev, dtype=dtype)
            dist.reduce_scatter_tensor(want, x.clone())
            xs = sc.empty((world * rows, row), dtype)
            xs.copy_(x)
            for mm in (False, True):
                sc.use_multimem = mm
                for src in (x, xs):  # staged and zero-copy
                    got = sc.reduce_scatter(src)
                    torch.testing.assert_close(got.float(), want.float(), rtol=tol, atol=tol * 4)
            wide = torch.zeros(rows, 2 * row + 8, device=dev, dtype=dtype)
            sc.reduce_scatter(xs, "avg", out=wide[:, 8 : 8 + row])
            torch.testing.assert_close(wide[:, 8 : 8 + row].float(), want.float() / world, rtol=tol, atol=tol * 4)
            assert wide[:, :8].abs().max().item() == 0 and wide[:, 8 + row :].abs().max().item() == 0
    # back-to-back calls reuse the same buffer without host synchronisation
    xs = sc.empty((world * 128, 512))
    for it in range(5):
        xs.fill_(float(it + rank))
        got = sc.reduce_scatter(xs)
        assert torch.all(got.float() == float(sum(it + r for r in range(world)))), it
    tp = FusedTP(mesh, 0, dev, rs_impl="nvls")
    M, Kr, N = 256 * world * 2, 512, 1024
    x = (torch.randn(M, Kr, device=dev, generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, Kr, device=dev, generator=g) * 0.05).bfloat16()
    want = torch.empty(M // world, N, device=dev, dtype=torch.float32)
    dist.reduce_scatter_tensor(want, x.float() @ w.float().t())
    for it in range(3):
        y = tp.gemm_rs(x, w)
        assert (y.float() - want).abs().max().item() < 0.02 * want.abs().max().item() + 0.05, it
    torch.cuda.synchronize()
    dist.barrier()


@pytest.mark.timeout(300)
def test_symm_reduce_scatter_and_nvls_gemm_rs():
    run_distributed(_reduce_scatter_nvls, min(torch.cuda.device_count(), 8), backend="nccl")


def _dtensor_fusion(rank, world):
    """``redistribute(Shard -> Replicate) -> mm`` and ``mm(Partial) -> redistribute(-> Shard)`` issued as plain DTensor ops hit the
    fused sm_100a kernels (dtensor/fusion.py) and match the unfused DTensor path."""
    import os

    from vescale_b200 import Shard, distribute_tensor, init_device_mesh
    from vescale_b200.dtensor import fusion
    from vescale_b200.ops import _ext

    _ext.load(required=True)
    dev = torch.device("cuda", rank)
    mesh = init_device_mesh("cuda", (world,), mesh_dim_names=("TP",))
    g = torch.Generator(device=dev).manual_seed(0)
    M, H, Fd = 512 * world, 1024, 512 * world
    x = (torch.randn(M, H, device=dev, generator=g) * 0.5).bfloat16()
    w1 = (torch.randn(Fd, H, device=dev, generator=g) * 0.05).bfloat16()
    w2 = (torch.randn(H, Fd, device=dev, generator=g) * 0.05).bfloat16()
    dist.broadcast(x, 0), dist.broadcast(w1, 0), dist.broadcast(w2, 0)
    outs = {}
    for mode in ("auto", "off"):
        os.environ["VESCALE_B200_FUSE_TP"] = mode
        dx = distribute_tensor(x, mesh, [Shard(0)], src_data_rank=None)
        d1 = distribute_tensor(w1, mesh, [Shard(0)], src_data_rank=None)
        d2 = distribute_tensor(w2, mesh, [Shard(1)], src_data_rank=None)
        fusion.reset_stats()
        h = torch.relu(dx @ d1.t())
        with fusion.fuse_reshard(mesh, [Shard(0)]):
            y = h @ d2.t()
        y = y.redistribute(mesh, [Shard(0)])
        torch.cuda.synchronize()
        if mode == "auto":
            assert fusion.stats["ag_gemm"] == 1 and fusion.stats["gemm_rs"] == 1, fusion.stats
        else:
            assert fusion.stats["ag_gemm"] == 0 and fusion.stats["gemm_rs"] == 0, fusion.stats
        outs[mode] = y.full_tensor().float()
    os.environ["VESCALE_B200_FUSE_TP"] = "auto"
    ref = torch.relu(x.float() @ w1.float().t()).bfloat16().float() @ w2.float().t()
    for mode, o in outs.items():
        assert (o - ref).abs().max().item() < 0.03 * ref.abs().max().item() + 0.05, mode


@pytest.mark.timeout(200)
def test_dtensor_level_fusion_hits_fused_kernels():
    run_distributed(_dtensor_fusion, min(torch.cuda.device_count(), 8), backend="nccl")


def _emulator_vs_nccl(rank, world):
    """The emulator against LIVE NCCL (legacy ``test/emulator/test_distributed.py:72-101`` asserts ``torch.equal`` between the
    two): ring all-reduce / reduce-scatter with the Simple protocol on one channel.  The ring order NCCL chose on this box is
    discovered once (the reference reads it from a graph dump, ``emulator/distributed.py:741-809``): the same candidate must
    reproduce every message size bit for bit."""
    import itertools

    from vescale_b200.emulator.collectives import ring_all_reduce, ring_reduce_scatter

    dev = torch.device("cuda", rank)
    g = torch.Generator(device=dev)
    cases = []
    for n_per in (1024, 4096, 65536):
        count = n_per * world
        g.manual_seed(1234 + rank + n_per)
        x = (torch.randn(count, device=dev, generator=g) * 3.0).float()
        gathered = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(gathered, x)
        y = x.clone()
        dist.all_reduce(y)
        rs = torch.empty(n_per, device=dev)
        dist.reduce_scatter_tensor(rs, x.clone())
        cases.append((n_per, [t.cpu() for t in gathered], y.cpu(), rs.cpu()))
    if rank != 0:
        return
    rings = [[0] + list(p) for p in itertools.permutations(range(1, world))] if world <= 5 else [list(range(world)), list(range(world - 1, -1, -1))]
    good = None
    for ring in rings:
        if all(torch.equal(ring_all_reduce(ins, "sum", ring=ring, nchannels=1, chunk_elems=n_per)[0], y) for n_per, ins, y, _ in cases):
            good = ring
            break
    assert good is not None, "no ring order reproduces NCCL's all-reduce bit for bit"
    for n_per, ins, _, rs in cases:
        assert torch.equal(ring_reduce_scatter(ins, "sum", ring=good)[0], rs), ("reduce_scatter", n_per, good)
    print(f"[emulator] NCCL ring order on this box: {good}; all_reduce / reduce_scatter bit-identical for {[c[0] * world for c in cases]} fp32 elements", flush=True)


@pytest.mark.timeout(240)
def test_emulator_bitwise_equals_live_nccl(monkeypatch):
    # one channel, ring, Simple protocol: the configuration whose schedule the emulator reproduces
    for k, v in (("NCCL_ALGO", "Ring"), ("NCCL_PROTO", "Simple"), ("NCCL_MAX_NCHANNELS", "1"), ("NCCL_MIN_NCHANNELS", "1"), ("NCCL_NVLS_ENABLE", "0")):
        monkeypatch.setenv(k, v)
    run_distributed(_emulator_vs_nccl, min(torch.cuda.device_count(), 4), backend="nccl")


# Written after this round's GPU budget was spent: these bodies have only ever run on gloo.  A test nobody has seen pass on NCCL is not
# evidence, so it is opt-in until it has been (VESCALE_B200_RUN_UNVALIDATED_GPU_TESTS=1 on a >= 2 GPU box).
_unvalidated = pytest.mark.skipif(
    __import__("os").environ.get("VESCALE_B200_RUN_UNVALIDATED_GPU_TESTS", "0") != "1",
    reason="never run on NCCL yet (added after the GPU budget was spent); set VESCALE_B200_RUN_UNVALIDATED_GPU_TESTS=1",
)


@_unvalidated
@pytest.mark.timeout(240)
@pytest.mark.parametrize("batch", [False, True])
def test_pipeline_engine_on_gpus_overlapped_nccl_p2p(batch):
    """The pipeline engine on real GPUs: 1F1B over NCCL p2p with receives posted ahead (and batched send+recv groups), loss equal to
    the single-device model, every instruction on the ndtimeline (CUDA-event timers).  Same body as the gloo test."""
    from test_pipe_dist import _pp_overlap

    run_distributed(_pp_overlap, min(torch.cuda.device_count(), 4), batch, backend="nccl")


@_unvalidated
@pytest.mark.timeout(240)
@pytest.mark.parametrize("sched", ["ZERO_BUBBLE", "INTERLEAVED_1F1B"])
def test_pipeline_schedules_on_gpus(sched):
    from test_pipe_dist import _pp

    run_distributed(_pp, min(torch.cuda.device_count(), 4) // 2 * 2, sched, "STRUCTURAL", backend="nccl")
