"""GPU tests added after the last hardware session of round 1 (fp8 GEMM path, fused sharded dropout kernel, world-size-1
run of the symmetric-memory kernels).  Kept in a file that sorts after the others so that ``pytest -x`` reaches every
hardware-validated test first; same conventions as ``test_kernels_gpu.py``."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fp8_block_scaled_gemm_on_cuda():
    """fp8 forward GEMM on the GPU (cuBLASLt through torch._scaled_mm when the build accepts the scale layout, emulation
    otherwise) against the fp32 product; the emulated path is the numerics specification."""
    from vescale_b200.ops import fp8

    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(512, 1024, device=dev, generator=g).bfloat16()
    w = (torch.randn(768, 1024, device=dev, generator=g) * 0.05).bfloat16()
    ref = x.float() @ w.float().t()
    xq, xs = fp8.quantize_blockwise(x, (1, 128))
    wq, ws = fp8.quantize_blockwise(w, (128, 128))
    emu = fp8._emulated_gemm_nt(xq, xs, wq, ws, torch.float32)
    assert ((emu - ref).norm() / ref.norm()).item() < 0.05
    got = fp8.fp8_gemm_nt(xq, xs, wq, ws, torch.bfloat16).float()
    assert ((got - ref).norm() / ref.norm()).item() < 0.08
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    fp8.fp8_linear(xr, wr).float().sum().backward()
    assert xr.grad is not None and wr.grad is not None and torch.isfinite(xr.grad.float()).all()


def test_fused_sharded_dropout_matches_composite():
    """The fused Philox dropout kernel draws the same mask and values as the specification path (uniform fill, compare,
    scale), for a full tensor and for a shard of it viewed through a fake 4-rank mesh."""
    from vescale_b200 import Replicate, Shard, init_device_mesh
    from vescale_b200.dtensor import DTensor
    from vescale_b200.dtensor import random as R
    from vescale_b200.ops import philox

    if not philox.dropout_available():
        pytest.skip("extension built without philox_dropout_box")
    dev = torch.device("cuda")
    full = torch.randn(64, 96, device=dev).bfloat16()
    outs = {}
    for rank in (0, 2):
        mesh = init_device_mesh("cuda", (4,), _rank=rank, _init_process_groups=False)
        for pl in ([Replicate()], [Shard(0)], [Shard(1)]):
            local = full if isinstance(pl[0], Replicate) else full.chunk(4, dim=pl[0].dim)[rank].contiguous()
            spec = DTensor.from_local(local, mesh, pl, run_check=False)._spec
            tr = R.ThreadBasedRNGTracker()
            R.manual_seed(123)
            fused_out, fused_mask = tr.run(torch.ops.aten.native_dropout.default, [local, 0.3, True], {}, spec)
            R.manual_seed(123)
            u = R.sharded_random_fill(torch.empty(local.shape, dtype=torch.float32, device=dev), spec, "uniform")
            mask = u >= 0.3
            ref = local * mask.to(local.dtype) * (1.0 / 0.7)
            assert torch.equal(fused_mask, mask) and torch.equal(fused_out, ref), (rank, pl)
            outs[(rank, str(pl))] = (fused_out, fused_mask)
    # shards agree with the replicated result at their global positions
    rep_out, rep_mask = outs[(0, str([Replicate()]))]
    assert torch.equal(outs[(2, str([Shard(0)]))][0], rep_out.chunk(4, 0)[2]) and torch.equal(outs[(2, str([Shard(1)]))][1], rep_mask.chunk(4, 1)[2])
    assert 0.2 < (~rep_mask).float().mean().item() < 0.4


def test_symmetric_memory_kernels_single_rank(tmp_path):
    """World-size-1 run of the symmetric-memory / fused-collective kernels on one GPU: every peer table has one entry (this
    GPU), so the flag protocols, copy engines and epilogues execute end to end and must reproduce the plain single-device
    result.  (The multi-GPU behaviour is covered by tests/test_symm_multigpu.py on >= 2 GPUs.)"""
    import torch.distributed as dist

    from vescale_b200 import init_device_mesh

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    created = False
    try:
        if not dist.is_initialized():
            dist.init_process_group("nccl", init_method=f"file://{tmp_path}/store", rank=0, world_size=1, device_id=dev)
            created = True
        mesh = init_device_mesh("cuda", (1,))
        from vescale_b200.comm.fused_tp import FusedTP
        from vescale_b200.comm.symm_collectives import SymmCollectives

        sc = SymmCollectives(mesh, 0, dev)
        tp = FusedTP(mesh, 0, dev)
    except Exception as e:  # noqa: BLE001  — symmetric memory unavailable in this environment
        if created:
            dist.destroy_process_group()
        pytest.skip(f"symmetric memory not available for a single-rank group: {type(e).__name__}: {e}")
    try:
        g = torch.Generator(device=dev).manual_seed(0)
        # all-reduce: one-shot, two-shot P2P, two-shot NVLS (if a multicast mapping exists), zero-copy
        for n in (1000, 1 << 20):
            x = torch.randn(n, device=dev, generator=g).bfloat16()
            for mm in (False, True):
                sc.use_multimem = mm
                torch.testing.assert_close(sc.all_reduce(x.clone(), "sum").float(), x.float(), rtol=1e-2, atol=1e-2)
            xs = sc.empty(n)
            xs.copy_(x)
            torch.testing.assert_close(sc.all_reduce(xs, "avg").float(), x.float(), rtol=1e-2, atol=1e-2)
        # reduce-scatter with one rank is a (scaled) copy: P2P and NVLS forms, staged and zero-copy, strided output
        xr = torch.randn(96, 256, device=dev, generator=g).bfloat16()
        xrs = sc.empty((96, 256))
        xrs.copy_(xr)
        for mm in (False, True):
            sc.use_multimem = mm
            assert torch.equal(sc.reduce_scatter(xr), xr) and torch.equal(sc.reduce_scatter(xrs), xr)
        wide = torch.zeros(96, 512, device=dev, dtype=torch.bfloat16)
        sc.reduce_scatter(xrs, out=wide[:, 128:384])
        assert torch.equal(wide[:, 128:384], xr) and wide[:, :128].abs().max().item() == 0
        # Shard(i) -> Shard(j) with one rank is the identity; ragged exchange likewise
        t = torch.randn(4, 6, 8, device=dev, generator=g).bfloat16()
        assert torch.equal(sc.all_to_all_permute(t, 0, 2), t) and torch.equal(sc.all_to_all_permute(t, 2, 1), t)
        flat = torch.randn(4096, device=dev, generator=g)
        assert torch.equal(sc.ragged_exchange(flat, [(0, 4096)], [(0, 4096)]), flat)
        # vocab-parallel CE with the whole vocabulary local == ordinary CE
        T, V = 300, 1024
        logits = (torch.randn(T, V, device=dev, generator=g) * 3).bfloat16()
        target = torch.randint(0, V, (T,), device=dev, generator=g)
        target[::7] = -100
        ref_l = logits.float().requires_grad_()
        ref = torch.nn.functional.cross_entropy(ref_l, target, ignore_index=-100)
        ref.backward()
        buf = logits.clone()
        loss = sc.vocab_parallel_cross_entropy(buf.requires_grad_() * 1.0, target)
        assert abs(loss.item() - ref.item()) < 2e-3
        # fused TP kernels: all-gather(x) @ W^T and reduce-scatter(x @ W^T) degenerate to the plain product
        x = (torch.randn(512, 1024, device=dev, generator=g) * 0.5).bfloat16()
        w = (torch.randn(768, 1024, device=dev, generator=g) * 0.05).bfloat16()
        want = x.float() @ w.float().t()
        y, x_full = tp.ag_gemm(x, w)
        assert torch.equal(x_full, x)
        assert (y.float() - want).abs().max().item() < 0.02 * want.abs().max().item() + 0.05
        y2 = tp.gemm_rs(x, w)
        assert (y2.float() - want).abs().max().item() < 0.02 * want.abs().max().item() + 0.05
        y3 = FusedTP(mesh, 0, dev, rs_impl="nvls").gemm_rs(x, w)  # GEMM into symmetric memory + reduce-scatter kernel
        assert (y3.float() - want).abs().max().item() < 0.02 * want.abs().max().item() + 0.05
        # FSDP all-gather ⊕ first GEMM: the kernel copies the (single) shard into the gathered buffer while multiplying
        from vescale_b200.models import LlamaConfig
        from vescale_b200.models.llama import LlamaBlock
        from vescale_b200.parallel.fsdp.unit import FSDPUnit, MixedPrecisionPolicy

        cfg = LlamaConfig(vocab_size=2048, hidden_size=512, intermediate_size=1024, num_layers=1, num_heads=8, num_kv_heads=2, head_dim=64, max_seq_len=256)
        blk = LlamaBlock(cfg, 0, device=dev)
        blk.reset_parameters(g)
        wq = blk.wqkv.detach().clone()
        from vescale_b200.comm.symm import get_unit_comm

        comm = get_unit_comm(mesh, 0, dev)
        unit = FSDPUnit(blk, list(blk.named_parameters()), mesh, 0, MixedPrecisionPolicy(), name="blk", comm=comm, block_rows=32)
        slot = comm.fusable_slot(unit, "wqkv")
        assert slot is not None
        full = torch.full((unit.S,), float("nan"), dtype=torch.bfloat16, device=dev)
        xin = (torch.randn(512, 512, device=dev, generator=g) * 0.5).bfloat16()
        yq = comm.fused_first_linear(xin, unit, slot, full)
        torch.cuda.synchronize()
        assert torch.equal(full[slot.offset : slot.end].view(slot.shape), wq)
        wantq = xin.float() @ wq.float().t()
        assert (yq.float() - wantq).abs().max().item() < 0.02 * wantq.abs().max().item() + 0.05
        # ---- FSDP default-path kernels with one rank: all-gather (SM pull and copy-engine forms, whole
        # unit / only-range / skip-range), reduce-scatter ⊕ scale ⊕ fp32 ⊕ sum-of-squares (P2P and NVLS forms), reduce-scatter ⊕ AdamW
        S = unit.S
        shard = unit.param_shard
        ref_shard = shard.clone()
        lo, hi = slot.offset, slot.end
        for impl in ("pull", "ce"):
            comm.ag_impl = impl
            full2 = torch.full((S,), float("nan"), dtype=torch.bfloat16, device=dev)
            comm.all_gather(shard, full2, unit)
            torch.cuda.synchronize()
            assert torch.equal(full2, ref_shard), impl
            full3 = torch.zeros(S, dtype=torch.bfloat16, device=dev)
            comm.all_gather(shard, full3, unit, only=(lo, hi))
            torch.cuda.synchronize()
            assert torch.equal(full3[lo:hi], ref_shard[lo:hi]) and full3[:lo].abs().sum().item() == 0 and full3[hi:].abs().sum().item() == 0, impl
            full4 = torch.zeros(S, dtype=torch.bfloat16, device=dev)
            comm.all_gather(shard, full4, unit, skip=(lo, hi))
            torch.cuda.synchronize()
            assert torch.equal(full4[:lo], ref_shard[:lo]) and torch.equal(full4[hi:], ref_shard[hi:]) and full4[lo:hi].abs().sum().item() == 0, impl
        fg = unit._alloc_full(torch.bfloat16, symmetric=True)
        fg.copy_((torch.randn(S, device=dev, generator=g) * 0.1).bfloat16())
        want_g = fg.float() * 0.5
        for mm in (False, True):
            comm.use_multimem = mm
            out = torch.zeros(S, dtype=torch.float32, device=dev)
            comm.reduce_scatter(fg, out, 0.5, unit)
            torch.cuda.synchronize()
            torch.testing.assert_close(out, want_g, rtol=1e-2, atol=1e-3)
            torch.testing.assert_close(unit.sumsq[0], want_g.pow(2).sum(), rtol=2e-2, atol=1e-3)
        comm.use_multimem = True
        # reduce-scatter ⊕ AdamW against torch.optim.AdamW on the same fp32 master
        unit.exp_avg = torch.zeros(S, dtype=torch.float32, device=dev)
        unit.exp_avg_sq = torch.zeros(S, dtype=torch.float32, device=dev)
        unit.wd_table = torch.tensor([[0, S, 1]], dtype=torch.int64, device=dev)
        ref_p = torch.nn.Parameter(unit.master.clone())
        ref_p.grad = want_g.clone()
        opt = torch.optim.AdamW([ref_p], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
        opt.step()
        hp = dict(coef=torch.ones(1, device=dev), lr=1e-2, b1=0.9, b2=0.95, eps=1e-8, wd=0.1, bc1=1.0 - 0.9, bc2=1.0 - 0.95)
        comm.reduce_scatter_adamw(fg, unit, 0.5, hp)
        torch.cuda.synchronize()
        torch.testing.assert_close(unit.master, ref_p.detach(), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(unit.param_shard.float(), ref_p.detach().bfloat16().float(), rtol=1e-2, atol=1e-3)
    finally:
        torch.cuda.synchronize()
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("M,N,K", [(128, 256, 128), (256, 512, 512), (300, 264, 384), (1024, 768, 2048), (4096, 4096, 4096)])
def test_mxfp8_native_kernel_matches_emulation(M, N, K):
    """tcgen05.mma.kind::mxf8f6f4.block_scale kernel (scales staged smem -> TMEM) against the dequantise-and-multiply
    specification of the same quantised operands: only accumulation order and the bf16 output rounding may differ."""
    from vescale_b200.ops import _ext, fp8

    _ext.load(required=True)
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    # wide dynamic range across K blocks so that a wrong scale byte / block mapping cannot hide
    x = torch.randn(M, K, device=dev, generator=g) * torch.exp2(torch.randint(-6, 6, (M, K // 32), device=dev, generator=g).float()).repeat_interleave(32, 1)
    w = torch.randn(N, K, device=dev, generator=g) * torch.exp2(torch.randint(-6, 6, (N, K // 32), device=dev, generator=g).float()).repeat_interleave(32, 1)
    xq, xs = fp8.quantize_mx(x)
    wq, ws = fp8.quantize_mx(w)
    ref = fp8.dequantize_mx(xq, xs) @ fp8.dequantize_mx(wq, ws).t()
    got = fp8.mxfp8_gemm_nt_native(xq, fp8.mx_scale_atoms(xs, 128), wq, fp8.mx_scale_atoms(ws, 256)).float()
    torch.cuda.synchronize()
    err = ((got - ref).norm() / ref.norm()).item()
    assert err < 6e-3, err
    torch.testing.assert_close(got, ref, rtol=2e-2, atol=2e-2 * ref.abs().max().item())
    # second call on the same operands is bit-identical (no stale TMEM / barrier state between launches)
    assert torch.equal(fp8.mxfp8_gemm_nt_native(xq, fp8.mx_scale_atoms(xs, 128), wq, fp8.mx_scale_atoms(ws, 256)).float(), got)


def test_mx_quantize_kernel_matches_specification():
    """Fused bf16 -> MXFP8 quantiser (elements + E8M0 scales in atom order) is bit-identical to quantize_mx + mx_scale_atoms."""
    from vescale_b200.ops import _ext, fp8

    _ext.load(required=True)
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(7)
    for R, K, mult in ((128, 128, 128), (300, 512, 128), (300, 512, 256), (4096, 4096, 256)):
        x = (torch.randn(R, K, device=dev, generator=g) * torch.exp2(torch.randint(-20, 20, (R, K // 32), device=dev, generator=g).float()).repeat_interleave(32, 1)).bfloat16()
        x[0, :32] = 0  # an all-zero block
        x[1, 32:64] = 1e-38  # below the smallest scale
        q_ref, s_ref = fp8.quantize_mx(x)
        q, sf = fp8.quantize_mx_fused(x, mult)
        torch.cuda.synchronize()
        assert torch.equal(sf, fp8.mx_scale_atoms(s_ref, mult)), (R, K, mult)
        assert torch.equal(q.view(torch.uint8), q_ref.view(torch.uint8)), (R, K, mult)
