"""Symmetric-memory tensor collectives and the FSDP all-gather⊕first-GEMM kernel vs NCCL(+cuBLAS) (torchrun, >= 2 GPUs).
Device-timed with CUDA events, median of 10 after 3 warm-ups, max over ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks.tp_bench import timeit  # noqa: E402


def main():
    dist.init_process_group("nccl")
    rank, W = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device())
    from vescale_b200 import init_device_mesh
    from vescale_b200.comm.symm_collectives import enable_symmetric_collectives
    from vescale_b200.models import LlamaConfig
    from vescale_b200.models.llama import LlamaBlock
    from vescale_b200.parallel.fsdp import fully_shard

    mesh = init_device_mesh("cuda", (W,))
    (sc,) = enable_symmetric_collectives(mesh, reserve_bytes=256 << 20)
    res = []

    # ---- all-reduce (TP without SP / norm-weight grads / tied embeddings): C11, C17, C19
    for n in (4096, 65536, 1 << 20, 8192 * 4096, 1 << 27):
        x = torch.randn(n, device=dev).bfloat16()
        y = x.clone()
        xs = sc.empty(n)  # symmetric-resident operand: zero-copy
        xs.copy_(x)
        for mm in (True, False):
            sc.use_multimem = mm
            t_s = timeit(lambda: sc.all_reduce(x))
            t_z = timeit(lambda: sc.all_reduce(xs))
            res.append({"op": f"all_reduce bf16 {'nvls' if mm else 'p2p'}", "bytes": n * 2, "symm_ms": t_s, "symm_zero_copy_ms": t_z})
        t_n = timeit(lambda: dist.all_reduce(y))
        for r in res[-2:]:
            r["zero_copy_speedup_vs_nccl"] = t_n / r["symm_zero_copy_ms"]
            r["nccl_ms"] = t_n
            r["speedup_vs_nccl"] = t_n / r["symm_ms"]
    sc.use_multimem = True

    # ---- Shard(seq) -> Shard(heads) (Ulysses-style swap): C12
    B, S, Hh, D = 1, 8192, 32, 128
    x = torch.randn(B, S // W, Hh, D, device=dev).bfloat16()

    def nccl_a2a():
        pieces = torch.stack([p.contiguous() for p in x.chunk(W, dim=2)], 0)
        out = torch.empty_like(pieces)
        dist.all_to_all_single(out, pieces)
        return torch.cat(list(out.unbind(0)), dim=1)

    t_s = timeit(lambda: sc.all_to_all_permute(x, 1, 2))
    t_n = timeit(nccl_a2a)
    res.append({"op": "a2a Shard(1)->Shard(2) [1,8192,32,128] bf16", "bytes": x.numel() * 2, "symm_ms": t_s, "nccl_ms": t_n, "speedup_vs_nccl": t_n / t_s})

    # ---- vocab-parallel cross entropy, Llama-3 vocabulary, 8192 tokens: C20
    T, V = 8192, 128256 // W // 8 * 8
    logits = torch.randn(T, V, device=dev).bfloat16()
    target = torch.randint(0, V * W, (T,), device=dev)
    nv = torch.tensor([float(T)], device=dev)

    def ref_vp():
        lmax = logits.amax(-1).float()
        dist.all_reduce(lmax, op=dist.ReduceOp.MAX)
        ex = (logits.float() - lmax[:, None]).exp()
        se = ex.sum(-1)
        dist.all_reduce(se)
        idx = (target - rank * V).clamp(0, V - 1)
        picked = logits.gather(1, idx[:, None]).squeeze(1).float()
        dist.all_reduce(picked)
        return ex.div_(se[:, None]).bfloat16()

    t_s = timeit(lambda: sc.vocab_ce_fwd_bwd_(logits, target, nv, rank * V))
    t_n = timeit(ref_vp)
    res.append({"op": f"vocab-parallel CE fwd+bwd [8192, {V}] per rank", "bytes": T * V * 2, "symm_ms": t_s, "nccl_ms": t_n, "speedup_vs_nccl": t_n / t_s,
                "frac_of_copy_bw": (3 * T * V * 2 / 6478e9 * 1e3) / t_s})

    # ---- FSDP unit all-gather ⊕ first GEMM (Llama-3-8B block, 8192 tokens): C1/C8
    cfg = LlamaConfig.llama3_8b()
    with torch.device("meta"):
        blk = LlamaBlock(cfg, 0)
    fully_shard(blk, mesh, comm_backend="symm", block_rows=32, init_fn=lambda m: m.reset_parameters(torch.Generator(device=dev).manual_seed(1)) if hasattr(m, "reset_parameters") else None)
    u = blk._fsdp_unit
    comm = u.comm
    slot = comm.fusable_slot(u, "wqkv")
    full = u._alloc_full(torch.bfloat16)
    xin = torch.randn(8192, cfg.hidden_size, device=dev).bfloat16()
    if slot is not None:
        w_view = full[slot.offset : slot.end].view(slot.shape)

        side = torch.cuda.Stream()

        def fused():  # as FSDP runs it: the rest of the unit streams in on the all-gather stream behind the first GEMM
            cur = torch.cuda.current_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                comm.all_gather(u.param_shard, full, u, skip=(slot.offset, slot.end))
            y = comm.fused_first_linear(xin, u, slot, full)
            cur.wait_stream(side)
            return y

        def serial():
            comm.all_gather(u.param_shard, full, u)
            return xin @ w_view.t()

        def serial_nccl():
            dist.all_gather_into_tensor(full, u.param_shard)
            return xin @ w_view.t()

        y1 = fused()
        y2 = serial()
        torch.cuda.synchronize()
        err = (y1.float() - y2.float()).abs().max().item()
        t_f, t_s2, t_n = timeit(fused), timeit(serial), timeit(serial_nccl)
        t_mm = timeit(lambda: xin @ w_view.t())
        t_first = timeit(lambda: comm.fused_first_linear(xin, u, slot, full))
        res.append({"op": "FSDP unit all-gather + qkv GEMM (Llama-3-8B block, 436 MB unit)", "fused_total_ms": t_f, "symm_ag_then_gemm_ms": t_s2, "nccl_ag_then_gemm_ms": t_n,
                    "gemm_only_ms": t_mm, "time_to_first_gemm_output_ms": t_first, "time_to_first_gemm_output_serial_ms": t_s2, "max_abs_err": err})
    if rank == 0:
        print(json.dumps({"world": W, "results": res}, indent=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
