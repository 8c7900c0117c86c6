"""Mixtral-8x7B expert-parallel training benchmark (BASELINE.json config 4): EP over all ranks, device-side symmetric-memory
dispatch / grouped tcgen05 GEMM / combine (``--dispatch symm``) vs the NCCL all_to_all_single path with its host sync
(``--dispatch nccl``, the legacy ``moe/_scheduler.py:165-215`` pattern).  Same JSON contract as ``bench.py``:

    torchrun --nproc-per-node 8 benchmarks/mixtral_bench.py --steps 5 --warmup 3 [--layers 8]

``--layers`` below 32 is for bring-up only and marks the record invalid for the named configuration.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="mixtral_8x7b", choices=["mixtral_8x7b", "tiny"])
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--seq-len", type=int, default=4096)
    ap.add_argument("--micro-batch", type=int, default=1)
    ap.add_argument("--dispatch", default="auto", choices=["auto", "symm", "nccl"])
    args = ap.parse_args()

    cuda = torch.cuda.is_available()
    dist.init_process_group("nccl" if cuda else "gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")

    from vescale_b200 import init_device_mesh
    from vescale_b200.models import MixtralConfig, MixtralModel
    from vescale_b200.parallel.moe import MoEOptimizer
    from vescale_b200.utils import mixtral_flops_per_token

    cfg = getattr(MixtralConfig, args.model)()
    invalid = None
    if args.layers is not None:
        cfg.num_layers = args.layers
        invalid = f"layers overridden to {args.layers}"
    if not cuda:
        cfg.dtype = torch.float32
    S, B = (args.seq_len, args.micro_batch) if args.model != "tiny" else (64, 2)
    cfg.max_seq_len = max(cfg.max_seq_len, S)
    mesh = init_device_mesh(dev.type, (world,), mesh_dim_names=("EP",))
    model = MixtralModel(cfg, ep_group=mesh.get_group(0), device=dev).reset_parameters(0)
    use_symm = cuda and world > 1 and args.dispatch in ("auto", "symm")
    if use_symm:
        from vescale_b200.parallel.moe.symm_dispatch import SymmMoEDispatcher

        disp = SymmMoEDispatcher(mesh, cfg.num_experts, cfg.hidden_size, cfg.intermediate_size, max_tokens=B * S, top_k=cfg.top_k, capacity_factor=2.0, device=dev)
        for blk in model.layers:
            blk.moe.use_symmetric_dispatch(disp)  # one set of symmetric buffers shared by all layers
    opt = MoEOptimizer(torch.optim.AdamW(model.parameters(), lr=1e-4, fused=cuda), model, ep_group=mesh.get_group(0), clip_grad=1.0)
    g = torch.Generator().manual_seed(1000 + rank)
    batches = [torch.randint(0, cfg.vocab_size, (B, S + 1), generator=g) for _ in range(4)]
    batches = [b.pin_memory() if cuda else b for b in batches]

    def step(i):
        tok = batches[i % 4].to(dev, non_blocking=True)
        loss = model(tok[:, :-1], tok[:, 1:])
        loss.backward()
        opt.step()
        opt.zero_grad()
        return loss

    def barrier():
        dist.barrier()
        if cuda:
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    if cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    for i in range(args.steps):
        last = step(i)
    if cuda:
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1) if cuda else (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    tokens = world * B * S * args.steps
    tps = tokens / (ms / 1e3)
    out = {
        "metric": "tokens/sec Mixtral-8x7B expert-parallel (bf16)", "value": tps, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "dtype": "bf16" if cuda else "fp32", "data": "synthetic tokens, random-init weights",
        "config": {"model": args.model, "layers": cfg.num_layers, "experts": cfg.num_experts, "top_k": cfg.top_k, "seq_len": S, "global_batch": world * B,
                   "parallelism": f"ep{world}", "dispatch": "symm (device-side counts / put / grouped tcgen05 GEMM / get)" if use_symm else "nccl all_to_all_single + host sync"},
        "model_tflops_per_gpu": tps / world * mixtral_flops_per_token(cfg, S) / 1e12, "final_loss": float(last.item()),
    }
    if invalid:
        out["invalid"] = invalid
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
