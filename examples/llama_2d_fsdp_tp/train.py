"""2-D Llama training: FSDP over the data-parallel mesh dim x tensor/sequence parallelism over the TP dim, with every TP
collective fused into its GEMM on B200 (``FusedTP``: all-gather ⊕ GEMM, GEMM ⊕ reduce-scatter, forward and backward), or the
same model on ordinary collectives (``--tp-impl plain``; also what runs on CPU/gloo).

    torchrun --nproc-per-node 8 examples/llama_2d_fsdp_tp/train.py --tp 2 --model llama3_8b --seq-len 8192
    torchrun --nproc-per-node 4 examples/llama_2d_fsdp_tp/train.py --tp 2 --model tiny --device cpu --steps 3

Reference counterpart: the 4-D Llama recipe ``legacy/examples/llama2_4D_finetune`` (TP+SP sharding plan + DDP/ZeRO); here
the data-parallel dimension is RaggedShard FSDP and TP/SP is the model's own fused-kernel path.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tp", type=int, default=2)
    ap.add_argument("--tp-impl", default="auto", choices=["auto", "fused", "plain"])
    ap.add_argument("--model", default="tiny")
    ap.add_argument("--seq-len", type=int, default=None)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    args = ap.parse_args()

    cuda = args.device == "cuda"
    dist.init_process_group("nccl" if cuda else "gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")

    from vescale_b200 import init_device_mesh
    from vescale_b200.comm.fused_tp import FusedTP, PlainTP
    from vescale_b200.models import LlamaConfig, LlamaTPModel, llama_flops_per_token
    from vescale_b200.optim import FSDPAdamW
    from vescale_b200.parallel.fsdp import MixedPrecisionPolicy, fully_shard

    cfg = getattr(LlamaConfig, args.model)()
    S = args.seq_len or min(cfg.max_seq_len, 64 if not cuda else cfg.max_seq_len)
    cfg.max_seq_len = max(cfg.max_seq_len, S)
    mesh = init_device_mesh(args.device, (world // args.tp, args.tp), mesh_dim_names=("dp", "tp"))
    impl = args.tp_impl if args.tp_impl != "auto" else ("fused" if cuda else "plain")
    tp = FusedTP(mesh, "tp", dev) if impl == "fused" else PlainTP(mesh, "tp")
    model = LlamaTPModel(cfg, tp, device=dev).reset_parameters(seed=0)
    mp = MixedPrecisionPolicy(param_dtype=cfg.dtype)
    for blk in model.layers:
        fully_shard(blk, mesh, mesh_dim="dp", mp_policy=mp)
    fully_shard(model.embed, mesh, mesh_dim="dp", mp_policy=mp)
    fully_shard(model.head, mesh, mesh_dim="dp", mp_policy=mp)
    fully_shard(model, mesh, mesh_dim="dp", mp_policy=mp)
    opt = FSDPAdamW(model, lr=3e-4, max_grad_norm=1.0, tp_group=mesh.get_group("tp"))
    dp_rank, dp = mesh.get_local_rank("dp"), world // args.tp
    g = torch.Generator().manual_seed(1000 + dp_rank)  # one batch per TP group
    for step in range(args.steps):
        tok = torch.randint(0, cfg.vocab_size, (args.batch, S + 1), generator=g).to(dev)
        t0 = time.perf_counter()
        share = model(tok[:, :-1], tok[:, 1:])
        share.backward()
        gnorm = opt.step()
        opt.zero_grad()
        loss = model.loss_for_logging(share)
        if dp > 1:
            dist.all_reduce(loss, group=mesh.get_group("dp"))
            loss /= dp
        if cuda:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            toks = dp * args.batch * S
            print(f"step {step} loss {loss.item():.4f} grad-norm {gnorm.item():.3f} {dt * 1e3:.1f} ms  {toks / dt:.0f} tok/s  "
                  f"{toks / dt * llama_flops_per_token(cfg, S) / world / 1e12:.1f} model-TFLOPS/GPU  [fsdp{dp} x tp{args.tp} ({impl})]", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
