"""Mixtral expert-parallel training with MFU print.  torchrun --nproc-per-node 8 examples/mixtral_ep/train.py"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from vescale_b200 import init_device_mesh  # noqa: E402
from vescale_b200.models import MixtralConfig, MixtralModel  # noqa: E402
from vescale_b200.parallel.moe import MoEOptimizer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="tiny")
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--seq-len", type=int, default=64)
    ap.add_argument("--batch", type=int, default=2)
    args = ap.parse_args()
    cuda = torch.cuda.is_available()
    dist.init_process_group("nccl" if cuda else "gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = "cuda" if cuda else "cpu"
    cfg = getattr(MixtralConfig, args.model)()
    if args.layers:
        cfg.num_layers = args.layers
    if not cuda:
        cfg.dtype = torch.float32
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("EP",))
    model = MixtralModel(cfg, ep_group=mesh.get_group(0), device=dev).reset_parameters(0)
    opt = MoEOptimizer(torch.optim.AdamW(model.parameters(), lr=1e-4, fused=cuda), model, ep_group=mesh.get_group(0), clip_grad=1.0)
    g = torch.Generator().manual_seed(rank)
    for step in range(args.steps):
        t0 = time.time()
        tok = torch.randint(0, cfg.vocab_size, (args.batch, args.seq_len + 1), generator=g).to(dev)
        loss = model(tok[:, :-1], tok[:, 1:])
        loss.backward()
        opt.step()
        opt.zero_grad()
        if cuda:
            torch.cuda.synchronize()
        if rank == 0:
            print(f"step {step} loss {loss.item():.4f} {(time.time()-t0)*1e3:.1f} ms tokens/rank routed {model.layers[0].moe.last_tokens_per_rank}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
