"""Llama-3 FSDP (RaggedShard) training.  torchrun --nproc-per-node 8 examples/llama_fsdp/train.py --model llama3_8b"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from vescale_b200 import init_device_mesh  # noqa: E402
from vescale_b200.models import LlamaConfig, LlamaModel, llama_flops_per_token  # noqa: E402
from vescale_b200.optim import FSDPAdamW  # noqa: E402
from vescale_b200.parallel.fsdp import fully_shard  # noqa: E402
import vescale_b200.checkpoint as ckpt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="tiny")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--seq-len", type=int, default=None)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--save", default=None)
    args = ap.parse_args()
    cuda = torch.cuda.is_available()
    dist.init_process_group("nccl" if cuda else "gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = "cuda" if cuda else "cpu"
    cfg = getattr(LlamaConfig, args.model)()
    if not cuda:
        cfg.dtype = torch.float32
    S = args.seq_len or min(cfg.max_seq_len, 8192)
    mesh = init_device_mesh(dev, (world,))
    with torch.device("meta"):
        model = LlamaModel(cfg)
    g = torch.Generator(device=dev).manual_seed(0)

    def init_fn(m):
        if hasattr(m, "reset_parameters"):
            m.reset_parameters(g)
        else:
            for p in m.parameters(recurse=False):
                with torch.no_grad():
                    p.normal_(0, cfg.init_std, generator=g) if p.ndim > 1 else p.fill_(1.0)

    from vescale_b200.parallel.fsdp import MixedPrecisionPolicy

    mp = MixedPrecisionPolicy(param_dtype=cfg.dtype)
    for m in [model.embed, *model.layers, model.head]:
        fully_shard(m, mesh, init_fn=init_fn, mp_policy=mp)
    fully_shard(model, mesh, mp_policy=mp)
    opt = FSDPAdamW(model, lr=3e-4, max_grad_norm=1.0)
    tok_gen = torch.Generator().manual_seed(rank)
    for step in range(args.steps):
        t0 = time.time()
        tok = torch.randint(0, cfg.vocab_size, (args.batch, S + 1), generator=tok_gen).to(dev)
        loss = model(tok[:, :-1], tok[:, 1:])
        loss.backward()
        gn = opt.step()
        opt.zero_grad()
        if cuda:
            torch.cuda.synchronize()
        dt = time.time() - t0
        if rank == 0:
            tps = world * args.batch * S / dt
            print(f"step {step} loss {loss.item():.4f} grad_norm {float(gn):.3f} {dt*1e3:.1f} ms {tps:,.0f} tok/s  model TFLOPS/GPU {tps/world*llama_flops_per_token(cfg, S)/1e12:.1f}")
    if args.save:
        ckpt.save(args.save, {"model": model, "optimizer": opt})
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
