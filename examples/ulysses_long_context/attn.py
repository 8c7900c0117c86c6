"""Ulysses context parallelism: the sequence is sharded outside attention, the heads inside it; the two swaps are single
all-to-alls (on B200 with ``enable_symmetric_collectives`` a put kernel with the permutes folded in).  Checks the result
against attention over the gathered sequence.

    torchrun --nproc-per-node 4 examples/ulysses_long_context/attn.py [--device cpu]
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--seq-len", type=int, default=None)
    args = ap.parse_args()
    cuda = args.device == "cuda"
    dist.init_process_group("nccl" if cuda else "gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.context import ulysses_attention

    mesh = init_device_mesh(args.device, (world,), mesh_dim_names=("cp",))
    if cuda:
        from vescale_b200.comm.symm_collectives import enable_symmetric_collectives

        enable_symmetric_collectives(mesh)
    B, S, Hq, Hk, D = 1, args.seq_len or (32768 if cuda else 64 * world), 4 * world, world, 128 if cuda else 16
    dt = torch.bfloat16 if cuda else torch.float32
    g = torch.Generator(device=dev).manual_seed(7)
    q, k, v = (torch.randn(B, S, h, D, device=dev, generator=g, dtype=dt) for h in (Hq, Hk, Hk))
    sl = slice(rank * S // world, (rank + 1) * S // world)
    out = ulysses_attention(q[:, sl].contiguous(), k[:, sl].contiguous(), v[:, sl].contiguous(), mesh, "cp", causal=True)
    ref = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True, enable_gqa=True).transpose(1, 2)[:, sl]
    err = (out.float() - ref.float()).abs().max().item()
    if rank == 0:
        print(f"ulysses attention over {world} ranks, seq {S}: max abs err vs full attention {err:.2e}")
    assert err < (2e-2 if cuda else 1e-4)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
