"""OpenLLaMA-7B-shape DP x TP benchmark (the reference's published benchmark script shape:
seq 2048, total batch 16, bf16, e.g. dp=4 tp=2, warmup 10 / iters 40; prints `1 iter time` and `mfu`;
``legacy/examples/open_llama_4D_benchmark/run_open_llama_w_vescale.py:35-123``).

TP/SP by DModule auto-plan over an nn.Module Llama; DP by DDP + DistributedOptimizer (ZeRO-2+).
    torchrun --nproc-per-node 8 examples/open_llama_4D_benchmark/run.py --dp 4 --tp 2
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from vescale_b200 import Replicate, Shard, init_device_mesh  # noqa: E402
from vescale_b200.optim import DistributedOptimizer  # noqa: E402
from vescale_b200.parallel.ddp import DistributedDataParallel as DDP  # noqa: E402
from vescale_b200.parallel.dmodule import parallelize_module  # noqa: E402


class Attn(nn.Module):
    def __init__(self, h, nh):
        super().__init__()
        self.nh, self.hd = nh, h // nh
        self.q_proj, self.k_proj, self.v_proj, self.o_proj = (nn.Linear(h, h, bias=False) for _ in range(4))

    def forward(self, x):
        B, S, _ = x.shape
        q, k, v = (p(x).view(B, S, -1, self.hd).transpose(1, 2) for p in (self.q_proj, self.k_proj, self.v_proj))
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        return self.o_proj(o.transpose(1, 2).reshape(B, S, -1))


class MLP(nn.Module):
    def __init__(self, h, f):
        super().__init__()
        self.gate_proj, self.up_proj, self.down_proj = nn.Linear(h, f, bias=False), nn.Linear(h, f, bias=False), nn.Linear(f, h, bias=False)

    def forward(self, x):
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


class Block(nn.Module):
    def __init__(self, h, f, nh):
        super().__init__()
        self.input_layernorm, self.post_attention_layernorm = nn.LayerNorm(h), nn.LayerNorm(h)
        self.self_attn, self.mlp = Attn(h, nh), MLP(h, f)

    def forward(self, x):
        x = x + self.self_attn(self.input_layernorm(x))
        return x + self.mlp(self.post_attention_layernorm(x))


class Model(nn.Module):
    def __init__(self, vocab, h, f, nh, layers):
        super().__init__()
        self.embed_tokens = nn.Embedding(vocab, h)
        self.layers = nn.ModuleList([Block(h, f, nh) for _ in range(layers)])
        self.norm = nn.LayerNorm(h)
        self.lm_head = nn.Linear(h, vocab, bias=False)

    def forward(self, ids):
        x = self.embed_tokens(ids)
        for l in self.layers:
            x = l(x)
        return self.lm_head(self.norm(x))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dp", type=int, default=None, help="default: world size / tp")
    ap.add_argument("--tp", type=int, default=None, help="default: 2 when the world size is even, else 1")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--hidden", type=int, default=4096)
    ap.add_argument("--ffn", type=int, default=11008)
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--vocab", type=int, default=32000)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--bsz", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--iters", type=int, default=40)
    args = ap.parse_args()
    cuda = torch.cuda.is_available()
    dist.init_process_group("nccl" if cuda else "gloo")
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = "cuda" if cuda else "cpu"
    ws = dist.get_world_size()
    args.tp = args.tp or (2 if ws % 2 == 0 else 1)
    args.dp = args.dp or ws // args.tp
    mesh = init_device_mesh(dev, (args.dp, args.tp), mesh_dim_names=("DP", "TP"))
    torch.manual_seed(0)
    model = Model(args.vocab, args.hidden, args.ffn, args.heads, args.layers).to(dev).to(torch.bfloat16 if cuda else torch.float32)
    L = r"layers\.\d+\."
    plan = {
        "parameter": {L + r"self_attn\.[qkv]_proj\.weight": [Shard(0)], L + r"self_attn\.o_proj\.weight": [Shard(1)], L + r"mlp\.(gate|up)_proj\.weight": [Shard(0)], L + r"mlp\.down_proj\.weight": [Shard(1)], r"lm_head\.weight": [Shard(0)], r"embed_tokens\.weight": [Shard(0)]},
        "forward": {r"input": [[Replicate()]], r"embed_tokens\.output": [[Shard(1)]], L + r"self_attn\.input": [[Replicate()]], L + r"self_attn\.output": [[Shard(1)]], L + r"mlp\.input": [[Replicate()]], L + r"mlp\.output": [[Shard(1)]], r"lm_head\.input": [[Replicate()]]},
    }
    parallelize_module(model, mesh["TP"], plan)
    ddp = DDP(model, mesh["DP"].get_group(0), use_distributed_optimizer=True, overlap_grad_reduce=True)
    opt = DistributedOptimizer(torch.optim.AdamW(model.parameters(), lr=3e-4), [ddp], clip_grad=1.0, overlap_param_gather=True)
    g = torch.Generator().manual_seed(mesh.get_local_rank("DP"))
    bs = args.bsz // args.dp
    times = []
    for it in range(args.warmup + args.iters):
        ids = torch.randint(0, args.vocab, (bs, args.seq + 1), generator=g).to(dev)
        if cuda:
            torch.cuda.synchronize()
        t0 = time.time()
        opt.zero_grad()
        logits = ddp(ids[:, :-1])
        from vescale_b200.dtensor import loss_parallel

        with loss_parallel():
            loss = F.cross_entropy(logits.view(-1, args.vocab), ids[:, 1:].reshape(-1))
            loss.backward()
        opt.step()
        if cuda:
            torch.cuda.synchronize()
        if it >= args.warmup:
            times.append(time.time() - t0)
    it_t = sum(times) / max(1, len(times))
    n_params = 12 * args.layers * args.hidden**2 * (1 + (args.ffn * 3 / (4 * args.hidden) - 1) / 3) + args.vocab * args.hidden
    flops = 3 * 2 * (args.layers * (4 * args.hidden**2 + 3 * args.hidden * args.ffn + 2 * args.seq * args.hidden / 2) + args.vocab * args.hidden) * args.bsz * args.seq
    if dist.get_rank() == 0:
        peak = 1462e12 if cuda else 1e12
        print(f"1 iter time: {it_t:.4f} s ; mfu: {flops / it_t / (peak * dist.get_world_size()) * 100:.2f}%")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
