"""Mixtral DP x TP(+SP) benchmark by DModule sharding plan (the reference's script shape: 16 layers, hidden 4096, ffn 14336,
32 heads / 8 KV heads, 8 experts top-2, bsz 16, seq 256, tp=8 dp=1, warmup 2 / iters 10, prints ``1 iter time`` and ``mfu``;
``legacy/examples/mixtral_4D_benchmark/mixtral_train.py:60-149``, plan ``sharding_plan.py:22-69``).

The model is a plain ``nn.Module`` Mixtral with token-choice sparse routing (``one_hot`` / ``where`` / ``index_add_``, the
HuggingFace formulation); nothing in it knows about parallelism.  The plan shards attention and every expert's w1 / w3 by
column and o_proj / w2 by row, keeps the router replicated (so every TP rank routes identically), runs the norms
sequence-parallel, and DDP + DistributedOptimizer (ZeRO-2+) handle the DP dim.  For expert parallelism on the fused dispatch
kernels see ``examples/mixtral_ep``.
    torchrun --nproc-per-node 8 examples/mixtral_4D_benchmark/run.py --tp 8 --dp 1
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


class RMSNorm(nn.Module):
    def __init__(self, h, eps=1e-5):
        super().__init__()
        self.weight, self.eps = nn.Parameter(torch.ones(h)), eps

    def forward(self, x):
        return self.weight * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.eps))


class Attn(nn.Module):
    def __init__(self, h, nh, nkv):
        super().__init__()
        self.hd, self.rep = h // nh, nh // nkv
        self.q_proj, self.o_proj = nn.Linear(h, h, bias=False), nn.Linear(h, h, bias=False)
        self.k_proj, self.v_proj = nn.Linear(h, nkv * self.hd, bias=False), nn.Linear(h, nkv * self.hd, bias=False)

    def forward(self, x):
        B, S, _ = x.shape
        q, k, v = (p(x).view(B, S, -1, self.hd).transpose(1, 2) for p in (self.q_proj, self.k_proj, self.v_proj))
        k, v = (t.repeat_interleave(self.rep, dim=1) for t in (k, v))
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        return self.o_proj(o.transpose(1, 2).reshape(B, S, -1))


class Expert(nn.Module):
    def __init__(self, h, f):
        super().__init__()
        self.w1, self.w3, self.w2 = nn.Linear(h, f, bias=False), nn.Linear(h, f, bias=False), nn.Linear(f, h, bias=False)

    def forward(self, x):
        return self.w2(F.silu(self.w1(x)) * self.w3(x))


class SparseMoE(nn.Module):
    def __init__(self, h, f, n_experts, top_k):
        super().__init__()
        self.gate = nn.Linear(h, n_experts, bias=False)
        self.experts = nn.ModuleList([Expert(h, f) for _ in range(n_experts)])
        self.top_k = top_k

    def forward(self, x):
        B, S, H = x.shape
        flat = x.reshape(-1, H)
        probs = F.softmax(self.gate(flat).float(), dim=-1)
        w, sel = torch.topk(probs, self.top_k, dim=-1)
        w = (w / w.sum(-1, keepdim=True)).to(x.dtype)
        out = torch.zeros_like(flat)
        mask = F.one_hot(sel, len(self.experts)).permute(2, 1, 0)  # (expert, k, token)
        for e, expert in enumerate(self.experts):
            k_idx, tok = torch.where(mask[e])
            if tok.numel() == 0:
                continue
            out.index_add_(0, tok, expert(flat[tok]) * w[tok, k_idx, None])
        return out.reshape(B, S, H)


class Block(nn.Module):
    def __init__(self, h, f, nh, nkv, n_experts, top_k):
        super().__init__()
        self.input_layernorm, self.post_attention_layernorm = RMSNorm(h), RMSNorm(h)
        self.self_attn, self.block_sparse_moe = Attn(h, nh, nkv), SparseMoE(h, f, n_experts, top_k)

    def forward(self, x):
        x = x + self.self_attn(self.input_layernorm(x))
        return x + self.block_sparse_moe(self.post_attention_layernorm(x))


class Mixtral(nn.Module):
    def __init__(self, a):
        super().__init__()
        self.embed_tokens = nn.Embedding(a.vocab_size, a.hidden_size)
        self.layers = nn.ModuleList([Block(a.hidden_size, a.intermediate_size, a.num_attention_heads, a.num_key_value_heads, a.num_experts, a.top_k) for _ in range(a.num_hidden_layers)])
        self.norm = RMSNorm(a.hidden_size)

    def forward(self, ids):
        x = self.embed_tokens(ids)
        for l in self.layers:
            x = l(x)
        return self.norm(x)


L = r"layers\.\d+\."
E = L + r"block_sparse_moe\.experts\.\d+\."
mixtral_plan = {
    "parameter": {
        L + r"self_attn\.[qkv]_proj\.weight": [Shard(0)], L + r"self_attn\.o_proj\.weight": [Shard(1)],
        E + r"w[13]\.weight": [Shard(0)], E + r"w2\.weight": [Shard(1)],
    },  # everything else (embedding, norms, router) stays replicated
    "forward": {
        r"input": [[Replicate()]],
        L + r"input_layernorm\.input": [[Shard(1)]], L + r"input_layernorm\.output": [[Shard(1)]],  # sequence parallel norms
        L + r"self_attn\.input": [[Replicate()]], L + r"self_attn\.output": [[Replicate()]],
        L + r"post_attention_layernorm\.input": [[Shard(1)]], L + r"post_attention_layernorm\.output": [[Shard(1)]],
        L + r"block_sparse_moe\.input": [[Replicate()]], L + r"block_sparse_moe\.gate\.output": [[Replicate()]],
        E + r"w[13]\.input": [[Replicate()]], E + r"w2\.output": [[Replicate()]],
        L + r"block_sparse_moe\.output": [[Replicate()]],
        r"norm\.input": [[Replicate()]],
    },
}


def fwd_flops(a, bsz, seq):
    """Forward FLOPs of one batch: attention projections + scores + top-k experts + router (2 x MACs)."""
    h, f, hd = a.hidden_size, a.intermediate_size, a.hidden_size // a.num_attention_heads
    per_tok = 2 * h * h * 2 + 2 * h * a.num_key_value_heads * hd * 2 + 2 * seq * h + a.top_k * 3 * 2 * h * f + 2 * h * a.num_experts
    return a.num_hidden_layers * per_tok * bsz * seq


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--iter", type=int, default=10)
    ap.add_argument("--vocab_size", type=int, default=32000)
    ap.add_argument("--hidden_size", type=int, default=4096)
    ap.add_argument("--intermediate_size", type=int, default=14336)
    ap.add_argument("--num_hidden_layers", type=int, default=16)
    ap.add_argument("--num_attention_heads", type=int, default=32)
    ap.add_argument("--num_key_value_heads", type=int, default=8)
    ap.add_argument("--num_experts", type=int, default=8)
    ap.add_argument("--top_k", type=int, default=2)
    ap.add_argument("--bsz", type=int, default=16)
    ap.add_argument("--seqlen", type=int, default=256)
    ap.add_argument("--dp", type=int, default=None)
    ap.add_argument("--tp", type=int, default=None)
    ap.add_argument("--dtype", default=None, help="bf16 (default on GPU) or fp32 (the reference's benchmark dtype)")
    a = ap.parse_args()
    cuda = torch.cuda.is_available()
    dist.init_process_group("nccl" if cuda else "gloo")
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev, ws = ("cuda" if cuda else "cpu"), dist.get_world_size()
    a.tp = a.tp or (ws if a.dp is None else ws // a.dp)
    a.dp = a.dp or ws // a.tp
    dtype = {"bf16": torch.bfloat16, "fp32": torch.float32, None: torch.bfloat16 if cuda else torch.float32}[a.dtype]
    mesh = init_device_mesh(dev, (a.dp, a.tp), mesh_dim_names=("DP", "TP"))
    torch.manual_seed(0)
    model = Mixtral(a).to(dev).to(dtype)
    parallelize_module(model, mesh["TP"], mixtral_plan)
    ddp = DDP(model, mesh["DP"].get_group(0), use_distributed_optimizer=True, overlap_grad_reduce=True)
    opt = DistributedOptimizer(torch.optim.AdamW(model.parameters(), lr=3e-4), [ddp], clip_grad=1.0, overlap_param_gather=True)
    g = torch.Generator().manual_seed(mesh.get_local_rank("DP"))
    bs = max(1, a.bsz // a.dp)
    times, loss = [], None
    for it in range(a.warmup + a.iter):
        ids = torch.randint(0, a.vocab_size, (bs, a.seqlen), generator=g).to(dev)
        if cuda:
            torch.cuda.synchronize()
        t0 = time.time()
        opt.zero_grad()
        loss = ddp(ids).to_local().float().pow(2).mean()
        loss.backward()
        model.finish_grad_sync()
        opt.step()
        if cuda:
            torch.cuda.synchronize()
        if it >= a.warmup:
            times.append(time.time() - t0)
    it_t = sum(times) / max(1, len(times))
    if dist.get_rank() == 0:
        peak = (2250e12 if dtype is torch.bfloat16 else 70e12) if cuda else 1e12  # B200 dense bf16 / non-tensor-core fp32
        print(f"1 iter time: {it_t:.4f} s ; loss {loss.item():.5f} ; mfu: {3 * fwd_flops(a, bs * a.dp, a.seqlen) / it_t / (peak * ws) * 100:.2f}%")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
