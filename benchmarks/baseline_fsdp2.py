"""Baseline B0 of BASELINE.md §3: the *upstream* stack the reference builds on — stock ``torch.distributed.fsdp.fully_shard``
(FSDP2) over NCCL, cuBLAS GEMMs, SDPA attention, ``torch.optim.AdamW(fused=True)`` — on the same Llama-3-8B configuration and
with the same JSON line as ``bench.py``, so the two can be compared on the same box:

    torchrun --nproc-per-node 8 benchmarks/baseline_fsdp2.py --gpus 8 --steps 5 --warmup 3 [--ac full]

Nothing from ``vescale_b200`` is on this path except the config dataclass and the FLOPs formula.  (The reference itself
cannot run: it ships no FSDP wrapper and does not import under torch 2.11 — ``bench.py --impl reference`` reports that.)
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class RMSNorm(nn.Module):
    def __init__(self, h, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(h))
        self.eps = eps

    def forward(self, x):
        return F.rms_norm(x, (x.shape[-1],), self.weight, self.eps)


def rope(x, cos, sin):  # x [B, S, H, D], rotate-half convention
    d = x.shape[-1] // 2
    x1, x2 = x[..., :d], x[..., d:]
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], -1).to(x.dtype)


class Block(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        h, f, d = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
        self.cfg = cfg
        self.attn_norm, self.mlp_norm = RMSNorm(h, cfg.rms_eps), RMSNorm(h, cfg.rms_eps)
        self.wq = nn.Linear(h, cfg.num_heads * d, bias=False)
        self.wk = nn.Linear(h, cfg.num_kv_heads * d, bias=False)
        self.wv = nn.Linear(h, cfg.num_kv_heads * d, bias=False)
        self.wo = nn.Linear(cfg.num_heads * d, h, bias=False)
        self.w_gate, self.w_up, self.w_down = nn.Linear(h, f, bias=False), nn.Linear(h, f, bias=False), nn.Linear(f, h, bias=False)

    def forward(self, x, cos, sin):
        cfg = self.cfg
        B, S, _ = x.shape
        y = self.attn_norm(x)
        q = rope(self.wq(y).view(B, S, cfg.num_heads, cfg.head_dim), cos, sin)
        k = rope(self.wk(y).view(B, S, cfg.num_kv_heads, cfg.head_dim), cos, sin)
        v = self.wv(y).view(B, S, cfg.num_kv_heads, cfg.head_dim)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True, enable_gqa=True)
        x = x + self.wo(o.transpose(1, 2).reshape(B, S, -1))
        y = self.mlp_norm(x)
        return x + self.w_down(F.silu(self.w_gate(y)) * self.w_up(y))


class Llama(nn.Module):
    def __init__(self, cfg, ac: str = "none"):
        super().__init__()
        self.cfg, self.ac = cfg, ac
        self.embed = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.layers = nn.ModuleList([Block(cfg) for _ in range(cfg.num_layers)])
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_eps)
        self.head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)

    def forward(self, tokens, labels):
        S = tokens.shape[1]
        d = self.cfg.head_dim
        inv = 1.0 / (self.cfg.rope_theta ** (torch.arange(0, d, 2, device=tokens.device, dtype=torch.float32) / d))
        ang = torch.arange(S, device=tokens.device, dtype=torch.float32)[:, None] * inv[None, :]
        cos, sin = ang.cos(), ang.sin()
        x = self.embed(tokens)
        for blk in self.layers:
            if self.ac == "full" and self.training:
                from torch.utils.checkpoint import checkpoint

                x = checkpoint(blk, x, cos, sin, use_reentrant=False)
            else:
                x = blk(x, cos, sin)
        logits = self.head(self.norm(x))
        return F.cross_entropy(logits.float().view(-1, logits.shape[-1]), labels.reshape(-1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--model", default="llama3_8b")
    ap.add_argument("--seq-len", type=int, default=8192)
    ap.add_argument("--micro-batch", type=int, default=1)
    ap.add_argument("--ac", default="none", choices=["none", "full"], help="activation checkpointing per block")
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    args = ap.parse_args()

    from torch.distributed.fsdp import MixedPrecisionPolicy, fully_shard

    from vescale_b200.models.llama import LlamaConfig, llama_flops_per_token

    cuda = args.device == "cuda"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or "RANK" in os.environ:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if cuda else "gloo")
    rank = dist.get_rank() if dist.is_initialized() else 0
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    cfg = getattr(LlamaConfig, args.model)()
    S, B = (args.seq_len, args.micro_batch) if args.model != "tiny" else (64, 2)
    torch.manual_seed(1234)
    with torch.device("meta"):
        model = Llama(cfg, args.ac)
    if dist.is_initialized():
        mp = MixedPrecisionPolicy(param_dtype=torch.bfloat16 if cuda else None, reduce_dtype=torch.float32)
        for blk in model.layers:
            fully_shard(blk, mp_policy=mp)
        fully_shard(model, mp_policy=mp)
    model.to_empty(device=dev)
    with torch.no_grad():
        for n, p in model.named_parameters():
            (p.fill_(1.0) if p.ndim == 1 else p.normal_(0, cfg.init_std))
    if not dist.is_initialized() and cuda:
        model = model.bfloat16()  # single process without FSDP: plain bf16 weights (no fp32 master) — noted in the record
    opt = torch.optim.AdamW(model.parameters(), lr=3e-4, betas=(0.9, 0.95), weight_decay=0.1, fused=cuda)
    g = torch.Generator().manual_seed(1000 + rank)
    host = [torch.randint(0, cfg.vocab_size, (B, S + 1), generator=g) for _ in range(4)]
    host = [t.pin_memory() if cuda else t for t in host]

    def step(i):
        t = host[i % 4].to(dev, non_blocking=True)
        loss = model(t[:, :-1], t[:, 1:])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if dist.is_initialized():
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
    if dist.is_initialized():
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    tps = world * B * S * args.steps / (ms / 1e3)
    out = {
        "metric": "tokens/sec Llama-3-8B FSDP (bf16) — torch FSDP2 + NCCL + cuBLAS baseline (B0)", "value": tps, "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "dtype": "bf16" if cuda else "fp32",
        "data": "synthetic tokens, random-init weights", "impl": "torch_fsdp2_baseline",
        "config": {"model": args.model, "layers": cfg.num_layers, "global_batch": world * B, "seq_len": S, "parallelism": f"fsdp{world}",
                   "activation_checkpointing": args.ac, "optimizer": "torch AdamW fused, clip_grad_norm_ 1.0",
                   "master_weights": "fp32 (FSDP2 mixed precision)" if dist.is_initialized() else "none (single process, bf16 weights)"},
        "model_tflops_per_gpu": tps / world * llama_flops_per_token(cfg, S) / 1e12,
        "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30 if cuda else None, "final_loss": float(last.item()),
    }
    if rank == 0:
        print(json.dumps(out))
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
