"""nanoGPT-style model with automatic TP/SP plan (MEGATRON policy) + DDP + BasicOptimizer + checkpoint save/load.
    torchrun --nproc-per-node 4 examples/nanogpt_4D_finetune/finetune.py --dp 2 --tp 2
(reference: ``legacy/examples/nanogpt_4D_finetune/finetune_4D.py`` — sharding plan, DDP, DistributedOptimizer, checkpoint.)"""
import argparse
import os
import sys
import tempfile

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import vescale_b200.checkpoint as ckpt  # noqa: E402
from vescale_b200 import Replicate, Shard  # noqa: E402
from vescale_b200.devicemesh_api import VESCALE_DEVICE_MESH  # noqa: E402
from vescale_b200.optim import BasicOptimizer  # noqa: E402
from vescale_b200.parallel.ddp import DistributedDataParallel as DDP  # noqa: E402
from vescale_b200.parallel.dmp import auto_parallelize_module  # noqa: E402


class CausalSelfAttention(nn.Module):
    def __init__(self, h, nh):
        super().__init__()
        self.nh = nh
        self.q_proj, self.k_proj, self.v_proj = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)
        self.c_proj = nn.Linear(h, h)

    def forward(self, x):
        B, T, C = x.shape
        q, k, v = (p(x).view(B, T, -1, C // self.nh).transpose(1, 2) for p in (self.q_proj, self.k_proj, self.v_proj))
        y = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        return self.c_proj(y.transpose(1, 2).reshape(B, T, -1))


class Block(nn.Module):
    def __init__(self, h, nh):
        super().__init__()
        self.ln_1, self.ln_2 = nn.LayerNorm(h), nn.LayerNorm(h)
        self.attn = CausalSelfAttention(h, nh)
        self.c_fc, self.c_proj2 = nn.Linear(h, 4 * h), nn.Linear(4 * h, h)

    def forward(self, x):
        x = x + self.attn(self.ln_1(x))
        return x + self.c_proj2(F.gelu(self.c_fc(self.ln_2(x))))


class GPT(nn.Module):
    def __init__(self, vocab=512, h=64, nh=4, layers=2, block=64):
        super().__init__()
        self.wte, self.wpe = nn.Embedding(vocab, h), nn.Embedding(block, h)
        self.h = nn.ModuleList([Block(h, nh) for _ in range(layers)])
        self.ln_f = nn.LayerNorm(h)
        self.lm_head = nn.Linear(h, vocab, bias=False)

    def forward(self, idx):
        pos = torch.arange(idx.shape[1], device=idx.device)
        x = self.wte(idx) + self.wpe(pos)
        for b in self.h:
            x = b(x)
        return self.lm_head(self.ln_f(x))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dp", type=int, default=2)
    ap.add_argument("--tp", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    cuda = torch.cuda.is_available()
    dist.init_process_group("nccl" if cuda else "gloo")
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = "cuda" if cuda else "cpu"
    VESCALE_DEVICE_MESH.init_device_mesh(dev, (args.dp, args.tp), mesh_dim_names=("DP", "TP"))
    torch.manual_seed(0)
    model = GPT().to(dev)
    over = {"parameter": {r"h\.\d+\.c_proj2\.weight": [Shard(1)], r"wpe\.weight": [Replicate()]},
            "forward": {r"input": [[Replicate()]], r"h\.\d+\.attn\.input": [[Replicate()]], r"h\.\d+\.c_fc\.input": [[Replicate()]], r"ln_f\.input": [[Replicate()]], r"wte\.output": [[Replicate()]], r"h\.\d+\.ln_\d\.input": [[Replicate()]]}}
    def build():
        torch.manual_seed(0)
        m = GPT().to(dev)
        auto_parallelize_module(m, VESCALE_DEVICE_MESH["TP"], "MEGATRON", plan_override=over, factory=True)
        d = DDP(m, VESCALE_DEVICE_MESH.get_data_parallel_group(), overlap_grad_reduce=True)
        o = BasicOptimizer(torch.optim.AdamW(m.parameters(), lr=1e-3), [d], clip_grad=1.0)
        return m, d, o

    def train_step(d, o, step):
        g = torch.Generator().manual_seed(1000 * step + VESCALE_DEVICE_MESH.get_data_parallel_rank())
        ids = torch.randint(0, 512, (4, 33), generator=g).to(dev)
        o.zero_grad()
        logits = d(ids[:, :-1])
        loss = F.cross_entropy(logits.full_tensor().view(-1, 512), ids[:, 1:].reshape(-1))
        loss.backward()
        o.step()
        return loss.item()

    del model
    model, ddp, opt = build()
    for step in range(args.steps):
        loss = train_step(ddp, opt, step)
        if dist.get_rank() == 0:
            print(f"step {step} loss {loss:.4f}")
    # ---- checkpoint (model + optimizer, asynchronous: pinned staging + worker processes), then RESUME into a fresh model /
    # optimizer and check that training continues exactly where it stopped (reference: finetune_4D.py:355,390 save / load)
    box = [tempfile.mkdtemp() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, 0)
    ckpt.save(box[0], {"model": model, "optimizer": opt}, async_checkpoint=True)
    ckpt.wait_for_async()
    cont = train_step(ddp, opt, args.steps)  # the original run, one step further
    model2, ddp2, opt2 = build()  # fresh weights and optimizer state ...
    train_step(ddp2, opt2, 12345)  # ... moved away from the initial state, so a silent no-op load would be caught
    ckpt.load(box[0], {"model": model2, "optimizer": opt2})
    resumed = train_step(ddp2, opt2, args.steps)
    assert abs(resumed - cont) < 1e-5, (resumed, cont)
    if dist.get_rank() == 0:
        print(f"checkpoint resume ok: next-step loss {resumed:.6f} == {cont:.6f} ({box[0]})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
