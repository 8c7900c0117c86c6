"""PP x DP Llama training with the pipeline engine (1F1B / interleaved / zero-bubble) over FSDP-free stages.
    torchrun --nproc-per-node 4 examples/llama_4D_finetune/pp_train.py --pp 2 --dp 2 --schedule ZERO_BUBBLE
(reference: ``legacy/examples/llama2_4D_finetune/llama_train.py``)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from vescale_b200.devicemesh_api import VESCALE_DEVICE_MESH  # noqa: E402
from vescale_b200.models import LlamaConfig, LlamaModel  # noqa: E402
from vescale_b200.parallel.pipe import PipeEngine, PipelineParallelPlan, PipelineScheduleType, construct_pipeline_stage  # noqa: E402
from vescale_b200 import ops as O  # noqa: E402


class _Embed(nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, tokens):
        h = self.m.embed(tokens)
        return h, torch.zeros_like(h)


class _Layer(nn.Module):
    def __init__(self, m, blk):
        super().__init__()
        self.blk, self.m = blk, [m]

    def forward(self, h, delta):
        cos, sin = self.m[0].rope(h.shape[1], h.device)
        return self.blk(h, delta, cos, sin)


class _Head(nn.Module):
    def __init__(self, m):
        super().__init__()
        self.head = m.head

    def forward(self, h, delta):
        return self.head(h, delta)


def units_of(model: LlamaModel):
    return [("embed", _Embed(model))] + [(f"layers.{i}", _Layer(model, b)) for i, b in enumerate(model.layers)] + [("head", _Head(model))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pp", type=int, default=2)
    ap.add_argument("--dp", type=int, default=2)
    ap.add_argument("--schedule", default="SIMPLE_1F1B")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    cuda = torch.cuda.is_available()
    dist.init_process_group("nccl" if cuda else "gloo")
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = "cuda" if cuda else "cpu"
    mesh = VESCALE_DEVICE_MESH.init_device_mesh(dev, (args.pp, args.dp), mesh_dim_names=("PP", "DP"))
    cfg = LlamaConfig.tiny(num_layers=6)
    model = LlamaModel(cfg, device=dev).reset_parameters(0)
    model.pipeline_units = lambda: units_of(model)
    plan = PipelineParallelPlan(num_stages=args.pp, schedule_type=PipelineScheduleType[args.schedule])
    pm = construct_pipeline_stage(model, plan, mesh)
    engine = PipeEngine(pm, mesh, lambda logits, y: O.cross_entropy(logits.reshape(-1, logits.shape[-1]).contiguous(), y.reshape(-1)), plan)
    opt = torch.optim.AdamW(pm.parameters(), lr=1e-3)
    dp_group = mesh.get_group("DP")
    g = torch.Generator().manual_seed(VESCALE_DEVICE_MESH.get_data_parallel_rank())
    for step in range(args.steps):
        toks = [torch.randint(0, cfg.vocab_size, (2, 17), generator=g).to(dev) for _ in range(4)]
        opt.zero_grad()
        loss, _ = engine([t[:, :-1] for t in toks], [t[:, 1:] for t in toks])
        for p in pm.parameters():  # data-parallel gradient average
            if p.grad is not None:
                dist.all_reduce(p.grad, group=dp_group)
                p.grad.div_(args.dp)
        opt.step()
        if engine.is_last_rank and VESCALE_DEVICE_MESH.get_data_parallel_rank() == 0:
            print(f"step {step} loss {loss.item():.4f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
