"""ndtimeline over a PP x DP x TP run: every rank records its forward / backward compute, pipeline p2p and gradient collectives on
the aligned global clock; the per-rank chrome traces are merged into one timeline (open in chrome://tracing or ui.perfetto.dev).

    python examples/ndtimeline_4d/trace.py --out profiles/ndtimeline_4d_trace.json      # 8 ranks (gloo on CPU, NCCL with >= 8 GPUs)

Parity: the reference wires ``@ndtimer(FORWARD_COMPUTE/BACKWARD_COMPUTE)`` and ``ndtimeit_p2p`` into its pipeline schedules
(``legacy/vescale/pipe/_schedules/pipedream_flush.py:1114,1190``, ``p2p_communication.py:624-847``); here the engine, the p2p layer,
the FSDP wrapper and the optimizers are instrumented (``parallel/pipe/engine.py``, ``p2p.py``, ``parallel/fsdp/api.py``).
"""
import argparse
import glob
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class Block(nn.Module):
    def __init__(self, h=64):
        super().__init__()
        self.ln_1 = nn.LayerNorm(h)
        self.c_fc = nn.Linear(h, 4 * h)
        self.c_proj = nn.Linear(4 * h, h)

    def forward(self, x):
        return x + self.c_proj(torch.nn.functional.gelu(self.c_fc(self.ln_1(x))))


class _LocalOut(nn.Module):
    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, x):
        from vescale_b200.dtensor import DTensor

        y = self.inner(x)
        return y.to_local() if isinstance(y, DTensor) else y


def worker(rank, world, out_dir):
    from common import device_type

    from vescale_b200 import Replicate
    from vescale_b200 import profiler as ndt
    from vescale_b200.devicemesh_api import VESCALE_DEVICE_MESH
    from vescale_b200.dtensor import DTensor
    from vescale_b200.parallel.dmp import auto_parallelize_module
    from vescale_b200.parallel.pipe import PipeEngine, PipelineParallelPlan, PipelineScheduleType, construct_pipeline_stage

    dev = device_type()
    mesh = VESCALE_DEVICE_MESH.init_device_mesh(dev, (2, 2, 2), mesh_dim_names=("PP", "DP", "TP"))
    ndt.init_ndtimers(rank=rank, world_size=world, handlers=[ndt.ChromeTraceNDHandler(out_dir, "trace")])
    torch.manual_seed(0)
    model = nn.Sequential(*[Block() for _ in range(4)]).to(dev)
    plan = PipelineParallelPlan(num_stages=2, schedule_type=PipelineScheduleType.SIMPLE_1F1B, batch_p2p_comm=True)
    pm = construct_pipeline_stage(model, plan, mesh)
    stage = pm.chunk(0)
    for blk in stage.mods:
        auto_parallelize_module(blk, mesh["TP"], "MEGATRON", plan_override={"forward": {r"input": [[Replicate()]], r"c_proj\.output": [[Replicate()]]}})
    pm.stage_modules["0"] = _LocalOut(stage)
    engine = PipeEngine(pm, mesh, lambda o, y: torch.nn.functional.mse_loss(o, y), plan)
    dp_group = mesh.get_group("DP")
    for step in range(3):
        g = torch.Generator().manual_seed(100 * step + mesh.get_local_rank("DP"))
        xs = [torch.randn(4, 16, 64, generator=g).to(dev) for _ in range(8)]
        ys = [torch.randn(4, 16, 64, generator=g).to(dev) for _ in range(8)]
        engine.zero_grad()
        engine(xs, ys)
        with ndt.ndtimeit(ndt.predefined.GRAD_AR):
            for blk in stage.mods:
                dm = getattr(blk, "_dmodule", None)
                if dm is not None:
                    dm.finish_grad_sync()
            for p in pm.parameters():
                if p.grad is not None:
                    gl = p.grad._local_tensor if isinstance(p.grad, DTensor) else p.grad
                    dist.all_reduce(gl, group=dp_group)
        with ndt.ndtimeit(ndt.predefined.OPTIMIZER_STEP):
            with torch.no_grad():
                for p in pm.parameters():
                    if p.grad is not None:
                        gl = p.grad._local_tensor if isinstance(p.grad, DTensor) else p.grad
                        pl = p._local_tensor if isinstance(p, DTensor) else p.data
                        pl.sub_(0.01 * gl / 2)
        ndt.inc_step()
    ndt.flush(asynchronous=False)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="ndtimeline_4d_trace.json")
    args = ap.parse_args()
    from common import run_distributed

    with tempfile.TemporaryDirectory() as d:
        run_distributed(worker, 8, d)
        events = []
        for f in sorted(glob.glob(os.path.join(d, "trace_rank*.json"))):
            events += json.load(open(f))["traceEvents"]
    t0 = min(e["ts"] for e in events)
    names = {0: "PP0 DP0 TP0", 1: "PP0 DP0 TP1", 2: "PP0 DP1 TP0", 3: "PP0 DP1 TP1", 4: "PP1 DP0 TP0", 5: "PP1 DP0 TP1", 6: "PP1 DP1 TP0", 7: "PP1 DP1 TP1"}
    for e in events:
        e["ts"] = round(e["ts"] - t0, 1)
        e["dur"] = round(e["dur"], 1)
    meta = [{"name": "process_name", "ph": "M", "pid": r, "args": {"name": f"rank {r} ({n})"}} for r, n in names.items()]
    with open(args.out, "w") as f:
        json.dump({"traceEvents": meta + sorted(events, key=lambda e: (e["pid"], e["ts"]))}, f)
    kinds = sorted({e["name"] for e in events})
    print(f"{len(events)} events from 8 ranks -> {args.out}; metrics: {kinds}")


if __name__ == "__main__":
    main()
