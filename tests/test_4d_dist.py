"""Pipeline x data x tensor parallelism composed (2 x 2 x 2 = 8 ranks): auto-planned TP/SP stages (DModule), data-parallel
gradient averaging over the DP group, 1F1B pipeline engine — against single-process training on the same data.
Strategy parity: the reference's 4-D examples (``legacy/examples/nanogpt_4D_finetune``, ``llama2_4D_finetune``) assert loss-curve
agreement with a single device."""
import copy

import torch
import torch.distributed as dist
import torch.nn as nn

from common import device_type, run_distributed


class GPTBlock(nn.Module):
    def __init__(self, h=32):
        super().__init__()
        self.ln_1 = nn.LayerNorm(h)
        self.c_fc = nn.Linear(h, 4 * h)
        self.c_proj = nn.Linear(4 * h, h)

    def forward(self, x):
        return x + self.c_proj(torch.nn.functional.gelu(self.c_fc(self.ln_1(x))))


class _LocalOut(nn.Module):
    """Stage wrapper: TP/SP DTensors stay inside the stage; the pipeline sees plain local tensors."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, x):
        from vescale_b200.dtensor import DTensor

        y = self.inner(x)
        return y.to_local() if isinstance(y, DTensor) else y


def _4d(rank, world):
    from vescale_b200 import Replicate
    from vescale_b200.devicemesh_api import VESCALE_DEVICE_MESH
    from vescale_b200.dtensor import DTensor
    from vescale_b200.parallel.dmp import auto_parallelize_module
    from vescale_b200.parallel.pipe import PipeEngine, PipelineParallelPlan, PipelineScheduleType, construct_pipeline_stage

    dev = device_type()
    mesh = VESCALE_DEVICE_MESH.init_device_mesh(dev, (2, 2, 2), mesh_dim_names=("PP", "DP", "TP"))
    pp_rank, dp_rank = mesh.get_local_rank("PP"), mesh.get_local_rank("DP")
    torch.manual_seed(0)
    ref = nn.Sequential(*[GPTBlock() for _ in range(4)]).to(dev)
    model = copy.deepcopy(ref)
    plan = PipelineParallelPlan(num_stages=2, schedule_type=PipelineScheduleType.SIMPLE_1F1B)
    pm = construct_pipeline_stage(model, plan, mesh)
    # tensor / sequence parallelism inside this rank's stage: Megatron plan per block, replicated activations at the borders
    stage = pm.chunk(0)
    for blk in stage.mods:
        auto_parallelize_module(blk, mesh["TP"], "MEGATRON", plan_override={"forward": {r"input": [[Replicate()]], r"c_proj\.output": [[Replicate()]]}})
    pm.stage_modules["0"] = _LocalOut(stage)
    loss_fn = lambda out, y: torch.nn.functional.mse_loss(out, y)  # noqa: E731
    engine = PipeEngine(pm, mesh, loss_fn, plan)
    dp_group = mesh.get_group("DP")
    params = [p for p in pm.parameters()]
    lr, M = 0.05, 4
    ref_params = dict(ref.named_parameters())
    for step in range(2):
        # golden: both DP replicas' micro-batches, averaged
        ref.zero_grad()
        ref_loss = 0.0
        for r in range(2):
            g = torch.Generator().manual_seed(100 * step + r)
            xs = [torch.randn(3, 5, 32, generator=g).to(dev) for _ in range(M)]
            ys = [torch.randn(3, 5, 32, generator=g).to(dev) for _ in range(M)]
            for x, y in zip(xs, ys):
                l = loss_fn(ref(x), y) / (M * 2)
                l.backward()
                ref_loss += l.item()
        with torch.no_grad():
            for p in ref.parameters():
                p -= lr * p.grad
        # 3-D parallel run: this DP replica's micro-batches through the pipeline
        g = torch.Generator().manual_seed(100 * step + dp_rank)
        xs = [torch.randn(3, 5, 32, generator=g).to(dev) for _ in range(M)]
        ys = [torch.randn(3, 5, 32, generator=g).to(dev) for _ in range(M)]
        engine.zero_grad()
        loss, _ = engine(xs, ys)
        for blk in stage.mods:  # TP: partial gradients of replicated parameters (LayerNorm under SP, row-parallel bias)
            dm = getattr(blk, "_dmodule", None)
            if dm is not None:
                dm.finish_grad_sync()
        with torch.no_grad():
            for p in params:
                if p.grad is None:
                    continue
                gl = p.grad._local_tensor if isinstance(p.grad, DTensor) else p.grad
                dist.all_reduce(gl, group=dp_group)  # data-parallel average
                gl.div_(2)
                pl = p._local_tensor if isinstance(p, DTensor) else p.data
                (pl.data if hasattr(pl, "data") else pl).sub_(lr * gl)
        if engine.is_last_rank:
            tot = loss.detach().clone()
            dist.all_reduce(tot, group=dp_group)
            assert abs(tot.item() / 2 - ref_loss) < 1e-5, (step, tot.item() / 2, ref_loss)
    # weights after two steps equal the golden model's (each rank checks its stage's blocks, gathered over TP)
    names = stage.names
    for i, blk in enumerate(stage.mods):
        for n, p in blk.named_parameters():
            full = p.full_tensor() if isinstance(p, DTensor) else (p.data.full_tensor() if isinstance(p.data, DTensor) else p.data)
            torch.testing.assert_close(full, ref_params[f"{names[i]}.{n}"].detach(), rtol=1e-4, atol=1e-5, msg=f"{names[i]}.{n}")


def test_pp_dp_tp_composition_matches_single_process():
    run_distributed(_4d, 8)
