"""DTensor-level fusion of redistribute -> mm / mm -> redistribute (dtensor/fusion.py) on 2 and 4 ranks: the pattern match fires,
results and gradients equal the unfused path and the single-device model.  On CPU the c10d back end stands in for the sm_100a
kernels (same handler, same control flow); tests/test_symm_multigpu.py runs the fused kernels on >= 2 GPUs."""
import os

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from common import device_type, run_distributed


def _direct(rank, world):
    os.environ["VESCALE_B200_FUSE_TP"] = "c10d"
    from vescale_b200 import Partial, Replicate, Shard, distribute_tensor, init_device_mesh
    from vescale_b200.dtensor import fusion

    dev = device_type()
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("TP",))
    g = torch.Generator().manual_seed(0)
    x, w1, w2 = torch.randn(8 * world, 12, generator=g).to(dev), torch.randn(16 * world, 12, generator=g).to(dev), torch.randn(12, 16 * world, generator=g).to(dev)
    # all-gather ⊕ GEMM: row-sharded activation into a column-parallel weight
    dx = distribute_tensor(x, mesh, [Shard(0)], src_data_rank=None)
    dw1 = distribute_tensor(w1, mesh, [Shard(0)], src_data_rank=None)
    fusion.reset_stats()
    h = dx @ dw1.t()
    assert fusion.stats["ag_gemm"] == 1 and h.placements == (Shard(1),), (fusion.stats, h.placements)
    torch.testing.assert_close(h.full_tensor(), x @ w1.t(), rtol=1e-5, atol=1e-5)
    # GEMM ⊕ reduce-scatter: needs the resharding hint; without it the result is Partial as always
    dw2 = distribute_tensor(w2, mesh, [Shard(1)], src_data_rank=None)
    y_plain = h @ dw2.t()
    assert any(p.is_partial() for p in y_plain.placements) and fusion.stats["gemm_rs"] == 0
    with fusion.fuse_reshard(mesh, [Shard(0)]):
        y = h @ dw2.t()
    assert fusion.stats["gemm_rs"] == 1 and y.placements == (Shard(0),), (fusion.stats, y.placements)
    ref = (x @ w1.t()) @ w2.t()
    torch.testing.assert_close(y.full_tensor(), ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(y_plain.full_tensor(), ref, rtol=1e-4, atol=1e-4)
    # F.linear on a (1, S, H) sequence-sharded activation: view -> mm -> view keeps the pattern
    x3 = x.view(1, 8 * world, 12)
    d3 = distribute_tensor(x3, mesh, [Shard(1)], src_data_rank=None)
    fusion.reset_stats()
    h3 = F.linear(d3, dw1)
    assert fusion.stats["ag_gemm"] == 1, fusion.stats
    torch.testing.assert_close(h3.full_tensor(), F.linear(x3, w1), rtol=1e-5, atol=1e-5)
    # off switch
    os.environ["VESCALE_B200_FUSE_TP"] = "off"
    fusion.reset_stats()
    h_off = dx @ dw1.t()
    assert fusion.stats["ag_gemm"] == 0
    torch.testing.assert_close(h_off.full_tensor(), x @ w1.t(), rtol=1e-5, atol=1e-5)
    os.environ["VESCALE_B200_FUSE_TP"] = "c10d"


class MLP(nn.Module):
    def __init__(self, h, f):
        super().__init__()
        self.fc1 = nn.Linear(h, f, bias=False)
        self.fc2 = nn.Linear(f, h, bias=False)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class Net(nn.Module):
    def __init__(self, h=16, f=32):
        super().__init__()
        self.norm = nn.LayerNorm(h)
        self.mlp = MLP(h, f)

    def forward(self, x):
        return x + self.mlp(self.norm(x))


def _dmodule(rank, world):
    """A Megatron sequence-parallel plan (activations Shard(0) over tokens between blocks, column -> row parallel MLP): the
    DModule output plan is the resharding hint, so fc1 runs as all-gather ⊕ GEMM and fc2 as GEMM ⊕ reduce-scatter."""
    os.environ["VESCALE_B200_FUSE_TP"] = "c10d"
    import copy

    from vescale_b200 import Replicate, Shard, init_device_mesh
    from vescale_b200.dtensor import fusion
    from vescale_b200.parallel.dmodule import parallelize_module

    dev = device_type()
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("TP",))
    torch.manual_seed(0)
    ref = Net().to(dev)
    plan = {
        "parameter": {r"mlp\.fc1\.weight": [Shard(0)], r"mlp\.fc2\.weight": [Shard(1)]},
        "forward": {r"input": [[Shard(0)]], r"mlp\.output": [[Shard(0)]], r"output": [[Shard(0)]]},
    }
    x = torch.randn(8 * world, 16, generator=torch.Generator().manual_seed(1)).to(dev)
    results = {}
    for mode in ("c10d", "off"):
        os.environ["VESCALE_B200_FUSE_TP"] = mode
        m = parallelize_module(copy.deepcopy(ref), mesh, plan)
        fusion.reset_stats()
        out = m(x.chunk(world)[rank].clone())
        if mode == "c10d":
            assert fusion.stats["ag_gemm"] == 1 and fusion.stats["gemm_rs"] == 1, fusion.stats
        else:
            assert fusion.stats["ag_gemm"] == 0 and fusion.stats["gemm_rs"] == 0, fusion.stats
        out.to_local().sum().backward()
        m.finish_grad_sync()
        results[mode] = (out.full_tensor(), {n: p.grad.full_tensor() for n, p in m.named_parameters()})
    xr = x.clone()
    o = ref(xr)
    o.sum().backward()
    for mode in results:
        torch.testing.assert_close(results[mode][0], o.detach(), rtol=1e-4, atol=1e-5, msg=lambda s: f"{mode}: {s}")
        for n, p in ref.named_parameters():
            torch.testing.assert_close(results[mode][1][n], p.grad, rtol=1e-4, atol=1e-5, msg=lambda s: f"{mode} {n}: {s}")


@pytest.mark.parametrize("world", [2, 4])
def test_mm_fusion_patterns(world):
    run_distributed(_direct, world)


def test_dmodule_plan_hits_fused_paths():
    run_distributed(_dmodule, 2)
