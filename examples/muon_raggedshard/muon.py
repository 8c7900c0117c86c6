"""Muon on sharded parameters with RaggedShard: gather each 2-D parameter to a root rank (one-hot ``local_units``),
run the Newton–Schulz orthogonalisation there, scatter the update back — three ``DTensor.redistribute`` calls
(``docs/texts/raggedshard.md:79-91``; one-hot units as in ``test/dtensor/ragged_shard/test_redistribute.py:147-150``).

    torchrun --nproc-per-node 4 examples/muon_raggedshard/muon.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from vescale_b200 import distribute_tensor, init_device_mesh  # noqa: E402
from vescale_b200.dtensor import RaggedShard  # noqa: E402


def newton_schulz(G: torch.Tensor, steps: int = 5) -> torch.Tensor:
    a, b, c = 3.4445, -4.7750, 2.0315
    X = G / (G.norm() + 1e-7)
    T = X.shape[0] > X.shape[1]
    X = X.t() if T else X
    for _ in range(steps):
        A = X @ X.t()
        X = a * X + (b * A + c * A @ A) @ X
    return X.t() if T else X


def muon_step(params, grads, lr=0.02, mesh=None):
    world = mesh.size()
    for i, (p, g) in enumerate(zip(params, grads)):
        root = i % world  # load-balance roots round-robin
        one_hot = RaggedShard((0,), tuple(1 if r == root else 0 for r in range(world)))
        g_root = g.redistribute(mesh, [one_hot])  # gather-to-root = uneven all-to-all
        local = g_root.to_local()
        if local.numel():
            upd = newton_schulz(local.view(p.shape)).reshape(-1)
        else:
            upd = local
        upd_dt = type(g_root).from_local(upd, mesh, [one_hot], shape=p.shape)
        upd_sharded = upd_dt.redistribute(mesh, p.placements)  # scatter back
        with torch.no_grad():
            p.to_local().add_(upd_sharded.to_local(), alpha=-lr)


def main():
    dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    world = dist.get_world_size()
    mesh = init_device_mesh(dev, (world,))
    torch.manual_seed(0)
    shapes = [(64, 32), (48, 96), (32, 32)]
    full = [torch.randn(s, device=dev) for s in shapes]
    units = tuple([1] * world)
    params = [distribute_tensor(w, mesh, [RaggedShard((0,), units)]) for w in full]
    grads = [distribute_tensor(torch.randn(s, device=dev), mesh, [RaggedShard((0,), units)]) for s in shapes]
    ref = [w - 0.02 * newton_schulz(g.full_tensor()) for w, g in zip(full, grads)]
    muon_step(params, grads, mesh=mesh)
    for p, r in zip(params, ref):
        torch.testing.assert_close(p.full_tensor(), r, rtol=1e-4, atol=1e-5)
    if dist.get_rank() == 0:
        print("muon over RaggedShard matches the single-device update")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
