"""Ulysses sequence<->head all-to-all attention vs attention on the gathered tensors (4 ranks)."""
import torch
import torch.nn.functional as F

from common import device_type, run_distributed


def _ulysses(rank, world):
    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.context import ulysses_attention

    dev = device_type()
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("cp",))
    B, S, Hq, Hk, D = 2, 8 * world, 2 * world, world, 16
    g = torch.Generator().manual_seed(3)
    q = torch.randn(B, S, Hq, D, generator=g).to(dev)
    k = torch.randn(B, S, Hk, D, generator=g).to(dev)
    v = torch.randn(B, S, Hk, D, generator=g).to(dev)
    w = torch.randn(B, S, Hq, D, generator=g).to(dev)
    qf, kf, vf = (t.clone().requires_grad_(True) for t in (q, k, v))
    ref = F.scaled_dot_product_attention(qf.transpose(1, 2), kf.transpose(1, 2), vf.transpose(1, 2), is_causal=True, enable_gqa=True).transpose(1, 2)
    (ref * w).sum().backward()
    sl = slice(rank * S // world, (rank + 1) * S // world)
    ql, kl, vl = (t[:, sl].clone().requires_grad_(True) for t in (q, k, v))
    out = ulysses_attention(ql, kl, vl, mesh, "cp", causal=True)
    (out * w[:, sl]).sum().backward()
    torch.testing.assert_close(out, ref[:, sl], rtol=1e-4, atol=1e-5)
    for a, b in ((ql, qf), (kl, kf), (vl, vf)):
        torch.testing.assert_close(a.grad, b.grad[:, sl], rtol=1e-4, atol=1e-5)
    # all-gather-KV context parallelism: same result, dK/dV come back through a reduce-scatter
    from vescale_b200.parallel.context import allgather_kv_attention

    ql2, kl2, vl2 = (t[:, sl].clone().requires_grad_(True) for t in (q, k, v))
    out2 = allgather_kv_attention(ql2, kl2, vl2, mesh, "cp", causal=True)
    (out2 * w[:, sl]).sum().backward()
    torch.testing.assert_close(out2, ref[:, sl], rtol=1e-4, atol=1e-5)
    for a, b in ((ql2, qf), (kl2, kf), (vl2, vf)):
        torch.testing.assert_close(a.grad, b.grad[:, sl], rtol=1e-4, atol=1e-5)


def test_ulysses_attention_matches_full():
    run_distributed(_ulysses, 4)
