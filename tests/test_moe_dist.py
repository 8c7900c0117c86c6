"""Expert parallelism: EP=4 Mixtral-tiny must match the same model with all experts on one device
(legacy ``test/parallel/ddp_optim/test_moe.py`` / ``test/model/mixtral`` strategy)."""
import torch
import torch.distributed as dist

from common import device_type, run_distributed


def _ep(rank, world):
    from vescale_b200 import init_device_mesh
    from vescale_b200.models import MixtralConfig, MixtralModel
    from vescale_b200.parallel.moe import MoEOptimizer, ragged_token_placement

    dev = device_type()
    cfg = MixtralConfig.tiny()
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("EP",))
    ref = MixtralModel(cfg).reset_parameters(seed=3).to(dev)  # all 8 experts local
    model = MixtralModel(cfg, ep_group=mesh.get_group(0)).reset_parameters(seed=3).to(dev)
    assert model.layers[0].moe.num_local == 2
    # same expert weights whatever the EP size
    e0 = rank * 2
    assert torch.equal(model.layers[0].moe.experts.w_down[0], ref.layers[0].moe.experts.w_down[e0])
    toks = []
    for r in range(world):
        g = torch.Generator().manual_seed(50 + r)
        toks.append(torch.randint(0, cfg.vocab_size, (2, 17), generator=g).to(dev))
    mine = toks[rank]
    loss = model(mine[:, :-1], mine[:, 1:])
    loss.backward()
    ref_losses = [ref(t[:, :-1], t[:, 1:]) for t in toks]
    (sum(ref_losses) / world).backward()
    torch.testing.assert_close(loss.detach(), ref_losses[rank].detach(), rtol=1e-4, atol=1e-5)
    tp = model.layers[0].moe.last_tokens_per_rank
    assert sum(tp) > 0 and ragged_token_placement(tp).total_units > 0
    opt = MoEOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), model, dp_group=None, ep_group=mesh.get_group(0))
    opt.step()  # averages dense grads over all ranks, scales expert grads
    gd = model.layers[1].wqkv.grad
    torch.testing.assert_close(gd, ref.layers[1].wqkv.grad, rtol=1e-3, atol=1e-6)
    ge = model.layers[1].moe.experts.w_gate_up.grad[1]
    torch.testing.assert_close(ge, ref.layers[1].moe.experts.w_gate_up.grad[e0 + 1], rtol=1e-3, atol=1e-6)


def test_expert_parallel_matches_local_experts():
    run_distributed(_ep, 4)


def _realloc(rank, world):
    """Experts (weights + Adam state) migrate between EP ranks mid-training; the function computed and the optimizer
    trajectory are unchanged (legacy ``_moe_param_buffer.py:183-337`` dynamic re-allocation)."""
    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.moe import MoEConfig, MoELayer
    from vescale_b200.parallel.moe.api import balanced_allocation, reallocate_experts

    dev = device_type()
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("EP",))
    cfg = MoEConfig(32, 64, 8, 2, dtype=torch.float32)
    layers, opts = [], []
    for _ in range(2):  # twin A is re-allocated, twin B is not
        l = MoELayer(cfg, mesh.get_group(0), device=dev)
        l.reset_parameters(torch.Generator().manual_seed(4))
        layers.append(l)
        opts.append(torch.optim.Adam(l.parameters(), lr=1e-2))

    def step(l, o, s):
        x = torch.randn(24, 32, generator=torch.Generator().manual_seed(100 * s + rank)).to(dev)
        y = l(x)
        o.zero_grad()
        y.pow(2).mean().backward()
        for p in l.parameters():  # data-parallel part of the step: the router is replicated over EP ranks
            if not getattr(p, "_is_expert_param", False):
                dist.all_reduce(p.grad)
        o.step()
        return y.detach()

    for s in range(2):
        ya, yb = step(layers[0], opts[0], s), step(layers[1], opts[1], s)
        assert torch.equal(ya, yb)
    load = [5, 1, 9, 2, 7, 3, 8, 4]
    slots = balanced_allocation(load, world)
    assert sorted(slots) == list(range(8))
    per_rank = [sum(load[e] for e in range(8) if slots[e] // 2 == r) for r in range(world)]
    assert max(per_rank) - min(per_rank) <= 2
    reallocate_experts(layers[0], slots, opts[0])
    assert layers[0].slot_of_expert.tolist() == slots
    for s in range(2, 5):
        ya, yb = step(layers[0], opts[0], s), step(layers[1], opts[1], s)
        torch.testing.assert_close(ya, yb, rtol=1e-5, atol=1e-6)
    # and back to the identity layout: weights return to their original owners bit for bit
    reallocate_experts(layers[0], list(range(8)), opts[0])
    torch.testing.assert_close(layers[0].experts.w_down, layers[1].experts.w_down, rtol=1e-5, atol=1e-6)


def test_dynamic_expert_reallocation():
    run_distributed(_realloc, 4)
