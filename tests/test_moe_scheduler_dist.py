"""MoE scheduling with per-expert DP x TP allocations (``parallel/moe/scheduler.py``): replicated and tensor-parallel experts, replica
gradient sync, re-allocation with optimizer-state migration, and the reference-shaped allocator / dispatcher contract
(legacy ``moe/_scheduler.py``, ``experts_allocator.py``, ``token_dispatcher.py``, ``_moe_param_buffer.py`` refresh)."""
import torch
import torch.distributed as dist
import torch.nn.functional as F

from common import device_type, run_distributed

E, H, I, K, T = 4, 32, 64, 2, 24

ALLOC_A = [  # TP = 2 everywhere; experts 0 and 3 have two replicas
    [[0, 1], [2, 3]],
    [[1, 2]],
    [[3, 0]],
    [[2, 3], [0, 1]],
]
ALLOC_B = [  # after re-allocation: different hosts, expert 1 becomes the replicated one
    [[1, 0]],
    [[0, 3], [2, 1]],
    [[3, 2]],
    [[1, 2]],
]


def _dense_moe(x, router_w, wgu, wdn):
    """Single-device reference: every expert whole, top-k softmax routing renormalised (TopKRouter)."""
    probs = torch.softmax(F.linear(x.float(), router_w), -1)
    topv, topi = torch.topk(probs, K, -1)
    topv = topv / topv.sum(-1, keepdim=True)
    out = torch.zeros_like(x)
    for e in range(E):
        gate_up = x @ wgu[e].t()
        y = (F.silu(gate_up[:, :I]) * gate_up[:, I:]) @ wdn[e].t()
        w = (topv * (topi == e)).sum(-1, keepdim=True)
        out = out + w * y
    return out


def _setup(rank, world, alloc_meshes):
    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.moe import MoEConfig
    from vescale_b200.parallel.moe.scheduler import ExpertsAllocation, ScheduledMoELayer

    dev = device_type()
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("EP",))
    cfg = MoEConfig(H, I, E, K, dtype=torch.float32)
    g = torch.Generator().manual_seed(11)
    router_w = torch.randn(E, H, generator=g) * 0.5
    wgu = torch.randn(E, 2 * I, H, generator=g) * 0.2
    wdn = torch.randn(E, H, I, generator=g) * 0.2
    xs = [torch.randn(T, H, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    layer = ScheduledMoELayer(cfg, ExpertsAllocation(alloc_meshes, world), mesh.get_group(0), device=dev)
    with torch.no_grad():
        layer.router.weight.copy_(router_w)
    layer.load_full_experts(wgu.to(dev), wdn.to(dev))
    return layer, router_w, wgu, wdn, xs, dev


def _golden_step(router_w, wgu, wdn, xs):
    ws = [t.clone().requires_grad_() for t in (router_w, wgu, wdn)]
    outs = [_dense_moe(x, *ws) for x in xs]
    sum((o * o).sum() for o in outs).backward()
    return [o.detach() for o in outs], [w.grad for w in ws], ws


def _check_shards(layer, full_gu, full_dn, what, tol=1e-4):
    from vescale_b200.parallel.moe.scheduler import ScheduledMoELayer

    for i, (e, _r, t) in enumerate(layer.alloc.hosted[layer.ep_rank]):
        gu, dn = ScheduledMoELayer.shard_of(full_gu[e], full_dn[e], t, layer.alloc.tp)
        torch.testing.assert_close(what(layer.experts.w_gate_up)[i].cpu(), gu, rtol=tol, atol=tol, msg=lambda m: f"gate_up shard {(e, t)}: {m}")
        torch.testing.assert_close(what(layer.experts.w_down)[i].cpu(), dn, rtol=tol, atol=tol, msg=lambda m: f"down shard {(e, t)}: {m}")


def _replicated_tp(rank, world):
    from vescale_b200.parallel.moe.scheduler import ExpertsAllocation

    layer, router_w, wgu, wdn, xs, dev = _setup(rank, world, ALLOC_A)
    assert layer.alloc.tp == 2 and layer.alloc.slots_per_rank == 3 and layer.alloc.dp_size.tolist() == [2, 1, 1, 2]
    opt = torch.optim.Adam(layer.parameters(), lr=1e-2)
    gold_opt = None
    gw = [router_w.clone(), wgu.clone(), wdn.clone()]
    for step in range(3):
        if step == 2:  # move to a different allocation between steps: weights AND Adam moments migrate
            layer.reallocate(ExpertsAllocation(ALLOC_B, world), opt)
            assert layer.alloc.dp_size.tolist() == [1, 2, 1, 1]
            _check_shards(layer, gw[1], gw[2], lambda p: p.data)
        outs, grads, leaves = _golden_step(*gw, xs)
        out = layer(xs[rank].to(dev))
        torch.testing.assert_close(out.cpu(), outs[rank], rtol=2e-4, atol=2e-4)
        opt.zero_grad()
        (out * out).sum().backward()
        layer.sync_replica_grads()
        dist.all_reduce(layer.router.weight.grad)  # the router is replicated on every rank
        _check_shards(layer, grads[1], grads[2], lambda p: p.grad, tol=5e-4)
        torch.testing.assert_close(layer.router.weight.grad.cpu(), grads[0], rtol=5e-4, atol=5e-4)
        opt.step()
        # the same Adam step on the full weights
        if gold_opt is None:
            gold_params = [torch.nn.Parameter(w.clone()) for w in gw]
            gold_opt = torch.optim.Adam(gold_params, lr=1e-2)
        for p, g_ in zip(gold_params, grads):
            p.grad = g_.clone()
        gold_opt.step()
        gw = [p.detach().clone() for p in gold_params]
        _check_shards(layer, gw[1], gw[2], lambda p: p.data, tol=5e-4)


def test_replicated_tensor_parallel_experts_and_reallocation():
    run_distributed(_replicated_tp, 4)


class _Allocator:
    """Reference-shaped allocator: a list of per-expert [DP, TP] meshes the first time a layer is seen, a new one at iteration 1,
    ``None`` (keep) otherwise; records the loads it is shown."""

    def __init__(self):
        self.seen, self.perf = set(), []

    def allocate_experts(self, layer_id, iter=-1):
        if iter == 1 and ("moved", layer_id) not in self.seen:
            self.seen.add(("moved", layer_id))
            return [torch.tensor(m) for m in ALLOC_B]
        return None

    def collect_performance(self, perf, iter=-1):
        self.perf.append((iter, perf["layer_id"], perf["tokens_per_expert"].tolist(), list(perf["tokens_per_rank"])))


class _RoundRobinDispatcher:
    """Reference-shaped dispatcher: replica = token id modulo the expert's replica count."""

    def __init__(self):
        self.calls = 0

    def set_experts_alloc(self, info):
        self.num_replicate = info["dp_size"]

    def assign_task(self, layer_id, token_id, expert_id, hidden_state, token_weight):
        self.token_id, self.expert_id = token_id, expert_id

    def collect_performance(self, perf, iter=-1):
        pass

    def dispatch_token(self, layer_id):
        self.calls += 1
        return self.expert_id, self.token_id % self.num_replicate.to(self.expert_id.device)[self.expert_id]


def _scheduler(rank, world):
    from vescale_b200.parallel.moe.scheduler import MoEScheduler

    layer, router_w, wgu, wdn, xs, dev = _setup(rank, world, ALLOC_A)
    opt = torch.optim.SGD(layer.parameters(), lr=0.0)
    alloc, disp = _Allocator(), _RoundRobinDispatcher()
    sched = MoEScheduler(alloc, disp, optimizer=opt).register(layer)
    gold = _dense_moe(xs[rank], router_w, wgu, wdn)
    for it in range(3):
        out = layer(xs[rank].to(dev))
        torch.testing.assert_close(out.detach().cpu(), gold, rtol=2e-4, atol=2e-4)
        (out * out).sum().backward()
        sched.step_end()
        want = ALLOC_A if it == 0 else ALLOC_B
        assert [m.tolist() for m in layer.alloc.meshes] == want, (it, layer.alloc.meshes)
    assert disp.calls == 3 and len(alloc.perf) == 3
    it, lid, per_expert, per_rank = alloc.perf[-1]
    assert it == 2 and lid == 0 and sum(per_expert) == T * K and len(per_rank) == world
    sched.remove()


def test_scheduler_with_reference_shaped_allocator_and_dispatcher():
    run_distributed(_scheduler, 4)


def _parallelize(rank, world):
    """``parallelize_experts`` with a reference-shaped allocator (overrides ``allocate_experts``) turns a local MoE layer into a
    scheduled one with replicated / sharded experts; the function computed does not change."""
    import torch.nn as nn

    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.moe import ExpertsAllocator, BasicTokenDispatcher, MoEConfig, MoELayer, ScheduledMoELayer, parallelize_experts

    dev = device_type()
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("EP",))

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.moe = MoELayer(MoEConfig(H, I, E, K, dtype=torch.float32), None, device=dev)

        def forward(self, x):
            return self.moe(x)

    torch.manual_seed(7)
    model = Block()
    model.moe.reset_parameters(torch.Generator(device=dev).manual_seed(5))
    x = torch.randn(T, H, generator=torch.Generator().manual_seed(40 + rank)).to(dev)
    want = model(x).detach()

    class Alloc(ExpertsAllocator):
        def allocate_experts(self, layer_id, iter=-1):
            return ALLOC_A if iter <= 0 else None

    parallelize_experts(model, r"moe", mesh, experts_allocator=Alloc(E, world), token_dispatcher=BasicTokenDispatcher())
    assert isinstance(model.moe, ScheduledMoELayer) and model.moe.alloc.tp == 2 and hasattr(model, "_moe_scheduler")
    torch.testing.assert_close(model(x).detach(), want, rtol=2e-4, atol=2e-4)
    info = Alloc(E, world).allocate_experts_internal(0)
    assert info["dp_size"].tolist() == [2, 1, 1, 2] and info["tp_size"].tolist() == [2] * E


def test_parallelize_experts_with_reference_shaped_allocator():
    run_distributed(_parallelize, 4)
