"""Expert parallelism: EP=4 Mixtral-tiny must match the same model with all experts on one device
(legacy ``test/parallel/ddp_optim/test_moe.py`` / ``test/model/mixtral`` strategy)."""
import os

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


def _ep_starved_rank(rank, world):
    """A rank whose experts receive NO token must still take part in the backward of the dispatch all-to-all (its peers run theirs):
    the router is biased so that every token goes to the experts of the last rank.  Outputs and gradients equal the dense layer."""
    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.moe import MoEConfig, MoELayer

    dev = device_type()
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("EP",))
    cfg = MoEConfig(16, 32, 2 * world, 2, dtype=torch.float32)
    ep, dense = MoELayer(cfg, mesh.get_group(0), device=dev), MoELayer(MoEConfig(16, 32, 2 * world, 2, dtype=torch.float32), None, device=dev)
    for l in (ep, dense):
        l.reset_parameters(torch.Generator().manual_seed(4))
        with torch.no_grad():
            l.router.weight.zero_()
            l.router.weight[-2:, 0] = 50.0  # with x[:, 0] = 1 below, the last two experts win every token
    for it in range(2):
        xs = [torch.randn(6, 16, generator=torch.Generator().manual_seed(10 * it + r)).to(dev) for r in range(world)]
        for x in xs:
            x[:, 0] = 1.0
        x = xs[rank].clone().requires_grad_(True)
        y = ep(x)
        assert sum(ep.last_tokens_per_rank) == (6 * 2 * world if rank == world - 1 else 0)
        y.pow(2).sum().backward()
        xd = [t.clone().requires_grad_(True) for t in xs]
        sum(dense(t).pow(2).sum() for t in xd).backward()
        torch.testing.assert_close(y.detach(), dense(xs[rank]).detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(x.grad, xd[rank].grad, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ep.experts.w_down.grad, dense.experts.w_down.grad[2 * rank : 2 * rank + 2], rtol=1e-5, atol=1e-6)
        assert ep.experts.w_gate_up.grad is not None and (rank == world - 1 or float(ep.experts.w_gate_up.grad.abs().sum()) == 0.0)
        for l in (ep, dense):
            for p in l.parameters():
                p.grad = None


def test_expert_parallel_rank_without_tokens():
    run_distributed(_ep_starved_rank, 3)


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


def _hijack_hf_mixtral(rank, world):
    """``parallelize_experts`` on an UNMODIFIED HuggingFace ``MixtralSparseMoeBlock`` (legacy ``moe/_moe_tensor.py:42-99``): the
    block keeps its class, router and signature; its expert container is swapped for the EP dispatch -> grouped GEMM -> combine
    path.  Output and expert gradients equal the single-process block (each rank routes different tokens)."""
    import copy

    from common import device_type
    from transformers import MixtralConfig
    from transformers.models.mixtral import modeling_mixtral as mm

    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.moe import is_experts_parallized, parallelize_experts

    cfg = MixtralConfig(hidden_size=32, intermediate_size=64, num_local_experts=4, num_experts_per_tok=2, num_hidden_layers=1, num_attention_heads=4,
                        num_key_value_heads=2, vocab_size=64)
    torch.manual_seed(0)
    blk = mm.MixtralSparseMoeBlock(cfg)
    for p in blk.parameters():
        torch.nn.init.normal_(p, 0, 0.1)
    ref = copy.deepcopy(blk)
    mesh = init_device_mesh(device_type(), (world,), mesh_dim_names=("EP",))
    holder = torch.nn.ModuleDict({"moe": blk})
    parallelize_experts(holder, r"moe", ep_mesh=mesh)
    assert is_experts_parallized(holder) and type(holder["moe"]) is mm.MixtralSparseMoeBlock
    per = 4 // world
    lay = holder["moe"]._vb_moe_layer
    assert lay.experts.w_gate_up.shape[0] == per and all(getattr(p, "_is_expert_param", False) for p in lay.experts.parameters())
    x = torch.randn(2, 8, 32, generator=torch.Generator().manual_seed(10 + rank))
    out = holder["moe"](x.clone())
    want = ref(x.clone())
    torch.testing.assert_close(out, want, rtol=1e-4, atol=1e-5)
    out.sum().backward()
    want.sum().backward()
    for mine, full in ((lay.experts.w_gate_up.grad, ref.experts.gate_up_proj.grad), (lay.experts.w_down.grad, ref.experts.down_proj.grad)):
        g = full.clone()
        dist.all_reduce(g)  # an expert's gradient sums the tokens every rank sent to it
        torch.testing.assert_close(mine, g[rank * per : (rank + 1) * per], rtol=1e-4, atol=1e-5)


class _OldStyleExpert(torch.nn.Module):
    def __init__(self, h, f):
        super().__init__()
        self.w1, self.w2, self.w3 = torch.nn.Linear(h, f, bias=False), torch.nn.Linear(f, h, bias=False), torch.nn.Linear(h, f, bias=False)

    def forward(self, x):
        return self.w2(torch.nn.functional.silu(self.w1(x)) * self.w3(x))


class _OldStyleBlock(torch.nn.Module):
    """transformers-4.x ``MixtralSparseMoeBlock``: per-expert loop with ``index_add_`` in the block's forward."""

    def __init__(self, h=16, f=32, e=4, k=2):
        super().__init__()
        self.top_k, self.num_experts = k, e
        self.gate = torch.nn.Linear(h, e, bias=False)
        self.experts = torch.nn.ModuleList([_OldStyleExpert(h, f) for _ in range(e)])

    def forward(self, hidden_states):
        b, s, h = hidden_states.shape
        x = hidden_states.view(-1, h)
        logits = self.gate(x)
        w = torch.softmax(logits, dim=1, dtype=torch.float)
        w, sel = torch.topk(w, self.top_k, dim=-1)
        w = (w / w.sum(-1, keepdim=True)).to(x.dtype)
        final = torch.zeros_like(x)
        mask = torch.nn.functional.one_hot(sel, self.num_experts).permute(2, 1, 0)
        for e in range(self.num_experts):
            idx, top_x = torch.where(mask[e])
            final.index_add_(0, top_x, self.experts[e](x[top_x]) * w[top_x, idx, None])
        return final.view(b, s, h), logits


def _hijack_modulelist(rank, world):
    import copy

    from common import device_type

    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.moe import parallelize_experts

    torch.manual_seed(1)
    blk = _OldStyleBlock()
    ref = copy.deepcopy(blk)
    mesh = init_device_mesh(device_type(), (world,), mesh_dim_names=("EP",))
    holder = torch.nn.ModuleDict({"block_sparse_moe": blk})
    parallelize_experts(holder, r"block_sparse_moe", ep_mesh=mesh)
    assert len(holder["block_sparse_moe"].experts) == 0  # remote experts' weights were released
    x = torch.randn(2, 6, 16, generator=torch.Generator().manual_seed(20 + rank))
    out, logits = holder["block_sparse_moe"](x.clone())
    want, want_logits = ref(x.clone())
    torch.testing.assert_close(out, want, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(logits, want_logits)


def test_hijack_unmodified_hf_mixtral_block():
    run_distributed(_hijack_hf_mixtral, 2)


def test_hijack_modulelist_style_block():
    run_distributed(_hijack_modulelist, 2)


def _param_buffer(rank, world):
    """``MoELayerParamBuffer`` on an EP=2 x DP=2 mesh: expert buffers sharded over DP (FSDP units sharing one state), gathered a
    layer ahead, gradients reduce-scattered from the backward hook; ``refresh_buffer`` moves experts with master weights and
    AdamW moments.  Golden = the same layers with un-sharded experts on the EP group only."""
    import copy

    from common import device_type
    from vescale_b200 import init_device_mesh
    from vescale_b200.optim import FSDPAdamW
    from vescale_b200.parallel.fsdp import MixedPrecisionPolicy
    from vescale_b200.parallel.moe import MoEConfig, MoELayer, MoELayerParamBuffer

    dev = device_type()
    mesh = init_device_mesh(dev, (2, 2), mesh_dim_names=("DP", "EP"))
    ep_group, dp = mesh.get_group("EP"), mesh.get_local_rank("DP")
    cfg = MoEConfig(16, 32, 4, 2, dtype=torch.float32)
    def build():
        g = torch.Generator().manual_seed(3)
        ls = torch.nn.ModuleList([MoELayer(cfg, ep_group, device=dev) for _ in range(2)])
        for l in ls:
            l.reset_parameters(g)
        return ls

    layers, golden = build(), build()
    buf = MoELayerParamBuffer(layers, mesh, mesh_dim="DP", mp_policy=MixedPrecisionPolicy(param_dtype=torch.float32, reduce_dtype=torch.float32), prefetch=1)
    assert len(buf.units) == 2 and buf.units[0]._state is buf.units[1]._state
    assert all(u.S * 2 >= sum(p.numel() for p in gl.experts.parameters()) for u, gl in zip(buf.units, golden))

    def run(ls, x):
        h = x
        for l in ls:
            h = h + l(h)
        return h.pow(2).mean()

    x = torch.randn(12, 16, generator=torch.Generator().manual_seed(50 + rank)).to(dev)
    loss = run(layers, x)
    ref = run(golden, x)
    torch.testing.assert_close(loss, ref, rtol=1e-5, atol=1e-6)
    loss.backward()
    ref.backward()
    buf.state.wait_grads()
    for u, gl in zip(buf.units, golden):
        flat = torch.cat([gl.experts.w_gate_up.grad.reshape(-1), gl.experts.w_down.grad.reshape(-1)])
        dist.all_reduce(flat, group=mesh.get_group("DP"))
        flat /= 2  # the reduce-scatter averages over the DP replicas
        got = torch.zeros(u.S * 2)
        mine = u.grad_shard.float() * getattr(u, "grad_scale_pending", 1.0)
        for slot in u.layout.slots:
            lo, hi = u.layout.rank_range(slot, u.rank)
            if hi > lo:
                got[u.rank * u.S + lo : u.rank * u.S + hi] = mine[lo:hi]
        want = torch.zeros(u.S * 2)
        off = 0
        for slot, p in zip(u.layout.slots, (gl.experts.w_gate_up.grad, gl.experts.w_down.grad)):
            want[slot.offset : slot.end] = flat[off : off + p.numel()]
            off += p.numel()
        sl = slice(u.rank * u.S, (u.rank + 1) * u.S)
        if os.environ.get("DBG"):
            print(rank, "got", got[sl][:4], got[sl].norm(), "want", want[sl][:4], want[sl].norm(), "ratio", (got[sl] / want[sl])[:6], flush=True)
        torch.testing.assert_close(got[sl], want[sl], rtol=1e-4, atol=1e-6)
    # ---- re-allocation: swap the homes of experts 0 and 3; forward is unchanged, moments travel with their expert
    holder = torch.nn.Module()
    holder.layers = layers
    opt = FSDPAdamW(holder, lr=1e-2, max_grad_norm=None)  # finds the buffer's FSDP state through the expert units
    u0 = buf.units[0]
    El = 2
    for slot in u0.layout.slots:  # tag the first moment of every element with the id of the expert it belongs to
        lo, hi = u0.layout.rank_range(slot, u0.rank)
        per = slot.end - slot.offset
        per //= El
        idx = torch.arange(u0.rank * u0.S + lo, u0.rank * u0.S + hi)
        local_slot = (idx - slot.offset) // per
        ep = mesh.get_local_rank("EP")
        u0.exp_avg[lo:hi] = (ep * El + local_slot).float()
    before = run(layers, x).detach()
    buf.refresh_buffer(0, [3, 1, 2, 0], optimizer=opt)
    after = run(layers, x).detach()
    torch.testing.assert_close(after, before, rtol=1e-5, atol=1e-6)
    ep = mesh.get_local_rank("EP")
    owner = {0: 3, 1: 1, 2: 2, 3: 0}  # slot -> expert now living there
    for slot in u0.layout.slots:
        lo, hi = u0.layout.rank_range(slot, u0.rank)
        per = (slot.end - slot.offset) // El
        idx = torch.arange(u0.rank * u0.S + lo, u0.rank * u0.S + hi)
        local_slot = (idx - slot.offset) // per
        want = torch.tensor([float(owner[ep * El + int(s)]) for s in local_slot])
        torch.testing.assert_close(u0.exp_avg[lo:hi].cpu(), want)


def test_moe_layer_param_buffer():
    run_distributed(_param_buffer, 4)
