"""DModule (TP/SP by plan) and deferred init on 4 ranks — golden = same module on one device
(legacy ``test/dmodule/test_fwd_plan.py``, ``test_initialize.py`` strategy)."""
import copy

import torch
import torch.nn as nn

from common import device_type, run_distributed


class Block(nn.Module):
    def __init__(self, h=32, f=64):
        super().__init__()
        self.ln = nn.LayerNorm(h)
        self.fc1 = nn.Linear(h, f)
        self.fc2 = nn.Linear(f, h)

    def forward(self, x):
        return x + self.fc2(torch.nn.functional.gelu(self.fc1(self.ln(x))))


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.blocks = nn.ModuleList([Block(), Block()])

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x


def _tp_sp(rank, world):
    from vescale_b200 import Replicate, Shard, init_device_mesh
    from vescale_b200.dtensor import DTensor
    from vescale_b200.dtensor.debug import CommDebugMode
    from vescale_b200.parallel.dmodule import parallelize_module

    torch.manual_seed(0)
    dev = device_type()
    ref = Net().to(dev)
    model = copy.deepcopy(ref)
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("TP",))
    plan = {
        "parameter": {
            r"blocks\.\d+\.fc1\.weight": [Shard(0)],
            r"blocks\.\d+\.fc1\.bias": [Shard(0)],
            r"blocks\.\d+\.fc2\.weight": [Shard(1)],
        },
        "forward": {
            r"input": [[Replicate()]],  # the caller passes the full (replicated) batch as a plain tensor
            r"blocks\.\d+\.input": [[Shard(1)]],  # sequence-parallel residual stream
            r"blocks\.\d+\.fc1\.input": [[Replicate()]],  # all-gather before the column-parallel GEMM
            r"blocks\.\d+\.fc2\.output": [[Shard(1)]],  # reduce-scatter after the row-parallel GEMM
            r"blocks\.1\.output": [[Replicate()]],
        },
    }
    parallelize_module(model, mesh, plan)
    assert all(isinstance(p, DTensor) for p in model.parameters())
    x = torch.randn(2, 8, 32, generator=torch.Generator().manual_seed(1)).to(dev)
    with CommDebugMode(model) as cm:
        out = model(x)
    counts = cm.get_comm_counts()
    assert counts.get("all_gather", 0) >= 2 and counts.get("reduce_scatter", 0) + counts.get("all_reduce", 0) >= 2, counts
    assert out.placements == (Replicate(),)
    torch.testing.assert_close(out.to_local(), ref(x), rtol=1e-4, atol=1e-5)
    out.to_local().sum().backward()
    ref(x).sum().backward()
    # LayerNorm grads are Partial under SP until synced
    n_partial = len(model.list_partial_grads())
    assert n_partial >= 4, n_partial
    model.finish_grad_sync()
    assert len(model.list_partial_grads()) == 0
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        torch.testing.assert_close(p.grad.full_tensor(), q.grad, rtol=1e-4, atol=1e-5, msg=n)


    # ---- factory mode: tensors created inside forward become DTensors; modules may return dicts / tuples of DTensors
    # (legacy ``dmodule/test_dfactory.py``, ``test_obj_return.py``)
    class WithFactory(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(32, 32)

        def forward(self, x):
            y = self.fc(x)
            bias = torch.ones(y.shape[-1], device=y.device if not isinstance(y, DTensor) else None)  # created inside forward
            return {"out": y + bias, "aux": (y * 2, 3)}

    torch.manual_seed(1)
    ref2 = WithFactory().to(dev)
    m2 = copy.deepcopy(ref2)
    parallelize_module(
        m2, mesh,
        {"parameter": {r"fc\.weight": [Shard(0)], r"fc\.bias": [Shard(0)]}, "forward": {r"input": [[Replicate()]], r"fc\.output": [[Replicate()]]}},
        factory=True,
    )
    res = m2(x)
    want = ref2(x)
    assert isinstance(res, dict) and isinstance(res["aux"], tuple) and res["aux"][1] == 3
    got_out = res["out"].full_tensor() if isinstance(res["out"], DTensor) else res["out"]
    got_aux = res["aux"][0].full_tensor() if isinstance(res["aux"][0], DTensor) else res["aux"][0]
    torch.testing.assert_close(got_out, want["out"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got_aux, want["aux"][0], rtol=1e-4, atol=1e-5)


def _deferred(rank, world):
    import vescale_b200.dtensor as vd
    from vescale_b200 import Replicate, Shard, init_device_mesh
    from vescale_b200.dtensor import RaggedShard
    from vescale_b200.initialize import deferred_init, materialize_dtensor, materialize_module

    dev = device_type()
    mesh = init_device_mesh(dev, (world,))
    vd.manual_seed(7)
    golden = materialize_module(deferred_init(Net), dev)
    vd.manual_seed(7)
    m = deferred_init(Net)
    assert all(p.is_meta for p in m.parameters())
    placements = {2: [Shard(0)], 1: [RaggedShard((0,), (1, 0, 2, 1))]}
    for (n, p), (_, g) in zip(m.named_parameters(), golden.named_parameters()):
        pl = placements.get(p.ndim, [Replicate()]) if p.shape[0] % 4 == 0 else [Replicate()]
        dt = materialize_dtensor(p, mesh, pl)
        assert dt.to_local().numel() <= p.numel()
        assert torch.equal(dt.full_tensor(), g.detach()), n  # only the shard was materialised, values match the single-device init


def test_tp_sp_plan():
    run_distributed(_tp_sp, 4)


def test_deferred_init_sharded_equals_single_device():
    run_distributed(_deferred, 4)


class _Head(nn.Module):
    def __init__(self, V=48, h=32):
        super().__init__()
        self.emb = nn.Embedding(V, h)
        self.proj = nn.Linear(h, h)  # row parallel, with bias
        self.lm = nn.Linear(h, V, bias=False)
        self.loss = nn.CrossEntropyLoss(ignore_index=-100)

    def forward(self, ids, labels):
        return self.loss(self.lm(self.proj(self.emb(ids))), labels)


def _model_patches(rank, world):
    """RowParallelLinear / VocabParallelEmbedding / VocabParallelCrossEntropy patches (legacy ``model/patch``): a vocab-parallel
    embedding, a row-parallel linear with bias and a vocab-parallel LM head + CrossEntropyLoss reproduce the single-device
    loss and gradients; the patched loss never gathers the logits."""
    from vescale_b200 import Replicate, Shard, init_device_mesh
    from vescale_b200.dtensor import DTensor
    from vescale_b200.dtensor.debug import CommDebugMode
    from vescale_b200.model.patch import VocabParallelCrossEntropy, get_all_model_patch
    from vescale_b200.parallel.dmodule import parallelize_module

    dev = device_type()
    mesh = init_device_mesh(dev, (world,))
    torch.manual_seed(0)
    ref = _Head().to(dev)
    m = copy.deepcopy(ref)
    parallelize_module(
        m, mesh,
        {"parameter": {r"emb\.weight": [Shard(0)], r"proj\.weight": [Shard(1)], r"proj\.bias": [Replicate()], r"lm\.weight": [Shard(0)]},
         "forward": {r"proj\.input": [[Shard(1)]], r"lm\.input": [[Replicate()]]}},
    )
    for patch in get_all_model_patch():
        patch(m)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 48, (24,), generator=g).to(dev)
    labels = torch.randint(0, 48, (24,), generator=g)
    labels[::5] = -100
    labels = labels.to(dev)
    with CommDebugMode() as comm:
        loss = m(ids, labels)
    want = ref(ids, labels)
    got = loss.full_tensor() if isinstance(loss, DTensor) else loss
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    counts = comm.get_comm_counts()
    assert not any("all_gather" in str(k) and v for k, v in counts.items() if "logits" in str(k)), counts
    got.backward() if not isinstance(loss, DTensor) else loss.backward()
    want.backward()
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        gp = p.grad.full_tensor() if isinstance(p.grad, DTensor) else p.grad
        torch.testing.assert_close(gp, q.grad, rtol=1e-3, atol=1e-5, msg=lambda s, n=n: f"{n}: {s}")
    # reductions of the patched loss module on class-sharded logits
    logits = torch.randn(24, 48, generator=g).to(dev)
    dl = DTensor.from_local(logits.chunk(world, 1)[rank].contiguous(), mesh, [Shard(1)])
    for red in ("mean", "sum", "none"):
        lossmod = nn.CrossEntropyLoss(ignore_index=-100, reduction=red)
        holder = nn.ModuleDict({"l": lossmod})
        VocabParallelCrossEntropy.patch(holder)
        torch.testing.assert_close(lossmod(dl, labels), torch.nn.functional.cross_entropy(logits, labels, ignore_index=-100, reduction=red), rtol=1e-4, atol=1e-5)


def test_model_patches():
    run_distributed(_model_patches, 4)


# ---------------------------------------------------------------------------------------------------------------------------
# forward plans bound to the signature, factory regions per module class (legacy ``test_fwd_plan.py``, ``test_dfactory.py``)
class _KwOnly(nn.Module):
    def forward(self, a, *rest, b=None, **extra):
        return a, rest, b, extra


class _InnerZ(nn.Module):
    def forward(self, x):
        return torch.zeros(x.shape, dtype=x.dtype)


class _OuterZ(nn.Module):
    def __init__(self):
        super().__init__()
        self.m = _InnerZ()

    def forward(self, x):
        return torch.zeros(x.shape), self.m(x), torch.ones(x.shape)


def _fwd_plan_and_factory(rank, world):
    import warnings

    from vescale_b200 import Replicate, Shard, init_device_mesh
    from vescale_b200 import dtensor as D
    from vescale_b200.dtensor import DTensor
    from vescale_b200.parallel.dmodule import _factory, parallelize_module

    mesh = init_device_mesh(device_type(), (world,))
    a, b = torch.ones(2, 2), torch.ones(2, 2) * 2

    # sequence plan runs over positional then keyword arguments; dict plan goes by name, *args takes a list, **kwargs is searched
    m = parallelize_module(_KwOnly(), mesh, {"forward": {".input": [[Shard(0)], None, [Replicate()]]}})
    o_a, o_rest, o_b, _ = m(a, a, b=b)
    assert isinstance(o_a, DTensor) and o_a.placements[0] == Shard(0) and tuple(o_a.shape) == (2 * world, 2)
    assert type(o_rest[0]) is torch.Tensor and isinstance(o_b, DTensor) and o_b.placements[0].is_replicate()
    m = parallelize_module(_KwOnly(), mesh, {"forward": {".input": {"a": [Shard(1)], "rest": [[Shard(0)]], "z": [Replicate()]}}})
    o_a, o_rest, o_b, o_extra = m(a, b, z=b)
    assert o_a.placements[0] == Shard(1) and o_rest[0].placements[0] == Shard(0) and o_b is None and isinstance(o_extra["z"], DTensor)
    # a wrong call fails as the bare module would; surplus / unknown plan entries warn and are ignored
    try:
        m(b=b)
        raise AssertionError("expected TypeError")
    except TypeError:
        pass
    for plan in ({".input": [[Shard(0)], None, None, None]}, {".input": {"a": [Shard(0)], "nope": None}}):
        m = parallelize_module(_KwOnly(), mesh, {"forward": plan})
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            m(a)
        assert any(issubclass(x.category, UserWarning) for x in w), plan

    # factory regions: references to the builtins taken outside still build DTensors; Off inside On inside Off
    f = torch.zeros
    adp = _factory._provide_args(mesh, {torch.zeros: [Shard(0)], "full": [Replicate()]})
    with _factory.FactoryDispatchModeOn(mesh, adp):
        z = f((2 * world, 3))
        assert isinstance(z, DTensor) and z.placements[0] == Shard(0) and tuple(z.to_local().shape) == (2, 3)
        assert D.equal(z, D.zeros((2 * world, 3), device_mesh=mesh, placements=[Shard(0)]))
        assert torch.full((3,), 2.0).placements[0].is_replicate() and torch.ones(3).placements[0].is_replicate()
        r = torch.arange(0, 2 * world, 1, dtype=torch.float32)
        assert isinstance(r, DTensor) and r.full_tensor().tolist() == list(range(2 * world))
        with _factory.FactoryDispatchModeOff():
            assert type(torch.zeros(3)) is torch.Tensor
            with _factory.FactoryDispatchModeOn(mesh, {}):
                assert isinstance(torch.empty(3), DTensor)
            assert type(torch.zeros(3)) is torch.Tensor
        assert isinstance(torch.zeros(3), DTensor)
    assert type(torch.zeros(3)) is torch.Tensor
    # per module class: {cls: True | False | {factory: placements}}
    x = torch.ones(4 * world)
    o = parallelize_module(_OuterZ(), mesh, {}, factory={_OuterZ: True, _InnerZ: False})(x)
    assert [isinstance(t, DTensor) for t in o] == [True, False, True]
    o = parallelize_module(_OuterZ(), mesh, {}, factory={_OuterZ: False, _InnerZ: {torch.zeros: [Shard(0)]}})(x)
    assert [isinstance(t, DTensor) for t in o] == [False, True, False] and o[1].placements[0] == Shard(0) and o[1].to_local().numel() == 4
    o = parallelize_module(_OuterZ(), mesh, {"forward": {".input": [[Shard(0)]]}}, factory=True)(torch.ones(4))
    assert tuple(o[0].shape) == (4 * world,) and isinstance(o[1], DTensor)  # inner module inherits the enclosing region


def test_forward_plan_binding_and_factory_regions():
    run_distributed(_fwd_plan_and_factory, 2)


class _TwoBranch(nn.Module):
    def __init__(self, h=16):
        super().__init__()
        self.m1 = nn.Sequential(nn.Linear(h, 4 * h, bias=False), nn.Linear(4 * h, h, bias=False))
        self.m2 = nn.Sequential(nn.Linear(h, 4 * h, bias=False), nn.Linear(4 * h, h, bias=False))

    def forward(self, x):
        return x + (self.m1(x) + self.m2(x))


def _deferred_output_reshard(rank, world):
    """``PlacementsInterface(defer_reshard=True)`` on two row-parallel outputs: they stay Partial, their sum is all-reduced once
    (legacy ``test/dtensor/general/test_defer_resharding.py``)."""
    from vescale_b200 import Replicate, Shard, init_device_mesh
    from vescale_b200.dtensor import DTensor
    from vescale_b200.dtensor.debug import CommDebugMode
    from vescale_b200.parallel.dmodule import PlacementsInterface as PI
    from vescale_b200.parallel.dmodule import parallelize_module

    torch.manual_seed(0)
    mesh = init_device_mesh(device_type(), (world,))
    m = _TwoBranch().to(device_type())
    golden = copy.deepcopy(m)
    x = torch.randn(4, 16, 16, device=device_type())
    plan = {
        "parameter": {r"m\d\.0\.weight": [Shard(0)], r"m\d\.1\.weight": [Shard(1)]},
        "forward": {r"m\d\.input": [[Replicate()]], r"m\d\.output": [PI([Replicate()], defer_reshard=True)]},
    }
    dm = parallelize_module(m, mesh, plan)
    seen = {}
    dm.m1.register_forward_hook(lambda mod, a, o: seen.update(m1=o.placements[0]))
    with CommDebugMode() as comm:
        out = dm(DTensor.from_local(x, mesh, [Replicate()]))
    assert seen["m1"].is_partial() and out.placements[0].is_replicate()
    assert comm.get_total_counts() == 1, comm.get_total_counts()
    ref = golden(x)
    assert torch.allclose(out.to_local(), ref, atol=1e-5)
    out.to_local().sum().backward()
    ref.sum().backward()
    assert torch.allclose(dm.m1[0].weight.grad.full_tensor(), golden.m1[0].weight.grad, atol=1e-4)


def test_deferred_output_reshard_single_allreduce():
    run_distributed(_deferred_output_reshard, 2)


def _mixtral_plan_matches_single_device(rank, world):
    """The sparse-MoE Mixtral of ``examples/mixtral_4D_benchmark`` under its TP+SP sharding plan (data-dependent routing on
    replicated router outputs, row-parallel experts) reproduces the unparallelised model: output and gradients."""
    import argparse
    import importlib.util
    import os

    from vescale_b200 import init_device_mesh
    from vescale_b200.parallel.dmodule import parallelize_module

    path = os.path.join(os.path.dirname(__file__), "..", "examples", "mixtral_4D_benchmark", "run.py")
    spec = importlib.util.spec_from_file_location("mixtral_4d_example", path)
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    a = argparse.Namespace(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, num_experts=4, top_k=2)
    torch.manual_seed(0)
    dev = device_type()
    model = ex.Mixtral(a).to(dev)
    golden = copy.deepcopy(model)
    ids = torch.randint(0, a.vocab_size, (2, 8), device=dev)
    probe = torch.randn(2, 8, a.hidden_size, device=dev)
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("TP",))
    parallelize_module(model, mesh, ex.mixtral_plan)
    out = model(ids)
    ref = golden(ids)
    assert out.placements[0].is_replicate() and torch.allclose(out.to_local(), ref, atol=1e-4), (out.to_local() - ref).abs().max()
    (out.to_local() * probe).sum().backward()
    model.finish_grad_sync()
    (ref * probe).sum().backward()
    for name in ("layers.0.block_sparse_moe.experts.1.w2.weight", "layers.1.self_attn.q_proj.weight", "layers.0.block_sparse_moe.gate.weight", "embed_tokens.weight", "layers.1.post_attention_layernorm.weight"):
        g, gg = dict(model.named_parameters())[name].grad, dict(golden.named_parameters())[name].grad
        assert torch.allclose(g.full_tensor(), gg, atol=1e-3), (name, (g.full_tensor() - gg).abs().max())


def test_mixtral_4d_plan_matches_single_device():
    run_distributed(_mixtral_plan_matches_single_device, 2)


class _FFNBlock(nn.Module):
    def __init__(self, d=16):
        super().__init__()
        self.up, self.gate, self.down = nn.Linear(d, 4 * d), nn.Linear(d, 4 * d), nn.Linear(4 * d, d)

    def forward(self, x):
        return self.down(torch.nn.functional.silu(self.gate(x)) * self.up(x))


class _TwoBlocks(nn.Module):
    def __init__(self):
        super().__init__()
        self.a, self.b = _FFNBlock(), _FFNBlock()

    def forward(self, x):
        return self.b(self.a(x))


def _dmp_pinned(rank, world):
    """Policy plans from a class-level provider, overridden for one subtree by a pinned plan (legacy ``dmp.py:37-184``)."""
    import copy

    from vescale_b200 import Replicate, Shard, distribute_tensor, init_device_mesh
    from vescale_b200.dtensor.api import DTensor
    from vescale_b200.parallel.dmp import PlanGenerator, auto_parallelize_module, get_plan_overriding_policy, set_plan_overriding_policy
    from vescale_b200.parallel.dmp.policies import REGISTRY
    from vescale_b200.parallel.dmp.policies.megatron import mlp_plan_provider
    from vescale_b200.parallel.dmp.policies.utils import validate_single_input

    if not REGISTRY.has_module("_FFNBLOCK"):
        REGISTRY.provide_register_for_policy("MEGATRON")(["_FFNBlock"])(mlp_plan_provider)
    dev = device_type()
    mesh = init_device_mesh(dev, (world,), mesh_dim_names=("TP",))
    torch.manual_seed(0)
    ref = _TwoBlocks().to(dev)
    model = copy.deepcopy(ref)
    assert validate_single_input(model.a) == "x"
    # pin block b: keep it entirely replicated (e.g. too small to be worth sharding); the policy still shards block a
    set_plan_overriding_policy(model.b, param_sharding_plan={r".*": [Replicate()]}, fwd_resharding_plan={"input": [[Replicate()]], "output": [[Replicate()]]})
    assert get_plan_overriding_policy(model.b)[0] is not None and get_plan_overriding_policy(model.a) == (None, None)
    try:
        set_plan_overriding_policy(model, {})
        raise AssertionError("pinning above a pinned subtree must be rejected")
    except NotImplementedError:
        pass
    p, f, pin_p, pin_f, pol_p, pol_f = PlanGenerator(model, "megatron").generate()
    assert any(k.startswith("b") for k in pol_p) and not any(k.startswith("b.") and v != [Replicate()] for k, v in p.items())
    assert p["a.up.weight"] == [Shard(0)] and p["a.down.weight"] == [Shard(1)] and p["b..*"] == [Replicate()]
    assert f["a.input"] == [[Replicate()]] and f["b.output"] == [[Replicate()]] and "b.input" in pin_f
    assert auto_parallelize_module(copy.deepcopy(ref), None, "MEGATRON", plan_only=True)[2]["parameter"]["a.up.weight"] == [Shard(0)]
    saved = {}
    auto_parallelize_module(model, mesh, "MEGATRON", plan_to_save=saved)
    assert saved["param_sharding_plan"] == p
    assert model.a.up.weight.placements == (Shard(0),) and model.a.down.weight.placements == (Shard(1),) and model.b.up.weight.placements == (Replicate(),)
    x = torch.randn(2, 8, 16).to(dev)
    out = model(distribute_tensor(x, mesh, [Replicate()]))
    out = out.full_tensor() if isinstance(out, DTensor) else out
    torch.testing.assert_close(out, ref(x), rtol=1e-4, atol=1e-5)
    try:
        PlanGenerator(model, "no_such_policy")
        raise AssertionError
    except ValueError:
        pass


def test_dmp_pinned_plans_and_class_level_providers():
    run_distributed(_dmp_pinned, 2)


def _w_weight_forward_plan(rank, world):
    """A forward plan on a PARAMETER: fc1 is stored column-sharded (ZeRO-3 style), computes with a replicated weight, and the
    gradient comes back in the parameter's own layout; ``grad=`` re-labels the gradient (legacy ``_hook.py:178-273``)."""
    from vescale_b200 import Replicate, Shard, init_device_mesh
    from vescale_b200.dtensor.api import DTensor
    from vescale_b200.parallel.dmodule import parallelize_module
    from vescale_b200.parallel.dmodule.api import PlacementsInterface as PI
    from vescale_b200.parallel.dmodule._hook import PostHookGrad, get_sig

    torch.manual_seed(0)
    mesh = init_device_mesh(device_type(), (world,))
    golden = Block().to(device_type())
    model = copy.deepcopy(golden)
    plan = {
        "parameter": {r"fc1\.weight": [Shard(0)], r"fc1\.bias": [Shard(0)]},
        "forward": {
            r"fc1\.weight": [Replicate()],
            r"fc1\.bias": PI([Replicate()]),
            r"input": [[Replicate()]],
            r"output": [[Replicate()]],
        },
    }
    model = parallelize_module(model, mesh, plan)
    w = model.fc1.weight
    assert isinstance(w, nn.Parameter) and w.placements == (Shard(0),)
    x = torch.randn(4, 8, 32, device=device_type())
    seen = {}
    model.fc1.register_forward_pre_hook(lambda m, a: seen.update(in_fwd=(m.weight.placements, m.bias.placements)))
    out = model(x)
    assert seen["in_fwd"] == ((Replicate(),), (Replicate(),))
    assert model.fc1.weight is w and model.fc1._parameters["weight"] is w  # the parameter is back in its slot
    ref = golden(x)
    assert torch.allclose(out.to_local(), ref, atol=1e-5)
    out.to_local().sum().backward()
    ref.sum().backward()
    assert w.grad.placements == (Shard(0),)
    assert torch.allclose(w.grad.full_tensor(), golden.fc1.weight.grad, atol=1e-5)
    assert torch.allclose(model.fc1.bias.grad.full_tensor(), golden.fc1.bias.grad, atol=1e-5)
    # a second step goes through the same hooks
    out2 = model(x)
    assert torch.allclose(out2.to_local(), ref, atol=1e-5) and model.fc1._parameters["weight"] is w
    # grad re-labelling
    from vescale_b200 import Partial

    g = DTensor.from_local(torch.ones(4, 4), mesh, [Partial()])
    relabelled = PostHookGrad.get_hook(mesh, [Replicate()])(g)
    assert relabelled.placements == (Replicate(),) and relabelled.to_local().data_ptr() == g.to_local().data_ptr()
    try:
        PostHookGrad.get_hook(mesh, [Shard(0)])(g)  # would claim a (4, 4) local piece of a (4, 4) tensor split two ways
        raise AssertionError("expected ValueError")
    except ValueError:
        pass
    assert list(get_sig(model).parameters) == ["x"]

    # through a plan: fc2's weight gradient is declared Replicate although the matmul with a sharded activation leaves it Partial
    m2 = copy.deepcopy(golden)
    m2 = parallelize_module(
        m2, mesh, {"parameter": {}, "forward": {r"fc2\.weight": PI(None, grad=[Replicate()]), r"input": [[Shard(0)]], r"output": [[Shard(0)]]}}
    )
    m2(x).to_local().sum().backward()
    assert m2.fc2.weight.grad.placements == (Replicate(),)
    assert any(p.is_partial() for p in m2.fc1.weight.grad.placements)  # the un-planned sibling keeps what autograd produced


def test_weight_forward_plan_and_grad_placements():
    run_distributed(_w_weight_forward_plan, 2)
