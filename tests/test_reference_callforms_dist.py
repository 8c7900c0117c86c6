"""Reference call forms kept for drop-in use: ``vescale.dtensor._collective_utils`` (legacy ``dtensor/_collective_utils.py``) and
``VESCALE_DEVICE_MESH`` accessors (legacy ``devicemesh_api/api.py``) on a 2 x 2 mesh."""
import torch

from common import device_type, run_distributed


def _collective_utils(rank, world):
    torch.set_default_device(device_type())
    import vescale
    from vescale.dtensor._collective_utils import mesh_all_gather, mesh_all_reduce, mesh_reduce_scatter, mesh_all_to_all, mesh_all_to_all_single, mesh_broadcast, mesh_scatter, broadcast_across_mesh
    from vescale.dtensor.device_mesh import DeviceMesh
    import torch.distributed.distributed_c10d as c10d
    mesh = DeviceMesh(device_type(), torch.arange(4).view(2, 2))
    x = torch.full((3, 2), float(rank))
    assert mesh_all_reduce(x, mesh, c10d.ReduceOp.SUM, 1)[0, 0].item() == [1, 1, 5, 5][rank]
    assert mesh_all_reduce(x, mesh, c10d.ReduceOp.MAX, 0)[0, 0].item() == [2, 3, 2, 3][rank]
    g = mesh_all_gather(x, (3, 4), mesh, 1, 1); assert g.shape == (3, 4) and g[0].tolist() == ([0, 0, 1, 1] if rank < 2 else [2, 2, 3, 3])
    # uneven: global 3 rows over 2 ranks -> 2 + 1
    c = mesh.get_coordinate()[1]
    loc = torch.arange(6.).view(3, 2)[[slice(0, 2), slice(2, 3)][c]]
    assert torch.equal(mesh_all_gather(loc, (3, 2), mesh, 0, 1), torch.arange(6.).view(3, 2))
    rs = mesh_reduce_scatter(torch.ones(4, 2) * (rank + 1), mesh, c10d.ReduceOp.SUM, 0, 1); assert rs.shape == (2, 2) and rs[0, 0].item() == (3 if rank < 2 else 7)
    outs = [torch.zeros(2), torch.zeros(2)]
    mesh_all_to_all(outs, [torch.full((2,), 10. * rank), torch.full((2,), 10. * rank + 1)], mesh, 1)
    peer = rank ^ 1
    assert outs[c][0].item() == 10. * rank + c and outs[1 - c][0].item() == 10. * peer + c
    t = mesh_all_to_all_single(torch.arange(8.).view(2, 4) + 100 * c, mesh, 0, 1, 1); assert t.shape == (4, 2)
    b = mesh_broadcast(torch.full((2,), float(rank)), mesh, 1); assert b[0].item() == (0 if rank < 2 else 2)
    o = torch.zeros(2); mesh_scatter(o, [torch.full((2,), 5.), torch.full((2,), 6.)] if c == 0 else None, mesh, 1); assert o[0].item() == 5 + c
    r = broadcast_across_mesh(torch.arange(3.) if rank == 2 else None, 2, (3,), torch.float32, mesh); assert r.tolist() == [0, 1, 2]

    # VeDeviceMesh: lazy get(**config), strategy size by index or name, coordinates of other ranks
    from vescale.devicemesh_api import VESCALE_DEVICE_MESH as V

    V._mesh = None
    m = V.get(device_type=device_type(), mesh_shape=(2, 2), mesh_dim_names=("DP", "TP"))
    assert m.mesh.tolist() == [[0, 1], [2, 3]] and V.get() is m
    assert V.get_strategy_size(0) == V.get_strategy_size("DP") == 2 and V.get_strategy_size("PP") == 1
    assert V.get_strategy_coordinate(local_rank=3) == [1, 1] and V.get_strategy_coordinate() == [rank // 2, rank % 2]
    assert V.lookup_rank("TP") == rank % 2 and "PP" not in V._MESH_DIM_NAMES_LOOKUP
    assert [x.mesh.tolist() for x in V.get_global_pipeline_parallel_meshes(device_type())] == [[0, 1], [2, 3]]


def test_reference_collective_utils_and_vedevicemesh_call_forms():
    run_distributed(_collective_utils, 4)


def _interleaved_through_matmul(rank, world):
    """A sequence-sharded activation stays sharded through flatten -> mm / addmm -> unflatten: ``Shard(1)`` becomes
    ``InterleavedShard(0, B)`` and back, with no communication (legacy ``test/dtensor/shard/test_interleaved_shard.py``)."""
    from vescale_b200 import InterleavedShard, Replicate, Shard, distribute_tensor, init_device_mesh
    from vescale_b200.dtensor import from_local
    from vescale_b200.dtensor.debug import CommDebugMode

    torch.set_default_device(device_type())
    mesh = init_device_mesh(device_type(), (world,))
    torch.manual_seed(0)
    lhs, rhs, bias = torch.rand(4, 3 * world, 5), torch.rand(5, 7), torch.rand(7)
    d, r, b = distribute_tensor(lhs, mesh, [Shard(1)]), distribute_tensor(rhs, mesh, [Replicate()]), distribute_tensor(bias, mesh, [Replicate()])
    with CommDebugMode() as comm:
        x = d.reshape((-1, 5))
        o = torch.mm(x, r)
        o2 = o.reshape((4, -1, 7))
        o3 = torch.addmm(b, x, r).reshape((4, -1, 7))
    assert x.placements[0] == InterleavedShard(0, 4) and o.placements[0] == InterleavedShard(0, 4), (x.placements, o.placements)
    assert o2.placements[0] == Shard(1) and o3.placements[0] == Shard(1) and comm.get_total_counts() == 0
    assert torch.allclose(o2.full_tensor(), lhs @ rhs, atol=1e-6) and torch.allclose(o3.full_tensor(), lhs @ rhs + bias, atol=1e-6)
    assert torch.allclose(x.full_tensor(), lhs.reshape(-1, 5))
    try:  # a local shard that cannot hold whole sections is rejected
        from_local(torch.rand(3, 3), mesh, [InterleavedShard(0, 4)])
        raise AssertionError("expected ValueError")
    except ValueError:
        pass


def test_interleaved_shard_through_flatten_matmul_unflatten():
    run_distributed(_interleaved_through_matmul, 2)


def _custom_rules(rank, world):
    """A user's custom op gets its sharding behaviour through the reference's two registration contracts."""
    import torch
    from vescale_b200 import DTensor, Replicate, Shard, distribute_tensor, init_device_mesh
    from vescale_b200.dtensor.op_schema import OpSchema, OpStrategy, OutputSharding, PlacementStrategy
    from vescale_b200.dtensor.ops.basic_strategy import gen_einsum_strategies
    from vescale_b200.dtensor.ops.common_rules import einop_rule, pointwise_rule
    from vescale_b200.dtensor.ops.utils import generate_redistribute_costs, is_tensor_partial, register_op_strategy, register_prop_rule
    from vescale_b200.placement import Partial
    from vescale_b200.spec import DTensorSpec

    lib = torch.library.Library("vb_test", "FRAGMENT")  # noqa: TOR901
    lib.define("scaled_mm(Tensor a, Tensor b, float s) -> Tensor")
    lib.define("shift(Tensor a, Tensor b) -> Tensor")
    lib.impl("scaled_mm", lambda a, b, s: (a @ b) * s, "CompositeExplicitAutograd")
    lib.impl("shift", lambda a, b: a + b, "CompositeExplicitAutograd")
    mm_op, shift_op = torch.ops.vb_test.scaled_mm.default, torch.ops.vb_test.shift.default

    @register_prop_rule(mm_op)
    def _mm_rule(schema: OpSchema) -> OutputSharding:
        return einop_rule("mk,kn->mn", schema, linearity=False)

    seen = {}

    @register_op_strategy(shift_op)
    def _shift_strategy(mesh, schema):
        a, b = schema.args_schema
        assert isinstance(a, OpStrategy) and isinstance(b, OpStrategy)
        seen["cur"] = (a.strategies[0].output_spec.placements, b.strategies[0].output_spec.placements)
        alts = []
        for pl in ([Shard(0)], [Replicate()]):
            spec = DTensorSpec(mesh, tuple(pl))
            alts.append(PlacementStrategy(output_spec=spec, input_specs=[spec, spec], redistribute_cost=[generate_redistribute_costs(a, spec), generate_redistribute_costs(b, spec)]))
        return OpStrategy(alts)

    mesh = init_device_mesh("cpu", (world,), mesh_dim_names=("TP",))
    torch.manual_seed(0)
    A, B = torch.randn(8, 12), torch.randn(12, 6)
    # row-parallel: contraction sharded -> the einop rule leaves the output Partial; the product is right after reduction
    dA, dB = distribute_tensor(A, mesh, [Shard(1)]), distribute_tensor(B, mesh, [Shard(0)])
    out = mm_op(dA, dB, 0.5)
    assert isinstance(out, DTensor) and out.placements[0].is_partial()
    torch.testing.assert_close(out.full_tensor(), (A @ B) * 0.5, rtol=1e-5, atol=1e-5)
    # conflicting layouts (both operands sharded along the SAME mesh dim on m and n): the rule answers with a suggestion, inputs are resharded
    out = mm_op(distribute_tensor(A, mesh, [Shard(0)]), distribute_tensor(B, mesh, [Shard(1)]), 2.0)
    torch.testing.assert_close(out.full_tensor(), (A @ B) * 2.0, rtol=1e-5, atol=1e-5)
    # a Partial operand into a non-linear rule is reduced first
    P = DTensor.from_local(A / world, mesh, [Partial("sum")], run_check=False)
    out = mm_op(P, distribute_tensor(B, mesh, [Replicate()]), 1.0)
    torch.testing.assert_close(out.full_tensor(), A @ B, rtol=1e-5, atol=1e-5)
    # strategy contract: the cheapest alternative given the current layouts wins (both sharded -> stays sharded, no communication)
    X, Y = torch.randn(8, 4), torch.randn(8, 4)
    out = shift_op(distribute_tensor(X, mesh, [Shard(0)]), distribute_tensor(Y, mesh, [Shard(0)]))
    assert out.placements == (Shard(0),) and seen["cur"] == ((Shard(0),), (Shard(0),))
    torch.testing.assert_close(out.full_tensor(), X + Y)
    out = shift_op(distribute_tensor(X, mesh, [Replicate()]), distribute_tensor(Y, mesh, [Shard(0)]))
    torch.testing.assert_close(out.full_tensor(), X + Y)

    # the building blocks on their own
    sa = DTensorSpec(mesh, (Shard(0),), dA._spec.tensor_meta)
    sb = DTensorSpec(mesh, (Replicate(),), dB._spec.tensor_meta)
    r = einop_rule("mk,kn->mn", OpSchema(mm_op, (sa, sb, 1.0), {}, mesh))
    assert r.output_spec.placements == (Shard(0),) and not is_tensor_partial(r.output_spec)
    r = einop_rule("mk,kn->mn", OpSchema(mm_op, (sa, sb, 1.0), {}, mesh), enforce_sharding={"m": -1})
    assert r.output_spec is None and r.schema_suggestions[0].args_spec[0].placements == (Replicate(),)
    from vescale_b200.spec import TensorMeta
    bias = DTensorSpec(mesh, (Shard(0),), TensorMeta((6,), (1,), torch.float32))
    act = DTensorSpec(mesh, (Shard(1),), TensorMeta((8, 6), (6, 1), torch.float32))
    r = pointwise_rule(OpSchema(shift_op, (act, bias), {}, mesh))
    assert r.output_spec.placements == (Shard(1),)  # the bias' dim 0 IS the activation's dim 1 after right-alignment
    one = DTensorSpec(mesh, (Replicate(),), TensorMeta((8, 1), (1, 1), torch.float32))
    assert pointwise_rule(OpSchema(shift_op, (act, one), {}, mesh)).output_spec.placements == (Shard(1),)
    strat = gen_einsum_strategies("bmk,bkn->bmn", mesh)
    outs = {s.output_spec.placements for s in strat.strategies}
    assert outs == {(Replicate(),), (Shard(0),), (Shard(1),), (Shard(2),), (Partial("sum"),)}
    assert len(gen_einsum_strategies("mk,kn->mn", mesh, linearity=True).strategies) == 5
    lib._destroy()


def test_custom_op_rules_through_reference_registration_contracts():
    run_distributed(_custom_rules, 4)


def _cross_mesh_autograd(rank, world):
    import torch
    from vescale_b200 import DeviceMesh, DTensor, Replicate, Shard, distribute_tensor
    from vescale_b200.dtensor.cross_mesh import CrossMeshRedistribute, cross_mesh_anchor

    src, dst = DeviceMesh("cpu", [0, 1], _validate_mesh=False), DeviceMesh("cpu", [2, 3], _validate_mesh=False)
    torch.manual_seed(3)
    full = torch.randn(6, 4)
    w = torch.arange(24.0).reshape(6, 4)
    if rank < 2:
        x = distribute_tensor(full, src, [Shard(0)]).detach().requires_grad_(True)
        stub = CrossMeshRedistribute.apply(x, src, [Shard(0)], dst, [Shard(1)])
        assert not isinstance(stub, DTensor)
        stub.backward()  # joins the backward: receives d(loss)/dx from the target mesh
        torch.testing.assert_close(x.grad.full_tensor(), w)
        assert x.grad.placements == (Shard(0),)
    else:
        y = CrossMeshRedistribute.apply(None, src, [Shard(0)], dst, [Shard(1)], cross_mesh_anchor())
        assert isinstance(y, DTensor) and y.placements == (Shard(1),) and y.requires_grad
        torch.testing.assert_close(y.full_tensor(), full)
        (y * distribute_tensor(w, dst, [Shard(1)])).sum().backward()


def test_cross_mesh_redistribute_is_differentiable():
    run_distributed(_cross_mesh_autograd, 4)
