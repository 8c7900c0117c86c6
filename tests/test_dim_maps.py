"""Dim-map algebra (``dtensor/rules/dim_maps.py``) against real tensors: apply the op to the full tensor and to every shard, and
compare; cross-check with the hand-written view rules (reference: ``legacy/test/dtensor/ops/test_view_ops.py`` checks the same
table by running the ops on DTensors)."""
import itertools

import pytest
import torch
from torch import Tensor

from vescale_b200.dtensor.op_schema import OpSchema
from vescale_b200.dtensor.rules import dim_maps as dm
from vescale_b200.dtensor.rules.dim_maps import (
    Broadcast, Flatten, InputDim, NewDim, Repeat, Singleton, Split, dim_view, propagate_shape_and_sharding,
)
from vescale_b200.dtensor.rules.view import map_view_placements
from vescale_b200.mesh import init_device_mesh
from vescale_b200.placement import Replicate, Shard
from vescale_b200.spec import DTensorSpec, TensorMeta


def test_constructors_simplify():
    assert Flatten.new((InputDim(2),)) == InputDim(2)
    assert Flatten.new(()) == Singleton()
    assert Split.new(InputDim(0), (6,), 0) == InputDim(0)
    assert Split.new(InputDim(0), (1, 6), 1) == InputDim(0)
    assert Split.new(InputDim(0), (2, 1, 3), 2) == Split(InputDim(0), (2, 3), 1)
    assert Split.new(InputDim(0), (2, 1, 3), 1) == Singleton()
    assert Repeat.new(InputDim(1), 1) == InputDim(1)
    assert Repeat.new(Singleton(), 3) == Broadcast(Singleton(), 3)
    assert NewDim.new(1) == Singleton() and NewDim.new(4) == NewDim(4)
    assert dm.normalize_sizes(((2, 3),)) == (2, 3) and dm.normalize_sizes((2, 3)) == (2, 3)


def test_builders():
    assert dm.dim_pad_left(1, 3) == (Singleton(), Singleton(), InputDim(0))
    assert dm.dim_atleast_3d(1) == (Singleton(), InputDim(0), Singleton())
    assert dm.dim_movedim(3, 0, 2) == (InputDim(1), InputDim(2), InputDim(0))
    assert dm.dim_transpose(3, -1, 0) == (InputDim(2), InputDim(1), InputDim(0))
    assert dm.dim_squeeze((1, 4, 1)) == (InputDim(1),)
    assert dm.dim_squeeze((1, 4, 1), 0) == (InputDim(1), InputDim(2))
    assert dm.dim_unsqueeze(2, -1) == (InputDim(0), InputDim(1), Singleton())
    assert dm.dim_reduction(3, 1, True) == (InputDim(0), Singleton(), InputDim(2))
    assert dm.dim_reduction(3, (0, 2), False) == (InputDim(1),)
    assert dm.dim_flatten(3) == (Flatten((InputDim(0), InputDim(1), InputDim(2))),)
    assert dm.dim_flatten(3, 1, 2) == (InputDim(0), Flatten((InputDim(1), InputDim(2))))
    assert dm.expand((1, 4), (3, 2, 4)) == (NewDim(3), Broadcast(InputDim(0), 2), InputDim(1))
    assert dm.dim_tile(3, (2,)) == (InputDim(0), InputDim(1), Repeat(InputDim(2), 2))


def _eval_shape(dmap, shape):
    return propagate_shape_and_sharding([Replicate()], shape, dmap, (1,))[0]


CASES = [
    (torch.atleast_2d, (6,), (), {}),
    (torch.atleast_3d, (4, 6), (), {}),
    (torch.broadcast_to, (1, 6), ((4, 4, 6),), {}),
    (Tensor.expand, (4, 1, 6), (4, 3, 6), {}),
    (Tensor.expand, (4, 1, 6), ((-1, 3, -1),), {}),
    (torch.flatten, (4, 2, 6), (), {}),
    (torch.flatten, (4, 2, 6), (1, 2), {}),
    (torch.movedim, (4, 2, 6), (0, 2), {}),
    (torch.permute, (4, 2, 6), ((2, 0, 1),), {}),
    (torch.ravel, (4, 6), (), {}),
    (Tensor.repeat, (4, 6), (2, 1, 3), {}),
    (torch.reshape, (4, 6, 8), ((24, 8),), {}),
    (torch.reshape, (24, 8), ((4, 6, 8),), {}),
    (torch.reshape, (4, 6, 8), ((4, 2, 3, 4, 2),), {}),
    (torch.reshape, (4, 1, 8), ((-1, 2),), {}),
    (torch.squeeze, (4, 1, 6), (), {}),
    (torch.squeeze, (4, 1, 6), (1,), {}),
    (torch.tile, (4, 6), ((2,),), {}),
    (torch.transpose, (4, 2, 6), (0, 2), {}),
    (torch.unsqueeze, (4, 6), (1,), {}),
    (Tensor.view, (4, 6, 8), (4, 48), {}),
    (Tensor.view, (8, 6), (2, 4, 6), {}),
]


@pytest.mark.parametrize("fn,shape,args,kwargs", CASES, ids=[f"{getattr(c[0], '__name__', str(c[0]))}-{i}" for i, c in enumerate(CASES)])
def test_ops_table_against_real_tensors(fn, shape, args, kwargs):
    """For every input dim and mesh size: if the map says the dim can stay sharded, running the op shard by shard (size argument
    rewritten to the local shape) and concatenating along the output placement reproduces the op on the full tensor."""
    x = torch.arange(float(torch.Size(shape).numel())).reshape(shape)
    ref = fn(x, *args, **kwargs)
    spec = dm.ops[fn]
    dmap = spec.dim_map(x.to("meta"), *args, **kwargs)
    assert _eval_shape(dmap, shape) == tuple(ref.shape)
    checked = 0
    for d, n in itertools.product(range(x.ndim), (2, 4)):
        if shape[d] % n or shape[d] == n:  # a local size of 1 would change what a dim-less squeeze removes
            continue
        out_shape, out_pl, shardable = propagate_shape_and_sharding([Shard(d)], shape, dmap, (n,))
        assert out_shape == tuple(ref.shape)
        if out_pl is None:
            assert not bool(shardable[d, 0])
            continue
        (p,) = out_pl
        assert isinstance(p, Shard) and out_shape[p.dim] % n == 0
        local_out = list(out_shape)
        local_out[p.dim] //= n
        pieces = []
        for shard in x.chunk(n, dim=d):
            a = list(args)
            if spec.shape_argnum is not None:
                a = [tuple(local_out)]
            pieces.append(fn(shard, *a, **kwargs))
        assert torch.equal(torch.cat(pieces, dim=p.dim), ref), (fn, d, n)
        checked += 1
    assert checked or fn in (torch.ravel,)


def test_agrees_with_handwritten_view_rule():
    """Where both produce plain ``Shard`` outputs they must name the same dim; where the dim map needs a replicate, the
    hand-written rule either replicates too or answers with an ``InterleavedShard`` (which the algebra does not model)."""
    mesh = init_device_mesh("cpu", (2,), _rank=0, _init_process_groups=False)
    for in_shape, out_shape in [((4, 6, 8), (24, 8)), ((24, 8), (4, 6, 8)), ((4, 6, 8), (4, 48)), ((4, 48), (4, 6, 8)), ((2, 4, 6), (8, 6)), ((8, 6), (8, 2, 3))]:
        for d in range(len(in_shape)):
            stride = torch.empty(in_shape, device="meta").stride()
            spec = DTensorSpec(mesh, (Shard(d),), TensorMeta(tuple(in_shape), tuple(stride), torch.float32))
            ins, outs = map_view_placements(spec, out_shape, mesh)
            _, pl, _ = propagate_shape_and_sharding([Shard(d)], in_shape, dim_view(in_shape, out_shape), (2,))
            if pl is not None:
                assert type(outs[0]) is Shard and outs[0].dim == pl[0].dim, (in_shape, out_shape, d, outs, pl)
            else:
                assert not (type(outs[0]) is Shard and ins[0] == Shard(d)), (in_shape, out_shape, d, outs)


def test_rule_from_map_replicates_what_cannot_stay():
    mesh = init_device_mesh("cpu", (2, 2), _rank=0, _init_process_groups=False)
    stride = torch.empty((4, 6, 8), device="meta").stride()
    spec = DTensorSpec(mesh, (Shard(0), Shard(1)), TensorMeta((4, 6, 8), tuple(stride), torch.float32))
    rule = dm.dim_map_rule(dm.ops[torch.reshape])
    res = rule(OpSchema(torch.ops.aten.reshape.default, (spec, [24, 8]), {}))
    assert tuple(res.ins[0]) == (Shard(0), Replicate()) and tuple(res.out) == (Shard(0), Replicate())
    assert res.local_args == {1: [12, 8]}
    res = rule(OpSchema(torch.ops.aten.reshape.default, (spec, [4, 3, 2, 8]), {}))
    assert tuple(res.out) == (Shard(0), Replicate())  # 3 pieces do not divide over 2 ranks
    res = rule(OpSchema(torch.ops.aten.reshape.default, (spec, [4, 2, 3, 8]), {}))
    assert tuple(res.out) == (Shard(0), Shard(1)) and res.local_args == {1: [2, 1, 3, 8]}


def _w_dtensor_ops(rank, world):
    """The composite ops of the table (movedim / tile / ravel / atleast_3d / broadcast_to) and view_as_real / view_as_complex on
    live DTensors, every shardable input dim, against the same op on the full tensor."""
    from common import device_type
    from vescale_b200 import distribute_tensor, init_device_mesh as idm

    mesh = idm(device_type(), (world,))
    torch.manual_seed(0)
    x = torch.randn(4, 6, 8)
    fns = [
        lambda t: torch.movedim(t, 0, 2),
        lambda t: torch.tile(t, (2,)),
        lambda t: torch.ravel(t),
        lambda t: torch.atleast_3d(t),
        lambda t: torch.broadcast_to(t.unsqueeze(0), (3, 4, 6, 8)),
        lambda t: t.repeat(2, 1, 1),
        lambda t: torch.flatten(t, 1, 2),
        lambda t: torch.view_as_real(torch.view_as_complex(t.reshape(4, 6, 4, 2))),
    ]
    for k, fn in enumerate(fns):
        ref = fn(x)
        for d in range(3):
            dx = distribute_tensor(x, mesh, [Shard(d)])
            out = fn(dx)
            assert torch.equal(out.full_tensor(), ref), (k, d, out.placements)
    # a rule built from the table, attached to a custom op
    lib = torch.library.Library("dimmaptest", "DEF")
    lib.define("fold(Tensor x, int[] shape) -> Tensor")
    lib.impl("fold", lambda t, shape: t.reshape(shape).clone(), "CompositeExplicitAutograd")
    dm.register_prop_rule_map(torch.ops.dimmaptest.fold.default, torch.reshape)
    dx = distribute_tensor(x, mesh, [Shard(0)])
    out = torch.ops.dimmaptest.fold(dx, [24, 8])
    assert out.placements == (Shard(0),) and out.to_local().shape == (24 // world, 8)
    assert torch.equal(out.full_tensor(), x.reshape(24, 8))
    out = torch.ops.dimmaptest.fold(distribute_tensor(x, mesh, [Shard(1)]), [24, 8])  # cannot stay sharded: replicated first
    assert torch.equal(out.full_tensor(), x.reshape(24, 8))


def test_dtensor_ops_2ranks():
    from common import run_distributed

    run_distributed(_w_dtensor_ops, 2)
