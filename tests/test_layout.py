"""Pure-math tests (no process group): placements, local shape/offset, ragged box decomposition.
Idea parity: reference ``test/dtensor/cpu_only/test_break_ragged_box.py`` (brute force all intervals)."""
import itertools
import math
import random

import pytest
import torch

from vescale_b200 import DeviceMesh
from vescale_b200.layout import (
    break_ragged_box,
    compute_local_shape_and_global_offset,
    flatten_index,
    local_boxes,
    ragged_flat_interval,
)
from vescale_b200.placement import InterleavedShard, RaggedShard, Replicate, Shard, _StridedRaggedShard, _StridedShard


def _check_boxes(shape, s, e):
    boxes = break_ragged_box(shape, s, e)
    assert len(boxes) <= 2 ** len(shape) - 1
    covered = torch.zeros(shape, dtype=torch.int32)
    order = []
    for off, sz in boxes:
        sl = tuple(slice(o, o + n) for o, n in zip(off, sz))
        covered[sl] += 1
        order.append(flatten_index(off, shape))
    flat = covered.view(-1)
    assert flat[s:e].eq(1).all() and flat[:s].eq(0).all() and flat[e:].eq(0).all(), (shape, s, e, boxes)
    assert order == sorted(order)


@pytest.mark.parametrize("shape", [(7,), (3, 4), (2, 3, 4)])
def test_break_ragged_box_bruteforce(shape):
    n = math.prod(shape)
    for s in range(n + 1):
        for e in range(s, n + 1):
            _check_boxes(shape, s, e)


def test_break_ragged_box_random_4d():
    rng = random.Random(0)
    for _ in range(2000):
        shape = tuple(rng.randint(1, 5) for _ in range(4))
        n = math.prod(shape)
        s = rng.randint(0, n)
        e = rng.randint(s, n)
        _check_boxes(shape, s, e)


def _mesh(shape, rank=0, names=None):
    return DeviceMesh("meta", torch.arange(math.prod(shape)).reshape(shape), mesh_dim_names=names, _rank=rank)


def test_shard_uneven_chunk_semantics():
    mesh = _mesh((4,))
    sizes = [compute_local_shape_and_global_offset((10, 3), mesh, [Shard(0)], (i,)) for i in range(4)]
    assert [s[0][0] for s in sizes] == [3, 3, 3, 1]
    assert [s[1][0] for s in sizes] == [0, 3, 6, 9]
    # size smaller than mesh: trailing empties
    sizes = [compute_local_shape_and_global_offset((2,), mesh, [Shard(0)], (i,))[0][0] for i in range(4)]
    assert sizes == [1, 1, 0, 0]


def test_ragged_local_shape_and_zero_units():
    mesh = _mesh((4,))
    p = RaggedShard((0,), (1, 2, 0, 1))
    shapes = [compute_local_shape_and_global_offset((8, 6), mesh, [p], (i,)) for i in range(4)]
    assert shapes[0] == ((2, 6), (0, 0))
    assert shapes[1] == ((4, 6), (2, 0))
    assert shapes[2] == ((0,), ())
    assert shapes[3] == ((2, 6), (6, 0))
    assert ragged_flat_interval((8, 6), mesh, [p], (1,)) == (12, 36)


def test_ragged_composed_with_shard():
    # (RaggedShard, Shard(1)): Shard(1) applies first, ragged flattens the remaining rows
    mesh = _mesh((2, 2))
    pl = [RaggedShard((0,), (3, 1)), Shard(1)]
    s, o = compute_local_shape_and_global_offset((8, 6), mesh, pl, (0, 1))
    assert s == (6, 3) and o == (0, 3)
    s, o = compute_local_shape_and_global_offset((8, 6), mesh, pl, (1, 0))
    assert s == (2, 3) and o == (6, 0)
    # (_StridedRaggedShard, Shard(0)): rows are cut by Shard(0) first
    pl = [_StridedRaggedShard((0,), (1, 3), split_factor=2), Shard(0)]
    s, o = compute_local_shape_and_global_offset((8, 6), mesh, pl, (1, 1))
    assert s == (3, 6) and o == (5, 0)


def test_strided_shard_offsets():
    # FSDP(dp) over TP on the same dim: [_StridedShard(0, sf=tp), Shard(0)] on (dp=2, tp=2)
    mesh = _mesh((2, 2))
    pl = [_StridedShard(0, split_factor=2), Shard(0)]
    got = {(i, j): compute_local_shape_and_global_offset((8,), mesh, pl, (i, j)) for i in range(2) for j in range(2)}
    assert got[(0, 0)] == ((2,), (0,)) and got[(1, 0)] == ((2,), (2,))
    assert got[(0, 1)] == ((2,), (4,)) and got[(1, 1)] == ((2,), (6,))


def test_local_boxes_tile_the_tensor():
    shape = (6, 8)
    for mesh_shape, pls in [
        ((4,), [Shard(0)]),
        ((4,), [RaggedShard((0,), (1, 0, 3, 2))]),
        ((4,), [RaggedShard((0, 1), (5, 1, 1, 1))]),
        ((2, 2), [RaggedShard((0,), (1, 2)), Shard(1)]),
        ((2, 2), [Shard(0), Shard(1)]),
        ((2,), [InterleavedShard(1, 2)]),
    ]:
        mesh = _mesh(mesh_shape)
        cover = torch.zeros(shape, dtype=torch.int32)
        for coord in itertools.product(*[range(s) for s in mesh_shape]):
            for off, sz, _ in local_boxes(shape, mesh, pls, coord):
                cover[tuple(slice(o, o + n) for o, n in zip(off, sz))] += 1
        assert cover.eq(1).all(), (mesh_shape, pls, cover)


def test_placement_reprs_and_hash():
    p = RaggedShard((0,), (1, 2))
    assert repr(p) == "RaggedShard(dims=(0,), local_units=(1, 2))"
    assert p == RaggedShard((0,), [1, 2]) and hash(p) == hash(RaggedShard((0,), (1, 2)))
    assert Shard(1) != _StridedShard(1, split_factor=2)
    assert Replicate() == Replicate() and p.is_ragged_shard() and not Shard(0).is_ragged_shard()


def test_general_break_ragged_box_reference_entry_point():
    """``vescale.dtensor.vescale_utils.checkpoint._break_ragged_box``: any contiguous dim range flattened, the shard an n-d box in
    that ragged view; the returned boxes tile it exactly (sampled version of the reference's brute-force test
    ``test/dtensor/cpu_only/test_break_ragged_box.py``, which passes in full against this implementation)."""
    import itertools
    import math
    import random

    import torch
    from vescale.dtensor.vescale_utils.checkpoint import _break_ragged_box

    rnd = random.Random(7)
    worst = 0
    for shape in [(5,), (3, 4), (4, 3, 5), (2, 3, 4, 3), (3, 2, 2, 3, 2)]:
        n = len(shape)
        for d0, d1 in itertools.combinations_with_replacement(range(n), 2):
            rdims = tuple(range(d0, d1 + 1))
            rshape = shape[:d0] + (math.prod(shape[d0 : d1 + 1]),) + shape[d1 + 1 :]
            for _ in range(60):
                rect = []
                for s in rshape:
                    a = rnd.randint(0, s - 1)
                    rect.append((a, rnd.randint(a, s)))
                sizes, offs = _break_ragged_box(tuple(e - s for s, e in rect), tuple(s for s, _ in rect), rdims, shape, shape, (0,) * n)
                worst = max(worst, len(sizes))
                t = torch.zeros(shape, dtype=torch.int64)
                for sz, of in zip(sizes, offs):
                    t[tuple(slice(o, o + z) for o, z in zip(of, sz))] += 1
                want = torch.zeros(rshape, dtype=torch.int64)
                want[tuple(slice(s, e) for s, e in rect)] = 1
                assert torch.equal(t.view(rshape), want), (shape, rdims, rect)
    assert worst <= 2 * 5 - 1
    # a shard of a row-sharded tensor: ragged offsets are global flat positions, results come back in global coordinates
    sizes, offs = _break_ragged_box((7,), (4 * 3 + 2,), (0, 1), (4, 3), (8, 3), (4, 0))
    t = torch.zeros(8, 3, dtype=torch.int64)
    for sz, of in zip(sizes, offs):
        t[tuple(slice(o, o + z) for o, z in zip(of, sz))] += 1
    assert t.view(-1)[14:21].eq(1).all() and t.sum() == 7
    assert _break_ragged_box((0,), (), (0,), (4,), (4,), (0,)) == ([], [])
