"""Distributed DTensor tests on 4 ranks (gloo/CPU here, NCCL when 4 GPUs are present).
Golden = the same op on the full tensor on one device, as in the reference's DTensorTestBase tests
(``test/dtensor/ragged_shard/test_redistribute.py``, ``legacy/test/dtensor/general/test_redistribute.py``)."""
import itertools

import torch

from common import device_type, run_distributed


def _mk(rank, shape, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g, dtype=dtype).to(device_type())


def _redistribute_all(rank, world):
    from vescale_b200 import DeviceMesh, Shard, Replicate, Partial, RaggedShard, distribute_tensor, init_device_mesh
    from vescale_b200.dtensor import DTensor, InterleavedShard

    mesh = init_device_mesh(device_type(), (world,))
    full = _mk(rank, (8, 12))
    opts = [Replicate(), Shard(0), Shard(1), RaggedShard((0,), (1, 3, 0, 4)), RaggedShard((0, 1), (5, 1, 1, 1)), InterleavedShard(1, 3)]
    for a, b in itertools.product(opts, opts):
        dt = distribute_tensor(full, mesh, [a])
        out = dt.redistribute(mesh, [b])
        assert out.placements == (b,)
        assert torch.equal(out.full_tensor(), full), (a, b)
        ref = distribute_tensor(full, mesh, [b], src_data_rank=None)
        assert torch.equal(out.to_local(), ref.to_local()), (a, b)
    # uneven Shard
    full2 = _mk(rank, (10, 7), 1)
    for a, b in itertools.product([Shard(0), Shard(1), Replicate()], repeat=2):
        out = distribute_tensor(full2, mesh, [a]).redistribute(mesh, [b])
        assert torch.equal(out.full_tensor(), full2), (a, b)
    # Partial -> {Replicate, Shard, Ragged}
    local = _mk(rank, (8, 12), 100 + rank)
    want = sum(_mk(r, (8, 12), 100 + r) for r in range(world))
    dt = DTensor.from_local(local, mesh, [Partial()], shape=(8, 12))
    for b in [Replicate(), Shard(0), Shard(1), RaggedShard((0,), (2, 2, 3, 1))]:
        out = dt.redistribute(mesh, [b])
        torch.testing.assert_close(out.full_tensor(), want)
    # 2-D mesh
    mesh2 = init_device_mesh(device_type(), (2, 2), mesh_dim_names=("dp", "tp"))
    pls = [[Shard(0), Shard(1)], [Shard(0), Shard(0)], [Replicate(), Shard(1)], [Shard(1), Replicate()], [RaggedShard((0,), (1, 3)), Shard(1)], [Replicate(), Replicate()]]
    for a, b in itertools.product(pls, pls):
        out = distribute_tensor(full, mesh2, a).redistribute(mesh2, b)
        assert torch.equal(out.full_tensor(), full), (a, b)
    assert mesh2["tp"].size() == 2 and mesh2["dp"].get_group() is not None
    # collective-named API (legacy dtensor/api.py:314-436), also through the `vescale` alias package
    from vescale.dtensor.api import vescale_all_gather, vescale_all_reduce, vescale_reduce_scatter

    sh = distribute_tensor(full, mesh2, [Shard(0), Shard(1)])
    g1 = vescale_all_gather(sh, mesh_dims=1)
    assert g1.placements == (Shard(0), Replicate()) and torch.equal(g1.full_tensor(), full)
    g2 = vescale_all_gather(sh)
    assert g2.placements == (Replicate(), Replicate()) and torch.equal(g2.to_local(), full)
    try:
        vescale_all_gather(g1, mesh_dims=[1])
        raise AssertionError("all-gather over an unsharded mesh dim must be rejected")
    except ValueError:
        pass
    part = DTensor.from_local(local, mesh2, [Partial(), Partial()], shape=(8, 12))
    r1 = vescale_all_reduce(part, mesh_dims="tp")
    assert r1.placements == (Partial(), Replicate())
    torch.testing.assert_close(vescale_all_reduce(part).to_local(), want)
    rs = vescale_reduce_scatter(part, scatter_dims=[0], mesh_dims=[1])
    assert rs.placements == (Replicate(), Shard(0))
    torch.testing.assert_close(rs.full_tensor(), want)


def _ops(rank, world):
    from vescale_b200 import DeviceMesh, Shard, Replicate, Partial, RaggedShard, distribute_tensor, init_device_mesh
    from vescale_b200.dtensor import DTensor, implicit_replication
    from vescale_b200.dtensor.debug import CommDebugMode

    mesh = init_device_mesh(device_type(), (world,))
    a, b = _mk(rank, (8, 16), 1), _mk(rank, (16, 12), 2)
    # matmul: BASELINE config #1 — Shard(0) x Replicate -> Shard(0) -> Replicate
    da, db = distribute_tensor(a, mesh, [Shard(0)]), distribute_tensor(b, mesh, [Replicate()])
    dc = torch.mm(da, db)
    assert dc.placements == (Shard(0),)
    torch.testing.assert_close(dc.redistribute(mesh, [Replicate()]).to_local(), a @ b)
    # k-sharded -> Partial -> reduce
    dc = torch.mm(distribute_tensor(a, mesh, [Shard(1)]), distribute_tensor(b, mesh, [Shard(0)]))
    assert dc.placements == (Partial(),)
    torch.testing.assert_close(dc.full_tensor(), a @ b)
    with CommDebugMode() as cm:
        dc = torch.mm(distribute_tensor(a, mesh, [Replicate()], src_data_rank=None), distribute_tensor(b, mesh, [Shard(1)], src_data_rank=None))
    assert dc.placements == (Shard(1),) and cm.get_total_counts() == 0
    # pointwise + broadcasting + reductions
    x = distribute_tensor(a, mesh, [Shard(0)])
    bias = distribute_tensor(a[0], mesh, [Replicate()])
    y = torch.nn.functional.gelu(x * 2 + bias) - x.mean(dim=1, keepdim=True)
    torch.testing.assert_close(y.full_tensor(), torch.nn.functional.gelu(a * 2 + a[0]) - a.mean(1, keepdim=True))
    torch.testing.assert_close(x.sum().full_tensor(), a.sum())
    assert x.sum().placements == (Partial(),)
    torch.testing.assert_close(x.max().full_tensor(), a.max())
    torch.testing.assert_close(torch.linalg.vector_norm(x).full_tensor(), torch.linalg.vector_norm(a))
    torch.testing.assert_close(x.softmax(-1).full_tensor(), a.softmax(-1))
    torch.testing.assert_close(x.softmax(0).full_tensor(), a.softmax(0))
    # views
    v = x.view(2, 4, 16)
    assert v.placements == (Replicate(),) or v.placements == (Shard(0),) or v.placements == (Shard(1),)
    torch.testing.assert_close(v.full_tensor(), a.view(2, 4, 16))
    v = distribute_tensor(a, mesh, [Shard(0)]).view(8, 4, 4).transpose(0, 1).contiguous()
    torch.testing.assert_close(v.full_tensor(), a.view(8, 4, 4).transpose(0, 1))
    assert v.placements == (Shard(1),)
    # cat / slice / index
    c = torch.cat([x, x], dim=1)
    torch.testing.assert_close(c.full_tensor(), torch.cat([a, a], 1))
    torch.testing.assert_close(x[:, 2:6].full_tensor(), a[:, 2:6])
    torch.testing.assert_close(x[1:5].full_tensor(), a[1:5])
    # autograd through redistribute + matmul (TP column->row parallel MLP)
    w1, w2 = _mk(rank, (16, 32), 3).requires_grad_(), _mk(rank, (32, 16), 4).requires_grad_()
    xin = _mk(rank, (8, 16), 5)
    ref = (torch.relu(xin @ w1) @ w2).sum()
    ref.backward()
    dw1 = distribute_tensor(w1.detach(), mesh, [Shard(1)]).requires_grad_()
    dw2 = distribute_tensor(w2.detach(), mesh, [Shard(0)]).requires_grad_()
    dx = distribute_tensor(xin, mesh, [Replicate()])
    out = torch.relu(dx @ dw1) @ dw2
    assert out.placements == (Partial(),)
    loss = out.redistribute(mesh, [Replicate()]).sum()
    loss.backward()
    torch.testing.assert_close(loss.full_tensor(), ref.detach())
    torch.testing.assert_close(dw1.grad.full_tensor(), w1.grad)
    torch.testing.assert_close(dw2.grad.full_tensor(), w2.grad)
    # ragged: elementwise in-place with implicit replication, zero-unit ranks, norm, fused adam
    rp = RaggedShard((0,), (3, 0, 4, 1))
    r = distribute_tensor(a, mesh, [rp])
    with implicit_replication():
        r.add_(1.0)
        r.mul_(torch.tensor(2.0).to(a.device))
    torch.testing.assert_close(r.full_tensor(), (a + 1) * 2)
    torch.testing.assert_close(torch.linalg.vector_norm(r).full_tensor(), torch.linalg.vector_norm((a + 1) * 2))
    for dim in (0, 1):
        torch.testing.assert_close(torch.linalg.vector_norm(r, 2, dim=[dim]).full_tensor(), torch.linalg.vector_norm((a + 1) * 2, 2, dim=[dim]), rtol=1e-5, atol=1e-5)
    r3 = distribute_tensor(_mk(rank, (4, 6, 5), 9), mesh, [RaggedShard((0, 1), (5, 1, 1, 5))])
    f3 = _mk(rank, (4, 6, 5), 9)
    for dims in ([0], [1], [2], [0, 2], [1, 2], [0, 1]):
        torch.testing.assert_close(torch.linalg.vector_norm(r3, 2, dim=dims).full_tensor(), torch.linalg.vector_norm(f3, 2, dim=dims), rtol=1e-5, atol=1e-5)
    # ragged -> replicate has zero comm in backward (reference test_redistribute.py:44-86)
    rr = distribute_tensor(a, mesh, [rp]).requires_grad_()
    full = rr.redistribute(mesh, [Replicate()])
    with CommDebugMode() as cm:
        full.to_local().sum().backward()
    assert cm.get_total_counts() == 0, cm.get_comm_counts()
    assert rr.grad.placements == (rp,)
    torch.testing.assert_close(rr.grad.full_tensor(), torch.ones_like(a))
    # optimizer on ragged DTensor params (foreach + fused paths)
    for kw in ({"foreach": True}, {"fused": True} if device_type() == "cuda" else {"foreach": False}):
        p_ref = a.clone().requires_grad_()
        p_dt = torch.nn.Parameter(distribute_tensor(a.clone(), mesh, [rp]))
        o_ref, o_dt = torch.optim.AdamW([p_ref], lr=0.1, **kw), torch.optim.AdamW([p_dt], lr=0.1, **kw)
        for step in range(2):
            g = _mk(rank, (8, 16), 50 + step)
            p_ref.grad = g.clone()
            p_dt.grad = distribute_tensor(g, mesh, [rp])
            o_ref.step(), o_dt.step()
        torch.testing.assert_close(p_dt.full_tensor(), p_ref.detach())
    # clip_grad_norm_ over mixed placements
    ps = [torch.nn.Parameter(distribute_tensor(a.clone(), mesh, [rp])), torch.nn.Parameter(distribute_tensor(b.clone(), mesh, [Shard(0)]))]
    for p, g in zip(ps, (a, b)):
        p.grad = distribute_tensor(g * 3, mesh, p.placements)
    total = torch.nn.utils.clip_grad_norm_(ps, 1.0, foreach=False)
    want = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(a * 3), torch.linalg.vector_norm(b * 3)]))
    total = total.full_tensor() if isinstance(total, DTensor) else total
    torch.testing.assert_close(total, want)
    # embedding with vocab-sharded weight
    emb = _mk(rank, (20, 6), 7)
    ids = torch.randint(0, 20, (3, 5), generator=torch.Generator().manual_seed(3)).to(a.device)
    de = distribute_tensor(emb, mesh, [Shard(0)]).requires_grad_()
    out = torch.nn.functional.embedding(distribute_tensor(ids, mesh, [Replicate()]), de)
    assert out.placements == (Partial(),)
    torch.testing.assert_close(out.full_tensor(), emb[ids])
    out.redistribute(mesh, [Replicate()]).sum().backward()
    eg = torch.zeros_like(emb).index_add_(0, ids.view(-1), torch.ones(15, 6, device=a.device))
    torch.testing.assert_close(de.grad.full_tensor(), eg)
    # loss parallel
    from vescale_b200.dtensor import loss_parallel

    logits = _mk(rank, (6, 20), 11).requires_grad_()
    tgt = torch.randint(0, 20, (6,), generator=torch.Generator().manual_seed(5)).to(a.device)
    ref = torch.nn.functional.cross_entropy(logits, tgt)
    ref.backward()
    dl = distribute_tensor(logits.detach(), mesh, [Shard(1)]).requires_grad_()
    with loss_parallel():
        l = torch.nn.functional.cross_entropy(dl, distribute_tensor(tgt, mesh, [Replicate()]))
        l.backward()
    torch.testing.assert_close(l.full_tensor(), ref.detach())
    torch.testing.assert_close(dl.grad.full_tensor(), logits.grad)


def _factories_random(rank, world):
    import vescale_b200.dtensor as vd
    from vescale_b200 import Shard, Replicate, RaggedShard, init_device_mesh

    mesh = init_device_mesh(device_type(), (world,))
    z = vd.zeros(8, 6, device_mesh=mesh, placements=[Shard(0)])
    assert z.to_local().shape == (2, 6) and z.shape == (8, 6)
    o = vd.full((8, 6), 3.0, device_mesh=mesh, placements=[RaggedShard((0,), (1, 1, 1, 5))])
    assert torch.equal(o.full_tensor(), torch.full((8, 6), 3.0, device=o.device))
    # single-device-equivalent randomness: any placement gives the same global tensor
    outs = []
    for pl in ([Replicate()], [Shard(0)], [Shard(1)], [RaggedShard((0,), (1, 0, 2, 1))]):
        vd.manual_seed(1234, mesh)
        outs.append(vd.randn(8, 6, device_mesh=mesh, placements=pl).full_tensor())
        u = vd.rand(8, 6, device_mesh=mesh, placements=pl).full_tensor()
        assert (u >= 0).all() and (u < 1).all()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    assert abs(outs[0].mean().item()) < 0.5 and 0.5 < outs[0].std().item() < 1.5
    # cross-rank equality helpers: a mismatch on one rank makes every rank say False
    from vescale_b200.dtensor import allclose, equal

    a = vd.ones(8, 6, device_mesh=mesh, placements=[Shard(0)])
    b = vd.ones(8, 6, device_mesh=mesh, placements=[Shard(0)])
    assert equal(a, b) and allclose(a, b)
    if rank == 2:
        b.to_local()[0, 0] = 1.0 + 1e-7
    assert not equal(a, b) and allclose(a, b, rtol=1e-5)
    assert not equal(a, vd.ones(8, 6, device_mesh=mesh, placements=[Shard(1)]))
    # single-device-equivalent aten random ops (ThreadBasedRNGTracker): the mask / values do not depend on the sharding
    from vescale_b200.dtensor.random import OffsetBasedRNGTracker, TensorParallelRNGTracker, ThreadBasedRNGTracker, set_rng_tracker

    set_rng_tracker(ThreadBasedRNGTracker())
    base = torch.arange(48.0).reshape(8, 6).to(z.device) + 1
    outs_d, outs_n = [], []
    for pl in ([Replicate()], [Shard(0)], [Shard(1)], [RaggedShard((0,), (1, 0, 2, 1))]):
        vd.manual_seed(99, mesh)
        dt = vd.distribute_tensor(base, mesh, pl, src_data_rank=None)
        outs_d.append(torch.nn.functional.dropout(dt, 0.4, training=True).full_tensor())
        outs_n.append(torch.empty_like(dt).normal_(1.0, 2.0).full_tensor() if not isinstance(pl[0], RaggedShard) else None)
    assert all(torch.equal(o, outs_d[0]) for o in outs_d[1:]) and 0.15 < (outs_d[0] == 0).float().mean().item() < 0.65
    assert all(torch.equal(o, outs_n[0]) for o in outs_n[1:] if o is not None)
    set_rng_tracker(TensorParallelRNGTracker(tp_mesh_dim=0))
    vd.manual_seed(5, mesh)
    y = torch.nn.functional.dropout(vd.ones(8, 64, device_mesh=mesh, placements=[Shard(0)]), 0.5, training=True).full_tensor()
    assert not torch.equal(y[:2], y[2:4])  # TP ranks draw different streams
    set_rng_tracker(OffsetBasedRNGTracker())
    # dropout on a sharded tensor: replicas agree, shards differ
    vd.manual_seed(7, mesh)
    x = vd.ones(8, 64, device_mesh=mesh, placements=[Shard(0)])
    y = torch.nn.functional.dropout(x, 0.5, training=True).full_tensor()
    assert 0.2 < (y == 0).float().mean().item() < 0.8
    assert not torch.equal(y[:2], y[2:4])


def test_redistribute_matrix():
    run_distributed(_redistribute_all, 4)


def test_ops_and_autograd():
    run_distributed(_ops, 4)


def test_factories_and_random():
    run_distributed(_factories_random, 4)


def _torch_mesh_strided_ragged_2d(rank, world):
    """The reference's ``test_ragged_shard_2d`` scenario at CPU-sized tensors (``test/dtensor/ragged_shard/test_redistribute.py:
    217-300``): meshes are ``torch.distributed.device_mesh`` objects (the reference's new package is built on them), a
    Shard(s) DTensor's local tensor is ragged-sharded over a second mesh dim, the 2-D spec ``[_StridedRaggedShard | RaggedShard,
    Shard(s)]`` is assembled by hand with ``DTensorSpec`` / ``DTensor(...)``, and ``full_tensor()`` must reproduce the original."""
    import math

    import numpy as np
    from torch.distributed.device_mesh import init_device_mesh as torch_init_device_mesh

    from vescale.dtensor import DTensor, distribute_tensor
    from vescale.dtensor.placement_types import DTensorSpec, RaggedShard, Shard, TensorMeta, _StridedRaggedShard

    np.random.seed(42)
    torch.manual_seed(42)

    def units(N, k):
        if np.random.randint(1, 5) == 1:  # everything on one rank (the Muon gather-to-root layout)
            rt = np.random.randint(0, k)
            return tuple(1 if i == rt else 0 for i in range(k))
        return tuple(int(x) for x in np.random.multinomial(N, np.full(k, 1.0 / k)))

    for mesh_size in [(1, world), (2, world // 2), (world, 1)]:
        dm = torch_init_device_mesh(device_type(), mesh_size, mesh_dim_names=["ragged", "other"])
        for _ in range(3):
            nums = [int(np.random.randint(3, 9)) * mesh_size[1] for _ in range(3)]
            for input_size in (tuple(nums[:1]), tuple(nums[:2]), tuple(nums[:3])):
                ndim = len(input_size)
                g = torch.randn(input_size).to(device_type())
                for rd, sd in [(d, 0) for d in range(ndim)] + ([(0, 1)] if ndim > 1 else []):
                    dt_shard = distribute_tensor(g, dm["other"], [Shard(sd)])
                    st = dt_shard._local_tensor
                    rdims = tuple(range(rd + 1))
                    lst = units(math.prod(st.shape[: rd + 1]), dm["ragged"].size())
                    dt2 = distribute_tensor(st, dm["ragged"], [RaggedShard(rdims, lst)])
                    first = _StridedRaggedShard(dims=rdims, local_units=lst, split_factor=dm["other"].size()) if sd == 0 else RaggedShard(dims=rdims, local_units=lst)
                    spec = DTensorSpec(dm, [first, Shard(sd)], tensor_meta=TensorMeta(dt_shard.size(), dt_shard.stride(), dt_shard.dtype))
                    d = DTensor(dt2._local_tensor, spec, requires_grad=True)
                    full = d.full_tensor()
                    assert full.shape == g.shape and torch.equal(full, g), (mesh_size, input_size, rd, sd, lst)
    # torch's own implicit_replication() context manager is honoured by the dispatcher
    from torch.distributed.tensor.experimental import implicit_replication as torch_implicit_replication

    mesh1 = torch_init_device_mesh(device_type(), (world,))
    z = distribute_tensor(torch.zeros(8, 6, dtype=torch.int64).to(device_type()), mesh1, (RaggedShard((0, 1), (0, 1, 5, 2)[:world] if world == 4 else (1,) * world),))
    add = torch.arange(48).view(8, 6).to(device_type())
    with torch_implicit_replication():
        z.add_(add)
    assert torch.equal(z.full_tensor(), add)


def test_torch_device_mesh_and_strided_ragged_2d():
    run_distributed(_torch_mesh_strided_ragged_2d, 4)
