"""Op-level sweep in the style of the reference's ``DTensorConverter`` (``test/common_dtensor.py:433-562``,
``legacy/test/dtensor/ops/test_{pointwise,math,matrix,tensor,view}_ops.py``): every op runs under every combination of
``Replicate`` / ``Shard(d)`` placements of its tensor arguments (even and uneven shards) on 4 ranks and must reproduce the
single-device result; a second pass checks gradients through a few representative ops and a 2-D mesh."""
import itertools

import torch
import torch.nn.functional as F

from common import device_type, run_distributed


def _t(shape, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    if dtype in (torch.int64, torch.int32):
        return torch.randint(0, 5, shape, generator=g, dtype=dtype).to(device_type())
    if dtype == torch.bool:
        return (torch.rand(shape, generator=g) > 0.5).to(device_type())
    return torch.randn(shape, generator=g, dtype=dtype).to(device_type())


def _choices(t, max_dims=3):
    from vescale_b200 import Replicate, Shard

    return [Replicate()] + [Shard(d) for d in range(min(t.ndim, max_dims))]


def _compare(out, ref, name, combo):
    from vescale_b200.dtensor import DTensor

    if isinstance(ref, (tuple, list)):
        assert isinstance(out, (tuple, list)) and len(out) == len(ref), (name, combo)
        for o, r in zip(out, ref):
            _compare(o, r, name, combo)
        return
    if isinstance(ref, torch.Tensor):
        got = out.full_tensor() if isinstance(out, DTensor) else out
        assert got.shape == ref.shape, (name, combo, got.shape, ref.shape)
        assert got.dtype == ref.dtype, (name, combo, got.dtype, ref.dtype)
        if ref.dtype.is_floating_point:
            torch.testing.assert_close(got, ref, rtol=2e-5, atol=2e-5, msg=lambda m: f"{name} {combo}: {m}")
        else:
            assert torch.equal(got, ref), (name, combo)
    else:
        assert out == ref, (name, combo)


def _sweep(name, fn, tensors, mesh, limit=20):
    """Run ``fn`` over the cartesian product of placements of ``tensors`` (sub-sampled deterministically to ``limit``)."""
    from vescale_b200 import distribute_tensor

    ref = fn(*tensors)
    combos = list(itertools.product(*[_choices(t) for t in tensors]))
    step = max(1, len(combos) // limit)
    failures = []
    for combo in combos[::step]:
        try:
            dts = [distribute_tensor(t, mesh, [p], src_data_rank=None) for t, p in zip(tensors, combo)]
            _compare(fn(*dts), ref, name, combo)
        except Exception as e:  # noqa: BLE001
            failures.append(f"{name} {combo}: {type(e).__name__}: {str(e)[:300]}")
    return failures


def _op_table():
    S = (8, 12)  # even over 4 ranks on both dims
    U = (6, 10)  # uneven on both dims
    ops = []

    def add(name, fn, *tensors):
        ops.append((name, fn, tensors))

    for shp, tag in ((S, "even"), (U, "uneven")):
        x, y, z = _t(shp, 1), _t(shp, 2), _t(shp, 3)
        # ---- pointwise, unary
        for nm in ("neg", "abs", "exp", "sigmoid", "tanh", "relu", "sin", "cos", "sign", "floor", "erf", "square"):
            add(f"{nm}/{tag}", getattr(torch, nm), x)
        add(f"log/{tag}", lambda a: torch.log(a.abs() + 1), x)
        add(f"sqrt/{tag}", lambda a: torch.sqrt(a.abs()), x)
        add(f"rsqrt/{tag}", lambda a: torch.rsqrt(a.abs() + 1), x)
        add(f"reciprocal/{tag}", lambda a: torch.reciprocal(a + 3), x)
        add(f"gelu/{tag}", F.gelu, x)
        add(f"silu/{tag}", F.silu, x)
        add(f"pow_scalar/{tag}", lambda a: a.pow(2) + 2**a.clamp(-1, 1), x)
        add(f"clamp/{tag}", lambda a: a.clamp(-0.5, 0.5), x)
        add(f"to_dtype/{tag}", lambda a: a.to(torch.float64).to(torch.bfloat16).float(), x)
        add(f"isfinite/{tag}", lambda a: torch.isfinite(a) & ~torch.isnan(a), x)
        add(f"scalar_mix/{tag}", lambda a: 2 * a - 1 + a / 3, x)
        # ---- pointwise, binary / ternary
        for nm in ("add", "sub", "mul", "maximum", "minimum", "atan2"):
            add(f"{nm}/{tag}", getattr(torch, nm), x, y)
        add(f"div/{tag}", lambda a, b: a / (b.abs() + 1), x, y)
        add(f"cmp/{tag}", lambda a, b: (a > b) | (a == b), x, y)
        add(f"where/{tag}", lambda a, b, c: torch.where(a > 0, b, c), x, y, z)
        add(f"addcmul/{tag}", lambda a, b, c: torch.addcmul(a, b, c, value=0.5), x, y, z)
        add(f"addcdiv/{tag}", lambda a, b, c: torch.addcdiv(a, b, c.abs() + 1, value=0.5), x, y, z)
        add(f"lerp/{tag}", lambda a, b: torch.lerp(a, b, 0.3), x, y)
        add(f"bcast_row/{tag}", lambda a, b: a + b[0], x, y)
        add(f"bcast_col/{tag}", lambda a, b: a * b[:, :1], x, y)
        add(f"masked_fill/{tag}", lambda a, b: a.masked_fill(b > 0, -1.0), x, y)
        # ---- reductions
        add(f"sum_all/{tag}", lambda a: a.sum(), x)
        for d in (0, 1, -1):
            add(f"sum{d}/{tag}", lambda a, d=d: a.sum(d), x)
            add(f"mean{d}k/{tag}", lambda a, d=d: a.mean(d, keepdim=True), x)
            add(f"amax{d}/{tag}", lambda a, d=d: a.amax(d), x)
            add(f"amin{d}/{tag}", lambda a, d=d: a.amin(d), x)
            add(f"max{d}/{tag}", lambda a, d=d: tuple(a.max(d)), x)
            add(f"argmax{d}/{tag}", lambda a, d=d: a.argmax(d), x)
            add(f"var{d}/{tag}", lambda a, d=d: a.var(d), x)
            add(f"std{d}/{tag}", lambda a, d=d: a.std(d, unbiased=False), x)
            add(f"logsumexp{d}/{tag}", lambda a, d=d: a.logsumexp(d), x)
            add(f"cumsum{d}/{tag}", lambda a, d=d: a.cumsum(d), x)
            add(f"softmax{d}/{tag}", lambda a, d=d: a.softmax(d), x)
            add(f"log_softmax{d}/{tag}", lambda a, d=d: a.log_softmax(d), x)
            add(f"norm{d}/{tag}", lambda a, d=d: torch.linalg.vector_norm(a, 2, dim=d), x)
        add(f"prod/{tag}", lambda a: (a * 0.5 + 1).prod(1), x)
        add(f"all_any/{tag}", lambda a: (torch.all(a > -10), torch.any(a > 10), (a > 0).any(1)), x)
        add(f"minmax_all/{tag}", lambda a: (a.max(), a.min(), a.mean()), x)
        # ---- views / shape
        add(f"transpose/{tag}", lambda a: a.t().contiguous(), x)
        add(f"permute3/{tag}", lambda a: a.unsqueeze(1).permute(2, 0, 1).contiguous(), x)
        add(f"unsqueeze_squeeze/{tag}", lambda a: a.unsqueeze(0).squeeze(0).unsqueeze(-1), x)
        add(f"expand/{tag}", lambda a: a.unsqueeze(0).expand(3, *a.shape).contiguous(), x)
        add(f"slice/{tag}", lambda a: (a[1:5], a[:, 2:7], a[::2, 1::3]), x)
        add(f"select/{tag}", lambda a: (a[2], a[:, 3], a.select(1, -1)), x)
        add(f"narrow/{tag}", lambda a: a.narrow(1, 2, 5), x)
        add(f"split/{tag}", lambda a: tuple(a.split(3, dim=0)) + tuple(a.chunk(2, dim=1)), x)
        add(f"unbind/{tag}", lambda a: a.unbind(0)[1] + a.unbind(1)[2].sum(), x)
        add(f"cat/{tag}", lambda a, b: (torch.cat([a, b], 0), torch.cat([a, b], 1)), x, y)
        add(f"stack/{tag}", lambda a, b: (torch.stack([a, b], 0), torch.stack([a, b], 2)), x, y)
        add(f"flatten/{tag}", lambda a: a.flatten(), x)
        add(f"tril_triu/{tag}", lambda a: a.tril() + a.triu(1), x)
        add(f"flip/{tag}", lambda a: a.flip(0) + a.flip(1), x)
        add(f"clone_detach/{tag}", lambda a: a.clone().detach() + 0, x)
        # ---- creation-like / indexing
        add(f"like/{tag}", lambda a: torch.zeros_like(a) + torch.ones_like(a) * 2 + torch.full_like(a, 3.0) + a.new_zeros(a.shape) + a.new_ones(a.shape), x)
        sel = torch.tensor([0, 3, 1, 2], device=x.device)
        add(f"index_select/{tag}", lambda a, i: (a.index_select(0, i), a.index_select(1, i + 2)), x, sel)
        idx = _t((shp[0], 4), 7, torch.int64)
        add(f"gather/{tag}", lambda a, i: a.gather(1, i), x, idx)
        add(f"inplace/{tag}", lambda a, b: a.clone().add_(b).mul_(2).sub_(1).div_(2), x, y)
        add(f"copy_fill/{tag}", lambda a, b: (a.clone().copy_(b), a.clone().fill_(1.5), a.clone().zero_()), x, y)
    # ---- data-dependent output shapes (run replicated, spec from the result), in-place ops that need a reshard, misc
    x, y = _t((8, 12), 1), _t((8, 12), 2)
    idx = _t((8, 4), 7, torch.int64)
    add("topk_sort", lambda a: tuple(a.topk(3, dim=1)) + tuple(a.sort(dim=0)) + (a.argsort(dim=1),), x)
    add("unique", lambda a: (a > 0).to(torch.int64).sum(1).unique(), x)
    add("nonzero", lambda a: (a > 0.5).nonzero(), x)
    add("masked_select", lambda a, b: a.masked_select(b > 0), x, y)
    add("scatter_add", lambda a, i: torch.zeros_like(a).scatter_add(1, i, a[:, :4]), x, idx)
    add("scatter_inplace", lambda a, i: a.clone().scatter_(1, i, 1.0), x, idx)
    add("index_put_inplace", lambda a, i: a.clone().index_put_((i[:, 0],), a * 0 + 1.0, accumulate=True), x, idx)
    add("adv_index", lambda a, i: a[i[:, 0]], x, idx)
    add("pad_roll_repeat", lambda a: (F.pad(a, (1, 2, 3, 0)), a.roll(2, 0) + a.roll(-1, 1), a.repeat(2, 3), a.repeat_interleave(2, dim=0)), x)
    add("one_hot", lambda i: F.one_hot(i, 6), idx)
    add("conv1d", lambda a, w: F.conv1d(a.unsqueeze(0), w[:4, :8].reshape(4, 8, 1).contiguous()), x, y)
    add("cdist_outer_dot", lambda a, b: (torch.cdist(a, b), torch.outer(a[:, 0], b[:, 1]), torch.dot(a[:, 0], b[:, 0])), x, y)
    add("var_mean_median", lambda a: tuple(torch.var_mean(a, dim=1)) + (a.median(1)[0], a.kthvalue(2, 0)[0], a.cummax(1)[0], a.cumprod(0)), x)
    add("group_norm", lambda a: F.group_norm(a.view(8, 4, 3), 2), x)
    add("activations", lambda a: F.gelu(a, approximate="tanh") + F.leaky_relu(a) + F.elu(a) + F.softplus(a) + F.hardtanh(a), x)
    add("losses", lambda a, b: F.l1_loss(a, b) + F.smooth_l1_loss(a, b) + F.binary_cross_entropy_with_logits(a, (b > 0).float()), x, y)
    # ---- view / reshape (shape arguments must be localised)
    v = _t((8, 12), 4)
    add("view_split", lambda a: a.view(2, 4, 12), v)
    add("view_merge", lambda a: a.view(2, 4, 3, 4).reshape(8, 12), v)
    add("reshape_flat", lambda a: a.reshape(-1), v)
    add("reshape_3d", lambda a: a.reshape(8, 3, 4).transpose(1, 2).reshape(8, 12), v)
    add("unflatten", lambda a: a.unflatten(1, (3, 4)), v)
    # ---- matrix
    a, b, c = _t((8, 12), 5), _t((12, 16), 6), _t((8, 16), 7)
    add("mm", torch.mm, a, b)
    add("matmul2d", torch.matmul, a, b)
    add("addmm", lambda m, p, q: torch.addmm(m, p, q, beta=0.5, alpha=2.0), c, a, b)
    add("linear_bias", lambda p, w, bias: F.linear(p, w, bias), a, _t((16, 12), 8), _t((16,), 9))
    add("linear_nobias", lambda p, w: F.linear(p, w), a, _t((16, 12), 8))
    ba, bb, bc = _t((4, 8, 12), 10), _t((4, 12, 8), 11), _t((4, 8, 8), 12)
    add("bmm", torch.bmm, ba, bb)
    add("baddbmm", lambda m, p, q: torch.baddbmm(m, p, q), bc, ba, bb)
    add("matmul_bcast", torch.matmul, ba, b[:, :8].contiguous())
    add("matmul_vec", lambda p, q: torch.matmul(p, q[:, 0]), a, b)
    add("einsum", lambda p, q: torch.einsum("bik,bkj->bij", p, q), ba, bb)
    # ---- nn
    h = _t((8, 16), 13)
    add("layer_norm", lambda p, w, bb_: F.layer_norm(p, (16,), w, bb_), h, _t((16,), 14), _t((16,), 15))
    add("mse", lambda p, q: F.mse_loss(p, q), h, _t((8, 16), 16))
    add("dropout0", lambda p: F.dropout(p, 0.0, training=True), h)
    emb, ids = _t((20, 8), 17), torch.randint(0, 20, (4, 6), generator=torch.Generator().manual_seed(18)).to(device_type())
    add("embedding", lambda w, i: F.embedding(i, w), emb, ids)
    logits, tgt = _t((8, 20), 19), torch.randint(0, 20, (8,), generator=torch.Generator().manual_seed(20)).to(device_type())
    add("cross_entropy", lambda l, t_: F.cross_entropy(l, t_), logits, tgt)
    add("nll", lambda l, t_: F.nll_loss(l.log_softmax(-1), t_), logits, tgt)
    q, k, vv = _t((2, 4, 8, 8), 21), _t((2, 4, 8, 8), 22), _t((2, 4, 8, 8), 23)
    add("sdpa", lambda q_, k_, v_: F.scaled_dot_product_attention(q_, k_, v_, is_causal=True), q, k, vv)
    return ops


def _op_sweep(rank, world, part, nparts):
    from vescale_b200 import init_device_mesh

    mesh = init_device_mesh(device_type(), (world,))
    ops = _op_table()
    failures = []
    for i, (name, fn, tensors) in enumerate(ops):
        if i % nparts != part:
            continue
        failures += _sweep(name, fn, list(tensors), mesh, limit=12)
    if failures and rank == 0:
        print(f"\n{len(failures)} failing op/placement combinations (part {part}):\n  " + "\n  ".join(failures[:60]), flush=True)
    assert not failures, f"{len(failures)} failures, first: {failures[0]}"


def _grads_and_2d(rank, world):
    from vescale_b200 import Partial, Replicate, Shard, distribute_tensor, init_device_mesh

    dev = device_type()
    mesh = init_device_mesh(dev, (world,))
    mesh2 = init_device_mesh(dev, (2, 2), mesh_dim_names=("a", "b"))
    x, w1, w2 = _t((8, 12), 1), _t((12, 16), 2), _t((16, 4), 3)

    def model(x_, w1_, w2_):
        hdn = F.gelu(x_ @ w1_)
        hdn = F.layer_norm(hdn, (16,))
        return ((hdn @ w2_).softmax(-1) * 2).sum()

    xr, w1r, w2r = (t.clone().requires_grad_() for t in (x, w1, w2))
    model(xr, w1r, w2r).backward()
    cases = [
        (mesh, [Shard(0)], [Replicate()], [Replicate()]),  # data parallel
        (mesh, [Replicate()], [Shard(1)], [Shard(0)]),  # tensor parallel (column -> row)
        (mesh, [Shard(0)], [Shard(1)], [Replicate()]),
        (mesh2, [Shard(0), Replicate()], [Replicate(), Shard(1)], [Replicate(), Shard(0)]),  # dp x tp
        (mesh2, [Shard(0), Shard(1)], [Replicate(), Replicate()], [Shard(1), Replicate()]),
    ]
    for m, px, p1, p2 in cases:
        dx, d1, d2 = (distribute_tensor(t, m, p, src_data_rank=None).requires_grad_() for t, p in ((x, px), (w1, p1), (w2, p2)))
        out = model(dx, d1, d2)
        out.redistribute(m, [Replicate()] * m.ndim).backward()
        for got, ref in ((dx, xr), (d1, w1r), (d2, w2r)):
            assert got.grad is not None, (px, p1, p2)
            torch.testing.assert_close(got.grad.full_tensor(), ref.grad, rtol=1e-4, atol=1e-5, msg=lambda s: f"{(px, p1, p2)}: {s}")
    # 2-D mesh forward sweep on a few ops
    a, b = _t((8, 12), 5), _t((12, 8), 6)
    pls = [[Replicate(), Replicate()], [Shard(0), Shard(1)], [Shard(1), Shard(0)], [Shard(0), Shard(0)], [Replicate(), Shard(1)]]
    for pa, pb in itertools.product(pls, pls):
        da, db = distribute_tensor(a, mesh2, pa, src_data_rank=None), distribute_tensor(b, mesh2, pb, src_data_rank=None)
        torch.testing.assert_close((da @ db).full_tensor(), a @ b, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close((da + db.t()).full_tensor(), a + b.t())
        torch.testing.assert_close(da.sum(0).full_tensor(), a.sum(0), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(torch.cat([da, db.t()], 1).full_tensor(), torch.cat([a, b.t()], 1))


def _view_2d_and_inplace_partial(rank, world):
    """(B,S,H)->(B*S,H) views under a DP x SP mesh ([Shard(0), Shard(1)]) and in-place ops on Partial tensors."""
    from vescale_b200 import Partial, Replicate, Shard, distribute_tensor, init_device_mesh
    from vescale_b200.dtensor import DTensor

    dev = device_type()
    mesh2 = init_device_mesh(dev, (2, 2), mesh_dim_names=("dp", "sp"))
    x, w = _t((4, 8, 3), 11), _t((5, 3), 12)
    y = torch.randint(0, 5, (4, 8), generator=torch.Generator().manual_seed(3)).to(dev)
    for pl in ([Shard(0), Shard(1)], [Shard(1), Shard(2)], [Replicate(), Shard(1)], [Shard(1), Shard(0)], [Shard(0), Shard(0)]):
        dx = distribute_tensor(x, mesh2, pl, src_data_rank=None)
        torch.testing.assert_close(dx.view(32, 3).full_tensor(), x.view(32, 3), msg=lambda m: f"view {pl}: {m}")
        torch.testing.assert_close(dx.reshape(-1).full_tensor(), x.reshape(-1), msg=lambda m: f"flatten {pl}: {m}")
        dw = distribute_tensor(w, mesh2, [Replicate(), Replicate()], src_data_rank=None)
        torch.testing.assert_close(torch.mm(dx.view(32, 3), dw.t()).full_tensor(), x.view(32, 3) @ w.t(), rtol=1e-5, atol=1e-5)
        xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
        ref = F.cross_entropy(F.linear(xr, wr).view(-1, 5), y.view(-1))
        ref.backward()
        dxg = distribute_tensor(x, mesh2, pl, src_data_rank=None).requires_grad_()
        dwg = distribute_tensor(w, mesh2, [Replicate(), Replicate()], src_data_rank=None).requires_grad_()
        dy = distribute_tensor(y, mesh2, [Replicate(), Replicate()], src_data_rank=None)
        loss = F.cross_entropy(F.linear(dxg, dwg).view(-1, 5), dy.view(-1))
        torch.testing.assert_close(loss.full_tensor(), ref.detach(), rtol=1e-5, atol=1e-5, msg=lambda m: f"loss {pl}: {m}")
        loss.redistribute(mesh2, [Replicate(), Replicate()]).backward()
        torch.testing.assert_close(dwg.grad.full_tensor(), wr.grad, rtol=1e-4, atol=1e-5, msg=lambda m: f"dW {pl}: {m}")
        torch.testing.assert_close(dxg.grad.full_tensor(), xr.grad, rtol=1e-4, atol=1e-5, msg=lambda m: f"dX {pl}: {m}")
    # in-place ops on a Partial(sum) tensor: non-linear / scalar ops must reduce first (or raise), linear ones may stay Partial
    mesh = init_device_mesh(dev, (world,))
    a, b = _t((6, 8), 21), _t((8, 4), 22)
    ref = a @ b
    for name, fn in (("relu_", lambda t: t.relu_()), ("add_1", lambda t: t.add_(1.0)), ("clamp_", lambda t: t.clamp_(min=0.1)), ("mul_2", lambda t: t.mul_(2.0)), ("neg_", lambda t: t.neg_())):
        o = distribute_tensor(a, mesh, [Shard(1)], src_data_rank=None) @ distribute_tensor(b, mesh, [Shard(0)], src_data_rank=None)
        assert any(isinstance(p, Partial) for p in o.placements)
        got = fn(o)
        torch.testing.assert_close(got.full_tensor(), fn(ref.clone()), rtol=1e-5, atol=1e-5, msg=lambda m: f"{name}: {m}")
        torch.testing.assert_close(o.full_tensor(), fn(ref.clone()), rtol=1e-5, atol=1e-5, msg=lambda m: f"{name} (self): {m}")
    # neg of Partial(max) must not stay Partial(max)
    loc = _t((4, 4), 30 + rank)
    pm = DTensor.from_local(loc, mesh, [Partial("max")], run_check=False)
    full = pm.full_tensor()
    torch.testing.assert_close((-pm).full_tensor(), -full)


def test_view_2d_mesh_and_inplace_partial():
    run_distributed(_view_2d_and_inplace_partial, 4)


def test_op_sweep_part0():
    run_distributed(_op_sweep, 4, 0, 2)


def test_op_sweep_part1():
    run_distributed(_op_sweep, 4, 1, 2)


def test_autograd_and_2d_mesh():
    run_distributed(_grads_and_2d, 4)


def _conv_rules(rank, world):
    """Convolution rules (legacy ``ops/conv_ops.py``): batch-parallel and output-channel-parallel conv2d keep their shards,
    weight gradients of the batch-parallel form are Partial, everything matches single-device forward and backward."""
    from vescale_b200 import Replicate, Shard, distribute_tensor, init_device_mesh

    mesh = init_device_mesh(device_type(), (world,))
    g = torch.Generator().manual_seed(0)
    x, w, b = torch.randn(8, 3, 10, 10, generator=g).to(device_type()), torch.randn(2 * world, 3, 3, 3, generator=g).to(device_type()), torch.randn(2 * world, generator=g).to(device_type())
    xr, wr, br = (t.clone().requires_grad_() for t in (x, w, b))
    yr = F.conv2d(xr, wr, br, stride=1, padding=1)
    yr.square().sum().backward()
    for px, pw, want in (([Shard(0)], [Replicate()], Shard(0)), ([Replicate()], [Shard(0)], Shard(1)), ([Replicate()], [Replicate()], Replicate()), ([Shard(2)], [Replicate()], None)):
        dx = distribute_tensor(x, mesh, px, src_data_rank=None).requires_grad_()
        dw = distribute_tensor(w, mesh, pw, src_data_rank=None).requires_grad_()
        db = distribute_tensor(b, mesh, pw, src_data_rank=None).requires_grad_()
        y = F.conv2d(dx, dw, db, stride=1, padding=1)
        if want is not None:
            assert y.placements == (want,), (px, pw, y.placements)
        torch.testing.assert_close(y.full_tensor(), yr.detach(), rtol=1e-4, atol=1e-5)
        y.square().sum().redistribute(mesh, [Replicate()]).backward()
        torch.testing.assert_close(dx.grad.full_tensor(), xr.grad, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(dw.grad.full_tensor(), wr.grad, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(db.grad.full_tensor(), br.grad, rtol=1e-4, atol=1e-4)
        if px == [Shard(0)]:
            assert dw.grad.placements[0].is_partial(), dw.grad.placements


def test_conv_rules():
    run_distributed(_conv_rules, 2)
