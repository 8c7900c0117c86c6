"""DDP + DistributedOptimizer / BasicOptimizer vs single-process golden (2 epochs x 2 micro-batches),
parametrised like ``legacy/test/parallel/ddp_optim/test_doptimizer.py:51-80``."""
import copy
import os

import pytest
import torch
import torch.distributed as dist
import torch.nn as nn

from common import device_type, run_distributed


class MLP(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(16, 64)
        self.n = nn.LayerNorm(64)
        self.b = nn.Linear(64, 16)

    def forward(self, x):
        return self.b(self.n(torch.relu(self.a(x))))


def _data(step, mb, rank):
    g = torch.Generator().manual_seed(1000 * step + 10 * mb + rank)
    return torch.randn(4, 16, generator=g), torch.randn(4, 16, generator=g)


def _run(rank, world, use_dopt, overlap_grad, overlap_gather):
    from vescale_b200.optim import BasicOptimizer, DistributedOptimizer
    from vescale_b200.parallel.ddp import DistributedDataParallel as DDP

    dev = device_type()
    torch.manual_seed(0)
    ref = MLP().to(dev)
    model = copy.deepcopy(ref)
    ref_opt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.05)
    ddp = DDP(model, dist.group.WORLD, overlap_grad_reduce=overlap_grad, use_distributed_optimizer=use_dopt, bucket_size=600)
    inner = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.05)
    opt = DistributedOptimizer(inner, [ddp], clip_grad=1.0, overlap_param_gather=overlap_gather) if use_dopt else BasicOptimizer(inner, [ddp], clip_grad=1.0)
    for step in range(2):
        ref_opt.zero_grad()
        for mb in range(2):
            for r in range(world):
                x, y = _data(step, mb, r)
                (torch.nn.functional.mse_loss(ref(x.to(dev)), y.to(dev)) / (2 * world)).backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        ref_opt.step()
        opt.zero_grad()
        for mb in range(2):
            x, y = _data(step, mb, rank)
            ctx = ddp.no_sync() if mb == 0 else torch.enable_grad()
            with ctx:
                (torch.nn.functional.mse_loss(ddp(x.to(dev)), y.to(dev)) / 2).backward()
        opt.step()
        if use_dopt:
            opt.finish_param_gather()
    for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=2e-4, atol=2e-5, msg=n)
    if use_dopt:
        sd = opt.state_dict()
        assert sd["state"] and all("exp_avg" in e and "main" in e for ents in sd["state"].values() for e in ents)
        opt.load_state_dict(sd)


@pytest.mark.parametrize("use_dopt,overlap_grad,overlap_gather", [(False, True, False), (True, True, False), (True, False, True), (True, True, True)])
def test_ddp_optimizers_match_single_process(use_dopt, overlap_grad, overlap_gather):
    run_distributed(_run, 4, use_dopt, overlap_grad, overlap_gather)


class Tied(nn.Module):
    def __init__(self):
        super().__init__()
        self.emb = nn.Embedding(20, 16)
        self.mid = nn.Linear(16, 16)
        self.head = nn.Linear(16, 20, bias=False)
        self.head.weight = self.emb.weight

    def forward(self, ids):
        return self.head(torch.tanh(self.mid(self.emb(ids))))


def _tied(rank, world):
    """Tied embedding / lm-head under DDP with BasicOptimizer and with DistributedOptimizer (legacy ``test_shared_weight.py``)."""
    from vescale_b200.optim import BasicOptimizer, DistributedOptimizer
    from vescale_b200.parallel.ddp import DistributedDataParallel as DDP

    dev = device_type()
    for use_dopt in (False, True):
        torch.manual_seed(0)
        ref = Tied().to(dev)
        model = copy.deepcopy(ref)
        assert model.head.weight is model.emb.weight
        ref_opt = torch.optim.SGD(ref.parameters(), lr=0.1)
        ddp = DDP(model, dist.group.WORLD, overlap_grad_reduce=True, use_distributed_optimizer=use_dopt, bucket_size=300)
        inner = torch.optim.SGD(model.parameters(), lr=0.1)
        opt = DistributedOptimizer(inner, [ddp]) if use_dopt else BasicOptimizer(inner, [ddp])
        for step in range(2):
            ref_opt.zero_grad()
            for r in range(world):
                g = torch.Generator().manual_seed(10 * step + r)
                ids = torch.randint(0, 20, (4, 6), generator=g).to(dev)
                tgt = torch.randint(0, 20, (4, 6), generator=g).to(dev)
                (torch.nn.functional.cross_entropy(ref(ids).view(-1, 20), tgt.view(-1)) / world).backward()
            ref_opt.step()
            g = torch.Generator().manual_seed(10 * step + rank)
            ids = torch.randint(0, 20, (4, 6), generator=g).to(dev)
            tgt = torch.randint(0, 20, (4, 6), generator=g).to(dev)
            opt.zero_grad()
            torch.nn.functional.cross_entropy(ddp(ids).view(-1, 20), tgt.view(-1)).backward()
            opt.step()
            if use_dopt:
                opt.finish_param_gather()
        for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-4, atol=1e-6, msg=f"{use_dopt} {n}")
        assert model.head.weight.data_ptr() == model.emb.weight.data_ptr()


def test_ddp_tied_weights():
    run_distributed(_tied, 4)


def _clip_model_parallel(rank, world):
    """the clip norm counts TP-replicated gradients once and TP-sharded ones over the TP mesh."""
    import torch

    from common import device_type
    from vescale_b200 import Replicate, Shard, distribute_tensor, init_device_mesh
    from vescale_b200.optim import clip_grad_norm_fp32, get_grad_norm_fp32

    mesh = init_device_mesh(device_type(), (2, 2), mesh_dim_names=("DP", "TP"))
    tp = mesh["TP"]
    g = torch.Generator().manual_seed(7)
    gw, gn, gp = torch.randn(8, 6, generator=g), torch.randn(6, generator=g), torch.randn(5, generator=g)
    ref = torch.cat([gw.flatten(), gn, gp]).norm()
    dw = distribute_tensor(gw.to(device_type()), tp, [Shard(0)], src_data_rank=None)
    dn = distribute_tensor(gn.to(device_type()), tp, [Replicate()], src_data_rank=None)
    plain = gp.to(device_type())
    got = get_grad_norm_fp32([dw, dn, plain])
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-5, atol=1e-5)
    params = [torch.nn.Parameter(t.clone()) for t in (dw, dn)]
    for p, t in zip(params, (dw, dn)):
        p.grad = t.clone()
    total = clip_grad_norm_fp32(params, max_norm=0.5)
    expect = torch.cat([gw.flatten(), gn]).norm()
    torch.testing.assert_close(total.cpu(), expect, rtol=1e-5, atol=1e-5)
    coef = 0.5 / (expect + 1e-6)
    torch.testing.assert_close(params[1].grad.full_tensor().cpu(), gn * coef, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(params[0].grad.full_tensor().cpu(), gw * coef, rtol=1e-5, atol=1e-6)


def test_clip_grad_norm_model_parallel():
    run_distributed(_clip_model_parallel, 4)


def _dmodule_ddp_dopt(rank, world, tmp):
    """TP-sharded DModule parameters (DTensors) under DDP + DistributedOptimizer: (1) the optimizer really updates the shards the
    model computes with (the parameter object's local tensor is re-pointed into the flat parameter buffer) and the trajectory equals
    the single-device one; (2) its checkpoint state — flat ranges of TP-local shards — is written as boxes of the global tensors
    (``checkpoint/flat_piece.py``) and reloads exactly, under the same layout and under a different DP x TP factorisation."""
    import torch
    import torch.nn as nn

    import vescale_b200.checkpoint as ckpt
    from vescale_b200 import Replicate, Shard, init_device_mesh
    from vescale_b200.dtensor import DTensor
    from vescale_b200.optim import DistributedOptimizer
    from vescale_b200.parallel.ddp import DistributedDataParallel as DDP
    from vescale_b200.parallel.dmodule import parallelize_module

    dev = device_type()

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.n = nn.Linear(8, 16, bias=False), nn.Linear(16, 8, bias=False), nn.LayerNorm(8)

        def forward(self, x):
            return self.n(self.b(torch.relu(self.a(x))))

    plan = {"parameter": {r"a\.weight": [Shard(0)], r"b\.weight": [Shard(1)]}, "forward": {r"input": [[Replicate()]], r"b\.output": [[Replicate()]]}}
    x = torch.randn(4, 8, generator=torch.Generator().manual_seed(3)).to(dev)

    def build(seed, dp, tp):
        mesh = init_device_mesh(dev, (dp, tp), mesh_dim_names=("DP", "TP"))
        torch.manual_seed(seed)
        m = M().to(dev)
        single = copy.deepcopy(m)
        parallelize_module(m, mesh["TP"], plan)
        ddp = DDP(m, mesh["DP"].get_group(0), use_distributed_optimizer=True, overlap_grad_reduce=True)
        opt = DistributedOptimizer(torch.optim.AdamW(m.parameters(), lr=1e-1), [ddp], clip_grad=1.0, overlap_param_gather=False)
        return m, ddp, opt, single

    def step(m, ddp, opt):
        opt.zero_grad()
        out = ddp(x)
        out = out.to_local() if isinstance(out, DTensor) else out
        loss = out.pow(2).mean()
        loss.backward()
        m.finish_grad_sync()
        opt.step()
        return float(loss)

    m, ddp, opt, single = build(0, 2, 2)
    sopt = torch.optim.AdamW(single.parameters(), lr=1e-1)
    for _ in range(2):
        got = step(m, ddp, opt)
        sopt.zero_grad()
        sl = single(x).pow(2).mean()
        sl.backward()
        torch.nn.utils.clip_grad_norm_(single.parameters(), 1.0)
        sopt.step()
        assert abs(got - float(sl)) < 2e-4, (got, float(sl))
    # (weights are not compared element-wise: AdamW at lr 0.1 turns the sign of a ~1e-9 gradient into a +-0.1 update; the loss of the
    # second step equals the single-device one only if the first update was applied to the shards the model computes with)
    path = os.path.join(tmp, "ck")
    ckpt.save(path, {"model": ddp, "optimizer": opt})
    cont = step(m, ddp, opt)
    cont2 = step(m, ddp, opt)
    for dp, tp in ((2, 2), (1, 4), (4, 1)):
        m2, ddp2, opt2, _ = build(1, dp, tp)
        ckpt.load(path, {"model": ddp2, "optimizer": opt2})
        r1, r2 = step(m2, ddp2, opt2), step(m2, ddp2, opt2)
        assert abs(r1 - cont) < 1e-6 and abs(r2 - cont2) < 1e-5, ((dp, tp), r1, cont, r2, cont2)


def test_dmodule_params_under_ddp_distributed_optimizer_and_checkpoint(tmp_path):
    run_distributed(_dmodule_ddp_dopt, 4, str(tmp_path))
