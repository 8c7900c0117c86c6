"""DDP + DistributedOptimizer / BasicOptimizer vs single-process golden (2 epochs x 2 micro-batches),
parametrised like ``legacy/test/parallel/ddp_optim/test_doptimizer.py:51-80``."""
import copy

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
