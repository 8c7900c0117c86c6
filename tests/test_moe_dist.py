"""Expert parallelism: EP=4 Mixtral-tiny must match the same model with all experts on one device
(legacy ``test/parallel/ddp_optim/test_moe.py`` / ``test/model/mixtral`` strategy)."""
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
