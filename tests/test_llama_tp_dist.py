"""2-D Llama (FSDP over dp x TP/SP over tp) vs single-process training: same data, same optimizer, compare losses.
Strategy parity: ``legacy/test/model/open_llama/test_attention.py`` / ``test_decoder_layer.py`` (TP+SP module vs golden) and
the 4-D example's loss-curve match (``legacy/examples/llama2_4D_finetune/README``)."""
import torch
import torch.distributed as dist

from common import device_type, run_distributed


def _train_ref(cfg, steps, dp, lr, wd, clip):
    from vescale_b200.models import LlamaModel

    m = LlamaModel(cfg).reset_parameters(seed=1)
    decay = [p for p in m.parameters() if p.ndim > 1]
    nodecay = [p for p in m.parameters() if p.ndim <= 1]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": wd}, {"params": nodecay, "weight_decay": 0.0}], lr=lr, betas=(0.9, 0.95), eps=1e-8)
    losses, norms = [], []
    for s in range(steps):
        opt.zero_grad()
        tot = 0.0
        for r in range(dp):
            g = torch.Generator().manual_seed(100 * s + r)
            tok = torch.randint(0, cfg.vocab_size, (2, 16), generator=g)
            lab = torch.randint(0, cfg.vocab_size, (2, 16), generator=g)
            loss = m(tok, lab) / dp
            loss.backward()
            tot += loss.item()
        norms.append(torch.nn.utils.clip_grad_norm_(m.parameters(), clip).item())
        opt.step()
        losses.append(tot)
    return losses, norms


def _tp_fsdp(rank, world, tp_size):
    from vescale_b200 import init_device_mesh
    from vescale_b200.comm.fused_tp import PlainTP
    from vescale_b200.models import LlamaConfig, LlamaModel
    from vescale_b200.models.llama_tp import LlamaTPModel, shard_llama_state_for_tp
    from vescale_b200.optim import FSDPAdamW
    from vescale_b200.parallel.fsdp import MixedPrecisionPolicy, fully_shard

    dev = device_type()
    cfg = LlamaConfig.tiny()
    dp = world // tp_size
    lr, wd, clip, steps = 1e-2, 0.1, 1.0, 3
    ref_losses, ref_norms = _train_ref(cfg, steps, dp, lr, wd, clip)
    mesh = init_device_mesh(dev, (dp, tp_size), mesh_dim_names=("dp", "tp"))
    tp = PlainTP(mesh, "tp")
    full = LlamaModel(cfg).reset_parameters(seed=1)
    model = LlamaTPModel(cfg, tp)
    model.load_state_dict(shard_llama_state_for_tp(full.state_dict(), cfg, tp.rank, tp.world))
    model = model.to(dev)
    mp = MixedPrecisionPolicy(param_dtype=torch.float32, reduce_dtype=torch.float32)
    for blk in model.layers:
        fully_shard(blk, mesh, mesh_dim="dp", mp_policy=mp)
    fully_shard(model.embed, mesh, mesh_dim="dp", mp_policy=mp)
    fully_shard(model.head, mesh, mesh_dim="dp", mp_policy=mp)
    fully_shard(model, mesh, mesh_dim="dp", mp_policy=mp)
    opt = FSDPAdamW(model, lr=lr, betas=(0.9, 0.95), eps=1e-8, weight_decay=wd, max_grad_norm=clip, tp_group=mesh.get_group("tp"))
    dp_rank = mesh.get_local_rank("dp")
    for s in range(steps):
        g = torch.Generator().manual_seed(100 * s + dp_rank)
        tok = torch.randint(0, cfg.vocab_size, (2, 16), generator=g).to(dev)
        lab = torch.randint(0, cfg.vocab_size, (2, 16), generator=g).to(dev)
        share = model(tok, lab)
        share.backward()
        norm = opt.step()
        opt.zero_grad()
        loss = model.loss_for_logging(share)
        if dp > 1:
            dist.all_reduce(loss, group=mesh.get_group("dp"))
            loss /= dp
        assert abs(loss.item() - ref_losses[s]) < 2e-4, (s, loss.item(), ref_losses[s])
        assert abs(norm.item() - ref_norms[s]) < 2e-3 * max(1.0, ref_norms[s]), (s, norm.item(), ref_norms[s])


def test_llama_tp2_fsdp2_matches_single_process():
    run_distributed(_tp_fsdp, 4, 2)


def test_llama_tp2_only_matches_single_process():
    run_distributed(_tp_fsdp, 2, 2)
