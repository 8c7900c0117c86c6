"""Finetune an unmodified HuggingFace ``LlamaForCausalLM`` with DP x TP(+SP): DModule sharding plan on the TP mesh dimension,
DDP + DistributedOptimizer (ZeRO-2+) on the DP dimension, cosine learning-rate schedule with warm-up, gradient clipping, periodic
evaluation on the validation split, periodic (asynchronous) checkpoints and exact resume — the reference's
``legacy/examples/llama2_4D_finetune/llama_train.py`` experiment.

    # 4 GPUs (or 4 CPU processes): dp2 x tp2, sequence parallel
    torchrun --standalone --nproc-per-node 4 examples/llama_4D_finetune/llama_train.py --dp 2 --tp 2 --max_iters 50
    # the single-device baseline of the same run (the curves must coincide; exp.py overlays them)
    python examples/llama_4D_finetune/llama_train.py --dp 1 --tp 1 --max_iters 50

Weights: ``--hf_path`` loads a local HuggingFace checkpoint directory (e.g. a downloaded ``open_llama_3b``); without it the model is
randomly initialised from ``--config`` (``tiny`` / ``open_llama_3b`` / ``llama2_7b`` shapes; there is no network in the build
environment).  Data: ``--data_dir`` with nanoGPT-style ``train.bin`` / ``val.bin`` (files prepared for the reference load
unchanged); missing files are prepared from the built-in character corpus (``vescale_b200.data.prepare_char_corpus``).
"""
import argparse
import inspect
import math
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
import vescale_b200.checkpoint as ckpt  # noqa: E402
from sharding_plan import llama_plan  # noqa: E402
from vescale_b200.data import DistributedTokenLoader, TokenBinDataset, prepare_char_corpus  # noqa: E402
from vescale_b200.devicemesh_api import VESCALE_DEVICE_MESH  # noqa: E402
from vescale_b200.dtensor import implicit_replication  # noqa: E402
from vescale_b200.dtensor.random import manual_seed  # noqa: E402
from vescale_b200.optim import DistributedOptimizer  # noqa: E402
from vescale_b200.parallel.ddp import DistributedDataParallel as DDP  # noqa: E402
from vescale_b200.parallel.dmodule import parallelize_module  # noqa: E402
from vescale_b200.utils import model_tflops  # noqa: E402

CONFIGS = {
    "tiny": dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2),
    "small": dict(hidden_size=512, intermediate_size=1376, num_hidden_layers=8, num_attention_heads=8, num_key_value_heads=8),
    "open_llama_3b": dict(hidden_size=3200, intermediate_size=8640, num_hidden_layers=26, num_attention_heads=32, num_key_value_heads=32),
    "llama2_7b": dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=32),
}


def build_model(a, vocab_size: int):
    from transformers import LlamaConfig, LlamaForCausalLM

    if a.hf_path:
        return LlamaForCausalLM.from_pretrained(a.hf_path, torch_dtype=a.ptdtype, attn_implementation="sdpa")
    cfg = LlamaConfig(vocab_size=vocab_size, max_position_embeddings=max(a.seqlen, 64), attn_implementation="sdpa", tie_word_embeddings=False, **CONFIGS[a.config])
    return LlamaForCausalLM(cfg).to(a.ptdtype)


def configure_optimizer(model, lr, weight_decay, betas, cuda):
    """Weights of matmuls and embeddings decay, norms and biases do not (the reference's grouping)."""
    params = {n: p for n, p in model.named_parameters() if p.requires_grad}
    decay = [p for p in params.values() if p.dim() >= 2]
    no_decay = [p for p in params.values() if p.dim() < 2]
    extra = dict(fused=True) if cuda and "fused" in inspect.signature(torch.optim.AdamW).parameters else {}
    return torch.optim.AdamW([{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}], lr=lr, betas=betas, **extra)


def lr_at(it, a):
    if it < a.warmup_iters:
        return a.lr * (it + 1) / a.warmup_iters
    if it >= a.lr_decay_iters:
        return a.min_lr
    r = (it - a.warmup_iters) / max(1, a.lr_decay_iters - a.warmup_iters)
    return a.min_lr + 0.5 * (1.0 + math.cos(math.pi * r)) * (a.lr - a.min_lr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dp", type=int, default=None)
    ap.add_argument("--tp", type=int, default=None)
    ap.add_argument("--no_sp", action="store_true", help="tensor parallel only (activations replicated between blocks)")
    ap.add_argument("--config", default="tiny", choices=sorted(CONFIGS))
    ap.add_argument("--hf_path", default=None)
    ap.add_argument("--dtype", default=None, choices=["float32", "bfloat16"])
    ap.add_argument("--data_dir", default=os.path.join(HERE, "data", "synthetic_char"))
    ap.add_argument("--bsz", type=int, default=8, help="global batch")
    ap.add_argument("--seqlen", type=int, default=32)
    ap.add_argument("--max_iters", type=int, default=30)
    ap.add_argument("--lr", type=float, default=3e-3)
    ap.add_argument("--min_lr", type=float, default=3e-4)
    ap.add_argument("--warmup_iters", type=int, default=3)
    ap.add_argument("--lr_decay_iters", type=int, default=1000)
    ap.add_argument("--weight_decay", type=float, default=0.1)
    ap.add_argument("--grad_clip", type=float, default=1.0)
    ap.add_argument("--eval_interval", type=int, default=0)
    ap.add_argument("--eval_iters", type=int, default=4)
    ap.add_argument("--save_interval", type=int, default=0)
    ap.add_argument("--ckpt_dir", default=None)
    ap.add_argument("--resume", action="store_true")
    ap.add_argument("--async_checkpoint", action="store_true")
    ap.add_argument("--no_DO", action="store_true", help="plain DDP all-reduce + the base optimizer instead of the DistributedOptimizer")
    ap.add_argument("--log_file", default=None)
    a = ap.parse_args()
    cuda = torch.cuda.is_available()
    distributed = "RANK" in os.environ
    if distributed:
        dist.init_process_group("nccl" if cuda else "gloo")
    rank, ws = (dist.get_rank(), dist.get_world_size()) if distributed else (0, 1)
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = "cuda" if cuda else "cpu"
    a.ptdtype = {"float32": torch.float32, "bfloat16": torch.bfloat16}[a.dtype or ("bfloat16" if cuda else "float32")]
    a.tp = a.tp or (ws if a.dp is None else ws // a.dp)
    a.dp = a.dp or ws // a.tp
    assert a.dp * a.tp == ws, f"dp {a.dp} x tp {a.tp} != world size {ws}"

    # ---- data (rank 0 prepares missing files, everybody waits)
    if rank == 0:
        prepare_char_corpus(a.data_dir)
    if distributed:
        dist.barrier()
    meta = prepare_char_corpus(a.data_dir)
    vocab = int(meta["vocab_size"])
    vocab_padded = (vocab + 63) // 64 * 64

    torch.manual_seed(0)
    model = build_model(a, vocab_padded).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    parallel = ws > 1
    if parallel:
        mesh = VESCALE_DEVICE_MESH.init_device_mesh(dev, (a.dp, a.tp), mesh_dim_names=("DP", "TP"))
        manual_seed(0, mesh)
        parallelize_module(model, mesh["TP"], llama_plan(sequence_parallel=not a.no_sp))
        ddp = DDP(model, mesh["DP"].get_group(0), use_distributed_optimizer=not a.no_DO, overlap_grad_reduce=True)
        base = configure_optimizer(model, a.lr, a.weight_decay, (0.9, 0.95), cuda)
        opt = DistributedOptimizer(base, [ddp], clip_grad=a.grad_clip, overlap_param_gather=False) if not a.no_DO else base
        dp_rank, dp_group = mesh.get_local_rank("DP"), mesh["DP"].get_group(0)
    else:
        ddp, opt = model, configure_optimizer(model, a.lr, a.weight_decay, (0.9, 0.95), cuda)
        dp_rank, dp_group = 0, None

    train = DistributedTokenLoader(TokenBinDataset(os.path.join(a.data_dir, "train.bin")), a.seqlen, a.bsz, dp_rank=dp_rank, dp_size=a.dp, device=dev, seed=1337)
    val = DistributedTokenLoader(TokenBinDataset(os.path.join(a.data_dir, "val.bin")), a.seqlen, a.bsz, dp_rank=dp_rank, dp_size=a.dp, device=dev, seed=4242, split="val")

    start = 0
    if a.resume and a.ckpt_dir and os.path.exists(os.path.join(a.ckpt_dir, "iter.txt")):
        state = {"model": ddp, "optimizer": opt} if parallel else {"model": model}
        ckpt.load(a.ckpt_dir, state)
        if not parallel:
            opt.load_state_dict(torch.load(os.path.join(a.ckpt_dir, "optim_single.pt")))
        with open(os.path.join(a.ckpt_dir, "iter.txt")) as f:
            start = int(f.read())
        train.load_state_dict({"step": start})
        if rank == 0:
            print(f"resumed from {a.ckpt_dir} at iteration {start}", flush=True)

    def loss_of(x, y):
        with implicit_replication():  # HF builds position ids / masks as plain tensors next to DTensor activations
            logits = ddp(input_ids=x).logits if parallel else model(input_ids=x).logits
        logits = logits.to_local() if hasattr(logits, "to_local") else logits
        return F.cross_entropy(logits.float().view(-1, logits.shape[-1]), y.reshape(-1))

    def dp_mean(t):
        t = t.detach().clone()
        if dp_group is not None:
            dist.all_reduce(t, group=dp_group)
        return float(t) / a.dp

    @torch.no_grad()
    def evaluate():
        model.eval()
        out = {}
        for name, loader in (("train", train), ("val", val)):
            out[name] = sum(dp_mean(loss_of(*loader.get_batch(10_000_000 + k))) for k in range(a.eval_iters)) / a.eval_iters
        model.train()
        return out

    log = open(a.log_file, "a") if (a.log_file and rank == 0) else None

    def say(msg):
        if rank == 0:
            print(msg, flush=True)
            if log:
                log.write(msg + "\n")
                log.flush()

    say(f"llama {a.config}: {n_params / 1e6:.2f} M parameters, vocab {vocab} (padded {vocab_padded}), dp {a.dp} x tp {a.tp}{'' if a.no_sp or not parallel else ' + sp'}, {a.ptdtype}")
    t_last = time.time()
    for it in range(start, a.max_iters):
        if a.eval_interval and it % a.eval_interval == 0:
            ev = evaluate()
            say(f"eval {it}: train loss {ev['train']:.4f}, val loss {ev['val']:.4f}")
        lr = lr_at(it, a)
        for g in (opt.optimizer.param_groups if hasattr(opt, "optimizer") else opt.param_groups):
            g["lr"] = lr
        x, y = next(train)
        opt.zero_grad()
        loss = loss_of(x, y)
        with implicit_replication():
            loss.backward()
        if parallel:
            model.finish_grad_sync()
            if a.no_DO and a.grad_clip:
                from vescale_b200.optim import clip_grad_norm_fp32

                clip_grad_norm_fp32(list(model.parameters()), a.grad_clip)
        elif a.grad_clip:
            torch.nn.utils.clip_grad_norm_(model.parameters(), a.grad_clip)
        opt.step()
        now = time.time()
        tok_s = a.bsz * a.seqlen / max(now - t_last, 1e-9)
        t_last = now
        say(f"iter {it}: loss {dp_mean(loss):.4f}, lr {lr:.2e}, {tok_s:,.0f} tok/s, {model_tflops(tok_s, 6 * n_params, ws):.4f} TFLOPS/device")
        if a.save_interval and a.ckpt_dir and (it + 1) % a.save_interval == 0:
            state = {"model": ddp, "optimizer": opt} if parallel else {"model": model}
            ckpt.save(a.ckpt_dir, state, async_checkpoint=a.async_checkpoint)
            if not parallel:
                torch.save(opt.state_dict(), os.path.join(a.ckpt_dir, "optim_single.pt"))
            if rank == 0:
                with open(os.path.join(a.ckpt_dir, "iter.txt"), "w") as f:
                    f.write(str(it + 1))
            if distributed:
                dist.barrier()
    if a.eval_interval:
        ev = evaluate()
        say(f"eval {a.max_iters}: train loss {ev['train']:.4f}, val loss {ev['val']:.4f}")
    ckpt.wait_for_async()
    train.close()
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
