"""Train a small Mixtral (sparse MoE, token-choice top-2 routing) with DP x TP(+SP): DModule sharding plan on the TP dim, DDP +
DistributedOptimizer (ZeRO-2+) on the DP dim, constant learning rate, gradient clipping at 1, periodic checkpoints and resume.
Reference: ``legacy/examples/mixtral_4D_training/mixtral_train.py`` + ``sharding_plan.py`` (Shakespeare characters, loss curve of
the 4-GPU run laid over the 1-GPU run).  There is no dataset download here: the corpus is a deterministic character stream with
real structure (``make_corpus``), and ``--compare-single`` trains an unparallelised twin on the global batch next to the parallel
model and checks that the two loss curves agree — the experiment of the reference's README as an assertion.

    torchrun --standalone --nproc-per-node 4 examples/mixtral_4D_training/mixtral_train.py --dp 2 --tp 2 --max_iters 20 --compare-single
"""
import argparse
import copy
import importlib.util
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import vescale_b200.checkpoint as ckpt  # noqa: E402
from vescale_b200 import Replicate, init_device_mesh  # noqa: E402
from vescale_b200.optim import DistributedOptimizer  # noqa: E402
from vescale_b200.parallel.ddp import DistributedDataParallel as DDP  # noqa: E402
from vescale_b200.parallel.dmodule import parallelize_module  # noqa: E402

_spec = importlib.util.spec_from_file_location("mixtral_4d_model", os.path.join(ROOT, "examples", "mixtral_4D_benchmark", "run.py"))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)


class MixtralLM(nn.Module):
    """The benchmark's Mixtral body + an untied LM head."""

    def __init__(self, a):
        super().__init__()
        self.model = _m.Mixtral(a)
        self.lm_head = nn.Linear(a.hidden_size, a.vocab_size, bias=False)

    def forward(self, ids):
        return self.lm_head(self.model(ids))


def lm_plan():
    """The benchmark's TP+SP plan under the ``model.`` prefix; the head stays replicated (vocab 96)."""
    p = {"parameter": {}, "forward": {}}
    for k, v in _m.mixtral_plan["parameter"].items():
        p["parameter"][r"model\." + k] = v
    for k, v in _m.mixtral_plan["forward"].items():
        p["forward"][(r"model\." + k) if k != "input" else r"model\.input"] = v
    p["forward"][r"lm_head\.input"] = [[Replicate()]]
    p["forward"][r"lm_head\.output"] = [[Replicate()]]
    return p


def make_corpus(n_chars: int = 60000, seed: int = 0) -> torch.Tensor:
    """A character stream with word / line structure: sentences drawn from a small grammar (learnable, unlike uniform noise)."""
    g = torch.Generator().manual_seed(seed)
    subj = ["the king", "my lord", "a fool", "thy brother", "the night", "sweet love", "old time", "the crown"]
    verb = ["doth speak", "shall fall", "will rise", "hath seen", "must die", "may weep", "did swear", "can wait"]
    tail = ["of war.", "in sorrow.", "to the sea.", "with grace.", "by night.", "for gold.", "no more.", "at dawn."]
    out = []
    while sum(len(s) for s in out) < n_chars:
        i, j, k = (int(torch.randint(0, 8, (1,), generator=g)) for _ in range(3))
        out.append(f"{subj[i]} {verb[j]} {tail[k]}\n")
    text = "".join(out)[:n_chars]
    return torch.tensor([min(95, max(0, ord(c) - 32)) if c != "\n" else 95 for c in text], dtype=torch.long)


def batch(corpus, bsz, seq, gen):
    ix = torch.randint(0, corpus.numel() - seq - 1, (bsz,), generator=gen)
    x = torch.stack([corpus[i : i + seq] for i in ix])
    y = torch.stack([corpus[i + 1 : i + seq + 1] for i in ix])
    return x, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dp", type=int, default=None)
    ap.add_argument("--tp", type=int, default=None)
    ap.add_argument("--max_iters", type=int, default=20)
    ap.add_argument("--bsz", type=int, default=8, help="global batch")
    ap.add_argument("--seqlen", type=int, default=32)
    ap.add_argument("--lr", type=float, default=3e-3)
    ap.add_argument("--hidden_size", type=int, default=64)
    ap.add_argument("--intermediate_size", type=int, default=128)
    ap.add_argument("--num_hidden_layers", type=int, default=2)
    ap.add_argument("--num_attention_heads", type=int, default=4)
    ap.add_argument("--num_key_value_heads", type=int, default=2)
    ap.add_argument("--num_experts", type=int, default=4)
    ap.add_argument("--top_k", type=int, default=2)
    ap.add_argument("--save_interval", type=int, default=0)
    ap.add_argument("--ckpt_dir", default=None)
    ap.add_argument("--resume", action="store_true")
    ap.add_argument("--compare-single", action="store_true", help="train an unparallelised twin on the global batch and assert equal loss curves")
    a = ap.parse_args()
    a.vocab_size = 96
    cuda = torch.cuda.is_available()
    dist.init_process_group("nccl" if cuda else "gloo")
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev, ws, rank = ("cuda" if cuda else "cpu"), dist.get_world_size(), dist.get_rank()
    a.tp = a.tp or (ws if a.dp is None else ws // a.dp)
    a.dp = a.dp or ws // a.tp
    mesh = init_device_mesh(dev, (a.dp, a.tp), mesh_dim_names=("DP", "TP"))
    torch.manual_seed(0)
    model = MixtralLM(a).to(dev)
    twin = copy.deepcopy(model) if a.compare_single else None
    parallelize_module(model, mesh["TP"], lm_plan())
    ddp = DDP(model, mesh["DP"].get_group(0), use_distributed_optimizer=True, overlap_grad_reduce=True)
    opt = DistributedOptimizer(torch.optim.AdamW(model.parameters(), lr=a.lr, weight_decay=0.0), [ddp], clip_grad=1.0, overlap_param_gather=False)
    twin_opt = torch.optim.AdamW(twin.parameters(), lr=a.lr, weight_decay=0.0) if twin is not None else None
    start = 0
    if a.resume and a.ckpt_dir and os.path.isdir(a.ckpt_dir):
        ckpt.load(a.ckpt_dir, {"model": ddp, "optimizer": opt})
        with open(os.path.join(a.ckpt_dir, "iter.txt")) as f:
            start = int(f.read())
    corpus = make_corpus()
    gen = torch.Generator().manual_seed(1234)
    for _ in range(start):  # the data stream is a function of the iteration only: a resumed run sees the same batches
        batch(corpus, a.bsz, a.seqlen, gen)
    per = a.bsz // a.dp
    dpr = mesh.get_local_rank("DP")
    losses, twin_losses = [], []
    for it in range(start, a.max_iters):
        x, y = batch(corpus, a.bsz, a.seqlen, gen)
        xl, yl = x[dpr * per : (dpr + 1) * per].to(dev), y[dpr * per : (dpr + 1) * per].to(dev)
        opt.zero_grad()
        logits = ddp(xl).to_local()
        loss = F.cross_entropy(logits.view(-1, a.vocab_size).float(), yl.view(-1))
        loss.backward()
        model.finish_grad_sync()
        opt.step()
        lt = loss.detach().clone()
        dist.all_reduce(lt, group=mesh["DP"].get_group(0))
        losses.append(float(lt) / a.dp)
        msg = f"iter {it}: loss {losses[-1]:.4f}"
        if twin is not None:
            twin_opt.zero_grad()
            tl = F.cross_entropy(twin(x.to(dev)).view(-1, a.vocab_size).float(), y.to(dev).view(-1))
            tl.backward()
            torch.nn.utils.clip_grad_norm_(twin.parameters(), 1.0)
            twin_opt.step()
            twin_losses.append(float(tl))
            msg += f"   single-device {twin_losses[-1]:.4f}"
        if rank == 0:
            print(msg, flush=True)
        if a.save_interval and a.ckpt_dir and (it + 1) % a.save_interval == 0:
            ckpt.save(a.ckpt_dir, {"model": ddp, "optimizer": opt})
            if rank == 0:
                with open(os.path.join(a.ckpt_dir, "iter.txt"), "w") as f:
                    f.write(str(it + 1))
            dist.barrier()
    if twin is not None:
        worst = max(abs(p - s) for p, s in zip(losses, twin_losses))
        assert worst < 5e-3 * max(1.0, max(twin_losses)), (worst, losses, twin_losses)
        if rank == 0:
            print(f"loss curves agree: max |dp{a.dp} x tp{a.tp} - single| = {worst:.2e}; first {losses[0]:.4f} last {losses[-1]:.4f}")
    assert losses[-1] < losses[0] or len(losses) < 5, "the loss must go down on a learnable corpus"
    ckpt.wait_for_async()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
