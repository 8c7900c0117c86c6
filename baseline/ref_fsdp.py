"""Reference arm of ``bench.py``: Llama-3-8B data-parallel-sharded training on the UNMODIFIED reference.

What runs here (none of ``vescale_b200`` is imported on this path):

* the reference package from ``baseline/_ref`` (volcengine/veScale @ 20cf5c7, ``pip --no-deps --target``), imported through
  ``baseline/ref_compat.py`` which only re-creates the three torch-private names torch 2.11 renamed;
* its public API — ``vescale.dtensor.DTensor.from_local / redistribute / to_local`` with ``RaggedShard`` placements —
  i.e. its stock collectives: list ``dist.all_gather`` + ``torch.cat`` for RaggedShard -> Replicate
  (``vescale/dtensor/placement_types.py:128-150``) and all-reduce-then-slice for Partial -> RaggedShard
  (``vescale/dtensor/_redistribute.py:111-120``);
* the stock model a reference user trains: HuggingFace ``LlamaForCausalLM`` (the reference's own examples wrap HF Llama /
  Mixtral, ``legacy/examples/open_llama_4D_benchmark``), torch SDPA attention, ``torch.optim.AdamW(fused=True)`` on the fp32
  local shards (what the reference's fused-adamw handler unwraps to, ``vescale/dtensor/_dispatch.py:118-132``).

The reference ships RaggedShard DTensor primitives but no FSDP wrapper (``docs/texts/raggedshard.md:67-71`` only describes
one), so the ~100-line unit loop below is the minimal driver a user of the reference writes: every parameter is a RaggedShard
DTensor; a unit's parameters are redistributed to Replicate before the unit runs and gradients are redistributed
Partial -> RaggedShard as they are produced.  Same model, sequence length, batch, dtype policy (bf16 compute, fp32 master +
AdamW state, global-norm clip 1.0), synthetic data and timing protocol as the ``--impl ours`` arm.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MODELS = {
    # name: (vocab, hidden, ffn, layers, heads, kv_heads, rope_theta)
    "llama3_8b": (128256, 4096, 14336, 32, 32, 8, 500000.0),
    "llama3_70b": (128256, 8192, 28672, 80, 64, 8, 500000.0),
    "open_llama_7b": (32000, 4096, 11008, 32, 32, 32, 10000.0),
    "tiny": (2048, 256, 512, 2, 8, 2, 10000.0),
}


def run(args, ClockSampler) -> int:
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = args.device == "cuda"
    if cuda:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank) if cuda else torch.device("cpu")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if not dist.is_initialized():
        if cuda:
            dist.init_process_group("nccl", device_id=dev, rank=rank, world_size=world)
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)

    sys.path.insert(0, ROOT)
    from baseline import ref_compat

    ref_compat.install()
    import vescale  # the reference, from baseline/_ref
    from vescale import Partial, Replicate, init_device_mesh
    from vescale.dtensor import DTensor, RaggedShard

    assert os.path.realpath(vescale.__file__).startswith(os.path.realpath(ref_compat.REF)), vescale.__file__
    from transformers import LlamaConfig, LlamaForCausalLM

    V, H, FF, L, NH, NKV, theta = MODELS[args.model]
    invalid = None
    if args.layers is not None:
        L, invalid = args.layers, f"layers overridden to {args.layers}"
    if not cuda:
        invalid = "cpu control-flow smoke test"
    S, B = args.seq_len, args.micro_batch
    dtype = torch.bfloat16 if cuda else torch.float32
    cfg = LlamaConfig(vocab_size=V, hidden_size=H, intermediate_size=FF, num_hidden_layers=L, num_attention_heads=NH, num_key_value_heads=NKV,
                      max_position_embeddings=max(S, 8192), rope_theta=theta, rms_norm_eps=1e-5, tie_word_embeddings=False, attn_implementation="sdpa", use_cache=False)
    torch.manual_seed(1234)
    with torch.device(dev):
        model = LlamaForCausalLM._from_config(cfg, dtype=dtype) if hasattr(LlamaForCausalLM, "_from_config") else LlamaForCausalLM(cfg).to(dtype)
    model.train()
    n_params = sum(p.numel() for p in model.parameters())
    mesh = init_device_mesh(args.device, (world,), mesh_dim_names=("dp",))
    # memory: 16 B/param of fp32 master+m+v+grad shard per rank, 2 B/param resident bf16 weights, ~1.2 GB/layer of activations at
    # 8192 tokens without recompute -> at N=1 a 8 B model does not fit 180 GB without checkpointing; N>=2 does.
    est = n_params * (16 / world + 2) + L * 36 * B * S * H * 2 / 8 * 1.0 + B * S * V * 10
    ac = bool(cuda and est > 150e9)
    if ac:
        model.gradient_checkpointing_enable()

    units = [[model.model.embed_tokens]] + [[blk] for blk in model.model.layers] + [[model.model.norm, model.lm_head]]
    place = lambda: RaggedShard(dims=(0,), local_units=(1,) * world)  # noqa: E731

    class Unit:
        def __init__(self, mods):
            self.params = [p for m in mods for p in m.parameters(recurse=True)]
            self.master = []  # fp32 local shards (torch Parameters handed to AdamW)
            self.stale = False
            for p in self.params:
                assert p.shape[0] % world == 0, p.shape
                dt = DTensor.from_local(p.data.view(-1).chunk(world)[rank].float().clone(), mesh, [place()], run_check=False, shape=p.shape, stride=p.stride())
                self.master.append(torch.nn.Parameter(dt.to_local()))
                p._ref_master = self.master[-1]
                p.register_post_accumulate_grad_hook(self.reduce_grad)
            mods[0].register_forward_pre_hook(lambda m, a: self.unshard())

        def unshard(self):
            """RaggedShard -> Replicate through the reference: uneven-capable list all_gather + cat, one per parameter."""
            if not self.stale:
                return
            for p, m in zip(self.params, self.master):
                dt = DTensor.from_local(m.data.to(dtype), mesh, [place()], run_check=False, shape=p.shape, stride=p.stride())
                p.data = dt.redistribute(mesh, [Replicate()]).to_local()
            self.stale = False

        def reduce_grad(self, p):
            """Partial -> RaggedShard through the reference: all-reduce, then slice this rank's rows."""
            g = DTensor.from_local(p.grad, mesh, [Partial()], run_check=False)
            gs = g.redistribute(mesh, [place()]).to_local()
            m = p._ref_master
            gs = gs.float().div_(world)
            m.grad = gs if m.grad is None else m.grad.add_(gs)
            p.grad = None

    funits = [Unit(u) for u in units]
    masters = [m for u in funits for m in u.master]
    opt = torch.optim.AdamW(masters, lr=3e-4, betas=(0.9, 0.95), weight_decay=0.1, fused=cuda, foreach=None if cuda else True)
    max_norm = args.max_grad_norm

    def optimizer_step():
        if max_norm:
            sq = torch.stack(torch._foreach_norm([m.grad for m in masters])).square().sum()
            if world > 1:
                dist.all_reduce(sq)  # what the reference's _NormPartial reduction issues
            coef = (max_norm / (sq.sqrt() + 1e-6)).clamp(max=1.0)
            torch._foreach_mul_([m.grad for m in masters], coef)
        opt.step()
        opt.zero_grad(set_to_none=True)
        for u in funits:
            u.stale = True

    n_batches = max(args.steps, 4)
    g = torch.Generator().manual_seed(1000 + rank)
    host_tok = [torch.randint(0, V, (B, S + 1), generator=g) for _ in range(n_batches)]
    host_tok = [t.pin_memory() for t in host_tok] if cuda else host_tok
    dev_tok = [t.to(dev) for t in host_tok[:4]]
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory() if cuda else torch.zeros(1)

    def fwd_bwd(t):
        logits = model(input_ids=t[:, :-1]).logits
        loss = F.cross_entropy(logits.float().view(-1, V), t[:, 1:].reshape(-1))
        loss.backward()
        optimizer_step()
        return loss

    def step_device(i):
        return fwd_bwd(dev_tok[i % len(dev_tok)])

    def step_e2e(i):
        loss = fwd_bwd(host_tok[i % n_batches].to(dev, non_blocking=True))
        loss_host.copy_(loss.detach().float().reshape(1), non_blocking=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank]) if cuda else dist.barrier()
        if cuda:
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0 and cuda:
        sampler.start()
    for i in range(args.warmup):
        step_device(i)
    barrier()
    mem_gb = torch.cuda.max_memory_allocated() / 2**30 if cuda else 0.0

    class _Wall:
        def __init__(self, enable_timing=True):
            self.t = 0.0

        def record(self):
            self.t = time.perf_counter()

        def elapsed_time(self, o):
            return (o.t - self.t) * 1e3

    Event = torch.cuda.Event if cuda else _Wall
    sampler.mark_begin()
    barrier()
    e0, e1 = Event(enable_timing=True), Event(enable_timing=True)
    marks = [Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    last = None
    for i in range(args.steps):
        last = step_device(i)
        marks[i].record()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    step_ms = [round((e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]), 2) for i in range(args.steps)]
    sampler.mark_end()
    clocks = sampler.stop() if (rank == 0 and cuda) else None
    final_loss = float(last.item())
    e2e_ms = 0.0
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_e2e(i)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        _ = float(loss_host[0])
    t = torch.tensor([ms, e2e_ms, mem_gb], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, mem_gb = t.tolist()
    tokens_per_step = world * B * S
    out = {
        "metric": "tokens/sec Llama-3-8B FSDP (bf16, RaggedShard veScale-FSDP)" if args.model == "llama3_8b" else f"tokens/sec {args.model} FSDP",
        "value": tokens_per_step * args.steps / (ms / 1e3),
        "unit": "tokens/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic tokens (uniform random ids), random-init weights of the named architecture",
        "impl": "reference",
        "config": {
            "model": args.model, "layers": L, "params_b": round(n_params / 1e9, 3), "global_batch": world * B, "seq_len": S,
            "tokens_per_gpu_per_step": B * S, "parallelism": f"fsdp{world}",
            "engine": "unmodified reference vescale 0.3.4a0 (baseline/_ref) RaggedShard DTensor redistribute (list all_gather+cat; all-reduce-then-slice) "
                      "driven by a minimal per-unit loop; HF LlamaForCausalLM + torch SDPA; torch.optim.AdamW(fused) on fp32 shards; NCCL + cuBLAS",
            "compat": "baseline/ref_compat.py re-creates torch-2.7 private names under torch 2.11; reference files unmodified",
            "reshard_after_forward": False,
            "optimizer": "AdamW fp32 master/m/v, global-norm clip 1.0" if max_norm else "AdamW fp32 master/m/v, no clip",
            "activation_memory": "full activation checkpointing per block (HF gradient_checkpointing; needed to fit)" if ac else "no recompute",
            "l2_policy": "no explicit flush: per-step working set is ~1000x the 126 MB L2",
        },
        "peak_mem_gb": mem_gb,
        "final_loss": final_loss,
        "step_ms": step_ms,
        "gpu_launches": 0,
        "clocks": clocks,
    }
    if not args.no_e2e:
        out["e2e"] = {"value": tokens_per_step * args.steps / (e2e_ms / 1e3), "unit": "tokens/s", "ms_per_step": e2e_ms / args.steps,
                      "h2d_bytes_per_step": host_tok[0].numel() * host_tok[0].element_size(), "d2h_bytes_per_step": 4,
                      "timing": "wall clock (perf_counter) bracketed by barrier+cuda synchronize, max over ranks"}
    if invalid:
        out["invalid"] = invalid
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    return 0
