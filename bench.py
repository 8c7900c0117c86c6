#!/usr/bin/env python
"""Headline benchmark: tokens/s of Llama-3-8B veScale-FSDP (RaggedShard) training in bf16 on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        bench.py --gpus 8 --steps 5 --warmup 3

Weak scaling: every GPU trains ``micro_batch x seq_len`` tokens per step (global batch = N x micro_batch).
Synthetic tokens, random-init weights of the named architecture (no network).  One JSON line on rank 0.

Two timed regions, both after ``--warmup`` untimed steps, both bracketed by barrier + cuda synchronize:
  * ``value`` — K full training steps (forward, backward, reduce-scatter, clip, AdamW, all-gather) with the
    batch already on the device, timed with CUDA events on the launching stream, max over ranks;
  * ``e2e``   — K steps through the public API including, every step, the host->device copy of that step's
    batch from pinned memory and a device->host read of the loss, wall clock, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--model", default="llama3_8b", choices=["llama3_8b", "llama3_70b", "open_llama_7b", "tiny"])
    p.add_argument("--seq-len", type=int, default=8192)
    p.add_argument("--micro-batch", type=int, default=1)
    p.add_argument("--layers", type=int, default=None, help="debug only: override the layer count (marks the run invalid)")
    p.add_argument("--comm", default="auto", choices=["auto", "nccl", "symm"])
    p.add_argument("--gemm", default="auto", choices=["auto", "tcgen05", "cublas"])
    p.add_argument("--reshard", default="auto", choices=["auto", "yes", "no"])
    p.add_argument("--attn", default="auto", choices=["auto", "tcgen05", "cudnn"], help="attention back end: hand-written tcgen05 flash attention, library SDPA (cuDNN), or the faster of the two")
    p.add_argument("--max-grad-norm", type=float, default=1.0)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--device", default="cuda", choices=["cuda", "cpu"], help="cpu = control-flow smoke test of this script on gloo (marks the record invalid)")
    p.add_argument("--tp", type=int, default=1, help="tensor/sequence-parallel degree (2-D: FSDP over world/tp x TP over tp, fused TP kernels)")
    p.add_argument("--tp-impl", default="fused", choices=["fused", "plain"], help="fused = ag_gemm/gemm_rs sm_100a kernels; plain = NCCL + library GEMMs")
    p.add_argument("--fused-reduce", action="store_true", help="reduce-scatter ⊕ AdamW in one kernel (no gradient shard is ever stored; implies no global-norm clip; symm backend, N > 1) — the memory-lean optimizer path for 70B-class models")
    p.add_argument("--ac", default="selective", choices=["selective", "full"], help="selective = recompute norm / SwiGLU outputs only (default); full = checkpoint every block (70B-class models)")
    p.add_argument("--fp8", nargs="?", const="block128", default=None, choices=["block128", "mx"],
                   help="block-scaled e4m3 forward GEMMs in the decoder blocks (BASELINE config 5; use with --model llama3_70b): block128 = 1x128 / 128x128 "
                        "fp32 scales through cuBLASLt, mx = OCP MXFP8 on the hand-written tcgen05 block-scaled kernel (sets VESCALE_B200_MXFP8_NATIVE=1)")
    p.add_argument("--prefetch", type=int, default=1, help="FSDP all-gather prefetch depth (0 = every all-gather exposed: the memory-lean mode)")
    p.add_argument("--fuse-first-gemm", action="store_true", help="exposed all-gathers: the unit's first GEMM gathers its own weight (wag_gemm)")
    p.add_argument("--profile", default=None, help="after the timed regions, run ONE extra step under torch.profiler and write the per-kernel table here")
    return p.parse_args()


def reference_arm(args):
    """The UNMODIFIED reference from ``baseline/_ref`` through its own public API (``baseline/ref_fsdp.py``): its RaggedShard
    DTensor redistribute collectives drive a minimal per-unit sharded-training loop over the stock HF Llama model (the
    reference ships no FSDP wrapper).  ``baseline/ref_compat.py`` re-creates the torch-2.7 private names the
    reference imports; if the reference still cannot run, one ``unavailable`` line is printed and the exit code is 0."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    rank = int(os.environ.get("RANK", "0"))
    if not os.path.isdir(os.path.join(ref, "vescale")):
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/vescale not installed (pip --no-index --no-deps --target baseline/_ref /root/reference)"}))
        return 0
    try:
        from baseline import ref_fsdp

        return ref_fsdp.run(args, ClockSampler)
    except Exception as e:  # noqa: BLE001
        import traceback

        traceback.print_exc()
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"reference arm failed under torch 2.11: {type(e).__name__}: {str(e)[:200]}"}))
        return 0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def _start_nvml(self) -> bool:
        """In-process NVML sampling (nvidia_ml_py): one nvmlInit, then three cheap queries per 200 ms sample.  Preferred over a
        `nvidia-smi -lms` child process: nvidia-smi re-enumerates every GPU of the host for each sample and was the prime suspect
        for the sporadic 150-550 ms stall of ONE step inside the device-timed region (it only ever appeared
        while the sampler ran)."""
        try:
            import pynvml as nv

            nv.nvmlInit()
            try:
                import torch

                uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
                h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {
                "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
            }
            self._stop = threading.Event()

            def loop():
                while not self._stop.is_set():
                    try:
                        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                        pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                        r = int(reasons_fn(h))
                        flags = ", ".join("Active" if r & bits[n] else "Not Active" for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"))
                        self.lines.append((time.time(), f"{self.idx}, {sm}, {mx}, {pw:.2f}, {flags}"))
                    except Exception:
                        pass
                    self._stop.wait(0.2)

            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
            self.nvml = True
            return True
        except Exception:
            return False

    def start(self):
        if os.environ.get("VESCALE_B200_CLOCK_SAMPLER", "nvml") == "nvml" and self._start_nvml():
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if getattr(self, "nvml", False):
            self._stop.set()
            self.t.join(timeout=2)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t_begin", 0.0), getattr(self, "t_end", float("inf"))
        inside = [(ts, l) for ts, l in self.lines if t0 <= ts <= t1 + 0.25]  # samples taken while the timed region was running
        for ts, l in inside or self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "power_w_max": max(pw) if pw else None,
            "samples": len(sm),
            "sampler": "nvml (in-process, 200 ms)" if getattr(self, "nvml", False) else "nvidia-smi -lms 200",
            "reasons": sorted(reasons),
        }


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    cuda = args.device == "cuda"
    if cuda:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank) if cuda else torch.device("cpu")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev) if cuda else dist.init_process_group("gloo")

    class _WallEvent:  # CPU stand-in for torch.cuda.Event (perf_counter based)
        def __init__(self, enable_timing=True):
            self.t = 0.0

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    Event = torch.cuda.Event if cuda else _WallEvent

    from vescale_b200 import init_device_mesh
    from vescale_b200.models import LlamaConfig, LlamaModel, llama_flops_per_token
    from vescale_b200.ops import _ext, functional as Fn
    from vescale_b200.optim import FSDPAdamW
    from vescale_b200.parallel.fsdp import fully_shard

    if cuda:
        _ext.load(required=True)
    Fn.set_gemm_backend(args.gemm)
    Fn.set_attention_backend(args.attn)
    cfg = getattr(LlamaConfig, args.model)()
    cfg.fp8 = args.fp8 or False
    if args.fp8 == "mx":
        os.environ.setdefault("VESCALE_B200_MXFP8_NATIVE", "1")
    invalid = None
    if args.layers is not None:
        cfg.num_layers = args.layers
        invalid = f"layers overridden to {args.layers}"
    S, B = args.seq_len, args.micro_batch
    if not cuda:
        invalid = "cpu control-flow smoke test"
        cfg.dtype = torch.float32
    cfg.max_seq_len = max(cfg.max_seq_len, S)
    tp_size = args.tp
    if world % tp_size:
        raise SystemExit("--tp must divide the number of GPUs")
    dp_size = world // tp_size
    tp = None
    if tp_size > 1:
        from vescale_b200.comm.fused_tp import FusedTP, PlainTP
        from vescale_b200.models.llama_tp import LlamaTPModel

        mesh = init_device_mesh(args.device, (dp_size, tp_size), mesh_dim_names=("dp", "tp"))
        tp = FusedTP(mesh, "tp", dev) if (args.tp_impl == "fused" and cuda) else PlainTP(mesh, "tp")
    else:
        mesh = init_device_mesh(args.device, (world,), mesh_dim_names=("dp",), **({} if world > 1 else {"_init_process_groups": False}))
    dp_rank = mesh.get_local_rank("dp") if world > 1 else 0

    # ---- build: meta-device model, materialised unit by unit straight into the sharded master weights
    torch.manual_seed(1234)
    with torch.device("meta"):
        model = LlamaModel(cfg) if tp is None else LlamaTPModel(cfg, tp)
    reshard = {"auto": None, "yes": True, "no": False}[args.reshard]
    if reshard is None:
        # keep gathered bf16 weights resident when they fit comfortably (ZeRO-2-style): saves the backward all-gather
        reshard = not (dp_size > 1 and cfg.num_params() * 2 / tp_size < 40e9)
    gens = {}

    def init_fn(mod):
        g = gens.setdefault("g", torch.Generator(device=dev).manual_seed(1234 + 7 * (mesh.get_local_rank("tp") if tp is not None else 0)))
        if hasattr(mod, "reset_parameters"):
            mod.reset_parameters(g) if isinstance(mod, type(model.layers[0])) else None
        if mod is model.embed:
            with torch.no_grad():
                mod.weight.normal_(0, cfg.init_std, generator=g)
        if mod is model.head:
            with torch.no_grad():
                mod.norm.fill_(1.0)
                mod.weight.normal_(0, cfg.init_std, generator=g)

    kw = dict(comm_backend=args.comm, reshard_after_forward=reshard, init_fn=init_fn, prefetch=args.prefetch, mesh_dim="dp")
    fully_shard(model.embed, mesh, **kw)
    for blk in model.layers:
        if args.ac == "full":
            from vescale_b200.parallel.fsdp import checkpoint_module

            checkpoint_module(blk)
        fully_shard(blk, mesh, fuse_first_gemm=bool(args.fuse_first_gemm and dp_size > 1 and tp is None), **kw)
    fully_shard(model.head, mesh, **kw)
    fully_shard(model, mesh, **kw)
    fused_reduce = bool(args.fused_reduce and dp_size > 1 and tp is None)
    if fused_reduce:
        args.max_grad_norm = 0.0
    opt = FSDPAdamW(model, lr=3e-4, betas=(0.9, 0.95), weight_decay=0.1, max_grad_norm=args.max_grad_norm or None, fused_reduce=fused_reduce,
                    tp_group=mesh.get_group("tp") if tp is not None else None)
    state = model._fsdp_state
    comm_name = type(state.comm).__name__ if state.comm is not None else ("none" if dp_size == 1 else "nccl")

    # ---- synthetic data: pinned host batches (e2e) and device-resident copies (kernel-timed)
    n_batches = max(args.steps, 4)
    g = torch.Generator().manual_seed(1000 + dp_rank)  # the ranks of one TP group see the same batch
    host_tok = [torch.randint(0, cfg.vocab_size, (B, S + 1), generator=g) for _ in range(n_batches)]
    host_tok = [t.pin_memory() for t in host_tok] if cuda else host_tok
    dev_tok = [t.to(dev) for t in host_tok[:4]]
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory() if cuda else torch.zeros(1, dtype=torch.float32)
    h2d_bytes = host_tok[0].numel() * host_tok[0].element_size()
    d2h_bytes = 4

    def step_device(i):
        t = dev_tok[i % len(dev_tok)]
        loss = model(t[:, :-1], t[:, 1:])
        loss.backward()
        opt.step()
        opt.zero_grad()
        return loss

    def step_e2e(i):
        t = host_tok[i % n_batches].to(dev, non_blocking=True)
        loss = model(t[:, :-1], t[:, 1:])
        loss.backward()
        opt.step()
        opt.zero_grad()
        loss_host.copy_(loss.detach().float().reshape(1), non_blocking=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank]) if cuda else dist.barrier()
        if cuda:
            torch.cuda.synchronize()

    # the clock sampler is started before the warm-up: spawning nvidia-smi (NVML init over all GPUs) stalls kernel launches for
    # a few hundred ms, which must not land inside the timed region; only samples taken during the region are reported
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    elif cuda and os.environ.get("VESCALE_B200_CLOCK_SAMPLER", "nvml") == "nvml":
        sampler._start_nvml()  # the other ranks sample their own GPU in-process too (per-rank diagnostics); never a child process
    import gc

    e0, e1 = Event(enable_timing=True), Event(enable_timing=True)
    marks = [Event(enable_timing=True) for _ in range(args.steps)]

    def housekeeping(waits_per_step: int) -> None:
        """Everything slow that has to happen before the timed region, done BEFORE THE LAST WARM-UP STEP so that the GPU goes from a
        full training step straight into the timed steps.  (a) A generation-2 garbage collection over the ~10^5 live Python objects
        of the model takes 150-250 ms: collect now, freeze the survivors, keep the collector off in the timed regions (Megatron's
        --manual-gc).  (b) No CUDA event is CREATED inside the timed region: torch creates events lazily at their first record and an
        event-pool growth in the driver synchronises the device; the exposed-communication events come from a pre-recorded pool,
        the step marks are recorded once here.  (c) The GPU must not sit idle right before the region: under the 1 kW cap the power
        controller lets the first step after an idle gap run fast (390-399 ms against 403-409 steady) and then over-corrects on
        the second one — the one slow step of the earlier records (464 ms at N=2, 644 ms at N=8, always the SECOND timed step,
        never in the end-to-end region, which starts without such a gap)."""
        gc.collect()
        gc.freeze()
        gc.disable()
        state.prepare_exposed_measure((waits_per_step + 8) * args.steps)
        if cuda:
            for ev in [e0, e1] + marks:
                ev.record()

    waits_per_step = 0
    for i in range(args.warmup):
        if i == 0:  # count the compute-stream waits of one step: the measured region draws its timing events from a pool
            state.measure_exposed, state._exposed_events = True, []
        if i == args.warmup - 1 and i > 0:
            housekeeping(waits_per_step)
        step_device(i)
        if i == 0:
            waits_per_step = len(state._exposed_events)
            state.measure_exposed, state._exposed_events = False, []
    if args.warmup <= 1:
        housekeeping(waits_per_step)
    mem_gb = torch.cuda.max_memory_allocated() / 2**30 if cuda else 0.0

    # ---- timed region 1: device-timed (barrier + synchronize on both sides; nothing else between the last warm-up step and it)
    sampler.mark_begin()
    state.measure_exposed = True
    state._exposed_events = []
    _ext.LAUNCH_COUNTER.update(n=0, enabled=True, by_op={})
    barrier()
    e0.record()
    last = None
    host_ms = []  # host time spent ENQUEUEING each step (diagnostic: a step whose host time jumps was blocked in the driver)
    for i in range(args.steps):
        th = time.perf_counter()
        last = step_device(i)
        marks[i].record()
        host_ms.append(round((time.perf_counter() - th) * 1e3, 1))
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    exposed_ms = state.exposed_comm_ms() / args.steps  # compute-stream stalls on all-gather / reduce-scatter completion
    state.measure_exposed = False
    step_ms = [round((e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]), 2) for i in range(args.steps)]
    launches = _ext.LAUNCH_COUNTER["n"]
    by_op = dict(_ext.LAUNCH_COUNTER["by_op"])
    _ext.LAUNCH_COUNTER["enabled"] = False
    sampler.mark_end()
    clocks = sampler.stop() if (rank == 0 or getattr(sampler, "nvml", False)) else None
    final_loss = float(last.item()) * tp_size  # TP: the model returns this rank's share (local mean / tp)

    # ---- timed region 2: end to end through the public API (pinned H2D of the batch + D2H of the loss every step)
    e2e_ms = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_e2e(i)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        _ = float(loss_host[0])

    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step_device(0)
            torch.cuda.synchronize()
        os.makedirs(os.path.dirname(os.path.abspath(args.profile)), exist_ok=True)
        with open(args.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
    elif args.profile:
        step_device(0)

    t = torch.tensor([ms, e2e_ms if e2e_ms is not None else 0.0, mem_gb, exposed_ms], device=dev, dtype=torch.float64)
    # per-rank diagnostics: the rank whose compute stream never waits for a collective is the straggler the others wait for, and
    # under the 1 kW power cap the GPUs of one box settle at different SM clocks (the step time of the job is the slowest GPU's)
    per_rank = torch.tensor([exposed_ms, float((clocks or {}).get("sm_mhz") or 0.0), float((clocks or {}).get("power_w_max") or 0.0)], device=dev, dtype=torch.float64)
    per_rank_all = [per_rank.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank_all, per_rank)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_rank_diag = {
        "exposed_comm_ms_per_step": [round(float(x[0]), 2) for x in per_rank_all],
        "sm_mhz": [float(x[1]) for x in per_rank_all],
        "power_w_max": [round(float(x[2]), 1) for x in per_rank_all],
    }
    ms, e2e_ms_max, mem_gb, exposed_ms = t.tolist()
    tokens_per_step = dp_size * B * S
    tps = tokens_per_step * args.steps / (ms / 1e3)
    flops_tok = llama_flops_per_token(cfg, S)
    peak = 1462.2e12
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f)["bf16_tflops_sustained"] * 1e12
    except Exception:
        pass
    out = {
        "metric": "tokens/sec Llama-3-8B FSDP (bf16, RaggedShard veScale-FSDP)" if args.model == "llama3_8b" else f"tokens/sec {args.model} FSDP",
        "value": tps,
        "unit": "tokens/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16" if not args.fp8 else ("fp8 e4m3 block-scaled forward GEMMs (1x128 / 128x128 scales), bf16 backward" if args.fp8 == "block128" else "MXFP8 (e4m3, 1x32 E8M0 scales) forward GEMMs, bf16 backward"),
        "data": "synthetic tokens (uniform random ids), random-init weights of the named architecture",
        "impl": "ours",
        "config": {
            "model": args.model,
            "layers": cfg.num_layers,
            "params_b": round(cfg.num_params() / 1e9, 3),
            "global_batch": dp_size * B,
            "seq_len": S,
            "tokens_per_gpu_per_step": B * S // tp_size,
            "parallelism": f"fsdp{world}" if tp is None else f"fsdp{dp_size}xtp{tp_size}({type(tp).__name__})",
            "comm_backend": comm_name,
            "gemm_backend": args.gemm,
            "attention_backend": args.attn,
            "reshard_after_forward": bool(reshard),
            "prefetch": args.prefetch,
            "fuse_first_gemm": bool(args.fuse_first_gemm),
            "optimizer": ("AdamW fp32 master/m/v, global-norm clip 1.0" if args.max_grad_norm else "AdamW fp32 master/m/v, no clip") + (", fused into the reduce-scatter kernel" if fused_reduce else ""),
            "activation_memory": "selective recompute (norm and SwiGLU outputs recomputed)" if args.ac == "selective" else "full activation checkpointing per block",
            "host_gc": "manual: gc.collect() + gc.freeze() after warm-up, automatic collection off during the timed regions",
            "l2_policy": "no explicit flush: per-step working set (>100 GB of weights/optimizer state/activations) is ~1000x the 126 MB L2",
        },
        "mfu_of_measured_cublas_sustained": tps / world * flops_tok / peak,
        "model_tflops_per_gpu": tps / world * flops_tok / 1e12,
        "exposed_comm_ms_per_step": exposed_ms,  # device-timed stall of the compute stream on FSDP collectives, max over ranks
        "peak_mem_gb": mem_gb,
        "final_loss": final_loss,
        "step_ms": step_ms,
        "host_enqueue_ms": host_ms,
        "per_rank": per_rank_diag,
        "gpu_launches": launches,
        "gpu_launches_by_op": by_op,
        "clocks": clocks,
    }
    if e2e_ms is not None:
        out["e2e"] = {
            "value": tokens_per_step * args.steps / (e2e_ms_max / 1e3),
            "unit": "tokens/s",
            "ms_per_step": e2e_ms_max / args.steps,
            "h2d_bytes_per_step": h2d_bytes,
            "d2h_bytes_per_step": d2h_bytes,
            "timing": "wall clock (perf_counter) bracketed by barrier+cuda synchronize, max over ranks",
        }
    if invalid:
        out["invalid"] = invalid
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
