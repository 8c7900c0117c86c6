"""Every environment switch the framework reads, in one table (name, default, meaning).  ``flag(name)`` returns the parsed
value; modules may still read ``os.environ`` directly on hot paths, but each name used anywhere must be listed here
(``tests/test_contracts.py`` greps for strays)."""
from __future__ import annotations

import os
from typing import Dict, Tuple

__all__ = ["FLAGS", "flag", "describe_flags"]

# name -> (default, description)
FLAGS: Dict[str, Tuple[str, str]] = {
    "VESCALE_DISABLE_REDISTRIBUTE": ("0", "1 = op dispatch raises instead of resharding operands implicitly (legacy default was 1)"),
    "VESCALE_DISABLE_RUN_CHECK": ("0", "1 = DTensor.from_local skips the cross-rank metadata check"),
    "VESCALE_STRICT_RULES": ("0", "1 = ops without a sharding rule raise instead of falling back to replicated execution"),
    "VESCALE_SINGLE_DEVICE_RAND": ("0", "1 = aten random ops on DTensors (dropout, uniform_, normal_, rand_like) are single-device-equivalent (ThreadBasedRNGTracker)"),
    "VESCALE_DEBUG_MODE": ("", "non-empty = DebugLogger prints every dispatched op and mesh collective (rank filter after ':')"),
    "VESCALE_DUMMY_P2P": ("0", "1 = pipeline p2p ops are logged, not executed (schedule dry run)"),
    "VESCALE_DUMP_INSTRUCTION": ("0", "1 = the pipeline engine dumps each rank's instruction list to a file"),
    "VESCALE_CHECKPOINT_LOGGING_LEVEL": ("", "level of the checkpoint logger: a logging level name (DEBUG, INFO, ...) or number; empty = WARNING"),
    "VESCALE_DEVICE_MESH": ("", "internal: name of the global VeDeviceMesh registry entry"),
    "VESCALE_B200_ALLOW_FALLBACK": ("0", "1 = allow PyTorch fallbacks on a CUDA device when vescale_b200/_C.so is missing (default: fail loudly)"),
    "VESCALE_B200_SYMM_DEBUG": ("0", "1 = poison symmetric buffers when they return to a pool and validate signal epochs (comm/symm_debug.py)"),
    "VESCALE_B200_MULTIMEM": ("1", "0 = never use NVLS multimem instructions in the symmetric-memory kernels"),
    "VESCALE_B200_SYMM_CHUNK_MB": ("2048", "size of one symmetric-memory arena chunk (one rendezvous per chunk)"),
    "VESCALE_B200_GEMM_VARIANT": ("2", "read by csrc/gemm_sm100.cu: 1 = 1-CTA tcgen05 kernel, 2 = CTA pairs (default), 3 = experimental 2x2 cluster with TMA multicast of B"),
    "VESCALE_B200_GEMM_GROUP_M": ("8", "read by csrc/gemm_sm100.cu: M-tiles per rasterisation group (L2 locality)"),
    "VESCALE_B200_AG_CTAS": ("0", "CTAs of the FSDP pull all-gather kernel (0 = a quarter of the SMs)"),
    "VESCALE_B200_RS_CTAS": ("0", "CTAs of the fused reduce-scatter kernels (0 = 32 on the NVLS path, a third of the SMs on the P2P path)"),
    "VESCALE_NDTIMELINE_LOG_LEVEL": ("INFO", "level of the ndtimeline logger (unknown names fall back to WARNING)"),
    "VESCALE_NDTIMELINE_SOCK_DIR": ("/tmp/ndtimeline", "directory of the default collector socket of the ndtimeline streamer"),
    "VESCALE_B200_MXFP8_NATIVE": ("0", "1 = MXFP8 GEMMs (ops.fp8.mxfp8_gemm_nt, fp8_linear(recipe='mx')) run on the tcgen05 block-scaled kernel csrc/gemm_mxfp8.cu instead of the emulation"),
    "VESCALE_B200_SYMM_RS": ("0", "1 = mesh_reduce_scatter (DTensor Partial -> Shard) uses the symmetric-memory reduce-scatter kernel (NVLS / P2P) instead of NCCL"),
    "VESCALE_B200_FUSE_TP": ("auto", "DTensor-level fusion of redistribute -> mm / mm -> redistribute onto ag_gemm / gemm_rs (dtensor/fusion.py): auto = fused sm_100a kernels on CUDA, c10d = same pattern match on ordinary collectives (CPU tests, baseline), off = generic path"),
    "VESCALE_B200_AG_IMPL": ("ce", "FSDP all-gather transport: ce = peer cudaMemcpyAsync on the copy engines (no SM), pull = SM pull kernel"),
    "VESCALE_B200_GEMM_SCHED": ("static", "read by csrc/gemm_sm100.cu: static = persistent grid, tile += grid; clc = cluster launch control (one cluster per tile, running clusters pull the remaining tiles)"),
    "VESCALE_B200_ATTN": ("auto", "attention back end of ops.packed_attention: tcgen05 = hand-written sm_100a flash attention (csrc/attention_sm100.cu), cudnn = library SDPA, auto = faster of the two per shape"),
    "VESCALE_B200_CLOCK_SAMPLER": ("nvml", "bench.py clock / throttle sampler: nvml = in-process NVML thread, smi = nvidia-smi -lms child process"),
    "VESCALE_CHECKPOINT_COORDINATOR": ("", "host:port of a checkpoint report service (checkpoint/server_lib.py): saves coordinate plans / results through it instead of process-group collectives"),
    "VESCALE_CHECKPOINT_WORKERS": ("2", "writer processes per rank that serialise and write checkpoint files (checkpoint/storage.py); 0 = in-process writer"),
    "VESCALE_B200_ATTN_FWD": ("3", "read by csrc/attention_sm100.cu: forward kernel variant (1 = four softmax warps, 2 = eight softmax warps splitting each row in halves, 3 = 2 with one 64-column TMEM load per tile)"),
    "VESCALE_B200_GEMM_RS": ("staged", "FusedTP.gemm_rs implementation: staged = partial tiles pushed into the owner's staging slots by the GEMM epilogue; nvls = GEMM into a symmetric buffer + switch-reduced pull of the owner's rows"),
}


def flag(name: str):
    default, _ = FLAGS[name]
    v = os.environ.get(name, default)
    if default in ("0", "1"):
        return v == "1"
    if default.isdigit():
        return int(v)
    return v


def describe_flags() -> str:
    w = max(len(k) for k in FLAGS)
    return "\n".join(f"{k.ljust(w)}  default={d!r:8}  {doc}" for k, (d, doc) in FLAGS.items())
