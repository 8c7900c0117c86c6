"""Kernel micro-benchmarks on one B200: hand-written kernels vs their library/eager counterparts.
CUDA-event timing, >=3 warm-ups, L2 flushed (256 MB write) between timed iterations.  Writes JSON."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/micro.json")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--gemm-once", action="store_true", help="launch each GEMM a few times only (for ncu)")
    args = ap.parse_args()
    from vescale_b200.ops import _ext

    _ext.load(required=True)
    ops = torch.ops.vescale_b200
    dev = "cuda"
    flush = torch.empty(256 * 2**20, dtype=torch.uint8, device=dev)
    peaks = {"hbm_gbs": 6478.3, "bf16_tflops": 1737.9, "bf16_tflops_sustained": 1462.2}
    try:
        peaks.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
    except Exception:
        pass
    res = {"gemm": [], "elementwise": [], "peaks": {k: peaks[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained")}}
    shapes = [(8192, 6144, 4096), (8192, 4096, 4096), (8192, 28672, 4096), (8192, 4096, 14336), (8192, 128256, 4096), (8192, 8192, 8192), (16384, 4096, 4096)]
    if args.quick:
        shapes = shapes[:2]
    for M, N, K in shapes:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn(N, K, device=dev).bfloat16()
        c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        if args.gemm_once:
            for _ in range(3):
                ops.gemm_nt(a, b, c, False)
            torch.cuda.synchronize()
            continue
        t_v1, _ = timeit(lambda: ops.gemm_nt(a, b, c, False, 1), flush=flush)
        t_ours, _ = timeit(lambda: ops.gemm_nt(a, b, c, False, 2), flush=flush)
        t_lib, _ = timeit(lambda: torch.mm(a, b.t(), out=c), flush=flush)
        fl = 2.0 * M * N * K
        r = {"M": M, "N": N, "K": K, "tcgen05_1cta_ms": t_v1, "tcgen05_2cta_ms": t_ours, "cublas_ms": t_lib, "tcgen05_1cta_tflops": fl / t_v1 / 1e9,
             "tcgen05_2cta_tflops": fl / t_ours / 1e9, "cublas_tflops": fl / t_lib / 1e9,
             "frac_of_measured_burst": fl / t_ours / 1e9 / peaks["bf16_tflops"]}
        # backward shapes of the same linear layer: dgrad [M,N]x[N,K] and wgrad [M,N]^T x [M,K]
        if hasattr(ops, "gemm_nn") and N <= 32768:
            dy = torch.randn(M, N, device=dev).bfloat16()
            dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
            dw = torch.empty(N, K, device=dev, dtype=torch.bfloat16)
            t1, _ = timeit(lambda: ops.gemm_nn(dy, b, dx), flush=flush)
            t1l, _ = timeit(lambda: torch.mm(dy, b, out=dx), flush=flush)
            t2, _ = timeit(lambda: ops.gemm_tn(dy, a, dw, False), flush=flush)
            t2l, _ = timeit(lambda: torch.mm(dy.t(), a, out=dw), flush=flush)
            r.update(dgrad_tcgen05_tflops=fl / t1 / 1e9, dgrad_cublas_tflops=fl / t1l / 1e9, wgrad_tcgen05_tflops=fl / t2 / 1e9, wgrad_cublas_tflops=fl / t2l / 1e9)
            del dy, dx, dw
        res["gemm"].append(r)
        print(r, flush=True)
        del a, b, c
    if args.gemm_once:
        return
    # ---- attention backends (library): pick the fastest for the Llama-3-8B shape
    from torch.nn.attention import SDPBackend, sdpa_kernel

    B_, S_, Hq, Hk, D = 1, 8192, 32, 8, 128
    qkv = torch.randn(B_, S_, (Hq + 2 * Hk) * D, device=dev).bfloat16().requires_grad_()
    res["attention"] = []
    for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
        try:
            def fwd():
                q = qkv[..., : Hq * D].unflatten(-1, (Hq, D)).transpose(1, 2)
                k = qkv[..., Hq * D : (Hq + Hk) * D].unflatten(-1, (Hk, D)).transpose(1, 2)
                v = qkv[..., (Hq + Hk) * D :].unflatten(-1, (Hk, D)).transpose(1, 2)
                with sdpa_kernel(be):
                    return torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)
            o = fwd()
            go = torch.randn_like(o)
            tf, _ = timeit(lambda: fwd(), iters=5)
            def fb():
                qkv.grad = None
                fwd().backward(go)
            tfb, _ = timeit(fb, iters=5)
            fl_f = 4 * B_ * Hq * S_ * S_ * D / 2
            r = {"backend": name, "fwd_ms": tf, "fwd_bwd_ms": tfb, "fwd_tflops": fl_f / tf / 1e9, "fwd_bwd_tflops": 3.5 * fl_f / tfb / 1e9}
        except Exception as e:  # noqa: BLE001
            r = {"backend": name, "error": str(e)[:200]}
        res["attention"].append(r)
        print(r, flush=True)
    del qkv
    # ---- bandwidth kernels
    T, H, F = 8192, 4096, 14336
    x = torch.randn(T, H, device=dev).bfloat16()
    y2 = torch.randn(T, H, device=dev).bfloat16()
    w = torch.ones(H, device=dev).bfloat16()
    gu = torch.randn(T, 2 * F, device=dev).bfloat16()
    dy = torch.randn(T, F, device=dev).bfloat16()

    def rec(name, fn, nbytes, ref=None):
        t, _ = timeit(fn, flush=flush)
        r = {"kernel": name, "ms": t, "GBps": nbytes / t / 1e6, "frac_of_measured_copy": nbytes / t / 1e6 / peaks["hbm_gbs"]}
        if ref is not None:
            tr, _ = timeit(ref, flush=flush)
            r["eager_ms"] = tr
            r["speedup_vs_eager"] = tr / t
        res["elementwise"].append(r)
        print(r, flush=True)

    rec("rms_norm_fwd", lambda: ops.rms_norm_fwd(x, w, 1e-5), 2 * T * H * 2, lambda: torch.nn.functional.rms_norm(x, (H,), w, 1e-5))
    yy, rstd = ops.rms_norm_fwd(x, w, 1e-5)
    rec("add_rms_norm_fwd", lambda: ops.add_rms_norm_fwd(x, y2, w, 1e-5), 4 * T * H * 2, lambda: torch.nn.functional.rms_norm(x + y2, (H,), w, 1e-5))
    rec("rms_norm_bwd", lambda: ops.rms_norm_bwd(y2, x, w, rstd), 3 * T * H * 2)
    rec("add_rms_norm_bwd", lambda: ops.add_rms_norm_bwd(y2, yy, x, w, rstd), 4 * T * H * 2)
    rec("swiglu_fwd", lambda: ops.swiglu_fwd(gu), 3 * T * F * 2, lambda: torch.nn.functional.silu(gu[:, :F]) * gu[:, F:])
    rec("swiglu_bwd", lambda: ops.swiglu_bwd(dy, gu), 5 * T * F * 2)
    qkv = torch.randn(T, 6144, device=dev).bfloat16()
    from vescale_b200.ops import functional as Fn

    cos, sin = Fn.rope_tables(T, 128, 500000.0, dev)
    rec("rope_qk", lambda: ops.rope_qk_(qkv, cos, sin, T, 32, 8, 128, 1.0), 2 * T * 5120 * 2)
    V = 128256
    logits = torch.randn(T, V, device=dev).bfloat16()
    tgt = torch.randint(0, V, (T,), device=dev)
    nv = torch.tensor([float(T)], device=dev)
    rec("cross_entropy_fwd_bwd", lambda: ops.cross_entropy_fwd_bwd_(logits, tgt, nv, -100), 2 * T * V * 2)
    n = 64 * 1024 * 1024
    master = torch.randn(n, device=dev)
    m_, v_ = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    g = torch.randn(n, device=dev).bfloat16()
    p = torch.empty(n, device=dev, dtype=torch.bfloat16)
    table = torch.tensor([[0, n, 1]], dtype=torch.int64, device=dev)
    coef = torch.ones(1, device=dev)
    rec("fused_adamw(bf16 grad)", lambda: ops.fused_adamw_(master, m_, v_, g, p, table, coef, 1e-3, 0.9, 0.95, 1e-8, 0.1, 0.1, 0.05, 1.0), n * (4 * 3 * 2 + 2 + 2))
    acc = torch.zeros(1, device=dev)
    rec("sumsq(bf16)", lambda: ops.sumsq_accumulate(g, acc, 1.0), n * 2)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
