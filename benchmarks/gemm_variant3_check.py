"""Bring-up check for the experimental 2x2-cluster multicast GEMM (variant 3): numerics against cuBLAS on a few shapes (edge
tiles included), then burst and sustained timing against variants 2 and cuBLAS.  The kernel has never run on hardware — run this
under a hard timeout first:

    timeout 120 python benchmarks/gemm_variant3_check.py

A protocol bug would show up as a hang (killed by the timeout) or a mismatch, not as a silent error.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from vescale_b200.ops import _ext

    _ext.load(required=True)
    ops = torch.ops.vescale_b200
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    out = {"numerics": [], "timing": []}
    for M, N, K in ((512, 256, 64), (512, 512, 256), (1024, 768, 512), (1000, 264, 192), (8192, 6144, 4096), (8192, 4096, 14336), (300, 136, 64)):
        a = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
        b = (torch.randn(N, K, device=dev, generator=g) * 0.5).bfloat16()
        c = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.gemm_nt(a, b, c, False, 3)
        torch.cuda.synchronize()
        ref = (a.float() @ b.float().t())
        err = (c.float() - ref).abs().max().item()
        ok = bool(torch.isfinite(c.float()).all()) and err <= 0.02 * ref.abs().max().item() + 0.05
        out["numerics"].append({"shape": [M, N, K], "max_abs_err": err, "ok": ok})
        print(f"variant 3 {M}x{N}x{K}: max err {err:.4f} {'ok' if ok else 'MISMATCH'}", flush=True)
        if not ok:
            print(json.dumps(out))
            sys.exit(1)

    def timeit(fn, n):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for M, N, K in ((8192, 6144, 4096), (8192, 28672, 4096), (8192, 4096, 14336)):
        a = torch.randn(M, K, device=dev, generator=g).bfloat16()
        b = torch.randn(N, K, device=dev, generator=g).bfloat16()
        c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        row = {"shape": [M, N, K]}
        for name, fn in (("v2", lambda: ops.gemm_nt(a, b, c, False, 2)), ("v3", lambda: ops.gemm_nt(a, b, c, False, 3)), ("cublas", lambda: torch.mm(a, b.t(), out=c))):
            row[f"{name}_burst_tflops"] = fl / timeit(fn, 10) / 1e9
            t_end = time.time() + 2.0  # sustained: ~2 s of back-to-back launches (power-capped clocks)
            n = 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            while time.time() < t_end:
                for _ in range(20):
                    fn()
                n += 20
            e1.record()
            torch.cuda.synchronize()
            row[f"{name}_sustained_tflops"] = fl * n / e0.elapsed_time(e1) / 1e9
        out["timing"].append(row)
        print(row, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gemm_variant3.json", "w"), indent=1)


if __name__ == "__main__":
    main()
