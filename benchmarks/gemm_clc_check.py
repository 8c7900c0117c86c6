"""Bring-up + measurement of the cluster-launch-control (CLC) tile scheduler of the 2-CTA tcgen05 GEMM.

    timeout 200 python benchmarks/gemm_clc_check.py

1. numerics: CLC result must be bit-identical to the static persistent schedule (same tiles, same math) for NT / NN / TN,
   edge tiles included;
2. isolated timing (burst): CLC vs static vs cuBLAS;
3. timing while a communication-like kernel occupies a quarter / half of the SMs (512-thread memory-bound CTAs on a side
   stream — the FSDP reduce-scatter / all-gather pull kernels look exactly like this): the static schedule waits for its slowest
   CTA pair, CLC lets the undisturbed pairs take the remaining tiles.
Writes gpurun_out/gemm_clc.json.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from vescale_b200.ops import _ext

    _ext.load(required=True)
    ops = torch.ops.vescale_b200
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    out = {"numerics": [], "timing": []}

    def run(kind, a, b, c, sched):
        ops.gemm_set_sched(sched)
        if kind == "nt":
            ops.gemm_nt(a, b, c, False, 2)
        elif kind == "nn":
            ops.gemm_nn(a, b, c)
        else:
            ops.gemm_tn(a, b, c, False)

    ok_all = True
    for M, N, K in ((512, 256, 64), (1024, 768, 512), (1000, 264, 192), (300, 136, 64), (8192, 6144, 4096), (8192, 4096, 14336)):
        for kind in ("nt", "nn", "tn"):
            if kind == "nt":
                a = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
                b = (torch.randn(N, K, device=dev, generator=g) * 0.5).bfloat16()
            elif kind == "nn":
                a = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
                b = (torch.randn(K, N, device=dev, generator=g) * 0.5).bfloat16()
            else:
                if M % 8 or N % 8:
                    continue
                a = (torch.randn(K, M, device=dev, generator=g) * 0.5).bfloat16()
                b = (torch.randn(K, N, device=dev, generator=g) * 0.5).bfloat16()
            c0 = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
            c1 = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
            run(kind, a, b, c0, 0)
            run(kind, a, b, c1, 1)
            torch.cuda.synchronize()
            same = bool(torch.equal(c0, c1)) and bool(torch.isfinite(c1.float()).all())
            out["numerics"].append({"kind": kind, "shape": [M, N, K], "bit_identical": same})
            print(f"clc {kind} {M}x{N}x{K}: {'identical' if same else 'MISMATCH'}", flush=True)
            ok_all &= same
    if not ok_all:
        print(json.dumps(out))
        ops.gemm_set_sched(0)
        sys.exit(1)

    def timeit(fn, n=20):
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

    # a "communication kernel": the FSDP pull all-gather on one rank = a 512-thread streaming copy on a fixed number of CTAs
    src = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    dst = torch.empty_like(src)
    side = torch.cuda.Stream()

    def noise(ctas):
        ops.symm_all_gather([src.data_ptr()], dst, src.numel(), 0, [], 0, 0, ctas, 0, 0, 0)

    for M, N, K in ((8192, 6144, 4096), (8192, 28672, 4096), (8192, 4096, 14336)):
        a = torch.randn(M, K, device=dev, generator=g).bfloat16()
        b = torch.randn(N, K, device=dev, generator=g).bfloat16()
        c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        row = {"shape": [M, N, K]}
        for name, fn in (("static", lambda: run("nt", a, b, c, 0)), ("clc", lambda: run("nt", a, b, c, 1)), ("cublas", lambda: torch.mm(a, b.t(), out=c))):
            row[f"{name}_alone_tflops"] = fl / (timeit(fn) * 1e-3) / 1e12
            for ctas in (37, 74):
                # keep the side stream busy for the whole measurement
                stop = torch.cuda.Event()
                with torch.cuda.stream(side):
                    for _ in range(40):
                        noise(ctas)
                ms = timeit(fn, 12)
                side.synchronize()
                row[f"{name}_with_{ctas}cta_copy_tflops"] = fl / (ms * 1e-3) / 1e12
        out["timing"].append(row)
        print(json.dumps(row), flush=True)
    ops.gemm_set_sched(0)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_clc.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
