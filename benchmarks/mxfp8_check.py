"""Bring-up check for the hand-written MXFP8 GEMM (csrc/gemm_mxfp8.cu: tcgen05.mma.kind::mxf8f6f4.block_scale, scale factors
staged smem -> TMEM by tcgen05.cp).  The kernel was written without hardware access — run it under a hard timeout first:

    timeout 150 python benchmarks/mxfp8_check.py

1. numerics against the dequantise-and-multiply specification (ops/fp8.py) on growing shapes (edge tiles included); a wrong
   scale-factor byte/row mapping shows up here because the operands carry a wide dynamic range across K blocks;
2. timing against the bf16 tcgen05 kernel, cuBLAS bf16 and (when the build accepts E8M0 scales) cuBLASLt MXFP8 through
   torch._scaled_mm on the same operands and the same scale-factor buffers.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from vescale_b200.ops import _ext, fp8

    _ext.load(required=True)
    ops = torch.ops.vescale_b200
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    out = {"numerics": [], "timing": []}

    def operands(M, N, K):
        x = torch.randn(M, K, device=dev, generator=g) * torch.exp2(torch.randint(-6, 6, (M, K // 32), device=dev, generator=g).float()).repeat_interleave(32, 1)
        w = torch.randn(N, K, device=dev, generator=g) * torch.exp2(torch.randint(-6, 6, (N, K // 32), device=dev, generator=g).float()).repeat_interleave(32, 1)
        xq, xs = fp8.quantize_mx(x)
        wq, ws = fp8.quantize_mx(w)
        return x, w, xq, xs, wq, ws, fp8.mx_scale_atoms(xs, 128), fp8.mx_scale_atoms(ws, 256)

    for M, N, K in ((128, 256, 128), (128, 256, 512), (256, 512, 256), (300, 264, 384), (1024, 768, 2048), (8192, 6144, 4096)):
        x, w, xq, xs, wq, ws, xa, wa = operands(M, N, K)
        ref = fp8.dequantize_mx(xq, xs) @ fp8.dequantize_mx(wq, ws).t()
        c = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        fp8.mxfp8_gemm_nt_native(xq, xa, wq, wa, out=c)
        torch.cuda.synchronize()
        err = ((c.float() - ref).norm() / ref.norm()).item()
        ok = bool(torch.isfinite(c.float()).all()) and err < 6e-3
        out["numerics"].append({"shape": [M, N, K], "rel_fro_err": err, "ok": ok})
        print(f"mxfp8 {M}x{N}x{K}: rel err {err:.2e} {'ok' if ok else 'MISMATCH'}", flush=True)
        if not ok:
            # diagnosis aid: which (row block, K block) pattern is off?
            d = (c.float() - ref).abs()
            print("worst rows:", d.amax(1).topk(min(8, M)).indices.tolist(), "worst cols:", d.amax(0).topk(min(8, N)).indices.tolist())
            print(json.dumps(out))
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

    for M, N, K in ((8192, 6144, 4096), (8192, 28672, 4096), (8192, 4096, 14336), (8192, 8192, 8192)):
        x, w, xq, xs, wq, ws, xa, wa = operands(M, N, K)
        a, b = x.bfloat16(), w.bfloat16()
        c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        row = {"shape": [M, N, K]}
        arms = [("mxfp8_tcgen05", lambda: fp8.mxfp8_gemm_nt_native(xq, xa, wq, wa, out=c)), ("bf16_tcgen05_2cta", lambda: ops.gemm_nt(a, b, c, False, 2)),
                ("bf16_cublas", lambda: torch.mm(a, b.t(), out=c))]
        try:  # cuBLASLt MXFP8: same e4m3 bytes, same swizzled E8M0 scale buffers (rows padded to 128 for both operands)
            e8 = getattr(torch, "float8_e8m0fnu")
            sa, sb = xa.view(e8), fp8.mx_scale_atoms(ws, 128).view(e8)
            torch._scaled_mm(xq, wq.t(), scale_a=sa, scale_b=sb, out_dtype=torch.bfloat16)
            arms.append(("mxfp8_cublaslt", lambda: torch._scaled_mm(xq, wq.t(), scale_a=sa, scale_b=sb, out_dtype=torch.bfloat16)))
        except Exception as e:  # noqa: BLE001
            row["cublaslt_mx_unavailable"] = f"{type(e).__name__}: {str(e)[:120]}"
        for name, fn in arms:
            t = timeit(fn)
            row[name + "_ms"] = t
            row[name + "_tflops"] = fl / t / 1e9
        out["timing"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/mxfp8_check.json", "w"), indent=1)


if __name__ == "__main__":
    main()
