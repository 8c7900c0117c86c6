"""Bring-up + timing of the hand-written tcgen05 flash-attention kernels (csrc/attention_sm100.cu) against SDPA.

    timeout 200 python benchmarks/attn_check.py [--bwd]
"""
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def ref_attention(qkv, hq, hk, d):
    B, S, _ = qkv.shape
    q = qkv[..., : hq * d].view(B, S, hq, d).transpose(1, 2).float()
    k = qkv[..., hq * d : (hq + hk) * d].view(B, S, hk, d).transpose(1, 2).float()
    v = qkv[..., (hq + hk) * d :].view(B, S, hk, d).transpose(1, 2).float()
    o = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=hq != hk)
    return o.transpose(1, 2).reshape(B, S, hq * d)


def main():
    from vescale_b200.ops import _ext

    _ext.load(required=True)
    ops = torch.ops.vescale_b200
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    d = 128
    out = {"numerics": [], "timing": []}
    ok_all = True
    bwd_only = "--bwd-only" in sys.argv  # quick mode: backward numerics + one backward timing row (own vs cuDNN)
    for B, S, hq, hk in (() if bwd_only else ((1, 128, 1, 1), (1, 256, 2, 1), (1, 512, 4, 2), (2, 1024, 8, 2), (1, 4096, 4, 1))):
        qkv = (torch.randn(B, S, (hq + 2 * hk) * d, device=dev, generator=g) * 1.0).bfloat16()
        ref = ref_attention(qkv, hq, hk, d)
        for variant in (1, 2, 3):
            o = torch.full((B, S, hq * d), float("nan"), device=dev, dtype=torch.bfloat16)
            lse = torch.full((B, hq, S), float("nan"), device=dev, dtype=torch.float32)
            ops.attn_fwd(qkv, o, lse, hq, hk, 1.0 / math.sqrt(d), variant)
            torch.cuda.synchronize()
            err = (o.float() - ref).abs().max().item()
            if not (bool(torch.isfinite(o.float()).all()) and err < 0.02):
                print(f"attn_fwd variant {variant} B{B} S{S} Hq{hq} Hkv{hk}: out err {err:.4f} MISMATCH", flush=True)
                ok_all = False
        # reference lse
        q = qkv[..., : hq * d].view(B, S, hq, d).transpose(1, 2).float()
        k = qkv[..., hq * d : (hq + hk) * d].view(B, S, hk, d).transpose(1, 2).float().repeat_interleave(hq // hk, 1)
        sc = (q @ k.transpose(-1, -2)) / math.sqrt(d)
        sc = sc.masked_fill(torch.triu(torch.ones(S, S, device=dev, dtype=torch.bool), 1), float("-inf"))
        lse_ref = torch.logsumexp(sc, -1)
        lerr = (lse - lse_ref).abs().max().item()
        ok = bool(torch.isfinite(o.float()).all()) and err < 0.02 and lerr < 0.02
        ok_all &= ok
        out["numerics"].append({"shape": [B, S, hq, hk], "max_abs_err": err, "lse_err": lerr, "ok": ok})
        print(f"attn_fwd B{B} S{S} Hq{hq} Hkv{hk}: out err {err:.4f} lse err {lerr:.4f} {'ok' if ok else 'MISMATCH'}", flush=True)
    if not ok_all:
        print(json.dumps(out))
        sys.exit(1)

    # ---- backward: dqkv against autograd through the fp32 reference
    for B, S, hq, hk in ((1, 128, 1, 1), (1, 256, 2, 1), (1, 512, 4, 2), (2, 1024, 8, 2)):
        qkv = (torch.randn(B, S, (hq + 2 * hk) * d, device=dev, generator=g) * 1.0).bfloat16()
        do = (torch.randn(B, S, hq * d, device=dev, generator=g) * 1.0).bfloat16()
        o = torch.empty(B, S, hq * d, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, hq, S, device=dev, dtype=torch.float32)
        ops.attn_fwd(qkv, o, lse, hq, hk, 1.0 / math.sqrt(d))
        dqkv = torch.full_like(qkv, float("nan"))
        dvec = torch.empty(2, B, hq, S, device=dev, dtype=torch.float32)
        dq_acc = torch.empty(B, S, hq * d, device=dev, dtype=torch.float32)
        ops.attn_bwd(qkv, o, do, lse, dqkv, dvec, dq_acc, hq, hk, 1.0 / math.sqrt(d))
        torch.cuda.synchronize()
        x = qkv.float().requires_grad_()
        ref_attention(x, hq, hk, d).backward(do.float())
        gref = x.grad
        errs = {}
        for name, lo, hi in (("dq", 0, hq * d), ("dk", hq * d, (hq + hk) * d), ("dv", (hq + hk) * d, (hq + 2 * hk) * d)):
            a, r_ = dqkv[..., lo:hi].float(), gref[..., lo:hi]
            errs[name] = ((a - r_).abs().max() / (r_.abs().max() + 1e-6)).item()
        ok = bool(torch.isfinite(dqkv.float()).all()) and max(errs.values()) < 0.03
        ok_all &= ok
        out["numerics"].append({"bwd_shape": [B, S, hq, hk], **errs, "ok": ok})
        print(f"attn_bwd B{B} S{S} Hq{hq} Hkv{hk}: rel err dq {errs['dq']:.4f} dk {errs['dk']:.4f} dv {errs['dv']:.4f} {'ok' if ok else 'MISMATCH'}", flush=True)
    if not ok_all:
        print(json.dumps(out))
        sys.exit(1)

    def timeit(fn, n=10):
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

    from torch.nn.attention import SDPBackend, sdpa_kernel

    for B, S, hq, hk in (((1, 8192, 32, 8),) if bwd_only else ((1, 8192, 32, 8), (2, 4096, 32, 8))):
        qkv = torch.randn(B, S, (hq + 2 * hk) * d, device=dev, generator=g).bfloat16()
        o = torch.empty(B, S, hq * d, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, hq, S, device=dev, dtype=torch.float32)
        fl = 4.0 * B * hq * S * S * d / 2  # causal
        ms1 = 0.0 if bwd_only else timeit(lambda: ops.attn_fwd(qkv, o, lse, hq, hk, 1.0 / math.sqrt(d), 1))
        ms2 = 0.0 if bwd_only else timeit(lambda: ops.attn_fwd(qkv, o, lse, hq, hk, 1.0 / math.sqrt(d), 2))
        ms = timeit(lambda: ops.attn_fwd(qkv, o, lse, hq, hk, 1.0 / math.sqrt(d), 3))
        q = qkv[..., : hq * d].view(B, S, hq, d).transpose(1, 2)
        k = qkv[..., hq * d : (hq + hk) * d].view(B, S, hk, d).transpose(1, 2)
        v = qkv[..., (hq + hk) * d :].view(B, S, hk, d).transpose(1, 2)
        with sdpa_kernel([SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION], set_priority=True):
            ms_c = timeit(lambda: F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True))
        do = torch.randn(B, S, hq * d, device=dev, generator=g).bfloat16()
        dqkv = torch.empty_like(qkv)
        dvec = torch.empty(2, B, hq, S, device=dev, dtype=torch.float32)
        dq_acc = torch.empty(B, S, hq * d, device=dev, dtype=torch.float32)
        ms_b = timeit(lambda: ops.attn_bwd(qkv, o, do, lse, dqkv, dvec, dq_acc, hq, hk, 1.0 / math.sqrt(d)))
        qg, kg, vg = (t.detach().clone().requires_grad_() for t in (q, k, v))
        with sdpa_kernel([SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION], set_priority=True):
            oc = F.scaled_dot_product_attention(qg, kg, vg, is_causal=True, enable_gqa=True)
            do4 = do.view(B, S, hq, d).transpose(1, 2)
            ms_cb = timeit(lambda: torch.autograd.grad(oc, (qg, kg, vg), do4, retain_graph=True))
        row = {"shape": [B, S, hq, hk], "ours_bwd_ms": ms_b, "ours_bwd_tflops": 2.5 * fl / ms_b / 1e9, "cudnn_bwd_ms": ms_cb, "cudnn_bwd_tflops": 2.5 * fl / ms_cb / 1e9, "ours_fwd_v1_ms": ms1, "ours_fwd_v2_ms": ms2, "ours_fwd_ms": ms, "ours_fwd_tflops": fl / ms / 1e9, "cudnn_fwd_ms": ms_c, "cudnn_fwd_tflops": fl / ms_c / 1e9}
        out["timing"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.environ.get("ATTN_CHECK_OUT", "gpurun_out/attn_check.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
