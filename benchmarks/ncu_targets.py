"""Small, deterministic launches of the hot kernels for `ncu --set full` captures (one GPU; see tools/ncu_capture.sh).

    python benchmarks/ncu_targets.py attn | gemm | elementwise
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from vescale_b200.ops import _ext

    _ext.load(required=True)
    ops = torch.ops.vescale_b200
    dev = torch.device("cuda")
    what = sys.argv[1] if len(sys.argv) > 1 else "attn"
    g = torch.Generator(device=dev).manual_seed(0)
    if what == "attn":
        B, S, hq, hk, d = 1, 8192, 32, 8, 128
        qkv = torch.randn(B, S, (hq + 2 * hk) * d, device=dev, generator=g).bfloat16()
        o = torch.empty(B, S, hq * d, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, hq, S, device=dev, dtype=torch.float32)
        do = torch.randn(B, S, hq * d, device=dev, generator=g).bfloat16()
        dqkv, dvec, dq = torch.empty_like(qkv), torch.empty(2, B, hq, S, device=dev, dtype=torch.float32), torch.empty(B, S, hq * d, device=dev, dtype=torch.float32)
        for _ in range(2):
            ops.attn_fwd(qkv, o, lse, hq, hk, 1.0 / math.sqrt(d), 2)
            ops.attn_bwd(qkv, o, do, lse, dqkv, dvec, dq, hq, hk, 1.0 / math.sqrt(d))
    elif what == "gemm":
        for M, N, K in ((8192, 6144, 4096), (8192, 4096, 14336)):
            a = torch.randn(M, K, device=dev, generator=g).bfloat16()
            b = torch.randn(N, K, device=dev, generator=g).bfloat16()
            c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            ops.gemm_nt(a, b, c, False, 2)
            ops.gemm_set_sched(1)
            ops.gemm_nt(a, b, c, False, 2)
            ops.gemm_set_sched(0)
            aq, bq = a.view(torch.uint8)[:, :K].contiguous(), b.view(torch.uint8)[:, :K].contiguous()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
