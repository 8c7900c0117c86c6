"""Differentiable front-ends of the hand-written sm_100a kernels (``csrc/*.cu``) with PyTorch fp32
reference fallbacks for CPU.  Numerics tests compare the two (tests/test_kernels_gpu.py).

Everything operates on *local* tensors: the FSDP/TP wrappers run model compute on local shards and keep
DTensor dispatch off the hot path.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from . import _ext

__all__ = [
    "linear",
    "gemm_nt",
    "rms_norm",
    "add_rms_norm",
    "swiglu",
    "rope_qk_",
    "attention",
    "packed_attention",
    "set_attention_backend",
    "cross_entropy",
    "set_gemm_backend",
]

# auto: the tcgen05 kernel and cuBLAS are timed once per (kind, M, N, K) and the faster one is kept (see ``_autotune``);
# tcgen05 / cublas force one back end.  wgrad (TN, both operands MN-major) stays on cuBLAS, which is ~7 % faster there.
_CFG = {"gemm": "auto", "wgrad": "cublas"}  # gemm: auto | tcgen05 | cublas ; wgrad: cublas | tcgen05


def set_gemm_backend(name: str, wgrad: Optional[str] = None) -> None:
    assert name in ("auto", "tcgen05", "cublas")
    _CFG["gemm"] = name
    _CFG["wgrad"] = wgrad or ("tcgen05" if name == "tcgen05" else "cublas")


def _use_kernels(t: torch.Tensor) -> bool:
    return t.is_cuda and _ext.available()


# ---- per-shape backend selection ("measure, don't guess"): under the 1 kW cap cuBLAS and the tcgen05 kernel are
# within a few percent of each other and the winner depends on the shape (profiles/sustained_r1.json), so in
# ``auto`` mode each (kind, M, N, K) is timed once with both back ends — 40 back-to-back launches each, CUDA events —
# and the faster one is used from then on.
_AUTOTUNE: dict = {}


def _autotune(kind: str, M: int, N: int, K: int, run_ours, run_lib) -> bool:
    key = (kind, M, N, K)
    hit = _AUTOTUNE.get(key)
    if hit is not None:
        return hit
    if torch.cuda.is_current_stream_capturing():
        return True

    def t(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    with torch.no_grad():  # the timing launches write into ``out=`` buffers, which autograd would reject for tracked inputs
        ours, lib = t(run_ours), t(run_lib)
        ours2, lib2 = t(run_ours), t(run_lib)  # second pass under warmed-up clocks/power
    use = min(ours, ours2) <= min(lib, lib2)
    _AUTOTUNE[key] = use
    return use


# =============================================================================== GEMM
def _tcgen05_ok(a: torch.Tensor, b: torch.Tensor) -> bool:
    if _CFG["gemm"] == "cublas" or not _use_kernels(a):
        return False
    M, K = a.shape
    N = b.shape[0]
    return (
        a.dtype == torch.bfloat16
        and b.dtype == torch.bfloat16
        and a.is_contiguous()
        and b.is_contiguous()
        and K % 64 == 0
        and N % 16 == 0
        and M >= 1
    )


def gemm_nt(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *, accumulate: bool = False) -> torch.Tensor:
    """``out (+)= a[M,K] @ b[N,K]^T`` — both operands K-major (the nn.Linear forward shape).
    tcgen05/TMA/TMEM kernel on sm_100a (``csrc/gemm_sm100.cu``); cuBLAS otherwise."""
    pending = getattr(b, "_vb_pending_gather", None)
    if pending is not None:  # FSDP: this weight is gathered by the GEMM kernel itself (all-gather ⊕ first GEMM of the unit)
        return pending(a, out)
    if _tcgen05_ok(a, b) and (out is None or out.is_contiguous()):
        use = True
        if _CFG["gemm"] == "auto" and not accumulate:
            M, K, N = a.shape[0], a.shape[1], b.shape[0]
            tmp = out if out is not None else torch.empty((M, N), dtype=a.dtype, device=a.device)
            use = _autotune("nt", M, N, K, lambda: _ext.ops().gemm_nt(a, b, tmp, False, 0), lambda: torch.mm(a, b.t(), out=tmp))
            out = tmp
        if use:
            _ext.count_launch("gemm_nt")
            if out is None:
                out = torch.empty((a.shape[0], b.shape[0]), dtype=a.dtype, device=a.device)
            _ext.ops().gemm_nt(a, b, out, bool(accumulate), 0)
            return out
    if out is None:
        return a @ b.t()
    if accumulate:
        return out.addmm_(a, b.t())
    return torch.mm(a, b.t(), out=out)


def gemm_nn(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``a[M,K] @ b[K,N]`` (dgrad shape).  Uses the MN-major-B tcgen05 variant when built, else cuBLAS."""
    if _use_kernels(a) and _CFG["gemm"] != "cublas" and hasattr(torch.ops.vescale_b200, "gemm_nn") and a.dtype == torch.bfloat16 and a.is_contiguous() and b.is_contiguous() and a.shape[1] % 64 == 0 and b.shape[1] % 64 == 0:
        if out is None:
            out = torch.empty((a.shape[0], b.shape[1]), dtype=a.dtype, device=a.device)
        use = True
        if _CFG["gemm"] == "auto":
            use = _autotune("nn", a.shape[0], b.shape[1], a.shape[1], lambda: _ext.ops().gemm_nn(a, b, out), lambda: torch.mm(a, b, out=out))
        if use:
            _ext.count_launch("gemm_nn")
            _ext.ops().gemm_nn(a, b, out)
            return out
        return torch.mm(a, b, out=out)
    return torch.mm(a, b, out=out) if out is not None else a @ b


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *, accumulate: bool = False) -> torch.Tensor:
    """``out (+)= a[K,M]^T @ b[K,N]`` (wgrad shape: dW = dY^T X)."""
    if _use_kernels(a) and _CFG["gemm"] != "cublas" and _CFG["wgrad"] == "tcgen05" and hasattr(torch.ops.vescale_b200, "gemm_tn") and a.dtype == torch.bfloat16 and a.is_contiguous() and b.is_contiguous() and a.shape[0] % 64 == 0 and a.shape[1] % 64 == 0 and b.shape[1] % 64 == 0 and (out is None or out.is_contiguous()):
        _ext.count_launch("gemm_tn")
        if out is None:
            out = torch.empty((a.shape[1], b.shape[1]), dtype=a.dtype, device=a.device)
        _ext.ops().gemm_tn(a, b, out, bool(accumulate))
        return out
    if out is None:
        return a.t() @ b
    if accumulate:
        return out.addmm_(a.t(), b)
    return torch.mm(a.t(), b, out=out)


class _Linear(torch.autograd.Function):
    """y = x @ W^T.  Backward writes dW straight into ``weight.main_grad`` when the FSDP/DDP wrapper
    provides one (a view into the unit's flat gradient buffer), so no gradient copy ever happens."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.save_for_backward(x, weight)
        x2 = x.reshape(-1, x.shape[-1])
        y = gemm_nt(x2, weight)
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = gemm_nn(dy2.contiguous(), weight).view(x.shape)
        if ctx.needs_input_grad[1]:
            mg = getattr(weight, "main_grad", None)
            if mg is not None:
                acc = getattr(weight, "_main_grad_initialised", False)
                gemm_tn(dy2.contiguous(), x2.contiguous(), out=mg, accumulate=acc)
                weight._main_grad_initialised = True
                hook = getattr(weight, "_post_main_grad_hook", None)
                if hook is not None:
                    hook(weight)
                dw = None
            else:
                dw = gemm_tn(dy2.contiguous(), x2.contiguous())
        return dx, dw


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    y = _Linear.apply(x, weight)
    return y if bias is None else y + bias


# =============================================================================== RMSNorm
def rms_norm_ref(x, w, eps):
    xf = x.float()
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (xf * rstd * w.float()).to(x.dtype), rstd.squeeze(-1)


def rms_norm_bwd_ref(dy, x, w, rstd):
    xf, dyf, wf = x.float(), dy.float(), w.float()
    r = rstd.unsqueeze(-1)
    xhat = xf * r
    g = dyf * wf
    dx = r * (g - xhat * (g * xhat).mean(-1, keepdim=True))
    dw = (dyf * xhat).reshape(-1, x.shape[-1]).sum(0)
    return dx.to(x.dtype), dw


class _RMSNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps):
        x2 = x.reshape(-1, x.shape[-1])
        if _use_kernels(x) and x.dtype == torch.bfloat16 and x2.is_contiguous():
            _ext.count_launch("rms_norm_fwd")
            y, rstd = _ext.ops().rms_norm_fwd(x2, w, float(eps))
        else:
            y, rstd = rms_norm_ref(x2, w, eps)
        ctx.save_for_backward(x2, w, rstd)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, w, rstd = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        if _use_kernels(dy2) and dy2.dtype == torch.bfloat16:
            _ext.count_launch("rms_norm_bwd", 2)
            dx, dw = _ext.ops().rms_norm_bwd(dy2, x2, w, rstd)
        else:
            dx, dw = rms_norm_bwd_ref(dy2, x2, w, rstd)
        _accumulate_small_grad(w, dw)
        return dx.view(ctx.shape), (None if getattr(w, "main_grad", None) is not None else dw.to(w.dtype)), None


def _accumulate_small_grad(w, dw) -> None:
    """Norm-weight grads also land in the unit gradient buffer when one exists."""
    mg = getattr(w, "main_grad", None)
    if mg is None:
        return
    if getattr(w, "_main_grad_initialised", False):
        mg.add_(dw.to(mg.dtype))
    else:
        mg.copy_(dw)
        w._main_grad_initialised = True
    hook = getattr(w, "_post_main_grad_hook", None)
    if hook is not None:
        hook(w)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return _RMSNorm.apply(x, w, eps)


class _AddRMSNorm(torch.autograd.Function):
    """h = a + b ; y = rmsnorm(h) * w.  Returns (h, y): the residual stream and the normed branch input,
    one pass over HBM instead of three."""

    @staticmethod
    def forward(ctx, a, b, w, eps):
        a2, b2 = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
        if _use_kernels(a) and a.dtype == torch.bfloat16 and a2.is_contiguous() and b2.is_contiguous():
            _ext.count_launch("add_rms_norm_fwd")
            h, y, rstd = _ext.ops().add_rms_norm_fwd(a2, b2, w, float(eps))
        else:
            h = a2 + b2
            y, rstd = rms_norm_ref(h, w, eps)
        ctx.save_for_backward(h, w, rstd)
        ctx.shape = a.shape
        return h.view(a.shape).detach(), y.view(a.shape).detach()

    @staticmethod
    def backward(ctx, dh, dy):
        h, w, rstd = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dh2 = dh.reshape(-1, dh.shape[-1]).contiguous()
        if _use_kernels(dy2) and dy2.dtype == torch.bfloat16:
            _ext.count_launch("add_rms_norm_bwd", 2)
            dx, dw = _ext.ops().add_rms_norm_bwd(dy2, dh2, h, w, rstd)
        else:
            dx, dw = rms_norm_bwd_ref(dy2, h, w, rstd)
            dx = dx + dh2
        _accumulate_small_grad(w, dw)
        dxv = dx.view(ctx.shape)
        return dxv, dxv, (None if getattr(w, "main_grad", None) is not None else dw.to(w.dtype)), None


def add_rms_norm(a, b, w, eps: float = 1e-5):
    return _AddRMSNorm.apply(a, b, w, eps)


# =============================================================================== SwiGLU
class _SwiGLU(torch.autograd.Function):
    """y = silu(gate) * up with gate|up packed along the last dim ([..., 2F] -> [..., F]).
    Only the packed input is saved; the product is recomputed in backward."""

    @staticmethod
    def forward(ctx, gu):
        g2 = gu.reshape(-1, gu.shape[-1])
        if _use_kernels(gu) and gu.dtype == torch.bfloat16 and g2.is_contiguous():
            _ext.count_launch("swiglu_fwd")
            y = _ext.ops().swiglu_fwd(g2)
        else:
            f = g2.shape[-1] // 2
            y = (F.silu(g2[:, :f].float()) * g2[:, f:].float()).to(gu.dtype)
        ctx.save_for_backward(g2)
        ctx.shape = gu.shape
        return y.view(*gu.shape[:-1], gu.shape[-1] // 2)

    @staticmethod
    def backward(ctx, dy):
        (g2,) = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        if _use_kernels(dy2) and dy2.dtype == torch.bfloat16:
            _ext.count_launch("swiglu_bwd")
            dgu = _ext.ops().swiglu_bwd(dy2, g2)
        else:
            f = g2.shape[-1] // 2
            g, u, d = g2[:, :f].float(), g2[:, f:].float(), dy2.float()
            s = torch.sigmoid(g)
            dg = d * u * s * (1 + g * (1 - s))
            du = d * g * s
            dgu = torch.cat([dg, du], -1).to(g2.dtype)
        return dgu.view(ctx.shape)


def swiglu(gate_up: torch.Tensor) -> torch.Tensor:
    return _SwiGLU.apply(gate_up)


# =============================================================================== RoPE
def rope_tables(seq_len: int, head_dim: int, theta: float, device, scaling: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32, device=device) / head_dim))
    if scaling:  # Llama-3.1 style frequency scaling
        factor, lo, hi, orig = scaling.get("factor", 8.0), scaling.get("low_freq_factor", 1.0), scaling.get("high_freq_factor", 4.0), scaling.get("original_max_position_embeddings", 8192)
        wavelen = 2 * math.pi / inv
        smooth = ((orig / wavelen) - lo) / (hi - lo)
        inv = torch.where(wavelen > orig / lo, inv / factor, torch.where(wavelen < orig / hi, inv, (1 - smooth) * inv / factor + smooth * inv))
    t = torch.arange(seq_len, dtype=torch.float32, device=device)
    fr = torch.outer(t, inv)
    return fr.cos().contiguous(), fr.sin().contiguous()


def _rope_ref_(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, sign: float) -> None:
    # x: [B, S, H, D] (view into the packed qkv tensor); rotate-half convention
    d2 = x.shape[-1] // 2
    c, s = cos[None, :, None, :].float(), sin[None, :, None, :].float() * sign
    x1, x2 = x[..., :d2].float(), x[..., d2:].float()
    o1, o2 = x1 * c - x2 * s, x2 * c + x1 * s
    x[..., :d2] = o1.to(x.dtype)
    x[..., d2:] = o2.to(x.dtype)


class _RopeQK(torch.autograd.Function):
    """In-place rotary embedding of the q and k slices of a packed qkv activation [B, S, (Hq+2Hk)*D]."""

    @staticmethod
    def forward(ctx, qkv, cos, sin, n_q, n_kv, head_dim):
        ctx.save_for_backward(cos, sin)
        ctx.cfg = (n_q, n_kv, head_dim)
        _rope_apply_(qkv, cos, sin, n_q, n_kv, head_dim, 1.0)
        ctx.mark_dirty(qkv)
        return qkv

    @staticmethod
    def backward(ctx, dqkv):
        cos, sin = ctx.saved_tensors
        n_q, n_kv, head_dim = ctx.cfg
        dqkv = dqkv.contiguous()
        _rope_apply_(dqkv, cos, sin, n_q, n_kv, head_dim, -1.0)
        return dqkv, None, None, None, None, None


def _rope_apply_(qkv, cos, sin, n_q, n_kv, head_dim, sign):
    B, S = qkv.shape[0], qkv.shape[1]
    if _use_kernels(qkv) and qkv.dtype == torch.bfloat16 and qkv.is_contiguous():
        _ext.count_launch("rope_qk")
        _ext.ops().rope_qk_(qkv.view(B * S, -1), cos, sin, int(S), int(n_q), int(n_kv), int(head_dim), float(sign))
        return
    qk = qkv[..., : (n_q + n_kv) * head_dim].unflatten(-1, (n_q + n_kv, head_dim))
    _rope_ref_(qk, cos[:S], sin[:S], sign)


def rope_qk_(qkv, cos, sin, n_q: int, n_kv: int, head_dim: int):
    return _RopeQK.apply(qkv, cos, sin, n_q, n_kv, head_dim)


# =============================================================================== attention
_SDPA_PRIORITY = None


def _sdpa_priority():
    """cuDNN's Blackwell attention first: measured on B200 at S=8192, 32q/8kv heads, d=128 it runs forward at
    1347 TFLOPS and fwd+bwd at 988 TFLOPS, against 331 / 291 for the FA2 (mma.sync) backend that SDPA picks by
    default (profiles/micro_r1.json)."""
    global _SDPA_PRIORITY
    if _SDPA_PRIORITY is None:
        from torch.nn.attention import SDPBackend

        _SDPA_PRIORITY = [SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION, SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]
    return _SDPA_PRIORITY


def attention(q, k, v, *, causal: bool = True) -> torch.Tensor:
    """q [B,Hq,S,D], k/v [B,Hkv,S,D] (strided views are fine).  Library attention through SDPA, as the
    reference uses SDPA / flash_attn; on CUDA the cuDNN backend is given priority."""
    gqa = q.shape[1] != k.shape[1]
    if q.is_cuda:
        from torch.nn.attention import sdpa_kernel

        with sdpa_kernel(_sdpa_priority(), set_priority=True):
            return F.scaled_dot_product_attention(q, k, v, is_causal=causal, enable_gqa=gqa)
    return F.scaled_dot_product_attention(q, k, v, is_causal=causal, enable_gqa=gqa)


class _PackedAttention(torch.autograd.Function):
    """Causal GQA attention straight from / to the packed qkv activation [B, S, (Hq+2Hk)*D].

    Autograd over three slices of one tensor materialises three zero-filled full-size gradients and adds them
    (9 elementwise launches and ~0.6 GB of traffic per layer at 8k tokens); here the backward assembles dqkv with
    one concatenation."""

    @staticmethod
    def forward(ctx, qkv, n_q, n_kv, head_dim, causal):
        B, S, _ = qkv.shape
        ctx.cfg = (n_q, n_kv, head_dim, causal)
        with torch.enable_grad():
            src = qkv.detach().requires_grad_(True)
            q = src[..., : n_q * head_dim].unflatten(-1, (n_q, head_dim)).transpose(1, 2)
            k = src[..., n_q * head_dim : (n_q + n_kv) * head_dim].unflatten(-1, (n_kv, head_dim)).transpose(1, 2)
            v = src[..., (n_q + n_kv) * head_dim :].unflatten(-1, (n_kv, head_dim)).transpose(1, 2)
            q, k, v = q.detach().requires_grad_(True), k.detach().requires_grad_(True), v.detach().requires_grad_(True)
            o = attention(q, k, v, causal=causal)
        ctx.graph = (q, k, v, o)
        return o.detach().transpose(1, 2).reshape(B, S, n_q * head_dim)

    @staticmethod
    def backward(ctx, do):
        q, k, v, o = ctx.graph  # kept: the node may run again over a retained graph (zero-bubble W pass)
        n_q, n_kv, head_dim, _ = ctx.cfg
        B, S = do.shape[0], do.shape[1]
        do4 = do.view(B, S, n_q, head_dim).transpose(1, 2)
        dq, dk, dv = torch.autograd.grad(o, (q, k, v), do4, retain_graph=True)
        dqkv = torch.cat(
            [dq.transpose(1, 2).reshape(B, S, -1), dk.transpose(1, 2).reshape(B, S, -1), dv.transpose(1, 2).reshape(B, S, -1)], dim=-1
        )
        return dqkv, None, None, None, None


# attention back end: "cudnn" = library SDPA (cuDNN's Blackwell kernels), "tcgen05" = the hand-written sm_100a kernels of
# csrc/attention_sm100.cu (forward + backward, head dim 128, causal, S % 128 == 0), "auto" = whichever is faster for the shape
# (timed once, like the GEMM back ends).  VESCALE_B200_ATTN / set_attention_backend().
_ATTN = {"backend": None}
_ATTN_TUNE: dict = {}


def set_attention_backend(name: str) -> None:
    assert name in ("auto", "cudnn", "tcgen05")
    _ATTN["backend"] = name


def _attn_backend() -> str:
    if _ATTN["backend"] is None:
        import os

        _ATTN["backend"] = os.environ.get("VESCALE_B200_ATTN", "auto")
    return _ATTN["backend"]


def _native_attn_ok(qkv: torch.Tensor, n_q: int, n_kv: int, head_dim: int, causal: bool) -> bool:
    return (causal and head_dim == 128 and qkv.is_cuda and qkv.dtype == torch.bfloat16 and qkv.dim() == 3 and qkv.shape[1] % 128 == 0 and n_q % n_kv == 0
            and _ext.available() and hasattr(_ext.ops(), "attn_bwd"))


class _NativeAttention(torch.autograd.Function):
    """Causal GQA flash attention on the hand-written tcgen05 kernels, straight from / to the packed qkv activation: the forward
    keeps only the output and the per-row log-sum-exp, the backward writes dq | dk | dv into one packed gradient."""

    @staticmethod
    def forward(ctx, qkv, n_q, n_kv, head_dim):
        B, S, _ = qkv.shape
        qkv = qkv.contiguous()
        out = torch.empty(B, S, n_q * head_dim, dtype=qkv.dtype, device=qkv.device)
        lse = torch.empty(B, n_q, S, dtype=torch.float32, device=qkv.device)
        _ext.count_launch("attn_fwd")
        _ext.ops().attn_fwd(qkv, out, lse, n_q, n_kv, 1.0 / math.sqrt(head_dim))
        ctx.save_for_backward(qkv, out, lse)
        ctx.cfg = (n_q, n_kv, head_dim)
        return out

    @staticmethod
    def backward(ctx, do):
        qkv, out, lse = ctx.saved_tensors
        n_q, n_kv, head_dim = ctx.cfg
        B, S, _ = qkv.shape
        dqkv = torch.empty_like(qkv)
        dvec = torch.empty(2 * lse.numel(), dtype=torch.float32, device=qkv.device)  # D | lse * log2(e)
        dq_acc = torch.empty(B, S, n_q * head_dim, dtype=torch.float32, device=qkv.device)
        _ext.count_launch("attn_bwd")
        _ext.ops().attn_bwd(qkv, out, do.contiguous(), lse, dqkv, dvec, dq_acc, n_q, n_kv, 1.0 / math.sqrt(head_dim))
        return dqkv, None, None, None


def _attn_pick_native(qkv, n_q, n_kv, head_dim) -> bool:
    """auto: time forward + backward of both back ends once per (B, S, Hq, Hkv) and keep the faster."""
    key = (tuple(qkv.shape), n_q, n_kv)
    hit = _ATTN_TUNE.get(key)
    if hit is not None:
        return hit
    if torch.cuda.is_current_stream_capturing():
        return False

    def t(fn):
        src = qkv.detach().clone().requires_grad_(True)
        for _ in range(2):
            fn(src).sum().backward()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(src).sum().backward()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    with torch.enable_grad():
        ours = t(lambda x: _NativeAttention.apply(x, n_q, n_kv, head_dim))
        lib = t(lambda x: _PackedAttention.apply(x, n_q, n_kv, head_dim, True))
    _ATTN_TUNE[key] = ours <= lib
    return _ATTN_TUNE[key]


def packed_attention(qkv: torch.Tensor, n_q: int, n_kv: int, head_dim: int, causal: bool = True) -> torch.Tensor:
    """qkv [B, S, (Hq+2Hk)*D] (RoPE already applied) -> attention output [B, S, Hq*D]."""
    be = _attn_backend()
    if be != "cudnn" and _native_attn_ok(qkv, n_q, n_kv, head_dim, causal):
        if be == "tcgen05" or _attn_pick_native(qkv, n_q, n_kv, head_dim):
            return _NativeAttention.apply(qkv, n_q, n_kv, head_dim)
    return _PackedAttention.apply(qkv, n_q, n_kv, head_dim, causal)


# =============================================================================== cross entropy
class _CrossEntropy(torch.autograd.Function):
    """Mean token cross-entropy over rows of ``logits`` [T, V]; the gradient is written *in place* into the
    logits buffer during forward (one read + one write of the 2 GB logits instead of materialising fp32
    log-probs, their grad, and a softmax).  ``ignore_index`` rows contribute nothing."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        T, V = logits.shape
        n_valid = (target != ignore_index).sum().clamp(min=1).to(torch.float32)
        if _use_kernels(logits) and logits.dtype == torch.bfloat16 and logits.is_contiguous():
            _ext.count_launch("cross_entropy")
            losses = _ext.ops().cross_entropy_fwd_bwd_(logits, target, n_valid, int(ignore_index))
            loss = losses.sum() / n_valid
        else:
            lf = logits.float()
            lse = torch.logsumexp(lf, -1)
            valid = target != ignore_index
            tgt = target.clamp(min=0)
            picked = lf.gather(1, tgt[:, None]).squeeze(1)
            loss = torch.where(valid, lse - picked, torch.zeros_like(lse)).sum() / n_valid
            grad = torch.softmax(lf, -1)
            grad.scatter_add_(1, tgt[:, None], -torch.ones_like(picked)[:, None])
            grad = torch.where(valid[:, None], grad / n_valid, torch.zeros_like(grad))
            logits.copy_(grad.to(logits.dtype))
        # ``logits`` now holds d loss / d logits; it is not an output, so it is stashed directly
        # (save_for_backward would trip the version check on the in-place write)
        ctx.grad_buf = logits
        return loss

    @staticmethod
    def backward(ctx, dloss):
        grad = ctx.grad_buf
        # dloss is 1.0 in ordinary training; the in-place scale keeps the 2 GB buffer single.  The node may
        # run twice (zero-bubble pipeline: input-gradient pass, then weight-gradient pass over the retained
        # graph) — scale once only.
        if not getattr(ctx, "scaled", False):
            grad.mul_(dloss.to(grad.dtype))
            ctx.scaled = True
        return grad, None, None


def cross_entropy(logits: torch.Tensor, target: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """NOTE: consumes ``logits`` (overwritten with d loss / d logits)."""
    return _CrossEntropy.apply(logits, target, ignore_index)


# =============================================================================== fused-with-recompute blocks
class _NormLinear(torch.autograd.Function):
    """y = rmsnorm(h) @ W^T where only ``h`` and ``rstd`` are kept: the normed activation is recomputed in
    backward (one cheap bandwidth pass) instead of being stored (64 MB/layer at 8k tokens x 4096)."""

    @staticmethod
    def forward(ctx, h, nw, weight, eps):
        h2 = h.reshape(-1, h.shape[-1])
        if _use_kernels(h) and h.dtype == torch.bfloat16 and h2.is_contiguous():
            _ext.count_launch("rms_norm_fwd")
            n, rstd = _ext.ops().rms_norm_fwd(h2, nw, float(eps))
        else:
            n, rstd = rms_norm_ref(h2, nw, eps)
        y = gemm_nt(n, weight)
        ctx.save_for_backward(h2, nw, weight, rstd)
        ctx.shape = h.shape
        ctx.eps = eps
        return y.view(*h.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        h2, nw, weight, rstd = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        kern = _use_kernels(h2) and h2.dtype == torch.bfloat16
        if kern:
            _ext.count_launch("rms_norm_fwd")
            n, _ = _ext.ops().rms_norm_fwd(h2, nw, float(ctx.eps))
        else:
            n, _ = rms_norm_ref(h2, nw, ctx.eps)
        _wgrad(weight, dy2, n)
        dn = gemm_nn(dy2, weight)
        del n
        if kern:
            _ext.count_launch("rms_norm_bwd", 2)
            dh, dnw = _ext.ops().rms_norm_bwd(dn, h2, nw, rstd)
        else:
            dh, dnw = rms_norm_bwd_ref(dn, h2, nw, rstd)
        _accumulate_small_grad(nw, dnw)
        return (
            dh.view(ctx.shape),
            None if getattr(nw, "main_grad", None) is not None else dnw.to(nw.dtype),
            None if getattr(weight, "main_grad", None) is not None else _pop_dw(weight),
            None,
        )


def _wgrad(weight, dy2, x2) -> None:
    mg = getattr(weight, "main_grad", None)
    if mg is not None:
        gemm_tn(dy2, x2, out=mg, accumulate=getattr(weight, "_main_grad_initialised", False))
        weight._main_grad_initialised = True
        hook = getattr(weight, "_post_main_grad_hook", None)
        if hook is not None:
            hook(weight)
    else:
        weight._tmp_dw = gemm_tn(dy2, x2)


def _pop_dw(weight):
    dw = weight._tmp_dw
    del weight._tmp_dw
    return dw


def norm_linear(h, norm_weight, weight, eps: float = 1e-5):
    return _NormLinear.apply(h, norm_weight, weight, eps)


class _SwiGLULinear(torch.autograd.Function):
    """y = (silu(g) * u) @ W^T keeping only the packed ``gate|up`` tensor; the product (224 MB/layer at
    8k tokens x 14336) is recomputed in backward."""

    @staticmethod
    def forward(ctx, gu, weight):
        g2 = gu.reshape(-1, gu.shape[-1])
        act = _swiglu_fwd(g2)
        y = gemm_nt(act, weight)
        ctx.save_for_backward(g2, weight)
        ctx.shape = gu.shape
        return y.view(*gu.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        g2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        act = _swiglu_fwd(g2)
        _wgrad(weight, dy2, act)
        del act
        dact = gemm_nn(dy2, weight)
        if _use_kernels(dact) and dact.dtype == torch.bfloat16:
            _ext.count_launch("swiglu_bwd")
            dgu = _ext.ops().swiglu_bwd(dact, g2)
        else:
            dgu = _swiglu_bwd_ref(dact, g2)
        return dgu.view(ctx.shape), (None if getattr(weight, "main_grad", None) is not None else _pop_dw(weight))


def _swiglu_fwd(g2):
    if _use_kernels(g2) and g2.dtype == torch.bfloat16 and g2.is_contiguous():
        _ext.count_launch("swiglu_fwd")
        return _ext.ops().swiglu_fwd(g2)
    f = g2.shape[-1] // 2
    return (F.silu(g2[:, :f].float()) * g2[:, f:].float()).to(g2.dtype)


def _swiglu_bwd_ref(dy2, g2):
    f = g2.shape[-1] // 2
    g, u, d = g2[:, :f].float(), g2[:, f:].float(), dy2.float()
    s = torch.sigmoid(g)
    return torch.cat([d * u * s * (1 + g * (1 - s)), d * g * s], -1).to(g2.dtype)


def swiglu_linear(gate_up, weight):
    return _SwiGLULinear.apply(gate_up, weight)


class _AddNormLinear(torch.autograd.Function):
    """(h', y) = (h + delta, rmsnorm(h + delta) @ W^T): residual add, norm and the unit's first GEMM in one
    autograd node.  Saves only h' and rstd; the normed activation is recomputed in backward."""

    @staticmethod
    def forward(ctx, h, delta, nw, weight, eps):
        h2, d2 = h.reshape(-1, h.shape[-1]), delta.reshape(-1, delta.shape[-1])
        kern = _use_kernels(h) and h.dtype == torch.bfloat16 and h2.is_contiguous() and d2.is_contiguous()
        if kern:
            _ext.count_launch("add_rms_norm_fwd")
            hn, n, rstd = _ext.ops().add_rms_norm_fwd(h2, d2, nw, float(eps))
        else:
            hn = h2 + d2
            n, rstd = rms_norm_ref(hn, nw, eps)
        y = gemm_nt(n, weight)
        ctx.save_for_backward(hn, nw, weight, rstd)
        ctx.shape = h.shape
        ctx.eps = eps
        # .detach() drops the view marker so downstream in-place ops (RoPE) on an output are legal
        return hn.view(h.shape).detach(), y.view(*h.shape[:-1], weight.shape[0]).detach()

    @staticmethod
    def backward(ctx, dhn, dy):
        hn, nw, weight, rstd = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dhn2 = dhn.reshape(-1, dhn.shape[-1]).contiguous()
        kern = _use_kernels(hn) and hn.dtype == torch.bfloat16
        if kern:
            _ext.count_launch("rms_norm_fwd")
            n, _ = _ext.ops().rms_norm_fwd(hn, nw, float(ctx.eps))
        else:
            n, _ = rms_norm_ref(hn, nw, ctx.eps)
        _wgrad(weight, dy2, n)
        del n
        dn = gemm_nn(dy2, weight)
        if kern:
            _ext.count_launch("add_rms_norm_bwd", 2)
            dh, dnw = _ext.ops().add_rms_norm_bwd(dn, dhn2, hn, nw, rstd)
        else:
            dh, dnw = rms_norm_bwd_ref(dn, hn, nw, rstd)
            dh = dh + dhn2
        _accumulate_small_grad(nw, dnw)
        dhv = dh.view(ctx.shape)
        return (
            dhv,
            dhv,
            None if getattr(nw, "main_grad", None) is not None else dnw.to(nw.dtype),
            None if getattr(weight, "main_grad", None) is not None else _pop_dw(weight),
            None,
        )


def add_norm_linear(h, delta, norm_weight, weight, eps: float = 1e-5):
    """Returns (h + delta, rmsnorm(h + delta) @ weight^T)."""
    return _AddNormLinear.apply(h, delta, norm_weight, weight, eps)
