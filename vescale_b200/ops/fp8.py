"""Block-scaled fp8 (e4m3) linear layers for the fp8 training configuration (BASELINE.json config 5: Llama-3-70B FSDP fp8).

Scaling scheme (DeepSeek-V3-style software block scaling): activations are quantised per token in 1×128 blocks along K,
weights in 128×128 blocks, both to ``float8_e4m3fn`` with fp32 scales ``amax / 448``; the GEMM accumulates in fp32 and the
two scale grids are applied to the partial sums of each 128-wide K block.  Backward runs in bf16 on the saved operands.

Execution:
  * CUDA: ``torch._scaled_mm`` (cuBLASLt fp8 tensor-core GEMM — a *library* call; the hand-written tcgen05 ``kind::f8f6f4``
    variant of ``csrc/gemm_sm100.cu`` is listed under DESIGN.md "known gaps") with the scale layout the installed build
    accepts — 1×128 / 128×128 blockwise first, row-wise as the fallback (scales are then folded per row / column, which is a
    coarser but still block-derived scaling); any failure falls through to the emulated path.
  * everywhere else (and for tests): the *emulated* path — dequantise the fp8 operands block by block and multiply in
    fp32/bf16 — which defines the numerics the tests check.

The reference has no fp8 path; this follows the north-star configuration list, not reference code.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

__all__ = ["quantize_blockwise", "dequantize_blockwise", "fp8_linear", "fp8_gemm_nt", "FP8_MAX", "set_fp8_backend", "quantize_mx", "dequantize_mx", "mxfp8_gemm_nt", "mx_scale_atoms", "mx_scale_from_atoms", "mxfp8_gemm_nt_native", "quantize_mx_fused"]

FP8 = torch.float8_e4m3fn
FP8_MAX = 448.0
_BACKEND = {"mode": "auto", "scaled_mm_ok": None}  # auto | emulate | scaled_mm


def set_fp8_backend(mode: Optional[str] = None, *, mx_native: Optional[bool] = None) -> None:
    """``mode``: block128 recipe backend (auto | emulate | scaled_mm).  ``mx_native``: run MXFP8 GEMMs on the hand-written
    tcgen05 block-scaled kernel (default: the ``VESCALE_B200_MXFP8_NATIVE`` environment flag)."""
    if mode is not None:
        assert mode in ("auto", "emulate", "scaled_mm")
        _BACKEND["mode"] = mode
        _BACKEND["scaled_mm_ok"] = None
    if mx_native is not None:
        _BACKEND["mx_native"] = bool(mx_native)


def _pad_to(x: torch.Tensor, dim: int, mult: int) -> torch.Tensor:
    n = x.shape[dim]
    pad = (-n) % mult
    if pad == 0:
        return x
    shp = list(x.shape)
    shp[dim] = pad
    return torch.cat([x, x.new_zeros(shp)], dim=dim)


def quantize_blockwise(x: torch.Tensor, block: Tuple[int, int] = (1, 128)) -> Tuple[torch.Tensor, torch.Tensor]:
    """2-D ``x`` [R, C] -> (fp8 tensor [R, C], fp32 scales [ceil(R/br), ceil(C/bc)]) with ``x ≈ q * scale`` per block."""
    assert x.dim() == 2
    br, bc = block
    R, C = x.shape
    xp = _pad_to(_pad_to(x.float(), 0, br), 1, bc)
    Rp, Cp = xp.shape
    blocks = xp.view(Rp // br, br, Cp // bc, bc)
    amax = blocks.abs().amax(dim=(1, 3)).clamp_(min=1e-12)
    scale = amax / FP8_MAX
    q = (blocks / scale[:, None, :, None]).clamp_(-FP8_MAX, FP8_MAX).to(FP8)
    return q.view(Rp, Cp)[:R, :C].contiguous(), scale


def dequantize_blockwise(q: torch.Tensor, scale: torch.Tensor, block: Tuple[int, int] = (1, 128), dtype=torch.float32) -> torch.Tensor:
    br, bc = block
    R, C = q.shape
    s = scale.repeat_interleave(br, 0)[:R].repeat_interleave(bc, 1)[:, :C]
    return (q.float() * s).to(dtype)


def _emulated_gemm_nt(xq, xs, wq, ws, out_dtype) -> torch.Tensor:
    """y = dequant(x) @ dequant(w)^T, accumulated in fp32 — what a block-scaled fp8 GEMM computes."""
    x = dequantize_blockwise(xq, xs, (1, 128))
    w = dequantize_blockwise(wq, ws, (128, 128))
    return (x @ w.t()).to(out_dtype)


def _scaled_mm(xq, xs, wq, ws, out_dtype) -> Optional[torch.Tensor]:
    """cuBLASLt fp8 GEMM.  Tries block-wise scales, then row-wise; returns None when the build / shape does not support it."""
    M, K = xq.shape
    N = wq.shape[0]
    if K % 16 or N % 16:
        return None
    b = wq.t()  # [K, N] column-major, as _scaled_mm wants the second operand
    try:  # 1x128 (activations) x 128x128 (weights) block scales
        return torch._scaled_mm(xq, b, scale_a=xs.contiguous(), scale_b=ws.t().contiguous(), out_dtype=out_dtype)
    except Exception:  # noqa: BLE001
        pass
    try:  # row-wise: re-derive one scale per row / per output column from the block grids (max over the K blocks)
        sa = xs.amax(dim=1, keepdim=True)  # [M, 1]
        sb = ws.amax(dim=1).repeat_interleave(128)[:N].view(1, N)  # [1, N]
        xa = (dequantize_blockwise(xq, xs, (1, 128)) / sa).clamp_(-FP8_MAX, FP8_MAX).to(FP8)
        wb = (dequantize_blockwise(wq, ws, (128, 128)) / sb.t()).clamp_(-FP8_MAX, FP8_MAX).to(FP8)
        return torch._scaled_mm(xa, wb.t(), scale_a=sa.contiguous(), scale_b=sb.contiguous(), out_dtype=out_dtype)
    except Exception:  # noqa: BLE001
        return None


def fp8_gemm_nt(xq: torch.Tensor, xs: torch.Tensor, wq: torch.Tensor, ws: torch.Tensor, out_dtype=torch.bfloat16) -> torch.Tensor:
    """``[M, K] fp8 (1x128 scales) x [N, K] fp8 (128x128 scales) -> [M, N]``."""
    mode = _BACKEND["mode"]
    if xq.is_cuda and mode != "emulate" and _BACKEND["scaled_mm_ok"] is not False:
        y = _scaled_mm(xq, xs, wq, ws, out_dtype)
        if y is not None:
            _BACKEND["scaled_mm_ok"] = True
            return y
        _BACKEND["scaled_mm_ok"] = False
        if mode == "scaled_mm":
            raise RuntimeError("torch._scaled_mm rejected the fp8 block-scaled GEMM on this build")
    return _emulated_gemm_nt(xq, xs, wq, ws, out_dtype)


class _Fp8Linear(torch.autograd.Function):
    """y = x @ W^T with both operands block-quantised to e4m3 for the forward GEMM.  Backward: bf16 dgrad / wgrad on the saved
    (unquantised) operands; dW goes straight into ``weight.main_grad`` when the FSDP/DDP wrapper provides one."""

    @staticmethod
    def forward(ctx, x, weight):
        x2 = x.reshape(-1, x.shape[-1])
        xq, xs = quantize_blockwise(x2, (1, 128))
        wq, ws = quantize_blockwise(weight, (128, 128))
        ctx.save_for_backward(x2, weight)
        ctx.shape = x.shape
        y = fp8_gemm_nt(xq, xs, wq, ws, x.dtype if x.dtype in (torch.bfloat16, torch.float16, torch.float32) else torch.bfloat16)
        return y.view(*x.shape[:-1], weight.shape[0]).detach()  # not an autograd view: downstream ops (RoPE) write in place

    @staticmethod
    def backward(ctx, dy):
        from .functional import gemm_nn, gemm_tn

        x2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dx = gemm_nn(dy2, weight).view(ctx.shape) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            mg = getattr(weight, "main_grad", None)
            if mg is not None:
                gemm_tn(dy2, x2.contiguous(), out=mg, accumulate=getattr(weight, "_main_grad_initialised", False))
                weight._main_grad_initialised = True
            else:
                dw = gemm_tn(dy2, x2.contiguous())
        return dx, dw


def fp8_linear(x: torch.Tensor, weight: torch.Tensor, recipe: str = "block128") -> torch.Tensor:
    """``recipe``: "block128" (1x128 activation / 128x128 weight fp32 scales) or "mx" (OCP MXFP8: 1x32 E8M0 scales on both)."""
    if recipe == "mx":
        return _MXFp8Linear.apply(x, weight)
    return _Fp8Linear.apply(x, weight)


# ------------------------------------------------------------------------------- MXFP8 (the tensor cores' native block scaling)
# OCP microscaling: 1x32 blocks along K, one power-of-two scale (E8M0, a biased exponent byte) per block, e4m3 elements.  This
# is the format ``tcgen05.mma.kind::mxf8f6f4.block_scale`` consumes (scale factors staged in TMEM): the quantiser and the
# emulated GEMM below are the numerics specification of the hand-written kernel ``csrc/gemm_mxfp8.cu`` (opt-in until it has
# been validated on hardware: ``set_fp8_backend(mx_native=True)``).
MX_BLOCK = 32


def quantize_mx(x: torch.Tensor):
    """[R, C] -> (e4m3 [R, C], E8M0 scales as uint8 [R, ceil(C/32)]).  scale = 2^(floor(log2(amax)) - 8) so that the block's
    largest element lands in e4m3's top binade (448 = 1.75 * 2^8); zero blocks get the smallest scale."""
    assert x.dim() == 2
    R, C = x.shape
    xp = _pad_to(x.float(), 1, MX_BLOCK)
    blocks = xp.view(R, -1, MX_BLOCK)
    amax = blocks.abs().amax(dim=2)
    exp = torch.floor(torch.log2(amax.clamp(min=2.0**-127))) - 8.0
    exp = exp.clamp(-127.0, 127.0)
    scale = torch.exp2(exp)
    q = (blocks / scale[:, :, None]).clamp_(-FP8_MAX, FP8_MAX).to(FP8)
    e8m0 = (exp + 127.0).to(torch.uint8)  # biased exponent byte
    return q.view(R, -1)[:, :C].contiguous(), e8m0


def dequantize_mx(q: torch.Tensor, e8m0: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    R, C = q.shape
    scale = torch.exp2(e8m0.float() - 127.0).repeat_interleave(MX_BLOCK, dim=1)[:, :C]
    return (q.float() * scale).to(dtype)


def mx_scale_atoms(e8m0: torch.Tensor, row_multiple: int = 128) -> torch.Tensor:
    """[R, K/32] E8M0 bytes -> flat uint8 in the tensor core's scale-factor order (``csrc/gemm_mxfp8.cu``): 512-byte atoms of
    128 rows x 4 K-blocks, byte ``(r % 32) * 16 + (r % 128 // 32) * 4 + k % 4``; the atoms of one 128-row block are consecutive
    along K.  Rows are padded to ``row_multiple`` (128 for the A operand, 256 for B) with the neutral exponent 127."""
    R, KB = e8m0.shape
    if KB % 4:
        raise ValueError("mx_scale_atoms needs K to be a multiple of 128")
    pad = (-R) % row_multiple
    if pad:
        e8m0 = torch.cat([e8m0, e8m0.new_full((pad, KB), 127)], dim=0)
    ra = e8m0.shape[0] // 128
    return e8m0.view(ra, 4, 32, KB // 4, 4).permute(0, 3, 2, 1, 4).contiguous().view(-1)


def mx_scale_from_atoms(atoms: torch.Tensor, rows: int, kblocks: int) -> torch.Tensor:
    """Inverse of ``mx_scale_atoms`` (drops the padding rows)."""
    ra = atoms.numel() // (kblocks // 4 * 512)
    return atoms.view(ra, kblocks // 4, 32, 4, 4).permute(0, 3, 2, 1, 4).reshape(ra * 128, kblocks)[:rows].contiguous()


def _mx_native_enabled() -> bool:
    import os

    return _BACKEND.get("mx_native", os.environ.get("VESCALE_B200_MXFP8_NATIVE", "0") == "1")


def quantize_mx_fused(x: torch.Tensor, row_multiple: int = 128):
    """bf16 [R, K] (K % 128 == 0) on CUDA -> (e4m3 [R, K], E8M0 scales already in atom order) in ONE kernel
    (``csrc/gemm_mxfp8.cu::mx_quantize_kernel``); bit-identical to ``quantize_mx`` + ``mx_scale_atoms``."""
    from . import _ext

    R, K = x.shape
    rp = (R + row_multiple - 1) // row_multiple * row_multiple
    q = torch.empty(R, K, dtype=torch.uint8, device=x.device)
    sf = torch.full((rp // 128 * (K // 128) * 512,), 127, dtype=torch.uint8, device=x.device)
    _ext.count_launch("mx_quantize")
    _ext.ops().mx_quantize(x.contiguous(), q, sf)
    return q.view(FP8), sf


def mxfp8_gemm_nt_native(xq, xs_atoms, wq, ws_atoms, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The hand-written ``tcgen05.mma.kind::mxf8f6f4.block_scale`` kernel; scales already in atom order."""
    from . import _ext

    if out is None:
        out = torch.empty(xq.shape[0], wq.shape[0], dtype=torch.bfloat16, device=xq.device)
    _ext.count_launch("mxfp8_gemm_nt")
    _ext.ops().mxfp8_gemm_nt(xq.view(torch.uint8), xs_atoms, wq.view(torch.uint8), ws_atoms, out)
    return out


def mxfp8_gemm_nt(xq, xs, wq, ws, out_dtype=torch.bfloat16) -> torch.Tensor:
    """MXFP8 GEMM, both operands with 1x32 E8M0 scales along K, fp32 accumulation.  On CUDA with the native kernel enabled
    (``set_fp8_backend(mx_native=True)`` or ``VESCALE_B200_MXFP8_NATIVE=1``) and K % 128 == 0, N % 8 == 0: the tcgen05
    block-scaled kernel; otherwise the emulation (dequantise, multiply in fp32), which is the numerics specification."""
    if xq.is_cuda and _mx_native_enabled() and xq.shape[1] % 128 == 0 and wq.shape[0] % 8 == 0 and out_dtype == torch.bfloat16:
        return mxfp8_gemm_nt_native(xq.contiguous(), mx_scale_atoms(xs, 128), wq.contiguous(), mx_scale_atoms(ws, 256))
    return (dequantize_mx(xq, xs) @ dequantize_mx(wq, ws).t()).to(out_dtype)


class _MXFp8Linear(torch.autograd.Function):
    """``recipe="mx"`` twin of ``_Fp8Linear``: forward GEMM on MXFP8 operands, bf16 backward on the saved operands."""

    @staticmethod
    def forward(ctx, x, weight):
        x2 = x.reshape(-1, x.shape[-1])
        ctx.save_for_backward(x2, weight)
        ctx.shape = x.shape
        K, N = x2.shape[1], weight.shape[0]
        if x2.is_cuda and _mx_native_enabled() and x2.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and K % 128 == 0 and N % 8 == 0:
            xq, xa = quantize_mx_fused(x2, 128)  # one quantise kernel per operand, scales written in the GEMM's atom order
            wq, wa = quantize_mx_fused(weight, 256)
            y = mxfp8_gemm_nt_native(xq, xa, wq, wa)
        else:
            xq, xs = quantize_mx(x2)
            wq, ws = quantize_mx(weight)
            y = mxfp8_gemm_nt(xq, xs, wq, ws, torch.bfloat16).to(x.dtype)
        return y.view(*x.shape[:-1], weight.shape[0]).detach()

    backward = _Fp8Linear.backward
