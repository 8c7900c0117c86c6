"""VocabParallelCrossEntropy: Megatron's algorithm on vocab-sharded logits with label smoothing
(legacy ``model/patch/vp_cross_entropy.py:43-147``): local max → AR(MAX); masked target logit → AR(SUM);
local sum-exp → AR(SUM).  ``loss_parallel`` is the DTensor-dispatch route to the same math."""
from __future__ import annotations

import torch

from ...comm import collectives as C
from ...dtensor.api import DTensor
from ...layout import compute_local_shape_and_global_offset
from ...placement import Shard


class _VPCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits_local, target, mesh, mesh_dim, vocab_start, label_smoothing, vocab_size):
        lmax = logits_local.amax(-1)
        gmax = C.mesh_all_reduce(lmax, mesh, "max", mesh_dim)
        x = logits_local.float() - gmax.unsqueeze(-1)
        n = x.shape[-1]
        mask = (target < vocab_start) | (target >= vocab_start + n)
        idx = (target - vocab_start).masked_fill(mask, 0)
        picked = x.gather(-1, idx.unsqueeze(-1)).squeeze(-1).masked_fill(mask, 0.0)
        picked = C.mesh_all_reduce(picked, mesh, "sum", mesh_dim)
        ex = x.exp()
        sumexp = C.mesh_all_reduce(ex.sum(-1), mesh, "sum", mesh_dim)
        loss = sumexp.log() - picked
        softmax = ex / sumexp.unsqueeze(-1)
        if label_smoothing > 0:
            s = label_smoothing * vocab_size / (vocab_size - 1)
            mean_logp = C.mesh_all_reduce((x - sumexp.log().unsqueeze(-1)).sum(-1), mesh, "sum", mesh_dim) / vocab_size
            loss = (1 - s) * loss - s * mean_logp
        ctx.save_for_backward(softmax, mask, idx)
        ctx.ls, ctx.V = label_smoothing, vocab_size
        return loss

    @staticmethod
    def backward(ctx, g):
        softmax, mask, idx = ctx.saved_tensors
        grad = softmax.clone()
        upd = (~mask).to(grad.dtype)
        if ctx.ls > 0:
            s = ctx.ls * ctx.V / (ctx.V - 1)
            grad.scatter_add_(-1, idx.unsqueeze(-1), -(1 - s) * upd.unsqueeze(-1))
            grad -= s / ctx.V
        else:
            grad.scatter_add_(-1, idx.unsqueeze(-1), -upd.unsqueeze(-1))
        return grad * g.unsqueeze(-1), None, None, None, None, None, None


class VocabParallelCrossEntropy:
    @staticmethod
    def apply(logits: DTensor, target, label_smoothing: float = 0.0) -> torch.Tensor:
        """``logits``: DTensor sharded on the last (vocab) dim; ``target``: replicated tensor/DTensor. Per-token loss."""
        mesh = logits.device_mesh
        md = next(i for i, p in enumerate(logits.placements) if isinstance(p, Shard) and p.dim == logits.ndim - 1)
        _, off = compute_local_shape_and_global_offset(logits.shape, mesh, logits.placements)
        t = target._local_tensor if isinstance(target, DTensor) else target
        return _VPCE.apply(logits.to_local(), t, mesh, md, off[-1], label_smoothing, logits.shape[-1])

    @staticmethod
    def mean(logits: DTensor, target, ignore_index: int = -100) -> torch.Tensor:
        """Mean token loss.  When the symmetric-memory collectives are enabled on the vocab mesh dim (and the logits are an
        evenly sharded bf16 CUDA tensor) this is ONE kernel launch — local max/sum-exp, a W-way statistics exchange through
        peer memory and the gradient written in place — and it CONSUMES the logits buffer.  Otherwise the
        three-all-reduce algorithm above."""
        from ...comm.symm_collectives import symm_backend_for

        mesh = logits.device_mesh
        md = next(i for i, p in enumerate(logits.placements) if isinstance(p, Shard) and p.dim == logits.ndim - 1)
        t = target._local_tensor if isinstance(target, DTensor) else target
        local = logits.to_local()
        sc = symm_backend_for(mesh.get_group(md), local) if mesh.size(md) > 1 else None
        V, W = logits.shape[-1], mesh.size(md)
        if sc is not None and local.dtype == torch.bfloat16 and local.is_contiguous() and V % (8 * W) == 0:
            return sc.vocab_parallel_cross_entropy(local, t, ignore_index)
        per_tok = VocabParallelCrossEntropy.apply(logits, t.clamp(min=0))
        valid = (t != ignore_index).to(per_tok.dtype)
        return (per_tok * valid).sum() / valid.sum().clamp(min=1)

    @staticmethod
    def patch(root: torch.nn.Module) -> None:
        """Post-patch every plain ``nn.CrossEntropyLoss`` under ``root`` (no class weights, no label smoothing — those keep the
        stock forward, with a warning): when its input arrives as a DTensor sharded on the class dim, the loss is computed
        vocab-parallel instead of gathering the logits (legacy ``model/patch/vp_cross_entropy.py:182-230``)."""
        import types
        import warnings

        for path, sub in root.named_modules():
            if not isinstance(sub, torch.nn.CrossEntropyLoss) or getattr(sub, "_vb_vp_patched", False):
                continue
            if sub.weight is not None or sub.label_smoothing != 0.0:
                warnings.warn(f"CrossEntropyLoss `{path}` has class weights or label smoothing: left unpatched (no vocab-parallel path)", UserWarning)
                continue
            stock = sub.forward

            def forward(self, input, target, _stock=stock):
                if not isinstance(input, DTensor) or input.ndim != 2 or not any(isinstance(p, Shard) and p.dim % input.ndim == 1 for p in input.placements):
                    return _stock(input, target)
                t = target._local_tensor if isinstance(target, DTensor) else target
                if self.reduction == "mean":
                    return VocabParallelCrossEntropy.mean(input, t, self.ignore_index)
                valid = t != self.ignore_index
                per_tok = VocabParallelCrossEntropy.apply(input, t.clamp(min=0)) * valid.to(input.dtype)
                return per_tok.sum() if self.reduction == "sum" else per_tok

            sub.forward = types.MethodType(forward, sub)
            sub._vb_vp_patched = True
