"""Dense IUV losses of the training step with the reference's call surface:

    loss_U, loss_V, loss_IndexUV, loss_segAnn = body_uv_losses(u_pred, v_pred, index_pred, ann_pred, uvia_list, has_iuv)
    loss_pU, loss_pV, loss_pIndexUV           = part_iuv_losses(part_iuv_pred, part_iuv_gt, has_iuv)

`body_uv_losses` is models/danet/iuv_estimator.py:304-341; `part_iuv_losses` is the loop over the 24 part crops of
iuv_estimator.py:232-255 as one launch.  Forward and backward are ONE fused CUDA pass (csrc/losses.cu,
danet_body_uv_losses): the gradients w.r.t. the predictions are produced with the losses and handed to autograd by a
torch.autograd.Function.  No CPU path.

Deviation from the reference, stated: when no image has IUV ground truth the reference returns `torch.zeros(1)` per
loss after a host synchronisation (`torch.sum(has_iuv) > 0`); here the losses are 0-dim zeros and nothing synchronises
(the count stays on the device)."""
import torch
from torch.autograd.function import once_differentiable

from . import _lib

POINT_REGRESSION_WEIGHTS = 0.5                  # configs/danet_default.yaml:23 (cfg.DANET.POINT_REGRESSION_WEIGHTS)


def _launch(N, C, Cann, HW, pred_stride, map_stride, u, v, idx, ann, U, V, I, A, has, batch_size, point_weight, dev,
            gu, gv, gi, ga):
    """Pointers are integer device addresses (or None); returns losses [4] on `dev`."""
    lib = _lib.load()
    p = lambda x: _lib.c_p(x if x else 0)
    with torch.cuda.device(dev):
        ws = torch.empty(int(lib.danet_body_uv_losses_workspace_bytes(N, HW)), dtype=torch.uint8, device=dev)
        losses = torch.empty(4, device=dev)
        _lib.check(lib.danet_body_uv_losses(N, C, Cann, HW, pred_stride, map_stride, p(u), p(v), p(idx), p(ann), p(U), p(V),
                                            p(I), p(A), p(has), float(batch_size), float(point_weight), _lib.ptr(losses),
                                            p(gu), p(gv), p(gi), p(ga), _lib.ptr(ws), _lib.stream_ptr(dev)),
                   "body_uv_losses")
    return losses


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _has_u8(has_iuv, dev, repeat=1):
    if has_iuv is None:
        return None
    h = (has_iuv.to(dev) != 0).to(torch.uint8)
    if repeat > 1:
        h = h.repeat_interleave(repeat)
    return h.contiguous()


class _BodyUvLosses(torch.autograd.Function):
    """losses [4] = (loss_U, loss_V, loss_IndexUV, loss_segAnn); the backward multiplies the gradients the fused pass
    already wrote by the incoming d/d losses[k]."""

    @staticmethod
    def forward(ctx, u, v, idx, ann, U, V, I, A, has, point_weight):
        dev = u.device
        B, C = u.shape[0], u.shape[1]
        HW = u.shape[2] * u.shape[3]
        u_, v_, i_ = _f32(u, dev), _f32(v, dev), _f32(idx, dev)
        U_, V_, I_ = _f32(U, dev), _f32(V, dev), _f32(I, dev)
        a_ = _f32(ann, dev) if ann is not None else None
        A_ = _f32(A, dev) if ann is not None else None
        need = [ctx.needs_input_grad[k] for k in range(4)]
        gu = torch.empty_like(u_) if need[0] else None
        gv = torch.empty_like(v_) if need[1] else None
        gi = torch.empty_like(i_) if need[2] else None
        ga = torch.empty_like(a_) if (ann is not None and need[3]) else None
        d = lambda t: t.data_ptr() if t is not None else 0
        losses = _launch(B, C, a_.shape[1] if a_ is not None else 0, HW, 0, 0, d(u_), d(v_), d(i_), d(a_), d(U_), d(V_),
                         d(I_), d(A_), d(has), float(B), point_weight, dev, d(gu), d(gv), d(gi), d(ga))
        ctx.grads = (gu, gv, gi, ga)
        ctx.dtypes = (u.dtype, v.dtype, idx.dtype, ann.dtype if ann is not None else None)
        return losses

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        out = []
        for k, t in enumerate(ctx.grads):
            out.append(None if t is None else (t * g[k]).to(ctx.dtypes[k]))
        return (*out, None, None, None, None, None, None)


def body_uv_losses(u_pred, v_pred, index_pred, ann_pred, uvia_list, has_iuv=None,
                   point_weight=POINT_REGRESSION_WEIGHTS):
    """models/danet/iuv_estimator.py:304-341.  u/v/index_pred [B,C,S,S]; ann_pred [B,Cann,S,S] or None;
    uvia_list = (Umap, Vmap, Imap, Annmap) as utils/iuvmap.py iuv_img2map returns them; has_iuv [B] or None.
    Returns (loss_U, loss_V, loss_IndexUV, loss_segAnn | None), differentiable w.r.t. the predictions."""
    _lib.require_cuda(u_pred, "u_pred")
    Umap, Vmap, Imap, Annmap = uvia_list
    if u_pred.dim() != 4 or u_pred.shape != v_pred.shape or u_pred.shape != index_pred.shape or u_pred.shape != Imap.shape:
        raise ValueError("body_uv_losses: u/v/index predictions and the target maps must share one [B,C,S,S] shape")
    if ann_pred is not None and (Annmap is None or ann_pred.shape != Annmap.shape):
        raise ValueError("body_uv_losses: ann_pred needs an Annmap of the same shape")
    if has_iuv is not None and has_iuv.shape[0] != u_pred.shape[0]:
        raise ValueError("body_uv_losses: has_iuv must have one entry per image")
    if u_pred.shape[0] == 0 or u_pred.shape[2] * u_pred.shape[3] == 0:      # nothing to sum: zeros that still carry a graph
        z = u_pred.sum() * 0 + v_pred.sum() * 0 + index_pred.sum() * 0
        return z, z, z, (ann_pred.sum() * 0 if ann_pred is not None else None)
    has = _has_u8(has_iuv, u_pred.device)
    L = _BodyUvLosses.apply(u_pred, v_pred, index_pred, ann_pred, Umap, Vmap, Imap, Annmap if ann_pred is not None else None,
                            has, point_weight)
    return L[0], L[1], L[2], (L[3] if ann_pred is not None else None)


class _PartIuvLosses(torch.autograd.Function):
    """The 24 body_uv_losses calls of iuv_estimator.py:232-255 (+ the /24 means) over part_iuv_pred [B,24,3,7,S,S]
    in place: image = (batch, part) row, u / v / index = the three 7-channel groups of a row."""

    @staticmethod
    def forward(ctx, pred, gt, has, point_weight):
        dev = pred.device
        B, P, three, C = pred.shape[:4]
        HW = pred.shape[4] * pred.shape[5]
        p_, g_ = _f32(pred, dev), _f32(gt, dev)
        grad = torch.empty_like(p_) if ctx.needs_input_grad[0] else None
        step = C * HW * 4                                         # bytes between the u, v and index groups of a row
        pb, gb, qb = p_.data_ptr(), g_.data_ptr(), (grad.data_ptr() if grad is not None else 0)
        q = lambda k: qb + k * step if qb else 0
        losses = _launch(B * P, C, 0, HW, three * C * HW, three * C * HW, pb, pb + step, pb + 2 * step, 0,
                         gb, gb + step, gb + 2 * step, 0, has.data_ptr() if has is not None else 0, float(B * P),
                         point_weight, dev, q(0), q(1), q(2), 0)
        ctx.grad = grad
        ctx.dtype = pred.dtype
        return losses[:3]

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        if ctx.grad is None:
            return None, None, None, None
        return (ctx.grad * g.reshape(1, 1, 3, 1, 1, 1)).to(ctx.dtype), None, None, None


def part_iuv_losses(part_iuv_pred, part_iuv_gt, has_iuv=None, point_weight=POINT_REGRESSION_WEIGHTS):
    """iuv_estimator.py:232-255: part_iuv_pred / part_iuv_gt [B,24,3,7,S,S] (u, v, index groups of the part crops).
    Returns (loss_pU, loss_pV, loss_pIndexUV) = the means over the 24 parts of body_uv_losses per part."""
    _lib.require_cuda(part_iuv_pred, "part_iuv_pred")
    if part_iuv_pred.dim() != 6 or part_iuv_pred.shape[2] != 3 or part_iuv_pred.shape != part_iuv_gt.shape:
        raise ValueError("part_iuv_losses: expected part_iuv_pred and part_iuv_gt of one shape [B,P,3,C,S,S]")
    if has_iuv is not None and has_iuv.shape[0] != part_iuv_pred.shape[0]:
        raise ValueError("part_iuv_losses: has_iuv must have one entry per image")
    if part_iuv_pred.numel() == 0:
        z = part_iuv_pred.sum() * 0
        return z, z, z
    has = _has_u8(has_iuv, part_iuv_pred.device, repeat=part_iuv_pred.shape[1])
    L = _PartIuvLosses.apply(part_iuv_pred, part_iuv_gt, has, point_weight)
    return L[0], L[1], L[2]
