"""CPU restatement (numpy, fp64) of the reference's dense IUV losses and their gradients -- TEST INFRASTRUCTURE ONLY
(imported by tests/ alone; the product path never touches it).

Follows models/danet/iuv_estimator.py:304-341 (IUV_Estimator.body_uv_losses) and the 24-part loop of
iuv_estimator.py:232-255.  Pinned by tests/golden/losses.npz, which oracle/gen_golden.py produces by calling the
reference's own function under torch autograd (tests/test_oracle_cpu.py)."""
import numpy as np


def _ce(logits, target_map, sel):
    """Mean cross-entropy over the pixels of the selected images (iuv_estimator.py:320-327) and d/dlogits."""
    x = logits[sel].astype(np.float64)                                    # [n,C,H,W]
    t = np.argmax(target_map[sel], axis=1)                                # first maximum, like torch.argmax
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m)
    s = e.sum(axis=1, keepdims=True)
    lse = (m + np.log(s))[:, 0]
    xt = np.take_along_axis(x, t[:, None], axis=1)[:, 0]
    npix = t.size
    g = np.zeros(logits.shape, np.float64)
    gs = e / s
    np.put_along_axis(gs, t[:, None], np.take_along_axis(gs, t[:, None], axis=1) - 1.0, axis=1)
    g[sel] = gs / npix
    return (lse - xt).sum() / npix, g


def _sl1(pred, tgt, mask, sel, scale):
    """scale * sum smooth_l1(pred - tgt) over mask (beta = 1; iuv_estimator.py:325-326,329-330) and d/dpred."""
    d = (pred.astype(np.float64) - tgt.astype(np.float64))
    on = mask & sel.reshape((-1,) + (1,) * (pred.ndim - 1))
    a = np.abs(d)
    l = np.where(a < 1.0, 0.5 * d * d, a - 0.5)
    g = np.where(on, np.clip(d, -1.0, 1.0), 0.0) * scale
    return (l * on).sum() * scale, g


def body_uv_losses(u_pred, v_pred, index_pred, ann_pred, uvia_list, has_iuv=None, point_weight=0.5):
    """Returns (losses [4] float64 -- loss_segAnn = 0 when ann_pred is None, grads dict of d loss_k / d prediction)."""
    Umap, Vmap, Imap, Annmap = uvia_list
    B = u_pred.shape[0]
    sel = np.ones(B, bool) if has_iuv is None else np.asarray(has_iuv).astype(bool)
    z = {"u": np.zeros(u_pred.shape), "v": np.zeros(v_pred.shape), "index": np.zeros(index_pred.shape),
         "ann": None if ann_pred is None else np.zeros(ann_pred.shape)}
    if not sel.any():                                                     # iuv_estimator.py:311-318
        return np.zeros(4), z
    mask = Imap > 0
    lu, gu = _sl1(u_pred, Umap, mask, sel, point_weight / B)
    lv, gv = _sl1(v_pred, Vmap, mask, sel, point_weight / B)
    li, gi = _ce(index_pred, Imap, sel)
    la, ga = (0.0, None) if ann_pred is None else _ce(ann_pred, Annmap, sel)
    return np.array([lu, lv, li, la]), {"u": gu, "v": gv, "index": gi, "ann": ga}


def part_iuv_losses(part_iuv_pred, part_iuv_gt, has_iuv=None, point_weight=0.5):
    """iuv_estimator.py:232-255: mean over the parts of body_uv_losses(pred[:, i, 0..2], None, gt[:, i, 0..2]).
    Returns (losses [3], gradient w.r.t. part_iuv_pred)."""
    P = part_iuv_pred.shape[1]
    tot = np.zeros(3)
    g = np.zeros(part_iuv_pred.shape)
    for i in range(P):
        l, gr = body_uv_losses(part_iuv_pred[:, i, 0], part_iuv_pred[:, i, 1], part_iuv_pred[:, i, 2], None,
                               [part_iuv_gt[:, i, 0], part_iuv_gt[:, i, 1], part_iuv_gt[:, i, 2], None], has_iuv, point_weight)
        tot += l[:3] / P
        g[:, i, 0], g[:, i, 1], g[:, i, 2] = gr["u"] / P, gr["v"] / P, gr["index"] / P
    return tot, g
