"""Shared helpers of the network parity tests."""
import os

import numpy as np
import torch
import torch.nn.functional as F

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_image(B, seed):
    """Same generator as oracle/gen_golden_net.py:make_image."""
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(B, 3, 7, 7, generator=g)
    return F.interpolate(low, size=224, mode="bilinear", align_corners=False) * 2 + 0.3 * torch.randn(B, 3, 224, 224, generator=g)


def build(width, device="cpu", conv_algo="simt", **kw):
    import danet_b200
    from danet_b200 import synthetic
    net = danet_b200.DaNet(None, synthetic.make_mean_params(0), pretrained=False, width=width,
                           smpl_model=synthetic.make_smpl_model(0), dp_mesh=synthetic.make_dp_mesh(0),
                           conv_algo=conv_algo, **kw)
    net.load_state_dict(synthetic.keyed_state_dict(net.state_dict(), 0), strict=True)
    return net.to(device).eval()


def infer_with_ops(net, image, ops):
    """infer_net with the kernel layer replaced (host-logic tests on CPU): drives a Plan directly."""
    B = image.shape[0]
    plan = net.plan_for(B, image.device, ops=ops)
    plan.run(image)
    return net.outputs_of(plan, B)


def golden_path(width, B=2):
    return os.path.join(GOLD, "net_w%d.npz" % width if B == 2 else "net_w%d_b%d.npz" % (width, B))


def check_against_golden(out, width, para_tol, kps_tol, margin_eps, min_agree=1.0, B=2, near_tie_ok=False):
    """Compares infer_net output with the reference-generated golden.  Integer decisions (argmax
    maps) must agree wherever the reference's own top-2 margin exceeds `margin_eps`."""
    g = np.load(golden_path(width, B))
    para = out["para"].cpu().numpy()
    kps = out["stn_kps_pred"].cpu().numpy()
    u, v, i, a = [t.cpu().numpy() for t in out["visualization"]["iuv_pred"]]
    nd = int(g["detail"]) if "detail" in g.files else para.shape[0]     # images with per-part / sum detail
    parts = out["visualization"]["part_iuv_pred"][:nd].cpu().numpy()     # [nd,24,3,7,S,S] raw predictions
    res = {}
    res["para_err"] = float(np.abs(para - g["para"]).max())
    res["kps_err"] = float(np.abs(kps - g["stn_kps"]).max())
    idx = i.argmax(1)
    if near_tie_ok:
        # an image whose Index / part argmax differs from the reference's at a pixel where the REFERENCE's own top-2
        # margin is below margin_eps took the other side of a near-tie: its `para` is not comparable at 1e-4
        # (see check_batch_against_golden); such flips at larger margins still fail below
        pidx_ = parts[:, :, 2].argmax(2)
        dirty = (idx != g["index_argmax"]).reshape(idx.shape[0], -1).any(1)[:nd] | (pidx_ != g["part_argmax"]).reshape(nd, -1).any(1)
        res["images_with_near_tie_flips"] = int(dirty.sum())
        clean = ~dirty
        assert clean.any(), "every image took a near-tie flip: nothing to compare"
        res["para_err"] = float(np.abs(para[:nd][clean] - g["para"][:nd][clean]).max())
    safe = g["index_margin"].astype(np.float32) > margin_eps
    res["index_agree_safe"] = float((idx == g["index_argmax"])[safe].mean())
    res["index_agree_all"] = float((idx == g["index_argmax"]).mean())
    ann = a.argmax(1)
    safe_a = g["ann_margin"].astype(np.float32) > margin_eps
    res["ann_agree_safe"] = float((ann == g["ann_argmax"])[safe_a].mean())
    pidx = parts[:, :, 2].argmax(2)
    safe_p = g["part_margin"].astype(np.float32) > margin_eps
    res["part_agree_safe"] = float((pidx == g["part_argmax"])[safe_p].mean())
    res["part_agree_all"] = float((pidx == g["part_argmax"]).mean())
    if "parts_sub" in g.files:
        res["parts_sub_err"] = float(np.abs(parts[:, ::5, :, :, ::8, ::8] - g["parts_sub"]).max())
    # one-hot structure of the cleaned maps
    assert ((i == 0) | (i == 1)).all() and (i.sum(1) == 1).all()
    assert ((a == 0) | (a == 1)).all() and (a.sum(1) == 1).all()
    same = (idx == g["index_argmax"])[:nd]
    res["u_err"] = float(np.abs(u[:nd].sum(1) - g["u_sum"])[same].max())
    res["v_err"] = float(np.abs(v[:nd].sum(1) - g["v_sum"])[same].max())
    print("golden check w%d:" % width, res)
    assert res["para_err"] < para_tol, res
    assert res["kps_err"] < kps_tol, res
    assert res["index_agree_safe"] >= min_agree and res["ann_agree_safe"] >= min_agree and res["part_agree_safe"] >= min_agree, res
    assert res["u_err"] < para_tol * 10 and res["v_err"] < para_tol * 10, res
    return res


def check_batch_against_golden(out, width, B, para_tol=1e-4, kps_tol=1e-4):
    """The benched configuration (B images) against the reference's outputs for the same images.
    The network makes integer decisions (iuvmap_clean argmax, utils/iuvmap.py:8) that feed the regressors: where the
    REFERENCE's own top-2 margin is below 1e-3 the decision is a near-tie, any implementation (the reference under
    another BLAS included) may flip it, and that image's `para` then moves by ~1e-3.  So:
      * every image whose three argmax maps equal the reference's  -> para within para_tol (1e-4, north_star);
      * every flipped pixel must be a reference near-tie (margin < 1e-3);
      * STN centres (computed before any per-part decision) within kps_tol for every image.
    Returns the summary (also used by bench.py's parity block)."""
    g = np.load(golden_path(width, B))
    para = out["para"].cpu().numpy()
    kps = out["stn_kps_pred"].cpu().numpy()
    u, v, i, a = [t.cpu().numpy() for t in out["visualization"]["iuv_pred"]]
    idx, ann = i.argmax(1), a.argmax(1)
    parts = out["visualization"]["part_iuv_pred"][:, :, 2].argmax(2).cpu().numpy()          # [B,24,S,S]
    tie = np.unpackbits(g["part_tie_bits"], axis=1)[:, :parts[0].size].reshape(parts.shape).astype(bool)
    flip_i = idx != g["index_argmax"]
    flip_a = ann != g["ann_argmax"]
    flip_p = parts != g["part_argmax_all"]
    # flips are only legitimate at near-ties of the reference
    assert not (flip_i & (g["index_margin"].astype(np.float32) > 1e-3)).any()
    assert not (flip_a & (g["ann_margin"].astype(np.float32) > 1e-3)).any()
    assert not (flip_p & ~tie).any()
    # an Index near-tie flip changes the body maps; an Ann flip only the visualisation
    dirty = flip_i.reshape(B, -1).any(1) | flip_p.reshape(B, -1).any(1)
    err = np.abs(para - g["para"]).max(1)
    res = {"images": B, "images_with_identical_integer_maps": int((~dirty).sum()),
           "para_max_abs_err_identical_maps": float(err[~dirty].max()) if (~dirty).any() else None,
           "images_with_near_tie_flips": int(dirty.sum()), "flipped_pixels": int(flip_i.sum() + flip_p.sum()),
           "para_max_abs_err_flipped_images": float(err[dirty].max()) if dirty.any() else 0.0,
           "stn_kps_max_abs_err": float(np.abs(kps - g["stn_kps"]).max())}
    print("batch golden check w%d b%d:" % (width, B), res)
    assert res["images_with_identical_integer_maps"] >= 0.75 * B, res
    assert res["para_max_abs_err_identical_maps"] < para_tol, res
    assert res["stn_kps_max_abs_err"] < kps_tol, res
    return res
