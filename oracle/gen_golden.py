"""Generates tests/golden/*.npz by running the REFERENCE'S OWN Python code (imported from
/root/reference with the shims in oracle/ref_import.py).  Run in this container only:

    python -m oracle.gen_golden [geometry] [iuvmap] [part_utils] [losses] [net]

The vectors pin the oracle restatements (oracle/lbs.py geometry helpers) and the CUDA kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def gen_geometry(ns):
    torch = ns.torch
    g = torch.Generator().manual_seed(1234)
    x6 = torch.randn(64, 6, generator=g)
    x6[0] = torch.tensor([1., 0., 0., 1., 0., 0.])          # identity pattern
    aa = torch.randn(64, 3, generator=g) * 0.7
    aa[0] = 0.0
    aa[1] = torch.tensor([3.1, 0.0, 0.0])
    pts = torch.randn(3, 17, 3, generator=g)
    pts[..., 2] += 5.0
    rot = ns.geometry.batch_rodrigues(torch.randn(3, 3, generator=g) * 0.2)
    tr = torch.randn(3, 3, generator=g) * 0.1
    ctr = torch.tensor([[112., 112.]] * 3)
    out = {
        "x6": x6.numpy(), "rot6d": ns.geometry.rot6d_to_rotmat(x6).numpy(),
        "aa": aa.numpy(), "rodrigues_quat": ns.geometry.batch_rodrigues(aa).numpy(),
        "pts": pts.numpy(), "rot": rot.numpy(), "tr": tr.numpy(), "ctr": ctr.numpy(),
        "persp": ns.geometry.perspective_projection(pts, rot, tr, 5000., ctr).numpy(),
    }
    np.savez_compressed(os.path.join(GOLD, "geometry.npz"), **out)
    print("geometry.npz written")


def gen_iuvmap(ns):
    torch = ns.torch
    g = torch.Generator().manual_seed(4321)
    U, V, I = (torch.randn(2, 25, 12, 12, generator=g) for _ in range(3))
    A = torch.randn(2, 15, 12, 12, generator=g)
    I[0, :, 0, 0] = 0.5                                   # exact tie -> first index
    I[0, 3, 0, 1] = I[0, 7, 0, 1] = 9.0                   # tie between 3 and 7
    cu, cv, ci, ca = ns.iuvmap.iuvmap_clean(U, V, I, A)
    pu, pv, pi_, _ = ns.iuvmap.iuvmap_clean(U[:, :7], V[:, :7], I[:, :7])
    part = torch.randint(0, 25, (2, 12, 12), generator=g).float()
    img = torch.stack([part / 24.0, torch.rand(2, 12, 12, generator=g), torch.rand(2, 12, 12, generator=g)], 1)
    img[:, 1:] *= (part > 0).float().unsqueeze(1)
    torch.Tensor.get_device = lambda self: -1 if not self.is_cuda else self.device.index
    mu, mv, mi, ma = ns.iuvmap.iuv_img2map(img)
    np.savez_compressed(os.path.join(GOLD, "iuvmap.npz"), U=U.numpy(), V=V.numpy(), I=I.numpy(), A=A.numpy(),
                        cu=cu.numpy(), cv=cv.numpy(), ci=ci.numpy(), ca=ca.numpy(),
                        pu=pu.numpy(), pv=pv.numpy(), pi=pi_.numpy(),
                        img=img.numpy(), mu=mu.numpy(), mv=mv.numpy(), mi=mi.numpy(), ma=ma.numpy())
    print("iuvmap.npz written")


def gen_part_utils(ns):
    """utils/part_utils.py:27-36 (PartRenderer.get_parts) and utils/iuvmap.py:41-70 (iuv_map2img, no-roi branch)."""
    torch = ns.torch
    sys.modules['neural_renderer'].Renderer = None
    import importlib
    part_utils = importlib.import_module("utils.part_utils")
    g = torch.Generator().manual_seed(99)
    pr = object.__new__(part_utils.PartRenderer)                 # the constructor needs a GPU and data files
    pr.cube_parts = torch.randint(0, 7, (100, 100, 100), generator=g).float()
    parts = (torch.randint(0, 100, (2, 3, 16, 16), generator=g).float() + 0.5) / 100.0
    mask = (torch.rand(2, 16, 16, generator=g) > 0.4).float()
    out = pr.get_parts(parts.clone(), mask.clone())
    U, V, I = (torch.randn(2, 25, 12, 12, generator=g) for _ in range(3))
    A = torch.randn(2, 15, 12, 12, generator=g)
    torch.Tensor.get_device = lambda self: -1 if not self.is_cuda else self.device.index
    img = ns.iuvmap.iuv_map2img(U, V, I)
    img_a = ns.iuvmap.iuv_map2img(U, V, I, A)
    np.savez_compressed(os.path.join(GOLD, "part_utils.npz"), cube=pr.cube_parts.numpy(), parts=parts.numpy(),
                        mask=mask.numpy(), out=out.numpy(), U=U.numpy(), V=V.numpy(), I=I.numpy(), A=A.numpy(),
                        img=img.numpy(), img_a=img_a.numpy())
    print("part_utils.npz written")


def gen_losses(ns):
    """models/danet/iuv_estimator.py:304-341 (body_uv_losses) under torch autograd, and the 24-part loop of
    iuv_estimator.py:232-255 (here 4 parts of 7 channels): losses and gradients w.r.t. the predictions."""
    torch = ns.torch
    F = torch.nn.functional
    g = torch.Generator().manual_seed(2718)
    B, C, CA, S, P, CP = 3, 25, 15, 6, 4, 7
    fn = ns.IUV_Estimator.body_uv_losses                         # `self` is unused apart from the module-level cfg

    def onehot(n, shape):
        return F.one_hot(torch.randint(0, n, shape, generator=g), n).movedim(-1, -3).float()

    out = {}
    u, v, i = (torch.randn(B, C, S, S, generator=g).mul_(1.5).requires_grad_() for _ in range(3))
    a = torch.randn(B, CA, S, S, generator=g).requires_grad_()
    I = onehot(C, (B, S, S))
    U, V = torch.rand(B, C, S, S, generator=g) * I, torch.rand(B, C, S, S, generator=g) * I
    A = onehot(CA, (B, S, S))
    wts = torch.tensor([1.0, 2.0, 3.0, 4.0])
    for tag, has in (("all", None), ("some", torch.tensor([True, False, True]))):
        for t in (u, v, i, a):
            t.grad = None
        L = fn(None, u, v, i, a, [U, V, I, A], has)
        sum(w * l for w, l in zip(wts, L)).backward()
        out["L_" + tag] = torch.stack([l.detach() for l in L]).numpy()
        for k, t in (("u", u), ("v", v), ("i", i), ("a", a)):       # gradient of sum_k (k+1) loss_k
            out["g%s_%s" % (k, tag)] = t.grad.numpy().copy()
    L = fn(None, u, v, i, a, [U, V, I, A], torch.tensor([False, False, False]))
    out["L_none"] = torch.stack([l.reshape(()) for l in L]).numpy()
    out.update(u=u.detach().numpy(), v=v.detach().numpy(), i=i.detach().numpy(), a=a.detach().numpy(),
               U=U.numpy(), V=V.numpy(), I=I.numpy(), A=A.numpy(), has_some=np.array([1, 0, 1], np.uint8),
               grad_weights=wts.numpy())
    # the per-part loop (iuv_estimator.py:232-255), no annotation head
    pp = torch.randn(B, P, 3, CP, S, S, generator=g).requires_grad_()
    pI = onehot(CP, (B, P, S, S))
    pg = torch.stack([torch.rand(B, P, CP, S, S, generator=g) * pI, torch.rand(B, P, CP, S, S, generator=g) * pI, pI], dim=2)
    has = torch.tensor([True, True, False])
    tot = None
    for k in range(P):
        Lk = fn(None, pp[:, k, 0], pp[:, k, 1], pp[:, k, 2], None, [pg[:, k, 0], pg[:, k, 1], pg[:, k, 2], None], has)[:3]
        tot = list(Lk) if tot is None else [x + y for x, y in zip(tot, Lk)]
    tot = [x / float(P) for x in tot]
    sum(w * l for w, l in zip(wts[:3], tot)).backward()
    out.update(part_pred=pp.detach().numpy(), part_gt=pg.numpy(), part_has=has.numpy().astype(np.uint8),
               part_L=torch.stack([l.detach() for l in tot]).numpy(), part_grad=pp.grad.numpy())
    np.savez_compressed(os.path.join(GOLD, "losses.npz"), **out)
    print("losses.npz written")


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_import
    what = sys.argv[1:] or ["geometry", "iuvmap", "part_utils", "losses", "net"]
    ns = ref_import.load(48)
    os.makedirs(GOLD, exist_ok=True)
    if "geometry" in what:
        gen_geometry(ns)
    if "iuvmap" in what:
        gen_iuvmap(ns)
    if "part_utils" in what:
        gen_part_utils(ns)
    if "losses" in what:
        gen_losses(ns)
    if "net" in what:
        from oracle import gen_golden_net
        gen_golden_net.main(ns)


if __name__ == "__main__":
    main()
