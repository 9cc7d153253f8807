"""GPU tests of the fused IUV-loss kernel (csrc/losses.cu through the C ABI) -- SURVEY section 8f-2.
Oracles: tests/golden/losses.npz (the reference's own body_uv_losses under torch autograd, oracle/gen_golden.py) and,
at the sizes of the training configuration, the same torch expressions (iuv_estimator.py:320-339) evaluated by torch on
the device.  Tolerances: golden (6 x 6 maps) losses 5e-6 relative, gradients 5e-7 absolute; training size (torch's own
fp32 tree sums on the other side) losses 2e-5 relative, gradients 2e-4 relative."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "losses.npz"))


def _dev(x, grad=False):
    t = torch.tensor(np.asarray(x), dtype=torch.float32, device="cuda:0")
    return t.requires_grad_() if grad else t


def test_body_uv_losses_match_reference_golden(gold):
    from danet_b200 import losses
    g = gold
    w = _dev(g["grad_weights"])
    maps = [_dev(g[k]) for k in ("U", "V", "I", "A")]
    for tag, has in (("all", None), ("some", torch.tensor(g["has_some"]).bool().cuda()), ("none", torch.zeros(3).cuda())):
        u, v, i, a = (_dev(g[k], True) for k in ("u", "v", "i", "a"))
        L = losses.body_uv_losses(u, v, i, a, maps, has)
        assert all(l.dim() == 0 for l in L)
        np.testing.assert_allclose(torch.stack(L).detach().cpu().numpy(), g["L_" + tag], rtol=5e-6)
        sum(wk * l for wk, l in zip(w, L)).backward()
        for t, name in ((u, "gu"), (v, "gv"), (i, "gi"), (a, "ga")):
            ref = g["%s_%s" % (name, tag)] if tag != "none" else 0.0
            np.testing.assert_allclose(t.grad.cpu().numpy(), ref, atol=5e-7)
    # no annotation head: three losses + None, gradients only where asked for
    u, i = _dev(g["u"], True), _dev(g["i"], True)
    L = losses.body_uv_losses(u, _dev(g["v"]), i, None, maps[:3] + [None])
    assert L[3] is None
    np.testing.assert_allclose(torch.stack(L[:3]).detach().cpu().numpy(), g["L_all"][:3], rtol=5e-6)
    (L[0] + 3 * L[2]).backward()
    np.testing.assert_allclose(u.grad.cpu().numpy(), g["gu_all"], atol=5e-7)
    np.testing.assert_allclose(i.grad.cpu().numpy(), g["gi_all"], atol=5e-7)


def test_part_iuv_losses_match_reference_golden(gold):
    from danet_b200 import losses
    g = gold
    pp = _dev(g["part_pred"], True)
    L = losses.part_iuv_losses(pp, _dev(g["part_gt"]), torch.tensor(g["part_has"]).cuda())
    np.testing.assert_allclose(torch.stack(L).detach().cpu().numpy(), g["part_L"], rtol=5e-6)
    sum(wk * l for wk, l in zip(_dev(g["grad_weights"][:3]), L)).backward()
    np.testing.assert_allclose(pp.grad.cpu().numpy(), g["part_grad"], atol=5e-7)


def _torch_body_uv(u, v, idx, ann, U, V, I, A, has, pw=0.5):
    """iuv_estimator.py:304-341 restated with current torch spellings (reduction='sum' for size_average=False)."""
    B = u.shape[0]
    if has is not None:
        u, v, idx, U, V, I = u[has], v[has], idx[has], U[has], V[has], I[has]
        if ann is not None:
            ann, A = ann[has], A[has]
    m = I > 0
    lu = F.smooth_l1_loss(u[m], U[m], reduction="sum") / B * pw
    lv = F.smooth_l1_loss(v[m], V[m], reduction="sum") / B * pw
    li = F.cross_entropy(idx.permute(0, 2, 3, 1).reshape(-1, I.shape[1]), I.argmax(1).reshape(-1))
    la = None if ann is None else F.cross_entropy(ann.permute(0, 2, 3, 1).reshape(-1, A.shape[1]), A.argmax(1).reshape(-1))
    return lu, lv, li, la


def test_losses_at_training_size_match_torch():
    """BASELINE configs[4] per-GPU batch (16 images, 56 x 56 maps): global heads and the 24 part crops."""
    from danet_b200 import losses
    gen = torch.Generator(device="cuda:0").manual_seed(5)
    B, S = 16, 56
    onehot = lambda n, shape: F.one_hot(torch.randint(0, n, shape, generator=gen, device="cuda:0"), n).movedim(-1, -3).float()
    rnd = lambda *s: torch.randn(*s, generator=gen, device="cuda:0")
    I, A = onehot(25, (B, S, S)), onehot(15, (B, S, S))
    U, V = torch.rand(B, 25, S, S, generator=gen, device="cuda:0") * I, torch.rand(B, 25, S, S, generator=gen, device="cuda:0") * I
    has = torch.rand(B, generator=gen, device="cuda:0") > 0.3
    has[0] = True
    preds = [rnd(B, 25, S, S) * 2, rnd(B, 25, S, S) * 2, rnd(B, 25, S, S) * 3, rnd(B, 15, S, S) * 3]
    ours = [p.clone().requires_grad_() for p in preds]
    ref = [p.clone().requires_grad_() for p in preds]
    L = losses.body_uv_losses(*ours, [U, V, I, A], has)
    R = _torch_body_uv(*ref, U, V, I, A, has)
    for l, r in zip(L, R):
        assert abs(l.item() - r.item()) <= 2e-5 * abs(r.item())
    sum((k + 1.0) * l for k, l in enumerate(L)).backward()
    sum((k + 1.0) * r for k, r in enumerate(R)).backward()
    for o, r in zip(ours, ref):
        torch.testing.assert_close(o.grad, r.grad, rtol=2e-4, atol=1e-9)
        assert r.grad.abs().max().item() > 0
    # bit-for-bit repeatable (fixed summation order, no float atomics)
    L2 = losses.body_uv_losses(*[p.detach() for p in preds], [U, V, I, A], has)
    assert all(torch.equal(a.detach(), b) for a, b in zip(L, L2))
    # part crops: one launch over 16 x 24 rows == the reference's loop of 24 calls (iuv_estimator.py:232-255)
    P, C = 24, 7
    pI = onehot(C, (B, P, S, S))
    gt = torch.stack([torch.rand(B, P, C, S, S, generator=gen, device="cuda:0") * pI,
                      torch.rand(B, P, C, S, S, generator=gen, device="cuda:0") * pI, pI], dim=2)
    pp = rnd(B, P, 3, C, S, S)
    a, b = pp.clone().requires_grad_(), pp.clone().requires_grad_()
    L = losses.part_iuv_losses(a, gt, has)
    tot = [0.0, 0.0, 0.0]
    for k in range(P):
        r = _torch_body_uv(b[:, k, 0], b[:, k, 1], b[:, k, 2], None, gt[:, k, 0], gt[:, k, 1], gt[:, k, 2], None, has)
        tot = [t + x / P for t, x in zip(tot, r[:3])]
    for l, r in zip(L, tot):
        assert abs(l.item() - r.item()) <= 2e-5 * abs(r.item())
    (L[0] + 2 * L[1] + 3 * L[2]).backward()
    (tot[0] + 2 * tot[1] + 3 * tot[2]).backward()
    torch.testing.assert_close(a.grad, b.grad, rtol=2e-4, atol=1e-9)


def test_losses_edge_cases():
    from danet_b200 import losses
    z = lambda *s: torch.zeros(*s, device="cuda:0")
    # empty batch: zeros, nothing launched over pixels
    L = losses.body_uv_losses(z(0, 25, 8, 8), z(0, 25, 8, 8), z(0, 25, 8, 8), None, [z(0, 25, 8, 8)] * 3 + [None])
    assert [float(l) for l in L[:3]] == [0.0, 0.0, 0.0]
    # shape / argument errors mirror a wrong call of the reference
    with pytest.raises(ValueError):
        losses.body_uv_losses(z(2, 25, 8, 8), z(2, 24, 8, 8), z(2, 25, 8, 8), None, [z(2, 25, 8, 8)] * 3 + [None])
    with pytest.raises(ValueError):
        losses.body_uv_losses(z(2, 25, 8, 8), z(2, 25, 8, 8), z(2, 25, 8, 8), z(2, 15, 8, 8), [z(2, 25, 8, 8)] * 3 + [None])
    with pytest.raises(ValueError):
        losses.part_iuv_losses(z(2, 24, 2, 7, 8, 8), z(2, 24, 2, 7, 8, 8))
    # uniform logits on a one-hot target: cross-entropy = log(C), smooth-L1 branches: |d| = 2 -> 1.5, d = 0.5 -> 0.125
    I = z(1, 4, 2, 2)
    I[:, 1] = 1
    u = z(1, 4, 2, 2)
    u[:, 1] = 2.0
    v = z(1, 4, 2, 2)
    v[:, 1] = 0.5
    L = losses.body_uv_losses(u, v, z(1, 4, 2, 2), None, [z(1, 4, 2, 2), z(1, 4, 2, 2), I, None])
    assert abs(float(L[0]) - 0.5 * 4 * 1.5) < 1e-6 and abs(float(L[1]) - 0.5 * 4 * 0.125) < 1e-6
    assert abs(float(L[2]) - np.log(4.0)) < 1e-6
