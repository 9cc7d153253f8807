"""Times the fused IUV-loss pass (csrc/losses.cu) at the training configuration's per-GPU batch (BASELINE configs[4]:
16 images) with CUDA events and rates it against the HBM roofline: algorithmic bytes = every prediction and target
read once + every gradient written once.  Also times the same losses + backward through torch ops (the reference's
expressions, iuv_estimator.py:320-339) on the same device for scale.  Dev tool; `one` = a single call each (for ncu)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from danet_b200 import losses

dev = torch.device("cuda:0")
one = "one" in sys.argv
gen = torch.Generator(device=dev).manual_seed(0)
B, S, P = 16, 56, 24
HW = S * S
oh = lambda n, shape: F.one_hot(torch.randint(0, n, shape, generator=gen, device=dev), n).movedim(-1, -3).float()
rnd = lambda *s: torch.randn(*s, generator=gen, device=dev)
peak = 6581.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(f, n):
    for _ in range(1 if one else 3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
# global heads: u, v, index [B,25,S,S] + ann [B,15,S,S]
I, A = oh(25, (B, S, S)), oh(15, (B, S, S))
U, V = torch.rand(B, 25, S, S, device=dev) * I, torch.rand(B, 25, S, S, device=dev) * I
preds = [rnd(B, 25, S, S).requires_grad_(), rnd(B, 25, S, S).requires_grad_(), rnd(B, 25, S, S).requires_grad_(),
         rnd(B, 15, S, S).requires_grad_()]
n = 1 if one else 50
ms = timed(lambda: losses.body_uv_losses(*preds, [U, V, I, A]), n)
by = B * HW * 4 * (9 * 25 + 3 * 15)                         # 6 reads + 3 writes of 25 channels, 2 reads + 1 write of 15
res["global"] = {"ms": ms, "bytes": by, "GB/s": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peak}
# part crops: [B,24,3,7,S,S]
pI = oh(7, (B, P, S, S))
gt = torch.stack([torch.rand(B, P, 7, S, S, device=dev) * pI, torch.rand(B, P, 7, S, S, device=dev) * pI, pI], dim=2)
pp = rnd(B, P, 3, 7, S, S).requires_grad_()
ms = timed(lambda: losses.part_iuv_losses(pp, gt), n)
by = B * P * HW * 4 * 9 * 7
res["parts"] = {"ms": ms, "bytes": by, "GB/s": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peak,
                "note": "includes the wrapper's fp32-contiguous pass-through and the gradient buffer allocation"}


def torch_parts():
    q = pp.detach().requires_grad_()
    tot = 0
    for k in range(P):
        u, v, i = q[:, k, 0], q[:, k, 1], q[:, k, 2]
        m = gt[:, k, 2] > 0
        tot = tot + F.smooth_l1_loss(u[m], gt[:, k, 0][m], reduction="sum") / B * 0.5 \
            + F.smooth_l1_loss(v[m], gt[:, k, 1][m], reduction="sum") / B * 0.5 \
            + F.cross_entropy(i.permute(0, 2, 3, 1).reshape(-1, 7), gt[:, k, 2].argmax(1).reshape(-1))
    (tot / P).backward()


res["parts_torch_ops_fwd_bwd"] = {"ms": timed(torch_parts, 1 if one else 5)}
res["hbm_peak_gbs"] = peak
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
if not one:
    json.dump(res, open("gpurun_out/losses_bench.json", "w"), indent=1)
