"""Role-cycle accounting and knock-outs of the conv engine on the SMPL blend-shape GEMM shape
(1024 bodies x 224 features x 20672 outputs, exact mode, fp32 output): python tools/gemm_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conv_tc_common import DEV, desc, launch, pack, problem, split
from danet_b200 import _lib
from tc_roles import NAMES


def run(case, exact=True, planes_out=False, iters=10):
    N, H, W, Cin, Cout, k, s, G, relu, has_res = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, H, W, Cin, generator=g).to(DEV)
    w = (torch.randn(G, k * k * Cin, Cout, generator=g) * 0.05).to(DEV)
    b = (torch.randn(G, Cout, generator=g) * 0.1).to(DEV)
    d = desc(case, exact)
    xp = split(x, want_lo=exact)
    wpk = pack(d, w)
    y = torch.empty(N, H, W, Cout, device=DEV)
    yp = None
    if planes_out:
        yp = (torch.empty(N, H, W, Cout, dtype=torch.float16, device=DEV), torch.empty(N, H, W, Cout, dtype=torch.float16, device=DEV))
    p = problem(d, xp, wpk, b, y_f32=None if planes_out else y, y_planes=yp)
    for _ in range(3):
        launch([p])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        launch([p])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    lib = _lib.load()
    buf = torch.zeros(16, dtype=torch.int64, device=DEV)
    lib.danet_conv_tc_set_profile_buffer(buf.data_ptr())
    launch([p])
    torch.cuda.synchronize()
    lib.danet_conv_tc_set_profile_buffer(None)
    v = buf.cpu().tolist()
    return us, v


if __name__ == "__main__":
    case = (1, 128, 8, 224, 20672, 1, 1, 1, 0, 0)
    for planes in (False, True):
        us, v = run(case, True, planes)
        print("%s exact %s: %.1f us  %.1f TF" % (case, "planes out" if planes else "f32 out", us, 2.0 * 1024 * 224 * 20672 / us / 1e6))
        print("   " + "  ".join("%s=%d" % (n, x) for n, x in zip(NAMES, v)))
