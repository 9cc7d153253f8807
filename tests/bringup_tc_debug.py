"""Bring-up harness for the tcgen05 conv kernel (TEST INFRASTRUCTURE: checks against oracle/net_ops.py): small isolating cases first, dumps for offline analysis.  Run as `python tests/bringup_tc_debug.py` on a GPU box."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from danet_b200.plan import CudaOps
from oracle.net_ops import TorchEmulOps

DEV = "cuda:0"
ops, ref = CudaOps(DEV), TorchEmulOps()
os.makedirs("gpurun_out", exist_ok=True)
cases = [
    # name, N,H,W,Cin,Cout,k,wsets,relu,res   (stride via name suffix _s2)
    ("1x1_c8_n16_onetile", 1, 16, 8, 8, 16, 1, 1, 0, 0),
    ("1x1_c32_n16", 1, 16, 8, 32, 16, 1, 1, 0, 0),
    ("1x1_c64_n48", 2, 16, 16, 64, 48, 1, 1, 0, 0),
    ("3x3_c8_n16_onetile", 1, 16, 8, 8, 16, 3, 1, 0, 0),
    ("3x3_c48_n48_56", 2, 56, 56, 48, 48, 3, 1, 1, 1),
    ("3x3_c96_n96_28", 2, 28, 28, 96, 96, 3, 1, 1, 1),
    ("3x3_c192_14", 3, 14, 14, 192, 192, 3, 1, 0, 0),
    ("3x3_c384_7", 2, 7, 7, 384, 384, 3, 1, 1, 1),
    ("1x1_c256_n64", 2, 56, 56, 256, 64, 1, 1, 1, 0),
    ("3x3_c48_n24_g24", 48, 56, 56, 48, 24, 3, 24, 0, 0),
    ("3x3_c48_n92", 2, 56, 56, 48, 92, 3, 1, 0, 0),
    ("3x3_c512", 2, 7, 7, 512, 512, 3, 1, 1, 1),
    ("3x3_c8_n16_s2", 1, 32, 16, 8, 16, 3, 1, 0, 0),
    ("3x3_c48_n96_s2", 2, 56, 56, 48, 96, 3, 1, 1, 0),
    ("7x7_c64_n64_s2", 4, 56, 56, 64, 64, 7, 1, 1, 0),
    ("1x1_c64_n128_s2", 2, 28, 28, 64, 128, 1, 1, 0, 0),
    ("3x3_c64_n128_s2_14", 3, 14, 14, 64, 128, 3, 1, 1, 0),
    ("3x3_c13x9", 5, 13, 9, 24, 36, 3, 1, 1, 1),
    # full-batch timings of the dominant shapes (B=64)
    ("T_3x3_c48_56_b64", 64, 56, 56, 48, 48, 3, 1, 1, 1),
    ("T_3x3_c96_28_b64", 64, 28, 28, 96, 96, 3, 1, 1, 1),
    ("T_3x3_c192_14_b64", 64, 14, 14, 192, 192, 3, 1, 1, 1),
    ("T_3x3_c384_7_b64", 64, 7, 7, 384, 384, 3, 1, 1, 1),
    ("T_7x7_c64_s2_b1536", 1536, 56, 56, 64, 64, 7, 1, 1, 0),
    ("T_3x3_c64_14_b1536", 1536, 14, 14, 64, 64, 3, 1, 1, 1),
    ("T_1x1_c24_56_b1536", 1536, 56, 56, 24, 64, 1, 1, 1, 0),
]
results = {}
for (name, N, H, W, Cin, Cout, k, G, relu, has_res) in cases:
    st = 2 if "_s2" in name else 1
    d = dict(N=N, H=H, W=W, Cin=Cin, Cout=Cout, ksize=k, stride=st, pad=k // 2, wsets=G, relu=relu)
    Ho, Wo = (H + 2 * (k // 2) - k) // st + 1, (W + 2 * (k // 2) - k) // st + 1
    big = name.startswith("T_")
    if not ops.conv_tc_supported(d):
        results[name] = "unsupported"; print(name, "unsupported", flush=True); continue
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, H, W, Cin, generator=g)
    w = torch.randn(G, k * k * Cin, Cout, generator=g) * (1.0 / (k * k * Cin)) ** 0.5
    b = torch.randn(G, Cout, generator=g) * 0.1
    res = torch.randn(N, Ho, Wo, Cout, generator=g) if has_res else None
    y = torch.full((N, Ho, Wo, Cout), float("nan"), device=DEV)
    if big:      # reference = the (already validated) fp32 FMA kernel on the GPU
        y_ref_d = torch.empty(N, Ho, Wo, Cout, device=DEV)
        ops.conv2d(d, 0, x.to(DEV), w.to(DEV), b.to(DEV), res.to(DEV) if has_res else None, y_ref_d)
        y_ref = y_ref_d.cpu()
    else:
        y_ref = torch.empty(N, Ho, Wo, Cout); ref.conv2d(d, 0, x, w, b, res, y_ref)
    wp = ops.conv_tc_pack(d, w.to(DEV))
    try:
        ops.conv2d(d, 1, x.to(DEV), wp, b.to(DEV), res.to(DEV) if has_res else None, y)
        torch.cuda.synchronize()
    except Exception as e:
        results[name] = "EXC " + str(e)[:200]; print(name, results[name], flush=True); break
    yc = y.cpu()
    err = (yc - y_ref).abs()
    nan = torch.isnan(yc).float().mean().item()
    results[name] = dict(max_err=float(err[~torch.isnan(err)].max()) if nan < 1 else None, nan_frac=nan,
                         mean_err=float(err[~torch.isnan(err)].mean()) if nan < 1 else None)
    print(name, results[name], flush=True)
    if "onetile" in name:
        np.savez("gpurun_out/tc_%s.npz" % name, y=yc.numpy(), y_ref=y_ref.numpy(), x=x.numpy(), w=w.numpy(), b=b.numpy())
    # timing for the big ones
    if big and results[name]["max_err"] is not None and results[name]["max_err"] < 1e-2:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        xc, bc = x.to(DEV), b.to(DEV)
        e0.record()
        for _ in range(20):
            ops.conv2d(d, 1, xc, wp, bc, None, y)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        fl = 2.0 * N * Ho * Wo * Cout * k * k * Cin
        results[name]["us"] = us; results[name]["tflops"] = fl / us / 1e6
        e0.record()
        wd = w.to(DEV)
        for _ in range(5):
            ops.conv2d(d, 0, xc, wd, bc, None, y)
        e1.record(); torch.cuda.synchronize()
        results[name]["simt_us"] = e0.elapsed_time(e1) / 5 * 1e3
        print("   tc us/launch %.1f  (%.1f TFLOP/s)   fp32-FMA kernel us %.1f" % (us, fl / us / 1e6, results[name]["simt_us"]), flush=True)
json.dump(results, open("gpurun_out/tc_debug.json", "w"), indent=1)
