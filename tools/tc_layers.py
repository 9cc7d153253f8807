"""Per-layer timing of the tcgen05 conv on the layer shapes that dominate the W48 B=64 step
(counts from profiles/r01_final_conv_per_layer.txt).  Prints us per launch (CUDA events, 3 rotating
buffer sets so that a launch does not find its own inputs in L2) and the weighted sum.
usage: python tools/tc_layers.py [quick]"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from danet_b200.plan import CudaOps
DEV = "cuda:0"
ops = CudaOps(DEV)
# (N, H, Cin, Cout, k, stride, wsets, residual, count per step)
LAYERS = [
    (64, 56, 48, 48, 3, 1, 1, 1, 32), (64, 56, 48, 48, 3, 1, 1, 0, 32),
    (64, 28, 96, 96, 3, 1, 1, 1, 32), (64, 28, 96, 96, 3, 1, 1, 0, 32),
    (64, 14, 192, 192, 3, 1, 1, 1, 28), (64, 14, 192, 192, 3, 1, 1, 0, 28),
    (64, 7, 384, 384, 3, 1, 1, 1, 12), (64, 7, 384, 384, 3, 1, 1, 0, 12),
    (1536, 56, 64, 64, 7, 2, 1, 0, 1), (1536, 56, 48, 24, 3, 1, 24, 0, 1),
    (64, 56, 64, 256, 1, 1, 1, 1, 4), (1536, 56, 24, 64, 1, 1, 1, 0, 1),
    (1536, 4, 256, 256, 3, 1, 1, 1, 2), (64, 56, 256, 64, 1, 1, 1, 0, 3),
    (64, 56, 48, 48, 3, 2, 1, 0, 8), (64, 56, 48, 96, 3, 2, 1, 0, 7),
    (64, 28, 96, 192, 3, 2, 1, 0, 7), (64, 28, 96, 48, 1, 1, 1, 0, 8),
    (1536, 14, 64, 64, 3, 1, 1, 1, 2), (1536, 7, 128, 128, 3, 1, 1, 1, 2),
    (64, 14, 192, 96, 1, 1, 1, 0, 6), (64, 14, 192, 48, 1, 1, 1, 0, 7),
]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    LAYERS = LAYERS[:8] + LAYERS[10:11]
total = 0.0
for (N, H, Cin, Cout, k, st, ws, has_res, cnt) in LAYERS:
    d = dict(N=N, H=H, W=H, Cin=Cin, Cout=Cout, ksize=k, stride=st, pad=k // 2, wsets=ws, relu=1)
    Ho = (H + 2 * (k // 2) - k) // st + 1
    big = N * H * H * Cin * 4 > 200e6
    nset = 1 if big else 3
    xs = [torch.randn(N, H, H, Cin, device=DEV) for _ in range(nset)]
    rs = [torch.randn(N, Ho, Ho, Cout, device=DEV) if has_res else None for _ in range(nset)]
    ys = [torch.empty(N, Ho, Ho, Cout, device=DEV) for _ in range(nset)]
    w = torch.randn(ws, k * k * Cin, Cout, device=DEV) * 0.05
    b = torch.randn(ws, Cout, device=DEV) * 0.1
    wp = ops.conv_tc_pack(d, w)
    reps = 4 if big else 15
    for i in range(3):
        ops.conv2d(d, 1, xs[i % nset], wp, b, rs[i % nset], ys[i % nset])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.conv2d(d, 1, xs[i % nset], wp, b, rs[i % nset], ys[i % nset])
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    # correctness spot check against the fp32 FMA kernel on the first set
    ws_ = w.contiguous()
    yref = torch.empty_like(ys[0])
    ops.conv2d(d, 0, xs[0], ws_, b, rs[0], yref)
    ops.conv2d(d, 1, xs[0], wp, b, rs[0], ys[0])
    torch.cuda.synchronize()
    err = (ys[0] - yref).abs().max().item() / max(1e-6, yref.abs().max().item())
    fl = 2.0 * N * Ho * Ho * Cin * Cout * k * k
    total += us * cnt
    print("N%-5d H%-3d Cin%-4d Cout%-4d k%d s%d g%-2d res%d  %8.1f us  %7.1f TFLOP/s  x%-3d rel.err %.1e" %
          (N, H, Cin, Cout, k, st, ws, has_res, us, fl / us / 1e6, cnt, err), flush=True)
    del xs, rs, ys
print("weighted sum: %.2f ms per step (these layer classes)" % (total / 1e3))
