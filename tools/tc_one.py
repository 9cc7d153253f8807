"""Runs a few representative tcgen05 conv launches (for `ncu --set full` captures)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from danet_b200.plan import CudaOps
DEV = "cuda:0"
ops = CudaOps(DEV)
cases = [(64, 56, 56, 48, 48, 3, 1), (64, 28, 28, 96, 96, 3, 1), (64, 14, 14, 192, 192, 3, 1), (1536, 56, 56, 64, 64, 7, 2)]
for (N, H, W, Cin, Cout, k, st) in cases:
    d = dict(N=N, H=H, W=W, Cin=Cin, Cout=Cout, ksize=k, stride=st, pad=k // 2, wsets=1, relu=1)
    Ho, Wo = (H + 2 * (k // 2) - k) // st + 1, (W + 2 * (k // 2) - k) // st + 1
    x = torch.randn(N, H, W, Cin, device=DEV)
    w = torch.randn(1, k * k * Cin, Cout, device=DEV) * 0.05
    b = torch.randn(1, Cout, device=DEV) * 0.1
    res = torch.randn(N, Ho, Wo, Cout, device=DEV) if st == 1 else None
    y = torch.empty(N, Ho, Wo, Cout, device=DEV)
    wp = ops.conv_tc_pack(d, w)
    prof = torch.zeros(16, dtype=torch.int64, device=DEV)
    ops.lib.danet_conv_tc_set_profile_buffer(ctypes.c_void_p(prof.data_ptr()))
    for _ in range(2):
        ops.conv2d(d, 1, x, wp, b, res, y)
    torch.cuda.synchronize()
    ops.lib.danet_conv_tc_set_profile_buffer(ctypes.c_void_p(0))
    pr = prof.cpu().tolist()
    names = ["mma_total", "mma_wait_acc_empty", "mma_wait_a_full", "mma_wait_b_full", "prod_total", "prod_wait_a_empty", "epi_total", "epi_wait_acc_full"]
    print((N, H, Cin, Cout, k, st), {n: v for n, v in zip(names, pr)}, flush=True)
print("done")
