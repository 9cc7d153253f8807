"""Runs a few representative tcgen05 conv launches (for `ncu --set full` captures)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from danet_b200.plan import CudaOps
DEV = "cuda:0"
ops = CudaOps(DEV)
cases = [(64, 56, 56, 48, 48, 3, 1), (64, 28, 28, 96, 96, 3, 1), (64, 14, 14, 192, 192, 3, 1), (1536, 56, 56, 64, 64, 7, 2)]
if len(sys.argv) > 1:      # e.g. 64,56,56,64,256,1,1 ...
    cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (N, H, W, Cin, Cout, k, st) in cases:
    d = dict(N=N, H=H, W=W, Cin=Cin, Cout=Cout, ksize=k, stride=st, pad=k // 2, wsets=1, relu=1)
    Ho, Wo = (H + 2 * (k // 2) - k) // st + 1, (W + 2 * (k // 2) - k) // st + 1
    x = torch.randn(N, H, W, Cin, device=DEV)
    w = torch.randn(1, k * k * Cin, Cout, device=DEV) * 0.05
    b = torch.randn(1, Cout, device=DEV) * 0.1
    res = torch.randn(N, Ho, Wo, Cout, device=DEV) if st == 1 else None
    y = torch.empty(N, Ho, Wo, Cout, device=DEV)
    wp = ops.conv_tc_pack(d, w)
    prof = torch.zeros(16 + 16 * 160, dtype=torch.int64, device=DEV)
    if not os.environ.get("TC_NOPROF"):          # TC_NOPROF=1: plain instantiation (for ncu captures)
        ops.lib.danet_conv_tc_set_profile_buffer(ctypes.c_void_p(prof.data_ptr()))
    for _ in range(2):
        ops.conv2d(d, 1, x, wp, b, res, y)
    torch.cuda.synchronize()
    ops.lib.danet_conv_tc_set_profile_buffer(ctypes.c_void_p(0))
    pr = prof.cpu().tolist()
    names = ["mma_total", "mma_wait_acc_empty", "mma_wait_a_full", "mma_wait_b_full", "prod_total", "prod_wait_a_empty", "epi_total", "epi_wait_acc_full",
             "epi_decode", "epi_fetch", "epi_wait", "epi_tmem_ld", "epi_finish", "epi_arrive"]
    print((N, H, Cin, Cout, k, st), {n: v for n, v in zip(names, pr)}, flush=True)
    if os.environ.get("TC_NOPROF"):
        continue
    import numpy as np
    tl = np.array(pr[16:16 + 16 * 148]).reshape(148, 16)
    act = tl[:, 0] > 0
    t0 = tl[act, 0].min()
    print("   grid %d: entry spread %.1f us, kernel span (first entry -> last exit) %.1f us" % (act.sum(), (tl[act, 0].max() - t0) / 1e3, (tl[act, 1].max() - t0) / 1e3))
    for name, col in (("prologue_done", 2), ("first_b_full", 3), ("first_a_full", 4), ("last_mma_commit", 5), ("epi_first_acc_full", 6), ("epi_done", 7), ("prod_loop_start", 8), ("prod_ldg_issued", 9), ("prod_ldg_returned", 10), ("prod_arrived", 11), ("b_first_issue", 12)):
        v = tl[act, col] / 1.9e3
        print("   %-20s us from CTA entry: min %.1f median %.1f max %.1f   (CTA0 %.1f)" % (name, v.min(), np.median(v), v.max(), v[0]))
    span = (tl[act, 1] - tl[act, 0]) / 1e3
    print("   CTA lifetime us: min %.1f median %.1f max %.1f" % (span.min(), np.median(span), span.max()))
print("done")
