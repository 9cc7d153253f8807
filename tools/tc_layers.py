"""Per-layer timing of the tensor-core convolution engine on the layer classes that make up the DaNet step
(B = 64, HRNet-W48), through the C ABI.  Usage: python tools/tc_layers.py [tag]
Env: DANET_TC_S, DANET_TC_VARIANT (knock-outs), DANET_TC_MMAORDER, DANET_TC_SWB.  Usage: tc_layers.py [tag] [exact|fast] [s1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conv_tc_common import DEV, desc, launch, make_case, pack, problem, split

# N, H, W, Cin, Cout, k, stride, wsets, relu, residual ; count = launches of this class per step
LAYERS = [
    ((64, 56, 56, 48, 48, 3, 1, 1, 1, 1), 32), ((64, 56, 56, 48, 48, 3, 1, 1, 1, 0), 32),
    ((64, 28, 28, 96, 96, 3, 1, 1, 1, 1), 32), ((64, 28, 28, 96, 96, 3, 1, 1, 1, 0), 32),
    ((64, 14, 14, 192, 192, 3, 1, 1, 1, 1), 28), ((64, 14, 14, 192, 192, 3, 1, 1, 1, 0), 28),
    ((64, 7, 7, 384, 384, 3, 1, 1, 1, 1), 12), ((64, 7, 7, 384, 384, 3, 1, 1, 1, 0), 12),
    ((64, 56, 56, 64, 256, 1, 1, 1, 1, 1), 4), ((64, 56, 56, 256, 64, 1, 1, 1, 1, 0), 3),
    ((1536, 56, 56, 64, 64, 7, 2, 1, 1, 0), 1), ((1536, 56, 56, 48, 24, 3, 1, 24, 0, 0), 1),
    ((1536, 56, 56, 24, 64, 1, 1, 1, 1, 0), 1), ((1536, 4, 4, 256, 256, 3, 1, 1, 1, 1), 3),
    ((1536, 14, 14, 64, 64, 3, 1, 1, 1, 1), 4), ((1536, 7, 7, 128, 128, 3, 1, 1, 1, 1), 3),
    ((64, 56, 56, 48, 96, 3, 2, 1, 0, 0), 7), ((64, 224, 224, 8, 64, 3, 2, 1, 1, 0), 1),
]


def time_layer(case, exact, iters=10):
    N, H, W, Cin, Cout, k, s, G, relu, has_res = case
    g = torch.Generator().manual_seed(1)
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    x = torch.randn(N, H, W, Cin, generator=g).to(DEV)
    w = (torch.randn(G, k * k * Cin, Cout, generator=g) * (1.0 / (k * k * Cin)) ** 0.5).to(DEV)
    b = (torch.randn(G, Cout, generator=g) * 0.1).to(DEV)
    d = desc(case, exact)
    xp = split(x, want_lo=exact)
    del x
    wpk = pack(d, w)
    rp = None
    if has_res:
        r = torch.randn(N, Ho, Wo, Cout, generator=g).to(DEV)
        rp = split(r, want_lo=exact)
        del r
    yh = torch.empty(N, Ho, Wo, Cout, dtype=torch.float16, device=DEV)
    yl = torch.empty(N, Ho, Wo, Cout, dtype=torch.float16, device=DEV) if exact else None
    p = problem(d, xp, wpk, b, res_planes=rp, y_planes=(yh, yl))
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
    flop = 2.0 * N * Ho * Wo * k * k * Cin * Cout
    return us, flop / us / 1e6


MODES = (0, 1)

if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else ""
    if len(sys.argv) > 2:
        MODES = (1,) if sys.argv[2] == "exact" else (0,)
    if len(sys.argv) > 3 and sys.argv[3] == "s1":
        # the classes whose exact-mode N tiles (96 / 128 channels) leave one accumulator chain per CTA, + one 48-channel control
        LAYERS = [l for l in LAYERS if l[0][4] in (96, 192, 384, 128, 256) and l[0][5] == 3 and l[0][6] == 1] + [LAYERS[0]]
    tot = {0: 0.0, 1: 0.0}
    print("# %s S=%s variant=%s" % (tag, os.environ.get("DANET_TC_S", "2"), os.environ.get("DANET_TC_VARIANT", "0")))
    for case, cnt in LAYERS:
        row = []
        for exact in MODES:
            us, tf = time_layer(case, bool(exact))
            tot[exact] += us * cnt
            row.append("%8.1f us %6.1f TF" % (us, tf))
        if len(MODES) == 2:
            print("%-46s x%-2d fast %s | exact %s" % (case, cnt, row[0], row[1]), flush=True)
        else:
            print("%-46s x%-2d %s %s" % (case, cnt, "exact" if MODES[0] else "fast", row[0]), flush=True)
    print("weighted ms: fast %.2f exact %.2f" % (tot[0] / 1e3, tot[1] / 1e3))
