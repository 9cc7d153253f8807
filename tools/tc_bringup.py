"""Bring-up of the tensor-core convolution kernel: every case in its own process (a trap kills the CUDA
context), both precisions, residual as fp32 and as planes.  Usage: python tools/tc_bringup.py [quick]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [
    (2, 16, 8, 16, 16, 1, 1, 1, 0, 0),           # one tile, one K step
    (2, 16, 8, 64, 48, 1, 1, 1, 0, 0),
    (2, 16, 16, 16, 16, 3, 1, 1, 0, 0),          # two sub-tiles, 3x3 taps
    (2, 56, 56, 48, 48, 3, 1, 1, 1, 1),
    (2, 28, 28, 96, 96, 3, 1, 1, 1, 1),
    (3, 14, 14, 192, 192, 3, 1, 1, 0, 0),
    (2, 7, 7, 384, 384, 3, 1, 1, 1, 1),
    (2, 56, 56, 64, 256, 1, 1, 1, 1, 1),
    (2, 56, 56, 256, 64, 1, 1, 1, 1, 0),
    (2, 56, 56, 48, 96, 3, 2, 1, 0, 0),
    (1, 224, 224, 8, 64, 3, 2, 1, 1, 0),
    (2, 56, 56, 64, 64, 7, 2, 1, 1, 0),
    (48, 56, 56, 48, 24, 3, 1, 24, 0, 0),
    (48, 4, 4, 256, 128, 3, 2, 24, 1, 0),
    (48, 4, 4, 256, 128, 1, 2, 24, 0, 0),
    (48, 2, 2, 128, 128, 3, 1, 24, 1, 1),
    (2, 2, 2, 512, 512, 3, 1, 1, 1, 1),
    (2, 56, 56, 48, 96, 3, 1, 1, 0, 0),
    (2, 56, 56, 80, 64, 1, 1, 1, 1, 0),
    (5, 13, 9, 24, 40, 3, 1, 1, 1, 1),            # ragged: nothing divides the tile sizes
    (2, 14, 14, 64, 64, 3, 1, 1, 1, 1),
    (2, 28, 28, 64, 128, 3, 2, 1, 1, 0),
    (2, 28, 28, 64, 128, 1, 2, 1, 0, 0),
    (13, 7, 7, 64, 64, 3, 1, 1, 1, 1),           # stacked small maps: 2 images per tile, ragged batch
    (7, 4, 4, 64, 64, 3, 1, 1, 1, 1),            # 3 images per tile (rows of 5)
    (11, 2, 2, 32, 48, 3, 1, 1, 1, 1),           # 5 images per tile
    (9, 4, 4, 32, 32, 1, 1, 1, 0, 0),            # 1x1: no zero rows, 4 images per tile
    (5, 5, 3, 16, 16, 3, 1, 1, 0, 1),            # rows of 6: 2 images per tile, narrow
]


if __name__ == "__main__":
    import torch
    from conv_tc_common import run_case
    quick = "quick" in sys.argv
    idxs = list(range(4)) if quick else list(range(len(CASES)))
    bad = 0
    for S in ("2", "1"):
        os.environ["DANET_TC_S"] = S
        for i in idxs:
            if S == "1" and CASES[i][2] <= 8:
                continue
            for exact in (0, 1):
                planes = (i + exact) % 2
                case = CASES[i]
                try:
                    y, ym, ref = run_case(case, bool(exact), res_as_planes=bool(planes))
                except Exception as e:       # a trap kills the context: nothing after this can run
                    print("FAIL S=%s case %r exact %d: %s" % (S, case, exact, str(e)[-300:]), flush=True)
                    print("bring-up: aborted")
                    sys.exit(1)
                e1 = (y.double() - ref).abs().max().item()
                e2 = (ym.double() - ref).abs().max().item()
                tol = 3e-5 if exact else 1.5e-2
                ok = e1 < tol and e2 < tol and not bool(torch.isnan(y).any())
                bad += 0 if ok else 1
                print("%s S=%s exact=%d planes=%d err_f32=%.3e err_planes=%.3e refmax=%.2f case=%r" %
                      ("OK  " if ok else "FAIL", S, exact, planes, e1, e2, ref.abs().max().item(), case), flush=True)
    print("bring-up: %d failures" % bad)
