"""Per-role cycle accounting of CTA 0 for a few layer classes (danet_conv_tc_set_profile_buffer).
python tools/tc_roles.py [layer indices...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import tc_layers
from danet_b200 import _lib

NAMES = ["A total", "A wait a_empty", "B total", "B wait b_empty", "B wait sched", "MMA total", "MMA wait a_full", "MMA wait b_full",
         "MMA wait acc_empty", "MMA wait sched", "EPI total", "EPI wait acc_full", "EPI wait sched", "EPI init(bias/res)", "EPI tmem segs", "EPI stores"]

if __name__ == "__main__":
    idxs = [int(a) for a in sys.argv[1:]] or [1, 0, 3, 5, 7, 8]
    lib = _lib.load()
    buf = torch.zeros(16, dtype=torch.int64, device="cuda:0")
    for i in idxs:
        case, _ = tc_layers.LAYERS[i]
        for exact in (0, 1):
            lib.danet_conv_tc_set_profile_buffer(None)
            us, tf = tc_layers.time_layer(case, bool(exact), iters=3)
            buf.zero_()
            lib.danet_conv_tc_set_profile_buffer(ctypes_ptr := buf.data_ptr())
            tc_layers.time_layer(case, bool(exact), iters=1)
            torch.cuda.synchronize()
            v = buf.cpu().tolist()
            lib.danet_conv_tc_set_profile_buffer(None)
            print("%s %s: %.1f us (%.0f cycles @1.965 GHz)" % (case, "exact" if exact else "fast", us, us * 1965))
            print("   " + "  ".join("%s=%d" % (n, x) for n, x in zip(NAMES, v)))
