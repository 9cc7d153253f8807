"""One layer class of tools/tc_layers.py, a few launches (for ncu): python tools/tc_one.py <index> <exact 0|1> [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tc_layers

if __name__ == "__main__":
    idx, exact = int(sys.argv[1]), int(sys.argv[2])
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    case, _ = tc_layers.LAYERS[idx]
    us, tf = tc_layers.time_layer(case, bool(exact), iters=iters)
    print(case, "exact" if exact else "fast", "%.1f us %.1f TFLOP/s" % (us, tf))
