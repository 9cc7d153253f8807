"""Map the conv launches of the LAST bench step in an ncu CSV (bench.py --no-graph) onto the graph's
conv ops (launch order = graph order) and aggregate device time / DRAM bytes per layer shape.
usage: python tools/ncu_per_layer.py gpurun_out/ncu_final.csv [width] [batch]"""
import collections
import csv
import sys

sys.path.insert(0, ".")
from danet_b200 import netgraph  # noqa: E402


def main():
    path = sys.argv[1]
    width = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    g = netgraph.danet_graph(width)
    convs = [op for op in g.ops if op["op"] == "conv"]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    d = collections.OrderedDict()
    for x in csv.DictReader(lines):
        kn = x["Kernel Name"]
        if "k_conv_tc" in kn or "k_conv_simt" in kn:
            e = d.setdefault(int(x["ID"]), {"k": "tc" if "k_conv_tc" in kn else "simt"})
            e[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    # launch order = graph order; the capture may end mid-step, so take the SECOND complete step
    # (the first one is the warm-up) when there is one, else the first
    allids = sorted(d)
    start = len(convs) if len(allids) >= 2 * len(convs) else 0
    ids = allids[start:start + len(convs)]
    agg = collections.OrderedDict()
    tot = 0.0
    for i, op in zip(ids, convs):
        x, y = op["x"], op["y"]
        key = (d[i]["k"], B * x.nmult, x.H, x.Cp, y.Cp, op["k"], op["stride"], op["groups"], op["res"] is not None)
        us = d[i]["gpu__time_duration.sum"] / 1e3
        Ho = y.H
        fl = 2.0 * B * x.nmult * Ho * Ho * x.Cp * y.Cp * op["k"] ** 2
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += fl
        a[3] += d[i].get("dram__bytes_read.sum", 0.0); a[4] += d[i].get("dram__bytes_write.sum", 0.0)
        tot += us
    print("# conv launches of one step under ncu (cold L2, serialised): %.1f us over %d launches" % (tot, len(convs)))
    print("%-5s %5s %4s %4s %4s %1s %1s %3s %3s %4s %9s %6s %8s %8s %8s" %
          ("algo", "N", "H", "Cin", "Cout", "k", "s", "g", "res", "n", "total_us", "share", "us/conv", "TFLOP/s", "GB/s"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-5s %5d %4d %4d %4d %1d %1d %3d %3d %4d %9.1f %5.1f%% %8.1f %8.1f %8.0f" %
              (key + (a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2] / a[1] / 1e6, (a[3] + a[4]) / a[1] / 1e3)))


if __name__ == "__main__":
    main()
