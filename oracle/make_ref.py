"""Recipe for oracle/_ref: the reference's OWN network-half modules, taken from /root/reference where they lie
(this container) into a git-ignored directory that travels to the GPU box with the snapshot, so that
`bench.py --impl reference` and the cpu_baseline leg time the reference's code there too ("kind": "reference").

TEST / BENCH INFRASTRUCTURE.  Nothing is copied into the repository's history (oracle/_ref is in .gitignore);
the product package never imports it.  What is taken = the import closure of oracle/ref_import.load():
every module file the reference's IUV_Estimator / DecomposedPredictor / iuvmap / geometry import from its own
tree, plus the two data files they open (configs/danet_default.yaml, data/pretrained_model/learned_ratio.pkl).
The third-party halves (smplx LBS, neural_renderer) are absent from /root/reference and stay restated
(oracle/lbs.py, oracle/raster.c).

    python -m oracle.make_ref          # no-op when /root/reference is absent
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
DATA = ["configs/danet_default.yaml", "data/pretrained_model/learned_ratio.pkl"]


def main():
    if not os.path.isdir(os.path.join(SRC, "models", "danet")):
        print("make_ref: %s not present, nothing to do" % SRC)
        return 0
    marker = os.path.join(DST, ".complete")
    if os.path.exists(marker):
        return 0
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    cwd = os.getcwd()
    os.environ["DANET_REFERENCE"] = SRC
    from oracle import ref_import
    ref_import.load(48)
    os.chdir(cwd)
    files = set()
    for m in list(sys.modules.values()):
        f = getattr(m, "__file__", None)
        if f and os.path.abspath(f).startswith(SRC + os.sep) and f.endswith(".py"):
            files.add(os.path.relpath(os.path.abspath(f), SRC))
    # package markers of every directory on the way
    for f in list(files):
        d = os.path.dirname(f)
        while d:
            init = os.path.join(d, "__init__.py")
            if os.path.exists(os.path.join(SRC, init)):
                files.add(init)
            d = os.path.dirname(d)
    files.update(DATA)
    for f in sorted(files):
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), dst)
    with open(marker, "w") as fh:
        fh.write("\n".join(sorted(files)) + "\n")
    print("make_ref: %d files -> %s" % (len(files), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
