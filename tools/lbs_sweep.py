"""Times the SMPL layer over batch sizes and routes (CUDA events): fused fp32 kernel (bodies_per_cta = 4/8/16; -1 forces it
for large batches) vs the tensor-core GEMM route (auto, B >= 512).  Dev tool; with `one` runs B = 8192 only (for ncu)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import danet_b200
from danet_b200 import synthetic as synth

dev = torch.device("cuda:0")
model = synth.make_smpl_model(0)
smpl = danet_b200.SMPL(model).to(dev)
res = []
one = "one" in sys.argv
for B in ((8192,) if one else (64, 512, 2048, 8192)):
    betas = torch.randn(B, 10, device=dev)
    x6 = torch.randn(B, 24, 6, device=dev)
    for nb in ((0,) if one else (-1, 4, 8, 16, 0)):
        if nb > 0 and B >= 512:
            continue                      # explicit blockings are the small-batch knob
        for _ in range(3):
            smpl(betas=betas, pose6d=x6, bodies_per_cta=nb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3 if one else 10
        e0.record()
        for _ in range(n):
            smpl(betas=betas, pose6d=x6, bodies_per_cta=nb)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        route = "gemm" if (nb == 0 and B >= 512) else "fused"
        res.append({"B": B, "nb": nb, "route": route, "ms": ms, "bodies_per_s": B / ms * 1e3, "verts_per_s": B * 6890 / ms * 1e3})
        print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
if not one:
    json.dump(res, open("gpurun_out/lbs_sweep.json", "w"), indent=1)
