"""N>1 host logic on CPU: world_size-2 gloo ranks shard a batch, run the plan through the torch
test double and all-gather `para`; the result must equal the single-process run."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from danet_b200 import parallel
    from net_common import build, infer_with_ops, make_image
    from oracle.net_ops import TorchEmulOps
    net, emul = build(32), TorchEmulOps()
    img = make_image(3, 100)                     # 3 images over 2 ranks: ragged shards (2 + 1)
    para = parallel.infer_sharded(net, img, infer=lambda x: infer_with_ops(net, x, emul)["para"])
    if rank == 0:
        torch.save(para, tmp)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    from danet_b200.parallel import shard_bounds
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            cover = []
            for r in range(w):
                lo, hi = shard_bounds(n, w, r)
                cover += list(range(lo, hi))
            assert cover == list(range(n))


def test_two_rank_gloo_matches_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from net_common import build, infer_with_ops, make_image
    from oracle.net_ops import TorchEmulOps
    out = str(tmp_path / "para.pt")
    mp.spawn(_worker, args=(2, 29541, out), nprocs=2, join=True)
    got = torch.load(out)
    want = infer_with_ops(build(32), make_image(3, 100), TorchEmulOps())["para"]
    assert got.shape == (3, 229)
    assert (got - want).abs().max() < 1e-5


def _eval_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = _run_eval_stub()
    if rank == 0:
        torch.save({k: res[k] for k in ("mpjpe", "recon_err", "mpjpe_mm", "recon_err_mm", "per_action")}, tmp)
    dist.barrier()
    dist.destroy_process_group()


def _eval_set(n=70):
    import numpy as np
    rng = np.random.default_rng(11)
    names = ["images/S9_%s_1.60457274_%06d.jpg" % (["Directions", "Eating", "Walking"][i % 3], i) for i in range(n)]
    return {"img": rng.normal(size=(n, 3, 8, 8)).astype(np.float32), "pose_3d": rng.normal(size=(n, 24, 4)).astype(np.float32),
            "imgname": np.array(names)}


def _run_eval_stub():
    """run_evaluation with the GPU pass replaced by a deterministic function of the inputs (host logic only)."""
    from danet_b200 import eval_h36m

    def batch_fn(img, gt):
        pj = gt + 0.05 * img[:, :, :, 0].mean(dim=(1, 2)).reshape(-1, 1, 1) + 0.01 * torch.sin(gt * 7)
        return {"mpjpe": torch.sqrt(((pj - gt) ** 2).sum(-1)).mean(-1), "pred_j14": pj}
    return eval_h36m.run_evaluation(None, "h36m-p2", eval_h36m.CachedPoseDataset(_eval_set()), batch_size=16, quiet=True,
                                    batch_fn=batch_fn)


def test_eval_loop_two_ranks_equals_one(tmp_path):
    """eval.py:142-216 loop, batches dealt r::world over 2 gloo ranks (ragged: 70 samples, batch 16) == single rank;
    per-action table (eval.py:150-157,283-299) and PA-MPJPE from the host SVD."""
    import numpy as np
    out = str(tmp_path / "eval.pt")
    mp.spawn(_eval_worker, args=(2, 29547, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    want = _run_eval_stub()
    np.testing.assert_allclose(got["mpjpe"], want["mpjpe"], atol=1e-7)
    np.testing.assert_allclose(got["recon_err"], want["recon_err"], atol=1e-7)
    assert abs(got["mpjpe_mm"] - want["mpjpe_mm"]) < 1e-6 and set(got["per_action"]) == {"Directions", "Eating", "Walking"}
    assert (want["recon_err"] <= want["mpjpe"] + 1e-9).all()          # Procrustes can only reduce the error


def test_reconstruction_error_matches_independent_procrustes():
    """utils/pose_utils.py:10-76 restated in eval_h36m: a similarity-transformed copy has zero PA error; against
    scipy's orthogonal Procrustes on centred, scale-solved data for random pairs."""
    import numpy as np
    from scipy.linalg import orthogonal_procrustes
    from scipy.spatial.transform import Rotation
    from danet_b200.eval_h36m import reconstruction_error
    rng = np.random.default_rng(2)
    S2 = rng.normal(size=(5, 14, 3))
    R = Rotation.random(5, random_state=3).as_matrix()
    S1 = np.einsum("bij,bkj->bki", R, S2) * 1.7 + rng.normal(size=(5, 1, 3))
    assert reconstruction_error(S1, S2).max() < 1e-9
    A, Bm = rng.normal(size=(14, 3)), rng.normal(size=(14, 3))
    Ac, Bc = A - A.mean(0), Bm - Bm.mean(0)
    Rp, sca = orthogonal_procrustes(Ac, Bc)
    if np.linalg.det(Rp) > 0:                                           # same optimum when no reflection is needed
        hat = sca / (Ac ** 2).sum() * Ac.dot(Rp) + Bm.mean(0)
        want = np.sqrt(((hat - Bm) ** 2).sum(-1)).mean()
        assert abs(reconstruction_error(A[None], Bm[None])[0] - want) < 1e-9


def _grad_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from danet_b200 import parallel
    g = torch.Generator().manual_seed(7)
    shapes = [(64, 3, 3, 3), (64,), (5,), (128, 64, 1, 1), (0,), (13, 512)]
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    base = [torch.randn(s, generator=g) for s in shapes]
    for p, b in zip(params, base):
        p.grad = b * (rank + 1)                       # rank r contributes (r + 1) * base
    params[2].grad = None                             # a parameter without a gradient is skipped on every rank
    half = torch.nn.Parameter(torch.zeros(7, dtype=torch.float64))
    half.grad = torch.full((7,), float(rank + 1), dtype=torch.float64)      # another dtype: its own bucket
    n = parallel.all_reduce_gradients(params + [half], bucket_bytes=40000)
    if rank == 0:
        torch.save({"n": n, "grads": [None if p.grad is None else p.grad.clone() for p in params], "half": half.grad.clone(),
                    "base": base}, tmp)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_all_reduce(tmp_path):
    """Training-step exchange: bucketed all-reduce = mean over ranks, tensor by tensor."""
    from danet_b200 import parallel
    out = str(tmp_path / "grads.pt")
    mp.spawn(_grad_worker, args=(2, 29547, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["n"] >= 3                                # 40 KB buckets split the ~70 KB of fp32 gradients; fp64 gets its own
    for got, b in zip(r["grads"], r["base"]):
        if got is None:
            continue
        assert torch.allclose(got, b * 1.5, rtol=1e-6, atol=1e-7)            # mean of 1x and 2x
    assert r["grads"][2] is None
    assert torch.equal(r["half"], torch.full((7,), 1.5, dtype=torch.float64))
    assert parallel.all_reduce_gradients([torch.nn.Parameter(torch.zeros(3))]) == 0      # no process group: nothing to do
