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
