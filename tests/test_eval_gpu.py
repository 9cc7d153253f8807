"""eval.py:57-299 loop (danet_b200.eval_h36m.run_evaluation) on the GPU over a synthetic cached set: per-sample MPJPE
against the oracle chain (reference para -> oracle SMPL -> H36M joints -> MPJPE), PA-MPJPE, per-action table, and the
image-sharded 2-GPU run against the single-GPU one."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from net_common import build, make_image

pytestmark = pytest.mark.gpu
N_SAMPLES = 256


def _dataset(n=N_SAMPLES):
    rng = np.random.default_rng(21)
    acts = ["Directions", "Eating", "Walking", "Sitting"]
    names = ["images/S9_%s_1.60457274_%06d.jpg" % (acts[i % 4], i) for i in range(n)]
    pose = (rng.normal(size=(n, 24, 4)) * 0.2).astype(np.float32)
    return {"img": make_image(n, 77).numpy(), "pose_3d": pose, "imgname": np.array(names)}


def test_eval_loop_matches_oracle_chain_over_256_samples():
    from danet_b200 import constants, eval_h36m
    from oracle import lbs, synth
    net = build(32, "cuda:0", conv_algo="auto")
    data = _dataset()
    res = eval_h36m.run_evaluation(net, "h36m-p2", eval_h36m.CachedPoseDataset(data), batch_size=32, quiet=True)
    assert res["mpjpe"].shape == (N_SAMPLES,) and set(res["per_action"]) == {"Directions", "Eating", "Walking", "Sitting"}
    # oracle chain from the para the network produced for the same batches (ragged last batch not needed: 256 = 8 x 32)
    para = np.concatenate([net.infer_net(torch.from_numpy(data["img"][lo:lo + 32]).cuda())["para"].cpu().numpy()
                           for lo in range(0, N_SAMPLES, 32)])
    R = para[:, 13:].reshape(-1, 24, 3, 3)
    ref = lbs.smpl_forward(synth.make_smpl_model(0), para[:, 3:13], R[:, 1:], R[:, :1], pose2rot=False, dtype=np.float64)
    gt = data["pose_3d"][:, constants.J24_TO_J14, :3].astype(np.float64)
    want = lbs.mpjpe_h36m(ref["joints_h36m"], gt)
    np.testing.assert_allclose(res["mpjpe"], want, atol=1e-5)
    assert abs(res["mpjpe_mm"] - 1000 * want.mean()) < 1e-2
    assert (res["recon_err"] <= res["mpjpe"] + 1e-6).all()              # Procrustes alignment can only reduce the error
    for a, (m_, r_) in res["per_action"].items():
        idx = [i for i, nm in enumerate(data["imgname"]) if "_%s_" % a in nm]
        assert abs(m_ - 1000 * res["mpjpe"][idx].mean()) < 1e-6


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from danet_b200 import eval_h36m
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    net = build(32, "cuda:%d" % rank, conv_algo="auto")
    res = eval_h36m.run_evaluation(net, "h36m-p2", eval_h36m.CachedPoseDataset(_dataset(96)), batch_size=16, quiet=True)
    if rank == 0:
        torch.save({k: res[k] for k in ("mpjpe", "recon_err", "pred_joints")}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_eval_loop_two_gpus_equals_one(tmp_path):
    from danet_b200 import eval_h36m
    out = str(tmp_path / "eval2.pt")
    mp.spawn(_worker, args=(2, 29561, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    net = build(32, "cuda:0", conv_algo="auto")
    want = eval_h36m.run_evaluation(net, "h36m-p2", eval_h36m.CachedPoseDataset(_dataset(96)), batch_size=16, quiet=True)
    # images are independent and the tensor-core path is batch-invariant: dealing batches to two GPUs changes nothing
    np.testing.assert_array_equal(got["mpjpe"], want["mpjpe"])
    np.testing.assert_array_equal(got["pred_joints"], want["pred_joints"])


def test_eval_loop_mpi_inf_3dhp_uses_17_joints():
    """eval.py:139-140,189-211 for mpi-inf-3dhp: H36M_TO_J17 / J24_TO_J17, torch path; against the oracle chain."""
    from danet_b200 import constants, eval_h36m
    from oracle import lbs, synth
    net = build(32, "cuda:0", conv_algo="auto")
    data = _dataset(48)
    res = eval_h36m.run_evaluation(net, "mpi-inf-3dhp", eval_h36m.CachedPoseDataset(data), batch_size=16, quiet=True)
    assert res["pred_joints"].shape == (48, 17, 3) and not res["per_action"]
    para = np.concatenate([net.infer_net(torch.from_numpy(data["img"][lo:lo + 16]).cuda())["para"].cpu().numpy() for lo in range(0, 48, 16)])
    R = para[:, 13:].reshape(-1, 24, 3, 3)
    ref = lbs.smpl_forward(synth.make_smpl_model(0), para[:, 3:13], R[:, 1:], R[:, :1], pose2rot=False, dtype=np.float64)
    jh = ref["joints_h36m"]
    pred = (jh - jh[:, :1])[:, constants.H36M_TO_J17]
    gt = data["pose_3d"][:, constants.J24_TO_J17, :3].astype(np.float64)
    want = np.sqrt(((pred - gt) ** 2).sum(-1)).mean(-1)
    np.testing.assert_allclose(res["mpjpe"], want, atol=1e-5)
