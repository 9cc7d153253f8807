"""Evaluation loop of the reference's eval.py (run_evaluation, eval.py:57-299) for the 3D-pose datasets
(h36m-p1 / h36m-p2 / mpi-inf-3dhp), over CACHED network inputs, image-sharded across ranks.

    python -m danet_b200.eval_h36m --checkpoint data/pretrained_model/danet_model_h36m_itw.pt \
           --cache data/dataset_extras/h36m_valid_protocol2_cached.npz --batch_size 32
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m danet_b200.eval_h36m ...    (r::world batches)

The cache is what the reference's BaseDataset hands to the model and the metric (eval.py:142-147,196-200): the
cropped / normalised image tensor and the 3D joints -- arrays `img` [N,3,224,224] float32 (or uint8 RGB crops, which
are normalised here like datasets/base_dataset.py does), `pose_3d` [N,24,4], optionally `imgname` [N].  Cropping
from raw frames (cv2, BaseDataset) is a data-loader concern and stays with the reference.

What runs where:
    model.infer_net -> SMPL -> J_regressor_h36m joints -> MPJPE      GPU (evaluate.evaluate_batch: one fused pass)
    reconstruction error (Procrustes / PA-MPJPE, utils/pose_utils.py:10-76)   host numpy SVD, as in the reference
    per-sample errors: one all_gather per batch round (parallel.gather_outputs)
Results are printed like the reference prints them (eval.py:271-299) and returned as a dict.
"""
import argparse
import os

import numpy as np
import torch

from . import constants
from .evaluate import evaluate_batch

IMG_NORM_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)       # constants.py:10-11 of the reference
IMG_NORM_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def compute_similarity_transform(S1, S2):
    """Similarity transform (s, R, t) taking S1 [N,3] closest to S2 [N,3] (orthogonal Procrustes, det R = +1);
    returns the transformed S1.  utils/pose_utils.py:10-57."""
    X1 = S1.T.astype(np.float64)
    X2 = S2.T.astype(np.float64)
    mu1, mu2 = X1.mean(axis=1, keepdims=True), X2.mean(axis=1, keepdims=True)
    X1c, X2c = X1 - mu1, X2 - mu2
    var1 = np.sum(X1c ** 2)
    K = X1c.dot(X2c.T)
    U, _s, Vh = np.linalg.svd(K)
    V = Vh.T
    Z = np.eye(U.shape[0])
    Z[-1, -1] *= np.sign(np.linalg.det(U.dot(V.T)))
    R = V.dot(Z.dot(U.T))
    scale = np.trace(R.dot(K)) / var1
    t = mu2 - scale * R.dot(mu1)
    return (scale * R.dot(X1) + t).T


def reconstruction_error(S1, S2):
    """Per-sample PA-MPJPE of S1 [B,J,3] against S2 [B,J,3] (utils/pose_utils.py:66-76, reduction=None)."""
    out = np.zeros(S1.shape[0])
    for i in range(S1.shape[0]):
        hat = compute_similarity_transform(S1[i], S2[i])
        out[i] = np.sqrt(((hat - S2[i]) ** 2).sum(axis=-1)).mean()
    return out


class CachedPoseDataset(object):
    """Cached inputs of a 3D-pose evaluation set (see module docstring)."""

    def __init__(self, path_or_dict):
        z = np.load(path_or_dict, allow_pickle=True) if isinstance(path_or_dict, str) else path_or_dict
        self.img = z["img"]
        self.pose_3d = np.asarray(z["pose_3d"], dtype=np.float32)
        self.imgname = [str(s) for s in z["imgname"]] if "imgname" in z else None
        if self.img.shape[0] != self.pose_3d.shape[0]:
            raise ValueError("cache: img and pose_3d disagree on the number of samples")

    def __len__(self):
        return self.img.shape[0]

    def batch(self, lo, hi):
        img = self.img[lo:hi]
        if img.dtype == np.uint8:                     # RGB crops [n,224,224,3] -> normalised NCHW (base_dataset.py:170-173)
            x = img.astype(np.float32) / 255.0
            x = (x - IMG_NORM_MEAN) / IMG_NORM_STD
            img = np.ascontiguousarray(x.transpose(0, 3, 1, 2))
        return torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)), self.pose_3d[lo:hi]


def run_evaluation(model, dataset_name, dataset, result_file=None, batch_size=32, log_freq=50, group=None, quiet=False,
                   batch_fn=None):
    """eval.py:57-299 for eval_pose datasets.  Every rank holds the dataset; batch b is evaluated by rank b % world
    (image-sharded) and the per-sample errors are all-gathered once at the end.  Returns
    dict(mpjpe [N] m, recon_err [N] m, mpjpe_mm, recon_err_mm, per_action).
    batch_fn(images, gt_j14) -> dict(mpjpe, pred_j14) replaces the GPU pass (host-logic tests)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    j17 = dataset_name == "mpi-inf-3dhp"                    # eval.py:139-140: 17 joints there, the 14 LSP joints otherwise
    if batch_fn is None:
        dev = next(model.parameters()).device
        smpl = model.iuv2smpl.smpl
        batch_fn = lambda im, gt: evaluate_batch(model, smpl, im.to(dev), gt.to(dev), shard=False, joints17=j17)
    else:
        dev = torch.device("cpu")
    nj = 17 if j17 else 14
    n = len(dataset)
    nb = (n + batch_size - 1) // batch_size
    mpjpe = np.zeros(n)
    recon = np.zeros(n)
    done = np.zeros(n, dtype=bool)
    pred_joints = np.zeros((n, nj, 3), dtype=np.float32)
    for b in range(rank, nb, world):
        lo, hi = b * batch_size, min(n, (b + 1) * batch_size)
        img, pose_3d = dataset.batch(lo, hi)
        gt = torch.from_numpy(pose_3d[:, constants.J24_TO_J17 if j17 else constants.J24_TO_J14, :3].copy())     # eval.py:189-190
        out = batch_fn(img, gt)                              # this rank's batch only (the loop deals the batches)
        e = out["mpjpe"].cpu().numpy()
        pj = out["pred_j14"].cpu().numpy()
        mpjpe[lo:hi] = e
        recon[lo:hi] = reconstruction_error(pj, gt.numpy())                        # host SVD, as the reference does
        pred_joints[lo:hi] = pj
        done[lo:hi] = True
        if not quiet and rank == 0 and (b // world) % log_freq == log_freq - 1:
            print("MPJPE: " + str(1000 * mpjpe[done].mean()))
            print("Reconstruction Error: " + str(1000 * recon[done].mean()))
            print()
    if world > 1:
        buf = torch.from_numpy(np.stack([mpjpe, recon, done.astype(np.float64)])).to(dev if dist.get_backend(group) == "nccl" else "cpu")
        pred_joints = pred_joints.astype(np.float32)
        dist.all_reduce(buf, group=group)                                             # disjoint supports: sum = merge
        pj = torch.from_numpy(pred_joints).to(buf.device)
        dist.all_reduce(pj, group=group)
        mpjpe, recon, cnt = buf[0].cpu().numpy(), buf[1].cpu().numpy(), buf[2].cpu().numpy()
        pred_joints = pj.cpu().numpy()
        assert (cnt == 1).all()
    res = {"mpjpe": mpjpe, "recon_err": recon, "mpjpe_mm": 1000 * mpjpe.mean(), "recon_err_mm": 1000 * recon.mean(),
           "pred_joints": pred_joints, "per_action": {}}
    if dataset_name == "h36m-p2" and dataset.imgname is not None:
        acts = {}
        for i, name in enumerate(dataset.imgname):
            a = name.split("/")[-1].split(".")[0].split("_")[1]                       # eval.py:150-157
            acts.setdefault(a, []).append(i)
        for a, idx in acts.items():
            res["per_action"][a] = (1000 * mpjpe[idx].mean(), 1000 * recon[idx].mean())
    if result_file is not None and rank == 0:
        np.savez(result_file, pred_joints=pred_joints, mpjpe=mpjpe, recon_err=recon)
    if not quiet and rank == 0:
        print("*** Final Results ***")
        print("MPJPE: " + str(res["mpjpe_mm"]))
        print("Reconstruction Error: " + str(res["recon_err_mm"]))
        print()
        if res["per_action"]:
            print(["action err"] + [str(v[1]) for v in res["per_action"].values()] + list(res["per_action"].keys()))
    return res


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--checkpoint", default=None, help="path to the network checkpoint (the reference's .pt with a 'model' entry)")
    ap.add_argument("--dataset", default="h36m-p2", choices=["h36m-p1", "h36m-p2", "mpi-inf-3dhp"])
    ap.add_argument("--cache", required=True, help="npz with img / pose_3d / imgname (see module docstring)")
    ap.add_argument("--batch_size", default=32, type=int)
    ap.add_argument("--log_freq", default=50, type=int)
    ap.add_argument("--result_file", default=None)
    ap.add_argument("--smpl_mean_params", default="data/smpl_mean_params.npz")
    ap.add_argument("--legacy_align_corners", action="store_true",
                    help="torch<=1.2 affine_grid/grid_sample semantics (the released checkpoint was trained under torch 1.1)")
    ap.add_argument("--precision", default="exact", choices=["exact", "fast"])
    ap.add_argument("--synthetic", action="store_true", help="random weights and synthetic SMPL assets (no licensed files)")
    args = ap.parse_args(argv)
    import torch.distributed as dist
    from .danet import DaNet, build_synthetic_danet
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.synthetic:
        model = build_synthetic_danet(width=48, device=dev, precision=args.precision, legacy_align_corners=args.legacy_align_corners)
    else:
        model = DaNet(args, args.smpl_mean_params, pretrained=False, precision=args.precision,
                      legacy_align_corners=args.legacy_align_corners)
        ck = torch.load(args.checkpoint, map_location="cpu")
        model.load_state_dict(ck["model"], strict=False)                               # eval.py:333-334
        model = model.to(dev).eval()
    res = run_evaluation(model, args.dataset, CachedPoseDataset(args.cache), args.result_file, args.batch_size, args.log_freq)
    if dist.is_initialized():
        dist.destroy_process_group()
    return res


if __name__ == "__main__":
    main()
