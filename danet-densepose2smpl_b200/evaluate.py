"""H36M-P2 style evaluation step of the reference (eval.py:142-216), GPU part only:

    pred = model.infer_net(images)                     eval.py:166
    verts = smpl(betas, rotmat[:,1:], rotmat[:,:1], pose2rot=False).vertices   eval.py:172-173
    J17  = J_regressor_h36m @ verts                    eval.py:186,202   (fused into the SMPL pass)
    J14  = (J17 - J17[:, [0]])[:, H36M_TO_J14]          eval.py:203-207
    mpjpe = ||J14 - gt||_2.mean(-1)                    eval.py:211       (csrc/lbs.cu k_mpjpe)

PA-MPJPE (per-sample numpy SVD, utils/pose_utils.py:10-76) stays on the host, as in the reference.
Image-sharded over ranks: every rank evaluates its contiguous slice and the per-sample errors are
gathered with one all_gather."""
import torch

from . import constants
from .parallel import gather_outputs, shard_bounds
from .smpl import mpjpe_h36m


@torch.no_grad()
def evaluate_batch(model, smpl, images, gt_keypoints_3d_j14, group=None, shard=True, joints17=False):
    """images [N,3,224,224]; gt_keypoints_3d_j14 [N,14,3] (pelvis-centred H36M joints in the
    H36M_TO_J14 order, eval.py:196-200).  Returns dict(mpjpe [N] in metres, pred_j14 [N,14,3],
    para [N,229]) on every rank.  shard=False: evaluate all N images on this rank, no collective
    (eval_h36m.run_evaluation deals whole batches to ranks instead)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if shard and dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n = images.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    dev = images.device
    if hi > lo:
        para = model.infer_net(images[lo:hi])["para"]
        rot = para[:, 13:].reshape(-1, 24, 3, 3)
        smpl(betas=para[:, 3:13].contiguous(), body_pose=rot[:, 1:], global_orient=rot[:, :1], pose2rot=False)
        j17 = smpl.joints_h36m()
        if j17 is None:
            raise ValueError("evaluate_batch: the SMPL layer was built without J_regressor_h36m (eval.py:78)")
        if joints17:
            # mpi-inf-3dhp (eval.py:139-140): all 17 H36M joints; gt is [N,17,3] in the J24_TO_J17 order.  Same expressions
            # as eval.py:202-211, in torch on the GPU (the fused kernel selects the 14 LSP joints)
            j14 = (j17 - j17[:, :1])[:, constants.H36M_TO_J17]
            err = torch.sqrt(((j14 - gt_keypoints_3d_j14[lo:hi].to(dev)) ** 2).sum(dim=-1)).mean(dim=-1)
        else:
            err = mpjpe_h36m(j17, gt_keypoints_3d_j14[lo:hi].to(dev))
            j14 = (j17 - j17[:, :1])[:, constants.H36M_TO_J14]
    else:
        para = torch.zeros(0, 229, device=dev)
        err = torch.zeros(0, device=dev)
        j14 = torch.zeros(0, 17 if joints17 else 14, 3, device=dev)
    if world == 1:
        return {"mpjpe": err, "pred_j14": j14, "para": para}
    return {"mpjpe": gather_outputs(err, n, group), "pred_j14": gather_outputs(j14, n, group),
            "para": gather_outputs(para, n, group)}
