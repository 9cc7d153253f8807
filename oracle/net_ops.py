"""Torch (CPU) restatement of the kernel-level ops behind plan.CudaOps -- TEST INFRASTRUCTURE
(oracle/__init__.py): the checker of the network-half kernels, the test double that lets the host
logic run without a GPU, and (driven through the same graph) the CPU port timed by bench.py's
cpu_baseline / --impl reference legs when /root/reference is absent.  Used only to
check the host logic (graph wiring, BN folding, weight packing, buffer planning) on machines
without a GPU.  Each method restates the semantics documented in include/danet_b200.h with plain
torch ops; it is never used by the package itself."""
import numpy as np
import torch
import torch.nn.functional as F

PARENTS0 = [0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]
CHILDREN1 = [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 10, 11, 15, 16, 17, 15, 18, 19, 20, 21, 22, 23, 22, 23]
SMPL2DP = [[1, 2], [8, 10], [7, 9], [1, 2], [8, 10, 12, 14], [7, 9, 11, 13], [1, 2], [12, 14, 5], [11, 13, 6],
           [1, 2], [12, 14, 5], [11, 13, 6], [1, 2, 23, 24], [15, 17], [16, 18], [23, 24], [15, 17], [16, 18],
           [15, 17, 19, 21], [16, 18, 20, 22], [19, 21, 4], [20, 22, 3], [19, 21, 4], [20, 22, 3]]


class TorchEmulOps(object):
    """Same interface as danet_b200.plan.CudaOps; activations arrive as plan.ActBuf objects whose fp32 view is
    the only one present (planes() == 0)."""

    def planes(self, precision):
        return 0

    def conv_tc_supported(self, d):
        return True

    def conv_tc_pack(self, d, w):
        return w

    def conv_group(self, convs):
        for cv in convs:
            self.conv2d(cv["d"], cv["x"], cv["w"], cv["b"], cv["res"], cv["y"])

    def conv2d(self, d, x, w, bias, res, y):
        x, y = x.f32, y.f32
        res = res.f32 if res is not None else None
        N, H, W, Cin = x.shape
        k, G = d["ksize"], d["wsets"]
        Cout = d["Cout"]
        xs = x.permute(0, 3, 1, 2)
        out = torch.empty(N, Cout, y.shape[1], y.shape[2])
        for g in range(G):
            wg = w[g].reshape(k, k, Cin, Cout).permute(3, 2, 0, 1)
            out[g::G] = F.conv2d(xs[g::G], wg, bias[g], stride=d["stride"], padding=d["pad"])
        out = out.permute(0, 2, 3, 1)
        if res is not None:
            out = out + res
        if d["relu"]:
            out = torch.relu(out)
        y.copy_(out)

    def nchw_to_nhwc(self, x, y):
        y = y.f32
        y.zero_()
        y[..., :x.shape[1]] = x.permute(0, 2, 3, 1)

    def fuse_sum(self, terms, factors, relu, y, shape=None):
        acc = None
        for t, f in zip(terms, factors):
            t = t.f32
            u = t.repeat_interleave(f, dim=1).repeat_interleave(f, dim=2) if f > 1 else t
            acc = u.clone() if acc is None else acc + u
        y.f32.copy_(torch.relu(acc) if relu else acc)

    def maxpool(self, x, y, shape=None):
        y.f32.copy_(F.max_pool2d(x.f32.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1))

    def avgpool(self, x, y, shape=None):
        x = x.f32
        y.reshape(x.shape[0], x.shape[-1]).copy_(x.reshape(x.shape[0], -1, x.shape[-1]).mean(1))

    def linear(self, x, w, b, add, y):
        y.reshape(-1)[:x.shape[0] * w.shape[0]].view(x.shape[0], w.shape[0]).copy_(x @ w.t() + b + add)

    def clean_global(self, heads, body, amax, vis, shape=None):
        heads, body = heads.f32, body.f32
        U, V, I, A = heads[..., 0:25], heads[..., 25:50], heads[..., 50:75], heads[..., 75:90]
        idx = I.argmax(-1)
        oh = F.one_hot(idx, 25).float()
        body.zero_()
        body[..., 0:25] = oh * U
        body[..., 25:50] = oh * V
        body[..., 50:75] = oh
        amax.copy_(idx.to(torch.uint8))
        if vis is not None:
            vis[0].copy_((oh * U).permute(0, 3, 1, 2)); vis[1].copy_((oh * V).permute(0, 3, 1, 2))
            vis[2].copy_(oh.permute(0, 3, 1, 2))
            vis[3].copy_(F.one_hot(A.argmax(-1), 15).float().permute(0, 3, 1, 2))

    def clean_parts(self, x, y, raw, shape=None):
        x, y = x.f32, y.f32
        U, V, I = x[..., 0:7], x[..., 7:14], x[..., 14:21]
        oh = F.one_hot(I.argmax(-1), 7).float()
        y.zero_()
        y[..., 0:7] = oh * U; y[..., 7:14] = oh * V; y[..., 14:21] = oh
        if raw is not None:
            raw.copy_(x[..., :21].permute(0, 3, 1, 2))

    def stn_params(self, hm, amax, ratio, offset, vis_thresh, align_corners, centers, theta):
        hm = hm.f32
        B, S = hm.shape[0], hm.shape[1]
        h = hm[..., :24].permute(0, 3, 1, 2).reshape(B, 24, -1)
        p = F.softmax(10 * h, 2).reshape(B, 24, S, S)
        ar = torch.arange(S, dtype=torch.float32)
        cx = (p.sum(2) * ar).sum(2) / (0.5 * S) - 1
        cy = (p.sum(3) * ar).sum(2) / (0.5 * S) - 1
        c = torch.stack([cx, cy], -1)
        box = c.max(1)[0] - c.min(1)[0]
        scale_box = box.max(1)[0] / 2
        th = torch.zeros(B, 24, 3)
        for i in range(24):
            if i == 0:
                s = scale_box.clone()
            else:
                sc = (c[:, CHILDREN1[i]] - c[:, i]).norm(dim=1) / 2
                sp = (c[:, PARENTS0[i]] - c[:, i]).norm(dim=1) / 2
                s = 2 * torch.max(sc, sp)
            s = s * torch.relu(ratio[i]) + torch.relu(offset[i])
            if i != 0 and vis_thresh > 0:
                m = torch.zeros(B, 1, S, S)
                for pid in SMPL2DP[i]:
                    m = torch.max(m, (amax == pid).float().reshape(B, 1, S, S))
                score = F.grid_sample(m, c[:, i].reshape(B, 1, 1, 2), align_corners=bool(align_corners)).reshape(B)
                s = torch.where(score < vis_thresh, 0.8 * scale_box, s)
            th[:, i, 0] = s
            th[:, i, 1:] = c[:, i]
        centers.reshape(-1)[:B * 48].copy_(c.reshape(-1))
        theta.reshape(-1)[:B * 72].copy_(th.reshape(-1))

    def stn_sample(self, xd, theta, align_corners, crops, shape=None):
        xd, crops = xd.f32, crops.f32
        B, S, _, C = xd.shape
        th = theta.reshape(-1)[:B * 72].view(B, 24, 3)
        x = xd.permute(0, 3, 1, 2)
        outs = []
        for i in range(24):
            t = torch.zeros(B, 2, 3)
            t[:, 0, 0] = th[:, i, 0]; t[:, 1, 1] = th[:, i, 0]; t[:, :, 2] = th[:, i, 1:]
            grid = F.affine_grid(t, x.size(), align_corners=bool(align_corners))
            outs.append(F.grid_sample(x, grid, align_corners=bool(align_corners)))
        o = torch.stack(outs, 1).reshape(B * 24, C, S, S).permute(0, 2, 3, 1)
        crops.copy_(o)

    def gcn_head(self, gp, rot_feats, gpara, para):
        B = rot_feats.shape[0] // 24
        x = rot_feats.reshape(B, 24, 128)
        adj = gp["adj"]

        def layer(l, x, a):
            y = torch.matmul(torch.matmul(a, x), gp["W"][l]) + gp["b"][l]
            return torch.relu(y * gp["bn_scale"][l].reshape(1, 24, 1) + gp["bn_shift"][l].reshape(1, 24, 1))
        p0 = layer(0, x, adj[0])
        h = layer(3, layer(2, layer(1, p0, adj[1]), adj[1]), adj[1])
        r = layer(4, p0 + h, adj[2])
        p6 = (r.reshape(B, 24, 1, 128) * gp["head_w"].reshape(1, 24, 6, 128)).sum(-1).reshape(B, 144)
        p6 = p6 + gp["head_b"] + gp["mean_pose"]
        v = p6.reshape(-1, 3, 2)
        a1, a2 = v[:, :, 0], v[:, :, 1]
        b1 = F.normalize(a1)
        b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1)
        b3 = torch.cross(b1, b2, dim=1)
        R = torch.stack([b1, b2, b3], -1).reshape(B, 216)
        out = torch.cat([gpara.reshape(-1)[:B * 13].view(B, 13), R], 1)
        para.reshape(-1)[:B * 229].copy_(out.reshape(-1))
