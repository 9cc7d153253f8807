"""Graph description of the DaNet network half.

One description serves two consumers: (a) parameter registration under the reference's
state_dict key names (`img2iuv.iuv_est.*`, `iuv2smpl.smpl_para_Outs.*`),
(b) the execution plan that launches the CUDA kernels (plan.py).

Structure restated from the reference (file:line per builder):
  HRNet              models/module/hr_module.py:188-378 (+ HighResolutionModule :15-179)
  Bottleneck/Basic   models/module/res_module.py:27-97
  IUV_predict_layer  models/module/res_module.py:281-390
  SmplResNet         models/module/res_module.py:393-464 ; LimbResLayers :500-535
  DecomposedPredictor (gcn strategy) models/danet/smpl_regressor.py:398-674
Activations are NHWC; "images" of the limb branch are the (batch, part)-flattened axis
(nmult = 24) so that grouped convolutions become per-image weight sets.
"""
from collections import OrderedDict

import numpy as np

# HR_MODEL.EXTRA of configs/danet_default.yaml:87-140 (NUM_CHANNELS = width * 2**i)
HR_STAGES = (
    dict(modules=1, branches=2, blocks=4),
    dict(modules=4, branches=3, blocks=4),
    dict(modules=3, branches=4, blocks=4),
)


def pad8(c):
    """Channel padding of NHWC activations: fp16 planes move in 16-byte rows (TMA strides, 8-byte vector access)."""
    return (c + 7) // 8 * 8


class Tensor(object):
    __slots__ = ("name", "nmult", "H", "W", "C", "Cp", "dtype")

    def __init__(self, name, nmult, H, W, C, dtype="f32", raw=False):
        self.name, self.nmult, self.H, self.W, self.C, self.dtype = name, nmult, H, W, C, dtype
        self.Cp = C if raw else pad8(C)      # raw: small per-sample vectors, no channel padding

    def __repr__(self):
        return "T(%s %dx[%d,%d,%d])" % (self.name, self.nmult, self.H, self.W, self.C)


class ParamSpec(object):
    """A parameter or buffer of the module tree: key, shape, kind ('param'|'buffer'), init."""
    __slots__ = ("key", "shape", "kind", "init")

    def __init__(self, key, shape, kind="param", init="conv"):
        self.key, self.shape, self.kind, self.init = key, tuple(shape), kind, init


class Graph(object):
    def __init__(self):
        self.ops = []
        self.params = OrderedDict()
        self.tensors = OrderedDict()
        self._n = 0

    # -- tensors / params ---------------------------------------------------------------------
    def tensor(self, nmult, H, W, C, name=None, dtype="f32", raw=False):
        self._n += 1
        t = Tensor(name or ("t%d" % self._n), nmult, H, W, C, dtype, raw)
        self.tensors[t.name] = t
        return t

    def param(self, key, shape, kind="param", init="conv"):
        if key not in self.params:
            self.params[key] = ParamSpec(key, shape, kind, init)
        return key

    def bn(self, prefix, c, dims=2):
        self.param(prefix + ".weight", (c,), init="ones")
        self.param(prefix + ".bias", (c,), init="zeros")
        self.param(prefix + ".running_mean", (c,), "buffer", "zeros")
        self.param(prefix + ".running_var", (c,), "buffer", "ones")
        self.param(prefix + ".num_batches_tracked", (), "buffer", "long0")
        return prefix

    # -- ops ----------------------------------------------------------------------------------
    def conv(self, x, wkey, cout, k, stride=1, bn=None, bias=False, relu=False, res=None, groups=1,
             extra_w=()):
        """nn.Conv2d(+BN folded)(+residual)(+ReLU).  groups>1 requires x.nmult == groups images per
        sample (weight set = image index % groups).  extra_w: more (wkey, cout, bias) convs on the
        same input whose outputs are concatenated along channels (prediction heads)."""
        pad = k // 2
        Ho, Wo = (x.H + 2 * pad - k) // stride + 1, (x.W + 2 * pad - k) // stride + 1
        cin = x.C
        parts = [(wkey, cout, bias)] + list(extra_w)
        for (wk, co, bs) in parts:
            self.param(wk + ".weight", (co * groups, cin, k, k), init="conv")
            if bs:
                self.param(wk + ".bias", (co * groups,), init="conv_bias:%d" % (cin * k * k))
        if bn:
            self.bn(bn, cout * groups)
        ctot = sum(p[1] for p in parts)
        y = self.tensor(x.nmult, Ho, Wo, ctot)
        self.ops.append(dict(op="conv", x=x, y=y, parts=parts, k=k, stride=stride, pad=pad, bn=bn, relu=relu,
                             res=res, groups=groups))
        return y

    def fuse(self, terms, relu=True):
        """terms: list of (tensor, upsample_factor); output has the shape of the factor-1 resolution."""
        t0, f0 = terms[0]
        H, W = t0.H * f0, t0.W * f0
        y = self.tensor(t0.nmult, H, W, t0.C)
        self.ops.append(dict(op="fuse", terms=terms, y=y, relu=relu))
        return y

    def simple(self, op, x, y, **kw):
        d = dict(op=op, x=x, y=y)
        d.update(kw)
        self.ops.append(d)
        return y


# ---------------------------------------------------------------------------------------------
# residual blocks
# ---------------------------------------------------------------------------------------------
def basic_block(g, x, prefix, planes, stride=1, groups=1):
    """res_module.py:27-56: conv3x3-bn-relu, conv3x3-bn, (+downsample), add, relu."""
    cin = x.C
    res = x
    if stride != 1 or cin != planes:
        res = g.conv(x, prefix + ".downsample.0", planes, 1, stride, bn=prefix + ".downsample.1", groups=groups)
    y = g.conv(x, prefix + ".conv1", planes, 3, stride, bn=prefix + ".bn1", relu=True, groups=groups)
    return g.conv(y, prefix + ".conv2", planes, 3, 1, bn=prefix + ".bn2", relu=True, res=res, groups=groups)


def bottleneck(g, x, prefix, planes):
    """res_module.py:59-97 (expansion 4, stride 1)."""
    cin = x.C
    res = x
    if cin != planes * 4:
        res = g.conv(x, prefix + ".downsample.0", planes * 4, 1, 1, bn=prefix + ".downsample.1")
    y = g.conv(x, prefix + ".conv1", planes, 1, 1, bn=prefix + ".bn1", relu=True)
    y = g.conv(y, prefix + ".conv2", planes, 3, 1, bn=prefix + ".bn2", relu=True)
    return g.conv(y, prefix + ".conv3", planes * 4, 1, 1, bn=prefix + ".bn3", relu=True, res=res)


# ---------------------------------------------------------------------------------------------
# HRNet backbone + IUV heads
# ---------------------------------------------------------------------------------------------
def hr_module(g, xs, prefix, chans, multi_scale_output):
    """One HighResolutionModule (hr_module.py:15-179): per-branch 4 BasicBlocks, then fuse."""
    nb = len(xs)
    for b in range(nb):
        for i in range(4):
            xs[b] = basic_block(g, xs[b], "%s.branches.%d.%d" % (prefix, b, i), chans[b])
    outs = []
    for i in range(nb if multi_scale_output else 1):
        terms = []
        for j in range(nb):
            fp = "%s.fuse_layers.%d.%d" % (prefix, i, j)
            if j == i:
                terms.append((xs[j], 1))
            elif j > i:      # 1x1 conv + BN at low resolution, nearest upsample by 2**(j-i)
                t = g.conv(xs[j], fp + ".0", chans[i], 1, 1, bn=fp + ".1")
                terms.append((t, 2 ** (j - i)))
            else:            # chain of stride-2 3x3 convs; ReLU on all but the last
                t = xs[j]
                for k in range(i - j):
                    last = k == i - j - 1
                    t = g.conv(t, "%s.%d.0" % (fp, k), chans[i] if last else chans[j], 3, 2,
                               bn="%s.%d.1" % (fp, k), relu=not last)
                terms.append((t, 1))
        outs.append(g.fuse(terms, relu=True))
    return outs


def hrnet(g, img, width, prefix="img2iuv.iuv_est."):
    """PoseHighResolutionNet.forward (hr_module.py:334-378) up to final_feat `xd`."""
    x = g.conv(img, prefix + "conv1", 64, 3, 2, bn=prefix + "bn1", relu=True)
    x = g.conv(x, prefix + "conv2", 64, 3, 2, bn=prefix + "bn2", relu=True)
    for i in range(4):
        x = bottleneck(g, x, "%slayer1.%d" % (prefix, i), 64)
    ys = [x]
    pre = [256]
    for s, st in enumerate(HR_STAGES):
        chans = [width * 2 ** i for i in range(st["branches"])]
        tp = "%stransition%d" % (prefix, s + 1)
        xs = []
        for i in range(st["branches"]):
            if i < len(pre):
                if chans[i] != pre[i]:
                    xs.append(g.conv(ys[i], "%s.%d.0" % (tp, i), chans[i], 3, 1, bn="%s.%d.1" % (tp, i), relu=True))
                else:
                    xs.append(ys[i])
            else:                                  # new lower-resolution branch from the last one
                t = ys[-1]
                for j in range(i + 1 - len(pre)):
                    last = j == i - len(pre)
                    t = g.conv(t, "%s.%d.%d.0" % (tp, i, j), chans[i] if last else pre[-1], 3, 2,
                               bn="%s.%d.%d.1" % (tp, i, j), relu=True)
                xs.append(t)
        for m in range(st["modules"]):
            multi = not (s == len(HR_STAGES) - 1 and m == st["modules"] - 1)
            xs = hr_module(g, xs, "%sstage%d.%d" % (prefix, s + 2, m), chans, multi)
        ys, pre = xs, chans
    return ys[0]


def iuv_heads(g, xd, width, prefix="img2iuv.iuv_est.final_pred."):
    """IUV_predict_layer.forward (res_module.py:375-390): U,V,Index,Ann heads as one conv with
    concatenated output channels [U 25 | V 25 | Index 25 | Ann 15], and the heat-map branch."""
    heads = g.conv(xd, prefix + "predict_u", 25, 3, 1, bias=True,
                   extra_w=[(prefix + "predict_v", 25, True), (prefix + "predict_uv_index", 25, True),
                            (prefix + "predict_ann_index", 15, True)])
    h = xd
    for i in range(3):
        h = bottleneck(g, h, "%spredict_hm.0.%d" % (prefix, i), width // 4)
    hm = g.conv(h, prefix + "predict_hm.1", 24, 3, 1, bias=True)
    return heads, hm


# ---------------------------------------------------------------------------------------------
# ResNet-18 regressors
# ---------------------------------------------------------------------------------------------
def smpl_resnet18(g, x, prefix, truncate):
    """SmplResNet.forward (res_module.py:444-464), BasicBlock x [2,2,2,2]."""
    x = g.conv(x, prefix + "conv1", 64, 7, 2, bn=prefix + "bn1", relu=True)
    y = g.tensor(x.nmult, (x.H - 1) // 2 + 1, (x.W - 1) // 2 + 1, x.C)
    x = g.simple("maxpool", x, y)
    for li, (planes, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2))):
        if truncate >= 1 and li == 3:
            break
        for b in range(2):
            x = basic_block(g, x, "%slayer%d.%d" % (prefix, li + 1, b), planes, stride if b == 0 else 1)
    return x


def danet_graph(width=48, img_size=224):
    """The whole `DaNet.infer_net` network half (danet.py:61-131) for INPUT_MODE='iuv',
    DECOMPOSED=True, REFINE_STRATEGY='gcn'."""
    g = Graph()
    S = img_size // 4
    img = g.tensor(1, img_size, img_size, 3, name="image")
    g.ops.append(dict(op="input", y=img))
    xd = hrnet(g, img, width)
    heads, hm = iuv_heads(g, xd, width)
    g.param("img2iuv.learned_ratio", (24,), "buffer", "learned_ratio")
    g.param("img2iuv.learned_offset", (24,), "buffer", "learned_offset")
    # iuvmap_clean of the global maps -> body_iuv [B,S,S,75] + argmax map
    body_iuv = g.tensor(1, S, S, 75, name="body_iuv")
    amax = g.tensor(1, S, S, 1, name="index_argmax", dtype="u8", raw=True)
    g.ops.append(dict(op="clean_global", x=heads, y=body_iuv, amax=amax))
    # STN: centres, visibility, thetas -> 24 crops of xd
    theta = g.tensor(1, 1, 24, 3, name="theta", raw=True)
    centers = g.tensor(1, 1, 24, 2, name="stn_centers", raw=True)
    g.ops.append(dict(op="stn_params", hm=hm, amax=amax, theta=theta, centers=centers))
    crops = g.tensor(24, S, S, xd.C, name="part_crops")
    g.ops.append(dict(op="stn_sample", x=xd, theta=theta, y=crops))
    fp = "img2iuv.iuv_est.final_pred."
    part_pred = g.conv(crops, fp + "predict_partial_iuv", 21, 3, 1, bias=True, groups=24)
    part_iuv = g.tensor(24, S, S, 21, name="part_iuv_clean")
    g.ops.append(dict(op="clean_parts", x=part_pred, y=part_iuv))
    # SMPL regressor, global branch
    rp = "iuv2smpl.smpl_para_Outs."
    b = g.conv(body_iuv, rp + "body_net.0", 64, 1, 1, bn=rp + "body_net.1", relu=True)
    b = smpl_resnet18(g, b, rp + "body_net.3.", truncate=0)
    g.param(rp + "body_net.3.final_layer.weight", (13, 512), init="linear")
    g.param(rp + "body_net.3.final_layer.bias", (13,), init="conv_bias:512")
    g.param(rp + "mean_cam_shape", (1, 13), "buffer", "mean_cam_shape")
    g.param(rp + "mean_pose", (1, 144), "buffer", "mean_pose")
    gpara = g.tensor(1, 1, 1, 13, name="global_para", raw=True)
    g.ops.append(dict(op="body_fc", x=b, y=gpara))
    # limb branch: (batch,part)-flattened images
    l = g.conv(part_iuv, rp + "limb_net.0", 64, 1, 1, bn=rp + "limb_net.1", relu=True)
    l = smpl_resnet18(g, l, rp + "limb_net.3.", truncate=1)
    for bi in range(2):
        l = basic_block(g, l, "%slimb_reslayer.layer4.%d" % (rp, bi), 128, 2 if bi == 0 else 1, groups=24)
    rot_feats = g.tensor(24, 1, 1, 128, name="rot_feats")
    g.simple("avgpool", l, rot_feats)
    # GCN refinement + pose head (parameters registered here; kernel consumes them pre-packed)
    for name, dims in (("r2p_gcn", [(128, 128)]), ("refine_gcn", [(128, 256), (256, 256), (256, 128)]),
                       ("p2r_gcn", [(128, 128)])):
        for i, (di, do) in enumerate(dims):
            g.param("%s%s.gc.%d.weight" % (rp, name, i), (di, do), init="xavier_relu")
            g.param("%s%s.gc.%d.bias" % (rp, name, i), (do,), init="zeros")
            g.bn("%s%s.act.%d.0" % (rp, name, i), 24, dims=1)
    g.param(rp + "edge_importance", (1, 24, 24), init="ones")
    for k in ("I_n", "A_link", "A", "A_mask", "r2p_A", "p2r_A"):
        g.param(rp + k, (1, 24, 24), "buffer", "graph:" + k)
    for i in range(2):
        g.param("%spose_regressors.%d.1.weight" % (rp, i), (144, 128, 1, 1), init="xavier_small")
        g.param("%spose_regressors.%d.1.bias" % (rp, i), (144,), init="conv_bias:128")
        g.param("%scoord_regressors.%d.1.weight" % (rp, i), (72, 128, 1, 1), init="conv")
        g.param("%scoord_regressors.%d.1.bias" % (rp, i), (72,), init="conv_bias:128")
    # constructed by the reference (smpl_regressor.py:583-600) but unused by the 'gcn' forward;
    # registered so that checkpoints load without unexpected keys
    for i in range(24):
        for (idx, shp, kind) in ((0, (512, 256, 1, 1), "c"), (1, 512, "bn"), (3, (128, 512, 1, 1), "c"), (4, 128, "bn")):
            key = "%srot2pos.%d.%d" % (rp, i, idx)
            if kind == "c":
                g.param(key + ".weight", shp, init="conv")
                g.param(key + ".bias", (shp[0],), init="conv_bias:%d" % shp[1])
            else:
                g.bn(key, shp)
    for (idx, shp, kind) in ((0, (1024, 384, 1, 1), "c"), (1, 1024, "bn"), (3, (128, 1024, 1, 1), "c"), (4, 128, "bn")):
        key = "%spos2rot.%d" % (rp, idx)
        if kind == "c":
            g.param(key + ".weight", shp, init="conv")
            g.param(key + ".bias", (shp[0],), init="conv_bias:%d" % shp[1])
        else:
            g.bn(key, shp)
    para = g.tensor(1, 1, 1, 229, name="para", raw=True)
    g.ops.append(dict(op="gcn_head", x=rot_feats, gpara=gpara, y=para))
    g.outputs = dict(para=para, heads=heads, hm=hm, xd=xd, body_iuv=body_iuv, part_pred=part_pred,
                     part_iuv=part_iuv, centers=centers, theta=theta, amax=amax, rot_feats=rot_feats,
                     global_para=gpara)
    return g


# ---------------------------------------------------------------------------------------------
# SMPL skeleton graphs (utils/graph.py:74-106,110-158 via smpl_regressor.py:626-672)
# ---------------------------------------------------------------------------------------------
LIMB_PAIRS = [(0, 1), (1, 4), (4, 7), (7, 10), (0, 2), (2, 5), (5, 8), (8, 11), (0, 3), (3, 6), (6, 9),
              (9, 13), (13, 16), (16, 18), (18, 20), (20, 22), (9, 14), (14, 17), (17, 19), (19, 21), (21, 23),
              (9, 12), (12, 15)]
NEIGH2_EXTRA = [(12, 17), (12, 16)]
NEIGH2_LINKS = [(0, 4), (0, 5), (0, 6), (2, 8), (1, 7), (5, 11), (4, 10), (3, 9), (6, 12), (9, 15),
                (6, 13), (9, 16), (13, 18), (16, 20), (18, 22), (6, 14), (9, 17), (14, 19), (17, 21), (19, 23)]
ADD_LINKS = [(1, 2), (1, 3), (2, 3), (13, 14), (12, 13), (12, 14)]
SMPL_PARENTS0 = [0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def _sym(pairs, n=24):
    A = np.eye(n)
    for i, j in pairs:
        A[i, j] = A[j, i] = 1
    return A


def _row_normalize(A):
    d = A.sum(1)
    out = A.copy()
    nz = d > 0
    out[nz] = A[nz] / d[nz, None]
    return out


def undigraph_normalize(A):
    """utils/graph.py:232-261: D^-1/2 A D^-1/2 with D = column sums (zeros stay zero)."""
    d = A.sum(0)
    dn = np.zeros_like(d)
    dn[d > 0] = d[d > 0] ** (-0.5)
    return (dn[:, None] * A) * dn[None, :]


def graph_buffers():
    """Structural buffers DecomposedPredictor registers (smpl_regressor.py:626-672)."""
    I = np.eye(24)
    A_link = _sym(LIMB_PAIRS) - I
    A_mask_full = _sym(LIMB_PAIRS + NEIGH2_EXTRA + NEIGH2_LINKS + ADD_LINKS)
    A = undigraph_normalize(A_mask_full)
    chains = []
    for i in range(24):
        c, p = [i], i
        while p != 0:
            p = SMPL_PARENTS0[p]
            c.append(p)
        chains.append(c)
    r2p = np.zeros((24, 24))
    p2r = np.zeros((24, 24))
    for i in range(24):
        r2p[i, chains[i]] = 1
        r2p[i, i] = 0
        kids = [k for k, v in enumerate(SMPL_PARENTS0) if v == i]
        p2r[i, kids] = 1
        p2r[i, SMPL_PARENTS0[i]] = 1
        p2r[i, i] = 1
    f = lambda a: a[None].astype(np.float32)
    return {"I_n": f(I), "A_link": f(A_link), "A": f(A), "A_mask": f(A_mask_full - I),
            "r2p_A": f(_row_normalize(r2p)), "p2r_A": f(_row_normalize(p2r))}
