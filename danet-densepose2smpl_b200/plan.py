"""Execution plan of the DaNet network half: folds BatchNorm into the convolutions, packs the
weights into the kernels' layouts, plans activation buffers (liveness-based reuse) and replays
the op list through the C ABI (optionally as one CUDA graph)."""
import ctypes

import numpy as np
import torch

from . import _lib
from . import netgraph as ng

BN_EPS = 1e-5


class CudaOps(object):
    """Thin tensor-level wrappers over libdanet_b200.so.  Fails loudly without the library / GPU."""

    def __init__(self, device):
        if not torch.cuda.is_available():
            raise RuntimeError("danet_b200: CUDA device required (there is no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device)

    @staticmethod
    def _desc(d):
        c = _lib.ConvDesc()
        for k in ("N", "H", "W", "Cin", "Cout", "ksize", "stride", "pad", "wsets", "relu"):
            setattr(c, k, int(d[k]))
        c.flags = int(d.get("flags", 0))
        return c

    supports_f16 = True      # tensor-core convs take / produce fp16 activation buffers (danet_conv_desc.flags)

    def conv_tc_supported(self, d):
        return bool(self.lib.danet_conv_tc_supported(ctypes.byref(self._desc(d))))

    def conv_tc_pack(self, d, w_simt):
        c = self._desc(d)
        nbytes = int(self.lib.danet_conv_tc_packed_bytes(ctypes.byref(c)))
        out = torch.empty(nbytes, dtype=torch.uint8, device=w_simt.device)
        _lib.check(self.lib.danet_conv_tc_pack(ctypes.byref(c), _lib.ptr(w_simt), _lib.ptr(out), _lib.stream_ptr()),
                   "conv_tc_pack")
        return out

    def conv2d(self, d, algo, x, w, bias, res, y):
        _lib.check(self.lib.danet_conv2d(ctypes.byref(self._desc(d)), algo, _lib.ptr(x), _lib.ptr(w), _lib.ptr(bias),
                                         _lib.ptr(res), _lib.ptr(y), _lib.stream_ptr()), "conv2d")

    def nchw_to_nhwc(self, x, y):
        N, C, H, W = x.shape
        _lib.check(self.lib.danet_nchw_to_nhwc(N, C, H * W, y.shape[-1], _lib.ptr(x), _lib.ptr(y), _lib.stream_ptr()),
                   "nchw_to_nhwc")

    def fuse_sum(self, terms, factors, relu, y):
        N, H, W, C = y.shape
        n = len(terms)
        arr = (ctypes.c_void_p * n)(*[t.data_ptr() for t in terms])
        fac = (ctypes.c_int32 * n)(*factors)
        _lib.check(self.lib.danet_fuse_sum(N, H, W, C, n, ctypes.cast(arr, ctypes.c_void_p),
                                           ctypes.cast(fac, ctypes.c_void_p), int(relu), _lib.ptr(y), _lib.stream_ptr()),
                   "fuse_sum")

    def maxpool(self, x, y):
        N, H, W, C = x.shape
        _lib.check(self.lib.danet_maxpool3x3s2(N, H, W, C, _lib.ptr(x), _lib.ptr(y), _lib.stream_ptr()), "maxpool")

    def avgpool(self, x, y):
        N, H, W, C = x.shape
        _lib.check(self.lib.danet_global_avgpool(N, H * W, C, _lib.ptr(x), _lib.ptr(y), _lib.stream_ptr()), "avgpool")

    def linear(self, x, w, b, add, y):
        N, In = x.shape[0], w.shape[1]
        _lib.check(self.lib.danet_linear(N, In, w.shape[0], _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(add),
                                         _lib.ptr(y), _lib.stream_ptr()), "linear")

    def clean_global(self, heads, body, amax, vis):
        B, H, W, Ch = heads.shape
        u, v, i, a = vis if vis is not None else (None, None, None, None)
        _lib.check(self.lib.danet_iuv_clean_global(B, H * W, Ch, 0, 25, 50, 75, body.shape[-1], _lib.ptr(heads),
                                                   _lib.ptr(body), _lib.ptr(amax), _lib.ptr(u), _lib.ptr(v), _lib.ptr(i),
                                                   _lib.ptr(a), _lib.stream_ptr()), "iuv_clean_global")

    def clean_parts(self, x, y, raw):
        N, H, W, Cx = x.shape
        _lib.check(self.lib.danet_iuv_clean_parts(N, H * W, Cx, y.shape[-1], _lib.ptr(x), _lib.ptr(y), _lib.ptr(raw),
                                                  int(y.dtype == torch.float16), _lib.stream_ptr()), "iuv_clean_parts")

    def stn_params(self, hm, amax, ratio, offset, vis_thresh, align_corners, centers, theta):
        B, S, _, Chm = hm.shape
        _lib.check(self.lib.danet_stn_params(B, S, Chm, _lib.ptr(hm), _lib.ptr(amax), _lib.ptr(ratio), _lib.ptr(offset),
                                             float(vis_thresh), int(align_corners), _lib.ptr(centers), _lib.ptr(theta),
                                             _lib.stream_ptr()), "stn_params")

    def stn_sample(self, xd, theta, align_corners, crops):
        B, S, _, C = xd.shape
        _lib.check(self.lib.danet_stn_sample(B, S, C, _lib.ptr(xd), _lib.ptr(theta), int(align_corners),
                                             _lib.ptr(crops), int(crops.dtype == torch.float16), _lib.stream_ptr()),
                   "stn_sample")

    def gcn_head(self, gp, rot_feats, gpara, para):
        B = para.shape[0]
        p = _lib.GcnParams()
        p.adj = gp["adj"].data_ptr()
        for l in range(5):
            p.W[l] = gp["W"][l].data_ptr(); p.b[l] = gp["b"][l].data_ptr()
            p.bn_scale[l] = gp["bn_scale"][l].data_ptr(); p.bn_shift[l] = gp["bn_shift"][l].data_ptr()
            p.dim_in[l], p.dim_out[l] = gp["W"][l].shape
        p.head_w = gp["head_w"].data_ptr(); p.head_b = gp["head_b"].data_ptr(); p.mean_pose = gp["mean_pose"].data_ptr()
        _lib.check(self.lib.danet_gcn_pose_head(B, ctypes.byref(p), _lib.ptr(rot_feats), _lib.ptr(gpara), _lib.ptr(para),
                                                _lib.stream_ptr()), "gcn_pose_head")


def fold_bn(sd, prefix, cout):
    """(scale, shift) of an eval-mode BatchNorm (running stats), fp32 like the reference computes."""
    if prefix is None:
        return torch.ones(cout), torch.zeros(cout)
    g, b = sd[prefix + ".weight"].float().cpu(), sd[prefix + ".bias"].float().cpu()
    m, v = sd[prefix + ".running_mean"].float().cpu(), sd[prefix + ".running_var"].float().cpu()
    scale = g / torch.sqrt(v + BN_EPS)
    return scale, b - m * scale


def pack_conv(sd, op):
    """Conv(+BN) parameters -> SIMT layout: w [wsets][k*k*Cin_p][Cout_p], bias [wsets][Cout_p]."""
    x, y, k, G = op["x"], op["y"], op["k"], op["groups"]
    cin, cin_p, cout_p = x.C, x.Cp, y.Cp
    ws, bs = [], []
    for (wkey, co, has_bias) in op["parts"]:
        w = sd[wkey + ".weight"].float().cpu()                        # [G*co, cin, k, k]
        b = sd[wkey + ".bias"].float().cpu() if has_bias else torch.zeros(G * co)
        ws.append(w.reshape(G, co, cin, k, k))
        bs.append(b.reshape(G, co))
    w = torch.cat(ws, dim=1)                                           # [G, ctot, cin, k, k]
    b = torch.cat(bs, dim=1)
    ctot = w.shape[1]
    scale, shift = fold_bn(sd, op["bn"], G * ctot)
    w = w * scale.reshape(G, ctot, 1, 1, 1)
    b = b * scale.reshape(G, ctot) + shift.reshape(G, ctot)
    wp = torch.zeros(G, k, k, cin_p, cout_p)
    wp[:, :, :, :cin, :ctot] = w.permute(0, 3, 4, 2, 1)
    bp = torch.zeros(G, cout_p)
    bp[:, :ctot] = b
    return wp.reshape(G, k * k * cin_p, cout_p).contiguous(), bp.contiguous()


def conv2x2_as_gemm(w, b, cin, cout):
    """A 3x3 / stride 1 / pad 1 convolution on 2x2-pixel maps touches every input pixel from every output
    pixel (|dy|,|dx| <= 1 always), so it IS a dense matrix product: x viewed as [images, 4*cin] times
    W' [4*cin, 4*cout] with W'[(iy,ix,ci), (oy,ox,co)] = w[ky=iy-oy+1, kx=ix-ox+1, ci, co]  (no padding taps:
    16 of the 36 (tap, pixel) pairs of the direct form multiply zeros).  NHWC memory of x / y / residual is
    already [image][(y,x,c)], so nothing moves; the images become the pixels of one 1x1 convolution.
    w [1][9*cin][cout] (SIMT layout), b [1][cout] -> (w' [1][4*cin][4*cout], b' [1][4*cout])."""
    w9 = w.reshape(3, 3, cin, cout)
    out = torch.zeros(2, 2, cin, 2, 2, cout, dtype=w.dtype)
    for iy in range(2):
        for ix in range(2):
            for oy in range(2):
                for ox in range(2):
                    out[iy, ix, :, oy, ox, :] = w9[iy - oy + 1, ix - ox + 1]
    return out.reshape(1, 4 * cin, 4 * cout).contiguous(), b.reshape(1, cout).repeat(1, 4).contiguous()


def refine_adjacency(sd, rp):
    """normalize_undigraph(I_n + A_mask * relu(edge_importance)) (smpl_regressor.py:870-871,
    utils/graph.py:232-261) -- parameter-only, so evaluated once per weight load."""
    I_n = sd[rp + "I_n"].float().cpu()[0]
    A = I_n + sd[rp + "A_mask"].float().cpu()[0] * torch.relu(sd[rp + "edge_importance"].float().cpu()[0])
    d = A.sum(0)
    dn = torch.zeros_like(d)
    dn[d > 0] = d[d > 0] ** (-0.5)
    return torch.matmul(torch.matmul(torch.diag(dn), A), torch.diag(dn))


def pack_gcn(sd, rp, device):
    layers = [("r2p_gcn", 0), ("refine_gcn", 0), ("refine_gcn", 1), ("refine_gcn", 2), ("p2r_gcn", 0)]
    gp = {"W": [], "b": [], "bn_scale": [], "bn_shift": []}
    for name, i in layers:
        gp["W"].append(sd["%s%s.gc.%d.weight" % (rp, name, i)].float().contiguous().to(device))
        gp["b"].append(sd["%s%s.gc.%d.bias" % (rp, name, i)].float().contiguous().to(device))
        s, t = fold_bn(sd, "%s%s.act.%d.0" % (rp, name, i), 24)
        gp["bn_scale"].append(s.contiguous().to(device)); gp["bn_shift"].append(t.contiguous().to(device))
    adj = torch.stack([sd[rp + "r2p_A"].float().cpu()[0], refine_adjacency(sd, rp), sd[rp + "p2r_A"].float().cpu()[0]])
    gp["adj"] = adj.contiguous().to(device)
    gp["head_w"] = sd[rp + "pose_regressors.1.1.weight"].float().reshape(144, 128).contiguous().to(device)
    gp["head_b"] = sd[rp + "pose_regressors.1.1.bias"].float().contiguous().to(device)
    gp["mean_pose"] = sd[rp + "mean_pose"].float().reshape(144).contiguous().to(device)
    return gp


class Plan(object):
    """Compiled forward for a fixed batch size B on one device."""

    RP = "iuv2smpl.smpl_para_Outs."

    def __init__(self, graph, state_dict, B, device, conv_algo="simt", align_corners=False, vis_thresh=0.5,
                 want_vis=True, ops=None, use_cuda_graph=False, f16_intermediates=True, gemm_2x2=False):
        self.g, self.B, self.device = graph, B, torch.device(device)
        self.ops = ops if ops is not None else CudaOps(device)
        self.align_corners, self.vis_thresh, self.want_vis = align_corners, vis_thresh, want_vis
        self.conv_algo = conv_algo
        self.gemm_2x2 = gemm_2x2          # DaNet passes True; validated by tests/test_kernels_gpu.py, test_net_gpu.py
        self.n_launch = 0
        self.n_tc = 0
        sd = state_dict
        dev = self.device
        self.steps = []
        self.f16 = self._f16_tensors() if f16_intermediates else set()
        self._plan_buffers()
        for op in graph.ops:
            kind = op["op"]
            if kind == "conv":
                self._add_conv(op, sd)
            elif kind == "input":
                self.steps.append(("input", op))
            elif kind == "fuse":
                self.steps.append(("fuse", op))
            elif kind in ("maxpool", "avgpool", "clean_global", "clean_parts", "stn_sample"):
                self.steps.append((kind, op))
            elif kind == "stn_params":
                self.ratio = sd["img2iuv.learned_ratio"].float().contiguous().to(dev)
                self.offset = sd["img2iuv.learned_offset"].float().contiguous().to(dev)
                self.steps.append((kind, op))
            elif kind == "body_fc":
                self.fc_w = sd[self.RP + "body_net.3.final_layer.weight"].float().contiguous().to(dev)
                self.fc_b = sd[self.RP + "body_net.3.final_layer.bias"].float().contiguous().to(dev)
                self.fc_add = sd[self.RP + "mean_cam_shape"].float().reshape(13).contiguous().to(dev)
                self.pooled = torch.empty(B, 512, device=dev)
                self.steps.append((kind, op))
            elif kind == "gcn_head":
                self.gcn = pack_gcn(sd, self.RP, dev)
                self.steps.append((kind, op))
            else:
                raise ValueError("unknown op %s" % kind)
        S = graph.outputs["heads"].H
        self.vis = None
        self.raw_parts = None
        if want_vis:
            self.vis = [torch.empty(B, c, S, S, device=dev) for c in (25, 25, 25, 15)]
            self.raw_parts = torch.empty(B * 24, 21, S, S, device=dev)
        self.graph_exec = None
        self.use_cuda_graph = use_cuda_graph
        self.static_in = None

    # tensors callers read after run() (infer_net, tests): never recycled, never fp16
    KEEP = ("para", "centers", "theta", "amax", "global_para", "rot_feats", "heads", "hm", "body_iuv")

    def _keep(self):
        return set(self.g.outputs[k].name for k in self.KEEP if k in self.g.outputs)

    # -- fp16 intermediates -------------------------------------------------------------------
    def _conv_desc(self, op):
        x, y = op["x"], op["y"]
        return dict(N=self.B * x.nmult, H=x.H, W=x.W, Cin=x.Cp, Cout=y.Cp, ksize=op["k"], stride=op["stride"],
                    pad=op["pad"], wsets=op["groups"], relu=int(op["relu"]))

    def _f16_tensors(self):
        """Tensors written by a tensor-core conv and read ONLY as the input of tensor-core convs are kept
        in fp16: that kernel rounds its activations to fp16 (RN) when it stages them, so the values the
        MMAs see are bit-identical and the tensor costs half the traffic.  Residuals, fuse terms, glue
        inputs and graph outputs stay fp32."""
        if self.conv_algo != "tc" or not getattr(self.ops, "supports_f16", False):
            return set()
        tc = {}
        for op in self.g.ops:
            if op["op"] == "conv":
                tc[id(op)] = self.ops.conv_tc_supported(self._conv_desc(op))
        produced_by_tc, bad = set(), set()
        for op in self.g.ops:
            if op["op"] == "conv":
                if tc[id(op)]:
                    produced_by_tc.add(op["y"].name)
                else:
                    bad.add(op["x"].name)
                if op["res"] is not None:
                    bad.add(op["res"].name)
            else:
                if op["op"] in ("stn_sample", "clean_parts"):
                    produced_by_tc.add(op["y"].name)       # these two glue kernels can write fp16 as well
                for key in ("x", "hm", "amax", "theta", "gpara"):
                    t = op.get(key)
                    if t is not None:
                        bad.add(t.name)
                for (t, _f) in op.get("terms", []):
                    bad.add(t.name)
        keep = self._keep()
        out = set()
        for name in produced_by_tc - bad - keep:
            t = self.g.tensors[name]
            if t.Cp % 8 == 0 and t.dtype == "f32":
                out.add(name)
        return out

    # -- buffers ------------------------------------------------------------------------------
    def _plan_buffers(self):
        g, B, dev = self.g, self.B, self.device
        last_use = {}
        produced = {}
        for idx, op in enumerate(g.ops):
            for key in ("x", "res", "hm", "amax", "theta", "gpara"):
                t = op.get(key)
                if t is not None:
                    last_use[t.name] = idx
            for (t, _f) in op.get("terms", []):
                last_use[t.name] = idx
            for key in ("y", "amax", "theta", "centers"):
                t = op.get(key)
                if t is not None and t.name not in produced:
                    produced[t.name] = idx
        keep = self._keep()
        free = {}
        self.buf = {}
        release_at = {}
        for name, idx in last_use.items():
            if name not in keep:
                release_at.setdefault(idx, []).append(name)
        order = sorted(produced.items(), key=lambda kv: kv[1])
        oi = 0
        for idx in range(len(g.ops)):
            while oi < len(order) and order[oi][1] == idx:
                name = order[oi][0]
                t = g.tensors[name]
                shape = (B * t.nmult, t.H, t.W, t.Cp)
                numel = int(np.prod(shape))
                if t.dtype == "u8":
                    self.buf[name] = torch.empty(shape[:3], dtype=torch.uint8, device=dev)
                elif name in self.f16:
                    pool = free.get(("h", numel))
                    self.buf[name] = pool.pop().view(shape) if pool else torch.empty(shape, dtype=torch.float16, device=dev)
                else:
                    pool = free.get(numel)
                    if pool and name not in keep:
                        self.buf[name] = pool.pop().view(shape)
                    else:
                        self.buf[name] = torch.empty(shape, device=dev)
                oi += 1
            for name in release_at.get(idx, []):
                t = self.buf.get(name)
                if t is not None and t.dtype == torch.float32:
                    free.setdefault(t.numel(), []).append(t)
                elif t is not None and t.dtype == torch.float16:
                    free.setdefault(("h", t.numel()), []).append(t)
        seen = {}
        for t in self.buf.values():
            seen[t.untyped_storage().data_ptr()] = t.untyped_storage().nbytes()
        self.bytes_alloc = sum(seen.values())

    def T(self, t):
        return self.buf[t.name]

    # -- conv ---------------------------------------------------------------------------------
    def _add_conv(self, op, sd):
        x, y = op["x"], op["y"]
        w, b = pack_conv(sd, op)
        dev = self.device
        d = self._conv_desc(op)
        d["flags"] = (1 if x.name in self.f16 else 0) | (2 if y.name in self.f16 else 0)
        if (self.gemm_2x2 and self.conv_algo == "tc" and d["ksize"] == 3 and d["stride"] == 1 and d["pad"] == 1 and
                d["H"] == 2 and d["W"] == 2 and d["wsets"] == 1 and d["N"] % 8 == 0 and d["N"] >= 32 and
                d["flags"] == 0):
            # the ResNet tail's 2x2-pixel layers as one dense product on the tensor-core path (see
            # conv2x2_as_gemm); the images become an (N/8) x 8 pixel map of a 1x1 convolution
            d2 = dict(N=1, H=d["N"] // 8, W=8, Cin=4 * d["Cin"], Cout=4 * d["Cout"], ksize=1, stride=1, pad=0,
                      wsets=1, relu=d["relu"], flags=0)
            if self.ops.conv_tc_supported(d2):
                w, b = conv2x2_as_gemm(w, b, d["Cin"], d["Cout"])
                d = d2
        w, b = w.to(dev), b.to(dev)
        algo = 0
        if d["flags"] and not self.ops.conv_tc_supported(d):
            raise RuntimeError("plan: fp16 tensor on a convolution the tensor-core path does not take: %r" % (d,))
        if self.conv_algo == "tc" and self.ops.conv_tc_supported(d):
            w = self.ops.conv_tc_pack(d, w)
            algo = 1
            self.n_tc += 1
        self.steps.append(("conv", dict(d=d, algo=algo, w=w, b=b, x=x, y=y, res=op["res"])))

    # -- run ----------------------------------------------------------------------------------
    def _run_steps(self, image):
        ops = self.ops
        n = 0
        for kind, op in self.steps:
            if kind == "conv":
                ops.conv2d(op["d"], op["algo"], self.T(op["x"]), op["w"], op["b"],
                           self.T(op["res"]) if op["res"] is not None else None, self.T(op["y"]))
            elif kind == "input":
                ops.nchw_to_nhwc(image, self.T(op["y"]))
            elif kind == "fuse":
                ops.fuse_sum([self.T(t) for t, _ in op["terms"]], [f for _, f in op["terms"]], op["relu"], self.T(op["y"]))
            elif kind == "maxpool":
                ops.maxpool(self.T(op["x"]), self.T(op["y"]))
            elif kind == "avgpool":
                ops.avgpool(self.T(op["x"]), self.T(op["y"]))
            elif kind == "clean_global":
                ops.clean_global(self.T(op["x"]), self.T(op["y"]), self.T(op["amax"]), self.vis)
            elif kind == "stn_params":
                ops.stn_params(self.T(op["hm"]), self.T(op["amax"]), self.ratio, self.offset, self.vis_thresh,
                               self.align_corners, self.T(op["centers"]), self.T(op["theta"]))
            elif kind == "stn_sample":
                ops.stn_sample(self.T(op["x"]), self.T(op["theta"]), self.align_corners, self.T(op["y"]))
            elif kind == "clean_parts":
                ops.clean_parts(self.T(op["x"]), self.T(op["y"]), self.raw_parts)
            elif kind == "body_fc":
                ops.avgpool(self.T(op["x"]), self.pooled)
                ops.linear(self.pooled, self.fc_w, self.fc_b, self.fc_add, self.T(op["y"]))
                n += 1
            elif kind == "gcn_head":
                ops.gcn_head(self.gcn, self.T(op["x"]), self.T(op["gpara"]), self.T(op["y"]))
            n += 1
        self.n_launch = n

    def run(self, image):
        """image [B,3,H,W] fp32 NCHW on the plan's device.  Results stay in the plan's buffers."""
        if image.shape[0] != self.B:
            raise ValueError("plan compiled for batch %d, got %d" % (self.B, image.shape[0]))
        image = image.detach().to(self.device, torch.float32).contiguous()
        if not self.use_cuda_graph:
            self._run_steps(image)
            return
        if self.graph_exec is None:
            self.static_in = image.clone()
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                self._run_steps(self.static_in)          # warm-up outside capture
            torch.cuda.current_stream(self.device).wait_stream(s)
            self.graph_exec = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_exec):
                self._run_steps(self.static_in)
        self.static_in.copy_(image, non_blocking=True)
        self.graph_exec.replay()

    def out(self, name):
        return self.T(self.g.outputs[name])
