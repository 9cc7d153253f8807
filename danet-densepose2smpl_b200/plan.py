"""Execution plan of the DaNet network half: folds BatchNorm into the convolutions, packs the
weights into the kernels' layouts, schedules the graph into launch steps (independent convolutions --
HRNet's parallel branches, the body / limb regressors -- share one tensor-core launch), plans
activation buffers (liveness-based reuse) and replays the steps through the C ABI (optionally as one
CUDA graph)."""
import ctypes

import numpy as np
import torch

from . import _lib
from . import netgraph as ng

BN_EPS = 1e-5
MAX_GROUP = 6            # problems per tensor-core launch (csrc/conv_tc.cu kMaxProb)


class ActBuf(object):
    """One activation tensor of the plan (include/danet_b200.h danet_act): an fp32 NHWC view and/or
    split-fp16 planes h[P,N,H,W,C] (P = 2: hi + lo, exact mode; P = 1: hi only, fast mode)."""
    __slots__ = ("f32", "h")

    def __init__(self, f32=None, h=None):
        self.f32, self.h = f32, h

    def c(self):
        a = _lib.Act()
        a.f32 = self.f32.data_ptr() if self.f32 is not None else None
        a.hi = self.h[0].data_ptr() if self.h is not None else None
        a.lo = self.h[1].data_ptr() if (self.h is not None and self.h.shape[0] > 1) else None
        return a

    def value(self):
        """fp32 torch tensor of the activation (tests / debugging)."""
        if self.f32 is not None:
            return self.f32
        v = self.h[0].float()
        return v + self.h[1].float() if self.h.shape[0] > 1 else v


def _null_act():
    return _lib.Act(None, None, None)


class CudaOps(object):
    """Thin tensor-level wrappers over libdanet_b200.so.  Fails loudly without the library / GPU."""

    def __init__(self, device):
        if not torch.cuda.is_available():
            raise RuntimeError("danet_b200: CUDA device required (there is no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device)

    def planes(self, precision):
        """fp16 planes per activation on the tensor-core path (0 would mean: fp32 buffers only)."""
        return 2 if precision == "exact" else 1

    def _sp(self):
        return _lib.stream_ptr(self.device)

    @staticmethod
    def _desc(d):
        c = _lib.ConvDesc()
        for k in ("N", "H", "W", "Cin", "Cout", "ksize", "stride", "pad", "wsets", "relu"):
            setattr(c, k, int(d[k]))
        c.flags = int(d.get("flags", 0))
        return c

    def tc_capable(self):
        """The tcgen05 kernels are sm_100a code: compute capability 10.x only."""
        return torch.cuda.get_device_capability(self.device)[0] == 10

    def conv_tc_supported(self, d):
        return bool(self.lib.danet_conv_tc_supported(ctypes.byref(self._desc(d))))

    def conv_tc_pack(self, d, w_simt):
        c = self._desc(d)
        nbytes = int(self.lib.danet_conv_tc_packed_bytes(ctypes.byref(c)))
        if nbytes <= 0:
            raise RuntimeError("danet_b200: convolution %r is not supported by the tensor-core path" % (d,))
        out = torch.empty(nbytes, dtype=torch.uint8, device=w_simt.device)
        _lib.check(self.lib.danet_conv_tc_pack(ctypes.byref(c), _lib.ptr(w_simt), _lib.ptr(out), self._sp()), "conv_tc_pack")
        return out

    def conv_config(self, descs):
        """(sub-tiles per problem, (activation stages, weight stages)) of a would-be launch over `descs`."""
        n = len(descs)
        arr = (_lib.ConvDesc * n)(*[self._desc(d) for d in descs])
        S = (ctypes.c_int32 * n)()
        st = (ctypes.c_int32 * 2)()
        _lib.check(self.lib.danet_conv_tc_config(n, arr, ctypes.cast(S, ctypes.c_void_p), ctypes.cast(st, ctypes.c_void_p)), "conv_tc_config")
        return list(S), (st[0], st[1])

    def conv_group(self, convs):
        """convs: list of dict(d, x, res, y (ActBuf), w (packed), b)."""
        arr = (_lib.ConvProblem * len(convs))()
        for i, cv in enumerate(convs):
            p = arr[i]
            p.d = self._desc(cv["d"])
            p.x = cv["x"].c()
            p.res = cv["res"].c() if cv["res"] is not None else _null_act()
            p.y = cv["y"].c()
            p.w_packed = cv["w"].data_ptr()
            p.bias = cv["b"].data_ptr() if cv["b"] is not None else None
        _lib.check(self.lib.danet_conv_tc_group(len(convs), arr, self._sp()), "conv_tc_group")

    def conv2d(self, d, x, w, bias, res, y):
        """fp32 FMA convolution on the fp32 views."""
        _lib.check(self.lib.danet_conv2d(ctypes.byref(self._desc(d)), 0, _lib.ptr(x.f32), _lib.ptr(w), _lib.ptr(bias),
                                         _lib.ptr(res.f32 if res is not None else None), _lib.ptr(y.f32), self._sp()), "conv2d")

    def nchw_to_nhwc(self, x, y):
        N, C, H, W = x.shape
        t = y.f32 if y.f32 is not None else y.h[0]
        _lib.check(self.lib.danet_nchw_to_nhwc(N, C, H * W, t.shape[-1], _lib.ptr(x), ctypes.byref(y.c()), self._sp()),
                   "nchw_to_nhwc")

    def fuse_sum(self, terms, factors, relu, y, shape):
        N, H, W, C = shape
        n = len(terms)
        arr = (_lib.Act * n)(*[t.c() for t in terms])
        fac = (ctypes.c_int32 * n)(*factors)
        _lib.check(self.lib.danet_fuse_sum(N, H, W, C, n, arr, ctypes.cast(fac, ctypes.c_void_p), int(relu),
                                           ctypes.byref(y.c()), self._sp()), "fuse_sum")

    def maxpool(self, x, y, shape):
        N, H, W, C = shape
        _lib.check(self.lib.danet_maxpool3x3s2(N, H, W, C, ctypes.byref(x.c()), ctypes.byref(y.c()), self._sp()), "maxpool")

    def avgpool(self, x, y, shape):
        N, H, W, C = shape
        _lib.check(self.lib.danet_global_avgpool(N, H * W, C, ctypes.byref(x.c()), _lib.ptr(y), self._sp()), "avgpool")

    def linear(self, x, w, b, add, y):
        N, In = x.shape[0], w.shape[1]
        _lib.check(self.lib.danet_linear(N, In, w.shape[0], _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(add),
                                         _lib.ptr(y), self._sp()), "linear")

    def clean_global(self, heads, body, amax, vis, shape):
        B, H, W, Ch, Cb = shape
        u, v, i, a = vis if vis is not None else (None, None, None, None)
        _lib.check(self.lib.danet_iuv_clean_global(B, H * W, Ch, 0, 25, 50, 75, Cb, _lib.ptr(heads.f32),
                                                   ctypes.byref(body.c()), _lib.ptr(amax), _lib.ptr(u), _lib.ptr(v), _lib.ptr(i),
                                                   _lib.ptr(a), self._sp()), "iuv_clean_global")

    def clean_parts(self, x, y, raw, shape):
        N, H, W, Cx, Cy = shape
        _lib.check(self.lib.danet_iuv_clean_parts(N, H * W, Cx, Cy, _lib.ptr(x.f32), ctypes.byref(y.c()), _lib.ptr(raw),
                                                  self._sp()), "iuv_clean_parts")

    def stn_params(self, hm, amax, ratio, offset, vis_thresh, align_corners, centers, theta):
        B, S, _, Chm = hm.f32.shape
        _lib.check(self.lib.danet_stn_params(B, S, Chm, _lib.ptr(hm.f32), _lib.ptr(amax), _lib.ptr(ratio), _lib.ptr(offset),
                                             float(vis_thresh), int(align_corners), _lib.ptr(centers), _lib.ptr(theta),
                                             self._sp()), "stn_params")

    def stn_sample(self, xd, theta, align_corners, crops, shape):
        B, S, C = shape
        _lib.check(self.lib.danet_stn_sample(B, S, C, ctypes.byref(xd.c()), _lib.ptr(theta), int(align_corners),
                                             ctypes.byref(crops.c()), self._sp()), "stn_sample")

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
                                                self._sp()), "gcn_pose_head")


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


# which views of its input tensors an op needs on the tensor-core path: "h" = fp16 planes, "f" = fp32
_F32_INPUTS = {("clean_global", "x"), ("clean_parts", "x"), ("stn_params", "hm"), ("gcn_head", "x"), ("gcn_head", "gpara"),
               ("stn_sample", "theta")}


def op_inputs(op):
    """(tensor, role) pairs an op reads (`amax` / `theta` are outputs of the op that makes them, inputs of the next)."""
    out = []
    keys = ["x", "res", "hm", "gpara"]
    if op["op"] == "stn_params":
        keys.append("amax")
    if op["op"] == "stn_sample":
        keys.append("theta")
    for key in keys:
        t = op.get(key)
        if t is not None:
            out.append((t, key))
    for (t, _f) in op.get("terms", []):
        out.append((t, "term"))
    return out


def op_outputs(op):
    out = []
    if op.get("y") is not None:
        out.append(op["y"])
    if op["op"] == "clean_global":
        out.append(op["amax"])
    if op["op"] == "stn_params":
        out += [op["theta"], op["centers"]]
    return out


class Plan(object):
    """Compiled forward for a fixed batch size B on one device."""

    RP = "iuv2smpl.smpl_para_Outs."

    def __init__(self, graph, state_dict, B, device, conv_algo="simt", precision="exact", align_corners=False,
                 vis_thresh=0.5, want_vis=True, ops=None, use_cuda_graph=False, group_convs=True, wcache=None,
                 keep_all=False):
        self.g, self.B, self.device = graph, B, torch.device(device)
        self.ops = ops if ops is not None else CudaOps(device)
        self.align_corners, self.vis_thresh, self.want_vis = align_corners, vis_thresh, want_vis
        if conv_algo not in ("simt", "tc"):
            raise ValueError("conv_algo must be 'simt' or 'tc'")
        if precision not in ("exact", "fast"):
            raise ValueError("precision must be 'exact' or 'fast'")
        self.conv_algo, self.precision = conv_algo, precision
        self.tc = conv_algo == "tc"
        if self.tc and hasattr(self.ops, "tc_capable") and not self.ops.tc_capable():
            raise RuntimeError("danet_b200: the tensor-core path is sm_100a code; device %s is not compute capability 10.x" % device)
        self.P = self.ops.planes(precision) if self.tc else 0          # fp16 planes per activation (0: fp32 buffers only)
        self.group_convs = group_convs and self.tc
        self.keep_all = keep_all                      # debugging: no buffer reuse, every intermediate stays readable
        self.wcache = wcache if wcache is not None else {}
        self.n_launch = 0
        self.n_tc = 0
        sd = state_dict
        dev = self.device
        with self._guard():
            self._schedule()
            self._formats()
            self._plan_buffers()
            self.steps = []
            for level_ops in self.schedule:
                group = []
                for op in level_ops:
                    kind = op["op"]
                    if kind == "conv":
                        cv = self._make_conv(op, sd)
                        if self.tc:
                            group.append(cv)
                        else:
                            self.steps.append(("conv_simt", cv))
                    else:
                        self._add_glue(op, sd)
                # tensor-core convolutions of one level: independent by construction, <= MAX_GROUP per launch,
                # most expensive tiles first (they start first inside the persistent grid)
                group.sort(key=lambda c: -c["cost"])
                for g in self._form_groups(group):
                    self.steps.append(("conv_group", g))
            S = graph.outputs["heads"].H
            self.vis = None
            self.raw_parts = None
            if want_vis:
                self.vis = [torch.empty(B, c, S, S, device=dev) for c in (25, 25, 25, 15)]
                self.raw_parts = torch.empty(B * 24, 21, S, S, device=dev)
        self.graph_exec = None
        self.use_cuda_graph = use_cuda_graph
        self.static_in = None

    def _form_groups(self, convs):
        """Partition one level's convolutions (cost-descending) into launches of <= MAX_GROUP.  The shared-memory rings
        of a launch are sized for its largest member, so a convolution joins a launch only if no member that carries
        a noticeable share of the launch's work loses its sub-tile pair (S = 2 -> 1) by it."""
        if not self.group_convs:
            return [[c] for c in convs]
        if not hasattr(self.ops, "conv_config"):
            return [convs[i:i + MAX_GROUP] for i in range(0, len(convs), MAX_GROUP)]
        solo = {id(c): self.ops.conv_config([c["d"]])[0][0] for c in convs}
        groups = []
        for c in convs:
            placed = False
            for g in groups:
                if len(g) >= MAX_GROUP:
                    continue
                S, _st = self.ops.conv_config([m["d"] for m in g] + [c["d"]])
                tot = sum(m["cost"] for m in g) + c["cost"]
                ok = all(s >= solo[id(m)] or m["cost"] < 0.05 * tot for m, s in zip(g + [c], S))
                if ok:
                    g.append(c)
                    placed = True
                    break
            if not placed:
                groups.append([c])
        return groups

    def _guard(self):
        """Kernels, buffers and the stream handle must belong to the plan's device, whatever the caller's
        current device is."""
        if self.device.type == "cuda":
            return torch.cuda.device(self.device)
        import contextlib
        return contextlib.nullcontext()

    # tensors callers read after run() (infer_net, tests): never recycled, always with an fp32 view
    KEEP = ("para", "centers", "theta", "amax", "global_para", "rot_feats", "heads", "hm", "body_iuv")

    def _keep(self):
        return set(self.g.outputs[k].name for k in self.KEEP if k in self.g.outputs)

    # -- scheduling ---------------------------------------------------------------------------
    def _schedule(self):
        """ASAP levels of the op DAG: every op of a level depends only on earlier levels, so the convolutions
        of a level (HRNet's parallel branches, hr_module.py:165-166; the fuse layers' 1x1 and stride-2
        convolutions; the body and limb regressors) may share one launch."""
        level_of_tensor = {}
        levels = []
        for op in self.g.ops:
            lv = 0
            for (t, _role) in op_inputs(op):
                lv = max(lv, level_of_tensor.get(t.name, -1) + 1)
            for t in op_outputs(op):
                level_of_tensor[t.name] = lv
            while len(levels) <= lv:
                levels.append([])
            levels[lv].append(op)
        if not self.group_convs:
            # graph order (one op per step); still expressed as levels
            levels = [[op] for op in self.g.ops]
        self.schedule = [l for l in levels if l]
        self.step_of = {}
        for idx, l in enumerate(self.schedule):
            for op in l:
                self.step_of[id(op)] = idx

    # -- tensor formats -----------------------------------------------------------------------
    def _formats(self):
        """Views each tensor carries.  fp32 path: fp32 only.  Tensor-core path: convolutions, fuse sums, pools and
        the STN sampler exchange split-fp16 planes; the kernels that work on fp32 (iuvmap_clean, soft-argmax,
        GCN head) and the tensors callers read keep an fp32 view."""
        self.fmt = {}
        g = self.g
        for name in g.tensors:
            self.fmt[name] = set()
        keep = self._keep()
        for op in g.ops:
            for (t, role) in op_inputs(op):
                if t.dtype != "f32":
                    continue
                if not self.tc or self.P == 0 or (op["op"], role) in _F32_INPUTS or t.Cp % 8 != 0:
                    self.fmt[t.name].add("f")
                else:
                    self.fmt[t.name].add("h")
        for name in g.tensors:
            t = g.tensors[name]
            if t.dtype != "f32":
                continue
            if name in keep or not self.fmt[name]:
                self.fmt[name].add("f")
        # producers that can only write fp32
        for op in g.ops:
            if op["op"] in ("avgpool", "body_fc", "gcn_head", "stn_params"):
                for t in op_outputs(op):
                    if t.dtype == "f32":
                        if "h" in self.fmt[t.name]:
                            raise RuntimeError("plan: %s output %s is needed as fp16 planes" % (op["op"], t.name))

    # -- buffers ------------------------------------------------------------------------------
    def _plan_buffers(self):
        g, B, dev = self.g, self.B, self.device
        last_use, produced = {}, {}
        for op in g.ops:
            idx = self.step_of[id(op)]
            for (t, _r) in op_inputs(op):
                last_use[t.name] = max(last_use.get(t.name, -1), idx)
            for t in op_outputs(op):
                if t.name not in produced:
                    produced[t.name] = idx
        keep = self._keep()
        free = {}
        self.buf = {}
        release_at = {}
        for name, idx in last_use.items():
            if name not in keep and not self.keep_all:
                release_at.setdefault(idx, []).append(name)
        by_step = {}
        for name, idx in produced.items():
            by_step.setdefault(idx, []).append(name)

        def take(key, maker):
            pool = free.get(key)
            return pool.pop() if pool else maker()

        self._pool_key = {}
        for idx in range(len(self.schedule)):
            for name in by_step.get(idx, []):
                t = g.tensors[name]
                shape = (B * t.nmult, t.H, t.W, t.Cp)
                numel = int(np.prod(shape))
                if t.dtype == "u8":
                    self.buf[name] = torch.empty(shape[:3], dtype=torch.uint8, device=dev)
                    continue
                fm = self.fmt[name]
                f32 = h = None
                if "f" in fm:
                    if name in keep:
                        f32 = torch.empty(shape, device=dev)
                    else:
                        f32 = take(("f", numel), lambda: torch.empty(numel, device=dev)).view(shape)
                if "h" in fm:
                    h = take(("h", numel), lambda: torch.empty(self.P * numel, dtype=torch.float16, device=dev)).view((self.P,) + shape)
                    if name in keep:
                        h = torch.empty((self.P,) + shape, dtype=torch.float16, device=dev)
                self.buf[name] = ActBuf(f32, h)
            for name in release_at.get(idx, []):
                a = self.buf.get(name)
                if not isinstance(a, ActBuf):
                    continue
                if a.f32 is not None:
                    free.setdefault(("f", a.f32.numel()), []).append(a.f32.reshape(-1))
                if a.h is not None:
                    free.setdefault(("h", a.h.numel() // self.P), []).append(a.h.reshape(-1))
        seen = {}
        for a in self.buf.values():
            for t in ([a] if torch.is_tensor(a) else [a.f32, a.h]):
                if t is not None:
                    seen[t.untyped_storage().data_ptr()] = t.untyped_storage().nbytes()
        self.bytes_alloc = sum(seen.values())

    def T(self, t):
        return self.buf[t.name]

    def shape(self, t):
        return (self.B * t.nmult, t.H, t.W, t.Cp)

    # -- conv ---------------------------------------------------------------------------------
    def _conv_desc(self, op):
        x, y = op["x"], op["y"]
        return dict(N=self.B * x.nmult, H=x.H, W=x.W, Cin=x.Cp, Cout=y.Cp, ksize=op["k"], stride=op["stride"],
                    pad=op["pad"], wsets=op["groups"], relu=int(op["relu"]),
                    flags=4 if (self.tc and self.precision == "exact") else 0)

    def _make_conv(self, op, sd):
        x, y = op["x"], op["y"]
        d = self._conv_desc(op)
        dev = self.device
        key = (op["parts"][0][0], "tc" if self.tc else "simt", d["flags"], x.Cp, y.Cp)
        if key not in self.wcache:
            w, b = pack_conv(sd, op)
            w, b = w.to(dev), b.to(dev)
            if self.tc:
                if not self.ops.conv_tc_supported(d):
                    raise RuntimeError("plan: convolution %r is not supported by the tensor-core path" % (d,))
                w = self.ops.conv_tc_pack(d, w)
            self.wcache[key] = (w, b)
        w, b = self.wcache[key]
        if self.tc:
            self.n_tc += 1
        Ho, Wo = y.H, y.W
        cost = float(d["N"]) * Ho * Wo * d["ksize"] ** 2 * d["Cin"] * d["Cout"]
        return dict(d=d, w=w, b=b, x=self.T(x), y=self.T(y), res=self.T(op["res"]) if op["res"] is not None else None,
                    cost=cost, op=op)

    def _add_glue(self, op, sd):
        kind = op["op"]
        dev, B = self.device, self.B
        if kind in ("input", "fuse", "maxpool", "avgpool", "clean_global", "clean_parts", "stn_sample"):
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

    # -- run ----------------------------------------------------------------------------------
    def _run_steps(self, image):
        ops = self.ops
        n = 0
        for kind, op in self.steps:
            if kind == "conv_group":
                ops.conv_group(op)
            elif kind == "conv_simt":
                ops.conv2d(op["d"], op["x"], op["w"], op["b"], op["res"], op["y"])
            elif kind == "input":
                ops.nchw_to_nhwc(image, self.T(op["y"]))
            elif kind == "fuse":
                ops.fuse_sum([self.T(t) for t, _ in op["terms"]], [f for _, f in op["terms"]], op["relu"], self.T(op["y"]),
                             self.shape(op["y"]))
            elif kind == "maxpool":
                ops.maxpool(self.T(op["x"]), self.T(op["y"]), self.shape(op["x"]))
            elif kind == "avgpool":
                ops.avgpool(self.T(op["x"]), self.T(op["y"]).f32, self.shape(op["x"]))
            elif kind == "clean_global":
                x, y = op["x"], op["y"]
                ops.clean_global(self.T(x), self.T(y), self.T(op["amax"]), self.vis, (self.B, x.H, x.W, x.Cp, y.Cp))
            elif kind == "stn_params":
                ops.stn_params(self.T(op["hm"]), self.T(op["amax"]), self.ratio, self.offset, self.vis_thresh,
                               self.align_corners, self.T(op["centers"]).f32, self.T(op["theta"]).f32)
            elif kind == "stn_sample":
                x = op["x"]
                ops.stn_sample(self.T(x), self.T(op["theta"]).f32, self.align_corners, self.T(op["y"]), (self.B, x.H, x.Cp))
            elif kind == "clean_parts":
                x, y = op["x"], op["y"]
                ops.clean_parts(self.T(x), self.T(y), self.raw_parts, (self.B * x.nmult, x.H, x.W, x.Cp, y.Cp))
            elif kind == "body_fc":
                ops.avgpool(self.T(op["x"]), self.pooled, self.shape(op["x"]))
                ops.linear(self.pooled, self.fc_w, self.fc_b, self.fc_add, self.T(op["y"]).f32)
                n += 1
            elif kind == "gcn_head":
                ops.gcn_head(self.gcn, self.T(op["x"]).f32, self.T(op["gpara"]).f32, self.T(op["y"]).f32)
            n += 1
        self.n_launch = n

    def run(self, image):
        """image [B,3,H,W] fp32 NCHW on the plan's device.  Results stay in the plan's buffers."""
        if image.shape[0] != self.B:
            raise ValueError("plan compiled for batch %d, got %d" % (self.B, image.shape[0]))
        image = image.detach().to(self.device, torch.float32).contiguous()
        with self._guard():
            if not self.use_cuda_graph:
                self._run_steps(image)
                return
            cur = torch.cuda.current_stream(self.device)
            if self.graph_exec is None:
                self.static_in = image.clone()
                s = torch.cuda.Stream(device=self.device)
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    self._run_steps(self.static_in)          # warm-up outside capture
                cur.wait_stream(s)
                self.graph_exec = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_exec):
                    self._run_steps(self.static_in)
            self.static_in.copy_(image, non_blocking=True)
            self.graph_exec.replay()

    # -- export ---------------------------------------------------------------------------------
    OPCODES = {"input": 1, "conv_group": 2, "conv_simt": 3, "fuse": 4, "maxpool": 5, "avgpool": 6, "clean_global": 7,
               "clean_parts": 8, "stn_params": 9, "stn_sample": 10, "linear": 11, "gcn_head": 12}
    _DESC_KEYS = ("N", "H", "W", "Cin", "Cout", "ksize", "stride", "pad", "wsets", "relu", "flags")

    def export(self, path=None):
        """Serialise this plan as a network program for danet_net_load (csrc/net.cu, include/danet_b200.h): the launch
        steps with their arguments, the activation buffer table and the folded / packed weights.  Returns the bytes
        (and writes them to `path` when given).  The program is tied to this plan's batch size and precision."""
        import struct
        bufs, consts = {}, {}                          # storage ptr -> (id, nbytes) ; tensor ptr -> (id, tensor)

        def bref(t):
            if t is None:
                return (0, 0, 0)
            st = t.untyped_storage()
            key = st.data_ptr()
            if key not in bufs:
                bufs[key] = (len(bufs), st.nbytes())
            return (1, bufs[key][0], t.data_ptr() - key)

        def cref(t):
            if t is None:
                return (0, 0, 0)
            if not t.is_contiguous():
                raise RuntimeError("export: constant tensors must be contiguous")
            key = t.data_ptr()
            if key not in consts:
                consts[key] = (len(consts), t)
            return (2, consts[key][0], 0)

        def aref(a):
            if a is None:
                return [(0, 0, 0)] * 3
            lo = a.h[1] if (a.h is not None and a.h.shape[0] > 1) else None
            return [bref(a.f32), bref(a.h[0] if a.h is not None else None), bref(lo)]

        def desc(d):
            return [int(d.get(k, 0)) for k in self._DESC_KEYS]

        steps = []                                      # (opcode, ints, floats, refs)
        OP = self.OPCODES
        for kind, op in self.steps:
            if kind == "conv_group":
                ints, refs = [len(op)], []
                for cv in op:
                    ints += desc(cv["d"])
                    refs += aref(cv["x"]) + aref(cv["res"]) + aref(cv["y"]) + [cref(cv["w"]), cref(cv["b"])]
                steps.append((OP[kind], ints, [], refs))
            elif kind == "conv_simt":
                res = op["res"].f32 if op["res"] is not None else None
                steps.append((OP[kind], desc(op["d"]), [], [bref(op["x"].f32), cref(op["w"]), cref(op["b"]), bref(res), bref(op["y"].f32)]))
            elif kind == "input":
                y = self.T(op["y"])
                t = y.f32 if y.f32 is not None else y.h[0]
                Hin, Win = op["y"].H, op["y"].W
                steps.append((OP[kind], [self.B, 3, Hin * Win, t.shape[-1]], [], [(3, 0, 0)] + aref(y)))
            elif kind == "fuse":
                N, H, W, C = self.shape(op["y"])
                terms = op["terms"]
                refs = []
                for t, _f in terms:
                    refs += aref(self.T(t))
                steps.append((OP[kind], [N, H, W, C, len(terms), int(op["relu"])] + [int(f) for _, f in terms], [],
                              refs + aref(self.T(op["y"]))))
            elif kind == "maxpool":
                steps.append((OP[kind], list(self.shape(op["x"])), [], aref(self.T(op["x"])) + aref(self.T(op["y"]))))
            elif kind == "avgpool":
                N, H, W, C = self.shape(op["x"])
                steps.append((OP[kind], [N, H * W, C], [], aref(self.T(op["x"])) + [bref(self.T(op["y"]).f32)]))
            elif kind == "clean_global":
                x, y = op["x"], op["y"]
                vis = self.vis if self.vis is not None else [None] * 4
                steps.append((OP[kind], [self.B, x.H * x.W, x.Cp, 0, 25, 50, 75, y.Cp], [],
                              [bref(self.T(x).f32)] + aref(self.T(y)) + [bref(self.T(op["amax"]))] + [bref(v) for v in vis]))
            elif kind == "stn_params":
                hm = self.T(op["hm"]).f32
                steps.append((OP[kind], [hm.shape[0], hm.shape[1], hm.shape[3], int(self.align_corners)], [float(self.vis_thresh)],
                              [bref(hm), bref(self.T(op["amax"])), cref(self.ratio), cref(self.offset),
                               bref(self.T(op["centers"]).f32), bref(self.T(op["theta"]).f32)]))
            elif kind == "stn_sample":
                x = op["x"]
                steps.append((OP[kind], [self.B, x.H, x.Cp, int(self.align_corners)], [],
                              aref(self.T(x)) + [bref(self.T(op["theta"]).f32)] + aref(self.T(op["y"]))))
            elif kind == "clean_parts":
                x, y = op["x"], op["y"]
                steps.append((OP[kind], [self.B * x.nmult, x.H * x.W, x.Cp, y.Cp], [],
                              [bref(self.T(x).f32)] + aref(self.T(y)) + [bref(self.raw_parts)]))
            elif kind == "body_fc":
                N, H, W, C = self.shape(op["x"])
                steps.append((OP["avgpool"], [N, H * W, C], [], aref(self.T(op["x"])) + [bref(self.pooled)]))
                steps.append((OP["linear"], [N, self.fc_w.shape[1], self.fc_w.shape[0]], [],
                              [bref(self.pooled), cref(self.fc_w), cref(self.fc_b), cref(self.fc_add), bref(self.T(op["y"]).f32)]))
            elif kind == "gcn_head":
                gp = self.gcn
                ints = [self.B] + [int(w.shape[0]) for w in gp["W"]] + [int(w.shape[1]) for w in gp["W"]]
                refs = [cref(gp["adj"])] + [cref(t) for t in gp["W"]] + [cref(t) for t in gp["b"]] + \
                       [cref(t) for t in gp["bn_scale"]] + [cref(t) for t in gp["bn_shift"]] + \
                       [cref(gp["head_w"]), cref(gp["head_b"]), cref(gp["mean_pose"]), bref(self.T(op["x"]).f32),
                        bref(self.T(op["gpara"]).f32), bref(self.T(op["y"]).f32)]
                steps.append((OP[kind], ints, [], refs))
            else:
                raise ValueError("export: unknown step %s" % kind)

        outs = []                                       # (name, ref, elem_bytes, dims)
        for k in self.KEEP:
            if k not in self.g.outputs:
                continue
            a = self.T(self.g.outputs[k])
            t = a if torch.is_tensor(a) else a.f32
            outs.append((k, bref(t), t.element_size(), list(t.shape)))
        if self.vis is not None:
            for nm, t in zip(("vis_u", "vis_v", "vis_i", "vis_a"), self.vis):
                outs.append((nm, bref(t), 4, list(t.shape)))
            outs.append(("part_iuv_raw", bref(self.raw_parts), 4, list(self.raw_parts.shape)))

        def pad16(b):
            return b + b"\0" * (-len(b) % 16)

        step_bytes = b""
        for (code, ints, floats, refs) in steps:
            step_bytes += struct.pack("<4I", code, len(ints), len(floats), len(refs))
            step_bytes += struct.pack("<%di" % len(ints), *ints) + struct.pack("<%df" % len(floats), *floats)
            for r in refs:
                step_bytes += struct.pack("<IIQ", *r)
        buf_list = sorted(bufs.values())
        const_list = sorted(consts.values(), key=lambda c: c[0])
        hdr_size = 8 + 12 * 4 + 4 * 8
        tables_size = 8 * len(buf_list) + 16 * len(const_list) + 72 * len(outs)
        steps_off = (hdr_size + tables_size + 15) // 16 * 16
        payload_off = (steps_off + len(step_bytes) + 15) // 16 * 16
        crecs, off = [], payload_off
        for (_i, t) in const_list:
            n = t.numel() * t.element_size()
            crecs.append((off, n))
            off += (n + 15) // 16 * 16
        payload_bytes = off - payload_off
        img = [op["y"] for kind, op in self.steps if kind == "input"][0]
        Hin, Win = img.H, img.W
        prec = 2 if not self.tc else (1 if self.precision == "exact" else 0)
        blob = bytearray()
        blob += b"DANETPRG" + struct.pack("<12I", 1, self.B, 3, Hin, Win, len(buf_list), len(const_list), len(outs), len(steps), prec, 0, 0)
        blob += struct.pack("<4Q", steps_off, len(step_bytes), payload_off, payload_bytes)
        for (_i, nbytes) in buf_list:
            blob += struct.pack("<Q", nbytes)
        for (o, n) in crecs:
            blob += struct.pack("<QQ", o, n)
        for (name, ref, eb, dims) in outs:
            d = (list(dims) + [1, 1, 1, 1])[:4]
            if len(dims) > 4:
                raise RuntimeError("export: output %s has more than 4 dims" % name)
            blob += name.encode()[:31].ljust(32, b"\0") + struct.pack("<IIQ", *ref) + struct.pack("<Ii4i", eb, len(dims), *d)
        blob += b"\0" * (steps_off - len(blob))
        blob += step_bytes
        blob += b"\0" * (payload_off - len(blob))
        with self._guard():
            for (_i, t), (o, n) in zip(const_list, crecs):
                assert len(blob) == o
                blob += pad16(t.detach().reshape(-1).view(torch.uint8).cpu().numpy().tobytes())
        blob = bytes(blob)
        if path is not None:
            with open(path, "wb") as f:
                f.write(blob)
        return blob

    def out(self, name):
        """fp32 tensor of a graph output."""
        a = self.T(self.g.outputs[name])
        return a if torch.is_tensor(a) else a.value()
