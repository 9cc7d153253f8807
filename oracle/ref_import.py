"""Import the reference's own network modules on CPU (SURVEY Appendix C shims).
TEST INFRASTRUCTURE; only usable where /root/reference exists (this container, not the GPU box).
Used by oracle/gen_golden.py to produce tests/golden/*.npz and by bench.py's reference arm when
the reference tree is present."""
import collections
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_ref():
    """The reference tree: $DANET_REFERENCE, else /root/reference (this container), else oracle/_ref (the
    git-ignored copy of the reference's own network modules that oracle/make_ref.py takes along to the GPU box)."""
    for cand in (os.environ.get("DANET_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "models", "danet")):
            return cand
    return "/root/reference"


REF = _find_ref()


def available():
    return os.path.isdir(os.path.join(REF, "models", "danet"))


_loaded = {}


def load(width=48):
    """Returns a namespace with the reference's cfg, IUV_Estimator, DecomposedPredictor, iuvmap_clean, ...
    NOTE: chdir()s into the reference tree (iuv_estimator.py:22 opens a relative path)."""
    if "ns" in _loaded:
        ns = _loaded["ns"]
        _set_width(ns.cfg, width)
        return ns
    if not available():
        raise RuntimeError("reference tree %s not present" % REF)
    sys.path.insert(0, REF)
    os.chdir(REF)
    for m in ['trimesh', 'neural_renderer', 'skimage', 'skimage.transform', 'smplx', 'smplx.body_models',
              'smplx.lbs', 'pyrender', 'torchgeometry', 'easydict']:
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['skimage.transform'].resize = None
    sys.modules['smplx'].SMPL = type('SMPL', (), {'__init__': lambda s, *a, **k: None})
    sys.modules['smplx.body_models'].ModelOutput = collections.namedtuple(
        'ModelOutput', ['vertices', 'joints', 'full_pose', 'betas', 'global_orient', 'body_pose'])
    sys.modules['smplx.lbs'].vertices2joints = None

    class ED(dict):
        __getattr__ = dict.__getitem__
    sys.modules['easydict'].EasyDict = ED
    import yaml
    _l = yaml.load
    yaml.load = lambda f, Loader=None: _l(f, Loader=yaml.SafeLoader)       # config.py:1070
    import torch
    import torch.cuda.comm
    torch.cuda.comm.broadcast = lambda t, devices=None: [t]                # keypoints.py:360-361
    torch.Tensor.cuda = lambda self, *a, **k: self                         # iuv_estimator.py:293 (CPU only)
    import warnings
    warnings.filterwarnings("ignore")
    from models.core.config import cfg, cfg_from_file
    cfg_from_file('configs/danet_default.yaml')
    cfg.DANET.REFINEMENT = ED(cfg.DANET.REFINEMENT)
    cfg.MSRES_MODEL.EXTRA = ED(cfg.MSRES_MODEL.EXTRA)                      # demo.py:63-64
    _set_width(cfg, width)
    from models.danet.iuv_estimator import IUV_Estimator
    from models.danet.smpl_regressor import DecomposedPredictor
    from utils import iuvmap, geometry
    ns = types.SimpleNamespace(cfg=cfg, IUV_Estimator=IUV_Estimator, DecomposedPredictor=DecomposedPredictor,
                               iuvmap=iuvmap, geometry=geometry, torch=torch)
    _loaded["ns"] = ns
    return ns


def _set_width(cfg, width):
    for st, n in (('STAGE2', 2), ('STAGE3', 3), ('STAGE4', 4)):
        cfg.HR_MODEL.EXTRA[st]['NUM_CHANNELS'] = [width * 2 ** i for i in range(n)]


def infer_para(ns, iuv_est, predictor, image):
    """The network half of DaNet.infer_net (danet.py:78-98,118) using the reference's own modules."""
    import torch
    with torch.no_grad():
        ret = iuv_est(image)
        u, v, i, a = ns.iuvmap.iuvmap_clean(*ret['uvia_pred'])
        iuv_map = torch.cat([u, v, i], dim=1)
        pp = ret['part_iuv_pred']
        parts = []
        for p in range(pp.size(1)):
            pu, pv, pi_, _ = ns.iuvmap.iuvmap_clean(pp[:, p, 0], pp[:, p, 1], pp[:, p, 2])
            parts.append(torch.stack([pu, pv, pi_], dim=1))
        part_iuv_map = torch.stack(parts, dim=1)
        out = predictor(iuv_map, part_iuv_map)
    return {"para": out['para'], "uvia_clean": (u, v, i, a), "ret": ret, "part_iuv_map": part_iuv_map}
