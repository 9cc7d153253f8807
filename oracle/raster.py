"""ctypes wrapper around oracle/raster.c -- TEST INFRASTRUCTURE (oracle/__init__.py).

Restates reference utils/renderer.py:256-298 (IUV_Renderer.verts2uvimg / camera_matrix)
on top of the restated neural_renderer forward pass.  Parity unpinned (see raster.c)."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "raster.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_raster_iuv.restype = None
    return _LIB


def verts2uvimg(verts, cam, mesh, textures, orig_size=224, out_size=56, focal=5000.0,
                near=0.1, far=100.0, tex_mode=0):
    """verts [B,6890,3] f32, cam [B,3] f32 -> (img [B,3,S,S] f32, face_idx [B,S,S] i32, depth)."""
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    cam = np.ascontiguousarray(cam, dtype=np.float32)
    B, nv = verts.shape[0], verts.shape[1]
    vm = np.ascontiguousarray(mesh["All_vertices"].astype(np.int64) - 1, dtype=np.int32)
    faces = np.ascontiguousarray(mesh["FacesDensePose"], dtype=np.int32)
    tex = np.ascontiguousarray(textures, dtype=np.float32)
    S = out_size
    img = np.zeros((B, 3, S, S), dtype=np.float32)
    fidx = np.zeros((B, S, S), dtype=np.int32)
    depth = np.zeros((B, S, S), dtype=np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib().oracle_raster_iuv(
        ctypes.c_int(B), ctypes.c_int(nv), p(verts), p(cam),
        ctypes.c_int(vm.shape[0]), p(vm), ctypes.c_int(faces.shape[0]), p(faces), p(tex),
        ctypes.c_int(orig_size), ctypes.c_int(S), ctypes.c_float(focal),
        ctypes.c_float(near), ctypes.c_float(far), ctypes.c_int(tex_mode),
        p(img), p(fidx), p(depth))
    return img, fidx, depth


def iuv_img2map(uvimages):
    """reference utils/iuvmap.py:103-151 (no-roi branch), numpy restatement."""
    Index2mask = [[0], [1, 2], [3], [4], [5], [6], [7, 9], [8, 10], [11, 13], [12, 14], [15, 17],
                  [16, 18], [19, 21], [20, 22], [23, 24]]
    part = np.round(uvimages[:, 0] * np.float32(24))
    I = np.stack([(part == i).astype(np.float32) for i in range(25)], 1)
    U = I * uvimages[:, 1:2]
    V = I * uvimages[:, 2:3]
    Ann = np.stack([sum(I[:, j] for j in m) for m in Index2mask], 1)
    return U, V, I, Ann
