"""PartRenderer (reference utils/part_utils.py:8-53): silhouette mask + 6-part segmentation of the
posed SMPL mesh at 224 x 224, the second consumer of the rasteriser (eval.py:219-266, LSP mask / part
F1).  Same call surface: `PartRenderer()(vertices, camera) -> (mask [B,R,R] float, parts [B,R,R] long)`.

The reference renders per-face constant colours with neural_renderer and maps `floor(100 * rgb)`
through the `cube_parts` lookup table; here the face winner comes from csrc/raster.cu (bit-exact
against oracle/raster.c) and the table lookup is one gather."""
import numpy as np
import torch

from .renderer import IUV_Renderer


class PartRenderer(object):
    def __init__(self, focal_length=5000., render_res=224, faces=None, textures=None, cube_parts=None,
                 num_smpl_verts=6890):
        """faces [F,3] int (SMPL(...).faces), textures [F,3] or the reference's [1,F,1,1,1,3] file
        layout (VERTEX_TEXTURE_FILE), cube_parts 3-D table (CUBE_PARTS_FILE).  When None the reference's
        files are read from the reference's locations (path_config.py:64-71)."""
        self.focal_length = focal_length
        self.render_res = render_res
        if faces is None:
            from .smpl import SMPL
            faces = SMPL("data/smpl").faces
        if textures is None:
            textures = np.load("data/vertex_texture.npy")
        if cube_parts is None:
            cube_parts = np.load("data/cube_parts.npy")
        faces = np.asarray(faces).astype(np.int64)
        textures = np.asarray(textures, dtype=np.float32).reshape(-1, 3)
        if textures.shape[0] != faces.shape[0]:
            raise ValueError("PartRenderer: %d face colours for %d faces" % (textures.shape[0], faces.shape[0]))
        mesh = {"All_vertices": np.arange(1, num_smpl_verts + 1), "FacesDensePose": faces,
                "FaceIndices": np.ones(faces.shape[0]), "U_norm": np.zeros(num_smpl_verts),
                "V_norm": np.zeros(num_smpl_verts)}
        self._r = IUV_Renderer(orig_size=render_res, out_size=render_res, focal_length=focal_length, mesh=mesh,
                               num_smpl_verts=num_smpl_verts)
        if render_res != 224:
            # part_utils.py:42-47 builds K from focal_length and render_res directly (no 224-relative rescale)
            self._r._focal_eff = float(focal_length)
        self._r.textures = torch.from_numpy(textures[None, :, None, None, None, :].copy())
        self.faces = torch.from_numpy(faces.astype(np.int32))
        self.textures = self._r.textures
        self.cube_parts = torch.as_tensor(np.asarray(cube_parts), dtype=torch.float32)

    def get_parts(self, parts, mask):
        """part_utils.py:27-36: rendered colour image -> body part indices."""
        bn, c, h, w = parts.shape
        cube = self.cube_parts.to(parts.device)
        idx = torch.floor(100 * parts.permute(0, 2, 3, 1).contiguous().view(-1, 3)).long()
        out = cube[idx[:, 0], idx[:, 1], idx[:, 2], None]
        out = out * mask.view(-1, 1)
        return out.view(bn, h, w).long()

    @torch.no_grad()
    def __call__(self, vertices, camera):
        """vertices [B,6890,3], camera [B,3] (s,tx,ty) -> (mask, parts)   (part_utils.py:38-53)."""
        img, fidx = self._r.verts2faceidx(vertices, camera)
        mask = (fidx >= 0).to(torch.float32)
        return mask, self.get_parts(img, mask)
