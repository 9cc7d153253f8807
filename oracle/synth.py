"""Synthetic stand-ins for the licensed assets (re-exported from the package's generator so the
oracle, the tests and the benchmark all see byte-identical arrays).  TEST INFRASTRUCTURE."""
from danet_b200.synthetic import (  # noqa: F401
    NV, NF, NJ, N_DP_V, N_DP_F, PARENTS, SELECTED_VERTS, make_smpl_model, make_dp_mesh,
    make_mean_params, dp_textures)
