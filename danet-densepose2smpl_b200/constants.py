"""Constants of the hot path (values follow the reference's constants.py / utils/smpl_utlis.py;
file:line cited per item)."""
import numpy as np

FOCAL_LENGTH = 5000.0            # constants.py:2
IMG_RES = 224                    # constants.py:3
IMG_NORM_MEAN = [0.485, 0.456, 0.406]
IMG_NORM_STD = [0.229, 0.224, 0.225]

# constants.py:15-69 JOINT_NAMES order mapped through constants.py:73-91 JOINT_MAP: index of each
# of the 49 output joints inside cat(45 smplx joints, 9 extra-regressor joints)
JOINT_MAP_49 = [24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 25, 26, 27, 28, 29, 30, 31, 32,
                33, 34, 8, 5, 45, 46, 4, 7, 21, 19, 17, 16, 18, 20, 47, 48, 49, 50, 51, 52, 53, 24,
                26, 25, 28, 27]
# constants.py:95-101
H36M_TO_J17 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9]
H36M_TO_J14 = H36M_TO_J17[:14]
J24_TO_J17 = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 18, 14, 16, 17]
J24_TO_J14 = J24_TO_J17[:14]
J24_TO_J19 = J24_TO_J17[:14] + [19, 20, 21, 22, 23]
J24_TO_JCOCO = [19, 20, 21, 22, 23, 9, 8, 10, 7, 11, 6, 3, 2, 4, 1, 5, 0]

# smplx VertexJointSelector vertex ids: face 5, feet 6, fingertips 10
SMPLX_SELECTED_VERTS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                        2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]

# utils/smpl_utlis.py:13 (row 0) with root -1: the kinematic tree smplx stores as `parents`
SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

# utils/smpl_utlis.py:55-79 dp2smpl_mapping (read by demo.py:139)
DP2SMPL_MAPPING = [[7, 8, 9, 10, 1, 2], [1, 2, 8, 10, 12, 14], [1, 2, 7, 9, 11, 13], [7, 8, 9, 10, 1, 2],
                   [1, 2, 8, 10, 12, 14], [1, 2, 7, 9, 11, 13], [7, 8, 9, 10, 1, 2], [8, 10, 12, 14, 5, 5],
                   [7, 9, 11, 13, 6, 6], [7, 8, 9, 10, 1, 2], [8, 10, 12, 14, 5, 5], [7, 9, 11, 13, 6, 6],
                   [1, 2, 23, 24, 23, 24], [1, 2, 15, 17, 19, 21], [1, 2, 16, 18, 20, 22], [1, 2, 23, 24, 23, 24],
                   [1, 2, 15, 17, 19, 21], [1, 2, 16, 18, 20, 22], [1, 2, 15, 17, 19, 21], [1, 2, 16, 18, 20, 22],
                   [15, 17, 19, 21, 4, 4], [16, 18, 20, 22, 3, 3], [15, 17, 19, 21, 4, 4], [16, 18, 20, 22, 3, 3]]

# data/pretrained_model/learned_ratio.pkl (the only asset the reference ships; loaded at
# iuv_estimator.py:21-31 into the learned_ratio / learned_offset buffers, which a checkpoint's
# state_dict overrides).  Values copied as data, 8 significant digits.
LEARNED_RATIO = np.array([0.6827488, 1.2050959, 1.1849039, 1.3892102, 1.0949879, 1.0947448, 1.6018374,
                          1.0222101, 1.0536219, 0.8735159, 0.35833353, 0.44389617, 1.0155953, 1.2463734,
                          1.2582259, 0.5802805, 1.1734062, 1.2033107, 1.1978842, 1.204344, 0.84852725,
                          0.8551517, 0.46325213, 0.3972259], dtype=np.float32)
LEARNED_OFFSET = np.array([0.09105359, 0.02297057, 0.02257976, 0.2006476, 0.01430975, 0.01649577,
                           0.11027719, 0.06102319, 0.06142722, 0.16606377, 0.7373183, 0.7548186,
                           0.07830715, 0.15315747, 0.14974837, 0.25240502, 0.06382725, 0.06352104,
                           0.046521, 0.0466027, 0.06009533, 0.05492286, 0.21719937, 0.21409516],
                          dtype=np.float32)
