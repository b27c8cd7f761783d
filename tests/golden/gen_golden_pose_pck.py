"""Golden vectors for the pose-PCK tracking cost (oracle/keypoints.py:pck_distance / pairwise_kpt_distance), from the
REFERENCE's own lib/utils/keypoints.py:266-291 and lib/core/tracking_engine.py:113-129.  Build container only.

    python tests/golden/gen_golden_pose_pck.py      -> tests/golden/pose_pck.npz

Shims: gen_golden._setup_reference_imports (np.float/np.int aliases, stubs); both modules import Caffe2-era code, so
the three functions are executed from their source text (keypoints.py:266-291, tracking_engine.py:113-129).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def main():
    import gen_golden as g
    g._setup_reference_imports()
    # utils/keypoints.py imports Caffe2 through utils/blob.py, so compute_head_size / pck_distance (:266-291) and
    # tracking_engine's pairwise loop (:113-129) are executed from their source text
    import types
    ksrc = open('/root/reference/lib/utils/keypoints.py').read().split('\n')
    kps_utils = types.ModuleType('kps_utils')
    kps_utils.np = np
    exec('\n'.join(ksrc[265:291]), kps_utils.__dict__)
    src = open('/root/reference/lib/core/tracking_engine.py').read().split('\n')
    ns = {'np': np, 'kps_utils': kps_utils}
    exec('\n'.join(src[112:129]), ns)                      # def _compute_pairwise_kpt_distance(a, b, kpt_names)
    pair = ns['_compute_pairwise_kpt_distance']
    names = ['nose', 'head_bottom', 'head_top', 'left_ear', 'right_ear', 'left_shoulder', 'right_shoulder', 'left_elbow',
             'right_elbow', 'left_wrist', 'right_wrist', 'left_hip', 'right_hip', 'left_knee', 'right_knee', 'left_ankle',
             'right_ankle']
    rng = np.random.default_rng(20260924)
    out = {'names': np.array(names)}
    for tag, (na, nb, jitter) in {'small': (5, 7, 6.0), 'frame': (40, 37, 12.0), 'far': (6, 6, 200.0)}.items():
        base = np.stack([rng.uniform(100, 1200, (na, 17)), rng.uniform(50, 750, (na, 17)), rng.normal(2, 2, (na, 17)),
                         rng.uniform(0, 1, (na, 17))], 1).astype(np.float32)                     # [na, 4, 17]
        idx = rng.integers(0, na, nb)
        other = base[idx].copy()
        other[:, :2] += rng.normal(0, jitter, (nb, 2, 17)).astype(np.float32)
        a = [base[i] for i in range(na)]
        b = [other[j] for j in range(nb)]
        out[tag + '_a'], out[tag + '_b'] = base, other
        out[tag + '_dist'] = pair(a, b, names)
        out[tag + '_head'] = np.array([kps_utils.compute_head_size(x, names) for x in a], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'pose_pck.npz'), **out)
    print('wrote pose_pck.npz', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
