"""Golden vectors for the PoseTrack result writer (detectandtrack_b200/core/mpii_eval_engine.py), produced by
running the REFERENCE's own lib/core/mpii_eval_engine.py (coco2posetrack, _convert_data_to_annorect_struct) on
seeded inputs.  Runs only in the build container (needs /root/reference).

    python tests/golden/gen_golden_posetrack.py     -> tests/golden/posetrack_writer.json

Shims (on top of gen_golden._setup_reference_imports): ``cPickle`` -> pickle, ``h5py`` stubbed (imported by
utils/video_io.py, never called).  The container's numpy (2.x, NEP 50) keeps float32 where the reference's
pinned numpy 1.14 promoted ``float32 op python_float`` to float64, so derived joints (neck, head_top) and the
'scaled' confidence agree with the reference environment only to fp32 rounding — the test compares those with a
1e-6 relative tolerance and everything else exactly.
"""
import json
import os
import pickle
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def main():
    import gen_golden as g
    cfg = g._setup_reference_imports()
    sys.modules['cPickle'] = pickle
    sys.modules.setdefault('h5py', types.ModuleType('h5py'))
    import core.mpii_eval_engine as ref
    rng = np.random.default_rng(20260923)
    cases = []
    for kp_conf_type, kpt_thr, drop_thr in [('global', -float('inf'), 0.5), ('local', 1.0, 0.5), ('scaled', -float('inf'), 0.9)]:
        cfg.TRACKING.KP_CONF_TYPE = kp_conf_type
        cfg.EVAL.EVAL_MPII_KPT_THRESHOLD = kpt_thr
        cfg.EVAL.EVAL_MPII_DROP_DETECTION_THRESHOLD = drop_thr
        for n in (0, 1, 7):
            boxes = np.hstack([rng.uniform(0, 800, (n, 4)), rng.uniform(0.3, 1.0, (n, 1))]).astype(np.float32)
            poses = [np.vstack([rng.uniform(0, 1333, 17), rng.uniform(0, 800, 17), rng.normal(2, 3, 17), rng.uniform(0, 1, 17)]).astype(np.float32)
                     for _ in range(n)]
            tracks = [int(t) for t in rng.integers(0, 999, n)]
            out = ref._convert_data_to_annorect_struct(boxes, poses, tracks)
            out = json.loads(json.dumps(out, default=float))          # numpy scalars -> python floats
            cases.append(dict(kp_conf_type=kp_conf_type, kpt_thr=(None if kpt_thr == -float('inf') else kpt_thr), drop_thr=drop_thr,
                              boxes=boxes.tolist(), poses=[p.tolist() for p in poses], tracks=tracks, annorect=out))
    with open(os.path.join(HERE, 'posetrack_writer.json'), 'w') as f:
        json.dump(dict(cases=cases, src_keypoints=ref.posetrack_src_keypoints, dst_keypoints=ref.dst_keypoints,
                       coco_src_keypoints=ref.coco_src_keypoints), f)
    print('wrote', len(cases), 'cases')


if __name__ == '__main__':
    main()
