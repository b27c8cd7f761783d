"""CPU: the PoseTrack result writer (core/mpii_eval_engine.py) against golden vectors
produced by the reference's own lib/core/mpii_eval_engine.py (tests/golden/gen_golden_posetrack.py)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def golden():
    return json.load(open(os.path.join(HERE, 'golden', 'posetrack_writer.json')))


def _set(cfg, case):
    cfg.TRACKING.KP_CONF_TYPE = case['kp_conf_type']
    cfg.EVAL.EVAL_MPII_KPT_THRESHOLD = -float('inf') if case['kpt_thr'] is None else case['kpt_thr']
    cfg.EVAL.EVAL_MPII_DROP_DETECTION_THRESHOLD = case['drop_thr']


def _close(a, b, exact):
    if exact:
        return a == b
    return abs(a - b) <= 1e-6 * max(1.0, abs(b))


def test_keypoint_tables_match_reference(golden):
    from detectandtrack_b200.core import mpii_eval_engine as me
    assert me.posetrack_src_keypoints == golden['src_keypoints']
    assert me.dst_keypoints == golden['dst_keypoints']
    assert me.coco_src_keypoints == golden['coco_src_keypoints']


def test_annorect_struct_matches_reference(golden):
    from detectandtrack_b200.core.config import cfg, reset_cfg
    from detectandtrack_b200.core import mpii_eval_engine as me
    reset_cfg()
    try:
        for case in golden['cases']:
            _set(cfg, case)
            boxes = np.asarray(case['boxes'], np.float32).reshape(-1, 5)
            poses = [np.asarray(p, np.float32) for p in case['poses']]
            got = json.loads(json.dumps(me._convert_data_to_annorect_struct(boxes, poses, case['tracks']), default=float))
            ref = case['annorect']
            assert len(got) == len(ref)
            for g, r in zip(got, ref):
                assert g['track_id'] == r['track_id'] and g['score'] == r['score']
                gp, rp = g['annopoints'][0]['point'], r['annopoints'][0]['point']
                assert [p['id'] for p in gp] == [p['id'] for p in rp]
                for a, b in zip(gp, rp):
                    # joints copied from the detector are exact; neck / head_top (ids 12, 14) and the 'scaled'
                    # confidence carry the numpy-1.14 float64 promotion the generator's numpy 2 does not have
                    derived = a['id'][0] in (12, 14)
                    assert _close(a['x'][0], b['x'][0], not derived) and _close(a['y'][0], b['y'][0], not derived)
                    assert _close(a['score'][0], b['score'][0], case['kp_conf_type'] != 'scaled' and not (derived and case['kp_conf_type'] == 'local'))
    finally:
        reset_cfg()


def test_write_posetrack_json_groups_by_video(tmp_path):
    from detectandtrack_b200.core.config import reset_cfg
    from detectandtrack_b200.core import mpii_eval_engine as me
    reset_cfg()
    rng = np.random.default_rng(0)
    roidb = [{'image': 'images/vidA/%05d.jpg' % i} for i in range(3)] + [{'image': ['images/vidB/00000.jpg', 'images/vidB/00001.jpg', 'images/vidB/00002.jpg']}]
    boxes, keyps, tracks = [], [], []
    for i in range(4):
        n = [2, 0, 1, 3][i]
        boxes.append(np.hstack([rng.uniform(0, 500, (n, 4)), rng.uniform(0.6, 1, (n, 1))]).astype(np.float32))
        keyps.append([rng.uniform(0, 500, (4, 17)).astype(np.float32) for _ in range(n)])
        tracks.append(list(range(n)))
    dets = {'all_boxes': [[], boxes], 'all_keyps': [[], keyps], 'all_tracks': [[], tracks]}
    paths = me.write_posetrack_json(roidb, dets, str(tmp_path))
    assert sorted(paths) == ['images/vidA', 'images/vidB']
    a = json.load(open(paths['images/vidA']))['annolist']
    assert [e['imagenum'] for e in a] == [[0], [1], [2]] and [len(e['annorect']) for e in a] == [2, 1, 1]
    assert a[1]['annorect'][0]['score'] == [0] and a[1]['annorect'][0]['track_id'] == [0]      # dummy for the empty frame
    b = json.load(open(paths['images/vidB']))['annolist']
    assert b[0]['image'] == 'images/vidB/00001.jpg' and b[0]['imagenum'] == [1] and len(b[0]['annorect']) == 3
    assert len(b[0]['annorect'][0]['annopoints'][0]['point']) == 15


def test_run_mpii_eval_reads_pickle_and_calls_evaluator(tmp_path):
    import pickle
    from detectandtrack_b200.core.config import reset_cfg
    from detectandtrack_b200.core import mpii_eval_engine as me
    reset_cfg()
    assert me.run_mpii_eval(str(tmp_path), []) is None                      # no detections_withTracks.pkl yet
    rng = np.random.default_rng(1)
    roidb = [{'image': 'images/v/%05d.jpg' % i, 'frame_id': i + 1} for i in range(2)]
    boxes = [np.hstack([rng.uniform(0, 300, (2, 4)), [[0.9], [0.2]]]).astype(np.float32), np.zeros((0, 5), np.float32)]
    keyps = [[rng.uniform(0, 300, (4, 17)).astype(np.float32) for _ in range(2)], []]
    dets = {'all_boxes': [[], boxes], 'all_keyps': [[], keyps], 'all_tracks': [[], [[5, 6], []]]}
    with open(tmp_path / 'detections_withTracks.pkl', 'wb') as f:
        pickle.dump(dets, f)
    seen = {}

    def evaluator(annot_dir, out_dir, eval_tracking):
        seen['args'] = (annot_dir, out_dir, eval_tracking)
        return 'scores'
    paths, res = me.run_mpii_eval(str(tmp_path), roidb, evaluator=evaluator)
    assert res == 'scores' and seen['args'][2] is True and seen['args'][1].endswith('detections_withTracks.pkl_json/')
    ann = json.load(open(paths['images/v']))['annolist']
    assert len(ann) == 2 and len(ann[0]['annorect']) == 1 and ann[0]['annorect'][0]['track_id'] == [5]   # 0.2 < drop threshold
    assert ann[1]['annorect'][0]['score'] == [0]                                                         # dummy for the empty frame
