"""CPU: clip assembly (detectandtrack_b200/utils/video.py get_clip) against outputs of the REFERENCE's own
lib/utils/video.py:149-201 on a synthetic roidb with a missing frame, video ends and several track ids
(tests/golden/video_clips.npz, tests/golden/gen_golden_video.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'video_clips.npz'))
sys.path.insert(0, os.path.join(HERE, 'golden'))


@pytest.mark.parametrize('tag,T,mid,interval,drop', [('t3', 3, 1, 1, False), ('t3m3', 3, 3, 1, False), ('t5i2', 5, 3, 2, False), ('t3drop', 3, 1, 1, True)])
def test_get_clip_equals_reference(tag, T, mid, interval, drop):
    from gen_golden_video import synth_roidb
    from detectandtrack_b200.core.config import cfg, reset_cfg
    from detectandtrack_b200.utils import video
    reset_cfg()
    try:
        cfg.VIDEO.NUM_FRAMES, cfg.VIDEO.NUM_FRAMES_MID, cfg.VIDEO.TIME_INTERVAL = T, mid, interval
        roidb = synth_roidb()
        for e in roidb:
            e['gt_overlaps'] = e.pop('gt_overlaps_dense')
        res = video.get_clip(roidb, remove_imperfect=drop)
        assert len(res) == int(G[tag + '_n'])
        for i, e in enumerate(res):
            assert e['image'] == G['%s_%d_image' % (tag, i)].tolist() and len(e['image']) == T
            assert e['all_frame_ids'] == G['%s_%d_frames' % (tag, i)].tolist()
            assert e['id'] == int(G['%s_%d_id' % (tag, i)])
            for k in ('tracks', 'boxes', 'gt_keypoints', 'track_visible', 'gt_classes'):
                assert np.array_equal(e[k], G['%s_%d_%s' % (tag, i, k)]), (i, k)
            assert e['boxes'].shape[1] == 4 * mid and e['gt_keypoints'].shape[2] == 17 * mid
    finally:
        reset_cfg()
