"""Golden vectors for the clip assembly (detectandtrack_b200/utils/video.py), from the REFERENCE's own lib/utils/video.py
get_clip / _combine_clips run unmodified on a synthetic PoseTrack-like roidb.  Build container only.

    python tests/golden/gen_golden_video.py      -> tests/golden/video_clips.npz

Shims: gen_golden._setup_reference_imports + the stub modules of gen_golden_targets (cPickle, h5py, caffe2.*); tqdm is
replaced by the identity (progress bar only)."""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


class _DS(object):
    frames_from_video = False


def synth_roidb(seed=4):
    rng = np.random.RandomState(seed)
    roidb = []
    for v, nf in (('vidA', 5), ('vidB', 3)):
        for f in range(1, nf + 1):
            if v == 'vidA' and f == 3:
                continue                                        # a missing frame in the middle of a video
            ids = sorted(rng.choice(6, size=rng.randint(1, 4), replace=False).tolist())
            n = len(ids)
            roidb.append(dict(image='%s/%06d.jpg' % (v, f), frame_id=f, flipped=False, height=240, width=320, id=len(roidb), nframes=nf,
                              is_labeled=True, has_visible_keypoints=True,
                              tracks=np.array(ids, np.int32).reshape(-1, 1), boxes=rng.uniform(0, 200, (n, 4)).astype(np.float32),
                              gt_keypoints=rng.randint(0, 300, (n, 3, 17)).astype(np.int32), gt_classes=np.ones(n, np.int32),
                              is_crowd=np.zeros(n, bool), box_to_gt_ind_map=np.arange(n, dtype=np.int32),
                              max_classes=np.ones(n, np.int64), max_overlaps=np.ones(n, np.float32), seg_areas=np.ones(n, np.float32),
                              gt_overlaps_dense=np.tile(np.array([[0., 1.]], np.float32), (n, 1)), head_boxes=-np.ones((n, 4), np.float32)))
    return roidb


def main():
    import gen_golden as g
    import gen_golden_targets as gt
    g._setup_reference_imports()
    sys.modules['cPickle'] = pickle
    for n in ['caffe2.proto.caffe2_pb2', 'caffe2.python.scope', 'caffe2.python.core', 'caffe2.python.workspace', 'caffe2.python.utils', 'h5py']:
        gt._stub(n)
    import scipy.sparse
    import utils.video as rv
    rv.tqdm = lambda x, **kw: x
    from core.config import cfg
    out = {}
    for tag, (T, mid, interval, drop) in {'t3': (3, 1, 1, False), 't3m3': (3, 3, 1, False), 't5i2': (5, 3, 2, False), 't3drop': (3, 1, 1, True)}.items():
        cfg.VIDEO.NUM_FRAMES, cfg.VIDEO.NUM_FRAMES_MID, cfg.VIDEO.TIME_INTERVAL = T, mid, interval
        roidb = synth_roidb()
        for e in roidb:
            e['dataset'] = _DS()
            e['gt_overlaps'] = scipy.sparse.csr_matrix(e.pop('gt_overlaps_dense'))
        res = rv.get_clip(roidb, remove_imperfect=drop)
        out[tag + '_n'] = np.int32(len(res))
        for i, e in enumerate(res):
            out['%s_%d_image' % (tag, i)] = np.array(e['image'])
            out['%s_%d_frames' % (tag, i)] = np.array(e['all_frame_ids'], np.int32)
            for k in ('tracks', 'boxes', 'gt_keypoints', 'track_visible', 'gt_classes'):
                out['%s_%d_%s' % (tag, i, k)] = e[k]
            out['%s_%d_id' % (tag, i)] = np.int32(e['id'])
    np.savez_compressed(os.path.join(HERE, 'video_clips.npz'), **out)
    print('wrote video_clips.npz', len(out))


if __name__ == '__main__':
    main()
