"""Clip assembly on the host (lib/utils/video.py:38-201): every roidb entry becomes a T-frame clip centred on its frame —
``image`` turns into the list of the T frame paths (neighbours at VIDEO.TIME_INTERVAL, clamped to the nearest existing frame
at the ends of a video) and the ground truth of the middle NUM_FRAMES_MID frames is merged into tubes: boxes [n, 4*Tm],
gt_keypoints [n, 3, K*Tm], track_visible [n, Tm], one row per track id seen in those frames.

The pixels themselves are read by core/test_engine.read_image_video on loader threads straight into the pipeline's pinned
staging buffers; resize / mean subtraction / padding run on the device (dt_prep_clip)."""
import math
import os

import numpy as np

from ..core.config import cfg


def get_video_info(roidb):
    """:38-55.  index -> (video name, key frame number, flipped)."""
    info = {}
    for i, e in enumerate(roidb):
        ds = e.get('dataset')
        if ds is not None and getattr(ds, 'frames_from_video', False):
            name, key = e['image'], e['frame_id']
        else:
            name = os.path.dirname(e['image'])
            key = int(os.path.splitext(os.path.basename(e['image']))[0])
        info[i] = (name, key, e.get('flipped', False))
    return info


def _center_crop_list(l, sz):
    assert len(l) >= sz
    start = (len(l) // 2) - (sz // 2)
    return l[start:start + sz]


_COPY = ('dataset', 'has_visible_keypoints', 'id', 'nframes', 'width', 'head_boxes', 'is_labeled', 'frame_id', 'height', 'flipped')


def _combine_clips(entry, clip):
    """:66-147.  clip: the T neighbouring entries."""
    new = {'image': [c['image'] for c in clip]}
    mid = _center_crop_list(clip, cfg.VIDEO.NUM_FRAMES_MID)
    for k in _COPY:
        if k in entry:
            new[k] = entry[k]
    Tm = len(mid)
    new['all_frame_ids'] = [c['frame_id'] for c in mid]
    if 'original_file_name' in entry:
        new['original_file_name'] = [c['original_file_name'] for c in mid]
    if 'tracks' not in entry:                                 # inference roidb without annotations: the frame list is all there is
        return new
    ids = np.array(list(set(t for c in mid for t in np.asarray(c['tracks']).reshape(-1).tolist())), dtype=entry['tracks'].dtype)
    n = len(ids)
    kp = entry['gt_keypoints']
    new['tracks'] = ids
    gk = np.zeros((n, Tm, kp.shape[-2], kp.shape[-1]), dtype=kp.dtype)
    new['boxes'] = np.zeros((n, 4 * Tm), dtype=entry['boxes'].dtype)
    new['is_crowd'] = np.zeros((n,), dtype=entry['is_crowd'].dtype)
    new['gt_classes'] = np.zeros((n,), dtype=entry['gt_classes'].dtype)
    new['track_visible'] = np.full((n, Tm), False)
    new['segms'] = [[]] * n
    new['box_to_gt_ind_map'] = np.arange(n, dtype=entry['box_to_gt_ind_map'].dtype) if 'box_to_gt_ind_map' in entry else np.arange(n, dtype=np.int32)
    new['max_classes'] = np.ones((n,), dtype=np.int64)
    new['max_overlaps'] = np.ones((n,), dtype=np.float32)
    new['seg_areas'] = np.ones((n,), dtype=np.float32)
    ncls = entry['gt_overlaps'].shape[1] if 'gt_overlaps' in entry else 2
    ov = np.zeros((n, ncls), dtype=np.float32)
    for i, tid in enumerate(ids):
        for f, c in enumerate(mid):
            tr = np.asarray(c['tracks']).reshape(-1).tolist()
            if tid in tr:
                p = tr.index(tid)
                new['boxes'][i, 4 * f:4 * f + 4] = c['boxes'][p]
                gk[i, f] = c['gt_keypoints'][p]
                new['track_visible'][i, f] = True
                new['gt_classes'][i] = c['gt_classes'][p]
                ov[i, 1] = 1.0
    new['gt_overlaps'] = ov
    new['gt_keypoints'] = gk.transpose((0, 2, 1, 3)).reshape((n, kp.shape[-2], Tm * kp.shape[-1]))     # NxTx3xK -> Nx3x(T*K)
    return new


def get_clip(roidb, remove_imperfect=False):
    """:149-201.  Returns the video-fied roidb."""
    info = get_video_info(roidb)
    pos_of = {v: k for k, v in info.items()}
    T = cfg.VIDEO.NUM_FRAMES
    half = (T - 1) / 2.0
    offsets = list(range(int(math.floor(-half)), int(math.floor(half)) + 1))
    assert len(offsets) == T and offsets[len(offsets) // 2] == 0
    out = []
    for i, entry in enumerate(roidb):
        name, key, flip = info[i]
        clip = [None] * T
        for j, d in enumerate(offsets):
            tgt = (name, key + d * cfg.VIDEO.TIME_INTERVAL, flip)
            if tgt in pos_of:
                clip[j] = roidb[pos_of[tgt]]
        if any(c is None for c in clip):
            if remove_imperfect:
                continue
            last = None
            for k in range(T // 2, -1, -1):                   # towards the start: repeat the nearest existing frame
                if clip[k] is not None:
                    last = clip[k]
                else:
                    clip[k] = last
            last = None
            for k in range(T // 2, T):
                if clip[k] is not None:
                    last = clip[k]
                else:
                    clip[k] = last
        out.append(_combine_clips(entry, clip))
    return out
