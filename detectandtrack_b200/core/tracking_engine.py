"""Per-video detection linking on the GPU — drop-in for the default path of the
reference's ``core.tracking_engine`` (lib/core/tracking_engine.py).

Kept surface: ``_compute_matches``, ``_compute_tracks_video``,
``compute_matches_tracks``, ``_prune_bad_detections``, ``_center_detections``,
``run_posetrack_tracking`` with the reference's argument meaning and the
``detections.pkl`` / ``detections_withTracks.pkl`` schema; the PoseTrack JSON
writer that follows it lives in ``core/mpii_eval_engine.py``.

What changed underneath: instead of one python loop per frame pair
(bbox_overlaps -> scipy LSA -> id list), all frames of all videos are packed
into one [F, Dmax, 4T+1] tensor and three launches do the work
(csrc/lsa.cu): fused (1 - IoU) cost + assignment for every consecutive pair
(pairs are independent), then the sequential id scan, one thread per video.
Only the 'bbox-overlap' cost runs on the device (the configuration of every shipped yaml; the
weights of the other costs are 0, config.py:550-551), with 'hungarian' or 'greedy' matching;
anything else raises NotImplementedError.
"""
import logging
import os.path as osp
import pickle

import numpy as np

from .config import cfg
from ..ops import box_ops

logger = logging.getLogger(__name__)

MAX_TRACK_IDS = box_ops.MAX_TRACK_IDS      # tracking_engine.py:45
FIRST_TRACK_ID = box_ops.FIRST_TRACK_ID    # tracking_engine.py:46


# ----------------------------------------------------------------- file I/O
def _load_det_file(det_fpath):
    with open(det_fpath, 'rb') as fin:
        try:
            return pickle.load(fin)
        except UnicodeDecodeError:          # pickles written by the py2 reference
            fin.seek(0)
            return pickle.load(fin, encoding='latin1')


def _write_det_file(dets, det_fpath):
    with open(det_fpath, 'wb') as fout:
        pickle.dump(dets, fout, pickle.HIGHEST_PROTOCOL)


def _image_path(entry):
    """lib/utils/image.py:44-48 get_image_path: the CENTRE frame of a clip entry."""
    im = entry['image']
    return im[len(im) // 2] if isinstance(im, (list, tuple)) else im


def _is_same_video(json1, json2):
    return osp.dirname(_image_path(json1)) == osp.dirname(_image_path(json2))


def _get_boxes(dets, img_id):
    return dets['all_boxes'][1][img_id]


def _get_poses(dets, img_id):
    return dets['all_keyps'][1][img_id]


# ------------------------------------------------------- centre / prune (host)
def _center_boxes(boxes):
    """:86-93"""
    if len(boxes) == 0:
        return boxes
    assert (boxes.shape[-1] - 1) % 4 == 0, 'Must contain scores in last col.'
    time_dim = (boxes.shape[-1] - 1) // 4
    center = time_dim // 2
    return boxes[:, np.array(list(range(center * 4, (center + 1) * 4)) + [-1])]


def _center_poses(poses):
    """:96-103"""
    if len(poses) == 0:
        return poses
    K = cfg.KRCNN.NUM_KEYPOINTS
    time_dim = poses[0].shape[-1] // K
    center = time_dim // 2
    return [el[..., center * K:(center + 1) * K] for el in poses]


def _center_detections(dets):
    """:751-755"""
    for img_id in range(len(dets['all_boxes'][1])):
        dets['all_boxes'][1][img_id] = _center_boxes(_get_boxes(dets, img_id))
        dets['all_keyps'][1][img_id] = _center_poses(_get_poses(dets, img_id))


def _pack(box_list, ld=None):
    """list of (n_i, ld) arrays -> padded [F, Dmax, ld] fp32 + counts."""
    F = len(box_list)
    counts = np.array([0 if b is None else len(b) for b in box_list], dtype=np.int32)
    if ld is None:
        ld = max([b.shape[1] for b in box_list if b is not None and len(b)] + [5])
    dmax = max(int(counts.max()) if F else 0, 1)
    packed = np.zeros((F, dmax, ld), dtype=np.float32)
    for i, b in enumerate(box_list):
        if counts[i]:
            packed[i, :counts[i]] = b
    return packed, counts


def _prune_bad_detections(dets, json_data, conf):
    """:731-748 on the device for all images at once (dt_prune_detections).
    Like the reference, the kept boxes are the clipped ones."""
    import torch
    boxes_l = dets['all_boxes'][1]
    N = len(boxes_l)
    if N == 0:
        return dets
    packed, counts = _pack(boxes_l)
    T = (packed.shape[2] - 1) // 4
    hw = np.array([[json_data[i]['height'], json_data[i]['width']] for i in range(N)], dtype=np.float32)
    out, counts_out, sel = box_ops.prune_detections(
        torch.from_numpy(packed).cuda(), torch.from_numpy(counts), torch.from_numpy(hw),
        conf, T=T, center_only=False)
    out, counts_out, sel = out.cpu().numpy(), counts_out.cpu().numpy(), sel.cpu().numpy()
    for i in range(N):
        n = int(counts_out[i])
        poses = dets['all_keyps'][1][i]
        dets['all_boxes'][1][i] = out[i, :n].copy()
        dets['all_keyps'][1][i] = [poses[j] for j in sel[i, :n].tolist()]
    return dets


# ------------------------------------------------------------------ matching
POSETRACK_KEYPOINTS = ['nose', 'head_bottom', 'head_top', 'left_ear', 'right_ear', 'left_shoulder', 'right_shoulder', 'left_elbow',
                       'right_elbow', 'left_wrist', 'right_wrist', 'left_hip', 'right_hip', 'left_knee', 'right_knee', 'left_ankle',
                       'right_ankle']          # person_cat_info['keypoints'] of the PoseTrack json datasets


def _cost_weights(cost_types, cost_weights):
    """(w_iou, w_pck) of TRACKING.DISTANCE_METRICS / _WTS (tracking_engine.py:158-181); 'cnn-cosdist' needs a CNN feature
    extractor (utils/pytorch_cnn_features, weight 0 in every shipped yaml) and is not on the device path."""
    assert len(cost_weights) == len(cost_types)
    w = {'bbox-overlap': 0.0, 'pose-pck': 0.0}
    for t, wt in zip(cost_types, cost_weights):
        if wt == 0:
            continue
        if t not in w:
            raise NotImplementedError('device tracking implements the bbox-overlap and pose-pck costs (got %r)' % (t,))
        w[t] += float(wt)
    if w['bbox-overlap'] == 0.0 and w['pose-pck'] == 0.0:
        raise NotImplementedError('no active tracking cost')
    return w['bbox-overlap'], w['pose-pck']


def _check_default_cost(cost_types, cost_weights):
    w_iou, w_pck = _cost_weights(cost_types, cost_weights)
    if w_pck != 0.0:
        raise NotImplementedError('this call path takes boxes only; pose-pck needs the poses (see _tracks_for_videos)')
    return w_iou


def _pack_poses(pose_list, dmax, K):
    """list (frames) of lists of [4|3, K] arrays -> [F, dmax, 4, K] fp32 (x row, y row used by the cost)."""
    out = np.zeros((len(pose_list), dmax, 4, K), dtype=np.float32)
    for f, poses in enumerate(pose_list):
        for i, p in enumerate(poses or []):
            p = np.asarray(p, dtype=np.float32)
            out[f, i, :p.shape[0]] = p
    return out


def _compute_matches(prev_frame_data, cur_frame_data, prev_boxes, cur_boxes,
                     prev_poses, cur_poses, cost_types, cost_weights,
                     bipart_match_algo, C=None):
    """:209-246.  Returns int32 matches[n_cur] = index into prev boxes, or -1."""
    import torch
    if bipart_match_algo not in box_ops.ALGOS:
        raise NotImplementedError('Unknown matching algo: {}'.format(bipart_match_algo))
    if C is not None:
        C = np.ascontiguousarray(C, dtype=np.float32)
        P, Q = C.shape
        if P == 0 or Q == 0:
            return -np.ones((Q,), dtype=np.int32)
        m, status = box_ops.lsa_batched(torch.from_numpy(C).cuda().unsqueeze(0),
                                        torch.tensor([P]), torch.tensor([Q]), bipart_match_algo)
        if int(status[0].item()) != 0:
            raise ValueError('cost matrix is infeasible')
        return m[0, :Q].cpu().numpy().astype(np.int32)
    weight = _check_default_cost(cost_types, cost_weights)
    nboxes = cur_boxes.shape[0]
    if nboxes == 0 or prev_boxes.shape[0] == 0:
        return -np.ones((nboxes,), dtype=np.int32)
    packed, counts = _pack([prev_boxes, cur_boxes])
    T = (packed.shape[2] - 1) // 4
    m, _ = box_ops.match_frames(torch.from_numpy(packed).cuda(), torch.from_numpy(counts), None, T=T, weight=weight,
                                algo=bipart_match_algo)
    return m[1, :nboxes].cpu().numpy().astype(np.int32)


def _tracks_for_videos(videos_boxes, videos_poses=None, kpt_names=None):
    """videos_boxes: list (videos) of lists (frames) of (n,4T+1) arrays; videos_poses (needed when the 'pose-pck' cost is
    active): same nesting of lists of [4, K] keypoint arrays.  Returns list of list of python-int id lists, one batched
    device pass."""
    import torch
    flat, first = [], []
    for vb in videos_boxes:
        first.append(len(flat))
        flat.extend(vb)
    if not flat:
        return [[] for _ in videos_boxes]
    packed, counts = _pack(flat)
    T = (packed.shape[2] - 1) // 4
    is_start = np.zeros(len(flat), dtype=np.uint8)
    is_start[np.array(first, dtype=np.int64)] = 1
    w_iou, w_pck = _cost_weights(cfg.TRACKING.DISTANCE_METRICS, cfg.TRACKING.DISTANCE_METRIC_WTS)
    algo = cfg.TRACKING.BIPARTITE_MATCHING_ALGO
    if algo not in box_ops.ALGOS:
        raise NotImplementedError('Unknown matching algo: {}'.format(algo))
    d_counts = torch.from_numpy(counts).cuda()
    d_start = torch.from_numpy(is_start).cuda()
    if w_pck == 0.0:
        matches, _ = box_ops.match_frames(torch.from_numpy(packed).cuda(), d_counts, d_start, T=T, weight=w_iou, algo=algo)
    else:
        # combined cost (tracking_engine.py:158-181): cost matrices of every frame pair in one launch, then the batched solver
        assert videos_poses is not None, "the 'pose-pck' cost needs the poses of every detection"
        names = list(kpt_names or POSETRACK_KEYPOINTS)
        K = cfg.KRCNN.NUM_KEYPOINTS
        if K <= 0:                                   # unset (config default -1): take it from the data
            K = next((np.asarray(p).shape[-1] for vp in videos_poses for fr in vp for p in (fr or [])), len(names))
        poses = _pack_poses([p for vp in videos_poses for p in vp], packed.shape[1], K)
        cost = box_ops.frame_costs(torch.from_numpy(packed).cuda(), d_counts, d_start, torch.from_numpy(poses).cuda(), T=T, w_iou=w_iou,
                                   w_pck=w_pck, head_top=names.index('head_top'), head_bottom=names.index('head_bottom'))
        nrows = torch.cat([d_counts[:1] * 0, d_counts[:-1]]) * (1 - d_start.to(torch.int32))
        matches, status = box_ops.lsa_batched(cost, nrows, d_counts, algo)
    tracks = box_ops.assign_track_ids(matches, d_counts, torch.tensor(first, dtype=torch.int32), d_start)
    tracks = tracks.cpu().numpy()
    out = []
    for v, vb in enumerate(videos_boxes):
        f0 = first[v]
        out.append([tracks[f0 + i, :counts[f0 + i]].tolist() for i in range(len(vb))])
    return out


def _compute_tracks_video(video_json_data, dets):
    """:272-350 (non-debug path) for one video: list of per-frame id lists."""
    if (cfg.TRACKING.DEBUG.UPPER_BOUND or cfg.TRACKING.DEBUG.UPPER_BOUND_3_SHOTS or
            cfg.TRACKING.DEBUG.UPPER_BOUND_5_GT_KPS_ONLY):
        raise NotImplementedError('TRACKING.DEBUG.* upper-bound modes are not on the device path')
    boxes = [_get_boxes(dets, det_id) for (_, det_id) in video_json_data]
    poses = [_get_poses(dets, det_id) for (_, det_id) in video_json_data]
    return _tracks_for_videos([boxes], [poses])[0]


def compute_matches_tracks(json_data, dets, lstm_model=None):
    """:669-708.  Splits images into videos by dirname, links every video in one
    batched device pass, writes dets['all_tracks'] = [[], per-image id lists]."""
    if cfg.TRACKING.LSTM_TEST.LSTM_TRACKING_ON or cfg.TRACKING.FLOW_SMOOTHING_ON:
        raise NotImplementedError('LSTM tracking / flow smoothing are outside the device path')
    num_imgs = len(json_data)
    all_tracks = [[]] * num_imgs
    all_video_roidb, video_entries = [], []
    for img_id in range(num_imgs):
        if img_id == 0 or _is_same_video(json_data[img_id - 1], json_data[img_id]):
            video_entries.append((json_data[img_id], img_id))
        else:
            all_video_roidb.append(sorted(video_entries, key=lambda x: _image_path(x[0])))
            video_entries = [(json_data[img_id], img_id)]
    if len(video_entries) > 0:
        all_video_roidb.append(video_entries)
    assert num_imgs == sum(len(v) for v in all_video_roidb)
    logger.info('Computing tracks for %d videos.', len(all_video_roidb))
    vids = [[_get_boxes(dets, det_id) for (_, det_id) in v] for v in all_video_roidb]
    vposes = [[_get_poses(dets, det_id) for (_, det_id) in v] for v in all_video_roidb]
    tracks = _tracks_for_videos(vids, vposes)
    for v, entries in enumerate(all_video_roidb):
        for i, (_, det_id) in enumerate(entries):
            all_tracks[det_id] = tracks[v][i]
    dets['all_tracks'] = [[], all_tracks]
    return dets


def run_posetrack_tracking(test_output_dir, json_data, write_json=True):
    """:758-795.  Writes detections_withTracks.pkl and then, like the reference's run_mpii_eval call
    (:792-795), the per-video PoseTrack JSON files (core/mpii_eval_engine.py).  The evaluation itself
    (vendored poseval + annotations) stays outside the package."""
    det_file = cfg.TRACKING.DETECTIONS_FILE if len(cfg.TRACKING.DETECTIONS_FILE) else \
        osp.join(test_output_dir, 'detections.pkl')
    out_det_file = osp.join(test_output_dir, 'detections_withTracks.pkl')
    if not osp.exists(det_file):
        raise ValueError('Output file not found {}'.format(det_file))
    logger.info('Tracking over %s', det_file)
    dets = _load_det_file(det_file)
    if cfg.TRACKING.KEEP_CENTER_DETS_ONLY:
        _center_detections(dets)
    assert len(json_data) == len(dets['all_boxes'][1])
    assert len(json_data) == len(dets['all_keyps'][1])
    conf = cfg.TRACKING.CONF_FILTER_INITIAL_DETS
    logger.info('Pruning detections with less than %s confidence', conf)
    dets = _prune_bad_detections(dets, json_data, conf)
    dets_withTracks = compute_matches_tracks(json_data, dets, None)
    _write_det_file(dets_withTracks, out_det_file)
    if write_json:
        from . import mpii_eval_engine
        mpii_eval_engine.run_mpii_eval(test_output_dir, json_data)
    return dets_withTracks
