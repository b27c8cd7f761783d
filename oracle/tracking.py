"""Oracle: per-video detection linking (TEST INFRASTRUCTURE ONLY).

Restates lib/core/tracking_engine.py:
  _center_boxes            :86-93      center_boxes
  _get_big_inside_image_boxes / _prune_bad_detections  :711-748   prune_boxes
  _compute_distance_matrix :158-181    distance_matrix ('bbox-overlap' cost only;
                                       the other two have weight 0 in every
                                       shipped yaml, config.py:550-551)
  _compute_matches         :209-246    compute_matches
  _compute_tracks_video    :272-350    compute_tracks_video (non-debug path)

``solver='scipy'`` calls the installed scipy (what the reference itself calls);
``solver='oracle'`` uses oracle.lsa.lsap_crouse (same indices, see oracle/lsa).
"""
import numpy as np

from . import boxes as obox
from .lsa import lsap_crouse, bipartite_matching_greedy

MAX_TRACK_IDS = 999     # tracking_engine.py:45
FIRST_TRACK_ID = 0      # tracking_engine.py:46


def center_boxes(boxes):
    """:86-93 — keep the centre frame's 4 columns + score."""
    if len(boxes) == 0:
        return boxes
    T = (boxes.shape[-1] - 1) // 4
    c = T // 2
    cols = list(range(c * 4, (c + 1) * 4)) + [-1]
    return boxes[:, np.array(cols)]


def prune_boxes(boxes, height, width, conf):
    """:711-748 — clips IN PLACE (as the reference does), returns kept rows.
    Note the area test has no '+1' (:716-717) and the clip is to [0,w]x[0,h]."""
    boxes[:, 0] = np.maximum(boxes[:, 0], 0)
    boxes[:, 1] = np.maximum(boxes[:, 1], 0)
    boxes[:, 2] = np.minimum(boxes[:, 2], width)
    boxes[:, 3] = np.minimum(boxes[:, 3], height)
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    sel = np.where(np.logical_and(boxes[:, -1] >= conf, area >= 50))[0]
    return sel


def distance_matrix(prev_boxes, cur_boxes, weight=1.0):
    """:158-181 with the default metric: (1 - IoU) * w, fp32."""
    C = (np.float32(1) - obox.bbox_overlaps(prev_boxes, cur_boxes)).astype(np.float32)
    C *= np.float32(weight)
    return C


def compute_matches(prev_boxes, cur_boxes, algo='hungarian', C=None, solver='oracle'):
    """:209-246.  matches[q] = index of the previous-frame box, or -1."""
    if C is None:
        C = distance_matrix(prev_boxes, cur_boxes)
    matches = -np.ones((C.shape[1],), dtype=np.int32)
    if algo == 'hungarian':
        if solver == 'scipy':
            import scipy.optimize
            prev_inds, next_inds = scipy.optimize.linear_sum_assignment(C)
        else:
            prev_inds, next_inds = lsap_crouse(C)
    elif algo == 'greedy':
        prev_inds, next_inds = bipartite_matching_greedy(C)
    else:
        raise NotImplementedError(algo)
    for p, q in zip(prev_inds, next_inds):
        matches[q] = p
    return matches


def compute_tracks_video(frames_boxes, algo='hungarian', solver='oracle'):
    """:272-350 for one video: list of (n_t, 5) fp32 arrays -> list of id lists.
    Track ids: next_track_id++ and wrap ``%= 999`` once it reaches 999 (:339-345)."""
    video_tracks = []
    next_id = FIRST_TRACK_ID
    for f, cur in enumerate(frames_boxes):
        if f == 0:
            matches = -np.ones((cur.shape[0],), dtype=np.int32)
        else:
            matches = compute_matches(frames_boxes[f - 1], cur, algo, solver=solver)
        prev_tracks = video_tracks[f - 1] if f > 0 else None
        tracks = []
        for m in matches:
            if m == -1:
                tracks.append(next_id)
                next_id += 1
                if next_id >= MAX_TRACK_IDS:
                    next_id %= MAX_TRACK_IDS
            else:
                tracks.append(prev_tracks[m])
        video_tracks.append(tracks)
    return video_tracks


def synth_video(rng, n_frames=30, n_dets=100, width=1333, height=800, T=1):
    """SURVEY.md §8(d) config 1 generator: frame 0 random boxes, frame t =
    frame t-1 + N(0,5^2) px jitter, rows permuted; scores U(0.95,1); fp32."""
    x1 = rng.uniform(0, 1133, n_dets); y1 = rng.uniform(0, 500, n_dets)
    w = rng.uniform(30, 180, n_dets); h = rng.uniform(50, 300, n_dets)
    b = np.stack([x1, y1, x1 + w, y1 + h], 1)
    frames = []
    for f in range(n_frames):
        if f > 0:
            b = b + rng.normal(0, 5.0, b.shape)
            b = b[rng.permutation(n_dets)]
        sc = rng.uniform(0.95, 1.0, (n_dets, 1))
        frames.append(np.hstack([np.tile(b, (1, T)), sc]).astype(np.float32))
    return frames
