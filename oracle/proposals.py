"""Oracle: anchors and RPN proposal generation (TEST INFRASTRUCTURE ONLY).

Restates
  lib/modeling/generate_anchors.py:42-140     generate_anchors ('replicate' tubes)
  lib/ops/generate_proposals.py:40-196        generate_proposals (one image)
  lib/ops/collect_and_distribute_fpn_rpn_proposals.py:44-87   collect / distribute
  lib/modeling/FPN.py:349-360                 map_rois_to_fpn_levels
  lib/ops/roi_blob_transforms.py:25-36        roi_to_batch_format
Pinned by tests/golden/{anchors,gp2d,gp3d,cd2d,cd3d,r2b}.npz, which were produced
by the reference's own modules (tests/golden/gen_golden.py).
"""
import numpy as np

from . import boxes as obox

F32 = np.float32


# ------------------------------------------------------------------ anchors
def _whctrs(a):
    w = a[2] - a[0] + 1
    h = a[3] - a[1] + 1
    return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)


def _mkanchors(ws, hs, xc, yc):
    ws = ws[:, None]; hs = hs[:, None]
    return np.hstack((xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1), xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)))


def generate_anchors(stride=16, sizes=(32, 64, 128, 256, 512), aspect_ratios=(0.5, 1, 2), time_dim=1):
    """generate_anchors.py:42-78 (float64, like the reference), tubes by 'replicate'."""
    scales = np.array(sizes, dtype=np.float64) / stride
    ratios = np.array(aspect_ratios, dtype=np.float64)
    base = np.array([1, 1, stride, stride], dtype=np.float64) - 1
    w, h, xc, yc = _whctrs(base)
    size_ratios = (w * h) / ratios
    ws = np.round(np.sqrt(size_ratios))
    hs = np.round(ws * ratios)
    ratio_anchors = _mkanchors(ws, hs, xc, yc)
    out = []
    for i in range(ratio_anchors.shape[0]):
        w, h, xc, yc = _whctrs(ratio_anchors[i])
        out.append(_mkanchors(w * scales, h * scales, xc, yc))
    return np.tile(np.vstack(out), [1, time_dim])


def shifted_anchors(anchors, height, width, feat_stride, T):
    """generate_proposals.py:135-161: (H*W*A, 4T) float64, rows ordered (H, W, A)."""
    sx = np.arange(0, width) * feat_stride
    sy = np.arange(0, height) * feat_stride
    sx, sy = np.meshgrid(sx, sy, copy=False)
    shifts = np.vstack((sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel())).transpose()
    shifts = np.tile(shifts, [1, T])
    A = anchors.shape[0]
    K = shifts.shape[0]
    return (anchors[None, :, :] + shifts[:, None, :]).reshape((K * A, 4 * T))


def filter_boxes(boxes, min_size, im_info):
    """generate_proposals.py:184-196."""
    min_size = F32(min_size) * im_info[2]
    ws = boxes[:, 2] - boxes[:, 0] + F32(1)
    hs = boxes[:, 3] - boxes[:, 1] + F32(1)
    x_ctr = boxes[:, 0] + ws / F32(2.)
    y_ctr = boxes[:, 1] + hs / F32(2.)
    return np.where((ws >= min_size) & (hs >= min_size) & (x_ctr < im_info[1]) & (y_ctr < im_info[0]))[0]


def generate_proposals(scores, bbox_deltas, im_info, anchors, feat_stride,
                       pre_nms_topN=1000, post_nms_topN=1000, nms_thresh=0.7, min_size=0,
                       return_intermediate=False):
    """generate_proposals.py:40-114 for ONE image.
    scores (A,H,W) f32, bbox_deltas (4*A*T,H,W) f32, im_info (3,) f32 [h,w,scale].
    Returns proposals (n,4T) f32, scores (n,1) f32 — (:163-174 prepend the batch idx)."""
    A = anchors.shape[0]
    T = bbox_deltas.shape[0] // (4 * A)
    H, W = scores.shape[-2:]
    all_anchors = shifted_anchors(anchors, H, W, feat_stride, T)
    bbox_deltas = bbox_deltas.transpose((1, 2, 0)).reshape((-1, 4 * T))
    scores = scores.transpose((1, 2, 0)).reshape((-1, 1))
    if pre_nms_topN <= 0 or pre_nms_topN > len(scores):
        order = np.argsort(-scores.squeeze(), kind='stable')
    else:
        # argpartition + argsort of the partition == top-k by descending score (distinct scores)
        order = np.argsort(-scores.squeeze(), kind='stable')[:pre_nms_topN]
    bbox_deltas = bbox_deltas[order, :]
    all_anchors = all_anchors[order, :]
    scores = scores[order]
    proposals = obox.bbox_transform(all_anchors, bbox_deltas, (1.0, 1.0, 1.0, 1.0))
    proposals = obox.clip_tiled_boxes(proposals, im_info[:2])
    keep = np.arange(proposals.shape[0])
    for t in range(T):
        keep = np.intersect1d(keep, filter_boxes(proposals[:, 4 * t:4 * t + 4], min_size, im_info))
    proposals = proposals[keep, :]
    scores = scores[keep]
    pre = (proposals.copy(), scores.copy())
    if nms_thresh > 0:
        keep = obox.nms(np.hstack((proposals, scores)), nms_thresh)
        if post_nms_topN > 0:
            keep = keep[:post_nms_topN]
        keep = np.asarray(keep, dtype=np.int64)
        proposals = proposals[keep, :]
        scores = scores[keep]
    if return_intermediate:
        return proposals, scores, pre, keep
    return proposals, scores


# --------------------------------------------------- collect / distribute
def map_rois_to_fpn_levels(rois, k_min=2, k_max=5, s0=224, lvl0=4):
    """FPN.py:349-360 (float64 when rois are fp32? no: np.sqrt(f32)->f32, '/ s0' stays f32,
    '+ 1e-6' stays f32 under numpy-1.14 value casting, log2 f32, 'lvl0 +' f32)."""
    s = np.sqrt(obox.boxes_area(rois))
    dt = s.dtype.type
    lvls = np.floor(dt(lvl0) + np.log2(s / dt(s0) + dt(1e-6)))
    return np.clip(lvls, k_min, k_max)


def collect(rois_per_level, scores_per_level, post_nms_topN=1000):
    """collect_and_distribute...py:44-62.  Distinct scores assumed (argsort tie order)."""
    rois = np.concatenate(rois_per_level)
    scores = np.concatenate(scores_per_level).squeeze()
    inds = np.argsort(-scores, kind='stable')[:post_nms_topN]
    return rois[inds, :]


def distribute(rois, lvl_min=2, lvl_max=5):
    """:65-87.  Returns (rois, [rois of level l], rois_idx_restore int32)."""
    lvls = map_rois_to_fpn_levels(rois[:, 1:], lvl_min, lvl_max)
    per_level = []
    order = np.empty((0,))
    for lvl in range(lvl_min, lvl_max + 1):
        idx = np.where(lvls == lvl)[0]
        per_level.append(rois[idx, :])
        order = np.concatenate((order, idx))
    restore = np.argsort(order, kind='stable').astype(np.int32)
    return rois, per_level, restore


def roi_to_batch_format(tubes):
    """roi_blob_transforms.py:25-36: N x (4T+1) -> N*T x 5, batch idx b*T + t."""
    T = (tubes.shape[1] - 1) // 4
    N = tubes.shape[0]
    out = np.zeros((N * T, 5))
    for t in range(T):
        out[t::T, 0] = tubes[:, 0] * T + t
        out[t::T, 1:] = tubes[:, 1 + 4 * t:1 + 4 * (t + 1)]
    return out
