"""Oracle: box / tube arithmetic (TEST INFRASTRUCTURE ONLY, see oracle/__init__).

Restates, in numpy fp32 with the reference's operation order:
  lib/utils/cython_bbox.pyx:16-56      bbox_overlaps_2d
  lib/utils/boxes.py:26-69             split_tube_into_boxes, bbox_overlaps
  lib/utils/boxes.py:72-78             boxes_area
  lib/utils/boxes.py:141-202           bbox_transform / tube_transform
  lib/utils/boxes.py:243-253           clip_tiled_boxes
  lib/utils/cython_nms.pyx:37-87       nms_2d           (suppress ovr >= thr)
  lib/nms/py_cpu_nms_tubes.py:17-53    nms_tubes        (suppress mean ovr > thr)
  lib/core/nms_wrapper.py:49-70        nms (dispatch)

Pinned against: oracle/_ref (reference Cython compiled unmodified / with the
np.int_t fix) and tests/golden/boxes_*.npz generated from the reference's own
python modules (tests/golden/gen_golden.py).
"""
import numpy as np

F32 = np.float32
# lib/core/config.py:672  __C.BBOX_XFORM_CLIP = np.log(1000. / 16.)
BBOX_XFORM_CLIP = float(np.log(1000. / 16.))


def bbox_overlaps_2d(boxes, query_boxes):
    """cython_bbox.pyx:16-56.  (N,4),(K,4) f32 -> (N,K) f32, '+1' convention.

    Mirrors the C that Cython emits for the .pyx (checked bit-for-bit against the
    compiled reference in oracle/_ref): coordinate differences are fp32, the
    literal ``1`` is emitted as the C double ``1.0`` so every ``+ 1`` and the
    area products run in fp64; box_area, iw, ih are rounded to fp32 when stored
    (:35-38,:42-50); ua = fp32(fp64(area_n*) + box_area - fp64(iw*ih [fp32]))
    (:51-55); result = fp32 iw*ih / ua (:56).  iw/ih tests are strict (> 0)."""
    b = np.ascontiguousarray(boxes[:, :4], dtype=F32)
    q = np.ascontiguousarray(query_boxes[:, :4], dtype=F32)
    F64 = np.float64
    qarea = (((q[:, 2] - q[:, 0]).astype(F64) + 1.0) *
             ((q[:, 3] - q[:, 1]).astype(F64) + 1.0)).astype(F32)           # (K,)
    barea64 = (((b[:, 2] - b[:, 0]).astype(F64) + 1.0) *
               ((b[:, 3] - b[:, 1]).astype(F64) + 1.0))                       # (N,) f64
    iw = ((np.minimum(b[:, None, 2], q[None, :, 2]) -
           np.maximum(b[:, None, 0], q[None, :, 0])).astype(F64) + 1.0).astype(F32)
    ih = ((np.minimum(b[:, None, 3], q[None, :, 3]) -
           np.maximum(b[:, None, 1], q[None, :, 1])).astype(F64) + 1.0).astype(F32)
    inter = iw * ih                                                          # f32
    ua = ((barea64[:, None] + qarea[None, :].astype(F64)) - inter.astype(F64)).astype(F32)
    with np.errstate(divide='ignore', invalid='ignore'):
        ov = inter / ua
    ok = (iw > 0) & (ih > 0)
    return np.where(ok, ov, F32(0)).astype(F32)


def split_tube_into_boxes(tube, T=None):
    """boxes.py:26-57.  dtype behaviour is part of the semantics: ``scores`` is a
    float64 (N,0) array and ``box_rep`` float64 zeros, so every part returned for a
    score-less tube is PROMOTED TO FLOAT64 by the final hstack (:56) — which is why
    tube_transform runs in fp64 while the 2-D bbox_transform runs in fp32."""
    N = tube.shape[0]
    if tube.shape[1] % 4 == 0:
        scores = np.zeros((N, 0))
    elif (tube.shape[1] - 1) % 4 == 0:
        scores = tube[:, (-1,)]
        tube = tube[:, :-1]
    else:
        raise ValueError('Invalid tube dimensions {}'.format(tube.shape))
    T = T or tube.shape[-1] // 4
    boxes = []
    if 4 * T != tube.shape[-1]:
        assert tube.shape[-1] % (4 * T) == 0
        num_classes = tube.shape[-1] // (4 * T)
        for t in range(T):
            box_rep = np.zeros((N, 4 * num_classes))
            for cid in range(num_classes):
                box_rep[:, cid * 4:(cid + 1) * 4] = \
                    tube[:, cid * 4 * T:(cid + 1) * 4 * T][:, t * 4:(t + 1) * 4]
            boxes.append(box_rep)
    else:
        for t in range(T):
            boxes.append(tube[..., t * 4:(t + 1) * 4])
    boxes = [np.hstack((box, scores)) for box in boxes]
    return boxes, T


def bbox_overlaps(boxes, query_boxes):
    """boxes.py:60-69: mean over the T frames of the per-frame IoU.
    np.mean over a stacked leading axis of fp32 = sequential fp32 adds, / T."""
    parts, _ = split_tube_into_boxes(np.asarray(boxes))
    qparts, _ = split_tube_into_boxes(np.asarray(query_boxes))
    acc = None
    for p, q in zip(parts, qparts):
        o = bbox_overlaps_2d(p.astype(F32, copy=False), q.astype(F32, copy=False))
        acc = o if acc is None else (acc + o).astype(F32)
    return (acc / F32(len(parts))).astype(F32)


def boxes_area(boxes):
    """boxes.py:72-78."""
    w = (boxes[:, 2::4] - boxes[:, 0::4] + 1)
    h = (boxes[:, 3::4] - boxes[:, 1::4] + 1)
    return np.mean(w * h, axis=1)


def bbox_transform(boxes, deltas, weights=(1.0, 1.0, 1.0, 1.0)):
    """boxes.py:141-183 (+ tube dispatch :148-149).  All arithmetic in
    deltas.dtype (fp32 on the hot path); BBOX_XFORM_CLIP is cast to that dtype
    (numpy-1.14 value-based casting of the np.float64 scalar)."""
    if boxes.shape[1] > 4:
        return tube_transform(boxes, deltas, weights)
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=deltas.dtype)
    dt = deltas.dtype.type
    boxes = boxes.astype(deltas.dtype, copy=False)
    widths = boxes[:, 2] - boxes[:, 0] + dt(1.0)
    heights = boxes[:, 3] - boxes[:, 1] + dt(1.0)
    ctr_x = boxes[:, 0] + dt(0.5) * widths
    ctr_y = boxes[:, 1] + dt(0.5) * heights
    wx, wy, ww, wh = [dt(w) for w in weights]
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = deltas[:, 2::4] / ww
    dh = deltas[:, 3::4] / wh
    dw = np.minimum(dw, dt(BBOX_XFORM_CLIP))
    dh = np.minimum(dh, dt(BBOX_XFORM_CLIP))
    pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
    pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
    pred_w = np.exp(dw) * widths[:, None]
    pred_h = np.exp(dh) * heights[:, None]
    pred = np.zeros(deltas.shape, dtype=deltas.dtype)
    pred[:, 0::4] = pred_ctr_x - dt(0.5) * pred_w
    pred[:, 1::4] = pred_ctr_y - dt(0.5) * pred_h
    pred[:, 2::4] = pred_ctr_x + dt(0.5) * pred_w
    pred[:, 3::4] = pred_ctr_y + dt(0.5) * pred_h
    return pred


def tube_transform(boxes, deltas, weights):
    """boxes.py:186-202.  Parts are float64 (see split_tube_into_boxes), so the
    per-frame transforms run in fp64 and are rounded once when stored into the
    deltas.dtype (fp32) result."""
    boxes_parts, T = split_tube_into_boxes(boxes)
    deltas_parts, _ = split_tube_into_boxes(deltas, T)
    all_tx = [bbox_transform(b, d, weights) for b, d in zip(boxes_parts, deltas_parts)]
    ncls = all_tx[0].shape[-1] // 4
    res = np.zeros(deltas.shape, dtype=deltas.dtype)
    for cid in range(ncls):
        for t in range(T):
            res[:, cid * 4 * T + t * 4: cid * 4 * T + (t + 1) * 4] = \
                all_tx[t][:, cid * 4:(cid + 1) * 4]
    return res


def clip_tiled_boxes(boxes, im_shape):
    """boxes.py:243-253 (in place, like the reference). im_shape=[h, w]."""
    dt = boxes.dtype.type
    h1 = dt(dt(im_shape[0]) - dt(1))
    w1 = dt(dt(im_shape[1]) - dt(1))
    boxes[:, 0::4] = np.maximum(np.minimum(boxes[:, 0::4], w1), dt(0))
    boxes[:, 1::4] = np.maximum(np.minimum(boxes[:, 1::4], h1), dt(0))
    boxes[:, 2::4] = np.maximum(np.minimum(boxes[:, 2::4], w1), dt(0))
    boxes[:, 3::4] = np.maximum(np.minimum(boxes[:, 3::4], h1), dt(0))
    return boxes


def score_order(scores):
    """``scores.argsort()[::-1]`` for DISTINCT scores (the reference's order is
    undefined under ties: numpy's default sort is unstable).  Tie rule used by
    oracle AND device so tests stay deterministic: descending score, then
    descending original index (what reversing a stable ascending sort gives)."""
    return np.argsort(scores, kind='stable')[::-1]


def nms_2d(dets, thresh):
    """cython_nms.pyx:37-87.  dets (N,5) f32.  Greedy; box j is suppressed when
    ovr >= thresh; returns surviving ORIGINAL indices in ascending order."""
    dets = np.ascontiguousarray(dets, dtype=F32)
    thresh = F32(thresh)
    x1, y1, x2, y2, sc = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3], dets[:, 4]
    one = F32(1)
    areas = (x2 - x1 + one) * (y2 - y1 + one)
    order = score_order(sc)
    n = dets.shape[0]
    suppressed = np.zeros(n, dtype=bool)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(F32(0), xx2 - xx1 + one)
        h = np.maximum(F32(0), yy2 - yy1 + one)
        inter = w * h
        with np.errstate(divide='ignore', invalid='ignore'):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr >= thresh]] = True
    return np.where(~suppressed)[0]


def nms_tubes(dets, thresh):
    """py_cpu_nms_tubes.py:17-53.  dets (N,4T+1) f32.  IoU summed over frames
    with sequential fp32 adds, /T; keeps ovT <= thresh; returns keep in score
    order (python list of indices)."""
    dets = np.ascontiguousarray(dets, dtype=F32)
    T = (dets.shape[1] - 1) // 4
    one = F32(1)
    areas = [(dets[:, 4 * t + 2] - dets[:, 4 * t + 0] + one) *
             (dets[:, 4 * t + 3] - dets[:, 4 * t + 1] + one) for t in range(T)]
    order = score_order(dets[:, -1])
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        rest = order[1:]
        ovT = np.zeros(rest.shape, dtype=F32)
        for t in range(T):
            xx1 = np.maximum(dets[i, 4 * t + 0], dets[rest, 4 * t + 0])
            yy1 = np.maximum(dets[i, 4 * t + 1], dets[rest, 4 * t + 1])
            xx2 = np.minimum(dets[i, 4 * t + 2], dets[rest, 4 * t + 2])
            yy2 = np.minimum(dets[i, 4 * t + 3], dets[rest, 4 * t + 3])
            w = np.maximum(F32(0), xx2 - xx1 + one)
            h = np.maximum(F32(0), yy2 - yy1 + one)
            inter = w * h
            with np.errstate(divide='ignore', invalid='ignore'):
                ovr = inter / (areas[t][i] + areas[t][rest] - inter)
            ovT = (ovT + ovr).astype(F32)
        ovT = ovT / F32(T)
        order = rest[np.where(ovT <= F32(thresh))[0]]
    return keep


def nms(dets, thresh):
    """nms_wrapper.py:49-57 dispatch."""
    if dets.shape[0] == 0:
        return []
    if dets.shape[1] > 5:
        return nms_tubes(dets, thresh)
    return nms_2d(dets, thresh)
