"""Oracle: box-head post-processing (TEST INFRASTRUCTURE ONLY).

Restates lib/core/test.py:
  im_detect_bbox decode      :211-252   decode_boxes
  box_results_with_nms_and_limit  :750-806
Softmax stands in for Caffe2's Softmax op (un-vendored; parity unpinned)."""
import numpy as np

from . import boxes as obox

F32 = np.float32


def softmax(logits):
    z = logits - logits.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True)).astype(F32)


def decode_boxes(rois, box_deltas, im_scale, im_shape, weights=(10., 10., 5., 5.)):
    """:211-237.  rois (R, 4T+1) fp32 (col 0 batch idx) in blob coordinates."""
    boxes = (rois[:, 1:] / F32(im_scale)).astype(F32)          # numpy-1.14: fp32 / python float -> fp32
    pred = obox.bbox_transform(boxes, box_deltas, weights)
    return obox.clip_tiled_boxes(pred, im_shape)


def box_results_with_nms_and_limit(scores, boxes, num_classes=2, score_thresh=0.05, nms_thresh=0.5,
                                   dets_per_im=100):
    """:750-806 (hard NMS, no voting)."""
    T = boxes.shape[-1] // (num_classes * 4)
    cls_boxes = [[] for _ in range(num_classes)]
    for j in range(1, num_classes):
        inds = np.where(scores[:, j] > F32(score_thresh))[0]
        dets_j = np.hstack((boxes[inds, j * 4 * T:(j + 1) * 4 * T], scores[inds, j][:, None])).astype(F32, copy=False)
        keep = obox.nms(dets_j, nms_thresh)
        cls_boxes[j] = dets_j[np.asarray(keep, dtype=np.int64), :] if len(keep) else dets_j[:0]
    if dets_per_im > 0:
        image_scores = np.hstack([cls_boxes[j][:, -1] for j in range(1, num_classes)])
        if len(image_scores) > dets_per_im:
            th = np.sort(image_scores)[-dets_per_im]
            for j in range(1, num_classes):
                cls_boxes[j] = cls_boxes[j][np.where(cls_boxes[j][:, -1] >= th)[0], :]
    im_results = np.vstack([cls_boxes[j] for j in range(1, num_classes)])
    return im_results[:, -1], im_results[:, :-1], cls_boxes
