"""Drop-in for lib/utils/cython_bbox.pyx: ``bbox_overlaps(boxes, query_boxes)``
on host numpy fp32 arrays, computed on the GPU (csrc/boxes.cu, bit-exact)."""
import numpy as np

from ..ops import box_ops


def bbox_overlaps(boxes, query_boxes):
    """(N,4) f32, (K,4) f32 -> (N,K) f32   [cython_bbox.pyx:16-56]"""
    import torch
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    query_boxes = np.ascontiguousarray(query_boxes, dtype=np.float32)
    if boxes.ndim != 2 or query_boxes.ndim != 2 or boxes.shape[1] < 4 or query_boxes.shape[1] < 4:
        raise RuntimeError('bbox_overlaps: expected (N,4) and (K,4) arrays, got %s and %s'
                           % (boxes.shape, query_boxes.shape))
    out = box_ops.bbox_overlaps(torch.from_numpy(boxes[:, :4].copy()).cuda(),
                                torch.from_numpy(query_boxes[:, :4].copy()).cuda(), T=1)
    return out.cpu().numpy()
