"""Host mirror of the lib/utils/boxes.py entry points that sit on the hot path.
Tube-aware IoU runs on the device; the tiny layout helpers stay numpy."""
import numpy as np

from ..ops import box_ops


def _strip_score(a):
    a = np.asarray(a)
    if a.shape[1] % 4 == 0:
        return a
    if (a.shape[1] - 1) % 4 == 0:
        return a[:, :-1]
    raise ValueError('Invalid tube dimensions {}'.format(a.shape))


def bbox_overlaps(boxes, query_boxes):
    """lib/utils/boxes.py:60-69: mean-over-frames IoU of (N,4T[+1]) vs (K,4T[+1])."""
    import torch
    b = np.ascontiguousarray(_strip_score(boxes), dtype=np.float32)
    q = np.ascontiguousarray(_strip_score(query_boxes), dtype=np.float32)
    T = b.shape[1] // 4
    if q.shape[1] // 4 != T:
        raise RuntimeError('bbox_overlaps: tube lengths differ: %s vs %s' % (b.shape, q.shape))
    return box_ops.bbox_overlaps(torch.from_numpy(b).cuda(), torch.from_numpy(q).cuda(), T=T).cpu().numpy()


def boxes_area(boxes):
    """lib/utils/boxes.py:72-78."""
    w = (boxes[:, 2::4] - boxes[:, 0::4] + 1)
    h = (boxes[:, 3::4] - boxes[:, 1::4] + 1)
    areas = np.mean(w * h, axis=1)
    assert np.all(areas >= 0), 'Negative areas founds'
    return areas
