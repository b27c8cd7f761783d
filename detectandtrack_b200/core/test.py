"""Per-clip inference entry point with the reference's signature and return containers
(lib/core/test.py:897-958): ``im_detect_all(model, im, box_proposals, timers)``.

``im`` is the reference's list of T BGR uint8 frames (HxWx3).  Everything between the H2D copy
of the frames and the D2H copy of the (<=100) detections runs on the device
(modeling/engine.py); test-time augmentation (COMPETITION_MODE) is outside the hot path."""
from collections import defaultdict

import numpy as np

from .config import cfg
from ..utils.timer import Timer


def im_detect_all(model, im, box_proposals=None, timers=None):
    import torch
    if timers is None:
        timers = defaultdict(Timer)
    if box_proposals is not None:
        raise NotImplementedError('precomputed proposals are not on the hot path (FASTER_RCNN models only)')
    if cfg.TEST.COMPETITION_MODE:
        raise NotImplementedError('test-time augmentation (TEST.COMPETITION_MODE) is not on the hot path')
    if not isinstance(im, (list, tuple)):
        im = [im]
    T = cfg.VIDEO.NUM_FRAMES if cfg.MODEL.VIDEO_ON else 1
    assert len(im) == T, 'expected {} frames, got {}'.format(T, len(im))
    timers['im_detect_bbox'].tic()
    frames = torch.from_numpy(np.ascontiguousarray(np.stack(im)[None])).cuda(non_blocking=True)
    res = model.engine.detect(frames)[0]
    boxes = res['boxes'].cpu().numpy()
    keyps = res['keyps'].cpu().numpy() if res['keyps'] is not None else None
    timers['im_detect_bbox'].toc()
    num_classes = cfg.MODEL.NUM_CLASSES
    cls_boxes = [[] for _ in range(num_classes)]
    cls_boxes[1] = boxes
    cls_keyps = None
    if cfg.MODEL.KEYPOINTS_ON and boxes.shape[0] > 0:
        cls_keyps = [[] for _ in range(num_classes)]
        cls_keyps[1] = [keyps[i] for i in range(keyps.shape[0])]
    return cls_boxes, None, cls_keyps
