"""Dataset-level inference driver with the reference's entry points
(lib/core/test_engine.py): ``initialize_model_from_cfg`` (:50-74),
``get_roidb_and_dataset`` (:77-103), ``test_net`` (:124-206),
``multi_gpu_test_net_on_dataset`` (:278-308), ``test_net_on_dataset`` (:311-333); same
``detections.pkl`` schema (dict(all_boxes, all_segms, all_keyps, cfg)).

Datasets: the reference's JsonDataset needs pycocotools and the PoseTrack images (absent
here).  Two dataset kinds are understood:
  * ``synthetic_<V>x<F>[_HxW]``: V videos x F clips of seeded uniform-noise frames
     — the benchmark / smoke data;
  * any other name is looked up as a JSON roidb file path (list of entries with
    'image' (path or list of T paths), 'height', 'width'), frames read with cv2.
"""
import datetime
import json
import logging
import os
import pickle
from collections import defaultdict

import numpy as np
import yaml

from .config import cfg, get_output_dir
from .test import im_detect_all, get_pipeline
from .tracking_engine import run_posetrack_tracking
from ..modeling import model_builder, params as P
from ..utils import subprocess as subprocess_utils
from ..utils.subprocess import _plain
from ..utils.timer import Timer

logger = logging.getLogger(__name__)


def initialize_model_from_cfg(dtype=None):
    """:50-74.  TEST.WEIGHTS '' or 'random' -> seeded synthetic weights (cfg.RNG_SEED).
    dtype None -> cfg.TEST.PRECISION (default 'bf16x3', the mode that meets the fp32 reference to 1e-3)."""
    dtype = dtype or cfg.TEST.PRECISION
    if cfg.TEST.WEIGHTS in ('', 'random'):
        blobs, _ = P.random_blobs(cfg)
    else:
        blobs, _ = P.load_weights_file(cfg, cfg.TEST.WEIGHTS)
    return model_builder.create(cfg.MODEL.TYPE, train=False, blobs=blobs, dtype=dtype)


class SyntheticDataset(object):
    def __init__(self, name):
        self.name = name
        parts = name.split('_')
        v, f = parts[1].split('x')
        self.V, self.F = int(v), int(f)
        self.H, self.W = (int(x) for x in parts[2].split('x')) if len(parts) > 2 else (800, 1333)

    def get_roidb(self, gt=False):
        T = cfg.VIDEO.NUM_FRAMES if cfg.MODEL.VIDEO_ON else 1
        roidb = []
        for v in range(self.V):
            for f in range(self.F):
                paths = ['synthetic/vid%03d/%06d.jpg' % (v, max(0, min(self.F - 1, f + d - T // 2))) for d in range(T)]
                roidb.append(dict(image=paths if T > 1 else paths[0], height=self.H, width=self.W,
                                  frame_id=f + 1, seed=1000003 * v + f, synthetic=True))
        return roidb


class JsonListDataset(object):
    def __init__(self, name):
        self.name = name
        with open(name) as f:
            self.entries = json.load(f)

    def get_roidb(self, gt=False):
        """Per-frame entries ('image' a path) of a video model are assembled into T-frame clips exactly like the reference's
        json_dataset -> utils/video.get_clip (neighbours at VIDEO.TIME_INTERVAL, clamped at the video ends); entries that
        already carry a frame list pass through."""
        roidb = self.entries
        if cfg.MODEL.VIDEO_ON and roidb and not isinstance(roidb[0]['image'], (list, tuple)):
            from ..utils import video
            for e in roidb:
                for k in ('boxes', 'gt_keypoints', 'tracks', 'gt_classes', 'is_crowd'):
                    if k in e and not isinstance(e[k], np.ndarray):
                        e[k] = np.asarray(e[k], dtype={'boxes': np.float32, 'is_crowd': bool}.get(k, np.int32))
            roidb = video.get_clip(roidb)
        return roidb


def get_dataset(name):
    return SyntheticDataset(name) if name.startswith('synthetic_') else JsonListDataset(name)


def get_roidb_and_dataset(ind_range, include_gt=False):
    dataset = get_dataset(cfg.TEST.DATASET)
    roidb = dataset.get_roidb(gt=include_gt)
    total = len(roidb)
    if ind_range is not None:
        start, end = ind_range
        roidb = roidb[start:end]
    else:
        start, end = 0, total
    return roidb, dataset, start, end, total


_NOISE_BANK = {}


def _synthetic_frame(h, w, seed, t):
    """Seeded uniform-noise frame as a window of a per-process noise bank (a fresh 3 MB RandomState draw per frame
    costs ~10 ms; a window is free and still deterministic in (h, w, seed, t))."""
    bank = _NOISE_BANK.get((h, w))
    if bank is None:
        bank = _NOISE_BANK[(h, w)] = np.random.RandomState(20260923).randint(0, 256, (h + 64, w + 64, 3)).astype(np.uint8)
    k = (seed * 2654435761 + t * 40503) & 0xffffffff
    oy, ox = (k >> 8) % 64, (k >> 16) % 64
    return bank[oy:oy + h, ox:ox + w]


def read_image_video(entry):
    """lib/utils/image.py:65-79: list of T BGR uint8 frames."""
    T = cfg.VIDEO.NUM_FRAMES if cfg.MODEL.VIDEO_ON else 1
    if entry.get('synthetic'):
        return [_synthetic_frame(entry['height'], entry['width'], entry['seed'], t) for t in range(T)]
    import cv2
    paths = entry['image'] if isinstance(entry['image'], list) else [entry['image']]
    ims = [cv2.imread(p) for p in paths]
    assert all(im is not None for im in ims), 'could not read {}'.format(paths)
    return ims


def empty_results(num_classes, num_images):
    mk = lambda: [[[] for _ in range(num_images)] for _ in range(num_classes)]
    return mk(), mk(), mk()


def extend_results(index, all_res, im_res):
    for j in range(1, len(im_res)):
        all_res[j][index] = im_res[j]


def _dump(obj, path):
    with open(path, 'wb') as f:
        pickle.dump(obj, f, pickle.HIGHEST_PROTOCOL)


def test_net(ind_range=None):
    assert cfg.TEST.DATASET != '', 'TEST.DATASET must be set to the dataset name to test'
    output_dir = get_output_dir(training=False)
    roidb, dataset, start_ind, end_ind, total_num_images = get_roidb_and_dataset(ind_range)
    model = initialize_model_from_cfg()
    num_images = len(roidb)
    all_boxes, all_segms, all_keyps = empty_results(cfg.MODEL.NUM_CLASSES, num_images)
    timers = defaultdict(Timer)
    detect_roidb(model, roidb, all_boxes, all_keyps, timers,
                 log=lambda i, ave: logger.info('im_detect: range [%d, %d] of %d: %d/%d %.3fs (eta: %s)', start_ind + 1, end_ind,
                                                total_num_images, start_ind + i + 1, start_ind + num_images, ave,
                                                str(datetime.timedelta(seconds=int(ave * (num_images - i - 1))))))
    det_name = 'detection_range_%s_%s.pkl' % tuple(ind_range) if ind_range is not None else 'detections.pkl'
    det_file = os.path.join(output_dir, det_name)
    _dump(dict(all_boxes=all_boxes, all_segms=all_segms, all_keyps=all_keyps, cfg=yaml.safe_dump(_plain(cfg))), det_file)
    logger.info('Wrote detections to: %s', os.path.abspath(det_file))
    return all_boxes, all_segms, all_keyps


def _all_jpeg(entries):
    for e in entries:
        if e.get('synthetic'):
            return False
        paths = e['image'] if isinstance(e['image'], (list, tuple)) else [e['image']]
        if not all(str(p).lower().endswith(('.jpg', '.jpeg')) for p in paths):
            return False
    return True


def detect_roidb(model, roidb, all_boxes, all_keyps, timers=None, log=None):
    """The loop of test_net (:142-165) on the batched device step: runs of same-sized entries go through one
    ClipPipeline (cfg.TEST.CLIPS_PER_STEP clips per captured graph replay; frames are read by loader threads straight
    into pinned memory while the previous step computes).  Results land at the entry's index, as in the reference."""
    timers = timers if timers is not None else defaultdict(Timer)
    T = cfg.VIDEO.NUM_FRAMES if cfg.MODEL.VIDEO_ON else 1
    B = max(1, int(cfg.TEST.CLIPS_PER_STEP))
    i0, n = 0, len(roidb)
    while i0 < n:
        hw = (roidb[i0]['height'], roidb[i0]['width'])
        i1 = i0
        while i1 < n and (roidb[i1]['height'], roidb[i1]['width']) == hw:
            i1 += 1
        pipe = get_pipeline(model, min(B, i1 - i0), T, hw[0], hw[1])
        base = i0

        def fill(i, dst):
            ims = read_image_video(roidb[base + i])
            for t in range(T):
                assert ims[t].shape == dst[t].shape, 'roidb entry {}: frame size {} != height/width fields {}'.format(
                    base + i, ims[t].shape, dst[t].shape)
                np.copyto(dst[t], ims[t])

        def on_result(i, cls_boxes_i, cls_segms_i, cls_keyps_i):
            extend_results(base + i, all_boxes, cls_boxes_i)
            if cls_keyps_i is not None:
                extend_results(base + i, all_keyps, cls_keyps_i)
            timers['im_detect_bbox'].toc()
            timers['im_detect_bbox'].tic()
            if log is not None and (base + i) % 10 == 0:
                log(base + i, timers['im_detect_bbox'].average_time)
        fill_device = None
        if cfg.TEST.DEVICE_JPEG_DECODE and _all_jpeg(roidb[i0:i1]):
            from ..ops import image_ops
            if not image_ops.jpeg_available():
                raise RuntimeError('TEST.DEVICE_JPEG_DECODE needs nvJPEG (libnvjpeg.so.12), which is not available here')

            def fill_device(i, dev_dst, stream):      # loader thread: file bytes -> frames decoded straight into the device buffer
                e = roidb[base + i]
                paths = e['image'] if isinstance(e['image'], (list, tuple)) else [e['image']]
                image_ops.jpeg_decode([open(p, 'rb').read() for p in paths], hw[0], hw[1], out=dev_dst, stream=stream)
        timers['im_detect_bbox'].tic()
        pipe.run(i1 - i0, fill, on_result, fill_device=fill_device)
        i0 = i1


def multi_gpu_test_net_on_dataset(num_images, output_dir):
    binary = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tools', 'test_net.py')
    assert os.path.exists(binary), 'Binary {} not found'.format(binary)
    outputs = subprocess_utils.process_in_parallel('detection', num_images, binary, output_dir)
    C = cfg.MODEL.NUM_CLASSES
    all_boxes, all_segms, all_keyps = [[] for _ in range(C)], [[] for _ in range(C)], [[] for _ in range(C)]
    for det in outputs:
        for j in range(1, C):
            all_boxes[j] += det['all_boxes'][j]
            all_segms[j] += det['all_segms'][j]
            all_keyps[j] += det['all_keyps'][j]
    det_file = os.path.join(output_dir, 'detections.pkl')
    _dump(dict(all_boxes=all_boxes, all_segms=all_segms, all_keyps=all_keyps, cfg=yaml.safe_dump(_plain(cfg))), det_file)
    logger.info('Wrote detections to: %s', os.path.abspath(det_file))
    return all_boxes, all_segms, all_keyps


def test_net_on_dataset(multi_gpu=False):
    output_dir = get_output_dir(training=False)
    dataset = get_dataset(cfg.TEST.DATASET)
    timer = Timer()
    timer.tic()
    if multi_gpu:
        res = multi_gpu_test_net_on_dataset(len(dataset.get_roidb()), output_dir)
    else:
        res = test_net()
    timer.toc()
    logger.info('Total inference time: %.3fs', timer.average_time)
    if os.path.basename(str(dataset.name)).lower().startswith(('posetrack', 'kinetics', 'synthetic')) or cfg.TRACKING.get('RUN_AFTER_TEST', False):
        roidb, _, _, _, _ = get_roidb_and_dataset(None)
        run_posetrack_tracking(output_dir, roidb)
    return res
