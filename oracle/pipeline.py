"""Oracle: ONE clip through the whole reference inference path on the CPU in fp32 (TEST INFRASTRUCTURE ONLY; also the
measured CPU arm of bench.py): image blob -> conv body -> FPN -> RPN -> GenerateProposals / collect / distribute ->
RoIAlign -> box head -> decode / NMS / limit -> keypoint head -> heat-map decode, following

  lib/core/test.py:897-958 (im_detect_all), :158-252 (im_detect_bbox), :584-627 (im_detect_keypoints),
  :750-806 (box_results_with_nms_and_limit), :865-894 (keypoint_results),
  lib/ops/generate_proposals.py:40-181, lib/ops/collect_and_distribute_fpn_rpn_proposals.py:44-87

with the stand-ins for the un-vendored Caffe2 operators documented in oracle/net.py (parity unpinned for those).
Only the 2-D-head FPN graphs (the benchmarked configs[3] and configs[1]) with frames already at the blob scale
(scale 1.0: the synthetic workload) are covered here; the tube graph is checked stage by stage in the tests."""
import numpy as np

from . import net as onet, proposals as oprop, detections as odet, keypoints as okp
from .graph import OracleSpec


def detect_clip(cfg, blobs, frames, spec=None, want_heatmaps=False, device='cpu', conv_flags=None):
    """frames [T, H, W, 3] uint8 BGR (T = 1 for 2-D models).  Returns dict(cls_boxes [n, 5], keyps [n, 4, K],
    rois [R, 5], heat (n, K, 56, 56) or None, feats = P2..P6 centre-frame maps).
    device='cuda' + conv_flags=dict(tf32=bool, autocast=None|'bf16') is bench.py's cuDNN stand-in leg (the same graph
    and host steps, dense ops through cuDNN on the GPU); the oracle proper is device='cpu'."""
    import contextlib
    import torch
    onet.set_device(device)
    try:
        ctx = contextlib.nullcontext()
        if device != 'cpu':
            flags = conv_flags or {}
            torch.backends.cudnn.allow_tf32 = bool(flags.get('tf32', False))
            torch.backends.cuda.matmul.allow_tf32 = bool(flags.get('tf32', False))
            if flags.get('autocast') == 'bf16':
                ctx = torch.autocast('cuda', dtype=torch.bfloat16)
        with ctx:
            return _detect_clip(cfg, blobs, frames, spec, want_heatmaps, device)
    finally:
        onet.set_device('cpu')


def _detect_clip(cfg, blobs, frames, spec, want_heatmaps, device):
    import torch
    npy = lambda t: t.detach().float().cpu().numpy()
    spec = spec or OracleSpec(cfg)
    assert spec.fpn and not spec.head3d, 'oracle pipeline: FPN graphs with 2-D heads'
    T, H, W = frames.shape[:3]
    stride = int(cfg.FPN.COARSEST_STRIDE)
    hp, wp = (H + stride - 1) // stride * stride, (W + stride - 1) // stride * stride      # blob.py:47-50
    means = np.asarray(cfg.PIXEL_MEANS, np.float32).reshape(1, 1, 1, 3)
    blob = np.zeros((T, hp, wp, 3), np.float32)
    blob[:, :H, :W] = frames.astype(np.float32) - means
    if spec.is3d:
        data = torch.from_numpy(blob)[None].permute(0, 4, 1, 2, 3).contiguous().to(device)   # (1, 3, T, H, W)
    else:
        assert T == 1
        data = torch.from_numpy(blob).permute(0, 3, 1, 2).contiguous().to(device)          # (1, 3, H, W)
    im_info = np.array([hp, wp, 1.0], np.float32)
    R, D = cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.DETECTIONS_PER_IM
    C, K = cfg.MODEL.NUM_CLASSES, cfg.KRCNN.NUM_KEYPOINTS
    with torch.no_grad():
        pyr = onet.fpn(blobs, spec, onet.conv_body(blobs, spec, data))
        feats = [onet.time_pool(p, spec.link, cfg.VIDEO.NUM_FRAMES_MID) for p in pyr][::-1]      # P2..P6
        rois_l, sc_l = [], []
        for l, (lg, dl) in enumerate(onet.rpn_heads_fpn(blobs, spec, feats[::-1])):
            lvl = spec.rpn_levels[l]
            anchors = oprop.generate_anchors(2. ** lvl, (cfg.FPN.RPN_ANCHOR_START_SIZE * 2. ** (lvl - spec.rpn_levels[0]),),
                                             cfg.FPN.RPN_ASPECT_RATIOS)
            probs = npy(torch.sigmoid(lg.float())[0])
            p, s = oprop.generate_proposals(probs, npy(dl[0]), im_info, anchors, 2. ** lvl, cfg.TEST.RPN_PRE_NMS_TOP_N,
                                            cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH, cfg.TEST.RPN_MIN_SIZE)
            rois_l.append(np.hstack([np.zeros((p.shape[0], 1), np.float32), p])); sc_l.append(s)
        rois = oprop.collect(rois_l, sc_l, R)
        nl = len(spec.roi_levels)
        scales = [1. / 2 ** l for l in spec.roi_levels]
        cls, bbox = onet.box_head_2mlp(blobs, onet.roi_features(feats[:nl], scales, rois, cfg.FAST_RCNN.ROI_XFORM_RESOLUTION,
                                                                cfg.FAST_RCNN.ROI_XFORM_SAMPLING_RATIO,
                                                                spec.roi_levels[0], spec.roi_levels[-1]))
        scores = odet.softmax(npy(cls))
        boxes = odet.decode_boxes(rois, npy(bbox), 1.0, (H, W), cfg.MODEL.BBOX_REG_WEIGHTS)
        _, det_boxes, cls_boxes = odet.box_results_with_nms_and_limit(scores, boxes, C, cfg.TEST.SCORE_THRESH, cfg.TEST.NMS, D)
        keyps, heat = None, None
        if cfg.MODEL.KEYPOINTS_ON and det_boxes.shape[0]:
            kr = np.hstack([np.zeros((det_boxes.shape[0], 1), np.float32), det_boxes]).astype(np.float32)
            heat_t, _ = onet.keypoint_head_2d(blobs, onet.roi_features(feats[:nl], scales, kr, cfg.KRCNN.ROI_XFORM_RESOLUTION,
                                                                       cfg.KRCNN.ROI_XFORM_SAMPLING_RATIO,
                                                                       spec.roi_levels[0], spec.roi_levels[-1]),
                                              cfg.KRCNN.NUM_STACKED_CONVS)
            keyps = okp.keypoint_results(npy(heat_t), det_boxes, K)
            heat = npy(heat_t) if want_heatmaps else None
    return dict(cls_boxes=cls_boxes[1], cls_boxes_all=cls_boxes, keyps=keyps, rois=rois, heat=heat, feats=feats)
