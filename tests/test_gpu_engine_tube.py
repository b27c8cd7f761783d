"""GPU parity of the shipped 3-D config (configs/video/3d/03_R-18-3D_PTFromCOCO.yaml semantics: R18
conv4 3-D body, 3-D RPN with tube anchors, res5 RoI head, 3-D keypoint head) against the torch-fp32
oracle graph, stage by stage with teacher forcing.  Modes: bf16x3 (the parity mode) and tf32x3 must meet the north
star's 1e-3 (asserted at 5e-4 .. 1e-3); tf32 keeps its single-MMA bars (2e-3 .. 2.5e-3, see test_gpu_engine.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import net as onet
from oracle import proposals as oprop


def _cfg():
    from detectandtrack_b200.core.config import cfg, reset_cfg, assert_and_infer_cfg
    reset_cfg()
    cfg.MODEL.TYPE = 'keypoint_rcnn'
    cfg.MODEL.CONV_BODY = 'ResNet3D.add_ResNet18_conv4_body'
    cfg.MODEL.ROI_HEAD = 'ResNet3D.add_ResNet18_roi_conv5_head'
    cfg.MODEL.NUM_CLASSES = 2
    cfg.MODEL.FASTER_RCNN = True; cfg.MODEL.KEYPOINTS_ON = True; cfg.MODEL.VIDEO_ON = True
    cfg.FAST_RCNN.ROI_XFORM_METHOD = 'RoIAlign'; cfg.FAST_RCNN.ROI_XFORM_RESOLUTION = 7; cfg.FAST_RCNN.ROI_XFORM_SAMPLING_RATIO = 2
    cfg.KRCNN.ROI_KEYPOINTS_HEAD = 'keypoint_rcnn_heads.add_roi_pose_head_v1convX_3d'
    cfg.KRCNN.NUM_STACKED_CONVS = 8; cfg.KRCNN.NUM_KEYPOINTS = 17; cfg.KRCNN.USE_DECONV_OUTPUT = True
    cfg.KRCNN.CONV_HEAD_DIM = 512; cfg.KRCNN.UP_SCALE = 2; cfg.KRCNN.HEATMAP_SIZE = 56
    cfg.KRCNN.ROI_XFORM_RESOLUTION = 14; cfg.KRCNN.ROI_XFORM_SAMPLING_RATIO = 2; cfg.KRCNN.NO_3D_DECONV_TIME_TO_CH = True
    cfg.VIDEO.NUM_FRAMES = 3; cfg.VIDEO.TIME_INTERVAL = 1
    for k in ('BODY', 'HEAD_RPN', 'HEAD_KPS', 'HEAD_DET'):
        cfg.VIDEO.TIME_KERNEL_DIM[k] = 3
    cfg.VIDEO.BODY_HEAD_LINK = ''
    cfg.TEST.SCALES = (128,); cfg.TEST.MAX_SIZE = 192
    cfg.TEST.NMS = 0.5; cfg.TEST.RPN_PRE_NMS_TOP_N = 1000; cfg.TEST.RPN_POST_NMS_TOP_N = 200
    cfg.TEST.COMPETITION_MODE = False
    assert_and_infer_cfg()
    return cfg


@pytest.fixture(scope='module')
def setup():
    import torch
    from detectandtrack_b200.modeling import params as P
    cfg = _cfg()
    blobs, spec = P.random_blobs(cfg, seed=5)
    rng = np.random.RandomState(1)
    frames = rng.randint(0, 256, (1, 3, 128, 160, 3)).astype(np.uint8)
    means = np.asarray(cfg.PIXEL_MEANS, np.float32).reshape(1, 1, 1, 1, 3)
    data = torch.from_numpy(frames.astype(np.float32) - means).permute(0, 4, 1, 2, 3).contiguous()
    with torch.no_grad():
        feat = onet.conv_body(blobs, spec, data)[spec.stage_blobs[-1]]            # (1, 256, 3, 8, 10)
        lg, dl = onet.rpn_heads_3d(blobs, spec, feat)
    return dict(cfg=cfg, blobs=blobs, spec=spec, frames=frames, feat=feat, lg=lg, dl=dl)


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


MODES = [('bf16x3', 5e-4, 5e-4, 1e-3), ('tf32x3', 5e-4, 5e-4, 1e-3), ('tf32', 2e-3, 1.5e-3, 2.5e-3)]


def _dev_feat(eng, feat):
    """oracle NCTHW feature -> the engine's NDHWC storage (split pairs in the x3 modes)."""
    from detectandtrack_b200.ops import conv as cv
    x = feat.permute(0, 2, 3, 4, 1).contiguous().cuda()
    return cv.split_for(eng.dtype, x) if eng.x3 else x.to(eng.act_dtype)


@pytest.mark.parametrize('mode,tol_feat,tol_rpn,tol_head', MODES)
def test_body_and_rpn_3d(setup, mode, tol_feat, tol_rpn, tol_head):
    import torch
    from detectandtrack_b200.modeling.engine import DetectionEngine
    eng = DetectionEngine(setup['cfg'], setup['blobs'], setup['spec'], dtype=mode)
    feats, im_info, scale = eng.forward_features(torch.from_numpy(setup['frames']).cuda())
    got = eng.plain(feats[0]).permute(0, 4, 1, 2, 3).float().cpu()
    assert got.shape == setup['feat'].shape and _rel(got, setup['feat']) <= tol_feat
    # RPN head on the oracle's feature (teacher forcing): per-frame outputs -> time pooled / folded
    x = _dev_feat(eng, setup['feat'])
    h = eng.rpn_conv(x)
    B, T, H, W, _ = h.shape
    o = torch.empty((B, T, H, W, eng.rpn_out_ld), dtype=torch.float32, device='cuda')
    eng.rpn_out(h, out_f32=True, out=o)
    A = setup['spec'].num_anchors
    lg = o[..., :A].mean(dim=1).permute(0, 3, 1, 2).cpu()
    assert _rel(lg, setup['lg']) <= tol_rpn
    dl = o[..., A:5 * A].view(B, T, H, W, A, 4).permute(0, 4, 1, 5, 2, 3).reshape(B, A * T * 4, H, W).cpu()
    assert _rel(dl, setup['dl']) <= tol_rpn
    # device proposals from the oracle's raw head outputs == oracle proposals (tube decode is fp64: bit-exact)
    from detectandtrack_b200.ops import rpn_ops, box_ops
    o2 = torch.zeros_like(o)
    o2[..., :A] = setup['lg'].permute(0, 2, 3, 1)[:, None].expand(B, T, H, W, A)        # equal per-frame logits -> same mean
    o2[..., A:5 * A] = setup['dl'].view(B, A, T, 4, H, W).permute(0, 2, 4, 5, 1, 3).reshape(B, T, H, W, 4 * A)
    anchors = eng.anchors[0]
    props, counts = rpn_ops.rpn_proposals(o2[..., :A], o2[..., A:5 * A], anchors, 16.0, im_info, 1000, 0.0, T, time_major=True)
    probs = torch.sigmoid(setup['lg'])[0].numpy()
    p_ref, s_ref, (pre_b, pre_s), keep_ref = oprop.generate_proposals(
        probs, setup['dl'][0].numpy(), im_info[0].cpu().numpy(), anchors.cpu().numpy(), 16.0, 1000, 200, 0.7, 0, True)
    n = int(counts[0])
    assert n == pre_b.shape[0]
    np.testing.assert_allclose(props[0, :n, -1].cpu().numpy(), pre_s[:, 0], rtol=2e-6, atol=1e-7)
    assert np.array_equal(props[0, :n, :-1].cpu().numpy(), pre_b)


@pytest.mark.parametrize('mode,tol_feat,tol_rpn,tol_head', MODES)
def test_tube_heads_given_oracle_rois(setup, mode, tol_feat, tol_rpn, tol_head):
    import torch
    from detectandtrack_b200.modeling.engine import DetectionEngine
    cfg, blobs, spec = setup['cfg'], setup['blobs'], setup['spec']
    eng = DetectionEngine(cfg, blobs, spec, dtype=mode)
    feat = setup['feat']
    x = _dev_feat(eng, feat)
    rng = np.random.RandomState(9)
    R, T = 48, 3
    x1 = rng.uniform(0, 100, R); y1 = rng.uniform(0, 80, R)
    b0 = np.stack([x1, y1, x1 + rng.uniform(8, 60, R), y1 + rng.uniform(8, 48, R)], 1)
    rois = np.hstack([np.zeros((R, 1))] + [b0 + rng.normal(0, 2, (R, 1)) * (t > 0) for t in range(T)]).astype(np.float32)
    with torch.no_grad():
        rf = onet.roi_features_tube(feat, 1 / 16., rois, 7, 2)
        cls_ref, bb_ref = onet.box_head_conv5_3d(blobs, rf)
        kf = onet.roi_features_tube(feat, 1 / 16., rois[:8], 14, 2)
        heat_ref = onet.keypoint_head_3d(blobs, kf)
    rois_d = torch.from_numpy(rois).cuda()
    xr = eng._roi_feats_tube(x, rois_d, 7, 2)
    assert _rel(eng.plain(xr).permute(0, 4, 1, 2, 3).cpu(), rf) <= (3e-5 if eng.x3 else 6e-4)
    from detectandtrack_b200.ops import dense_ops
    y = xr
    for blk in eng.res5:
        y = eng._run_block(blk, y)
    n, _, hh, ww, ch = y.shape
    y = dense_ops.spatial_mean(y.view(n * T, hh, ww, ch), round_tf32=(mode == 'tf32'), x3=eng.x3)
    o = torch.empty((1, 1, 1, n * T, eng.cls_bbox_ld), dtype=torch.float32, device='cuda')
    eng.cls_bbox(y.view(1, 1, 1, n * T, ch), out_f32=True, out=o)
    cls, bbox = dense_ops.fold_tube_heads(o.view(n * T, eng.cls_bbox_ld), n, T, 2)
    assert _rel(cls.cpu(), cls_ref) <= min(tol_head, 2e-3) and _rel(bbox.cpu(), bb_ref) <= min(tol_head, 2e-3)
    boxes = rois_d[:8, 1:].contiguous()
    xy, heat = eng.keypoint_head([x], boxes, torch.zeros(8, device='cuda'), 1.0, want_heatmaps=True)
    assert heat.shape == heat_ref.shape == (8, 51, 56, 56)
    assert _rel(heat.cpu(), heat_ref) <= tol_head
    assert xy.shape == (8, 4, 51) and torch.isfinite(xy).all()


def test_detect_end_to_end_tubes(setup):
    import torch
    from detectandtrack_b200.modeling.engine import DetectionEngine
    eng = DetectionEngine(setup['cfg'], setup['blobs'], setup['spec'], dtype='bf16')
    res = eng.detect(torch.from_numpy(setup['frames']).cuda())
    b = res[0]['boxes'].cpu().numpy()
    assert b.shape[1] == 13 and b.shape[0] > 0
    assert b[:, :12].min() >= 0 and b[:, 0:12:4].max() <= 159 and b[:, 1:12:4].max() <= 127
    assert res[0]['keyps'].shape == (b.shape[0], 4, 51)


def test_odd_blob_size_single_level_body():
    """Single-level bodies do no /32 blob padding (blob.py:47-50 pads only when FPN is on): a 1280x720 PoseTrack frame
    resizes to 750x1333.  conv1 (7x7/2, pad 3) then yields ceil(h/2) x ceil(w/2); the engine pads the physical blob to
    even dims (zero row / column == the conv's own padding) and keeps im_info at the unpadded size."""
    import torch
    from detectandtrack_b200.core.config import cfg
    from detectandtrack_b200.modeling import params as P
    from detectandtrack_b200.modeling.engine import DetectionEngine
    _cfg()
    try:
        cfg.TEST.SCALES = (127,); cfg.TEST.MAX_SIZE = 161
        blobs, spec = P.random_blobs(cfg, seed=5)
        frames = np.random.RandomState(3).randint(0, 256, (1, 3, 127, 161, 3)).astype(np.uint8)
        means = np.asarray(cfg.PIXEL_MEANS, np.float32).reshape(1, 1, 1, 1, 3)
        data = torch.from_numpy(frames.astype(np.float32) - means).permute(0, 4, 1, 2, 3).contiguous()
        with torch.no_grad():
            feat = onet.conv_body(blobs, spec, data)[spec.stage_blobs[-1]]
        eng = DetectionEngine(cfg, blobs, spec, dtype='bf16x3')
        feats, im_info, scale = eng.forward_features(torch.from_numpy(frames).cuda())
        assert scale == 1.0 and im_info.cpu().numpy().tolist() == [[127.0, 161.0, 1.0]]
        got = eng.plain(feats[0]).permute(0, 4, 1, 2, 3).float().cpu()
        assert got.shape == feat.shape and _rel(got, feat) <= 5e-4
        res = eng.detect(torch.from_numpy(frames).cuda())[0]          # the whole tube path runs at this geometry
        assert res['boxes'].shape[1] == 13
    finally:
        _cfg()
