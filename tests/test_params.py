"""CPU: parameter inventory, pkl wire format and 2-D -> 3-D weight inflation
(lib/utils/net.py:95-161,164-294)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _cfg3d():
    from test_gpu_engine import _cfg
    return _cfg()


def test_param_inventory_names_and_shapes():
    from detectandtrack_b200.modeling import params as P
    cfg = _cfg3d()
    shapes, spec = P.param_shapes(cfg)
    assert shapes['conv1_w'] == (64, 3, 1, 7, 7)
    assert shapes['res2_0_branch2b_w'] == (64, 64, 1, 3, 3)              # res2: no temporal kernel
    assert shapes['res3_0_branch2b_w'] == (128, 128, 3, 3, 3)            # TIME_KERNEL_DIM.BODY
    assert shapes['res3_0_branch1_w'] == (512, 256, 1, 1, 1)
    assert shapes['fpn_inner_res5_2_sum_w'] == (256, 2048, 1, 1, 1)
    assert shapes['fpn_inner_res4_5_sum_lateral_w'] == (256, 1024, 1, 1, 1)
    assert shapes['fpn_res2_2_sum_w'] == (256, 256, 3, 3, 3)
    assert shapes['conv_rpn_fpn2_w'] == (256, 256, 3, 3) and shapes['rpn_bbox_pred_fpn2_w'] == (12, 256, 1, 1)
    assert shapes['fc6_w'] == (1024, 256 * 7 * 7) and shapes['bbox_pred_w'] == (8, 1024)
    assert shapes['conv_fcn1_w'] == (512, 256, 3, 3) and shapes['kps_score_lowres_w'] == (512, 17, 4, 4)
    assert sum(int(np.prod(s)) for s in shapes.values()) > 80e6
    assert spec.stage_blobs == ['res2_2_sum', 'res3_3_sum', 'res4_5_sum', 'res5_2_sum']


def test_inflate_modes():
    from detectandtrack_b200.modeling import params as P
    w2 = np.arange(2 * 3 * 3 * 3, dtype=np.float32).reshape(2, 3, 3, 3)
    for mode in ('center-only', 'mean-repeat', 'repeat'):
        w3 = P.inflate_weights(w2, (2, 3, 3, 3, 3), mode)
        assert w3.shape == (2, 3, 3, 3, 3)
        if mode == 'center-only':
            assert np.array_equal(w3[:, :, 1], w2) and not w3[:, :, 0].any() and not w3[:, :, 2].any()
        elif mode == 'mean-repeat':
            np.testing.assert_allclose(w3.sum(2), w2, rtol=1e-6)
        else:
            assert all(np.array_equal(w3[:, :, t], w2) for t in range(3))
    with pytest.raises(ValueError):
        P.inflate_weights(w2, (2, 3, 3, 3, 3), 'nope')


def test_weights_file_roundtrip_with_inflation(tmp_path):
    """A 2-D (COCO-style) pkl loads into the 3-D graph through center-only inflation; unknown blobs keep init."""
    from detectandtrack_b200.modeling import params as P
    cfg = _cfg3d()
    cfg.VIDEO.WEIGHTS_INFLATE_MODE = 'center-only'
    blobs, _ = P.random_blobs(cfg, seed=1)
    two_d = {}
    for k, v in blobs.items():
        if k.startswith(('res3_0', 'conv1', 'fc7')):
            two_d[k] = v[:, :, v.shape[2] // 2] if v.ndim == 5 else v          # 4-D filters as a 2-D model stores them
    path = str(tmp_path / 'w.pkl')
    P.save_weights_file(two_d, 'cfg: yaml', path)
    loaded, _ = P.load_weights_file(cfg, path)
    w = loaded['res3_0_branch2b_w']
    assert w.shape == (128, 128, 3, 3, 3)
    assert np.array_equal(w[:, :, 1], two_d['res3_0_branch2b_w']) and not w[:, :, 0].any()
    assert np.array_equal(loaded['fc7_w'], blobs['fc7_w'])
    assert loaded['res4_0_branch2a_w'].shape == blobs['res4_0_branch2a_w'].shape      # missing in file: kept


SHIPPED_3D = '''
MODEL:
  TYPE: keypoint_rcnn
  CONV_BODY: ResNet3D.add_ResNet18_conv4_body
  ROI_HEAD: ResNet3D.add_ResNet18_roi_conv5_head
  NUM_CLASSES: 2
  FASTER_RCNN: True
  KEYPOINTS_ON: True
  VIDEO_ON: True
NUM_GPUS: 8
FAST_RCNN:
  ROI_XFORM_METHOD: RoIAlign
  ROI_XFORM_RESOLUTION: 7
  ROI_XFORM_SAMPLING_RATIO: 2
KRCNN:
  ROI_KEYPOINTS_HEAD: keypoint_rcnn_heads.add_roi_pose_head_v1convX_3d
  NUM_STACKED_CONVS: 8
  NUM_KEYPOINTS: 17
  USE_DECONV_OUTPUT: True
  CONV_INIT: MSRAFill
  CONV_HEAD_DIM: 512
  UP_SCALE: 2
  HEATMAP_SIZE: 56
  ROI_XFORM_METHOD: RoIAlign
  ROI_XFORM_RESOLUTION: 14
  ROI_XFORM_SAMPLING_RATIO: 2
  KEYPOINT_CONFIDENCE: bbox
  NO_3D_DECONV_TIME_TO_CH: True
VIDEO:
  NUM_FRAMES: 3
  TIME_INTERVAL: 1
  WEIGHTS_INFLATE_MODE: center-only
  TIME_KERNEL_DIM: 3
  BODY_HEAD_LINK: ''
  PREDICT_RPN_BOX_VIS: False
TEST:
  DATASET: posetrack_v1.0_val
  SCALES: (256,)
  MAX_SIZE: 333
  NMS: 0.5
  RPN_PRE_NMS_TOP_N: 1000
  RPN_POST_NMS_TOP_N: 1000
  COMPETITION_MODE: False
TRACKING:
  CONF_FILTER_INITIAL_DETS: 0.95
  DISTANCE_METRICS: ('bbox-overlap', 'cnn-cosdist')
  DISTANCE_METRIC_WTS: (1.0, 0.0)
  BIPARTITE_MATCHING_ALGO: 'hungarian'
EVAL:
  EVAL_MPII_KPT_THRESHOLD: 1.95
USE_NCCL: False
OUTPUT_DIR: .
'''


def test_shipped_style_3d_yaml_builds_the_tube_graph(tmp_path):
    """The option set of the reference's shipped 3-D config (configs/video/3d/03_R-18-3D_PTFromCOCO.yaml,
    restated here because /root/reference does not exist on the GPU box) resolves to the tube graph:
    basic-block conv4 body, single-level 3-D RPN with 12 tube anchors, res5 RoI head, 3-D keypoint head."""
    from detectandtrack_b200.core.config import cfg, reset_cfg, cfg_from_file, assert_and_infer_cfg
    from detectandtrack_b200.modeling import params as P
    y = tmp_path / 'c.yaml'
    y.write_text(SHIPPED_3D)
    reset_cfg(); cfg_from_file(str(y)); assert_and_infer_cfg()
    shapes, spec = P.param_shapes(cfg)
    assert spec.block == 'basic' and spec.counts == (2, 2, 2) and not spec.fpn and spec.head3d
    assert spec.T_head == 3 and spec.num_anchors == 12
    assert shapes['conv_rpn_w'] == (256, 256, 3, 3, 3) and shapes['rpn_bbox_pred_1_w'] == (48, 256, 1, 1, 1)
    assert shapes['res5_0_branch1_w'] == (512, 256, 1, 1, 1) and shapes['cls_score_1_w'] == (2, 512, 1, 1, 1)
    assert shapes['conv_fcn1_w'] == (512, 256, 3, 3, 3) and shapes['kps_score_lowres_w'] == (512, 17, 4, 4)
    reset_cfg()


BODIES = ['FPN3D.add_fpn_ResNet50_conv5_body', 'FPN3D.add_fpn_ResNet101_conv5_body', 'FPN3D.add_fpn_ResNet152_conv5_body',
          'FPN.add_fpn_ResNet50_conv5_body', 'FPN.add_fpn_ResNet101_conv5_body', 'ResNet3D.add_ResNet18_conv4_body',
          'ResNet3D.add_ResNet34_conv4_body', 'ResNet3D.add_ResNet50_conv4_body']


@pytest.mark.parametrize('body', BODIES)
def test_graphspec_matches_reference_tables(body):
    """The product's reading of the reference builders (modeling/params.GraphSpec) against oracle/graph_tables.json, which
    tests/golden/gen_golden_graph.py extracted from the reference's own source (ResNet3D.py:334-394, ResNet.py:298-397),
    and against the oracle's independent OracleSpec."""
    from detectandtrack_b200.core.config import cfg, reset_cfg, assert_and_infer_cfg
    from detectandtrack_b200.modeling import params as P
    from oracle import graph as og
    reset_cfg()
    cfg.MODEL.TYPE = 'keypoint_rcnn'; cfg.MODEL.CONV_BODY = body; cfg.MODEL.NUM_CLASSES = 2
    cfg.MODEL.FASTER_RCNN = True; cfg.MODEL.KEYPOINTS_ON = True
    fpn = 'fpn' in body
    cfg.MODEL.VIDEO_ON = body.split('.')[0].endswith('3D')
    cfg.FPN.FPN_ON = fpn; cfg.FPN.MULTILEVEL_ROIS = fpn; cfg.FPN.MULTILEVEL_RPN = fpn
    cfg.KRCNN.USE_DECONV_OUTPUT = True; cfg.KRCNN.UP_SCALE = 2
    cfg.MODEL.ROI_HEAD = 'head_builder.add_roi_2mlp_head' if fpn else 'ResNet3D.add_ResNet18_roi_conv5_head'
    cfg.VIDEO.BODY_HEAD_LINK = 'slice-center' if fpn else ''
    cfg.VIDEO.NUM_FRAMES = 3; cfg.VIDEO.NUM_FRAMES_MID = 1 if fpn else -1
    for k in ('BODY', 'HEAD_RPN', 'HEAD_KPS', 'HEAD_DET'):
        cfg.VIDEO.TIME_KERNEL_DIM[k] = 3
    if not fpn:
        cfg.KRCNN.ROI_KEYPOINTS_HEAD = 'keypoint_rcnn_heads.add_roi_pose_head_v1convX_3d'; cfg.KRCNN.NO_3D_DECONV_TIME_TO_CH = True
    assert_and_infer_cfg()
    try:
        g, o = P.GraphSpec(cfg), og.OracleSpec(cfg)
        key = body.split('add_fpn_')[-1].split('add_')[-1][:-len('_body')]
        tab = og.tables()['ResNet3D' if g.is3d else 'ResNet']['bodies'][key]
        assert list(g.counts) == tab['counts'] and list(g.dims[:len(g.counts) + 1]) == tab['dims']
        assert g.block == ('bottleneck' if tab['trans_func'].startswith('bottleneck') else 'basic')
        for a in ('counts', 'block', 'tk_body', 'stride_1x1', 'stage_blobs', 'num_anchors', 'link', 'head3d', 'T_head', 'fpn', 'is3d'):
            assert getattr(g, a) == getattr(o, a), (a, getattr(g, a), getattr(o, a))
        assert tuple(g.dims[:len(g.counts) + 1]) == tuple(o.dims)
        if fpn:
            assert g.rpn_levels == o.rpn_levels and g.roi_levels == o.roi_levels
            lv = og.tables()['fpn_levels'][key]
            assert g.stage_blobs == lv['blobs'][::-1]
        else:
            head = og.tables()['ResNet3D']['roi_conv5_heads']['ResNet18']
            shapes, _ = P.param_shapes(cfg)
            assert shapes['res5_%d_branch2b_w' % (head['block_counts'] - 1)][0] == head['dim_out']
            assert 'res5_%d_branch2b_w' % head['block_counts'] not in shapes
    finally:
        reset_cfg()


def test_inflate_weights_matches_reference_goldens():
    """modeling/params.inflate_weights against outputs of the REFERENCE's own lib/utils/net.py:95-161, source-executed
    by tests/golden/gen_golden_graph.py for every VIDEO.WEIGHTS_INFLATE_MODE and time sizes 1 / 3 (bit-equal; the
    random mode with the same numpy seed)."""
    from detectandtrack_b200.modeling import params as P
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'inflate_weights.npz'))
    w2d = g['w2d']
    for mode in ('mean-repeat', 'repeat', 'center-only', 'center-only-rest-rand'):
        for kt in (1, 3):
            np.random.seed(1234)
            got = P.inflate_weights(w2d, (6, 4, kt, 3, 3), mode)
            ref = g['%s_%d' % (mode, kt)]
            assert got.shape == ref.shape and np.array_equal(np.asarray(got, np.float32), ref), (mode, kt)
    assert np.array_equal(P.inflate_weights(w2d, w2d.shape, 'center-only'), g['same_rank'])
