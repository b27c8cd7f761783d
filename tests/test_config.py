"""CPU: cfg surface of lib/core/config.py (merge rules, legacy key fan-out, CLI list)."""
import os

import numpy as np
import pytest

from detectandtrack_b200.core import config as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _reset():
    C.reset_cfg()
    yield
    C.reset_cfg()


def test_defaults_match_reference_values():
    cfg = C.cfg
    assert cfg.TEST.RPN_PRE_NMS_TOP_N == 12000 and cfg.TEST.DETECTIONS_PER_IM == 100
    assert cfg.MODEL.BBOX_REG_WEIGHTS == (10., 10., 5., 5.)
    assert abs(cfg.BBOX_XFORM_CLIP - np.log(1000. / 16.)) < 1e-12
    assert cfg.PIXEL_MEANS.shape == (1, 1, 3) and cfg.RNG_SEED == 3
    assert cfg.TRACKING.DISTANCE_METRIC_WTS == (1.0, 0.0, 0.0)
    assert cfg.VIDEO.TIME_KERNEL_DIM == {'BODY': 1, 'HEAD_RPN': 1, 'HEAD_KPS': 1, 'HEAD_DET': 1}


def test_shipped_style_yaml_merges(tmp_path):
    y = tmp_path / 'c.yaml'
    y.write_text('''
MODEL:
  TYPE: keypoint_rcnn
  CONV_BODY: ResNet3D.add_ResNet18_conv4_body
  NUM_CLASSES: 2
  FASTER_RCNN: True
  KEYPOINTS_ON: True
  VIDEO_ON: True
VIDEO:
  NUM_FRAMES: 3
  TIME_KERNEL_DIM: 3
  BODY_HEAD_LINK: ''
TEST:
  SCALES: (256,)
  NMS: 0.5
TRACKING:
  DISTANCE_METRICS: ('bbox-overlap', 'cnn-cosdist')
  DISTANCE_METRIC_WTS: (1.0, 0.0)
SOLVER:
  STEPS: [0, 10000, 12000]
OUTPUT_DIR: .
''')
    C.cfg_from_file(str(y))
    C.assert_and_infer_cfg()
    cfg = C.cfg
    assert cfg.VIDEO.TIME_KERNEL_DIM == {'BODY': 3, 'HEAD_RPN': 3, 'HEAD_KPS': 3, 'HEAD_DET': 3}
    assert cfg.TEST.SCALES == (256,) and cfg.RPN.ON is True
    assert cfg.VIDEO.NUM_FRAMES_MID == 3
    assert cfg.TRACKING.DISTANCE_METRICS == ('bbox-overlap', 'cnn-cosdist')


def test_unknown_key_and_type_mismatch(tmp_path):
    y = tmp_path / 'bad.yaml'
    y.write_text('TEST:\n  NOT_A_KEY: 1\n')
    with pytest.raises(KeyError):
        C.cfg_from_file(str(y))
    y.write_text('TEST:\n  NMS: [1, 2]\n')
    with pytest.raises(ValueError):
        C.cfg_from_file(str(y))
    y.write_text('USE_GPU_NMS: True\n')      # deprecated twin exists -> ignored
    C.cfg_from_file(str(y))


def test_cfg_from_list():
    C.cfg_from_list(['TEST.NMS', '0.4', 'NUM_GPUS', '1', 'TEST.WEIGHTS', '/tmp/w.pkl', 'TEST.SCALES', '(800,)'])
    assert C.cfg.TEST.NMS == 0.4 and C.cfg.TEST.WEIGHTS == '/tmp/w.pkl' and C.cfg.TEST.SCALES == (800,)
    with pytest.raises(AssertionError):
        C.cfg_from_list(['TEST.NOPE', '1'])
    with pytest.raises(AssertionError):
        C.cfg_from_list(['NUM_GPUS', 'abc'])


def test_get_output_dir(tmp_path):
    C.cfg.OUTPUT_DIR = str(tmp_path)
    C.cfg.TEST.DATASET = 'posetrack_v1.0_val'
    C.cfg.MODEL.TYPE = 'keypoint_rcnn'
    d = C.get_output_dir(training=False)
    assert d.endswith(os.path.join('test', 'posetrack_v1.0_val', 'keypoint_rcnn')) and os.path.isdir(d)


def test_output_dir_of_a_json_roidb_given_by_path(tmp_path):
    """TEST.DATASET may be the path of a JSON roidb (test_engine.JsonListDataset): its file stem names the directory level the
    reference fills with the dataset name (config.py:777-785), so the output stays under OUTPUT_DIR."""
    import os
    from detectandtrack_b200.core.config import cfg, reset_cfg, get_output_dir
    reset_cfg()
    try:
        cfg.OUTPUT_DIR = str(tmp_path / 'out')
        cfg.MODEL.TYPE = 'keypoint_rcnn'
        cfg.TEST.DATASET = str(tmp_path / 'lists' / 'posetrack_val.json')
        d = get_output_dir(training=False)
        assert d == os.path.join(str(tmp_path / 'out'), 'test', 'posetrack_val', 'keypoint_rcnn') and os.path.isdir(d)
        cfg.TEST.DATASET = 'synthetic_2x3'
        assert get_output_dir(training=False).endswith(os.path.join('test', 'synthetic_2x3', 'keypoint_rcnn'))
    finally:
        reset_cfg()


def test_device_jpeg_decode_only_for_jpeg_files():
    from detectandtrack_b200.core.test_engine import _all_jpeg
    assert _all_jpeg([dict(image=['a/1.jpg', 'a/2.JPEG']), dict(image='b/3.jpg')])
    assert not _all_jpeg([dict(image=['a/1.jpg', 'a/2.png'])])
    assert not _all_jpeg([dict(image='x.jpg', synthetic=True)])
