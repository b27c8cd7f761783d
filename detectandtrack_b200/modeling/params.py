"""Parameter inventory of the detection graphs, in the REFERENCE's blob names and Caffe2
filter layouts (so weight pkls written by / for the reference load unchanged):

    conv filters   <name>_w  (Cout, Cin, kT, kH, kW) for ConvNd, (Cout, Cin, kH, kW) for Conv
    conv biases    <name>_b  (Cout,)
    AffineChannel  <name>_bn_s, <name>_bn_b  (C,)            lib/modeling/detector.py:89-109
    FC             <name>_w  (Cout, Cin), <name>_b
    ConvTranspose  <name>_w  (Cin, Cout, kH, kW)

``param_shapes(cfg)`` walks the same builders the reference walks
(lib/modeling/ResNet3D.py:251-327, FPN3D.py:109-222, FPN.py:205-279,
head_builder.py:17-37, keypoint_rcnn_heads.py:39-73, model_builder.py:426-478,755-870).
``random_blobs`` gives the seeded synthetic weights of SURVEY.md §8(d);
``load_weights_file`` mirrors lib/utils/net.py:95-249 incl. the 2-D -> 3-D inflation.
"""
import logging
import pickle
from collections import OrderedDict

import numpy as np

logger = logging.getLogger(__name__)

_BLOCKS = {  # ResNet3D.py:334-394 / ResNet.py
    'ResNet18': ((2, 2, 2, 2), 'basic', (64, 64, 128, 256, 512)),
    'ResNet34': ((3, 4, 6, 3), 'basic', (64, 64, 128, 256, 512)),
    'ResNet50': ((3, 4, 6, 3), 'bottleneck', (64, 256, 512, 1024, 2048)),
    'ResNet101': ((3, 4, 23, 3), 'bottleneck', (64, 256, 512, 1024, 2048)),
    'ResNet152': ((3, 8, 36, 3), 'bottleneck', (64, 256, 512, 1024, 2048)),
}


def parse_conv_body(name):
    """'FPN3D.add_fpn_ResNet50_conv5_body' -> dict(module, fpn, arch, nstages)."""
    module, func = name.split('.')
    fpn = func.startswith('add_fpn_')
    core = func[len('add_fpn_'):] if fpn else func[len('add_'):]
    arch, conv, _ = core.split('_')[:3]
    assert arch in _BLOCKS, 'unknown backbone {}'.format(name)
    nstages = 4 if conv == 'conv5' else 3
    return dict(module=module, fpn=fpn, arch=arch, nstages=nstages, is3d=module.endswith('3D'))


class GraphSpec(object):
    """Static description of the graph derived from cfg (shared by engine and oracle)."""

    def __init__(self, cfg):
        b = parse_conv_body(cfg.MODEL.CONV_BODY)
        self.body = b
        self.counts, self.block, self.dims = _BLOCKS[b['arch']]
        self.counts = self.counts[:b['nstages']]
        self.video = bool(cfg.MODEL.VIDEO_ON)
        self.T = cfg.VIDEO.NUM_FRAMES if self.video else 1
        self.is3d = b['is3d']
        self.tk_body = cfg.VIDEO.TIME_KERNEL_DIM.BODY if self.is3d else 1
        self.link = cfg.VIDEO.BODY_HEAD_LINK if self.video else 'none2d'
        self.head3d = self.video and cfg.VIDEO.BODY_HEAD_LINK == ''
        self.T_head = (cfg.VIDEO.NUM_FRAMES_MID if cfg.VIDEO.NUM_FRAMES_MID > 0 else self.T) if self.head3d else 1
        self.fpn = b['fpn']
        self.fpn_dim = cfg.FPN.DIM
        self.num_classes = cfg.MODEL.NUM_CLASSES
        self.stride_1x1 = bool(cfg.RESNETS.STRIDE_1X1)
        self.roi_head = cfg.MODEL.ROI_HEAD
        self.kps_head = cfg.KRCNN.ROI_KEYPOINTS_HEAD
        self.K = cfg.KRCNN.NUM_KEYPOINTS
        if self.fpn:
            assert self.counts.__len__() == 4, 'FPN needs a conv5 body'
            self.rpn_levels = list(range(cfg.FPN.RPN_MIN_LEVEL, cfg.FPN.RPN_MAX_LEVEL + 1))
            self.roi_levels = list(range(cfg.FPN.ROI_MIN_LEVEL, cfg.FPN.ROI_MAX_LEVEL + 1))
            self.num_anchors = len(cfg.FPN.RPN_ASPECT_RATIOS)
            if self.head3d:
                raise NotImplementedError('3-D FPN RPN heads are unimplemented in the reference too '
                                          '(lib/modeling/FPN3D.py:228)')
        else:
            self.num_anchors = len(cfg.RPN.SIZES) * len(cfg.RPN.ASPECT_RATIOS)
        self.dim_inner = cfg.RESNETS.NUM_GROUPS * cfg.RESNETS.WIDTH_PER_GROUP
        # blob names of the stage outputs (stage_info_*: res{s}_{n-1}_sum)
        self.stage_blobs = ['res%d_%d_sum' % (s + 2, n - 1) for s, n in enumerate(self.counts)]

    def conv_shape(self, cout, cin, k):
        """k = (kT, kH, kW); 3-D bodies use 5-D filters, 2-D ones 4-D."""
        return (cout, cin) + (tuple(k) if self.is3d else tuple(k[1:]))


def _add_conv_affine(P, spec, prefix, cin, cout, k):
    P[prefix + '_w'] = spec.conv_shape(cout, cin, k)
    P[prefix + '_bn_s'] = (cout,)
    P[prefix + '_bn_b'] = (cout,)


def param_shapes(cfg):
    spec = GraphSpec(cfg)
    P = OrderedDict()
    d = spec.dims
    P['conv1_w'] = spec.conv_shape(d[0], 3, (1, 7, 7))
    P['res_conv1_bn_s'] = (d[0],); P['res_conv1_bn_b'] = (d[0],)
    dim_in = d[0]
    for s, n in enumerate(spec.counts):
        dim_out = d[s + 1]
        inner = spec.dim_inner * (2 ** s)
        tk = 1 if s == 0 else spec.tk_body            # res2 has no temporal kernel (ResNet3D.py:269-272)
        for i in range(n):
            pre = 'res%d_%d' % (s + 2, i)
            if spec.block == 'bottleneck':
                _add_conv_affine(P, spec, pre + '_branch2a', dim_in, inner, (1, 1, 1))
                _add_conv_affine(P, spec, pre + '_branch2b', inner, inner, (tk, 3, 3))
                _add_conv_affine(P, spec, pre + '_branch2c', inner, dim_out, (1, 1, 1))
            else:
                _add_conv_affine(P, spec, pre + '_branch2a', dim_in, dim_out, (tk, 3, 3))
                _add_conv_affine(P, spec, pre + '_branch2b', dim_out, dim_out, (tk, 3, 3))
            if dim_in != dim_out:
                P[pre + '_branch1_w'] = spec.conv_shape(dim_out, dim_in, (1, 1, 1))
                P[pre + '_branch1_bn_s'] = (dim_out,); P[pre + '_branch1_bn_b'] = (dim_out,)
            dim_in = dim_out
    if spec.fpn:
        fd = spec.fpn_dim
        blobs = spec.stage_blobs[::-1]                 # coarsest first, like stage_info.blobs
        dims = list(d[1:len(spec.counts) + 1])[::-1]
        P['fpn_inner_%s_w' % blobs[0]] = spec.conv_shape(fd, dims[0], (1, 1, 1)); P['fpn_inner_%s_b' % blobs[0]] = (fd,)
        for i in range(1, len(blobs)):
            P['fpn_inner_%s_lateral_w' % blobs[i]] = spec.conv_shape(fd, dims[i], (1, 1, 1))
            P['fpn_inner_%s_lateral_b' % blobs[i]] = (fd,)
        for bname in blobs:
            P['fpn_%s_w' % bname] = spec.conv_shape(fd, fd, (spec.tk_body, 3, 3)); P['fpn_%s_b' % bname] = (fd,)
        # RPN heads are the 2-D FPN ones, shared across levels (FPN.py:205-279)
        k = str(spec.rpn_levels[0])
        A = spec.num_anchors
        P['conv_rpn_fpn%s_w' % k] = (fd, fd, 3, 3); P['conv_rpn_fpn%s_b' % k] = (fd,)
        P['rpn_cls_logits_fpn%s_w' % k] = (A, fd, 1, 1); P['rpn_cls_logits_fpn%s_b' % k] = (A,)
        P['rpn_bbox_pred_fpn%s_w' % k] = (4 * A, fd, 1, 1); P['rpn_bbox_pred_fpn%s_b' % k] = (4 * A,)
        dim_conv = fd
    else:
        dim_conv = dim_in
        A, Th = spec.num_anchors, spec.T_head
        tk = cfg.VIDEO.TIME_KERNEL_DIM.HEAD_RPN if spec.head3d else 1
        if spec.head3d:                                   # blob names of model_builder.py:509-547
            P['conv_rpn_w'] = (dim_conv, dim_conv, tk, 3, 3); P['conv_rpn_b'] = (dim_conv,)
            P['rpn_cls_logits_1_w'] = (A, dim_conv, 1, 1, 1); P['rpn_cls_logits_1_b'] = (A,)
            P['rpn_bbox_pred_1_w'] = (4 * A, dim_conv, 1, 1, 1); P['rpn_bbox_pred_1_b'] = (4 * A,)   # time fold :553-563
        else:
            P['conv_rpn_w'] = (dim_conv, dim_conv, 3, 3); P['conv_rpn_b'] = (dim_conv,)
            P['rpn_cls_logits_w'] = (A, dim_conv, 1, 1); P['rpn_cls_logits_b'] = (A,)
            P['rpn_bbox_pred_w'] = (4 * A, dim_conv, 1, 1); P['rpn_bbox_pred_b'] = (4 * A,)
    # ---- box head ----
    C = spec.num_classes
    if spec.roi_head.endswith('add_roi_2mlp_head'):
        res = cfg.FAST_RCNN.ROI_XFORM_RESOLUTION
        hid = cfg.FAST_RCNN.MLP_HEAD_DIM
        P['fc6_w'] = (hid, spec.T_head * dim_conv * res * res); P['fc6_b'] = (hid,)
        P['fc7_w'] = (hid, hid); P['fc7_b'] = (hid,)
        dim_box = hid
        if spec.head3d:
            raise NotImplementedError('2mlp head on 3-D RoI features is not exercised by any shipped config')
        P['cls_score_w'] = (C, dim_box); P['cls_score_b'] = (C,)
        P['bbox_pred_w'] = (4 * C, dim_box); P['bbox_pred_b'] = (4 * C,)
    elif 'roi_conv5_head' in spec.roi_head:
        arch = spec.roi_head.split('add_')[1].split('_')[0]
        cnts, block, dd = _BLOCKS[arch]
        n5, dout = cnts[3], dd[4]
        inner = spec.dim_inner * 8
        din = dim_conv
        saved = (spec.is3d,)
        for i in range(n5):
            pre = 'res5_%d' % i
            if block == 'bottleneck':
                _add_conv_affine(P, spec, pre + '_branch2a', din, inner, (1, 1, 1))
                _add_conv_affine(P, spec, pre + '_branch2b', inner, inner, (1, 3, 3))
                _add_conv_affine(P, spec, pre + '_branch2c', inner, dout, (1, 1, 1))
            else:
                _add_conv_affine(P, spec, pre + '_branch2a', din, dout, (1, 3, 3))
                _add_conv_affine(P, spec, pre + '_branch2b', dout, dout, (1, 3, 3))
            if din != dout:
                P[pre + '_branch1_w'] = spec.conv_shape(dout, din, (1, 1, 1))
                P[pre + '_branch1_bn_s'] = (dout,); P[pre + '_branch1_bn_b'] = (dout,)
            din = dout
        dim_box = dout
        if spec.head3d:
            P['cls_score_1_w'] = (C, dim_box, 1, 1, 1); P['cls_score_1_b'] = (C,)
            P['bbox_pred_1_w'] = (4 * C, dim_box, 1, 1, 1); P['bbox_pred_1_b'] = (4 * C,)
        else:
            P['cls_score_w'] = (C, dim_box); P['cls_score_b'] = (C,)
            P['bbox_pred_w'] = (4 * C, dim_box); P['bbox_pred_b'] = (4 * C,)
    else:
        raise NotImplementedError('ROI_HEAD {}'.format(spec.roi_head))
    # ---- keypoint head ----
    if cfg.MODEL.KEYPOINTS_ON:
        hd = cfg.KRCNN.CONV_HEAD_DIM
        ks = cfg.KRCNN.CONV_HEAD_KERNEL
        nd = spec.kps_head.endswith('_3d')
        tk = cfg.VIDEO.TIME_KERNEL_DIM.HEAD_KPS if nd else 1
        din = dim_conv
        for i in range(cfg.KRCNN.NUM_STACKED_CONVS):
            P['conv_fcn%d_w' % (i + 1)] = (hd, din, tk, ks, ks) if nd else (hd, din, ks, ks)
            P['conv_fcn%d_b' % (i + 1)] = (hd,)
            din = hd
        assert not cfg.KRCNN.USE_DECONV and cfg.KRCNN.USE_DECONV_OUTPUT and cfg.KRCNN.UP_SCALE == 2, \
            'only the shipped keypoint output stack (deconv output + 2x bilinear) is implemented'
        if nd and not cfg.KRCNN.NO_3D_DECONV_TIME_TO_CH:
            raise NotImplementedError('time-in-channel keypoint deconv (KRCNN.NO_3D_DECONV_TIME_TO_CH False) is not '
                                      'used by any shipped config')
        kt = 1
        P['kps_score_lowres_w'] = (din * kt, spec.K * kt, cfg.KRCNN.DECONV_KERNEL, cfg.KRCNN.DECONV_KERNEL)
        P['kps_score_lowres_b'] = (spec.K * kt,)
    return P, spec


def random_blobs(cfg, seed=None):
    """Seeded synthetic weights (He-normal filters, affine s~U(0.5,1.5), b~N(0,0.1))
    conditioned like a trained network so the synthetic workload is the one BASELINE.json names
    (R = 1000 proposals, D = DETECTIONS_PER_IM detections per clip):
      * conv1 absorbs the un-normalised pixel scale (He std / 64);
      * the last affine of every residual branch is damped (s~U(0.2,0.4)) so activations stay O(1);
      * FPN convs Xavier, heads with the reference's own init stds (GaussianFill 0.01 for RPN /
        cls_score, 0.001 for bbox_pred: FPN.py:223-262, model_builder.py:431-478; MSRAFill for the
        keypoint convs, keypoint_rcnn_heads.py:53-65), biases 0."""
    shapes, spec = param_shapes(cfg)
    rng = np.random.RandomState(cfg.RNG_SEED if seed is None else seed)
    blobs = OrderedDict()
    for name, shp in shapes.items():
        fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else 1
        if name.endswith('_bn_s'):
            last = name.endswith('_branch2c_bn_s') or (spec.block == 'basic' and name.endswith('_branch2b_bn_s'))
            v = rng.uniform(0.2, 0.4, shp) if last else rng.uniform(0.5, 1.5, shp)
        elif name.endswith('_bn_b'):
            v = rng.normal(0, 0.1, shp)
        elif name.endswith('_b'):
            v = np.zeros(shp)
        elif name == 'conv1_w':
            v = rng.normal(0, np.sqrt(2.0 / fan_in) / 64.0, shp)
        elif name.startswith('fpn_') or name.startswith(('fc6', 'fc7')):
            v = rng.uniform(-np.sqrt(3.0 / fan_in), np.sqrt(3.0 / fan_in), shp)          # XavierFill
        elif name.startswith(('conv_rpn', 'rpn_cls', 'rpn_bbox', 'cls_score')):
            v = rng.normal(0, 0.01, shp)
        elif name.startswith('bbox_pred'):
            v = rng.normal(0, 0.001, shp)
        elif name.startswith('kps_score_lowres'):
            v = rng.normal(0, np.sqrt(2.0 / (shp[0] * shp[2] * shp[3] / 4.0)), shp)      # ConvTranspose: Cin first
        else:
            v = rng.normal(0, np.sqrt(2.0 / fan_in), shp)                                # MSRAFill
        blobs[name] = v.astype(np.float32)
    return blobs, spec


def inflate_weights(pretrained_w, shape, mode, name=''):
    """lib/utils/net.py:95-161: 4-D filter -> 5-D by repeating on the time axis (-3)."""
    if len(shape) != 5 or pretrained_w.ndim != 4:
        return pretrained_w
    ncopies = float(shape[-3])
    w = np.repeat(np.expand_dims(pretrained_w, axis=-3), int(ncopies), axis=-3)
    if mode == 'mean-repeat':
        w = w / ncopies
    elif mode == 'repeat':
        pass
    elif mode == 'center-only':
        w[..., :int(ncopies / 2), :, :] = 0
        w[..., int(ncopies / 2) + 1:, :, :] = 0
    elif mode == 'center-only-rest-rand':
        w[..., :int(ncopies / 2), :, :] = 0.001 * np.random.randn(*w[..., :int(ncopies / 2), :, :].shape)
        w[..., int(ncopies / 2) + 1:, :, :] = 0.001 * np.random.randn(*w[..., int(ncopies / 2) + 1:, :, :].shape)
        w = w / ncopies
    else:
        raise ValueError('Invalid INFLATE_MODE: {}'.format(mode))
    if tuple(w.shape) != tuple(shape):
        logger.error('blob %s with shape %s does not match weights file shape %s (even after inflating to %s)',
                     name, shape, pretrained_w.shape, w.shape)
    return w


def load_weights_file(cfg, weights_file, init_missing=True):
    """lib/utils/net.py:164-249: {'blobs': {name: ndarray}, 'cfg': yaml} (or a bare dict).
    Names with a '_[xyz]_' prefix fall back to the un-prefixed source blob; shape mismatches
    go through inflate_weights; blobs missing from the file keep their random init."""
    with open(weights_file, 'rb') as f:
        try:
            src = pickle.load(f)
        except UnicodeDecodeError:
            f.seek(0)
            src = pickle.load(f, encoding='latin1')
    if 'blobs' in src:
        src = src['blobs']
    src = {(k.decode() if isinstance(k, bytes) else k): v for k, v in src.items()}
    blobs, spec = random_blobs(cfg)
    if not init_missing:
        blobs = OrderedDict((k, None) for k in blobs)
    shapes, _ = param_shapes(cfg)
    for name, shp in shapes.items():
        src_name = name[name.find(']_') + 2:] if (name.find(']_') >= 0 and name not in src) else name
        if src_name not in src:
            logger.info('%s not found', src_name)
            continue
        w = np.asarray(src[src_name])
        if tuple(w.shape) != tuple(shp):
            w = inflate_weights(w, shp, cfg.VIDEO.WEIGHTS_INFLATE_MODE, src_name)
        if tuple(w.shape) != tuple(shp):
            raise RuntimeError('weights blob {} has shape {} but the graph needs {}'.format(src_name, w.shape, shp))
        blobs[name] = w.astype(np.float32, copy=False)
    return blobs, spec


def save_weights_file(blobs, cfg_yaml, path):
    """lib/utils/net.py:252-294 wire format."""
    with open(path, 'wb') as f:
        pickle.dump(dict(blobs={k: np.asarray(v) for k, v in blobs.items()}, cfg=cfg_yaml), f, pickle.HIGHEST_PROTOCOL)
