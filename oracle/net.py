"""Oracle: the detection graph as a plain PyTorch fp32 CPU functional (TEST
INFRASTRUCTURE ONLY).  NC[T]HW blobs and reference blob names, one function per
reference builder:

  conv_body      lib/modeling/ResNet3D.py:251-298 / ResNet.py:230-263 (+ :21-152 blocks)
  fpn            lib/modeling/FPN3D.py:109-222 / FPN.py:114-203
  time_pool      lib/modeling/model_builder.py:1024-1042, detector.py:559-576
  rpn_heads_fpn  lib/modeling/FPN.py:205-279
  box_head_2mlp  lib/modeling/head_builder.py:17-37 + model_builder.py:426-478
  keypoint_head  lib/modeling/keypoint_rcnn_heads.py:39-73 + model_builder.py:755-870
  roi_features   lib/modeling/detector.py:216-310 (RoIAlign per level + BatchPermutation)

Stand-ins for un-vendored Caffe2/cuDNN operators ("parity unpinned" by the
reference): torch.nn.functional.conv2d/conv3d (cross-correlation, like
Caffe2), max_pool (floor mode, like Caffe2 legacy pooling), conv_transpose2d (weight
(Cin, Cout, kh, kw), like Caffe2), linear, torchvision.ops.roi_align(aligned=False)
(same lineage as the Detectron RoIAlign op), nearest 2x upsampling as index // 2.
"""
import numpy as np

from . import proposals as oprop
from . import keypoints as okp


_DEVICE = {'device': 'cpu', 'cache': {}}


def set_device(device):
    """'cpu' (the oracle proper) or 'cuda' (bench.py's cuDNN stand-in leg: the same graph through cuDNN on the GPU)."""
    _DEVICE['device'] = device
    if device == 'cpu':
        _DEVICE['cache'].clear()


def _t(blobs, name):
    import torch
    if torch.is_tensor(blobs[name]):          # a dict of (possibly requires_grad) tensors: the autograd oracle of the training tests
        return blobs[name]
    if _DEVICE['device'] == 'cpu':
        return torch.from_numpy(np.ascontiguousarray(blobs[name]))
    key = (id(blobs), name)
    t = _DEVICE['cache'].get(key)
    if t is None:
        t = _DEVICE['cache'][key] = torch.from_numpy(np.ascontiguousarray(blobs[name])).to(_DEVICE['device'])
    return t


def _conv(x, blobs, name, stride, pad, bias=False):
    import torch.nn.functional as F
    w = _t(blobs, name + '_w')
    b = _t(blobs, name + '_b') if bias else None
    if w.dim() == 5:
        return F.conv3d(x, w, b, stride, pad)
    return F.conv2d(x, w, b, stride[1:], pad[1:])


def _affine(x, blobs, name):
    s, b = _t(blobs, name + '_s'), _t(blobs, name + '_b')
    shape = [1, -1] + [1] * (x.dim() - 2)
    return x * s.view(shape) + b.view(shape)            # affine_channel_nd_op.cu:19-32


def conv_body(blobs, spec, data):
    """data (B,3,T,H,W) [3-D] or (B,3,H,W) [2-D] -> dict of stage outputs res{s}_{n-1}_sum."""
    import torch
    import torch.nn.functional as F
    x = _conv(data, blobs, 'conv1', (1, 2, 2), (0, 3, 3))
    x = torch.relu(_affine(x, blobs, 'res_conv1_bn'))
    if x.dim() == 5:
        x = F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    else:
        x = F.max_pool2d(x, 3, 2, 1)
    outs = {}
    dim_in = spec.dims[0]
    for s, n in enumerate(spec.counts):
        dim_out = spec.dims[s + 1]
        tk = 1 if s == 0 else spec.tk_body
        for i in range(n):
            pre = 'res%d_%d' % (s + 2, i)
            stride = 2 if (dim_in != dim_out and s != 0) else 1          # ResNet3D.py:131-132
            st = (1, stride, stride)
            if spec.block == 'bottleneck':
                s1, s3 = (st, (1, 1, 1)) if spec.stride_1x1 else ((1, 1, 1), st)
                y = torch.relu(_affine(_conv(x, blobs, pre + '_branch2a', s1, (0, 0, 0)), blobs, pre + '_branch2a_bn'))
                y = torch.relu(_affine(_conv(y, blobs, pre + '_branch2b', s3, (tk // 2, 1, 1)), blobs, pre + '_branch2b_bn'))
                y = _affine(_conv(y, blobs, pre + '_branch2c', (1, 1, 1), (0, 0, 0)), blobs, pre + '_branch2c_bn')
            else:
                y = torch.relu(_affine(_conv(x, blobs, pre + '_branch2a', st, (tk // 2, 1, 1)), blobs, pre + '_branch2a_bn'))
                y = _affine(_conv(y, blobs, pre + '_branch2b', (1, 1, 1), (tk // 2, 1, 1)), blobs, pre + '_branch2b_bn')
            if dim_in != dim_out:
                sc = _affine(_conv(x, blobs, pre + '_branch1', st, (0, 0, 0)), blobs, pre + '_branch1_bn')
            else:
                sc = x
            x = torch.relu(y + sc)
            dim_in = dim_out
        outs['res%d_%d_sum' % (s + 2, n - 1)] = x
    return outs


def _up2(x):
    return x.repeat_interleave(2, dim=-2).repeat_interleave(2, dim=-1)


def fpn(blobs, spec, stage_outs, with_p6=True):
    """Returns [P_max ... P2] (coarsest first, like blobs_fpn) incl. P6 when asked."""
    import torch.nn.functional as F
    names = spec.stage_blobs[::-1]
    inner = {}
    inner[names[0]] = _conv(stage_outs[names[0]], blobs, 'fpn_inner_' + names[0], (1, 1, 1), (0, 0, 0), bias=True)
    for i in range(1, len(names)):
        lat = _conv(stage_outs[names[i]], blobs, 'fpn_inner_%s_lateral' % names[i], (1, 1, 1), (0, 0, 0), bias=True)
        inner[names[i]] = lat + _up2(inner[names[i - 1]])
    tk = spec.tk_body
    outs = [_conv(inner[n], blobs, 'fpn_' + n, (1, 1, 1), (tk // 2, 1, 1), bias=True) for n in names]
    if with_p6:
        p5 = outs[0]
        outs.insert(0, p5[..., ::2, ::2])              # MaxPool k=1 s=2 == subsample (FPN3D.py:157-160)
    return outs


def time_pool(x, link, num_frames_mid):
    if x.dim() == 4 or link in ('', 'none2d'):
        return x
    if link == 'slice-center':
        return x[:, :, int(num_frames_mid / 2)]
    if link == 'avg':
        return x.mean(dim=2)
    raise NotImplementedError(link)


def rpn_heads_fpn(blobs, spec, feats):
    """feats: [P6..P2] 2-D maps -> list over levels (finest first) of (logits (B,A,H,W), deltas (B,4A,H,W))."""
    import torch
    import torch.nn.functional as F
    k = str(spec.rpn_levels[0])
    out = []
    for f in feats[::-1]:
        h = torch.relu(F.conv2d(f, _t(blobs, 'conv_rpn_fpn%s_w' % k), _t(blobs, 'conv_rpn_fpn%s_b' % k), 1, 1))
        out.append((F.conv2d(h, _t(blobs, 'rpn_cls_logits_fpn%s_w' % k), _t(blobs, 'rpn_cls_logits_fpn%s_b' % k)),
                    F.conv2d(h, _t(blobs, 'rpn_bbox_pred_fpn%s_w' % k), _t(blobs, 'rpn_bbox_pred_fpn%s_b' % k))))
    return out


def roi_features(feats_fine_first, scales, rois, resolution, sampling_ratio, k_min=2, k_max=5):
    """detector.py:256-310 for 2-D heads: per-level RoIAlign, concat, un-shuffle.  rois (R,5)."""
    import torch
    from torchvision.ops import roi_align
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    if len(feats_fine_first) == 1:
        return roi_align(feats_fine_first[0].float(), torch.from_numpy(rois).to(feats_fine_first[0].device), (resolution, resolution), scales[0], sampling_ratio, aligned=False)
    _, per_level, restore = oprop.distribute(rois, k_min, k_max)
    parts = []
    for l, r in enumerate(per_level):
        f = feats_fine_first[l]
        if r.shape[0] == 0:
            parts.append(torch.zeros((0, f.shape[1], resolution, resolution), device=f.device))
        else:
            parts.append(roi_align(f.float(), torch.from_numpy(np.ascontiguousarray(r)).to(f.device), (resolution, resolution), scales[l], sampling_ratio, aligned=False))
    return torch.cat(parts, 0)[torch.from_numpy(restore.astype(np.int64)).to(parts[0].device)]


def box_head_2mlp(blobs, roi_feat):
    import torch
    import torch.nn.functional as F
    x = roi_feat.reshape(roi_feat.shape[0], -1)
    x = torch.relu(F.linear(x, _t(blobs, 'fc6_w'), _t(blobs, 'fc6_b')))
    x = torch.relu(F.linear(x, _t(blobs, 'fc7_w'), _t(blobs, 'fc7_b')))
    return (F.linear(x, _t(blobs, 'cls_score_w'), _t(blobs, 'cls_score_b')),
            F.linear(x, _t(blobs, 'bbox_pred_w'), _t(blobs, 'bbox_pred_b')))


def keypoint_head_2d(blobs, roi_feat, num_convs=8):
    """-> kps_score (D, K, 4S, 4S) and kps_score_lowres (D, K, 2S, 2S)."""
    import torch
    import torch.nn.functional as F
    x = roi_feat
    for i in range(num_convs):
        x = torch.relu(F.conv2d(x, _t(blobs, 'conv_fcn%d_w' % (i + 1)), _t(blobs, 'conv_fcn%d_b' % (i + 1)), 1, 1))
    low = okp.deconv_k4s2p1(x, _t(blobs, 'kps_score_lowres_w'), _t(blobs, 'kps_score_lowres_b'))
    return okp.bilinear_upsample2x(low), low


# ------------------------------------------------------------------ 3-D (tube) heads
def rpn_heads_3d(blobs, spec, feat):
    """model_builder.py:500-563.  feat (B,C,T,H,W) -> logits (B,A,H,W) [TimePool avg of the per-frame
    logits], deltas (B, A*T*4, H, W) with channel a*4T + t*4 + k."""
    import torch
    import torch.nn.functional as F
    w = _t(blobs, 'conv_rpn_w')
    tk = w.shape[2]
    h = torch.relu(F.conv3d(feat, w, _t(blobs, 'conv_rpn_b'), 1, (tk // 2, 1, 1)))
    lg = F.conv3d(h, _t(blobs, 'rpn_cls_logits_1_w'), _t(blobs, 'rpn_cls_logits_1_b')).mean(dim=2)
    d = F.conv3d(h, _t(blobs, 'rpn_bbox_pred_1_w'), _t(blobs, 'rpn_bbox_pred_1_b'))        # (B, 4A, T, H, W)
    B, A4, T, H, W = d.shape
    d = d.view(B, A4 // 4, 4, T, H, W).permute(0, 1, 3, 2, 4, 5).reshape(B, -1, H, W)
    return lg, d


def roi_features_tube(feat, scale, rois, resolution, sampling_ratio):
    """detector.py:216-254 for 3-D heads: RoIToBatchFormat + time->batch + RoIAlign + inverse.
    feat (B,C,T,H,W), rois (R, 4T+1) -> (R, C, T, res, res)."""
    import torch
    from torchvision.ops import roi_align
    B, C, T, H, W = feat.shape
    f = feat.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
    bt = oprop.roi_to_batch_format(np.asarray(rois, dtype=np.float32)).astype(np.float32)       # rows n*T + t
    out = roi_align(f, torch.from_numpy(bt), (resolution, resolution), scale, sampling_ratio, aligned=False)
    R = rois.shape[0]
    return out.view(R, T, C, resolution, resolution).permute(0, 2, 1, 3, 4).contiguous()


def _basic_block_3d(blobs, x, pre, has_sc):
    import torch
    y = torch.relu(_affine(_conv(x, blobs, pre + '_branch2a', (1, 1, 1), (0, 1, 1)), blobs, pre + '_branch2a_bn'))
    y = _affine(_conv(y, blobs, pre + '_branch2b', (1, 1, 1), (0, 1, 1)), blobs, pre + '_branch2b_bn')
    sc = _affine(_conv(x, blobs, pre + '_branch1', (1, 1, 1), (0, 0, 0)), blobs, pre + '_branch1_bn') if has_sc else x
    return torch.relu(y + sc)


def box_head_conv5_3d(blobs, roi_feat, n_blocks=2):
    """ResNet3D.add_ResNet18_roi_conv5_head (:301-327, stride_init 1, time kernel 1) +
    add_fast_rcnn_outputs 3-D (:427-473).  roi_feat (R,C,T,7,7) -> cls (R,Cls), bbox (R, Cls*T*4)."""
    import torch.nn.functional as F
    x = roi_feat
    for i in range(n_blocks):
        x = _basic_block_3d(blobs, x, 'res5_%d' % i, i == 0)
    x = x.mean(dim=4).mean(dim=3)                                   # ReduceBackMean W then H -> (R, C, T)
    x = x[..., None, None]
    cls = F.conv3d(x, _t(blobs, 'cls_score_1_w'), _t(blobs, 'cls_score_1_b')).mean(4).mean(3).mean(2)
    bb = F.conv3d(x, _t(blobs, 'bbox_pred_1_w'), _t(blobs, 'bbox_pred_1_b'))                   # (R, 4C, T, 1, 1)
    R, C4, T = bb.shape[:3]
    bb = bb.view(R, C4 // 4, 4, T, 1, 1).permute(0, 1, 3, 2, 4, 5).reshape(R, -1, 1, 1).mean(3).mean(2)
    return cls, bb


def keypoint_head_3d(blobs, roi_feat, num_convs=8):
    """keypoint_rcnn_heads.add_roi_pose_head_v1convX_3d + add_heatmap_outputs with
    NO_3D_DECONV_TIME_TO_CH (time -> batch for the deconv, back to channel t*K + k).
    roi_feat (D,C,T,14,14) -> kps_score (D, T*K, 56, 56)."""
    import torch
    import torch.nn.functional as F
    x = roi_feat
    for i in range(num_convs):
        w = _t(blobs, 'conv_fcn%d_w' % (i + 1))
        x = torch.relu(F.conv3d(x, w, _t(blobs, 'conv_fcn%d_b' % (i + 1)), 1, (w.shape[2] // 2, 1, 1)))
    D, C, T, H, W = x.shape
    xb = x.permute(0, 2, 1, 3, 4).reshape(D * T, C, H, W)
    low = okp.deconv_k4s2p1(xb, _t(blobs, 'kps_score_lowres_w'), _t(blobs, 'kps_score_lowres_b'))
    up = okp.bilinear_upsample2x(low)                                # (D*T, K, 56, 56)
    K = up.shape[1]
    return up.view(D, T * K, up.shape[2], up.shape[3])
