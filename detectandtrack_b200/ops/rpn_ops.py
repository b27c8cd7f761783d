"""Device-tensor API over csrc/proposals.cu (RPN proposals, collect/distribute, box decode,
detection limit).  Tensors stay on the GPU; counts are device int32 (CUDA-graph friendly)."""
import ctypes as C

import numpy as np

from .. import _lib as L
from . import box_ops

BBOX_XFORM_CLIP = float(np.log(1000. / 16.))     # lib/core/config.py:672


def _nhwc_ld(t):
    """Leading dim of a [B,H,W,C'] tensor that may be a channel slice of a wider contiguous one."""
    B, H, W, _ = t.shape
    ld = t.stride(2)
    assert t.stride(3) == 1 and t.stride(1) == W * ld and (B == 1 or t.stride(0) == H * W * ld), \
        'expected an NHWC tensor (or a channel slice of one)'
    return ld


def rpn_proposals(logits, deltas, anchors, feat_stride, im_info, pre_nms_topn, min_size=0.0,
                  T=1, out=None, counts=None, clip=BBOX_XFORM_CLIP, time_major=False):
    """logits [B,H,W,A], deltas [B,H,W,4AT] (fp32 or bf16; channel slices of a wider NHWC tensor
    are fine), anchors [A,4T] fp64 cuda, im_info [B,3] fp32 cuda.
    Returns (props [B,K,4T+1] fp32, counts [B] int32)."""
    torch = L.require_cuda()
    if time_major:          # [B, T, H, W, C'] channel slices of contiguous 5-D tensors
        B, Tt, H, W, _ = logits.shape
        assert Tt == T and deltas.shape[:4] == logits.shape[:4]
        ld_s, ld_d = logits.stride(3), deltas.stride(3)
        assert logits.stride(4) == 1 and deltas.stride(4) == 1 and logits.stride(1) == H * W * ld_s
    else:
        B, H, W, _ = logits.shape
        ld_s, ld_d = _nhwc_ld(logits), _nhwc_ld(deltas)
    A = anchors.shape[0]
    assert logits.dtype == deltas.dtype
    act_f32 = int(logits.dtype == torch.float32)
    n = H * W * A
    K = n if (pre_nms_topn <= 0 or pre_nms_topn > n) else pre_nms_topn
    if out is None:
        out = torch.zeros((B, K, 4 * T + 1), dtype=torch.float32, device='cuda')
    if counts is None:
        counts = torch.zeros((B,), dtype=torch.int32, device='cuda')
    assert out.shape[-1] == 4 * T + 1 and out.shape[-2] >= K
    L.call('dt_rpn_proposals', C.c_void_p(logits.data_ptr()), ld_s, C.c_void_p(deltas.data_ptr()), ld_d, act_f32,
           B, H, W, A, T, L.ptr(anchors), float(feat_stride), L.ptr(im_info), int(pre_nms_topn), float(min_size),
           float(clip), C.c_void_p(out.data_ptr()), out.stride(0), C.c_void_p(counts.data_ptr()), counts.stride(0),
           int(bool(time_major)), L.stream_ptr())
    return out, counts


def collect(props, keep, nkeep, post_nms_topn, R=None):
    """props [B,L,K,4T+1], keep [B*L,K], nkeep [B*L] -> rois [B,R,4T+1], scores [B,R], counts [B]."""
    torch = L.require_cuda()
    B, Lv, K, ld = props.shape
    T = (ld - 1) // 4
    R = R or (post_nms_topn if post_nms_topn > 0 else Lv * K)
    rois = torch.zeros((B, R, ld), dtype=torch.float32, device='cuda')
    scores = torch.zeros((B, R), dtype=torch.float32, device='cuda')
    counts = torch.zeros((B,), dtype=torch.int32, device='cuda')
    L.call('dt_collect_rpn', L.ptr(props.contiguous()), L.ptr(keep.contiguous()), L.ptr(nkeep.contiguous()), B, Lv, K, T,
           int(post_nms_topn), L.ptr(rois), L.ptr(scores), L.ptr(counts), R, L.stream_ptr())
    return rois, scores, counts


def distribute(rois, n_dev=None, col0=1, T=1, k_min=2, k_max=5, s0=224.0, lvl0=4.0, want_restore=True):
    """rois [n, ld] -> (levels [n] int32, idx_restore [n] int32 or None, level_counts)."""
    torch = L.require_cuda()
    rois = rois.contiguous()
    n, ld = rois.shape
    levels = torch.zeros((max(n, 1),), dtype=torch.int32, device='cuda')
    restore = torch.zeros((max(n, 1),), dtype=torch.int32, device='cuda') if want_restore else None
    lc = torch.zeros((k_max - k_min + 1,), dtype=torch.int32, device='cuda')
    L.call('dt_distribute_fpn', L.ptr(rois), n, L.ptr(n_dev), ld, col0, T, k_min, k_max, float(s0), float(lvl0),
           L.ptr(levels), L.ptr(restore), L.ptr(lc), L.stream_ptr())
    return levels[:n], (restore[:n] if want_restore else None), lc


def box_decode(rois, roi_counts, cls_logits, bbox_deltas, num_classes, im_info, im_hw,
               weights=(10., 10., 5., 5.), score_thresh=0.05, T=1, clip=BBOX_XFORM_CLIP):
    """rois [B,R,4T+1]; cls_logits [B*R, >=C]; bbox_deltas [B*R, >=4TC] fp32.
    Returns dets [B, C-1, R, 4T+1], det_counts [B*(C-1)]."""
    torch = L.require_cuda()
    B, R, _ = rois.shape
    dets = torch.zeros((B, num_classes - 1, R, 4 * T + 1), dtype=torch.float32, device='cuda')
    cnt = torch.zeros((B * (num_classes - 1),), dtype=torch.int32, device='cuda')
    w4 = (C.c_float * 4)(*[float(w) for w in weights])
    assert cls_logits.dtype == torch.float32 and bbox_deltas.dtype == torch.float32
    assert cls_logits.stride(1) == 1 and bbox_deltas.stride(1) == 1
    L.call('dt_box_decode', L.ptr(rois.contiguous()), L.ptr(roi_counts), B, R, T, C.c_void_p(cls_logits.data_ptr()),
           cls_logits.stride(0), C.c_void_p(bbox_deltas.data_ptr()), bbox_deltas.stride(0), num_classes, L.ptr(im_info), L.ptr(im_hw), w4, float(clip),
           float(score_thresh), L.ptr(dets), L.ptr(cnt), L.stream_ptr())
    return dets, cnt


def limit_detections(dets, keep, nkeep, max_per_im, cap=None):
    """dets [B,C-1,R,ld], keep [B*(C-1), R], nkeep -> out [B, C-1, cap, ld], counts [B*(C-1)]."""
    torch = L.require_cuda()
    B, C1, R, ld = dets.shape
    T = (ld - 1) // 4
    cap = cap or R
    out = torch.zeros((B, C1, cap, ld), dtype=torch.float32, device='cuda')
    cnt = torch.zeros((B * C1,), dtype=torch.int32, device='cuda')
    L.call('dt_limit_detections', L.ptr(dets), L.ptr(keep.contiguous()), L.ptr(nkeep), B, C1 + 1, R, T, int(max_per_im),
           L.ptr(out), L.ptr(cnt), cap, L.stream_ptr())
    return out, cnt


def generate_proposals_level(logits, deltas, anchors, feat_stride, im_info, pre_nms_topn, post_nms_topn,
                             nms_thresh, min_size=0.0, T=1):
    """Full GenerateProposalsOp for one level (generate_proposals.py:40-181): returns
    (props [B,K,4T+1], keep [B,K], nkeep [B]) — rois = props[b, keep[b, :nkeep[b]]]."""
    props, counts = rpn_proposals(logits, deltas, anchors, feat_stride, im_info, pre_nms_topn, min_size, T)
    cmp_mode = box_ops.NMS_2D_GE if T == 1 else box_ops.NMS_TUBE_GT
    order = box_ops.ORDER_INDEX if T == 1 else box_ops.ORDER_SCORE
    keep, nkeep = box_ops.nms_batched(props, counts, nms_thresh, cmp_mode, order, max_keep=post_nms_topn)
    return props, keep, nkeep
