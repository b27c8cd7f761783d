"""Device-tensor API over the box / NMS / assignment kernels (torch tensors are
only the memory container; arithmetic is in csrc/boxes.cu and csrc/lsa.cu)."""
import ctypes as C

from .. import _lib as L

NMS_2D_GE, NMS_TUBE_GT = 0, 1
ORDER_SCORE, ORDER_INDEX = 0, 1
MAX_TRACK_IDS = 999   # lib/core/tracking_engine.py:45
FIRST_TRACK_ID = 0    # lib/core/tracking_engine.py:46


def _f32c(t, torch):
    return t.to(device='cuda', dtype=torch.float32).contiguous()


def bbox_overlaps(boxes, query, T=None):
    """lib/utils/boxes.py:60-69.  boxes [N, >=4T], query [K, >=4T] cuda fp32 -> [N, K]."""
    torch = L.require_cuda()
    boxes, query = _f32c(boxes, torch), _f32c(query, torch)
    if T is None:
        T = boxes.shape[1] // 4
    n, k = boxes.shape[0], query.shape[0]
    out = torch.empty((n, k), dtype=torch.float32, device='cuda')
    L.call('dt_bbox_overlaps', L.ptr(boxes), n, boxes.shape[1], L.ptr(query), k, query.shape[1],
           T, L.ptr(out), max(k, 1), L.stream_ptr())
    return out


def _workspace(nbytes, torch, slot='nms'):
    """Scratch buffer of a call.  Allocated per call (the caching allocator makes that free): a process-wide cached
    buffer that is re-allocated when a later call needs more would be freed under a previously CAPTURED CUDA graph
    that still points at it (two ClipPipelines of different batch sizes) — under capture the allocation comes from
    that graph's own pool and lives as long as the graph."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device='cuda')


def nms_batched(dets, counts=None, thresh=0.5, cmp_mode=None, out_order=None, max_keep=0):
    """dets [B, Nmax, 4T+1] cuda fp32, counts [B] int32 (or None).
    Returns (keep [B, Nmax] int32, num_keep [B] int32).  cmp_mode/out_order default to the
    reference's dispatch (nms_wrapper.py:49-57): T == 1 -> '>=' + index order, else tube."""
    torch = L.require_cuda()
    dets = _f32c(dets, torch)
    B, nmax, ld = dets.shape
    T = (ld - 1) // 4
    if cmp_mode is None:
        cmp_mode = NMS_2D_GE if T == 1 else NMS_TUBE_GT
    if out_order is None:
        out_order = ORDER_INDEX if cmp_mode == NMS_2D_GE else ORDER_SCORE
    keep = torch.empty((B, max(nmax, 1)), dtype=torch.int32, device='cuda')
    num = L.zeros((B,), torch.int32)
    if counts is not None:
        counts = counts.to(device='cuda', dtype=torch.int32).contiguous()
    need = C.c_size_t(0)
    L.call('dt_nms_workspace_bytes', B, nmax, C.byref(need))
    ws = _workspace(need.value, torch)
    L.call('dt_nms_batched', L.ptr(dets), B, nmax, ld, T, L.ptr(counts), float(thresh), cmp_mode,
           out_order, int(max_keep), L.ptr(keep), L.ptr(num), L.ptr(ws), ws.numel(), L.stream_ptr())
    return keep, num


ALGOS = {'hungarian': 0, 'greedy': 1}


def lsa_batched(cost, nrows, ncols, algo='hungarian'):
    """cost [B, D, D'] cuda fp32 (rows prev, cols cur); returns matches [B, dmax] int32."""
    torch = L.require_cuda()
    cost = _f32c(cost, torch)
    B, P, Q = cost.shape
    dmax = max(P, Q)
    if P != dmax or Q != dmax:
        pad = torch.zeros((B, dmax, dmax), dtype=torch.float32, device='cuda')
        pad[:, :P, :Q] = cost
        cost = pad
    nrows = nrows.to(device='cuda', dtype=torch.int32).contiguous()
    ncols = ncols.to(device='cuda', dtype=torch.int32).contiguous()
    matches = torch.empty((B, dmax), dtype=torch.int32, device='cuda')
    status = torch.empty((B,), dtype=torch.int32, device='cuda')
    L.call('dt_lsa_batched', L.ptr(cost), B, dmax, dmax, L.ptr(nrows), L.ptr(ncols), ALGOS[algo], L.ptr(matches),
           L.ptr(status), L.stream_ptr())
    return matches, status


def match_frames(frames, counts, is_start=None, T=1, weight=1.0, algo='hungarian'):
    """frames [F, Dmax, ld] cuda fp32, counts [F] int32 -> matches [F, Dmax] int32."""
    torch = L.require_cuda()
    frames = _f32c(frames, torch)
    F, dmax, ld = frames.shape
    counts = counts.to(device='cuda', dtype=torch.int32).contiguous()
    if is_start is not None:
        is_start = is_start.to(device='cuda', dtype=torch.uint8).contiguous()
    matches = torch.empty((F, dmax), dtype=torch.int32, device='cuda')
    status = torch.empty((F,), dtype=torch.int32, device='cuda')
    L.call('dt_match_frames', L.ptr(frames), F, dmax, ld, T, L.ptr(counts), L.ptr(is_start), float(weight), ALGOS[algo],
           L.ptr(matches), L.ptr(status), L.stream_ptr())
    return matches, status


def pose_pck_cost(a, b, head_top, head_bottom, dist_thresh=0.5):
    """a [P, 4|3, K], b [Q, 4|3, K] fp32 poses -> [P, Q] fp64 'pose-pck' cost (tracking_engine.py:113-129)."""
    torch = L.require_cuda()
    a, b = _f32c(a, torch), _f32c(b, torch)
    P, _, K = a.shape
    Q = b.shape[0]
    out = torch.empty((P, Q), dtype=torch.float64, device='cuda')
    L.call('dt_pose_pck_cost', L.ptr(a), P, L.ptr(b), Q, a.shape[1] * K, K, int(head_top), int(head_bottom), float(dist_thresh),
           L.ptr(out), L.stream_ptr())
    return out


def frame_costs(boxes, counts, is_start, poses=None, T=1, w_iou=1.0, w_pck=0.0, head_top=2, head_bottom=1, dist_thresh=0.5):
    """boxes [F, Dmax, ld], poses [F, Dmax, 4|3, K] (or None) -> cost [F, Dmax, Dmax] fp32 (rows = previous frame)."""
    torch = L.require_cuda()
    boxes = _f32c(boxes, torch)
    F, dmax, ld = boxes.shape
    counts = counts.to(device='cuda', dtype=torch.int32).contiguous()
    if is_start is not None:
        is_start = is_start.to(device='cuda', dtype=torch.uint8).contiguous()
    K = ldp = 0
    if poses is not None:
        poses = _f32c(poses, torch)
        K, ldp = poses.shape[3], poses.shape[2] * poses.shape[3]
    cost = torch.empty((F, dmax, dmax), dtype=torch.float32, device='cuda')
    L.call('dt_frame_costs', L.ptr(boxes), ld, T, L.ptr(poses), ldp, K, int(head_top), int(head_bottom), float(dist_thresh),
           L.ptr(counts), L.ptr(is_start), F, dmax, float(w_iou), float(w_pck), L.ptr(cost), L.stream_ptr())
    return cost


def assign_track_ids(matches, counts, video_first, is_start=None,
                     first_id=FIRST_TRACK_ID, max_ids=MAX_TRACK_IDS):
    torch = L.require_cuda()
    F, dmax = matches.shape
    counts = counts.to(device='cuda', dtype=torch.int32).contiguous()
    video_first = video_first.to(device='cuda', dtype=torch.int32).contiguous()
    tracks = torch.empty((F, dmax), dtype=torch.int32, device='cuda')
    L.call('dt_assign_track_ids', L.ptr(matches.contiguous()), L.ptr(counts), L.ptr(is_start), F, dmax,
           L.ptr(video_first), video_first.numel(), first_id, max_ids, L.ptr(tracks), L.stream_ptr())
    return tracks


def prune_detections(boxes, counts, hw, conf, T=1, center_only=True, min_area=50.0):
    """boxes [F, Dmax, 4T+1]; hw [F, 2] (height, width).  Returns (out, counts_out, sel)."""
    torch = L.require_cuda()
    boxes = _f32c(boxes, torch)
    F, dmax, ld = boxes.shape
    counts = counts.to(device='cuda', dtype=torch.int32).contiguous()
    hw = _f32c(hw, torch)
    T_out = 1 if center_only else T
    out = torch.zeros((F, dmax, 4 * T_out + 1), dtype=torch.float32, device='cuda')
    counts_out = torch.empty((F,), dtype=torch.int32, device='cuda')
    sel = torch.full((F, dmax), -1, dtype=torch.int32, device='cuda')
    L.call('dt_prune_detections', L.ptr(boxes), F, dmax, ld, T, int(bool(center_only)), L.ptr(counts), L.ptr(hw),
           float(conf), float(min_area), L.ptr(out), L.ptr(counts_out), L.ptr(sel), L.stream_ptr())
    return out, counts_out, sel
