"""Oracle: the training target generators (TEST INFRASTRUCTURE ONLY; never imported by the product).

CPU / numpy restatement of
    lib/roi_data/rpn.py:206-381            _get_rpn_blobs        (RPN anchor labels / box targets / weights)
    lib/datasets/json_dataset.py:423-534   _merge_proposal_boxes_into_roidb + _add_class_assignments + add_proposals
    lib/roi_data/fast_rcnn.py:118-238      _sample_rois, _compute_targets, _expand_bbox_targets
    lib/roi_data/keypoint_rcnn.py:24-99    add_keypoint_rcnn_blobs, _within_box
    lib/utils/keypoints.py:152-207         keypoints_to_heatmap_labels
    lib/utils/boxes.py:205-230             bbox_transform_inv
    lib/ops/collect_and_distribute_fpn_rpn_proposals.py:44-62   collect (training: top-N over the WHOLE batch)

The reference draws its random subsets from numpy's global Mersenne twister (npr.choice / npr.randint), which a device
kernel cannot replay.  The product (csrc/targets.cu) and this oracle therefore share ONE documented counter-based
generator, `hash_u32(seed, stream, image, i)` (splitmix64 finaliser), and define

    choice(a, k)   = the k elements of `a` with the smallest (hash(a_i), a_i), in that order   (npr.choice, replace=False)
    randint(n, k)  = [(hash(j) * n) >> 32 for j < k]                                           (npr.randint(n, size=k))

streams: 0 RPN fg disable, 1 RPN bg enable, 2 RoI fg, 3 RoI bg, 4 keypoint RoIs.  Pinning: tests/golden/targets.npz holds
outputs of the REFERENCE's own functions run with npr.choice / npr.randint patched to exactly these two definitions
(tests/golden/gen_golden_targets.py); everything else in those functions runs unmodified.
"""
import numpy as np

from . import boxes as obox

M64 = (1 << 64) - 1


def hash_u32(seed, stream, image, i):
    """splitmix64 finaliser over a linear combination of the four counters -> the high 32 bits (vectorised over i)."""
    i = np.asarray(i, dtype=np.uint64)
    with np.errstate(over='ignore'):
        x = (np.uint64((int(seed) * 0x9E3779B97F4A7C15) & M64) + np.uint64((int(stream) * 0xBF58476D1CE4E5B9) & M64) +
             np.uint64((int(image) * 0x94D049BB133111EB) & M64) + i)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(32)).astype(np.uint32)


def choice(a, size, seed, stream, image):
    a = np.asarray(a)
    keys = hash_u32(seed, stream, image, a).astype(np.uint64)
    order = np.lexsort((a, keys))
    return a[order[:int(size)]]


def randint(n, size, seed, stream, image):
    h = hash_u32(seed, stream, image, np.arange(int(size))).astype(np.uint64)
    return ((h * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def _transform_inv_frame(ex, gt, weights, dt):
    one, half = dt(1.0), dt(0.5)
    ew = ex[:, 2] - ex[:, 0] + one
    eh = ex[:, 3] - ex[:, 1] + one
    ecx = ex[:, 0] + half * ew
    ecy = ex[:, 1] + half * eh
    gw = gt[:, 2] - gt[:, 0] + one
    gh = gt[:, 3] - gt[:, 1] + one
    gcx = gt[:, 0] + half * gw
    gcy = gt[:, 1] + half * gh
    wx, wy, ww, wh = [dt(w) for w in weights]
    return np.vstack((wx * (gcx - ecx) / ew, wy * (gcy - ecy) / eh, ww * np.log(gw / ew), wh * np.log(gh / eh))).transpose()


def bbox_transform_inv(ex, gt, weights):
    """utils/boxes.py:205-240 -> fp32 targets.  Boxes [n, 4]: fp32 arithmetic (numpy weak python-float scalars).  Tubes
    [n, 4T] go through split_tube_into_boxes (:28-58), whose np.hstack with an empty fp64 score column promotes every frame
    to fp64, so the tube targets are fp64 arithmetic on the fp32 coordinates, rounded to fp32 once (rpn.py:384-387)."""
    ex = np.asarray(ex, np.float32); gt = np.asarray(gt, np.float32)
    if ex.shape[1] > 4:
        e64, g64 = ex.astype(np.float64), gt.astype(np.float64)
        return np.concatenate([_transform_inv_frame(e64[:, 4 * t:4 * t + 4], g64[:, 4 * t:4 * t + 4], weights, np.float64)
                               for t in range(ex.shape[1] // 4)], axis=1).astype(np.float32)
    return _transform_inv_frame(ex, gt, weights, np.float32)


def field_of_anchors(cell_anchors, stride, H, W):
    """rpn.py:160-203 on an H x W grid: [(H*W*A), 4T] fp32, enumeration (h, w, a); cell anchors [A, 4T] (tube anchors are the
    2-D anchor replicated over the T frames, generate_anchors.py:64-67, and so are the shifts)."""
    sx, sy = np.meshgrid(np.arange(W) * stride, np.arange(H) * stride)
    shifts = np.vstack((sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel())).transpose()
    A, D = cell_anchors.shape
    shifts = np.tile(shifts, [1, D // 4])
    f = cell_anchors.reshape((1, A, D)) + shifts.reshape((1, -1, D)).transpose((1, 0, 2))
    return f.reshape((-1, D)).astype(np.float32)


def rpn_targets(levels, gt_boxes, im_h, im_w, seed, image, straddle=0.0, pos=0.7, neg=0.3, batch=256, fg_frac=0.5, visible=None):
    """rpn.py:206-381 for one image.  levels: list of (cell_anchors [A,4T], stride, H, W); gt_boxes [G,4T] fp32 already scaled;
    visible [G,T] bool (track_visible; all True when None).  Returns per level dict(labels [H,W,A] i32, bbox_targets / inside /
    outside [H,W,4T*A] f32, vis_labels [H,W,T*A] i32) and the pre-sampling diagnostics dict(max, argmax, fg, bgcand)."""
    all_anchors = np.concatenate([field_of_anchors(c, s, H, W) for (c, s, H, W) in levels])
    total = all_anchors.shape[0]
    gt_boxes = np.asarray(gt_boxes, np.float32)
    T = all_anchors.shape[1] // 4
    visible = np.full((len(gt_boxes), T), True) if visible is None else np.asarray(visible, bool)
    inside = np.where(np.all(all_anchors[:, 0::4] >= -straddle, axis=1) & np.all(all_anchors[:, 1::4] >= -straddle, axis=1) &
                      np.all(all_anchors[:, 2::4] < im_w + straddle, axis=1) & np.all(all_anchors[:, 3::4] < im_h + straddle, axis=1))[0]
    anchors = all_anchors[inside]
    n = len(inside)
    labels = np.full((n,), -1, np.int32)
    ov = obox.bbox_overlaps(anchors, gt_boxes)
    argmax = ov.argmax(axis=1)
    amax = ov[np.arange(n), argmax]
    gmax = ov[ov.argmax(axis=0), np.arange(ov.shape[1])]
    labels[np.where(ov == gmax)[0]] = 1
    labels[amax >= np.float32(pos)] = 1
    diag = dict(inside=inside, max=amax.copy(), argmax=argmax.copy(), fg=np.where(labels == 1)[0])
    num_fg = int(fg_frac * batch)
    fg_inds = np.where(labels == 1)[0]
    if len(fg_inds) > num_fg:
        labels[choice(fg_inds, len(fg_inds) - num_fg, seed, 0, image)] = -1
    fg_inds = np.where(labels == 1)[0]
    num_bg = batch - np.sum(labels == 1)
    bg_inds = np.where(amax < np.float32(neg))[0]
    diag['bgcand'] = bg_inds.copy()
    if len(bg_inds) > num_bg:
        labels[bg_inds[randint(len(bg_inds), num_bg, seed, 1, image)]] = 0
    bt = np.zeros((n, 4 * T), np.float32)
    bt[fg_inds] = bbox_transform_inv(anchors[fg_inds], gt_boxes[argmax[fg_inds]], (1.0, 1.0, 1.0, 1.0))
    iw = np.zeros((n, 4 * T), np.float32)
    iw[fg_inds] = np.repeat(visible[argmax[fg_inds]], 4, axis=1).astype(np.float32)     # invisible frames of a tube: no box loss
    ow = np.zeros((n, 4 * T), np.float32)
    nex = np.sum(labels >= 0)
    ow[labels == 1] = 1.0 / nex
    ow[labels == 0] = 1.0 / nex
    vis = np.tile(labels[:, None], [1, T])
    vis[fg_inds] *= visible[argmax[fg_inds]].astype(np.int32)

    def unmap(d, fill):
        r = np.full((total,) + d.shape[1:], fill, d.dtype)
        r[inside] = d
        return r
    labels, bt, iw, ow, vis = unmap(labels, -1), unmap(bt, 0), unmap(iw, 0), unmap(ow, 0), unmap(vis, -1)
    out, s = [], 0
    for (c, st, H, W) in levels:
        A = c.shape[0]
        e = s + H * W * A
        out.append(dict(labels=labels[s:e].reshape(H, W, A), bbox_targets=bt[s:e].reshape(H, W, 4 * T * A),
                        inside=iw[s:e].reshape(H, W, 4 * T * A), outside=ow[s:e].reshape(H, W, 4 * T * A),
                        vis_labels=vis[s:e].reshape(H, W, T * A)))
        s = e
    return out, diag


def collect_train(rois_per_image, scores_per_image, post_nms_topn):
    """collect() in training (collect_and_distribute_fpn_rpn_proposals.py:44-62): the top post_nms_topN of the WHOLE batch.
    Inputs are each image's proposals in descending score; ties across images resolve to the lower image index (the
    reference's np.argsort on ties is unspecified).  Returns the kept rows per image (order preserved)."""
    sc = np.concatenate([np.asarray(s, np.float32) for s in scores_per_image]) if scores_per_image else np.zeros((0,), np.float32)
    img = np.concatenate([np.full(len(s), b) for b, s in enumerate(scores_per_image)])
    rank = np.concatenate([np.arange(len(s)) for s in scores_per_image])
    order = np.lexsort((rank, img, -sc.astype(np.float64)))
    keep = np.zeros(len(sc), bool)
    keep[order[:post_nms_topn]] = True
    out, o = [], 0
    for b, r in enumerate(rois_per_image):
        out.append(np.asarray(r)[keep[o:o + len(r)]])
        o += len(r)
    return out


def merge_proposals(gt_boxes, gt_classes, is_crowd, proposals):
    """json_dataset.py:423-512 for one entry whose roidb rows are its gt boxes: boxes = [gt ; proposals], max_overlaps,
    max_classes, box_to_gt_ind_map."""
    gt_boxes = np.asarray(gt_boxes, np.float32)
    G = gt_boxes.shape[0]
    P = proposals.shape[0]
    ncls = int(max(2, gt_classes.max() + 1)) if G else 2
    gt_ov = np.zeros((G + P, ncls), np.float32)
    bmap = -np.ones((G + P,), np.int32)
    for i in range(G):                                   # the dataset's own rows (json_dataset.py:303-323)
        if is_crowd[i]:
            gt_ov[i, :] = -1.0
        else:
            gt_ov[i, gt_classes[i]] = 1.0
        bmap[i] = i
    if G > 0 and P > 0:
        ov = obox.bbox_overlaps(proposals.astype(np.float32), gt_boxes)
        am = ov.argmax(axis=1)
        mx = ov.max(axis=1)
        I = np.where(mx > 0)[0]
        gt_ov[G + I, gt_classes[am[I]]] = mx[I]
        bmap[G + I] = am[I]
    boxes = np.concatenate([gt_boxes, proposals.astype(np.float32)], 0)
    return boxes, gt_ov.max(axis=1), gt_ov.argmax(axis=1), bmap


def within_box(points, boxes):
    x = np.logical_and(points[:, 0, :] >= boxes[:, 0:1], points[:, 0, :] <= boxes[:, 2:3])
    y = np.logical_and(points[:, 1, :] >= boxes[:, 1:2], points[:, 1, :] <= boxes[:, 3:4])
    return np.logical_and(x, y)


def keypoints_to_heatmap_labels(kps, rois, M):
    """utils/keypoints.py:152-207.  kps [n,3,K] int32, rois [n,4] fp32 -> (locations [n,K] f32, weights [n,K] f32)."""
    n, _, K = kps.shape
    heat = np.zeros((n, K), np.float32); wts = np.zeros((n, K), np.float32)
    ox, oy = rois[:, 0], rois[:, 1]
    sx = np.float32(M) / (rois[:, 2] - rois[:, 0] + np.float32(1))
    sy = np.float32(M) / (rois[:, 3] - rois[:, 1] + np.float32(1))
    for k in range(K):
        vis = kps[:, 2, k] > 0
        x = kps[:, 0, k].astype(np.float32); y = kps[:, 1, k].astype(np.float32)
        xb = np.where(x == rois[:, 2])[0]; yb = np.where(y == rois[:, 3])[0]
        x = np.floor((x - ox) * sx); x[xb] = M - 1
        y = np.floor((y - oy) * sy); y[yb] = M - 1
        valid = np.logical_and(np.logical_and(np.logical_and(x >= 0, y >= 0), np.logical_and(x < M, y < M)), vis).astype(np.int32)
        heat[:, k] = (y * np.float32(M) + x) * valid
        wts[:, k] = valid
    return heat, wts


def sample_rois(gt_boxes, gt_classes, is_crowd, gt_keypoints, proposals_scaled, im_scale, image, seed, num_classes=2,
                batch=512, fg_frac=0.25, fg_thresh=0.5, bg_hi=0.5, bg_lo=0.0, weights=(10., 10., 5., 5.), M=56):
    """add_proposals + _sample_rois + add_keypoint_rcnn_blobs for one image.  proposals_scaled [P,4T] fp32 in blob coordinates
    (the RPN's rois / tubes); gt_boxes [G,4T]; gt_keypoints [G,3,K*T]; im_scale fp32.  Returns the blobs of
    fast_rcnn.py:118-186 / keypoint_rcnn.py:24-83 (tubes: box targets 4T per class, heat-map labels frame by frame)."""
    s = np.float32(im_scale)
    inv = np.float32(1.0) / s
    props = (np.asarray(proposals_scaled, np.float32) * inv).astype(np.float32)
    boxes, max_ov, max_cls, bmap = merge_proposals(gt_boxes, gt_classes, is_crowd, props)
    G = len(gt_boxes)
    fg_per = int(np.round(fg_frac * batch))
    fg = np.where(max_ov >= np.float32(fg_thresh))[0]
    nfg = min(fg_per, fg.size)
    if fg.size > 0:
        fg = choice(fg, nfg, seed, 2, image)
    bg = np.where((max_ov < np.float32(bg_hi)) & (max_ov >= np.float32(bg_lo)))[0]
    nbg = min(batch - nfg, bg.size)
    if bg.size > 0:
        bg = choice(bg, nbg, seed, 3, image)
    keep = np.append(fg, bg).astype(np.int64)
    labels = max_cls[keep].copy()
    labels[nfg:] = 0
    sb = boxes[keep]
    gt_assign = bmap[keep]                               # gt_inds[...]: the gt rows are rows 0..G-1 (negative index -> last gt)
    tg = bbox_transform_inv(sb, np.asarray(gt_boxes, np.float32)[gt_assign], weights).astype(np.float32)
    D = boxes.shape[1]                                   # tube_dim = 4T (fast_rcnn.py:213)
    T = D // 4
    bt = np.zeros((len(keep), D * num_classes), np.float32)
    iw = np.zeros_like(bt)
    for i in np.where(labels > 0)[0]:
        c = int(labels[i])
        bt[i, D * c:D * c + D] = tg[i]
        iw[i, D * c:D * c + D] = 1.0
    ow = (iw > 0).astype(np.float32)
    rois = np.hstack((np.full((len(keep), 1), image, np.float32), sb * s)).astype(np.float32)
    # keypoints (keypoint_rcnn.py:24-83)
    kp = np.asarray(gt_keypoints)
    ind_kp = bmap.copy()
    ind_kp[ind_kp < 0] += G                              # python negative indexing of gt_inds[-1]
    wb = within_box(kp[ind_kp], boxes)
    vis = kp[ind_kp, 2, :] > 0
    is_vis = np.sum(np.logical_and(vis, wb), axis=1) > 0
    kfg = np.where(np.logical_and(max_ov >= np.float32(fg_thresh), is_vis))[0]
    nk = min(fg_per, kfg.size)
    if kfg.size > nk:
        kfg = choice(kfg, nk, seed, 4, image)
    if kfg.shape[0] == 0:
        kfg = np.arange(G)
    kb = boxes[kfg]
    km = bmap[kfg]
    sk = -np.ones((len(kb), 3, kp.shape[2]), kp.dtype)
    for ii in range(len(kb)):
        if km[ii] >= 0:
            sk[ii] = kp[km[ii]]
    Kf = kp.shape[2] // T                                # per-frame joints (keypoint_rcnn.py:62-70)
    hw = [keypoints_to_heatmap_labels(sk[..., t * Kf:(t + 1) * Kf], kb[:, 4 * t:4 * t + 4], M) for t in range(T)]
    heat = np.concatenate([h for h, _ in hw], axis=-1)
    wts = np.concatenate([w for _, w in hw], axis=-1)
    krois = np.hstack((np.full((len(kb), 1), image, np.float32), kb * s)).astype(np.float32)
    return dict(rois=rois, labels=labels.astype(np.int32), bbox_targets=bt, inside=iw, outside=ow, keypoint_rois=krois,
                keypoint_locations=heat.astype(np.int32), keypoint_weights=wts,
                diag=dict(boxes=boxes, max_overlaps=max_ov, max_classes=max_cls, box_to_gt=bmap))
