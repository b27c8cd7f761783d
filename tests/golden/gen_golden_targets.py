"""Golden vectors for the training target generators (oracle/targets.py), from the REFERENCE's own modules.  Build container
only (needs /root/reference):

    python tests/golden/gen_golden_targets.py      -> tests/golden/targets.npz

What runs, unmodified, from /root/reference/lib:
    roi_data/rpn.py            _get_field_of_anchors, _get_rpn_blobs
    datasets/json_dataset.py   add_proposals (_merge_proposal_boxes_into_roidb, _add_class_assignments)
    roi_data/fast_rcnn.py      _sample_rois (+ _compute_targets, _expand_bbox_targets)
    roi_data/keypoint_rcnn.py  add_keypoint_rcnn_blobs, _within_box
    utils/keypoints.py         keypoints_to_heatmap_labels
    utils/boxes.py             bbox_transform_inv, bbox_overlaps (compiled reference Cython, oracle/_ref)
Shims, all recorded here: gen_golden._setup_reference_imports (np.float/np.int aliases, compiled Cython, caffe2 /
pycocotools stubs) plus empty stub modules for cPickle (-> pickle), h5py and caffe2.proto / caffe2.python.* (imported at
module scope by utils/blob.py etc., never called on these paths).  The ONE behavioural patch: `numpy.random.choice` /
`numpy.random.randint` are replaced by the counter-based definitions of oracle/targets.py (`choice`, `randint`), keyed by
(seed, stream, image) — stream chosen from the calling function (see _stream) — because a device kernel cannot replay
numpy's Mersenne twister; everything else is the reference's arithmetic.
"""
import os
import pickle
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

CTX = dict(seed=3, image=0)


def _stub(name):
    parts = name.split('.')
    for i in range(1, len(parts) + 1):
        n = '.'.join(parts[:i])
        m = sys.modules.setdefault(n, types.ModuleType(n))
        if i > 1:
            setattr(sys.modules['.'.join(parts[:i - 1])], parts[i - 1], m)


def _stream():
    """Which of the reference's five random draws is calling (frame 2 = the reference function)."""
    f = sys._getframe(2)
    name = f.f_code.co_name
    if name == '_get_rpn_blobs':
        return None                      # resolved by the patched function itself (choice -> 0, randint -> 1)
    if name == '_sample_rois':
        return 3 if 'bg_inds' in f.f_locals else 2
    if name == 'add_keypoint_rcnn_blobs':
        return 4
    raise RuntimeError('unexpected random draw from ' + name)


def main():
    import gen_golden as g
    g._setup_reference_imports()
    sys.modules['cPickle'] = pickle
    for n in ['caffe2.proto.caffe2_pb2', 'caffe2.python.scope', 'caffe2.python.core', 'caffe2.python.workspace',
              'caffe2.python.utils', 'caffe2.python.muji', 'pycocotools.mask', 'pycocotools.coco', 'h5py']:
        _stub(n)
    sys.modules['pycocotools.coco'].COCO = object
    from oracle import targets as ot
    import numpy.random as npr

    def p_choice(a, size=None, replace=True, p=None):
        assert replace is False and p is None
        s = _stream()
        return ot.choice(a, size, CTX['seed'], 0 if s is None else s, CTX['image'])

    def p_randint(low, high=None, size=None, dtype=int):
        assert high is None and _stream() is None
        return ot.randint(low, size, CTX['seed'], 1, CTX['image'])
    npr.choice = p_choice
    npr.randint = p_randint
    np.random.choice = p_choice

    from core.config import cfg
    import roi_data.rpn as rrpn
    import roi_data.fast_rcnn as rfr
    import datasets.json_dataset as jd
    import scipy.sparse
    import utils.keypoints as kpu
    from modeling.generate_anchors import generate_anchors
    cfg.MODEL.KEYPOINTS_ON = True
    cfg.MODEL.NUM_CLASSES = 2
    cfg.KRCNN.NUM_KEYPOINTS = 17
    cfg.KRCNN.HEATMAP_SIZE = 56
    cfg.FPN.FPN_ON = True
    cfg.FPN.MULTILEVEL_RPN = True
    cfg.FPN.MULTILEVEL_ROIS = True
    cfg.TRAIN.MAX_SIZE = 320                       # field of anchors: 320 / stride per side (small maps, seconds on the CPU)
    K = 17
    rng = np.random.default_rng(20260925)
    out = dict(seed=np.int64(CTX['seed']))

    def synth_entry(H, W, G, crowd=False):
        x1 = rng.uniform(0, W - 60, G); y1 = rng.uniform(0, H - 60, G)
        w = rng.uniform(20, min(160, W) - 1, G); h = rng.uniform(30, min(200, H) - 1, G)
        boxes = np.stack([x1, y1, np.minimum(x1 + w, W - 1), np.minimum(y1 + h, H - 1)], 1).astype(np.float32)
        kps = np.zeros((G, 3, K), np.int32)
        for i in range(G):
            kps[i, 0] = rng.integers(int(boxes[i, 0]) - 5, int(boxes[i, 2]) + 6, K)
            kps[i, 1] = rng.integers(int(boxes[i, 1]) - 5, int(boxes[i, 3]) + 6, K)
            kps[i, 2] = rng.integers(0, 3, K)
            kps[i, 0, 0] = int(boxes[i, 2]); kps[i, 2, 0] = 2          # a joint on the right boundary
        kps[G - 1, 2, :] = 0                                          # one person without visible joints
        is_crowd = np.zeros(G, bool)
        if crowd:
            is_crowd[G - 2] = True
        gt_ov = np.zeros((G, 2), np.float32)
        for i in range(G):
            if is_crowd[i]:
                gt_ov[i, :] = -1.0
            else:
                gt_ov[i, 1] = 1.0
        return dict(boxes=boxes, gt_classes=np.ones(G, np.int32), is_crowd=is_crowd, gt_keypoints=kps,
                    gt_overlaps=scipy.sparse.csr_matrix(gt_ov), box_to_gt_ind_map=np.arange(G, dtype=np.int32),
                    seg_areas=np.zeros(G, np.float32), height=H, width=W)

    # ------------------------------------------------------------------ RPN targets (rpn.py:206-381)
    for tag, (H, W, G, scale, batch) in {'rpnA': (192, 256, 3, 1.25, 256), 'rpnB': (200, 300, 6, 1.0, 64), 'rpnC': (240, 320, 1, 1.0, 256)}.items():
        cfg.TRAIN.RPN_BATCH_SIZE_PER_IM = batch
        rrpn._threadlocal_foa.__dict__.pop('cache', None)
        e = synth_entry(H, W, G)
        foas = []
        for lvl in range(2, 7):
            foas.append(rrpn._get_field_of_anchors(2. ** lvl, (cfg.FPN.RPN_ANCHOR_START_SIZE * 2. ** (lvl - 2),), cfg.FPN.RPN_ASPECT_RATIOS, 1))
        all_anchors = np.concatenate([f.field_of_anchors for f in foas])
        im_h, im_w = np.round(H * scale), np.round(W * scale)
        gt = e['boxes'] * scale
        CTX['image'] = 1
        blobs = rrpn._get_rpn_blobs(im_h, im_w, foas, all_anchors, gt, np.full((G, 1), True))
        out[tag + '_gt'] = gt.astype(np.float32); out[tag + '_im'] = np.array([im_h, im_w, scale], np.float32)
        out[tag + '_field'] = np.array([f.field_size for f in foas], np.int32)
        out[tag + '_batch'] = np.int32(batch)
        for l, b in enumerate(blobs):
            out['%s_labels%d' % (tag, l)] = b['rpn_labels_int32_wide'][0].transpose(1, 2, 0)              # [H, W, A]
            for k, n in (('rpn_bbox_targets_wide', 'bt'), ('rpn_bbox_inside_weights_wide', 'iw'), ('rpn_bbox_outside_weights_wide', 'ow')):
                out['%s_%s%d' % (tag, n, l)] = b[k][0].transpose(1, 2, 0)                                   # [H, W, 4A]
    # tube anchors / tube ground truth (T = 3, the reference's 3-D RPN: rpn.py with time_dim > 1, per-frame visibility)
    for tag, (H, W, G, batch) in {'rpnT3': (200, 300, 4, 128)}.items():
        T = 3
        cfg.TRAIN.RPN_BATCH_SIZE_PER_IM = batch
        rrpn._threadlocal_foa.__dict__.pop('cache', None)
        e = synth_entry(H, W, G)
        tubes = np.concatenate([np.clip(e['boxes'] + rng.normal(0, 4, e['boxes'].shape), 0, [W - 1, H - 1, W - 1, H - 1]) for _ in range(T)], 1).astype(np.float32)
        vis = rng.uniform(0, 1, (G, T)) > 0.25
        vis[0] = True
        tubes[np.repeat(~vis, 4, axis=1)] = 0                        # utils/video._combine_clips leaves the boxes of missing frames at 0
        foas = [rrpn._get_field_of_anchors(2. ** lvl, (cfg.FPN.RPN_ANCHOR_START_SIZE * 2. ** (lvl - 2),), cfg.FPN.RPN_ASPECT_RATIOS, T)
                for lvl in range(2, 7)]
        all_anchors = np.concatenate([f.field_of_anchors for f in foas])
        CTX['image'] = 2
        blobs = rrpn._get_rpn_blobs(float(H), float(W), foas, all_anchors, tubes, vis)
        out[tag + '_gt'], out[tag + '_vis'] = tubes, vis
        out[tag + '_im'] = np.array([H, W, 1.0], np.float32)
        out[tag + '_field'] = np.array([f.field_size for f in foas], np.int32)
        out[tag + '_batch'] = np.int32(batch)
        for l, b in enumerate(blobs):
            out['%s_labels%d' % (tag, l)] = b['rpn_labels_int32_wide'][0].transpose(1, 2, 0)
            out['%s_vis%d' % (tag, l)] = b['rpn_vis_labels_int32_wide'][0].transpose(1, 2, 0)
            for k, n in (('rpn_bbox_targets_wide', 'bt'), ('rpn_bbox_inside_weights_wide', 'iw'), ('rpn_bbox_outside_weights_wide', 'ow')):
                out['%s_%s%d' % (tag, n, l)] = b[k][0].transpose(1, 2, 0)
        for lvl in range(2, 7):
            out['cell_anchors_T3_%d' % lvl] = generate_anchors(stride=2. ** lvl, sizes=(cfg.FPN.RPN_ANCHOR_START_SIZE * 2. ** (lvl - 2),),
                                                               aspect_ratios=cfg.FPN.RPN_ASPECT_RATIOS, time_dim=T)
    rrpn._threadlocal_foa.__dict__.pop('cache', None)
    for lvl in range(2, 7):
        out['cell_anchors%d' % lvl] = generate_anchors(stride=2. ** lvl, sizes=(cfg.FPN.RPN_ANCHOR_START_SIZE * 2. ** (lvl - 2),),
                                                       aspect_ratios=cfg.FPN.RPN_ASPECT_RATIOS, time_dim=1)

    # ------------------------------------------------------------------ RoI sampling + keypoint targets
    for tag, (H, W, G, P, scale, crowd, batch) in {'roiA': (300, 400, 4, 300, 1.6, False, 64), 'roiB': (240, 320, 3, 40, 2.0, True, 512),
                                                   'roiC': (240, 320, 2, 900, 1.0, False, 512)}.items():
        cfg.TRAIN.BATCH_SIZE_PER_IM = batch
        out[tag + '_batch'] = np.int32(batch)
        e0 = synth_entry(H, W, G, crowd)
        # proposals in blob coordinates: jittered copies of the gts (foreground) + random boxes
        gtb = e0['boxes']
        nf = P // 3
        src = gtb[rng.integers(0, G, nf)]
        jit = src + rng.normal(0, 6, (nf, 4)).astype(np.float32)
        x1 = rng.uniform(0, W - 20, P - nf); y1 = rng.uniform(0, H - 20, P - nf)
        rnd = np.stack([x1, y1, np.minimum(x1 + rng.uniform(8, 200, P - nf), W - 1), np.minimum(y1 + rng.uniform(8, 200, P - nf), H - 1)], 1)
        props = np.concatenate([jit, rnd], 0).astype(np.float32)
        props[:, 2:] = np.maximum(props[:, 2:], props[:, :2] + 1)
        props = props[rng.permutation(P)]
        rois = np.hstack([np.zeros((P, 1), np.float32), props * np.float32(scale)]).astype(np.float32)
        im_scales = np.array([scale], np.float32)
        roidb = [dict((k, (v.copy() if hasattr(v, 'copy') else v)) for k, v in e0.items())]
        jd.add_proposals(roidb, rois, im_scales)
        CTX['image'] = 0
        blobs = rfr._sample_rois(roidb[0], im_scales[0], 0)
        for k in ('boxes', 'gt_classes', 'is_crowd', 'gt_keypoints'):
            out['%s_gt_%s' % (tag, k)] = e0[k]
        out[tag + '_rois_in'] = rois
        out[tag + '_scale'] = np.float32(scale)
        out[tag + '_max_overlaps'] = roidb[0]['max_overlaps']
        out[tag + '_max_classes'] = roidb[0]['max_classes']
        out[tag + '_box_to_gt'] = roidb[0]['box_to_gt_ind_map']
        for k in ('rois', 'labels_int32', 'bbox_targets', 'bbox_inside_weights', 'bbox_outside_weights', 'keypoint_rois',
                  'keypoint_locations_int32', 'keypoint_weights'):
            out['%s_%s' % (tag, k)] = blobs[k]
    # tubes (T = 3): tube proposals, tube IoU, 4T box targets per class, per-frame heat-map labels
    for tag, (H, W, G, P, scale, batch) in {'roiT3': (240, 320, 3, 200, 1.5, 128)}.items():
        T = 3
        cfg.TRAIN.BATCH_SIZE_PER_IM = batch
        out[tag + '_batch'] = np.int32(batch)
        e0 = synth_entry(H, W, G)
        jit = lambda b, s_: np.clip(b + rng.normal(0, s_, b.shape), 0, [W - 1, H - 1, W - 1, H - 1]).astype(np.float32)
        e0['boxes'] = np.concatenate([jit(e0['boxes'], 3) for _ in range(T)], 1)
        kp1 = e0['gt_keypoints']
        e0['gt_keypoints'] = np.concatenate([kp1 + rng.integers(-2, 3, kp1.shape).astype(np.int32) * np.array([1, 1, 0], np.int32)[None, :, None]
                                             for _ in range(T)], 2)
        nf = P // 3
        src = e0['boxes'][rng.integers(0, G, nf)]
        props = np.concatenate([src + rng.normal(0, 5, src.shape),
                                np.tile(np.stack([rng.uniform(0, W - 40, P - nf), rng.uniform(0, H - 40, P - nf),
                                                  rng.uniform(0, W - 40, P - nf) + 40, rng.uniform(0, H - 40, P - nf) + 40], 1), [1, T])], 0).astype(np.float32)
        for t in range(T):
            props[:, 4 * t + 2:4 * t + 4] = np.maximum(props[:, 4 * t + 2:4 * t + 4], props[:, 4 * t:4 * t + 2] + 1)
        props = props[rng.permutation(P)]
        rois = np.hstack([np.zeros((P, 1), np.float32), props * np.float32(scale)]).astype(np.float32)
        im_scales = np.array([scale], np.float32)
        roidb = [dict((k, (v.copy() if hasattr(v, 'copy') else v)) for k, v in e0.items())]
        jd.add_proposals(roidb, rois, im_scales)
        CTX['image'] = 0
        blobs = rfr._sample_rois(roidb[0], im_scales[0], 0)
        for k in ('boxes', 'gt_classes', 'is_crowd', 'gt_keypoints'):
            out['%s_gt_%s' % (tag, k)] = e0[k]
        out[tag + '_rois_in'] = rois
        out[tag + '_scale'] = np.float32(scale)
        for k in ('rois', 'labels_int32', 'bbox_targets', 'bbox_inside_weights', 'bbox_outside_weights', 'keypoint_rois',
                  'keypoint_locations_int32', 'keypoint_weights'):
            out['%s_%s' % (tag, k)] = blobs[k]
    # heat-map labels alone (utils/keypoints.py:152-207), incl. boundary and out-of-box joints
    n = 64
    rois = np.stack([rng.uniform(0, 200, n), rng.uniform(0, 200, n), rng.uniform(210, 400, n), rng.uniform(210, 400, n)], 1).astype(np.float32)
    rois[:8] = np.round(rois[:8])
    kps = np.stack([rng.integers(-10, 420, (n, K)), rng.integers(-10, 420, (n, K)), rng.integers(0, 3, (n, K))], 1).astype(np.int32)
    kps[:8, 0, 0] = rois[:8, 2].astype(np.int32); kps[:8, 1, 1] = rois[:8, 3].astype(np.int32)
    heat, wts = kpu.keypoints_to_heatmap_labels(kps, rois)
    out['heat_rois'], out['heat_kps'], out['heat_loc'], out['heat_w'] = rois, kps, heat, wts
    np.savez_compressed(os.path.join(HERE, 'targets.npz'), **out)
    print('wrote targets.npz:', len(out), 'arrays')


if __name__ == '__main__':
    main()
