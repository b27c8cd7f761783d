"""GPU: the device target generators (csrc/targets.cu) against (a) outputs of the REFERENCE's own lib/roi_data functions
(tests/golden/targets.npz) and (b) the oracle restatement on fresh random batches.  Everything is compared bit for bit except
the two log() columns of bbox_transform_inv: numpy's SIMD float32 log is itself only ~4-ulp accurate (documented max 3.83),
CUDA's logf 1 ulp, so those columns are compared to 8 ulp."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import targets as ot

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'targets.npz'))
SEED = int(G['seed'])


def _ulp_close(a, b, ulps=8):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    return np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)))


def _cmp_box_targets(got, ref, name):
    got = got.reshape(-1, 4); ref = ref.reshape(-1, 4)
    assert np.array_equal(got[:, :2], ref[:, :2]), name + ' dx/dy'
    assert _ulp_close(got[:, 2:], ref[:, 2:]), name + ' dw/dh'


class _TrainCfg(object):
    RPN_STRADDLE_THRESH = 0
    RPN_POSITIVE_OVERLAP = 0.7
    RPN_NEGATIVE_OVERLAP = 0.3
    RPN_FG_FRACTION = 0.5

    def __init__(self, batch):
        self.RPN_BATCH_SIZE_PER_IM = batch


def _anchors(torch):
    return [torch.from_numpy(np.ascontiguousarray(G['cell_anchors%d' % lvl], dtype=np.float64)).cuda() for lvl in range(2, 7)]


@pytest.mark.parametrize('tag', ['rpnA', 'rpnB', 'rpnC'])
def test_rpn_targets_equal_reference_goldens(tag):
    import torch
    from detectandtrack_b200.ops import target_ops
    gt = G[tag + '_gt']                                   # already in blob coordinates -> scale 1 on the device
    im_h, im_w, _ = G[tag + '_im']
    field = [int(f) for f in G[tag + '_field']]
    Gn = gt.shape[0]
    boxes = np.zeros((2, 8, 4), np.float32); boxes[:, :Gn] = gt
    out = target_ops.rpn_targets([(f, f) for f in field], _anchors(torch), [2. ** l for l in range(2, 7)], 3,
                                 torch.from_numpy(boxes).cuda(), torch.tensor([Gn, Gn], dtype=torch.int32).cuda(),
                                 torch.tensor([[im_h, im_w, 1.0]] * 2, dtype=torch.float32).cuda(), _TrainCfg(int(G[tag + '_batch'])), SEED)
    for l, o in enumerate(out):
        assert np.array_equal(o['labels'][1].cpu().numpy(), G['%s_labels%d' % (tag, l)]), l      # the goldens were drawn for image 1
        _cmp_box_targets(o['bbox_targets'][1].cpu().numpy(), G['%s_bt%d' % (tag, l)], 'level %d' % l)
        assert np.array_equal(o['inside'][1].cpu().numpy(), G['%s_iw%d' % (tag, l)])
        assert np.array_equal(o['outside'][1].cpu().numpy(), G['%s_ow%d' % (tag, l)])
    # image 0 has the same boxes but its own draws
    assert any(not np.array_equal(o['labels'][0].cpu().numpy(), o['labels'][1].cpu().numpy()) for o in out)


def test_rpn_targets_vs_oracle_fresh_batch_with_scale():
    import torch
    from detectandtrack_b200.ops import target_ops
    rng = np.random.default_rng(5)
    shapes = [(48, 64), (24, 32), (12, 16), (6, 8), (3, 4)]
    B, Gmax = 3, 8
    scale = np.float32(1.3)
    boxes = np.zeros((B, Gmax, 4), np.float32); counts = np.array([1, 4, 7], np.int32)
    for b in range(B):
        x1 = rng.uniform(0, 120, counts[b]); y1 = rng.uniform(0, 80, counts[b])
        boxes[b, :counts[b]] = np.stack([x1, y1, x1 + rng.uniform(8, 70, counts[b]), y1 + rng.uniform(8, 60, counts[b])], 1)
    im = np.array([[140, 190, scale]] * B, np.float32)
    out = target_ops.rpn_targets(shapes, _anchors(torch), [2. ** l for l in range(2, 7)], 3, torch.from_numpy(boxes).cuda(),
                                 torch.from_numpy(counts).cuda(), torch.from_numpy(im).cuda(), _TrainCfg(64), 11)
    for b in range(B):
        lv = [(G['cell_anchors%d' % lvl], 2. ** lvl, H, W) for lvl, (H, W) in zip(range(2, 7), shapes)]
        ref, _ = ot.rpn_targets(lv, boxes[b, :counts[b]] * scale, 140., 190., 11, b, batch=64)
        for l, o in enumerate(out):
            assert np.array_equal(o['labels'][b].cpu().numpy(), ref[l]['labels']), (b, l)
            _cmp_box_targets(o['bbox_targets'][b].cpu().numpy(), ref[l]['bbox_targets'], 'img %d level %d' % (b, l))
            assert np.array_equal(o['inside'][b].cpu().numpy(), ref[l]['inside'])
            assert np.array_equal(o['outside'][b].cpu().numpy(), ref[l]['outside'])


class _Cfg(object):
    class MODEL:
        NUM_CLASSES = 2
        BBOX_REG_WEIGHTS = (10., 10., 5., 5.)

    class KRCNN:
        NUM_KEYPOINTS = 17
        HEATMAP_SIZE = 56

    def __init__(self, batch, topn=100000):
        class TRAIN:
            BATCH_SIZE_PER_IM = batch
            FG_FRACTION = 0.25
            FG_THRESH = 0.5
            BG_THRESH_HI = 0.5
            BG_THRESH_LO = 0.0
            RPN_POST_NMS_TOP_N = topn
        self.TRAIN = TRAIN


def _gt_tensors(torch, boxes, classes, crowd, kps, Gmax=8):
    B = len(boxes)
    gb = np.zeros((B, Gmax, 4), np.float32); gc = np.zeros((B, Gmax), np.int32); gcr = np.zeros((B, Gmax), np.int32)
    gk = np.zeros((B, Gmax, 3, 17), np.int32); cnt = np.zeros((B,), np.int32)
    for b in range(B):
        g = len(boxes[b]); cnt[b] = g
        gb[b, :g] = boxes[b]; gc[b, :g] = classes[b]; gcr[b, :g] = crowd[b]; gk[b, :g] = kps[b]
    t = lambda a: torch.from_numpy(a).cuda()
    return dict(boxes=t(gb), classes=t(gc), crowd=t(gcr), keypoints=t(gk), counts=t(cnt))


def _check_sampled(o, b, ref, batch):
    n = len(ref['rois'])
    assert int(o['counts'][b]) == n
    assert np.array_equal(o['rois'][b, :n].cpu().numpy(), ref['rois'])
    lab = o['labels'][b].cpu().numpy()
    assert np.array_equal(lab[:n], ref['labels']) and np.all(lab[n:] == -1)
    bt = o['bbox_targets'][b, :n].cpu().numpy()
    for c in range(2):
        _cmp_box_targets(bt[:, 4 * c:4 * c + 4], ref['bbox_targets'][:, 4 * c:4 * c + 4], 'class %d' % c)
    assert np.array_equal(o['inside'][b, :n].cpu().numpy(), ref['inside']) and np.array_equal(o['outside'][b, :n].cpu().numpy(), ref['outside'])
    assert float(o['inside'][b, n:].abs().sum()) == 0 and float(o['outside'][b, n:].abs().sum()) == 0
    nk = len(ref['keypoint_rois'])
    assert int(o['kp_counts'][b]) == nk
    assert np.array_equal(o['kp_rois'][b, :nk].cpu().numpy(), ref['keypoint_rois'])
    assert np.array_equal(o['kp_locations'][b, :nk].cpu().numpy(), ref['keypoint_locations'].reshape(nk, 17))
    assert np.array_equal(o['kp_weights'][b, :nk].cpu().numpy(), ref['keypoint_weights'].reshape(nk, 17))
    assert float(o['kp_weights'][b, nk:].abs().sum()) == 0


@pytest.mark.parametrize('tag', ['roiA', 'roiB', 'roiC'])
def test_sample_rois_equal_reference_goldens(tag):
    import torch
    from detectandtrack_b200.ops import target_ops
    rois_in = G[tag + '_rois_in']
    P = rois_in.shape[0]
    batch = int(G[tag + '_batch'])
    gt = _gt_tensors(torch, [G[tag + '_gt_boxes']], [G[tag + '_gt_gt_classes']], [G[tag + '_gt_is_crowd'].astype(np.int32)], [G[tag + '_gt_gt_keypoints']])
    scores = np.linspace(1.0, 0.1, P).astype(np.float32)[None]
    o = target_ops.sample_rois(torch.from_numpy(rois_in[None].copy()).cuda(), torch.from_numpy(scores).cuda(),
                               torch.tensor([P], dtype=torch.int32).cuda(), gt,
                               torch.tensor([[0, 0, float(G[tag + '_scale'])]], dtype=torch.float32).cuda(), _Cfg(batch), SEED)
    ref = dict(rois=G[tag + '_rois'], labels=G[tag + '_labels_int32'], bbox_targets=G[tag + '_bbox_targets'], inside=G[tag + '_bbox_inside_weights'],
               outside=G[tag + '_bbox_outside_weights'], keypoint_rois=G[tag + '_keypoint_rois'],
               keypoint_locations=G[tag + '_keypoint_locations_int32'], keypoint_weights=G[tag + '_keypoint_weights'])
    _check_sampled(o, 0, ref, batch)
    tot = o['totals'].cpu().numpy()
    assert tot[0] == len(ref['rois']) and tot[1] == ref['keypoint_weights'].sum()


def test_sample_rois_batch_wide_topn_vs_oracle():
    """Two images, collect's training branch keeps the top-N of the WHOLE batch; one image has no visible foreground RoI."""
    import torch
    from detectandtrack_b200.ops import target_ops
    rng = np.random.default_rng(9)
    B, R, topn, batch = 2, 300, 350, 128
    gts, cls, crowd, kps, rois, scores, cnts = [], [], [], [], [], [], []
    scale = np.float32(1.7)
    for b in range(B):
        g = 3 + b
        x1 = rng.uniform(0, 300, g); y1 = rng.uniform(0, 200, g)
        gb = np.stack([x1, y1, x1 + rng.uniform(30, 150, g), y1 + rng.uniform(40, 180, g)], 1).astype(np.float32)
        k = np.zeros((g, 3, 17), np.int32)
        for i in range(g):
            k[i, 0] = rng.integers(int(gb[i, 0]), int(gb[i, 2]) + 1, 17); k[i, 1] = rng.integers(int(gb[i, 1]), int(gb[i, 3]) + 1, 17)
            k[i, 2] = rng.integers(0, 3, 17) if b == 0 else 0          # image 1: nothing visible -> trains on its gt boxes
        n = 250 + 40 * b
        src = gb[rng.integers(0, g, n)] + rng.normal(0, 8, (n, 4)).astype(np.float32)
        src[n // 2:] = np.stack([rng.uniform(0, 400, n - n // 2), rng.uniform(0, 300, n - n // 2), rng.uniform(410, 500, n - n // 2), rng.uniform(310, 400, n - n // 2)], 1)
        src[:, 2:] = np.maximum(src[:, 2:], src[:, :2] + 1)
        r = np.zeros((R, 5), np.float32); r[:, 0] = b; r[:n, 1:] = src * scale
        sc = np.zeros((R,), np.float32); sc[:n] = np.sort(rng.uniform(0, 1, n).astype(np.float32))[::-1]
        gts.append(gb); cls.append(np.ones(g, np.int32)); crowd.append(np.zeros(g, np.int32)); kps.append(k)
        rois.append(r); scores.append(sc); cnts.append(n)
    gt = _gt_tensors(torch, gts, cls, crowd, kps)
    o = target_ops.sample_rois(torch.from_numpy(np.stack(rois)).cuda(), torch.from_numpy(np.stack(scores)).cuda(),
                               torch.tensor(cnts, dtype=torch.int32).cuda(), gt, torch.tensor([[0, 0, float(scale)]] * B, dtype=torch.float32).cuda(),
                               _Cfg(batch, topn), 21)
    kept = ot.collect_train([rois[b][:cnts[b]] for b in range(B)], [scores[b][:cnts[b]] for b in range(B)], topn)
    assert sum(len(k) for k in kept) == topn and all(0 < len(k) < cnts[b] for b, k in enumerate(kept))
    for b in range(B):
        ref = ot.sample_rois(gts[b], cls[b], crowd[b], kps[b], kept[b][:, 1:], scale, b, 21, batch=batch)
        _check_sampled(o, b, ref, batch)
    assert int(o['kp_counts'][1]) == len(gts[1])                      # the no-visible-keypoint fallback (keypoint_rcnn.py:45-46)


def test_targets_edge_cases_empty_inputs():
    """Empty inputs: an image without ground truth (the reference's loader filters those out; here every anchor is ignored
    and every proposal is background) and an image without proposals (its gt boxes are the only RoIs), in one batch."""
    import torch
    from detectandtrack_b200.ops import target_ops
    rng = np.random.default_rng(3)
    shapes = [(24, 32), (12, 16), (6, 8), (3, 4), (2, 2)]
    gb = np.zeros((2, 8, 4), np.float32)
    gb[1, :2] = [[10, 12, 60, 90], [40, 20, 100, 80]]
    counts = np.array([0, 2], np.int32)
    im = np.array([[96, 128, 1.0]] * 2, np.float32)
    out = target_ops.rpn_targets(shapes, _anchors(torch), [2. ** l for l in range(2, 7)], 3, torch.from_numpy(gb).cuda(),
                                 torch.from_numpy(counts).cuda(), torch.from_numpy(im).cuda(), _TrainCfg(64), 1)
    for o in out:
        assert int((o['labels'][0] != -1).sum()) == 0 and float(o['outside'][0].abs().sum()) == 0 and float(o['inside'][0].abs().sum()) == 0
    assert sum(int((o['labels'][1] == 1).sum()) for o in out) > 0
    # RoI sampling: image 0 has proposals but no gt, image 1 has gt but no proposals
    R = 50
    x1 = rng.uniform(0, 60, R); y1 = rng.uniform(0, 40, R)
    rois = np.zeros((2, R, 5), np.float32); rois[1, :, 0] = 1
    rois[0, :, 1:] = np.stack([x1, y1, x1 + 30, y1 + 40], 1)
    scores = np.zeros((2, R), np.float32); scores[0] = np.linspace(1, 0.1, R)
    kps = np.zeros((2, 8, 3, 17), np.int32)
    kps[1, :2, 0] = 50; kps[1, :2, 1] = 50; kps[1, :2, 2] = 2
    gt = dict(boxes=torch.from_numpy(gb).cuda(), classes=torch.ones((2, 8), dtype=torch.int32).cuda(), crowd=torch.zeros((2, 8), dtype=torch.int32).cuda(),
              keypoints=torch.from_numpy(kps).cuda(), counts=torch.from_numpy(counts).cuda())
    o = target_ops.sample_rois(torch.from_numpy(rois).cuda(), torch.from_numpy(scores).cuda(), torch.tensor([R, 0], dtype=torch.int32).cuda(), gt,
                               torch.from_numpy(im).cuda(), _Cfg(32), 2)
    assert o['counts'].tolist() == [32, 2]                               # 32 background RoIs; the two gt boxes as foreground
    assert int((o['labels'][0, :32] != 0).sum()) == 0 and o['labels'][1, :2].tolist() == [1, 1] and int((o['labels'][1, 2:] != -1).sum()) == 0
    assert o['kp_counts'].tolist() == [0, 2]
    ref = ot.sample_rois(gb[1, :2], np.ones(2, np.int32), np.zeros(2, np.int32), kps[1, :2], np.zeros((0, 4), np.float32), 1.0, 1, 2, batch=32)
    _check_sampled(o, 1, ref, 32)


def test_rpn_tube_targets_equal_reference_goldens():
    """T = 3 tubes: tube IoU, per-frame box targets in the reference's fp64-promoted arithmetic (bit-exact: the fp64 log rounds
    to the same fp32), inside weights and vis labels from the track visibility."""
    import torch
    from detectandtrack_b200.ops import target_ops
    tag, T = 'rpnT3', 3
    gt, vis = G[tag + '_gt'], G[tag + '_vis']
    im_h, im_w, _ = G[tag + '_im']
    field = [int(f) for f in G[tag + '_field']]
    Gn = gt.shape[0]
    boxes = np.zeros((3, 8, 4 * T), np.float32); boxes[:, :Gn] = gt
    v = np.zeros((3, 8, T), np.uint8); v[:, :Gn] = vis
    anchors = [torch.from_numpy(np.ascontiguousarray(G['cell_anchors_T3_%d' % lvl], dtype=np.float64)).cuda() for lvl in range(2, 7)]
    out = target_ops.rpn_targets([(f, f) for f in field], anchors, [2. ** l for l in range(2, 7)], 3, torch.from_numpy(boxes).cuda(),
                                 torch.tensor([Gn] * 3, dtype=torch.int32).cuda(), torch.tensor([[im_h, im_w, 1.0]] * 3, dtype=torch.float32).cuda(),
                                 _TrainCfg(int(G[tag + '_batch'])), SEED, gt_visible=torch.from_numpy(v).cuda())
    for l, o in enumerate(out):
        assert np.array_equal(o['labels'][2].cpu().numpy(), G['%s_labels%d' % (tag, l)]), l      # the goldens were drawn for image 2
        assert np.array_equal(o['vis_labels'][2].cpu().numpy(), G['%s_vis%d' % (tag, l)]), l
        got, ref = o['bbox_targets'][2].cpu().numpy(), G['%s_bt%d' % (tag, l)]
        assert got.shape == ref.shape and _ulp_close(got, ref, 1), l
        assert np.array_equal(o['inside'][2].cpu().numpy(), G['%s_iw%d' % (tag, l)])
        assert np.array_equal(o['outside'][2].cpu().numpy(), G['%s_ow%d' % (tag, l)])


def test_sample_tube_rois_equal_reference_goldens():
    """T = 3: tube proposals merged with gt tubes (tube IoU), 4T box targets per class, per-frame heat-map labels."""
    import torch
    from detectandtrack_b200.ops import target_ops
    tag, T = 'roiT3', 3
    rois_in = G[tag + '_rois_in']
    P = rois_in.shape[0]
    batch = int(G[tag + '_batch'])
    gb, gk = G[tag + '_gt_boxes'], G[tag + '_gt_gt_keypoints']
    g = gb.shape[0]
    boxes = np.zeros((1, 8, 4 * T), np.float32); boxes[0, :g] = gb
    kps = np.zeros((1, 8, 3, 17 * T), np.int32); kps[0, :g] = gk
    t = lambda a: torch.from_numpy(a).cuda()
    gt = dict(boxes=t(boxes), classes=torch.ones((1, 8), dtype=torch.int32).cuda(), crowd=torch.zeros((1, 8), dtype=torch.int32).cuda(),
              keypoints=t(kps), counts=torch.tensor([g], dtype=torch.int32).cuda())
    scores = np.linspace(1.0, 0.1, P).astype(np.float32)[None]
    o = target_ops.sample_rois(t(rois_in[None].copy()), t(scores), torch.tensor([P], dtype=torch.int32).cuda(), gt,
                               torch.tensor([[0, 0, float(G[tag + '_scale'])]], dtype=torch.float32).cuda(), _Cfg(batch), SEED)
    n = len(G[tag + '_rois'])
    assert int(o['counts'][0]) == n and o['rois'].shape[2] == 13
    assert np.array_equal(o['rois'][0, :n].cpu().numpy(), G[tag + '_rois'])
    assert np.array_equal(o['labels'][0, :n].cpu().numpy(), G[tag + '_labels_int32'])
    assert _ulp_close(o['bbox_targets'][0, :n].cpu().numpy(), G[tag + '_bbox_targets'], 1)
    assert np.array_equal(o['inside'][0, :n].cpu().numpy(), G[tag + '_bbox_inside_weights'])
    assert np.array_equal(o['outside'][0, :n].cpu().numpy(), G[tag + '_bbox_outside_weights'])
    nk = len(G[tag + '_keypoint_rois'])
    assert int(o['kp_counts'][0]) == nk
    assert np.array_equal(o['kp_rois'][0, :nk].cpu().numpy(), G[tag + '_keypoint_rois'])
    assert np.array_equal(o['kp_locations'][0, :nk].cpu().numpy().reshape(-1, 1), G[tag + '_keypoint_locations_int32'])
    assert np.array_equal(o['kp_weights'][0, :nk].cpu().numpy().reshape(-1, 1), G[tag + '_keypoint_weights'])
