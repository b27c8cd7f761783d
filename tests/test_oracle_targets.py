"""CPU: oracle/targets.py (the checker of the device target generators) against outputs of the REFERENCE's own
lib/roi_data + lib/datasets functions (tests/golden/targets.npz, tests/golden/gen_golden_targets.py)."""
import os

import numpy as np
import pytest

from oracle import targets as ot

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'targets.npz'))
SEED = int(G['seed'])


def _levels(tag):
    return [(G['cell_anchors%d' % lvl], 2. ** lvl, int(G[tag + '_field'][l]), int(G[tag + '_field'][l])) for l, lvl in enumerate(range(2, 7))]


def test_hash_is_a_fixed_function():
    # known-answer values of the shared counter-based generator (also restated in csrc/targets.cu)
    h = ot.hash_u32(3, 1, 2, np.arange(4))
    assert h.dtype == np.uint32 and len(set(h.tolist())) == 4
    assert np.array_equal(h, ot.hash_u32(3, 1, 2, np.arange(4)))
    assert not np.array_equal(h, ot.hash_u32(3, 2, 2, np.arange(4)))
    a = np.arange(100)
    c = ot.choice(a, 10, 3, 0, 0)
    assert len(set(c.tolist())) == 10 and np.array_equal(c, ot.choice(a[::-1].copy(), 10, 3, 0, 0))
    r = ot.randint(7, 1000, 3, 1, 0)
    assert r.min() >= 0 and r.max() == 6 and abs(np.mean(r) - 3.0) < 0.3


@pytest.mark.parametrize('tag', ['rpnA', 'rpnB', 'rpnC'])
def test_rpn_targets_equal_reference(tag):
    im_h, im_w, _ = G[tag + '_im']
    out, diag = ot.rpn_targets(_levels(tag), G[tag + '_gt'], float(im_h), float(im_w), SEED, 1, batch=int(G[tag + '_batch']))
    nfg = nbg = 0
    for l, o in enumerate(out):
        assert np.array_equal(o['labels'], G['%s_labels%d' % (tag, l)])
        assert np.array_equal(o['bbox_targets'], G['%s_bt%d' % (tag, l)])
        assert np.array_equal(o['inside'], G['%s_iw%d' % (tag, l)])
        assert np.array_equal(o['outside'], G['%s_ow%d' % (tag, l)])
        nfg += int((o['labels'] == 1).sum()); nbg += int((o['labels'] == 0).sum())
    assert nfg > 0 and nbg > 0 and nfg + nbg <= int(G[tag + '_batch'])


@pytest.mark.parametrize('tag', ['roiA', 'roiB', 'roiC'])
def test_sample_rois_equal_reference(tag):
    rois_in = G[tag + '_rois_in']
    r = ot.sample_rois(G[tag + '_gt_boxes'], G[tag + '_gt_gt_classes'], G[tag + '_gt_is_crowd'], G[tag + '_gt_gt_keypoints'],
                       rois_in[:, 1:], G[tag + '_scale'], 0, SEED, batch=int(G[tag + '_batch']))
    assert np.array_equal(r['diag']['max_overlaps'], G[tag + '_max_overlaps'])
    assert np.array_equal(r['diag']['max_classes'], G[tag + '_max_classes'])
    assert np.array_equal(r['diag']['box_to_gt'], G[tag + '_box_to_gt'])
    for k, gk in (('rois', 'rois'), ('labels', 'labels_int32'), ('bbox_targets', 'bbox_targets'), ('inside', 'bbox_inside_weights'),
                  ('outside', 'bbox_outside_weights'), ('keypoint_rois', 'keypoint_rois')):
        assert np.array_equal(r[k], G['%s_%s' % (tag, gk)]), k
    K = 17
    assert np.array_equal(r['keypoint_locations'].reshape(-1, 1), G[tag + '_keypoint_locations_int32'])
    assert np.array_equal(r['keypoint_weights'].reshape(-1, 1), G[tag + '_keypoint_weights'])
    assert r['keypoint_weights'].shape[1] == K and r['keypoint_weights'].sum() > 0


def test_heatmap_labels_equal_reference():
    heat, w = ot.keypoints_to_heatmap_labels(G['heat_kps'], G['heat_rois'], 56)
    assert np.array_equal(heat, G['heat_loc']) and np.array_equal(w, G['heat_w'])
    assert 0 < w.sum() < w.size


def test_collect_train_is_batch_wide():
    rng = np.random.default_rng(0)
    sc = [np.sort(rng.uniform(0, 1, n).astype(np.float32))[::-1] for n in (50, 30)]
    rois = [np.arange(len(s))[:, None] + 100 * b for b, s in enumerate(sc)]
    kept = ot.collect_train(rois, sc, 40)
    assert sum(len(k) for k in kept) == 40
    thr = np.sort(np.concatenate(sc))[::-1][39]
    for b in range(2):
        assert np.array_equal(kept[b][:, 0] - 100 * b, np.where(sc[b] >= thr)[0])


def test_rpn_tube_targets_equal_reference():
    """T = 3: tube anchors, tube IoU (mean over frames), per-frame targets, inside weights and vis labels from track_visible."""
    tag = 'rpnT3'
    im_h, im_w, _ = G[tag + '_im']
    lv = [(G['cell_anchors_T3_%d' % lvl], 2. ** lvl, int(G[tag + '_field'][l]), int(G[tag + '_field'][l])) for l, lvl in enumerate(range(2, 7))]
    out, _ = ot.rpn_targets(lv, G[tag + '_gt'], float(im_h), float(im_w), SEED, 2, batch=int(G[tag + '_batch']), visible=G[tag + '_vis'])
    some_invisible = 0
    for l, o in enumerate(out):
        assert np.array_equal(o['labels'], G['%s_labels%d' % (tag, l)])
        assert np.array_equal(o['vis_labels'], G['%s_vis%d' % (tag, l)])
        assert np.array_equal(o['bbox_targets'], G['%s_bt%d' % (tag, l)]) and o['bbox_targets'].shape[-1] == 3 * 4 * 3
        assert np.array_equal(o['inside'], G['%s_iw%d' % (tag, l)]) and np.array_equal(o['outside'], G['%s_ow%d' % (tag, l)])
        fg = o['labels'] == 1
        some_invisible += int((o['vis_labels'].reshape(o['labels'].shape + (3,))[fg] == 0).sum())
    assert sum(int((o['labels'] == 1).sum()) for o in out) > 0 and some_invisible > 0


def test_sample_tube_rois_equal_reference():
    tag = 'roiT3'
    rois_in = G[tag + '_rois_in']
    r = ot.sample_rois(G[tag + '_gt_boxes'], G[tag + '_gt_gt_classes'], G[tag + '_gt_is_crowd'], G[tag + '_gt_gt_keypoints'],
                       rois_in[:, 1:], G[tag + '_scale'], 0, SEED, batch=int(G[tag + '_batch']))
    for k, gk in (('rois', 'rois'), ('labels', 'labels_int32'), ('bbox_targets', 'bbox_targets'), ('inside', 'bbox_inside_weights'),
                  ('outside', 'bbox_outside_weights'), ('keypoint_rois', 'keypoint_rois')):
        assert np.array_equal(r[k], G['%s_%s' % (tag, gk)]), k
    assert r['rois'].shape[1] == 13 and r['bbox_targets'].shape[1] == 24
    assert np.array_equal(r['keypoint_locations'].reshape(-1, 1), G[tag + '_keypoint_locations_int32'])
    assert np.array_equal(r['keypoint_weights'].reshape(-1, 1), G[tag + '_keypoint_weights'])
    assert r['keypoint_weights'].shape[1] == 51 and (r['labels'] == 1).sum() > 0
