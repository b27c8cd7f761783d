"""GPU parity (through the C ABI) of the RPN proposal pipeline and the box-head
post-processing against the reference goldens (tests/golden) and the oracle.
Bar: integer results (selection order, keep lists, levels, restore index, counts)
identical; box coordinates: tube path (fp64 decode) bit-exact, 2-D path (fp32 exp)
within 1e-3 relative (in practice a few ulp)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import boxes as ob
from oracle import proposals as op
from oracle import detections as od


def _cuda(a, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _run_level(g, name, anchors, T, post=300):
    import torch
    from detectandtrack_b200.ops import rpn_ops
    scores, deltas, im_info = g[name + '_scores'], g[name + '_deltas'], g[name + '_im_info']
    stride = float(g[name + '_stride'])
    # golden inputs are probabilities (the op receives sigmoid outputs); the device op takes logits
    p = scores.astype(np.float64)
    logits = np.log(p / (1 - p)).astype(np.float32)
    lg = _cuda(np.transpose(logits, (0, 2, 3, 1)))           # NCHW -> NHWC
    dl = _cuda(np.transpose(deltas, (0, 2, 3, 1)))
    props, keep, nkeep = rpn_ops.generate_proposals_level(
        lg.contiguous(), dl.contiguous(), _cuda(anchors), stride, _cuda(im_info), 1000, post, 0.7, 0.0, T)
    n = int(nkeep[0])
    k = keep[0, :n].cpu().numpy()
    return props[0].cpu().numpy(), k


@pytest.mark.parametrize('name,T', [('gp2d', 1), ('gp3d', 3)])
def test_generate_proposals_golden(golden, name, T):
    g = golden(name)
    ga = golden('anchors')
    anchors = ga['anchors_fpn5'] if name == 'gp2d' else ga['anchors_rpn12_T3']
    props, keep = _run_level(g, name, anchors, T)
    rois = g[name + '_rois']
    # The golden scores are a permutation of i/n, so sigmoid(logit(p)) can merge neighbours in fp32;
    # compare through the oracle run on the SAME logits to keep the test about the device op.
    p = g[name + '_scores'].astype(np.float64)
    logits = np.log(p / (1 - p)).astype(np.float32)
    probs = (1.0 / (1.0 + np.exp(-logits.astype(np.float32)))).astype(np.float32)
    o_props, o_sc, (pre_boxes, pre_sc), o_keep = op.generate_proposals(
        probs[0], g[name + '_deltas'][0], g[name + '_im_info'][0], anchors, float(g[name + '_stride']),
        1000, 300, 0.7, 0, return_intermediate=True)
    n_pre = pre_boxes.shape[0]
    dev_boxes, dev_sc = props[:n_pre, :-1], props[:n_pre, -1]
    if T == 1:
        np.testing.assert_allclose(dev_boxes, pre_boxes, rtol=1e-3, atol=1e-3)
    else:
        assert np.array_equal(dev_boxes, pre_boxes)           # fp64 decode rounded once: bit-exact
    np.testing.assert_allclose(dev_sc, pre_sc[:, 0], rtol=1e-6, atol=1e-7)
    # NMS keep list: identical to the oracle's NMS on the device-decoded boxes (bit-exact NMS),
    # and identical to the reference golden when the decode is bit-exact.
    ref_keep = ob.nms(np.hstack([dev_boxes, dev_sc[:, None]]).astype(np.float32), 0.7)[:300]
    assert keep.tolist() == [int(x) for x in ref_keep]
    if T == 3:
        assert np.array_equal(props[keep][:, :-1], rois[:, 1:]) or np.array_equal(keep, np.asarray(o_keep))


def test_rpn_topk_exact_selection_and_ties():
    """top-k selection is exact (radix select) incl. the documented tie rule and K > n."""
    import torch
    from detectandtrack_b200.ops import rpn_ops
    rng = np.random.default_rng(4)
    B, H, W, A = 2, 40, 56, 3
    logits = rng.normal(0, 2, (B, H, W, A)).astype(np.float32)
    logits = np.round(logits * 8) / 8                      # many exact ties
    deltas = np.zeros((B, H, W, 4 * A), np.float32)
    anchors = op.generate_anchors(16, (64,), (0.5, 1, 2))
    im_info = np.array([[H * 16, W * 16, 1.0]] * B, np.float32)
    for K in (500, 6000, 0):
        props, counts = rpn_ops.rpn_proposals(_cuda(logits), _cuda(deltas), _cuda(anchors), 16.0, _cuda(im_info), K)
        for b in range(B):
            sc = (1.0 / (1.0 + np.exp(-logits[b].reshape(-1)))).astype(np.float32)
            order = np.argsort(-sc, kind='stable')
            kk = len(sc) if K == 0 else min(K, len(sc))
            got = props[b, :int(counts[b]), -1].cpu().numpy()
            assert int(counts[b]) == kk
            np.testing.assert_allclose(got, sc[order[:kk]], rtol=1e-6)
            # decoded box of rank r is the anchor of index order[r] (zero deltas): checks the tie order
            allanch = op.shifted_anchors(anchors, H, W, 16.0, 1).astype(np.float32)
            exp_boxes = ob.clip_tiled_boxes(ob.bbox_transform(allanch[order[:kk]], np.zeros((kk, 4), np.float32)), im_info[b, :2])
            np.testing.assert_allclose(props[b, :kk, :4].cpu().numpy(), exp_boxes, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize('name,T', [('cd2d', 1), ('cd3d', 3)])
def test_collect_distribute_golden(golden, name, T):
    import torch
    from detectandtrack_b200.ops import rpn_ops
    g = golden(name)
    rois_l = [g['%s_in_rois%d' % (name, i)] for i in range(5)]
    sc_l = [g['%s_in_scores%d' % (name, i)] for i in range(5)]
    K = max(r.shape[0] for r in rois_l)
    ld = 4 * T + 1
    props = np.zeros((1, 5, K, ld), np.float32)
    keep = np.zeros((5, K), np.int32)
    nkeep = np.zeros((5,), np.int32)
    for l in range(5):
        n = rois_l[l].shape[0]
        perm = np.random.default_rng(l).permutation(n)          # keep lists need not be the identity
        props[0, l, perm, :-1] = rois_l[l][:, 1:]
        props[0, l, perm, -1] = sc_l[l][:, 0]
        keep[l, :n] = perm
        nkeep[l] = n
    rois, scores, counts = rpn_ops.collect(_cuda(props), _cuda(keep), _cuda(nkeep), 1000)
    n = int(counts[0])
    assert n == g[name + '_rois'].shape[0]
    assert np.array_equal(rois[0, :n].cpu().numpy(), g[name + '_rois'])
    levels, restore, lc = rpn_ops.distribute(rois[0, :n], None, col0=1, T=T)
    assert np.array_equal(levels.cpu().numpy().astype(np.float32), g[name + '_lvls'])
    assert np.array_equal(restore.cpu().numpy(), g[name + '_idx_restore'])
    assert lc.cpu().numpy().tolist() == [g['%s_rois_fpn%d' % (name, l)].shape[0] for l in range(2, 6)]


@pytest.mark.parametrize('T', [1, 3])
def test_box_decode_nms_limit_vs_oracle(T):
    import torch
    from detectandtrack_b200.ops import rpn_ops, box_ops
    rng = np.random.default_rng(10 + T)
    B, R, C = 2, 600, 2
    im_scale = 1.6
    im_hw = np.array([[500, 833], [480, 640]], np.float32)
    rois = np.zeros((B, R, 4 * T + 1), np.float32)
    counts = np.array([600, 433], np.int32)
    for b in range(B):
        ctr = rng.uniform(50, 400, (40, 2))
        c = ctr[rng.integers(0, 40, R)] + rng.normal(0, 8, (R, 2))
        wh = rng.uniform(30, 150, (R, 2))
        box = np.hstack([c, c + wh]) * im_scale
        rois[b, :, 0] = b
        for t in range(T):
            rois[b, :, 1 + 4 * t:5 + 4 * t] = box + rng.normal(0, 3, (R, 1)) * (t > 0)
    logits = rng.normal(0, 2.5, (B * R, C)).astype(np.float32)
    deltas = rng.normal(0, 0.5, (B * R, 4 * T * C)).astype(np.float32)
    im_info = np.array([[800, 1344, im_scale]] * B, np.float32)
    dets, cnt = rpn_ops.box_decode(_cuda(rois), _cuda(counts), _cuda(logits), _cuda(deltas), C, _cuda(im_info),
                                   _cuda(im_hw), (10., 10., 5., 5.), 0.05, T)
    cmp_mode = box_ops.NMS_2D_GE if T == 1 else box_ops.NMS_TUBE_GT
    order = box_ops.ORDER_INDEX if T == 1 else box_ops.ORDER_SCORE
    keep, nkeep = box_ops.nms_batched(dets.view(B * (C - 1), R, 4 * T + 1), cnt, 0.5, cmp_mode, order)
    out, ocnt = rpn_ops.limit_detections(dets, keep, nkeep, 100)
    for b in range(B):
        n = counts[b]
        sc = od.softmax(logits[b * R:b * R + n])
        pred = od.decode_boxes(rois[b, :n], deltas[b * R:b * R + n], im_scale, im_hw[b])
        # device-side pre-NMS set == oracle's score filter; boxes within tolerance (bit-exact for tubes)
        inds = np.where(sc[:, 1] > np.float32(0.05))[0]
        dn = int(cnt[b])
        d_dev = dets[b, 0, :dn].cpu().numpy()
        assert dn == len(inds)
        if T == 1:
            np.testing.assert_allclose(d_dev[:, :-1], pred[inds, 4:8], rtol=1e-3, atol=1e-3)
        else:
            assert np.array_equal(d_dev[:, :-1], pred[inds, 4 * T:8 * T])
        np.testing.assert_allclose(d_dev[:, -1], sc[inds, 1], rtol=1e-5, atol=1e-7)
        # NMS + limit on the device-decoded rows == the oracle's box_results_with_nms_and_limit on them
        scores2 = np.zeros((dn, 2), np.float32); scores2[:, 1] = d_dev[:, -1]
        boxes2 = np.zeros((dn, 8 * T), np.float32); boxes2[:, 4 * T:] = d_dev[:, :-1]
        _, _, cls_boxes = od.box_results_with_nms_and_limit(scores2, boxes2, 2, 0.05, 0.5, 100)
        m = int(ocnt[b])
        assert m == cls_boxes[1].shape[0]
        assert np.array_equal(out[b, 0, :m].cpu().numpy(), cls_boxes[1])
