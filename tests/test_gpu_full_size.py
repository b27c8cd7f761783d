"""GPU: size-independent properties at BASELINE.json's full sizes (800x1333 clips, T=3) where the oracle would
take minutes: impulse response / shift equivariance / batch independence of the big convs, sortedness and the
exact selected set of the P2-level RPN top-k (604 800 anchors), and the whole detection step (bounds, counts,
determinism, batch consistency)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _conv(x, wp, k, pad, **kw):
    from detectandtrack_b200.ops import conv as cv
    return cv.conv3d(x, wp, k, (1, 1, 1), pad, relu=False, out_f32=False, dtype=cv.BF16, **kw)


def test_conv_impulse_response_full_size():
    """A unit impulse at (t, h, w, c) must reproduce the filter slice w[:, c, ::-1, ::-1, ::-1] around it and
    zeros elsewhere (exact: products with 1.0), at the FPN post-hoc shape 3x200x336x256, incl. clip corners."""
    import torch
    from detectandtrack_b200.ops import conv as cv
    g = torch.Generator().manual_seed(1)
    C = 256
    w = (torch.randn((C, C, 3, 3, 3), generator=g) * 0.05).bfloat16()
    wp = cv.pack_weight(w.float(), cv.BF16)
    T, H, W = 3, 200, 336
    for (t, h, wq, c) in [(0, 0, 0, 0), (1, 100, 167, 77), (2, 199, 335, 255), (1, 7, 15, 3), (0, 8, 16, 200), (2, 127, 128, 129)]:
        x = torch.zeros((1, T, H, W, C), dtype=torch.bfloat16, device='cuda')
        x[0, t, h, wq, c] = 1.0
        y = _conv(x, wp, (3, 3, 3), (1, 1, 1)).float().cpu()
        exp = torch.zeros_like(y)
        for kt in range(3):
            for kh in range(3):
                for kw in range(3):
                    ot, oh, ow = t - kt + 1, h - kh + 1, wq - kw + 1          # output position that sees the impulse at tap (kt,kh,kw)
                    if 0 <= ot < T and 0 <= oh < H and 0 <= ow < W:
                        exp[0, ot, oh, ow] = w[:, c, kt, kh, kw].float()
        assert torch.equal(y, exp), (t, h, wq, c)


def test_conv_shift_and_batch_invariance_full_size():
    """Bit-exact: (1) clip b of a batched launch equals the same clip alone (tiles, stacking and scheduling
    differ); (2) shifting the input by one tile-misaligned offset shifts the interior of the output."""
    import torch
    from detectandtrack_b200.ops import conv as cv
    g = torch.Generator().manual_seed(2)
    w = (torch.randn((128, 128, 3, 3, 3), generator=g) * 0.05).bfloat16()
    wp = cv.pack_weight(w.float(), cv.BF16)
    x = torch.randn((2, 3, 100, 168, 128), generator=g).bfloat16().cuda()
    yb = _conv(x, wp, (3, 3, 3), (1, 1, 1))
    for b in range(2):
        assert torch.equal(yb[b:b + 1], _conv(x[b:b + 1].contiguous(), wp, (3, 3, 3), (1, 1, 1)))
    dh, dw = 5, 3
    xs = torch.zeros_like(x)
    xs[:, :, dh:, dw:] = x[:, :, :-dh, :-dw]
    ys = _conv(xs, wp, (3, 3, 3), (1, 1, 1))
    # rows / columns whose 3x3 window touches the zero padding of either tensor are excluded
    assert torch.equal(ys[:, :, dh + 1:-1, dw + 1:-1], yb[:, :, 1:-dh - 1, 1:-dw - 1])


def test_rpn_topk_full_size_properties():
    """P2-sized level (200x336x3 = 201 600 positions x 3 anchors per image, 2 images): the K=1000 selected
    scores are sorted, equal torch.topk's values exactly, and no unselected score beats the last one."""
    import torch
    from detectandtrack_b200.ops import rpn_ops
    from detectandtrack_b200.modeling.generate_anchors import generate_anchors
    rng = np.random.default_rng(7)
    B, H, W, A, K = 2, 200, 336, 3, 1000
    logits = torch.from_numpy(rng.normal(-3, 2, (B, H, W, A)).astype(np.float32)).cuda()
    deltas = torch.zeros((B, H, W, 4 * A), dtype=torch.float32, device='cuda')
    anchors = torch.from_numpy(np.ascontiguousarray(generate_anchors(4, (32,), (0.5, 1, 2)), dtype=np.float64)).cuda()
    im_info = torch.tensor([[800.0, 1344.0, 1.0]] * B, device='cuda')        # the padded blob: nothing is filtered
    props, counts = rpn_ops.rpn_proposals(logits, deltas, anchors, 4.0, im_info, K)
    sc_all = torch.sigmoid(logits.double()).float().view(B, -1)              # monotone in the logit: same ranking
    for b in range(B):
        n = int(counts[b])
        assert n == K                                                         # zero deltas, min_size 0: nothing filtered
        got = props[b, :n, -1]
        assert torch.all(got[:-1] >= got[1:])
        ref_vals, ref_idx = torch.topk(logits[b].view(-1), K)
        # the kernel's sigmoid is Caffe2's 1/(1+exp(-x)) in fp32: compare through the logits it selected
        kth = got[-1].item()
        assert (sc_all[b] > kth + 1e-6).sum().item() <= K
        np.testing.assert_allclose(got.cpu().numpy(), torch.sigmoid(ref_vals.double()).float().cpu().numpy(), rtol=2e-6)
        # boxes inside the image after clipping
        bx = props[b, :n, :4]
        assert bx[:, 0].min() >= 0 and bx[:, 1].min() >= 0 and bx[:, 2].max() <= 1343 and bx[:, 3].max() <= 799


def test_detect_step_full_size_properties():
    """The benchmarked step at 800x1333, T=3 (R50-FPN-3D, R=1000, D<=100): counts within limits, boxes inside the
    image, keypoints finite, bit-identical when repeated, and identical clips in one batch give identical results."""
    import torch
    import bench
    from detectandtrack_b200.modeling import params as P
    from detectandtrack_b200.modeling.engine import DetectionEngine
    from detectandtrack_b200.core.config import reset_cfg
    cfg = bench.bench_cfg(800, 1333)
    blobs, spec = P.random_blobs(cfg)
    eng = DetectionEngine(cfg, blobs, spec, dtype='bf16')
    fr = torch.from_numpy(bench.synth_frames(1, 3, 800, 1333, 7))
    two = torch.cat([fr, fr], 0).cuda()
    a = eng.detect_static(two)
    b = eng.detect_static(two)
    torch.cuda.synchronize()
    for k in ('dets', 'det_counts', 'xy'):
        assert torch.equal(a[k], b[k]), k                                    # deterministic
    D = cfg.TEST.DETECTIONS_PER_IM
    cnt = a['det_counts'].view(2, -1)[:, 0]
    assert int(cnt.min()) >= 1 and int(cnt.max()) <= D
    assert int(cnt[0]) == int(cnt[1])
    n = int(cnt[0])
    dets = a['dets'].view(2, -1, a['dets'].shape[-1])
    assert torch.equal(dets[0, :n], dets[1, :n])                             # same clip twice in the batch
    bx = dets[0, :n, :4]
    assert bx.min().item() >= 0 and bx[:, 2].max().item() <= 1332.0 and bx[:, 3].max().item() <= 799.0
    assert dets[0, :n, 4].min().item() >= cfg.TEST.SCORE_THRESH                # (rows keep the reference's NMS order, not score order)
    cap = dets.shape[1]
    assert torch.equal(a['xy'][:n], a['xy'][cap:cap + n])
    assert torch.isfinite(a['xy'][:n]).all()
    reset_cfg()                                                               # the global cfg is shared between test modules


def test_training_step_full_size_properties():
    """BASELINE.json configs[4] at its full size (2 clips of 3 x 800 x 1333, 512 RoIs per image, 2000 proposals per level):
    the oracle graph would take many minutes on the CPU, so size-independent properties are checked — the target generators
    are deterministic in (seed, image) and independent of the other images of the batch, every sampled RoI / anchor count
    obeys the reference's quotas, losses are finite and the wgrad of a linear layer is linear in its upstream gradient."""
    import torch
    import bench
    from detectandtrack_b200.modeling import params as P
    from detectandtrack_b200.modeling.trainer import KeypointRcnnTrainer, pack_gt
    from detectandtrack_b200.ops import train_ops as to
    cfg = bench.bench_cfg(800, 1333, 'r50fpn3d')
    cfg.TRAIN.BATCH_SIZE_PER_IM = 512; cfg.TRAIN.RPN_PRE_NMS_TOP_N = 2000
    try:
        blobs, spec = P.random_blobs(cfg)
        tr = KeypointRcnnTrainer(cfg, blobs, spec)
        frames = torch.from_numpy(bench.synth_frames(2, 3, 800, 1333, 5)).cuda()
        entries = bench.synth_gt(2, 800, 1333, 11)
        gt = pack_gt(entries)
        outs = tr.forward_all(frames)
        rt, smp = tr.make_targets(outs, gt, 2, 800, 1333, seed=9)
        rt2, smp2 = tr.make_targets(outs, gt, 2, 800, 1333, seed=9)
        for a, b in zip(rt, rt2):
            for k in a:
                assert torch.equal(a[k], b[k]), k                         # same (seed, image) -> same draws
        for k in ('rois', 'labels', 'bbox_targets', 'kp_rois', 'kp_locations', 'kp_weights'):
            assert torch.equal(smp[k], smp2[k]), k
        # RPN quotas (rpn.py:266-283): <= 128 foreground, <= 256 labelled anchors per image, outside weight = 1 / #labelled
        for b in range(2):
            nfg = sum(int((t['labels'][b] == 1).sum()) for t in rt)
            nlab = sum(int((t['labels'][b] >= 0).sum()) for t in rt)
            assert 0 < nfg <= 128 and nfg < nlab <= 256
            ow = torch.cat([t['outside'][b].reshape(-1) for t in rt])
            assert torch.equal(torch.unique(ow[ow > 0]), torch.tensor([np.float32(1.0 / nlab)], device='cuda'))
        # image 1's RPN targets do not depend on image 0 (swap the batch order, same per-image seeds are used by position)
        gt_one = pack_gt([entries[1], entries[1]])
        rt3, _ = tr.make_targets(outs, gt_one, 2, 800, 1333, seed=9)
        for a, c in zip(rt, rt3):
            assert torch.equal(a['labels'][1], c['labels'][1]) and torch.equal(a['bbox_targets'][1], c['bbox_targets'][1])
        # RoI quotas (fast_rcnn.py:118-147): 512 RoIs per image, <= 128 foreground first, labels in {0, 1}
        assert smp['counts'].tolist() == [512, 512] and float(smp['totals'][0]) == 1024.0
        lab = smp['labels']
        for b in range(2):
            nf = int((lab[b] == 1).sum())
            assert 0 < nf <= 128 and int((lab[b, :nf] == 1).sum()) == nf and int((lab[b, nf:] == 0).sum()) == 512 - nf
            assert float(smp['inside'][b, nf:].abs().sum()) == 0.0
        assert 0 < int(smp['kp_counts'].min()) and int(smp['kp_counts'].max()) <= 128
        loc = smp['kp_locations']
        assert int(loc.min()) >= 0 and int(loc.max()) < 56 * 56
        l1, l2 = tr.step(frames, gt, seed=9)
        torch.cuda.synchronize()
        assert torch.isfinite(l1).all() and torch.isfinite(l2).all() and float(l2[2]) > 0
        # linearity of the filter gradient in the upstream gradient at a full-size layer (P3 post-hoc conv shape)
        g = torch.Generator().manual_seed(3)
        x = torch.randn((2, 3, 100, 168, 256), generator=g).bfloat16().cuda()
        g1 = (torch.randint(-4, 5, (2, 3, 100, 168, 256), generator=g).float() / 8).bfloat16().cuda()
        g2 = (torch.randint(-4, 5, (2, 3, 100, 168, 256), generator=g).float() / 8).bfloat16().cuda()
        d1, d2 = to.wgrad_nhwc(g1, x, (3, 3, 3)), to.wgrad_nhwc(g2, x, (3, 3, 3))
        d12 = to.wgrad_nhwc((g1.float() + g2.float()).bfloat16(), x, (3, 3, 3))      # sums of eighths stay exact in bf16
        ref = d1 + d2
        assert float((d12 - ref).abs().max()) <= 2e-3 * float(ref.abs().max())
    finally:
        from test_gpu_engine import _cfg
        _cfg()
