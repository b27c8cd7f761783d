"""GPU: OUTPUT-LEVEL parity of the whole path in the headline arithmetic (bf16x3) against the oracle's full CPU
pipeline (oracle/pipeline.py: torch-fp32 graph + the reference's host ops) — no teacher forcing: pixels in, the final
cls_boxes / keypoints out.

Asserted (north star: bit-exact NMS / assignment indices, boxes and keypoint heat-map values within 1e-3 relative):
  * identical proposal count and identical detection count (=> the same RPN / per-class NMS keep sets and the same
    DETECTIONS_PER_IM cut), rows in the same order;
  * RPN rois, detection boxes, scores:            max|a - b| <= 1e-3 * max|b|   (max-norm relative)
  * keypoint logits (row 2) and probabilities (row 3): <= 1e-3 * max|ref| resp. 1e-3 absolute;
  * keypoint positions: identical arg-max pixel (|dx|,|dy| <= 1e-3 px) for >= 99 % of the keypoints — the arg-max of a
    bicubic-resized map may legitimately flip between two pixels whose values differ by less than the arithmetic noise.
Seeds were chosen once; a seed with a borderline IoU / score tie would show up as a count mismatch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pipeline as opipe


def _compare(cfg, blobs, frames, mode, kp_tol=1e-3):
    import torch
    from detectandtrack_b200.modeling import params as P
    from detectandtrack_b200.modeling.engine import DetectionEngine
    ref = opipe.detect_clip(cfg, blobs, frames)
    eng = DetectionEngine(cfg, blobs, P.GraphSpec(cfg), dtype=mode)
    fr = torch.from_numpy(frames[None]).cuda()
    # the RPN stage on its own (identical keep sets <=> identical roi rows)
    g = eng._geom_tensors(1, frames.shape[1], frames.shape[2])
    feats = eng.link(eng.fpn(eng.body(eng._blob(fr, g['scale'], g['hr'], g['wr'], g['hp'], g['wp']))))
    rois, _, roi_counts = eng.rpn(feats, g['im_info'])
    n_roi = int(roi_counts[0])
    assert n_roi == ref['rois'].shape[0], ('proposal count', n_roi, ref['rois'].shape[0])
    got_rois = rois[0, :n_roi].cpu().numpy()
    assert np.abs(got_rois - ref['rois']).max() <= 1e-3 * np.abs(ref['rois']).max(), 'rpn rois'
    res = eng.detect(fr)[0]
    b = res['boxes'].cpu().numpy()
    rb = ref['cls_boxes']
    assert b.shape == rb.shape, ('detection count', b.shape, rb.shape)
    assert np.abs(b[:, :4] - rb[:, :4]).max() <= 1e-3 * np.abs(rb[:, :4]).max(), ('boxes', np.abs(b[:, :4] - rb[:, :4]).max())
    assert np.abs(b[:, 4] - rb[:, 4]).max() <= 1e-3, ('scores', np.abs(b[:, 4] - rb[:, 4]).max())
    k = res['keyps'].cpu().numpy()
    rk = ref['keyps']
    assert k.shape == rk.shape
    assert np.abs(k[:, 2] - rk[:, 2]).max() <= kp_tol * np.abs(rk[:, 2]).max(), ('keypoint logits', np.abs(k[:, 2] - rk[:, 2]).max())
    same = (np.abs(k[:, 0] - rk[:, 0]) <= 1e-3) & (np.abs(k[:, 1] - rk[:, 1]) <= 1e-3)
    assert same.mean() >= 0.99, ('keypoint arg-max positions', same.mean())
    assert np.abs(k[:, 3] - rk[:, 3])[same].max() <= 1e-3, 'keypoint probabilities'
    return dict(n_roi=n_roi, n_det=b.shape[0], box_err=float(np.abs(b[:, :4] - rb[:, :4]).max() / np.abs(rb[:, :4]).max()),
                logit_err=float(np.abs(k[:, 2] - rk[:, 2]).max() / np.abs(rk[:, 2]).max()), same=float(same.mean()))


@pytest.mark.parametrize('mode', ['bf16x3', 'tf32x3'])
def test_small_clip_outputs_match_oracle_pipeline(mode):
    from test_gpu_engine import _cfg
    from detectandtrack_b200.modeling import params as P
    cfg = _cfg()
    blobs, _ = P.random_blobs(cfg, seed=3)
    frames = np.random.RandomState(0).randint(0, 256, (3, 96, 128, 3)).astype(np.uint8)
    r = _compare(cfg, blobs, frames, mode)
    assert r['n_det'] > 0 and r['n_roi'] > 20


def test_full_size_clip_outputs_match_oracle_pipeline():
    """BASELINE.json configs[3] at its real size: ONE 800x1333 clip (R = 1000, D = 100), headline mode, vs the oracle's
    values (the CPU side takes a minute on the GPU box)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from detectandtrack_b200.modeling import params as P
    from detectandtrack_b200.core.config import reset_cfg
    import torch
    cfg = bench.bench_cfg(800, 1333, 'r50fpn3d')
    try:
        torch.set_num_threads(min(os.cpu_count() or 8, 32))
        blobs, _ = P.random_blobs(cfg)
        frames = bench.synth_frames(1, 3, 800, 1333, 7)[0]
        r = _compare(cfg, blobs, frames, 'bf16x3')
        assert r['n_roi'] == 1000 and r['n_det'] == 100
    finally:
        reset_cfg()


def test_2d_r50fpn_outputs_match_oracle_pipeline():
    """BASELINE.json configs[1]: 2-D R50-FPN keypoint R-CNN (MODEL.VIDEO_ON False; lib/modeling/ResNet.py:230-263,
    FPN.py:114-184), one 256x320 frame, output-level parity."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from detectandtrack_b200.modeling import params as P
    from detectandtrack_b200.core.config import reset_cfg
    cfg = bench.bench_cfg(256, 320, 'r50fpn2d')
    try:
        cfg.TEST.RPN_POST_NMS_TOP_N = 300
        blobs, _ = P.random_blobs(cfg, seed=11)
        assert blobs['conv1_w'].ndim == 4 and blobs['res3_0_branch2b_w'].shape == (128, 128, 3, 3)      # 2-D filters
        frames = np.random.RandomState(2).randint(0, 256, (1, 256, 320, 3)).astype(np.uint8)
        r = _compare(cfg, blobs, frames, 'bf16x3')
        assert r['n_det'] > 0
    finally:
        reset_cfg()
