"""GPU: OUTPUT-LEVEL parity of the whole path in the headline arithmetic (bf16x3) against the oracle's full CPU
pipeline (oracle/pipeline.py: torch-fp32 graph + the reference's host ops) — no teacher forcing: pixels in, the final
cls_boxes / keypoints out.

Rows are compared as SETS: the reference orders RPN rois by score, and two fp32 scores that differ by less than the
arithmetic noise (1e-5) may legitimately swap places; a row is "matched" when some row of the other side lies within the
tolerance in every coordinate (max-norm relative to the image size).
Asserted (north star: bit-exact NMS / assignment indices, boxes and keypoint heat-map values within 1e-3 relative):
  * identical proposal count and identical detection count;
  * RPN rois: >= 99 % of the rows matched in both directions within 1e-3 (a roi exactly at a top-k / NMS decision boundary
    may flip on a 1e-5 score difference; with the seeds used here the match is 100 %, printed by the test);
  * detections: EVERY row matched within 1e-3, scores within 1e-3 absolute => the same per-class NMS keep set and the
    same DETECTIONS_PER_IM cut;
  * keypoint HEAT MAPS (kps_score, [n, 17, 56, 56]) of matched detections <= 1e-3 * max|ref|, and the decoded logit at
    the maximum (row 2 of the keypoint array) <= 1e-3 * max|ref|; probabilities <= 1e-3 absolute where the arg-max pixel
    agrees.  The arg-max POSITION itself is reported, not asserted: with seeded random weights the heat maps are nearly
    flat (the maximum exceeds thousands of other pixels by less than the 1e-4 arithmetic noise), so which of the
    near-equal pixels wins is ill-conditioned; arg-max equality on peaked maps, given identical inputs, is asserted in
    tests/test_gpu_dense_ops.py::test_keypoint_decode_vs_cv2_oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pipeline as opipe


def _match(a, b, tol):
    """For each row of a: index of the nearest row of b (max-abs distance) and whether it is within tol."""
    d = np.abs(a[:, None, :] - b[None, :, :]).max(-1)
    j = d.argmin(1)
    return j, d[np.arange(a.shape[0]), j] <= tol


_ORACLE = {}


def _compare(cfg, blobs, frames, mode, kp_tol=1e-3, key=None):
    import torch
    from detectandtrack_b200.modeling import params as P
    from detectandtrack_b200.modeling.engine import DetectionEngine
    ref = _ORACLE.get(key) if key else None           # the CPU oracle of a workload is computed once per test session
    if ref is None:
        ref = opipe.detect_clip(cfg, blobs, frames, want_heatmaps=True)
        ref.pop('feats', None)
        if key:
            _ORACLE[key] = ref
    eng = DetectionEngine(cfg, blobs, P.GraphSpec(cfg), dtype=mode)
    fr = torch.from_numpy(frames[None]).cuda()
    size = float(max(frames.shape[1:3]))
    # the RPN stage on its own
    g = eng._geom_tensors(1, frames.shape[1], frames.shape[2])
    feats = eng.link(eng.fpn(eng.body(eng._blob(fr, g['scale'], g['hr'], g['wr'], g['hp'], g['wp']))))
    rois, _, roi_counts = eng.rpn(feats, g['im_info'])
    n_roi = int(roi_counts[0])
    assert n_roi == ref['rois'].shape[0], ('proposal count', n_roi, ref['rois'].shape[0])
    got_rois = rois[0, :n_roi].cpu().numpy()
    _, ok_ab = _match(got_rois[:, 1:], ref['rois'][:, 1:], 1e-3 * size)
    _, ok_ba = _match(ref['rois'][:, 1:], got_rois[:, 1:], 1e-3 * size)
    assert ok_ab.mean() >= 0.99 and ok_ba.mean() >= 0.99, ('rpn rois matched', ok_ab.mean(), ok_ba.mean())
    res = eng.detect(fr, want_heatmaps=True)[0]
    b = res['boxes'].cpu().numpy()
    rb = ref['cls_boxes']
    assert b.shape == rb.shape, ('detection count', b.shape, rb.shape)
    j, ok = _match(b[:, :4], rb[:, :4], 1e-3 * size)
    assert ok.all() and len(set(j.tolist())) == len(j), ('detections matched', ok.mean(), len(set(j.tolist())), len(j))
    assert np.abs(b[:, 4] - rb[j, 4]).max() <= 1e-3, ('scores', np.abs(b[:, 4] - rb[j, 4]).max())
    k = res['keyps'].cpu().numpy()
    rk = ref['keyps'][j]
    assert k.shape == rk.shape
    heat = res['heatmaps'].cpu().numpy()
    rheat = ref['heat'][j]
    assert heat.shape == rheat.shape
    heat_err = np.abs(heat - rheat).max() / np.abs(rheat).max()
    assert heat_err <= kp_tol, ('keypoint heat maps', heat_err)
    assert np.abs(k[:, 2] - rk[:, 2]).max() <= kp_tol * np.abs(rk[:, 2]).max(), ('keypoint logits', np.abs(k[:, 2] - rk[:, 2]).max())
    same = (np.abs(k[:, 0] - rk[:, 0]) <= 1e-3) & (np.abs(k[:, 1] - rk[:, 1]) <= 1e-3)
    if same.any():
        assert np.abs(k[:, 3] - rk[:, 3])[same].max() <= 1e-3, 'keypoint probabilities'
    r = dict(n_roi=n_roi, n_det=b.shape[0], rois_matched=float(min(ok_ab.mean(), ok_ba.mean())),
             box_err=float(np.abs(b[:, :4] - rb[j, :4]).max() / size), score_err=float(np.abs(b[:, 4] - rb[j, 4]).max()),
             heat_err=float(heat_err), logit_err=float(np.abs(k[:, 2] - rk[:, 2]).max() / np.abs(rk[:, 2]).max()),
             argmax_same=float(same.mean()))
    print('e2e parity', mode, r)
    return r


@pytest.mark.parametrize('mode', ['bf16x3', 'tf32x3'])
def test_small_clip_outputs_match_oracle_pipeline(mode):
    from test_gpu_engine import _cfg
    from detectandtrack_b200.modeling import params as P
    cfg = _cfg()
    blobs, _ = P.random_blobs(cfg, seed=3)
    frames = np.random.RandomState(0).randint(0, 256, (3, 96, 128, 3)).astype(np.uint8)
    r = _compare(cfg, blobs, frames, mode, key='small')
    assert r['n_det'] > 0 and r['n_roi'] > 20


def test_fp16_posthoc_mode_is_not_a_parity_mode():
    """'bf16x3h' (bf16x3 with the four post-hoc FPN convs as ONE fp16 MMA per product: +20 % clips/s) is measured here so that
    its status is a test result, not a claim: on the 96x128 clip the heat maps agree to ~1e-3 (9.6e-4 measured: no margin)
    and at 800x1333 one of the 100 detections differs (round-2 gpurun).  It therefore stays a labelled extra in bench.py;
    this test only bounds it at 2e-3 so a regression of the fp16 path is still caught."""
    from test_gpu_engine import _cfg
    from detectandtrack_b200.modeling import params as P
    cfg = _cfg()
    blobs, _ = P.random_blobs(cfg, seed=3)
    frames = np.random.RandomState(0).randint(0, 256, (3, 96, 128, 3)).astype(np.uint8)
    r = _compare(cfg, blobs, frames, 'bf16x3h', kp_tol=2e-3, key='small')
    assert r['n_det'] > 0


@pytest.mark.parametrize('mode', ['bf16x3'])
def test_full_size_clip_outputs_match_oracle_pipeline(mode):
    """BASELINE.json configs[3] at its real size: ONE 800x1333 clip (R = 1000, D = 100), the headline mode, vs the oracle's
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
        r = _compare(cfg, blobs, frames, mode, key='full')
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
