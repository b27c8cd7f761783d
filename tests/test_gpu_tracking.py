"""GPU parity: batched assignment and track-id propagation vs scipy / the oracle.
Bar: identical indices (bit-exact), on tie-dominated costs too."""
import numpy as np
import pytest
import scipy.optimize

pytestmark = pytest.mark.gpu

from oracle import tracking as ot
from oracle.lsa import lsap_crouse


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _matches_from(r, c, Q):
    m = -np.ones(Q, np.int32)
    m[c] = r
    return m


def test_lsa_golden(golden):
    from detectandtrack_b200.core import tracking_engine as te
    g = golden('lsa')
    for k in [k[len('lsa_C'):] for k in g if k.startswith('lsa_C')]:
        C = g['lsa_C' + k]
        m = te._compute_matches(None, None, None, None, None, None, None, None, 'hungarian', C=C)
        assert np.array_equal(m, _matches_from(g['lsa_r' + k], g['lsa_c' + k], C.shape[1])), k


def test_lsa_batched_random_vs_scipy():
    import torch
    from detectandtrack_b200.ops import box_ops
    rng = np.random.default_rng(21)
    B, D = 300, 48
    cost = np.zeros((B, D, D), np.float32)
    nr = rng.integers(0, D + 1, B).astype(np.int32)
    nc = rng.integers(0, D + 1, B).astype(np.int32)
    for b in range(B):
        P, Q = nr[b], nc[b]
        kind = b % 4
        if kind == 0:
            C = rng.random((P, Q))
        elif kind == 1:
            C = rng.integers(0, 3, (P, Q))
        elif kind == 2:
            C = np.ones((P, Q)); m = rng.random((P, Q)) < 0.1; C[m] = rng.random(int(m.sum()))
        else:
            C = np.ones((P, Q))
        cost[b, :P, :Q] = C
    m, st = box_ops.lsa_batched(_cuda(cost), torch.from_numpy(nr), torch.from_numpy(nc))
    m, st = m.cpu().numpy(), st.cpu().numpy()
    assert not st.any()
    for b in range(B):
        P, Q = nr[b], nc[b]
        r, c = scipy.optimize.linear_sum_assignment(cost[b, :P, :Q])
        assert np.array_equal(m[b, :Q], _matches_from(r, c, Q)), (b, P, Q)
        assert np.all(m[b, Q:] == -1)


@pytest.mark.parametrize('T', [1, 3])
def test_match_frames_and_tracks_vs_oracle(T):
    from detectandtrack_b200.core import tracking_engine as te
    from detectandtrack_b200.core.config import cfg, reset_cfg
    reset_cfg()
    rng = np.random.default_rng(3)
    videos = [ot.synth_video(rng, n_frames=12, n_dets=100, T=T), ot.synth_video(rng, n_frames=5, n_dets=37, T=T)]
    # ragged: drop / add detections in some frames, one empty frame
    videos[0][4] = videos[0][4][:80]
    videos[0][7] = videos[0][7][:0]
    videos[1][2] = np.vstack([videos[1][2], videos[1][2][:9] + 3])
    json_data, boxes = [], []
    for v, frames in enumerate(videos):
        for f, b in enumerate(frames):
            json_data.append({'image': 'vid%02d/%06d.jpg' % (v, f), 'height': 800, 'width': 1333})
            boxes.append(b)
    dets = {'all_boxes': [[], boxes], 'all_keyps': [[], [[np.zeros((4, 17 * T))] * len(b) for b in boxes]]}
    out = te.compute_matches_tracks(json_data, dets, None)
    i = 0
    for frames in videos:
        ref = ot.compute_tracks_video(frames, solver='scipy')
        for f in range(len(frames)):
            assert out['all_tracks'][1][i] == [int(x) for x in ref[f]], (i, f)
            i += 1
    # single pair through the reference-shaped entry point
    m = te._compute_matches(None, None, videos[0][0], videos[0][1], None, None,
                            cfg.TRACKING.DISTANCE_METRICS, cfg.TRACKING.DISTANCE_METRIC_WTS, 'hungarian')
    assert np.array_equal(m, ot.compute_matches(videos[0][0], videos[0][1], solver='scipy'))


def test_prune_and_center_vs_oracle():
    from detectandtrack_b200.core import tracking_engine as te
    from detectandtrack_b200.core.config import cfg, reset_cfg
    reset_cfg()
    cfg.KRCNN.NUM_KEYPOINTS = 17
    rng = np.random.default_rng(8)
    T = 3
    boxes, json_data = [], []
    for i in range(7):
        n = [0, 1, 50, 100, 33, 64, 65][i]
        b = ot.synth_video(rng, 1, max(n, 1), T=T)[0][:n]
        b[:, :-1] += rng.normal(0, 40, b[:, :-1].shape).astype(np.float32)       # some leave the image / shrink
        b[:, -1] = rng.uniform(0.85, 1.0, n)
        boxes.append(b.astype(np.float32))
        json_data.append({'image': 'v/%d.jpg' % i, 'height': 720, 'width': 1280})
    dets = {'all_boxes': [[], [b.copy() for b in boxes]],
            'all_keyps': [[], [[np.full((4, 17 * T), j, np.float32) for j in range(len(b))] for b in boxes]]}
    te._center_detections(dets)
    te._prune_bad_detections(dets, json_data, 0.95)
    for i, b in enumerate(boxes):
        c = ot.center_boxes(b.copy())
        if len(c):
            sel = ot.prune_boxes(c, 720, 1280, 0.95)
            ref = c[sel]
        else:
            sel, ref = np.zeros(0, int), c.reshape(0, 5)
        got = dets['all_boxes'][1][i]
        assert got.shape == ref.shape and np.array_equal(got, ref), i
        assert [int(p[0, 0]) for p in dets['all_keyps'][1][i]] == sel.tolist()


def test_assignment_properties_full_size():
    """At the ABI's largest frame pair (224 x 224, tie-heavy): a permutation with the
    same total cost as scipy's optimum (size-independent optimality check) and equal indices."""
    import torch
    from detectandtrack_b200.ops import box_ops
    rng = np.random.default_rng(2)
    D = 224
    C = np.ones((4, D, D), np.float32)
    for b in range(4):
        m = rng.random((D, D)) < 0.08
        C[b][m] = rng.random(int(m.sum()))
    n = torch.full((4,), D, dtype=torch.int32)
    m, st = box_ops.lsa_batched(_cuda(C), n, n)
    m = m.cpu().numpy()
    for b in range(4):
        assert sorted(m[b].tolist()) == list(range(D))
        r, c = scipy.optimize.linear_sum_assignment(C[b])
        assert np.isclose(C[b][m[b], np.arange(D)].sum(), C[b][r, c].sum(), rtol=0, atol=1e-4)
        assert np.array_equal(m[b], _matches_from(r, c, D))


def test_greedy_matching_vs_oracle():
    """TRACKING.BIPARTITE_MATCHING_ALGO greedy (tracking_engine.py:184-206): identical pairs, incl. ties."""
    import torch
    from detectandtrack_b200.ops import box_ops
    from detectandtrack_b200.core import tracking_engine as te
    from oracle.lsa import bipartite_matching_greedy
    rng = np.random.default_rng(31)
    B, D = 60, 40
    cost = np.zeros((B, D, D), np.float32)
    nr = rng.integers(0, D + 1, B).astype(np.int32); nc = rng.integers(0, D + 1, B).astype(np.int32)
    for b in range(B):
        C = rng.random((nr[b], nc[b])) if b % 3 else np.round(rng.random((nr[b], nc[b])) * 4) / 4     # ties
        cost[b, :nr[b], :nc[b]] = C
    m, _ = box_ops.lsa_batched(_cuda(cost), torch.from_numpy(nr), torch.from_numpy(nc), algo='greedy')
    m = m.cpu().numpy()
    for b in range(B):
        P, Q = nr[b], nc[b]
        ref = -np.ones(Q, np.int32)
        if P and Q:
            pi, qi = bipartite_matching_greedy(cost[b, :P, :Q])
            ref[qi] = pi
        assert np.array_equal(m[b, :Q], ref), b
    vid = ot.synth_video(rng, n_frames=3, n_dets=50)
    got = te._compute_matches(None, None, vid[0], vid[1], None, None, ('bbox-overlap',), (1.0,), 'greedy')
    assert np.array_equal(got, ot.compute_matches(vid[0], vid[1], algo='greedy'))


# ---- 'pose-pck' tracking cost on the device (lib/core/tracking_engine.py:113-129,158-181) ----
POSE_NAMES = ['nose', 'head_bottom', 'head_top', 'left_ear', 'right_ear', 'left_shoulder', 'right_shoulder', 'left_elbow',
              'right_elbow', 'left_wrist', 'right_wrist', 'left_hip', 'right_hip', 'left_knee', 'right_knee', 'left_ankle', 'right_ankle']


@pytest.mark.parametrize('tag', ['small', 'frame', 'far'])
def test_pose_pck_cost_golden(tag):
    """dt_pose_pck_cost == the reference's own _compute_pairwise_kpt_distance (goldens generated by source-executing
    tracking_engine.py:113-129 + keypoints.py:266-291): bit-equal fp64 cost matrix."""
    import os
    import torch
    from detectandtrack_b200.ops import box_ops
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'pose_pck.npz'))
    a, b, ref = g[tag + '_a'], g[tag + '_b'], g[tag + '_dist']
    got = box_ops.pose_pck_cost(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), POSE_NAMES.index('head_top'),
                                POSE_NAMES.index('head_bottom')).cpu().numpy()
    assert got.dtype == np.float64 and got.shape == ref.shape and np.array_equal(got, ref)


def test_tracks_with_pose_pck_cost_vs_reference_loop():
    """TRACKING.DISTANCE_METRICS ('bbox-overlap', 'pose-pck') with weights (1.0, 0.5): ids == the reference's loop restated
    with the oracle costs (fp64 sum of the two weighted matrices, scipy solver)."""
    import scipy.optimize
    from detectandtrack_b200.core import tracking_engine as te
    from detectandtrack_b200.core.config import cfg, reset_cfg
    from oracle import keypoints as okp
    reset_cfg()
    try:
        cfg.TRACKING.DISTANCE_METRICS = ('bbox-overlap', 'cnn-cosdist', 'pose-pck'); cfg.TRACKING.DISTANCE_METRIC_WTS = (1.0, 0.0, 0.5)
        rng = np.random.default_rng(11)
        videos, vposes = [], []
        for nf, nd in ((8, 40), (5, 17)):
            frames = ot.synth_video(rng, n_frames=nf, n_dets=nd)
            poses = []
            for b in frames:                                  # a pose per box: 17 joints inside the box + jitter
                cx = b[:, 0:1] + rng.uniform(0, 1, (b.shape[0], 17)) * (b[:, 2:3] - b[:, 0:1])
                cy = b[:, 1:2] + rng.uniform(0, 1, (b.shape[0], 17)) * (b[:, 3:4] - b[:, 1:2])
                p = np.stack([cx, cy, rng.normal(2, 1, cx.shape), rng.uniform(0, 1, cx.shape)], 1).astype(np.float32)
                poses.append([p[i] for i in range(p.shape[0])])
            videos.append(frames); vposes.append(poses)
        got = te._tracks_for_videos(videos, vposes, POSE_NAMES)
        for v, (frames, poses) in enumerate(zip(videos, vposes)):
            tracks, next_id = [], 0
            for f, cur in enumerate(frames):
                ids = []
                if f == 0:
                    m = -np.ones(cur.shape[0], np.int32)
                else:
                    # tracking_engine.py:166-181: float32 (1 - IoU) * w stacked with the float64 pck * w, summed in float64
                    C = np.sum(np.stack([ot.distance_matrix(frames[f - 1], cur, 1.0),
                                         okp.pairwise_kpt_distance(poses[f - 1], poses[f], POSE_NAMES) * 0.5], 0), 0)
                    m = -np.ones(cur.shape[0], np.int32)
                    pi, qi = scipy.optimize.linear_sum_assignment(C)
                    m[qi] = pi
                for mm in m:
                    if mm == -1:
                        ids.append(next_id); next_id += 1
                        if next_id >= 999:
                            next_id %= 999
                    else:
                        ids.append(tracks[f - 1][mm])
                tracks.append(ids)
            assert got[v] == [[int(x) for x in t] for t in tracks], v
    finally:
        reset_cfg()
