"""CPU: oracle.lsa.lsap_crouse must return the SAME indices as the installed
scipy (the implementation the reference's call at tracking_engine.py:237 reaches
in this image), including on tie-dominated and rectangular matrices."""
import numpy as np
import pytest
import scipy.optimize

from oracle.lsa import lsap_crouse, bipartite_matching_greedy
from oracle import tracking as ot


def test_lsa_golden(golden):
    g = golden('lsa')
    keys = [k[len('lsa_C'):] for k in g if k.startswith('lsa_C')]
    assert len(keys) >= 7
    for k in keys:
        r, c = lsap_crouse(g['lsa_C' + k])
        assert np.array_equal(r, g['lsa_r' + k]) and np.array_equal(c, g['lsa_c' + k])


def test_lsa_matches_installed_scipy_random():
    rng = np.random.default_rng(11)
    for trial in range(400):
        P, Q = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        kind = trial % 4
        if kind == 0:
            C = rng.random((P, Q)).astype(np.float32)
        elif kind == 1:
            C = rng.integers(0, 3, (P, Q)).astype(np.float32)
        elif kind == 2:
            C = np.ones((P, Q), np.float32); m = rng.random((P, Q)) < 0.1; C[m] = rng.random(int(m.sum()))
        else:
            C = np.full((P, Q), 1.0, np.float32)
        r, c = scipy.optimize.linear_sum_assignment(C)
        r2, c2 = lsap_crouse(C)
        assert np.array_equal(r, r2) and np.array_equal(c, c2), (trial, P, Q)


def test_lsa_empty():
    r, c = lsap_crouse(np.zeros((0, 5)))
    assert r.size == 0 and c.size == 0


def test_tracking_oracle_equals_scipy_path():
    frames = ot.synth_video(np.random.default_rng(3), n_frames=8, n_dets=60)
    a = ot.compute_tracks_video(frames, solver='oracle')
    b = ot.compute_tracks_video(frames, solver='scipy')
    assert a == b
    # every detection of frame t>0 is matched when counts are equal (no gating, :244-246)
    assert sorted(a[1]) == sorted(a[0])


def test_track_id_wraparound():
    # 10 frames of 120 brand-new boxes each would exceed 999 ids -> wraps with '%=' (:341-345)
    rng = np.random.default_rng(0)
    frames = []
    for f in range(10):
        x1 = rng.uniform(0, 1000, 120); y1 = rng.uniform(0, 600, 120)
        n = 120 if f % 2 == 0 else 0
        frames.append(np.stack([x1, y1, x1 + 50, y1 + 80, np.ones(120)], 1).astype(np.float32)[:n])
    tr = ot.compute_tracks_video(frames)
    flat = [i for t in tr for i in t]
    assert max(flat) <= 998 and len(flat) == 600 and flat[599] == 599


def test_greedy_matches_reference_semantics():
    C = np.array([[0.9, 0.1, 0.5], [0.2, 0.8, 0.3]], np.float32)
    p, c = bipartite_matching_greedy(C)
    assert (p, c) == ([0, 1], [1, 0])
