"""tracking plumbing throughput (frame pairs / s).
V videos x 30 frames x 100 detections (synthetic, oracle.tracking.synth_video), linked by
core.tracking_engine (fused (1 - IoU) cost + Hungarian assignment for every frame pair + id scan on the GPU),
next to the CPU restatement of the reference loop (bbox_overlaps -> scipy linear_sum_assignment, 1 core).
Checks that both give identical track ids.   python tests/perf_tracking.py [--videos 64]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--videos', type=int, default=64)
    ap.add_argument('--frames', type=int, default=30)
    ap.add_argument('--dets', type=int, default=100)
    ap.add_argument('--cpu-videos', type=int, default=4)
    a = ap.parse_args()
    import torch
    from detectandtrack_b200.core import tracking_engine as te
    from detectandtrack_b200.core.config import cfg, reset_cfg
    from oracle import tracking as ot                     # checker / CPU baseline only
    reset_cfg()
    cfg.TRACKING.DISTANCE_METRICS = ('bbox-overlap',); cfg.TRACKING.DISTANCE_METRIC_WTS = (1.0,)
    cfg.TRACKING.BIPARTITE_MATCHING_ALGO = 'hungarian'
    rng = np.random.default_rng(3)
    vids = [ot.synth_video(rng, n_frames=a.frames, n_dets=a.dets) for _ in range(a.videos)]
    pairs = a.videos * (a.frames - 1)
    te._tracks_for_videos(vids[:2])                       # warm-up (library load, allocator)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        got = te._tracks_for_videos(vids)                 # host lists in -> host id lists out
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t_gpu = sorted(ts)[len(ts) // 2]
    # CPU baseline: the reference's per-pair loop, scipy solver, one core; also the exactness check
    t0 = time.perf_counter()
    ref = [ot.compute_tracks_video(v, solver='scipy') for v in vids[:a.cpu_videos]]
    t_cpu = time.perf_counter() - t0
    same = all(ref[i] == got[i] for i in range(a.cpu_videos))
    print(json.dumps(dict(metric='tracking frame pairs/s (config 1: %d videos x %d frames x %d dets)' % (a.videos, a.frames, a.dets),
                          value=pairs / t_gpu, unit='frame-pairs/s', wall_ms=1000 * t_gpu, includes='H2D of the boxes, 3 launches, D2H of the ids',
                          cpu_baseline=dict(value=a.cpu_videos * (a.frames - 1) / t_cpu, unit='frame-pairs/s', cores=1, kind='port',
                                            sample='%d videos, oracle.tracking.compute_tracks_video with scipy %s' % (a.cpu_videos, __import__('scipy').__version__)),
                          ids_identical_to_cpu=bool(same))))


if __name__ == '__main__':
    main()
