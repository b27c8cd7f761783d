"""Micro-benchmark of the tcgen05 filter-gradient kernel (dt_wgrad) on the trainable layer shapes of the R50-FPN-3D trunk
at TRAIN.IMS_PER_BATCH = 2 clips (T = 3, 800x1344 blob).  CUDA events, median of --iters; TFLOP/s = 2*MACs of the conv.
    python tools/bench_wgrad.py [--iters 5] [--only res4]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from detectandtrack_b200.ops import train_ops as to

LAYERS = [
    # name, N, T, H, W, Cin, Cout, k
    ('res3 branch2b 128>128 3x3x3', 2, 3, 100, 168, 128, 128, (3, 3, 3)),
    ('res3 branch2c 128>512 1x1x1', 2, 3, 100, 168, 128, 512, (1, 1, 1)),
    ('res4 branch2b 256>256 3x3x3', 2, 3, 50, 84, 256, 256, (3, 3, 3)),
    ('res4 branch2a 1024>256 1x1x1', 2, 3, 50, 84, 1024, 256, (1, 1, 1)),
    ('res5 branch2b 512>512 3x3x3', 2, 3, 25, 42, 512, 512, (3, 3, 3)),
    ('fpn posthoc P2 256>256 3x3x3', 2, 3, 200, 336, 256, 256, (3, 3, 3)),
    ('fpn posthoc P3 256>256 3x3x3', 2, 3, 100, 168, 256, 256, (3, 3, 3)),
    ('rpn conv P2 256>256 1x3x3', 2, 1, 200, 336, 256, 256, (1, 3, 3)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--only', default='')
    a = ap.parse_args()
    rows = []
    for name, N, T, H, W, Cin, Cout, k in LAYERS:
        if a.only and a.only not in name:
            continue
        x = torch.randn((N, T, H, W, Cin), device='cuda').bfloat16()
        gz = torch.randn((N, T, H, W, Cout), device='cuda').bfloat16()
        pad = (k[1] // 2, k[2] // 2)
        xp = to.to_planes(x, pad=pad, copies=True)
        gp = to.to_planes(gz, pad=pad)
        dW = to.wgrad(gp, xp, (H, W), k)
        torch.cuda.synchronize()
        ts, tp = [], []
        for _ in range(a.iters):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            xp = to.to_planes(x, pad=pad, copies=True); gp = to.to_planes(gz, pad=pad)
            e1.record()
            to.wgrad(gp, xp, (H, W), k, dW)
            e2.record(); torch.cuda.synchronize()
            tp.append(e0.elapsed_time(e1)); ts.append(e1.elapsed_time(e2))
        ms, mp = sorted(ts)[len(ts) // 2], sorted(tp)[len(tp) // 2]
        fl = 2.0 * N * T * H * W * Cin * Cout * k[0] * k[1] * k[2]
        rows.append(dict(layer=name, wgrad_ms=round(ms, 4), planes_ms=round(mp, 4), tflops=round(fl / ms / 1e9, 1), gflop=round(fl / 1e9, 1)))
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(rows, open('gpurun_out/bench_wgrad.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
