"""Per-stage relative error of the engine vs the torch-fp32 oracle graph (small clip)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from test_gpu_engine import _cfg
from oracle import net as onet
from detectandtrack_b200.modeling import params as P
from detectandtrack_b200.modeling.engine import DetectionEngine
from detectandtrack_b200.ops import dense_ops, conv as cv

cfg = _cfg()
blobs, spec = P.random_blobs(cfg, seed=3)
rng = np.random.RandomState(0)
frames = rng.randint(0, 256, (1, 3, 96, 128, 3)).astype(np.uint8)
means = np.asarray(cfg.PIXEL_MEANS, np.float32).reshape(1, 1, 1, 1, 3)
data = torch.from_numpy((frames.astype(np.float32) - means)).permute(0, 4, 1, 2, 3).contiguous()
with torch.no_grad():
    stages = onet.conv_body(blobs, spec, data)
    pyr = onet.fpn(blobs, spec, stages)


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item(), b.abs().max().item()


for mode in ('tf32', 'bf16'):
    eng = DetectionEngine(cfg, blobs, spec, dtype=mode)
    fr = torch.from_numpy(frames).cuda()
    x = dense_ops.prep_clip(fr.view(3, 96, 128, 3), eng.pixel_means, 1.0, (96, 128), (96, 128), cpad=eng.cin_pad, out_f32=(mode == 'tf32'), border=(3, 4), row_planes=True).view(1, 3, 2, 51, 136, eng.cin_pad)
    outs = eng.body(x)
    for i, o in enumerate(outs):
        r = stages[spec.stage_blobs[i]]
        print(mode, 'stage', spec.stage_blobs[i], rel(o.permute(0, 4, 1, 2, 3).float().cpu(), r))
    f = eng.fpn(outs)
    for i, o in enumerate(f):
        r = pyr[::-1][i]
        print(mode, 'fpn level', i + 2, rel(o.permute(0, 4, 1, 2, 3).float().cpu(), r))
