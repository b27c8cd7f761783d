"""Step-by-step smoke of the training kernels with a synchronize after each launch (bring-up aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from detectandtrack_b200.ops import train_ops as to

def step(name, fn):
    try:
        r = fn(); torch.cuda.synchronize(); print('ok  ', name, flush=True); return r
    except Exception as e:
        print('FAIL', name, str(e)[:300], flush=True); raise

x = torch.randn((2, 3, 20, 28, 128), device='cuda').bfloat16()
gz = torch.randn((2, 3, 20, 28, 128), device='cuda').bfloat16()
xp = step('to_planes x', lambda: to.to_planes(x, pad=(1, 1), copies=True))
gp = step('to_planes gz', lambda: to.to_planes(gz, pad=(1, 1)))
ref = xp[1].float().sum().item()
print('planes checksum', ref, x.float().sum().item())
dW = step('wgrad 3x3x3', lambda: to.wgrad(gp, xp, (20, 28), (3, 3, 3)))
print('dW', dW.abs().max().item())
xp1 = step('to_planes 1x1', lambda: to.to_planes(x, copies=True))
gp1 = step('to_planes 1x1 gz', lambda: to.to_planes(gz))
dW1 = step('wgrad 1x1', lambda: to.wgrad(gp1, xp1, (20, 28), (1, 1, 1)))
ref1 = torch.einsum('ntwhc,ntwhd->cd', gz.float(), x.float())
print('1x1 err', (dW1[0] - ref1).abs().max().item() / ref1.abs().max().item())
