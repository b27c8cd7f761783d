"""GPU: the RoI-head training kernels (csrc/train_ops.cu: dt_roi_align_bwd, dt_frcnn_loss_grad, dt_kps_loss_grad,
dt_subpixel_grad_fix, dt_grad_join_f32) against torch autograd of the oracle's operators in fp32
(torchvision roi_align(aligned=False); Caffe2's SoftmaxWithLoss / Detectron's SmoothL1Loss restated with torch ops:
lib/modeling/model_builder.py:481-493,873-888)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import keypoints as okp


def _rel(got, ref):
    return float((got - ref).abs().max() / max(float(ref.abs().max()), 1e-30))


def test_roi_align_bwd_vs_torchvision_autograd():
    import torch
    from torchvision.ops import roi_align
    from detectandtrack_b200.ops import train_ops as to, rpn_ops
    rng = np.random.RandomState(0)
    B, C, P = 2, 16, 7
    shapes = [(40, 56), (20, 28), (10, 14), (5, 7)]
    feats = [torch.from_numpy(rng.randn(B, C, h, w).astype(np.float32)).requires_grad_(True) for h, w in shapes]
    n = 60
    x1 = rng.uniform(-10, 180, n); y1 = rng.uniform(-10, 120, n)
    side = np.exp(rng.uniform(np.log(8), np.log(700), n))                         # areas that reach every level P2..P5
    rois = np.stack([rng.randint(0, B, n), x1, y1, x1 + side, y1 + side * rng.uniform(0.5, 1.5, n)], 1).astype(np.float32)
    rois_t = torch.from_numpy(rois).cuda()
    lv, _, _ = rpn_ops.distribute(rois_t, None, col0=1, T=1, k_min=2, k_max=5, want_restore=False)
    lvn = lv.cpu().numpy()
    g = torch.from_numpy(rng.randn(n, P, P, C).astype(np.float32)).to(torch.bfloat16)        # bf16-exact upstream gradient
    loss = 0.
    for l in range(4):
        idx = np.where(lvn == l + 2)[0]
        if len(idx) == 0:
            continue
        out = roi_align(feats[l], torch.from_numpy(rois[idx]), (P, P), 1. / 2 ** (l + 2), 2, aligned=False)     # [n, C, P, P]
        loss = loss + (out * g[idx].float().permute(0, 3, 1, 2)).sum()
    loss.backward()
    dfe = [torch.zeros((B, h, w, C), dtype=torch.float32, device='cuda') for h, w in shapes]
    to.roi_align_bwd(g.cuda().view(n, 1, P, P, C), dfe, [1. / 2 ** l for l in range(2, 6)], rois_t, lv, P, 2, T=1, k_min=2)
    for l in range(4):
        ref = (feats[l].grad if feats[l].grad is not None else torch.zeros_like(feats[l])).permute(0, 2, 3, 1)   # no RoI on this level
        assert _rel(dfe[l].cpu(), ref) <= 1e-5, (l, _rel(dfe[l].cpu(), ref))
    # join with a bf16 gradient
    gb = torch.from_numpy(rng.randn(*dfe[0].shape).astype(np.float32)).to(torch.bfloat16).cuda()
    j = to.grad_join_f32(dfe[0], gb)
    assert torch.equal(j, (dfe[0] + gb.float()).to(torch.bfloat16))


def test_frcnn_loss_grad_vs_autograd():
    import torch
    import torch.nn.functional as F
    from detectandtrack_b200.ops import train_ops as to
    rng = np.random.RandomState(1)
    rows, C, ld = 200, 2, 16
    out = torch.from_numpy(rng.randn(rows, ld).astype(np.float32) * 2)
    labels = rng.randint(0, C, rows).astype(np.int32); labels[150:] = -1                     # padding rows
    tg = rng.randn(rows, 4 * C).astype(np.float32) * 2
    iw = np.zeros((rows, 4 * C), np.float32); iw[labels == 1, 4:] = 1.0
    ow = (iw > 0).astype(np.float32)
    N = 150.
    o = out.clone().requires_grad_(True)
    live = torch.from_numpy(labels >= 0)
    lc = F.cross_entropy(o[live][:, :C], torch.from_numpy(labels[labels >= 0].astype(np.int64)), reduction='sum') * 0.5 / N
    d = torch.from_numpy(iw) * (o[:, C:5 * C] - torch.from_numpy(tg))
    ad = d.abs()
    lb = (torch.from_numpy(ow) * torch.where(ad < 1, 0.5 * d * d, ad - 0.5))[live].sum() * 0.5 / N
    (lc + lb).backward()
    loss = torch.zeros(2, device='cuda'); acc = torch.zeros(1, device='cuda')
    totals = torch.tensor([N, 0.], device='cuda')
    g = to.frcnn_loss_grad(out.cuda(), torch.from_numpy(labels).cuda(), torch.from_numpy(tg).cuda(), torch.from_numpy(iw).cuda(),
                           torch.from_numpy(ow).cuda(), C, totals, 0.5, 0.5, ld, loss=loss, accuracy=acc)
    assert abs(float(loss[0]) - float(lc)) <= 1e-5 * abs(float(lc)) and abs(float(loss[1]) - float(lb)) <= 1e-5 * abs(float(lb))
    ref = o.grad
    assert _rel(g.float().cpu()[:, :5 * C], ref[:, :5 * C]) <= 4e-3                          # bf16 storage of the gradient
    assert float(g[:, 5 * C:].float().abs().sum()) == 0 and float(g[150:].float().abs().sum()) == 0
    pred = out[:150, :C].argmax(1).numpy()
    assert float(acc[0]) == float((pred == labels[:150]).sum())


def _unpack_low(low, K):
    """packed [D, S, S, >=4K] -> (D, K, 2S, 2S)"""
    D, S = low.shape[0], low.shape[1]
    x = low[..., :4 * K].reshape(D, S, S, 2, 2, K)                      # (y, x, py, px, k)
    return x.permute(0, 5, 1, 3, 2, 4).reshape(D, K, 2 * S, 2 * S)


def test_kps_loss_grad_vs_autograd():
    import torch
    import torch.nn.functional as F
    from detectandtrack_b200.ops import train_ops as to
    rng = np.random.RandomState(2)
    D, S, K, ld = 6, 14, 17, 72
    low = torch.from_numpy(rng.randn(D, S, S, ld).astype(np.float32) * 3)
    loc = rng.randint(0, 56 * 56, (D, K)).astype(np.int32)
    w = (rng.uniform(0, 1, (D, K)) > 0.3).astype(np.float32); w[D - 1] = 0
    loc[w == 0] = 0
    tw = float(w.sum())
    lo = low.clone().requires_grad_(True)
    up = okp.bilinear_upsample2x(_unpack_low(lo, K))                    # (D, K, 56, 56)
    lp = F.log_softmax(up.reshape(D * K, -1), dim=1)
    nll = -lp[torch.arange(D * K), torch.from_numpy(loc.reshape(-1).astype(np.int64))]
    ref_loss = (nll * torch.from_numpy(w.reshape(-1))).sum() / tw * 0.7
    ref_loss.backward()
    loss = torch.zeros(1, device='cuda')
    totals = torch.tensor([0., tw], device='cuda')
    g = to.kps_loss_grad(low.cuda(), K, torch.from_numpy(loc).cuda(), torch.from_numpy(w).cuda(), totals, 0.7, ld, loss=loss)
    assert abs(float(loss[0]) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss)), (float(loss[0]), float(ref_loss))
    assert _rel(g.float().cpu(), lo.grad) <= 4e-3
    assert float(g[..., 4 * K:].float().abs().sum()) == 0 and float(g[D - 1].float().abs().sum()) == 0


def test_subpixel_grad_fix():
    import torch
    from detectandtrack_b200.ops import train_ops as to
    K, Cin, ldc = 17, 8, 72
    gW = torch.ones((9, ldc, Cin), device='cuda')
    gb = torch.arange(ldc, dtype=torch.float32, device='cuda')
    to.subpixel_grad_fix(gW, gb, K)
    m = gW.cpu().numpy()
    for co in range(ldc):
        for tap in range(9):
            live = False
            if co < 4 * K:
                py, px = (co // K) >> 1, (co // K) & 1
                dy, dx = tap // 3 - 1, tap % 3 - 1
                live = 0 <= py + 1 - 2 * dy <= 3 and 0 <= px + 1 - 2 * dx <= 3
            assert np.all(m[tap, co] == (1.0 if live else 0.0)), (co, tap)
    assert m.sum() == 16 * K * Cin                                      # every one of the 4x4 deconv taps appears exactly once
    b = gb.cpu().numpy()
    for k in range(K):
        assert np.all(b[[k, K + k, 2 * K + k, 3 * K + k]] == 4 * k + 6 * K)
    assert np.all(b[4 * K:] == 0)


def test_roi_align_forward_backward_ignore_non_finite_rois():
    """A diverging training run can emit NaN / inf proposal coordinates: such a RoI must pool zeros and scatter nothing
    (no wild address), leaving the other RoIs untouched."""
    import torch
    from detectandtrack_b200.ops import dense_ops, train_ops as to
    feat = torch.randn((1, 20, 28, 16), device='cuda').to(torch.bfloat16)
    rois = torch.tensor([[0, 4, 4, 40, 60], [0, float('nan'), 3, 50, 50], [0, 2, float('inf'), 30, 40], [0, 10, 8, 70, 44]],
                        dtype=torch.float32, device='cuda')
    out = dense_ops.roi_align([feat], [0.25], rois, None, 7, 2)
    torch.cuda.synchronize()
    assert float(out[1].float().abs().sum()) == 0.0 and torch.isfinite(out.float()).all()
    ref = dense_ops.roi_align([feat], [0.25], rois[[0, 3]].contiguous(), None, 7, 2)
    assert torch.equal(out[[0, 3]], ref)
    d = [torch.zeros((1, 20, 28, 16), dtype=torch.float32, device='cuda')]
    g = torch.ones((4, 1, 7, 7, 16), dtype=torch.bfloat16, device='cuda')
    to.roi_align_bwd(g, d, [0.25], rois, None, 7, 2)
    d2 = [torch.zeros((1, 20, 28, 16), dtype=torch.float32, device='cuda')]
    to.roi_align_bwd(g[[0, 3]].contiguous(), d2, [0.25], rois[[0, 3]].contiguous(), None, 7, 2)
    torch.cuda.synchronize()
    assert torch.isfinite(d[0]).all() and torch.allclose(d[0], d2[0], rtol=1e-5, atol=1e-5)
