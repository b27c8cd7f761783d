"""GPU parity of the HBM-bound NDHWC kernels vs plain PyTorch / torchvision / OpenCV fp32
references (stand-ins for un-vendored Caffe2 / OpenCV 3.4 arithmetic; tolerance 1e-3 rel)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import keypoints as ok


def test_prep_clip_identity_scale_and_padding():
    import torch
    from detectandtrack_b200.ops import dense_ops
    rng = np.random.default_rng(0)
    F, H, W = 3, 50, 70
    fr = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    means = (102.9801, 115.9465, 122.7717)
    out = dense_ops.prep_clip(torch.from_numpy(fr).cuda(), means, 1.0, (H, W), (64, 96), cpad=8, out_f32=True).cpu().numpy()
    ref = np.zeros((F, 64, 96, 8), np.float32)
    ref[:, :H, :W, :3] = fr.astype(np.float32) - np.array(means, np.float32)
    # fp32 output is rounded to tf32 (10-bit mantissa) for the kind::tf32 consumer: |err| <= 2^-11 |x|
    np.testing.assert_allclose(out, ref, rtol=2.0 ** -11, atol=0)
    out_bf = dense_ops.prep_clip(torch.from_numpy(fr).cuda(), means, 1.0, (H, W), (64, 96), cpad=8, out_f32=False).float().cpu().numpy()
    assert np.array_equal(out_bf, torch.from_numpy(ref).bfloat16().float().numpy())
    # physical zero border (conv1 wants 3 rows / 4 pixels)
    ob = dense_ops.prep_clip(torch.from_numpy(fr).cuda(), means, 1.0, (H, W), (64, 96), cpad=8, out_f32=False, border=(3, 4)).float().cpu().numpy()
    assert ob.shape == (F, 70, 104, 8) and np.array_equal(ob[:, 3:67, 4:100], out_bf)
    assert np.all(ob[:, :3] == 0) and np.all(ob[:, 67:] == 0) and np.all(ob[:, :, :4] == 0) and np.all(ob[:, :, 100:] == 0)
    # rows de-interleaved by parity (the layout conv1 reads): padded row r lives at [r & 1][r >> 1]
    op = dense_ops.prep_clip(torch.from_numpy(fr).cuda(), means, 1.0, (H, W), (64, 96), cpad=8, out_f32=False, border=(3, 4),
                             row_planes=True).float().cpu().numpy()
    assert op.shape == (F, 2, 35, 104, 8)
    assert np.array_equal(op[:, 0], ob[:, 0::2]) and np.array_equal(op[:, 1], ob[:, 1::2])


@pytest.mark.parametrize('scale', [1.6, 0.53])
def test_prep_clip_resize_vs_cv2(scale):
    import cv2
    import torch
    from detectandtrack_b200.ops import dense_ops
    rng = np.random.default_rng(1)
    H, W = 45, 61
    fr = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    means = np.array([[[102.9801, 115.9465, 122.7717]]])
    refs = []
    for f in range(2):
        im = fr[f].astype(np.float32, copy=True)
        im -= means
        refs.append(cv2.resize(im, None, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR))
    Hr, Wr = refs[0].shape[:2]
    Hp, Wp = (Hr + 31) // 32 * 32, (Wr + 31) // 32 * 32
    out = dense_ops.prep_clip(torch.from_numpy(fr).cuda(), means.ravel(), scale, (Hr, Wr), (Hp, Wp), cpad=4, out_f32=True).cpu().numpy()
    for f in range(2):
        np.testing.assert_allclose(out[f, :Hr, :Wr, :3], refs[f], rtol=1e-3, atol=2e-3)
        assert np.all(out[f, Hr:] == 0) and np.all(out[f, :, Wr:] == 0) and np.all(out[f, ..., 3:] == 0)


@pytest.mark.parametrize('k,s,p', [(3, 2, 1), (1, 2, 0)])
@pytest.mark.parametrize('dt', ['f32', 'bf16'])
def test_maxpool(k, s, p, dt):
    import torch
    import torch.nn.functional as Fn
    from detectandtrack_b200.ops import dense_ops
    x = torch.randn((3, 25, 42, 64))
    if dt == 'bf16':
        x = x.bfloat16()
    y = dense_ops.maxpool2d(x.cuda().contiguous(), k, s, p).cpu().float()
    ref = Fn.max_pool2d(x.float().permute(0, 3, 1, 2), k, s, p).permute(0, 2, 3, 1)
    assert y.shape == ref.shape and torch.equal(y, ref)


@pytest.mark.parametrize('T', [1, 3])
def test_roi_align_vs_torchvision(T):
    import torch
    from torchvision.ops import roi_align as tv_roi_align
    from detectandtrack_b200.ops import dense_ops, rpn_ops
    g = torch.Generator().manual_seed(3)
    nimg, C = 2, 64
    sizes = [(48, 80), (24, 40), (12, 20), (6, 10)]
    scales = [1 / 4., 1 / 8., 1 / 16., 1 / 32.]
    feats = [torch.randn((nimg * T, h, w, C), generator=g) for (h, w) in sizes]
    R = 150
    rng = np.random.default_rng(5)
    x1 = rng.uniform(-10, 300, R); y1 = rng.uniform(-10, 170, R)
    w = rng.uniform(2, 330, R); h = rng.uniform(2, 200, R)
    rois = np.zeros((R, 4 * T + 1), np.float32)
    rois[:, 0] = rng.integers(0, nimg, R)
    for t in range(T):
        rois[:, 1 + 4 * t:5 + 4 * t] = np.stack([x1, y1, x1 + w, y1 + h], 1) + rng.normal(0, 2, (R, 4)) * (t > 0)
    rois_d = torch.from_numpy(rois).cuda()
    levels, _, _ = rpn_ops.distribute(rois_d, None, col0=1, T=T)
    out = dense_ops.roi_align([f.cuda().contiguous() for f in feats], scales, rois_d, levels, 7, 2, T=T).cpu()
    lv = levels.cpu().numpy()
    for t in range(T):
        for l in range(4):
            idx = np.where(lv == l + 2)[0]
            if not len(idx):
                continue
            bt = np.hstack([rois[idx, :1] * T + t, rois[idx, 1 + 4 * t:5 + 4 * t]]).astype(np.float32)   # RoIToBatchFormat
            ref = tv_roi_align(feats[l].permute(0, 3, 1, 2).contiguous(), torch.from_numpy(bt), (7, 7), scales[l], 2, aligned=False)
            got = out[idx, t].permute(0, 3, 1, 2)
            assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), (t, l, (got - ref).abs().max())   # fp32 FMA/order noise


def test_keypoint_decode_vs_cv2_oracle():
    import torch
    from detectandtrack_b200.ops import dense_ops
    g = torch.Generator().manual_seed(9)
    D, K, S, T = 12, 17, 14, 1
    lowres = torch.randn((D, K, 2 * S, 2 * S), generator=g) * 2.0            # kps_score_lowres (28x28)
    # smooth the maps a little so maxima are well separated (the argmax under cubic resize is then stable)
    lowres = torch.nn.functional.avg_pool2d(lowres, 3, 1, 1) * 3
    heat_ref = ok.bilinear_upsample2x(lowres)                                  # (D, K, 56, 56)
    rng = np.random.default_rng(2)
    x1 = rng.uniform(0, 600, D); y1 = rng.uniform(0, 400, D)
    bw = rng.uniform(0.4, 260, D); bh = rng.uniform(20, 300, D)
    boxes = np.stack([x1, y1, x1 + bw, y1 + bh], 1).astype(np.float32)
    # pack to the sub-pixel layout [D, S, S, 4K]: channel (py*2+px)*K + k
    packed = lowres.view(D, K, S, 2, S, 2).permute(0, 2, 4, 3, 5, 1).reshape(D, S, S, 4 * K).contiguous()
    xy, heat = dense_ops.keypoint_decode(packed.cuda(), torch.from_numpy(boxes).cuda(), K, T, want_heatmaps=True)
    np.testing.assert_allclose(heat.cpu().numpy(), heat_ref.numpy(), rtol=1e-5, atol=1e-5)
    ref = ok.heatmaps_to_keypoints(heat_ref.numpy(), boxes, K)
    got = xy.cpu().numpy()
    np.testing.assert_allclose(got[:, 2], ref[:, 2], rtol=1e-3, atol=1e-3)      # logit at the max
    np.testing.assert_allclose(got[:, 3], ref[:, 3], rtol=2e-3, atol=1e-6)      # prob
    # positions: identical unless two resized samples tie within float noise (none expected here)
    np.testing.assert_allclose(got[:, :2], ref[:, :2], rtol=0, atol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['bf16', 'f32', 'x3', 'bf16x3'])
def test_time_mean(mode):
    """dt_time_mean (the 'avg' body/head link): mean over T of [B, T, H, W, C]."""
    import torch
    from detectandtrack_b200.ops import dense_ops, conv as cv
    g = torch.Generator().manual_seed(5)
    x = torch.randn((2, 3, 5, 7, 64), generator=g)
    if mode == 'bf16':
        xd = x.bfloat16().cuda()
        y = dense_ops.time_mean(xd).float().cpu()
        ref = x.bfloat16().float().mean(dim=1, keepdim=True)
        tol = 1e-2
    elif mode == 'f32':
        y = dense_ops.time_mean(x.cuda()).cpu()
        ref = x.mean(dim=1, keepdim=True)
        tol = 1e-6
    elif mode == 'bf16x3':
        y = cv.join_split(dense_ops.time_mean(cv.split_bf16(x.cuda()), x3=True)).cpu()
        ref = x.mean(dim=1, keepdim=True)
        tol = 3e-5          # bf16 pairs: 2^-17 on the way in, 2^-17 on the way out
    else:
        y = cv.join_tf32(dense_ops.time_mean(cv.split_tf32(x.cuda()), x3=True)).cpu()
        ref = x.mean(dim=1, keepdim=True)
        tol = 1e-6
    assert y.shape == (2, 1, 5, 7, 64)
    assert (y - ref).abs().max().item() <= tol * ref.abs().max().item()
