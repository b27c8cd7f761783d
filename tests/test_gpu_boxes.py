"""GPU parity (through the C ABI): IoU and NMS vs the oracle and the goldens.
Bar: bit-exact IoU values (fp32) and identical NMS index lists."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import boxes as ob


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _clustered(rng, n, T, spread=10.0):
    c = np.hstack([rng.uniform(0, 1200, (max(n // 8, 1), 1)), rng.uniform(0, 700, (max(n // 8, 1), 1))])
    ctr = c[rng.integers(0, c.shape[0], n)] + rng.normal(0, spread, (n, 2))
    wh = rng.uniform(20, 150, (n, 2))
    b = np.hstack([ctr, ctr + wh])
    parts = [b]
    for t in range(1, T):
        s = rng.normal(0, 5, (n, 2))
        parts.append(parts[-1] + np.hstack([s, s]))
    sc = rng.permutation(n)[:, None] / float(n) + 1e-4
    return np.hstack(parts + [sc]).astype(np.float32)


def test_bbox_overlaps_golden(golden):
    from detectandtrack_b200.utils import boxes as B, cython_bbox
    g = golden('iou2d')
    assert np.array_equal(cython_bbox.bbox_overlaps(g['iou2d_a'], g['iou2d_b']), g['iou2d_out'])
    g = golden('iou3')
    assert np.array_equal(B.bbox_overlaps(g['iou3_a'], g['iou3_b']), g['iou3_out'])
    g = golden('iou5')
    assert np.array_equal(B.bbox_overlaps(g['iou5_a'], g['iou5_b']), g['iou5_out'])


@pytest.mark.parametrize('n,k,T', [(1, 1, 1), (333, 1000, 1), (100, 100, 3), (64, 31, 2), (2000, 1500, 1)])
def test_bbox_overlaps_vs_oracle(n, k, T):
    from detectandtrack_b200.ops import box_ops
    rng = np.random.default_rng(n * 7 + k)
    a = _clustered(rng, n, T)[:, :-1]
    b = _clustered(rng, k, T)[:, :-1]
    out = box_ops.bbox_overlaps(_cuda(a), _cuda(b), T=T).cpu().numpy()
    assert np.array_equal(out, ob.bbox_overlaps(a, b))


def test_bbox_overlaps_degenerate_and_empty():
    from detectandtrack_b200.ops import box_ops
    a = np.array([[0, 0, 10, 10], [5, 5, 5, 5], [20, 20, 10, 10], [0, 0, 0, 0]], np.float32)
    out = box_ops.bbox_overlaps(_cuda(a), _cuda(a), T=1).cpu().numpy()
    assert np.array_equal(out, ob.bbox_overlaps_2d(a, a))
    assert box_ops.bbox_overlaps(_cuda(a[:0]), _cuda(a), T=1).shape == (0, 4)


@pytest.mark.parametrize('name', ['nms2d', 'nms2d_small', 'nmst3', 'nmst2'])
def test_nms_golden(golden, name):
    from detectandtrack_b200.core import nms_wrapper
    g = golden(name.split('_')[0])
    d = g[name + '_dets']
    for th in (0.3, 0.5, 0.7):
        keep = np.asarray(nms_wrapper.nms(d, th), dtype=np.int64)
        assert np.array_equal(keep, g['%s_keep_%d' % (name, int(th * 10))])


@pytest.mark.parametrize('n,T,th', [(1, 1, 0.5), (63, 1, 0.5), (64, 1, 0.3), (65, 3, 0.5), (1000, 1, 0.7),
                                    (1000, 3, 0.7), (3000, 1, 0.5), (777, 2, 0.4)])
def test_nms_vs_oracle(n, T, th):
    from detectandtrack_b200.core import nms_wrapper
    d = _clustered(np.random.default_rng(n + 13 * T), n, T)
    keep = nms_wrapper.nms(d, th)
    ref = ob.nms(d, th)
    assert list(map(int, keep)) == list(map(int, ref))


def test_nms_batched_counts_and_max_keep():
    import torch
    from detectandtrack_b200.ops import box_ops
    rng = np.random.default_rng(5)
    B, nmax, T = 6, 700, 3
    counts = np.array([700, 0, 1, 64, 333, 129], np.int32)
    dets = np.stack([_clustered(rng, nmax, T) for _ in range(B)])
    keep, num = box_ops.nms_batched(_cuda(dets), torch.from_numpy(counts), 0.6, max_keep=50)
    keep, num = keep.cpu().numpy(), num.cpu().numpy()
    for b in range(B):
        ref = ob.nms(dets[b, :counts[b]], 0.6)[:50] if counts[b] else []
        assert num[b] == len(ref)
        assert keep[b, :num[b]].tolist() == list(map(int, ref))
    # 2-D, index order: truncation applies to the ascending-index list
    d2 = np.stack([_clustered(rng, 500, 1) for _ in range(3)])
    keep, num = box_ops.nms_batched(_cuda(d2), None, 0.5, max_keep=20)
    for b in range(3):
        ref = ob.nms_2d(d2[b], 0.5)[:20]
        assert keep[b, :int(num[b])].cpu().numpy().tolist() == ref.tolist()


def test_nms_score_ties_follow_documented_rule():
    from detectandtrack_b200.core import nms_wrapper
    d = _clustered(np.random.default_rng(1), 300, 1)
    d[:, 4] = np.round(d[:, 4] * 10) / 10          # heavy score ties
    assert list(nms_wrapper.nms(d, 0.5)) == list(ob.nms_2d(d, 0.5))


def test_nms_properties_full_size():
    """Size-independent properties at the largest supported problem (8192 boxes):
    survivors are mutually below the threshold, every suppressed box overlaps an
    earlier survivor, and NMS of the survivors is the identity."""
    import torch
    from detectandtrack_b200.ops import box_ops
    n, th = 8192, 0.5
    d = _clustered(np.random.default_rng(9), n, 1, spread=25.0)
    keep, num = box_ops.nms_batched(_cuda(d[None]), None, th, box_ops.NMS_2D_GE, box_ops.ORDER_SCORE)
    keep = keep[0, :int(num[0])].cpu().numpy()
    assert len(np.unique(keep)) == len(keep) and np.all(np.diff(d[keep, 4]) < 0)
    iou = box_ops.bbox_overlaps(_cuda(d[keep, :4]), _cuda(d[keep, :4]), T=1).cpu().numpy()
    # NMS IoU (fp32 path) and cython_bbox IoU can differ in the last ulp: compare off the threshold
    np.fill_diagonal(iou, 0)
    assert iou.max() < th + 1e-5
    sup = np.setdiff1d(np.arange(n), keep)
    iou2 = box_ops.bbox_overlaps(_cuda(d[sup, :4]), _cuda(d[keep, :4]), T=1).cpu().numpy()
    higher = d[keep, 4][None, :] > d[sup, 4][:, None]
    assert np.all((np.where(higher, iou2, 0)).max(1) >= th - 1e-5)
    keep2, num2 = box_ops.nms_batched(_cuda(d[keep][None]), None, th, box_ops.NMS_2D_GE, box_ops.ORDER_SCORE)
    assert int(num2[0]) == len(keep) and keep2[0, :len(keep)].cpu().numpy().tolist() == list(range(len(keep)))


def test_nms_rejects_bad_arguments():
    import torch
    from detectandtrack_b200.ops import box_ops
    with pytest.raises(RuntimeError, match='exceeds'):
        box_ops.nms_batched(torch.zeros((1, 9000, 5), device='cuda'), None, 0.5)
    with pytest.raises(RuntimeError, match='T=1'):
        box_ops.nms_batched(torch.zeros((1, 10, 13), device='cuda'), None, 0.5, cmp_mode=box_ops.NMS_2D_GE)
