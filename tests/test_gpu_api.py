"""GPU: the fast path behind the reference's surface (core/test.py).  im_detect_all (1-clip captured graph) and the
batched ClipPipeline that test_engine.test_net drives must return exactly what the eager engine returns, clip by clip,
in the reference's containers (lib/core/test.py:897-958)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def model():
    from test_gpu_engine import _cfg
    from detectandtrack_b200.modeling import params as P, model_builder
    cfg = _cfg()
    blobs, _ = P.random_blobs(cfg, seed=3)
    m = model_builder.create(cfg.MODEL.TYPE, train=False, blobs=blobs)        # dtype None -> cfg.TEST.PRECISION
    assert m.engine.dtype_name == 'bf16x3' == cfg.TEST.PRECISION
    return m


def _clips(n, seed=0):
    rng = np.random.RandomState(seed)
    return [[rng.randint(0, 256, (96, 128, 3)).astype(np.uint8) for _ in range(3)] for _ in range(n)]


def test_im_detect_all_equals_eager_engine(model):
    import torch
    from detectandtrack_b200.core.test import im_detect_all
    im = _clips(1)[0]
    cls_boxes, cls_segms, cls_keyps = im_detect_all(model, im)
    ref = model.engine.detect(torch.from_numpy(np.stack(im)[None]).cuda())[0]
    assert cls_segms is None and len(cls_boxes) == 2 and cls_boxes[0] == []
    assert cls_boxes[1].dtype == np.float32 and cls_boxes[1].shape[1] == 5 and cls_boxes[1].shape[0] > 0
    assert np.array_equal(cls_boxes[1], ref['boxes'].cpu().numpy())
    k = ref['keyps'].cpu().numpy()
    assert len(cls_keyps[1]) == k.shape[0] and all(np.array_equal(a, b) for a, b in zip(cls_keyps[1], k))
    # a second call reuses the captured graph and the pinned buffers: same answer
    again = im_detect_all(model, im)
    assert np.array_equal(again[0][1], cls_boxes[1])


def test_batched_pipeline_equals_per_clip_calls(model):
    """5 clips through a 2-clip captured step (ragged last batch, loader threads, double buffering) == 5 im_detect_all calls."""
    from detectandtrack_b200.core.test import im_detect_all, im_detect_all_batch
    ims = _clips(5, seed=4)
    batched = im_detect_all_batch(model, ims, clips_per_step=2)
    assert len(batched) == 5
    for im, (cb, cs, ck) in zip(ims, batched):
        rb, _, rk = im_detect_all(model, im)
        assert np.array_equal(cb[1], rb[1])
        assert (ck is None) == (rk is None)
        if ck is not None:
            assert len(ck[1]) == len(rk[1]) and all(np.array_equal(a, b) for a, b in zip(ck[1], rk[1]))


def test_detect_roidb_places_results_by_index(model):
    """test_engine.detect_roidb (the loop of test_net): mixed frame sizes split into runs, results land at the entry's index."""
    from detectandtrack_b200.core import test_engine as te
    from detectandtrack_b200.core.config import cfg
    from detectandtrack_b200.core.test import im_detect_all
    old = cfg.TEST.CLIPS_PER_STEP
    cfg.TEST.CLIPS_PER_STEP = 2
    try:
        roidb = [dict(image=['v/%d.jpg' % i] * 3, height=96, width=128, seed=100 + i, synthetic=True) for i in range(3)]
        roidb += [dict(image=['w/%d.jpg' % i] * 3, height=64, width=96, seed=200 + i, synthetic=True) for i in range(2)]
        all_boxes, _, all_keyps = te.empty_results(2, len(roidb))
        te.detect_roidb(model, roidb, all_boxes, all_keyps)
        for i, e in enumerate(roidb):
            rb, _, rk = im_detect_all(model, te.read_image_video(e))
            assert np.array_equal(all_boxes[1][i], rb[1]), i
            assert len(all_keyps[1][i]) == (0 if rk is None else len(rk[1]))
    finally:
        cfg.TEST.CLIPS_PER_STEP = old
