"""CPU: the resume logic of tools/train_net.py (reference tools/train_net.py:79-105): model_final.pkl ends training, else the
newest model_iter<N>.pkl restarts at N + 1."""
import importlib.util
import os


def _mod():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('train_net', os.path.join(root, 'tools', 'train_net.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_find_resume_point(tmp_path):
    tn = _mod()
    d = str(tmp_path)
    assert tn.find_resume_point(os.path.join(d, 'missing')) == (0, None)
    assert tn.find_resume_point(d) == (0, None)
    for n in ('model_iter19999.pkl', 'model_iter39999.pkl', 'model_iter7.pkl', 'net.pbtxt'):
        open(os.path.join(d, n), 'wb').close()
    assert tn.find_resume_point(d) == (40000, os.path.join(d, 'model_iter39999.pkl'))
    open(os.path.join(d, 'model_final.pkl'), 'wb').close()
    assert tn.find_resume_point(d) == ('final', os.path.join(d, 'model_final.pkl'))


def test_train_loop_snapshots_and_resume_with_a_stub_model(tmp_path, monkeypatch):
    """The host side of tools/train_net.py without a GPU: the device trainer is replaced by a stub; the loop must log every
    20 iterations, write model_iter<N>.pkl every TRAIN.SNAPSHOT_ITERS / NUM_GPUS iterations and model_final.pkl at the end,
    skip training when the final model exists, and resume after the newest snapshot (tools/train_net.py:79-105,208-218)."""
    import sys
    import numpy as np
    import torch
    from detectandtrack_b200.core.config import cfg, cfg_from_file, cfg_from_list, assert_and_infer_cfg, reset_cfg
    from detectandtrack_b200.modeling import model_builder, trainer
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_tools import TRAIN_YAML
    tn = _mod()
    steps = []

    class Stub(object):
        blobs0 = {'a_w': np.zeros((2, 2), np.float32)}
        totals = torch.tensor([10., 5.])

        def step(self, frames, gt):
            assert tuple(frames.shape) == (2, 3, 96, 128, 3) and frames.dtype == torch.uint8 and len(gt) == 2
            steps.append(self.lr)
            return torch.tensor([0.5, 0.1]), torch.tensor([0.2, 0.1, 3.0, 8.0])

        def export_blobs(self, b):
            return dict(b)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    monkeypatch.setattr(model_builder, 'create', lambda *a, **k: Stub())
    monkeypatch.setattr(trainer, 'pack_gt', lambda gts, K=17: gts)
    reset_cfg()
    try:
        y = tmp_path / 'cfg.yaml'
        y.write_text(TRAIN_YAML)
        cfg_from_file(str(y))
        cfg_from_list(['OUTPUT_DIR', str(tmp_path / 'out'), 'SOLVER.MAX_ITER', '45', 'TRAIN.SNAPSHOT_ITERS', '20'])
        assert_and_infer_cfg()
        d = os.path.join(str(tmp_path / 'out'), 'train', 'synthetic_1x2_96x128', 'keypoint_rcnn')
        assert len(tn.train_model()) == 4 and len(steps) == 45
        assert sorted(os.listdir(d)) == ['model_final.pkl', 'model_iter19.pkl', 'model_iter39.pkl']
        assert steps[0] < steps[4] < steps[10] and steps[44] < steps[10]            # linear warm-up, then the decay step at 30
        cfg_from_list(['CLUSTER.ON_CLUSTER', 'True'])
        assert tn.train_model() == [] and len(steps) == 45                          # model_final.pkl exists: nothing to do
        os.remove(os.path.join(d, 'model_final.pkl'))
        tn.train_model()
        assert len(steps) == 50 and os.path.exists(os.path.join(d, 'model_final.pkl'))     # resumed at iteration 40
    finally:
        reset_cfg()
