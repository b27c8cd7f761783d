"""CPU: the learning-rate schedule of the training loop against the worked examples in the reference's docstrings
(lib/utils/lr_policy.py:41-76) and its warm-up formula (:24-33)."""
import numpy as np


def test_lr_policies_match_reference_examples():
    from detectandtrack_b200.core.config import cfg, reset_cfg
    from detectandtrack_b200.utils.lr_policy import get_lr_at_iter
    reset_cfg()
    try:
        cfg.SOLVER.WARM_UP_ITERS = 0
        cfg.SOLVER.MAX_ITER = 90; cfg.SOLVER.STEPS = [0, 60, 80]; cfg.SOLVER.BASE_LR = 0.02; cfg.SOLVER.GAMMA = 0.1
        cfg.SOLVER.LR_POLICY = 'steps_with_decay'
        assert [float(get_lr_at_iter(i)) for i in (0, 59, 60, 79, 80, 89)] == [np.float32(v) for v in (0.02, 0.02, 0.002, 0.002, 0.0002, 0.0002)]
        cfg.SOLVER.LR_POLICY = 'steps_with_lrs'; cfg.SOLVER.LRS = [0.02, 0.002, 0.0002]
        assert [float(get_lr_at_iter(i)) for i in (0, 60, 80)] == [np.float32(v) for v in (0.02, 0.002, 0.0002)]
        cfg.SOLVER.LR_POLICY = 'step'; cfg.SOLVER.STEP_SIZE = 30
        assert float(get_lr_at_iter(65)) == np.float32(0.02 * 0.1 ** 2)
        cfg.SOLVER.WARM_UP_ITERS = 10; cfg.SOLVER.WARM_UP_METHOD = 'linear'; cfg.SOLVER.WARM_UP_FACTOR = 1.0 / 3
        assert abs(float(get_lr_at_iter(0)) - 0.02 / 3) < 1e-9 and abs(float(get_lr_at_iter(5)) - 0.02 * (1.0 / 3 * 0.5 + 0.5)) < 1e-9
        assert float(get_lr_at_iter(10)) == np.float32(0.02)
    finally:
        reset_cfg()
