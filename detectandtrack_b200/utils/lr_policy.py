"""Learning-rate schedule of the training loop (lib/utils/lr_policy.py:19-125): cfg.SOLVER.LR_POLICY in
{'step', 'steps_with_decay', 'steps_with_lrs'} with the constant / linear warm-up of the first WARM_UP_ITERS iterations."""
import numpy as np

from ..core.config import cfg


def _step_index(it):
    steps = list(cfg.SOLVER.STEPS)
    assert steps and steps[0] == 0, 'The first step should always start at 0.'
    bounds = steps + [cfg.SOLVER.MAX_ITER]
    ind = 0
    for ind, b in enumerate(bounds):
        if it < b:
            break
    return ind - 1


def _base_lr(it):
    pol = cfg.SOLVER.LR_POLICY
    if pol == 'step':
        return cfg.SOLVER.BASE_LR * cfg.SOLVER.GAMMA ** (it // cfg.SOLVER.STEP_SIZE)
    if pol == 'steps_with_decay':
        return cfg.SOLVER.BASE_LR * cfg.SOLVER.GAMMA ** _step_index(it)
    if pol == 'steps_with_lrs':
        return cfg.SOLVER.LRS[_step_index(it)]
    raise NotImplementedError('Unknown LR policy: {}'.format(pol))


def get_lr_at_iter(it):
    lr = _base_lr(it)
    if it < cfg.SOLVER.WARM_UP_ITERS:
        m = cfg.SOLVER.WARM_UP_METHOD
        if m == 'constant':
            f = cfg.SOLVER.WARM_UP_FACTOR
        elif m == 'linear':
            a = it / float(cfg.SOLVER.WARM_UP_ITERS)
            f = cfg.SOLVER.WARM_UP_FACTOR * (1 - a) + a
        else:
            raise KeyError('Unknown SOLVER.WARM_UP_METHOD: {}'.format(m))
        lr *= f
    return np.float32(lr)
