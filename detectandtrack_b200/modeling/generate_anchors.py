"""Anchor (tube) enumeration with the reference's interface
(lib/modeling/generate_anchors.py:42-53,134-140).  Host-side float64 table, A x 4T."""
import itertools

import numpy as np

from ..core.config import cfg


def _centered(ws, hs, xc, yc):
    half_w, half_h = 0.5 * (ws - 1), 0.5 * (hs - 1)
    return np.stack([xc - half_w, yc - half_h, xc + half_w, yc + half_h], axis=1)


def generate_anchors(stride=16, sizes=(32, 64, 128, 256, 512), aspect_ratios=(0.5, 1, 2), time_dim=1):
    """Windows centred on (stride-1)/2 with sqrt-areas `sizes` and the given aspect ratios;
    rows ordered ratio-major then size, as the reference enumerates them.  Tubes repeat the
    window per frame ('replicate'), or enumerate combinations / permutations."""
    stride = float(stride)
    ratios = np.asarray(aspect_ratios, dtype=np.float64)
    scales = np.asarray(sizes, dtype=np.float64) / stride
    ctr = 0.5 * (stride - 1)
    # one window per ratio with (rounded) equal area
    ws = np.round(np.sqrt(stride * stride / ratios))
    hs = np.round(ws * ratios)
    rows = []
    for w, h in zip(ws, hs):
        rows.append(_centered(w * scales, h * scales, ctr, ctr))
    anchors = np.vstack(rows)
    style = cfg.VIDEO.RPN_TUBE_GEN_STYLE
    if style == 'replicate':
        return np.tile(anchors, [1, time_dim])
    lst = anchors.tolist()
    if style == 'combinations':
        it = itertools.combinations_with_replacement(lst, time_dim)
    elif style == 'permutations':
        it = itertools.permutations(lst, time_dim)
    else:
        raise NotImplementedError('Unknown {}'.format(style))
    return np.array([sum(item, []) for item in it])


def time_extend_shifts(shifts, time_dim):
    return np.tile(shifts, [1, time_dim])
