"""Inference data parallelism, as lib/utils/subprocess.py:27-74: the index range is split into
NUM_GPUS contiguous chunks (np.array_split), one child ``tools/test_net.py --range s e`` per GPU
via CUDA_VISIBLE_DEVICES, each child writes ``<tag>_range_s_e.pkl``; the parent concatenates in
rank order.  No collective is involved."""
import logging
import os
import pickle
import subprocess
import sys

import numpy as np
import yaml

from ..core.config import cfg

logger = logging.getLogger(__name__)


def split_ranges(total, n):
    parts = np.array_split(range(total), n)
    return [(int(p[0]), int(p[-1]) + 1) for p in parts if len(p)]


def process_in_parallel(tag, total_range_size, binary, output_dir, extra_opts=()):
    cfg_file = os.path.join(output_dir, '{}_range_config.yaml'.format(tag))
    with open(cfg_file, 'w') as f:
        yaml.safe_dump(_plain(cfg), f)
    env = os.environ.copy()
    procs = []
    for i, (s, e) in enumerate(split_ranges(total_range_size, cfg.NUM_GPUS)):
        env['CUDA_VISIBLE_DEVICES'] = str(i)
        cmd = [sys.executable, binary, '--range', str(s), str(e), '--cfg', cfg_file, 'NUM_GPUS', '1'] + list(extra_opts)
        log = open(os.path.join(output_dir, '%s_range_%s_%s.stdout' % (tag, s, e)), 'w')
        logger.info('%s range command %d: %s', tag, i, ' '.join(cmd))
        procs.append((i, subprocess.Popen(cmd, env=env, stdout=log, stderr=subprocess.STDOUT, bufsize=1), s, e, log))
    outputs = []
    for i, p, s, e, log in procs:
        rc = p.wait()
        log.close()
        assert rc == 0, 'Range subprocess {} [{}, {}) failed (exit code {})'.format(i, s, e, rc)
        with open(os.path.join(output_dir, '%s_range_%s_%s.pkl' % (tag, s, e)), 'rb') as f:
            outputs.append(pickle.load(f))
    return outputs


def _plain(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out[k] = _plain(v)
        elif isinstance(v, np.ndarray):
            out[k] = v.tolist()
        elif isinstance(v, tuple):
            out[k] = str(v)
        elif isinstance(v, (np.floating, np.integer)):
            out[k] = v.item()
        elif v is None or (isinstance(v, float) and v in (float('inf'), float('-inf'))):
            continue
        else:
            out[k] = v
    return out
