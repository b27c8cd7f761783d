"""Global option tree, drop-in for the reference's ``core.config``.

Public surface kept (lib/core/config.py:764-900): ``cfg``, ``cfg_default``,
``cfg_from_file``, ``cfg_from_cfg``, ``cfg_from_list``, ``assert_and_infer_cfg``,
``get_output_dir``.  Behaviour kept: unknown keys raise KeyError unless a
``<KEY>_deprecated`` twin exists (then the rest of that sub-tree is skipped, as
the reference returns early), string values go through ``literal_eval``, type
mismatches raise ValueError (ndarray targets coerce), the legacy scalar
``VIDEO.TIME_KERNEL_DIM`` fans out to BODY/HEAD_RPN/HEAD_KPS/HEAD_DET.
"""
import copy
import logging
import os
import os.path as osp
from ast import literal_eval

import numpy as np

from .config_defaults import DEFAULTS

logger = logging.getLogger(__name__)


class AttrDict(dict):
    """dict with attribute access (lib/utils/collections.py:15-27)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _to_attr(d):
    return AttrDict({k: _to_attr(v) if isinstance(v, dict) else v for k, v in d.items()})


cfg = _to_attr(DEFAULTS)
cfg_default = copy.deepcopy(cfg)


def reset_cfg():
    """Restore defaults in place (tests use this; the reference has no equivalent)."""
    fresh = copy.deepcopy(cfg_default)
    cfg.clear()
    cfg.update(fresh)


def assert_and_infer_cfg():
    """lib/core/config.py:764-774."""
    if cfg.MODEL.RPN_ONLY or cfg.MODEL.FASTER_RCNN:
        cfg.RPN.ON = True
    if cfg.MODEL.RPN_ONLY:
        cfg.TRAIN.BBOX_REG = False
    if cfg.VIDEO.NUM_FRAMES_MID == -1:
        cfg.VIDEO.NUM_FRAMES_MID = cfg.VIDEO.NUM_FRAMES
    assert (not cfg.MODEL.USE_BN_TESTMODE_ONLY) or cfg.MODEL.USE_BN


def get_output_dir(training=True):
    """<OUTPUT_DIR>/<train|test>/<dataset>/<model-type>/  (config.py:777-785)."""
    dataset = cfg.TRAIN.DATASET if training else cfg.TEST.DATASET
    if os.sep in str(dataset):                   # a JSON roidb given by path (test_engine.JsonListDataset): its file stem names the directory
        dataset = osp.splitext(osp.basename(str(dataset)))[0]
    outdir = osp.join(cfg.OUTPUT_DIR, 'train' if training else 'test', dataset, cfg.MODEL.TYPE)
    os.makedirs(outdir, exist_ok=True)
    return outdir


def _coerce(key, new, old):
    """Type rule of config.py:812-822, py3 flavoured."""
    if new is None or type(old) is type(new) or old is None:
        return new
    if isinstance(old, np.ndarray):
        return np.array(new, dtype=old.dtype)
    if isinstance(old, str) and isinstance(new, (str, bytes)):
        return new.decode() if isinstance(new, bytes) else str(new)
    if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
        # yaml writes 1 for 1.0; the py2 reference would raise here, we accept the widening
        return float(new)
    raise ValueError('Type mismatch ({} vs. {}) for config key: {}'.format(type(old), type(new), key))


def _merge(src, dst, path=''):
    for k, v in src.items():
        if k not in dst:
            if k + '_deprecated' in dst:
                logger.warning('Config key %s%s is deprecated, ignoring', path, k)
                return
            raise KeyError('{} is not a valid config key'.format(path + k))
        if isinstance(v, dict):
            if not isinstance(dst[k], dict):
                raise ValueError('Type mismatch (dict vs. {}) for config key: {}'.format(type(dst[k]), path + k))
            _merge(v, dst[k], path + k + '.')
            continue
        if isinstance(v, str):
            try:
                v = literal_eval(v)
            except Exception:
                pass
        dst[k] = _coerce(path + k, v, dst[k])


def _config_mapping_rules(a):
    """config.py:839-861: legacy int VIDEO.TIME_KERNEL_DIM -> per-part dict."""
    vid = a.get('VIDEO') if isinstance(a, dict) else None
    if isinstance(vid, dict) and isinstance(vid.get('TIME_KERNEL_DIM'), int):
        val = vid['TIME_KERNEL_DIM']
        vid['TIME_KERNEL_DIM'] = {k: val for k in cfg_default.VIDEO.TIME_KERNEL_DIM.keys()}
    return a


def cfg_from_file(filename):
    """Load a yaml file and merge it into ``cfg`` (config.py:865-873)."""
    import yaml
    with open(filename, 'r') as f:
        y = yaml.safe_load(f) or {}
    _merge(_config_mapping_rules(y), cfg)


def cfg_from_cfg(other):
    _merge(_config_mapping_rules(dict(other)), cfg)


def cfg_from_list(cfg_list):
    """KEY VALUE KEY VALUE ... overrides from the command line (config.py:880-900)."""
    assert len(cfg_list) % 2 == 0, 'expected KEY VALUE pairs'
    for k, v in zip(cfg_list[0::2], cfg_list[1::2]):
        d = cfg
        parts = k.split('.')
        for sub in parts[:-1]:
            assert sub in d, 'Config key {} not found'.format(sub)
            d = d[sub]
        sub = parts[-1]
        assert sub in d, 'Config key {} not found'.format(sub)
        try:
            value = literal_eval(v) if isinstance(v, str) else v
        except Exception:
            value = v
        old = d[sub]
        if isinstance(old, float) and isinstance(value, int) and not isinstance(value, bool):
            value = float(value)
        assert old is None or isinstance(value, type(old)), \
            'type {} does not match original type {}'.format(type(value), type(old))
        d[sub] = value
