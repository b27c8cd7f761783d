#!/usr/bin/env python
"""Inference entry point with the reference's command line (tools/test_net.py:40-63,107-143):

    python tools/test_net.py --cfg X.yaml [--multi-gpu-testing] [--range S E] [KEY VALUE ...]
"""
import argparse
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from detectandtrack_b200.core.config import cfg, cfg_from_file, cfg_from_list, assert_and_infer_cfg  # noqa: E402


def parse_args():
    p = argparse.ArgumentParser(description='Test a Fast R-CNN network')
    p.add_argument('--cfg', dest='cfg_file', help='optional config file', default=None, type=str)
    p.add_argument('--wait', dest='wait', help='wait until net file exists', default=True, type=bool)
    p.add_argument('--vis', dest='vis', help='visualize detections', action='store_true')
    p.add_argument('--multi-gpu-testing', dest='multi_gpu_testing', help='using cfg.NUM_GPUS for inference',
                   action='store_true')
    p.add_argument('--range', dest='range', help='start (inclusive) and end (exclusive) indices', default=None,
                   type=int, nargs=2)
    p.add_argument('opts', help='See core/config.py for all options', default=None, nargs=argparse.REMAINDER)
    if len(sys.argv) == 1:
        p.print_help()
        sys.exit(1)
    return p.parse_args()


def main(ind_range=None, multi_gpu_testing=False):
    from detectandtrack_b200.core import test_engine as engine
    if cfg.MODEL.RPN_ONLY:
        raise NotImplementedError('RPN-only proposal dumping is outside the hot path')
    if ind_range is not None:
        engine.test_net(ind_range=ind_range)                 # child of the multi-GPU fan-out
    else:
        if len(cfg.TEST.DATASETS) == 0:
            cfg.TEST.DATASETS = (cfg.TEST.DATASET,)
        for ds in cfg.TEST.DATASETS:
            cfg.TEST.DATASET = ds
            engine.test_net_on_dataset(multi_gpu=multi_gpu_testing)


if __name__ == '__main__':
    logging.basicConfig(level=logging.INFO, format='%(levelname)s %(filename)s:%(lineno)4d: %(message)s', stream=sys.stdout)
    args = parse_args()
    if args.cfg_file is not None:
        cfg_from_file(args.cfg_file)
    if args.opts:
        cfg_from_list(args.opts)
    assert_and_infer_cfg()
    if args.vis:
        cfg.VIS = True
    w = cfg.TEST.WEIGHTS
    while w not in ('', 'random') and not os.path.exists(w) and args.wait:      # test_net.py:139-141
        logging.info('Waiting for %s to exist...', w)
        time.sleep(10)
    main(ind_range=args.range, multi_gpu_testing=args.multi_gpu_testing)
