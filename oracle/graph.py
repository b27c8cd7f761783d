"""Oracle: the graph description the torch-fp32 oracle network (oracle/net.py, oracle/pipeline.py) is built from
(TEST INFRASTRUCTURE ONLY).

Independent of the product's reading (detectandtrack_b200/modeling/params.py:GraphSpec): block counts, feature
dims, block type and the FPN level tables come from ``graph_tables.json``, which tests/golden/gen_golden_graph.py
extracts with ``ast`` from the reference's own builder source (lib/modeling/ResNet3D.py:334-394,
lib/modeling/ResNet.py:298-397); the cfg-dependent rest follows the reference call sites cited per attribute.
tests/test_params.py::test_graphspec_matches_reference_tables checks the product against the same tables."""
import json
import os

_TABLES = None


def tables():
    global _TABLES
    if _TABLES is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'graph_tables.json')) as f:
            _TABLES = json.load(f)
    return _TABLES


class OracleSpec(object):
    def __init__(self, cfg):
        module, func = cfg.MODEL.CONV_BODY.split('.')            # model_builder.get_func (model_builder.py:39-49)
        self.fpn = func.startswith('add_fpn_')                   # FPN[3D].add_fpn_<body>: FPN.py:32-75 / FPN3D.py:27-70
        body = func[len('add_fpn_'):] if self.fpn else func[len('add_'):]
        assert body.endswith('_body'), cfg.MODEL.CONV_BODY
        body = body[:-len('_body')]
        self.is3d = module.endswith('3D')
        t = tables()['ResNet3D' if self.is3d else 'ResNet']
        b = t['bodies'][body]
        self.counts = tuple(b['counts'])
        self.dims = tuple(b['dims'])
        self.block = 'bottleneck' if b['trans_func'].startswith('bottleneck') else 'basic'   # cfg.RESNETS.TRANS_FUNC
        video = bool(cfg.MODEL.VIDEO_ON)
        self.T = cfg.VIDEO.NUM_FRAMES if video else 1
        self.tk_body = cfg.VIDEO.TIME_KERNEL_DIM.BODY if self.is3d else 1          # ResNet3D.py:276-296
        self.stride_1x1 = bool(cfg.RESNETS.STRIDE_1X1)                             # ResNet3D.py:186-189
        self.link = cfg.VIDEO.BODY_HEAD_LINK if video else 'none2d'               # model_builder.py:1024-1042
        self.head3d = video and cfg.VIDEO.BODY_HEAD_LINK == ''
        self.T_head = (cfg.VIDEO.NUM_FRAMES_MID if cfg.VIDEO.NUM_FRAMES_MID > 0 else self.T) if self.head3d else 1
        self.num_classes = cfg.MODEL.NUM_CLASSES
        self.K = cfg.KRCNN.NUM_KEYPOINTS
        if self.fpn:
            lv = tables()['fpn_levels'][body]                    # ResNet.py:363-397 (coarsest first)
            self.stage_blobs = list(lv['blobs'])[::-1]           # finest first: res2_x_sum ... res5_x_sum
            assert list(lv['dims'])[::-1] == list(self.dims[1:]), (lv['dims'], self.dims)
            self.rpn_levels = list(range(cfg.FPN.RPN_MIN_LEVEL, cfg.FPN.RPN_MAX_LEVEL + 1))       # FPN.py:205-279
            self.roi_levels = list(range(cfg.FPN.ROI_MIN_LEVEL, cfg.FPN.ROI_MAX_LEVEL + 1))       # FPN.py:349-381
            self.num_anchors = len(cfg.FPN.RPN_ASPECT_RATIOS)                                     # FPN.py:213
        else:
            self.stage_blobs = ['res%d_%d_sum' % (s + 2, n - 1) for s, n in enumerate(self.counts)]
            self.num_anchors = len(cfg.RPN.SIZES) * len(cfg.RPN.ASPECT_RATIOS)                    # model_builder.py:505
        head = cfg.MODEL.ROI_HEAD.split('.')[1]
        self.roi_conv5 = None
        if 'roi_conv5_head' in head:                             # ResNet3D.py:301-331 (dim_out / block_counts)
            arch = head[len('add_'):].split('_')[0]
            self.roi_conv5 = dict(t['roi_conv5_heads'][arch])
