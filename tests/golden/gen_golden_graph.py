"""Graph tables + weight-inflation goldens taken from the REFERENCE'S OWN SOURCE.  Build container only.

    python tests/golden/gen_golden_graph.py   -> oracle/graph_tables.json, tests/golden/inflate_weights.npz

(1) graph_tables.json: block counts / feature dims / block type of every ``add_ResNet*_conv{4,5}_body`` and
    ``add_ResNet*_roi_conv5_head`` builder, and the FPN level tables, read with ``ast`` from
    lib/modeling/ResNet3D.py:334-394, lib/modeling/ResNet.py:298-397 (nothing is imported: the modules need
    Caffe2).  oracle/graph.py builds the oracle's graph description from this file, so the oracle does not share
    the product's reading of the builders (modeling/params.py), and tests/test_params.py checks the product's
    GraphSpec against it.
(2) inflate_weights.npz: lib/utils/net.py:95-161 ``inflate_weights`` executed from its source text (the module
    imports Caffe2) for every VIDEO.WEIGHTS_INFLATE_MODE, time sizes 1 and 3; 'center-only-rest-rand' with a fixed
    numpy seed.
"""
import ast
import json
import os
import types

import numpy as np

REF = '/root/reference/lib'
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _lit(node):
    return ast.literal_eval(node)


def parse_bodies(path):
    """{'ResNet50_conv5': dict(counts, dims, trans_func)} and {'ResNet18': dict(dim_out, block_counts)} heads."""
    tree = ast.parse(open(path).read())
    default_dims = None
    for fn in tree.body:
        if isinstance(fn, ast.FunctionDef) and fn.name == 'add_ResNet_convX_body':
            names = [a.arg for a in fn.args.args]
            defaults = dict(zip(names[len(names) - len(fn.args.defaults):], fn.args.defaults))
            default_dims = list(_lit(defaults['feat_dims']))
    bodies, heads = {}, {}
    for fn in tree.body:
        if not isinstance(fn, ast.FunctionDef) or not fn.name.startswith('add_ResNet') or fn.name.startswith('add_ResNet_'):
            continue
        key = fn.name[len('add_'):]
        if key.endswith('_body'):
            trans = None
            for st in ast.walk(fn):
                if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Attribute) and st.targets[0].attr == 'TRANS_FUNC':
                    trans = _lit(st.value)
                if isinstance(st, ast.Call) and getattr(st.func, 'id', '') == 'add_ResNet_convX_body':
                    counts = list(_lit(st.args[1]))
                    dims = default_dims
                    for kw in st.keywords:
                        if kw.arg == 'feat_dims':
                            dims = list(_lit(kw.value))
            bodies[key[:-len('_body')]] = dict(counts=counts, dims=dims[:len(counts) + 1], trans_func=trans)
        elif key.endswith('_roi_conv5_head'):
            d = {}
            for st in ast.walk(fn):
                if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Subscript):
                    d[_lit(st.targets[0].slice)] = _lit(st.value)
            heads[key[:-len('_roi_conv5_head')]] = d
    return bodies, heads


def parse_fpn_levels(path):
    """stage_info_ResNet*_conv5 (ResNet.py:363-397): blobs / dims / spatial scales, coarsest first."""
    tree = ast.parse(open(path).read())
    out = {}
    for fn in tree.body:
        if isinstance(fn, ast.FunctionDef) and fn.name.startswith('stage_info_'):
            for st in ast.walk(fn):
                if isinstance(st, ast.Call) and getattr(st.func, 'attr', getattr(st.func, 'id', '')) == 'ConvStageInfo':
                    kw = {k.arg: k.value for k in st.keywords}
                    scales = [float(eval(compile(ast.Expression(e), '<scale>', 'eval'))) for e in kw['spatial_scales'].elts]
                    out[fn.name[len('stage_info_'):]] = dict(blobs=list(_lit(kw['blobs'])), dims=list(_lit(kw['dims'])), spatial_scales=scales)
    return out


def gen_tables():
    b3, h3 = parse_bodies(os.path.join(REF, 'modeling', 'ResNet3D.py'))
    b2, h2 = parse_bodies(os.path.join(REF, 'modeling', 'ResNet.py'))
    tables = dict(source='lib/modeling/ResNet3D.py:334-394, lib/modeling/ResNet.py:298-397 (parsed with ast)',
                  ResNet3D=dict(bodies=b3, roi_conv5_heads=h3), ResNet=dict(bodies=b2, roi_conv5_heads=h2),
                  fpn_levels=parse_fpn_levels(os.path.join(REF, 'modeling', 'ResNet.py')))
    with open(os.path.join(ROOT, 'oracle', 'graph_tables.json'), 'w') as f:
        json.dump(tables, f, sort_keys=True, separators=(",", ":"))
    return tables


def gen_inflate():
    src = open(os.path.join(REF, 'utils', 'net.py')).read().split('\n')
    # lines 54-161: inflate_weights_2d + inflate_weights (1-based 54..161)
    start = next(i for i, l in enumerate(src) if l.startswith('def inflate_weights_2d'))
    end = next(i for i, l in enumerate(src) if l.startswith('def initialize_gpu_0_from_weights_file'))
    import logging
    cfg = types.SimpleNamespace(VIDEO=types.SimpleNamespace(WEIGHTS_INFLATE_MODE='center-only'))
    ns = {'np': np, 'cfg': cfg, 'logger': logging.getLogger('ref')}
    # numpy >= 1.x: np.repeat needs an int count; the reference passes the float `ncopies` (numpy 1.14 accepted it)
    body = '\n'.join(src[start:end]).replace('pretrained_w, axis=-3), ncopies, axis=-3)', 'pretrained_w, axis=-3), int(ncopies), axis=-3)')
    exec(body, ns)
    rng = np.random.RandomState(7)
    w2d = rng.randn(6, 4, 3, 3).astype(np.float32)
    out = {'w2d': w2d}
    for mode in ('mean-repeat', 'repeat', 'center-only', 'center-only-rest-rand'):
        for kt in (1, 3):
            cfg.VIDEO.WEIGHTS_INFLATE_MODE = mode
            np.random.seed(1234)
            ws = np.zeros((6, 4, kt, 3, 3), np.float32)
            out['%s_%d' % (mode, kt)] = np.asarray(ns['inflate_weights'](w2d, ws, 'w', {'w': w2d}), dtype=np.float32)
    # non-5-D target: returned as is (2-D conv blob of the same rank goes through inflate_weights_2d)
    out['same_rank'] = np.asarray(ns['inflate_weights'](w2d, np.zeros_like(w2d), 'w', {'w': w2d}), dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, 'inflate_weights.npz'), **out)


if __name__ == '__main__':
    t = gen_tables()
    print(json.dumps(t['ResNet3D']['bodies'], sort_keys=True))
    gen_inflate()
    print('ok')
