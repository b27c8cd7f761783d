"""GPU: one training step of the RPN model trunk (modeling/trainer.py: res3..res5 + FPN + RPN heads, bf16) against
torch autograd of the oracle graph in fp32 on the CPU (oracle/net.py with requires_grad tensors; Detectron's
SigmoidCrossEntropyLoss / SmoothL1Loss restated with torch ops, lib/modeling/FPN.py:282-321).

Tolerance.  The kernels themselves are checked tightly op by op (tests/test_gpu_train_ops.py: wgrad / dgrad <= 2e-3 on
identical bf16 inputs).  HERE the device forward is bf16 (config 5 trains in bf16) and the oracle's is fp32, so besides the
2^-9 rounding of every stored activation / gradient the ReLU masks differ wherever a pre-activation lies within the
forward error of zero (~1 % of the units per layer); measured per-layer gradient agreement on this graph: cosine 0.993 ..
1.000, max-norm relative error 0.4 % (rpn_out) .. 15 % (res4_0_branch2b).  Asserted: losses <= 2e-2 relative; every
checked filter / bias gradient has cosine >= 0.99 and max-norm error <= 0.2 * max|ref| — a wiring error (a missing
branch, a wrong level, a sign) shows up as a cosine far below that."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import net as onet
from oracle.graph import OracleSpec


def _oracle_step(cfg, blobs, frames, targets_np, world=1):
    import torch
    import torch.nn.functional as F
    spec = OracleSpec(cfg)
    tb = {k: torch.from_numpy(np.ascontiguousarray(v)).clone() for k, v in blobs.items()}
    train = [k for k in tb if (k.startswith(('res3', 'res4', 'res5', 'fpn_', 'conv_rpn', 'rpn_')) and not k.endswith(('_bn_s', '_bn_b')))]
    for k in train:
        tb[k].requires_grad_(True)
    means = np.asarray(cfg.PIXEL_MEANS, np.float32).reshape(1, 1, 1, 1, 3)
    data = torch.from_numpy(frames.astype(np.float32) - means).permute(0, 4, 1, 2, 3).contiguous()
    pyr = onet.fpn(tb, spec, onet.conv_body(tb, spec, data))                      # [P6..P2]
    feats = [onet.time_pool(p, 'slice-center', cfg.VIDEO.NUM_FRAMES_MID) for p in pyr]
    heads = onet.rpn_heads_fpn(tb, spec, feats)                                   # finest first: (logits (B,A,H,W), deltas (B,4A,H,W))
    B = frames.shape[0]
    s_cls = 1.0 / world / cfg.TRAIN.RPN_BATCH_SIZE_PER_IM / cfg.TRAIN.IMS_PER_BATCH
    s_box = 1.0 / world / B
    beta = 1.0 / 9.0
    lc, lb = 0., 0.
    for (lg, dl), t in zip(heads, targets_np):
        lab = torch.from_numpy(t['labels']).permute(0, 3, 1, 2)                   # (B,A,H,W)
        m = lab >= 0
        lc = lc + s_cls * F.binary_cross_entropy_with_logits(lg[m], lab[m].float(), reduction='sum')
        tg, iw, ow = (torch.from_numpy(t[k]).permute(0, 3, 1, 2) for k in ('bbox_targets', 'inside', 'outside'))
        d = iw * (dl - tg)
        ad = d.abs()
        lb = lb + s_box * (ow * torch.where(ad < beta, 0.5 * d * d / beta, ad - 0.5 * beta)).sum()
    (lc + lb).backward()
    return float(lc), float(lb), {k: tb[k].grad for k in train}


def test_rpn_trunk_training_step_vs_autograd():
    import torch
    from test_gpu_engine import _cfg
    from detectandtrack_b200.modeling import params as P
    from detectandtrack_b200.modeling.trainer import RpnTrainer
    cfg = _cfg()
    blobs, spec = P.random_blobs(cfg, seed=3)
    frames = np.random.RandomState(0).randint(0, 256, (2, 3, 64, 96, 3)).astype(np.uint8)
    cfg.TEST.SCALES = (64,); cfg.TEST.MAX_SIZE = 96
    try:
        tr = RpnTrainer(cfg, blobs, spec, lr=0.01, weight_decay=1e-4)
        targets = tr.synthetic_targets(2, 64, 96, seed=1)
        tnp = [{k: v.cpu().numpy() for k, v in t.items()} for t in targets]
        outs = tr.forward_all(torch.from_numpy(frames).cuda())
        loss = tr.backward(outs, targets).cpu().numpy()
        lc, lb, grads = _oracle_step(cfg, blobs, frames, tnp)
        assert abs(loss[0] - lc) <= 2e-2 * abs(lc) and abs(loss[1] - lb) <= 2e-2 * abs(lb), (loss, lc, lb)

        def dev_grad(c):
            kT, kH, kW = c.k
            return c.g.view(kT, kH, kW, c.cout, c.cin).permute(3, 4, 0, 1, 2).cpu()

        bad = []

        def check(name, got, ref, tol=0.2):
            ref = ref.reshape(got.shape)
            err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
            cos = float((got.flatten() * ref.flatten()).sum() / (got.norm() * ref.norm() + 1e-30))
            if err > tol or cos < 0.99:
                bad.append((name, round(err, 4), round(cos, 5)))
            return round(err, 4), round(cos, 5)

        errs = {}
        errs['conv_rpn'] = check('conv_rpn_fpn2_w', dev_grad(tr.rpn_conv), grads['conv_rpn_fpn2_w'])
        A = tr.A
        g_out = dev_grad(tr.rpn_out)[:5 * A]
        ref_out = torch.cat([grads['rpn_cls_logits_fpn2_w'], grads['rpn_bbox_pred_fpn2_w']], 0)
        errs['rpn_out'] = check('rpn_out', g_out[:, :, 0], ref_out)
        errs['rpn_out_b'] = check('rpn_out_b', tr.rpn_out.bias_g[:5 * A].cpu(), torch.cat([grads['rpn_cls_logits_fpn2_b'], grads['rpn_bbox_pred_fpn2_b']], 0))
        names = spec.stage_blobs[::-1]
        for i, n in enumerate(names):
            errs['post%d' % i] = check('fpn_' + n, dev_grad(tr.post[i]), grads['fpn_%s_w' % n])
            errs['post%d_b' % i] = check('fpn_%s_b' % n, tr.post[i].bias_g.cpu(), grads['fpn_%s_b' % n])
            lname = 'fpn_inner_%s%s' % (n, '' if i == 0 else '_lateral')
            errs['lat%d' % i] = check(lname, dev_grad(tr.lat[i]), grads[lname + '_w'])
        stage_names = ['res3', 'res4', 'res5']
        for si, blocks in enumerate(tr.stages):
            for bi in (0, len(blocks) - 1):
                pre = '%s_%d' % (stage_names[si], bi)
                for br, key in (('a', '_branch2a'), ('b', '_branch2b'), ('c', '_branch2c'), ('sc', '_branch1')):
                    if blocks[bi][br] is not None:
                        errs[pre + key] = check(pre + key, dev_grad(blocks[bi][br]), grads[pre + key + '_w'])
        # the update moves the master weights by lr * (g + wd * w) (momentum buffer starts at 0) and refreshes the bf16 filters
        c = tr.stages[-1][-1]['c']
        w0, g0 = c.w.clone(), c.g.clone()
        tr.update()
        torch.cuda.synchronize()
        assert torch.allclose(c.w, w0 - 0.01 * (g0 + 1e-4 * w0), rtol=1e-5, atol=1e-8)
        assert torch.equal(c.w_fwd, c.w.to(torch.bfloat16))
        print('grad (max-norm rel err, cosine) per layer', errs)
        assert not bad, bad
    finally:
        _cfg()
