"""Data-parallel training step on the device (BASELINE.json configs[4]; reference: tools/train_net.py:129-229,
lib/modeling/model_builder.py:908-985 build_data_parallel_model / add_parameter_update_ops).

Implemented graph: the reference's RPN training model (``MODEL.TYPE rpn``, model_builder.py ``rpn`` + FPN.add_fpn_rpn_outputs
+ FPN.add_fpn_rpn_losses, lib/modeling/FPN.py:205-321) on the FPN / FPN3D ResNet-50/101/152 bodies: conv1 / pool1 / res2
frozen (freeze_at=2, ResNet3D.py:273-274; AffineChannel parameters frozen everywhere), res3..res5 + FPN + RPN heads
trained: forward in bf16 (fp32 accumulate), backward =
    dgrad   dt_conv3d with the flipped / transposed filter (+ dt_scatter_stride2 for the stride-2 1x1 convs)
    wgrad   dt_wgrad on channel-major planes (tcgen05, split-K)
    joins   dt_bwd_pointwise (Relu / Sum / AffineChannelNd gradients), dt_upsample_add_bwd (FPN top-down), dt_bias_grad
    losses  dt_rpn_loss_grad per level (SigmoidCrossEntropyLoss + SmoothL1Loss)
then a bucketed gradient SUM all-reduce over NCCL issued per bucket as soon as its filter gradients are enqueued
(overlapping the rest of the backward pass; losses carry 1/NUM_GPUS like the reference, model_builder.py:484) and the fused
MomentumSGDUpdate on fp32 master weights that re-emits the bf16 forward and dgrad filters.

The RoI heads of the keypoint R-CNN training graph (RoIAlign backward, Fast R-CNN / keypoint losses, the target
generators of lib/roi_data) are NOT implemented: config 5 proper is this trunk plus those heads."""
import numpy as np

from .. import _lib as L
from ..ops import conv as cv, dense_ops, train_ops as to
from . import params as P
from .engine import DetectionEngine
from .generate_anchors import generate_anchors


def plan_buckets(counts, nbuckets):
    """counts: parameter counts in BACKWARD order -> (cumulative ends per parameter, bucket end offsets).  Buckets hold
    ~equal parameter counts and are cut at parameter boundaries, so a bucket is complete as soon as the gradient of its
    last parameter has been enqueued."""
    bounds, off = [], 0
    for n in counts:
        off += int(n)
        bounds.append(off)
    total = off
    ends, target = [], total / float(max(1, nbuckets))
    for b in bounds:
        if b >= target * (len(ends) + 1) or b == total:
            if not ends or b > ends[-1]:
                ends.append(b)
    return bounds, ends


class BucketReducer(object):
    """Gradient SUM all-reduce of a flat buffer in buckets, each launched (async, on the process group's own stream) as soon
    as the producer says the buffer is complete up to an offset — the data-parallel exchange of
    lib/modeling/model_builder.py:922-942 (one NCCLAllreduce per parameter there), overlapped with the backward pass."""

    def __init__(self, flat, bucket_ends, world):
        self.flat, self.ends, self.world = flat, list(bucket_ends), world
        self.reset()

    def reset(self):
        self.next, self.pending = 0, []

    def ready(self, upto):
        if self.world <= 1:
            return
        import torch.distributed as dist
        while self.next < len(self.ends) and self.ends[self.next] <= upto:
            lo = self.ends[self.next - 1] if self.next else 0
            self.pending.append(dist.all_reduce(self.flat[lo:self.ends[self.next]], op=dist.ReduceOp.SUM, async_op=True))
            self.next += 1

    def wait(self):
        for w in self.pending:
            w.wait()
        assert self.world <= 1 or self.next == len(self.ends), 'a gradient bucket was never marked ready'
        self.pending = []


class TrainConv(object):
    """One trainable conv: packed fp32 master filter [taps, Cout, Cin], momentum, gradient (a view into the trainer's
    flat gradient buffer), the bf16 forward / dgrad filters, and either a frozen AffineChannel (scale, bias) or a
    trainable bias."""

    def __init__(self, torch, w, scale=None, shift=None, bias=None, stride=(1, 1, 1), relu=False):
        w = torch.from_numpy(np.ascontiguousarray(w)).float()
        if w.dim() == 4:
            w = w[:, :, None]
        self.k = tuple(w.shape[2:])
        self.pad = tuple(x // 2 for x in self.k)
        self.stride, self.relu = stride, relu
        self.cout, self.cin = w.shape[0], w.shape[1]
        self.taps = self.k[0] * self.k[1] * self.k[2]
        self.w = w.permute(2, 3, 4, 0, 1).reshape(self.taps, self.cout, self.cin).contiguous().cuda()
        self.m = torch.zeros_like(self.w)
        self.g = None                                     # assigned by the trainer (view into the flat buffer)
        self.w_fwd = self.w.to(torch.bfloat16)
        self.w_dg = torch.empty((self.taps, self.cin, self.cout), dtype=torch.bfloat16, device='cuda')
        self.w_dg.copy_(to.pack_dgrad_weight(w))
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda() if a is not None else None
        self.scale, self.shift = f32(scale), f32(shift)
        self.bias = f32(bias)                             # trainable (FPN / RPN convs); None for AffineChannel convs
        self.bias_m = torch.zeros_like(self.bias) if self.bias is not None else None
        self.bias_g = None

    def nparams(self):
        return self.w.numel() + (self.bias.numel() if self.bias is not None else 0)

    def forward(self, x, residual=None, res_mode=0, out_f32=False, out=None):
        return cv.conv3d(x, self.w_fwd, self.k, self.stride, self.pad, self.scale, self.shift if self.bias is None else self.bias,
                         residual, res_mode, self.relu, out_f32=out_f32, dtype=cv.BF16, out=out)

    def backward(self, gz, x, x_planes=None, need_dx=True):
        """gz: gradient wrt the conv's raw output (after the pointwise joins), x: the saved input.  Accumulates dW (and db),
        returns (dx or None, x_planes)."""
        N, T, Ho, Wo, _ = gz.shape
        sp = (self.pad[1], self.pad[2])
        if x_planes is None:
            x_planes = to.to_planes(x, pad=sp, stride=self.stride[1:], channels=self.cin, copies=True)
        assert self.cout % 8 == 0 and gz.shape[-1] == self.cout
        to.wgrad(to.to_planes(gz, pad=sp), x_planes, (Ho, Wo), self.k, self.g)
        if self.bias is not None:
            L.call('dt_bias_grad', L.ptr(gz), gz.numel() // gz.shape[-1], self.cout, gz.shape[-1], L.ptr(self.bias_g), L.stream_ptr())
        dx = None
        if need_dx:
            dx = cv.conv3d(gz, self.w_dg, self.k, (1, 1, 1), self.pad, out_f32=False, dtype=cv.BF16, cin=self.cout)
            if self.stride[1] == 2:
                dx = to.scatter_stride2(dx, (x.shape[2], x.shape[3]))
        return dx, x_planes

    def update(self, lr, momentum, wd, grad_scale):
        to.sgd_update(self.w, self.g, self.m, lr, momentum, wd, grad_scale, self.w_fwd, self.w_dg)
        if self.bias is not None:     # biases: no weight decay, 2x learning rate (model_builder.py:971-976)
            to.sgd_update(self.bias.view(1, 1, -1), self.bias_g.view(1, 1, -1), self.bias_m.view(1, 1, -1), 2.0 * lr, momentum, 0.0,
                          grad_scale)


class RpnTrainer(object):
    def __init__(self, cfg, blobs, spec=None, world=1, buckets=4, lr=None, momentum=None, weight_decay=None):
        torch = L.require_cuda()
        self.torch, self.cfg, self.world = torch, cfg, world
        self.spec = s = spec or P.GraphSpec(cfg)
        assert s.fpn and s.block == 'bottleneck' and not s.head3d, 'RpnTrainer: FPN bottleneck bodies with 2-D RPN heads'
        self.lr = cfg.SOLVER.BASE_LR if lr is None else lr
        self.momentum = cfg.SOLVER.MOMENTUM if momentum is None else momentum
        self.wd = cfg.SOLVER.WEIGHT_DECAY if weight_decay is None else weight_decay
        # frozen stem (conv1, pool1, res2): the inference engine's bf16 kernels
        self.eng = DetectionEngine(cfg, blobs, s, dtype='bf16')
        self.convs = []                                   # forward order
        mk = self._mk
        self.stages = []
        dim_in = s.dims[1]
        for si in range(1, len(s.counts)):
            dim_out, inner = s.dims[si + 1], s.dim_inner * (2 ** si)
            blocks = []
            for i in range(s.counts[si]):
                pre = 'res%d_%d' % (si + 2, i)
                st = (1, 2, 2) if i == 0 else (1, 1, 1)
                s1, s3 = (st, (1, 1, 1)) if s.stride_1x1 else ((1, 1, 1), st)
                assert s3 == (1, 1, 1), 'RpnTrainer: stride on the 1x1 conv (RESNETS.STRIDE_1X1 True, the default)'
                blk = dict(a=mk(blobs, pre + '_branch2a', affine=True, stride=s1, relu=True),
                           b=mk(blobs, pre + '_branch2b', affine=True, relu=True),
                           c=mk(blobs, pre + '_branch2c', affine=True, relu=True),
                           sc=mk(blobs, pre + '_branch1', affine=True, stride=st) if dim_in != dim_out else None)
                blocks.append(blk)
                dim_in = dim_out
            self.stages.append(blocks)
        names = s.stage_blobs[::-1]                       # coarsest first
        self.lat = [mk(blobs, 'fpn_inner_' + names[0])] + [mk(blobs, 'fpn_inner_%s_lateral' % n) for n in names[1:]]
        self.post = [mk(blobs, 'fpn_' + n) for n in names]
        k = str(s.rpn_levels[0])
        A = self.A = s.num_anchors
        self.rpn_conv = mk(blobs, 'conv_rpn_fpn' + k, relu=True)
        w = np.concatenate([blobs['rpn_cls_logits_fpn%s_w' % k], blobs['rpn_bbox_pred_fpn%s_w' % k]], 0)
        b = np.concatenate([blobs['rpn_cls_logits_fpn%s_b' % k], blobs['rpn_bbox_pred_fpn%s_b' % k]], 0)
        ld = self.rpn_ld = (5 * A + 7) // 8 * 8           # padded with zero filters so the planes are 16-byte rows
        wp = np.zeros((ld,) + w.shape[1:], np.float32); wp[:5 * A] = w
        bp = np.zeros((ld,), np.float32); bp[:5 * A] = b
        self.rpn_out = TrainConv(torch, wp, bias=bp)
        self.convs.append(self.rpn_out)
        # ---- flat gradient buffer in BACKWARD order (so a bucket finished early in the backward pass is contiguous)
        order = self.convs[::-1]
        total = sum(c.nparams() for c in order)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device='cuda')
        off = 0
        for c in order:
            n = c.w.numel()
            c.g = self.flat_g[off:off + n].view_as(c.w); off += n
            if c.bias is not None:
                c.bias_g = self.flat_g[off:off + c.bias.numel()]; off += c.bias.numel()
        bounds, self.bucket_ends = plan_buckets([c.nparams() for c in order], buckets)
        self._order_end = {id(c): e for c, e in zip(order, bounds)}
        self.reducer = BucketReducer(self.flat_g, self.bucket_ends, world)
        self.loss = torch.zeros(2, dtype=torch.float32, device='cuda')

    def _mk(self, blobs, name, affine=False, stride=(1, 1, 1), relu=False):
        c = TrainConv(self.torch, blobs[name + '_w'], scale=blobs[name + '_bn_s'] if affine else None,
                      shift=blobs[name + '_bn_b'] if affine else None, bias=None if affine else blobs[name + '_b'],
                      stride=stride, relu=relu)
        self.convs.append(c)
        return c

    # ------------------------------------------------------------------ forward (activations saved for the backward)
    def forward_all(self, frames_u8):
        """stem -> res2 (frozen) -> res3..5 -> FPN -> RPN heads; returns per level the fp32 RPN outputs [B,1,H,W,ld]."""
        torch, eng, s, cfg = self.torch, self.eng, self.spec, self.cfg
        B, T, H, W, _ = frames_u8.shape
        g = eng._geom_tensors(B, H, W)
        x = eng._blob(frames_u8, g['scale'], g['hr'], g['wr'], g['hp'], g['wp'])
        xs = x.view((B * T,) + tuple(x.shape[2:]))
        y = cv.conv1_7x7s2(xs, eng.conv1_w, (g['hp'], g['wp']), eng.conv1_s, eng.conv1_b, relu=True, dtype=cv.BF16)
        y = dense_ops.maxpool2d(y, 3, 2, 1)
        y = y.view((B, T) + tuple(y.shape[1:]))
        for blk in eng.stages[0]:
            y = eng._run_block(blk, y)
        sv = self.saved = dict(blocks=[], C=[y])                                  # C2 (frozen producer)
        for blocks in self.stages:
            for blk in blocks:
                a = blk['a'].forward(y)
                b = blk['b'].forward(a)
                sc = blk['sc'].forward(y) if blk['sc'] is not None else y
                out = blk['c'].forward(b, residual=sc, res_mode=1)
                sv['blocks'].append(dict(blk=blk, x=y, a=a, b=b, y=out))
                y = out
            sv['C'].append(y)
        Cs = sv['C'][::-1]                                                        # coarsest first
        inner = [self.lat[0].forward(Cs[0])]
        for i in range(1, len(Cs)):
            inner.append(self.lat[i].forward(Cs[i], residual=inner[i - 1], res_mode=2))
        Ps = [self.post[i].forward(inner[i]) for i in range(len(inner))]          # [B, T, h, w, 256], coarsest first
        sv['inner'], sv['P'] = inner, Ps
        c = int(cfg.VIDEO.NUM_FRAMES_MID / 2) if (s.link == 'slice-center' and T > 1) else 0
        sv['center'] = c
        feats = [p[:, c:c + 1].contiguous() if T > 1 else p for p in Ps]          # slice-center link
        p6 = dense_ops.maxpool2d(feats[0].view((B,) + tuple(feats[0].shape[2:])), 1, 2, 0)
        feats = [p6.view((B, 1) + tuple(p6.shape[1:]))] + feats                   # P6 first (coarsest)
        sv['feats'] = feats
        outs, hs = [], []
        for f in feats:
            h = self.rpn_conv.forward(f)
            o = torch.empty(tuple(h.shape[:4]) + (self.rpn_ld,), dtype=torch.float32, device='cuda')
            self.rpn_out.forward(h, out_f32=True, out=o)
            hs.append(h); outs.append(o)
        sv['rpn_h'] = hs
        return outs[::-1]                                                          # finest (P2) first, like spec.rpn_levels

    # ------------------------------------------------------------------ backward
    def _bucket_ready(self, conv, pending=None):
        """Launch the all-reduce of every bucket whose last filter gradient has just been enqueued."""
        self.reducer.ready(self._order_end[id(conv)])

    def backward(self, rpn_outs, targets):
        """targets: per level (finest first) dict(labels [B,H,W,A] i32, bbox_targets / inside / outside [B,H,W,4A] f32)."""
        torch, cfg, s, sv = self.torch, self.cfg, self.spec, self.saved
        A = self.A
        L.call('dt_memset', L.ptr(self.flat_g), 0, self.flat_g.numel() * 4, L.stream_ptr())
        L.call('dt_memset', L.ptr(self.loss), 0, 8, L.stream_ptr())
        self.reducer.reset()
        pending = None
        B = rpn_outs[0].shape[0]
        s_cls = 1.0 / self.world / cfg.TRAIN.RPN_BATCH_SIZE_PER_IM / cfg.TRAIN.IMS_PER_BATCH
        s_box = 1.0 / self.world / B
        feats, hs = sv['feats'], sv['rpn_h']                                    # coarsest first
        nl = len(feats)
        gP6 = None
        gfeat = [None] * nl
        # RPN heads, level by level (shared filters: their gradients accumulate across levels)
        for li in range(nl):
            o = rpn_outs[nl - 1 - li]                                           # level of feats[li]
            t = targets[nl - 1 - li]
            go = torch.empty(tuple(o.shape[:4]) + (self.rpn_ld,), dtype=torch.bfloat16, device='cuda')
            rows = o.numel() // o.shape[-1]
            L.call('dt_rpn_loss_grad', L.ptr(o), o.shape[-1], L.ptr(t['labels']), L.ptr(t['bbox_targets']), L.ptr(t['inside']),
                   L.ptr(t['outside']), rows, A, s_cls, s_box, 1.0 / 9.0, L.ptr(go), self.rpn_ld, L.ptr(self.loss), L.stream_ptr())
            gh, _ = self.rpn_out.backward(go, hs[li])
            gz = to.bwd_pointwise(gh, None, hs[li], None)                      # Relu of conv_rpn (bias conv: no scale)
            gfeat[li], _ = self.rpn_conv.backward(gz, feats[li])
        self._bucket_ready(self.rpn_conv, pending)
        # P6 = stride-2 subsample of P5's centre frame: its gradient lands on P5's even positions
        Ps, inner = sv['P'], sv['inner']                                         # coarsest first (P5 ... P2)
        T = Ps[0].shape[1]
        c = sv['center']
        gP = []
        for i in range(len(Ps)):
            g = gfeat[i + 1]                                                     # [B,1,h,w,256]
            if i == 0:
                g = to.bwd_pointwise(g, to.scatter_stride2(gfeat[0], (g.shape[2], g.shape[3])))
            if T > 1:                                                            # slice-center: the other frames get zero
                full = L.zeros(tuple(Ps[i].shape), torch.bfloat16)
                full[:, c:c + 1].copy_(g)
                g = full
            gP.append(g)
        # FPN: finest level first (its inner gradient flows into the next coarser one through the top-down add)
        g_inner = [None] * len(Ps)
        for i in range(len(Ps) - 1, -1, -1):
            gi, _ = self.post[i].backward(gP[i], inner[i])
            if i < len(Ps) - 1:                                                  # add the 2x2 sums of the finer level's inner gradient
                B_, T_, h, w, C_ = gi.shape
                gi = to.upsample_add_bwd(g_inner[i + 1], gi)
            g_inner[i] = gi
        Cs = sv['C'][::-1]
        gC = []
        for i in range(len(Ps)):
            need_dx = i < len(Ps) - 1                                            # C2's producer is frozen
            dx, _ = self.lat[i].backward(g_inner[i], Cs[i], need_dx=need_dx)
            gC.append(dx)
        self._bucket_ready(self.lat[0], pending)
        # body: res5 -> res3
        blocks = sv['blocks']
        bi = len(blocks)
        g_parts = (gC[0], None)                                                  # gradient wrt C5 from its lateral
        stage_of = []
        for si, st_blocks in enumerate(self.stages):
            stage_of += [si] * len(st_blocks)
        for bidx in range(len(blocks) - 1, -1, -1):
            sb = blocks[bidx]
            blk = sb['blk']
            g1, g2 = g_parts
            last_of_trunk = bidx == 0
            gz3 = to.bwd_pointwise(g1, g2, sb['y'], blk['c'].scale)
            gb, _ = blk['c'].backward(gz3, sb['b'])
            gz2 = to.bwd_pointwise(gb, None, sb['b'], blk['b'].scale)
            ga, _ = blk['b'].backward(gz2, sb['a'])
            gz1 = to.bwd_pointwise(ga, None, sb['a'], blk['a'].scale)
            need_dx = not last_of_trunk                                          # res2 is frozen: stop at the input of res3_0
            gx1, xpl = blk['a'].backward(gz1, sb['x'], need_dx=need_dx)
            if blk['sc'] is not None:
                gz4 = to.bwd_pointwise(g1, g2, sb['y'], blk['sc'].scale)
                gx2, _ = blk['sc'].backward(gz4, sb['x'], x_planes=xpl, need_dx=need_dx)
            else:
                gx2 = to.bwd_pointwise(g1, g2, sb['y'], None)
            for cnv in (blk['c'], blk['b'], blk['a'], blk['sc']):
                if cnv is not None:
                    self._bucket_ready(cnv, pending)
            g_parts = (gx1, gx2)
            # a stage boundary: the lateral's gradient wrt this stage output joins the two parts
            if bidx > 0 and stage_of[bidx - 1] != stage_of[bidx]:
                lvl = len(self.stages) - stage_of[bidx - 1] - 1                 # index into gC (coarsest first)
                g_parts = (to.bwd_pointwise(gx1, gx2), gC[lvl])
        self.reducer.wait()
        return self.loss

    def update(self):
        gs = 1.0            # losses already carry 1/NUM_GPUS; the all-reduce is a SUM (model_builder.py:484,938-942)
        for c in self.convs:
            c.update(self.lr, self.momentum, self.wd, gs)

    def step(self, frames_u8, targets):
        outs = self.forward_all(frames_u8)
        loss = self.backward(outs, targets)
        self.update()
        return loss

    # ------------------------------------------------------------------ synthetic targets (the f1 loaders are out of scope)
    def synthetic_targets(self, B, H, W, seed=0):
        """Per level random RPN targets of the reference's shapes (lib/roi_data/rpn.py): 256 sampled anchors per image
        (half foreground) spread over the levels, unit inside weights on foreground, outside weights 1 / 256."""
        torch = self.torch
        eng = self.eng
        g = eng._geom_tensors(B, H, W)
        rng = np.random.RandomState(seed)
        A = self.A
        out = []
        for lvl in self.spec.rpn_levels:
            h, w = int(np.ceil(g['hp'] / 2. ** lvl)), int(np.ceil(g['wp'] / 2. ** lvl))
            labels = -np.ones((B, h, w, A), np.int32)
            n = max(2, int(256 * (h * w) / float(sum(int(np.ceil(g['hp'] / 2. ** l)) * int(np.ceil(g['wp'] / 2. ** l)) for l in self.spec.rpn_levels))))
            bt = np.zeros((B, h, w, 4 * A), np.float32); iw = np.zeros_like(bt); ow = np.zeros_like(bt)
            for b in range(B):
                idx = rng.choice(h * w * A, size=min(n, h * w * A), replace=False)
                fg = idx[:len(idx) // 2]
                lab = labels[b].reshape(-1)
                lab[idx] = 0; lab[fg] = 1
                pos, a = fg // A, fg % A
                for k in range(4):
                    bt[b].reshape(-1, 4 * A)[pos, a * 4 + k] = rng.normal(0, 0.5, len(fg))
                    iw[b].reshape(-1, 4 * A)[pos, a * 4 + k] = 1.0
                pos_all, a_all = idx // A, idx % A
                for k in range(4):
                    ow[b].reshape(-1, 4 * A)[pos_all, a_all * 4 + k] = 1.0 / 256.0
            out.append(dict(labels=torch.from_numpy(labels).cuda(), bbox_targets=torch.from_numpy(bt).cuda(),
                            inside=torch.from_numpy(iw).cuda(), outside=torch.from_numpy(ow).cuda()))
        return out
