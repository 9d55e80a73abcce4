"""CPU interpreter of a launch plan -- TEST INFRASTRUCTURE ONLY.

Executes the ops of hcflow_b200.plan with torch on the CPU, using the same packed weights
(hcflow_b200.prep) and the same NHWC channel-slice views the CUDA engine uses.  It exists so
that the host logic (plan construction, weight packing, view/segment bookkeeping, log-det
constants) can be checked against the oracle in the GPU-less authoring container.  It is
not importable from the package and is never a product path.
"""
import math

import torch
import torch.nn.functional as F

from hcflow_b200 import plan as P
from hcflow_b200 import prep
from hcflow_b200 import rewrite


class Emulator:
    def __init__(self, net, plan, dtype=torch.float32, ops=None, extra_bufs=None):
        """ops / extra_bufs: an engine-level rewrite of plan.ops (hcflow_b200.rewrite) to interpret instead."""
        self.plan = plan
        self.ops = list(plan.ops) if ops is None else ops
        self.dtype = dtype
        self.sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        allb = dict(plan.bufs)
        allb.update(extra_bufs or {})
        self.bufs = {n: torch.zeros(plan.B, b.H, b.W, b.C, dtype=dtype) for n, b in allb.items()}
        self.ext = {}
        self.logdet = torch.zeros(plan.B, dtype=torch.float64)
        self.quant = getattr(net, "quant", 256)

    def view(self, v):
        return self.bufs[v.buf.name][..., v.off:v.off + v.C]

    def _conv(self, op):
        npad = prep.npad_for(op.cout)
        segc = [v.C for v, _ in op.segs]
        wp = prep.pack_conv_weight(rewrite.raw_weight(self.sd, op), segc, npad).to(self.dtype)  # [taps, kpad, npad]
        ins = []
        for (v, up), c in zip(op.segs, segc):
            x = self.view(v).permute(0, 3, 1, 2)
            if up:
                x = x.repeat_interleave(1 << up, dim=2).repeat_interleave(1 << up, dim=3)
            assert x.shape[2] == op.H and x.shape[3] == op.W, (x.shape, op.H, op.W)
            pad = prep.seg_pad(c) - c
            if pad:
                x = F.pad(x, (0, 0, 0, 0, 0, pad))
            ins.append(x)
        x = torch.cat(ins, 1)
        ks = op.ks
        w = wp.permute(2, 1, 0).reshape(npad, x.shape[1], ks, ks)
        y = F.conv2d(x, w, None, padding=ks // 2)
        if op.pre is not None:   # pre-activation addend (hcf_conv_args.pre)
            y[:, :op.cout] = y[:, :op.cout] + self.view(op.pre).permute(0, 3, 1, 2)
        if op.bias:
            y = y + prep.pad_vec(prep.derive(self.sd, op.bias), npad, 0.0).to(self.dtype).view(1, -1, 1, 1)
        if op.scale:
            y = y * prep.pad_vec(prep.derive(self.sd, op.scale), npad, 1.0).to(self.dtype).view(1, -1, 1, 1)
        if op.act == P.ACT_RELU:
            y = F.relu(y)
        elif op.act == P.ACT_LRELU:
            y = F.leaky_relu(y, 0.2)
        if op.raw2 is not None:   # columns [32, 64) leave raw (before bias / activation): recompute them without
            raw = F.conv2d(x, w, None, padding=ks // 2)[:, 32:64].permute(0, 2, 3, 1)
            self.view(op.raw2).copy_(raw)
            y = y[:, :32].permute(0, 2, 3, 1).clone()
            self.view(op.out).copy_(y)
            return
        y = y[:, :op.cout].permute(0, 2, 3, 1)
        if op.res1 is not None:
            y = y * op.alpha1 + self.view(op.res1)
        if op.res2 is not None:
            y = y * op.alpha2 + self.view(op.res2)
        y = y.clone()
        self.view(op.out).copy_(y)
        if op.out2 is not None:
            self.view(op.out2).copy_(y)
        if op.step is not None:   # fused FlowStep tail (hcf_conv_step): h is consumed right away
            self._step(op.step)

    def _step(self, op):
        z = self.view(op.z)
        C = op.z.C
        if op.variant == "forward_head":
            sc = prep.derive(self.sd, op.an_scale).to(self.dtype)
            bs = prep.derive(self.sd, op.an_bias).to(self.dtype)
            y = (z + bs) * sc
            if op.w:
                y = y @ prep.derive(self.sd, op.w).to(self.dtype).t()
            z.copy_(y)
            return
        h = self.view(op.h) if op.h is not None else None
        y = z.clone()
        if op.mode == "affine":
            shift, scale = h[..., 0::2], h[..., 1::2]
            ls = 0.318 * torch.atan(2 * scale)
            if op.variant == "inverse":
                y[..., op.n_pass:] = y[..., op.n_pass:] * torch.exp(-ls) - shift
            else:
                y[..., op.n_pass:] = (y[..., op.n_pass:] + shift) * torch.exp(ls)
                if self.plan.uses_logdet:
                    self.logdet += ls.double().sum(dim=(1, 2, 3))
        else:
            y[..., :3] = y[..., :3] - h if op.variant == "inverse" else y[..., :3] + h
        if op.variant == "inverse":
            if op.w:
                y = y @ prep.derive(self.sd, op.w).to(self.dtype).t()
            y = y * prep.derive(self.sd, op.an_scale).to(self.dtype) - prep.derive(self.sd, op.an_bias).to(self.dtype)
        z.copy_(y)

    def _prior(self, op):
        h = self.view(op.h)
        mean, logs = h[..., 0::2], h[..., 1::2]
        if op.atan_logscale:
            logs = 0.318 * torch.atan(2 * logs)
        z = self.view(op.z)
        if op.variant == "sample":
            eps = self.ext["eps{}".format(op.eps_index)].to(self.dtype).permute(0, 2, 3, 1)
            z.copy_(mean + torch.exp(logs) * eps)
        elif op.variant == "logp":
            ll = -0.5 * (logs * 2.0 + (z - mean) ** 2 / torch.exp(logs * 2.0) + math.log(2 * math.pi))
            self.logdet += ll.double().sum(dim=(1, 2, 3))
        else:
            self.ext[op.out_name] = ((z - mean) * torch.exp(-logs)).permute(0, 3, 1, 2).contiguous()

    def _layout(self, op):
        if op.variant == "ingest":
            x = self.ext[op.src].to(self.dtype)
            if op.noise:
                x = x + self.ext[op.noise].to(self.dtype) * op.noise_scale
            self.view(op.dst).copy_(x.permute(0, 2, 3, 1))
        elif op.variant == "egress":
            x = self.view(op.src).permute(0, 3, 1, 2).contiguous()
            if op.post >= 1:
                x = torch.clamp(x, 0, 1)
            if op.post == 2:
                x = (x * 255.0).round() / 255.0
            self.ext[op.dst] = x
        else:
            src, dst = self.view(op.src), self.view(op.dst)
            C = op.C
            if op.variant == "copy":
                dst.copy_(src.clone())
            elif op.variant == "upsample":   # nearest, factor 2^post (hcf_upsample_nearest)
                f = 1 << op.post
                dst.copy_(src.repeat_interleave(f, dim=1).repeat_interleave(f, dim=2))
            elif op.variant == "squeeze":   # src hi-res C -> dst low-res 4C
                B, H2, W2, _ = src.shape
                x = src.reshape(B, H2 // 2, 2, W2 // 2, 2, C).permute(0, 1, 3, 5, 2, 4)  # b,y,x,c,i,j
                dst.copy_(x.reshape(B, H2 // 2, W2 // 2, 4 * C))
            elif op.variant == "unsqueeze":
                B, H, W, _ = src.shape
                x = src[..., :4 * C].reshape(B, H, W, C, 2, 2).permute(0, 1, 4, 2, 5, 3)  # b,y,i,x,j,c
                dst.copy_(x.reshape(B, 2 * H, 2 * W, C))
            elif op.variant == "haar_fwd":
                a, b = src[:, 0::2, 0::2], src[:, 0::2, 1::2]
                c, d = src[:, 1::2, 0::2], src[:, 1::2, 1::2]
                dst.copy_(torch.cat([(a + b + c + d) / 4, (a - b + c - d) / 4, (a + b - c - d) / 4,
                                     (a - b - c + d) / 4], -1))
            elif op.variant == "haar_inv":
                k0, k1, k2, k3 = [src[..., i * C:(i + 1) * C] for i in range(4)]
                dst[:, 0::2, 0::2] = k0 + k1 + k2 + k3
                dst[:, 0::2, 1::2] = k0 - k1 + k2 - k3
                dst[:, 1::2, 0::2] = k0 + k1 - k2 - k3
                dst[:, 1::2, 1::2] = k0 - k1 - k2 + k3
            else:
                raise ValueError(op.variant)

    def run(self, **ext):
        self.ext.update(ext)
        const = prep.logdet_constant(self.sd, self.plan.logdet_terms)
        if self.plan.direction == "forward" and self.plan.sr:
            s = 2 ** (len(self.plan.noise_shapes) or 0)
            hr = self.ext["hr"]
            const += prep.quant_logdet(self.quant, hr.shape[2] * hr.shape[3])
        self.logdet.fill_(const)
        flat = []
        for op in self.ops:     # a fused FlowStep chain computes exactly what the ops it replaces compute
            flat.extend(op.orig if isinstance(op, P.FlowChainOp) else [op])
        for op in flat:
            if isinstance(op, P.ConvOp):
                self._conv(op)
            elif isinstance(op, P.StepOp):
                self._step(op)
            elif isinstance(op, P.PriorOp):
                self._prior(op)
            elif isinstance(op, P.LayoutOp):
                self._layout(op)
            elif isinstance(op, P.DiracLogpOp):
                x, m = self.ext[op.x_name].double(), self.ext[op.mean_name].double()
                ll = -0.5 * (op.logs * 2.0 + (x - m) ** 2 / math.exp(op.logs * 2.0) + math.log(2 * math.pi))
                self.logdet += ll.sum(dim=(1, 2, 3))
        return self.ext
