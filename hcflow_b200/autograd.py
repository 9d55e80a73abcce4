"""Training path: the hot-path ops as ``torch.autograd.Function`` extensions.

The reference trains through torch autograd over its Python modules (HCFlow_SR_model.py:184-218
``optimize_parameters``: ``_, nll = netG(hr=..., lr=..., reverse=False); nll.backward()``).  Here every op of that
graph is a Function whose forward AND backward are CUDA kernels of libhcflow_b200.so (csrc/grad_ops.cu, plus the
engine's own fp32 conv / squeeze / up-sample kernels): exact fp32 arithmetic over NHWC tensors, no torch math on
activations -- torch only moves data (channel slices / concatenations), keeps the autograd tape and does the O(C^2)
parameter-side scalar work the reference also does with torch (exp(logs), slogdet(W)).

``sr_forward_nll(net, hr, lr, dequant_noise)`` builds the same graph as HCFlowNet_SR.normal_flow_diracLR
(HCFlowNet_SR_arch.py:47-67 -> FlowNet_SR_x4.py:84-101 / _x8.py:91-118) and returns ``(clamp(fake_lr), nll)`` with
gradients to every parameter and to ``hr``.  The fused tensor-core engine stays the inference path; arch.py routes a
call here when autograd is recording.
"""
import ctypes as C
import math

import torch
from torch.autograd import Function

from . import _lib as L
from . import modules as M
from . import prep

ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2


def _st(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _chk(t):
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), (t.device, t.dtype, t.is_contiguous())
    return t


# ------------------------------------------------------------------------------------------------ raw kernel calls
def _conv_raw(x, w, ks):
    """y[B,H,W,Cout] = conv_ks(x[B,H,W,Cin], w[Cout,Cin,ks,ks]), zero padding, no bias (hcf_conv_fp32)."""
    lib = L.load()
    _chk(x)
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    npad = prep.npad_for(Cout)
    wp = prep.pack_conv_weight(w.detach(), [Cin], npad)
    y = torch.empty(B, H, W, Cout, dtype=torch.float32, device=x.device)
    a = L.ConvArgs()
    a.B, a.H, a.W, a.nseg = B, H, W, 1
    a.seg[0].ptr, a.seg[0].ld, a.seg[0].C, a.seg[0].up_shift = x.data_ptr(), Cin, Cin, 0
    a.ks, a.kpad, a.cout, a.npad = ks, prep.seg_pad(Cin), Cout, npad
    a.w = wp.data_ptr()
    a.act = ACT_NONE
    a.out, a.out_ld = y.data_ptr(), Cout
    with torch.cuda.device(x.device):
        L.check(lib.hcf_conv_fp32(C.byref(a), _st(x)), "conv_fp32")
    return y


class Conv2dFn(Function):
    """F.conv2d(x, w, padding=ks//2) on NHWC (Basic.py:14-72, 360-383; Permutations.py:100 for the 1x1 mix)."""

    @staticmethod
    def forward(ctx, x, w):
        ks = w.shape[2]
        x = x.contiguous()
        ctx.save_for_backward(x, w)
        ctx.ks = ks
        return _conv_raw(x, w, ks)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        lib = L.load()
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # dx = conv(dy, w flipped in space, transposed in channels)
            dx = _conv_raw(dy, w.detach().flip(2, 3).transpose(0, 1).contiguous(), ctx.ks)
        if ctx.needs_input_grad[1]:
            B, H, W, Cin = x.shape
            Cout = w.shape[0]
            dw = torch.empty(w.shape, dtype=torch.float32, device=w.device)   # (contiguous: torch.inverse returns column-major)
            with torch.cuda.device(x.device):
                L.check(lib.hcf_conv_wgrad(x.data_ptr(), Cin, dy.data_ptr(), Cout, B, H, W, Cin, Cout, ctx.ks, dw.data_ptr(),
                                           _st(x)), "conv_wgrad")
        return dx, dw


class AffineActFn(Function):
    """y = act((x + bias[c]) * scale[c]): ActNorm forward (ActNorms.py:66-69), conv bias + LeakyReLU / ReLU
    (Basic.py:349-355, 377-381, 442-446), Conv2dZeros' (. + bias) * exp(3 logs) (Basic.py:70-72)."""

    @staticmethod
    def forward(ctx, x, bias, scale, act):
        lib = L.load()
        x = _chk(x.contiguous())
        Cc = x.shape[-1]
        npix = x.numel() // Cc
        y = torch.empty_like(x)
        b = None if bias is None else bias.detach().reshape(-1).contiguous()
        s = None if scale is None else scale.detach().reshape(-1).contiguous()
        with torch.cuda.device(x.device):
            L.check(lib.hcf_affine_act_fwd(x.data_ptr(), None if b is None else b.data_ptr(), None if s is None else s.data_ptr(),
                                           act, y.data_ptr(), npix, Cc, _st(x)), "affine_act_fwd")
        ctx.save_for_backward(x, b if b is not None else x.new_empty(0), s if s is not None else x.new_empty(0))
        ctx.act, ctx.has = act, (bias is not None, scale is not None)
        ctx.shapes = (None if bias is None else bias.shape, None if scale is None else scale.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        x, b, s = ctx.saved_tensors
        hb, hs = ctx.has
        dy = dy.contiguous()
        Cc = x.shape[-1]
        npix = x.numel() // Cc
        dx = torch.empty_like(x)
        db = torch.empty(Cc, dtype=torch.float32, device=x.device) if hb else None
        ds = torch.empty(Cc, dtype=torch.float32, device=x.device) if hs else None
        with torch.cuda.device(x.device):
            L.check(lib.hcf_affine_act_bwd(dy.data_ptr(), x.data_ptr(), b.data_ptr() if hb else None, s.data_ptr() if hs else None,
                                           ctx.act, dx.data_ptr(), db.data_ptr() if hb else None, ds.data_ptr() if hs else None,
                                           npix, Cc, _st(x)), "affine_act_bwd")
        return (dx, db.reshape(ctx.shapes[0]) if hb else None, ds.reshape(ctx.shapes[1]) if hs else None, None)


class CouplingFn(Function):
    """Affine coupling on the coupled channels (AffineCouplings.py:52-61 forward, :78-85 reverse):
    (z2 [npix, nc], h [npix, 2 nc]) -> (z2', per-image sum of log-scales as fp64 [B])."""

    @staticmethod
    def forward(ctx, z2, h, pix_per_img, inverse):
        lib = L.load()
        z2, h = _chk(z2.contiguous()), _chk(h.contiguous())
        nc = z2.shape[-1]
        npix = z2.numel() // nc
        out = torch.empty_like(z2)
        lsum = torch.zeros(npix // pix_per_img, dtype=torch.float64, device=z2.device)
        with torch.cuda.device(z2.device):
            L.check(lib.hcf_coupling_fwd(z2.data_ptr(), h.data_ptr(), nc, int(inverse), out.data_ptr(), lsum.data_ptr(), npix,
                                         pix_per_img, _st(z2)), "coupling_fwd")
        ctx.save_for_backward(z2, h)
        ctx.cfg = (pix_per_img, int(inverse))
        return out, lsum

    @staticmethod
    def backward(ctx, dout, dlsum):
        lib = L.load()
        z2, h = ctx.saved_tensors
        pix_per_img, inverse = ctx.cfg
        nc = z2.shape[-1]
        npix = z2.numel() // nc
        dout = dout.contiguous()
        gl = None if (dlsum is None or inverse) else dlsum.to(torch.float32).contiguous()
        dz2, dh = torch.empty_like(z2), torch.empty_like(h)
        with torch.cuda.device(z2.device):
            L.check(lib.hcf_coupling_bwd(dout.data_ptr(), None if gl is None else gl.data_ptr(), z2.data_ptr(), h.data_ptr(), nc,
                                         inverse, dz2.data_ptr(), dh.data_ptr(), npix, pix_per_img, _st(z2)), "coupling_bwd")
        return dz2, dh, None, None


class GaussLogpFn(Function):
    """GaussianDiag.logp (Basic.py:79-93) summed per image -> fp64 [B]; ``logs`` a tensor or a python constant."""

    @staticmethod
    def forward(ctx, x, mean, logs, logs_const):
        lib = L.load()
        x, mean = _chk(x.contiguous()), _chk(mean.contiguous())
        lg = None if logs is None else _chk(logs.contiguous())
        B = x.shape[0]
        per = x.numel() // B
        out = torch.zeros(B, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            L.check(lib.hcf_gauss_logp_fwd(x.data_ptr(), mean.data_ptr(), None if lg is None else lg.data_ptr(), float(logs_const),
                                           B, per, out.data_ptr(), _st(x)), "gauss_logp_fwd")
        ctx.save_for_backward(x, mean, lg if lg is not None else x.new_empty(0))
        ctx.cfg = (lg is not None, float(logs_const))
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        x, mean, lg = ctx.saved_tensors
        has_l, lc = ctx.cfg
        B = x.shape[0]
        per = x.numel() // B
        g32 = g.to(torch.float32).contiguous()
        need = ctx.needs_input_grad
        dx = torch.empty_like(x) if need[0] else None
        dm = torch.empty_like(x) if need[1] else None
        dl = torch.empty_like(x) if (has_l and need[2]) else None
        with torch.cuda.device(x.device):
            L.check(lib.hcf_gauss_logp_bwd(g32.data_ptr(), x.data_ptr(), mean.data_ptr(), lg.data_ptr() if has_l else None, lc, B, per,
                                           None if dx is None else dx.data_ptr(), None if dm is None else dm.data_ptr(),
                                           None if dl is None else dl.data_ptr(), _st(x)), "gauss_logp_bwd")
        return dx, dm, dl, None


class AxpbyFn(Function):
    """y = alpha a + beta b: the residual scale-adds of the RRDB encoder (Basic.py:383, 398), trunk skip, dequantisation."""

    @staticmethod
    def forward(ctx, a, alpha, b, beta):
        lib = L.load()
        a, b = _chk(a.contiguous()), _chk(b.contiguous())
        y = torch.empty_like(a)
        with torch.cuda.device(a.device):
            L.check(lib.hcf_axpby(a.data_ptr(), float(alpha), b.data_ptr(), float(beta), y.data_ptr(), a.numel(), _st(a)), "axpby")
        ctx.ab = (float(alpha), float(beta))
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        alpha, beta = ctx.ab
        dy = dy.contiguous()
        outs = []
        for k, need in ((alpha, ctx.needs_input_grad[0]), (beta, ctx.needs_input_grad[2])):
            if not need:
                outs.append(None)
                continue
            g = torch.empty_like(dy)
            with torch.cuda.device(dy.device):
                L.check(lib.hcf_axpby(dy.data_ptr(), k, None, 0.0, g.data_ptr(), dy.numel(), _st(dy)), "axpby")
            outs.append(g)
        return outs[0], None, outs[1], None


class Quant8Fn(Function):
    """Quant (Basic.py:186-198): clamp to [0,1], round to 8 bit; straight-through backward."""

    @staticmethod
    def forward(ctx, x):
        lib = L.load()
        x = _chk(x.contiguous())
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.check(lib.hcf_quantize8(x.data_ptr(), y.data_ptr(), x.numel(), _st(x)), "quantize8")
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy


def _sq_args(src, dst, B, Cc, H, W):
    a = L.SqueezeArgs()
    a.B, a.C, a.H, a.W = B, Cc, H, W
    a.src, a.src_ld = src.data_ptr(), src.shape[-1]
    a.dst, a.dst_ld = dst.data_ptr(), dst.shape[-1]
    return a


class SqueezeFn(Function):
    """squeeze2d (Basic.py:127-140) on NHWC: [B,2H,2W,C] -> [B,H,W,4C]; backward = unsqueeze2d (:143-157)."""

    @staticmethod
    def forward(ctx, x):
        lib = L.load()
        x = _chk(x.contiguous())
        B, H2, W2, Cc = x.shape
        y = torch.empty(B, H2 // 2, W2 // 2, 4 * Cc, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            L.check(lib.hcf_squeeze2d(C.byref(_sq_args(x, y, B, Cc, H2 // 2, W2 // 2)), _st(x)), "squeeze2d")
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        dy = dy.contiguous()
        B, H, W, C4 = dy.shape
        dx = torch.empty(B, 2 * H, 2 * W, C4 // 4, dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            L.check(lib.hcf_unsqueeze2d(C.byref(_sq_args(dy, dx, B, C4 // 4, H, W)), _st(dy)), "unsqueeze2d")
        return dx


class UpsampleFn(Function):
    """F.interpolate(scale_factor=2^shift, mode='nearest') (FlowNet_SR_x4.py:98, _x8.py:109-113); backward = block sums."""

    @staticmethod
    def forward(ctx, x, shift):
        lib = L.load()
        x = _chk(x.contiguous())
        B, H, W, Cc = x.shape
        y = torch.empty(B, H << shift, W << shift, Cc, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            L.check(lib.hcf_upsample_nearest(C.byref(_sq_args(x, y, B, Cc, H << shift, W << shift)), shift, _st(x)), "upsample")
        ctx.shift = shift
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        dy = dy.contiguous()
        s = ctx.shift
        B, HH, WW, Cc = dy.shape
        dx = torch.empty(B, HH >> s, WW >> s, Cc, dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            L.check(lib.hcf_downsample_sum(dy.data_ptr(), dx.data_ptr(), B, HH >> s, WW >> s, Cc, s, _st(dy)), "downsample_sum")
        return dx, None


class UnsqueezeFn(Function):
    """unsqueeze2d (Basic.py:143-157): [B,H,W,4C] -> [B,2H,2W,C]; backward = squeeze2d."""

    @staticmethod
    def forward(ctx, x):
        lib = L.load()
        x = _chk(x.contiguous())
        B, H, W, C4 = x.shape
        y = torch.empty(B, 2 * H, 2 * W, C4 // 4, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            L.check(lib.hcf_unsqueeze2d(C.byref(_sq_args(x, y, B, C4 // 4, H, W)), _st(x)), "unsqueeze2d")
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        dy = dy.contiguous()
        B, H2, W2, Cc = dy.shape
        dx = torch.empty(B, H2 // 2, W2 // 2, 4 * Cc, dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            L.check(lib.hcf_squeeze2d(C.byref(_sq_args(dy, dx, B, Cc, H2 // 2, W2 // 2)), _st(dy)), "squeeze2d")
        return dx


class GaussSampleFn(Function):
    """GaussianDiag.sample with the noise given (Basic.py:96-100): mean + exp(logs) * eps."""

    @staticmethod
    def forward(ctx, mean, logs, eps):
        lib = L.load()
        mean, logs, eps = _chk(mean.contiguous()), _chk(logs.contiguous()), _chk(eps.contiguous())
        out = torch.empty_like(mean)
        with torch.cuda.device(mean.device):
            L.check(lib.hcf_gauss_sample(mean.data_ptr(), logs.data_ptr(), eps.data_ptr(), None, out.data_ptr(), None, mean.numel(),
                                         _st(mean)), "gauss_sample")
        ctx.save_for_backward(logs, eps)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        logs, eps = ctx.saved_tensors
        g = g.contiguous()
        dlogs = torch.empty_like(logs)
        with torch.cuda.device(g.device):
            L.check(lib.hcf_gauss_sample(None, logs.data_ptr(), eps.data_ptr(), g.data_ptr(), None, dlogs.data_ptr(), g.numel(),
                                         _st(g)), "gauss_sample")
        return g, dlogs, None


# ------------------------------------------------------------------------------------------------ module-level graph
def _conv(x, mod_w, bias=None, scale=None, act=ACT_NONE):
    y = Conv2dFn.apply(x, mod_w)
    if bias is not None or scale is not None or act != ACT_NONE:
        y = AffineActFn.apply(y, bias, scale, act)
    return y


def _cat(ts):
    return torch.cat(ts, dim=-1)        # channel concat on NHWC: data movement only


def _actnorm_data_init(an, x):
    """ActNorms.py:29-43 + :79-80, on an NHWC input: a not-yet-initialised ActNorm in TRAINING mode whose bias is still
    all zero takes bias = -mean and logs = log(scale / (std + 1e-6)) of its input over (batch, pixels); a non-zero bias
    counts as initialised (loaded weights); eval mode never initialises.  The reference's training loop clears `inited`
    on every step below act_norm_start_step (HCFlow_SR_model.py:184-187), so on a net trained from scratch exactly the
    first step initialises.  Parameter-side work on a few hundred channels: torch ops, like the reference."""
    if an.inited or not an.training:
        return
    if bool((an.bias != 0).any()):
        an.inited = True
        return
    with torch.no_grad():
        xf = x.detach().to(torch.float32)
        bias = -xf.mean(dim=(0, 1, 2))
        var = ((xf + bias) ** 2).mean(dim=(0, 1, 2))
        logs = torch.log(an.scale / (torch.sqrt(var) + 1e-6))
        an.bias.data.copy_(bias.view_as(an.bias))
        an.logs.data.copy_(logs.view_as(an.logs))
        an.inited = True


def _actnorm_conv(x, conv, act):
    """Basic.py:49-53: bias-free conv, then its ActNorm (data-initialised on the conv's output when due), then act."""
    y = Conv2dFn.apply(x, conv.weight)
    _actnorm_data_init(conv.actnorm, y)
    return AffineActFn.apply(y, conv.actnorm.bias, torch.exp(conv.actnorm.logs), act)


def _fcn(x, f):
    """Basic.py:442-447 (ActNorm inside Conv2d: Basic.py:49-53)."""
    h = _actnorm_conv(x, f.conv1, ACT_RELU)
    h = _actnorm_conv(h, f.conv2, ACT_RELU)
    return _conv(h, f.conv3.weight, f.conv3.bias, torch.exp(f.conv3.logs * 3.0), ACT_NONE)


def _dense5(x, blk):
    x1 = _conv(x, blk.conv1.weight, blk.conv1.bias, None, ACT_LRELU)
    x2 = _conv(_cat((x, x1)), blk.conv2.weight, blk.conv2.bias, None, ACT_LRELU)
    x3 = _conv(_cat((x, x1, x2)), blk.conv3.weight, blk.conv3.bias, None, ACT_LRELU)
    x4 = _conv(_cat((x, x1, x2, x3)), blk.conv4.weight, blk.conv4.bias, None, ACT_LRELU)
    return _conv(_cat((x, x1, x2, x3, x4)), blk.conv5.weight, blk.conv5.bias, None, ACT_NONE)


def _rrdb(x, m):
    """Basic.py:377-398."""
    out = x
    for rdb in (m.RDB1, m.RDB2, m.RDB3):
        out = AxpbyFn.apply(_dense5(out, rdb), 0.2, out, 1.0)
    return AxpbyFn.apply(out, 0.2, x, 1.0)


def _cond_feature(u, cf):
    """ConditionalFlow.py:99-110."""
    first = _conv(u, cf.conv_first.weight, cf.conv_first.bias, None, ACT_NONE)
    x = first
    for m in cf.RRDB_trunk0:
        x = _rrdb(x, m)
    f1 = x
    for m in cf.RRDB_trunk1:
        x = _rrdb(x, m)
    f2 = AxpbyFn.apply(_conv(x, cf.trunk_conv1.weight, cf.trunk_conv1.bias, None, ACT_NONE), 1.0, first, 1.0)
    return _cat((f1, f2)) if cf.SR else f2


def _flow_step_forward(z, u, step, logdet):
    """FlowStep.normal_flow (FlowStep.py:40-51): ActNorm -> invconv -> affine coupling; logdet fp64 [B]."""
    B, H, W, Cc = z.shape
    pixels = H * W
    _actnorm_data_init(step.actnorm, z)
    z = AffineActFn.apply(z, step.actnorm.bias, torch.exp(step.actnorm.logs), ACT_NONE)
    logdet = logdet + step.actnorm.logs.sum().double() * pixels
    if step.permute is not None:
        Wm = step.permute.weight
        z = Conv2dFn.apply(z, Wm.view(Cc, Cc, 1, 1))
        logdet = logdet + torch.slogdet(Wm)[1].double() * pixels      # (the reference does this on the CPU: Permutations.py:70)
    aff = step.affine
    assert aff.mode == "affine", "training path: AffineCoupling only (the SR nets)"
    n = aff.n_pass
    z1, z2 = z[..., :n].contiguous(), z[..., n:].contiguous()
    h = _fcn(z1 if u is None else _cat((z1, u)), aff.f)
    z2, lsum = CouplingFn.apply(z2, h, pixels, False)
    return _cat((z1, z2)), logdet + lsum


def sr_forward_nll(net, hr, lr, dequant_noise=None):
    """HCFlowNet_SR.normal_flow_diracLR with autograd (HCFlowNet_SR_arch.py:47-67).  hr [B,3,H,W], lr [B,3,h,w] CUDA fp32.
    Returns (clamp(fake_lr) NCHW, nll 0-dim) -- differentiable w.r.t. every parameter of ``net`` and ``hr``."""
    flow = net.flow
    assert flow.SR, "training path is implemented for HCFlowNet_SR"
    B, _, H, W = hr.shape
    pixels = H * W
    if dequant_noise is None:
        dequant_noise = torch.rand(hr.shape, device=hr.device)
    x = AxpbyFn.apply(hr.permute(0, 2, 3, 1).contiguous(), 1.0,
                      dequant_noise.to(hr.device, torch.float32).permute(0, 2, 3, 1).contiguous(), 1.0 / float(net.quant))
    logdet = torch.zeros(B, dtype=torch.float64, device=hr.device) + float(-math.log(net.quant) * pixels)
    z = x
    keep_y, keep_a = {}, {}
    for lay in flow.layers:
        if isinstance(lay, M.SqueezeLayer):
            z = SqueezeFn.apply(z)
        elif isinstance(lay, M.FlowStep):
            z, logdet = _flow_step_forward(z, None, lay, logdet)
        elif isinstance(lay, M.Split):
            n = lay.num_channels_split
            keep_y[lay.level], keep_a[lay.level] = z[..., :n].contiguous(), z[..., n:].contiguous()
            z = keep_y[lay.level]
        else:
            raise NotImplementedError(type(lay).__name__)
    feats = {}
    for level in range(flow.L - 1, -1, -1):
        cf = flow.cond_flow(level)
        u = _cat([keep_y[level]] + [UpsampleFn.apply(feats[l], l - level) for l in range(level + 1, flow.L)])
        feat = _cond_feature(u, cf)
        feats[level] = feat
        a = keep_a[level]
        for st in cf.additional_flow_steps:
            a, logdet = _flow_step_forward(a, feat, st, logdet)
        hp = _conv(feat, cf.f.weight, cf.f.bias, torch.exp(cf.f.logs * 3.0), ACT_NONE)
        mean, logs = hp[..., 0::2].contiguous(), hp[..., 1::2].contiguous()
        logdet = logdet + GaussLogpFn.apply(a, mean, logs, 0.0)
    fake_lr = Quant8Fn.apply(z)                                    # z: [B,h,w,3]
    lr_nhwc = lr.permute(0, 2, 3, 1).contiguous()
    objective = logdet + GaussLogpFn.apply(fake_lr, lr_nhwc, None, -6.0)      # logp(mean=lr, logs=-6, x=fake_lr)
    nll = ((-objective) / float(math.log(2.0) * pixels)).mean().to(torch.float32)
    return torch.clamp(fake_lr.permute(0, 3, 1, 2), 0, 1), nll


def _flow_step_reverse(z, u, step):
    """FlowStep.reverse_flow (FlowStep.py:53-64): coupling^-1 -> invconv^-1 -> ActNorm^-1."""
    B, H, W, Cc = z.shape
    aff = step.affine
    assert aff.mode == "affine", "training path: AffineCoupling only (the SR nets)"
    n = aff.n_pass
    z1, z2 = z[..., :n].contiguous(), z[..., n:].contiguous()
    h = _fcn(z1 if u is None else _cat((z1, u)), aff.f)
    z2, _ = CouplingFn.apply(z2, h, H * W, True)
    z = _cat((z1, z2))
    if step.permute is not None:
        # the reference inverts in fp64 on every call (Permutations.py:74); torch.inverse is differentiable
        winv = torch.inverse(step.permute.weight.double()).float()
        z = Conv2dFn.apply(z, winv.view(Cc, Cc, 1, 1))
    # x * exp(-logs) - bias  ==  (x + (-bias * exp(logs))) * exp(-logs)      (ActNorms.py:90-93)
    _actnorm_data_init(step.actnorm, z)     # (ActNorms.py:79-80 runs in either direction)
    logs, bias = step.actnorm.logs, step.actnorm.bias
    return AffineActFn.apply(z, -bias * torch.exp(logs), torch.exp(-logs), ACT_NONE)


def sr_reverse(net, lr, eps_std=0.0, eps=None):
    """HCFlowNet_SR.reverse_flow_diracLR with autograd (HCFlowNet_SR_arch.py:70-75 -> FlowNet_SR_x4.py:106-123): the
    inverse-path loss of optimize_parameters (HCFlow_SR_model.py:207-218) back-propagates through it.  lr [B,3,h,w] CUDA;
    eps: unit-normal tensors per level (deepest first), else drawn with torch like the reference.  Returns clamp(fake_hr)."""
    flow = net.flow
    assert flow.SR, "training path is implemented for HCFlowNet_SR"
    std = 0.0 if eps_std is None else float(eps_std)
    z = lr.to(torch.float32).permute(0, 2, 3, 1).contiguous()
    feats, draw = {}, 0
    for lay in reversed(list(flow.layers)):
        if isinstance(lay, M.SqueezeLayer):
            z = UnsqueezeFn.apply(z)
        elif isinstance(lay, M.FlowStep):
            z = _flow_step_reverse(z, None, lay)
        elif isinstance(lay, M.Split):
            level = lay.level
            cf = flow.cond_flow(level)
            u = _cat([z] + [UpsampleFn.apply(feats[l], l - level) for l in range(level + 1, flow.L)])
            feat = _cond_feature(u, cf)
            feats[level] = feat
            hp = _conv(feat, cf.f.weight, cf.f.bias, torch.exp(cf.f.logs * 3.0), ACT_NONE)
            mean, logs = hp[..., 0::2].contiguous(), hp[..., 1::2].contiguous()
            if eps is not None:
                e = eps[draw].to(lr.device, torch.float32).permute(0, 2, 3, 1).contiguous() * std
            else:
                e = torch.empty_like(mean.permute(0, 3, 1, 2)).normal_(0.0, 1.0).mul_(std).permute(0, 2, 3, 1).contiguous()
            draw += 1
            a = GaussSampleFn.apply(mean, logs, e)
            for st in reversed(list(cf.additional_flow_steps)):
                a = _flow_step_reverse(a, feat, st)
            z = _cat((z, a))
        else:
            raise NotImplementedError(type(lay).__name__)
    return torch.clamp(z.permute(0, 3, 1, 2), 0, 1)
