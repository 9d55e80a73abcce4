"""CPU oracle for the HCFlow hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (no nn.Module, no parameters of its own) restatement of the reference's
flow-step stack, squeeze/Haar/split plumbing, RRDB conditional encoder, Gaussian
prior and arch wrappers, written against plain torch CPU tensors.  It takes a
reference-layout ``state_dict`` + an ``opt`` dict + inputs and returns what the
reference returns.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this file.

Pinning: ``oracle/make_golden.py`` runs the UNMODIFIED reference modules (imported
from /root/reference in the authoring container) on the same synthetic weights and
noise and stores their outputs under ``tests/golden/``; ``tests/test_oracle.py``
checks this restatement against those vectors.  The reference itself ships no
tests or golden vectors, so that is the strongest pin there is.

Every function cites the reference lines it follows (paths relative to
/root/reference/codes/models/modules/).  All tensors are NCHW like the reference.
The compute dtype follows the inputs (float32 = reference behaviour, float64 = truth).
"""
import math

import torch
import torch.nn.functional as F

LOG2PI = float(math.log(2 * math.pi))


def _get(opt, keys, default=None):
    cur = opt
    for k in keys:
        if cur is None:
            return default
        cur = cur.get(k, None) if isinstance(cur, dict) else None
    return default if cur is None else cur


# ----------------------------------------------------------------------------- thops
def sum_chw(t):
    """thops.sum(t, dim=[1,2,3]) -- one dim at a time, C then H then W (thops.py:4-17)."""
    return t.sum(dim=1).sum(dim=1).sum(dim=1)


# ----------------------------------------------------------------------------- Basic
def squeeze2d(x, factor=2):
    """Basic.py:127-140: out[b, c*4+i*2+j, h, w] = in[b, c, 2h+i, 2w+j]."""
    B, C, H, W = x.shape
    assert H % factor == 0 and W % factor == 0
    x = x.reshape(B, C, H // factor, factor, W // factor, factor)
    x = x.permute(0, 1, 3, 5, 2, 4)
    return x.reshape(B, C * factor * factor, H // factor, W // factor)


def unsqueeze2d(x, factor=2):
    """Basic.py:143-157 (inverse of squeeze2d)."""
    B, C, H, W = x.shape
    f2 = factor * factor
    assert C % f2 == 0
    x = x.reshape(B, C // f2, factor, factor, H, W)
    x = x.permute(0, 1, 4, 2, 5, 3)
    return x.reshape(B, C // f2, H * factor, W * factor)


def haar_forward(x, haar_weights):
    """Basic.py:470-478: grouped 2x2 stride-2 conv / 4, then band-major channel order."""
    B, C, H, W = x.shape
    out = F.conv2d(x, haar_weights.to(x.dtype), None, stride=2, groups=C) / 4.0
    out = out.reshape(B, C, 4, H // 2, W // 2).transpose(1, 2)
    return out.reshape(B, C * 4, H // 2, W // 2)


def haar_reverse(x, haar_weights):
    """Basic.py:479-487: undo the channel transpose, then conv_transpose2d."""
    B, C4, H, W = x.shape
    C = C4 // 4
    out = x.reshape(B, 4, C, H, W).transpose(1, 2).reshape(B, C4, H, W)
    return F.conv_transpose2d(out, haar_weights.to(x.dtype), None, stride=2, groups=C)


class _Quant(torch.autograd.Function):
    """Basic.py:186-198: clamp to [0,1], round to 8 bit; the backward passes the gradient straight through."""

    @staticmethod
    def forward(ctx, x):
        return (torch.clamp(x, 0, 1) * 255.0).round() / 255.0

    @staticmethod
    def backward(ctx, g):
        return g


def quantize(x):
    """Basic.py:186-202 (Quantization module = Quant.apply)."""
    return _Quant.apply(x)


def gaussian_logp(mean, logs, x):
    """Basic.py:79-93."""
    ll = -0.5 * (logs * 2.0 + ((x - mean) ** 2) / torch.exp(logs * 2.0) + LOG2PI)
    return sum_chw(ll)


def conv3x3(x, w, b=None):
    return F.conv2d(x, w.to(x.dtype), None if b is None else b.to(x.dtype), padding=1)


def actnorm_conv(x, sd, pre, ksize):
    """Basic.py:49-53: bias-free conv, then ActNorm forward (x + bias) * exp(logs)."""
    w = sd[pre + ".weight"].to(x.dtype)
    y = F.conv2d(x, w, None, padding=(ksize - 1) // 2)
    return (y + sd[pre + ".actnorm.bias"].to(x.dtype)) * torch.exp(sd[pre + ".actnorm.logs"].to(x.dtype))


def zero_conv(x, sd, pre):
    """Basic.py:70-72: conv(+bias) * exp(3 * logs)."""
    y = conv3x3(x, sd[pre + ".weight"], sd[pre + ".bias"])
    return y * torch.exp(sd[pre + ".logs"].to(x.dtype) * 3.0)


def fcn(x, sd, pre):
    """Basic.py:442-447."""
    x = F.relu(actnorm_conv(x, sd, pre + ".conv1", 3))
    x = F.relu(actnorm_conv(x, sd, pre + ".conv2", 1))
    return zero_conv(x, sd, pre + ".conv3")


def _dense5(x, sd, pre):
    """The shared 5-conv dense pattern (Basic.py:349-356, 377-382)."""
    def c(i, inp):
        return conv3x3(inp, sd["{}.conv{}.weight".format(pre, i)], sd["{}.conv{}.bias".format(pre, i)])
    x1 = F.leaky_relu(c(1, x), 0.2)
    x2 = F.leaky_relu(c(2, torch.cat((x, x1), 1)), 0.2)
    x3 = F.leaky_relu(c(3, torch.cat((x, x1, x2), 1)), 0.2)
    x4 = F.leaky_relu(c(4, torch.cat((x, x1, x2, x3), 1)), 0.2)
    return c(5, torch.cat((x, x1, x2, x3, x4), 1))


def dense_block(x, sd, pre):
    return _dense5(x, sd, pre)


def rdb(x, sd, pre):
    """Basic.py:377-383."""
    return _dense5(x, sd, pre) * 0.2 + x


def rrdb(x, sd, pre):
    """Basic.py:394-398."""
    out = rdb(x, sd, pre + ".RDB1")
    out = rdb(out, sd, pre + ".RDB2")
    out = rdb(out, sd, pre + ".RDB3")
    return out * 0.2 + x


# ----------------------------------------------------------------------------- flow step
def actnorm(x, sd, pre, logdet, reverse):
    """ActNorms.py:45-94."""
    bias = sd[pre + ".bias"].to(x.dtype)
    logs = sd[pre + ".logs"].to(x.dtype)
    pixels = x.shape[2] * x.shape[3]
    if not reverse:
        x = (x + bias) * torch.exp(logs)
        if logdet is not None:
            logdet = logdet + logs.sum() * pixels
    else:
        x = x * torch.exp(-logs) - bias
        if logdet is not None:
            logdet = logdet - logs.sum() * pixels
    return x, logdet


def invconv(x, sd, pre, logdet, reverse):
    """Permutations.py:61-76, 94-108 (non-LU): slogdet on forward, fp64 inverse on reverse."""
    w = sd[pre + ".weight"]
    C = w.shape[0]
    pixels = x.shape[2] * x.shape[3]
    if not reverse:
        dlogdet = torch.slogdet(w.to(x.dtype))[1] * pixels
        z = F.conv2d(x, w.to(x.dtype).view(C, C, 1, 1))
        if logdet is not None:
            logdet = logdet + dlogdet
    else:
        winv = torch.inverse(w.double())
        winv = winv.float().to(x.dtype) if x.dtype == torch.float32 else winv.to(x.dtype)
        z = F.conv2d(x, winv.view(C, C, 1, 1))
    return z, logdet


def _subnet(x, sd, pre, kind):
    return fcn(x, sd, pre) if kind == "FCN" else dense_block(x, sd, pre)


def _subnet_kind(sd, pre):
    return "FCN" if (pre + ".conv1.actnorm.bias") in sd else "DenseBlock"


def coupling(z, u, sd, pre, logdet, reverse, mode, n_pass):
    """AffineCouplings.py:28-87 (mode 'affine', n_pass=C//2) and :117-160 (3shift)."""
    fpre = pre + ".f"
    kind = _subnet_kind(sd, fpre)
    if mode == "affine":
        z1, z2 = z[:, :n_pass], z[:, n_pass:]
        h = _subnet(z1 if u is None else torch.cat((z1, u), 1), sd, fpre, kind)
        shift, scale = h[:, 0::2], h[:, 1::2]
        logscale = 0.318 * torch.atan(2 * scale)
        if not reverse:
            z2 = (z2 + shift) * torch.exp(logscale)
            if logdet is not None:
                logdet = logdet + sum_chw(logscale)
        else:
            z2 = z2 * torch.exp(-logscale) - shift
        return torch.cat((z1, z2), 1), logdet
    # shift_first3: the last C-3 channels shift the first 3
    z2, z1 = z[:, :3], z[:, 3:]
    if not reverse:
        shift = _subnet(z1 if u is None else torch.cat((z1, u), 1), sd, fpre, kind)
        z2 = z2 + shift
    else:
        shift = _subnet(z1, sd, fpre, kind)  # AffineCouplings.py:154 ignores u on reverse
        z2 = z2 - shift
    return torch.cat((z2, z1), 1), logdet


def flow_step(z, u, sd, pre, logdet, reverse, mode, n_pass):
    """FlowStep.py:40-64."""
    has_perm = (pre + ".permute.weight") in sd
    if not reverse:
        z, logdet = actnorm(z, sd, pre + ".actnorm", logdet, False)
        if has_perm:
            z, logdet = invconv(z, sd, pre + ".permute", logdet, False)
        z, logdet = coupling(z, u, sd, pre + ".affine", logdet, False, mode, n_pass)
    else:
        z, _ = coupling(z, u, sd, pre + ".affine", None, True, mode, n_pass)
        if has_perm:
            z, _ = invconv(z, sd, pre + ".permute", None, True)
        z, _ = actnorm(z, sd, pre + ".actnorm", None, True)
    return z, logdet


# ----------------------------------------------------------------------------- conditional flow
def _count(sd, pre):
    n = 0
    while any(k.startswith("{}.{}.".format(pre, n)) for k in sd):
        n += 1
    return n


def cond_feature(u, sd, pre, SR):
    """ConditionalFlow.py:99-110."""
    first = conv3x3(u, sd[pre + ".conv_first.weight"], sd[pre + ".conv_first.bias"])
    x = first
    for j in range(_count(sd, pre + ".RRDB_trunk0")):
        x = rrdb(x, sd, "{}.RRDB_trunk0.{}".format(pre, j))
    f1 = x
    for j in range(_count(sd, pre + ".RRDB_trunk1")):
        x = rrdb(x, sd, "{}.RRDB_trunk1.{}".format(pre, j))
    f2 = conv3x3(x, sd[pre + ".trunk_conv1.weight"], sd[pre + ".trunk_conv1.bias"]) + first
    return torch.cat([f1, f2], 1) if SR else f2


def cond_flow(z, u, sd, pre, eps, logdet, reverse, SR):
    """ConditionalFlow.py:44-96.  ``eps``: noise tensor already scaled by eps_std
    (stands in for the torch.normal draw of Basic.py:96-100)."""
    cf = cond_feature(u, sd, pre, SR)
    nstep = _count(sd, pre + ".additional_flow_steps")
    h = zero_conv(cf, sd, pre + ".f")
    mean, second = h[:, 0::2], h[:, 1::2]
    zc = second.shape[1]
    steps = ["{}.additional_flow_steps.{}".format(pre, j) for j in range(nstep)]
    if SR:
        if not reverse:
            for s in steps:
                z, logdet = flow_step(z, cf, sd, s, logdet, False, "affine", zc // 2)
            logdet = logdet + gaussian_logp(mean, second, z)
            return logdet, cf
        z = mean + torch.exp(second) * eps
        for s in reversed(steps):
            z, _ = flow_step(z, cf, sd, s, None, True, "affine", zc // 2)
        return z, cf
    logscale = 0.318 * torch.atan(2 * second)
    if not reverse:
        for s in steps:
            z, logdet = flow_step(z, cf, sd, s, logdet, False, "affine", zc // 2)
        return (z - mean) * torch.exp(-logscale), cf
    z = mean + torch.exp(logscale) * eps
    for s in reversed(steps):
        z, _ = flow_step(z, cf, sd, s, None, True, "affine", zc // 2)
    return z, cf


# ----------------------------------------------------------------------------- flow graphs
def layer_list(opt, SR):
    """The (kind, ...) list the FlowNet constructors build (FlowNet_SR_x4.py:28-62,
    FlowNet_SR_x8.py:28-71, FlowNet_Rescaling_x4.py:30-71)."""
    fd = ["network_G", "flowDownsampler"]
    L = _get(opt, fd + ["L"])
    K = _get(opt, fd + ["K"])
    K = [K] * (L + 1) if isinstance(K, int) else list(K)
    after = _get(opt, fd + ["splitOff", "after_flowstep"], 0)
    after = [after] * (L + 1) if isinstance(after, int) else list(after)
    squeeze = "checkerboard" if SR else _get(opt, fd + ["squeeze"], "checkerboard")
    three = _get(opt, fd + ["flow_coupling"], "Affine") == "Affine3shift"
    C = _get(opt, ["network_G", "in_nc"], 3)
    layers = []
    for level in range(L):
        layers.append(("squeeze" if squeeze == "checkerboard" else "haar", C))
        C *= 4
        for k in range(K[level] - after[level]):
            if not three:
                layers.append(("step", "affine", C // 2))
            elif k % 2 == 0:
                layers.append(("step", "affine", 3))
            else:
                layers.append(("step", "shift_first3", C - 3))
        n_split = C // 2 if level < L - 1 else 3
        layers.append(("split", level, n_split))
        C = n_split
    return layers, L


def _up(x, s):
    return F.interpolate(x, scale_factor=s, mode="nearest")


def flownet_forward(x, sd, opt, logdet, SR):
    """normal_flow of FlowNet_SR_x4.py:84-101 / _x8.py:91-118 / Rescaling_x4.py:90-108."""
    layers, L = layer_list(opt, SR)
    z = x
    keep_y, keep_a = {}, {}
    for i, lay in enumerate(layers):
        pre = "flow.layers.{}".format(i)
        if lay[0] == "squeeze":
            z = squeeze2d(z)
        elif lay[0] == "haar":
            z = haar_forward(z, sd[pre + ".haar_weights"])
        elif lay[0] == "step":
            z, logdet = flow_step(z, None, sd, pre, logdet, False, lay[1], lay[2])
        else:
            level, n = lay[1], lay[2]
            z, a = z[:, :n], z[:, n:]
            keep_y[level], keep_a[level] = z, a
    feats, fake_z = {}, {}
    for level in range(L - 1, -1, -1):
        u = keep_y[level]
        ups = [_up(feats[l], 2 ** (l - level)) for l in range(level + 1, L)]
        u = torch.cat([u] + ups, 1) if ups else u
        pre = "flow.level{}_condFlow".format(level)
        if SR:
            logdet, feats[level] = cond_flow(keep_a[level], u, sd, pre, None, logdet, False, True)
        else:
            fake_z[level], feats[level] = cond_flow(keep_a[level], u, sd, pre, None, logdet, False, False)
    if SR:
        return z, logdet
    return z, fake_z[0], fake_z[1]


def flownet_reverse(lr, sd, opt, eps_list, SR):
    """reverse_flow of FlowNet_SR_x4.py:106-123 / _x8.py:123-144 / Rescaling_x4.py:113-128.
    ``eps_list``: scaled noise per level in draw order (deepest level first)."""
    layers, L = layer_list(opt, SR)
    z = lr
    feats = {}
    draw = 0
    for i in range(len(layers) - 1, -1, -1):
        lay = layers[i]
        pre = "flow.layers.{}".format(i)
        if lay[0] == "squeeze":
            z = unsqueeze2d(z)
        elif lay[0] == "haar":
            z = haar_reverse(z, sd[pre + ".haar_weights"])
        elif lay[0] == "step":
            z, _ = flow_step(z, None, sd, pre, None, True, lay[1], lay[2])
        else:
            level = lay[1]
            ups = [_up(feats[l], 2 ** (l - level)) for l in range(level + 1, L)]
            u = torch.cat([z] + ups, 1) if ups else z
            a, feats[level] = cond_flow(None, u, sd, "flow.level{}_condFlow".format(level),
                                        eps_list[draw], None, True, SR)
            draw += 1
            z = torch.cat((z, a), 1)
    return z


def noise_shapes(opt, B, h, w, SR=True):
    """Shapes of the per-level noise draws for an LR input of size h x w (draw order)."""
    layers, L = layer_list(opt, SR)
    shapes = []
    hh, ww = h, w
    splits = [l for l in layers if l[0] == "split"]
    # channel count before each split
    C = _get(opt, ["network_G", "in_nc"], 3)
    chans = []
    for level in range(L):
        C *= 4
        n = splits[level][2]
        chans.append(C - n)
        C = n
    for level in range(L - 1, -1, -1):
        shapes.append((B, chans[level], hh, ww))
        hh, ww = hh * 2, ww * 2
    return shapes


# ----------------------------------------------------------------------------- arch wrappers
def sr_reverse(lr, sd, opt, eps_list):
    """HCFlowNet_SR_arch.py:70-75: clamp(flow(z=lr, reverse), 0, 1). Returns (clamped, raw)."""
    raw = flownet_reverse(lr, sd, opt, eps_list, True)
    return torch.clamp(raw, 0, 1), raw


def sr_forward(hr, lr, sd, opt, dequant_noise):
    """HCFlowNet_SR_arch.py:47-67.  ``dequant_noise``: the U[0,1) draw of :52.
    Returns (clamp(fake_lr), nll, z_raw, logdet)."""
    quant = _get(opt, ["quant"], 256)
    pixels = hr.shape[2] * hr.shape[3]
    x = hr + dequant_noise / quant
    logdet = torch.zeros_like(x[:, 0, 0, 0]) + float(-math.log(quant) * pixels)
    z, logdet = flownet_forward(x, sd, opt, logdet, True)
    fake_lr = quantize(z)
    objective = logdet + gaussian_logp(lr, -torch.ones_like(lr) * 6, fake_lr)
    nll = ((-objective) / float(math.log(2.0) * pixels)).mean()
    return torch.clamp(fake_lr, 0, 1), nll, z, logdet


def rescaling_forward(hr, sd, opt):
    """HCFlowNet_Rescaling_arch.py:39-46. Returns (clamp(lr), z1, z2, raw_lr)."""
    z, fz1, fz2 = flownet_forward(hr, sd, opt, None, False)
    return torch.clamp(z, 0, 1), fz1, fz2, z


def rescaling_reverse(lr, sd, opt, eps_list):
    """HCFlowNet_Rescaling_arch.py:49-54. Returns (clamped, raw)."""
    raw = flownet_reverse(lr, sd, opt, eps_list, False)
    return torch.clamp(raw, 0, 1), raw
