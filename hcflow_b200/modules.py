"""Parameter tree of the HCFlow networks (state_dict-compatible with the reference).

These classes own ``nn.Parameter``s under exactly the key names / shapes the
reference checkpoints use (they are loaded with ``strict=True``, reference
codes/models/base_model.py:96-120), and nothing else: there is no per-module
``forward``.  The arithmetic is done by the CUDA engine (``engine.py``), which
walks this tree once to build a launch plan.

Key layout mirrored (reference files under codes/models/modules/):
  FlowStep.py:8-32            actnorm / permute / affine
  ActNorms.py:16-24           bias, logs  [1,C,1,1]
  Permutations.py:33-39       permute.weight [C,C]
  AffineCouplings.py:10-26, 92-115   affine.f  (FCN | DenseBlock)
  Basic.py:14-72              Conv2d(+actnorm), Conv2dZeros(weight,bias,logs[C,1,1])
  Basic.py:329-398            DenseBlock / ResidualDenseBlock / RRDB
  Basic.py:450-468            HaarDownsampling.haar_weights [4C,1,2,2]
  ConditionalFlow.py:15-41    conv_first, RRDB_trunk{0,1}, trunk_conv1, additional_flow_steps, f
  FlowNet_SR_x4.py:12-71, FlowNet_SR_x8.py:12-78, FlowNet_Rescaling_x4.py:12-79
"""
import math

import torch
from torch import nn

from .options import opt_get


def _xavier_normal_(w, scale):
    fan_out = w.shape[0] * w[0][0].numel()
    fan_in = w.shape[1] * w[0][0].numel()
    std = math.sqrt(2.0 / float(fan_in + fan_out)) * scale
    with torch.no_grad():
        w.normal_(0.0, std)


class ActNorm2d(nn.Module):
    """bias/logs of shape [1,C,1,1]; ``inited`` mirrors ActNorms.py:24,79-80."""

    def __init__(self, num_features, scale=1.0):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(1, num_features, 1, 1))
        self.logs = nn.Parameter(torch.zeros(1, num_features, 1, 1))
        self.num_features = num_features
        self.scale = float(scale)
        self.inited = False


class InvertibleConv1x1(nn.Module):
    """C x C mixing matrix, initialised to a random rotation (Permutations.py:36-40)."""

    def __init__(self, num_channels):
        super().__init__()
        q, _ = torch.linalg.qr(torch.randn(num_channels, num_channels, dtype=torch.float64))
        self.weight = nn.Parameter(q.float().contiguous())
        self.w_shape = [num_channels, num_channels]


class ActNormConv2d(nn.Module):
    """Bias-free conv followed by ActNorm (Basic.py:14-53); keys: weight, actnorm.{bias,logs}."""

    def __init__(self, cin, cout, ksize):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, ksize, ksize))
        self.actnorm = ActNorm2d(cout)
        self.ksize = ksize


class ZeroConv2d(nn.Module):
    """3x3 conv whose output is scaled by exp(3*logs), all-zero at init (Basic.py:57-72)."""

    logscale_factor = 3.0

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(cout, cin, 3, 3))
        self.bias = nn.Parameter(torch.zeros(cout))
        self.logs = nn.Parameter(torch.zeros(cout, 1, 1))


class PlainConv2d(nn.Module):
    """3x3 conv with bias (nn.Conv2d keys: weight, bias)."""

    def __init__(self, cin, cout, init="xavier", scale=0.1):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, 3, 3))
        self.bias = nn.Parameter(torch.zeros(cout))
        if init == "xavier":
            _xavier_normal_(self.weight, scale)
        elif init == "zero":
            with torch.no_grad():
                self.weight.zero_()
        else:  # torch's nn.Conv2d default
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            bound = 1.0 / math.sqrt(cin * 9)
            nn.init.uniform_(self.bias, -bound, bound)


class FCN(nn.Module):
    """conv3x3+actnorm+relu -> conv1x1+actnorm+relu -> zero conv3x3 (Basic.py:426-447)."""

    kind = "FCN"

    def __init__(self, cin, cout, hidden):
        super().__init__()
        self.conv1 = ActNormConv2d(cin, hidden, 3)
        self.conv2 = ActNormConv2d(hidden, hidden, 1)
        self.conv3 = ZeroConv2d(hidden, cout)
        _xavier_normal_(self.conv1.weight, 0.1)
        _xavier_normal_(self.conv2.weight, 0.1)
        self.cin, self.cout, self.hidden = cin, cout, hidden


class DenseBlock(nn.Module):
    """5 densely connected 3x3 convs, last one zero-init (Basic.py:329-356)."""

    kind = "DenseBlock"

    def __init__(self, cin, cout, gc):
        super().__init__()
        self.conv1 = PlainConv2d(cin, gc)
        self.conv2 = PlainConv2d(cin + gc, gc)
        self.conv3 = PlainConv2d(cin + 2 * gc, gc)
        self.conv4 = PlainConv2d(cin + 3 * gc, gc)
        self.conv5 = PlainConv2d(cin + 4 * gc, cout, init="zero")
        self.cin, self.cout, self.gc = cin, cout, gc


class ResidualDenseBlock(nn.Module):
    """Basic.py:360-383."""

    def __init__(self, nf, gc):
        super().__init__()
        self.conv1 = PlainConv2d(nf, gc)
        self.conv2 = PlainConv2d(nf + gc, gc)
        self.conv3 = PlainConv2d(nf + 2 * gc, gc)
        self.conv4 = PlainConv2d(nf + 3 * gc, gc)
        self.conv5 = PlainConv2d(nf + 4 * gc, nf)
        self.nf, self.gc = nf, gc


class RRDB(nn.Module):
    """Basic.py:385-398."""

    def __init__(self, nf, gc):
        super().__init__()
        self.RDB1 = ResidualDenseBlock(nf, gc)
        self.RDB2 = ResidualDenseBlock(nf, gc)
        self.RDB3 = ResidualDenseBlock(nf, gc)


class AffineCoupling(nn.Module):
    """Holder for the coupling sub-net ``f`` plus the split description.

    mode "affine":       z1 = z[:, :n_pass] conditions an affine map of the rest
                         (AffineCouplings.py:28-87; with n_pass=3 it is the
                         LRvsothers=True branch of AffineCoupling3shift, :120-129).
    mode "shift_first3": the last C-3 channels condition a pure shift of the
                         first 3 (AffineCoupling3shift LRvsothers=False, :130-133).
    """

    def __init__(self, in_channels, cond_channels, opt, three_shift=False, lr_vs_others=True):
        super().__init__()
        hidden = opt_get(opt, ["hidden_channels"], 64)
        nn_module = opt_get(opt, ["nn_module"], "FCN")
        cc = 0 if cond_channels is None else cond_channels
        if not three_shift:
            self.mode, self.n_pass = "affine", in_channels // 2
            f_in, f_out = self.n_pass + cc, (in_channels - self.n_pass) * 2
        elif lr_vs_others:
            self.mode, self.n_pass = "affine", 3
            f_in, f_out = 3 + cc, (in_channels - 3) * 2
        else:
            self.mode, self.n_pass = "shift_first3", in_channels - 3
            f_in, f_out = in_channels - 3 + cc, 3
        if nn_module == "DenseBlock":
            self.f = DenseBlock(f_in, f_out, hidden)
        elif nn_module == "FCN":
            self.f = FCN(f_in, f_out, hidden)
        else:
            raise NotImplementedError("nn_module {}".format(nn_module))
        self.in_channels, self.cond_channels = in_channels, cc


class FlowStep(nn.Module):
    """actnorm -> (invconv | none) -> coupling (FlowStep.py:8-32)."""

    def __init__(self, in_channels, cond_channels=None, flow_permutation="invconv",
                 flow_coupling="Affine", LRvsothers=True, opt=None):
        super().__init__()
        self.actnorm = ActNorm2d(in_channels)
        if flow_permutation == "invconv":
            self.permute = InvertibleConv1x1(in_channels)
        elif flow_permutation == "none":
            self.permute = None
        else:
            raise NotImplementedError("flow_permutation {}".format(flow_permutation))
        if flow_coupling == "Affine":
            self.affine = AffineCoupling(in_channels, cond_channels, opt)
        elif flow_coupling == "Affine3shift":
            self.affine = AffineCoupling(in_channels, cond_channels, opt, three_shift=True,
                                         lr_vs_others=LRvsothers)
        else:
            raise NotImplementedError("flow_coupling {}".format(flow_coupling))
        self.in_channels = in_channels


class SqueezeLayer(nn.Module):
    def __init__(self, factor=2):
        super().__init__()
        self.factor = factor


class HaarDownsampling(nn.Module):
    """Owns the fixed +-1 Haar filter bank under the reference's key (Basic.py:450-468)."""

    def __init__(self, channel_in):
        super().__init__()
        k = torch.ones(4, 1, 2, 2)
        k[1, 0, :, 1] = -1.0
        k[2, 0, 1, :] = -1.0
        k[3, 0, 1, 0] = -1.0
        k[3, 0, 0, 1] = -1.0
        self.haar_weights = nn.Parameter(k.repeat(channel_in, 1, 1, 1), requires_grad=False)
        self.channel_in = channel_in


class Split(nn.Module):
    def __init__(self, num_channels_split, level):
        super().__init__()
        self.num_channels_split = num_channels_split
        self.level = level


class ConditionalFlow(nn.Module):
    """RRDB encoder + prior conv + conditional FlowSteps (ConditionalFlow.py:15-41)."""

    def __init__(self, num_channels, num_channels_split, n_flow_step, opt, num_levels_condition, SR):
        super().__init__()
        self.SR = SR
        n_feat = 2 if SR else 1
        nb = opt_get(opt, ["RRDB_nb"], [5, 5])
        nf = opt_get(opt, ["RRDB_nf"], 64)
        gc = opt_get(opt, ["RRDB_gc"], 32)
        self.nb, self.nf, self.gc = list(nb), nf, gc
        self.in_channels = num_channels_split + nf * n_feat * num_levels_condition
        self.cond_channels = nf * n_feat
        self.z_channels = num_channels - num_channels_split
        self.conv_first = PlainConv2d(self.in_channels, nf, init="default")
        self.RRDB_trunk0 = nn.Sequential(*[RRDB(nf, gc) for _ in range(nb[0])])
        self.RRDB_trunk1 = nn.Sequential(*[RRDB(nf, gc) for _ in range(nb[1])])
        self.trunk_conv1 = PlainConv2d(nf, nf, init="default")
        self.additional_flow_steps = nn.ModuleList([
            FlowStep(self.z_channels, cond_channels=self.cond_channels,
                     flow_permutation=opt["flow_permutation"], flow_coupling=opt["flow_coupling"], opt=opt)
            for _ in range(n_flow_step)])
        self.f = ZeroConv2d(self.cond_channels, self.z_channels * 2)


class FlowNet(nn.Module):
    """Layer list for SR x4 / x8 and Rescaling x4; the level count L selects the graph."""

    def __init__(self, image_shape, opt, SR=True):
        super().__init__()
        H, W, C = image_shape
        assert C in (1, 3)
        fd = ["network_G", "flowDownsampler"]
        self.L = opt_get(opt, fd + ["L"])
        K = opt_get(opt, fd + ["K"])
        self.K = [K] * (self.L + 1) if isinstance(K, int) else list(K)
        squeeze = opt_get(opt, fd + ["squeeze"], "checkerboard") if not SR else "checkerboard"
        perm = opt_get(opt, fd + ["flow_permutation"], "invconv")
        coup = opt_get(opt, fd + ["flow_coupling"], "Affine")
        cond = opt_get(opt, fd + ["cond_channels"], None)
        split_on = opt_get(opt, fd + ["splitOff", "enable"], False)
        after = opt_get(opt, fd + ["splitOff", "after_flowstep"], 0)
        after = [after] * (self.L + 1) if isinstance(after, int) else list(after)
        self.SR = SR
        self.layers = nn.ModuleList()
        self.output_shapes = []
        for level in range(self.L):
            self.layers.append(SqueezeLayer(2) if squeeze == "checkerboard" else HaarDownsampling(C))
            C, H, W = C * 4, H // 2, W // 2
            self.output_shapes.append([-1, C, H, W])
            for k in range(self.K[level] - after[level]):
                self.layers.append(FlowStep(C, cond_channels=cond, flow_permutation=perm, flow_coupling=coup,
                                            LRvsothers=(k % 2 == 0),
                                            opt=opt["network_G"]["flowDownsampler"]))
                self.output_shapes.append([-1, C, H, W])
            if split_on:
                n_split = C // 2 if level < self.L - 1 else 3
                self.layers.append(Split(n_split, level))
                cf = ConditionalFlow(C, n_split, after[level], opt["network_G"]["flowDownsampler"]["splitOff"],
                                     num_levels_condition=self.L - 1 - level, SR=SR)
                setattr(self, "level{}_condFlow".format(level), cf)
                C = n_split
                self.output_shapes.append([-1, C, H, W])
        self.C, self.H, self.W = C, H, W

    def cond_flow(self, level):
        return getattr(self, "level{}_condFlow".format(level))
