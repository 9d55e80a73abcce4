"""Weight preparation: reference-layout parameters -> kernel-ready device arrays.

Runs once per weight version (the engine caches on the parameters' version counters),
replacing work the reference redoes on EVERY call:
  - W^-1 through an fp64 inverse (Permutations.py:74 does it per reverse call, 52x per pass);
  - log|det W| (Permutations.py:70 does an fp32 slogdet on the CPU per forward call);
  - exp(+-logs) of every ActNorm (ActNorms.py:62-66) and exp(3*logs) of Conv2dZeros (Basic.py:72);
  - conv weights repacked from [Cout,Cin,kh,kw] to the kernels' [tap][k][n] layout with the
    concat segments padded to 8 channels.
"""
import math

import torch


def npad_for(cout):
    if cout <= 16:
        return 16
    if cout <= 32:
        return 32
    return (cout + 63) // 64 * 64


def seg_pad(c):
    return (c + 7) // 8 * 8


def pack_conv_weight(w, seg_channels, npad):
    """w [Cout, Cin, ks, ks] -> [ks*ks, kpad, npad] fp32 (zero padded)."""
    cout, cin, ks, _ = w.shape
    assert sum(seg_channels) == cin, (seg_channels, cin)
    kpad = sum(seg_pad(c) for c in seg_channels)
    out = torch.zeros(ks * ks, kpad, npad, dtype=torch.float32, device=w.device)   # (device-agnostic: the training path packs on the GPU)
    wt = w.detach().float().permute(2, 3, 1, 0).reshape(ks * ks, cin, cout)  # [tap, cin, cout]
    src = 0
    dst = 0
    for c in seg_channels:
        out[:, dst:dst + c, :cout] = wt[:, src:src + c, :]
        src += c
        dst += seg_pad(c)
    return out.contiguous()


def pad_vec(v, npad, fill):
    out = torch.full((npad,), float(fill), dtype=torch.float32)
    flat = v.detach().float().reshape(-1)
    out[: flat.numel()] = flat
    return out


def derive(sd, key):
    """Resolve a derived-parameter key ("name#op") against a state dict -> fp32 CPU tensor."""
    if "#" not in key:
        return sd[key].detach().float()
    name, op = key.split("#")
    t = sd[name].detach()
    if op == "exp":
        return torch.exp(t.float()).reshape(-1)
    if op == "exp3":
        return torch.exp(t.float() * 3.0).reshape(-1)
    if op == "exppos":
        return torch.exp(t.float()).reshape(-1)
    if op == "expneg":
        return torch.exp(-t.float()).reshape(-1)
    if op == "vec":
        return t.float().reshape(-1)
    if op == "mat":
        return t.float().contiguous()
    if op == "inv":
        # same arithmetic as the reference: fp64 inverse rounded to fp32 (Permutations.py:74)
        return torch.inverse(t.double().cpu()).float().contiguous()
    raise KeyError(key)


def logdet_constant(sd, terms):
    """Sum over forward FlowSteps of (sum(actnorm.logs) + log|det W|) * pixels, in fp64
    (ActNorms.py:72-75, Permutations.py:70)."""
    total = 0.0
    for pre, has_perm, pixels in terms:
        v = float(sd[pre + ".actnorm.logs"].detach().double().sum())
        if has_perm:
            v += float(torch.slogdet(sd[pre + ".permute.weight"].detach().double().cpu())[1])
        total += v * pixels
    return total


def quant_logdet(quant, pixels):
    """HCFlowNet_SR_arch.py:53: -log(quant) * H*W  (H*W, not C*H*W)."""
    return float(-math.log(quant) * pixels)


def tc_pad(c):
    return (c + 31) // 32 * 32


def pad_weight_for_tc(w, seg_channels, chunk=32):
    """w [Cout, Cin, ks, ks] -> [Cout, sum(ceil_chunk(C_seg)), ks, ks] with zero rows after each segment
    (the tensor-core kernel walks K in 32-channel chunks per segment; 64 for the fp16 kernels)."""
    cout, cin, ks, _ = w.shape
    assert sum(seg_channels) == cin, (seg_channels, cin)

    def pad(c):
        return (c + chunk - 1) // chunk * chunk
    out = torch.zeros(cout, sum(pad(c) for c in seg_channels), ks, ks, dtype=torch.float32)
    src = dst = 0
    for c in seg_channels:
        out[:, dst:dst + c] = w[:, src:src + c].detach().float()
        src += c
        dst += pad(c)
    return out.contiguous()
