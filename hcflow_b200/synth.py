"""Deterministic synthetic weights and inputs.

There are no trained checkpoints offline, and at the reference's random init every
coupling / prior is the identity (zero-initialised last convs, Basic.py:67-68,
346-347, 439-440; ActNorm bias=logs=0, ActNorms.py:21-22), which would make any
parity test vacuous.  ``synthetic_state_dict`` therefore fills a state_dict with
values that exercise every term while keeping the 52-step flow O(1)
(perturbation recipe):

  actnorm.{bias,logs}            0.05 * N(0,1)
  zero-init conv weights         0.002 * N(0,1);  their bias / logs  0.01 * N(0,1)
  invertible 1x1 weights         random rotation (QR) times (1 + 0.05*N(0,1)) per row
  every other conv               Xavier-normal * 0.1, bias 0.01 * N(0,1)

The generator is keyed by (sorted key name, seed) so that the product modules, the
oracle and the reference (in oracle/make_golden.py) all see the same numbers.
"""
import hashlib
import math

import torch


def _gen(key, seed):
    h = hashlib.sha256("{}|{}".format(seed, key).encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def _randn(shape, key, seed):
    return torch.randn(*shape, generator=_gen(key, seed), dtype=torch.float32)


def synthetic_state_dict(template, seed=1):
    """template: mapping key -> tensor (only shapes are used). Returns a new state dict."""
    out = {}
    for key in sorted(template.keys()):
        shape = tuple(template[key].shape)
        leaf = key.rsplit(".", 1)[-1]
        zero_conv = (".affine.f.conv3." in key and "actnorm" not in key) or key.endswith("_condFlow.f.weight") \
            or key.endswith("_condFlow.f.bias") or key.endswith("_condFlow.f.logs") \
            or (".affine.f.conv5." in key)
        if key.endswith("haar_weights"):
            out[key] = template[key].detach().clone().float()
        elif ".actnorm." in key:
            out[key] = 0.05 * _randn(shape, key, seed)
        elif key.endswith("permute.weight"):
            c = shape[0]
            q, _ = torch.linalg.qr(torch.randn(c, c, generator=_gen(key, seed), dtype=torch.float64))
            s = 1.0 + 0.05 * torch.randn(c, 1, generator=_gen(key + "#s", seed), dtype=torch.float64)
            out[key] = (q * s).float().contiguous()
        elif zero_conv:
            out[key] = (0.002 if leaf == "weight" else 0.01) * _randn(shape, key, seed)
        elif leaf == "weight" and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            fan_out = shape[0] * shape[2] * shape[3]
            std = 0.1 * math.sqrt(2.0 / (fan_in + fan_out))
            if "conv_first" in key or "trunk_conv1" in key:
                std = math.sqrt(1.0 / (3.0 * fan_in))  # ~ torch default conv init
            out[key] = std * _randn(shape, key, seed)
        elif leaf == "bias":
            out[key] = 0.01 * _randn(shape, key, seed)
        else:
            raise KeyError("synthetic_state_dict: no rule for key {}".format(key))
    return out


# stress fixtures (tests/golden/*_stress.pt, oracle/make_golden.py): flow depth and scales chosen so that the coupling
# outputs reach |h| ~ 0.8 .. 0.9 (a trained flow's order of magnitude) while the un-clamped HR stays O(10)
STRESS = {
    "sr_x4_stress": {"cfg": "sr_x4", "K": 8, "after": [4, 4], "s_weight": 100.0, "s_bias": 5.0, "s_prior_mean": 5.0},
    "sr_x8_stress": {"cfg": "sr_x8", "K": 8, "after": [4, 4, 4], "s_weight": 100.0, "s_bias": 5.0, "s_prior_mean": 5.0},
    "rescaling_x4_stress": {"cfg": "rescaling_x4", "K": None, "after": None, "s_weight": 8.0, "s_bias": 3.0, "s_prior_mean": 3.0},
}


def stress_state_dict(sd, s_weight, s_bias=1.0, s_prior=1.0, s_prior_mean=1.0):
    """Stress variant of a synthetic state dict: the coupling sub-nets' last conv (zero-initialised in the reference,
    0.002 * N(0,1) here) scaled by ``s_weight`` (its bias / logs by ``s_bias``) so that the coupling outputs h become
    O(1) like a trained flow's, the prior conv's weights by ``s_prior`` and its MEAN rows (output channels 0::2,
    ConditionalFlow.py:62) by ``s_prior_mean`` on top, so that the encoder feature's rounding reaches z through an O(1)
    prior mean as well.  Only usable on a SHALLOW flow
    (options.shrink_config): 52 steps with O(1) couplings diverge."""
    out = {}
    for k, v in sd.items():
        leaf = k.rsplit(".", 1)[-1]
        if (".affine.f.conv3." in k and "actnorm" not in k) or ".affine.f.conv5." in k:
            out[k] = v * (s_weight if leaf == "weight" else s_bias)
        elif k.endswith("_condFlow.f.weight"):
            w = v * s_prior
            if s_prior_mean != 1.0:
                w = w.clone()
                w[0::2] = w[0::2] * s_prior_mean
            out[k] = w
        else:
            out[k] = v
    return out


def fingerprint(sd):
    """Order-independent checksum of a state dict (detects RNG drift between boxes)."""
    acc = 0.0
    for k in sorted(sd.keys()):
        t = sd[k].double()
        acc += float(t.sum()) + 0.5 * float((t * t).sum())
    return acc


def synthetic_lr(batch, h, w, seed=0):
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    return torch.rand(batch, 3, h, w, generator=g, dtype=torch.float32)


def synthetic_hr(batch, H, W, seed=0):
    g = torch.Generator(device="cpu")
    g.manual_seed(2000 + seed)
    return torch.rand(batch, 3, H, W, generator=g, dtype=torch.float32)


def synthetic_noise(shapes, seed=123):
    """Unit normal draws, one tensor per shape, from a dedicated generator."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return [torch.randn(*s, generator=g, dtype=torch.float32) for s in shapes]
