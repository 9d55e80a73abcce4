"""CPU emulation of the tensor-core modes' operand rounding, through the oracle (TEST / ANALYSIS TOOL).

    python tools/emulate_precision.py [--cfg sr_x4] [--stress S] [--K 4] [--after 2,2] [--policy f16x3 ...]

Every conv of the oracle is replaced by the arithmetic the fp16 chains perform: operands rounded to fp16
(round-to-nearest, 11 significant bits), products accumulated in fp32; a SPLIT conv uses hi + lo / 2048 on both
operands (a_hi b_hi + (a_hi b_lo + a_lo b_hi) / 2048, the lo x lo term is dropped as on the device).  Which convs
are split is a policy (name -> predicate on the state-dict key), so that candidate policies can be compared against
the exact fp32 oracle before any GPU time is spent.  Prints max-abs / mean errors on the un-clamped HR (inverse) and
on z / log-det (forward).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as TF  # noqa: E402

from hcflow_b200 import options as popt, synth  # noqa: E402
from hcflow_b200.arch import build_net  # noqa: E402
from oracle import hcflow_oracle as orc  # noqa: E402


def _hi(t):
    return t.half().float()


def _lo(t, hi):
    return ((t - hi) * 2048.0).half().float() / 2048.0


def conv_emulated(x, w, kind, split_in=None, **kw):
    """kind: 'exact' | 'one' | 'split' | 'split_in' (split on the first split_in input channels only)."""
    if kind == "exact":
        return TF.conv2d(x, w, None, **kw)
    xh, wh = _hi(x), _hi(w)
    if kind == "one":
        return TF.conv2d(xh, wh, None, **kw)
    xl, wl = _lo(x, xh), _lo(w, wh)
    if kind == "split_in":
        xl = xl.clone()
        wl = wl.clone()
        xl[:, split_in:] = 0
        wl[:, split_in:] = 0
    return TF.conv2d(xh, wh, None, **kw) + (TF.conv2d(xh, wl, None, **kw) + TF.conv2d(xl, wh, None, **kw))


def _rdb_kind(k):
    if ".RDB" in k:
        return "split_in" if ".conv5." in k else "one"
    return None


def _pol_f16x3(k, fcn_split=()):
    r = _rdb_kind(k)
    if r:
        return r
    if _is_fcn(k):
        return "split" if any(".conv{}.".format(i) in k for i in fcn_split) else "one"
    return "split"          # conv_first, trunk_conv, prior conv, dense sub-nets


POLICIES = {
    "f16": lambda k: "one",
    "f16x3": lambda k: _pol_f16x3(k),
    "f16x3_fcn13": lambda k: _pol_f16x3(k, (1, 3)),      # FCN conv1 and conv3 split (conv2 1x1 stays one pass)
    "f16x3_fcn3": lambda k: _pol_f16x3(k, (3,)),
    "f16x3_fcn": lambda k: _pol_f16x3(k, (1, 2, 3)),
    "split_all": lambda k: "split",
    # what if conv5 of the RDBs ran one pass (everything else as f16x3_fcn)?
    "fcn_split_conv5_one": lambda k: ("one" if ".RDB" in k else _pol_f16x3(k, (1, 2, 3))),
}
_FCN_KEYS = set()


def _is_fcn(k):
    return k.rsplit(".conv", 1)[0] in _FCN_KEYS


class FShim:
    """stands in for torch.nn.functional inside the oracle module"""

    def __init__(self, sd, policy):
        self.ids = {id(v): k for k, v in sd.items()}
        self.policy = policy
        self.stats = {}

    def __getattr__(self, name):
        return getattr(TF, name)

    def conv2d(self, x, w, b=None, **kw):
        key = self.ids.get(id(w))
        kind = "exact" if key is None or "haar" in key else self.policy(key)
        self.stats[kind] = self.stats.get(kind, 0) + 1
        y = conv_emulated(x, w, kind, split_in=64, **kw)
        if b is not None:
            y = y + b.view(1, -1, 1, 1)
        return y


def scale_zero_convs(sd, s, s_bias=None, s_prior=1.0):
    """multiply the 'zero-init' conv parameters of synth.synthetic_state_dict: the coupling sub-nets' last conv by s
    (weights) / s_bias (bias, logs), the prior conv's weights by s_prior"""
    return synth.stress_state_dict(sd, s, s_bias if s_bias is not None else 1.0, s_prior)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="sr_x4")
    ap.add_argument("--stress", type=float, default=1.0)
    ap.add_argument("--stress-bias", type=float, default=None)
    ap.add_argument("--stress-prior", type=float, default=1.0)
    ap.add_argument("--stress-prior-mean", type=float, default=1.0, help="scale only the MEAN rows (0::2) of the prior conv")
    ap.add_argument("--K", type=int, default=None)
    ap.add_argument("--after", default=None)
    ap.add_argument("--nb", default=None)
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--hw", type=int, default=12)
    ap.add_argument("--heat", type=float, default=0.8)
    ap.add_argument("--policy", nargs="*", default=["f16x3"])
    ap.add_argument("--forward", action="store_true")
    args = ap.parse_args()
    opt = popt.load_config(args.cfg)
    if args.K is not None or args.after or args.nb:
        opt = popt.shrink_config(opt, K=args.K, after=[int(a) for a in args.after.split(",")] if args.after else None,
                                 rrdb_nb=[int(a) for a in args.nb.split(",")] if args.nb else None)
    net = build_net(opt)
    sd = synth.synthetic_state_dict(net.state_dict(), seed=1)
    sd = scale_zero_convs(sd, args.stress, args.stress_bias, args.stress_prior)
    if args.stress_prior_mean != 1.0:
        for k in list(sd):
            if k.endswith("_condFlow.f.weight"):
                w = sd[k].clone()
                w[0::2] *= args.stress_prior_mean
                sd[k] = w
    for k in sd:
        if k.endswith(".conv1.actnorm.bias"):
            _FCN_KEYS.add(k[:-len(".conv1.actnorm.bias")])
    SR = opt["network_G"]["which_model_G"] == "HCFlowNet_SR"
    B, h = args.B, args.hw
    lr = synth.synthetic_lr(B, h, h, seed=0)
    unit = synth.synthetic_noise(orc.noise_shapes(opt, B, h, h, SR), seed=123)
    eps = [args.heat * e for e in unit]
    s = opt["scale"]
    hr = synth.synthetic_hr(B, h * s, h * s, seed=0)
    with torch.no_grad():
        if SR:
            _, want = orc.sr_reverse(lr, sd, opt, eps)
        else:
            _, want = orc.rescaling_reverse(lr, sd, opt, eps)
        # coupling strength of the fixture: max |h| over every coupling sub-net call of this pass
        hmax = []
        fcn0, dense0 = orc.fcn, orc.dense_block

        def rec(fn):
            def wrapped(*a, **k):
                y = fn(*a, **k)
                hmax.append(float(y.abs().max()))
                return y
            return wrapped
        orc.fcn, orc.dense_block = rec(fcn0), rec(dense0)
        try:
            (orc.sr_reverse if SR else orc.rescaling_reverse)(lr, sd, opt, eps)
        finally:
            orc.fcn, orc.dense_block = fcn0, dense0
        print(json.dumps({"cfg": args.cfg, "stress": args.stress, "hr_range": [float(want.min()), float(want.max())],
                          "h_absmax": max(hmax), "h_absmax_median": sorted(hmax)[len(hmax) // 2]}))
        if args.forward and SR:
            dq = torch.rand(hr.shape, generator=torch.Generator().manual_seed(77))
            _, nll_w, z_w, ld_w = orc.sr_forward(hr, lr, sd, opt, dq)
        for pol in args.policy:
            shim = FShim(sd, POLICIES[pol])
            orc.F = shim
            try:
                if SR:
                    _, got = orc.sr_reverse(lr, sd, opt, eps)
                else:
                    _, got = orc.rescaling_reverse(lr, sd, opt, eps)
                res = {"policy": pol, "inv_max": float((got - want).abs().max()),
                       "inv_mean": float((got - want).abs().mean()), "convs": dict(shim.stats)}
                if args.forward and SR:
                    _, nll, z, ld = orc.sr_forward(hr, lr, sd, opt, dq)
                    res["fwd_z_max"] = float((z - z_w).abs().max())
                    res["fwd_logdet_rel"] = float(((ld - ld_w).abs() / ld_w.abs()).max())
            finally:
                orc.F = TF
            print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
