"""Host-logic tests (no GPU): the launch plan, interpreted on the CPU by tests/plan_emulator.py
with the engine's packed weights, must reproduce the oracle (and therefore the reference
goldens) for every network and both directions."""
import pytest
import torch

from hcflow_b200 import plan as P
from hcflow_b200 import synth
from oracle import hcflow_oracle as orc
from tests.helpers import is_sr, load_golden, maxabs, net_and_weights
from tests.plan_emulator import Emulator

CASES = ["sr_x4", "sr_x8", "rescaling_x4"]


def _inputs(g, opt):
    B, h, w, heat = g["B"], g["h"], g["w"], g["heat"]
    s = opt["scale"]
    lr = synth.synthetic_lr(B, h, w)
    hr = synth.synthetic_hr(B, h * s, w * s)
    unit = synth.synthetic_noise(orc.noise_shapes(opt, B, h, w, is_sr(opt)))
    return lr, hr, [heat * e for e in unit]


@pytest.mark.parametrize("cfg", CASES)
def test_reverse_plan_matches_golden(cfg):
    g = load_golden(cfg)
    opt, net, sd = net_and_weights(cfg)
    assert abs(synth.fingerprint(sd) - g["fingerprint"]) < 1e-6 * abs(g["fingerprint"])
    lr, hr, eps = _inputs(g, opt)
    plan = P.build_plan(net, "reverse", g["B"], g["h"], g["w"])
    assert [tuple(s) for s in plan.noise_shapes] == [tuple(e.shape[1:]) for e in eps]
    em = Emulator(net, plan)
    out = em.run(lr=lr, **{"eps{}".format(i): e for i, e in enumerate(eps)})
    assert maxabs(out["hr_raw"], g["inv_raw"]) < 2e-4
    assert maxabs(out["hr"], g["inv_hr"]) < 2e-4


@pytest.mark.parametrize("cfg", CASES)
def test_forward_plan_matches_golden(cfg):
    g = load_golden(cfg)
    opt, net, sd = net_and_weights(cfg)
    lr, hr, eps = _inputs(g, opt)
    plan = P.build_plan(net, "forward", g["B"], g["h"], g["w"])
    em = Emulator(net, plan)
    if is_sr(opt):
        dq = torch.rand(hr.shape, generator=torch.Generator().manual_seed(77), dtype=torch.float32)
        out = em.run(hr=hr, lr=lr, dequant=dq)
        assert maxabs(out["z_raw"], g["fwd_z"]) < 2e-4
        H, W = hr.shape[2], hr.shape[3]
        import math
        nll = float(((-em.logdet) / (math.log(2.0) * H * W)).mean())
        assert abs(nll - float(g["fwd_nll"])) < 1e-4 * abs(float(g["fwd_nll"]))
        # logdet (without the dirac term): recompute through the oracle for the split
        _, _, _, ld = orc.sr_forward(hr, lr, sd, opt, dq)
        dirac = orc.gaussian_logp(lr, -torch.ones_like(lr) * 6, out["fake_lr"])
        rel = ((em.logdet - dirac.double() - ld.double()).abs() / ld.double().abs()).max()
        assert float(rel) < 1e-5
    else:
        out = em.run(hr=hr)
        assert maxabs(out["fake_lr"], g["fwd_fake_lr"]) < 2e-4
        assert maxabs(out["fake_z1"], g["fwd_z1"]) < 2e-3
        assert maxabs(out["fake_z2"], g["fwd_z2"]) < 2e-3


@pytest.mark.parametrize("cfg", CASES)
def test_rewritten_reverse_plan_matches_golden(cfg):
    """The engine-level rewrites of the tensor-core modes (hcflow_b200/rewrite.py: materialised up-sampled segments,
    shared conditioning convs + pre-activation addend, FlowStep tails fused into the sub-net's last conv), interpreted
    on the CPU, must still reproduce the reference goldens -- and must actually have rewritten something."""
    from hcflow_b200 import rewrite
    g = load_golden(cfg)
    opt, net, sd = net_and_weights(cfg)
    lr, hr, eps = _inputs(g, opt)
    plan = P.build_plan(net, "reverse", g["B"], g["h"], g["w"])
    ops, extra = rewrite.rewrite_ops(plan.ops, "tf32x3")   # (the fp16 modes apply everything but the conv pairing)
    assert not any(isinstance(o, P.ConvOp) and o.raw2 is not None for o in rewrite.rewrite_ops(plan.ops, "f16x3")[0])
    assert rewrite.rewrite_ops(plan.ops, "fp32")[0] == list(plan.ops)
    n_step = sum(isinstance(o, P.StepOp) for o in plan.ops)
    n_left = sum(isinstance(o, P.StepOp) for o in ops)
    n_fused = sum(isinstance(o, P.ConvOp) and o.step is not None for o in ops)
    assert n_fused + n_left == n_step
    assert not any(isinstance(o, P.ConvOp) and any(up for _, up in o.segs) for o in ops)
    if is_sr(opt):   # (the rescaling net's dense blocks grow by 16 channels: not paired)
        assert any(isinstance(o, P.ConvOp) and o.raw2 is not None for o in ops)
        assert n_fused > 0 and any(isinstance(o, P.ConvOp) and o.pre is not None for o in ops)
        assert any(isinstance(o, P.LayoutOp) and o.variant == "upsample" for o in ops)
    em = Emulator(net, plan, ops=ops, extra_bufs=extra)
    out = em.run(lr=lr, **{"eps{}".format(i): e for i, e in enumerate(eps)})
    assert maxabs(out["hr_raw"], g["inv_raw"]) < 2e-4
