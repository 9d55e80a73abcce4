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

CASES = ["sr_x4", "sr_x8", "rescaling_x4", "sr_x4_stress", "sr_x8_stress", "rescaling_x4_stress"]


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
    assert net.noise_shapes(g["B"], g["h"], g["w"]) == [tuple(e.shape) for e in eps]   # the public helper bench.py uses
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
        # (fp64: the dirac term is a sum of ~1e4-sized squares, its fp32 rounding alone is ~0.5 absolute)
        dirac = orc.gaussian_logp(lr.double(), -torch.ones_like(lr).double() * 6, out["fake_lr"].double())
        rel = ((em.logdet - dirac - ld.double()).abs() / ld.double().abs()).max()
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


def test_fp16_chain_layout_of_the_x4_encoder_and_flowstep_chains():
    """rewrite.chain16_layout (the data-flow decisions behind hcf_conv_chain16_create) on the real x4 plan: the RDB
    growth channels exist only as an fp16 hi plane, the residual stream keeps fp32 + hi + lo, inputs produced outside
    the chain are converted once, and a fused FlowStep leaves its z1 copy where the next step's first conv reads it."""
    from hcflow_b200 import _lib, rewrite
    assert (rewrite.OUT_F32, rewrite.OUT_HI, rewrite.OUT_LO) == (_lib.OUT_F32, _lib.OUT_HI, _lib.OUT_LO)
    opt, net, sd = net_and_weights("sr_x4")
    plan = P.build_plan(net, "reverse", 2, 16, 16)
    ops, _ = rewrite.rewrite_ops(plan.ops, "f16x3", flowchain=False)   # (the generic chained-conv lowering of FlowSteps)
    # runs of consecutive convs on one grid = the chains the engine builds
    runs, cur = [], []
    for i, o in enumerate(ops):
        if isinstance(o, P.ConvOp) and (not cur or (ops[cur[-1]].H, ops[cur[-1]].W) == (o.H, o.W)) and (not cur or cur[-1] == i - 1):
            cur.append(i)
        else:
            if len(cur) > 1:
                runs.append(cur)
            cur = [i] if isinstance(o, P.ConvOp) else []
    if len(cur) > 1:
        runs.append(cur)
    assert len(runs) == 6          # per level: encoder, conditional FlowSteps, main FlowSteps

    def passes_of(o):              # Engine._passes_for / _split_channels of the default mode
        one = o.tag in ("enc.rdb.conv1", "enc.rdb.conv2", "enc.rdb.conv3", "enc.rdb.conv4")
        ps = 1 if one else 3
        sp = 0 if ps == 1 else (o.res1.C if o.tag == "enc.rdb.conv5" else -1)
        return ps, sp
    for run in runs:
        chain = [ops[i] for i in run]
        ps, sp = zip(*[passes_of(o) for o in chain])
        lay = rewrite.chain16_layout(chain, list(ps), list(sp), ops[run[-1] + 1:])
        assert lay is not None
        for o, fl in zip(chain, lay["flags"]):
            if o.tag in ("enc.rdb.conv1", "enc.rdb.conv2", "enc.rdb.conv3", "enc.rdb.conv4"):
                assert fl == rewrite.OUT_HI, (o.tag, fl)                       # growth channels: hi plane only
            if o.tag == "enc.rdb.conv5":
                assert fl == rewrite.OUT_F32 | rewrite.OUT_HI | rewrite.OUT_LO  # residual stream
            if o.tag == "fcn.ucond":
                assert fl == rewrite.OUT_F32                                    # read back as a pre-activation addend
            if o.tag in ("fcn.conv1", "fcn.conv2"):
                assert fl == rewrite.OUT_HI | rewrite.OUT_LO                    # the sub-nets run split
        if chain[0].tag == "enc.conv_first":
            assert len(lay["external"]) == len(chain[0].segs)                   # only the first conv's inputs come from outside
            assert all(need_lo for _, need_lo, _ in lay["external"].values())  # conv_first is a split layer
        if any(o.step is not None for o in chain):
            tg = [t for o, t in zip(chain, lay["step_target"]) if o.step is not None]
            assert all(t is not None for t in tg[:-1])                          # every step but the last feeds a next conv1
            lo = [w for o, w in zip(chain, lay["step_lo"]) if o.step is not None]
            assert all(lo[:-1])                                                 # ... which reads hi + lo
            assert all(need_lo for _, need_lo, _ in lay["external"].values())


@pytest.mark.parametrize("cfg", ["sr_x4", "sr_x8"])
def test_rewritten_forward_plan_matches_golden(cfg):
    """The forward (NLL) plan as the tensor-core modes rewrite it (shared conditioning convs, materialised up-sampled
    segments; forward FlowSteps stay separate ops) still reproduces the reference's z and log-det."""
    import math
    from hcflow_b200 import rewrite
    g = load_golden(cfg)
    opt, net, sd = net_and_weights(cfg)
    lr, hr, eps = _inputs(g, opt)
    plan = P.build_plan(net, "forward", g["B"], g["h"], g["w"])
    ops, extra = rewrite.rewrite_ops(plan.ops, "tf32x3")
    assert any(isinstance(o, P.ConvOp) and o.pre is not None for o in ops)
    assert not any(isinstance(o, P.ConvOp) and o.step is not None for o in ops)       # only inverse steps are fused
    assert sum(isinstance(o, P.StepOp) for o in ops) == sum(isinstance(o, P.StepOp) for o in plan.ops)
    em = Emulator(net, plan, ops=ops, extra_bufs=extra)
    dq = torch.rand(hr.shape, generator=torch.Generator().manual_seed(77), dtype=torch.float32)
    out = em.run(hr=hr, lr=lr, dequant=dq)
    assert maxabs(out["z_raw"], g["fwd_z"]) < 2e-4
    H, W = hr.shape[2], hr.shape[3]
    nll = float(((-em.logdet) / (math.log(2.0) * H * W)).mean())
    assert abs(nll - float(g["fwd_nll"])) < 1e-4 * abs(float(g["fwd_nll"]))


@pytest.mark.parametrize("direction", ["reverse", "forward"])
@pytest.mark.parametrize("cfg", ["sr_x4", "sr_x8", "rescaling_x4", "sr_x4_stress"])
def test_flowstep_chains_are_grouped_and_still_reproduce_the_golden(cfg, direction):
    """rewrite.group_flowsteps (fp16 modes): every FlowStep with an FCN sub-net over z1 becomes part of a FlowChainOp --
    the work list of the fused-FlowStep kernel (csrc/flowstep_tc.cu) -- and the plan still computes the reference's
    outputs (the emulator interprets a chain through the ops it replaces)."""
    import math
    from hcflow_b200 import rewrite
    g = load_golden(cfg)
    opt, net, sd = net_and_weights(cfg)
    lr, hr, eps = _inputs(g, opt)
    plan = P.build_plan(net, direction, g["B"], g["h"], g["w"])
    ops, extra = rewrite.rewrite_ops(plan.ops, "f16x3")
    chains = [o for o in ops if isinstance(o, P.FlowChainOp)]
    assert chains and all(c.forward == (direction == "forward") for c in chains)
    n_steps_plan = sum(isinstance(o, P.StepOp) and o.variant in ("inverse", "forward_coupling") for o in plan.ops)
    n_in_chains = sum(len(c.steps) for c in chains)
    if cfg.startswith("sr_x4"):
        assert n_in_chains == n_steps_plan                      # x4: every FlowStep runs in the fused kernel
        assert not any(isinstance(o, P.StepOp) for o in ops)
    else:
        assert 0 < n_in_chains < n_steps_plan                   # x8: C = 48 steps; rescaling: dense sub-nets stay generic
    for c in chains:
        assert c.z.C in rewrite.FLOWCHAIN_C and c.n_pass == c.z.C // 2
        for c1, c2, c3, tail, head in c.steps:
            assert c1.segs[0][0] == c.z.sub(0, c.n_pass) and (head is not None) == c.forward
    em = Emulator(net, plan, ops=ops, extra_bufs=extra)
    if direction == "reverse":
        out = em.run(lr=lr, **{"eps{}".format(i): e for i, e in enumerate(eps)})
        assert maxabs(out["hr_raw"], g["inv_raw"]) < 2e-4
    elif is_sr(opt):
        dq = torch.rand(hr.shape, generator=torch.Generator().manual_seed(77), dtype=torch.float32)
        out = em.run(hr=hr, lr=lr, dequant=dq)
        assert maxabs(out["z_raw"], g["fwd_z"]) < 2e-4
        nll = float(((-em.logdet) / (math.log(2.0) * hr.shape[2] * hr.shape[3])).mean())
        assert abs(nll - float(g["fwd_nll"])) < 1e-4 * abs(float(g["fwd_nll"]))
    else:
        out = em.run(hr=hr)
        assert maxabs(out["fake_lr"], g["fwd_fake_lr"]) < 2e-4
