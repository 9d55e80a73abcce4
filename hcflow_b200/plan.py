"""Launch-plan IR for the HCFlow hot path.

``build_plan`` walks the parameter tree (modules.py) once for a given direction and input
size and emits a flat list of ops over named NHWC buffers.  The plan is pure data: the
engine (engine.py) lowers it to C-ABI calls on device pointers; tests/plan_emulator.py
interprets the same ops with torch on the CPU to check the host logic without a GPU.

Reference control flow mirrored (codes/models/modules/):
  reverse:  FlowNet_SR_x4.py:106-123, FlowNet_SR_x8.py:123-144, FlowNet_Rescaling_x4.py:113-128
  forward:  FlowNet_SR_x4.py:84-101,  FlowNet_SR_x8.py:91-118,  FlowNet_Rescaling_x4.py:90-108
  FlowStep.py:40-64, ConditionalFlow.py:44-110, Basic.py:349-398, 426-447
torch.cat / Split / dense-concat / F.interpolate never materialise: they are channel-slice
views and conv input segments.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from . import modules as M


@dataclass(frozen=True)
class Buf:
    name: str
    H: int
    W: int
    C: int


@dataclass(frozen=True)
class View:
    buf: Buf
    off: int
    C: int

    def sub(self, off, C):
        assert off + C <= self.C
        return View(self.buf, self.off + off, C)


@dataclass
class ConvOp:
    kind = "conv"
    H: int
    W: int
    segs: List[Tuple[View, int]]  # (view, up_shift)
    ks: int
    cout: int
    weight: str                   # key into the weight store (prep.py packs it)
    bias: Optional[str]
    scale: Optional[str]
    act: int
    out: View
    out2: Optional[View] = None
    res1: Optional[View] = None
    alpha1: float = 1.0
    res2: Optional[View] = None
    alpha2: float = 1.0
    tag: str = ""
    # engine-level rewrites only (engine.py, shared conditioning of the coupling sub-nets): `weight` may be a
    # tuple of keys (concatenated along Cout), `w_in` an input-channel slice of it, `pre` an fp32 view added to
    # the accumulator before bias / scale / activation
    w_in: Optional[Tuple[int, int]] = None
    pre: Optional[View] = None
    step: Optional[object] = None   # a StepOp("inverse") executed by this conv's epilogue (fused FlowStep tail)
    raw2: Optional[View] = None     # cout == 64: accumulator columns [32, 64) stored raw (fp32) here, [0, 32) -> out


@dataclass
class StepOp:
    kind = "step"
    variant: str            # "inverse" | "forward_head" | "forward_coupling"
    H: int
    W: int
    z: View
    h: Optional[View]
    mode: str               # "affine" | "shift_first3"
    n_pass: int
    w: Optional[str]        # mixing matrix key (W or W^-1), None = no permutation
    an_scale: Optional[str]
    an_bias: Optional[str]
    tag: str = ""


@dataclass
class PriorOp:
    kind = "prior"
    variant: str            # "sample" | "logp" | "standardize"
    H: int
    W: int
    h: View
    z: View
    atan_logscale: bool
    eps_index: int = -1     # which noise tensor (sample)
    out_name: str = ""      # NCHW output name (standardize)


@dataclass
class LayoutOp:
    kind = "layout"
    variant: str            # "ingest" | "egress" | "squeeze" | "unsqueeze" | "haar_fwd" | "haar_inv" | "copy"
    H: int                  # low-res size for squeeze-like ops, tensor size for ingest/egress
    W: int
    C: int                  # channels on the high-res side (squeeze-like) or tensor channels
    src: object             # View, or external tensor name (ingest)
    dst: object             # View, or output tensor name (egress)
    post: int = 0           # egress: 0 raw, 1 clamp, 2 quantize
    noise: Optional[str] = None  # ingest: external NCHW noise name
    noise_scale: float = 0.0


@dataclass
class DiracLogpOp:
    kind = "dirac_logp"
    x_name: str             # NCHW output tensor (quantised fake LR)
    mean_name: str          # external NCHW tensor (lr)
    logs: float
    n: int


@dataclass
class FlowChainOp:
    """engine-level rewrite only (rewrite.group_flowsteps): consecutive FlowSteps with FCN sub-nets on the same z,
    executed by ONE fused-FlowStep launch (csrc/flowstep_tc.cu).  ``steps``: per step (conv1, conv2, conv3, tail StepOp,
    head StepOp or None); ``orig``: the ops it replaces, in order (tests/plan_emulator.py interprets those)."""
    kind = "flowchain"
    H: int
    W: int
    z: View
    n_pass: int
    forward: bool
    steps: list
    orig: list
    tag: str = ""


@dataclass
class Plan:
    direction: str
    sr: bool
    B: int
    h: int
    w: int
    ops: list = field(default_factory=list)
    bufs: dict = field(default_factory=dict)
    outputs: dict = field(default_factory=dict)       # name -> (C, H, W)
    noise_shapes: list = field(default_factory=list)  # per draw (C, H, W)
    logdet_const: float = 0.0                          # data-independent log-det (per image)
    uses_logdet: bool = False

    def buf(self, name, H, W, C):
        b = self.bufs.get(name)
        if b is None:
            b = Buf(name, H, W, C)
            self.bufs[name] = b
        assert (b.H, b.W, b.C) == (H, W, C), (name, b, H, W, C)
        return b

    def view(self, name, H, W, C):
        return View(self.buf(name, H, W, C), 0, C)


ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2


def _layers(flow):
    """[(index, module)] of flow.layers."""
    return list(enumerate(flow.layers))


def _level_info(flow):
    """Per level: (C before split, n_split, [step layer indices], squeeze layer index)."""
    info = []
    cur = None
    for i, lay in _layers(flow):
        if isinstance(lay, (M.SqueezeLayer, M.HaarDownsampling)):
            cur = {"squeeze": i, "haar": isinstance(lay, M.HaarDownsampling), "steps": []}
        elif isinstance(lay, M.FlowStep):
            cur["steps"].append(i)
            cur["C"] = lay.in_channels
        elif isinstance(lay, M.Split):
            cur["n_split"] = lay.num_channels_split
            if "C" not in cur:
                cur["C"] = None
            info.append(cur)
    return info


def _emit_encoder(plan, pre, cf_mod, segs, H, W, lvl):
    """RRDB conditional encoder (ConditionalFlow.py:99-110). Returns the feature view."""
    nf, gc = cf_mod.nf, cf_mod.gc
    wide = nf + 4 * gc
    ring = [plan.view("enc{}_ring{}".format(lvl, i), H, W, wide) for i in range(4)]
    ff = plan.view("enc{}_first".format(lvl), H, W, nf)
    cf = plan.view("enc{}_feat".format(lvl), H, W, cf_mod.cond_channels)
    plan.ops.append(ConvOp(H, W, segs, 3, nf, pre + ".conv_first.weight", pre + ".conv_first.bias", None,
                           ACT_NONE, ff, out2=ring[0].sub(0, nf), tag="enc.conv_first"))
    r = 0
    trunks = [("RRDB_trunk0", cf_mod.nb[0]), ("RRDB_trunk1", cf_mod.nb[1])]
    for t, (tname, count) in enumerate(trunks):
        for j in range(count):
            q0 = (3 * r) % 4
            for d in range(3):
                cur, nxt = ring[(q0 + d) % 4], ring[(q0 + d + 1) % 4]
                p = "{}.{}.{}.RDB{}".format(pre, tname, j, d + 1)
                for i in range(1, 5):
                    cin = nf + (i - 1) * gc
                    plan.ops.append(ConvOp(H, W, [(cur.sub(0, cin), 0)], 3, gc, "{}.conv{}.weight".format(p, i),
                                           "{}.conv{}.bias".format(p, i), None, ACT_LRELU, cur.sub(cin, gc),
                                           tag="enc.rdb.conv{}".format(i)))
                op = ConvOp(H, W, [(cur.sub(0, wide), 0)], 3, nf, p + ".conv5.weight", p + ".conv5.bias", None,
                            ACT_NONE, nxt.sub(0, nf), res1=cur.sub(0, nf), alpha1=0.2, tag="enc.rdb.conv5")
                if d == 2:
                    op.res2, op.alpha2 = ring[q0].sub(0, nf), 0.2
                    if cf_mod.SR and t == 0 and j == count - 1:
                        op.out2 = cf.sub(0, nf)
                plan.ops.append(op)
            r += 1
    last = ring[(3 * r) % 4].sub(0, nf)
    if cf_mod.SR and cf_mod.nb[0] == 0:
        raise NotImplementedError("RRDB_nb[0] == 0 with SR features")
    out = cf.sub(nf, nf) if cf_mod.SR else cf.sub(0, nf)
    plan.ops.append(ConvOp(H, W, [(last, 0)], 3, nf, pre + ".trunk_conv1.weight", pre + ".trunk_conv1.bias", None,
                           ACT_NONE, out, res1=ff, alpha1=1.0, tag="enc.trunk_conv"))
    return cf


def _emit_subnet(plan, pre, aff, z, cond, H, W, lvl):
    """Coupling sub-net f(cat(z1, u)) -> h view (AffineCouplings.py:31,68; Basic.py:349-356,442-447)."""
    z1 = z.sub(0, aff.n_pass) if aff.mode == "affine" else z.sub(3, aff.n_pass)
    if z1.off % 4 != 0 or z1.buf.C % 4 != 0:
        # TMA needs 16-byte aligned views: stage the conditioning slice once per step
        z1s = plan.view("z1s_l{}_c{}".format(lvl, z1.C), H, W, (z1.C + 3) // 4 * 4).sub(0, z1.C)
        plan.ops.append(LayoutOp("copy", H, W, z1.C, z1, z1s))
        z1 = z1s
    segs = [(z1, 0)] + ([(cond, 0)] if cond is not None else [])
    f = aff.f
    fp = pre + ".affine.f"
    hout = plan.view("h3_l{}_c{}".format(lvl, f.cout), H, W, (f.cout + 3) // 4 * 4).sub(0, f.cout)
    if f.kind == "FCN":
        h1 = plan.view("h1_l{}".format(lvl), H, W, f.hidden)
        h2 = plan.view("h2_l{}".format(lvl), H, W, f.hidden)
        plan.ops.append(ConvOp(H, W, segs, 3, f.hidden, fp + ".conv1.weight", fp + ".conv1.actnorm.bias",
                               fp + ".conv1.actnorm.logs#exp", ACT_RELU, h1, tag="fcn.conv1"))
        plan.ops.append(ConvOp(H, W, [(h1, 0)], 1, f.hidden, fp + ".conv2.weight", fp + ".conv2.actnorm.bias",
                               fp + ".conv2.actnorm.logs#exp", ACT_RELU, h2, tag="fcn.conv2"))
        plan.ops.append(ConvOp(H, W, [(h2, 0)], 3, f.cout, fp + ".conv3.weight", fp + ".conv3.bias",
                               fp + ".conv3.logs#exp3", ACT_NONE, hout, tag="fcn.conv3"))
    else:
        g = plan.view("dense_l{}".format(lvl), H, W, 4 * f.gc)
        for i in range(1, 5):
            s = list(segs) + ([(g.sub(0, (i - 1) * f.gc), 0)] if i > 1 else [])
            plan.ops.append(ConvOp(H, W, s, 3, f.gc, "{}.conv{}.weight".format(fp, i), "{}.conv{}.bias".format(fp, i),
                                   None, ACT_LRELU, g.sub((i - 1) * f.gc, f.gc), tag="dense.conv{}".format(i)))
        plan.ops.append(ConvOp(H, W, list(segs) + [(g, 0)], 3, f.cout, fp + ".conv5.weight", fp + ".conv5.bias",
                               None, ACT_NONE, hout, tag="dense.conv5"))
    return hout


def _emit_step(plan, pre, step, z, cond, H, W, lvl, reverse):
    aff = step.affine
    has_perm = step.permute is not None
    if reverse:
        # AffineCoupling3shift ignores u on its reverse shift branch (AffineCouplings.py:154)
        c = None if aff.mode == "shift_first3" else cond
        h = _emit_subnet(plan, pre, aff, z, c, H, W, lvl)
        plan.ops.append(StepOp("inverse", H, W, z, h, aff.mode, aff.n_pass,
                               pre + ".permute.weight#inv" if has_perm else None,
                               pre + ".actnorm.logs#expneg", pre + ".actnorm.bias#vec", tag=pre))
    else:
        plan.ops.append(StepOp("forward_head", H, W, z, None, aff.mode, aff.n_pass,
                               pre + ".permute.weight#mat" if has_perm else None,
                               pre + ".actnorm.logs#exppos", pre + ".actnorm.bias#vec", tag=pre))
        h = _emit_subnet(plan, pre, aff, z, cond, H, W, lvl)
        plan.ops.append(StepOp("forward_coupling", H, W, z, h, aff.mode, aff.n_pass, None, None, None, tag=pre))
        plan.logdet_terms.append((pre, has_perm, H * W))


def _cond_view(plan, lvl, H, W, a):
    """The conditional flow's variable `a` = z[:, n_split:].  TMA needs 16-byte aligned views, so
    when the slice starts at a channel offset that is not a multiple of 4 the conditional steps
    run on a private buffer (copied to / from the slice once per level)."""
    if (a.off % 4 == 0) and (a.buf.C % 4 == 0):
        return a
    return plan.view("acond_l{}".format(lvl), H, W, (a.C + 3) // 4 * 4).sub(0, a.C)


def build_plan(net, direction, B, h, w):
    """net: HCFlowNet_SR / HCFlowNet_Rescaling (arch.py); direction: "reverse" | "forward";
    (h, w): LR size.  Returns a Plan."""
    flow = net.flow
    sr = flow.SR
    L = flow.L
    info = _level_info(flow)
    assert len(info) == L
    plan = Plan(direction, sr, B, h, w)
    plan.logdet_terms = []
    size = [(h << (L - 1 - l), w << (L - 1 - l)) for l in range(L)]
    Cl = []
    c = 3
    for l in range(L):
        c *= 4
        Cl.append(c)
        c = info[l]["n_split"]
    zb = [plan.view("z{}".format(l), size[l][0], size[l][1], Cl[l]) for l in range(L)]
    H0, W0 = size[0][0] * 2, size[0][1] * 2
    feats = {}

    def enc_segs(l):
        segs = [(zb[l].sub(0, info[l]["n_split"]), 0)]
        for l2 in range(l + 1, L):
            segs.append((feats[l2], l2 - l))
        return segs

    if direction == "reverse":
        plan.ops.append(LayoutOp("ingest", h, w, 3, "lr", zb[L - 1].sub(0, 3)))
        draw = 0
        for l in range(L - 1, -1, -1):
            H, W = size[l]
            ns = info[l]["n_split"]
            cfm = flow.cond_flow(l)
            pre = "flow.level{}_condFlow".format(l)
            feats[l] = _emit_encoder(plan, pre, cfm, enc_segs(l), H, W, l)
            a_dst = zb[l].sub(ns, Cl[l] - ns)
            a = _cond_view(plan, l, H, W, a_dst)
            hp = plan.view("prior_l{}".format(l), H, W, (2 * cfm.z_channels + 3) // 4 * 4).sub(0, 2 * cfm.z_channels)
            plan.ops.append(ConvOp(H, W, [(feats[l], 0)], 3, 2 * cfm.z_channels, pre + ".f.weight", pre + ".f.bias",
                                   pre + ".f.logs#exp3", ACT_NONE, hp, tag="prior.conv"))
            plan.ops.append(PriorOp("sample", H, W, hp, a, not sr, eps_index=draw))
            plan.noise_shapes.append((cfm.z_channels, H, W))
            draw += 1
            for j in range(len(cfm.additional_flow_steps) - 1, -1, -1):
                _emit_step(plan, "{}.additional_flow_steps.{}".format(pre, j), cfm.additional_flow_steps[j],
                           a, feats[l], H, W, l, True)
            if a is not a_dst:
                plan.ops.append(LayoutOp("copy", H, W, a.C, a, a_dst))
            for i in reversed(info[l]["steps"]):
                _emit_step(plan, "flow.layers.{}".format(i), flow.layers[i], zb[l], None, H, W, l, True)
            variant = "haar_inv" if info[l]["haar"] else "unsqueeze"
            if l > 0:
                plan.ops.append(LayoutOp(variant, H, W, Cl[l] // 4, zb[l], zb[l - 1].sub(0, Cl[l] // 4)))
            else:
                x0 = plan.view("x0", H0, W0, 3)
                plan.ops.append(LayoutOp(variant, H, W, 3, zb[0], x0))
                plan.ops.append(LayoutOp("egress", H0, W0, 3, x0, "hr_raw", post=0))
                plan.ops.append(LayoutOp("egress", H0, W0, 3, x0, "hr", post=1))
                plan.outputs["hr_raw"] = (3, H0, W0)
                plan.outputs["hr"] = (3, H0, W0)
        return plan

    # ---------------- forward
    x0 = plan.view("x0", H0, W0, 3)
    plan.ops.append(LayoutOp("ingest", H0, W0, 3, "hr", x0, noise="dequant" if sr else None,
                             noise_scale=(1.0 / float(net.quant)) if sr else 0.0))
    plan.uses_logdet = sr
    for l in range(L):
        H, W = size[l]
        src = x0 if l == 0 else zb[l - 1].sub(0, info[l - 1]["n_split"])
        plan.ops.append(LayoutOp("haar_fwd" if info[l]["haar"] else "squeeze", H, W, Cl[l] // 4, src, zb[l]))
        for i in info[l]["steps"]:
            _emit_step(plan, "flow.layers.{}".format(i), flow.layers[i], zb[l], None, H, W, l, False)
    for l in range(L - 1, -1, -1):
        H, W = size[l]
        ns = info[l]["n_split"]
        cfm = flow.cond_flow(l)
        pre = "flow.level{}_condFlow".format(l)
        feats[l] = _emit_encoder(plan, pre, cfm, enc_segs(l), H, W, l)
        a_src = zb[l].sub(ns, Cl[l] - ns)
        a = _cond_view(plan, l, H, W, a_src)
        if a is not a_src:
            plan.ops.append(LayoutOp("copy", H, W, a.C, a_src, a))
        for j in range(len(cfm.additional_flow_steps)):
            _emit_step(plan, "{}.additional_flow_steps.{}".format(pre, j), cfm.additional_flow_steps[j],
                       a, feats[l], H, W, l, False)
        hp = plan.view("prior_l{}".format(l), H, W, (2 * cfm.z_channels + 3) // 4 * 4).sub(0, 2 * cfm.z_channels)
        plan.ops.append(ConvOp(H, W, [(feats[l], 0)], 3, 2 * cfm.z_channels, pre + ".f.weight", pre + ".f.bias",
                               pre + ".f.logs#exp3", ACT_NONE, hp, tag="prior.conv"))
        if sr:
            plan.ops.append(PriorOp("logp", H, W, hp, a, False))
        else:
            name = "fake_z{}".format(l + 1)
            plan.ops.append(PriorOp("standardize", H, W, hp, a, True, out_name=name))
            plan.outputs[name] = (cfm.z_channels, H, W)
    zl = zb[L - 1].sub(0, 3)
    plan.ops.append(LayoutOp("egress", h, w, 3, zl, "z_raw", post=0))
    plan.outputs["z_raw"] = (3, h, w)
    if sr:
        plan.ops.append(LayoutOp("egress", h, w, 3, zl, "fake_lr", post=2))
        plan.outputs["fake_lr"] = (3, h, w)
        plan.ops.append(DiracLogpOp("fake_lr", "lr", -6.0, 3 * h * w))
    else:
        plan.ops.append(LayoutOp("egress", h, w, 3, zl, "fake_lr", post=1))
        plan.outputs["fake_lr"] = (3, h, w)
    return plan
