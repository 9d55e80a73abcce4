"""Engine-level rewrites of a launch plan for the tensor-core modes (pure functions over plan ops: no CUDA, so
the host logic is covered by the CPU tests through tests/plan_emulator.py).

  1. up-sampled conv segments (cat[z, up2(cf2), up4(cf3)] of an encoder's first conv, FlowNet_SR_x4.py:98,117) are
     materialised once, so that conv joins the level's chained tensor-core launch;
  2. shared conditioning: the conditional FlowSteps of a level all run their sub-net's first conv on cat(z1, u) with
     the SAME encoder feature u (ConditionalFlow.py:62-66, AffineCouplings.py:31): W * cat(z1, u) = W_z * z1 + W_u * u,
     and the W_u * u parts of all steps are wide convs over u (Cout = 2 steps x 64 = UMMA N 128) computed once per
     level; each step convolves only its few z1 channels and adds its slice before ActNorm + ReLU (hcf_conv_args.pre);
  3. the growth convs of every residual dense block in pairs (pair_rdb_convs);
  4. the FlowStep tail (FlowStep.py:55-64, reverse) fused into the sub-net's last conv (hcf_conv_step): removes one
     launch per step and lets the convs of consecutive steps of a level run as ONE chained launch.
"""
import torch

from . import plan as P


def raw_weight(sd, op):
    """[Cout, Cin, ks, ks] fp32 CPU weight of a conv op; rewrites may slice the input-channel axis (op.w_in) or
    concatenate several parameters along Cout (op.weight = tuple of (key, lo, hi))."""
    if isinstance(op.weight, str):
        w = sd[op.weight].detach().float()
        return w[:, op.w_in[0]:op.w_in[1]].contiguous() if op.w_in else w
    return torch.cat([sd[k].detach().float()[:, lo:hi] for k, lo, hi in op.weight], 0).contiguous()


def materialise_upsampled(ops, bufs):
    out = []
    for op in ops:
        if isinstance(op, P.ConvOp) and any(up > 0 for _, up in op.segs) and all(
                v.C % 4 == 0 and v.off % 4 == 0 and v.buf.C % 4 == 0 for v, up in op.segs if up > 0):
            segs = []
            for v, up in op.segs:
                if up > 0:
                    ub = bufs.setdefault("up{}_{}_{}".format(up, v.buf.name, v.off),
                                         P.Buf("up{}_{}_{}".format(up, v.buf.name, v.off), op.H, op.W, v.C))
                    dst = P.View(ub, 0, v.C)
                    out.append(P.LayoutOp("upsample", op.H, op.W, v.C, v, dst, post=up))
                    segs.append((dst, 0))
                else:
                    segs.append((v, 0))
            op = P.ConvOp(op.H, op.W, segs, op.ks, op.cout, op.weight, op.bias, op.scale, op.act, op.out, op.out2,
                          op.res1, op.alpha1, op.res2, op.alpha2, op.tag)
        out.append(op)
    return out


def share_conditioning(ops, bufs, max_cout=128):
    """max_cout: widest shared conv (UMMA N); 64 when the sub-nets run split (the [hi ; lo] weight rows double N)."""
    ops = list(ops)
    groups = {}
    for i, op in enumerate(ops):
        if isinstance(op, P.ConvOp) and op.tag == "fcn.conv1" and len(op.segs) == 2 and op.segs[1][1] == 0:
            v = op.segs[1][0]
            groups.setdefault((v.buf.name, v.off, v.C, op.H, op.W, op.cout, op.segs[0][0].C), []).append(i)
    inserts = {}
    for (bname, off, cc, H, W, cout, zc), idxs in groups.items():
        if len(idxs) < 2 or cout > 128 or cout % 4 != 0:
            continue
        n = len(idxs)
        per = max(1, max_cout // cout)
        ubuf = bufs.setdefault("ucond_{}_{}".format(bname, off), P.Buf("ucond_{}_{}".format(bname, off), H, W, n * cout))
        cond = ops[idxs[0]].segs[1][0]
        uops = []
        for g in range(0, n, per):
            mem = idxs[g:g + per]
            uops.append(P.ConvOp(H, W, [(cond, 0)], 3, cout * len(mem),
                                 tuple((ops[m].weight, zc, zc + cc) for m in mem), None, None, P.ACT_NONE,
                                 P.View(ubuf, g * cout, cout * len(mem)), tag="fcn.ucond"))
        # in front of the first step -- and of its forward ActNorm / W head, so that [head, conv1, conv2, conv3, coupling]
        # stays contiguous (group_flowsteps); the shared convs depend on the encoder feature only
        at = idxs[0]
        if at > 0 and isinstance(ops[at - 1], P.StepOp) and ops[at - 1].variant == "forward_head":
            at -= 1
        inserts[at] = uops
        for j, m in enumerate(idxs):
            o = ops[m]
            ops[m] = P.ConvOp(o.H, o.W, [o.segs[0]], o.ks, o.cout, o.weight, o.bias, o.scale, o.act, o.out, tag=o.tag,
                              w_in=(0, zc), pre=P.View(ubuf, j * cout, cout))
    out = []
    for i, op in enumerate(ops):
        out.extend(inserts.get(i, []))
        out.append(op)
    return out


def pair_rdb_convs(ops, bufs):
    """Dense-block growth convs in pairs (Basic.py:377-381): conv_{k+1} reads everything conv_k reads plus conv_k's
    32 new channels.  On the tensor cores a conv with 32 output channels pays for fetching the activation operand
    (UMMA N = 32 keeps the tensor pipe 40 % busy at best, N = 64: 67 %), so conv_k also accumulates, in 32 more
    accumulator columns, conv_{k+1}'s partial sum over the shared input channels (raw fp32, `raw2`), and conv_{k+1}
    shrinks to a conv over the 32 new channels that adds the partial before its bias (`pre`).  Pairs (conv1, conv2)
    and (conv3, conv4) of every RDB with 32 growth channels."""
    ops = list(ops)
    i = 0
    while i + 1 < len(ops):
        a, b = ops[i], ops[i + 1]
        ok = (isinstance(a, P.ConvOp) and isinstance(b, P.ConvOp) and a.tag in ("enc.rdb.conv1", "enc.rdb.conv3")
              and b.tag == {"enc.rdb.conv1": "enc.rdb.conv2", "enc.rdb.conv3": "enc.rdb.conv4"}.get(a.tag)
              and a.cout == 32 and b.cout == 32 and a.ks == b.ks == 3 and len(a.segs) == 1 and len(b.segs) == 1
              and a.segs[0][1] == 0 and b.segs[0][1] == 0 and (a.H, a.W) == (b.H, b.W)
              and isinstance(a.weight, str) and isinstance(b.weight, str) and a.w_in is None and b.w_in is None
              and a.out2 is None and b.out2 is None and a.res1 is None and b.res1 is None and a.pre is None and b.pre is None
              and a.act == b.act and a.scale is None and b.scale is None)
        if ok:
            va, vb = a.segs[0][0], b.segs[0][0]
            # conv_{k+1}'s input = conv_k's input followed by conv_k's output, contiguous in one buffer
            ok = (va.buf.name == vb.buf.name == a.out.buf.name and va.off == vb.off and vb.C == va.C + 32
                  and a.out.off == va.off + va.C and a.out.C == 32 and va.C % 32 == 0)
        if not ok:
            i += 1
            continue
        name = "rdbpart_{}x{}".format(a.H, a.W)
        pbuf = bufs.setdefault(name, P.Buf(name, a.H, a.W, 32))
        part = P.View(pbuf, 0, 32)
        cin = a.segs[0][0].C
        ops[i] = P.ConvOp(a.H, a.W, a.segs, 3, 64, ((a.weight, 0, cin), (b.weight, 0, cin)), a.bias, None, a.act, a.out,
                          tag=a.tag, raw2=part)
        ops[i + 1] = P.ConvOp(b.H, b.W, [(a.out, 0)], 3, 32, b.weight, b.bias, None, b.act, b.out, tag=b.tag,
                              w_in=(cin, cin + 32), pre=part)
        i += 2
    return ops


STEP_MAXC = 24   # csrc/conv_tc_kernel.cuh
# precision modes whose coupling sub-nets (FCN) run with the operand split: a coupling turns an error dh of the
# sub-net output into |z2| * 0.64 * dh, so with trained-size couplings (|h| ~ 1) one fp16 / TF32 pass per sub-net
# conv leaves 3e-4 .. 8e-4 on the un-clamped HR after only 8 steps (tests/golden/*_stress.pt; CPU emulation in
# tools/emulate_precision.py), the split 5e-6
SPLIT_FCN_MODES = ("f16x3", "tf32x3", "tf32x3_all")


def fuse_steps(ops):
    fused = []
    for op in ops:
        prev = fused[-1] if fused else None
        if (isinstance(op, P.StepOp) and op.variant == "inverse" and op.mode == "affine" and op.h is not None
                and isinstance(prev, P.ConvOp) and prev.tag == "fcn.conv3" and prev.step is None
                and prev.out == op.h and prev.out2 is None and prev.res1 is None and prev.res2 is None
                and prev.cout == 2 * (op.z.C - op.n_pass) and prev.cout <= 32 and op.z.C <= STEP_MAXC
                and (prev.H, prev.W) == (op.H, op.W)):
            prev.step = op
            continue
        if isinstance(op, P.ConvOp) and op.tag == "fcn.conv3":
            op = P.ConvOp(op.H, op.W, op.segs, op.ks, op.cout, op.weight, op.bias, op.scale, op.act, op.out,
                          tag=op.tag, w_in=op.w_in, pre=op.pre)   # private copy: the plan's op stays untouched
        fused.append(op)
    return fused


FLOWCHAIN_C = (6, 12, 21, 24)   # csrc/flowstep_tc.cu


def _fcn_triplet(ops, i):
    """ops[i:i+3] = conv1 (3x3 over z1 only, optional pre addend), conv2 (1x1), conv3 (3x3) of one FCN sub-net?"""
    if i + 2 >= len(ops):
        return None
    c1, c2, c3 = ops[i], ops[i + 1], ops[i + 2]
    if not all(isinstance(c, P.ConvOp) for c in (c1, c2, c3)):
        return None
    ok = (c1.tag == "fcn.conv1" and c2.tag == "fcn.conv2" and c3.tag == "fcn.conv3"
          and len(c1.segs) == 1 and c1.segs[0][1] == 0 and c1.ks == 3 and c1.cout == 64 and c1.act == P.ACT_RELU
          and c1.bias and c1.scale and c1.res1 is None and c1.res2 is None and c1.out2 is None and c1.raw2 is None
          and len(c2.segs) == 1 and c2.segs[0][0] == c1.out and c2.ks == 1 and c2.cout == 64 and c2.act == P.ACT_RELU
          and c2.bias and c2.scale and c2.pre is None
          and len(c3.segs) == 1 and c3.segs[0][0] == c2.out and c3.ks == 3 and c3.cout <= 32 and c3.act == P.ACT_NONE
          and c3.bias and c3.scale and c3.pre is None and c3.res1 is None and c3.out2 is None
          and (c1.H, c1.W) == (c2.H, c2.W) == (c3.H, c3.W)
          and (c1.pre is None or (c1.pre.C == 64 and c1.pre.off % 4 == 0 and c1.pre.buf.C % 4 == 0)))
    return (c1, c2, c3) if ok else None


def group_flowsteps(ops):
    """Runs of FlowSteps whose sub-net is an FCN over z1 alone (after share_conditioning the conditional steps
    qualify too) become ONE FlowChainOp each: the fused-FlowStep kernel runs a whole step per work item.
      reverse:  [conv1, conv2, conv3 + fused StepOp("inverse")] per step           (after fuse_steps)
      forward:  [StepOp("forward_head"), conv1, conv2, conv3, StepOp("forward_coupling")] per step"""
    out, i = [], 0
    while i < len(ops):
        steps, orig, j = [], [], i
        z = None
        forward = None
        while True:
            head = None
            k = j
            if k < len(ops) and isinstance(ops[k], P.StepOp) and ops[k].variant == "forward_head":
                head = ops[k]
                k += 1
            tri = _fcn_triplet(ops, k)
            if tri is None:
                break
            c1, c2, c3 = tri
            if head is not None:
                tail = ops[k + 3] if k + 3 < len(ops) else None
                if not (isinstance(tail, P.StepOp) and tail.variant == "forward_coupling" and tail.h == c3.out
                        and tail.z == head.z and c3.step is None):
                    break
                fw, end = True, k + 4
            else:
                tail = c3.step
                if tail is None or tail.variant != "inverse":
                    break
                fw, end = False, k + 3
            zz = tail.z
            if not (tail.mode == "affine" and zz.C in FLOWCHAIN_C and tail.n_pass == zz.C // 2
                    and c3.cout == 2 * (zz.C - tail.n_pass) and c1.segs[0][0] == zz.sub(0, tail.n_pass)
                    and (tail.H, tail.W) == (c1.H, c1.W)):
                break
            if steps and (zz != z or fw != forward):
                break
            z, forward = zz, fw
            steps.append((c1, c2, c3, tail, head))
            orig.extend(ops[j:end])
            j = end
        if steps:
            c1 = steps[0][0]
            out.append(P.FlowChainOp(c1.H, c1.W, z, steps[0][3].n_pass, forward, steps, orig,
                                     tag="flowsteps[{}{}]x{}".format("cond" if c1.pre is not None else "main",
                                                                    ",fwd" if forward else "", len(steps))))
            i = j
        else:
            out.append(ops[i])
            i += 1
    return out


def rewrite_ops(plan_ops, precision, share_cond=True, fuse=True, pair=True, flowchain=True):
    """-> (ops, {name: Buf} of the extra fp32 buffers the rewritten ops use)."""
    ops = list(plan_ops)
    bufs = {}
    if precision == "fp32":
        return ops, bufs
    ops = materialise_upsampled(ops, bufs)
    # measured (B200, x4 B=16): pairing cuts the TF32 chains by 9 % (12.25 -> 11.09 ms/step) but not the fp16 ones
    # (10.65 -> 10.82): their growth convs are no longer bound by the MMA operand fetch
    if pair and not precision.startswith("f16"):
        ops = pair_rdb_convs(ops, bufs)
    if share_cond:
        ops = share_conditioning(ops, bufs, 64 if precision in SPLIT_FCN_MODES else 128)
    if fuse:
        ops = fuse_steps(ops)
    if flowchain and precision in ("f16", "f16x3"):
        ops = group_flowsteps(ops)
    return ops, bufs


# ---------------------------------------------------------------------------------------------------------------
# fp16 chains: which representation every conv writes, which inputs need converting, where fused steps leave z1
OUT_F32, OUT_HI, OUT_LO = 1, 2, 4   # == _lib.OUT_* == HCF_OUT_* (include/hcflow_b200.h)


def _overlap(a, b):
    return a.buf.name == b.buf.name and a.off < b.off + b.C and b.off < a.off + a.C


def op_reads(op):
    if isinstance(op, P.FlowChainOp):
        return [op.z] + [st[0].pre for st in op.steps if st[0].pre is not None]
    if isinstance(op, P.ConvOp):
        return ([v for v, _ in op.segs] + [v for v in (op.res1, op.res2, op.pre) if v is not None]
                + ([op.step.z] if op.step is not None else []))
    if isinstance(op, P.StepOp):
        return [v for v in (op.z, op.h) if v is not None]
    if isinstance(op, P.PriorOp):
        return [op.h, op.z]
    if isinstance(op, P.LayoutOp):
        return [op.src] if isinstance(op.src, P.View) else []
    return []


def op_writes(op):
    if isinstance(op, P.FlowChainOp):
        return [op.z]
    if isinstance(op, P.ConvOp):
        if op.step is not None:    # h is consumed in the epilogue, z is updated in place
            return [op.step.z]
        return [v for v in (op.out, op.out2, op.raw2) if v is not None]
    if isinstance(op, P.StepOp):
        return [op.z]
    if isinstance(op, P.PriorOp):
        return [op.z] if op.variant == "sample" else []
    if isinstance(op, P.LayoutOp):
        return [op.dst] if isinstance(op.dst, P.View) else []
    return []


def read_later(view, later_ops):
    """Is `view` (written by a chain conv) read by one of the ops that follow the chain before it is fully
    overwritten?  (buffers are reused by later steps, so a plain overlap test would be too conservative)"""
    for o in later_ops:
        if any(_overlap(v, view) for v in op_reads(o)):
            return True
        if any(w.buf.name == view.buf.name and w.off <= view.off and w.off + w.C >= view.off + view.C
               for w in op_writes(o)):
            return False
    return False


def split_views(op, passes, split_ch):
    """the input views of a conv that its split (hi + lo on both operands) covers"""
    if passes != 3:
        return []
    if split_ch < 0:
        return [v for v, _ in op.segs]
    out, left = [], split_ch
    for v, _ in op.segs:
        if left <= 0:
            break
        out.append(v.sub(0, min(v.C, left)))
        left -= (v.C + 63) // 64 * 64
    return out


def chain16_layout(ops, passes, splits, later_ops):
    """Data-flow decisions for running the convs `ops` as one fp16 chain (engine._try_chain16 turns them into
    pointers).  Returns None when the run does not qualify, else a dict:
      flags[k]        OUT_F32 | OUT_HI | OUT_LO of conv k's output(s)
      segs[k][s]      {"key": (buffer, offset, C), "staged": bool}  -- staged: the fp32 geometry breaks TMA's 16-byte
                      rules in fp16 (ld % 8, offset % 8), the conv reads a private padded fp16 copy
      external        {key: (view, need_lo, staged)} inputs that no conv / fused step of the chain produced: converted
                      (hcf_split16) right before the launch
      step_target[k]  key of the fp16 copy that conv k's fused FlowStep must leave for the next step's first conv (or None)
      step_lo[k]      that copy needs its lo plane too (a split conv of the chain reads it)
    """
    n = len(ops)
    sv = [split_views(op, ps, sp) for op, ps, sp in zip(ops, passes, splits)]
    flags = []
    for k, op in enumerate(ops):
        outs = [v for v in (op.out, op.out2) if v is not None]
        hi = lo = f32 = False
        for j in range(k + 1, n):
            if any(_overlap(v, o) for v, _ in ops[j].segs for o in outs):
                hi = True
                lo = lo or any(_overlap(v, o) for v in sv[j] for o in outs)
        for j in range(n):   # residual sources and pre-activation addends stay fp32
            if any(_overlap(v, o) for v in (ops[j].res1, ops[j].res2, ops[j].pre) if v is not None for o in outs):
                f32 = True
        if any(read_later(o, later_ops) for o in outs):
            f32 = True
        if not (hi or f32):
            f32 = True
        flags.append((OUT_F32 if f32 else 0) | (OUT_HI if hi else 0) | (OUT_LO if lo else 0))
    external, segs, known = {}, [], set()
    for k, op in enumerate(ops):
        row = []
        for v, _ in op.segs:
            by_conv = any(_overlap(v, o) for j in range(k) if ops[j].step is None
                          for o in (ops[j].out, ops[j].out2) if o is not None)
            by_step = any(ops[j].step is not None and _overlap(v, ops[j].step.z) for j in range(k))
            aligned = v.buf.C % 8 == 0 and v.off % 8 == 0
            key = (v.buf.name, v.off, v.C)
            if by_conv and not aligned:
                return None
            # the lo plane is needed as soon as ANY split conv of the chain reads these channels (e.g. the first
            # RDB's conv5 reads x0 inside a wider view that is otherwise produced by the chain)
            need_lo = any(_overlap(v, s) for j in range(n) for s in sv[j])
            if not (by_conv or by_step) and key not in external:
                external[key] = (v, need_lo, not aligned)
            known.add(key)
            row.append({"key": key, "staged": not aligned})
        segs.append(row)
    step_target, step_lo = [], []
    for op in ops:
        if op.step is not None:
            key = (op.step.z.buf.name, op.step.z.off, op.step.n_pass)
            step_target.append(key if key in known else None)
            step_lo.append(any(_overlap(op.step.z.sub(0, op.step.n_pass), s) for j in range(n) for s in sv[j]))
        else:
            step_target.append(None)
            step_lo.append(False)
    return {"flags": flags, "segs": segs, "external": external, "step_target": step_target, "step_lo": step_lo}
