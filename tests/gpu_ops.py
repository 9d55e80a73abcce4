"""Thin test-side wrappers that call single C-ABI operators on torch CUDA tensors."""
import ctypes as C

import torch

from hcflow_b200 import _lib as L
from hcflow_b200 import prep


def _stream():
    return torch.cuda.current_stream().cuda_stream


def to_nhwc(x, ld=None, off=0):
    """NCHW cpu/cuda tensor -> (NHWC cuda buffer with `ld` channels, view pointer)."""
    B, Cc, H, W = x.shape
    ld = ld or Cc
    buf = torch.zeros(B, H, W, ld, dtype=torch.float32, device="cuda")
    buf[..., off:off + Cc] = x.to("cuda").permute(0, 2, 3, 1)
    return buf


def from_nhwc(buf, off, Cc):
    return buf[..., off:off + Cc].permute(0, 3, 1, 2).contiguous().cpu()


def conv(segs, w, bias=None, scale=None, act=0, res1=None, alpha1=1.0, res2=None, alpha2=1.0,
         out_ld=None, out_off=0, precision="fp32", want_out2=False):
    """segs: list of (nchw tensor, up_shift, ld, off).  w: [Cout, Cin, ks, ks] cpu.
    Returns NCHW cpu result (and the out2 copy when asked)."""
    lib = L.load()
    cout, cin, ks, _ = w.shape
    up0 = segs[0][1]
    B, _, Hs, Ws = segs[0][0].shape
    H, W = Hs << up0, Ws << up0
    npad = prep.npad_for(cout)
    a = L.ConvArgs()
    a.B, a.H, a.W, a.nseg = B, H, W, len(segs)
    keep = []
    kpad = 0
    segc = []
    for i, (x, up, ld, off) in enumerate(segs):
        buf = to_nhwc(x, ld, off)
        keep.append(buf)
        a.seg[i].ptr = buf.data_ptr() + 4 * off
        a.seg[i].ld, a.seg[i].C, a.seg[i].up_shift = buf.shape[3], x.shape[1], up
        kpad += prep.seg_pad(x.shape[1])
        segc.append(x.shape[1])
    wp = prep.pack_conv_weight(w, segc, npad).cuda()
    a.ks, a.kpad, a.cout, a.npad = ks, kpad, cout, npad
    a.w = wp.data_ptr()
    if bias is not None:
        bp = prep.pad_vec(bias, npad, 0.0).cuda()
        keep.append(bp)
        a.bias = bp.data_ptr()
    if scale is not None:
        sp = prep.pad_vec(scale, npad, 1.0).cuda()
        keep.append(sp)
        a.scale = sp.data_ptr()
    a.act = act
    out_ld = out_ld or cout
    out = torch.zeros(B, H, W, out_ld, dtype=torch.float32, device="cuda")
    a.out, a.out_ld = out.data_ptr() + 4 * out_off, out_ld
    out2 = None
    if want_out2:
        out2 = torch.zeros(B, H, W, cout, dtype=torch.float32, device="cuda")
        a.out2, a.out2_ld = out2.data_ptr(), cout
    if res1 is not None:
        r1 = to_nhwc(res1)
        keep.append(r1)
        a.res1, a.res1_ld, a.alpha1 = r1.data_ptr(), r1.shape[3], alpha1
    if res2 is not None:
        r2 = to_nhwc(res2)
        keep.append(r2)
        a.res2, a.res2_ld, a.alpha2 = r2.data_ptr(), r2.shape[3], alpha2
    if precision == "fp32":
        L.check(lib.hcf_conv_fp32(C.byref(a), _stream()), "conv_fp32")
    else:
        assert lib.hcf_conv_tc_supported(C.byref(a)), "tc kernel does not support this shape"
        wt = prep.pad_weight_for_tc(w, segc)
        kin = wt.shape[1]
        npass = {"tf32": 1, "tf32x3": 3}[precision]
        nbytes = lib.hcf_conv_tc_weight_bytes(kin, cout, ks, npass)
        img = torch.zeros(nbytes // 4, dtype=torch.float32)
        L.check(lib.hcf_conv_tc_pack_weights(wt.data_ptr(), kin, cout, ks, npass, img.data_ptr()), "tc_pack")
        img = img.cuda()
        h = C.c_void_p()
        L.check(lib.hcf_conv_tc_plan_create(C.byref(a), img.data_ptr(), {"tf32": 1, "tf32x3": 3}[precision],
                                            C.byref(h)), "tc_plan_create")
        L.check(lib.hcf_conv_tc_run(h, _stream()), "conv_tc_run")
        torch.cuda.synchronize()
        lib.hcf_conv_tc_plan_destroy(h)
    torch.cuda.synchronize()
    res = from_nhwc(out, out_off, cout)
    if want_out2:
        return res, from_nhwc(out2, 0, cout)
    return res


def step(variant, z, h, mode, n_pass, w, an_scale, an_bias, ld=None, off=0, logdet=None):
    """z NCHW cpu -> transformed NCHW cpu (in place on the device buffer)."""
    lib = L.load()
    B, Cc, H, W = z.shape
    zb = to_nhwc(z, ld, off)
    a = L.StepArgs()
    a.npix, a.pix_per_img = B * H * W, H * W
    a.z, a.z_ld, a.C = zb.data_ptr() + 4 * off, zb.shape[3], Cc
    keep = []
    if h is not None:
        hb = to_nhwc(h)
        keep.append(hb)
        a.h, a.h_ld = hb.data_ptr(), hb.shape[3]
    a.mode, a.n_pass = (0 if mode == "affine" else 1), n_pass
    for name, t in (("w", w), ("an_scale", an_scale), ("an_bias", an_bias)):
        if t is not None:
            d = t.float().contiguous().cuda()
            keep.append(d)
            setattr(a, name, d.data_ptr())
    if logdet is not None:
        a.logdet = logdet.data_ptr()
    fn = {"inverse": lib.hcf_step_inverse, "forward_head": lib.hcf_step_forward_head,
          "forward_coupling": lib.hcf_step_forward_coupling}[variant]
    L.check(fn(C.byref(a), _stream()), variant)
    torch.cuda.synchronize()
    return from_nhwc(zb, off, Cc)


def flowstep_chain(z, sd, pres, forward, split, u=None, logdet=None):
    """n = len(pres) FlowSteps (state-dict prefixes in EXECUTION order) on z (NCHW cpu) through ONE fused-FlowStep
    launch (hcf_flowstep_chain_*).  u: conditional feature [B,128,H,W] (its W_u * u part of conv1 is computed here on
    the CPU in fp64 and handed to the kernel as the fp32 `pre` addend, as the engine's shared-conditioning conv does).
    forward: the first step's ActNorm + W head is applied here on the CPU (the engine runs hcf_step_forward_head).
    Returns the transformed z (NCHW cpu)."""
    import torch.nn.functional as F
    lib = L.load()
    st = _stream()
    B, Cc, H, W = z.shape
    n_pass = Cc // 2
    n = len(pres)
    keep = []

    def dev(t):
        d = t.float().contiguous().cuda()
        keep.append(d)
        return d

    def img16(w, ks, skin):
        cout = w.shape[0]
        wt = prep.pad_weight_for_tc(w, [w.shape[1]], chunk=64)
        img = torch.zeros(lib.hcf_conv_tc16_weight_bytes(wt.shape[1], cout, ks, skin) // 2, dtype=torch.float16)
        L.check(lib.hcf_conv_tc16_pack_weights(wt.data_ptr(), wt.shape[1], cout, ks, skin, img.data_ptr()), "pack16")
        return dev_raw(img)

    def dev_raw(t):
        d = t.cuda()
        keep.append(d)
        return d
    zin = z.clone()
    if forward:   # head of the first step
        p0 = pres[0]
        zin = (zin + sd[p0 + ".actnorm.bias"]) * torch.exp(sd[p0 + ".actnorm.logs"])
        if (p0 + ".permute.weight") in sd:
            zin = F.conv2d(zin, sd[p0 + ".permute.weight"].view(Cc, Cc, 1, 1))
    zb = to_nhwc(zin, Cc + 3, 0)    # odd row pitch on purpose
    steps = (L.FlowStep * n)()
    pre_buf = None
    if u is not None:
        pre_buf = torch.zeros(B, H, W, 64 * n, dtype=torch.float32, device="cuda")
    for i, pre in enumerate(pres):
        fp = pre + ".affine.f"
        w1 = sd[fp + ".conv1.weight"].float()
        s = steps[i]
        w1z = w1[:, :n_pass].contiguous()
        img = torch.zeros(lib.hcf_flowstep_w1_bytes() // 2, dtype=torch.float16)
        L.check(lib.hcf_flowstep_pack_w1(w1z.data_ptr(), n_pass, img.data_ptr()), "pack_w1")
        s.w1 = dev_raw(img).data_ptr()
        s.w2 = img16(sd[fp + ".conv2.weight"].float(), 1, 64 if split else 0).data_ptr()
        w3 = sd[fp + ".conv3.weight"].float()
        s.w3 = img16(w3, 3, 64 if split else 0).data_ptr()
        s.bias1 = dev(prep.pad_vec(prep.derive(sd, fp + ".conv1.actnorm.bias"), 64, 0.0)).data_ptr()
        s.scale1 = dev(prep.pad_vec(prep.derive(sd, fp + ".conv1.actnorm.logs#exp"), 64, 1.0)).data_ptr()
        s.bias2 = dev(prep.pad_vec(prep.derive(sd, fp + ".conv2.actnorm.bias"), 64, 0.0)).data_ptr()
        s.scale2 = dev(prep.pad_vec(prep.derive(sd, fp + ".conv2.actnorm.logs#exp"), 64, 1.0)).data_ptr()
        npad = prep.npad_for(w3.shape[0])
        s.bias3 = dev(prep.pad_vec(prep.derive(sd, fp + ".conv3.bias"), npad, 0.0)).data_ptr()
        s.scale3 = dev(prep.pad_vec(prep.derive(sd, fp + ".conv3.logs#exp3"), npad, 1.0)).data_ptr()
        has_perm = (pre + ".permute.weight") in sd
        if forward:
            s.w = dev(prep.derive(sd, pre + ".permute.weight#mat")).data_ptr() if has_perm else None
            s.an_scale = dev(prep.derive(sd, pre + ".actnorm.logs#exppos")).data_ptr()
        else:
            s.w = dev(prep.derive(sd, pre + ".permute.weight#inv")).data_ptr() if has_perm else None
            s.an_scale = dev(prep.derive(sd, pre + ".actnorm.logs#expneg")).data_ptr()
        s.an_bias = dev(prep.derive(sd, pre + ".actnorm.bias#vec")).data_ptr()
        if u is not None:
            wu = F.conv2d(u.double(), w1[:, n_pass:].double(), None, padding=1).float()
            pre_buf[..., 64 * i:64 * (i + 1)] = wu.cuda().permute(0, 2, 3, 1)
            s.pre, s.pre_ld = pre_buf.data_ptr() + 4 * 64 * i, 64 * n
    a = L.FlowStepChainArgs()
    a.B, a.H, a.W, a.C, a.n_pass, a.n_steps, a.split, a.forward = B, H, W, Cc, n_pass, n, int(split), int(forward)
    a.z, a.z_ld = zb.data_ptr(), zb.shape[3]
    za = torch.zeros(B, H, W, 32, dtype=torch.float16, device="cuda")
    zc = torch.zeros(B, H, W, 32, dtype=torch.float16, device="cuda")
    a.z16_a, a.z16_b = za.data_ptr(), zc.data_ptr()
    done = torch.zeros(B * ((H + 15) // 16) * ((W + 7) // 8), dtype=torch.int32, device="cuda")
    a.done = done.data_ptr()
    if logdet is not None:
        a.logdet = logdet.data_ptr()
    a.steps = steps
    h = C.c_void_p()
    L.check(lib.hcf_flowstep_chain_create(C.byref(a), C.byref(h)), "flowstep_chain_create")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.check(lib.hcf_flowstep_chain_set_status(h, status.data_ptr()), "set_status")
    L.check(lib.hcf_flowstep_stage_z1(zb.data_ptr(), zb.shape[3], n_pass, B * H * W, za.data_ptr(), st), "stage_z1")
    L.check(lib.hcf_flowstep_chain_run(h, st), "flowstep_chain_run")
    torch.cuda.synchronize()
    lib.hcf_flowstep_chain_destroy(h)
    assert int(status.item()) == 0, int(status.item())
    assert int(done.min()) == n and int(done.max()) == n
    return from_nhwc(zb, 0, Cc)
