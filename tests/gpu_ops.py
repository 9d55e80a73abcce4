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
