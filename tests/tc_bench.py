"""Micro-benchmark of single conv launches (GPU box): per-shape time, TFLOP/s and L2->SM bytes.
    python tests/tc_bench.py [--precision tf32] [--only NAME] [--reps 20] [--mt 0|1|2]
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hcflow_b200 import _lib as L  # noqa: E402
from hcflow_b200 import prep  # noqa: E402

SHAPES = {
    # name: (B, H, W, Cin, ld, cout)
    "L0_conv1": (16, 80, 80, 64, 192, 32),
    "L0_conv2": (16, 80, 80, 96, 192, 32),
    "L0_conv3": (16, 80, 80, 128, 192, 32),
    "L0_conv4": (16, 80, 80, 160, 192, 32),
    "L0_conv5": (16, 80, 80, 192, 192, 64),
    "L1_conv1": (16, 40, 40, 64, 192, 32),
    "L1_conv4": (16, 40, 40, 160, 192, 32),
    "L1_conv5": (16, 40, 40, 192, 192, 64),
    "L0_fcn3": (16, 80, 80, 64, 64, 12),
    "L0_prior": (16, 80, 80, 128, 128, 12),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--only", default=None)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--mt", default="0")
    ap.add_argument("--flush", action="store_true")
    args = ap.parse_args()
    if args.mt != "0":
        os.environ["HCF_TC_MT"] = args.mt
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    for name, (B, H, W, cin, ld, cout) in SHAPES.items():
        if args.only and args.only != name:
            continue
        g = torch.Generator().manual_seed(0)
        xin = torch.randn(B, H, W, ld, generator=g).cuda()
        out = torch.zeros(B, H, W, 192, device="cuda")
        w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
        npad = prep.npad_for(cout)
        wp = prep.pack_conv_weight(w, [cin], npad).cuda()
        bias = prep.pad_vec(torch.randn(cout, generator=g) * 0.1, npad, 0.0).cuda()
        a = L.ConvArgs()
        a.B, a.H, a.W, a.nseg = B, H, W, 1
        a.seg[0].ptr, a.seg[0].ld, a.seg[0].C, a.seg[0].up_shift = xin.data_ptr(), ld, cin, 0
        a.ks, a.kpad, a.cout, a.npad = 3, cin, cout, npad
        a.w, a.bias, a.act = wp.data_ptr(), bias.data_ptr(), 2
        a.out, a.out_ld = out.data_ptr() + 4 * 64, 192
        if args.precision == "fp32":
            def run():
                return lib.hcf_conv_fp32(C.byref(a), st)
        else:
            npass = 1 if args.precision == "tf32" else 3
            img = torch.zeros(lib.hcf_conv_tc_weight_bytes(cin, cout, 3, npass) // 4, dtype=torch.float32)
            L.check(lib.hcf_conv_tc_pack_weights(w.contiguous().data_ptr(), cin, cout, 3, npass, img.data_ptr()), "pack")
            img = img.cuda()
            h = C.c_void_p()
            L.check(lib.hcf_conv_tc_plan_create(C.byref(a), img.data_ptr(), npass,
                                                C.byref(h)), "plan")

            def run():
                return lib.hcf_conv_tc_run(h, st)
        for _ in range(3):
            assert run() == 0
        torch.cuda.synchronize()
        times = []
        for _ in range(args.reps):
            if args.flush:
                flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            assert run() == 0
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3)
        times.sort()
        med = times[len(times) // 2]
        flop = 2.0 * B * H * W * 9 * cin * cout
        print(json.dumps({"shape": name, "precision": args.precision, "mt": args.mt, "flush": args.flush,
                          "us_med": round(med, 2), "us_min": round(times[0], 2),
                          "tflops": round(flop / med / 1e6, 1),
                          "act_MB": round(B * H * W * (cin + cout) * 4 / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
