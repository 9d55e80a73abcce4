"""Bring-up probe for the tcgen05 conv kernel (run on the GPU box; each case in its own
process so that a hung kernel only costs one case).  Usage:
    python tests/tc_probe.py --case rdb1 --passes 1 [--safe]
    python tests/tc_probe.py --all          # spawns one subprocess per (case, mode, passes)
"""
import argparse
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (B, H, W, Cin, ld, cout, bias, scale, act, res1, res2)
    "tiny": (1, 16, 8, 32, 32, 16, False, False, 0, False, False),
    "rdb1": (2, 40, 40, 64, 192, 32, True, False, 2, False, False),
    "rdb3": (1, 32, 24, 128, 192, 32, True, False, 2, False, False),
    "rdb5": (1, 20, 20, 192, 192, 64, True, False, 0, True, True),
    "prior42": (1, 16, 16, 128, 128, 42, True, True, 0, False, False),
    "fcn3_22": (2, 13, 9, 64, 64, 22, True, True, 0, False, False),
    "c48": (1, 16, 8, 64, 64, 48, True, False, 0, False, False),
}


def run_case(name, passes, safe):
    import torch
    import torch.nn.functional as F
    from tests import gpu_ops
    if safe:
        os.environ["HCF_TC_SAFE_A"] = "1"
    B, H, W, cin, ld, cout, hb, hs, act, r1, r2 = CASES[name]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    bias = 0.1 * torch.randn(cout, generator=g) if hb else None
    scale = torch.exp(0.2 * torch.randn(cout, generator=g)) if hs else None
    res1 = torch.randn(B, cout, H, W, generator=g) if r1 else None
    res2 = torch.randn(B, cout, H, W, generator=g) if r2 else None
    prec = "tf32" if passes == 1 else "tf32x3"
    got = gpu_ops.conv([(x, 0, ld, 0)], w, bias, scale, act, res1, 0.2, res2, 0.2, out_ld=cout + 8, out_off=4,
                       precision=prec)
    ref = F.conv2d(x.double(), w.double(), None, padding=1)
    if hb:
        ref = ref + bias.double().view(1, -1, 1, 1)
    if hs:
        ref = ref * scale.double().view(1, -1, 1, 1)
    ref = F.relu(ref) if act == 1 else (F.leaky_relu(ref, 0.2) if act == 2 else ref)
    if r1:
        ref = ref * 0.2 + res1.double()
    if r2:
        ref = ref * 0.2 + res2.double()
    err = float((got.double() - ref).abs().max())
    # where is the error? (helps to tell a layout bug from rounding)
    d = (got.double() - ref).abs()
    worst = [int(i) for i in torch.nonzero(d == d.max())[0]]
    frac_bad = float((d > 1e-2).double().mean())
    print(json.dumps({"case": name, "passes": passes, "safe": safe, "max_err": err, "ref_absmax": float(ref.abs().max()),
                      "worst_idx_bchw": worst, "frac_gt_1e-2": frac_bad}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--safe", action="store_true")
    ap.add_argument("--all", action="store_true")
    args = ap.parse_args()
    if not args.all:
        run_case(args.case, args.passes, args.safe)
        return
    for safe in (False, True):
        for passes in (1, 3):
            for name in CASES:
                cmd = ["timeout", "-k", "5", "90", sys.executable, os.path.abspath(__file__), "--case", name,
                       "--passes", str(passes)] + (["--safe"] if safe else [])
                r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
                out = [l for l in r.stdout.splitlines() if l.startswith("{")]
                if out:
                    print(out[-1], flush=True)
                else:
                    print(json.dumps({"case": name, "passes": passes, "safe": safe, "rc": r.returncode,
                                      "stderr": r.stderr[-400:]}), flush=True)


if __name__ == "__main__":
    main()
