#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs (one JSON line each; not the driver's bench):
  configs[2]  8x face SR, B=32, 20x20 LR <-> 160x160 HR: forward NLL and inverse
  configs[3]  4x rescaling round trip, B=64, 256x256 HR: encode + decode (T=1.0)
  configs[1]  4x SR forward NLL, B=16 (the inverse is bench.py's metric)
Device-resident inputs, CUDA-graph replay, CUDA events, median of --steps runs."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hcflow_b200 import options as popt, synth  # noqa: E402
from hcflow_b200.arch import build_net  # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=7)
    ap.add_argument("--precision", default="tf32x3")
    args = ap.parse_args()
    cases = [("sr_x4", 16, 40, "forward"), ("sr_x8", 32, 20, "reverse"), ("sr_x8", 32, 20, "forward"),
             ("rescaling_x4", 64, 64, "roundtrip")]
    for cfg, B, lrhw, what in cases:
        opt = popt.load_config(cfg)
        net = build_net(opt)
        net.load_state_dict(synth.synthetic_state_dict(net.state_dict(), seed=1), strict=True)
        net = net.cuda().eval()
        net.set_precision(args.precision)
        s = opt["scale"]
        lr = synth.synthetic_lr(B, lrhw, lrhw).cuda()
        hr = synth.synthetic_hr(B, lrhw * s, lrhw * s).cuda()
        with torch.no_grad():
            if what == "reverse":
                eng = net.engine("reverse", B, lrhw, lrhw, lr.device)
                eng.ext["lr"].copy_(lr)
                ms = timed(eng.run, args.steps)
                launches = eng.launches_per_run
            elif what == "forward":
                eng = net.engine("forward", B, lrhw, lrhw, lr.device)
                eng.ext["hr"].copy_(hr)
                eng.ext["lr"].copy_(lr)
                ms = timed(eng.run, args.steps)
                launches = eng.launches_per_run
            else:
                ef = net.engine("forward", B, lrhw, lrhw, lr.device)
                er = net.engine("reverse", B, lrhw, lrhw, lr.device)
                ef.ext["hr"].copy_(hr)

                def both():
                    ef.run()
                    er.ext["lr"].copy_(ef.ext["fake_lr"])
                    er.run()
                ms = timed(both, args.steps)
                launches = ef.launches_per_run + er.launches_per_run
        mp = B * (lrhw * s) ** 2 / 1e6
        print(json.dumps({"config": cfg, "pass": what, "B": B, "hr": lrhw * s, "precision": args.precision,
                          "ms": round(ms, 3), "hr_mp_per_s": round(mp / (ms / 1e3), 2), "launches": launches}), flush=True)
        del net
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
