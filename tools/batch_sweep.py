"""HR MP/s of the 4x SR inverse pass vs per-GPU batch (GPU box): configs[4] shards B=512 as 64 per GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hcflow_b200 import options as popt, synth  # noqa: E402
from hcflow_b200.arch import build_net  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
    opt = popt.load_config("sr_x4")
    net = build_net(opt)
    sd = synth.synthetic_state_dict(net.state_dict(), seed=1)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    net.set_precision(prec)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    ref = None
    for B in (1, 4, 16, 64):
        eng = net.engine("reverse", B, 40, 40, torch.device("cuda", 0))
        eng.ext["lr"].copy_(synth.synthetic_lr(B, 40, 40, seed=0))
        unit = synth.synthetic_noise(net.noise_shapes(B, 40, 40), seed=123)
        for i, e in enumerate(unit):
            eng.ext["eps{}".format(i)].copy_(0.8 * e)
        for _ in range(3):
            eng.run()
        torch.cuda.synchronize()
        ts = []
        for k in range(6):
            flush.fill_(float(k))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.run()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        ms = ts[len(ts) // 2]
        first = eng.ext["hr_raw"][0].clone()
        if ref is None:
            ref = first      # image 0 has the same LR / noise seed prefix only for B=1; used as a finiteness check
        print(json.dumps({"precision": prec, "B": B, "ms": round(ms, 3), "hr_mp_per_s": round(B * 160 * 160 / 1e6 / (ms / 1e3), 2),
                          "finite": bool(torch.isfinite(eng.ext["hr_raw"]).all()), "launches": eng.launches_per_run}), flush=True)
        net.clear_engines()
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
