"""Per-launch device times of one eager pass (GPU box): python tools/launch_times.py [precision [config direction B lr_hw]]
(defaults: f16x3 sr_x4 reverse 16 40; with HCF_TC_DEBUG set the results are wrong on purpose -- timing experiments of
the chained conv kernel)."""
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
    cfg = sys.argv[2] if len(sys.argv) > 2 else "sr_x4"
    direction = sys.argv[3] if len(sys.argv) > 3 else "reverse"
    opt = popt.load_config(cfg)
    net = build_net(opt)
    net.load_state_dict(synth.synthetic_state_dict(net.state_dict(), seed=1), strict=True)
    net = net.cuda().eval()
    net.set_precision(prec)
    net.use_graph = False
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    hw = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    eng = net.engine(direction, B, hw, hw, torch.device("cuda", 0))
    s_ = opt["scale"]
    if "lr" in eng.ext:
        eng.ext["lr"].copy_(synth.synthetic_lr(B, hw, hw, seed=0))
    if "hr" in eng.ext:
        eng.ext["hr"].copy_(synth.synthetic_hr(B, hw * s_, hw * s_, seed=0))
    if cfg != "rescaling_x4" and direction == "reverse":
        for i, e in enumerate(synth.synthetic_noise(net.noise_shapes(B, hw, hw), seed=123)):
            eng.ext["eps{}".format(i)].copy_(0.8 * e)
    st = torch.cuda.current_stream()
    acc, count = {}, {}
    for rep in range(5):
        pairs = []
        for (fn, arg, what), info in zip(eng.calls, eng.call_info):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            fn(arg, st.cuda_stream)
            b.record(st)
            pairs.append((info["tag"], a, b))
        torch.cuda.synchronize()
        if rep >= 2:
            per_tag = {}
            for tag, a, b in pairs:      # launches that share a tag (the dense sub-net chains of a level) are summed
                per_tag[tag] = per_tag.get(tag, 0.0) + a.elapsed_time(b)
                count[tag] = count.get(tag, 0) + 1
            for tag, v in per_tag.items():
                acc.setdefault(tag, []).append(v)
    thr = float(os.environ.get("LT_MIN_MS", "0.05"))
    out = {t: round(sum(v) / len(v), 3) for t, v in acc.items() if sum(v) / len(v) > thr}
    out = {("{} (x{})".format(t, count[t] // 3) if count[t] > 3 else t): v for t, v in out.items()}
    out["n_launches"] = len(eng.calls)
    out["fallbacks"] = eng.fallbacks
    out["total"] = round(sum(sum(v) / len(v) for v in acc.values()), 3)
    print(json.dumps({"debug": os.environ.get("HCF_TC_DEBUG", "0"), "rings": os.environ.get("HCF_TC_RINGS", ""), "ms": out}))


if __name__ == "__main__":
    main()
