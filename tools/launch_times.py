"""Per-launch device times of one eager inverse pass (GPU box): python tools/launch_times.py [precision]
(with HCF_TC_DEBUG set the results are wrong on purpose -- timing experiments of the chained conv kernel)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hcflow_b200 import options as popt, synth  # noqa: E402
from hcflow_b200.arch import build_net  # noqa: E402
from oracle import hcflow_oracle as orc  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
    opt = popt.load_config("sr_x4")
    net = build_net(opt)
    net.load_state_dict(synth.synthetic_state_dict(net.state_dict(), seed=1), strict=True)
    net = net.cuda().eval()
    net.set_precision(prec)
    net.use_graph = False
    B, hw = 16, 40
    eng = net.engine("reverse", B, hw, hw, torch.device("cuda", 0))
    eng.ext["lr"].copy_(synth.synthetic_lr(B, hw, hw, seed=0))
    for i, e in enumerate(synth.synthetic_noise(orc.noise_shapes(opt, B, hw, hw, True), seed=123)):
        eng.ext["eps{}".format(i)].copy_(0.8 * e)
    st = torch.cuda.current_stream()
    acc = {}
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
            for tag, a, b in pairs:
                acc.setdefault(tag, []).append(a.elapsed_time(b))
    out = {t: round(sum(v) / len(v), 3) for t, v in acc.items() if sum(v) / len(v) > 0.05}
    out["total"] = round(sum(sum(v) / len(v) for v in acc.values()), 3)
    print(json.dumps({"debug": os.environ.get("HCF_TC_DEBUG", "0"), "rings": os.environ.get("HCF_TC_RINGS", ""), "ms": out}))


if __name__ == "__main__":
    main()
