"""rescaling round trip: graph vs eager, engines alone and together (GPU box)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hcflow_b200 import options as popt, synth
from hcflow_b200.arch import build_net
from bench_configs import timed

prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
opt = popt.load_config("rescaling_x4")
net = build_net(opt)
net.load_state_dict(synth.synthetic_state_dict(net.state_dict(), seed=1), strict=True)
net = net.cuda().eval()
net.set_precision(prec)
B, hw = 64, 64
out = {}
for graph in (True,):
    net.use_graph = graph
    net.clear_engines() if hasattr(net, "clear_engines") else None
    with torch.no_grad():
        ef = net.engine("forward", B, hw, hw, torch.device("cuda", 0))
        ef.ext["hr"].copy_(synth.synthetic_hr(B, 4 * hw, 4 * hw).cuda())
        out["fwd_graph%d" % graph] = round(timed(ef.run, 5), 3)
        er = net.engine("reverse", B, hw, hw, torch.device("cuda", 0))
        er.ext["lr"].copy_(ef.ext["fake_lr"])
        out["rev_graph%d" % graph] = round(timed(er.run, 5), 3)
        out["fwd_again_graph%d" % graph] = round(timed(ef.run, 5), 3)
        def both():
            ef.run(); er.ext["lr"].copy_(ef.ext["fake_lr"]); er.run()
        out["both_graph%d" % graph] = round(timed(both, 5), 3)
    out["mem_gb_graph%d" % graph] = round(torch.cuda.max_memory_allocated() / 2**30, 2)
    tags = {}
    for e, nm in ((ef, "fwd"), (er, "rev")):
        for info in e.call_info:
            t = info["tag"]
            if t.startswith("chain"):
                tags.setdefault(nm, []).append(t)
    out["fallbacks"] = ef.fallbacks + er.fallbacks
print(json.dumps(out))
