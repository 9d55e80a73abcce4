"""In-kernel cycle profile of the chained conv launches (GPU box):
    HCF_TC_PROF=1 python tools/prof_chain.py [precision] [HCF_TC_DEBUG]
prints, per chain, the average cycles per work item each warp role spends in each wait class."""
import gc
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HCF_TC_PROF", "1")
if len(sys.argv) > 2:
    os.environ["HCF_TC_DEBUG"] = sys.argv[2]
import torch  # noqa: E402

from hcflow_b200 import options as popt, synth  # noqa: E402
from hcflow_b200.arch import build_net  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
    opt = popt.load_config("sr_x4")
    net = build_net(opt)
    net.load_state_dict(synth.synthetic_state_dict(net.state_dict(), seed=1), strict=True)
    net = net.cuda().eval()
    net.set_precision(prec)
    net.use_graph = False
    B, hw = 16, 40
    eng = net.engine("reverse", B, hw, hw, torch.device("cuda", 0))
    eng.ext["lr"].copy_(synth.synthetic_lr(B, hw, hw, seed=0))
    for i, e in enumerate(synth.synthetic_noise(net.noise_shapes(B, hw, hw), seed=123)):
        eng.ext["eps{}".format(i)].copy_(0.8 * e)
    for _ in range(4):
        eng.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        eng.run()
    e1.record()
    torch.cuda.synchronize()
    print("precision {} debug {} eager ms/step {:.3f}".format(prec, os.environ.get("HCF_TC_DEBUG", "0"),
                                                            e0.elapsed_time(e1) / 4), file=sys.stderr)
    net.clear_engines()
    del eng
    gc.collect()


if __name__ == "__main__":
    main()
