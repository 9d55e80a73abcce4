"""Measure end-to-end error of each precision mode against the reference goldens (GPU box)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hcflow_b200 import synth  # noqa: E402
from hcflow_b200.arch import build_net  # noqa: E402
from oracle import hcflow_oracle as orc  # noqa: E402
from tests.helpers import is_sr, load_golden, maxabs, net_and_weights  # noqa: E402


def main():
    precisions = sys.argv[1:] or ["fp32", "tf32x3", "tf32"]
    for cfg in ["sr_x4", "sr_x8", "rescaling_x4"]:
        g = load_golden(cfg)
        opt, _, sd = net_and_weights(cfg)
        B, h, w, heat = g["B"], g["h"], g["w"], g["heat"]
        s = opt["scale"]
        lr, hr = synth.synthetic_lr(B, h, w), synth.synthetic_hr(B, h * s, w * s)
        unit = synth.synthetic_noise(orc.noise_shapes(opt, B, h, w, is_sr(opt)))
        for prec in precisions:
            net = build_net(opt)
            net.load_state_dict(sd, strict=True)
            net = net.cuda().eval()
            net.set_precision(prec)
            with torch.no_grad():
                net(lr=lr.cuda(), eps_std=heat, reverse=True, eps=unit)
                raw = net.last["hr_raw"].cpu()
                eng = list(net._engines.values())[0]
                res = {"cfg": cfg, "precision": prec, "tc_convs": eng.n_tc, "fp32_convs": eng.n_fp32_conv,
                       "inv_raw_maxabs": maxabs(raw, g["inv_raw"]),
                       "inv_raw_mean": float((raw.double() - g["inv_raw"].double()).abs().mean())}
                if is_sr(opt):
                    dq = torch.rand(hr.shape, generator=torch.Generator().manual_seed(77), dtype=torch.float32)
                    _, nll = net(hr=hr.cuda(), lr=lr.cuda(), reverse=False, dequant_noise=dq)
                    res["fwd_z_maxabs"] = maxabs(net.last["z_raw"].cpu(), g["fwd_z"])
                    res["nll_rel"] = abs(float(nll) - float(g["fwd_nll"])) / abs(float(g["fwd_nll"]))
                else:
                    flr, z1, z2 = net(hr=hr.cuda(), reverse=False)
                    res["fwd_lr_maxabs"] = maxabs(net.last["z_raw"].cpu(), g["fwd_raw_lr"])
                    res["fwd_z1_maxabs"] = maxabs(z1.cpu(), g["fwd_z1"])
            print(json.dumps(res), flush=True)
            del net
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
