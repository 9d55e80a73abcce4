"""Shared test helpers: golden loading, synthetic weights for a config, oracle runs."""
import os

import torch

from hcflow_b200 import options as popt
from hcflow_b200 import synth
from hcflow_b200.arch import build_net

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLD_DIR, name + ".pt"), weights_only=False)


_CACHE = {}


def net_and_weights(cfg, seed=1):
    """(opt, product net on CPU with synthetic weights loaded, state_dict).  ``cfg`` may name a stress fixture
    (synth.STRESS: shallow flow, coupling outputs of order 1)."""
    key = (cfg, seed)
    if key not in _CACHE:
        stress = synth.STRESS.get(cfg)
        opt = popt.load_config(stress["cfg"] if stress else cfg)
        if stress:
            opt = popt.shrink_config(opt, K=stress["K"], after=stress["after"])
        net = build_net(opt)
        sd = synth.synthetic_state_dict(net.state_dict(), seed=seed)
        if stress:
            sd = synth.stress_state_dict(sd, stress["s_weight"], stress["s_bias"], s_prior_mean=stress["s_prior_mean"])
        net.load_state_dict(sd, strict=True)
        net.eval()
        _CACHE[key] = (opt, net, sd)
    return _CACHE[key]


def is_sr(opt):
    return opt["network_G"]["which_model_G"] == "HCFlowNet_SR"


def maxabs(a, b):
    return float((a.double() - b.double()).abs().max())
