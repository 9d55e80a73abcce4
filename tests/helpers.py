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
    """(opt, product net on CPU with synthetic weights loaded, state_dict)."""
    key = (cfg, seed)
    if key not in _CACHE:
        opt = popt.load_config(cfg)
        net = build_net(opt)
        sd = synth.synthetic_state_dict(net.state_dict(), seed=seed)
        net.load_state_dict(sd, strict=True)
        net.eval()
        _CACHE[key] = (opt, net, sd)
    return _CACHE[key]


def is_sr(opt):
    return opt["network_G"]["which_model_G"] == "HCFlowNet_SR"


def maxabs(a, b):
    return float((a.double() - b.double()).abs().max())
