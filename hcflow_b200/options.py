"""Option handling for the drop-in boundary.

The reference feeds its arch classes an ``opt`` dict parsed from YAML and turned
into a "missing key -> None" dict (reference: codes/options/options.py:10-90,
106-121) and reads nested keys through ``opt_get`` (codes/utils/util.py:1153-1161).
The product only needs the network-relevant part of that plumbing, so that
``HCFlowNet_SR(opt, step)`` can be constructed from either the reference's own
parsed ``opt`` or from one of the bundled configs.
"""
import os
from collections import OrderedDict

import yaml

_CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")

CONFIGS = {
    "sr_x4": "sr_x4.yml",
    "sr_x8": "sr_x8.yml",
    "rescaling_x4": "rescaling_x4.yml",
}


class NoneDict(dict):
    """dict whose missing keys read as None (same contract as the reference's)."""

    def __missing__(self, key):
        return None


def to_nonedict(opt):
    if isinstance(opt, dict):
        return NoneDict(**{k: to_nonedict(v) for k, v in opt.items()})
    if isinstance(opt, list):
        return [to_nonedict(v) for v in opt]
    return opt


def opt_get(opt, keys, default=None):
    """Nested lookup; any missing / None hop yields ``default``."""
    cur = opt
    if cur is None:
        return default
    for k in keys:
        cur = cur.get(k, None) if isinstance(cur, dict) else None
        if cur is None:
            return default
    return cur


def load_config(name_or_path):
    """Load a bundled config by name ("sr_x4", "sr_x8", "rescaling_x4") or a YAML path."""
    path = name_or_path
    if name_or_path in CONFIGS:
        path = os.path.join(_CONFIG_DIR, CONFIGS[name_or_path])
    with open(path, "r") as f:
        raw = yaml.safe_load(f)
    opt = to_nonedict(raw)
    if opt.get("distortion") == "sr" and opt.get("network_G") is not None:
        # the reference copies the top-level scale into network_G (options.py:72-73)
        opt["network_G"]["scale"] = opt.get("scale")
    return opt


def shrink_config(opt, K=None, after=None, rrdb_nb=None):
    """Return a copy of ``opt`` with a shallower flow / encoder (used by fast tests)."""
    import copy

    o = copy.deepcopy(opt)
    fd = o["network_G"]["flowDownsampler"]
    if K is not None:
        fd["K"] = K
    if after is not None:
        fd["splitOff"]["after_flowstep"] = list(after)
    if rrdb_nb is not None:
        fd["splitOff"]["RRDB_nb"] = list(rrdb_nb)
    return o
