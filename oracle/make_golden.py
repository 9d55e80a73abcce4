"""Generate tests/golden/*.pt by running the UNMODIFIED reference (authoring container only).

TEST INFRASTRUCTURE.  Usage:  python oracle/make_golden.py

For each case the reference net is built with ``networks.define_G`` (the same entry
the reference's model wrappers use, codes/models/networks.py:36-41) from the bundled
config, loaded ``strict=True`` with ``hcflow_b200.synth.synthetic_state_dict`` (which
also proves the product's key layout), ActNorm marked initialised as
``HCFlowSRModel.load`` does (codes/models/HCFlow_SR_model.py:462-465), and run in
eval / no_grad / fp32 on CPU.  The two RNG draws on the path are replaced by
committed-seed tensors: ``GaussianDiag.sample`` (Basic.py:96-100) consumes
``eps_std * synthetic_noise`` and ``torch.rand`` in ``normal_flow_diracLR``
(HCFlowNet_SR_arch.py:52) returns the synthetic dequantisation noise.

Only outputs + a weight fingerprint are stored (weights are regenerated from the
seed on the test box).  Also stores per-module vectors for the oracle's unit pins.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from hcflow_b200 import modules as pm  # noqa: E402
from hcflow_b200 import options as popt  # noqa: E402
from hcflow_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle import hcflow_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (config, B, lr_h, lr_w, eps_std)
    "sr_x4": ("sr_x4", 2, 12, 12, 0.8),
    "sr_x8": ("sr_x8", 1, 8, 8, 0.8),
    "rescaling_x4": ("rescaling_x4", 1, 10, 10, 1.0),
    # stress fixtures (hcflow_b200.synth.STRESS): coupling outputs h of order 1 like a trained flow's, on a flow shallow
    # enough not to diverge -- the regular fixtures keep every coupling within 2 % of the identity, which hides
    # operand-rounding error of the coupling sub-nets
    "sr_x4_stress": ("sr_x4", 2, 12, 12, 0.8),
    "sr_x8_stress": ("sr_x8", 1, 8, 8, 0.8),
    "rescaling_x4_stress": ("rescaling_x4", 1, 10, 10, 1.0),
}


def build_reference(networks, opt):
    torch.manual_seed(0)
    np.random.seed(0)
    net = networks.define_G(opt, 0)
    net.eval()
    return net


def mark_inited(net):
    for m in net.modules():
        if hasattr(m, "inited"):
            m.inited = True


class NoiseQueue:
    def __init__(self, tensors):
        self.q = list(tensors)

    def sample(self, mean, logs, eps_std=None):
        eps = self.q.pop(0)
        assert eps.shape == mean.shape, (eps.shape, mean.shape)
        return mean + torch.exp(logs) * eps


def check_yaml_matches_reference(name, opt):
    """The bundled config's network section must equal the reference YAML's."""
    import yaml
    ref_yaml = {"sr_x4": "test_SR_DF2K_4X_HCFlow.yml", "sr_x8": "test_SR_CelebA_8X_HCFlow.yml",
                "rescaling_x4": "test_Rescaling_DF2K_4X_HCFlow.yml"}[name]
    with open(os.path.join(ref_loader.REF_CODES, "options", "test", ref_yaml)) as f:
        ref = yaml.safe_load(f)
    mine = {k: v for k, v in opt["network_G"].items() if k != "scale"}
    assert dict(ref["network_G"]) == _plain(mine), (ref["network_G"], mine)
    assert ref["scale"] == opt["scale"] and ref.get("quant") == opt.get("quant")


def _plain(d):
    if isinstance(d, dict):
        return {k: _plain(v) for k, v in d.items()}
    if isinstance(d, list):
        return [_plain(v) for v in d]
    return d


def main():
    networks = ref_loader.load()
    from models.modules import Basic as RB  # reference module
    os.makedirs(GOLD, exist_ok=True)
    only = set(sys.argv[1:])
    for name, (cfg, B, h, w, heat) in CASES.items():
        if only and name not in only:
            continue
        opt = popt.load_config(cfg)
        check_yaml_matches_reference(cfg, opt)
        stress = synth.STRESS.get(name)
        if stress:
            opt = popt.shrink_config(opt, K=stress["K"], after=stress["after"])
        SR = opt["network_G"]["which_model_G"] == "HCFlowNet_SR"
        net = build_reference(networks, opt)
        sd = synth.synthetic_state_dict(net.state_dict(), seed=1)
        if stress:
            sd = synth.stress_state_dict(sd, stress["s_weight"], stress["s_bias"], s_prior_mean=stress["s_prior_mean"])
        net.load_state_dict(sd, strict=True)
        mark_inited(net)
        scale = opt["scale"]
        lr = synth.synthetic_lr(B, h, w, seed=0)
        hr = synth.synthetic_hr(B, h * scale, w * scale, seed=0)
        shapes = orc.noise_shapes(opt, B, h, w, SR)
        unit = synth.synthetic_noise(shapes, seed=123)
        eps = [heat * e for e in unit]
        out = {"fingerprint": synth.fingerprint(sd), "B": B, "h": h, "w": w, "heat": heat, "config": cfg}

        orig_sample = RB.GaussianDiag.sample
        orig_rand = torch.rand
        try:
            q = NoiseQueue(eps)
            RB.GaussianDiag.sample = staticmethod(q.sample)
            with torch.no_grad():
                # raw (un-clamped) HR from the inner FlowNet, and the public clamped output
                raw = net.flow(z=lr, u=None, eps_std=heat, reverse=True, training=False)
                q.q = list(eps)
                hr_out = net(lr=lr, z=None, u=None, eps_std=heat, reverse=True, training=False)
                # heat 0 (deterministic branch of the YAML's heats)
                q.q = [0.0 * e for e in unit]
                hr0 = net(lr=lr, z=None, u=None, eps_std=0.0, reverse=True, training=False)
            out["inv_raw"], out["inv_hr"], out["inv_hr_heat0"] = raw.clone(), hr_out.clone(), hr0.clone()
            assert len(q.q) == 0

            if SR:
                dq = torch.rand(hr.shape, generator=torch.Generator().manual_seed(77), dtype=torch.float32)
                torch.rand = lambda *a, **k: dq.clone()
                with torch.no_grad():
                    fake_lr, nll = net(hr=hr, lr=lr, u=None, reverse=False, training=False)
                    quant = opt["quant"]
                    x = hr + dq / quant
                    ld0 = torch.zeros(B) + float(-np.log(quant) * hr.shape[2] * hr.shape[3])
                    z, logdet = net.flow(hr=x, u=None, logdet=ld0, reverse=False, training=False)
                torch.rand = orig_rand
                out.update(fwd_fake_lr=fake_lr.clone(), fwd_nll=nll.clone(), fwd_z=z.clone(),
                           fwd_logdet=logdet.clone())
            else:
                with torch.no_grad():
                    flr, z1, z2 = net(hr=hr, reverse=False, training=False)
                    zraw, _, _ = net.flow(hr=hr, u=None, logdet=None, reverse=False, training=False)
                out.update(fwd_fake_lr=flr.clone(), fwd_z1=z1.clone(), fwd_z2=z2.clone(), fwd_raw_lr=zraw.clone())
        finally:
            RB.GaussianDiag.sample = orig_sample
            torch.rand = orig_rand

        # per-module pins (reference sub-modules called directly)
        with torch.no_grad():
            g = torch.Generator().manual_seed(5)
            mods = {}
            step = net.flow.layers[1]
            C = step.actnorm.bias.shape[1]
            zt = torch.randn(B, C, 8, 8, generator=g)
            ld = torch.zeros(B)
            zf, ldf = step(zt.clone(), None, logdet=ld.clone(), reverse=False)
            zr, _ = step(zt.clone(), None, reverse=True)
            mods["step_in"], mods["step_fwd"], mods["step_fwd_logdet"], mods["step_rev"] = zt, zf.clone(), ldf.clone(), zr.clone()
            cfl = net.flow.level0_condFlow
            cs = cfl.additional_flow_steps[0]
            Cz = cs.actnorm.bias.shape[1]
            zt2 = torch.randn(B, Cz, 8, 8, generator=g)
            ncond = cfl.f.weight.shape[1]
            ut = 0.3 * torch.randn(B, ncond, 8, 8, generator=g)
            zf2, ldf2 = cs(zt2.clone(), u=ut, logdet=ld.clone(), reverse=False)
            zr2, _ = cs(zt2.clone(), u=ut, reverse=True)
            mods["cstep_in"], mods["cstep_u"] = zt2, ut
            mods["cstep_fwd"], mods["cstep_fwd_logdet"], mods["cstep_rev"] = zf2.clone(), ldf2.clone(), zr2.clone()
            xr = 0.5 * torch.randn(B, 64, 8, 8, generator=g)
            mods["rrdb_in"], mods["rrdb_out"] = xr, cfl.RRDB_trunk0[0](xr.clone()).clone()
            lastlvl = getattr(net.flow, "level{}_condFlow".format(net.flow.L - 1))
            ulr = torch.rand(B, 3, 8, 8, generator=g)
            feat = lastlvl.get_conditional_feature_SR(ulr) if SR else lastlvl.get_conditional_feature_Rescaling(ulr)
            mods["feat_in"], mods["feat_out"] = ulr, feat.clone()
            xs = torch.randn(B, 3, 8, 8, generator=g)
            mods["squeeze_in"] = xs
            mods["squeeze_out"] = RB.squeeze2d(xs, 2).clone()
            if not SR:
                haar = net.flow.layers[0]
                hf, _ = haar(xs.clone(), reverse=False)
                hb, _ = haar(hf.clone(), reverse=True)
                mods["haar_fwd"], mods["haar_rev"] = hf.clone(), hb.clone()
        out["modules"] = mods
        path = os.path.join(GOLD, name + ".pt")
        torch.save(out, path)
        hmax = []
        if True:   # coupling strength of the fixture (max |h| over the sub-net calls of the inverse pass), for the record
            f0, d0 = orc.fcn, orc.dense_block

            def rec(fn):
                def wrapped(*a, **k):
                    y = fn(*a, **k)
                    hmax.append(float(y.abs().max()))
                    return y
                return wrapped
            orc.fcn, orc.dense_block = rec(f0), rec(d0)
            try:
                with torch.no_grad():
                    orc.flownet_reverse(lr, sd, opt, eps, SR)
            finally:
                orc.fcn, orc.dense_block = f0, d0
            out["h_absmax"] = max(hmax)
            torch.save(out, path)
        print("wrote", path, os.path.getsize(path), "bytes; |h|max {:.3f};".format(max(hmax)),
              "inv_raw range [{:.3f}, {:.3f}]".format(float(out["inv_raw"].min()), float(out["inv_raw"].max())),
              "fwd_nll" if SR else "", float(out["fwd_nll"]) if SR else "")


if __name__ == "__main__":
    main()
