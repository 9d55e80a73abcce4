"""Worker of the drop-in tests (run as a subprocess: it puts the UNMODIFIED reference on sys.path): the reference's own
model factory and wrapper -- create_model(opt) -> HCFlowSRModel (codes/models/__init__.py:40-55, HCFlow_SR_model.py:18-36,
which wraps the net in nn.DataParallel over every visible GPU) -> feed_data -> test() (:296-316: forward NLL, then
heats x n_sample inverse passes) -> get_current_visuals -- with hcflow_b200.install() as the only change, compared with
the same calls made directly on the module; then the same for HCFlowRescalingModel (HCFlow_Rescaling_model.py:306-324).
TEST INFRASTRUCTURE.  argv[1] = repo root."""
import sys, os
ROOT = sys.argv[1]
sys.path.insert(0, ROOT)
import torch
from oracle import ref_loader
ref_loader.load()                                   # UNMODIFIED reference on sys.path (models, utils, options)
import hcflow_b200
hcflow_b200.install()
from hcflow_b200 import options, synth
from hcflow_b200.arch import HCFlowNet_SR
from models import create_model                     # the reference's own model factory (codes/models/__init__.py)

gpu = torch.cuda.is_available()
opt = options.load_config("sr_x4")
opt["model"] = "HCFlow_SR"
opt["is_train"] = False
opt["dist"] = False
opt["gpu_ids"] = list(range(torch.cuda.device_count())) if gpu else None
opt["path"] = {"pretrain_model_G": None, "strict_load": True, "resume_state": None, "root": "/tmp", "models": "/tmp",
               "training_state": "/tmp"}
opt["val"] = {"heats": [0.0, 0.8], "n_sample": 2, "sr_mode": "bicubic"}
model = create_model(opt)
net = model.netG.module
assert type(net) is HCFlowNet_SR, type(net)
net.load_state_dict(synth.synthetic_state_dict(net.state_dict(), seed=1), strict=True)
print("model", type(model).__name__, "netG", type(model.netG).__name__, "->", type(net).__name__, "device", model.device)
if not gpu:
    from hcflow_b200.arch import HCFlowNet_Rescaling
    ropt = options.load_config("rescaling_x4")
    ropt["model"] = "HCFlow_Rescaling"
    for k in ("is_train", "dist", "gpu_ids", "path", "val"):
        ropt[k] = opt[k]
    rmodel = create_model(ropt)
    assert type(rmodel.netG.module) is HCFlowNet_Rescaling
    print("model", type(rmodel).__name__, "->", type(rmodel.netG.module).__name__)
    print("OK (construction only: no GPU)")
    sys.exit(0)
B = 4
lr = synth.synthetic_lr(B, 16, 16, seed=2)
hr = synth.synthetic_hr(B, 64, 64, seed=3)
model.feed_data({"LQ": lr, "GT": hr})
torch.manual_seed(7)
nll = model.test()                                  # HCFlow_SR_model.py:296-316: forward NLL, then heats x n_sample inverse passes
assert sorted(model.fake_H) == [(0.0, 0), (0.0, 1), (0.8, 0), (0.8, 1)]
vis = model.get_current_visuals()
assert vis["LQ"].shape == (3, 16, 16) and vis[("SR", 0.8, 1)].shape == (3, 64, 64)
# the same calls made directly on the module, in the same order with the same seed (same RNG stream)
net.eval()
with torch.no_grad():
    torch.manual_seed(7)
    _, nll2 = net(hr=hr.cuda(), lr=lr.cuda(), u=None, reverse=False, training=False)
    direct = {}
    for heat in (0.0, 0.8):
        for s in range(2):
            direct[(heat, s)] = net(lr=lr.cuda(), z=None, u=None, eps_std=heat, reverse=True, training=False)
one_gpu = torch.cuda.device_count() == 1
for k, v in direct.items():
    got = model.fake_H[k]
    assert torch.isfinite(got).all() and got.shape == v.shape
    if one_gpu or k[0] == 0.0:          # (several GPUs: each replica draws its own noise, only heat 0 is comparable)
        assert torch.equal(got, v), (k, float((got - v).abs().max()))
assert not torch.equal(model.fake_H[(0.8, 0)], model.fake_H[(0.8, 1)])      # two samples, two draws
assert torch.equal(model.fake_H[(0.0, 0)], model.fake_H[(0.0, 1)])
if one_gpu:
    assert abs(nll - float(nll2)) < 1e-6 * abs(nll), (nll, float(nll2))
print("nll", nll, "SR wrapper ok")

# ---- the rescaling wrapper (HCFlow_Rescaling_model.py:306-324): encode, quantise the LR, decode at every heat
from hcflow_b200.arch import HCFlowNet_Rescaling
ropt = options.load_config("rescaling_x4")
ropt["model"] = "HCFlow_Rescaling"
for k in ("is_train", "dist", "gpu_ids", "path", "val"):
    ropt[k] = opt[k]
rmodel = create_model(ropt)
rnet = rmodel.netG.module
assert type(rnet) is HCFlowNet_Rescaling, type(rnet)
rnet.load_state_dict(synth.synthetic_state_dict(rnet.state_dict(), seed=1), strict=True)
rmodel.feed_data({"LQ": lr, "GT": hr})
torch.manual_seed(11)
z1_mean = rmodel.test()
rnet.eval()
with torch.no_grad():
    torch.manual_seed(11)
    fl, z1, z2 = rnet(hr=hr.cuda(), lr=lr.cuda(), u=None, reverse=False, training=False)
    flq = rmodel.Quantization(fl)
    assert torch.equal(rmodel.fake_L_from_H, flq)
    d0 = rnet(lr=flq, z=None, u=None, eps_std=0.0, reverse=True, training=False)
assert torch.equal(rmodel.fake_H[(0.0, 0)], d0) and torch.equal(rmodel.fake_H[(0.0, 1)], d0)
assert abs(z1_mean - float(z1.mean())) < 1e-6 * max(1.0, abs(z1_mean))
assert all(torch.isfinite(v).all() for v in rmodel.fake_H.values())
assert rmodel.get_current_visuals()[("SR", 0.8, 1)].shape == (3, 64, 64)
print("rescaling wrapper ok; OK")
