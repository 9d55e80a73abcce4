"""Worker of the training drop-in tests (a subprocess: it puts the UNMODIFIED reference on sys.path).  TEST INFRASTRUCTURE.
argv[1] = repo root.
 (1) ActNorm data initialisation (ActNorms.py:29-43): the reference net (stock PyTorch, fp32 convs) and this package's
     net start from the same weights with every ActNorm bias / logs zeroed, in train mode; one forward NLL pass must
     leave the same ActNorm parameters, the same NLL and the same gradients.
 (2) the reference's own training loop on the drop-in: options.parse(train YAML) -> create_model -> HCFlowSRModel ->
     feed_data -> optimize_parameters(step) x 3 (HCFlow_SR_model.py:184-218: NLL, backward, gradient clipping, Adam):
     the logged NLL equals a direct evaluation before the step and the loss goes down."""
import copy
import os
import sys

ROOT = sys.argv[1]
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import ref_loader  # noqa: E402

networks = ref_loader.load()
from hcflow_b200 import options, synth  # noqa: E402
from hcflow_b200.arch import build_net  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
opt = options.shrink_config(options.load_config("sr_x4"), K=8, after=[4, 4])     # shallow flow: fp32 noise stays small
ref = networks.define_G(copy.deepcopy(opt), 0)            # the reference's class (install() not called yet)
ours = build_net(copy.deepcopy(opt))
assert type(ref).__module__.startswith("models.modules") and type(ours).__module__.startswith("hcflow_b200")
sd = synth.synthetic_state_dict(ours.state_dict(), seed=1)
for k in sd:
    if ".actnorm." in k:
        sd[k] = torch.zeros_like(sd[k])
ref.load_state_dict(sd, strict=True)
ours.load_state_dict(sd, strict=True)
ref, ours = ref.to(dev).train(), ours.to(dev).train()
B = 4
lr = synth.synthetic_lr(B, 16, 16, seed=2).to(dev)
hr = synth.synthetic_hr(B, 64, 64, seed=3).to(dev)
torch.manual_seed(5)
_, nll_ref = ref(hr=hr, lr=lr, u=None, reverse=False)
torch.manual_seed(5)
_, nll = ours(hr=hr, lr=lr, u=None, reverse=False)
rel = abs(float(nll) - float(nll_ref.mean())) / abs(float(nll_ref.mean()))
sr, so = ref.state_dict(), ours.state_dict()
worst, n_an = 0.0, 0
for k in sd:
    if ".actnorm." in k:
        n_an += 1
        assert float(so[k].abs().max()) > 0, k                         # it was initialised from data
        worst = max(worst, float((so[k] - sr[k]).abs().max()) / (1.0 + float(sr[k].abs().max())))
assert all(m.inited for m in ours.modules() if type(m).__name__ == "ActNorm2d")
nll_ref.mean().backward()
nll.backward()
pr, po = dict(ref.named_parameters()), dict(ours.named_parameters())
# gradients: with freshly centred activations many ReLU / LeakyReLU inputs sit at +-1 ulp of zero, so two fp32
# implementations disagree on a few gates and single tensors differ by percents (measured: the reference's cuDNN path is
# off by 7.7e-2 on one conv weight, this package by 7.7e-3) -- the referee is fp64 autograd over the oracle with the
# data-initialised weights and the same dequantisation noise
from oracle import hcflow_oracle as orc  # noqa: E402
torch.manual_seed(5)
dq = torch.rand(hr.shape, device=hr.device)
sd64 = {k: v.detach().double().clone().requires_grad_(v.is_floating_point() and "haar" not in k) for k, v in so.items()}
_, nll64, _, _ = orc.sr_forward(hr.double(), lr.double(), sd64, opt, dq.double())
nll64.backward()
stats = {}
for nm, gsrc in (("ours", po), ("reference", pr)):
    worst_t, num, den2 = (0.0, ""), 0.0, 0.0
    for k in po:
        if gsrc[k].grad is None or sd64[k].grad is None:
            continue
        g64 = sd64[k].grad.view_as(gsrc[k].grad)
        d = (gsrc[k].grad.double() - g64)
        e = float(d.abs().max()) / (float(g64.abs().max()) + 1e-12)
        worst_t = max(worst_t, (e, k))
        num += float((d * d).sum())
        den2 += float((g64 * g64).sum())
    stats[nm] = (worst_t[0], worst_t[1], (num / den2) ** 0.5)
    print("gradients vs fp64 autograd, {}: worst tensor {:.1e} ({}), global L2 {:.1e}".format(nm, *stats[nm]))
print("actnorm data init: {} tensors, worst rel {:.2e}; nll rel {:.2e} (fp64 oracle nll {:.6f})".format(n_an, worst, rel, float(nll64)))
assert n_an > 50 and worst < 2e-4 and rel < 1e-5, (n_an, worst, rel)
assert abs(float(nll) - float(nll64)) < 1e-5 * abs(float(nll64))
assert stats["ours"][0] < 3e-2 and stats["ours"][2] < 2e-3, stats
assert stats["ours"][2] <= 2.0 * stats["reference"][2] + 1e-4, stats       # no worse than the reference's own fp32 path

# ---- (2) the reference's training loop on the drop-in
import hcflow_b200  # noqa: E402
hcflow_b200.install()
from options import options as ref_options  # noqa: E402  (the reference's option parser)
from models import create_model  # noqa: E402
yml = os.path.join(ref_loader.REF_CODES, "options", "train", "train_SR_DF2K_4X_HCFlow.yml")
topt = ref_options.parse(yml, is_train=True)
topt["dist"] = False
topt["path"]["resume_state"] = None
topt["path"]["training_state"] = "/tmp/hcflow_b200_no_such_dir"
topt = ref_options.dict_to_nonedict(topt)
topt["network_G"]["flowDownsampler"]["K"] = 8                     # the shallow flow of the stress fixtures: seconds, not minutes
topt["network_G"]["flowDownsampler"]["splitOff"]["after_flowstep"] = [4, 4]
model = create_model(topt)
net = model.netG.module
assert type(net).__module__.startswith("hcflow_b200"), type(net)
net.load_state_dict(synth.synthetic_state_dict(net.state_dict(), seed=1), strict=True)
model.feed_data({"LQ": lr.cpu(), "GT": hr.cpu()})
losses = []
for step in range(200, 203):                                     # past act_norm_start_step, like a fine-tuning run
    torch.manual_seed(100 + step)
    with torch.no_grad():
        net.eval()
        _, before = net(hr=hr, lr=lr, u=None, reverse=False, training=False)
        net.train()
    torch.manual_seed(100 + step)
    model.optimize_parameters(step)
    logged = model.get_current_log()["nll"]
    assert abs(logged - float(before)) < 2e-4 * abs(float(before)), (step, logged, float(before))
    losses.append(logged)
assert losses[-1] < losses[0], losses
print("reference optimize_parameters on the drop-in: nll", ["{:.4f}".format(v) for v in losses], "OK")
