"""Worker of the script-level drop-in test (a subprocess; TEST INFRASTRUCTURE).  Runs the UNMODIFIED reference script
codes/test_HCFlow.py (its option parser, image-folder dataset + dataloader, create_model, HCFlowSRModel.test(),
tensor2img, PSNR / SSIM, PNG writer) on a generated two-image dataset with hcflow_b200.install() as the only addition
(INTEGRATION.md section 1), then checks the PNGs it wrote against the same samples computed directly.
argv[1] = repo root, argv[2] = scratch directory."""
import glob
import os
import runpy
import sys

ROOT, TMP = sys.argv[1], sys.argv[2]
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
import yaml  # noqa: E402

from oracle import ref_loader  # noqa: E402

ref_loader.load()                                       # reference packages + the import-time stubs on sys.path
import hcflow_b200  # noqa: E402
from hcflow_b200 import options, synth  # noqa: E402
from hcflow_b200.arch import build_net  # noqa: E402

# ---- a two-image dataset (HR 64x64, bicubic LR 16x16) and a checkpoint in the reference's format
os.makedirs(os.path.join(TMP, "HR"), exist_ok=True)
os.makedirs(os.path.join(TMP, "LR"), exist_ok=True)
hr = synth.synthetic_hr(2, 64, 64, seed=3)
for i in range(2):
    img = (hr[i].permute(1, 2, 0).numpy()[:, :, ::-1] * 255.0).round().clip(0, 255).astype(np.uint8)     # BGR uint8
    cv2.imwrite(os.path.join(TMP, "HR", "img{}.png".format(i)), img)
    cv2.imwrite(os.path.join(TMP, "LR", "img{}.png".format(i)), cv2.resize(img, (16, 16), interpolation=cv2.INTER_CUBIC))
opt0 = options.shrink_config(options.load_config("sr_x4"), K=8, after=[4, 4])
net = build_net(opt0)
sd = synth.synthetic_state_dict(net.state_dict(), seed=1)
ckpt = os.path.join(TMP, "synthetic_G.pth")
torch.save(sd, ckpt)
with open(os.path.join(ref_loader.REF_CODES, "options", "test", "test_SR_DF2K_4X_HCFlow.yml")) as f:
    y = yaml.safe_load(f)
y["name"] = "hcflow_b200_dropin_script_test"
y["datasets"] = {"test0": {"name": "synthetic2", "mode": "GTLQ", "dataroot_GT": os.path.join(TMP, "HR"),
                           "dataroot_LQ": os.path.join(TMP, "LR")}}
y["network_G"]["flowDownsampler"]["K"] = 8
y["network_G"]["flowDownsampler"]["splitOff"]["after_flowstep"] = [4, 4]
y["val"] = {"heats": [0.0, 0.8], "n_sample": 2}
y["path"]["pretrain_model_G"] = ckpt
yml = os.path.join(TMP, "test_dropin.yml")
with open(yml, "w") as f:
    yaml.safe_dump(y, f)

# ---- the one-line binding, then the unmodified script
hcflow_b200.install()
script = os.path.join(ref_loader.REF_CODES, "test_HCFlow.py")
os.chdir(ref_loader.REF_CODES)
sys.argv = [script, "--opt", yml]
ns = runpy.run_path(script, run_name="__main__")

# ---- what it wrote, against the same samples computed directly
res = os.path.join(ns["opt"]["path"]["results_root"], "synthetic2")
pngs = sorted(glob.glob(os.path.join(res, "SR_*.png")))
assert len(pngs) == 2 * 2 * 2, pngs
model = ns["model"]
assert type(model.netG.module).__module__.startswith("hcflow_b200")
net = model.netG.module.eval()
util = ns["util"]
lr0 = cv2.imread(os.path.join(TMP, "LR", "img0.png"), cv2.IMREAD_UNCHANGED).astype(np.float32) / 255.0
lr0 = torch.from_numpy(np.ascontiguousarray(lr0[:, :, ::-1].transpose(2, 0, 1))).unsqueeze(0).cuda()
with torch.no_grad():
    want = net(lr=lr0, z=None, u=None, eps_std=0.0, reverse=True, training=False)
want_png = util.tensor2img(want[0].float().cpu())
got_png = cv2.imread(os.path.join(res, "SR_img0_0.0_0.png"), cv2.IMREAD_UNCHANGED)
assert got_png.shape == (64, 64, 3) and np.array_equal(got_png, want_png), int(np.abs(got_png.astype(int) - want_png.astype(int)).max())
assert np.array_equal(got_png, cv2.imread(os.path.join(res, "SR_img0_0.0_1.png"), cv2.IMREAD_UNCHANGED))       # heat 0: same image
assert not np.array_equal(cv2.imread(os.path.join(res, "SR_img0_0.8_0.png")), cv2.imread(os.path.join(res, "SR_img0_0.8_1.png")))
print("test_HCFlow.py ran on the drop-in:", len(pngs), "PNGs, heat-0 output identical to the direct call; OK")
