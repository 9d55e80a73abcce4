"""Data-parallel training step on real NCCL: gradient all-reduce overlapped with backward.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/nccl_grad_check.py

Every rank runs the NLL forward + backward (hcflow_b200/autograd.py: CUDA forward and backward kernels) on its shard of a
global batch; hcflow_b200.dist.GradientReducer all-reduces the gradients bucket by bucket while backward is still
running.  Rank 0 compares every parameter's reduced gradient with torch autograd over the oracle on the CONCATENATED
batch on the CPU (the reference's optimize_parameters under DDP, HCFlow_SR_model.py:33-36,195-203) and prints one JSON
line.  Exit code 0 = match."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hcflow_b200 import dist as hd, options as popt, synth  # noqa: E402
from hcflow_b200.arch import build_net  # noqa: E402
from oracle import hcflow_oracle as orc  # noqa: E402


def main():
    rank, local, world = hd.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    opt = popt.shrink_config(popt.load_config("sr_x4"), K=4, after=[2, 2], rrdb_nb=[1, 1])
    net = build_net(opt)
    sd = synth.synthetic_state_dict(net.state_dict(), seed=3)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).train()
    per, h = 2, 8
    Bg = per * world
    lr = synth.synthetic_lr(Bg, h, h, seed=91)
    hr = synth.synthetic_hr(Bg, 4 * h, 4 * h, seed=92)
    dq = torch.rand(hr.shape, generator=torch.Generator().manual_seed(93), dtype=torch.float32)
    lo, hi = hd.shard_range(Bg, rank, world)
    red = hd.GradientReducer(net.parameters(), bucket_bytes=256 << 10)     # small buckets: several in flight on this net
    _, nll = net(hr=hr[lo:hi].to(dev), lr=lr[lo:hi].to(dev), u=None, reverse=False, training=True, dequant_noise=dq[lo:hi])
    nll.backward()
    early = red.launched_early
    n_buckets = red.finish()
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        sd_r = {k: v.clone().requires_grad_(v.is_floating_point() and "haar" not in k) for k, v in sd.items()}
        _, nll_ref, _, _ = orc.sr_forward(hr, lr, sd_r, opt, dq)
        nll_ref.backward()
        worst, bad, n = 0.0, [], 0
        for k, p in net.named_parameters():
            if not p.requires_grad:
                continue
            g_ref = sd_r[k].grad
            sc = float(g_ref.abs().max())
            err = float((p.grad.cpu() - g_ref).abs().max())
            n += 1
            worst = max(worst, err / (sc + 1e-12)) if sc > 1e-9 else worst
            if err > 2e-3 * sc + 1e-6:
                bad.append((k, err, sc))
        ok = not bad and n > 100 and early > 0
        print(json.dumps({"world": world, "backend": torch.distributed.get_backend() if world > 1 else "none",
                          "global_batch": Bg, "params_checked": n, "worst_param_grad_rel": worst, "buckets": n_buckets,
                          "buckets_reduced_during_backward": early, "bad": bad[:4], "ok": ok}), flush=True)
    hd.barrier()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
