"""configs[4]'s only collective, on real NCCL: the batch NLL mean over a batch sharded across ranks.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_nll_check.py

Every rank runs the forward (NLL) pass on its shard of a global batch (codes/data/__init__.py:13-14: batch // world)
through the CUDA engine and hcflow_b200.dist.batch_mean_nll all-reduces (sum nll_i, count) over NCCL; rank 0 compares
the result with the oracle's nll.mean() over the CONCATENATED batch (HCFlowNet_SR_arch.py:65) on the CPU and prints
one JSON line.  Exit code 0 = match."""
import json
import math
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
    opt = popt.load_config("sr_x4")
    net = build_net(opt)
    sd = synth.synthetic_state_dict(net.state_dict(), seed=1)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    net.set_precision(os.environ.get("HCFLOW_PRECISION", "f16x3"))
    Bg, h = 2 * world + 1, 12          # uneven shards on purpose: (sum, count) must weight them correctly
    lr = synth.synthetic_lr(Bg, h, h, seed=7)
    hr = synth.synthetic_hr(Bg, 4 * h, 4 * h, seed=7)
    dq = torch.rand(hr.shape, generator=torch.Generator().manual_seed(70), dtype=torch.float32)
    lo, hi = hd.shard_range(Bg, rank, world)
    with torch.no_grad():
        net(hr=hr[lo:hi].to(dev), lr=lr[lo:hi].to(dev), reverse=False, training=False, dequant_noise=dq[lo:hi])
        per_image = (-net.last["objective"]) / float(math.log(2.0) * hr.shape[2] * hr.shape[3])
        got = float(hd.batch_mean_nll(per_image))
    ok = True
    if rank == 0:
        with torch.no_grad():
            _, want, _, _ = orc.sr_forward(hr, lr, sd, opt, dq)
        rel = abs(got - float(want)) / abs(float(want))
        ok = rel < 2e-5
        print(json.dumps({"world": world, "backend": torch.distributed.get_backend() if world > 1 else "none",
                          "global_batch": Bg, "shards": [hd.shard_range(Bg, r, world) for r in range(world)],
                          "batch_mean_nll": got, "oracle_nll_mean": float(want), "rel_err": rel, "ok": ok}), flush=True)
    hd.barrier()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
