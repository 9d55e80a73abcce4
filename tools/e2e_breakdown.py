"""Where does the end-to-end step spend its time beyond the graph replay? (GPU box)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from hcflow_b200 import options as popt, synth  # noqa: E402
from hcflow_b200.arch import build_net  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    t_issue = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_issue


def main():
    opt = popt.load_config("sr_x4")
    net = build_net(opt)
    net.load_state_dict(synth.synthetic_state_dict(net.state_dict(), seed=1), strict=True)
    net = net.cuda().eval()
    net.set_precision(sys.argv[1] if len(sys.argv) > 1 else "f16x3")
    B, hw = 16, 40
    lr = synth.synthetic_lr(B, hw, hw)
    lr_host = lr.pin_memory()
    hr_host = torch.empty(B, 3, 160, 160).pin_memory()
    lr_dev = torch.empty_like(lr, device="cuda")
    eng = net.engine("reverse", B, hw, hw, torch.device("cuda", 0))
    with torch.no_grad():
        def full():
            lr_dev.copy_(lr_host, non_blocking=True)
            out = net(lr=lr_dev, eps_std=0.8, reverse=True)
            hr_host.copy_(out, non_blocking=True)

        def no_copies():
            net(lr=lr_dev, eps_std=0.8, reverse=True)

        def replay_only():
            eng.run()

        def draw_only():
            net._draw_eps(eng, 0.8, None, lr_dev.device)

        def sig_only():
            eng.weight_signature()
        for name, fn in (("full e2e step", full), ("module call, no host copies", no_copies), ("graph replay only", replay_only),
                         ("eps draw only", draw_only), ("weight signature only", sig_only)):
            gpu, cpu = timeit(fn)
            print("{:32s} device {:7.3f} ms/step   host issue {:7.3f} ms/step".format(name, gpu, cpu), flush=True)


if __name__ == "__main__":
    main()
