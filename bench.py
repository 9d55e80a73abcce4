#!/usr/bin/env python
"""Benchmark of the HCFlow inverse (sampling) pass -- BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision f16x3|f16|tf32x3|tf32|fp32]
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores

Workload (configs[1] of BASELINE.json): 4x SR, B=16 synthetic 40x40 LR tiles per GPU ->
160x160 HR, temperature 0.8, synthetic weights (hcflow_b200.synth, seed 1).  With N>1 every
rank runs its own shard of a 16*N batch (weak scaling, no data-path collective: images are
independent) and the time is the max over ranks.

One JSON line on stdout (rank 0).  `value` = device-resident throughput (inputs already in
HBM, whole pass replayed as one CUDA graph, per-step CUDA events, L2 flushed between steps);
`e2e` = the same metric through the public module call with pinned-host LR in and HR out.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "HR megapixels/sec inverse pass (4x SR)"
UNIT = "MP/s"
LR_HW = 40
SCALE = 4
B_PER_GPU = 16
HEAT = 0.8
WORKLOAD = "configs[1]: 4x SR inverse, B=16/GPU, 40x40 LR -> 160x160 HR, T=0.8"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.mask, self.stop_flag, self.max_mhz, self.ok = index, [], 0, False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": [n for b, n in self.REASONS.items() if self.mask & b]}


# --------------------------------------------------------------------------- reference arms
def reference_net(device):
    """The UNMODIFIED reference net (codes/models/networks.py:36-41 define_G) with the synthetic weights, from the copy
    oracle/build_ref.py staged under oracle/_ref (or /root/reference in the authoring container).  None if absent."""
    from hcflow_b200 import options as popt, synth
    from oracle import ref_loader
    if not ref_loader.available():
        return None
    import numpy as np
    networks = ref_loader.load()
    opt = popt.load_config("sr_x4")
    torch.manual_seed(0)
    np.random.seed(0)
    net = networks.define_G(opt, 0)
    net.load_state_dict(synth.synthetic_state_dict(net.state_dict(), seed=1), strict=True)
    for m in net.modules():     # HCFlowSRModel.load marks ActNorm initialised (HCFlow_SR_model.py:462-465)
        if hasattr(m, "inited"):
            m.inited = True
    return net.to(device).eval()


def cpu_inverse_mps(batch, budget_s, warmup, steps=None):
    """Times the reference's CPU path for the inverse pass on all host threads: the unmodified reference modules when
    they are staged (kind "reference"; test_HCFlow.py's call, HCFlow_SR_model.py:308-312), else the oracle port."""
    from hcflow_b200 import options as popt, synth
    from hcflow_b200.arch import build_net
    from oracle import hcflow_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lr = synth.synthetic_lr(batch, LR_HW, LR_HW, seed=0)
    ref = reference_net("cpu")
    if ref is not None:
        kind = "reference"

        def one():
            return ref(lr=lr, z=None, u=None, eps_std=HEAT, reverse=True, training=False)
    else:
        kind = "port"
        opt = popt.load_config("sr_x4")
        sd = synth.synthetic_state_dict(build_net(opt).state_dict(), seed=1)
        eps = [HEAT * e for e in synth.synthetic_noise(orc.noise_shapes(opt, batch, LR_HW, LR_HW, True))]

        def one():
            return orc.sr_reverse(lr, sd, opt, eps)
    times = []
    with torch.no_grad():
        for _ in range(warmup):
            one()
        t_end = time.perf_counter() + budget_s
        while True:
            t0 = time.perf_counter()
            one()
            times.append(time.perf_counter() - t0)
            if steps is not None and len(times) >= steps:
                break
            if steps is None and (time.perf_counter() > t_end or len(times) >= 50):
                break
    mp = batch * (LR_HW * SCALE) ** 2 / 1e6
    return mp, times, cores, kind


def gpu_eager_baseline(dev, lr, steps=5):
    """SURVEY 2a's comparator: the unmodified reference modules on the SAME B200 through stock PyTorch / cuDNN (TF32
    convs allowed = PyTorch's default), same batch, device-resident input, CUDA events.  None if not staged."""
    ref = reference_net(dev)
    if ref is None:
        return None
    torch.backends.cudnn.allow_tf32 = True
    x = lr.to(dev)
    with torch.no_grad():
        for _ in range(2):
            ref(lr=x, z=None, u=None, eps_std=HEAT, reverse=True, training=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            out = ref(lr=x, z=None, u=None, eps_std=HEAT, reverse=True, training=False)
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    B = x.shape[0]
    del ref
    torch.cuda.empty_cache()
    return {"value": B * (LR_HW * SCALE) ** 2 / 1e6 / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "what": "unmodified reference modules (define_G) on this GPU, stock PyTorch {} eager, cuDNN TF32 convs "
                    "allowed, B={}".format(torch.__version__, B), "finite": bool(torch.isfinite(out).all())}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = B_PER_GPU    # the workload's own batch: one step = one full configs[1] pass on the host cores
    mp, times, cores, kind = cpu_inverse_mps(batch, budget_s=0, warmup=1, steps=args.steps)
    total = sum(times)
    value = mp * len(times) / total
    sample = "{} timed passes of B={} (the full workload batch), 40x40 LR -> 160x160 HR, T=0.8".format(len(times), batch)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "same_config": True,
        "note": ("the UNMODIFIED reference modules (networks.define_G, staged by oracle/build_ref.py) on torch CPU fp32, "
                 "all host threads" if kind == "reference" else
                 "reference copy not staged: oracle/hcflow_oracle.py (same torch CPU ops, bit-exact vs the reference goldens)"),
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def conv_flops(plan, B):
    from hcflow_b200 import plan as P
    total = 0.0
    for op in plan.ops:
        if isinstance(op, P.ConvOp):
            cin = sum(v.C for v, _ in op.segs)
            total += 2.0 * B * op.H * op.W * op.ks * op.ks * cin * op.cout
    return total


def ncu_dram_bytes(precision):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch (chained encoder kernel, 80x80 level) from the
    newest committed `ncu --set full` summary of this precision mode under profiles/ -> (bytes, file name) or (None, None)."""
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_prof_chain_L0_{}_summary.csv".format(precision))))
    for path in reversed(cands):
        tot, unit_scale = 0.0, {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        found = 0
        with open(path) as f:
            for line in f:
                parts = line.strip().split(",")
                if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(parts[2]) * unit_scale.get(parts[1], 1.0)
                    found += 1
        if found == 2:
            return tot, os.path.basename(path)
    return None, None


def parity_from_report():
    """Measured parity numbers of the newest committed GPU parity report (written by pytest -m gpu) -- never literals."""
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_parity_report.json")))
    if not cands:
        return None
    with open(cands[-1]) as f:
        rep = json.load(f)
    keep = {}
    for k, v in rep.items():
        if k.startswith(("e2e_reverse", "stress_reverse", "config", "e2e_forward", "stress_forward")):
            keep[k] = v
    return {"file": os.path.basename(cands[-1]), "measured": keep}


def per_class_times(eng):
    """One eager (non-graph) pass with a CUDA-event pair around every launch, on the launching
    stream; returns {class: (n_launches, ms, flops)} and fills per_class_times.by_tag."""
    st = torch.cuda.current_stream()
    pairs = []
    if eng.plan.uses_logdet:
        eng.logdet.copy_(eng.logdet_init)
    for (fn, arg, what), info in zip(eng.calls, eng.call_info):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        rc = fn(arg, st.cuda_stream)
        b.record(st)
        assert rc == 0, what
        pairs.append((info["cls"], a, b, info["flops"], info["tag"]))
    torch.cuda.synchronize()
    out, by_tag = {}, {}
    for cls, a, b, fl, tag in pairs:
        n, ms, f = out.get(cls, (0, 0.0, 0.0))
        out[cls] = (n + 1, ms + a.elapsed_time(b), f + fl)
        if cls.startswith("conv"):
            n, ms, f = by_tag.get(tag, (0, 0.0, 0.0))
            by_tag[tag] = (n + 1, ms + a.elapsed_time(b), f + fl)
    per_class_times.by_tag = by_tag
    return out


def run_configs4(net, opt, dev, rank, world, hd, flush, steps):
    """configs[4] as BASELINE.json states it: 4x SR, global batch 64 x N sharded over the N ranks (64 per GPU; the
    reference's batch_size // world_size, codes/data/__init__.py:13-14), inverse pass, plus one forward NLL step whose
    batch mean (HCFlowNet_SR_arch.py:65 nll.mean()) is the path's only collective: ONE all-reduce of (sum nll_i, count)
    over NCCL, inside the timed region.  Every rank takes part; times are the max over ranks."""
    from hcflow_b200 import synth
    B4, HR = 64, LR_HW * SCALE
    lr = synth.synthetic_lr(B4, LR_HW, LR_HW, seed=100 + rank)
    hr = synth.synthetic_hr(B4, HR, HR, seed=100 + rank)
    unit = synth.synthetic_noise(net.noise_shapes(B4, LR_HW, LR_HW), seed=200 + rank)
    er = net.engine("reverse", B4, LR_HW, LR_HW, dev)
    er.ext["lr"].copy_(lr)
    for i, e in enumerate(unit):
        er.ext["eps{}".format(i)].copy_(HEAT * e)
    for _ in range(3):
        er.run()
    torch.cuda.synchronize()
    hd.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(1.0)
        a.record()
        er.run()
        b.record()
    torch.cuda.synchronize()
    hd.barrier()
    inv_ms = hd.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), dev) / steps
    # forward NLL with the all-reduce of the batch mean
    ef = net.engine("forward", B4, LR_HW, LR_HW, dev)
    ef.ext["hr"].copy_(hr)
    ef.ext["lr"].copy_(lr)
    ef.ext["dequant"].copy_(torch.rand(hr.shape, generator=torch.Generator().manual_seed(300 + rank)))
    ln2hw = float(math.log(2.0) * HR * HR)

    def fwd_step():
        ef.run()
        per_image = (-ef.logdet) / ln2hw            # fp64 [B4]: the reference's per-image nll before .mean()
        return hd.batch_mean_nll(per_image)
    for _ in range(3):
        nll = fwd_step()
    torch.cuda.synchronize()
    hd.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(steps)]
    for a, m, b in ev:
        flush.fill_(1.0)
        a.record()
        ef.run()
        m.record()
        nll = hd.batch_mean_nll((-ef.logdet) / ln2hw)
        b.record()
    torch.cuda.synchronize()
    hd.barrier()
    fwd_ms = hd.max_over_ranks(sum(a.elapsed_time(b) for a, _, b in ev), dev) / steps
    ar_us = hd.max_over_ranks(sorted(m.elapsed_time(b) for _, m, b in ev)[steps // 2], dev) * 1e3
    mp = world * B4 * HR * HR / 1e6
    out = {"workload": "configs[4]: 4x SR, global batch {} = 64/GPU x {} GPU(s), 160x160 HR".format(B4 * world, world),
           "inverse": {"value": mp / (inv_ms / 1e3), "unit": UNIT, "ms_per_step": inv_ms, "steps": steps},
           "forward_nll": {"value": mp / (fwd_ms / 1e3), "unit": UNIT, "ms_per_step": fwd_ms, "steps": steps,
                           "batch_mean_nll": float(nll), "allreduce_us_median": ar_us,
                           "collective": "torch.distributed all_reduce(SUM) of 2 doubles, backend {}".format(
                               torch.distributed.get_backend() if torch.distributed.is_initialized() else "none (1 rank)")}}
    net.clear_engines()
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    from hcflow_b200 import dist as hd, options as popt, synth
    from hcflow_b200.arch import build_net
    rank, local, world = hd.init_from_env()
    assert world == args.gpus or world == 1, (world, args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    opt = popt.load_config("sr_x4")
    net = build_net(opt)
    net.load_state_dict(synth.synthetic_state_dict(net.state_dict(), seed=1), strict=True)
    net = net.to(dev).eval()
    net.set_precision(args.precision)
    net.use_graph = not args.no_graph
    B = B_PER_GPU
    HR = LR_HW * SCALE
    lr = synth.synthetic_lr(B, LR_HW, LR_HW, seed=rank)
    unit = synth.synthetic_noise(net.noise_shapes(B, LR_HW, LR_HW), seed=123 + rank)
    eng = net.engine("reverse", B, LR_HW, LR_HW, dev)
    eng.ext["lr"].copy_(lr)
    for i, e in enumerate(unit):
        eng.ext["eps{}".format(i)].copy_(HEAT * e)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    eng.lib.hcf_launch_count_reset()
    eng.run()  # eager warm-up + graph capture
    launches_per_step = int(eng.lib.hcf_launch_count()) // (1 if args.no_graph else 2)
    for _ in range(max(args.warmup, 3)):
        eng.run()
    torch.cuda.synchronize()

    # ---- device-resident timing
    sampler = ClockSampler(local)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    hd.barrier()
    torch.cuda.synchronize()
    sampler.start()
    for k in range(args.steps):
        flush.fill_(float(k))
        starts[k].record()
        eng.run()
        ends[k].record()
    torch.cuda.synchronize()
    hd.barrier()
    sampler.stop_flag = True
    t_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t_ms = hd.max_over_ranks(t_ms, dev)
    mp_step = world * B * HR * HR / 1e6
    value = mp_step * args.steps / (t_ms / 1e3)

    if args.skip_e2e:   # profiling runs (ncu): only the device-resident loop
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "profiling_run": True,
                              "ms_per_step": t_ms / args.steps, "dtype": args.precision}), flush=True)
        return
    # ---- end to end through the public module call, pinned host buffers
    # a serving loop: every step copies its LR batch from pinned host memory, calls the module, and reads the HR batch
    # back into pinned host memory; the read-back runs on a copy stream (two host buffers) under the next step's compute
    lr_host = lr.pin_memory()
    hr_host = [torch.empty(B, 3, HR, HR, dtype=torch.float32).pin_memory() for _ in range(2)]
    lr_dev = torch.empty_like(lr, device=dev)
    copy_stream = torch.cuda.Stream(dev)
    main_stream = torch.cuda.current_stream(dev)
    done = [None, None]

    def e2e_step(i):
        lr_dev.copy_(lr_host, non_blocking=True)
        out = net(lr=lr_dev, z=None, u=None, eps_std=HEAT, reverse=True, training=False)
        if done[i & 1] is not None:
            done[i & 1].synchronize()          # the host buffer's previous read-back has landed (host-side consumer)
        copy_stream.wait_stream(main_stream)
        with torch.cuda.stream(copy_stream):
            hr_host[i & 1].copy_(out, non_blocking=True)
            out.record_stream(copy_stream)
            done[i & 1] = torch.cuda.Event()
            done[i & 1].record(copy_stream)

    with torch.no_grad():
        for i in range(4):
            e2e_step(i)
        torch.cuda.synchronize()
        hd.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            e2e_step(i)
        main_stream.wait_stream(copy_stream)      # the last read-back is inside the timed region
        e1.record()
        torch.cuda.synchronize()
        hd.barrier()
    e2e_ms = hd.max_over_ranks(e0.elapsed_time(e1), dev)
    e2e_value = mp_step * args.steps / (e2e_ms / 1e3)

    cfg4 = None
    if world > 1 or args.configs4:
        cfg4 = run_configs4(net, opt, dev, rank, world, hd, flush, steps=max(3, min(args.steps, 10)))

    if rank != 0:
        return
    # ---- roofline of the dominant kernel class (eager pass, per-launch events)
    net.use_graph = False
    eng2 = net.engine("reverse", B, LR_HW, LR_HW, dev)
    eng2.ext["lr"].copy_(lr)
    for i, e in enumerate(unit):
        eng2.ext["eps{}".format(i)].copy_(HEAT * e)
    per_class_times(eng2)
    classes = per_class_times(eng2)
    total_ms = sum(v[1] for c, v in classes.items() if c != "conv_tc_all")
    # dominant kernel = the single largest launch (the chained tcgen05 encoder kernel of the 80x80 level);
    # the aggregate over every tcgen05 conv launch is reported next to it
    tc_keys = [c for c in classes if c.startswith("conv_tc")]
    peaks = load_peaks()
    by_tag = per_class_times.by_tag
    top_tag = max(by_tag, key=lambda t: by_tag[t][1] / by_tag[t][0])
    n, ms, fl = by_tag[top_tag]
    achieved = fl / (ms / 1e3) / 1e12
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/README.md)
    ncu_traffic, ncu_file = ncu_dram_bytes(args.precision) if "chain" in top_tag and "enc." in top_tag else (None, None)
    f16_mode = args.precision in ("f16", "f16x3")
    # the launch is timed inside the step (per-launch events of an eager pass): the sustained figure applies
    peak_used = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
    roofline = {
        "bound": "tensor", "kernel": "conv_tc_kernel " + top_tag, "achieved": achieved, "peak": peak_used,
        "unit": "TFLOP/s", "frac": achieved / peak_used, "traffic": ncu_traffic, "peak_burst": peaks["bf16_tflops"],
        "traffic_note": "dram__bytes_read+write per launch read from profiles/{} (ncu --set full of this launch); "
                        "algorithmic bytes of that launch = 2.34e9".format(ncu_file),
        "peak_source": "{} bf16 sustained from MEASURED_PEAKS.json ({})".format(
            peaks["source"], "the kernel computes on fp16 operands: same tensor rate; split layers issue 2 MMAs per useful "
            "K-step" if f16_mode else "the kernel computes in TF32, nominal peak = half of it; tf32x3 issues 2 MMAs per "
            "useful K-step"),
        # what actually binds this launch: the SM's shared-memory data pipe -- per (tile, layer) work item
        # ~3.5 k MMA operand wavefronts (ncu: 605 M per launch) + ~1.3 k TMA fill + ~0.4 k staging wavefronts against
        # ~1.5 k cycles of tensor work at the nominal rate, because an RDB conv has only N = 32 / 64 output channels
        "binding_resource": {
            "what": "shared-memory operand pipe of an implicit GEMM with N = 32 / 64 (not HBM, not instruction issue)",
            "smem_wavefronts_per_item": 5300, "nominal_tensor_cycles_per_item": 1500,
            "formulation_ceiling_frac_of_sustained_peak": 0.49,
            "frac_of_formulation_ceiling": (achieved / peak_used) / 0.49,
            "evidence": "profiles/r02_prof_chain_L0_f16x3_summary.csv, r02_wait_profile_f16x3.log, r02_ws_*, "
                        "r02_unrolled_issue_*, r02_split_producers_single_plane_*"},
        "launches": n, "avg_launch_ms": ms / n, "share_of_step": ms / total_ms,
        "classes": {c: {"n": v[0], "ms": round(v[1], 4), "gflop": round(v[2] / 1e9, 3)} for c, v in classes.items()},
        "conv_by_layer": {t: {"n": v[0], "ms": round(v[1], 3), "tflops": round(v[2] / max(v[1], 1e-9) / 1e9, 1)}
                          for t, v in sorted(by_tag.items(), key=lambda kv: -kv[1][1])},
    }
    # the fused FlowStep launches (north_star: per fused FlowStep, fraction of the HBM roofline): algorithmic bytes of
    # SURVEY 8d's table (x4 inverse, per image: conditional / main FlowSteps of both levels = 14.1 + 4.0 + 46.6 + 8.0 MB)
    fs = [(t, v) for t, v in by_tag.items() if "fcn." in t or t.startswith("flowsteps")]
    if fs:
        fs_ms = sum(v[1] for _, v in fs)
        fs_bytes = B * (680 * 13 * 1600 + 192 * 13 * 1600 + 560 * 13 * 6400 + 96 * 13 * 6400)
        roofline["flowstep_chains"] = {
            "bound": "hbm", "launches": len(fs), "ms": round(fs_ms, 3), "algorithmic_bytes": fs_bytes,
            "achieved": fs_bytes / (fs_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": fs_bytes / (fs_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
            "flops": sum(v[2] for _, v in fs), "tflops": sum(v[2] for _, v in fs) / (fs_ms / 1e3) / 1e12,
            "frac_tensor": sum(v[2] for _, v in fs) / (fs_ms / 1e3) / 1e12 / peak_used,
            "by_launch": {t: round(v[1], 3) for t, v in fs},
            "note": "52 FlowSteps: per level one launch of shared-conditioning convs (W_u * u of all 13 conditional steps) and two "
                    "fused-FlowStep launches (csrc/flowstep_tc.cu: one work item = sub-net + ActNorm / 1x1 / coupling tail of one "
                    "step on one tile); algorithmic bytes = SURVEY 8d's per-FlowStep figures; ncu summaries under "
                    "profiles/r02_prof_flowstep_*"}
    if tc_keys:
        a_n, a_ms, a_fl = (sum(classes[c][k] for c in tc_keys) for k in range(3))
        roofline["all_tcgen05_convs"] = {"launches": a_n, "ms": round(a_ms, 3), "tflops": round(a_fl / a_ms / 1e9, 1),
                                         "frac": a_fl / a_ms / 1e9 / peak_used, "share_of_step": a_ms / total_ms}
    # ---- CPU baseline (rank 0, N=1 only): the oracle on the host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        mp, times, cores, kind = cpu_inverse_mps(1, budget_s=12.0, warmup=1)
        cpu = {"value": mp * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "{} passes of configs[0] (B=1, 40x40 LR -> 160x160 HR, T=0.8), {} on torch CPU fp32".format(
                   len(times), "unmodified reference modules" if kind == "reference" else "oracle port")}
    eager = None
    if world == 1 and not args.no_eager_baseline:
        eager = gpu_eager_baseline(dev, lr)
    # ---- the other precision modes, device-resident, short (same inputs, same graph-replay method)
    modes = {args.precision: {"value": value, "ms_per_step": t_ms / args.steps}}
    if world == 1 and not args.no_modes:
        net.use_graph = True
        for prec in ("f16x3", "f16", "tf32x3", "tf32", "fp32"):
            if prec in modes:
                continue
            net.set_precision(prec)
            e3 = net.engine("reverse", B, LR_HW, LR_HW, dev)
            e3.ext["lr"].copy_(lr)
            for i, e in enumerate(unit):
                e3.ext["eps{}".format(i)].copy_(HEAT * e)
            for _ in range(3):
                e3.run()
            torch.cuda.synchronize()
            n3 = 5
            s3 = [torch.cuda.Event(enable_timing=True) for _ in range(n3)]
            f3 = [torch.cuda.Event(enable_timing=True) for _ in range(n3)]
            for k in range(n3):
                flush.fill_(float(k))
                s3[k].record()
                e3.run()
                f3[k].record()
            torch.cuda.synchronize()
            ms3 = sum(a.elapsed_time(b) for a, b in zip(s3, f3)) / n3
            modes[prec] = {"value": B * HR * HR / 1e6 / (ms3 / 1e3), "ms_per_step": ms3}
            net.clear_engines()
            torch.cuda.empty_cache()
        net.set_precision(args.precision)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": "dp{}".format(world),
                   "l2": "256 MiB flush between timed steps; per-step CUDA events",
                   "conv_kernels": {"tcgen05": eng.n_tc, "fp32": eng.n_fp32_conv, "chained_launches": eng.n_chains},
                   "conv_tflop_per_step": conv_flops(eng.plan, B) / 1e12},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(lr.numel() * 4),
                "d2h_bytes_per_step": int(hr_host[0].numel() * 4), "ms_per_step": e2e_ms / args.steps,
                "note": "pinned-host LR in, HR out through net(lr=..., reverse=True) every step; the read-back of step i runs on a copy stream under step i+1 (two pinned buffers), the last one inside the timed region"},
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "clocks": sampler.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "gpu_eager_baseline": eager,
        "modes": modes,
        "parity": parity_from_report(),
    }
    if cfg4 is not None:
        line["configs4"] = cfg4
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("HCFLOW_PRECISION", "f16x3"),
                    choices=["fp32", "tf32", "tf32x3", "tf32x3_all", "f16", "f16x3"],
                    help="f16x3 (default): tcgen05 on fp16 hi/lo operand planes, split (hi+lo on both operands) for the "
                         "convs that write the encoder's residual stream, fp32-level parity (7e-5 measured, 2e-3 "
                         "tolerance); f16: one fp16 pass everywhere (1.7e-3); tf32x3 / tf32: the same on fp32 words read "
                         "as TF32; fp32: CUDA-core kernels")
    ap.add_argument("--no-modes", action="store_true", help="skip the short runs of the other precision modes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the stock-PyTorch run of the reference modules on the GPU")
    ap.add_argument("--configs4", action="store_true", help="also run configs[4]'s per-GPU shard (B=64) at N=1 (always on for N>1)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches (for ncu)")
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    with torch.no_grad():
        run_ours(args)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
