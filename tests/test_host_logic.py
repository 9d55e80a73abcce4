"""CPU tests of the host side: C-ABI exports, state_dict layout, the install() hook,
option handling, weight packing and the world_size-2 sharding / NLL reduction (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

import hcflow_b200
from hcflow_b200 import _lib, options as popt, prep, synth
from hcflow_b200.arch import HCFlowNet_Rescaling, HCFlowNet_SR, build_net
from tests.helpers import net_and_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hcflow_b200.h")).read()
    declared = set(re.findall(r"\b(hcf_[a-z0-9_]+)\s*\(", header))
    declared -= {"hcf_conv_tc_plan"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()  # raises if the .so is missing or a symbol is not exported
    assert lib.hcf_abi_version() == _lib.ABI_VERSION
    assert lib.hcf_last_error() is not None


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhcflow_b200.so")
    with pytest.raises(_lib.HcfError):
        _lib.load()


def test_no_cpu_fallback():
    opt, net, sd = net_and_weights("sr_x4")
    with pytest.raises(RuntimeError):
        net(lr=torch.zeros(1, 3, 8, 8), eps_std=0.0, reverse=True)


@pytest.mark.parametrize("cfg,ntensors,nparams", [("sr_x4", 1478, 23232539), ("rescaling_x4", 532, None)])
def test_state_dict_layout(cfg, ntensors, nparams):
    """Tensor / parameter counts of the reference nets; key-by-key equality
    with the reference is enforced when oracle/make_golden.py loads these weights strict=True."""
    opt, net, sd = net_and_weights(cfg)
    assert len(sd) == ntensors
    if nparams:
        assert sum(v.numel() for v in sd.values()) == nparams
    assert "flow.layers.1.actnorm.bias" in sd and tuple(sd["flow.layers.1.actnorm.bias"].shape) == (1, 12, 1, 1)
    assert tuple(sd["flow.level0_condFlow.f.logs"].shape) == (12, 1, 1)


def test_install_hook_resolves_like_the_reference_factory():
    import importlib
    names = hcflow_b200.install()
    try:
        for which, cls in (("HCFlowNet_SR", HCFlowNet_SR), ("HCFlowNet_Rescaling", HCFlowNet_Rescaling)):
            lib = importlib.import_module("models.modules." + which + "_arch")
            target = which.replace("_Net", "").lower()
            found = [c for n, c in lib.__dict__.items() if n.lower() == target]
            assert found == [cls]
    finally:
        hcflow_b200.uninstall()
    assert len(names) == 2


def test_forward_signature_matches_reference():
    import inspect
    sig = inspect.signature(HCFlowNet_SR.forward)
    names = list(sig.parameters)
    assert names[:10] == ["self", "hr", "lr", "z", "u", "eps_std", "add_gt_noise", "step", "reverse", "training"]
    assert sig.parameters["reverse"].default is False and sig.parameters["training"].default is True


def test_pack_conv_weight_roundtrip():
    w = torch.randn(22, 138, 3, 3)
    p = prep.pack_conv_weight(w, [10, 128], 32)
    assert tuple(p.shape) == (9, 16 + 128, 32)
    assert torch.equal(p[4, 3, :22], w[:, 3, 1, 1]) and torch.equal(p[0, 16 + 5, :22], w[:, 15, 0, 0])
    assert float(p[:, 10:16].abs().max()) == 0.0 and float(p[:, :, 22:].abs().max()) == 0.0


def test_derive_inverse_matches_reference_arithmetic():
    sd = {"w": torch.randn(12, 12)}
    assert torch.equal(prep.derive(sd, "w#inv"), torch.inverse(sd["w"].double()).float())


def test_shard_range_partitions():
    from hcflow_b200.dist import shard_range
    for n in (1, 7, 16, 512):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


_WORKER = r"""
import os, sys, torch
sys.path.insert(0, {root!r})
from hcflow_b200 import dist as hd
rank, local, world = hd.init_from_env(backend="gloo")
lo, hi = hd.shard_range(10, rank, world)
nll = torch.arange(10, dtype=torch.float32)[lo:hi]
m = hd.batch_mean_nll(nll)
t = hd.max_over_ranks(1.0 + rank, "cpu")
hd.barrier()
assert abs(float(m) - 4.5) < 1e-6, float(m)
assert t == float(world), t
print("OK", rank)
"""


def test_world_size_2_gloo_nll_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("OK") == 2


_GRAD_WORKER = r"""
import os, sys, torch
sys.path.insert(0, {root!r})
from hcflow_b200 import dist as hd
rank, local, world = hd.init_from_env(backend="gloo")
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
unused = torch.nn.Parameter(torch.ones(5))            # never reached by backward: reduced as zero, stays None
x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
lo, hi = hd.shard_range(8, rank, world)
red = hd.GradientReducer(list(net.parameters()) + [unused], bucket_bytes=300)
assert len(red.buckets) > 2
for it in range(2):                                    # buckets are reusable across steps
    net.zero_grad()
    net(x[lo:hi]).pow(2).mean().backward()
    early = red.launched_early
    red.finish()
    got = [p.grad.clone() for p in net.parameters()]
    net.zero_grad()
    red.remove() if it == 1 else None
assert early > 0
net.zero_grad()
net(x).pow(2).mean().backward()                        # whole batch, no reduction: the hooks are gone
for g, p in zip(got, net.parameters()):
    assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-7), float((g - p.grad).abs().max())
assert unused.grad is None or float(unused.grad.abs().max()) == 0.0
print("OK", rank)
"""


def test_world_size_2_gloo_gradient_reducer(tmp_path):
    """dist.GradientReducer (bucketed all-reduce issued from inside backward): mean of the shard gradients == gradient
    of the whole-batch mean loss; buckets left incomplete by unused parameters are flushed by finish()."""
    script = tmp_path / "gworker.py"
    script.write_text(_GRAD_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29615", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("OK") == 2


_REF_WORKER = r"""
import sys
sys.path.insert(0, {root!r})
from oracle import ref_loader
networks = ref_loader.load()                      # the UNMODIFIED reference factory
import hcflow_b200
hcflow_b200.install()
from hcflow_b200 import options
from hcflow_b200.arch import HCFlowNet_SR, HCFlowNet_Rescaling
for cfg, cls in (("sr_x4", HCFlowNet_SR), ("sr_x8", HCFlowNet_SR), ("rescaling_x4", HCFlowNet_Rescaling)):
    net = networks.define_G(options.load_config(cfg), 0)
    assert type(net) is cls, type(net)
print("OK")
"""


@pytest.mark.skipif(not os.path.isdir("/root/reference/codes"), reason="reference checkout not present")
def test_reference_factory_builds_the_dropin_after_install(tmp_path):
    """with install(), the reference's own networks.define_G returns our classes."""
    script = tmp_path / "w.py"
    script.write_text(_REF_WORKER.format(root=ROOT))
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/codes"), reason="reference checkout not present")
def test_reference_create_model_wraps_the_dropin():
    """The reference's own model factory and wrapper (create_model -> HCFlowSRModel, which wraps the net in
    nn.DataParallel and calls print_network / load) construct on top of the drop-in; the GPU half of this worker
    (feed_data -> test() -> get_current_visuals) runs in tests/test_gpu_parity.py."""
    worker = os.path.join(ROOT, "tests", "ref_model_worker.py")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, worker, ROOT], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "OK" in out.stdout and "HCFlowSRModel" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("cout,kin,ks,split_kin", [(32, 64, 3, 0), (64, 192, 3, 64), (64, 192, 3, 192), (22, 64, 3, 0),
                                                   (64, 64, 1, 0)])
def test_fp16_weight_image_layout(cout, kin, ks, split_kin):
    """hcf_conv_tc16_pack_weights (host code, no GPU): per 64-channel chunk [tap][rows][64 fp16], rows = [hi N ; lo N]
    for the chunks below split_kin, 16-byte groups XOR-swizzled by (row & 7) like a TMA SWIZZLE_128B write, and
    hi + lo / 2048 reproduces the fp32 weight to ~2^-22 relative."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(cout + kin)
    w = (torch.randn(cout, kin, ks, ks, generator=g) * 0.05).contiguous()
    N = (cout + 15) // 16 * 16
    nbytes = lib.hcf_conv_tc16_weight_bytes(kin, cout, ks, split_kin)
    assert nbytes == (kin + split_kin) // 64 * ks * ks * N * 128
    img = torch.zeros(nbytes // 2, dtype=torch.float16)
    assert lib.hcf_conv_tc16_pack_weights(w.data_ptr(), kin, cout, ks, split_kin, img.data_ptr()) == 0
    base = 0
    for kc in range(kin // 64):
        parts = 2 if kc * 64 < split_kin else 1
        NB = N * parts
        blk = img[base:base + ks * ks * NB * 64].view(ks * ks, NB, 8, 8)       # [tap][row][16-byte group][8]
        rows = torch.arange(NB).view(1, NB, 1, 1).expand(ks * ks, NB, 8, 8)
        grp = torch.arange(8).view(1, 1, 8, 1).expand(ks * ks, NB, 8, 8)
        unsw = torch.gather(blk, 2, (grp ^ (rows & 7)))                            # undo the swizzle
        unsw = unsw.reshape(ks * ks, NB, 64).float()
        ref = w[:, kc * 64:(kc + 1) * 64].permute(2, 3, 0, 1).reshape(ks * ks, cout, 64)
        hi = unsw[:, :cout]
        assert torch.equal(hi, ref.half().float())
        assert float(unsw[:, cout:N].abs().max()) == 0.0 if N > cout else True
        if parts == 2:
            lo = unsw[:, N:N + cout]
            assert torch.equal(lo, ((ref - hi) * 2048.0).half().float())
            rel = ((hi + lo / 2048.0) - ref).abs().max() / ref.abs().max()
            assert float(rel) < 2.0 ** -20
        base += ks * ks * NB * 64
    assert base == img.numel()
    assert lib.hcf_conv_tc16_weight_bytes(kin + 1, cout, ks, 0) == 0
    assert lib.hcf_conv_tc16_pack_weights(w.data_ptr(), kin, cout, ks, 32, img.data_ptr()) != 0   # split not a multiple of 64


@pytest.mark.parametrize("cout,kin,ks,passes", [(32, 64, 3, 1), (64, 96, 3, 3), (22, 32, 3, 1), (64, 64, 1, 3)])
def test_tf32_weight_image_layout(cout, kin, ks, passes):
    """hcf_conv_tc_pack_weights (host code): per 32-channel chunk [tap][rows][32 fp32], rows = [raw N ; lo N] for the
    3xTF32 split (lo = w - trunc_tf32(w), the part the tensor core ignores when it reads an fp32 word as TF32),
    16-byte groups XOR-swizzled by (row & 7)."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(cout * 7 + kin)
    w = (torch.randn(cout, kin, ks, ks, generator=g) * 0.05).contiguous()
    N = (cout + 15) // 16 * 16
    parts = 2 if passes == 3 else 1
    NB = N * parts
    nbytes = lib.hcf_conv_tc_weight_bytes(kin, cout, ks, passes)
    assert nbytes == kin // 32 * ks * ks * NB * 128
    img = torch.zeros(nbytes // 4, dtype=torch.float32)
    assert lib.hcf_conv_tc_pack_weights(w.data_ptr(), kin, cout, ks, passes, img.data_ptr()) == 0
    blk = img.view(kin // 32, ks * ks, NB, 8, 4)                                   # [chunk][tap][row][16-byte group][4]
    rows = torch.arange(NB).view(1, 1, NB, 1, 1).expand_as(blk)
    grp = torch.arange(8).view(1, 1, 1, 8, 1).expand_as(blk)
    unsw = torch.gather(blk, 3, (grp ^ (rows & 7))).reshape(kin // 32, ks * ks, NB, 32)
    ref = w.view(cout, kin // 32, 32, ks * ks).permute(1, 3, 0, 2)                 # [chunk][tap][cout][32]
    assert torch.equal(unsw[:, :, :cout], ref)
    if N > cout:
        assert float(unsw[:, :, cout:N].abs().max()) == 0.0
    if parts == 2:
        hi = (ref.contiguous().view(torch.int32) & -8192).view(torch.float32)     # TF32 read = drop 13 mantissa bits
        assert torch.equal(unsw[:, :, N:N + cout], ref - hi)
    assert lib.hcf_conv_tc_weight_bytes(kin + 8, cout, ks, passes) == 0


def test_dataparallel_replicas_share_the_master_and_load_state_dict_invalidates():
    """CPU side of two ADVICE items.  nn.DataParallel replicas carry no Parameters: a replica remembers the module it
    was copied from (its engine cache and weights), a replica of a replica still points at the root, training through a
    replica is refused with a clear error; load_state_dict bumps the weight epoch the engines' signature includes."""
    import torch
    opt, net, sd = net_and_weights("sr_x4")
    rep = net._replicate_for_data_parallel()
    assert rep._master() is net and net._master() is net
    assert rep._replicate_for_data_parallel()._master() is net
    assert rep._engines is net._engines and rep._stores is net._stores     # one cache, keyed by device
    with pytest.raises(RuntimeError, match="one process per GPU"):
        rep._refuse_replica_training()
    with torch.no_grad():
        rep._refuse_replica_training()          # inference through replicas is the supported case
    net._refuse_replica_training()              # a plain module is never refused
    e0 = net._weights_epoch
    net.load_state_dict(sd, strict=True)
    assert net._weights_epoch == e0 + 1
    net.invalidate_weights()
    assert net._weights_epoch == e0 + 2


def test_default_precision_is_the_validated_mode_and_env_overrides_it(monkeypatch):
    """A drop-in user never calls set_precision: the module default is the validated tensor-core mode (f16x3), and
    HCFLOW_PRECISION picks another one for callers that only see the reference's scripts."""
    from hcflow_b200 import options
    from hcflow_b200.arch import build_net
    opt = options.shrink_config(options.load_config("sr_x4"), K=4, after=[2, 2], rrdb_nb=[1, 1])
    monkeypatch.delenv("HCFLOW_PRECISION", raising=False)
    assert build_net(opt).precision == "f16x3"
    monkeypatch.setenv("HCFLOW_PRECISION", "fp32")
    assert build_net(opt).precision == "fp32"
    monkeypatch.setenv("HCFLOW_PRECISION", "bogus")
    with pytest.raises(AssertionError):
        build_net(opt)
