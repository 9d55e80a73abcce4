"""Drop-in network classes: HCFlowNet_SR and HCFlowNet_Rescaling.

Same constructor ``(opt, step=None)``, same keyword ``forward`` and same return values as
the reference's codes/models/modules/HCFlowNet_SR_arch.py:12-75 and
HCFlowNet_Rescaling_arch.py:14-54, and the same state_dict layout (modules.py), so the
reference's model wrappers (HCFlow_SR_model.py:195,208,305,311) and checkpoints work
unchanged.  The arithmetic runs in the CUDA engine; the modules refuse to run on a CPU.

Extensions (keyword-only, all optional, ignored by the reference's callers):
  gt=            alias of hr= (the task statement's spelling)
  eps=           list of unit-normal noise tensors [B,Cz,H,W], deepest level first, used
                 instead of the torch.normal draws of Basic.py:96-100 (parity / reproducibility)
  dequant_noise= the U[0,1) tensor used instead of torch.rand of HCFlowNet_SR_arch.py:52
After a call ``net.last`` holds the un-clamped tensors (hr_raw / z_raw / logdet / ...).
"""
import collections
import contextlib
import math
import os
import threading

import torch
from torch import nn

from . import modules as M
from .options import opt_get


# nn.DataParallel runs its replicas in threads of one process: the engine cache of the master module is shared by all of
# them (one engine per device), so cache look-ups, engine construction and the host side of a launch (incl. CUDA-graph
# capture, which is process-global) are serialised; the device work itself is asynchronous and overlaps across the GPUs.
_DP_LOCK = threading.RLock()


class _HCFlowBase(nn.Module):
    SR = True

    def __init__(self, opt, step=None):
        super().__init__()
        self.opt = opt
        hr_size = opt_get(opt, ["datasets", "train", "GT_size"], 160)
        hr_channel = opt_get(opt, ["network_G", "in_nc"], 3)
        self.flow = M.FlowNet((hr_size, hr_size, hr_channel), opt, SR=self.SR)
        # arithmetic of the inference engine: the validated default is "f16x3" (fp16 hi / lo operand planes, 2e-4
        # max-abs parity tolerance on the un-clamped HR); HCFLOW_PRECISION picks another mode for
        # callers that never see the module (the reference's scripts after install()); set_precision() at run time
        self.precision = os.environ.get("HCFLOW_PRECISION", "f16x3")
        assert self.precision in ("fp32", "tf32", "tf32x3", "tf32x3_all", "f16", "f16x3"), self.precision
        self.use_graph = True
        self.use_chains = True   # fuse runs of tensor-core convs into one persistent chained launch
        self.share_cond = True   # tensor-core modes: the sub-nets' shared conditioning conv once per level (engine.py)
        self.fuse_steps = True   # tensor-core modes, inverse pass: FlowStep tail in the last sub-net conv's epilogue
        self.pair_convs = True   # tensor-core modes: RDB growth convs in pairs (rewrite.pair_rdb_convs)
        self.flowchain = True    # fp16 modes: one fused kernel per FlowStep chain (csrc/flowstep_tc.cu)
        # opt-in: when the SAME lr tensor (same storage, same version counter) is sampled again, keep the
        # deepest level's encoder features of the previous call instead of recomputing them
        self.reuse_lr_features = False
        self._last_lr_key = None
        self._last_lr_ref = None
        # compiled engines, least recently used first.  An engine owns activation buffers, fp16 planes, tensor-core plans
        # and a captured graph for ONE (direction, B, h, w): a test loop over variable-size images would otherwise grow
        # GPU memory by gigabytes per new size.  Packed weights are shared (one WeightStore per device).
        self.max_engines = 4
        self._engines = collections.OrderedDict()
        self._stores = {}
        self._weights_epoch = 0
        self.last = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_weights())

    def invalidate_weights(self):
        """Tell the engines that parameter VALUES changed in a way autograd's version counters do not see
        (``p.data.copy_()``, ``p.data.mul_()``, EMA swaps ...).  load_state_dict calls it; optimizers and plain in-place
        ops on the Parameters are detected without it."""
        self._weights_epoch += 1

    def _replicate_for_data_parallel(self):
        """nn.DataParallel over several GPUs (the reference's default wrapper when no launcher is used,
        HCFlow_SR_model.py:33-36): a replica carries no Parameters of its own (torch attaches broadcast copies as plain
        attributes), so it remembers the module it was copied from and runs INFERENCE through that module's engine cache
        -- one engine, one packed-weight store per device, weights read from the master's state_dict.  Training under
        multi-device DataParallel is refused (use one process per GPU: torchrun / DistributedDataParallel,
        hcflow_b200.dist)."""
        replica = super()._replicate_for_data_parallel()
        replica.__dict__["_dp_master"] = self.__dict__.get("_dp_master") or self
        return replica

    def _master(self):
        return self.__dict__.get("_dp_master") or self

    def _guard(self):
        """Replicas (threads of one process) serialise the host side of their calls; a plain module pays nothing."""
        return _DP_LOCK if self.__dict__.get("_dp_master") is not None else contextlib.nullcontext()

    def _refuse_replica_training(self):
        m = self.__dict__.get("_dp_master")
        if m is not None and torch.is_grad_enabled() and any(p.requires_grad for p in m.parameters()):
            raise RuntimeError(
                "hcflow_b200: training through nn.DataParallel replicas on several GPUs is not supported (replicas carry "
                "no Parameters).  Use one process per GPU (torchrun / DistributedDataParallel, hcflow_b200.dist), or call "
                "the wrapped module under torch.no_grad() for inference.")

    # ---- engine management -------------------------------------------------------------
    def set_precision(self, precision):
        """"fp32" (CUDA-core exact), "tf32" (tcgen05, 1 pass), "tf32x3" (tcgen05; 3xTF32 split for the convs that
        write the encoder's residual stream, the prior and the dense sub-nets, 1 pass elsewhere -- see
        Engine._passes_for) or "tf32x3_all" (3xTF32 everywhere).  "f16" / "f16x3" are "tf32" / "tf32x3" with the
        chained encoder convs on FP16 operands (hi / lo planes, kind::f16: half the operand bytes per MAC; the split
        layers of "f16x3" use hi + lo on both operands, ~fp32 accuracy; operand range |x| < 65504)."""
        assert precision in ("fp32", "tf32", "tf32x3", "tf32x3_all", "f16", "f16x3")
        if precision != self.precision:
            self.precision = precision
            self.clear_engines()

    def noise_shapes(self, B, h, w):
        """[(B, C, H, W)] of the prior draws of an inverse pass on an h x w LR input, in draw order (deepest level first:
        the order in which the reference calls torch.normal, ConditionalFlow.py:44-110) -- what `eps=` expects."""
        from .plan import build_plan
        return [(B, c, H, W) for c, H, W in build_plan(self, "reverse", B, h, w).noise_shapes]

    def clear_engines(self):
        for eng in self._engines.values():
            eng.close()
        self._engines.clear()
        self._stores.clear()

    def engine(self, direction, B, h, w, device, io="f32"):
        from .engine import Engine, WeightStore
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        m = self._master()          # replicas share the master's cache (keyed by device) and read its weights
        key = (direction, B, h, w, str(device), self.precision, self.use_graph, self.use_chains, self.share_cond,
               self.fuse_steps, io, self.pair_convs, self.flowchain)
        with self._guard():
            eng = m._engines.get(key)
            if eng is not None:
                m._engines.move_to_end(key)
                return eng
            # the bound is per device: a replica must not evict the engine another GPU is running
            mine = [k for k in m._engines if k[4] == str(device)]
            while len(mine) >= max(1, m.max_engines):
                m._engines.pop(mine.pop(0)).close()          # least recently used of this device
            store = m._stores.setdefault(str(device), WeightStore())
            with torch.cuda.device(device):    # the plans allocate and set function attributes on the CURRENT device
                eng = Engine(m, direction, B, h, w, device, precision=self.precision, use_graph=self.use_graph,
                             use_chains=self.use_chains, share_cond=self.share_cond, fuse_steps=self.fuse_steps, io=io,
                             pair_convs=self.pair_convs, store=store, flowchain=self.flowchain)
            m._engines[key] = eng
            return eng

    def check_status(self):
        """Synchronous check of the engines' device status words (fp16 range guard): raises FP16RangeError /
        HcfError if a previous pass tripped one.  (Every call also checks, without synchronising, the passes that
        have already completed.)"""
        for eng in self._engines.values():
            eng.check_status(wait=True)

    def _check(self, t, name, c=3):
        if t is None:
            raise ValueError("{} is required".format(name))
        if t.dim() != 4 or t.shape[1] != c:
            raise ValueError("{} must be [B,{},H,W], got {}".format(name, c, tuple(t.shape)))
        if not t.is_cuda:
            raise RuntimeError("hcflow_b200 runs on CUDA tensors only (no CPU fallback); got {} on {}".format(
                name, t.device))
        return t.detach().to(torch.float32).contiguous()

    def _draw_eps(self, eng, eps_std, eps, device):
        B = eng.B
        std = 0.0 if eps_std is None else float(eps_std)
        for i, (c, H, W) in enumerate(eng.plan.noise_shapes):
            dst = eng.ext["eps{}".format(i)]
            if eps is not None:
                e = eps[i].to(device=device, dtype=torch.float32)
                assert tuple(e.shape) == (B, c, H, W), (tuple(e.shape), (B, c, H, W))
                dst.copy_(e * std)
            else:
                # the reference's own draw, in its order (Basic.py:96-100): torch.normal(mean=zeros, std=ones * std)
                # is normal_(0, 1) * std + mean inside ATen (normal_out_impl), drawn here without the host
                # synchronisation of its `std >= 0` check (tests/test_gpu_parity.py checks the streams are identical)
                dst.normal_(0.0, 1.0).mul_(std)

    def _reverse(self, lr, eps_std, eps):
        self._refuse_replica_training()
        if self.SR and torch.is_grad_enabled() and lr is not None and lr.is_cuda and (
                any(p.requires_grad for p in self.parameters()) or lr.requires_grad):
            # inverse-path loss of the reference's training loop (HCFlow_SR_model.py:207-218): differentiable graph
            from . import autograd as ag
            with torch.cuda.device(lr.device):
                return ag.sr_reverse(self, lr, eps_std, eps)
        lr_arg = lr
        lr = self._check(lr, "lr")
        B, _, h, w = lr.shape
        eng = self.engine("reverse", B, h, w, lr.device)
        # identity of the CALLER's tensor (not of a converted temporary, whose address the caching allocator recycles),
        # and a reference to it, so that its storage cannot be freed and handed to a different image in between
        key = (id(eng), id(lr_arg), lr_arg.data_ptr(), lr_arg._version, tuple(lr_arg.shape), lr_arg.dtype)
        same = self.reuse_lr_features and key == self._last_lr_key and self._last_lr_ref is lr_arg
        self._last_lr_key = key
        self._last_lr_ref = lr_arg if self.reuse_lr_features else None
        with self._guard(), torch.cuda.device(lr.device):
            eng.ext["lr"].copy_(lr)
            self._draw_eps(eng, eps_std, eps, lr.device)
            eng.run(reuse_lr_features=same)
            self.last = {"hr_raw": eng.ext["hr_raw"].clone()}
            return eng.ext["hr"].clone()


    def sample_many(self, lr, heats, n_sample=1):
        """All samples of one LR batch in ONE pass: what the reference's test loop computes with
        len(heats) * n_sample separate calls (HCFlow_SR_model.py:308-312: for heat: for sample: netG(lr, eps_std=heat,
        reverse=True)).  The LR batch is replicated along the batch axis, every replica gets its own heat, and the noise is
        drawn slice by slice in the loop's order (heat-major, then sample, deepest level first), so a seeded call returns
        exactly the tensors the loop returns.  Returns {(heat, sample): clamp(fake_hr) [B,3,H,W]}."""
        lr = self._check(lr, "lr")
        heats = [float(h) for h in heats]
        B, _, h, w = lr.shape
        combos = [(ht, i) for ht in heats for i in range(int(n_sample))]
        eng = self.engine("reverse", B * len(combos), h, w, lr.device)
        with self._guard(), torch.cuda.device(lr.device):
            eng.ext["lr"].copy_(lr.repeat(len(combos), 1, 1, 1))
            for k, (ht, _) in enumerate(combos):                  # the reference's draw order
                for i in range(len(eng.plan.noise_shapes)):
                    eng.ext["eps{}".format(i)][k * B:(k + 1) * B].normal_(0.0, 1.0).mul_(ht)
            eng.run()
            hr = eng.ext["hr"]
            self.last = {"hr_raw": eng.ext["hr_raw"].clone()}
            return {c: hr[k * B:(k + 1) * B].clone() for k, c in enumerate(combos)}

    def sample_uint8(self, lr_u8, eps_std=None, eps=None, bgr=True):
        """Inverse pass on 8-bit images: lr_u8 uint8 [B,h,w,3] as cv2 / the LMDB reader deliver it (HWC, BGR
        unless bgr=False) -> uint8 [B,H,W,3] in the same convention.  Replaces, on the device, the reference's
        read_img / BGR->RGB / HWC->CHW (codes/data/util.py:72-86, GTLQ_dataset.py:109-115) in front of the net and
        tensor2img (codes/utils/util.py:790-816) behind it; the flow itself is the same launch plan."""
        if lr_u8.dtype != torch.uint8 or lr_u8.dim() != 4 or lr_u8.shape[3] != 3:
            raise ValueError("lr_u8 must be uint8 [B,h,w,3], got {} {}".format(lr_u8.dtype, tuple(lr_u8.shape)))
        if not lr_u8.is_cuda:
            raise RuntimeError("hcflow_b200 runs on CUDA tensors only (no CPU fallback)")
        B, h, w, _ = lr_u8.shape
        eng = self.engine("reverse", B, h, w, lr_u8.device, io="u8" if bgr else "u8rgb")
        with self._guard(), torch.cuda.device(lr_u8.device):
            eng.ext["lr_u8"].copy_(lr_u8)
            self._draw_eps(eng, eps_std, eps, lr_u8.device)
            eng.run()
            return eng.ext["hr_u8"].clone()


class HCFlowNet_SR(_HCFlowBase):
    SR = True

    def __init__(self, opt, step=None):
        scale = opt_get(opt, ["scale"])
        if scale not in (4, 8):
            raise NotImplementedError("Scale {} is not implemented".format(scale))
        super().__init__(opt, step)
        self.quant = opt_get(opt, ["quant"], 256)
        self.quantization = nn.Identity()  # parameter-free in the reference too (Basic.py:194-199)

    def forward(self, hr=None, lr=None, z=None, u=None, eps_std=None, add_gt_noise=False, step=None,
                reverse=False, training=True, *, gt=None, eps=None, dequant_noise=None):
        if hr is None and gt is not None:
            hr = gt
        if reverse:
            return self._reverse(lr, eps_std, eps)
        return self._forward_nll(hr, lr, dequant_noise)

    def _forward_nll(self, hr, lr, dequant_noise):
        self._refuse_replica_training()
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or (hr is not None and hr.requires_grad)):
            # training (HCFlow_SR_model.py:184-203 optimize_parameters): the same ops as torch.autograd.Function
            # extensions with CUDA forward + backward kernels (hcflow_b200/autograd.py); the fused engine is inference-only
            from . import autograd as ag
            if hr is None or lr is None:
                raise ValueError("hr and lr are required")
            if not hr.is_cuda:
                raise RuntimeError("hcflow_b200 runs on CUDA tensors only (no CPU fallback)")
            with torch.cuda.device(hr.device):
                return ag.sr_forward_nll(self, hr.to(torch.float32), lr.to(torch.float32), dequant_noise)
        hr = self._check(hr, "hr")
        lr = self._check(lr, "lr")
        B, _, H, W = hr.shape
        s = 2 ** self.flow.L
        if H % s or W % s:
            raise ValueError("HR size {}x{} not divisible by {}".format(H, W, s))
        h, w = H // s, W // s
        assert tuple(lr.shape) == (B, 3, h, w), (tuple(lr.shape), (B, 3, h, w))
        eng = self.engine("forward", B, h, w, hr.device)
        with self._guard(), torch.cuda.device(hr.device):
            eng.ext["hr"].copy_(hr)
            eng.ext["lr"].copy_(lr)
            if dequant_noise is None:
                dequant_noise = torch.rand(hr.shape, device=hr.device)
            eng.ext["dequant"].copy_(dequant_noise.to(device=hr.device, dtype=torch.float32))
            eng.run()
            objective = eng.logdet.clone()  # fp64 [B]: logdet + log N(fake_lr; lr, e^-6)
            nll = ((-objective) / float(math.log(2.0) * H * W)).mean().to(torch.float32)
            self.last = {"z_raw": eng.ext["z_raw"].clone(), "objective": objective}
            return eng.ext["fake_lr"].clone(), nll


class HCFlowNet_Rescaling(_HCFlowBase):
    SR = False

    def __init__(self, opt, step=None):
        super().__init__(opt, step)
        self.quant = opt_get(opt, ["datasets", "train", "quant"], 256)

    def forward(self, hr=None, lr=None, z=None, u=None, eps_std=None, add_gt_noise=False, step=None,
                reverse=False, training=True, *, gt=None, eps=None):
        if hr is None and gt is not None:
            hr = gt
        if reverse:
            return self._reverse(lr, eps_std, eps)
        hr = self._check(hr, "hr")
        B, _, H, W = hr.shape
        s = 2 ** self.flow.L
        if H % s or W % s:
            raise ValueError("HR size {}x{} not divisible by {}".format(H, W, s))
        eng = self.engine("forward", B, H // s, W // s, hr.device)
        with self._guard(), torch.cuda.device(hr.device):
            eng.ext["hr"].copy_(hr)
            eng.run()
            self.last = {"z_raw": eng.ext["z_raw"].clone()}
            return eng.ext["fake_lr"].clone(), eng.ext["fake_z1"].clone(), eng.ext["fake_z2"].clone()


def build_net(opt, step=None):
    """Equivalent of the reference's networks.define_G (codes/models/networks.py:36-41)."""
    which = opt["network_G"]["which_model_G"]
    cls = {"hcflownet_sr": HCFlowNet_SR, "hcflownet_rescaling": HCFlowNet_Rescaling}.get(
        which.replace("_Net", "").lower())
    if cls is None:
        raise NotImplementedError(which)
    return cls(opt=opt, step=step)
