"""Tiled (patch-wise) inference for images larger than one compiled engine.

Same tiling and seam handling as the reference's ``test_patchwise`` (codes/data/util.py:489-514): patches of
``patchsize`` with ``overlapsize`` overlap, the last row / column aligned to the image border, overlapping outputs
averaged.  Differences are in the scheduling only: all patches of an image go through ONE compiled engine (fixed
[tile_batch, 3, patchsize, patchsize]) in batches instead of one module call per patch, and the accumulate / normalise
arithmetic runs in CUDA kernels (hcf_tile_accumulate / hcf_tile_normalize)."""
import torch

from . import _lib as L


def patch_origins(size, patchsize, overlapsize):
    """codes/data/util.py:503-505: list(range(0, size - patchsize, stride)) + [size - patchsize]"""
    stride = patchsize - overlapsize
    return list(range(0, size - patchsize, stride)) + [size - patchsize]


def sample_patchwise(net, lr, patchsize=48, overlapsize=16, eps_std=0.0, tile_batch=16, eps_fn=None):
    """Inverse (sampling) pass of ``net`` on an LR image of any size >= patchsize.  lr: [b, 3, h, w] CUDA.
    Returns the HR image [b, 3, h * sf, w * sf] (clamped, like the module call) with overlaps averaged.

    Noise: with eps_std > 0 every patch gets its own draws, made in the order the reference's loop would make them
    (patch by patch, deepest level first) from the current CUDA generator; ``eps_fn(patch_index)`` may supply the
    unit-normal tensors instead."""
    if not lr.is_cuda:
        raise RuntimeError("hcflow_b200 runs on CUDA tensors only (no CPU fallback)")
    lib = L.load()
    b, c, h, w = lr.shape
    sf = 2 ** net.flow.L
    if h < patchsize or w < patchsize:
        raise ValueError("image {}x{} smaller than the patch size {}".format(h, w, patchsize))
    ys, xs = patch_origins(h, patchsize, overlapsize), patch_origins(w, patchsize, overlapsize)
    coords = [(bi, y, x) for y in ys for x in xs for bi in range(b)]   # the reference calls the model on all b images of a window
    dev = lr.device
    E = torch.zeros(b, c, h * sf, w * sf, dtype=torch.float32, device=dev)
    cnt = torch.zeros(b, h * sf, w * sf, dtype=torch.float32, device=dev)
    eng = net.engine("reverse", tile_batch, patchsize, patchsize, dev)
    shapes = eng.plan.noise_shapes
    st = torch.cuda.current_stream(dev).cuda_stream
    with torch.no_grad(), torch.cuda.device(dev):
        for k0 in range(0, len(coords), tile_batch):
            chunk = coords[k0:k0 + tile_batch]
            n = len(chunk)
            patches = torch.stack([lr[bi, :, y:y + patchsize, x:x + patchsize] for bi, y, x in chunk])
            if n < tile_batch:    # pad the last batch (its extra outputs are dropped)
                patches = torch.cat([patches, patches[-1:].expand(tile_batch - n, -1, -1, -1)])
            eps = None
            if eps_std and eps_std > 0:
                eps = [torch.zeros(tile_batch, cc, hh, ww, dtype=torch.float32, device=dev) for cc, hh, ww in shapes]
                for i in range(n):
                    if eps_fn is not None:
                        for lvl, e in enumerate(eps_fn(k0 + i)):
                            eps[lvl][i].copy_(e)
                    else:
                        for lvl in range(len(shapes)):
                            eps[lvl][i].normal_(0.0, 1.0)
            out = net(lr=patches.contiguous(), eps_std=eps_std, reverse=True, eps=eps)
            for bi in range(b):
                sel = [i for i, (bj, _, _) in enumerate(chunk) if bj == bi]
                if not sel:
                    continue
                idx = torch.tensor(sel, device=dev)
                sub = out.index_select(0, idx).contiguous()
                y0 = torch.tensor([chunk[i][1] * sf for i in sel], dtype=torch.int32, device=dev)
                x0 = torch.tensor([chunk[i][2] * sf for i in sel], dtype=torch.int32, device=dev)
                L.check(lib.hcf_tile_accumulate(sub.data_ptr(), len(sel), c, patchsize * sf, patchsize * sf, y0.data_ptr(),
                                                x0.data_ptr(), E[bi].data_ptr(), cnt[bi].data_ptr(), h * sf, w * sf, st),
                        "tile_accumulate")
        for bi in range(b):
            L.check(lib.hcf_tile_normalize(E[bi].data_ptr(), cnt[bi].data_ptr(), c, h * sf, w * sf, st), "tile_normalize")
    return E
