"""Evaluation metrics on the device: PSNR / SSIM and their Y-channel variants exactly as the reference's
``util.calculate_psnr_ssim`` computes them on the host with numpy / cv2 (codes/utils/util.py:902-982,
codes/data/util.py:209-230), so that test_HCFlow.py's metric loop (codes/test_HCFlow.py:100-155) needs no device->host
copy of the images -- only the four numbers come back."""
import ctypes as C
import math

import torch

from . import _lib as L


def _gauss_window():
    """cv2.getGaussianKernel(11, 1.5) outer product (codes/utils/util.py:926-927), fp64."""
    x = torch.arange(11, dtype=torch.float64) - 5.0
    k = torch.exp(-(x * x) / (2.0 * 1.5 * 1.5))
    k = k / k.sum()
    return torch.outer(k, k).contiguous()


_WIN = {}


def psnr_ssim(img1, img2, crop_border=0):
    """img1, img2: [H, W, C] CUDA tensors, uint8 or float in [0, 1], channel order BGR (what util.tensor2img(...)/255
    gives the reference).  Returns (psnr, ssim, psnr_y, ssim_y) as python floats; psnr_y = ssim_y = 0 unless C == 3."""
    if not (img1.is_cuda and img2.is_cuda):
        raise RuntimeError("hcflow_b200 metrics run on CUDA tensors only (no CPU fallback)")
    if img1.shape != img2.shape or img1.dim() != 3:
        raise ValueError("Input images must have the same [H, W, C] dimensions.")
    if img1.dtype != img2.dtype or img1.dtype not in (torch.uint8, torch.float32):
        raise ValueError("images must both be uint8 or both be float32")
    lib = L.load()
    H, W, Cc = img1.shape
    a, b = img1.contiguous(), img2.contiguous()
    dev = a.device
    if dev not in _WIN:
        _WIN[dev] = _gauss_window().to(dev)
    out = torch.zeros(2 * Cc + 2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.hcf_image_metrics(a.data_ptr(), b.data_ptr(), int(a.dtype == torch.float32), H, W, Cc, int(crop_border),
                                      _WIN[dev].data_ptr(), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                "image_metrics")
    v = out.cpu().tolist()
    h, w = H - 2 * crop_border, W - 2 * crop_border

    def psnr_of(sq, n):
        mse = sq / n
        return float("inf") if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))
    psnr = psnr_of(sum(v[:Cc]), h * w * Cc)
    ssim = sum(v[Cc:2 * Cc]) / ((h - 10) * (w - 10)) / Cc
    if Cc == 3:
        return psnr, ssim, psnr_of(v[2 * Cc], h * w), v[2 * Cc + 1] / ((h - 10) * (w - 10))
    return psnr, ssim, 0, 0
