"""Stub: the reference's test_HCFlow.py builds lpips.LPIPS(net='alex') (test_HCFlow.py:14,48) for a perceptual score that
is outside the hot path (and needs downloaded AlexNet weights); this stand-in returns 0 so the unmodified script runs."""
import torch


class LPIPS(torch.nn.Module):
    def __init__(self, net="alex", **kwargs):
        super().__init__()

    def forward(self, a, b):
        return torch.zeros(1, device=a.device)
