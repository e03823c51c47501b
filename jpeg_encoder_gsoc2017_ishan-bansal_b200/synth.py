"""Seeded synthetic frames for tests and benchmarks.

photo : bilinear upsample of a (H/16+1) x (W/16+1) x 3 uniform u8 grid plus uniform
        per-pixel noise in [-12, 12], clamped to u8 (smooth regions, soft edges);
noise : i.i.d. uniform u8 (worst case for the scan / bit packer);
grey  : constant 128 (EOB-only blocks).
Generated on the CPU with torch.Generator().manual_seed(20171019 + 1000*config_idx + frame_idx).
"""
import torch
import torch.nn.functional as F


def frame(height, width, config_idx=0, frame_idx=0, kind="photo"):
    g = torch.Generator().manual_seed(20171019 + 1000 * int(config_idx) + int(frame_idx))
    if kind == "grey":
        return torch.full((height, width, 3), 128, dtype=torch.uint8)
    if kind == "noise":
        return torch.randint(0, 256, (height, width, 3), generator=g, dtype=torch.uint8)
    if kind != "photo":
        raise ValueError(kind)
    gh, gw = height // 16 + 1, width // 16 + 1
    grid = torch.randint(0, 256, (1, 3, gh, gw), generator=g, dtype=torch.int32).float()
    up = F.interpolate(grid, size=(height, width), mode="bilinear", align_corners=True)[0]
    noise = torch.randint(-12, 13, (3, height, width), generator=g, dtype=torch.int32).float()
    return (up + noise).round().clamp_(0, 255).to(torch.uint8).permute(1, 2, 0).contiguous()


def batch(n, height, width, config_idx=0, kind="photo", unique=None, out=None):
    """(n, H, W, 3) uint8.  `unique` < n generates that many distinct frames and tiles them."""
    unique = n if unique is None else min(unique, n)
    if out is None:
        out = torch.empty((n, height, width, 3), dtype=torch.uint8)
    for i in range(unique):
        out[i] = frame(height, width, config_idx, i, kind)
    for i in range(unique, n):
        out[i] = out[i % unique]
    return out
