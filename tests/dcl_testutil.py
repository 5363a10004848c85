"""Seeded synthetic inputs shared by the CPU and GPU tests."""
import types

import numpy as np
import torch

GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden")


def uniform_cloud(seed, b, n):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, n, 3, generator=g)


def cad_like_cloud(seed, b, n, n_unique=None):
    """Non-uniform surface-like cloud in metres, resampled WITH replacement => duplicated points => ties
    (the dataloader does this when a crop has too few points, YCBV/dataloader_train_YCBV.py:195-198)."""
    g = torch.Generator().manual_seed(seed)
    n_unique = n_unique or max(8, n // 3)
    u = torch.rand(b, n_unique, 2, generator=g)
    theta, z = u[..., 0] * 6.2831853, (u[..., 1] - 0.5) * 0.2
    base = torch.stack([0.05 * torch.cos(theta), 0.05 * torch.sin(theta), z], -1)
    base = (base * 512).round() / 512  # coarse grid => exact distance ties between distinct points too
    pick = torch.randint(0, n_unique, (b, n), generator=g)
    return torch.gather(base, 1, pick.unsqueeze(-1).expand(b, n, 3)).contiguous()


def flat_bxyz(seed, b, n_per, shuffle=True, scale=0.2):
    g = torch.Generator().manual_seed(seed)
    ids = torch.arange(b).repeat_interleave(n_per).float().unsqueeze(1)
    pts = (torch.rand(b * n_per, 3, generator=g) - 0.5) * scale
    rows = torch.cat([ids, pts], 1)
    if shuffle:
        rows = rows[torch.randperm(rows.shape[0], generator=g)]
    return rows.contiguous()


from dcl_net_b200.synthetic import backbone_levels as synthetic_backbone_levels, levels_to  # noqa: E402,F401


def rel_err(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def np_t(x):
    return x.detach().cpu().numpy()
