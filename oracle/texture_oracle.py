"""CPU restatement of the stage-2 colour field -- TEST INFRASTRUCTURE ONLY.

What it follows
  * /root/reference/geo/texture.py:47-111 -- `MLPTexture3D`: AABB normalisation + clamp (:102-103), the encoder
    configuration (:57-72: HashGrid, 16 levels, 2 features per level, log2_hashmap_size 19, base resolution 16,
    per_level_scale = exp(ln(4096 / 16) / 15)), the bias-free fp32 MLP `_MLP` (:18-41: Linear(32,32) ReLU
    Linear(32,32) ReLU Linear(32,C)), `sample_noact` (:99-111), and the two backward hooks (:71, :30-31) whose net
    effect is: gradients of the ENCODER PARAMETERS are multiplied by 128, every other gradient is the true one.
  * the hash-grid arithmetic itself lives in the third-party `tinycudann` (README.md:15: installed from the
    un-pinned git master of NVlabs/tiny-cuda-nn with --no-networks; absent from /root/reference and from this
    image).  Its published algorithm is restated here (include/tiny-cuda-nn/encodings/grid.h: `grid_scale`,
    `grid_resolution`, the level offset table of `GridEncodingTemplated`, `grid_index` with the coherent prime
    hash {1, 2654435761, 805459861}, trilinear `kernel_grid`; parameters and encoder output in fp16).

PARITY STATUS: **unpinned** -- there is no tinycudann to execute and the reference holds no test vector for the
texture.  The MLP half is pinned by construction (plain torch.nn.functional.linear, the reference's own ops).

Everything is differentiable torch (fp32 like the reference, or fp64 for gradient truth); the fp16 roundings of
tinycudann (parameters, encoder output) are applied with a straight-through estimator.
"""
from __future__ import annotations

import math

import numpy as np
import torch

N_LEVELS, N_FEATURES, LOG2_HASHMAP, BASE_RES = 16, 2, 19, 16
PER_LEVEL_SCALE = float(np.exp(np.log(4096 / 16) / (N_LEVELS - 1)))  # geo/texture.py:54-55
GRAD_SCALE = 128.0  # geo/texture.py:69
PRIMES = (1, 2654435761, 805459861)


def level_table():
    """-> (scales f32[16], resolutions int[16], offsets int[17]) exactly as tinycudann derives them (fp32 math)."""
    log2_pls = np.float32(np.log2(np.float32(PER_LEVEL_SCALE)))
    scales, res, offs = [], [], [0]
    for l in range(N_LEVELS):
        s = np.float32(np.exp2(np.float32(l) * log2_pls)) * np.float32(BASE_RES) - np.float32(1.0)
        r = int(np.ceil(s)) + 1
        n = r ** 3
        n = (n + 7) // 8 * 8
        n = min(n, 1 << LOG2_HASHMAP)
        scales.append(np.float32(s))
        res.append(r)
        offs.append(offs[-1] + n)
    return np.array(scales, np.float32), np.array(res, np.int64), np.array(offs, np.int64)


def n_grid_params() -> int:
    return int(level_table()[2][-1]) * N_FEATURES


def init_params(seed: int = 0, channels: int = 48):
    """tinycudann initialises the grid uniformly in [-1e-4, 1e-4]; _MLP uses kaiming_uniform_(relu) (texture.py:37-41)."""
    g = torch.Generator().manual_seed(seed)
    grid = (torch.rand(n_grid_params(), generator=g) * 2 - 1) * 1e-4
    ws = []
    for fan_out, fan_in in ((32, 32), (32, 32), (channels, 32)):
        bound = math.sqrt(2.0) * math.sqrt(3.0 / fan_in)
        ws.append((torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound)
    return grid, ws


def _ste_half(x):
    return x + (x.detach().half().to(x.dtype) - x.detach())


def _ste_f32(x):
    return x + (x.detach().float().to(x.dtype) - x.detach())


def encode(t, grid, half=True, f32_coords=False):
    """t [N,3] in [0,1] -> [N,32].  grid: flat [n_params] (level-major, then cell, then feature).
    f32_coords (t in float64): the cell position is rounded to fp32 as `fmaf(scale, t, 0.5)` gives it -- at the finest
    level one fp32 ulp of the position is 2.4e-4 of a cell, which moves the trilinear weights far more than fp32
    arithmetic anywhere else does."""
    scales, res, offs = level_table()
    N = t.shape[0]
    dt = t.dtype
    g = _ste_half(grid) if half else grid
    outs = []
    for l in range(N_LEVELS):
        s = float(scales[l])
        hsize = int(offs[l + 1] - offs[l])
        r = int(res[l])
        pos = t * s + 0.5  # fmaf(scale, x, 0.5)
        if f32_coords:
            pos = _ste_f32(pos)
        pg = torch.floor(pos.detach())
        w = pos - pg
        pgi = pg.to(torch.int64)
        acc = torch.zeros(N, N_FEATURES, dtype=dt)
        for corner in range(8):
            wgt = torch.ones(N, dtype=dt)
            c = []
            for d in range(3):
                if corner & (1 << d):
                    wgt = wgt * w[:, d]
                    c.append(pgi[:, d] + 1)
                else:
                    wgt = wgt * (1 - w[:, d])
                    c.append(pgi[:, d])
            # grid_index: dense while the strides fit the level's table, hashed otherwise
            stride, idx, dense = 1, torch.zeros(N, dtype=torch.int64), True
            for d in range(3):
                if stride > hsize:
                    break
                idx = idx + c[d] * stride
                stride *= r
            if hsize < stride:
                h = torch.zeros(N, dtype=torch.int64)
                for d in range(3):
                    h = h ^ ((c[d] * PRIMES[d]) & 0xFFFFFFFF)
                idx = h
            idx = (idx & 0xFFFFFFFF) % hsize
            base = (int(offs[l]) + idx) * N_FEATURES
            vals = torch.stack([g[base + f] for f in range(N_FEATURES)], 1).to(dt)
            acc = acc + wgt[:, None] * vals
        outs.append(acc)
    enc = torch.cat(outs, 1)
    return _ste_half(enc) if half else enc


def sample_noact(xyz, aabb, grid, ws, half=True, f32_coords=False):
    """geo/texture.py:99-111.  xyz [N,3]; aabb [2,3]; -> [N, C].  Gradients: true ones (multiply grid.grad by
    GRAD_SCALE to get what the reference's hooks hand to its optimiser).  f32_coords: see `encode` (the normalised
    coordinate is rounded to fp32 too, as the reference's fp32 tensors are)."""
    if f32_coords:  # every fp32 operation of (x - lo) / (hi - lo) rounds on its own
        t = _ste_f32(_ste_f32(xyz - aabb[0][None]) / _ste_f32(aabb[1][None] - aabb[0][None]))
    else:
        t = (xyz - aabb[0][None]) / (aabb[1][None] - aabb[0][None])
    t = torch.clamp(t, min=0, max=1)
    x = encode(t, grid, half, f32_coords)
    h = torch.relu(torch.nn.functional.linear(x, ws[0]))
    h = torch.relu(torch.nn.functional.linear(h, ws[1]))
    return torch.nn.functional.linear(h, ws[2])
