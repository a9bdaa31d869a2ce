"""`MLPTexture3D` -- the stage-2 colour field with the reference's interface (geo/texture.py:47-118), backed by
libdmgs_raster.so (dmgs_texture_forward / dmgs_texture_backward; csrc/texture.cu).

The reference builds it from `tinycudann.Encoding(3, HashGrid)` + a torch `_MLP` and samples it once per step at
every Gaussian centre (scene/gaussian_geo_model_mlp_flex.py:313).  Here the encoder, the MLP and their backward are
two kernels; module structure and parameter names are the reference's, so its checkpoints load unchanged:
`encoder.params` (fp32 master copy of the fp16 hash grid), `net.net.{0,2,4}.weight`.

Gradient scaling: the reference registers two backward hooks (a 1/128 on the encoder input, a x128 on the MLP
input, geo/texture.py:30-31, :69-71); their net effect is that `encoder.params.grad` arrives multiplied by 128 and
every other gradient is the true one.  `grid_grad_scale` (default 128) reproduces that.

There is no CPU path."""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib as L

N_LEVELS, N_FEATURES, ENC_DIMS = 16, 2, 32


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _HashGrid(nn.Module):
    """Parameter holder with tinycudann's name (`params`) and initialisation (uniform in [-1e-4, 1e-4])."""

    def __init__(self, n_params: int, device=None, generator=None):
        super().__init__()
        self.n_input_dims, self.n_output_dims = 3, ENC_DIMS
        p = (torch.rand(n_params, generator=generator, dtype=torch.float32) * 2 - 1) * 1e-4
        self.params = nn.Parameter(p.to(device) if device is not None else p)


class _MLP(nn.Module):
    """geo/texture.py:18-41: bias-free Linear / ReLU stack, kaiming-uniform(relu) weights.  The modules only hold the
    weights (state-dict names as the reference); the arithmetic runs in the fused kernels."""

    def __init__(self, cfg, device=None):
        super().__init__()
        net = (nn.Linear(cfg["n_input_dims"], cfg["n_neurons"], bias=False), nn.ReLU())
        for _ in range(cfg["n_hidden_layers"] - 1):
            net = net + (nn.Linear(cfg["n_neurons"], cfg["n_neurons"], bias=False), nn.ReLU())
        net = net + (nn.Linear(cfg["n_neurons"], cfg["n_output_dims"], bias=False),)
        self.net = nn.Sequential(*net)
        for m in self.net:
            if isinstance(m, nn.Linear):
                nn.init.kaiming_uniform_(m.weight, nonlinearity="relu")
        if device is not None:
            self.net.to(device)

    def weights(self):
        return [m.weight for m in self.net if isinstance(m, nn.Linear)]


class _SampleNoAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, grid, W0, W1, W2, aabb6, grid_half, grid_grad_scale):
        if xyz.device.type != "cuda":
            raise RuntimeError("dmgs_b200 texture needs CUDA tensors; there is no CPU path")
        dev = xyz.device
        x = xyz.detach().float().contiguous()
        N, Cc = int(x.shape[0]), int(W2.shape[0])
        out = torch.empty(N, Cc, dtype=torch.float32, device=dev)
        enc = torch.empty(N, ENC_DIMS, dtype=torch.float16, device=dev)
        w = [t.detach().float().contiguous() for t in (W0, W1, W2)]
        with torch.cuda.device(dev):
            L.check(L.lib().dmgs_texture_forward(N, Cc, aabb6, L.ptr(x), L.ptr(grid_half), L.ptr(w[0]), L.ptr(w[1]), L.ptr(w[2]),
                                                 L.ptr(out), L.ptr(enc), _stream(dev)), "dmgs_texture_forward")
        ctx.save_for_backward(x, enc, *w)
        ctx.meta = (aabb6, grid_half, float(grid_grad_scale), int(grid.numel()))
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, enc, W0, W1, W2 = ctx.saved_tensors
        aabb6, grid_half, gscale, n_params = ctx.meta
        dev = x.device
        N, Cc = int(x.shape[0]), int(W2.shape[0])
        need_xyz, need_grid = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g = g_out.detach().float().contiguous()
        d_grid = torch.zeros(n_params, dtype=torch.float32, device=dev) if need_grid else None
        dW = [torch.zeros_like(t) for t in (W0, W1, W2)]
        d_xyz = torch.empty(N, 3, dtype=torch.float32, device=dev) if need_xyz else None
        scratch = torch.empty(int(L.lib().dmgs_texture_backward_scratch_bytes(N)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L.check(L.lib().dmgs_texture_backward(N, Cc, aabb6, L.ptr(x), L.ptr(grid_half), L.ptr(enc), L.ptr(W0), L.ptr(W1),
                                                  L.ptr(W2), L.ptr(g), gscale, L.ptr(d_grid), L.ptr(dW[0]), L.ptr(dW[1]),
                                                  L.ptr(dW[2]), L.ptr(d_xyz), L.ptr(scratch), _stream(dev)), "dmgs_texture_backward")
        return d_xyz, d_grid, dW[0], dW[1], dW[2], None, None, None


class MLPTexture3D(nn.Module):
    def __init__(self, AABB, channels=3, internal_dims=32, hidden=2, min_max=None, grid_grad_scale=128.0, device="cuda",
                 generator=None):
        super().__init__()
        if internal_dims != 32 or hidden != 2:
            raise NotImplementedError("the fused kernels implement the reference's fixed configuration (32 neurons, 2 hidden layers)")
        if channels % 4 or not 4 <= channels <= 64:
            raise NotImplementedError("channels must be a multiple of 4 in [4, 64] (DMGS uses 3 * 16 = 48)")
        self.channels, self.internal_dims, self.AABB, self.min_max = channels, internal_dims, AABB, min_max
        self.grid_grad_scale = float(grid_grad_scale)
        n_params = int(L.lib().dmgs_texture_grid_params())
        self.encoder = _HashGrid(n_params, device=device, generator=generator)
        self.net = _MLP({"n_input_dims": ENC_DIMS, "n_output_dims": channels, "n_hidden_layers": hidden,
                         "n_neurons": internal_dims}, device=device)
        self._half = None       # fp16 copy of the grid the kernels read

    def _aabb6(self):
        lo, hi = self.AABB[0], self.AABB[1]
        vals = [float(v) for v in torch.as_tensor(lo).reshape(-1).tolist()] + [float(v) for v in torch.as_tensor(hi).reshape(-1).tolist()]
        return (C.c_float * 6)(*vals)

    def _grid_half(self):
        """fp16 copy of `encoder.params` for the kernels, refreshed on every call (tinycudann also casts its
        parameters every step; 73 MB of traffic, ~15 us -- and raw-pointer optimisers such as FusedAdam do not
        bump the tensor version a cache could key on)."""
        p = self.encoder.params
        if p.device.type != "cuda":
            raise RuntimeError("dmgs_b200 texture needs CUDA tensors; there is no CPU path")
        if self._half is None or self._half.device != p.device:
            self._half = torch.empty(p.numel(), dtype=torch.float16, device=p.device)
        with torch.cuda.device(p.device):
            L.check(L.lib().dmgs_texture_cast_params(p.numel(), L.ptr(p.detach()), L.ptr(self._half), _stream(p.device)),
                    "dmgs_texture_cast_params")
        return self._half

    def sample_noact(self, texc):
        """geo/texture.py:99-111: raw MLP output at the given positions, [..., 3] -> [..., channels]."""
        if texc.numel() == 0:
            return torch.zeros(*texc.shape[:-1], self.channels, device=texc.device)
        W0, W1, W2 = self.net.weights()
        out = _SampleNoAct.apply(texc.reshape(-1, 3), self.encoder.params, W0, W1, W2, self._aabb6(), self._grid_half(),
                                 self.grid_grad_scale)
        return out.view(*texc.shape[:-1], self.channels)

    def sample(self, texc):
        """geo/texture.py:84-96: sigmoid-limited output scaled to [min_max[0], min_max[1]]."""
        out = self.sample_noact(texc).view(-1, self.channels)
        out = torch.sigmoid(out) * (self.min_max[1][None, :] - self.min_max[0][None, :]) + self.min_max[0][None, :]
        return out.view(*texc.shape[:-1], self.channels)

    def clamp_(self):
        pass

    def cleanup(self):
        pass
