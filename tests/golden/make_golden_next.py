"""Generates tests/golden/{loss_l1_ssim,frustum,adam}.npz by EXECUTING THE REFERENCE'S OWN PYTHON
(or, for Adam, the library call the reference makes) on the CPU.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden_next.py
  * utils/loss_utils.py imports cleanly (pure torch): l1_loss, ssim and autograd gradients;
  * the two `in_frustum` functions live in model files that cannot be imported (open3d, pytorch3d,
    tinycudann ... missing), so their source lines are sliced out and exec()ed unchanged (TorchScript
    decorator dropped; the literal device='cuda' tensors of the COLMAP variant are created on the CPU);
  * Adam: torch.optim.Adam(l, lr=0.0, eps=1e-15) with per-group learning rates, as constructed at
    scene/gaussian_geo_model_finetune.py:526-537.
Nothing is copied into the repo except the resulting numbers.
"""
import os
import sys
import textwrap

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(OUT, "..", ".."))
sys.path.insert(0, REF)


def ref_lines(rel, lo, hi):
    with open(os.path.join(REF, rel)) as fh:
        lines = fh.readlines()
    return textwrap.dedent("".join(lines[lo - 1:hi]))


def loss_golden():
    from utils import loss_utils as RL  # the reference module
    gen = torch.Generator().manual_seed(11)
    out = {}
    for tag, (H, W) in {"a": (37, 53), "b": (64, 96)}.items():
        gt = torch.rand(3, H, W, generator=gen)
        img = (gt + 0.25 * torch.randn(3, H, W, generator=gen)).clamp(0, 1)
        img[:, :4, :5] = gt[:, :4, :5]  # exact ties: sign(0) = 0 in the L1 gradient
        img.requires_grad_()
        l1 = RL.l1_loss(img, gt)
        ss = RL.ssim(img, gt)
        lam = 0.2
        loss = (1.0 - lam) * l1 + lam * (1.0 - ss)  # train_geo_stage2.py:116
        (g,) = torch.autograd.grad(loss, img, retain_graph=True)
        (g_l1,) = torch.autograd.grad(l1, img, retain_graph=True)
        (g_ss,) = torch.autograd.grad(ss, img)
        out.update({f"img_{tag}": img.detach().numpy(), f"gt_{tag}": gt.numpy(), f"l1_{tag}": l1.item(),
                    f"ssim_{tag}": ss.item(), f"loss_{tag}": loss.item(), f"grad_{tag}": g.numpy(),
                    f"grad_l1_{tag}": g_l1.numpy(), f"grad_ssim_{tag}": g_ss.numpy()})
    # batched input, size_average=False (loss_utils.py:60-63)
    a, b = torch.rand(2, 3, 40, 48, generator=gen), torch.rand(2, 3, 40, 48, generator=gen)
    out["img_batch"], out["gt_batch"] = a.numpy(), b.numpy()
    out["ssim_batch"] = RL.ssim(a, b, size_average=False).numpy()
    out["window"] = RL.gaussian(11, 1.5).numpy()
    np.savez_compressed(os.path.join(OUT, "loss_l1_ssim.npz"), **out)
    print("loss:", {k: v for k, v in out.items() if np.ndim(v) == 0})


def frustum_golden():
    from dmgs_b200 import synthetic as S
    env = {"torch": torch}
    src = ref_lines("scene/gaussian_geo_model_finetune.py", 35, 48)  # def in_frustum ... return mask
    exec(src, env)
    in_frustum_ft = env["in_frustum"]
    src = ref_lines("scene/gaussian_geo_model_mlp_flex_colmap.py", 34, 76).replace("device='cuda'", "device='cpu'")
    env2 = {"torch": torch}
    exec(src, env2)
    in_frustum_cm = env2["in_frustum"]
    cam = S.look_at_camera([1.2, 0.4, 0.9], 320, 200, fovx=0.6)  # close to the mesh: part of it is off-screen
    verts, faces = S.jittered_sphere_mesh(6000, seed=4, jitter=0.05)
    proj = cam.full_proj_transform
    out = {"proj": proj.numpy(), "verts": verts.numpy(), "faces": faces.numpy()}
    query = verts[faces].mean(dim=1)  # finetune.py:405
    m = in_frustum_ft(proj, query)
    out["face_mask"] = m.numpy()
    out["faces_visible"] = faces[m].numpy()  # finetune.py:408
    out["vert_mask"] = in_frustum_ft(proj, verts).numpy()
    gen = torch.Generator().manual_seed(5)
    grid = torch.rand(20000, 3, generator=gen) * 3.0 - 1.5
    out["grid"] = grid.numpy()
    cube_len = 0.37
    out["cube_len"] = cube_len
    for pid, npc in [(-1, 1), (0, 2), (1, 2), (0, 4), (1, 4), (2, 4), (3, 4)]:
        out[f"grid_mask_{pid}_{npc}"] = in_frustum_cm(proj, grid, cube_len, pid, npc).numpy()
    np.savez_compressed(os.path.join(OUT, "frustum.npz"), **out)
    print("frustum: visible faces", int(m.sum()), "of", faces.shape[0], "; grid", {k: int(v.sum()) for k, v in out.items() if k.startswith("grid_mask")})


def adam_golden():
    gen = torch.Generator().manual_seed(3)
    P = 257
    shapes = {"scaling": (P, 2), "rotation": (P, 2), "opacity": (P, 1), "f_dc": (P, 1, 3), "f_rest": (P, 15, 3)}
    lrs = {"scaling": 0.005, "rotation": 0.001, "opacity": 0.05, "f_dc": 0.0025, "f_rest": 0.0025 / 20.0}
    params = {k: torch.randn(*s, generator=gen).requires_grad_() for k, s in shapes.items()}
    l = [{"params": [params[k]], "lr": lrs[k], "name": k} for k in shapes]
    opt = torch.optim.Adam(l, lr=0.0, eps=1e-15)  # finetune.py:537
    out = {f"p0_{k}": v.detach().numpy().copy() for k, v in params.items()}
    out["names"] = np.array(list(shapes))
    out["lrs"] = np.array([lrs[k] for k in shapes])
    steps = 4
    for t in range(steps):
        for k, p in params.items():
            g = torch.randn(*shapes[k], generator=gen) * (10.0 ** (t - 2))
            if t == 2:
                g[::3] = 0.0  # exact zeros
            p.grad = g
            out[f"g{t}_{k}"] = g.numpy().copy()
        opt.step()
        opt.zero_grad(set_to_none=True)
        for k, p in params.items():
            out[f"p{t + 1}_{k}"] = p.detach().numpy().copy()
    for k, p in params.items():
        out[f"m_{k}"] = opt.state[p]["exp_avg"].numpy().copy()
        out[f"v_{k}"] = opt.state[p]["exp_avg_sq"].numpy().copy()
    out["steps"] = steps
    np.savez_compressed(os.path.join(OUT, "adam.npz"), **out)
    print("adam: steps", steps)


if __name__ == "__main__":
    loss_golden()
    frustum_golden()
    adam_golden()
