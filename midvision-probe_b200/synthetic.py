"""Seeded synthetic inputs of the shapes named in BASELINE.json (no datasets or checkpoints exist offline).

All tensors are generated on the CPU with torch.Generator().manual_seed(1000 * config + index) so that the
CUDA path, the oracle and the committed golden vectors see bit-identical inputs on every machine.

    C1 spair_pair    feats (2, 768, 14, 14), 20 keypoints            DINO ViT-B/16 @ 224
    C2 navi_pair     feats (3072, 28, 28) x2, xyz grid (3, 112, 112) DINO ViT-B/16 @ 448, 4-block concat
    C3 scannet_pair  feats (2048, 15, 20) x2, depth (1, 120, 160)    ResNet-50 layer4 @ 480x640
    C5 stress_rows   A, B (19200, 768) L2-normalised rows

`coherent=True` makes image 1 a noisy copy of image 0 over a smooth surface, so that recall is neither 0 nor 100,
near-miss matches land between the thresholds and a wrong match changes the counts.
"""
import math

import torch


def _gen(config, index):
    g = torch.Generator()
    g.manual_seed(1000 * config + index)
    return g


def random_rt(g, max_deg=120.0, t_sigma=0.1):
    """(3, 4) rigid transform: rotation about a random axis by <= max_deg, translation ~ N(0, t_sigma)."""
    axis = torch.randn(3, generator=g)
    axis = axis / axis.norm()
    ang = (torch.rand(1, generator=g).item() * max_deg) * math.pi / 180.0
    Kx = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = torch.eye(3) + math.sin(ang) * Kx + (1 - math.cos(ang)) * (Kx @ Kx)
    t = torch.randn(3, generator=g) * t_sigma
    return torch.cat((R, t[:, None]), dim=1).float()


def _feature_pair(g, C, h, w, coherent, noise=3.0):
    f0 = torch.randn(C, h, w, generator=g)
    if coherent:
        f1 = f0 + noise * torch.randn(C, h, w, generator=g)
    else:
        f1 = torch.randn(C, h, w, generator=g)
    return f0.contiguous(), f1.contiguous()


def scannet_pair(index, C=2048, h=15, w=20, H=120, W=160, zero_frac=0.05, coherent=True, config=3):
    """ScanNet-shaped pair (render_scannet_correspondence.py:188-208 after the 0.25 rescale)."""
    g = _gen(config, index)
    f0, f1 = _feature_pair(g, C, h, w, coherent)
    ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    # smooth surface + a little roughness: a one-pixel mismatch is centimetres, not metres
    d0 = (2.0 + 0.8 * torch.sin(xs / 17.0) * torch.cos(ys / 13.0) + 0.01 * torch.rand(H, W, generator=g))[None]
    if coherent:
        d1 = d0.clone()
        Rt = torch.cat((torch.eye(3), torch.zeros(3, 1)), dim=1)
    else:
        d1 = (2.2 + 0.7 * torch.cos(xs / 15.0) * torch.sin(ys / 11.0) + 0.01 * torch.rand(H, W, generator=g))[None]
        Rt = random_rt(g)
    d0[torch.rand(1, H, W, generator=g) < zero_frac] = 0.0
    d1[torch.rand(1, H, W, generator=g) < zero_frac] = 0.0
    K = torch.tensor([[1165.7, 0.0, 649.1], [0.0, 1165.7, 484.8], [0.0, 0.0, 1.0]])
    K[:2] *= 0.25  # K_mat[:2, :] *= cfg.scale_factor (render_scannet_correspondence.py:203)
    return {"feat_0": f0, "feat_1": f1, "depth_0": d0, "depth_1": d1, "K": K, "Rt": Rt.float()}


def navi_pair(index, C=3072, h=28, w=28, H=112, W=112, radius=40.0, coherent=True, config=2):
    """NAVI-shaped pair (evaluate_navi_correspondence.py:143-181): xyz grid valid inside a disc."""
    g = _gen(config, index)
    f0, f1 = _feature_pair(g, C, h, w, coherent, noise=0.2 * math.sqrt(C))  # ~11 at C = 3072: recall@1cm ~ 80 %
    ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")

    def grid(cx, cy):
        # a smooth object surface seen through a pinhole of focal length 1.2 * W (in grid pixels): neighbouring
        # pixels are ~5 mm apart, so the 1 / 2 / 5 cm thresholds separate exact, near and far matches
        inside = ((xs - cx) ** 2 + (ys - cy) ** 2) < radius * radius
        z = 0.6 + 0.15 * torch.sin(xs / 9.0) * torch.cos(ys / 7.0) + 0.002 * torch.rand(H, W, generator=g)
        f = 1.2 * W
        xyz = torch.stack(((xs + 0.5 - W / 2) * z / f, (ys + 0.5 - H / 2) * z / f, z), dim=0)
        xyz[:, ~inside] = 0.0
        return xyz.contiguous()

    x0 = grid(W / 2 - 0.5, H / 2 - 0.5)
    if coherent:
        x1 = x0.clone()
        Rt = torch.cat((torch.eye(3), torch.zeros(3, 1)), dim=1)
    else:
        x1 = grid(W / 2 + 3.5, H / 2 - 2.5)
        Rt = random_rt(g)
    fx = 4.0 * 1.2 * W * (0.9 + 0.2 * torch.rand(1, generator=g).item())  # full-resolution intrinsics (grid = 1/4 scale)
    intr = torch.tensor([[fx, 0.0, 2.0 * W], [0.0, fx, 2.0 * H], [0.0, 0.0, 1.0]])
    return {"feat_0": f0, "feat_1": f1, "xyz_grid_0": x0, "xyz_grid_1": x1, "Rt": Rt.float(), "intrinsics": intr}


def spair_pair(index, C=768, h=14, w=14, K=20, image_size=224, valid_p=0.8, coherent=True, config=1):
    """SPair-shaped pair (evaluate_spair_correspondence.py:45-79): features of both images + keypoints."""
    g = _gen(config, index)
    f0, f1 = _feature_pair(g, C, h, w, coherent, noise=1.0)
    kps_i = torch.zeros(K, 3)
    kps_j = torch.zeros(K, 3)
    kps_i[:, :2] = torch.randint(0, image_size, (K, 2), generator=g).float()
    kps_j[:, :2] = kps_i[:, :2] if coherent else torch.randint(0, image_size, (K, 2), generator=g).float()
    kps_i[:, 2] = (torch.rand(K, generator=g) < valid_p).float()
    kps_j[:, 2] = (torch.rand(K, generator=g) < valid_p).float()
    thresh_scale = 0.3 + 0.7 * torch.rand(1, generator=g).item()
    return {"feats": torch.stack((f0, f1)).contiguous(), "kps_i": kps_i, "kps_j": kps_j,
            "thresh_scale": thresh_scale, "image_size": image_size}


def stress_rows(index, n=19200, m=19200, C=768, variant="iid", config=5):
    """Rows for the kernel-2 stress case: i.i.d. normal, or bilinear 8x upsamples of a 15 x 20 map
    (near-collinear neighbours, the regime of the real evaluations); L2-normalised fp32."""
    g = _gen(config, index)
    if variant == "iid":
        A = torch.randn(n, C, generator=g)
        B = torch.randn(m, C, generator=g)
    elif variant == "upsampled":
        def up(rows):
            side_h, side_w = 15, 20
            base = torch.randn(1, C, side_h, side_w, generator=g)
            Hh = max(1, int(round(math.sqrt(rows * side_h / side_w))))
            Ww = (rows + Hh - 1) // Hh
            big = torch.nn.functional.interpolate(base, size=(Hh, Ww), mode="bilinear", align_corners=False)
            return big[0].reshape(C, -1).t()[:rows].contiguous()
        A, B = up(n), up(m)
    else:
        raise ValueError(variant)
    A = torch.nn.functional.normalize(A, dim=1)
    B = torch.nn.functional.normalize(B, dim=1)
    return A.contiguous(), B.contiguous()
