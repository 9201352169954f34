"""Random-init frozen backbones of the shapes BASELINE.json names -- TEST / BENCH INPUT GENERATORS ONLY.

The matching path (kernels 1-3) takes the backbone's output; the backbone forward itself stays in PyTorch
(north star).  There is no network here for checkpoints, so the benchmarks and parity tests use the same
architectures with seeded random weights, which gives feature maps with the statistics of a real forward
(spatially smooth, low-rank, all-positive for the ResNet) instead of i.i.d. Gaussian maps:

    vit_b16            DINO / iBOT ViT-B/16 (evals/models/ibot_transformers.py:225-357, vit_base :432-442; the DINO
                       hub model is the same code lineage): conv patch embedding, cls token, bicubic-interpolated
                       position embedding, 12 pre-norm blocks (12 heads, qkv bias, GELU MLP x4, LayerNorm eps 1e-6).
                       Initialisation as the reference: trunc_normal(0.02) for Linear / pos_embed / cls_token, zero
                       biases, LayerNorm (1, 0), torch's default for the patch convolution.
    DenseViT           the wrapper of evals/models/dino.py:164-210 + evals/models/utils.py:105-124: the un-normed
                       residual stream after blocks [2, 5, 8, 11] (return_multilayer) or the last one, spatial tokens
                       -> (B, C, h, w).  `channel_last=True` skips tokens_to_output's .contiguous() and returns the
                       permuted view (the layout kernel 1 reads without a transpose).
    resnet50_layer4    torchvision resnet50(weights=None), fc dropped, stem + layer1..4 (evals/models/mocov2.py:46-59,
                       :91-109) -> (B, 2048, H/32, W/32).

`smooth_images` makes seeded low-pass images (natural images are smooth; N(0,1) pixels would turn the patch
embedding into i.i.d. noise again).  Nothing on the product path imports this module.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["vit_b16", "DenseViT", "resnet50_layer4", "smooth_images", "navi_backbone_pair", "scannet_backbone_pair",
           "spair_backbone_pair"]


class _Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio=4.0):
        super().__init__()
        self.heads = heads
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.fc1 = nn.Linear(dim, int(dim * mlp_ratio))
        self.fc2 = nn.Linear(int(dim * mlp_ratio), dim)

    def forward(self, x):
        B, N, D = x.shape
        q, k, v = self.qkv(self.norm1(x)).view(B, N, 3, self.heads, D // self.heads).permute(2, 0, 3, 1, 4)
        y = F.scaled_dot_product_attention(q, k, v)  # softmax(q k^T / sqrt(d)) v
        x = x + self.proj(y.transpose(1, 2).reshape(B, N, D))
        return x + self.fc2(F.gelu(self.fc1(self.norm2(x))))


class _ViT(nn.Module):
    def __init__(self, img_size=224, patch=16, dim=768, depth=12, heads=12):
        super().__init__()
        self.patch = patch
        self.grid0 = img_size // patch
        self.patch_embed = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.grid0 * self.grid0 + 1, dim))
        self.blocks = nn.ModuleList([_Block(dim, heads) for _ in range(depth)])
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def pos_encoding(self, n_tokens, H, W):
        """ibot_transformers.py:311-336: bicubic resize of the patch position grid, +0.1 on the target size."""
        N = self.pos_embed.shape[1] - 1
        if n_tokens == N and H == W:
            return self.pos_embed
        h0, w0 = H // self.patch + 0.1, W // self.patch + 0.1
        side = int(math.sqrt(N))
        grid = self.pos_embed[:, 1:].reshape(1, side, side, -1).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, scale_factor=(h0 / side, w0 / side), mode="bicubic")
        return torch.cat((self.pos_embed[:, :1], grid.permute(0, 2, 3, 1).flatten(1, 2)), dim=1)

    def prepare_tokens(self, img):
        B, _, H, W = img.shape
        x = self.patch_embed(img).flatten(2).transpose(1, 2)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        return x + self.pos_encoding(x.shape[1] - 1, H, W)


def vit_b16(seed=0, img_size=224):
    """seeded random-init ViT-B/16 in eval mode (85.8 M parameters)."""
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(10_000 + seed)
        return _ViT(img_size=img_size).eval().requires_grad_(False)


class DenseViT(nn.Module):
    """output="dense" wrapper (dino.py:164-210): list of 4 maps for multilayer, else the last block's map."""

    def __init__(self, vit, multilayer=False, channel_last=False):
        super().__init__()
        self.vit = vit
        n = len(vit.blocks)
        self.layers = [n // 4 - 1, n // 2 - 1, n // 4 * 3 - 1, n - 1] if multilayer else [n - 1]
        self.multilayer = multilayer
        self.channel_last = channel_last

    @torch.no_grad()
    def forward(self, img):
        p = self.vit.patch
        pad_h, pad_w = (-img.shape[-2]) % p, (-img.shape[-1]) % p  # center_padding (utils.py)
        if pad_h or pad_w:
            img = F.pad(img, (pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2))
        h, w = img.shape[-2] // p, img.shape[-1] // p
        x = self.vit.prepare_tokens(img)
        outs = []
        for i, blk in enumerate(self.vit.blocks):
            x = blk(x)
            if i in self.layers:
                t = x[:, -h * w:].reshape(x.shape[0], h, w, -1).permute(0, 3, 1, 2)  # b (h w) c -> b c h w
                outs.append(t if self.channel_last else t.contiguous())
        return outs if self.multilayer else outs[0]

    def features(self, img):
        """(B, C_total, h, w): the multilayer maps concatenated on channels (evaluate_navi_correspondence.py:146-148)."""
        o = self.forward(img)
        if not self.multilayer:
            return o
        if self.channel_last:  # concatenate in token layout, hand back the permuted view
            return torch.cat([t.permute(0, 2, 3, 1) for t in o], dim=-1).permute(0, 3, 1, 2)
        return torch.cat(o, dim=1)


class _ResNetLayer4(nn.Module):
    def __init__(self, net):
        super().__init__()
        self.stem = nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool)
        self.stages = nn.Sequential(net.layer1, net.layer2, net.layer3, net.layer4)

    @torch.no_grad()
    def forward(self, img):
        return self.stages(self.stem(img))

    features = forward


def resnet50_layer4(seed=0):
    """seeded random-init ResNet-50 trunk (MoCo v2 architecture), eval mode: (B, 3, H, W) -> (B, 2048, H/32, W/32)."""
    import torchvision

    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(20_000 + seed)
        net = torchvision.models.resnet50(weights=None)
    net.fc = nn.Identity()
    return _ResNetLayer4(net).eval().requires_grad_(False)


def smooth_images(seed, B, H, W, cell_px=(128, 32, 8), fine=0.15):
    """(B, 3, H, W) seeded low-pass images in roughly [-2, 2]: a sum of bicubically upsampled noise grids (cells of
    128 / 32 / 8 pixels, amplitudes 1, 1/2, 1/3) plus a little pixel noise, generated on the CPU so that every
    machine sees the same bits."""
    g = torch.Generator().manual_seed(30_000 + seed)
    img = torch.zeros(B, 3, H, W)
    for i, px in enumerate(cell_px):
        low = torch.randn(B, 3, max(2, H // px), max(2, W // px), generator=g)
        img += F.interpolate(low, size=(H, W), mode="bicubic", align_corners=False) / (i + 1)
    return img + fine * torch.randn(B, 3, H, W, generator=g)


def _pair_images(seed, H, W, coherent, noise):
    a = smooth_images(2 * seed, 1, H, W)
    if coherent:  # image 1 = image 0 seen again: small photometric change + sensor noise
        g = torch.Generator().manual_seed(40_000 + seed)
        b = 0.95 * a + 0.05 + noise * torch.randn(1, 3, H, W, generator=g)
    else:
        b = smooth_images(2 * seed + 1, 1, H, W)
    return torch.cat((a, b))


def _run(model, imgs, device):
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False  # fp32 like the reference's CPU forward
    try:
        return model.features(imgs.to(device)).float().cpu()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def navi_backbone_pair(index, model=None, device="cpu", coherent=True, noise=0.25, **geom):
    """NAVI-shaped pair (BASELINE configs[1]) whose features come from a random-init ViT-B/16 @ 448, blocks
    [2, 5, 8, 11] concatenated: feat (3072, 28, 28) per image.  Geometry as synthetic.navi_pair."""
    from . import synthetic as syn

    model = model or DenseViT(vit_b16(0), multilayer=True).to(device)
    f = _run(model, _pair_images(index, 448, 448, coherent, noise), device)
    p = syn.navi_pair(index, C=8, coherent=coherent, **geom)
    p["feat_0"], p["feat_1"] = f[0].contiguous(), f[1].contiguous()
    return p


def scannet_backbone_pair(index, model=None, device="cpu", coherent=True, noise=0.25, **geom):
    """ScanNet-shaped pair (configs[2]): random-init ResNet-50 layer4 @ 480 x 640 -> feat (2048, 15, 20)."""
    from . import synthetic as syn

    model = model or resnet50_layer4(0).to(device)
    f = _run(model, _pair_images(1000 + index, 480, 640, coherent, noise), device)
    p = syn.scannet_pair(index, C=8, coherent=coherent, **geom)
    p["feat_0"], p["feat_1"] = f[0].contiguous(), f[1].contiguous()
    return p


def spair_backbone_pair(index, model=None, device="cpu", coherent=True, noise=0.25):
    """SPair-shaped pair (configs[0]): random-init ViT-B/16 @ 224, last block: feats (2, 768, 14, 14)."""
    from . import synthetic as syn

    model = model or DenseViT(vit_b16(0), multilayer=False).to(device)
    f = _run(model, _pair_images(2000 + index, 224, 224, coherent, noise), device)
    p = syn.spair_pair(index, C=8, coherent=coherent)
    p["feats"] = f.contiguous()
    return p
