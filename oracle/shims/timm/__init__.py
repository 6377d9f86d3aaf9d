"""TEST INFRASTRUCTURE ONLY -- import shim standing in for the `timm` package.

`timm` is a third-party dependency of the reference (requirements.txt:7, `timm>=0.6.7`,
model id `deit3_medium_patch16_224.fb_in22k_ft_in1k`, allied_files.py:27) and is NOT
installed in this image (no network).  The reference reaches it from
model.py:17-18,22 and axial_model.py:47-48,52.  This shim restates timm's published
`VisionTransformer` algorithm for the DeiT-III family so that /root/reference/model.py
can be imported and executed VERBATIM as the parity oracle:

  patch_size 16, qkv_bias, mlp_ratio 4, pre-norm blocks with LayerScale (init 1e-6),
  LayerNorm eps 1e-6, exact-erf GELU, no_embed_class=True (pos_embed has one row per
  patch and is added BEFORE the cls token is concatenated), final norm, and with
  `num_classes=0, global_pool=''` the full (B, 1+n, D) token matrix is returned.

State-dict key names follow timm so reference checkpoints load.
PARITY UNPINNED at this boundary: real timm cannot be cross-checked offline.
Nothing outside tests/, bench.py's cpu_baseline/reference arm and
__graft_entry__.smoke() may import this.
"""
import math
import torch
from torch import nn
import torch.nn.functional as F

__version__ = "0.0-shim"

_FAMILY = {
    # name -> (embed_dim, depth, heads)
    "deit3_small_patch16_224": (384, 12, 6),
    "deit3_medium_patch16_224": (512, 12, 8),
    "deit3_base_patch16_224": (768, 12, 12),
    "deit3_large_patch16_224": (1024, 24, 16),
}


class _PatchEmbed(nn.Module):
    def __init__(self, img_size, patch, in_chans, dim):
        super().__init__()
        self.img_size = img_size
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        assert x.shape[-2] == self.img_size and x.shape[-1] == self.img_size, \
            "Input size doesn't match model"
        return self.proj(x).flatten(2).transpose(1, 2)


class _Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        a = (q * (C // self.num_heads) ** -0.5) @ k.transpose(-2, -1)
        a = a.softmax(dim=-1)
        return self.proj((a @ v).transpose(1, 2).reshape(B, N, C))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _LayerScale(nn.Module):
    def __init__(self, dim, init_values):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class _Block(nn.Module):
    def __init__(self, dim, heads, init_values):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, heads)
        self.ls1 = _LayerScale(dim, init_values)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, dim * 4)
        self.ls2 = _LayerScale(dim, init_values)

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        return x + self.ls2(self.mlp(self.norm2(x)))


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=512, depth=12,
                 num_heads=8, init_values=1e-6):
        super().__init__()
        self.embed_dim = self.num_features = embed_dim
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, n, embed_dim) * .02)
        self.blocks = nn.Sequential(*[_Block(embed_dim, num_heads, init_values) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        x = self.patch_embed(x) + self.pos_embed          # no_embed_class: pos before cls
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        return self.norm(self.blocks(x))                  # global_pool='', head = Identity


def create_model(model_name, pretrained=False, num_classes=0, global_pool='', img_size=224, **kw):
    if pretrained:
        raise RuntimeError("timm shim: no pretrained weights offline")
    assert num_classes == 0 and global_pool == '', "shim restates only the feature path"
    base = model_name.split('.')[0]
    dim, depth, heads = _FAMILY[base]
    return VisionTransformer(img_size=img_size, embed_dim=dim, depth=depth, num_heads=heads)
