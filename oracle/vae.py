"""ORACLE (test infrastructure, never imported by the product path).

Plain-PyTorch fp32 CPU restatement of the SD1.5 `AutoencoderKL` the reference drives through diffusers (third-party, absent
from /root/reference: `diffusers`, recommended ==0.19.3 in README.md:66, unpinned in requirements.txt:2).  Restated from the
published diffusers-0.19.3 algorithm with runwayml/stable-diffusion-v1-5 `vae/config.json` values (block_out_channels
(128, 256, 512, 512), layers_per_block 2, latent_channels 4, norm_num_groups 32, act silu, scaling_factor 0.18215);
anchored on the reference's call sites:
    latents = self.vae.encode(images).latent_dist.sample() * 0.18215      mixofshow/pipelines/trainer_edlora.py:203-204
    image = self.vae.decode(latents / 0.18215).sample                     mixofshow/pipelines/pipeline_edlora.py:303-313
                                                                          (diffusers decode_latents)
Module / parameter names equal diffusers' (0.19 attention names `to_q / to_k / to_v / to_out.0 / group_norm`).

PARITY PINNING: the reference holds no test or golden vector for this boundary and diffusers is not
installed here, so this restatement is "parity unpinned" against diffusers itself, exactly like oracle/unet.py.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

SD15_VAE = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215)


class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D with temb_channels=None, eps 1e-6, output_scale_factor 1."""

    def __init__(self, cin, cout, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """diffusers `Attention` as built by UNetMidBlock2D for the VAE: one head of dim C, GroupNorm(32, eps 1e-6) on the
    input, biased projections, residual connection, rescale_output_factor 1."""

    def __init__(self, channels, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        res = x
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        p = (torch.bmm(q, k.transpose(1, 2)) * c ** -0.5).softmax(-1)
        o = self.to_out[0](torch.bmm(p, v))
        return o.transpose(1, 2).reshape(b, c, h, w) + res


class MidBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attentions = nn.ModuleList([Attention(c)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c), ResnetBlock2D(c, c)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Downsample2D(nn.Module):
    """diffusers Downsample2D(padding=0): pad (0,1,0,1) then a stride-2 3x3 convolution without padding."""

    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode='constant', value=0))


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode='nearest'))


class _Block(nn.Module):
    def __init__(self, cin, cout, n, down=False, up=False):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(n)])
        if down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
        if up:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if hasattr(self, 'downsamplers'):
            x = self.downsamplers[0](x)
        if hasattr(self, 'upsamplers'):
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch, L = cfg['block_out_channels'], cfg['layers_per_block']
        self.conv_in = nn.Conv2d(cfg['in_channels'], ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, c in enumerate(ch):
            cin, out = out, c
            self.down_blocks.append(_Block(cin, out, L, down=i < len(ch) - 1))
        self.mid_block = MidBlock(ch[-1])
        self.conv_norm_out = nn.GroupNorm(cfg['norm_num_groups'], ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg['latent_channels'], 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch, L = cfg['block_out_channels'], cfg['layers_per_block']
        rev = list(reversed(ch))
        self.conv_in = nn.Conv2d(cfg['latent_channels'], rev[0], 3, padding=1)
        self.mid_block = MidBlock(rev[0])
        self.up_blocks = nn.ModuleList()
        out = rev[0]
        for i, c in enumerate(rev):
            cin, out = out, c
            self.up_blocks.append(_Block(cin, out, L + 1, up=i < len(ch) - 1))
        self.conv_norm_out = nn.GroupNorm(cfg['norm_num_groups'], ch[0], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[0], cfg['out_channels'], 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKL(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        cfg = dict(SD15_VAE, **(cfg or {}))
        self.cfg = cfg
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg['latent_channels'], 2 * cfg['latent_channels'], 1)
        self.post_quant_conv = nn.Conv2d(cfg['latent_channels'], cfg['latent_channels'], 1)

    def moments(self, x):
        """(mean, logvar) of the diagonal Gaussian posterior; logvar clamped to [-30, 20] (DiagonalGaussianDistribution)."""
        m = self.quant_conv(self.encoder(x))
        mean, logvar = m.chunk(2, dim=1)
        return mean, logvar.clamp(-30.0, 20.0)

    def encode_sample(self, x, noise):
        """`vae.encode(x).latent_dist.sample()` with the standard-normal draw supplied by the caller."""
        mean, logvar = self.moments(x)
        return mean + torch.exp(0.5 * logvar) * noise

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))


def build_vae(seed=0, cfg=None):
    torch.manual_seed(seed)
    return AutoencoderKL(cfg).eval()


TINY_VAE = dict(block_out_channels=(128, 256), layers_per_block=1)   # small variant for fast tests (one 2x down / up level)
