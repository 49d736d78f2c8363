"""Drop-in for the `vae` object the reference passes around (diffusers `AutoencoderKL`, loaded at
mixofshow/pipelines/trainer_edlora.py:39 and used at :203-204 and pipeline_edlora.py:303-313): same call shapes —
`vae.encode(images).latent_dist.sample()` and `vae.decode(z).sample` — running on `mos_b200.vae_engine.VAEEngine`
(inference only: the VAE is frozen in every reference workflow, trainer_edlora.py:73-76).  Engines are built per input shape on
first use (buffers are static)."""
from types import SimpleNamespace

import torch

# diffusers < 0.18 attention parameter names -> 0.19 names (checkpoints in the wild carry either)
_OLD_ATTN = {'query': 'to_q', 'key': 'to_k', 'value': 'to_v', 'proj_attn': 'to_out.0'}


def normalise_keys(state_dict):
    out = {}
    for k, v in state_dict.items():
        parts = k.split('.')
        if 'attentions' in parts and len(parts) >= 2 and parts[-2] in _OLD_ATTN:
            parts[-2:-1] = _OLD_ATTN[parts[-2]].split('.')
            k = '.'.join(parts)
        out[k] = v
    return out


class DiagonalGaussian:
    """`latent_dist` of `AutoencoderKLOutput`: mean / logvar (clamped) with `sample(generator)` and `mode()`."""

    def __init__(self, mean, logvar):
        self.mean, self.logvar = mean, logvar
        self.std = torch.exp(0.5 * logvar)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device if generator is None or
                            generator.device.type != 'cpu' else 'cpu').to(self.mean.device)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class AutoencoderKL:
    def __init__(self, state_dict, *, block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4,
                 scaling_factor=0.18215, device='cuda'):
        self._sd = {k: v.detach().to(torch.float32) for k, v in normalise_keys(state_dict).items()}
        self.config = SimpleNamespace(block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      latent_channels=latent_channels, scaling_factor=scaling_factor, in_channels=3,
                                      out_channels=3)
        self.device = torch.device(device)
        self.dtype = torch.float32
        self._engines = {}

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder='vae', **kw):
        """diffusers call shape (`AutoencoderKL.from_pretrained(path, subfolder='vae')`, trainer_edlora.py:39)."""
        from mixofshow.utils.model_io import load_vae
        return load_vae(pretrained_model_name_or_path, subfolder, **{k: v for k, v in kw.items() if k == 'device'})

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def state_dict(self):
        return dict(self._sd)

    def parameters(self):
        return iter(self._sd.values())

    def _engine(self, B, H, W):
        from mos_b200.vae_engine import VAEEngine
        key = (B, H, W)
        if key not in self._engines:
            c = self.config
            self._engines[key] = VAEEngine(self._sd, B, H, W, device=self.device, block_out=c.block_out_channels,
                                           layers=c.layers_per_block, latent_channels=c.latent_channels,
                                           scaling_factor=c.scaling_factor)
        return self._engines[key]

    @torch.no_grad()
    def encode(self, x, return_dict=True):
        B, _, H, W = x.shape
        mean, logvar = self._engine(B, H, W).encode(x)
        dist = DiagonalGaussian(mean.clone(), logvar.clone())
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    @torch.no_grad()
    def decode(self, z, return_dict=True):
        B, _, h, w = z.shape
        d = 2 ** (len(self.config.block_out_channels) - 1)
        img = self._engine(B, h * d, w * d).decode(z).clone()
        return SimpleNamespace(sample=img) if return_dict else (img,)
