"""Drop-in for the reference's `mixofshow/pipelines/pipeline_edlora.py`: `bind_concept_prompt` and `EDLoRAPipeline`
with the same constructor / `set_new_concept_cfg` / `set_controller` / `__call__` surface.

The denoise loop (reference :271-301) runs on the B200 engine: per step one captured UNet graph (CFG batch 2) and
ONE fused kernel for CFG combine + DPM-Solver++(2M) update + re-duplication of the latents (`mos_cfg_dpmpp_step`).
Prompt encoding and image decoding run on the B200 CLIP / VAE engines (mixofshow/models/clip_b200.py, vae_b200.py) when
`from_pretrained` finds `text_encoder/` and `vae/`; without them pass `prompt_embeds` and request `output_type='latent'`.
"""
from types import SimpleNamespace
from typing import List, Optional, Union

import torch

from mixofshow.models.edlora import (revise_edlora_unet_attention_controller_forward,
                                     revise_edlora_unet_attention_forward)
from mos_b200 import ops
from mos_b200.scheduler import DPMSolverPP2M


def numpy_to_pil(images):
    """[N,H,W,3] floats in [0,1] -> list of PIL images (diffusers `DiffusionPipeline.numpy_to_pil`, reached from the
    reference at pipeline_edlora.py:313), so that `.images[0].save(...)` works as in the reference's scripts."""
    from PIL import Image
    if images.ndim == 3:
        images = images[None]
    return [Image.fromarray(im) for im in (images * 255).round().astype('uint8')]


def bind_concept_prompt(prompts, new_concept_cfg):
    """Each prompt becomes 16 layer-specific prompts; `concept_name` -> `concept_token_names[layer]`
    (reference :18-29). Output order: prompt-major, layer fastest."""
    if isinstance(prompts, str):
        prompts = [prompts]
    new_prompts = []
    for prompt in prompts:
        per_layer = [prompt] * 16
        for concept_name, new_token_cfg in new_concept_cfg.items():
            per_layer = [p.replace(concept_name, new_name)
                         for p, new_name in zip(per_layer, new_token_cfg['concept_token_names'])]
        new_prompts.extend(per_layer)
    return new_prompts


class EDLoRAPipeline:
    def __init__(self, vae=None, text_encoder=None, tokenizer=None, unet=None, scheduler=None, safety_checker=None,
                 feature_extractor=None, requires_safety_checker: bool = False):
        assert unet is not None, 'EDLoRAPipeline needs the B200 UNet'
        revise_edlora_unet_attention_forward(unet)          # reference :93
        self.vae, self.text_encoder, self.tokenizer, self.unet = vae, text_encoder, tokenizer, unet
        self.scheduler = scheduler if scheduler is not None else DPMSolverPP2M()
        # diffusers: 2 ** (len(vae.config.block_out_channels) - 1); 8 for SD1.5 (and when no VAE is attached)
        self.vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1) if vae is not None else 8
        self.new_concept_cfg = None
        self.device = torch.device('cuda')

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, scheduler=None, vae=None, tokenizer=None, device='cuda',
                        **unused):
        """diffusers call shape (`EDLoRAPipeline.from_pretrained(path, scheduler=..., torch_dtype=...)`, test_edlora.py /
        README.md:146): loads `unet/` and `text_encoder/` of a diffusers-layout directory into the B200 containers
        (mixofshow/utils/model_io.py), `vae/` (when present) into the B200 VAE and `tokenizer/` through transformers."""
        import os
        from mixofshow.utils import model_io
        unet = model_io.load_unet(pretrained_model_name_or_path)
        text_encoder = model_io.load_text_encoder(pretrained_model_name_or_path, device=device)
        if vae is None and os.path.isdir(os.path.join(pretrained_model_name_or_path, 'vae')):
            vae = model_io.load_vae(pretrained_model_name_or_path, device=device)
        if tokenizer is None:
            from transformers import CLIPTokenizer
            tokenizer = CLIPTokenizer.from_pretrained(pretrained_model_name_or_path, subfolder='tokenizer')
        pipe = cls(vae=vae, text_encoder=text_encoder, tokenizer=tokenizer, unet=unet, scheduler=scheduler)
        if os.path.exists(os.path.join(pretrained_model_name_or_path, 'new_concept_cfg.json')):   # a fused model
            cfg = model_io.load_new_concept_cfg(pretrained_model_name_or_path)
            model_io.ensure_concept_tokens(tokenizer, cfg)
            pipe.set_new_concept_cfg(cfg)
        return pipe.to(device)

    def to(self, device):
        self.device = torch.device(device)
        return self

    def set_new_concept_cfg(self, new_concept_cfg=None):
        self.new_concept_cfg = new_concept_cfg

    def set_controller(self, controller):
        self.controller = controller
        revise_edlora_unet_attention_controller_forward(self.unet, controller)

    # reference :111-190
    def _encode_prompt(self, prompt, new_concept_cfg, device, num_images_per_prompt, do_classifier_free_guidance,
                       negative_prompt=None, prompt_embeds=None, negative_prompt_embeds=None):
        assert num_images_per_prompt == 1, 'only support num_images_per_prompt=1 now'
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        if prompt_embeds is None:
            if self.tokenizer is None or self.text_encoder is None:
                raise ValueError('no tokenizer / text_encoder supplied: pass prompt_embeds [B,16,77,768]')
            prompt_extend = bind_concept_prompt(prompt, new_concept_cfg)
            ids = self.tokenizer(prompt_extend, padding='max_length', max_length=self.tokenizer.model_max_length,
                                 truncation=True, return_tensors='pt').input_ids
            prompt_embeds = self.text_encoder(ids.to(device))[0]
            prompt_embeds = prompt_embeds.reshape(batch_size, -1, *prompt_embeds.shape[1:])   # '(b n) m c -> b n m c'
        prompt_embeds = prompt_embeds.to(device)
        bs_embed, layer_num, seq_len, _ = prompt_embeds.shape
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            if self.tokenizer is None or self.text_encoder is None:
                raise ValueError('classifier-free guidance needs negative_prompt_embeds [B,77,768] when no text '
                                 'encoder is supplied')
            if negative_prompt is None:
                uncond_tokens = [''] * batch_size
            elif type(prompt) is not type(negative_prompt):
                raise TypeError(f'`negative_prompt` should be the same type to `prompt`, but got '
                                f'{type(negative_prompt)} != {type(prompt)}.')
            elif isinstance(negative_prompt, str):
                uncond_tokens = [negative_prompt]
            elif batch_size != len(negative_prompt):
                raise ValueError(f'`negative_prompt`: {negative_prompt} has batch size {len(negative_prompt)}, but '
                                 f'`prompt`: {prompt} has batch size {batch_size}.')
            else:
                uncond_tokens = negative_prompt
            ids = self.tokenizer(uncond_tokens, padding='max_length', max_length=seq_len, truncation=True,
                                 return_tensors='pt').input_ids
            negative_prompt_embeds = self.text_encoder(ids.to(device))[0]
        if do_classifier_free_guidance:
            seq_len = negative_prompt_embeds.shape[1]
            negative_prompt_embeds = negative_prompt_embeds.to(device)
            negative_prompt_embeds = negative_prompt_embeds.view(batch_size, 1, seq_len, -1).repeat(1, layer_num, 1, 1)
            prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds])
        return prompt_embeds

    @torch.no_grad()
    def __call__(self, prompt: Union[str, List[str]] = None, height: Optional[int] = None,
                 width: Optional[int] = None, num_inference_steps: int = 50, guidance_scale: float = 7.5,
                 negative_prompt=None, num_images_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None,
                 latents: Optional[torch.Tensor] = None, prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_embeds: Optional[torch.Tensor] = None, output_type: Optional[str] = 'pil',
                 return_dict: bool = True, callback=None, callback_steps: int = 1, cross_attention_kwargs=None):
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f'`height` and `width` have to be divisible by 8 but are {height} and {width}.')
        if prompt is not None and isinstance(prompt, str):
            batch_size = 1
        elif prompt is not None and isinstance(prompt, list):
            batch_size = len(prompt)
        else:
            batch_size = prompt_embeds.shape[0]
        device = self.device
        do_cfg = guidance_scale > 1.0
        assert self.new_concept_cfg is not None
        prompt_embeds = self._encode_prompt(prompt, self.new_concept_cfg, device, num_images_per_prompt, do_cfg,
                                            negative_prompt, prompt_embeds=prompt_embeds,
                                            negative_prompt_embeds=negative_prompt_embeds)
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        timesteps = [int(t) for t in self.scheduler.timesteps]
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        shape = (batch_size, self.unet.in_channels, h, w)
        if latents is None:
            latents = torch.randn(shape, generator=generator, device=generator.device if generator is not None
                                  else 'cpu').to(device)
        latents = (latents.to(device, torch.float32) * self.scheduler.init_noise_sigma).contiguous()
        assert tuple(latents.shape) == shape, f'latents shape {tuple(latents.shape)} != {shape}'

        controller = getattr(self, 'controller', None)
        x0_prev = torch.zeros_like(latents)
        nb = 2 * batch_size if do_cfg else batch_size
        # one prepared session: weights verified / packed and the step-invariant inputs uploaded ONCE; a step is then one
        # graph replay + one fused kernel, with no host synchronisation inside the loop (reference loop :271-301)
        sess = self.unet.session(nb, h, w, device, prompt_embeds, cross_attention_kwargs)
        unet_in = sess.latents_in
        unet_in.copy_(torch.cat([latents] * 2) if do_cfg else latents)
        sess.t_in.fill_(float(timesteps[0]))
        for i, t in enumerate(timesteps):
            noise_pred = sess.step()
            # CFG combine + scheduler.step + cat([latents]*2) + next timestep, one kernel (reference :285-290, :273)
            t_next = float(timesteps[i + 1]) if i + 1 < len(timesteps) else 0.0
            ops.cfg_dpmpp_step(noise_pred, latents, x0_prev, unet_in.view(-1), cfg=do_cfg,
                               guidance=float(guidance_scale), coef=self.scheduler.coefficients(i), t_out=sess.t_in,
                               t_next=t_next)
            if controller is not None and hasattr(controller, 'step_callback'):
                new_latents = controller.step_callback(latents)
                if new_latents is not latents:      # a controller may return edited latents (reference :293-295)
                    latents.copy_(new_latents.to(latents.dtype))
                    unet_in.copy_(torch.cat([latents] * 2) if do_cfg else latents)
            if callback is not None and i % callback_steps == 0:
                callback(i, t, latents)
        if output_type == 'latent':
            image = latents
        else:
            if self.vae is None:
                raise ValueError("no VAE supplied: use output_type='latent'")
            image = self.vae.decode(latents / 0.18215).sample
            image = (image / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).float().numpy()
            if output_type == 'pil':
                image = numpy_to_pil(image)
        if not return_dict:
            return (image)
        return SimpleNamespace(images=image, nsfw_content_detected=None)
