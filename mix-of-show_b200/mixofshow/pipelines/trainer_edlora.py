"""B200 mirror of the reference trainer (mixofshow/pipelines/trainer_edlora.py:20-380) for the UNet half of ED-LoRA
training.  Same public names (`EDLoRATrainer`, `set_finetune_cfg`, `get_params_to_optimize`,
`get_all_concept_token_ids`, `forward`, `delta_state_dict`, `load_delta_state_dict`), same checkpoint layout
({'new_concept_embedding', 'text_encoder', 'unet'}, keys f'{module}.lora_down.weight' / '.lora_up.weight', :371-378).

Scope: the VAE encoder and the CLIP text encoder are "next", so `forward` takes what they produce —
latents (already x 0.18215, :204) and the layer-wise text embeddings `[b, 16, 77, 768]` (:232-234) — instead of images
and prompts, and only the UNet LoRA group is trained.  `forward` runs forward + loss + backward in one captured CUDA
graph (the gradient lands in the flat buffer the single NCCL all-reduce uses); there is no autograd graph to call
`.backward()` on.
"""
import math

import torch

from mos_b200.engine import ehs_to_layer_major
from mos_b200.train_engine import TrainEngine


class EDLoRATrainer:
    def __init__(self, unet_state_dict, batch_size_per_gpu, new_concept_cfg=None, finetune_cfg=None, noise_offset=None,
                 attn_reg_weight=None, reg_full_identity=True, use_mask_loss=True, latent_size=(64, 64),
                 unet_topology=None, device='cuda', seed=0, lora_state=None):
        if finetune_cfg is None:
            raise ValueError('finetune_cfg is required (trainer_edlora.py:66-67)')
        self.new_concept_cfg = dict(new_concept_cfg or {})
        self.noise_offset = noise_offset
        self.attn_reg_weight = attn_reg_weight
        self.reg_full_identity = reg_full_identity
        self.use_mask_loss = use_mask_loss
        self.device = torch.device(device)
        self.batch = int(batch_size_per_gpu)
        self.latent_size = tuple(latent_size)
        self._topo = dict(unet_topology or {})
        self._gen = torch.Generator(device='cpu').manual_seed(seed)
        self._sd = unet_state_dict
        self.set_finetune_cfg(finetune_cfg, lora_state)

    # ------------------------------------------------------------------------------------------ configuration
    def set_finetune_cfg(self, finetune_cfg, lora_state=None):
        """trainer_edlora.py:71-142.  Only the `unet` group exists on this path."""
        for part in ('text_embedding', 'text_encoder'):
            if finetune_cfg.get(part, {}).get('enable_tuning'):
                raise NotImplementedError(f"finetune_cfg['{part}'].enable_tuning: the CLIP side of ED-LoRA training is "
                                          'not built on the B200 path yet (SURVEY.md §8f); disable it')
        ucfg = finetune_cfg['unet']
        if not (ucfg.get('enable_tuning') and ucfg.get('lora_cfg')):
            raise ValueError("finetune_cfg['unet'] must enable tuning with a lora_cfg")
        lora_cfg = dict(ucfg['lora_cfg'])
        where = lora_cfg.pop('where')
        if where != 'Attention':
            raise NotImplementedError("lora_cfg.where: only 'Attention' (the shipped ED-LoRA configs) is supported")
        self.rank = int(lora_cfg.get('rank', 4))
        self.alpha = float(lora_cfg.get('alpha', 1.0))
        if not 1 <= self.rank <= 4:
            raise ValueError('LoRA rank must be in 1..4 (fused epilogue)')
        self.unet_lr = float(ucfg['lr'])
        H, W = self.latent_size
        probe_names = TrainEngine.lora_module_names.__get__(_NameProbe(self._topo))()
        if lora_state is None:
            lora_state = self._init_lora(probe_names)
        self.engine = TrainEngine(self._sd, self.batch, H, W, lora=lora_state, lora_alpha=self.alpha,
                                  attn_reg_weight=self.attn_reg_weight, reg_full_identity=self.reg_full_identity,
                                  lr=self.unet_lr, device=self.device, **self._topo)
        self._sd = None
        self.params_to_optimize_iterator = [{'params': [self.engine.state.params], 'lr': self.unet_lr}]

    def _init_lora(self, names):
        """LoRALinearLayer init (edlora.py:238-239): down ~ kaiming_uniform(a=sqrt(5)), up = 0."""
        state = {}
        for m in names:
            w = self._sd[m + '.weight']
            N, K = w.shape[0], w.reshape(w.shape[0], -1).shape[1]
            bound = 1.0 / math.sqrt(K)
            state[f'{m}.lora_down.weight'] = (torch.rand(self.rank, K, generator=self._gen) * 2 - 1) * bound
            state[f'{m}.lora_up.weight'] = torch.zeros(N, self.rank)
        return state

    def get_params_to_optimize(self):
        return self.params_to_optimize_iterator

    def get_all_concept_token_ids(self):
        ids = []
        for _, cfg in self.new_concept_cfg.items():
            ids.extend(cfg['concept_token_ids'])
        return ids

    # ------------------------------------------------------------------------------------------ step
    def concept_token_positions(self, text_input_ids):
        """trainer_edlora.py:270-279 (host integers): positions of the concept tokens in each sample's layer-0 prompt."""
        b = self.batch
        ids = text_input_ids.reshape(b, -1, text_input_ids.shape[-1]).cpu()
        concept = set(int(i) for i in self.get_all_concept_token_ids())
        pos = []
        for text in ids:
            p = [i for i in range(text.shape[-1]) if int(text[0][i]) in concept]
            if len(p) != 2:
                raise ValueError(f'cal_attn_reg assumes exactly two concept tokens per prompt (:298), found {len(p)}')
            pos.append(p)
        return pos

    def forward(self, latents, encoder_hidden_states, masks, img_masks, text_input_ids=None, noise=None,
                timesteps=None, accumulate=False):
        """latents fp32 [b,4,h,w] (VAE output x 0.18215); encoder_hidden_states [b,16,77,768]; masks / img_masks
        [b,1,h,w].  noise / timesteps may be given (tests); otherwise sampled as trainer_edlora.py:207-214."""
        b = latents.shape[0]
        if b != self.batch:
            raise ValueError(f'batch {b} != batch_size_per_gpu {self.batch} the engine was built for')
        if noise is None:
            noise = torch.randn(latents.shape, generator=self._gen)
            if self.noise_offset is not None:
                noise = noise + self.noise_offset * torch.randn((b, latents.shape[1], 1, 1), generator=self._gen)
        if timesteps is None:
            timesteps = torch.randint(0, 1000, (b,), generator=self._gen)
        pos = None
        if self.attn_reg_weight is not None:
            if text_input_ids is None:
                raise ValueError('the attention regulariser needs text_input_ids to locate the concept tokens (:270)')
            pos = self.concept_token_positions(text_input_ids)
        loss_mask = masks if self.use_mask_loss else img_masks
        eng = self.engine
        n_layers = len(eng.xattn_names)
        out = eng.forward_backward(latents.to(self.device), noise.to(self.device), timesteps.to(self.device),
                                   ehs_to_layer_major(encoder_hidden_states.to(self.device), n_layers),
                                   masks.to(self.device), loss_mask=loss_mask.to(self.device), token_pos=pos,
                                   accumulate=accumulate)
        return out[0]

    __call__ = forward

    # ------------------------------------------------------------------------------------------ checkpoints
    def delta_state_dict(self):
        """trainer_edlora.py:358-378 layout; the text-side sections stay empty on this path."""
        delta = {'new_concept_embedding': {}, 'text_encoder': {}, 'unet': {}}
        for k, v in self.engine.lora_state_dict().items():
            v = v.cpu()
            delta['unet'][k] = v[:self.rank].clone() if k.endswith('lora_down.weight') else v[:, :self.rank].clone()
        return delta

    def load_delta_state_dict(self, delta_state_dict):
        """trainer_edlora.py:315-356 (unet section)."""
        unet = delta_state_dict.get('unet', {})
        if len(unet) == 0:
            return
        views = self.engine.lora_views
        if len(unet) != 2 * len(views):
            raise ValueError(f'checkpoint has {len(unet)} unet tensors, the model has {2 * len(views)} LoRA tensors')
        for m, (D, U, _, _, K, N) in views.items():
            d = unet[f'{m}.lora_down.weight'].to(self.device, torch.float32).reshape(-1, K)
            u = unet[f'{m}.lora_up.weight'].to(self.device, torch.float32).reshape(N, -1)
            D.zero_()
            U.zero_()
            D[:d.shape[0]] = d
            U[:, :u.shape[1]] = u
        self.engine.refresh_lora()


class _NameProbe:
    """enough of the engine surface for TrainEngine.lora_module_names before the engine exists"""

    def __init__(self, topo):
        from mos_b200.engine import cross_attention_names
        self.xattn_names = cross_attention_names(topo.get('block_out', (320, 640, 1280, 1280)), topo.get('layers', 2))
