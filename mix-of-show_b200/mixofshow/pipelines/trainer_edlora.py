"""B200 mirror of the reference trainer (mixofshow/pipelines/trainer_edlora.py:20-380).

`EDLoRATrainer` keeps the reference's constructor (`EDLoRATrainer(**opt['models'])`, train_edlora.py:50), its public names
(`init_new_concept`, `set_finetune_cfg`, `get_params_to_optimize`, `get_all_concept_token_ids`, `forward`,
`delta_state_dict`, `load_delta_state_dict`) and its checkpoint layout ({'new_concept_embedding', 'text_encoder', 'unet'},
keys f'{module}.lora_down.weight' / '.lora_up.weight', :371-378), and trains all THREE parameter groups of :82-139 — the
new-concept embedding rows, the CLIPAttention LoRA and the UNet Attention LoRA — in one captured CUDA graph: text encoder
forward -> UNet forward -> masked MSE + attention regulariser -> UNet backward -> text encoder backward
(mos_b200/train_engine.py + mos_b200/clip_train_engine.py).  The gradients land in ONE flat fp32 buffer (the payload of the
step's single NCCL all-reduce); there is no autograd graph to call `.backward()` on.  `forward` takes images (encoded by
the B200 VAE engine, :203-204) or already encoded latents.

`UNetLoRATrainer` is the latents-and-embeddings level trainer of the UNet LoRA group alone (text encoder frozen and run
upstream)."""
import math
import re

import torch

from mos_b200.engine import ehs_to_layer_major
from mos_b200.train_engine import TrainEngine


class UNetLoRATrainer:
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


# ================================================================================================ full ED-LoRA trainer
class EDLoRATrainer:
    """Reference constructor (trainer_edlora.py:21-68).  `pretrained_path`: diffusers-layout directory with unet/,
    text_encoder/ and tokenizer/.  Keyword extras of the B200 path (all optional, so that `EDLoRATrainer(**opt['models'])`
    works): `tokenizer` (an already constructed tokenizer), `latent_size`, `device`, `seed`.  `enable_xformers` /
    `gradient_checkpoint` are accepted and ignored (the attention kernels are this library's own; activations of one step
    fit HBM many times over)."""

    def __init__(self, pretrained_path, new_concept_token, initializer_token, enable_edlora, finetune_cfg=None,
                 noise_offset=None, attn_reg_weight=None, reg_full_identity=True, use_mask_loss=True,
                 enable_xformers=False, gradient_checkpoint=False, *, tokenizer=None, latent_size=(64, 64), device='cuda',
                 seed=0):
        from mixofshow.utils import model_io
        if not enable_edlora:
            raise NotImplementedError('enable_edlora=False (vanilla LoRA with one embedding per concept) is not built on '
                                      'the B200 path: the cross-attention kernels take layer-wise embeddings')
        self.device = torch.device(device)
        self.enable_edlora = True
        self.unet = model_io.load_unet(pretrained_path)                               # :44
        self.text_encoder = model_io.load_text_encoder(pretrained_path, device=device)   # :41
        if tokenizer is None:
            from transformers import CLIPTokenizer
            tokenizer = CLIPTokenizer.from_pretrained(pretrained_path, subfolder='tokenizer')   # :40
        self.tokenizer = tokenizer
        import os
        self.vae = model_io.load_vae(pretrained_path, device=device) if os.path.isdir(os.path.join(pretrained_path, 'vae')) \
            else None                                                                     # :39
        self._gen = torch.Generator(device='cpu').manual_seed(seed)
        self.new_concept_cfg = self.init_new_concept(new_concept_token, initializer_token, enable_edlora=True)   # :55
        self.attn_reg_weight = attn_reg_weight
        self.reg_full_identity = reg_full_identity
        self.noise_offset = noise_offset
        self.use_mask_loss = use_mask_loss
        self.latent_size = tuple(latent_size)
        self.engine = self.text_engine = self.state = None
        self._batch = None
        self._loaded = None
        if finetune_cfg:
            self.set_finetune_cfg(finetune_cfg)

    # ------------------------------------------------------------------------------------------ new concept tokens
    def init_new_concept(self, new_concept_tokens, initializer_tokens, enable_edlora=True):
        """trainer_edlora.py:144-194: 16 tokens `<new{k}>` per concept word, embedding rows initialised from
        `<rand-sigma>` or from an existing single token."""
        new_concept_cfg = {}
        new_concept_tokens = new_concept_tokens.split('+')
        if initializer_tokens is None:
            initializer_tokens = ['<rand-0.017>'] * len(new_concept_tokens)
        else:
            initializer_tokens = initializer_tokens.split('+')
        assert len(new_concept_tokens) == len(initializer_tokens), 'concept token should match init token.'
        for idx, (concept_name, init_token) in enumerate(zip(new_concept_tokens, initializer_tokens)):
            num_new_embedding = 16 if enable_edlora else 1
            new_token_names = [f'<new{idx * num_new_embedding + layer_id}>' for layer_id in range(num_new_embedding)]
            num_added_tokens = self.tokenizer.add_tokens(new_token_names)
            assert num_added_tokens == len(new_token_names), 'some token is already in tokenizer'
            new_token_ids = [self.tokenizer.convert_tokens_to_ids(token_name) for token_name in new_token_names]
            self.text_encoder.resize_token_embeddings(len(self.tokenizer))
            token_embeds = self.text_encoder.get_input_embeddings().weight.data
            if init_token.startswith('<rand'):
                sigma_val = float(re.findall(r'<rand-(.*)>', init_token)[0])
                init_feature = torch.randn(token_embeds[0].shape, generator=self._gen) * sigma_val
            else:
                init_token_ids = self.tokenizer.encode(init_token, add_special_tokens=False)
                if len(init_token_ids) > 1 or init_token_ids[0] == 40497:        # sic, :179
                    raise ValueError('The initializer token must be a single existing token.')
                init_feature = token_embeds[init_token_ids[0]]
            for token_id in new_token_ids:
                token_embeds[token_id] = init_feature.clone()
            new_concept_cfg.update({concept_name: {'concept_token_ids': new_token_ids,
                                                   'concept_token_names': new_token_names}})
        return new_concept_cfg

    def get_all_concept_token_ids(self):
        ids = []
        for _, cfg in self.new_concept_cfg.items():
            ids.extend(cfg['concept_token_ids'])
        return ids

    # ------------------------------------------------------------------------------------------ configuration
    def set_finetune_cfg(self, finetune_cfg):
        """trainer_edlora.py:70-142: three parameter groups with their own learning rates."""
        te, tx, un = finetune_cfg['text_embedding'], finetune_cfg['text_encoder'], finetune_cfg['unet']
        if not (te.get('enable_tuning') and tx.get('enable_tuning') and tx.get('lora_cfg') and un.get('enable_tuning')
                and un.get('lora_cfg')):
            raise NotImplementedError('the B200 trainer trains the three groups of the shipped ED-LoRA configs together '
                                      '(text_embedding, text_encoder LoRA, unet LoRA); use UNetLoRATrainer for the UNet '
                                      'group alone')
        tcfg, ucfg = dict(tx['lora_cfg']), dict(un['lora_cfg'])
        if tcfg.pop('where') != 'CLIPAttention' or ucfg.pop('where') != 'Attention':
            raise NotImplementedError("lora_cfg.where: 'CLIPAttention' / 'Attention' (every shipped ED-LoRA config) only")
        for c in (tcfg, ucfg):
            if not 1 <= int(c.get('rank', 4)) <= 4:
                raise ValueError('LoRA rank must be in 1..4 (fused epilogue)')
        if 'weight_decay' in te:
            raise NotImplementedError('a per-group weight_decay for the embeddings is not supported by the flat AdamW')
        self.text_rank, self.text_alpha = int(tcfg.get('rank', 4)), float(tcfg.get('alpha', 1.0))
        self.unet_rank, self.unet_alpha = int(ucfg.get('rank', 4)), float(ucfg.get('alpha', 1.0))
        self.lrs = (float(te['lr']), float(tx['lr']), float(un['lr']))
        self.params_to_optimize_iterator = [{'lr': self.lrs[0]}, {'lr': self.lrs[1]}, {'lr': self.lrs[2]}]

    def get_params_to_optimize(self):
        return self.params_to_optimize_iterator

    # ------------------------------------------------------------------------------------------ engines
    def _kaiming(self, rank, K):
        return (torch.rand(rank, K, generator=self._gen) * 2 - 1) / math.sqrt(K)       # edlora.py:238

    def _build(self, batch):
        from mos_b200.clip_train_engine import CLIPTrainEngine
        from mos_b200.dp import FlatTrainState
        c = self.unet.config
        topo = dict(block_out=tuple(c.block_out_channels), layers=c.layers_per_block, heads=c.attention_head_dim,
                    cross_dim=c.cross_attention_dim)
        usd = {k: v.detach() for k, v in self.unet.state_dict().items()}
        tsd = self.text_encoder.state_dict()
        probe = _NameProbe(topo)
        unet_names = TrainEngine.lora_module_names.__get__(probe)()
        n_layers = 1 + max(int(k.split('.layers.')[1].split('.')[0]) for k in tsd if '.layers.' in k)
        text_names = [f'text_model.encoder.layers.{i}.self_attn.{p}' for i in range(n_layers)
                      for p in ('q_proj', 'k_proj', 'v_proj', 'out_proj')]
        ulora, tlora = {}, {}
        for m in unet_names:                                   # LoRALinearLayer init: down kaiming, up zeros (:238-239)
            w = usd[m + '.weight']
            ulora[f'{m}.lora_down.weight'] = self._kaiming(self.unet_rank, w.reshape(w.shape[0], -1).shape[1])
            ulora[f'{m}.lora_up.weight'] = torch.zeros(w.shape[0], self.unet_rank)
        for m in text_names:
            w = tsd[m + '.weight']
            tlora[f'{m}.lora_down.weight'] = self._kaiming(self.text_rank, w.shape[1])
            tlora[f'{m}.lora_up.weight'] = torch.zeros(w.shape[0], self.text_rank)
        ids = self.get_all_concept_token_ids()
        C = tsd['text_model.embeddings.token_embedding.weight'].shape[1]
        heads = getattr(self.text_encoder, 'hf_config', {}).get('num_attention_heads', 12)
        n_text = CLIPTrainEngine.lora_param_count(n_layers, C, heads * 80)
        n_unet = sum(4 * (usd[m + '.weight'].reshape(usd[m + '.weight'].shape[0], -1).shape[1] + usd[m + '.weight'].shape[0])
                     for m in unet_names)
        self.state = FlatTrainState(len(ids), C, n_text, n_unet, lrs=self.lrs, device=self.device)
        H, W = self.latent_size
        self.engine = TrainEngine(usd, batch, H, W, lora=ulora, lora_alpha=self.unet_alpha,
                                  attn_reg_weight=self.attn_reg_weight, reg_full_identity=self.reg_full_identity,
                                  state=self.state, state_offset=self.state.group_end[1], text_grad=True,
                                  device=self.device, **topo)
        n_x = len(self.engine.xattn_names)
        self.text_engine = CLIPTrainEngine(tsd, n_x * batch, lora=tlora, lora_alpha=self.text_alpha,
                                           concept_token_ids=ids, state=self.state, emb_offset=0,
                                           lora_offset=self.state.group_end[0], device=self.device, heads=heads)
        self.engine.attach_text_engine(self.text_engine)
        self._batch = batch
        if self._loaded is not None:
            self._apply_delta(self._loaded)
            self._loaded = None

    # ------------------------------------------------------------------------------------------ step
    def tokenize_layerwise(self, prompts):
        """prompts (b strings) -> token ids [(b 16), 77] as the reference builds them (:223-231) and the same ids in the
        engines' layer-major order [16 * b, 77]."""
        from mixofshow.pipelines.pipeline_edlora import bind_concept_prompt
        bound = bind_concept_prompt(list(prompts), new_concept_cfg=self.new_concept_cfg)
        ids = self.tokenizer(bound, padding='max_length', max_length=self.tokenizer.model_max_length, truncation=True,
                             return_tensors='pt').input_ids
        b = len(prompts)
        n_x = ids.shape[0] // b
        return ids, ids.view(b, n_x, -1).permute(1, 0, 2).reshape(n_x * b, -1).contiguous()

    def forward(self, images, prompts, masks, img_masks, noise=None, timesteps=None, accumulate=False):
        """trainer_edlora.py:202-261.  `images`: [b,3,H,W] in [-1,1] (encoded by the B200 VAE, :203-204) or already encoded
        latents [b,4,h,w] (x 0.18215).  Runs forward + loss + backward of the whole step (text encoder and UNet) and returns
        the loss as a device scalar."""
        if images.shape[1] == 3:                  # :203-204: latents = vae.encode(images).latent_dist.sample() * 0.18215
            if self.vae is None:
                raise ValueError(f'images were given but the model directory has no vae/: pass latents [b,4,h,w] (x 0.18215)')
            dist = self.vae.encode(images.to(self.device)).latent_dist
            latents = (dist.mean + dist.std * torch.randn(dist.mean.shape, generator=self._gen).to(self.device)) * 0.18215
        else:
            latents = images
        b = latents.shape[0]
        if self.engine is None:
            self._build(b)
        if b != self._batch:
            raise ValueError(f'batch {b} != the batch {self._batch} the engines were built for')
        if noise is None:
            noise = torch.randn(latents.shape, generator=self._gen)
            if self.noise_offset is not None:
                noise = noise + self.noise_offset * torch.randn((b, latents.shape[1], 1, 1), generator=self._gen)
        if timesteps is None:
            timesteps = torch.randint(0, 1000, (b,), generator=self._gen)
        ids, ids_lm = self.tokenize_layerwise(prompts)
        n_x = len(self.engine.xattn_names)
        if ids.shape[0] // b != n_x:          # a smaller topology uses the first n_x layer prompts of every sample
            ids_lm = ids.view(b, -1, ids.shape[-1])[:, :n_x].permute(1, 0, 2).reshape(n_x * b, -1).contiguous()
        pos = None
        if self.attn_reg_weight is not None:
            concept = set(int(i) for i in self.get_all_concept_token_ids())
            pos = []
            for text in ids.view(b, -1, ids.shape[-1]):
                p = [i for i in range(text.shape[-1]) if int(text[0][i]) in concept]        # :270-279
                if len(p) != 2:
                    raise ValueError(f'cal_attn_reg assumes exactly two concept tokens per prompt (:298), found {len(p)}')
                pos.append(p)
        loss_mask = masks if self.use_mask_loss else img_masks
        dev = self.device
        out = self.engine.forward_backward(latents.to(dev), noise.to(dev), timesteps.to(dev), None, masks.to(dev),
                                           loss_mask=loss_mask.to(dev), token_pos=pos, accumulate=accumulate,
                                           text_ids=ids_lm)
        return out[0]

    __call__ = forward

    def refresh(self):
        """after an optimiser step on the flat state: re-pack both LoRA sets and write the embedding rows back"""
        self.engine.refresh_lora()
        self.text_engine.refresh_lora()

    # ------------------------------------------------------------------------------------------ checkpoints
    def delta_state_dict(self):
        """trainer_edlora.py:358-378."""
        if self.engine is None:
            raise RuntimeError('delta_state_dict before the first forward: the engines are built on the first batch')
        delta = {'new_concept_embedding': {}, 'text_encoder': {}, 'unet': {}}
        rows = self.text_engine.emb_view.detach().cpu()
        k = 0
        for concept_name, cfg in self.new_concept_cfg.items():
            n = len(cfg['concept_token_ids'])
            delta['new_concept_embedding'][concept_name] = rows[k:k + n].clone()
            k += n
        for key, v in self.text_engine.lora_state_dict().items():
            v = v.cpu()
            delta['text_encoder'][key] = (v[:self.text_rank] if key.endswith('lora_down.weight') else v[:, :self.text_rank]).clone()
        for key, v in self.engine.lora_state_dict().items():
            v = v.cpu()
            delta['unet'][key] = (v[:self.unet_rank] if key.endswith('lora_down.weight') else v[:, :self.unet_rank]).clone()
        return delta

    def load_delta_state_dict(self, delta_state_dict):
        """trainer_edlora.py:315-356 (applied when the engines exist, i.e. at the first batch at the latest)."""
        if self.engine is None:
            self._loaded = delta_state_dict
        else:
            self._apply_delta(delta_state_dict)

    def _apply_delta(self, delta):
        emb = delta.get('new_concept_embedding') or {}
        if emb:
            k = 0
            for concept_name, cfg in self.new_concept_cfg.items():
                n = len(cfg['concept_token_ids'])
                if concept_name in emb:
                    self.text_engine.emb_view[k:k + n].copy_(emb[concept_name].to(self.device, torch.float32))
                k += n
        if delta.get('text_encoder'):
            self.text_engine.load_lora_state_dict(delta['text_encoder'])
        unet = delta.get('unet') or {}
        if unet:
            for m, (D, U, _, _, K, N) in self.engine.lora_views.items():
                d = unet[f'{m}.lora_down.weight'].to(self.device, torch.float32).reshape(-1, K)
                u = unet[f'{m}.lora_up.weight'].to(self.device, torch.float32).reshape(N, -1)
                D.zero_()
                U.zero_()
                D[:d.shape[0]] = d
                U[:, :u.shape[1]] = u
        self.refresh()
