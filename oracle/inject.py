"""ORACLE (test infrastructure). Restated installers that put ED-LoRA behaviour on the oracle UNet skeleton:

  install_edlora_processors   ~ revise_edlora_unet_attention_forward              mixofshow/models/edlora.py:176-190
  install_control_processors  ~ revise_edlora_unet_attention_controller_forward   mixofshow/models/edlora.py:193-218
  install_region_processors   ~ revise_regionally_t2iadapter_attention_forward    pipeline_regionally_t2iadapter.py:148-163
  inject_lora                 ~ LoRALinearLayer installation                      trainer_edlora.py:121-133, edlora.py:221-246
  random_lora_state           synthetic ED-LoRA weights in the reference's checkpoint key layout (trainer_edlora.py:371-378)

Pinned against the reference's own classes by tests/test_oracle_vs_reference.py and tests/golden/*.pt.
"""
import math

import torch
import torch.nn.functional as F

from . import edlora_ref as er


class EDLoRAProcessor:
    """edlora.py:103-173 restated (layer-wise embedding pick at :129-131); optional controller (:22-100)."""

    def __init__(self, cross_attention_idx, place_in_unet=None, controller=None):
        self.cross_attention_idx = cross_attention_idx
        self.place_in_unet = place_in_unet
        self.controller = controller

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **kw):
        is_cross = encoder_hidden_states is not None
        ehs = hidden_states if not is_cross else encoder_hidden_states
        if is_cross and ehs.ndim == 4:
            ehs = ehs[:, self.cross_attention_idx, ...]
        q = attn.head_to_batch_dim(attn.to_q(hidden_states))
        k = attn.head_to_batch_dim(attn.to_k(ehs))
        v = attn.head_to_batch_dim(attn.to_v(ehs))
        probs = er.attention_probs(q, k, attn.scale)
        if self.controller is not None:
            probs = self.controller(probs, is_cross, self.place_in_unet)
        out = attn.batch_to_head_dim(torch.bmm(probs, v))
        return attn.to_out[1](attn.to_out[0](out))


class RegionProcessor:
    """RegionT2I_AttnProcessor.__call__ (pipeline_regionally_t2iadapter.py:88-145) restated."""

    def __init__(self, cross_attention_idx):
        self.cross_attention_idx = cross_attention_idx

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 region_list=None, height=None, width=None):
        is_cross = encoder_hidden_states is not None
        ehs = hidden_states if not is_cross else encoder_hidden_states
        if is_cross and ehs.ndim == 4:
            ehs = ehs[:, self.cross_attention_idx, ...]
        q = attn.head_to_batch_dim(attn.to_q(hidden_states))
        k = attn.head_to_batch_dim(attn.to_k(ehs))
        v = attn.head_to_batch_dim(attn.to_v(ehs))
        if not is_cross and q.shape[1] > 1024:
            out = F.scaled_dot_product_attention(q, k, v, scale=attn.scale)
        else:
            out = torch.bmm(er.attention_probs(q, k, attn.scale), v)
        if is_cross:
            kv, boxes = [], []
            for emb, box in region_list:
                e = emb[:, self.cross_attention_idx, ...] if emb.ndim == 4 else emb
                kv.append((attn.head_to_batch_dim(attn.to_k(e)), attn.head_to_batch_dim(attn.to_v(e))))
                boxes.append(box)
            out = er.region_rewrite(out, q, kv, boxes, height, width, attn.scale)
        out = attn.batch_to_head_dim(out)
        return attn.to_out[1](attn.to_out[0](out))


def _walk(unet, fn):
    count = 0
    for root, place in ((unet.down_blocks, 'down'), (unet.mid_block, 'mid'), (unet.up_blocks, 'up')):
        def rec(mod):
            nonlocal count
            for name, child in mod.named_children():
                if child.__class__.__name__ == 'Attention':
                    count = fn(name, child, count, place)
                else:
                    rec(child)
        rec(root)
    return count


def install_edlora_processors(unet):
    def fn(name, layer, count, place):
        if 'attn2' in name:
            layer.set_processor(EDLoRAProcessor(count))
            return count + 1
        return count
    return _walk(unet, fn)


def install_control_processors(unet, controller):
    def fn(name, layer, count, place):
        if 'attn2' in name:
            layer.set_processor(EDLoRAProcessor(count, place, controller))
            return count + 1
        return count
    n = _walk(unet, fn)
    if controller is not None:
        controller.num_att_layers = n
    return n


def install_region_processors(unet):
    def fn(name, layer, count, place):
        layer.set_processor(RegionProcessor(count))
        return count + 1 if 'attn2' in name else count
    return _walk(unet, fn)


def lora_target_modules(unet, where='Attention'):
    """Module names that get a LoRA for `where: Attention` (trainer_edlora.py:121-133): every Linear/Conv2d child
    of every module whose class name is `where`."""
    names = []
    for name, module in unet.named_modules():
        if module.__class__.__name__ == where:
            for child_name, child in module.named_modules():
                if child.__class__.__name__ in ('Linear', 'Conv2d') and child_name != '':
                    if child.__class__.__name__ == 'Conv2d' and child.kernel_size != (1, 1):
                        continue
                    names.append(f'{name}.{child_name}')
    return names


def random_lora_state(unet, seed=10, rank=4, where='Attention', up_std=0.02):
    """Synthetic ED-LoRA in the reference checkpoint layout; down ~ kaiming_uniform(a=sqrt(5)) as edlora.py:238,
    up ~ N(0, up_std^2) instead of zeros so the low-rank path is exercised."""
    g = torch.Generator().manual_seed(seed)
    mods = dict(unet.named_modules())
    state = {}
    for n in lora_target_modules(unet, where):
        m = mods[n]
        conv = m.__class__.__name__ == 'Conv2d'
        cin = m.in_channels if conv else m.in_features
        cout = m.out_channels if conv else m.out_features
        bound = 1.0 / math.sqrt(cin)  # kaiming_uniform_(a=sqrt(5)) on [rank, cin]
        down = (torch.rand(rank, cin, generator=g) * 2 - 1) * bound
        up = torch.randn(cout, rank, generator=g) * up_std
        if conv:
            down, up = down[:, :, None, None], up[:, :, None, None]
        state[f'{n}.lora_down.weight'] = down
        state[f'{n}.lora_up.weight'] = up
    return state


def inject_lora(unet, lora_state, alpha=1.0):
    """Wrap the forward of every module that has a LoRA pair: y = orig(x) + alpha*up(down(x)) (edlora.py:244-246)."""
    mods = dict(unet.named_modules())
    n = 0
    for key in lora_state:
        if not key.endswith('.lora_down.weight'):
            continue
        name = key[:-len('.lora_down.weight')]
        m = mods[name]
        down, up = lora_state[key], lora_state[name + '.lora_up.weight']

        def fwd(x, m=m, down=down, up=up, orig=m.forward):
            if down.ndim == 4:
                return orig(x) + alpha * F.conv2d(F.conv2d(x, down), up)
            return orig(x) + alpha * F.linear(F.linear(x, down), up)
        m.forward = fwd
        n += 1
    return n
