"""ORACLE (test infrastructure, never imported by the product path).

CPU fp32 restatement, in plain tensor math, of everything the reference itself owns on the hot path
(SURVEY.md §8a rows P1-P6, R1, G1, G3).  Each function cites the reference lines it follows.  Pinned against the
reference's own modules executed under import shims: tests/golden/make_golden.py -> tests/golden/*.pt and
tests/test_oracle_vs_reference.py (runs where /root/reference exists).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------- P1  LoRALinearLayer.forward  edlora.py:244-246
def lora_linear(x, weight, bias, down, up, alpha):
    """y = orig(x) + alpha * up(down(x)); Linear, or 1x1 Conv2d when weight is 4-D (edlora.py:227-234)."""
    if weight.ndim == 4:
        y = F.conv2d(x, weight, bias)
        return y + alpha * F.conv2d(F.conv2d(x, down), up)
    y = F.linear(x, weight, bias)
    return y + alpha * F.linear(F.linear(x, down), up)


# ---------------------------------------------------------------- U4  head split helpers (diffusers Attention)
def head_to_batch(t, heads):
    b, n, c = t.shape
    return t.reshape(b, n, heads, c // heads).permute(0, 2, 1, 3).reshape(b * heads, n, c // heads)


def batch_to_head(t, heads):
    bh, n, d = t.shape
    return t.reshape(bh // heads, heads, n, d).permute(0, 2, 1, 3).reshape(bh // heads, n, d * heads)


def attention_probs(q, k, scale):
    """attn.get_attention_scores: baddbmm(beta=0, alpha=scale) -> softmax(-1)   (edlora.py:81,155)."""
    return (torch.bmm(q, k.transpose(1, 2)) * scale).softmax(dim=-1)


# ---------------------------------------------------------------- P2/P3  EDLoRA_AttnProcessor  edlora.py:108-173
def edlora_cross_attention(hidden_states, encoder_hidden_states, cross_attention_idx, wq, wk, wv, wo, bo, heads,
                           lora=None, return_probs=False):
    """hidden_states [B,N,C]; encoder_hidden_states [B,16,77,768] (layer-wise, edlora.py:129-131) or [B,77,768].

    lora: optional dict name -> (down [r,in], up [out,r], alpha) for 'to_q','to_k','to_v','to_out.0'.
    """
    ehs = encoder_hidden_states
    if ehs is None:
        ehs = hidden_states
    elif ehs.ndim == 4:
        ehs = ehs[:, cross_attention_idx, ...]

    def proj(x, w, b, name):
        if lora is not None and name in lora:
            d, u, a = lora[name]
            return lora_linear(x, w, b, d, u, a)
        return F.linear(x, w, b)

    q = head_to_batch(proj(hidden_states, wq, None, 'to_q'), heads)
    k = head_to_batch(proj(ehs, wk, None, 'to_k'), heads)
    v = head_to_batch(proj(ehs, wv, None, 'to_v'), heads)
    scale = (wq.shape[0] // heads) ** -0.5
    probs = attention_probs(q, k, scale)
    out = batch_to_head(torch.bmm(probs, v), heads)
    out = proj(out, wo, bo, 'to_out.0')
    return (out, probs) if return_probs else out


# ---------------------------------------------------------------- P4  installer ordering  edlora.py:176-190
def cross_attention_layer_order(unet):
    """Names of the attn2 modules in the order the reference numbers them (0..15):
    DFS over named_children of down_blocks, then mid_block, then up_blocks (edlora.py:176-190)."""
    order = []

    def walk(mod, prefix):
        for name, child in mod.named_children():
            full = f'{prefix}.{name}' if prefix else name
            if child.__class__.__name__ == 'Attention' and 'attn2' in name:
                order.append(full)
            else:
                walk(child, full)

    walk(unet.down_blocks, 'down_blocks')
    walk(unet.mid_block, 'mid_block')
    walk(unet.up_blocks, 'up_blocks')
    return order


# ---------------------------------------------------------------- P5  bind_concept_prompt  pipeline_edlora.py:18-29
def bind_concept_prompt(prompts, new_concept_cfg):
    if isinstance(prompts, str):
        prompts = [prompts]
    out = []
    for prompt in prompts:
        layer_prompts = [prompt for _ in range(16)]
        for concept_name, cfg in new_concept_cfg.items():
            names = cfg['concept_token_names']
            layer_prompts = [p.replace(concept_name, names[i]) for i, p in enumerate(layer_prompts)]
        out.extend(layer_prompts)
    return out


# ---------------------------------------------------------------- P6  CFG combine  pipeline_edlora.py:285-287
def cfg_combine(noise_pred, guidance_scale):
    u, c = noise_pred.chunk(2)
    return u + guidance_scale * (c - u)


# ---------------------------------------------------------------- R1  region mask / rewrite  regional :32-86
def region_box_indices(box, feat_height, feat_width):
    """ceil/floor in Python float64 exactly as pipeline_regionally_t2iadapter.py:37-39,67-68."""
    sh, sw, eh, ew = box
    return (math.ceil(sh * feat_height), math.ceil(sw * feat_width), math.floor(eh * feat_height),
            math.floor(ew * feat_width))


def region_feat_size(height, width, seq_len):
    """pipeline_regionally_t2iadapter.py:43-48."""
    downscale = math.sqrt(height * width / seq_len)
    return int(height // downscale), int(width // downscale)


def region_count_mask(boxes, feat_height, feat_width):
    """get_region_mask (regional :34-41): number of regions covering each feature pixel (int32)."""
    mask = np.zeros((feat_height, feat_width), dtype=np.int32)
    for box in boxes:
        sh, sw, eh, ew = region_box_indices(box, feat_height, feat_width)
        mask[sh:eh, sw:ew] += 1
    return mask


def region_rewrite(global_out, query, region_kv, boxes, height, width, scale):
    """global_out, query: [B*heads, N, d]; region_kv: list of (key [B*heads,77,d], value); boxes: fractions.
    Returns the rewritten hidden state (regional :32-86 with replace_ratio = 1.0)."""
    n = query.shape[1]
    fh, fw = region_feat_size(height, width, n)
    count = torch.from_numpy(region_count_mask(boxes, fh, fw)).to(query.device)   # (device placement only)
    q = query.reshape(query.shape[0], fh, fw, -1)
    out = global_out.reshape(global_out.shape[0], fh, fw, -1).clone()
    out[:, count != 0, :] = 0
    for (rk, rv), box in zip(region_kv, boxes):
        sh, sw, eh, ew = region_box_indices(box, fh, fw)
        qq = q[:, sh:eh, sw:ew, :]
        p = (torch.einsum('bhwc,bnc->bhwn', qq, rk) * scale).softmax(dim=-1)
        o = torch.einsum('bhwn,bnc->bhwc', p, rv)
        out[:, sh:eh, sw:ew, :] += o / count[sh:eh, sw:ew].reshape(1, eh - sh, ew - sw, 1).to(o.dtype)
    return out.reshape(global_out.shape[0], n, -1)


# ---------------------------------------------------------------- G3  merge  convert_edlora_to_diffusers.py:67-73
def merge_lora_weight(weight, down, up, alpha):
    """W' = W + alpha * up @ down (4-D 1x1 conv weights squeezed)   gradient_fusion.py:133-140."""
    if weight.ndim == 4:
        return weight + alpha * (up.squeeze() @ down.squeeze()).unsqueeze(-1).unsqueeze(-1)
    return weight + alpha * (up @ down)


# ---------------------------------------------------------------- G1  update_quasi_newton  gradient_fusion.py:38-96
def update_quasi_newton(K_target, V_target, W, iters):
    """min_W mean((K W^T - V)^2) by one torch.optim.LBFGS.step (lr 1, max_iter=iters, history 25, strong Wolfe,
    tolerances 1e-16), returning the best W over all closure evaluations (gradient_fusion.py:62-85)."""
    W = W.detach().clone().requires_grad_(True)
    K_target, V_target = K_target.detach(), V_target.detach()
    best = {'loss': float('inf'), 'W': None}
    opt = torch.optim.LBFGS([W], lr=1, max_iter=iters, history_size=25, line_search_fn='strong_wolfe',
                            tolerance_grad=1e-16, tolerance_change=1e-16)

    def closure():
        opt.zero_grad()
        loss = F.mse_loss(F.linear(K_target, W), V_target)
        if loss < best['loss']:
            best['loss'] = loss.item()
            best['W'] = W.detach().clone()
        loss.backward()
        return loss

    opt.step(closure)
    return best['W']
