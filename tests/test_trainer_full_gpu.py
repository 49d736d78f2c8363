"""The full ED-LoRA training step through the reference-shaped `EDLoRATrainer` (mixofshow/pipelines/trainer_edlora.py):
prompts -> bind_concept_prompt -> tokenizer -> CLIP text encoder (LoRA) -> UNet (LoRA) -> masked MSE + attention regulariser
-> backward through BOTH networks, against fp32 autograd through transformers' CLIPTextModel chained into the oracle UNet
(reference LoRA formula injected in both, oracle/train_ref.py loss).  Checks the three parameter groups of
trainer_edlora.py:82-139: new-concept embedding rows, CLIPAttention LoRA, UNet Attention LoRA.

Tolerances: bf16 operands through two networks forward and backward: whole-group gradient rel-L2 <= 4e-2, cosine >= 0.998;
loss within 2 %."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.flatten().float().cpu(), b.flatten().float().cpu(), dim=0).item()


FINETUNE = {'text_embedding': {'enable_tuning': True, 'lr': 1e-3},
            'text_encoder': {'enable_tuning': True, 'lora_cfg': {'rank': 4, 'alpha': 1.0, 'where': 'CLIPAttention'}, 'lr': 1e-5},
            'unet': {'enable_tuning': True, 'lora_cfg': {'rank': 4, 'alpha': 1.0, 'where': 'Attention'}, 'lr': 1e-4}}


def _base_dir(tmp_path, clip_layers=2):
    from transformers import CLIPTextConfig, CLIPTextModel
    from mixofshow.models.unet_b200 import UNet2DConditionModel
    from mixofshow.utils import model_io
    from oracle import unet as ou
    torch.manual_seed(0)
    ref_unet = ou.build_unet(0, ou.TINY)
    unet = UNet2DConditionModel(block_out_channels=ou.TINY['block_out_channels'], layers_per_block=ou.TINY['layers_per_block'])
    unet.load_state_dict(ref_unet.state_dict())
    base = str(tmp_path / 'base')
    model_io.save_unet(unet, base)
    clip = CLIPTextModel(CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072,
                                        num_hidden_layers=clip_layers, num_attention_heads=12,
                                        max_position_embeddings=77)).eval()
    clip.save_pretrained(os.path.join(base, 'text_encoder'))
    return base, ref_unet, clip


def test_full_trainer_step_vs_autograd(cuda, tmp_path):
    from test_fusion_orchestration import WordTokenizer
    from mixofshow.pipelines.pipeline_edlora import bind_concept_prompt
    from mixofshow.pipelines.trainer_edlora import EDLoRATrainer
    from mixofshow.utils.ptp_util import AttentionStore
    from oracle import inject, train_ref
    from oracle.schedulers import DDPMScheduler
    base, ref_unet, clip = _base_dir(tmp_path)
    tok = WordTokenizer()
    reg_w = 0.05
    tr = EDLoRATrainer(base, '<c1>+<c2>', '<rand-0.02>+<rand-0.02>', True, finetune_cfg=json.loads(json.dumps(FINETUNE)),
                       noise_offset=None, attn_reg_weight=reg_w, reg_full_identity=False, use_mask_loss=True,
                       enable_xformers=True, tokenizer=tok, latent_size=(16, 16))
    assert tr.new_concept_cfg['<c2>']['concept_token_names'] == [f'<new{16 + i}>' for i in range(16)]
    ids_concept = tr.get_all_concept_token_ids()
    assert ids_concept == list(range(49408, 49408 + 32))
    g = torch.Generator().manual_seed(5)
    delta = {'new_concept_embedding': {'<c1>': torch.randn(16, 768, generator=g) * 0.02,
                                       '<c2>': torch.randn(16, 768, generator=g) * 0.02},
             'text_encoder': inject.random_lora_state(clip, seed=3, where='CLIPAttention', up_std=0.05),
             'unet': inject.random_lora_state(ref_unet, seed=10)}
    tr.load_delta_state_dict(delta)
    B, H = 2, 16
    prompts = ['photo of a <c1> <c2>', 'the <c1> <c2> on a beach']
    lat = torch.randn(B, 4, H, H, generator=g)
    noise = torch.randn(B, 4, H, H, generator=g)
    t = torch.tensor([130, 811])
    masks = (torch.rand(B, 1, H, H, generator=g) > 0.5).float()
    masks[:, :, 4:9, 4:9] = 1.0
    masks[:, :, 0, 0] = 0.0
    loss = tr(lat, prompts, masks, torch.ones_like(masks), noise=noise, timesteps=t)
    torch.cuda.synchronize()
    # ---------------- reference chain in fp32 autograd
    clip.resize_token_embeddings(49408 + 32)
    emb = clip.get_input_embeddings().weight
    with torch.no_grad():
        emb[49408:49408 + 16] = delta['new_concept_embedding']['<c1>']
        emb[49408 + 16:49408 + 32] = delta['new_concept_embedding']['<c2>']
    emb.requires_grad_(True)
    t_leaves = {k: v.clone().requires_grad_(True) for k, v in delta['text_encoder'].items()}
    u_leaves = {k: v.clone().requires_grad_(True) for k, v in delta['unet'].items()}
    inject.inject_lora(clip, t_leaves, 1.0)
    inject.inject_lora(ref_unet, u_leaves, 1.0)
    ctl = AttentionStore(training=True)
    n_x = inject.install_control_processors(ref_unet, ctl)
    ids = tok(bind_concept_prompt(prompts, tr.new_concept_cfg), padding='max_length', max_length=77,
              return_tensors='pt').input_ids
    ehs = clip(ids)[0].view(B, 16, 77, 768)
    pos = train_ref.concept_token_positions(ids, B, ids_concept)
    noisy = DDPMScheduler().add_noise(lat, noise, t)
    loss_ref, _, _ = train_ref.train_loss(ref_unet, ctl, noisy, t, ehs, noise, masks, masks, pos, reg_full_identity=False,
                                          attn_reg_weight=reg_w)
    loss_ref.backward()
    print(f'full trainer step: loss {loss.item():.6f} vs autograd {loss_ref.item():.6f}  ({n_x} cross-attention layers)')
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    # group 0: embedding rows
    g_emb_ref = emb.grad[49408:49408 + 32]
    e0, c0 = rel_l2(tr.text_engine.emb_grad, g_emb_ref), _cos(tr.text_engine.emb_grad, g_emb_ref)
    # group 1: text LoRA, group 2: UNet LoRA
    res = {}
    for name, eng, leaves in (('text', tr.text_engine, t_leaves), ('unet', tr.engine, u_leaves)):
        fg, fr = [], []
        for m, (gd, gu) in eng.lora_grad_dict().items():
            fg += [gd.flatten().cpu(), gu.flatten().cpu()]
            fr += [leaves[m + '.lora_down.weight'].grad.reshape(gd.shape).flatten(),
                   leaves[m + '.lora_up.weight'].grad.reshape(gu.shape).flatten()]
        fg, fr = torch.cat(fg), torch.cat(fr)
        res[name] = (rel_l2(fg, fr), _cos(fg, fr), fg.numel())
    print(f'  embedding rows: rel-L2 {e0:.3e} cos {c0:.5f};  text LoRA ({res["text"][2]}): rel-L2 {res["text"][0]:.3e} cos '
          f'{res["text"][1]:.5f};  unet LoRA ({res["unet"][2]}): rel-L2 {res["unet"][0]:.3e} cos {res["unet"][1]:.5f}')
    assert e0 < 4e-2 and c0 > 0.998
    for name in ('text', 'unet'):
        assert res[name][0] < 4e-2 and res[name][1] > 0.998
    # checkpoint layout of the reference (trainer_edlora.py:358-378) and round trip
    d = tr.delta_state_dict()
    assert set(d) == {'new_concept_embedding', 'text_encoder', 'unet'} and set(d['new_concept_embedding']) == {'<c1>', '<c2>'}
    assert sorted(d['text_encoder']) == sorted(delta['text_encoder']) and sorted(d['unet']) == sorted(delta['unet'])
    for k in delta['text_encoder']:
        assert rel_l2(d['text_encoder'][k], delta['text_encoder'][k]) < 1e-6
    assert rel_l2(d['new_concept_embedding']['<c2>'], delta['new_concept_embedding']['<c2>']) < 1e-6


def test_train_loop_three_groups(cuda, tmp_path):
    """train() (train_edlora.py:105-158 mirror) with the full trainer: all three groups move, the learning rates decay
    linearly, the loss on a repeated batch goes down, the embedding rows freeze once Norm_mean crosses the threshold."""
    import train_edlora as te
    from test_fusion_orchestration import WordTokenizer
    from mixofshow.pipelines.trainer_edlora import EDLoRATrainer
    base, _, _ = _base_dir(tmp_path, clip_layers=1)
    tr = EDLoRATrainer(base, '<c1>+<c2>', '<rand-0.013>+<rand-0.013>', True, finetune_cfg=json.loads(json.dumps(FINETUNE)),
                       attn_reg_weight=0.01, reg_full_identity=False, tokenizer=WordTokenizer(), latent_size=(16, 16))
    g = torch.Generator().manual_seed(1)
    m = torch.zeros(2, 1, 16, 16)
    m[:, :, 3:12, 4:13] = 1
    batch = {'images': torch.randn(2, 4, 16, 16, generator=g), 'prompts': ['photo of a <c1> <c2>', 'a <c1> <c2> smiling'],
             'masks': m, 'img_masks': torch.ones(2, 1, 16, 16)}
    logs = []
    losses = te.train(tr, [batch] * 12, dataset_len=24, batch_size_per_gpu=2, print_freq=1, log=logs.append,
                      emb_norm_threshold=0.41)
    assert len(losses) == 12
    d = tr.delta_state_dict()
    assert any(v.abs().max().item() > 0 for k, v in d['unet'].items() if k.endswith('lora_up.weight'))
    assert any(v.abs().max().item() > 0 for k, v in d['text_encoder'].items() if k.endswith('lora_up.weight'))
    norms = [float(l.split('Norm_mean ')[1]) for l in logs]
    print('    losses', ' '.join(f'{x:.4f}' for x in losses), '| Norm_mean', ' '.join(f'{x:.4f}' for x in norms))
    assert norms[0] != norms[1]                                   # the rows train (lr 1e-3) ...
    crossed = [i for i, n in enumerate(norms) if n >= 0.41]
    if crossed:                                                   # ... and freeze for good after crossing the threshold
        assert all(abs(n - norms[crossed[0]]) < 1e-6 for n in norms[crossed[0]:])
