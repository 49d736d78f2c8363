"""Drop-in surface on the GPU: the B200 `mixofshow` package, driven exactly like the reference drives its own modules,
against the golden vectors produced by the reference's modules (tests/golden/reference_golden.pt).

Tolerances: golden = fp32 reference; B200 path = bf16 operands / fp32 accumulation -> rel-L2 <= 1.5e-2 on layer
outputs (two chained bf16 GEMMs + attention), <= 2e-2 on a whole UNet; integer quantities bit exact.
"""
import os

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_golden.pt')


def _f32(obj):
    if torch.is_tensor(obj):
        return obj.float() if obj.dtype == torch.bfloat16 else obj
    if isinstance(obj, dict):
        return {k: _f32(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_f32(v) for v in obj)
    return obj


@pytest.fixture(scope='module')
def G():
    return _f32(torch.load(GOLD, weights_only=False))


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def test_lora_linear_layer_linear_and_conv(cuda):
    """LoRALinearLayer (edlora.py:221-246) with the reference's constructor and patched-forward protocol."""
    from mixofshow.models.edlora import LoRALinearLayer
    from oracle import edlora_ref as er
    torch.manual_seed(0)
    lin = nn.Linear(640, 320).cuda()
    layer = LoRALinearLayer('x.to_q', lin, rank=4, alpha=0.7).cuda()
    assert layer.lora_up.weight.abs().max().item() == 0            # zeros init as the reference
    layer.lora_up.weight.data.normal_(0, 0.1)
    x = torch.randn(2, 77, 640, device=cuda)
    y = lin(x)                                                      # patched forward -> fused CUDA kernel
    ref = er.lora_linear(x.cpu(), lin.weight.cpu(), lin.bias.cpu(), layer.lora_down.weight.cpu(),
                         layer.lora_up.weight.cpu(), 0.7)
    assert y.shape == (2, 77, 320) and y.dtype == x.dtype
    assert rel_l2(y, ref) < 6e-3
    conv = nn.Conv2d(320, 320, 1).cuda()
    lc = LoRALinearLayer('x.proj_in', conv, rank=4, alpha=1.3).cuda()
    lc.lora_up.weight.data.normal_(0, 0.1)
    xc = torch.randn(2, 320, 8, 8, device=cuda)
    yc = conv(xc)
    refc = er.lora_linear(xc.cpu(), conv.weight.cpu(), conv.bias.cpu(), lc.lora_down.weight.cpu(),
                          lc.lora_up.weight.cpu(), 1.3)
    assert rel_l2(yc, refc) < 6e-3
    with pytest.raises(ValueError):
        LoRALinearLayer('bad', nn.Linear(64, 64), rank=8)


def _attention(state, cross):
    from mixofshow.models.unet_b200 import Attention
    a = Attention(320, cross, heads=8, dim_head=40)
    a.load_state_dict(state)
    return a.cuda()


def test_edlora_processors_vs_reference_golden(cuda, G):
    from mixofshow.models.edlora import EDLoRA_AttnProcessor, EDLoRA_Control_AttnProcessor
    g = G['attn_proc']
    attn = _attention(g['state'], 128)
    hs, ehs = g['hs'].cuda(), g['ehs'].cuda()
    out = EDLoRA_AttnProcessor(g['idx'])(attn, hs, encoder_hidden_states=ehs)
    assert rel_l2(out, g['out']) < 1.5e-2
    seen = {}

    class Ctl:
        def __call__(self, probs, is_cross, place):
            seen.update(probs=probs, is_cross=is_cross, place=place)
            return probs
    out2 = EDLoRA_Control_AttnProcessor(g['idx'], 'down', Ctl())(attn, hs, encoder_hidden_states=ehs)
    assert rel_l2(out2, g['out_ctl']) < 1.5e-2
    assert seen['is_cross'] is True and seen['place'] == 'down'
    assert tuple(seen['probs'].shape) == tuple(g['probs'].shape)            # [B*heads, N, 77]
    assert rel_l2(seen['probs'], g['probs']) < 1.5e-2
    assert (seen['probs'].sum(-1) - 1).abs().max().item() < 1e-4


@pytest.mark.parametrize('tag', ['abut', 'overlap'])
def test_region_processor_vs_reference_golden(cuda, G, tag):
    from mixofshow.pipelines.pipeline_regionally_t2iadapter import RegionT2I_AttnProcessor
    g = G['region']
    attn = _attention(g['state'], 128)
    boxes = g['boxes'] if tag == 'abut' else g['boxes_overlap']
    rl = [(g['region_embs'][i].cuda(), boxes[i]) for i in range(3)]
    out = RegionT2I_AttnProcessor(g['idx'])(attn, g['hs'].cuda(), encoder_hidden_states=g['ehs'].cuda(),
                                            region_list=rl, height=g['height'], width=g['width'])
    assert rel_l2(out, g['out'][tag]) < 1.5e-2
    attn_s = _attention(g['self_state'], None)
    so = RegionT2I_AttnProcessor(0)(attn_s, g['hs'].cuda(), encoder_hidden_states=None, region_list=[], height=96,
                                    width=192)
    assert rel_l2(so, g['self_out']) < 1.5e-2


def test_region_box_indices_bit_exact(G):
    from mixofshow.pipelines.pipeline_regionally_t2iadapter import region_box_indices, region_feat_size
    g = G['region']
    for (H, W, ds, tag), idx in g['box_index_kat'].items():
        boxes = g['boxes'] if tag == 'abut' else g['boxes_overlap']
        fh, fw = region_feat_size(H, W, (H // ds) * (W // ds))
        assert (fh, fw) == (H // ds, W // ds)
        assert [region_box_indices(b, fh, fw) for b in boxes] == [tuple(i) for i in idx]


def _tiny_b200_unet(seed=0):
    from mixofshow.models.unet_b200 import UNet2DConditionModel
    from oracle import unet as ou
    ref = ou.build_unet(seed, ou.TINY)
    u = UNet2DConditionModel(block_out_channels=ou.TINY['block_out_channels'],
                             layers_per_block=ou.TINY['layers_per_block'])
    u.load_state_dict(ref.state_dict())
    return u


def test_b200_unet_driven_like_the_reference_trainer(cuda, G):
    """The reference's own recipe (trainer_edlora.py:121-133 + pipeline_edlora.py:93) on the B200 UNet container:
    install processors, inject LoRALinearLayer on every Linear under every `Attention`, call unet(...).sample."""
    from mixofshow.models.edlora import LoRALinearLayer, revise_edlora_unet_attention_forward
    from oracle import inject
    g = G['tiny_unet']
    unet = _tiny_b200_unet(g['unet_seed'])
    revise_edlora_unet_attention_forward(unet)
    lora = inject.random_lora_state(unet, seed=g['lora_seed'])
    keep = []
    for name, module in unet.named_modules():                       # trainer_edlora.py:121-133 (where: Attention)
        if module.__class__.__name__ == 'Attention':
            for child_name, child in module.named_modules():
                if child.__class__.__name__ in ('Linear', 'Conv2d'):
                    full = name + '.' + child_name
                    layer = LoRALinearLayer(full, child, rank=4, alpha=1.0)
                    layer.lora_down.weight.data = lora[full + '.lora_down.weight'].clone()
                    layer.lora_up.weight.data = lora[full + '.lora_up.weight'].clone()
                    keep.append(layer)
    assert len(keep) == g['n_lora']
    out = unet(g['latents'].cuda(), torch.tensor([g['t'], g['t']]).cuda(), g['ehs'].cuda()).sample
    e = rel_l2(out, g['out'])
    print(f'B200 UNet (reference recipe) vs reference golden: rel-L2 {e:.3e}')
    assert e < 2e-2
    # a LoRA update (an optimiser step in training) must be picked up by the next call
    with torch.no_grad():
        keep[0].lora_up.weight.mul_(3.0)          # what optimizer.step() does (bumps the tensor version)
    out2 = unet(g['latents'].cuda(), torch.tensor([g['t'], g['t']]).cuda(), g['ehs'].cuda()).sample
    assert not torch.equal(out2, out)


@pytest.mark.parametrize('tag', ['abut', 'overlap'])
def test_b200_unet_regional_with_adapters(cuda, G, tag):
    """RegionallyT2IAdapterPipeline's UNet call (pipeline_regionally_t2iadapter.py:556-566): 3 regions + adapter
    residuals, vs the golden produced by the reference's RegionT2I processors."""
    from mixofshow.pipelines.pipeline_regionally_t2iadapter import revise_regionally_t2iadapter_attention_forward
    g = G['tiny_unet_region']
    unet = _tiny_b200_unet(g['unet_seed'])
    revise_regionally_t2iadapter_attention_forward(unet)
    boxes = g['boxes'] if tag == 'abut' else g['boxes_overlap']
    rl = [(g['region_embs'][i].cuda(), boxes[i]) for i in range(3)]
    kw = {'region_list': rl, 'height': g['height'], 'width': g['width']}
    ad = [a.cuda() for a in g['adapters']]
    t = torch.tensor([g['t'], g['t']]).cuda()
    out = unet(g['latents'].cuda(), t, g['ehs'].cuda(), cross_attention_kwargs=kw,
               down_block_additional_residuals=[a.clone() for a in ad]).sample
    e = rel_l2(out, g['out'][tag])
    print(f'regional B200 UNet [{tag}] vs reference golden: rel-L2 {e:.3e}')
    assert e < 2e-2
    out2 = unet(g['latents'].cuda(), t, g['ehs'].cuda(), cross_attention_kwargs=kw,
                down_block_additional_residuals=[a.clone() for a in ad]).sample      # captured-graph replay
    assert torch.equal(out, out2)


def test_edlora_pipeline_loop(cuda):
    """EDLoRAPipeline.__call__ (pipeline_edlora.py:193-322) with prompt_embeds, 4 steps, vs the oracle loop."""
    from mixofshow.pipelines.pipeline_edlora import EDLoRAPipeline
    from oracle import edlora_ref as er
    from oracle import inject
    from oracle import unet as ou
    from oracle.schedulers import DPMSolverMultistepScheduler
    ref = ou.build_unet(0, ou.TINY)
    inject.install_edlora_processors(ref)
    unet = _tiny_b200_unet(0)
    pipe = EDLoRAPipeline(unet=unet).to('cuda')
    pipe.set_new_concept_cfg({})
    lat = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    pe = torch.randn(1, 16, 77, 768, generator=torch.Generator().manual_seed(4)).to(torch.bfloat16).float()
    ne = torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16).float()
    steps, gs = 4, 3.0
    res = pipe(prompt_embeds=pe.cuda(), negative_prompt_embeds=ne.cuda(), latents=lat.clone(), height=128, width=128,
               num_inference_steps=steps, guidance_scale=gs, output_type='latent').images
    sched = DPMSolverMultistepScheduler()
    sched.set_timesteps(steps)
    x = lat.clone()
    emb = torch.cat([ne.view(1, 1, 77, 768).repeat(1, 16, 1, 1), pe])
    for t in sched.timesteps:
        with torch.no_grad():
            eps = ref(torch.cat([x, x]), torch.tensor([int(t), int(t)]), emb).sample
        x = sched.step(er.cfg_combine(eps, gs), int(t), x).prev_sample
    e = rel_l2(res, x)
    print(f'EDLoRAPipeline 4-step loop vs oracle: latents rel-L2 {e:.3e}')
    assert e < 2e-2


def test_attention_store_controller_whole_unet(cuda):
    """revise_edlora_unet_attention_controller_forward + AttentionStore(training=True) (trainer_edlora.py:96-101 recipe):
    after one UNet call the store holds one probability map per cross-attention layer, grouped by place, equal to the
    fp32 oracle's maps (same controller class driven by the oracle's processors)."""
    from mixofshow.models.edlora import revise_edlora_unet_attention_controller_forward
    from mixofshow.utils.ptp_util import AttentionStore
    from oracle import inject
    from oracle import unet as ou
    ref = ou.build_unet(0, ou.TINY)
    ctl_ref = AttentionStore(training=True)
    n_ref = inject.install_control_processors(ref, ctl_ref)
    unet = _tiny_b200_unet(0)
    ctl = AttentionStore(training=True)
    revise_edlora_unet_attention_controller_forward(unet, ctl)
    assert ctl.num_att_layers == n_ref == ctl_ref.num_att_layers
    g = torch.Generator().manual_seed(11)
    lat = torch.randn(2, 4, 16, 16, generator=g)
    ehs = torch.randn(2, n_ref, 77, 768, generator=g).to(torch.bfloat16).float()
    t = torch.tensor([441, 441])
    with torch.no_grad():
        ref(lat, t, ehs)
    unet(lat.cuda(), t.cuda(), ehs.cuda())
    assert ctl.cur_step == ctl_ref.cur_step == 1
    for place in AttentionStore.PLACES:
        a, b = ctl.attention_store[place], ctl_ref.attention_store[place]
        assert len(a) == len(b)
        for ma, mb in zip(a, b):
            assert tuple(ma.shape) == tuple(mb.shape)
            assert rel_l2(ma, mb) < 2e-2
            assert (ma.sum(-1) - 1).abs().max().item() < 1e-4


def test_control_processor_head_dim_160(cuda):
    """The deepest SD1.5 cross-attention (1280 channels, head_dim 160, 8x8 tokens) with the control processor: the
    probability maps come from the single-tile d=160 attention variant."""
    from mixofshow.models.edlora import EDLoRA_Control_AttnProcessor
    from mixofshow.models.unet_b200 import Attention
    torch.manual_seed(5)
    attn = Attention(1280, 768, heads=8, dim_head=160).cuda()
    hs = torch.randn(2, 64, 1280, device=cuda)
    ehs = torch.randn(2, 16, 77, 768, device=cuda).to(torch.bfloat16).float()
    seen = {}

    class Ctl:
        def __call__(self, probs, is_cross, place):
            seen['probs'] = probs
            return probs
    out = EDLoRA_Control_AttnProcessor(6, 'mid', Ctl())(attn, hs, encoder_hidden_states=ehs)
    e = ehs[:, 6].cpu()
    w = {k: v.detach().cpu() for k, v in attn.state_dict().items()}
    q = (hs.cpu() @ w['to_q.weight'].T).view(2, 64, 8, 160).transpose(1, 2)
    k = (e @ w['to_k.weight'].T).view(2, 77, 8, 160).transpose(1, 2)
    v = (e @ w['to_v.weight'].T).view(2, 77, 8, 160).transpose(1, 2)
    p = (q @ k.transpose(-1, -2) * 160 ** -0.5).softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(2, 64, 1280) @ w['to_out.0.weight'].T + w['to_out.0.bias']
    assert rel_l2(seen['probs'], p.reshape(16, 64, 77)) < 1.5e-2
    assert rel_l2(out, o) < 1.5e-2


def test_pipeline_loop_has_no_host_sync(cuda):
    """The denoise loop of EDLoRAPipeline.__call__ must not synchronise the host with the device (the per-step
    fingerprint walk did 128 device reads): from the first callback to the last, torch's sync debug mode is 'error', which
    raises on any implicit synchronisation (.item(), float(cuda_tensor), blocking copies, ...)."""
    from mixofshow.models.edlora import LoRALinearLayer
    from mixofshow.pipelines.pipeline_edlora import EDLoRAPipeline
    unet = _tiny_b200_unet(0)
    for name, module in list(unet.named_modules()):
        if module.__class__.__name__ == 'Attention':
            for child_name, child in module.named_modules():
                if child.__class__.__name__ == 'Linear':
                    LoRALinearLayer(name + '.' + child_name, child, rank=4, alpha=1.0).lora_up.weight.data.normal_(0, 0.02)
    pipe = EDLoRAPipeline(unet=unet).to('cuda')
    pipe.set_new_concept_cfg({})
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(1, 4, 16, 16, generator=g).cuda()
    pe, ne = torch.randn(1, 16, 77, 768, generator=g).cuda(), torch.randn(1, 77, 768, generator=g).cuda()
    steps = 6
    seen = []

    def cb(i, t, latents):
        seen.append(i)
        if i == 0:
            torch.cuda.set_sync_debug_mode('error')
        if i == steps - 1:
            torch.cuda.set_sync_debug_mode('default')

    kw = dict(prompt_embeds=pe, negative_prompt_embeds=ne, height=128, width=128, num_inference_steps=steps, guidance_scale=7.5,
              output_type='latent')
    pipe(latents=lat.clone(), **kw)                       # first call builds / captures (synchronises, by design)
    try:
        out = pipe(latents=lat.clone(), callback=cb, **kw).images
    finally:
        torch.cuda.set_sync_debug_mode('default')
    assert seen == list(range(steps)) and torch.isfinite(out).all()
