"""BASELINE config 4 (regionally_controlable_sampling: 768 x 1536, 3 regions, 4 adapter maps) at FULL size, and
`RegionallyT2IAdapterPipeline.__call__` end to end, against the fp32 oracle.

The oracle (oracle/unet.py skeleton + oracle/inject.py RegionProcessor = restatement of RegionT2I_AttnProcessor, pinned to
the reference's own processor by tests/golden) is plain PyTorch and runs here in fp32 ON THE GPU (TF32 off, the
attention through the MATH backend) so that the 11 TFLOP of one full-size CFG step take a second instead of a minute.

Tolerances (rel-L2 = ||a-b|| / ||b||): eps <= 5e-3 (fp16 operands, see tests/test_unet_gpu.py);
post-scheduler latents at guidance 7.5 <= 1e-3 (BASELINE.json); multi-step pipeline latents <= 5e-3 per 3 steps.
"""
import math

import pytest
import torch

gpu = pytest.mark.gpu

# regionally_sample.sh:66-74 scaled x0.75 to 768 x 1536: pixel [h0, w0, h1, w1]
BOXES_PX = [[3, 5, 768, 368], [11, 368, 768, 690], [2, 977, 768, 1494]]
BOXES_PX_OVERLAP = [[3, 5, 768, 368], [11, 330, 768, 690], [2, 977, 768, 1494]]


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _fractions(boxes, height, width):
    return [(b[0] / height, b[1] / width, b[2] / height, b[3] / width) for b in boxes]


def _math_sdpa():
    from torch.nn.attention import SDPBackend, sdpa_kernel
    return sdpa_kernel(SDPBackend.MATH)


def _b200_unet(ref, cfg):
    from mixofshow.models.unet_b200 import UNet2DConditionModel
    kw = dict(block_out_channels=cfg['block_out_channels'], layers_per_block=cfg['layers_per_block']) if cfg else {}
    u = UNet2DConditionModel(**kw)
    u.load_state_dict(ref.state_dict())
    return u


@gpu
@pytest.mark.parametrize('tag', ['abut', 'overlap'])
def test_config4_full_size_step(cuda, tag):
    """One CFG denoise step of the FULL SD1.5 topology at latent 96 x 192 (N = 18432 / 4608 / 1152 / 288 tokens), context +
    3 region embeddings [2,16,77,768], boxes of regionally_sample.sh, 4 adapter residual maps, then CFG 7.5 +
    DPM-Solver++ (first step of the 30-step schedule)."""
    from mixofshow.pipelines.pipeline_regionally_t2iadapter import revise_regionally_t2iadapter_attention_forward
    from mos_b200 import ops
    from mos_b200.scheduler import DPMSolverPP2M
    from oracle import edlora_ref as er
    from oracle import inject
    from oracle import unet as ou
    from oracle.schedulers import DPMSolverMultistepScheduler
    height, width = 768, 1536
    h, w = height // 8, width // 8
    ref = ou.build_unet(0, None)
    inject.install_region_processors(ref)
    unet = _b200_unet(ref, None)
    revise_regionally_t2iadapter_attention_forward(unet)
    g = lambda s: torch.Generator().manual_seed(s)
    lat1 = torch.randn(1, 4, h, w, generator=g(14))
    ehs = torch.randn(2, 16, 77, 768, generator=g(20))
    regs = [torch.randn(2, 16, 77, 768, generator=g(21 + i)) for i in range(3)]
    boxes = _fractions(BOXES_PX if tag == 'abut' else BOXES_PX_OVERLAP, height, width)
    ad = [torch.randn(2, c, h // d, w // d, generator=g(30 + i)) * 0.1
          for i, (c, d) in enumerate([(320, 1), (640, 2), (1280, 4), (1280, 8)])]
    sched = DPMSolverMultistepScheduler()
    sched.set_timesteps(30)
    t0 = int(sched.timesteps[0])
    lat2 = torch.cat([lat1, lat1])
    kw_ref = {'region_list': [(r.cuda(), b) for r, b in zip(regs, boxes)], 'height': height, 'width': width}
    ref = ref.cuda()
    with torch.no_grad(), _math_sdpa():
        eps_ref = ref(lat2.cuda(), torch.tensor([t0, t0]).cuda(), ehs.cuda(), cross_attention_kwargs=kw_ref,
                      down_block_additional_residuals=[a.cuda() for a in ad]).sample.cpu()
    del ref
    torch.cuda.empty_cache()
    prev_ref = sched.step(er.cfg_combine(eps_ref, 7.5), t0, lat1).prev_sample
    eps = unet(lat2.cuda(), torch.tensor([float(t0)] * 2).cuda(), ehs.cuda(), cross_attention_kwargs=kw_ref,
               down_block_additional_residuals=[a.cuda() for a in ad]).sample
    s2 = DPMSolverPP2M()
    s2.set_timesteps(30)
    latents = lat1.cuda().clone()
    ops.cfg_dpmpp_step(eps.float().contiguous(), latents, torch.zeros_like(latents), None, cfg=True, guidance=7.5,
                       coef=s2.coefficients(0))
    torch.cuda.synchronize()
    e_eps, e_lat = rel_l2(eps, eps_ref), rel_l2(latents, prev_ref)
    print(f'config 4 [{tag}] 768x1536, 3 regions + adapters: eps rel-L2 {e_eps:.3e}, latents (CFG 7.5) rel-L2 {e_lat:.3e}')
    assert e_eps < 5e-3 and e_lat < 1e-3


def _oracle_spatial_weight(feat, base, spec, height, width):
    """pipeline_regionally_t2iadapter.py:490-510 restated (the reference `eval`s the two halves of 'region-weight')."""
    fh, fw = feat.shape[2:]
    wmap = base * torch.ones(fh, fw)
    if spec != '':
        for item in spec.split('|'):
            region, weight = item.split('-')
            sh, sw, eh, ew = [float(v) for v in region.strip('[]').split(',')]
            a, b = math.ceil(sh / height * fh), math.ceil(sw / width * fw)
            c, d = math.floor(eh / height * fh), math.floor(ew / width * fw)
            wmap[a:c, b:d] = float(weight)
    return wmap * feat


@gpu
def test_regional_pipeline_call_vs_oracle_loop(cuda):
    """RegionallyT2IAdapterPipeline.__call__ (reference :303-599; loop :548-580) with `region_list`, key-pose AND sketch
    adapter states and a `region_sketch_adaptor_weight` override string, 3 DPM-Solver++ steps at guidance 7.5, against
    the oracle loop (oracle UNet with the restated RegionProcessor, reference adapter mixing :484-546)."""
    from mixofshow.pipelines.pipeline_regionally_t2iadapter import RegionallyT2IAdapterPipeline
    from oracle import edlora_ref as er
    from oracle import inject
    from oracle import unet as ou
    from oracle.schedulers import DPMSolverMultistepScheduler
    height, width = 192, 384
    h, w = height // 8, width // 8
    ref = ou.build_unet(0, ou.TINY)
    inject.install_region_processors(ref)
    unet = _b200_unet(ref, ou.TINY)
    pipe = RegionallyT2IAdapterPipeline(unet=unet).to('cuda')
    pipe.set_new_concept_cfg({})
    g = lambda s: torch.Generator().manual_seed(s)
    lat = torch.randn(1, 4, h, w, generator=g(3))
    ehs = torch.randn(2, 16, 77, 768, generator=g(4))
    boxes = _fractions([[0, 2, 192, 92], [3, 92, 192, 172], [1, 244, 192, 373]], height, width)
    regs = [(torch.randn(2, 16, 77, 768, generator=g(5 + i)), boxes[i]) for i in range(3)]
    chans = [(320, 1), (640, 2)]
    kp = [torch.randn(1, c, h // d, w // d, generator=g(40 + i)) * 0.1 for i, (c, d) in enumerate(chans)]
    sk = [torch.randn(1, c, h // d, w // d, generator=g(50 + i)) * 0.1 for i, (c, d) in enumerate(chans)]
    spec = '[0,16,192,92]-0.3|[8,244,184,373]-1.5'
    steps, gs = 3, 7.5
    seen = []
    res = pipe(prompt_embeds=ehs.cuda(), region_list=[(r.cuda(), b) for r, b in regs], latents=lat.clone(), height=height,
               width=width, num_inference_steps=steps, guidance_scale=gs, output_type='latent',
               keypose_adapter_state=[a.cuda() for a in kp], keypose_adaptor_weight=0.8,
               sketch_adapter_state=[a.cuda() for a in sk], sketch_adaptor_weight=0.5, region_sketch_adaptor_weight=spec,
               callback=lambda i, t, x: seen.append((i, int(t)))).images
    sched = DPMSolverMultistepScheduler()
    sched.set_timesteps(steps)
    assert seen == [(i, int(t)) for i, t in enumerate(sched.timesteps)]
    adapter = [torch.cat([_oracle_spatial_weight(kp[i], 0.8, '', height, width)
                          + _oracle_spatial_weight(sk[i], 0.5, spec, height, width)] * 2) for i in range(len(kp))]
    x = lat.clone()
    kw = {'region_list': regs, 'height': height, 'width': width}
    for t in sched.timesteps:
        with torch.no_grad():
            eps = ref(torch.cat([x, x]), torch.tensor([int(t), int(t)]), ehs, cross_attention_kwargs=kw,
                      down_block_additional_residuals=[a.clone() for a in adapter]).sample
        x = sched.step(er.cfg_combine(eps, gs), int(t), x).prev_sample
    e = rel_l2(res, x)
    print(f'RegionallyT2IAdapterPipeline 3-step loop (regions + 2 adapters + region weight string) vs oracle: rel-L2 {e:.3e}')
    assert e < 5e-3
    # the per-region override must matter (guards against the string being ignored)
    res2 = pipe(prompt_embeds=ehs.cuda(), region_list=[(r.cuda(), b) for r, b in regs], latents=lat.clone(),
                height=height, width=width, num_inference_steps=steps, guidance_scale=gs, output_type='latent',
                keypose_adapter_state=[a.cuda() for a in kp], keypose_adaptor_weight=0.8,
                sketch_adapter_state=[a.cuda() for a in sk], sketch_adaptor_weight=0.5).images
    assert rel_l2(res2, res) > 1e-3


def test_spatial_weight_box_indices_bit_exact():
    """`_spatial_weight` (host logic): the weight map equals the reference's ceil / floor construction exactly."""
    from mixofshow.pipelines.pipeline_regionally_t2iadapter import _spatial_weight
    feat = torch.ones(1, 2, 24, 48)
    spec = '[3,5,768,368]-0.25|[11,368,768,690]-2|[2,977,768,1494]-0'
    got = _spatial_weight(feat, 0.7, spec, 768, 1536)
    exp = _oracle_spatial_weight(feat, 0.7, spec, 768, 1536)
    assert torch.equal(got, exp)
