"""Live cross-check: the oracle against the reference's own modules executed in place (skipped where /root/reference
does not exist, e.g. on the GPU box; the committed golden vectors cover that case)."""
import pytest
import torch

from oracle import edlora_ref as er
from oracle import inject, ref_shims
from oracle import unet as ou

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason='reference checkout not present')


def test_reference_installer_and_lora_match_oracle():
    ed = ref_shims.load_reference_module('mixofshow/models/edlora.py')
    a = ou.build_unet(3, ou.TINY)
    b = ou.build_unet(3, ou.TINY)
    ed.revise_edlora_unet_attention_forward(a)
    inject.install_edlora_processors(b)
    lora = inject.random_lora_state(a, seed=4)
    mods = dict(a.named_modules())
    keep = []
    for k in lora:
        if k.endswith('.lora_down.weight'):
            n = k[:-len('.lora_down.weight')]
            layer = ed.LoRALinearLayer(n, mods[n], rank=4, alpha=0.8)
            layer.lora_down.weight.data = lora[k].clone()
            layer.lora_up.weight.data = lora[n + '.lora_up.weight'].clone()
            keep.append(layer)
    inject.inject_lora(b, lora, 0.8)
    x = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(5))
    ehs = torch.randn(2, 4, 77, 768, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        ya = a(x, torch.tensor([500, 500]), ehs).sample
        yb = b(x, torch.tensor([500, 500]), ehs).sample
    assert ((ya - yb).norm() / ya.norm()).item() < 1e-5


def test_reference_lora_target_selection_matches_trainer_rule():
    """trainer_edlora.py:121-133: every Linear/Conv2d under a module whose class name is `Attention`."""
    u = ou.build_unet(0, ou.TINY)
    names = inject.lora_target_modules(u, 'Attention')
    assert len(names) == 4 * 2 * 4  # 4 transformer blocks x (attn1, attn2) x (to_q,to_k,to_v,to_out.0)
    with torch.device('meta'):
        full = ou.UNet2DConditionModel()
    assert len(inject.lora_target_modules(full, 'Attention')) == 128


def test_reference_bind_and_quasi_newton():
    pe = ref_shims.load_reference_module('mixofshow/pipelines/pipeline_edlora.py')
    gf = ref_shims.load_reference_module('gradient_fusion.py')
    cfg = {'<a>': {'concept_token_names': [f'<n{i}>' for i in range(16)]}}
    assert pe.bind_concept_prompt(['x <a> y', '<a><a>'], cfg) == er.bind_concept_prompt(['x <a> y', '<a><a>'], cfg)
    K = torch.randn(18, 32, generator=torch.Generator().manual_seed(1))
    W0 = torch.randn(24, 32, generator=torch.Generator().manual_seed(2)) * 0.1
    V = K @ (W0 + 0.05 * torch.randn(24, 32, generator=torch.Generator().manual_seed(3))).t()
    a = gf.update_quasi_newton(K, V, W0.clone(), 20, 'cpu')
    b = er.update_quasi_newton(K, V, W0, 20)
    assert ((a - b).norm() / a.norm()).item() < 1e-5
