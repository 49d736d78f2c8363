"""On-disk model formats, CPU only: diffusers-layout directories round-trip through this repo's
containers, and the text-encoder folder is interoperable with the real transformers library in both directions."""
import json
import os

import pytest
import torch

from oracle import unet as ou


def _clip(vocab=320):
    from transformers import CLIPTextConfig, CLIPTextModel
    cfg = CLIPTextConfig(vocab_size=vocab, hidden_size=768, intermediate_size=3072, num_hidden_layers=1,
                         num_attention_heads=12, max_position_embeddings=77)
    torch.manual_seed(0)
    return CLIPTextModel(cfg).eval()


def test_unet_directory_roundtrip_and_config_checks(tmp_path):
    from mixofshow.models.unet_b200 import UNet2DConditionModel
    from mixofshow.utils import model_io
    torch.manual_seed(1)
    unet = UNet2DConditionModel(block_out_channels=ou.TINY['block_out_channels'],
                                layers_per_block=ou.TINY['layers_per_block'])
    model_io.save_unet(unet, str(tmp_path))
    cfg = json.load(open(tmp_path / 'unet' / 'config.json'))
    assert cfg['_class_name'] == 'UNet2DConditionModel' and cfg['block_out_channels'] == [320, 640]
    assert cfg['down_block_types'] == ['CrossAttnDownBlock2D', 'DownBlock2D']
    assert os.path.isfile(tmp_path / 'unet' / 'diffusion_pytorch_model.safetensors')
    back = UNet2DConditionModel.from_pretrained(str(tmp_path), subfolder='unet')      # the diffusers call shape
    a, b = unet.state_dict(), back.state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    # the oracle (diffusers parameter names) loads the same file: the key set is the diffusers one
    from safetensors.torch import load_file
    oracle = ou.UNet2DConditionModel(ou.TINY)
    oracle.load_state_dict(load_file(str(tmp_path / 'unet' / 'diffusion_pytorch_model.safetensors')))
    # unsupported options are rejected, not ignored
    cfg['use_linear_projection'] = True
    json.dump(cfg, open(tmp_path / 'unet' / 'config.json', 'w'))
    with pytest.raises(ValueError, match='use_linear_projection'):
        model_io.load_unet(str(tmp_path))
    cfg['use_linear_projection'] = False
    cfg['up_block_types'] = ['UpBlock2D', 'UpBlock2D']
    json.dump(cfg, open(tmp_path / 'unet' / 'config.json', 'w'))
    with pytest.raises(ValueError, match='block layout'):
        model_io.load_unet(str(tmp_path))


def test_sd15_config_is_accepted():
    from mixofshow.utils import model_io
    model_io.check_unet_config(model_io.UNET_CONFIG_SD15)


def test_text_encoder_folder_interoperates_with_transformers(tmp_path):
    from transformers import CLIPTextModel as HFCLIP
    from mixofshow.utils import model_io
    hf = _clip()
    hf.save_pretrained(str(tmp_path / 'a' / 'text_encoder'))             # the real library writes ...
    mine = model_io.load_text_encoder(str(tmp_path / 'a'), device='cpu')   # ... this repo reads
    sd_hf = {k: v for k, v in hf.state_dict().items() if not k.endswith('position_ids')}
    sd = mine.state_dict()
    assert sd.keys() == sd_hf.keys() and all(torch.equal(sd[k], sd_hf[k]) for k in sd)
    mine.get_input_embeddings().weight.data[7] = 0.25                     # e.g. a learned concept row
    model_io.save_combined_model(str(tmp_path / 'b'), _tiny_unet(), mine, {'<c>': {'concept_token_ids': [1, 2],
                                                                               'concept_token_names': ['<new0>', '<new1>']}})
    back = HFCLIP.from_pretrained(str(tmp_path / 'b' / 'text_encoder'))    # this repo writes, the real library reads
    assert torch.all(back.state_dict()['text_model.embeddings.token_embedding.weight'][7] == 0.25)
    ids = torch.randint(0, 320, (2, 77))
    with torch.no_grad():
        ya = hf(ids)[0]
        hf.get_input_embeddings().weight.data[7] = 0.25
        yb = back(ids)[0]
        yc = hf(ids)[0]
    assert torch.allclose(yb, yc, atol=1e-6) and ya.shape == yb.shape
    assert model_io.load_new_concept_cfg(str(tmp_path / 'b'))['<c>']['concept_token_ids'] == [1, 2]
    assert os.path.isfile(tmp_path / 'b' / 'unet' / 'config.json')


def _tiny_unet():
    from mixofshow.models.unet_b200 import UNet2DConditionModel
    torch.manual_seed(2)
    return UNet2DConditionModel(block_out_channels=ou.TINY['block_out_channels'],
                                layers_per_block=ou.TINY['layers_per_block'])


def test_pipeline_from_pretrained_builds_the_b200_containers(tmp_path):
    """EDLoRAPipeline.from_pretrained on a diffusers-layout directory (construction only: sampling needs the GPU)."""
    from mixofshow.models.clip_b200 import CLIPTextModel
    from mixofshow.models.unet_b200 import UNet2DConditionModel
    from mixofshow.pipelines.pipeline_edlora import EDLoRAPipeline
    from mixofshow.utils import model_io
    base = str(tmp_path)
    model_io.save_unet(_tiny_unet(), base)
    _clip().save_pretrained(os.path.join(base, 'text_encoder'))
    pipe = EDLoRAPipeline.from_pretrained(base, tokenizer=object(), device='cpu')
    assert isinstance(pipe.unet, UNet2DConditionModel) and isinstance(pipe.text_encoder, CLIPTextModel)
    # the reference's installer ran (pipeline_edlora.py:93): every attn2 got its cross_attention_idx in DFS order
    idx = [m.processor.cross_attention_idx for n, m in pipe.unet.named_modules()
           if m.__class__.__name__ == 'Attention' and n.endswith('attn2')]
    assert idx == list(range(len(idx))) and len(idx) == 4
