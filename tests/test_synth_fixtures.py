"""CPU checks of the synthetic fixtures the GPU workflow test relies on (tests/synth.py): the character-level CLIP tokenizer
directory must load with transformers' own CLIPTokenizer and behave like the real vocabulary where the ED-LoRA code depends
on it (special ids, padding, ids of added tokens, save / reload), and the synthetic model directory must load through this
repo's own readers."""
import os

import torch

from synth import make_clip_tokenizer_dir, make_pretrained_dir


def test_synthetic_clip_tokenizer(tmp_path):
    from transformers import CLIPTokenizer
    tok = CLIPTokenizer.from_pretrained(make_clip_tokenizer_dir(str(tmp_path / 'tok')))
    assert len(tok) == 49408 and tok.model_max_length == 77
    ids = tok('photo of a cat', padding='max_length', max_length=77, truncation=True, return_tensors='pt').input_ids[0]
    assert ids[0].item() == 49406 and ids[-1].item() == 49407 and ids.numel() == 77           # BOS ... EOS padding
    assert len(tok.encode('a', add_special_tokens=False)) == 1                                  # a usable initializer token
    assert tok.add_tokens([f'<new{i}>' for i in range(16)]) == 16
    assert [tok.convert_tokens_to_ids(f'<new{i}>') for i in (0, 15)] == [49408, 49423]          # trainer_edlora.py:160-166
    un = tok('photo of a <new3>', truncation=True, max_length=77, padding='do_not_pad').input_ids
    assert un[0] == 49406 and un[-1] == 49407 and un[-2] == 49411 and len(un) < 20
    tok.save_pretrained(str(tmp_path / 'saved'))
    tok2 = CLIPTokenizer.from_pretrained(str(tmp_path / 'saved'))
    assert tok2.convert_tokens_to_ids('<new15>') == 49423 and len(tok2) == 49424


def test_synthetic_pretrained_dir_loads(tmp_path):
    from mixofshow.utils import model_io
    base = make_pretrained_dir(str(tmp_path / 'base'))
    assert sorted(os.listdir(base)) == ['text_encoder', 'tokenizer', 'unet', 'vae']
    unet = model_io.load_unet(base)
    assert tuple(unet.config.block_out_channels) == (320, 640) and unet.config.layers_per_block == 1
    w = unet.state_dict()['down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight']
    assert tuple(w.shape) == (320, 768) and torch.isfinite(w).all()
