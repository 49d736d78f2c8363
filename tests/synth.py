"""Synthetic stand-ins for the artefacts the reference downloads (no network here): a CLIP tokenizer directory that
transformers' own `CLIPTokenizer.from_pretrained` loads, and a diffusers-layout "pretrained model" directory with random
weights.  Test infrastructure only."""
import json
import os

import torch


def _bytes_to_unicode():
    """the GPT-2 / CLIP byte -> printable unicode table (published algorithm)"""
    bs = list(range(ord('!'), ord('~') + 1)) + list(range(ord('¡'), ord('¬') + 1)) + list(range(ord('®'), ord('ÿ') + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def make_clip_tokenizer_dir(path):
    """vocab.json + merges.txt + tokenizer_config.json of a character-level BPE (no merges) with CLIP's layout: 49408
    entries, <|startoftext|> = 49406, <|endoftext|> = 49407 (= pad), model_max_length 77.  Added tokens (`<new0>` ...) get
    ids 49408.. exactly as with the real CLIP vocabulary."""
    os.makedirs(path, exist_ok=True)
    chars = list(_bytes_to_unicode().values())
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + '</w>'] = len(vocab)
    i = 0
    while len(vocab) < 49406:
        vocab[f'filler{i}</w>'] = len(vocab)
        i += 1
    vocab['<|startoftext|>'] = 49406
    vocab['<|endoftext|>'] = 49407
    json.dump(vocab, open(os.path.join(path, 'vocab.json'), 'w'))
    open(os.path.join(path, 'merges.txt'), 'w').write('#version: 0.2\n')
    json.dump({'bos_token': '<|startoftext|>', 'eos_token': '<|endoftext|>', 'unk_token': '<|endoftext|>',
               'pad_token': '<|endoftext|>', 'model_max_length': 77, 'tokenizer_class': 'CLIPTokenizer'},
              open(os.path.join(path, 'tokenizer_config.json'), 'w'))
    return path


def make_pretrained_dir(path, *, clip_layers=2, with_vae=True, seed=0):
    """unet/ (tiny 2-level SD1.5-style UNet), text_encoder/ (CLIP text model, `clip_layers` layers, SD1.5 widths), vae/ (tiny,
    one 2x level) and tokenizer/ in the diffusers layout - written with this repo's own savers and transformers."""
    from mixofshow.models.unet_b200 import UNet2DConditionModel
    from mixofshow.models.vae_b200 import AutoencoderKL
    from mixofshow.utils import model_io
    from transformers import CLIPTextConfig, CLIPTextModel
    torch.manual_seed(seed)
    unet = UNet2DConditionModel(block_out_channels=(320, 640), layers_per_block=1)
    model_io.save_unet(unet, path)
    clip = CLIPTextModel(CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=clip_layers,
                                        num_attention_heads=12, max_position_embeddings=77, hidden_act='quick_gelu')).eval()
    clip.save_pretrained(os.path.join(path, 'text_encoder'))
    if with_vae:
        from oracle import vae as ov                 # only for a state_dict with the right keys / shapes (random init)
        ref = ov.build_vae(seed, ov.TINY_VAE)
        vae = AutoencoderKL({k: v.detach() for k, v in ref.state_dict().items()},
                            block_out_channels=ov.TINY_VAE['block_out_channels'], layers_per_block=1, device='cpu')
        model_io.save_vae(vae, path)
    make_clip_tokenizer_dir(os.path.join(path, 'tokenizer'))
    return path
