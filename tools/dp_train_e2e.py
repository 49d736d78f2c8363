"""`train_edlora.py -opt <yml>` under torchrun on N GPUs, end to end, on a synthetic model directory (tests/synth.py):
    python tools/dp_train_e2e.py prepare /tmp/dp_e2e            # writes base/, data.pt, train.yml
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        mix-of-show_b200/train_edlora.py -opt /tmp/dp_e2e/train.yml
    python tools/dp_train_e2e.py check /tmp/dp_e2e              # the checkpoint rank 0 saved
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mix-of-show_b200'), os.path.join(ROOT, 'tests')]
import torch  # noqa: E402
import yaml  # noqa: E402

mode, work = sys.argv[1], sys.argv[2]
if mode == 'prepare':
    from synth import make_pretrained_dir
    os.makedirs(work, exist_ok=True)
    base = make_pretrained_dir(os.path.join(work, 'base'))
    g = torch.Generator().manual_seed(1)
    n = 16
    masks = torch.zeros(n, 1, 32, 32)
    masks[:, :, 4:28, 8:24] = 1.0
    torch.save({'latents': torch.randn(n, 4, 32, 32, generator=g) * 0.8, 'prompts': ['photo of a <TOK>'] * n, 'masks': masks},
               os.path.join(work, 'data.pt'))
    finetune = {'text_embedding': {'enable_tuning': True, 'lr': 1e-3},
                'text_encoder': {'enable_tuning': True, 'lora_cfg': {'rank': 4, 'alpha': 1.0, 'where': 'CLIPAttention'}, 'lr': 1e-5},
                'unet': {'enable_tuning': True, 'lora_cfg': {'rank': 4, 'alpha': 1.0, 'where': 'Attention'}, 'lr': 1e-4}}
    opt = {'name': 'dp_e2e', 'manual_seed': 0, 'gradient_accumulation_steps': 1,
           'datasets': {'train': {'path': os.path.join(work, 'data.pt'), 'replace_mapping': {'<TOK>': '<cat1> <cat2>'},
                                  'batch_size_per_gpu': 2, 'dataset_enlarge_ratio': 1}},
           'models': {'pretrained_path': base, 'enable_edlora': True, 'new_concept_token': '<cat1>+<cat2>',
                      'initializer_token': '<rand-0.013>+a', 'finetune_cfg': finetune, 'noise_offset': 0.01, 'attn_reg_weight': 0.01,
                      'reg_full_identity': False, 'use_mask_loss': True, 'latent_size': [32, 32]},
           'train': {'optim_g': {'type': 'AdamW', 'lr': 0.0, 'weight_decay': 0.01, 'betas': [0.9, 0.999]},
                     'emb_norm_threshold': 0.55},
           'path': {'models': os.path.join(work, 'models')}, 'logger': {'print_freq': 1}}
    yaml.safe_dump(opt, open(os.path.join(work, 'train.yml'), 'w'))
    print('prepared', work)
else:
    p = torch.load(os.path.join(work, 'models', 'edlora_model-latest.pth'))['params']
    ups = [v for k, v in p['unet'].items() if k.endswith('lora_up.weight')]
    ok = (list(p['new_concept_embedding']) == ['<cat1>', '<cat2>'] and all(torch.isfinite(v).all() for v in ups)
          and sum(float(v.abs().sum()) for v in ups) > 0 and len(p['text_encoder']) == 16)
    print('checkpoint ok' if ok else 'checkpoint BAD', {k: len(v) for k, v in p.items()})
    sys.exit(0 if ok else 1)
