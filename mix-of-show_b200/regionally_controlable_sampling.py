"""B200 mirror of the reference's `regionally_controlable_sampling.py` entry script (BASELINE config 4): region-string
parsing, model loading from a fused `combined_model_*` directory, and the sampling call.  Host logic only; the UNet loop runs
on `RegionallyT2IAdapterPipeline` (mixofshow/pipelines/pipeline_regionally_t2iadapter.py).

Out of scope here: the T2I-Adapter networks and the VAE.  Conditions are therefore passed as
pre-computed adapter feature maps (`--keypose_adapter_state` / `--sketch_adapter_state file.pt`: the 4 maps a T2IAdapter
returns) or skipped, and the result is
written as latents unless a VAE object is supplied."""
import argparse
import ast
import json
import os

import torch


def prepare_text(prompt, region_prompts, height, width):
    """regionally_controlable_sampling.py:67-94.  region_prompts:
    '[subject1]-*-[negative1]-*-[h0, w0, h1, w1]|[subject2]-*-[negative2]-*-[...]' (pixel boxes; '[]' = whole image) ->
    (prompt, [(region prompt, region negative prompt, [h0/H, w0/W, h1/H, w1/W]), ...]).  The box arithmetic is Python
    float division exactly as in the reference (the fractions feed the bit-exact ceil/floor of the region masks)."""
    region_collection = []
    for region in region_prompts.split('|'):
        if region == '':
            break
        prompt_region, neg_prompt_region, pos = region.split('-*-')
        prompt_region = prompt_region.replace('[', '').replace(']', '')
        neg_prompt_region = neg_prompt_region.replace('[', '').replace(']', '')
        pos = list(ast.literal_eval(pos))          # the reference uses eval(); the strings are list literals
        if len(pos) == 0:
            pos = [0, 0, 1, 1]
        else:
            pos[0], pos[2] = pos[0] / height, pos[2] / height
            pos[1], pos[3] = pos[1] / width, pos[3] / width
        region_collection.append((prompt_region, neg_prompt_region, pos))
    return (prompt, region_collection)


def build_model(pretrained_model, device='cuda', tokenizer=None):
    """reference :55-64: pipeline + new_concept_cfg.json from a fused model directory."""
    from mixofshow.pipelines.pipeline_regionally_t2iadapter import RegionallyT2IAdapterPipeline
    from mixofshow.utils import model_io
    assert os.path.exists(os.path.join(pretrained_model, 'new_concept_cfg.json'))
    unet = model_io.load_unet(pretrained_model)
    text_encoder = model_io.load_text_encoder(pretrained_model, device=device)
    if tokenizer is None:
        from transformers import CLIPTokenizer
        tokenizer = CLIPTokenizer.from_pretrained(pretrained_model, subfolder='tokenizer')
    new_concept_cfg = model_io.load_new_concept_cfg(pretrained_model)
    model_io.ensure_concept_tokens(tokenizer, new_concept_cfg)      # the fused model's added `<new{k}>` tokens
    pipe = RegionallyT2IAdapterPipeline(text_encoder=text_encoder, tokenizer=tokenizer, unet=unet).to(device)
    pipe.set_new_concept_cfg(new_concept_cfg)
    return pipe


def sample_image(pipe, input_prompt, input_neg_prompt=None, generator=None, num_inference_steps=50, guidance_scale=7.5,
                 **extra_kargs):
    """reference :14-52 (adapter states / weights travel in extra_kargs: keypose_adapter_state=..., sketch_adaptor_weight=...)."""
    return pipe(prompt=input_prompt, negative_prompt=input_neg_prompt, generator=generator, guidance_scale=guidance_scale,
                num_inference_steps=num_inference_steps, **extra_kargs).images


def parse_args(argv=None):
    parser = argparse.ArgumentParser('', add_help=False)
    parser.add_argument('--pretrained_model', required=True, type=str)
    parser.add_argument('--sketch_adapter_state', default=None, type=str, help='torch file: 4 pre-computed sketch adapter maps')
    parser.add_argument('--sketch_adaptor_weight', default=1.0, type=float)
    parser.add_argument('--region_sketch_adaptor_weight', default='', type=str)
    parser.add_argument('--keypose_adapter_state', default=None, type=str, help='torch file: 4 pre-computed keypose adapter maps')
    parser.add_argument('--keypose_adaptor_weight', default=1.0, type=float)
    parser.add_argument('--region_keypose_adaptor_weight', default='', type=str)
    parser.add_argument('--height', default=768, type=int)
    parser.add_argument('--width', default=1536, type=int)
    parser.add_argument('--save_dir', default=None, type=str)
    parser.add_argument('--prompt', default='photo of a toy', type=str)
    parser.add_argument('--negative_prompt', default='', type=str)
    parser.add_argument('--prompt_rewrite', default='', type=str)
    parser.add_argument('--seed', default=16141, type=int)
    parser.add_argument('--suffix', default='', type=str)
    parser.add_argument('--num_inference_steps', default=50, type=int)       # the reference samples with 50 steps (:38)
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    device = torch.device('cuda')
    pipe = build_model(args.pretrained_model, device)
    kwargs = {'height': args.height, 'width': args.width, 'output_type': 'latent'}
    for kind in ('sketch', 'keypose'):
        path = getattr(args, f'{kind}_adapter_state')
        if path is not None:
            kwargs[f'{kind}_adapter_state'] = torch.load(path)
        kwargs[f'{kind}_adaptor_weight'] = getattr(args, f'{kind}_adaptor_weight')
        kwargs[f'region_{kind}_adaptor_weight'] = getattr(args, f'region_{kind}_adaptor_weight')
    input_prompt = [prepare_text(args.prompt, args.prompt_rewrite, args.height, args.width)]
    latents = sample_image(pipe, input_prompt=input_prompt, input_neg_prompt=[args.negative_prompt],
                           generator=torch.Generator('cpu').manual_seed(args.seed),
                           num_inference_steps=args.num_inference_steps, **kwargs)
    if args.save_dir is not None:
        os.makedirs(args.save_dir, exist_ok=True)
        out = os.path.join(args.save_dir, f'latents---{args.seed}{"---" + args.suffix if args.suffix else ""}.pt')
        torch.save({'latents': latents.cpu(), 'config': vars(args)}, out)
        with open(os.path.join(args.save_dir, 'config.json'), 'w') as f:
            json.dump(vars(args), f)
        print(f'save to: {out}')
    return latents


if __name__ == '__main__':
    main()
