"""Drop-in for the `text_encoder` object the reference passes around (transformers `CLIPTextModel`, loaded at
mixofshow/pipelines/trainer_edlora.py:41 and used at pipeline_edlora.py:133-145, trainer_edlora.py:220-234,
gradient_fusion.py:182-199): same call shape — `text_encoder(input_ids)[0]` is the last hidden state — running on
`mos_b200.clip_engine.CLIPTextEngine` (forward only; the training backward of the text encoder is SURVEY.md §8f)."""
from types import SimpleNamespace

import torch

from mos_b200.clip_engine import CLIPTextEngine


class CLIPTextModel:
    def __init__(self, state_dict, *, lora=None, lora_alpha=1.0, merge_lora=False, device='cuda'):
        """state_dict: transformers CLIPTextModel parameters (`text_model.*`).  Engines are built per batch size on
        first use (buffers are static)."""
        self._sd = {k: v.detach() for k, v in state_dict.items()}
        self._kw = dict(lora=lora, lora_alpha=lora_alpha, merge_lora=merge_lora, device=device)
        self._engines = {}
        self.device = torch.device(device)
        self.dtype = torch.float32
        self.config = SimpleNamespace(hidden_size=self._sd['text_model.embeddings.token_embedding.weight'].shape[1],
                                      max_position_embeddings=self._sd['text_model.embeddings.position_embedding.weight'].shape[0])

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def state_dict(self):
        return self._sd

    def __call__(self, input_ids, attention_mask=None, **kw):
        assert attention_mask is None, 'the ED-LoRA pipelines never pass an attention mask to the text encoder'
        n = input_ids.shape[0]
        eng = self._engines.get(n)
        if eng is None:
            eng = self._engines[n] = CLIPTextEngine(self._sd, n, **self._kw)
        return (eng(input_ids).clone(),)
