"""Drop-in for the `text_encoder` object the reference passes around (transformers `CLIPTextModel`, loaded at
mixofshow/pipelines/trainer_edlora.py:41 and used at pipeline_edlora.py:133-145, trainer_edlora.py:220-234,
gradient_fusion.py:182-199, convert_edlora_to_diffusers.py:4-31,93-97): same call shape — `text_encoder(input_ids)[0]` is
the last hidden state — running on `mos_b200.clip_engine.CLIPTextEngine` (forward only; the training backward of the text
encoder is SURVEY.md §8f).  The surface the reference's checkpoint utilities touch is kept: `state_dict()`,
`load_state_dict()`, `get_input_embeddings().weight`, `resize_token_embeddings(n)`."""
from types import SimpleNamespace

import torch

TOKEN_KEY = 'text_model.embeddings.token_embedding.weight'


class CLIPTextModel:
    def __init__(self, state_dict, *, lora=None, lora_alpha=1.0, merge_lora=False, device='cuda'):
        """state_dict: transformers CLIPTextModel parameters (`text_model.*`).  Engines are built per batch size on
        first use (buffers are static); `load_state_dict` / `resize_token_embeddings` drop them, and the token table is
        re-uploaded after every `get_input_embeddings()` / `state_dict()` hand-out (the reference writes concept rows through
        `.weight.data[...]`)."""
        self._sd = {k: v.detach().clone() for k, v in state_dict.items()}
        self._kw = dict(lora=lora, lora_alpha=lora_alpha, merge_lora=merge_lora, device=device)
        self._engines = {}
        self._emb_dirty = False
        self.device = torch.device(device)
        self.dtype = torch.float32
        self.config = SimpleNamespace(hidden_size=self._sd[TOKEN_KEY].shape[1],
                                      max_position_embeddings=self._sd['text_model.embeddings.position_embedding.weight'].shape[0],
                                      vocab_size=self._sd[TOKEN_KEY].shape[0])

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder='text_encoder', **kw):
        """transformers call shape (`CLIPTextModel.from_pretrained(path, subfolder='text_encoder')`, trainer_edlora.py:41)."""
        from mixofshow.utils.model_io import load_text_encoder
        kw = {k: v for k, v in kw.items() if k in ('lora', 'lora_alpha', 'merge_lora', 'device')}
        return load_text_encoder(pretrained_model_name_or_path, subfolder, **kw)

    def save_pretrained(self, save_directory, **unused):
        from mixofshow.utils.model_io import save_text_encoder
        save_text_encoder(self, save_directory, subfolder=None)

    # ------------------------------------------------------------------ module-like surface
    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def state_dict(self):
        self._emb_dirty = True      # the caller may write into the returned tensors
        return dict(self._sd)

    def load_state_dict(self, state_dict, strict=True):
        missing = [k for k in self._sd if k not in state_dict]
        unexpected = [k for k in state_dict if k not in self._sd]
        if strict and (missing or unexpected):
            raise RuntimeError(f'load_state_dict: missing {missing[:3]}, unexpected {unexpected[:3]}')
        for k, v in state_dict.items():
            if k in self._sd:
                if tuple(v.shape) != tuple(self._sd[k].shape):
                    raise RuntimeError(f'size mismatch for {k}: {tuple(v.shape)} vs {tuple(self._sd[k].shape)}')
                self._sd[k] = v.detach().to(torch.float32).clone()
        self._engines = {}

    def get_input_embeddings(self):
        """Object with a `.weight` tensor that can be indexed / written in place (trainer_edlora.py:77-82,
        convert_edlora_to_diffusers.py:18-20)."""
        self._emb_dirty = True      # rows are written through `.weight.data[...]`, which no version counter sees
        return SimpleNamespace(weight=self._sd[TOKEN_KEY])

    def resize_token_embeddings(self, new_num_tokens):
        """transformers semantics: keep the existing rows, append rows for the new tokens (they are overwritten by the
        caller with the learned concept embeddings, convert_edlora_to_diffusers.py:17-20)."""
        old = self._sd[TOKEN_KEY]
        n0, c = old.shape
        if new_num_tokens != n0:
            new = torch.zeros(new_num_tokens, c, dtype=old.dtype)
            k = min(n0, new_num_tokens)
            new[:k] = old[:k]
            if new_num_tokens > n0:
                new[n0:] = old.mean(0, keepdim=True)     # placeholder rows until the caller writes them
            self._sd[TOKEN_KEY] = new
            self.config.vocab_size = new_num_tokens
            self._engines = {}
        return self.get_input_embeddings()

    def __call__(self, input_ids, attention_mask=None, **kw):
        from mos_b200.clip_engine import CLIPTextEngine
        assert attention_mask is None, 'the ED-LoRA pipelines never pass an attention mask to the text encoder'
        if self._emb_dirty:                 # the token table may have been edited since the last call: re-upload it
            for eng in self._engines.values():
                eng.set_token_embedding(self._sd[TOKEN_KEY])
            self._emb_dirty = False
        n, L = input_ids.shape
        eng = self._engines.get(n)
        if eng is None:
            eng = self._engines[n] = CLIPTextEngine(self._sd, n, **self._kw)
        T = eng.T
        if L > T:
            raise ValueError(f'sequence length {L} exceeds max_position_embeddings {T}')
        if L < T:
            # un-padded prompts (gradient_fusion.py:190-199 feeds them one by one): the encoder is causal, so the hidden
            # states of the first L positions do not depend on what follows - run the fixed-length engine on the sequence
            # padded with its own last id and return the first L positions
            pad = input_ids[:, -1:].expand(n, T - L)
            return (eng(torch.cat([input_ids, pad], 1))[:, :L].clone(),)
        return (eng(input_ids).clone(),)
