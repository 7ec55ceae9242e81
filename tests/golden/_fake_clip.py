"""Stand-in for the frozen CLIP encoders + image datasets so that the reference trainers' OWN ``train()`` methods run
on cached features in the build container (TEST INFRASTRUCTURE; golden generation only).

Out of scope for the hot path (clip/, datasets/, utils/data_manager.py) and impossible to run here (no
weights, no images, no network): ``load_clip`` returns a model whose image encoder is the identity on already-extracted
feature vectors and whose text side looks prompts up in a synthetic text bank ``E[class, template]``.  Everything
downstream of the encoders -- the code this repo re-implements -- is the reference's, unmodified.
"""
from __future__ import annotations

import math
import types

import torch
import torch.nn as nn
from torch.utils.data import DataLoader, Dataset


class _Visual(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.output_dim = dim
        self._anchor = nn.Parameter(torch.zeros(1), requires_grad=False)     # so that next(parameters()) has a device

    def forward(self, x):
        return x.float()


class FakeClip(nn.Module):
    """`clip_model` as the trainers use it: .visual, .encode_image, .encode_text, .token_embedding, .logit_scale, .dtype and the
    attributes utils.trainer.TextEncoder pulls out (its forward then returns E[class, template] exactly)."""

    def __init__(self, E: torch.Tensor, logit_scale: float = math.log(100.0)):
        super().__init__()
        K, M, D = E.shape
        self.register_buffer("bank", E.clone())
        self.visual = _Visual(D)
        self.logit_scale = nn.Parameter(torch.tensor(float(logit_scale)))
        self.transformer = nn.Identity()
        self.ln_final = nn.Identity()
        self.positional_embedding = nn.Parameter(torch.zeros(2, D), requires_grad=False)
        self.text_projection = nn.Parameter(torch.eye(D), requires_grad=False)
        self.embed_dim = D

    @property
    def dtype(self):
        return torch.float32

    def encode_image(self, x):
        return self.visual(x)

    def encode_text(self, tokens):
        return self.bank[tokens[:, 0], tokens[:, 1]]

    def token_embedding(self, tokens):
        return self.bank[tokens[:, 0], tokens[:, 1]].unsqueeze(1).expand(-1, tokens.shape[1], -1)


class FeatureDataset(Dataset):
    def __init__(self, feats, labels):
        self.feats, self.labels = feats, labels

    def __len__(self):
        return self.feats.shape[0]

    def __getitem__(self, i):
        return {"img": self.feats[i], "label": int(self.labels[i])}


class FakeDataManager:
    """utils/data_manager.py surface: loaders yield {"img": cached feature, "label": int}; the train loader shuffles and drops the
    last partial batch when N_tr >= batch size (utils/data_manager.py:79)."""

    def __init__(self, classnames, f_tr, y_tr, f_te, y_te, f_val=None, y_val=None, bs_train=16, bs_test=32):
        self.dataset = types.SimpleNamespace(classnames=list(classnames))
        self.num_classes = len(classnames)
        self.lab2cname = {i: c for i, c in enumerate(classnames)}
        self.train_loader_x = DataLoader(FeatureDataset(f_tr, y_tr), batch_size=bs_train, shuffle=True, num_workers=0,
                                         drop_last=f_tr.shape[0] >= bs_train)
        self.test_loader = DataLoader(FeatureDataset(f_te, y_te), batch_size=bs_test, shuffle=False, num_workers=0)
        self.val_loader = None if f_val is None else DataLoader(FeatureDataset(f_val, y_val), batch_size=bs_test, shuffle=False,
                                                                num_workers=0)


def install(ref_modules, E: torch.Tensor, classnames, templates_of):
    """Patch the CLIP loader / tokenizer seen by the given reference modules.  `templates_of(config)` must be the reference's own
    `_get_templates` (the prompts are only used as keys into the synthetic bank)."""
    import clip.clip as clip_mod                              # the reference's clip package

    lookup = {}

    def fake_tokenize(texts, context_length=77, truncate=False):
        if isinstance(texts, str):
            texts = [texts]
        return torch.tensor([lookup[t] for t in texts], dtype=torch.long)

    def register_prompts(config):
        for ti, t in enumerate(templates_of(config)):
            for ci, c in enumerate(classnames):
                lookup[t.format(c)] = (ci, ti)

    clip_mod.tokenize = fake_tokenize
    for m in ref_modules:
        m.load_clip = lambda config, device, _E=E: FakeClip(_E).to(device)
    return register_prompts
