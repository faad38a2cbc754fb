"""CLIP zero-shot tagger — mirrors models/tagger.py:13-158 of the reference.

The reference multiplies one stored image embedding by the text-embedding matrix per call
(`image_features @ self.text_embeddings.T`, tagger.py:99-101, batch 1, inside the hot loop of
processing/batch_processor.py:205-213).  Here the same matmul is part of the tail kernel of the
ViT forward pass for the whole batch (csrc/vit.cu vit_tail_kernel); this class keeps the
reference's interface and does the per-tag max / threshold / top-k selection on the similarities.

The text tower / tokenizer of open_clip are not available offline, so `text_embeddings`
([n_prompts,768] L2-normalised) and `tag_names` (one name per prompt) are injected by the caller
— exactly the two attributes `_precompute_text_embeddings` (tagger.py:51-75) would fill.
"""
from __future__ import annotations

import numpy as np

from ..utils.embedding import bytes_to_embedding


def select_tags(tag_names, similarities, threshold=0.25, max_tags=5):
    """tagger.py:103-114: best similarity per tag name, >= threshold, sorted descending, top-N."""
    scores = {}
    for name, sim in zip(tag_names, similarities):
        if name not in scores or sim > scores[name]:
            scores[name] = float(sim)
    kept = [(t, s) for t, s in scores.items() if s >= threshold]
    kept.sort(key=lambda x: x[1], reverse=True)
    return [t for t, _ in kept[:max_tags]]


class CLIPTagger:
    def __init__(self, clip_model=None, device="cuda", config=None, text_embeddings=None, tag_names=None):
        self.model = clip_model
        self.device = device
        self.config = config
        self.text_embeddings = None
        self.tag_names = None
        self.tag_vocabulary = config.get_tag_vocabulary() if config is not None else {}
        self.art_tags = config.get_art_tags() if (config is not None and hasattr(config, "get_art_tags")) else set()
        if text_embeddings is not None:
            self.set_text_embeddings(text_embeddings, tag_names)

    def prompt_tag_names(self):
        """One tag name per prompt, in vocabulary order (tagger.py:62-66)."""
        names = []
        for tag, descriptions in self.tag_vocabulary.items():
            names.extend([tag] * len(descriptions))
        return names

    def set_text_embeddings(self, text_embeddings, tag_names=None):
        emb = np.ascontiguousarray(np.asarray(text_embeddings, dtype=np.float32))
        names = list(tag_names) if tag_names is not None else self.prompt_tag_names()
        if len(names) != emb.shape[0]:
            raise ValueError("one tag name per text embedding row is required")
        self.text_embeddings = emb
        self.tag_names = names
        self._dev = None

    def _device_matrix(self):
        import torch
        if self._dev is None:
            self._dev = torch.from_numpy(self.text_embeddings).to(self.device)
        return self._dev

    def similarities(self, clip_embedding_bytes):
        """emb[1,768] @ T[n,768]^T on the GPU (float32), as tagger.py:99-101 (fb_embedding_heads, no library GEMV)."""
        import torch
        from .. import ops
        emb = torch.from_numpy(bytes_to_embedding(clip_embedding_bytes).copy()).to(self.device).reshape(1, -1)
        _, sims = ops.embedding_heads(emb, tag_embeddings=self._device_matrix())
        return sims[0].cpu().numpy()

    def get_tags_from_embedding(self, clip_embedding_bytes, threshold=0.25, max_tags=5):
        if self.text_embeddings is None or clip_embedding_bytes is None:
            return []
        return select_tags(self.tag_names, self.similarities(clip_embedding_bytes), threshold, max_tags)

    def get_tags_from_similarities(self, sims_row, threshold=0.25, max_tags=5):
        """Same selection on a row of the batched similarity matrix the ViT tail kernel produced."""
        if self.tag_names is None:
            return []
        return select_tags(self.tag_names, sims_row, threshold, max_tags)

    def get_tags_with_scores(self, clip_embedding_bytes, threshold=0.20):
        if self.text_embeddings is None or clip_embedding_bytes is None:
            return {}
        scores = {}
        for name, sim in zip(self.tag_names, self.similarities(clip_embedding_bytes)):
            if name not in scores or sim > scores[name]:
                scores[name] = float(sim)
        return {t: round(s, 3) for t, s in scores.items() if s >= threshold}

    def is_artwork(self, clip_embedding_bytes, threshold=0.24):
        """tagger.py:146-158: any art-category tag among the ten best tags at this threshold."""
        tags = self.get_tags_from_embedding(clip_embedding_bytes, threshold=threshold, max_tags=10)
        return bool(set(tags) & self.art_tags)
