"""float32 ndarray <-> bytes, the wire/DB format of `clip_embedding` (utils/embedding.py:10,26)."""
import numpy as np


def embedding_to_bytes(embedding):
    if embedding is None:
        return None
    return np.asarray(embedding).astype(np.float32).tobytes()


def bytes_to_embedding(data, dim=None):
    if data is None:
        return None
    emb = np.frombuffer(data, dtype=np.float32)
    if dim is not None and len(emb) != dim:
        return None
    return emb
