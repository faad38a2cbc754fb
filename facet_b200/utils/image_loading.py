"""Orientation / channel order of decoded frames (the tail of utils/image_loading.py:44-112).

The reference decodes on the CPU (`Image.open` / rawpy), applies `ImageOps.exif_transpose`, converts to
RGB and hands the analyzers a BGR copy (`cv2.cvtColor(RGB2BGR)`, :106).  File decoding stays outside this
path; what is here is the pixel work after it, for frames that are uploaded as
decoded: one byte-moving kernel does the transpose method PIL would pick for the EXIF orientation and the
channel swap (`csrc/orient.cu`).
"""
from __future__ import annotations

EXIF_ORIENTATION_TAG = 0x0112

# (swap, flip_x, flip_y) of out(x', y') = in(sx, sy), (u, v) = swap ? (y', x') : (x', y'),
# sx = flip_x ? W-1-u : u, sy = flip_y ? H-1-v : v -- the table fb_orient uses, keyed by the EXIF code
METHODS = {1: (0, 0, 0), 2: (0, 1, 0), 3: (0, 1, 1), 4: (0, 0, 1), 5: (1, 0, 0), 6: (1, 0, 1), 7: (1, 1, 1), 8: (1, 1, 0)}


def exif_orientation(pil_img) -> int:
    """EXIF orientation code of a PIL image (1 when absent or invalid), as `exif_transpose` reads it."""
    try:
        code = int(pil_img.getexif().get(EXIF_ORIENTATION_TAG, 1))
    except Exception:
        return 1
    return code if 1 <= code <= 8 else 1


def orient_frames(frames_rgb, orientation: int = 1, to_bgr: bool = True):
    """Decoded RGB frames ([n,H,W,3] or [H,W,3] uint8, host or device) -> device frames in the layout the
    scoring pass takes (`img_cv`: upright, BGR).  Returns a CUDA uint8 tensor [n,H',W',3]."""
    from .. import ops
    return ops.orient(frames_rgb, orientation, swap_rb=to_bgr)


def decode_on_host(data):
    """The reference's own loader for one JPEG stream (utils/image_loading.py:90-106): Pillow decode, `exif_transpose`,
    RGB, then the BGR copy the analyzers take.  Used by the streamed pass for streams the device decoder does not accept
    (progressive, no restart markers, CMYK ...): file loading is outside the scoring path, exactly as in the reference.
    Returns an [H,W,3] uint8 BGR array, or None when Pillow cannot read the data either."""
    import io

    import numpy as np
    from PIL import Image, ImageOps
    try:
        img = Image.open(io.BytesIO(bytes(data)))
        img = ImageOps.exif_transpose(img).convert("RGB")
        return np.ascontiguousarray(np.asarray(img)[:, :, ::-1])
    except Exception:
        return None


def read_jpeg_item(path, pinned: bool = False):
    """One loader item for `BatchProcessor.process_items_streamed`: {'path', 'jpeg'} with the file's bytes (a view of pinned
    host memory when `pinned`, so that the upload is asynchronous).  The bytes are not decoded here."""
    import os

    import numpy as np
    size = os.path.getsize(path)
    if pinned:
        import torch
        buf = torch.empty(size, dtype=torch.uint8, pin_memory=True)
        arr = buf.numpy()
        with open(path, "rb") as f:
            f.readinto(memoryview(arr))
        return {"path": str(path), "jpeg": arr, "_pinned": buf}
    with open(path, "rb") as f:
        return {"path": str(path), "jpeg": np.frombuffer(f.read(), np.uint8)}


JPEG_EXTENSIONS = (".jpg", ".jpeg", ".jpe", ".jfif")


def load_any(path):
    """Host loader for files that are not baseline JPEG: what `load_image_from_path` (utils/image_loading.py:44-112) does
    for them with Pillow — open, `exif_transpose`, RGB, BGR copy.  (RAW files need rawpy, which is outside this path.)
    Returns an [H,W,3] uint8 BGR array or None."""
    import numpy as np
    from PIL import Image, ImageOps
    try:
        with Image.open(path) as img:
            img = ImageOps.exif_transpose(img).convert("RGB")
            return np.ascontiguousarray(np.asarray(img)[:, :, ::-1])
    except Exception:
        return None


def load_item(path):
    """One loader item for a path: JPEG files as undecoded bytes (`read_jpeg_item`), everything else decoded on the host;
    unreadable files become error items like `_load_image` produces (batch_processor.py:92,110)."""
    import os
    p = str(path)
    try:
        if os.path.splitext(p)[1].lower() in JPEG_EXTENSIONS:
            return read_jpeg_item(p)
        img = load_any(p)
        if img is None:
            return {"path": p, "error": "Failed to load image"}
        return {"path": p, "img_cv": img}
    except Exception as exc:
        return {"path": p, "error": str(exc)}
