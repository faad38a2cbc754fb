"""Recipe for oracle/_ref: the reference's OWN technical analyzers as the CPU arm.

    python -m oracle.build_ref          (also run by __graft_entry__.build() when /root/reference exists)

The reference is pure Python, so "building" it means placing its two self-contained analyzer modules —
analyzers/technical.py and analyzers/image_cache.py (377 lines; imports: cv2, numpy, scipy, struct) — UNMODIFIED
under oracle/_ref/analyzers/, from where they lie in /root/reference.  oracle/_ref/ is git-ignored (no reference
source enters the history) but not gpurun-ignored, so the copy travels to the GPU box, where `bench.py --impl
reference` and the `cpu_baseline` leg time it (`kind: "reference"`).  Nothing else of the reference is placed there:
the CLIP tower needs open_clip (third-party, not installable here), so that part of the CPU arm stays the oracle's
fp32 tower (oracle/vit_torch.py).  A manifest with the sha256 of each file is written beside them.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
FILES = ("analyzers/technical.py", "analyzers/image_cache.py")
DEST = os.path.join(HERE, "_ref")


def build(ref_root: str = REF_ROOT) -> bool:
    """Returns True when oracle/_ref is in place (freshly copied or already there), False when there is no
    reference tree to copy from and nothing was placed before."""
    have = all(os.path.isfile(os.path.join(DEST, f)) for f in FILES)
    if not os.path.isdir(ref_root):
        return have
    manifest = {}
    for f in FILES:
        src, dst = os.path.join(ref_root, f), os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    # the package __init__ of the reference imports face / composition analyzers (onnxruntime etc.): not wanted here
    with open(os.path.join(DEST, "analyzers", "__init__.py"), "w") as fh:
        fh.write("# placed by oracle/build_ref.py: makes the two copied modules importable as a package\n")
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": ref_root, "files": manifest}, fh, indent=1)
    return True


def load():
    """(ImageCache, TechnicalAnalyzer) of the reference from oracle/_ref, or None when it is not in place."""
    if not all(os.path.isfile(os.path.join(DEST, f)) for f in FILES):
        return None
    import importlib
    import sys
    # analyzers/technical.py imports `analyzers.image_cache`, so the copied package is imported under its own name
    if "analyzers" in sys.modules and not getattr(sys.modules["analyzers"], "__file__", "").startswith(DEST):
        raise RuntimeError("another `analyzers` package is already imported; load oracle/_ref in a fresh process")
    sys.path.insert(0, DEST)
    try:
        ic = importlib.import_module("analyzers.image_cache")
        ta = importlib.import_module("analyzers.technical")
    finally:
        sys.path.remove(DEST)
    return ic.ImageCache, ta.TechnicalAnalyzer


if __name__ == "__main__":
    print("oracle/_ref in place" if build() else "no reference tree at " + REF_ROOT)
