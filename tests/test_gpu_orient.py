"""fb_orient against PIL's exif_transpose + cv2's RGB2BGR byte for byte (utils/image_loading.py:101-106)."""
import numpy as np
import pytest
from PIL import Image, ImageOps

from facet_b200.utils.image_loading import EXIF_ORIENTATION_TAG

pytestmark = pytest.mark.gpu

SHAPES = [(64, 64), (128, 192), (37, 53), (1, 7), (5, 1), (65, 130), (200, 300), (683, 1024), (96, 112), (100, 80), (132, 208), (64, 16)]


def _reference(img, code, to_bgr):
    pil = Image.fromarray(img)
    exif = pil.getexif()
    exif[EXIF_ORIENTATION_TAG] = code
    pil.info["exif"] = exif.tobytes()
    out = np.asarray(ImageOps.exif_transpose(pil))
    return out[..., ::-1] if to_bgr else out


@pytest.mark.parametrize("shape", SHAPES)
def test_orient_matches_pil(shape):
    from facet_b200 import ops
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    imgs = rng.integers(0, 256, size=(3, *shape, 3), dtype=np.uint8)
    for code in range(1, 9):
        for to_bgr in (False, True):
            got = ops.orient(imgs, code, swap_rb=to_bgr).cpu().numpy()
            for k in range(3):
                want = _reference(imgs[k], code, to_bgr)
                assert got[k].shape == want.shape and np.array_equal(got[k], want), (shape, code, to_bgr, k)


def test_orient_24mp_round_trips_and_feeds_the_pass():
    """Full size: every method followed by its inverse is the identity; an oriented RGB upload scores like
    the upright BGR frame."""
    import torch
    from facet_b200 import ops
    from facet_b200.synth import synth_image_bgr
    g = torch.Generator(device="cuda").manual_seed(5)
    frame = torch.randint(0, 256, (1, 4000, 6000, 3), dtype=torch.uint8, device="cuda", generator=g)
    inverse = {1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 8, 7: 7, 8: 6}
    for code, inv in inverse.items():
        fwd = ops.orient(frame, code)
        assert tuple(fwd.shape[1:3]) == ((6000, 4000) if code >= 5 else (4000, 6000))
        assert torch.equal(ops.orient(fwd, inv), frame), code
    bgr = synth_image_bgr(7, 256, 384)
    sideways_rgb = np.ascontiguousarray(np.rot90(bgr[..., ::-1], k=1))         # what a decoder hands over for code 6
    upright = ops.orient(sideways_rgb, 6, swap_rb=True)
    assert np.array_equal(upright[0].cpu().numpy(), bgr)
    a, b = ops.tech_stats(upright)[0], ops.tech_stats(bgr)[0]
    assert a.hist256.tolist() == b.hist256.tolist() and a.sum_lap_sq == b.sum_lap_sq


def test_orient_rejects_bad_arguments():
    from facet_b200 import ops
    with pytest.raises(ValueError):
        ops.orient(np.zeros((4, 4, 3), np.uint8), 0)
