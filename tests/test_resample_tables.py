"""CPU checks of the Pillow-exact coefficient tables (bicubic for the CLIP preprocess, Lanczos for pHash)."""
import numpy as np
from PIL import Image

from facet_b200.synth import synth_image_bgr
from facet_b200.utils import resample as rs


def _apply(L, hb, hc, vb, vc, out):
    h, w = L.shape
    tmp = np.zeros((h, out), np.uint8)
    for xo in range(out):
        f, n = hb[xo]
        acc = (1 << 21) + L[:, f:f + n].astype(np.int64) @ hc[xo, :n].astype(np.int64)
        tmp[:, xo] = np.clip(acc >> 22, 0, 255)
    res = np.zeros((out, out), np.uint8)
    for yo in range(out):
        f, n = vb[yo]
        acc = (1 << 21) + vc[yo, :n].astype(np.int64) @ tmp[f:f + n].astype(np.int64)
        res[yo] = np.clip(acc >> 22, 0, 255)
    return res


def test_lanczos_tables_match_pillow():
    for idx, (h, w) in enumerate([(683, 1024), (97, 131), (400, 600), (33, 35)]):
        bgr = synth_image_bgr(idx, h, w)
        pil = Image.fromarray(np.ascontiguousarray(bgr[..., ::-1])).convert("L")
        want = np.asarray(pil.resize((32, 32), Image.LANCZOS))
        hb, hc, hk, vb, vc, vk = rs.phash_plan(h, w)
        assert np.array_equal(_apply(np.asarray(pil), hb, hc, vb, vc, 32), want)


def test_bicubic_plan_matches_torchvision():
    import torchvision.transforms as T
    tf = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224)])
    for idx, (h, w) in enumerate([(683, 1024), (1024, 683), (225, 300), (500, 231)]):
        rgb = np.ascontiguousarray(synth_image_bgr(idx, h, w)[..., ::-1])
        want = np.asarray(tf(Image.fromarray(rgb)))
        assert np.array_equal(rs.resample_reference_numpy(rgb), want)
        p = rs.plan(h, w)
        # the padded, 4-pixel re-based table is the same taps
        for xo in (0, 100, 223):
            off = int(p.hbounds[xo, 0] - p.hp0[xo])
            cnt = int(p.hbounds[xo, 1])
            assert np.array_equal(p.hcpad[off:off + cnt, xo], p.hcoef[xo, :cnt])
            assert not p.hcpad[:off, xo].any() and not p.hcpad[off + cnt:, xo].any()
