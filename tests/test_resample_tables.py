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


def test_tensor_core_limb_tables_reproduce_taps():
    """The banded int8 matrices of csrc/resample_tc.cu recombine to the exact 22-bit taps."""
    rng = np.random.default_rng(0)
    for (h, w) in [(4000, 6000), (683, 1024), (225, 304)]:
        p = rs.plan(h, w)
        assert p.tc_coef is not None and p.tc_kw % 128 == 0 and p.tc_limbs in (3, 4)
        row = rng.integers(0, 256, w * 3).astype(np.int64)
        tbl = p.tc_coef.reshape(p.out // 8, 96, p.tc_kw).astype(np.int64)
        for j in (0, p.out // 16, p.out // 8 - 1):
            lo = int(p.tc_kb0[j])
            seg = np.zeros(p.tc_kw, np.int64)
            hi = min(lo + p.tc_kw, w * 3)
            seg[:hi - lo] = row[lo:hi]
            d = tbl[j] @ seg
            for xl in range(8):
                xo = 8 * j + xl
                f, c = p.hbounds[xo]
                for ch in range(3):
                    want = int(row[(f + np.arange(c)) * 3 + ch] @ p.hcoef[xo, :c].astype(np.int64))
                    got = sum(int(d[L * 24 + xl * 3 + ch]) * 128 ** L for L in range(p.tc_limbs))
                    assert want == got
    c = np.array([0, 1, -1, 63, 64, -64, -65, 8191, -8192, 4194304, -4194304, 262143])
    for limbs in (3, 4):
        d = rs.split_limbs(c, limbs)
        assert np.array_equal(sum(d[i] * 128 ** i for i in range(limbs)), c)
