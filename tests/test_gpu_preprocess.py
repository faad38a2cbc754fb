"""GPU parity: CLIP preprocess kernels vs torchvision/Pillow (bit-exact uint8 stage, float32 output)."""
import numpy as np
import pytest

from facet_b200.synth import synth_image_bgr

pytestmark = pytest.mark.gpu


def _torchvision_preprocess(rgb, mean, std):
    import torchvision.transforms as T
    from PIL import Image
    tf = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224), T.ToTensor(),
                    T.Normalize(mean, std)])
    return tf(Image.fromarray(rgb)).numpy()


@pytest.mark.parametrize("shape", [(683, 1024), (1024, 683), (400, 600), (225, 300), (224, 224), (1000, 3000)])
def test_preprocess_matches_torchvision(shape):
    from facet_b200 import ops
    from facet_b200.utils import resample as rs
    h, w = shape
    bgr = np.stack([synth_image_bgr(i, h, w) for i in range(2)])
    for mean, std in ((rs.LAION_MEAN, rs.LAION_STD), (rs.OPENAI_MEAN, rs.OPENAI_STD)):
        got = ops.clip_preprocess(bgr, mean=mean, std=std).cpu().numpy()
        for i in range(2):
            want = _torchvision_preprocess(np.ascontiguousarray(bgr[i][..., ::-1]), mean, std)
            assert got[i].shape == (3, 224, 224)
            # same uint8 pixel => same float up to the last ulp of the two float32 divisions
            np.testing.assert_allclose(got[i], want, rtol=0, atol=3e-7)
    # RGB-order input gives the same planes
    got_rgb = ops.clip_preprocess(np.ascontiguousarray(bgr[..., ::-1]), rgb_order=True).cpu().numpy()
    np.testing.assert_array_equal(got_rgb, ops.clip_preprocess(bgr).cpu().numpy())


def test_preprocess_24mp():
    from facet_b200 import ops
    bgr = synth_image_bgr(4, 4000, 6000)
    got = ops.clip_preprocess(bgr).cpu().numpy()[0]
    want = _torchvision_preprocess(np.ascontiguousarray(bgr[..., ::-1]), (0.5, 0.5, 0.5), (0.5, 0.5, 0.5))
    np.testing.assert_allclose(got, want, rtol=0, atol=3e-7)


@pytest.mark.parametrize("shape", [(683, 1024), (400, 608), (4000, 6000), (1024, 1024)])
def test_tensor_core_and_cuda_core_horizontal_pass_agree(shape):
    """csrc/resample_tc.cu (u8 x s8 tcgen05 product with base-128 coefficient limbs) is exact: identical
    output to the CUDA-core kernel and to torchvision."""
    import torch
    from facet_b200 import ops
    from facet_b200.utils import resample as rs
    h, w = shape
    assert rs.plan(h, w).tc_coef is not None
    bgr = np.stack([synth_image_bgr(30 + i, h, w) for i in range(3)])
    a = ops.clip_preprocess(bgr, tensor_cores=True)
    b = ops.clip_preprocess(bgr, tensor_cores=False)
    assert torch.equal(a, b)
    want = _torchvision_preprocess(np.ascontiguousarray(bgr[1][..., ::-1]), (0.5, 0.5, 0.5), (0.5, 0.5, 0.5))
    np.testing.assert_allclose(a[1].cpu().numpy(), want, rtol=0, atol=3e-7)
