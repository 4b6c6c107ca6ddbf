"""GPU: image pyramids built on the device equal the scalar vk::halfSample rule bit for bit."""
import numpy as np
import pytest

from rpg_svo_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,levels", [((640, 480), 5), ((752, 480), 5), ((1920, 1080), 6), ((70, 50), 3)])
def test_device_pyramid_bit_exact(ctx, oracle, size, levels):
    w, h = size
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    fr = ctx.frame_from_level0(img, levels)
    ref = img
    for l in range(levels):
        got = fr.download_level(l)
        assert np.array_equal(got, ref), f"level {l}"
        if l + 1 < levels:
            ref2 = oracle.half_sample(ref)
            assert np.array_equal(ref2, synth.half_sample(ref))  # oracle == numpy restatement
            ref = ref2
    fr.destroy()


@pytest.mark.parametrize("size,levels", [((640, 480), 5), ((752, 480), 5), ((1920, 1080), 6), ((130, 34), 4)])
def test_pool_fused_pyramid_bit_exact(ctx, size, levels):
    w, h = size
    rng = np.random.default_rng(w + h)
    imgs = rng.integers(0, 256, (3, h, w), dtype=np.uint8)
    pool = capi.FramePool(ctx, w, h, levels, 4)
    pool.upload_array(imgs, first=1)
    for i in range(3):
        pyr = synth.build_pyramid(imgs[i], levels)
        for l in range(levels):
            assert np.array_equal(pool.frames[1 + i].download_level(l), pyr[l]), (i, l)
    pool.destroy()
