"""GPU: image pyramids built on the device equal vk::halfSample [EXT] bit for bit -- both of vikit's branches: the SSE2
rounding an x86 build of the reference takes when the source width is a multiple of 16 (the default rule), and the scalar
rule.  The CPU side of the comparison is the oracle, which tests/test_oracle_pins.py pins against oracle/_ref (the
reference's own frame.cpp calling the restated halfSample with real SSE2 intrinsics)."""
import numpy as np
import pytest

from rpg_svo_b200 import capi, synth

pytestmark = pytest.mark.gpu

SIZES = [((640, 480), 5), ((752, 480), 5), ((1920, 1080), 6), ((70, 50), 3), ((656, 490), 5)]


@pytest.fixture(params=[synth.PYR_X86, synth.PYR_SCALAR], ids=["x86-sse2-rule", "scalar-rule"])
def rule(request, ctx):
    ctx.set_pyramid_rule(request.param)
    yield request.param
    ctx.set_pyramid_rule(synth.PYR_X86)


@pytest.mark.parametrize("size,levels", SIZES)
def test_device_pyramid_bit_exact(ctx, oracle, rule, size, levels):
    w, h = size
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    fr = ctx.frame_from_level0(img, levels)
    ref = img
    for l in range(levels):
        got = fr.download_level(l)
        assert np.array_equal(got, ref), f"level {l}"
        if l + 1 < levels:
            ref2 = oracle.half_sample(ref, rule)
            assert np.array_equal(ref2, synth.half_sample(ref, rule))  # oracle == numpy restatement
            ref = ref2
    fr.destroy()


@pytest.mark.parametrize("size,levels", [((640, 480), 5), ((752, 480), 5), ((1920, 1080), 6), ((130, 34), 4), ((656, 490), 5)])
def test_pool_fused_pyramid_bit_exact(ctx, rule, size, levels):
    w, h = size
    rng = np.random.default_rng(w + h)
    imgs = rng.integers(0, 256, (3, h, w), dtype=np.uint8)
    pool = capi.FramePool(ctx, w, h, levels, 4)
    pool.upload_array(imgs, first=1)
    for i in range(3):
        pyr = synth.build_pyramid(imgs[i], levels, rule)
        for l in range(levels):
            assert np.array_equal(pool.frames[1 + i].download_level(l), pyr[l]), (i, l)
    pool.destroy()


def test_the_two_rules_really_differ(ctx):
    """0,0,0,1 -> scalar 0, SSE2 1: the default rule is not the scalar one wherever the width is a multiple of 16."""
    img = np.zeros((32, 64), np.uint8)
    img[1::2, 1::2] = 1
    a = ctx.frame_from_level0(img, 2)
    ctx.set_pyramid_rule(synth.PYR_SCALAR)
    b = ctx.frame_from_level0(img, 2)
    ctx.set_pyramid_rule(synth.PYR_X86)
    assert np.all(a.download_level(1) == 1) and np.all(b.download_level(1) == 0)
    a.destroy(); b.destroy()
