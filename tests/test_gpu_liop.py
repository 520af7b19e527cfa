"""LIOP-144 on the GPU against the reference's own arithmetic.

`r3d_vl_liopdesc_process` is pinned by the COMPILED REFERENCE (oracle/_ref from /root/reference/src/thirdparty/liop/
vl_liop.c): its outputs are the committed golden vectors tests/golden/liop_ref_v1.npz, so the comparison below is
GPU vs reference, bit for bit -- not GPU vs our own restatement.  The patch extraction (OpenCV warpAffine + GaussianBlur)
is compared with the oracle's restatement (bit-exact, same evaluation order) and with cv2's golden patches."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx(r3dlib):
    c = r3dlib.Context((0,))
    yield c
    c.close()


def test_liop_process_equals_compiled_reference_golden(ctx):
    g = np.load(os.path.join(GOLD, "liop_ref_v1.npz"))
    got = ctx.debug_liop_process(g["patches"])
    assert np.array_equal(got.view(np.uint32), g["desc"].view(np.uint32))      # incl. flat / quantised (tie-heavy) patches


def test_liop_process_random_patches_equal_oracle(ctx, oracle):
    rng = np.random.default_rng(3)
    patches = rng.random((300, 41, 41)).astype(np.float32)
    patches[100:200] = np.floor(patches[100:200] * 5) / 5                      # exact ties
    patches[200:] = np.cumsum(patches[200:], 2) / 41
    got = ctx.debug_liop_process(patches)
    want = np.stack([oracle.liop_process(p) for p in patches])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_liop_describe_equals_oracle_and_cv2_patches(ctx, oracle):
    g = np.load(os.path.join(GOLD, "liop_patch_cv2_v1.npz"))
    img, kps, factor = g["img"], g["kps"], float(g["factor"])
    got = ctx.liop_describe(img, kps, factor)
    want = oracle.liop_describe(img, kps, factor)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))           # same evaluation order: bit-exact
    # against descriptors of cv2's own patches (reference's call sequence run in the build container): LIOP is order
    # based, so the float round-off of OpenCV's SIMD blur almost never moves a rank
    ref = np.stack([oracle.liop_process(p) for p in g["blurred"]])
    same = int(np.sum(np.all(got == ref, axis=1)))
    assert same >= 0.9 * len(got), same
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)


def test_liop_describe_full_image_size_and_borders(ctx, oracle):
    rng = np.random.default_rng(8)
    h, w = 1080, 1920
    img = rng.random((h, w)).astype(np.float32)
    img[:, : w // 3] = 0.5                                                      # a flat third: all-equal patches
    n = 4000
    kps = np.stack([rng.uniform(-20, w + 20, n), rng.uniform(-20, h + 20, n), rng.uniform(2, 40, n), rng.uniform(0, 360, n)], 1)
    got = ctx.liop_describe(img, kps.astype(np.float32), 8.0)
    want = oracle.liop_describe(img, kps.astype(np.float32), 8.0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
