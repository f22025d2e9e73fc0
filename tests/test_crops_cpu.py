"""CPU: the TTA view generator's host side and its oracle.  Pillow is the third-party carrier of the
resample arithmetic (jittor.transform delegates to it); the numpy restatement in oracle/crops.py -- the
algorithm the CUDA kernels implement -- is pinned against Pillow bit for bit here."""
import numpy as np
import pytest


def _img(rng, H, W, smooth):
    a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    if smooth:
        a = (np.cumsum(np.cumsum(a.astype(np.float64), 0), 1) % 256).astype(np.uint8)
    return a


@pytest.mark.parametrize("seed", range(6))
def test_restatement_equals_pillow(seed):
    from PIL import Image
    from oracle import crops as C
    rng = np.random.default_rng(seed)
    H, W = int(rng.integers(40, 360)), int(rng.integers(40, 360))
    img = _img(rng, H, W, seed % 2)
    ow, oh = int(rng.integers(16, 300)), int(rng.integers(16, 300))
    for kind, pk in ((C.BILINEAR, Image.BILINEAR), (C.BICUBIC, Image.BICUBIC)):
        ref = np.asarray(Image.fromarray(img, "RGB").resize((ow, oh), pk))
        assert np.array_equal(C.resample_restatement(img, ow, oh, kind), ref)


def test_identity_resize_is_exact():
    from oracle import crops as C
    img = _img(np.random.default_rng(1), 50, 70, 0)
    for kind in (C.BILINEAR, C.BICUBIC):
        assert np.array_equal(C.resample_restatement(img, 70, 50, kind), img)


def test_param_generators_match_the_oracle(jb):
    """The product's box generator (tta.py) and the oracle's restatement draw the same boxes from the same stream."""
    from oracle import crops as C
    for (W, H) in ((500, 375), (224, 224), (90, 400), (1024, 300), (31, 29)):
        assert jb.tta.centre_view_params(W, H) == C.centre_view_params(W, H)
        r1, r2 = np.random.default_rng(7), np.random.default_rng(7)
        for scale in ((0.5, 1.0), (0.2, 1.0), (0.05, 1.0)):
            for _ in range(20):
                assert jb.tta.random_resized_crop_params(r1, W, H, scale) == C.random_resized_crop_params(r2, W, H, scale)


def test_centre_view_params_follow_the_reference_resize():
    from oracle import crops as C
    assert C.centre_view_params(500, 375) == (341, 256, 58, 16)      # int(256 * 500 / 375) = 341
    assert C.centre_view_params(375, 500) == (256, 341, 16, 58)
    assert C.centre_view_params(256, 256) == (256, 256, 16, 16)
    assert C.centre_view_params(300, 256)[:2] == (300, 256)          # short side already 256: no resize


def test_crop_boxes_stay_inside_and_respect_scale(jb):
    rng = np.random.default_rng(3)
    W, H = 640, 480
    for _ in range(200):
        t, l, h, w = jb.tta.random_resized_crop_params(rng, W, H, (0.5, 1.0))
        assert 0 <= t and 0 <= l and t + h <= H and l + w <= W and h > 0 and w > 0
        assert 0.45 <= h * w / (W * H) <= 1.0 and 0.7 <= w / h <= 1.4


def test_draw_jobs_layout(jb):
    gen = jb.TTAViews(n_crops=5, seed=1)
    jobs = gen.draw_jobs([(375, 500), (300, 260)])
    assert len(jobs) == 12
    t = jb.tta.jobs_to_tuples(jobs)
    assert t[0] == (0, 0, 0, 375, 500, 256, 341, 16, 58, 1, 0)       # centre view first: bicubic, no flip
    assert t[6][0] == 1 and t[6][9] == 1 and all(x[9] == 0 and x[5:9] == (224, 224, 0, 0) for x in t[1:6])
    with pytest.raises(ValueError):
        jb.TTAViews(n_crops=1, resize=200).draw_jobs([(100, 2000)])   # Resize(200) cannot hold a 224-pixel centre crop


def test_fast_job_draw_has_the_same_layout_and_bounds(jb):
    import time
    gen = jb.TTAViews(n_crops=64, seed=3)
    shapes = [(375, 500)] * 64 + [(600, 400)] * 64
    t0 = time.perf_counter()
    jobs = gen.draw_jobs_fast(shapes)
    dt = time.perf_counter() - t0
    assert jobs.shape == (128 * 65,) and jobs.dtype == jb.tta.JOB_DTYPE and dt < 1.0
    t = jb.tta.jobs_to_tuples(jobs)
    assert t[0] == (0, 0, 0, 375, 500, 256, 341, 16, 58, 1, 0) and t[65 * 64][3:5] == (600, 400)
    for (img, top, left, h, w, oh, ow, oy, ox, filt, flip) in t:
        H, W = shapes[img]
        assert 0 <= top and 0 <= left and top + h <= H and left + w <= W and h > 0 and w > 0
        if filt == 0:
            assert (oh, ow, oy, ox) == (224, 224, 0, 0) and 0.45 <= h * w / (H * W) <= 1.0
    flips = np.mean([x[10] for x in t if x[9] == 0])
    assert 0.4 < flips < 0.6
