"""GPU parity of the TTA view generator (jcb_tta_views) against Pillow itself, BIT FOR BIT on uint8:
centre view = Resize(256, BICUBIC) + CenterCrop(224) (jclip/clip.py:130-135), crops = RandomResizedCrop
(BILINEAR) + RandomHorizontalFlip (test.py:1898-1903)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _img(rng, H, W, smooth=True):
    a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    if smooth:
        a = (np.cumsum(np.cumsum(a.astype(np.float64), 0), 1) % 256).astype(np.uint8)
    return a


def _check(jb, imgs, gen):
    from oracle import crops as C
    jobs = gen.draw_jobs([im.shape[:2] for im in imgs])
    out = gen(imgs, jobs=jobs).cpu().numpy()
    out = out.reshape(-1, 3, gen.size, gen.size)
    bad = 0
    for k, j in enumerate(jb.tta.jobs_to_tuples(jobs)):
        im = imgs[j[0]]
        if j[9] == 1:
            ref = C.pil_centre_view(im, gen.resize, gen.size)
        else:
            ref = C.pil_crop_view(im, j[1], j[2], j[3], j[4], j[10], gen.size)
        if not np.array_equal(out[k], ref):
            bad += 1
            d = np.abs(out[k].astype(int) - ref.astype(int))
            print("job", k, j, "max diff", d.max(), "frac", (d > 0).mean())
    return bad, len(jobs)


def test_views_match_pillow_bit_for_bit(jb, cuda_dev):
    rng = np.random.default_rng(0)
    imgs = [_img(rng, 375, 500), _img(rng, 500, 333), _img(rng, 256, 256), _img(rng, 240, 1000, smooth=False),
            _img(rng, 1200, 900)]
    gen = jb.TTAViews(n_crops=12, scale=(0.2, 1.0), seed=5)
    bad, n = _check(jb, imgs, gen)
    assert n == 5 * 13 and bad == 0


def test_upscaling_and_tiny_crops(jb, cuda_dev):
    """scale 0.05: crops far smaller than 224 px are magnified (filterscale = 1, two taps) -- reference
    lora_train_vlp.py uses scale=(0.05, 1)."""
    rng = np.random.default_rng(1)
    imgs = [_img(rng, 120, 160), _img(rng, 300, 300, smooth=False)]
    bad, n = _check(jb, imgs, jb.TTAViews(n_crops=20, scale=(0.05, 0.3), seed=9))
    assert bad == 0 and n == 42


def test_explicit_boxes_on_the_image_border(jb, cuda_dev):
    from oracle import crops as C
    rng = np.random.default_rng(2)
    im = _img(rng, 400, 600)
    gen = jb.TTAViews(n_crops=0)
    boxes = [(0, 0, 400, 600, 0), (0, 0, 224, 224, 1), (176, 376, 224, 224, 0), (399 - 50, 599 - 70, 51, 71, 1),
             (0, 0, 1, 1, 0), (10, 20, 300, 225, 1)]
    jobs = (jb._capi.ViewJob * len(boxes))()
    for j, (t, l, h, w, f) in zip(jobs, boxes):
        j.image, j.top, j.left, j.crop_h, j.crop_w = 0, t, l, h, w
        j.out_h, j.out_w, j.off_y, j.off_x, j.filter, j.flip = 224, 224, 0, 0, 0, f
    out = gen([im], jobs=jobs).cpu().numpy().reshape(-1, 3, 224, 224)
    for k, (t, l, h, w, f) in enumerate(boxes):
        assert np.array_equal(out[k], C.pil_crop_view(im, t, l, h, w, f)), boxes[k]


def test_byte_paths_unaligned_source_and_odd_size(jb, cuda_dev):
    """The word-wide resample passes need a 4-byte aligned source and size % 4 == 0; everything else takes the
    byte-wise bodies of the same kernels.  Both must equal Pillow: a source pointer 1 byte off, and size = 222."""
    import ctypes
    from oracle import crops as C
    get_context, check = jb.get_context, jb._capi.check
    rng = np.random.default_rng(7)
    im = _img(rng, 333, 421)
    boxes = [(0, 0, 333, 421, 0), (5, 7, 200, 260, 1), (100, 120, 233, 301, 0), (3, 1, 64, 80, 1)]
    ctx = get_context(cuda_dev)
    for size, shift in [(224, 1), (224, 3), (222, 0), (222, 2)]:
        jobs = (jb._capi.ViewJob * len(boxes))()
        for j, (t, l, h, w, f) in zip(jobs, boxes):
            j.image, j.top, j.left, j.crop_h, j.crop_w = 0, t, l, h, w
            j.out_h, j.out_w, j.off_y, j.off_x, j.filter, j.flip = size, size, 0, 0, 0, f
        buf = torch.zeros(im.size + 64, dtype=torch.uint8, device=cuda_dev)
        buf[shift:shift + im.size] = torch.from_numpy(im.reshape(-1)).to(cuda_dev)
        desc = (jb._capi.SrcImage * 1)()
        desc[0].offset, desc[0].height, desc[0].width = 0, im.shape[0], im.shape[1]
        out = torch.empty((len(boxes), 3, size, size), dtype=torch.uint8, device=cuda_dev)
        with torch.cuda.device(cuda_dev):
            ctx.bind_current_stream()
            check(ctx.lib.jcb_tta_views(ctx.handle, None, ctypes.c_void_p(buf.data_ptr() + shift), desc, 1, jobs, len(boxes), size,
                                        ctypes.c_void_p(out.data_ptr())), ctx.handle)
        got = out.cpu().numpy()
        for k, (t, l, h, w, f) in enumerate(boxes):
            assert np.array_equal(got[k], C.pil_crop_view(im, t, l, h, w, f, size)), (size, shift, boxes[k])


def test_rejects_bad_jobs(jb, cuda_dev):
    im = np.zeros((100, 100, 3), np.uint8)
    gen = jb.TTAViews(n_crops=0)
    jobs = (jb._capi.ViewJob * 1)()
    j = jobs[0]
    j.image, j.top, j.left, j.crop_h, j.crop_w, j.out_h, j.out_w = 0, 50, 50, 60, 60, 224, 224   # box leaves the image
    with pytest.raises(jb.JcbError):
        gen([im], jobs=jobs)


def test_views_feed_the_hot_path(jb, cuda_dev):
    """images -> TTAViews -> HotPath.evaluate_base equals feeding the Pillow-built views (as uint8/255 floats)."""
    from oracle import crops as C
    rng = np.random.default_rng(3)
    imgs = [_img(rng, 300, 400), _img(rng, 420, 380)]
    gen = jb.TTAViews(n_crops=6, seed=2)
    jobs = gen.draw_jobs([im.shape[:2] for im in imgs])
    views = gen(imgs, jobs=jobs)
    assert views.shape == (2, 7, 3, 224, 224) and views.dtype == torch.uint8 and views.is_cuda
    ref = np.stack([C.pil_centre_view(imgs[j[0]]) if j[9] == 1 else C.pil_crop_view(imgs[j[0]], j[1], j[2], j[3], j[4], j[10])
                    for j in jb.tta.jobs_to_tuples(jobs)]).reshape(2, 7, 3, 224, 224)
    sd = jb.synth.make_vit_state_dict(seed=4, layers=2)
    model = jb.jclip.build_model(sd)
    Ts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
    lp = jb.Channel_LP()
    lp.scale1.data, lp.bias1.data, lp.fc.weight.data, lp.fc.bias.data = jb.synth.make_head(2, Ts[2].numpy())
    hp = jb.HotPath(model, jb.TextBank(Ts[0], Ts[1], Ts[2], cuda_dev), lp)
    a = hp.evaluate_base(views)
    b = hp.evaluate_base(torch.from_numpy(ref.astype(np.float32) / 255.0).to(cuda_dev))
    assert torch.equal(a, b)


@pytest.mark.parametrize("op", ["f16", "bf16"])
@pytest.mark.parametrize("norm", [True, False])
def test_fused_patches_equal_views_plus_im2col(jb, cuda_dev, op, norm):
    """jcb_tta_patches (resampler's last pass writes the conv1 patch matrix: ToTensor, tfm_clip, im2col fused) is
    bit-identical to jcb_tta_views (Pillow-exact uint8) followed by jcb_im2col -- word-wide and byte-wise bodies,
    bicubic centre view, flipped and unflipped bilinear crops, image sizes that are not multiples of four."""
    rng = np.random.default_rng(11)
    imgs = [_img(rng, 300, 400), _img(rng, 421, 383), _img(rng, 257, 999)]
    ctx = jb.get_context(cuda_dev)
    prev = ctx.operand_type
    try:
        ctx.set_operand_type(op)
        dt = ctx.operand_torch_dtype
        gv = jb.TTAViews(n_crops=9, seed=5)
        jobs = gv.draw_jobs_fast([im.shape[:2] for im in imgs])
        views = gv(imgs, jobs=jobs)                                            # [3, 10, 3, 224, 224] uint8
        want = torch.empty(30 * 49, 3072, dtype=dt, device=cuda_dev)
        jb.blocks.im2col(views.view(30, 3, 224, 224), 224, 32, norm, want)
        gp = jb.TTAViews(n_crops=9, seed=5, emit="patches", apply_clip_norm=norm)
        got = gp(imgs, jobs=jobs)
        assert got.shape == (3, 10, 49, 3072) and got.dtype == dt
        assert torch.equal(got.view(30 * 49, 3072), want)
        # the byte-wise bodies: an unaligned source buffer
        from jclip_b200._capi import check
        im = imgs[1]
        buf = torch.zeros(im.size + 64, dtype=torch.uint8, device=cuda_dev)
        buf[1:1 + im.size] = torch.from_numpy(im.reshape(-1)).to(cuda_dev)
        desc = (jb._capi.SrcImage * 1)()
        desc[0].offset, desc[0].height, desc[0].width = 0, im.shape[0], im.shape[1]
        j1 = np.ascontiguousarray(jobs[10:20].copy())
        j1["image"] = 0
        out_b = torch.empty(10 * 49, 3072, dtype=dt, device=cuda_dev)
        check(ctx.lib.jcb_tta_patches(ctx.handle, None, ctypes.c_void_p(buf.data_ptr() + 1), desc, 1,
                                      j1.ctypes.data_as(ctypes.POINTER(jb._capi.ViewJob)), 10, 224, 32, int(norm),
                                      jb._capi.OPERAND_NAMES[op], ctypes.c_void_p(out_b.data_ptr())), ctx.handle)
        ctx.sync()
        assert torch.equal(out_b, want[10 * 49:20 * 49])
    finally:
        ctx.set_operand_type(prev)


def test_patches_feed_the_hot_path_and_image_stream(jb, cuda_dev):
    """HotPath.evaluate_base(patches) == evaluate_base(views), bit for bit, and HotPath.evaluate_image_stream (views or
    patches generated on a second stream while the towers run the previous batch) returns the same top-k per batch."""
    rng = np.random.default_rng(4)
    batches = [[_img(rng, 300 + 7 * i, 400 - 5 * i) for i in range(n)] for n in (3, 2, 4, 1, 3)]
    sd = jb.synth.make_vit_state_dict(seed=4, layers=2)
    model = jb.jclip.build_model(sd)
    Ts = [torch.from_numpy(jb.synth.make_text_features(seed=10 + i)) for i in range(3)]
    lp = jb.Channel_LP()
    lp.scale1.data, lp.bias1.data, lp.fc.weight.data, lp.fc.bias.data = jb.synth.make_head(2, Ts[2].numpy())
    hp = jb.HotPath(model, jb.TextBank(Ts[0], Ts[1], Ts[2], cuda_dev), lp, rank_by="cs5")
    want = []
    for emit in ("views", "patches"):
        gen = jb.TTAViews(n_crops=6, seed=2, emit=emit)
        got = [hp.evaluate_base(gen(b)).cpu() for b in batches]
        if not want:
            want = got
        assert all(torch.equal(g, w) for g, w in zip(got, want)), emit
        gen = jb.TTAViews(n_crops=6, seed=2, emit=emit)                  # same seed -> same boxes
        streamed = list(hp.evaluate_image_stream(iter(batches), gen))
        assert len(streamed) == len(want)
        assert all(torch.equal(g, w) for g, w in zip(streamed, want)), emit
    # a patch matrix of the wrong operand type is refused
    ctx = jb.get_context(cuda_dev)
    other = "bf16" if ctx.operand_type == "f16" else "f16"
    p = jb.TTAViews(n_crops=1, seed=0, emit="patches")(batches[0])
    wrong = p.to(torch.bfloat16 if p.dtype == torch.float16 else torch.float16)
    with pytest.raises(jb.JcbError):
        hp.evaluate_base(wrong)
    assert other
