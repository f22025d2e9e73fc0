"""Seeded synthetic weights / images / text embeddings / head for the hot path.

No checkpoint, dataset or text template is available offline (the reference needs `ViT-B-32.pkl`,
`text_template/` and the competition images), so every test and the benchmark run on the same
deterministic generators defined here.  State-dict key names and shapes are the ones
`build_model` consumes (reference jclip/model.py:235-285); init scales are the OpenAI-CLIP ones the
reference applies in `initialize_parameters` (jclip/model.py:172-187) and in
`VisionTransformer.__init__` (jclip/model.py:93-102), so activations stay O(1) through 12 layers.
"""
import numpy as np

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
NUM_CLASSES = 403          # reference classes.txt
EMBED_DIM = 512


def make_vit_state_dict(seed=0, layers=12, width=768, patch=32, resolution=224, embed_dim=EMBED_DIM,
                        with_text_stub=True, text_layers=1, vpt_tokens=0, trained_like=False):
    """Random-init CLIP state dict (numpy fp32) with the reference's key names (vision tower + the
    few text-side keys `build_model` reads shapes from)."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    n = lambda shape, std: (rng.standard_normal(shape, dtype=f32) * f32(std))
    grid = resolution // patch
    scale = width ** -0.5
    attn_std = width ** -0.5
    proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
    fc_std = (2 * width) ** -0.5
    sd = {
        "visual.conv1.weight": n((width, 3, patch, patch), 0.02),
        "visual.class_embedding": n((width,), scale),
        "visual.positional_embedding": n((grid * grid + 1, width), scale),
        "visual.ln_pre.weight": (1.0 + n((width,), 0.05)),
        "visual.ln_pre.bias": n((width,), 0.02),
        "visual.ln_post.weight": (1.0 + n((width,), 0.05)),
        "visual.ln_post.bias": n((width,), 0.02),
        "visual.proj": n((width, embed_dim), scale),
    }
    if vpt_tokens:
        sd["visual.VPT"] = n((vpt_tokens, width), 0.02)      # reference jclip/model1.py:161-164
    if trained_like:
        # Activation statistics of TRAINED CLIP towers that random init never produces: "offset" = every residual row
        # carries a common offset (|row mean| = 20 x the spread of its channels); "outliers" = three "massive
        # activation" channels 100 x above the rest; True / "both" = both.  They enter through ln_pre's bias, i.e.
        # straight into the residual stream, and stay there through all blocks (LayerNorm removes them from what the
        # blocks compute, the residual adds keep them).
        kind = "both" if trained_like is True else str(trained_like)
        if kind not in ("offset", "outliers", "both"):
            raise ValueError(f"trained_like must be 'offset', 'outliers' or 'both', got {trained_like!r}")
        if kind in ("offset", "both"):
            sd["visual.ln_pre.bias"] = sd["visual.ln_pre.bias"] + f32(20.0)
        if kind in ("outliers", "both"):
            for ch, sign in ((7, 1.0), (300, -1.0), (511, 1.0)):
                sd["visual.ln_pre.bias"][ch] += f32(sign * 100.0)
    for i in range(layers):
        p = f"visual.transformer.resblocks.{i}."
        sd[p + "attn.in_proj_weight"] = n((3 * width, width), attn_std)
        sd[p + "attn.in_proj_bias"] = n((3 * width,), 0.02)
        sd[p + "attn.out_proj.weight"] = n((width, width), proj_std)
        sd[p + "attn.out_proj.bias"] = n((width,), 0.02)
        sd[p + "ln_1.weight"] = 1.0 + n((width,), 0.05)
        sd[p + "ln_1.bias"] = n((width,), 0.02)
        sd[p + "ln_2.weight"] = 1.0 + n((width,), 0.05)
        sd[p + "ln_2.bias"] = n((width,), 0.02)
        sd[p + "mlp.c_fc.weight"] = n((4 * width, width), fc_std)
        sd[p + "mlp.c_fc.bias"] = n((4 * width,), 0.02)
        sd[p + "mlp.c_proj.weight"] = n((width, 4 * width), proj_std)
        sd[p + "mlp.c_proj.bias"] = n((width,), 0.02)
    if with_text_stub:
        # shape carriers only: build_model reads these shapes (jclip/model.py:267-276).  The text
        # tower itself is out of scope of the hot path; a 1-layer stub (text_layers=1) keeps the dict small.
        tw = 512
        sd["text_projection"] = n((tw, embed_dim), tw ** -0.5)
        sd["positional_embedding"] = n((77, tw), 0.01)
        sd["token_embedding.weight"] = n((64, tw), 0.02)
        sd["ln_final.weight"] = np.ones((tw,), f32)
        sd["ln_final.bias"] = np.zeros((tw,), f32)
        sd["logit_scale"] = np.array(np.log(1 / 0.07), f32)
        for li in range(text_layers):
            p = f"transformer.resblocks.{li}."
            sd[p + "attn.in_proj_weight"] = n((3 * tw, tw), tw ** -0.5)
            sd[p + "attn.in_proj_bias"] = np.zeros((3 * tw,), f32)
            sd[p + "attn.out_proj.weight"] = n((tw, tw), tw ** -0.5 * (2 * text_layers) ** -0.5)
            sd[p + "attn.out_proj.bias"] = np.zeros((tw,), f32)
            sd[p + "ln_1.weight"] = np.ones((tw,), f32)
            sd[p + "ln_1.bias"] = np.zeros((tw,), f32)
            sd[p + "ln_2.weight"] = np.ones((tw,), f32)
            sd[p + "ln_2.bias"] = np.zeros((tw,), f32)
            sd[p + "mlp.c_fc.weight"] = n((4 * tw, tw), (2 * tw) ** -0.5)
            sd[p + "mlp.c_fc.bias"] = np.zeros((4 * tw,), f32)
            sd[p + "mlp.c_proj.weight"] = n((tw, 4 * tw), tw ** -0.5 * (2 * text_layers) ** -0.5)
            sd[p + "mlp.c_proj.bias"] = np.zeros((tw,), f32)
    return sd


def make_lora(seed=7, layers=12, width=768, r=4, params=("q", "k", "v"), b_std=0.005):
    """Synthetic adapters {layer: {'q_proj': (A[r,W], B[W,r]), ...}}.  A ~ U(+-1/sqrt(W)) is what
    kaiming_uniform(a=sqrt 5) gives (reference test.py:304); B is zero at init in the reference
    (test.py:305), which would make LoRA a no-op, so B ~ N(0, 0.005^2) matching the magnitudes in
    the shipped lora_weights1/lora_weights.pkl."""
    rng = np.random.default_rng(seed)
    names = {"q": "q_proj", "k": "k_proj", "v": "v_proj", "o": "proj"}
    bound = 1.0 / np.sqrt(width)
    out = {}
    for i in range(layers):
        out[i] = {}
        for p in params:
            A = rng.uniform(-bound, bound, size=(r, width)).astype(np.float32)
            B = (rng.standard_normal((width, r)) * b_std).astype(np.float32)
            out[i][names[p]] = (A, B)
    return out


def _upsample_bilinear(grid, size):
    """grid [..., g, g] -> [..., size, size], align_corners=False style bilinear (numpy)."""
    g = grid.shape[-1]
    pos = (np.arange(size, dtype=np.float32) + 0.5) * (g / size) - 0.5
    pos = np.clip(pos, 0, g - 1)
    i0 = np.floor(pos).astype(np.int64)
    i1 = np.minimum(i0 + 1, g - 1)
    w = (pos - i0).astype(np.float32)
    rows = grid[..., i0, :] * (1 - w)[:, None] + grid[..., i1, :] * w[:, None]
    return rows[..., :, i0] * (1 - w) + rows[..., :, i1] * w


def make_views(seed, n_images, n_views, resolution=224, noise=0.5):
    """[n_images, n_views, 3, R, R] float32 in [0,1] (pre CLIP-normalisation, as `test.py` feeds
    `tfm_clip`).  Per image: a smooth random field (14x14 control grid); per view: an independent
    perturbation field (28x28 grid), so that a view set has the spread an augmented crop set has
    (oracle cosine ~0.995 within an image, ~0.985 between images with random-init weights).  View 0
    is the unperturbed 'centre' view (reference test.py:1700 puts the un-augmented image first)."""
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((n_images, 1, 3, 14, 14)).astype(np.float32)
    pert = rng.standard_normal((n_images, n_views, 3, 28, 28)).astype(np.float32) * np.float32(noise)
    pert[:, 0] = 0
    img = _upsample_bilinear(base, resolution) + _upsample_bilinear(pert, resolution)
    img = 1.0 / (1.0 + np.exp(-1.5 * img))
    return np.ascontiguousarray(img, dtype=np.float32)


def clip_normalize(images):
    mean = np.asarray(CLIP_MEAN, np.float32).reshape(3, 1, 1)
    std = np.asarray(CLIP_STD, np.float32).reshape(3, 1, 1)
    return ((images - mean) / std).astype(np.float32)


def make_text_features(seed=1, num_classes=NUM_CLASSES, dim=EMBED_DIM, project_out=None):
    """[C,dim] unit rows.  `project_out` ([dim]) removes one direction from every row before
    normalising: a random-init tower maps every image close to one common direction, and text rows
    orthogonal to it make the class ranking depend on the image-specific part of the embedding (as
    it does with trained weights) instead of being the same for every image."""
    rng = np.random.default_rng(seed)
    t = rng.standard_normal((num_classes, dim)).astype(np.float32)
    if project_out is not None:
        u = np.asarray(project_out, np.float32).reshape(-1)
        u = u / np.linalg.norm(u)
        t = t - (t @ u)[:, None] * u[None, :]
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    return t.astype(np.float32)


def make_structured_text_banks(centre_feats, seed=10, num_classes=NUM_CLASSES, per_image=6, n_banks=3):
    """Three text banks [C, dim] (prompt-tuned / hand / zero-shot variants of the SAME classes) whose class scores
    are separated the way trained CLIP text features separate them, built from the images' own centre-view
    embeddings `centre_feats` [I, dim] (unit rows, e.g. from the fp32 oracle tower).

    A random-init tower maps every image next to one common direction m, so against independent random unit rows all
    403 scores of an image sit within a few logits and neighbouring scores at rank 5 are ~0.2 apart: any 0.05 logit
    perturbation flips a label there, which says nothing about the path under test.  Here image i owns `per_image`
    classes c = per_image * i + k whose text rows point along the image-specific part d_i = f_i - m with decreasing
    weight a_k = 1, 0.8, 0.64, ... (plus per-bank noise), so its top scores are whole logits apart -- as they are for a
    trained model and its class prompts; the remaining classes are random rows orthogonal to m."""
    f = np.asarray(centre_feats, np.float32)
    I, dim = f.shape
    if per_image * I > num_classes:
        raise ValueError(f"{I} images x {per_image} classes do not fit {num_classes} classes")
    rng = np.random.default_rng(seed)
    m = f.mean(0)
    m /= np.linalg.norm(m)
    d = f - (f @ m)[:, None] * m[None, :]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    base = make_text_features(seed + 100, num_classes, dim, project_out=m)
    banks = []
    for b in range(n_banks):
        t = base + 0.05 * rng.standard_normal(base.shape).astype(np.float32)
        for i in range(I):
            for k in range(per_image):
                noise = rng.standard_normal(dim).astype(np.float32)
                noise -= (noise @ d[i]) * d[i]
                noise /= np.linalg.norm(noise)
                a = 0.8 ** k
                t[per_image * i + k] = a * d[i] + np.sqrt(max(1.0 - a * a, 0.0)) * noise
        t /= np.linalg.norm(t, axis=1, keepdims=True)
        banks.append(t.astype(np.float32))
    return banks


def make_head(seed=2, text_zs=None, num_classes=NUM_CLASSES, dim=EMBED_DIM):
    """Channel_LP parameters (reference test.py:1223-1228): scale1, bias1 [512], fc [403,512]+[403].
    fc is initialised from the zero-shot text features in the reference (slow_pace.py:1539)."""
    rng = np.random.default_rng(seed)
    if text_zs is None:
        text_zs = make_text_features(3, num_classes, dim)
    scale1 = (1.0 + 0.1 * rng.standard_normal(dim)).astype(np.float32)
    bias1 = (0.01 * rng.standard_normal(dim)).astype(np.float32)
    fc_w = (text_zs + 0.01 * rng.standard_normal((num_classes, dim))).astype(np.float32)
    fc_b = np.zeros((num_classes,), np.float32)
    return scale1, bias1, fc_w, fc_b


def make_unit_views(seed, n_images, n_views, dim=EMBED_DIM, spread=0.35):
    """Synthetic *embeddings* for MTA/head-only tests: per image a unit centre plus per-view
    perturbations, re-normalised.  [n_images, n_views, dim] float32 unit rows."""
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((n_images, 1, dim)).astype(np.float32)
    c /= np.linalg.norm(c, axis=-1, keepdims=True)
    e = rng.standard_normal((n_images, n_views, dim)).astype(np.float32) / np.sqrt(dim)
    # a few outlier views per image, which is what MTA's inlierness scores are for
    out_mask = rng.random((n_images, n_views, 1)) < 0.1
    x = c + spread * e * np.where(out_mask, 4.0, 1.0).astype(np.float32)
    x[:, 0] = c[:, 0] + 0.05 * e[:, 0]
    x /= np.linalg.norm(x, axis=-1, keepdims=True)
    return x.astype(np.float32)


def make_views_torch(seed, n_images, n_views, device, resolution=224, noise=0.5, dtype=None):
    """Device-side twin of make_views for benchmark-sized batches (tens of thousands of views): same
    construction (smooth per-image field + per-view perturbation field, squashed to [0,1]) drawn from a
    torch generator on `device`.  Returns [n_images, n_views, 3, R, R] float32 (or `dtype`)."""
    import torch
    g = torch.Generator(device=device).manual_seed(int(seed))
    out = torch.empty((n_images, n_views, 3, resolution, resolution), dtype=dtype or torch.float32, device=device)
    step = max(1, 4096 // max(n_views, 1))
    for i0 in range(0, n_images, step):
        n = min(step, n_images - i0)
        base = torch.randn((n, 3, 14, 14), generator=g, device=device)
        pert = torch.randn((n * n_views, 3, 28, 28), generator=g, device=device) * noise
        pert = pert.view(n, n_views, 3, 28, 28)
        pert[:, 0] = 0
        up = lambda t: torch.nn.functional.interpolate(t, size=(resolution, resolution), mode="bilinear",
                                                       align_corners=False)
        img = up(base).unsqueeze(1) + up(pert.view(n * n_views, 3, 28, 28)).view(n, n_views, 3, resolution, resolution)
        out[i0:i0 + n] = torch.sigmoid(1.5 * img).to(out.dtype)
    return out


def make_tokens(seed, n, vocab=64, context=77, min_len=3, max_len=20):
    """Synthetic token ids [n, context] int64 in the layout `clip.tokenize` produces: <sot> words... <eot> then zero
    padding, with <sot> = vocab - 2 and <eot> = vocab - 1 (the highest id, which `encode_text` locates by argmax)."""
    rng = np.random.default_rng(seed)
    t = np.zeros((n, context), np.int64)
    for i in range(n):
        L = int(rng.integers(min_len, max_len + 1))
        t[i, 0] = vocab - 2
        t[i, 1:1 + L] = rng.integers(1, vocab - 2, L)
        t[i, 1 + L] = vocab - 1
    return t
