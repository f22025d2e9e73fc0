"""Oracle: Channel_LP (LP++ channel re-weighting head), logit_normalize, score fusion, top-5.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py; parity unpinned).

Restates test.py:1223-1234 (Channel_LP), test.py:1304-1308 (logit_normalize),
test.py:1708-1738 (per-image fusion in evaluate_base), test.py:1766-1774 (evaluate_new),
ood.py:875-883 (OOD routing).
"""
import torch

from .mta import solve_mta

OOD_BASE_MAX = 372   # ood.py:880 routes pred <= 372 to the base list (bug-compatible; SURVEY C-2)


def channel_lp(features, scale1, bias1, fc_w, fc_b):
    # test.py:1229-1234: fc(scale1 * f + bias1); fc = nn.Linear(512, 403)
    f = scale1.unsqueeze(0) * features + bias1.unsqueeze(0)
    return f @ fc_w.t() + fc_b


def logit_normalize(logit):
    # test.py:1304-1308.  jt.std(logit) with no dim = one global scalar; Jittor's std is the
    # unbiased (n-1) estimator (assumed semantics).  Row mean is per row.
    std = logit.std(unbiased=True)
    mean = logit.mean(dim=1, keepdim=True)
    return (logit - mean) / std


def topk_lowest_index_first(scores, k=5):
    """Var.topk(k) descending (test.py:1738); ties broken by lowest index (oracle definition)."""
    s = scores.to(torch.float64)
    order = torch.argsort(-s, dim=-1, stable=True)
    return order[..., :k]


def fuse_scores(m_pt, m_hand, m_zs, T_pt, T_hand, T_zs, lp):
    """test.py:1710-1736 for one image.  m_* are [1,512] modes, T_* are [C,512] unit rows,
    lp = (scale1, bias1, fc_w, fc_b).  Returns dict of every intermediate score the reference names."""
    combine = (m_pt + m_hand) / 2                                  # :1710 (not re-normalised)
    logits1 = logit_normalize(channel_lp(combine, *lp))            # :1715,:1717
    logits2 = logit_normalize(channel_lp(m_zs, *lp))               # :1716,:1718
    logits = logit_normalize((logits1 + logits2) / 2)              # :1721-1722
    cs = 100.0 * m_hand @ T_hand.t()                               # :1729
    cs1 = 100.0 * m_pt @ T_pt.t()                                  # :1730
    cs3 = 100.0 * m_zs @ T_zs.t()                                  # :1731
    cs2 = (cs + cs1) / 2                                           # :1733
    cs4 = (cs2 + cs3) / 2                                          # :1734
    cs5 = (cs2 + cs3) / 2 + 0.5 * logits                           # :1735
    return {"logits": logits, "cs": cs, "cs1": cs1, "cs2": cs2, "cs3": cs3, "cs4": cs4, "cs5": cs5}


@torch.no_grad()
def pipeline_image(feats, feats_zs, T_pt, T_hand, T_zs, lp, score="cs5"):
    """evaluate_base body for one image (test.py:1705-1742) given the unit view embeddings.

    feats    [V,512]: LoRA tower embeddings of the V views (row 0 = centre view)
    feats_zs [V,512]: second tower's embeddings (pass `feats` again for the single-tower bench)
    Returns (top5 indices [5], scores dict, modes dict).
    """
    m_pt = solve_mta(feats, T_pt.t())                              # :1708
    m_hand = solve_mta(feats, T_hand.t())                          # :1709
    m_zs = solve_mta(feats_zs, T_zs.t())                           # :1713
    sc = fuse_scores(m_pt, m_hand, m_zs, T_pt, T_hand, T_zs, lp)
    top5 = topk_lowest_index_first(sc[score], 5)[0]                # :1738 (reference ranks cs1; BASELINE names cs5)
    return top5, sc, {"pt": m_pt, "hand": m_hand, "zs": m_zs}
