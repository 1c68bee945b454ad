"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's Hungarian matcher (sam3/train/matcher.py:15-29, 431-668;
box arithmetic sam3/model/box_ops.py:11-14, 91-142) in numpy float32 + scipy.optimize.linear_sum_assignment (the
third-party solver the reference itself calls; any SciPy >= 1.4 ships the same rectangular_lsap implementation).
Pinned by tests/golden/matcher_small.npz, produced by running the reference's own BinaryHungarianMatcherV2
(tests/golden/make_golden_matcher.py, tests/test_matcher_oracle.py)."""
from __future__ import annotations

import numpy as np
from scipy.optimize import linear_sum_assignment

F = np.float32


def _sigmoid(x):
    return (F(1) / (F(1) + np.exp(-x, dtype=F))).astype(F)


def _logsigmoid(x):
    return (np.minimum(x, F(0)) - np.log1p(np.exp(-np.abs(x), dtype=F), dtype=F)).astype(F)


def cost_matrix(logits, pred_boxes, tgt_boxes, w_class=1.0, w_bbox=1.0, w_giou=1.0, focal=False, alpha=0.25, gamma=2.0,
                stable=False, out_valid=None, tgt_valid=None):
    """logits [B,Q], pred_boxes [B,Q,4], tgt_boxes [B,T,4] (cxcywh) -> C [B,Q,T] float32 (matcher.py:571-617)."""
    s, p, t = logits.astype(F), pred_boxes.astype(F)[:, :, None, :], tgt_boxes.astype(F)[:, None, :, :]
    cost_bbox = np.abs(p - t).sum(-1, dtype=F)

    def xyxy(b):
        return np.stack([b[..., 0] - F(0.5) * b[..., 2], b[..., 1] - F(0.5) * b[..., 3], b[..., 0] + F(0.5) * b[..., 2],
                         b[..., 1] + F(0.5) * b[..., 3]], -1)

    a, b = xyxy(p), xyxy(t)
    area1 = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area2 = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    wh = np.clip(np.minimum(a[..., 2:], b[..., 2:]) - np.maximum(a[..., :2], b[..., :2]), 0, None)
    inter = wh[..., 0] * wh[..., 1]
    union = area1 + area2 - inter
    iou = inter / union
    whc = np.clip(np.maximum(a[..., 2:], b[..., 2:]) - np.minimum(a[..., :2], b[..., :2]), 0, None)
    areac = whc[..., 0] * whc[..., 1]
    cost_giou = -(iou - (areac - union) / areac)
    prob = _sigmoid(s)[:, :, None]
    if not focal:
        cost_class = np.broadcast_to(-prob, cost_bbox.shape)
    elif stable:
        pr = prob * ((-cost_giou + F(1)) / F(2))
        cost_class = -F(alpha) * (F(1) - pr) ** F(gamma) * np.log(pr) + (F(1) - F(alpha)) * pr ** F(gamma) * np.log(F(1) - pr)
    else:
        ls, l1s = _logsigmoid(s)[:, :, None], _logsigmoid(-s)[:, :, None]
        cost_class = np.broadcast_to(-F(alpha) * (F(1) - prob) ** F(gamma) * ls + (F(1) - F(alpha)) * prob ** F(gamma) * l1s,
                                     cost_bbox.shape)
    C = (F(w_bbox) * cost_bbox + F(w_class) * cost_class + F(w_giou) * cost_giou).astype(F)
    if out_valid is not None:
        C = np.where(out_valid[:, :, None], C, F(1e9))
    if tgt_valid is not None:
        C = np.where(tgt_valid[:, None, :], C, F(1e9))
    return C.astype(F)


def match_image(cost, repeats=1, do_filtering=False):
    """_do_matching (matcher.py:15-29) on one image's [Q, T] cost: (query indices, column indices), query-sorted."""
    if repeats > 1:
        cost = np.tile(cost, (1, repeats))
    i, j = linear_sum_assignment(cost)
    if do_filtering:
        keep = cost[i, j] < 1e8
        i, j = i[keep], j[keep]
    return i.astype(np.int64), j.astype(np.int64)


def match(C, num_boxes, repeats=1, do_filtering=False):
    """(batch_idx, src_idx, tgt_idx) as BinaryHungarianMatcherV2.forward returns them (matcher.py:618-668)."""
    B, Q, _ = C.shape
    rep = max(repeats, 1)
    return_tgt = do_filtering or any(Q < n * rep for n in num_boxes)
    bi, si, ti = [], [], []
    off = 0
    for b, n in enumerate(num_boxes):
        if n > 0:
            i, j = match_image(C[b, :, :n], repeats, do_filtering)
            if not return_tgt:
                i = i[np.argsort(j)]
            bi.append(np.full(len(i), b, np.int64)); si.append(i); ti.append(j + off)
        off += n
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)   # noqa: E731
    return cat(bi), cat(si), (cat(ti) if return_tgt else None)
