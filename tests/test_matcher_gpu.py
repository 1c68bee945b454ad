"""f2: the GPU Hungarian matcher against the reference's own results (tests/golden/matcher_small.npz) and against the
oracle (scipy) on its own cost matrix: index-exact where the optimum is unique, set-exact for tiled (repeats > 1) columns."""
import numpy as np
import pytest
import torch

from oracle import matcher_oracle as MO
from tests.helpers import GOLDEN
from tests.matcher_cases import CASES

pytestmark = pytest.mark.gpu


def _run(z, name):
    from sam3_lora_b200.matcher import BinaryHungarianMatcherV2

    kw, B, Q, nb, rep, rb, masks = CASES[name]
    outs = {"pred_logits": torch.from_numpy(z[f"{name}.logits"]).cuda(), "pred_boxes": torch.from_numpy(z[f"{name}.pred_boxes"]).cuda()}
    tg = {"boxes_padded": torch.from_numpy(z[f"{name}.boxes_padded"]).cuda(), "num_boxes": torch.tensor(nb)}
    ov = torch.from_numpy(z[f"{name}.out_valid"]).cuda() if masks else None
    tv = torch.from_numpy(z[f"{name}.tgt_valid"]).cuda() if masks else None
    m = BinaryHungarianMatcherV2(**kw)
    return m, outs, tg, ov, tv, m(outs, tg, repeats=rep, repeat_batch=rb, out_is_valid=ov, target_is_valid_padded=tv)


@pytest.mark.parametrize("name", list(CASES))
def test_gpu_matcher_matches_reference_golden(name):
    z = np.load(GOLDEN / "matcher_small.npz")
    kw, B, Q, nb, rep, rb, masks = CASES[name]
    m, outs, tg, ov, tv, (bi, si, ti) = _run(z, name)
    bi, si = bi.cpu().numpy(), si.cpu().numpy()
    assert bi.dtype == np.int64 and si.dtype == np.int64
    assert np.array_equal(bi, z[f"{name}.batch_idx"])
    if rep == 1:
        assert np.array_equal(si, z[f"{name}.src_idx"])
    else:
        assert sorted(zip(bi.tolist(), si.tolist())) == sorted(zip(z[f"{name}.batch_idx"].tolist(), z[f"{name}.src_idx"].tolist()))
    assert (ti is not None) == bool(z[f"{name}.has_tgt"])
    if ti is not None:
        assert np.array_equal(ti.cpu().numpy(), z[f"{name}.tgt_idx"])
    # cost matrix vs the oracle's float32 restatement
    tb = tg["boxes_padded"].repeat(rb, 1, 1)
    cost, qoc, coq = m.match(outs["pred_logits"].squeeze(-1), outs["pred_boxes"], tb, torch.tensor(nb * rb), rep, ov, tv)
    Cref = MO.cost_matrix(z[f"{name}.logits"][..., 0], z[f"{name}.pred_boxes"], tb.cpu().numpy(), w_class=kw.get("cost_class", 1),
                          w_bbox=kw.get("cost_bbox", 1), w_giou=kw.get("cost_giou", 1), focal=kw.get("focal", False),
                          alpha=kw.get("alpha", 0.25), gamma=kw.get("gamma", 2.0), stable=kw.get("stable", False),
                          out_valid=None if ov is None else ov.cpu().numpy(), tgt_valid=None if tv is None else tv.cpu().numpy())
    C = cost.cpu().numpy()
    assert np.allclose(C, Cref, rtol=2e-5, atol=2e-6)
    # optimality on the GPU's own cost matrix: same total as scipy, image by image
    for b, n in enumerate(nb * rb):
        if n == 0:
            assert (coq[b] < 0).all()
            continue
        cb = np.tile(C[b, :, :n].astype(np.float64), (1, rep))
        i, j = MO.linear_sum_assignment(cb)
        cols = coq[b].cpu().numpy()
        q = np.nonzero(cols >= 0)[0]
        if not masks:
            assert len(q) == min(Q, n * rep)
            assert abs(cb[q, cols[q]].sum() - cb[i, j].sum()) < 1e-9 * max(1.0, abs(cb[i, j].sum()))
        assert len(set(cols[q].tolist())) == len(q)                       # a matching: no column used twice
        assert all(qoc[b, c].item() == qq for qq, c in zip(q.tolist(), cols[q].tolist()))


def test_gpu_matcher_random_problems_equal_scipy_and_empty_batch():
    from sam3_lora_b200.matcher import BinaryHungarianMatcherV2

    g = torch.Generator().manual_seed(9)
    m = BinaryHungarianMatcherV2(focal=True, cost_class=2.0, cost_bbox=5.0, cost_giou=2.0)
    B, Q, Tmax = 16, 200, 64
    nb = torch.randint(0, Tmax + 1, (B,), generator=g)
    c = torch.rand(B, Q, 2, generator=g) * 0.8 + 0.1
    pb = torch.cat([c, torch.rand(B, Q, 2, generator=g) * 0.3 + 0.02], -1)
    tb = torch.cat([torch.rand(B, Tmax, 2, generator=g) * 0.8 + 0.1, torch.rand(B, Tmax, 2, generator=g) * 0.3 + 0.02], -1)
    lg = torch.randn(B, Q, generator=g) * 2
    cost, qoc, coq = m.match(lg.cuda(), pb.cuda(), tb.cuda(), nb, 1)
    C = cost.cpu().numpy().astype(np.float64)
    for b in range(B):
        n = int(nb[b])
        if n == 0:
            continue
        i, j = MO.linear_sum_assignment(C[b, :, :n])
        ref = np.full(n, -1)
        ref[j] = i
        assert np.array_equal(qoc[b, :n].cpu().numpy(), ref), b        # identical assignment (unique optimum)
    # no targets anywhere / empty batch
    bi, si, ti = m({"pred_logits": lg[:2, :, None].cuda(), "pred_boxes": pb[:2].cuda()},
                   {"boxes_padded": tb[:2].cuda(), "num_boxes": torch.tensor([0, 0])})
    assert bi.numel() == 0 and si.numel() == 0 and ti is None
