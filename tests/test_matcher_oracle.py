"""The matcher oracle (numpy cost matrix + scipy LSA) is pinned against tests/golden/matcher_small.npz, produced by the
reference's own BinaryHungarianMatcherV2 (tests/golden/make_golden_matcher.py)."""
import numpy as np
import pytest

from oracle import matcher_oracle as MO
from tests.helpers import GOLDEN
from tests.matcher_cases import CASES


def oracle_run(z, name):
    kw, B, Q, nb, rep, rb, masks = CASES[name]
    ov = z[f"{name}.out_valid"] if masks else None
    tv = z[f"{name}.tgt_valid"] if masks else None
    tb = np.tile(z[f"{name}.boxes_padded"], (rb, 1, 1))
    C = MO.cost_matrix(z[f"{name}.logits"][..., 0], z[f"{name}.pred_boxes"], tb, w_class=kw.get("cost_class", 1), w_bbox=kw.get("cost_bbox", 1),
                       w_giou=kw.get("cost_giou", 1), focal=kw.get("focal", False), alpha=kw.get("alpha", 0.25), gamma=kw.get("gamma", 2.0),
                       stable=kw.get("stable", False), out_valid=ov, tgt_valid=tv)
    return C, MO.match(C, nb * rb, rep, do_filtering=masks)


def pair_set(bi, si, ti, nb, rep):
    """{(image, query, target)} with tiled columns folded back onto their target (copies of a target are interchangeable)."""
    return sorted(zip(bi.tolist(), si.tolist(), ti.tolist()))


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_matcher(name):
    z = np.load(GOLDEN / "matcher_small.npz")
    _, (bi, si, ti) = oracle_run(z, name)
    assert np.array_equal(bi, z[f"{name}.batch_idx"])
    if CASES[name][4] == 1:                      # unique optimum: identical indices, identical order
        assert np.array_equal(si, z[f"{name}.src_idx"])
    else:                                        # tiled columns: the same queries per image (copies of a target tie exactly)
        assert sorted(zip(bi.tolist(), si.tolist())) == sorted(zip(z[f"{name}.batch_idx"].tolist(), z[f"{name}.src_idx"].tolist()))
    assert (ti is not None) == bool(z[f"{name}.has_tgt"])
    if ti is not None:
        assert np.array_equal(ti, z[f"{name}.tgt_idx"])
