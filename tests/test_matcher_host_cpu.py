"""Host logic of the GPU matcher's drop-in forward() on the CPU: the device kernels are replaced by the oracle (SciPy) through
`match`, so what is exercised is the index assembly of BinaryHungarianMatcherV2.forward — packed target offsets, repeat_batch,
validity filtering, the more-targets-than-queries branch — against the reference's own results (tests/golden/matcher_small.npz)."""
import numpy as np
import pytest
import torch

from oracle import matcher_oracle as MO
from tests.helpers import GOLDEN
from tests.matcher_cases import CASES


def _oracle_match(self, out_score, out_bbox, tgt_bbox, num_boxes, repeats=1, out_is_valid=None, target_is_valid_padded=None):
    """Stand-in for the CUDA kernels with the same outputs: (cost, query_of_col, col_of_query)."""
    kw = dict(w_class=self.cost_class, w_bbox=self.cost_bbox, w_giou=self.cost_giou, focal=self.focal, alpha=self.alpha,
              gamma=self.gamma, stable=self.stable)
    ov = None if out_is_valid is None else out_is_valid.numpy().astype(bool)
    tv = None if target_is_valid_padded is None else target_is_valid_padded.numpy().astype(bool)
    C = MO.cost_matrix(out_score.numpy(), out_bbox.numpy(), tgt_bbox.numpy(), out_valid=ov, tgt_valid=tv, **kw)
    B, Q, Tmax = C.shape
    rep = max(int(repeats), 1)
    qoc = -np.ones((B, max(1, Tmax * rep)), np.int32)
    coq = -np.ones((B, Q), np.int32)
    for b, n in enumerate(num_boxes.tolist()):
        if n == 0:
            continue
        i, j = MO.match_image(C[b, :, :n], rep, do_filtering=ov is not None or tv is not None)
        qoc[b, j] = i
        coq[b, i] = j
    return torch.from_numpy(C), torch.from_numpy(qoc), torch.from_numpy(coq)


@pytest.mark.parametrize("name", list(CASES))
def test_forward_index_assembly_matches_the_reference(name, monkeypatch):
    from sam3_lora_b200.matcher import BinaryHungarianMatcherV2

    monkeypatch.setattr(BinaryHungarianMatcherV2, "match", _oracle_match)
    z = np.load(GOLDEN / "matcher_small.npz")
    kw, B, Q, nb, rep, rb, masks = CASES[name]
    outs = {"pred_logits": torch.from_numpy(z[f"{name}.logits"]), "pred_boxes": torch.from_numpy(z[f"{name}.pred_boxes"])}
    tg = {"boxes_padded": torch.from_numpy(z[f"{name}.boxes_padded"]), "num_boxes": torch.tensor(nb)}
    ov = torch.from_numpy(z[f"{name}.out_valid"]) if masks else None
    tv = torch.from_numpy(z[f"{name}.tgt_valid"]) if masks else None
    bi, si, ti = BinaryHungarianMatcherV2(**kw)(outs, tg, repeats=rep, repeat_batch=rb, out_is_valid=ov, target_is_valid_padded=tv)
    assert bi.dtype == torch.long and si.dtype == torch.long
    assert np.array_equal(bi.numpy(), z[f"{name}.batch_idx"])
    if rep == 1:
        assert np.array_equal(si.numpy(), z[f"{name}.src_idx"])
    else:
        assert sorted(zip(bi.tolist(), si.tolist())) == sorted(zip(z[f"{name}.batch_idx"].tolist(), z[f"{name}.src_idx"].tolist()))
    assert (ti is not None) == bool(z[f"{name}.has_tgt"])
    if ti is not None:
        assert np.array_equal(ti.numpy(), z[f"{name}.tgt_idx"])
