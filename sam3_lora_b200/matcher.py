"""Drop-in for `sam3.train.matcher.BinaryHungarianMatcherV2` (sam3/train/matcher.py:431-668) on the GPU (next-row f2).

Same constructor and `forward(outputs, batched_targets, repeats, repeat_batch, out_is_valid, target_is_valid_padded)`
-> `(batch_idx, src_idx, tgt_idx)`.  The reference builds the cost matrix with ~15 torch kernels, copies it to the host
(`C.cpu().numpy()`, a device sync) and calls `scipy.optimize.linear_sum_assignment` once per image; here the cost matrix and
the assignment are two CUDA kernels (csrc/matcher.cu) and the matching never leaves the device.  When every image has at
most as many (repeated) targets as queries and no validity mask is given — SAM3's training case — the number of matches
is known from `num_boxes` alone and the index tensors are assembled without any read-back of results.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib as L
from ._abi import MatcherDesc


class BinaryHungarianMatcherV2(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, focal: bool = False,
                 alpha: float = 0.25, gamma: float = 2.0, stable: bool = False, remove_samples_with_0_gt: bool = True):
        super().__init__()
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self.cost_class, self.cost_bbox, self.cost_giou = cost_class, cost_bbox, cost_giou
        self.focal = focal
        self.alpha, self.gamma, self.stable = alpha, gamma, stable
        self.remove_samples_with_0_gt = remove_samples_with_0_gt   # images without targets produce no pairs either way

    @torch.no_grad()
    def match(self, out_score, out_bbox, tgt_bbox, num_boxes, repeats=1, out_is_valid=None, target_is_valid_padded=None):
        """Raw device result: (cost [B,Q,Tmax] fp32, query_of_col [B, Tmax*repeats] int32, col_of_query [B,Q] int32)."""
        if not out_score.is_cuda:
            raise L.Sam3bError("matcher: inputs are on the CPU; the GPU matcher has no CPU fallback")
        dev = out_score.device
        B, Q = out_score.shape
        Tmax = tgt_bbox.shape[1] if tgt_bbox.dim() == 3 else 0
        rep = max(int(repeats), 1)
        d = MatcherDesc()
        score = out_score.detach().float().contiguous()
        pbox = out_bbox.detach().float().contiguous()
        tbox = tgt_bbox.detach().float().contiguous()
        nb = num_boxes.to(device=dev, dtype=torch.int32).contiguous()
        ov = None if out_is_valid is None else out_is_valid.to(device=dev, dtype=torch.uint8).contiguous()
        tv = None if target_is_valid_padded is None else target_is_valid_padded.to(device=dev, dtype=torch.uint8).contiguous()
        d.B, d.Q, d.Tmax, d.repeats = B, Q, Tmax, rep
        d.logits, d.pred_boxes, d.tgt_boxes, d.num_boxes = L.ptr(score), L.ptr(pbox), L.ptr(tbox), L.ptr(nb)
        d.out_valid, d.tgt_valid = L.ptr(ov), L.ptr(tv)
        d.w_class, d.w_bbox, d.w_giou = float(self.cost_class), float(self.cost_bbox), float(self.cost_giou)
        d.focal, d.stable, d.alpha, d.gamma = int(self.focal), int(self.stable), float(self.alpha), float(self.gamma)
        cost = torch.empty(B, Q, max(Tmax, 1), device=dev, dtype=torch.float32)
        qoc = torch.empty(B, max(1, Tmax * rep), device=dev, dtype=torch.int32)
        coq = torch.empty(B, Q, device=dev, dtype=torch.int32)
        L.check(L.load().sam3b_matcher(C.byref(d), L.ptr(cost), L.ptr(qoc), L.ptr(coq), L.current_stream()))
        return cost[:, :, :Tmax], qoc, coq

    @torch.no_grad()
    def forward(self, outputs, batched_targets, repeats=1, repeat_batch=1, out_is_valid=None, target_is_valid_padded=None):
        out_score = outputs["pred_logits"].squeeze(-1)          # (B, Q)
        out_bbox = outputs["pred_boxes"]                        # (B, Q, 4)
        dev = out_score.device
        B, Q = out_score.shape
        num_boxes_host = batched_targets["num_boxes"].cpu()     # host copy of an INPUT (the collator's), as in the reference
        tgt_bbox = batched_targets["boxes_padded"]
        if repeat_batch > 1:                                    # concatenated final + auxiliary outputs (matcher.py:548-555)
            num_boxes_host = num_boxes_host.repeat(repeat_batch)
            tgt_bbox = tgt_bbox.repeat(repeat_batch, 1, 1)
            if target_is_valid_padded is not None:
                target_is_valid_padded = target_is_valid_padded.repeat(repeat_batch, 1)
        assert out_bbox.shape[0] == tgt_bbox.shape[0] == num_boxes_host.shape[0]
        rep = max(int(repeats), 1)
        _, qoc, coq = self.match(out_score, out_bbox, tgt_bbox, num_boxes_host, rep, out_is_valid, target_is_valid_padded)
        nb = num_boxes_host.tolist()
        do_filtering = out_is_valid is not None or target_is_valid_padded is not None
        return_tgt = do_filtering or any(Q < n * rep for n in nb)
        offsets = [0]
        for n in nb[:-1]:
            offsets.append(offsets[-1] + n)
        if not return_tgt:
            # every column of every image is matched: sizes are known on the host, nothing is read back
            keep = [b for b in range(B) if nb[b] > 0]
            if not keep:
                z = torch.empty(0, dtype=torch.long, device=dev)
                return z, z.clone(), None
            src_idx = torch.cat([qoc[b, : nb[b] * rep] for b in keep]).long()
            batch_idx = torch.cat([torch.full((nb[b] * rep,), b, dtype=torch.long, device=dev) for b in keep])
            return batch_idx, src_idx, None
        # general case (validity masks, or more targets than queries): pairs in query order, as scipy returns them
        valid = coq >= 0                                        # [B, Q]
        batch_idx, src_idx = valid.nonzero(as_tuple=True)       # row-major: by image, then by query  (sizes unknown: one sync)
        cols = coq[batch_idx, src_idx].long()
        off_dev = torch.tensor(offsets, device=dev, dtype=torch.long)
        return batch_idx, src_idx, cols + off_dev[batch_idx]    # column index + packed-target offset (matcher.py:637-640)
