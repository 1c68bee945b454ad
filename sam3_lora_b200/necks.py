"""Drop-in for `sam3.model.necks.Sam3DualViTDetNeck` (sam3/model/necks.py:13-125): the ViTDet "simple FPN" that turns the
trunk's last feature map into d_model-channel maps at 4x / 2x / 1x / 0.5x resolution.

Same constructor, same sub-module names (`convs.{i}.dconv_2x2_0 | gelu | dconv_2x2_1 | dconv_2x2 | maxpool_2x2 | conv_1x1 |
conv_3x3`, `sam2_convs`) so reference checkpoints load unchanged; the arithmetic runs in `conv_ops._NeckFn` (tcgen05 GEMMs
+ the conv.cu kernels), forward and backward, instead of cuDNN.  The nn.Conv2d / nn.ConvTranspose2d children only hold
the parameters.
"""
from __future__ import annotations

from copy import deepcopy
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import conv_ops

# per scale factor: (layers before the 1x1 projection, divisor of the trunk width that reaches conv_1x1)
_BRANCH_LAYOUT = {
    4.0: (("dconv_2x2_0", 1, 2), ("gelu", 0, 0), ("dconv_2x2_1", 2, 4)),
    2.0: (("dconv_2x2", 1, 2),),
    1.0: (),
    0.5: (("maxpool_2x2", 0, 0),),
}


def _make_branch(dim: int, d_model: int, scale: float) -> nn.Sequential:
    if scale not in _BRANCH_LAYOUT:
        raise NotImplementedError(f"scale_factor={scale} is not supported yet.")
    seq = nn.Sequential()
    width = dim
    for name, div_in, div_out in _BRANCH_LAYOUT[scale]:
        if name.startswith("dconv"):
            seq.add_module(name, nn.ConvTranspose2d(dim // div_in, dim // div_out, kernel_size=2, stride=2))
            width = dim // div_out
        elif name == "gelu":
            seq.add_module(name, nn.GELU())
        else:
            seq.add_module(name, nn.MaxPool2d(kernel_size=2, stride=2))
    seq.add_module("conv_1x1", nn.Conv2d(width, d_model, kernel_size=1, bias=True))
    seq.add_module("conv_3x3", nn.Conv2d(d_model, d_model, kernel_size=3, padding=1, bias=True))
    return seq


class Sam3DualViTDetNeck(nn.Module):
    def __init__(self, trunk: nn.Module, position_encoding: nn.Module, d_model: int,
                 scale_factors=(4.0, 2.0, 1.0, 0.5), add_sam2_neck: bool = False):
        super().__init__()
        self.trunk = trunk
        self.position_encoding = position_encoding
        self.scale_factors = scale_factors
        dim: int = self.trunk.channel_list[-1]
        self.convs = nn.ModuleList(_make_branch(dim, d_model, float(s)) for s in scale_factors)
        self.sam2_convs = deepcopy(self.convs) if add_sam2_neck else None   # a clone with its own weights (necks.py:94-97)

    def _run(self, x: torch.Tensor, branches) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        feats = conv_ops.neck_forward(x, branches)
        pos = [self.position_encoding(f).to(f.dtype) for f in feats]
        return feats, pos

    def forward(self, tensor_list) -> Tuple[List[torch.Tensor], List[torch.Tensor], Optional[List[torch.Tensor]],
                                             Optional[List[torch.Tensor]]]:
        x = self.trunk(tensor_list)[-1]
        sam3_out, sam3_pos = self._run(x, self.convs)
        sam2_out = sam2_pos = None
        if self.sam2_convs is not None:
            sam2_out, sam2_pos = self._run(x, self.sam2_convs)
        return sam3_out, sam3_pos, sam2_out, sam2_pos
