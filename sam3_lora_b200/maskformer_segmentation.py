"""Drop-in for the mask head of `sam3.model.maskformer_segmentation` (sam3/model/maskformer_segmentation.py): `PixelDecoder`
(:172-219), `MaskPredictor` (:23-51), `SegmentationHead` (:54-169), `UniversalSegmentationHead` (:222-336), `LinearPresenceHead`
(:14-20) and the `MLP` they use (sam3/model/model_misc.py:160-195).

Same constructors, parameter names and return dictionaries as the reference, so its checkpoints load and its callers
(`Sam3Image._run_segmentation_heads`, sam3/model/sam3_image.py) need no change.  The arithmetic runs on the sm_100a
kernels through `conv_ops` (3x3 convs as im2col + tcgen05 GEMM, GroupNorm+ReLU kernels, 1x1 heads and the mask einsum
as GEMMs) and `ops.lora_linear` (the Linear layers of the mask-embedding MLP); there is no CPU path.
Not mirrored: `compile_mode` (torch.compile is not used on this path) and `act_ckpt` (nothing is recomputed: the decoder
saves one fp32 map per stage).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import conv_ops, ops


class MLP(nn.Module):
    """Linear -> ReLU -> ... -> Linear with `layers.{i}` parameter names (model_misc.py:160-195)."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_layers: int, dropout: float = 0.0,
                 residual: bool = False, out_norm: Optional[nn.Module] = None):
        super().__init__()
        if residual and input_dim != output_dim:
            raise ValueError("residual is only supported if input_dim == output_dim")
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.num_layers = num_layers
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.drop = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        self.residual = residual
        self.out_norm = out_norm if out_norm is not None else nn.Identity()

    @staticmethod
    def _linear(layer, x):
        if hasattr(layer, "original_layer"):      # LoRALinear-wrapped: its own fused forward
            return layer(x)
        return ops.lora_linear(x, layer.weight, layer.bias, None, None, 1.0, 0.0)

    def forward(self, x):
        y = x
        for i, layer in enumerate(self.layers):
            y = self._linear(layer, y)
            if i < self.num_layers - 1:
                y = self.drop(F.relu(y))
        if self.residual:
            y = y + x
        return self.out_norm(y)


class LinearPresenceHead(nn.Sequential):
    def __init__(self, d_model):
        super().__init__(nn.Identity(), nn.Identity(), nn.Linear(d_model, 1))   # index 2 keeps old checkpoints loadable

    def forward(self, hs, prompt, prompt_mask):
        return super().forward(hs)


class MaskPredictor(nn.Module):
    def __init__(self, hidden_dim, mask_dim):
        super().__init__()
        self.mask_embed = MLP(hidden_dim, hidden_dim, mask_dim, 3)

    def forward(self, obj_queries, pixel_embed):
        """obj_queries [B,Q,C] or [L,B,Q,C]; pixel_embed [B,C,H,W] or [C,H,W] -> mask logits [(L,)B,Q,H,W]."""
        me = self.mask_embed(obj_queries)
        layered = me.dim() == 4
        if layered:                                   # fold the decoder-layer axis into the query axis of each image
            Lyr, B, Q, Cc = me.shape
            me = me.permute(1, 0, 2, 3).reshape(B, Lyr * Q, Cc)
        if pixel_embed.dim() == 3:                    # batch size was omitted: one map for every image
            pixel_embed = pixel_embed.unsqueeze(0).expand(me.shape[0], -1, -1, -1)
        masks = conv_ops.mask_einsum(me, pixel_embed)
        if layered:
            H, W = masks.shape[-2:]
            masks = masks.view(B, Lyr, Q, H, W).permute(1, 0, 2, 3, 4)
        return masks


class PixelDecoder(nn.Module):
    def __init__(self, hidden_dim, num_upsampling_stages, interpolation_mode="nearest", shared_conv=False, compile_mode=None):
        super().__init__()
        if interpolation_mode != "nearest":
            raise NotImplementedError("the fused up-sample + add kernel implements mode='nearest' (the reference default)")
        self.hidden_dim = hidden_dim
        self.num_upsampling_stages = num_upsampling_stages
        self.interpolation_mode = interpolation_mode
        self.shared_conv = shared_conv
        n = 1 if shared_conv else num_upsampling_stages
        self.conv_layers = nn.ModuleList(nn.Conv2d(hidden_dim, hidden_dim, 3, 1, 1) for _ in range(n))
        self.norms = nn.ModuleList(nn.GroupNorm(8, hidden_dim) for _ in range(n))
        self.out_dim = self.conv_layers[-1].out_channels

    def forward(self, backbone_feats: List[torch.Tensor]):
        return conv_ops.pixel_decoder_forward(backbone_feats, self.conv_layers, self.norms, self.shared_conv)


class SegmentationHead(nn.Module):
    def __init__(self, hidden_dim, upsampling_stages, use_encoder_inputs=False, aux_masks=False, no_dec=False,
                 pixel_decoder=None, act_ckpt=False, shared_conv=False, compile_mode_pixel_decoder=None):
        super().__init__()
        self.use_encoder_inputs = use_encoder_inputs
        self.aux_masks = aux_masks
        self.no_dec = no_dec
        self.act_ckpt = act_ckpt
        self.pixel_decoder = pixel_decoder if pixel_decoder is not None else PixelDecoder(
            hidden_dim, upsampling_stages, shared_conv=shared_conv)
        if no_dec:
            raise NotImplementedError("no_dec=True (3x3 conv mask predictor) is not used by SAM3's image model")
        self.mask_predictor = MaskPredictor(hidden_dim, mask_dim=hidden_dim)
        self.instance_keys = ["pred_masks"]

    @property
    def device(self):
        return next(self.parameters()).device

    def _embed_pixels(self, backbone_feats: List[torch.Tensor], image_ids, encoder_hidden_states) -> torch.Tensor:
        """maskformer_segmentation.py:102-147: per-query copies of the backbone maps, the coarsest one replaced by the
        encoder's visual tokens, through the pixel decoder."""
        dev = self.device
        if not self.use_encoder_inputs:
            pixel_embed = self.pixel_decoder([f.to(dev) for f in backbone_feats])
            return pixel_embed.squeeze(0) if pixel_embed.shape[0] == 1 else pixel_embed[image_ids, ...]
        ids = image_ids.to(backbone_feats[0].device)
        if backbone_feats[0].shape[0] > 1:
            feats = [f[ids, ...].to(dev) for f in backbone_feats]
        else:
            feats = [f.clone() for f in backbone_feats]
        ref = backbone_feats[-1]
        tokens = encoder_hidden_states.permute(1, 2, 0)                       # [B, C, tokens]
        feats[-1] = tokens[..., : math.prod(ref.shape[-2:])].reshape(-1, *ref.shape[1:])
        return self.pixel_decoder(feats)

    def _predict(self, obj_queries, pixel_embed):
        return self.mask_predictor(obj_queries if self.aux_masks else obj_queries[-1], pixel_embed)

    def forward(self, backbone_feats: List[torch.Tensor], obj_queries: torch.Tensor, image_ids,
                encoder_hidden_states: Optional[torch.Tensor] = None, **kwargs) -> Dict[str, torch.Tensor]:
        if self.use_encoder_inputs:
            assert encoder_hidden_states is not None
        pixel_embed = self._embed_pixels(backbone_feats, image_ids, encoder_hidden_states)
        return {"pred_masks": self._predict(obj_queries, pixel_embed)}


class UniversalSegmentationHead(SegmentationHead):
    """Semantic + instance segmentation head (maskformer_segmentation.py:222-336)."""

    def __init__(self, hidden_dim, upsampling_stages, pixel_decoder, aux_masks=False, no_dec=False, act_ckpt=False,
                 presence_head: bool = False, dot_product_scorer=None, cross_attend_prompt=None):
        super().__init__(hidden_dim=hidden_dim, upsampling_stages=upsampling_stages, use_encoder_inputs=True,
                         aux_masks=aux_masks, no_dec=no_dec, pixel_decoder=pixel_decoder, act_ckpt=act_ckpt)
        self.d_model = hidden_dim
        if dot_product_scorer is not None:
            assert presence_head, "Specifying a dot product scorer without a presence head is likely a mistake"
        self.presence_head = None
        if presence_head:
            self.presence_head = dot_product_scorer if dot_product_scorer is not None else LinearPresenceHead(self.d_model)
        self.cross_attend_prompt = cross_attend_prompt
        if cross_attend_prompt is not None:
            self.cross_attn_norm = nn.LayerNorm(self.d_model)
        self.semantic_seg_head = nn.Conv2d(self.pixel_decoder.out_dim, 1, kernel_size=1)
        self.instance_seg_head = nn.Conv2d(self.pixel_decoder.out_dim, self.d_model, kernel_size=1)

    def forward(self, backbone_feats: List[torch.Tensor], obj_queries: torch.Tensor, image_ids,
                encoder_hidden_states: Optional[torch.Tensor] = None, prompt: Optional[torch.Tensor] = None,
                prompt_mask: Optional[torch.Tensor] = None, **kwargs) -> Dict[str, Optional[torch.Tensor]]:
        assert encoder_hidden_states is not None
        bs = encoder_hidden_states.shape[1]
        if self.cross_attend_prompt is not None:
            attended = self.cross_attend_prompt(query=self.cross_attn_norm(encoder_hidden_states), key=prompt, value=prompt,
                                                key_padding_mask=prompt_mask)[0]
            encoder_hidden_states = attended + encoder_hidden_states
        presence_logit = None
        if self.presence_head is not None:
            pooled = encoder_hidden_states.mean(0).view(1, bs, 1, self.d_model)
            presence_logit = self.presence_head(pooled, prompt=prompt, prompt_mask=prompt_mask).squeeze(0).squeeze(1)
        pixel_embed = self._embed_pixels(backbone_feats, image_ids, encoder_hidden_states)
        instance_embeds = conv_ops.conv1x1_forward(pixel_embed, self.instance_seg_head)
        return {
            "pred_masks": self._predict(obj_queries, instance_embeds),
            "semantic_seg": conv_ops.conv1x1_forward(pixel_embed, self.semantic_seg_head),
            "presence_logit": presence_logit,
        }
