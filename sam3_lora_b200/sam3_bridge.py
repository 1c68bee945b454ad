"""Row a9: the SAM3 detector around the native hot path.

`Sam3Image.forward` / `forward_grounding` (sam3/model/sam3_image.py:442-576) is orchestration and stays the
reference's own Python: this module imports the reference's `sam3` package (from `SAM3_REFERENCE_ROOT`, from
`baseline/_ref` next to this repo, or from wherever `import sam3` already resolves), builds the model with
the reference's `build_sam3_image_model`, and swaps the hot-path modules in place for the native ones:

    backbone.vision_backbone.trunk      sam3.model.vitdet.ViT                 -> sam3_lora_b200.vit.ViT
    backbone.vision_backbone (neck)     sam3.model.necks.Sam3DualViTDetNeck   -> sam3_lora_b200.necks.Sam3DualViTDetNeck
    every nn.MultiheadAttention         (DETR encoder / decoder, seg head,    -> sam3_lora_b200.mha.MultiheadAttention
                                         geometry encoder, text tower)
    segmentation_head.pixel_decoder     maskformer_segmentation.PixelDecoder  -> sam3_lora_b200.maskformer_segmentation.PixelDecoder

Parameter names do not change, so `sam3.pt` and adapter checkpoints load into the swapped model unchanged.
Nothing here computes: the arithmetic is either the reference's eager PyTorch (the parts outside SURVEY §8) or
libsam3b.so.  Third-party packages the reference imports at module scope but that this image lacks
(iopath, timm, ftfy, decord, pycocotools, torchmetrics) are given import-time stand-ins ONLY when the real package is
missing; with the real packages installed nothing is touched.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.util
import os
import sys
import types
from pathlib import Path
from typing import Dict, Iterable, Optional

import torch
import torch.nn as nn

_REPO = Path(__file__).resolve().parents[1]


# ------------------------------------------------------------------------------------------------
# locating / importing the reference
# ------------------------------------------------------------------------------------------------
def reference_root() -> Optional[Path]:
    """Directory that holds the reference's `sam3/` package, or None when `import sam3` must resolve by itself."""
    env = os.environ.get("SAM3_REFERENCE_ROOT")
    for cand in ([Path(env)] if env else []) + [_REPO / "baseline" / "_ref"]:
        if (cand / "sam3" / "model_builder.py").exists():
            return cand
    return None


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install_shims() -> Dict[str, bool]:
    """Stand-ins for third-party imports the reference makes at module scope; installed only for missing packages.
    Returns {package: shimmed?}."""
    done = {}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__sam3b_shim__ = True
        sys.modules[name] = m
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(sys.modules[parent], leaf, m)
        return m

    done["timm"] = _missing("timm")
    if done["timm"]:
        class DropPath(nn.Module):
            """Per-sample stochastic depth (timm.layers.DropPath, scale_by_keep)."""

            def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
                super().__init__()
                self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

            def forward(self, x):
                if self.drop_prob == 0.0 or not self.training:
                    return x
                keep = 1 - self.drop_prob
                mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
                if keep > 0.0 and self.scale_by_keep:
                    mask.div_(keep)
                return x * mask

        class Mlp(nn.Module):
            """fc1 -> act -> drop1 -> norm -> fc2 -> drop2 (timm.layers.Mlp attribute names)."""

            def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None,
                         bias=True, drop=0.0, use_conv=False):
                super().__init__()
                out_features = out_features or in_features
                hidden_features = hidden_features or in_features
                drops = drop if isinstance(drop, tuple) else (drop, drop)
                self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
                self.act = act_layer()
                self.drop1 = nn.Dropout(drops[0])
                self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
                self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
                self.drop2 = nn.Dropout(drops[1])

            def forward(self, x):
                return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))

        mod("timm")
        layers = dict(DropPath=DropPath, Mlp=Mlp, trunc_normal_=nn.init.trunc_normal_)
        mod("timm.layers", **layers)
        mod("timm.models")
        mod("timm.models.layers", **layers)

    done["iopath"] = _missing("iopath")
    if done["iopath"]:
        class PathManager:
            def open(self, path, mode="r", **kw):
                return open(path, mode, **kw)

            def exists(self, path):
                return os.path.exists(path)

            def isfile(self, path):
                return os.path.isfile(path)

            def isdir(self, path):
                return os.path.isdir(path)

            def ls(self, path):
                return os.listdir(path)

            def mkdirs(self, path):
                os.makedirs(path, exist_ok=True)

            def get_local_path(self, path, **kw):
                return path

            def register_handler(self, *a, **kw):
                pass

        mod("iopath")
        mod("iopath.common")
        mod("iopath.common.file_io", PathManager=PathManager, g_pathmgr=PathManager())

    done["ftfy"] = _missing("ftfy")
    if done["ftfy"]:
        mod("ftfy", fix_text=lambda s: s)

    done["decord"] = _missing("decord")
    if done["decord"]:
        def _no_video(*a, **kw):
            raise RuntimeError("decord is not installed (video decoding is outside the image training path)")

        mod("decord", cpu=_no_video, VideoReader=_no_video)

    done["pycocotools"] = _missing("pycocotools")
    if done["pycocotools"]:
        def _no_coco(*a, **kw):
            raise RuntimeError("pycocotools is not installed; the native data path (sam3_lora_b200.data) decodes masks itself")

        m = mod("pycocotools")
        msk = mod("pycocotools.mask", decode=_no_coco, encode=_no_coco, frPyObjects=_no_coco, area=_no_coco, toBbox=_no_coco,
                  merge=_no_coco, iou=_no_coco)
        mod("pycocotools._mask", decode=_no_coco, encode=_no_coco)
        mod("pycocotools.coco", COCO=_no_coco)
        mod("pycocotools.cocoeval", COCOeval=_no_coco)
        m.mask = msk

    done["torchmetrics"] = _missing("torchmetrics")
    if done["torchmetrics"]:
        class Metric(nn.Module):
            def __init__(self, *a, **kw):
                super().__init__()

            def add_state(self, name, default, dist_reduce_fx=None):
                setattr(self, name, default)

        def f1_score(preds, target, task="binary", **kw):
            preds, target = preds.bool(), target.bool()
            tp = (preds & target).sum().float()
            den = preds.sum().float() + target.sum().float()
            return torch.where(den > 0, 2 * tp / den.clamp_min(1), torch.zeros_like(tp))

        tm = mod("torchmetrics", Metric=Metric)
        fn = mod("torchmetrics.functional", f1_score=f1_score)
        tm.functional = fn
    return done


_imported = None


def import_reference():
    """Returns the reference's `sam3.model_builder` module (imported once)."""
    global _imported
    if _imported is not None:
        return _imported
    root = reference_root()
    if root is not None and str(root) not in sys.path:
        sys.path.insert(0, str(root))
    install_shims()
    try:
        _imported = importlib.import_module("sam3.model_builder")
    except ImportError as e:  # noqa: PERF203
        raise ImportError(
            "the SAM3 detector wiring (row a9) is the reference's own Python: put Sompote/sam3_lora on sys.path, set "
            "SAM3_REFERENCE_ROOT, or install it into baseline/_ref (tools/install_reference.sh)") from e
    return _imported


def default_bpe_path() -> str:
    mb = import_reference()
    here = Path(mb.__file__).resolve().parent
    for cand in (here / "assets" / "bpe_simple_vocab_16e6.txt.gz", here.parent / "assets" / "bpe_simple_vocab_16e6.txt.gz"):
        if cand.exists():
            return str(cand)
    raise FileNotFoundError("bpe_simple_vocab_16e6.txt.gz not found next to the reference's sam3 package")


@contextlib.contextmanager
def cpu_compat():
    """The reference hard-codes device="cuda" in a few constructors (position_encoding.py:47, decoder.py:281) and calls
    Tensor.pin_memory in the geometry encoder's forward (geometry_encoders.py:659); on a machine WITHOUT a GPU (golden
    generation, CPU tests) those are routed to the CPU while this context is active.  No-op when CUDA is available."""
    if torch.cuda.is_available():
        yield
        return
    names = ("zeros", "arange", "ones", "empty", "tensor", "linspace")
    saved = {n: getattr(torch, n) for n in names}

    def wrap(fn):
        def inner(*a, **kw):
            if str(kw.get("device", "")).startswith("cuda"):
                kw["device"] = "cpu"
            return fn(*a, **kw)
        return inner

    pin = torch.Tensor.pin_memory
    try:
        for n in names:
            setattr(torch, n, wrap(saved[n]))
        torch.Tensor.pin_memory = lambda self, *a, **kw: self
        yield
    finally:
        for n in names:
            setattr(torch, n, saved[n])
        torch.Tensor.pin_memory = pin


def seed_parameters(model: nn.Module, seed: int = 0) -> None:
    """Deterministic synthetic weights (no checkpoint is reachable offline): every floating parameter ~ N(0, 0.02) from
    its own generator keyed by (seed, name) — so sub-module swaps do not shift the stream — except 1-D norm-like weights
    (= 1 + N(0, 0.02)).  The two text-tower parameters the reference leaves uninitialised (text_encoder_ve.py:196,218)
    get values as well."""
    import zlib

    with torch.no_grad():
        for name, p in model.named_parameters():
            if not p.is_floating_point():
                continue
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
            v = torch.randn(p.shape, generator=g, dtype=torch.float32) * 0.02
            if p.dim() == 1 and name.endswith("weight") and ("norm" in name or "ln_" in name or ".ln" in name):
                v += 1.0
            p.copy_(v.to(p.dtype))


def build_reference_model(device: str = "cpu", eval_mode: bool = False, checkpoint_path: Optional[str] = None,
                          seed: Optional[int] = 0, enable_segmentation: bool = True):
    """The UNMODIFIED reference model (`build_sam3_image_model`, sam3/model_builder.py:558-641).  `checkpoint_path=None`
    means seeded synthetic weights (`seed`), never a download."""
    mb = import_reference()
    with cpu_compat():
        model = mb.build_sam3_image_model(bpe_path=default_bpe_path(), device="cpu", eval_mode=eval_mode,
                                          checkpoint_path=checkpoint_path, load_from_HF=False,
                                          enable_segmentation=enable_segmentation, compile=False)
    if checkpoint_path is None and seed is not None:
        seed_parameters(model, seed)
    if device != "cpu":
        model = model.to(device)
    return model


# ------------------------------------------------------------------------------------------------
# module surgery
# ------------------------------------------------------------------------------------------------
def _copy_state(dst: nn.Module, src: nn.Module, what: str) -> None:
    missing, unexpected = dst.load_state_dict(src.state_dict(), strict=False)
    missing = [k for k in missing if "freqs_cis" not in k]
    if missing or unexpected:
        raise RuntimeError(f"{what}: state dict mismatch, missing={missing[:5]} unexpected={unexpected[:5]}")


def swap_trunk(model: nn.Module, max_batch: int = 8, operand_dtype=torch.float16, cuda_graphs: bool = False) -> nn.Module:
    """backbone.vision_backbone.trunk -> sam3_lora_b200.vit.ViT with the same weights (vitdet.py:813-859)."""
    from .vit import ViT  # noqa: PLC0415

    neck = model.backbone.vision_backbone
    ref = neck.trunk
    if isinstance(ref, ViT):
        return model
    blk0 = ref.blocks[0]
    patch = ref.patch_embed.proj.kernel_size[0]
    grid = int(round(ref.blocks[ref.full_attn_ids[0]].attn.freqs_cis.shape[0] ** 0.5))     # global blocks: rope over the full grid
    pos_side = int(round((ref.pos_embed.shape[1] - 1) ** 0.5))                             # cls slot + pretrain grid (vitdet.py:715-731)
    native = ViT(img_size=patch * grid, patch_size=patch, in_chans=ref.patch_embed.proj.in_channels,
                 embed_dim=blk0.attn.qkv.in_features, depth=len(ref.blocks), num_heads=blk0.attn.num_heads,
                 mlp_ratio=blk0.mlp.fc1.out_features / blk0.attn.qkv.in_features, window_size=blk0.window_size,
                 global_att_blocks=tuple(ref.full_attn_ids), pretrain_img_size=pos_side * patch,
                 ln_eps=blk0.norm1.eps, drop_path_rate=_drop_path_rate(ref), operand_dtype=operand_dtype, max_batch=max_batch,
                 cuda_graphs=cuda_graphs)
    _copy_state(native, ref, "trunk")
    native.train(ref.training)
    for p_new, p_old in zip(native.parameters(), ref.parameters()):
        p_new.requires_grad = p_old.requires_grad
    native.to(next(ref.parameters()).device)
    neck.trunk = native
    return model


def _drop_path_rate(ref_vit) -> float:
    last = ref_vit.blocks[-1].drop_path
    return float(getattr(last, "drop_prob", 0.0))


def swap_neck(model: nn.Module) -> nn.Module:
    """backbone.vision_backbone -> sam3_lora_b200.necks.Sam3DualViTDetNeck around the SAME trunk / position encoding."""
    from .necks import Sam3DualViTDetNeck  # noqa: PLC0415

    ref = model.backbone.vision_backbone
    if isinstance(ref, Sam3DualViTDetNeck):
        return model
    if getattr(ref, "sam2_convs", None) is not None:
        raise RuntimeError("swap_neck: the SAM2 side neck (enable_inst_interactivity) is outside the training path")
    native = Sam3DualViTDetNeck(ref.trunk, ref.position_encoding, d_model=ref.convs[0].conv_1x1.out_channels,
                                scale_factors=tuple(ref.scale_factors))
    _copy_state(native.convs, ref.convs, "neck")
    for p_new, p_old in zip(native.convs.parameters(), ref.convs.parameters()):
        p_new.requires_grad = p_old.requires_grad
    native.train(ref.training)
    native.to(next(ref.convs.parameters()).device)
    model.backbone.vision_backbone = native
    return model


def swap_pixel_decoder(model: nn.Module) -> nn.Module:
    """segmentation_head.pixel_decoder -> the native PixelDecoder (maskformer_segmentation.py:172-219)."""
    from .maskformer_segmentation import PixelDecoder  # noqa: PLC0415

    head = getattr(model, "segmentation_head", None)
    if head is None or isinstance(head.pixel_decoder, PixelDecoder):
        return model
    ref = head.pixel_decoder
    native = PixelDecoder(hidden_dim=ref.conv_layers[0].in_channels, num_upsampling_stages=len(ref.conv_layers),
                          interpolation_mode=ref.interpolation_mode, shared_conv=ref.shared_conv)
    _copy_state(native, ref, "pixel decoder")
    for p_new, p_old in zip(native.parameters(), ref.parameters()):
        p_new.requires_grad = p_old.requires_grad
    native.train(ref.training)
    native.to(next(ref.parameters()).device)
    head.pixel_decoder = native
    return model


def swap_mha(model: nn.Module, skip: Iterable[str] = ()) -> int:
    """Every nn.MultiheadAttention (incl. the reference's MultiheadAttentionWrapper, model_misc.py:31-34) -> the native
    module; returns how many were replaced."""
    from .mha import replace_torch_mha  # noqa: PLC0415

    return replace_torch_mha(model, skip=tuple(skip))


def swap_matcher(model: nn.Module) -> nn.Module:
    """model.matcher (the in-forward Hungarian pass, sam3_image.py:578-581; SciPy + 3 host syncs per call in the reference,
    sam3/train/matcher.py:539-617) -> the GPU-resident matcher with the same cost weights."""
    from .matcher import BinaryHungarianMatcherV2  # noqa: PLC0415

    ref = getattr(model, "matcher", None)
    if ref is None or isinstance(ref, BinaryHungarianMatcherV2):
        return model
    if type(ref).__name__ != "BinaryHungarianMatcherV2":
        raise RuntimeError(f"swap_matcher: unexpected matcher {type(ref).__name__}")
    norm = getattr(ref, "norm", None)
    focal = bool(getattr(ref, "focal", type(norm).__name__ == "Sigmoid" if norm is not None else True))
    model.matcher = BinaryHungarianMatcherV2(cost_class=ref.cost_class, cost_bbox=ref.cost_bbox, cost_giou=ref.cost_giou, focal=focal,
                                             alpha=getattr(ref, "alpha", 0.25), gamma=getattr(ref, "gamma", 2.0),
                                             stable=bool(getattr(ref, "stable", False)))
    return model


def disable_activation_checkpointing() -> int:
    """The reference recomputes every DETR encoder / decoder / geometry layer in the backward (`activation_ckpt_wrapper`,
    sam3/model/act_ckpt_utils.py:17-114, enabled by sam3/model_builder.py:146,184) to fit 24-32 GB cards.  With 180 GB of HBM
    the activations of a batch-8 step fit several times over, so the swapped model keeps them instead (the same decision as
    the trunk's): the wrapper is replaced, in the modules that imported it, by a direct call.  Outputs and
    gradients are unchanged.  Returns the number of module namespaces patched."""
    import_reference()

    def direct(module):
        def call(*args, act_ckpt_enable: bool = True, use_reentrant: bool = False, **kwargs):
            return module(*args, **kwargs)
        return call

    n = 0
    for name in ("sam3.model.encoder", "sam3.model.decoder", "sam3.model.geometry_encoders"):
        mod = sys.modules.get(name) or importlib.import_module(name)
        if hasattr(mod, "activation_ckpt_wrapper"):
            if not hasattr(mod, "_sam3b_ref_ckpt_wrapper"):
                mod._sam3b_ref_ckpt_wrapper = mod.activation_ckpt_wrapper
            mod.activation_ckpt_wrapper = direct
            n += 1
    return n


def restore_activation_checkpointing() -> None:
    for name in ("sam3.model.encoder", "sam3.model.decoder", "sam3.model.geometry_encoders"):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "_sam3b_ref_ckpt_wrapper"):
            mod.activation_ckpt_wrapper = mod._sam3b_ref_ckpt_wrapper


def build_native_model(device="cuda", *, checkpoint_path: Optional[str] = None, seed: Optional[int] = 0, max_batch: int = 8,
                       operand_dtype=torch.float16, cuda_graphs: bool = False, eval_mode: bool = False,
                       parts: Iterable[str] = ("trunk", "neck", "pixel_decoder", "mha", "matcher", "no_recompute"), reference_model: Optional[nn.Module] = None):
    """Reference `Sam3Image` with the hot-path modules swapped for the native ones.  `parts` selects which."""
    model = reference_model if reference_model is not None else build_reference_model(
        "cpu", eval_mode=eval_mode, checkpoint_path=checkpoint_path, seed=seed)
    parts = set(parts)
    if "trunk" in parts:
        swap_trunk(model, max_batch=max_batch, operand_dtype=operand_dtype, cuda_graphs=cuda_graphs)
    if "neck" in parts:
        swap_neck(model)
    if "pixel_decoder" in parts:
        swap_pixel_decoder(model)
    if "mha" in parts:
        swap_mha(model)
    if "matcher" in parts:
        swap_matcher(model)
    if "no_recompute" in parts:
        disable_activation_checkpointing()
    return model.to(device)
