"""sam3_lora_b200 — B200-native hot path behind the sam3_lora LoRA-training surface.

Only what the path needs lives here: `csrc/` (sm_100a CUDA kernels + the C ABI of
libsam3b.so) and the host-side mirror of the reference's LoRA / ViT interface.
"""
__version__ = "0.1.0"
