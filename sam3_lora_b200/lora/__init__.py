"""`sam3_lora.lora` surface (sam3_lora/lora/__init__.py) on the fused sm_100a LoRA-Linear op."""
from .lora_layer import LinearWithLoRA, LoRALayer
from .lora_utils import (LoRAConfig, get_lora_parameters, get_lora_state_dict, inject_lora_into_model, load_lora_state_dict,
                         merge_lora_weights, print_trainable_parameters)

__all__ = ["LoRALayer", "LinearWithLoRA", "LoRAConfig", "inject_lora_into_model", "get_lora_parameters", "get_lora_state_dict",
           "load_lora_state_dict", "merge_lora_weights", "print_trainable_parameters"]
