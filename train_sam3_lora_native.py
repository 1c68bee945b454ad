#!/usr/bin/env python
"""Same command line as the reference's train_sam3_lora_native.py:

    python train_sam3_lora_native.py --config configs/full_lora_config.yaml
    torchrun --nproc-per-node 8 train_sam3_lora_native.py --config configs/full_lora_config.yaml
"""
from sam3_lora_b200.train_native import main

if __name__ == "__main__":
    main()
