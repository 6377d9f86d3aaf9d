"""Global configuration mirroring the reference's `CFG` class (allied_files.py:15-36).

The reference reads `CFG.device, CFG.max_len, CFG.pad_idx, CFG.bos_idx` INSIDE the hot path
(model.py:32,60,94,117; utils.py:8,29) and its scripts add attributes at import time
(train_val_epoch.py:26-27, inference_p.py:122); the same globals are honoured here.
"""
import torch


class CFG:
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")   # allied_files.py:16 hard-codes cuda:1 (Q14)

    max_len = 100            # allied_files.py:18
    img_size = 224           # :19
    num_bins = img_size      # :20
    batch_size = 64          # :24
    model_name = "deit3_medium_patch16_224.fb_in22k_ft_in1k"   # :27
    num_patches = 196        # :28
    generation_steps = 101   # :32 (unusable past max_len-1, Q6)

    # set by the reference's scripts from the tokenizer (train_val_epoch.py:26-27)
    pad_idx = 302
    bos_idx = 300

    # B200 path only: arithmetic of the GEMM/attention kernels ("bf16" fast path or "fp32" token-exact path)
    precision = "bf16"
    kv_page_tokens = 16
