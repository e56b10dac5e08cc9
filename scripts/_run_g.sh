python - <<'P'
import os, sys
sys.path.insert(0, "lightning-generative-models_b200")
import torch
from b200dm import _lib as L
for ctas in (1, 148):
    for mode in (1, 3, 4):
        out = torch.zeros(ctas, dtype=torch.int64, device="cuda")
        iters = 64 if mode < 3 else 14
        per = 8 if mode < 3 else 36
        for _ in range(2):
            L.call("b200dm_debug_umma_rate", 64, iters, mode, ctas, out.data_ptr())
        torch.cuda.synchronize()
        print(f"ctas {ctas} mode {mode}: {out.float().mean().item() / (iters * per):.1f} cycles/MMA  ({out.float().mean().item()/iters:.0f} cycles per group of {per})")
P
