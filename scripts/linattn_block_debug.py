"""Debug aid for csrc/linattn_tc.cu: runs one case of the fused LinearAttention block and compares the intermediate
workspace (row maxima, softmax denominators, contexts) and the output with a torch restatement.

    python scripts/linattn_block_debug.py [--B 2 --S 16 --C 64]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--S", type=int, default=16)
    ap.add_argument("--C", type=int, default=64)
    a = ap.parse_args()
    import test_kernels_gpu as T
    from b200dm import _lib as L
    y, ref, ws, d, (xv, wq, wo, sd) = T._linattn_block_case(a.B, a.S, a.C, seed=1)
    B, n, C = a.B, a.S * a.S, a.C
    split = (ws.numel() - B * n - B * C * 64) // (B * 4736)
    w = ws[:B * split * 4736].view(B, split, 4736)
    rn = ws[B * split * 4736:B * split * 4736 + B * n]
    x = xv.to_nchw().float()
    xn = torch.nn.functional.normalize(x, dim=1) * sd["a.norm.g"] * C ** 0.5
    rn_ref = 1.0 / x.permute(0, 2, 3, 1).reshape(B * n, C).norm(dim=1)
    print("rn rel err", ((rn - rn_ref).abs().max() / rn_ref.abs().max()).item())
    qkv = torch.nn.functional.conv2d(xn, sd["a.to_qkv.weight"])
    q, k, v = (t.reshape(B, 128, n) for t in qkv.chunk(3, dim=1))
    kmax = w[:, :, :128].max(dim=1).values
    print("split", split, "kmax err", (kmax - k.max(dim=2).values).abs().max().item(), "ref scale", k.abs().max().item())
    mk, mv = sd["a.mem_kv"][0].reshape(128, 4), sd["a.mem_kv"][1].reshape(128, 4)
    m = torch.maximum(k.max(dim=2).values, mk.max(dim=1).values[None])
    p = torch.exp(k - m[:, :, None])
    s_ref = p.sum(dim=2)
    s = w[:, :, 128:640].reshape(B, split, 4, 128).sum(dim=(1, 2))
    print("s rel err", ((s - s_ref).abs().max() / s_ref.abs().max()).item())
    ctx_ref = torch.einsum("bhdn,bhen->bhed", p.view(B, 4, 32, n), v.reshape(B, 4, 32, n))
    ctx = w[:, :, 640:].reshape(B, split, 4, 32, 32).sum(dim=1)
    print("ctx rel err", ((ctx - ctx_ref).abs().max() / ctx_ref.abs().max()).item())
    print("y rel err", ((y - ref).norm() / ref.norm()).item(), "finite", bool(torch.isfinite(y).all()))
    bad = (y - ref).abs().amax(dim=(1,))        # [B, S, S]
    print("worst pixels", bad.flatten().topk(4).indices.tolist(), "max abs", bad.max().item())


if __name__ == "__main__":
    main()
