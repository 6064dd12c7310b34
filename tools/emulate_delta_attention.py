"""CPU emulation (torch, bf16 operand rounding + fp32 accumulation) of the perturbation-form paired attention that
attention_perturb_kernel implements, against an fp64 evaluation of attn(e + d) - attn(e). Development tool: it fixes the
arithmetic (which quantities are rounded to bf16, the row-max shift of the score difference, expm1) before the CUDA
kernel is written, and shows where the independent-pair formulation (two bf16 forward passes, subtract) breaks down.

    python tools/emulate_delta_attention.py [tiny|small|base]
"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import vit_oracle as O  # noqa: E402

bf = lambda t: t.to(torch.bfloat16).to(torch.float32)


def heads(t, h):
    n, l, e = t.shape
    return t.view(n, l, h, e // h).transpose(1, 2)


def expm1_kernel(w):
    """what the kernel evaluates: a 4th-order Taylor polynomial below 1/8, exp2 - 1 above"""
    poly = w * (1 + w * (0.5 + w * (1 / 6 + w * (1 / 24))))
    big = torch.exp2(w * 1.4426950408889634) - 1
    return torch.where(w > -0.125, poly, big)


def delta_path(e32, d32, wqkv, bqkv, wo, h):
    """e32: fp32 tokens of x, d32: fp32 token difference. Returns ||W_o (attn_core(b) - attn_core(a))|| per sample."""
    E = e32.shape[-1]
    w16 = bf(wqkv)
    qkv_a = bf(bf(e32) @ w16.T + bqkv)       # GEMM epilogue rounds to bf16
    dqkv = bf(bf(d32) @ w16.T)
    qa, ka, va = (heads(t, h) for t in qkv_a.chunk(3, -1))
    dq, dk, dv = (heads(t, h) for t in dqkv.chunk(3, -1))
    S = qa @ ka.transpose(-1, -2)
    dS = qa @ dk.transpose(-1, -2) + dq @ ka.transpose(-1, -2) + dq @ dk.transpose(-1, -2)
    c = 0.125
    m = S.max(-1, keepdim=True).values
    pa = torch.exp((S - m) * c)
    la = pa.sum(-1, keepdim=True)
    mw = dS.max(-1, keepdim=True).values
    g = pa * expm1_kernel((dS - mw) * c)
    dl = g.sum(-1, keepdim=True)
    lb = (pa + g).sum(-1, keepdim=True)
    H = bf((g - (dl / la) * pa) / lb)  # P_b - P_a formed in fp32, then ONE bf16 rounding
    Pa = bf(pa / la)
    D = H @ va + H @ dv + Pa @ dv
    D = bf(D).transpose(1, 2).reshape(e32.shape)
    out = D @ bf(wo).T
    return (out ** 2).flatten(1).sum(-1).sqrt()


def pair_path(e32, d32, wqkv, bqkv, wo, h):
    """the round-1 formulation: two independent bf16 evaluations, outputs subtracted in fp32"""
    w16 = bf(wqkv)

    def core(t):
        qkv = bf(bf(t) @ w16.T + bqkv)
        q, k, v = (heads(x, h) for x in qkv.chunk(3, -1))
        S = q @ k.transpose(-1, -2) * 0.125
        p = torch.exp(S - S.max(-1, keepdim=True).values)
        return (bf(p) @ v) / p.sum(-1, keepdim=True)

    D = bf(core(e32 + d32) - core(e32)).transpose(1, 2).reshape(e32.shape)
    return ((D @ bf(wo).T) ** 2).flatten(1).sum(-1).sqrt()


def truth(e, d, wqkv, bqkv, wo, h):
    e, d, wqkv, bqkv, wo = (t.double() for t in (e, d, wqkv, bqkv, wo))

    def core(t):
        q, k, v = (heads(x, h) for x in (t @ wqkv.T + bqkv).chunk(3, -1))
        return torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v

    D = (core(e + d) - core(e)).transpose(1, 2).reshape(e.shape)
    return ((D @ wo.T) ** 2).flatten(1).sum(-1).sqrt()


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "small"
    arch = {"tiny": O.Arch(emb_dim=128, n_heads=2, n_layers=2, ffn_dim=512, image_dim=(3, 32, 32)),
            "small": O.Arch(emb_dim=256, n_heads=4, n_layers=3, ffn_dim=1024, image_dim=(3, 64, 64)),
            "base": O.vit_arch("base")}[name]
    sd = O.init_state_dict(arch, 42)
    n = 2
    x, noise = O.synthetic_images(n, arch, 10), O.synthetic_images(n, arch, 12)
    e = O.embedding(sd, x, arch)
    P = arch.patch_size
    w = sd["embedding.patching.patching.0.weight"]
    dn = torch.nn.functional.conv2d(bf(noise), bf(w), None, stride=P).flatten(2).transpose(1, 2)
    dn = torch.cat((torch.zeros(n, 1, arch.emb_dim), dn), 1)  # cls row of the difference is zero
    for eps in (10.0, 1.0, 0.3, 0.1, 1e-2, 1e-3, 1e-4):
        worst_d = worst_p = 0.0
        for i in range(min(arch.n_layers, 3)):
            b = f"blocks.{i}.attn."
            args = (sd[b + "qkv_mat.weight"], sd[b + "qkv_mat.bias"], sd[b + "output.weight"], arch.n_heads)
            t = truth(e, eps * dn, *args)
            worst_d = max(worst_d, float(((delta_path(e, eps * dn, *args).double() - t) / t).abs().max()))
            worst_p = max(worst_p, float(((pair_path(e, eps * dn, *args).double() - t) / t).abs().max()))
        print(f"{name} eps {eps:8.0e}: delta form rel err {worst_d:.2e}   independent-pair form rel err {worst_p:.2e}")


if __name__ == "__main__":
    main()
