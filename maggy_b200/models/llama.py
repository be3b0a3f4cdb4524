"""Llama-3 style decoder (RMSNorm, rotary embeddings, grouped-query attention, SwiGLU) on the maggy_b200
kernels — the model of BASELINE.json config 5 ("Llama-3 1B, param all-gather + GEMM fused").

Every projection is the tcgen05 GEMM (bias-free), norms are the fused RMSNorm kernel, the loss is the fused
softmax-CE. Rotary embedding, SiLU*up and attention (torch SDPA) are ATen/cuDNN ops in this round.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn

from ..ops import functional as F


@dataclass
class LlamaConfig:
    vocab_size: int = 128256
    dim: int = 2048
    n_layers: int = 16
    n_heads: int = 32
    n_kv_heads: int = 8
    ffn_dim: int = 8192
    max_seq_len: int = 2048
    rope_theta: float = 500000.0
    norm_eps: float = 1e-5
    tie_embeddings: bool = True
    # forward qkv and gate-up projections in fp8 (bf16 backward): True / "tensor" = e4m3 with per-tensor scales on
    # tcgen05 kind::f8f6f4; "mx" = block-scaled MXFP8 (e4m3 + one UE8M0 scale per 32 elements) on
    # tcgen05 kind::mxf8f6f4.block_scale
    fp8: object = False

    @classmethod
    def named(cls, name: str) -> "LlamaConfig":
        table = {
            "llama3-1b": dict(),
            "llama-tiny": dict(vocab_size=1024, dim=128, n_layers=2, n_heads=2, n_kv_heads=1, ffn_dim=256, max_seq_len=128),
            # small enough for tests, wide enough to shard over 2 or 4 ranks (rows per shard stay multiples of 128)
            "llama-small": dict(vocab_size=4096, dim=512, n_layers=4, n_heads=8, n_kv_heads=4, ffn_dim=1024,
                                max_seq_len=512),
        }
        return cls(**table[name])


def _rope_tables(seq_len, head_dim, theta, device):
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=device, dtype=torch.float32) / head_dim))
    ang = torch.outer(torch.arange(seq_len, device=device, dtype=torch.float32), inv)
    return ang.cos(), ang.sin()


def apply_rope(x, cos, sin):
    """x: [B, T, H, D]; rotate pairs (x[..., ::2], x[..., 1::2])."""
    x1, x2 = x[..., ::2].float(), x[..., 1::2].float()
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    out = torch.stack((x1 * c - x2 * s, x1 * s + x2 * c), dim=-1).flatten(-2)
    return out.to(x.dtype)


from .gpt2 import _attn_impl  # noqa: E402  (MAGGY_B200_ATTN / cfg.attn_impl, default "native")


class LlamaBlock(nn.Module):
    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.cfg = cfg
        hd = cfg.dim // cfg.n_heads
        self.attn_norm = nn.Parameter(torch.ones(cfg.dim))
        self.wqkv = nn.Parameter(torch.empty((cfg.n_heads + 2 * cfg.n_kv_heads) * hd, cfg.dim))
        self.wo = nn.Parameter(torch.empty(cfg.dim, cfg.dim))
        self.ffn_norm = nn.Parameter(torch.ones(cfg.dim))
        self.w13 = nn.Parameter(torch.empty(2 * cfg.ffn_dim, cfg.dim))   # gate and up fused into one GEMM
        self.w2 = nn.Parameter(torch.empty(cfg.dim, cfg.ffn_dim))

    def forward(self, x, cos, sin):
        cfg = self.cfg
        B, T, C = x.shape
        H, KV, hd = cfg.n_heads, cfg.n_kv_heads, C // cfg.n_heads
        h = F.rms_norm(x, self.attn_norm, cfg.norm_eps)
        lin = F.linear_mxfp8 if cfg.fp8 == "mx" else (F.linear_fp8 if cfg.fp8 else F.linear)
        sharded = getattr(self, "sharded_mx", False)   # ZeRO-3: projections are ShardedLinearMX modules (K3-MX)
        qkv = self.wqkv_s(h) if sharded else lin(h, self.wqkv)
        q, k, v = qkv.split([H * hd, KV * hd, KV * hd], dim=-1)
        q = apply_rope(q.view(B, T, H, hd), cos, sin)
        k = apply_rope(k.view(B, T, KV, hd), cos, sin)
        v = v.view(B, T, KV, hd)
        if hd == 64 and x.is_cuda and _attn_impl(cfg) == "native":
            # K10 reads q / k / v out of one fused [B, T, (H + 2 KV) * 64] buffer (grouped-query aware)
            y = F.attention(torch.cat([q.reshape(B, T, H * hd), k.reshape(B, T, KV * hd), v.reshape(B, T, KV * hd)],
                                      dim=-1), H, KV, causal=True)
        else:
            y = torch.nn.functional.scaled_dot_product_attention(
                q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True, enable_gqa=True)
            y = y.transpose(1, 2).reshape(B, T, C)
        if sharded:
            x = x + self.wo_s(y)
            h = F.rms_norm(x, self.ffn_norm, cfg.norm_eps)
            gate, up = self.w13_s(h).chunk(2, dim=-1)
            return x + self.w2_s(torch.nn.functional.silu(gate) * up)
        x = F.linear(y, self.wo, residual=x)
        h = F.rms_norm(x, self.ffn_norm, cfg.norm_eps)
        gate, up = lin(h, self.w13).chunk(2, dim=-1)
        x = F.linear(torch.nn.functional.silu(gate) * up, self.w2, residual=x)
        return x


class Llama(nn.Module):
    maggy_b200_native = True

    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.cfg = cfg
        self.padded_vocab = (cfg.vocab_size + 127) // 128 * 128
        self.tok = nn.Parameter(torch.empty(self.padded_vocab, cfg.dim))
        self.blocks = nn.ModuleList([LlamaBlock(cfg) for _ in range(cfg.n_layers)])
        self.norm = nn.Parameter(torch.ones(cfg.dim))
        self.lm_head = None if cfg.tie_embeddings else nn.Parameter(torch.empty(self.padded_vocab, cfg.dim))
        std = 0.02
        nn.init.normal_(self.tok, std=std)
        for b in self.blocks:
            for w in (b.wqkv, b.w13):
                nn.init.normal_(w, std=std)
            for w in (b.wo, b.w2):
                nn.init.normal_(w, std=std / (2 * cfg.n_layers) ** 0.5)
        if self.lm_head is not None:
            nn.init.normal_(self.lm_head, std=std)

    def num_params(self):
        return sum(p.numel() for p in self.parameters())

    def forward(self, idx, targets=None):
        cfg = self.cfg
        B, T = idx.shape
        x = F.embedding(idx, self.tok)
        cos, sin = _rope_tables(T, cfg.dim // cfg.n_heads, cfg.rope_theta, idx.device)
        for blk in self.blocks:
            x = blk(x, cos, sin)
        x = F.rms_norm(x, self.norm, cfg.norm_eps)
        logits = F.linear(x, self.tok if self.lm_head is None else self.lm_head)
        if targets is None:
            return logits
        return F.cross_entropy(logits, targets, v_real=cfg.vocab_size, reduction="mean_all")


def shard_llama_mx(model: "Llama", group=None) -> int:
    """ZeRO-3 + block-scaled fp8 for the decoder blocks (BASELINE.json config 5): every projection weight becomes a
    :class:`maggy_b200.parallel.fsdp.ShardedLinearMX` — each rank keeps ``rows / world`` rows as a bf16 master shard, the
    forward gathers the peers' MXFP8 images INSIDE the block-scaled GEMM kernel (K3-MX), the backward reduce-scatters
    the weight gradient with the peer-memory kernel. Embedding / norms stay replicated. Returns the number of sharded
    projections; call after the process group exists and the model sits on this rank's GPU."""
    from ..parallel.fsdp import ShardedLinearMX

    n = 0
    for blk in model.blocks:
        if not isinstance(blk, LlamaBlock):
            continue   # a LOCO trial replaced this block by Identity
        for name in ("wqkv", "wo", "w13", "w2"):
            w = getattr(blk, name)
            setattr(blk, name + "_s", ShardedLinearMX(w.detach(), None, group=group, device=w.device))
            delattr(blk, name)
            n += 1
        blk.sharded_mx = True
    return n
