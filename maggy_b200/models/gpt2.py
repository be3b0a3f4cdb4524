"""GPT-2 (decoder-only transformer) on the hand-written sm_100a kernels.

The flagship model of the distributed-training config in BASELINE.json ("GPT-2 small bf16
data-parallel 8xB200, fused reduce-scatter+Adam"). Architecture = the public GPT-2: learned position
embeddings, pre-LN blocks, tanh-GELU MLP, tied input/output embedding. In the reference this model
would be the user's ``module`` handed to ``TorchDistributedConfig`` and run through cuBLAS/cuDNN
(maggy/core/executors/torch_dist_executor.py:145-159); here every GEMM is the tcgen05 kernel with
fused bias/GELU/dGELU/residual epilogues, LayerNorm / cross-entropy / embedding / AdamW are the
bandwidth kernels in csrc/ops, and wgrads are written straight into the flat gradient arena.
Attention itself uses torch's fused SDPA (library flash attention) in this round.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn

from ..ops import functional as F


@dataclass
class GPT2Config:
    vocab_size: int = 50257
    block_size: int = 1024
    n_layer: int = 12
    n_head: int = 12
    n_embd: int = 768
    padded_vocab_size: int = 0  # 0 -> round vocab up to a multiple of 128 (tile / vector friendly)
    layer_norm_eps: float = 1e-5
    attn_impl: str = ""  # "" -> MAGGY_B200_ATTN or "native" (maggy_b200 tcgen05 attention, K10); "sdpa" = torch SDPA

    def __post_init__(self):
        if self.padded_vocab_size == 0:
            self.padded_vocab_size = (self.vocab_size + 127) // 128 * 128

    @classmethod
    def named(cls, name: str) -> "GPT2Config":
        table = {
            # head_dim 64 like every real GPT-2 size: the tiny config runs the native attention kernels too
            "gpt2-tiny": dict(n_layer=2, n_head=2, n_embd=128, block_size=128, vocab_size=1000),
            "gpt2-small": dict(n_layer=12, n_head=12, n_embd=768),
            "gpt2-medium": dict(n_layer=24, n_head=16, n_embd=1024),
            "gpt2-large": dict(n_layer=36, n_head=20, n_embd=1280),
            "gpt2-xl": dict(n_layer=48, n_head=25, n_embd=1600),
        }
        return cls(**table[name])


def _attn_impl(cfg) -> str:
    """"native" (default): the maggy_b200 tcgen05 attention kernels (K10, csrc/ops/attention_v2.cu) — no library kernel
    is left in the training step. "sdpa": torch SDPA (cuDNN flash attention), kept as a comparison arm only
    (MAGGY_B200_ATTN=sdpa / cfg.attn_impl="sdpa"); numbers for both in profiles/attention.md."""
    import os

    return getattr(cfg, "attn_impl", None) or os.environ.get("MAGGY_B200_ATTN", "native")


class Block(nn.Module):
    def __init__(self, cfg: GPT2Config):
        super().__init__()
        C = cfg.n_embd
        self.n_head = cfg.n_head
        self.eps = cfg.layer_norm_eps
        self.attn_impl = _attn_impl(cfg)
        self.fold_residual_grad = __import__("os").environ.get("MAGGY_B200_LN_RES", "1") != "0"
        # bias gradients of proj / fc2 come out of the norm backward that consumes their output (no colsum pass)
        self.fold_bias_grad = self.fold_residual_grad and __import__("os").environ.get("MAGGY_B200_LN_COL", "1") != "0"
        self.ln1_w = nn.Parameter(torch.ones(C))
        self.ln1_b = nn.Parameter(torch.zeros(C))
        self.qkv_w = nn.Parameter(torch.empty(3 * C, C))
        self.qkv_b = nn.Parameter(torch.zeros(3 * C))
        self.proj_w = nn.Parameter(torch.empty(C, C))
        self.proj_b = nn.Parameter(torch.zeros(C))
        self.ln2_w = nn.Parameter(torch.ones(C))
        self.ln2_b = nn.Parameter(torch.zeros(C))
        self.fc_w = nn.Parameter(torch.empty(4 * C, C))
        self.fc_b = nn.Parameter(torch.zeros(4 * C))
        self.fc2_w = nn.Parameter(torch.empty(C, 4 * C))
        self.fc2_b = nn.Parameter(torch.zeros(C))

    def forward(self, x, prev_bias=None):
        """``prev_bias``: bias of the layer that produced ``x`` (the previous block's ``fc2_b``), or None."""
        B, T, C = x.shape
        H = self.n_head
        # x_res aliases x: routing the residual through the norm op folds the residual-gradient add into its backward
        fold = self.fold_residual_grad
        colb = self.fold_bias_grad
        h, x_res = F.layer_norm(x, self.ln1_w, self.ln1_b, self.eps, with_residual=True,
                                col_bias=prev_bias if colb else None) if fold else \
            (F.layer_norm(x, self.ln1_w, self.ln1_b, self.eps), x)
        qkv = F.linear(h, self.qkv_w, self.qkv_b)
        if self.attn_impl == "native" and C // H == 64 and qkv.is_cuda:
            # K10: reads q/k/v straight out of the fused projection, writes [B, T, C]; no split / transpose / cat
            y = F.attention(qkv, H, causal=True)
        elif self.attn_impl == "math":
            # unfused reference arm (matmul, mask, softmax, matmul): what the tests compare the kernels against
            q, k, v = (t.transpose(1, 2) for t in qkv.view(B, T, 3, H, C // H).unbind(dim=2))
            att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(C // H))
            mask = torch.ones(T, T, dtype=torch.bool, device=x.device).tril()
            att = att.masked_fill(~mask, float("-inf")).float().softmax(dim=-1).to(q.dtype)
            y = (att @ v).transpose(1, 2).reshape(B, T, C)
        else:
            q, k, v = qkv.view(B, T, 3, H, C // H).unbind(dim=2)
            y = torch.nn.functional.scaled_dot_product_attention(
                q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True)
            y = y.transpose(1, 2).reshape(B, T, C)
        x = F.linear(y, self.proj_w, self.proj_b, residual=x_res)   # x + attn (residual fused)
        h, x_res = F.layer_norm(x, self.ln2_w, self.ln2_b, self.eps, with_residual=True,
                                col_bias=self.proj_b if colb else None) if fold else \
            (F.layer_norm(x, self.ln2_w, self.ln2_b, self.eps), x)
        x = F.mlp(h, self.fc_w, self.fc_b, self.fc2_w, self.fc2_b, residual=x_res)  # x + mlp
        return x


class GPT2(nn.Module):
    maggy_b200_native = True

    def __init__(self, cfg: GPT2Config):
        super().__init__()
        self.cfg = cfg
        C = cfg.n_embd
        self.wte = nn.Parameter(torch.empty(cfg.padded_vocab_size, C))
        self.wpe = nn.Parameter(torch.empty(cfg.block_size, C))
        self.blocks = nn.ModuleList([Block(cfg) for _ in range(cfg.n_layer)])
        self.lnf_w = nn.Parameter(torch.ones(C))
        self.lnf_b = nn.Parameter(torch.zeros(C))
        self.reset_parameters()

    def reset_parameters(self):
        cfg = self.cfg
        std = 0.02
        resid_std = std / math.sqrt(2 * cfg.n_layer)
        nn.init.normal_(self.wte, std=std)
        nn.init.normal_(self.wpe, std=std)
        for b in self.blocks:
            nn.init.normal_(b.qkv_w, std=std)
            nn.init.normal_(b.fc_w, std=std)
            nn.init.normal_(b.proj_w, std=resid_std)
            nn.init.normal_(b.fc2_w, std=resid_std)

    def num_params(self) -> int:
        return sum(p.numel() for p in self.parameters())

    def flops_per_token(self, T: int) -> float:
        """Training FLOPs per token (fwd+bwd = 3x fwd), matmuls + causal-free attention term."""
        cfg = self.cfg
        C, L = cfg.n_embd, cfg.n_layer
        mm = L * (2 * C * 3 * C + 2 * C * C + 2 * 2 * C * 4 * C) + 2 * C * cfg.padded_vocab_size
        attn = L * 2 * 2 * T * C
        return 3.0 * (mm + attn)

    def forward(self, idx, targets=None):
        x = F.embedding(idx, self.wte, self.wpe)
        prev_bias = None
        for blk in self.blocks:
            x = blk(x, prev_bias)
            prev_bias = blk.fc2_b
        last = self.blocks[-1] if len(self.blocks) else None
        x = F.layer_norm(x, self.lnf_w, self.lnf_b, self.cfg.layer_norm_eps,
                         col_bias=last.fc2_b if last is not None and last.fold_bias_grad else None)
        logits = F.linear(x, self.wte)  # tied LM head: [B, T, padded_vocab]
        if targets is None:
            return logits
        return F.cross_entropy(logits, targets, v_real=self.cfg.vocab_size, reduction="mean_all")


class TorchGPT2(nn.Module):
    """Plain-PyTorch twin with the same parameter names/shapes — the numerics oracle for tests and the
    stock-library arm of benchmarks (cuBLAS GEMMs, ATen LayerNorm/GELU/CE, torch.optim.AdamW)."""

    def __init__(self, cfg: GPT2Config):
        super().__init__()
        self.inner = GPT2(cfg)
        self.cfg = cfg

    def forward(self, idx, targets=None):
        m, cfg = self.inner, self.cfg
        tf = torch.nn.functional
        B, T = idx.shape
        C, H = cfg.n_embd, cfg.n_head
        x = m.wte[idx] + m.wpe[:T]
        for b in m.blocks:
            h = tf.layer_norm(x, (C,), b.ln1_w, b.ln1_b, cfg.layer_norm_eps)
            qkv = tf.linear(h, b.qkv_w, b.qkv_b).view(B, T, 3, H, C // H)
            q, k, v = (t.transpose(1, 2) for t in qkv.unbind(dim=2))
            # written out (matmul, mask, softmax, matmul) instead of F.scaled_dot_product_attention: the oracle must not
            # depend on a fused library attention kernel either
            att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(C // H))
            mask = torch.ones(T, T, dtype=torch.bool, device=x.device).tril()
            att = att.masked_fill(~mask, float("-inf")).float().softmax(dim=-1).to(q.dtype)
            y = (att @ v).transpose(1, 2).reshape(B, T, C)
            x = x + tf.linear(y, b.proj_w, b.proj_b)
            h = tf.layer_norm(x, (C,), b.ln2_w, b.ln2_b, cfg.layer_norm_eps)
            x = x + tf.linear(tf.gelu(tf.linear(h, b.fc_w, b.fc_b), approximate="tanh"), b.fc2_w, b.fc2_b)
        x = tf.layer_norm(x, (C,), m.lnf_w, m.lnf_b, cfg.layer_norm_eps)
        logits = tf.linear(x, m.wte)
        if targets is None:
            return logits
        logits = logits[..., :cfg.vocab_size]
        return tf.cross_entropy(logits.reshape(-1, cfg.vocab_size).float(), targets.reshape(-1))
