"""BERT-base style encoder for masked-LM pre-training on the maggy_b200 kernels — the model of
BASELINE.json config 4 ("Bayesian optimizer + median earlystop, BERT-base synthetic, 4 trials x 2 GPUs").

Post-LayerNorm encoder blocks as in the original BERT; projections / MLP (fused bias+GELU and
bias+residual epilogues) run on the tcgen05 GEMM, LayerNorm and the MLM cross-entropy on the fused kernels.
The activation is the tanh-approximated GELU (the fused epilogue), attention is torch SDPA.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn

from ..ops import functional as F


@dataclass
class BertConfig:
    vocab_size: int = 30522
    hidden: int = 768
    n_layers: int = 12
    n_heads: int = 12
    ffn: int = 3072
    max_pos: int = 512
    type_vocab: int = 2
    eps: float = 1e-12

    @classmethod
    def named(cls, name: str) -> "BertConfig":
        table = {"bert-base": dict(), "bert-tiny": dict(vocab_size=1000, hidden=128, n_layers=2, n_heads=2, ffn=256, max_pos=128)}
        return cls(**table[name])


from .gpt2 import _attn_impl  # noqa: E402


class BertLayer(nn.Module):
    def __init__(self, cfg: BertConfig):
        super().__init__()
        C = cfg.hidden
        self.n_heads, self.eps = cfg.n_heads, max(cfg.eps, 1e-6)
        self.qkv_w, self.qkv_b = nn.Parameter(torch.empty(3 * C, C)), nn.Parameter(torch.zeros(3 * C))
        self.o_w, self.o_b = nn.Parameter(torch.empty(C, C)), nn.Parameter(torch.zeros(C))
        self.ln1_w, self.ln1_b = nn.Parameter(torch.ones(C)), nn.Parameter(torch.zeros(C))
        self.fc_w, self.fc_b = nn.Parameter(torch.empty(cfg.ffn, C)), nn.Parameter(torch.zeros(cfg.ffn))
        self.fc2_w, self.fc2_b = nn.Parameter(torch.empty(C, cfg.ffn)), nn.Parameter(torch.zeros(C))
        self.ln2_w, self.ln2_b = nn.Parameter(torch.ones(C)), nn.Parameter(torch.zeros(C))
        for w in (self.qkv_w, self.o_w, self.fc_w, self.fc2_w):
            nn.init.normal_(w, std=0.02)

    def forward(self, x, attn_mask=None):
        B, T, C = x.shape
        H = self.n_heads
        qkv = F.linear(x, self.qkv_w, self.qkv_b)
        if attn_mask is None and C // H == 64 and x.is_cuda and _attn_impl(None) == "native":
            y = F.attention(qkv, H, causal=False)   # K10, bidirectional
        else:
            q, k, v = qkv.view(B, T, 3, H, C // H).unbind(dim=2)
            y = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2),
                                                                 v.transpose(1, 2), attn_mask=attn_mask)
            y = y.transpose(1, 2).reshape(B, T, C)
        # post-LN: each norm is the only consumer of the linear / MLP output before it, so its backward also yields that
        # layer's bias gradient (col_bias) and no separate column-sum pass runs
        x = F.layer_norm(F.linear(y, self.o_w, self.o_b, residual=x), self.ln1_w, self.ln1_b, self.eps,
                         col_bias=self.o_b)
        x = F.layer_norm(F.mlp(x, self.fc_w, self.fc_b, self.fc2_w, self.fc2_b, residual=x), self.ln2_w, self.ln2_b,
                         self.eps, col_bias=self.fc2_b)
        return x


class BertForMaskedLM(nn.Module):
    maggy_b200_native = True

    def __init__(self, cfg: BertConfig):
        super().__init__()
        self.cfg = cfg
        C = cfg.hidden
        self.padded_vocab = (cfg.vocab_size + 127) // 128 * 128
        self.word = nn.Parameter(torch.empty(self.padded_vocab, C))
        self.pos = nn.Parameter(torch.empty(cfg.max_pos, C))
        self.type_emb = nn.Parameter(torch.empty(cfg.type_vocab, C))
        self.emb_ln_w, self.emb_ln_b = nn.Parameter(torch.ones(C)), nn.Parameter(torch.zeros(C))
        self.layers = nn.ModuleList([BertLayer(cfg) for _ in range(cfg.n_layers)])
        self.head_w, self.head_b = nn.Parameter(torch.empty(C, C)), nn.Parameter(torch.zeros(C))
        self.head_ln_w, self.head_ln_b = nn.Parameter(torch.ones(C)), nn.Parameter(torch.zeros(C))
        for w in (self.word, self.pos, self.type_emb, self.head_w):
            nn.init.normal_(w, std=0.02)

    def num_params(self):
        return sum(p.numel() for p in self.parameters())

    def forward(self, idx, labels=None, token_type=None):
        cfg = self.cfg
        B, T = idx.shape
        x = F.embedding(idx, self.word, self.pos)
        if token_type is not None:
            x = x + self.type_emb[token_type]
        else:
            x = x + self.type_emb[0]
        x = F.layer_norm(x, self.emb_ln_w, self.emb_ln_b, max(cfg.eps, 1e-6))
        for layer in self.layers:
            x = layer(x)
        # MLM head: dense + GELU + LayerNorm, decoder tied to the word embedding
        h = F.linear(x, self.head_w, self.head_b)
        h = F.layer_norm(torch.nn.functional.gelu(h, approximate="tanh"), self.head_ln_w, self.head_ln_b, max(cfg.eps, 1e-6))
        logits = F.linear(h, self.word)
        if labels is None:
            return logits
        return F.cross_entropy(logits, labels, v_real=cfg.vocab_size)
