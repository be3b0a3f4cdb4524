"""Numerics of every hand-written sm_100a kernel vs a plain PyTorch fp32 reference of the same op."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.fixture(scope="module")
def F():
    from maggy_b200.ops import functional

    return functional


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (1024, 768, 768), (384, 2304, 768),
                                   (1000, 520, 200), (4096, 3072, 768), (2048, 50304, 768)])
@pytest.mark.parametrize("block_n", [0, 128, 192, 256])
def test_gemm_kmajor(F, M, N, K, block_n):
    torch.manual_seed(0)
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=DEV, dtype=torch.bfloat16)
    c = F.gemm(a, b, block_n=block_n)
    ref = a.float() @ b.float().t()
    assert _rel_err(c, ref) < 1e-2


@pytest.mark.parametrize("a_major,b_major", [(0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (768, 3072, 4096), (520, 264, 1000), (128, 128, 1024)])
def test_gemm_mn_major(F, a_major, b_major, M, N, K):
    torch.manual_seed(1)
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=DEV, dtype=torch.bfloat16)
    a_in = a.t().contiguous() if a_major else a
    b_in = b.t().contiguous() if b_major else b
    c = F.gemm(a_in, b_in, a_major=a_major, b_major=b_major)
    ref = a.float() @ b.float().t()
    assert _rel_err(c, ref) < 1e-2


@pytest.mark.parametrize("splits", [None, 1, 2, 5])
@pytest.mark.parametrize("M,N,K,am,bm", [(768, 768, 8192, 1, 1), (3072, 768, 4096, 1, 1), (256, 520, 2048, 0, 0)])
def test_gemm_split_k(F, splits, M, N, K, am, bm):
    torch.manual_seed(11)
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=DEV, dtype=torch.bfloat16)
    a_in = a.t().contiguous() if am else a
    b_in = b.t().contiguous() if bm else b
    ref = a.float() @ b.float().t()
    c = F.gemm(a_in, b_in, a_major=am, b_major=bm, splits=splits)
    assert _rel_err(c, ref) < 1e-2
    acc = torch.ones(M, N, device=DEV, dtype=torch.bfloat16)
    F.gemm(a_in, b_in, a_major=am, b_major=bm, splits=splits, out=acc, accumulate=True)
    assert _rel_err(acc, ref + 1.0) < 1e-2


@pytest.mark.parametrize("M,N,K", [(1000, 520, 200), (130, 264, 64), (4096, 2304, 768)])
def test_gemm_epilogues_tails(F, M, N, K):
    """fused epilogues through the TMA-store path with M / N tails (clipped by the tensor map)."""
    torch.manual_seed(12)
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16) * 0.5
    b = torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.1
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    res = torch.randn(M, N, device=DEV, dtype=torch.bfloat16)
    base = a.float() @ b.float().t()
    guard = torch.full((M + 8, N + 8), 7.0, device=DEV, dtype=torch.bfloat16)
    out = guard[:M, :N]
    F.gemm(a, b, bias=bias, aux=res, epi=F.EPI_BIAS_RESIDUAL, out=out)
    assert _rel_err(out, base + bias.float() + res.float()) < 1e-2
    assert (guard[M:, :] == 7.0).all() and (guard[:, N:] == 7.0).all()  # nothing written out of bounds
    pre = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    act = F.gemm(a, b, bias=bias, aux=pre, epi=F.EPI_BIAS_GELU)
    assert _rel_err(pre, base + bias.float()) < 1e-2
    assert _rel_err(act, torch.nn.functional.gelu(base + bias.float(), approximate="tanh")) < 1e-2
    x = res.float().requires_grad_(True)
    torch.nn.functional.gelu(x, approximate="tanh").sum().backward()
    assert _rel_err(F.gemm(a, b, aux=res, epi=F.EPI_DGELU), base * x.grad) < 1.5e-2


def test_gemm_epilogues(F):
    torch.manual_seed(2)
    M, N, K = 512, 768, 256
    a = torch.randn(M, K, device=DEV, dtype=torch.bfloat16) * 0.5
    b = torch.randn(N, K, device=DEV, dtype=torch.bfloat16) * 0.1
    bias = torch.randn(N, device=DEV, dtype=torch.bfloat16)
    res = torch.randn(M, N, device=DEV, dtype=torch.bfloat16)
    base = a.float() @ b.float().t()
    assert _rel_err(F.gemm(a, b, bias=bias, epi=F.EPI_BIAS), base + bias.float()) < 1e-2
    pre = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    act = F.gemm(a, b, bias=bias, aux=pre, epi=F.EPI_BIAS_GELU)
    ref_pre = base + bias.float()
    assert _rel_err(pre, ref_pre) < 1e-2
    assert _rel_err(act, torch.nn.functional.gelu(ref_pre, approximate="tanh")) < 1e-2
    out = F.gemm(a, b, bias=bias, aux=res, epi=F.EPI_BIAS_RESIDUAL)
    assert _rel_err(out, ref_pre + res.float()) < 1e-2
    # dGELU epilogue: acc * gelu'(aux)
    x = res.float().requires_grad_(True)
    torch.nn.functional.gelu(x, approximate="tanh").sum().backward()
    out = F.gemm(a, b, aux=res, epi=F.EPI_DGELU)
    assert _rel_err(out, base * x.grad) < 1.5e-2
    # fp32 output + accumulate
    c32 = torch.ones(M, N, device=DEV, dtype=torch.float32)
    F.gemm(a, b, out=c32, accumulate=True)
    assert _rel_err(c32, base + 1.0) < 1e-2
    cb = torch.ones(M, N, device=DEV, dtype=torch.bfloat16)
    F.gemm(a, b, out=cb, accumulate=True)
    assert _rel_err(cb, base + 1.0) < 1e-2


@pytest.mark.parametrize("rows,C", [(64, 768), (1000, 768), (333, 1024), (128, 2048), (77, 128)])
@pytest.mark.parametrize("rms", [False, True])
def test_norm_fwd_bwd(F, rows, C, rms):
    torch.manual_seed(3)
    x = torch.randn(rows, C, device=DEV, dtype=torch.bfloat16) * 2 + 0.5
    w = (torch.randn(C, device=DEV) * 0.5 + 1).to(torch.bfloat16)
    b = (torch.randn(C, device=DEV) * 0.1).to(torch.bfloat16)
    dy = torch.randn(rows, C, device=DEV, dtype=torch.bfloat16)
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    br = b.float().requires_grad_(True)
    if rms:
        ref = xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-5) * wr
    else:
        ref = torch.nn.functional.layer_norm(xr, (C,), wr, br, 1e-5)
    ref.backward(dy.float())
    y, mean, rstd = F.layer_norm_fwd(x, w, None if rms else b, 1e-5, rms)
    assert _rel_err(y, ref) < 1e-2
    dg = torch.empty_like(w)
    db = torch.empty_like(b)
    dx = F.layer_norm_bwd(dy, x, w, mean, rstd, dg, None if rms else db, rms=rms)
    assert _rel_err(dx, xr.grad) < 1e-2
    assert _rel_err(dg, wr.grad) < 1.5e-2
    if not rms:
        assert _rel_err(db, br.grad) < 1.5e-2


def test_colsum_gelu(F):
    torch.manual_seed(4)
    x = torch.randn(1000, 776, device=DEV, dtype=torch.bfloat16)
    out = torch.empty(776, device=DEV, dtype=torch.bfloat16)
    F.colsum(x, out)
    assert _rel_err(out, x.float().sum(0)) < 1e-2
    F.colsum(x, out, accumulate=True)
    assert _rel_err(out, 2 * x.float().sum(0)) < 2e-2
    y = F.gelu_fwd(x)
    assert _rel_err(y, torch.nn.functional.gelu(x.float(), approximate="tanh")) < 1e-2
    xr = x.float().requires_grad_(True)
    torch.nn.functional.gelu(xr, approximate="tanh").backward(torch.ones_like(xr) * 0.5)
    dx = F.gelu_bwd(x, torch.full_like(x, 0.5))
    assert _rel_err(dx, xr.grad) < 1.5e-2


def test_cross_entropy(F):
    torch.manual_seed(5)
    rows, v_real, v_pad = 300, 1000, 1024
    logits = torch.randn(rows, v_pad, device=DEV, dtype=torch.bfloat16) * 3
    tgt = torch.randint(0, v_real, (rows,), device=DEV)
    tgt[5] = -100
    lr = logits[:, :v_real].float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, tgt, ignore_index=-100, reduction="sum") / rows
    ref.backward()
    work = logits.clone().requires_grad_(True)
    loss = F.cross_entropy(work, tgt, v_real=v_real, reduction="mean_all")
    assert abs(loss.item() - ref.item()) < 2e-2 * abs(ref.item())
    loss.backward()
    g = work.grad
    assert _rel_err(g[:, :v_real], lr.grad) < 1e-2
    assert g[:, v_real:].abs().max().item() == 0.0


def test_cross_entropy_mean_over_valid_rows_matches_torch(F):
    """Default reduction = torch's ignore_index semantics (BERT-MLM-style labels, mostly -100)."""
    torch.manual_seed(6)
    rows, v_real, v_pad = 256, 1000, 1024
    logits = torch.randn(rows, v_pad, device=DEV, dtype=torch.bfloat16) * 3
    tgt = torch.randint(0, v_real, (rows,), device=DEV)
    tgt[torch.rand(rows, device=DEV) < 0.85] = -100
    lr = logits[:, :v_real].float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, tgt, ignore_index=-100)
    ref.backward()
    work = logits.clone().requires_grad_(True)
    loss = F.cross_entropy(work, tgt, v_real=v_real)
    assert abs(loss.item() - ref.item()) < 2e-2 * abs(ref.item())
    loss.backward()
    assert _rel_err(work.grad[:, :v_real], lr.grad) < 1e-2


def test_cross_entropy_large_vocab_streaming_path(F):
    """Rows that do not fit in shared memory (Llama-3 vocabulary) take the streaming kernel."""
    torch.manual_seed(55)
    rows, v_real, v_pad = 64, 128256, 128256
    logits = torch.randn(rows, v_pad, device=DEV, dtype=torch.bfloat16) * 2
    tgt = torch.randint(0, v_real, (rows,), device=DEV)
    tgt[3] = -100
    lr = logits.float().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, tgt, ignore_index=-100, reduction="sum") / rows
    ref.backward()
    work = logits.clone().requires_grad_(True)
    loss = F.cross_entropy(work, tgt, v_real=v_real, reduction="mean_all")
    assert abs(loss.item() - ref.item()) < 2e-2 * abs(ref.item())
    loss.backward()
    assert _rel_err(work.grad, lr.grad) < 1e-2


def test_embedding(F):
    torch.manual_seed(6)
    V, T, C, B = 500, 64, 128, 4
    wte = torch.randn(V, C, device=DEV, dtype=torch.bfloat16).requires_grad_(True)
    wpe = torch.randn(T * 2, C, device=DEV, dtype=torch.bfloat16).requires_grad_(True)
    idx = torch.randint(0, V, (B, T), device=DEV)
    out = F.embedding(idx, wte, wpe)
    ref_w = wte.detach().float().requires_grad_(True)
    ref_p = wpe.detach().float().requires_grad_(True)
    ref = ref_w[idx] + ref_p[:T]
    assert _rel_err(out, ref) < 1e-2
    dout = torch.randn_like(out)
    out.backward(dout)
    ref.backward(dout.float())
    assert _rel_err(wte.grad, ref_w.grad) < 2e-2
    assert _rel_err(wpe.grad, ref_p.grad) < 1e-2


def test_adam_flat(F):
    torch.manual_seed(7)
    n = 128 * 1000
    p0 = torch.randn(n, device=DEV)
    master, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pb = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    for step in range(1, 4):
        g = torch.randn(n, device=DEV).to(torch.bfloat16)
        ref.grad = g.float() * 0.5
        opt.step()
        F.adam_flat_step(master, m, v, g, pb, decay_boundary=n, lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8,
                         weight_decay=0.1, step=step, grad_scale=0.5)
    assert _rel_err(master, ref.detach()) < 1e-5
    assert _rel_err(pb, ref.detach()) < 1e-2
    # device-resident step counter path (CUDA-graph safe) gives the same bias correction
    master2, m2, v2 = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    st = torch.tensor([1], device=DEV, dtype=torch.int32)
    g = torch.randn(n, device=DEV).to(torch.bfloat16)
    F.adam_flat_step(master2, m2, v2, g, pb, decay_boundary=0, lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8,
                     weight_decay=0.1, step=99, grad_scale=1.0, step_tensor=st)
    master3, m3, v3 = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    F.adam_flat_step(master3, m3, v3, g, pb, decay_boundary=0, lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8,
                     weight_decay=0.1, step=1, grad_scale=1.0)
    assert _rel_err(master2, master3) < 1e-5


def test_autograd_linear_mlp_layernorm(F):
    torch.manual_seed(8)
    B, T, C = 4, 96, 256
    x = torch.randn(B, T, C, device=DEV, dtype=torch.bfloat16).requires_grad_(True)
    w1 = (torch.randn(4 * C, C, device=DEV) * 0.05).to(torch.bfloat16).requires_grad_(True)
    b1 = (torch.randn(4 * C, device=DEV) * 0.05).to(torch.bfloat16).requires_grad_(True)
    w2 = (torch.randn(C, 4 * C, device=DEV) * 0.05).to(torch.bfloat16).requires_grad_(True)
    b2 = (torch.randn(C, device=DEV) * 0.05).to(torch.bfloat16).requires_grad_(True)
    g = torch.ones(C, device=DEV, dtype=torch.bfloat16).requires_grad_(True)
    be = torch.zeros(C, device=DEV, dtype=torch.bfloat16).requires_grad_(True)
    h = F.layer_norm(x, g, be)
    y = F.mlp(h, w1, b1, w2, b2, residual=x)
    dy = torch.randn_like(y)
    y.backward(dy)
    got = [t.grad.clone() for t in (x, w1, b1, w2, b2, g, be)]
    tf = torch.nn.functional
    refs = [t.detach().float().requires_grad_(True) for t in (x, w1, b1, w2, b2, g, be)]
    xr, w1r, b1r, w2r, b2r, gr, ber = refs
    hr = tf.layer_norm(xr, (C,), gr, ber, 1e-5)
    yr = xr + tf.linear(tf.gelu(tf.linear(hr, w1r, b1r), approximate="tanh"), w2r, b2r)
    assert _rel_err(y, yr) < 1.5e-2
    yr.backward(dy.float())
    for a, r, name in zip(got, refs, "x w1 b1 w2 b2 g be".split()):
        assert _rel_err(a, r.grad) < 3e-2, name


def test_gpt2_step_matches_torch():
    """Full model: loss and one AdamW step vs the plain-PyTorch twin (cuBLAS/ATen, fp32 master)."""
    from maggy_b200.models.gpt2 import GPT2, GPT2Config, TorchGPT2
    from maggy_b200.parallel.arena import FlatParams, FusedAdamW

    cfg = GPT2Config.named("gpt2-tiny")
    torch.manual_seed(0)
    ref = TorchGPT2(cfg).to(DEV)
    torch.manual_seed(0)
    model = GPT2(cfg)
    flat = FlatParams(model, device=DEV)
    opt = FusedAdamW(flat, lr=1e-3, weight_decay=0.1)
    ropt = torch.optim.AdamW(
        [{"params": [p for p in ref.parameters() if p.dim() >= 2], "weight_decay": 0.1},
         {"params": [p for p in ref.parameters() if p.dim() < 2], "weight_decay": 0.0}],
        lr=1e-3, betas=(0.9, 0.95), eps=1e-8)
    idx = torch.randint(0, cfg.vocab_size, (4, cfg.block_size), device=DEV)
    tgt = torch.randint(0, cfg.vocab_size, (4, cfg.block_size), device=DEV)
    losses, rlosses = [], []
    for _ in range(5):
        opt.zero_grad()
        loss = model(idx, tgt)
        loss.backward()
        opt.step()
        losses.append(loss.item())
        ropt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            rl = ref(idx, tgt)
        rl.backward()
        ropt.step()
        rlosses.append(rl.item())
    assert abs(losses[0] - rlosses[0]) < 0.05, (losses, rlosses)
    assert abs(losses[0] - math.log(cfg.vocab_size)) < 0.5
    assert losses[-1] < losses[0]  # it learns the fixed batch
    assert abs(losses[-1] - rlosses[-1]) < 0.15, (losses, rlosses)


def test_gpt2_small_50_step_loss_parity_with_fp32_master_twin():
    """GPT-2 small (the benchmarked architecture), 50 optimizer steps: the native engine (tcgen05 GEMMs, K10 attention,
    fused norms / CE / AdamW on a bf16 arena with fp32 master shards) tracks a plain-PyTorch twin that starts from the
    SAME weights (copied, fp32 parameters under bf16 autocast, torch.optim.AdamW with the same decay grouping)."""
    from maggy_b200.models.gpt2 import GPT2, GPT2Config, TorchGPT2
    from maggy_b200.parallel.arena import FlatParams, FusedAdamW

    cfg = GPT2Config.named("gpt2-small")
    B, T, steps = 4, 512, 50
    torch.manual_seed(0)
    model = GPT2(cfg)
    flat = FlatParams(model, device=DEV)
    opt = FusedAdamW(flat, lr=3e-4, weight_decay=0.1)
    ref = TorchGPT2(cfg).to(DEV)
    with torch.no_grad():
        src = dict(model.named_parameters())
        for name, p in ref.inner.named_parameters():
            p.copy_(src[name].detach().float())
    ropt = torch.optim.AdamW(
        [{"params": [p for p in ref.parameters() if p.dim() >= 2], "weight_decay": 0.1},
         {"params": [p for p in ref.parameters() if p.dim() < 2], "weight_decay": 0.0}],
        lr=3e-4, betas=(0.9, 0.95), eps=1e-8)
    g = torch.Generator(device="cpu").manual_seed(7)
    batches = [(torch.randint(0, cfg.vocab_size, (B, T), generator=g).to(DEV),
                torch.randint(0, cfg.vocab_size, (B, T), generator=g).to(DEV)) for _ in range(4)]
    losses, rlosses = [], []
    for i in range(steps):
        idx, tgt = batches[i % 4]
        opt.zero_grad()
        loss = model(idx, tgt)
        loss.backward()
        opt.step()
        losses.append(loss.item())
        ropt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            rl = ref(idx, tgt)
        rl.backward()
        ropt.step()
        rlosses.append(rl.item())
    diffs = sorted(abs(a - b) for a, b in zip(losses, rlosses))
    assert abs(losses[0] - rlosses[0]) < 0.03, (losses[0], rlosses[0])          # same weights, same batch
    assert losses[-1] < losses[0] - 1.0 and rlosses[-1] < rlosses[0] - 1.0        # both memorise the 4 batches
    # two bf16 training runs are chaotic step by step (one of them may take a loss spike a step earlier than the
    # other): the trajectories must agree in the bulk and at the end, not at every single step
    assert diffs[len(diffs) // 2] < 0.05 and diffs[int(len(diffs) * 0.9)] < 0.3, (diffs[-5:], losses[-5:], rlosses[-5:])
    assert abs(sum(losses[-8:]) - sum(rlosses[-8:])) / 8 < 0.1, (losses[-8:], rlosses[-8:])


def test_train_step_cuda_graph_matches_eager():
    from maggy_b200.models.gpt2 import GPT2Config
    from maggy_b200.parallel.trainer import build_gpt2_trainer

    cfg = GPT2Config.named("gpt2-tiny")
    idx = torch.randint(0, cfg.vocab_size, (2, 128)).pin_memory()
    tgt = torch.randint(0, cfg.vocab_size, (2, 128)).pin_memory()
    out = {}
    for graph in (False, True):
        _, flat, opt, step = build_gpt2_trainer(cfg, 2, 128, DEV, use_cuda_graph=graph, seed=7)
        ls = [step(idx, tgt).item() for _ in range(6)]
        out[graph] = ls
        assert step.kernels_per_step > 0
    # the graph path runs 3 untimed warm-up steps before capture, so compare the common trajectory
    assert out[True][0] < out[False][0] + 1e-3
    assert out[False][-1] < out[False][0]
    assert out[True][-1] < out[True][0]


def test_llama_and_bert_train_steps():
    """The other model families run forward/backward/optimizer on the same kernels and learn a fixed batch."""
    from maggy_b200.models.bert import BertConfig, BertForMaskedLM
    from maggy_b200.models.llama import Llama, LlamaConfig
    from maggy_b200.parallel.arena import FlatParams, FusedAdamW

    for build, vocab in ((lambda: Llama(LlamaConfig.named("llama-tiny")), 1024),
                         (lambda: BertForMaskedLM(BertConfig.named("bert-tiny")), 1000)):
        torch.manual_seed(0)
        model = build()
        flat = FlatParams(model, device=DEV)
        opt = FusedAdamW(flat, lr=2e-3, weight_decay=0.01)
        idx = torch.randint(0, vocab, (4, 64), device=DEV)
        tgt = torch.randint(0, vocab, (4, 64), device=DEV)
        losses = []
        for _ in range(8):
            opt.zero_grad()
            loss = model(idx, tgt)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        assert abs(losses[0] - math.log(vocab)) < 0.7, losses
        assert losses[-1] < losses[0] - 0.05, losses


@pytest.mark.gpu
def test_layer_norm_residual_gradient_is_folded_into_backward():
    """layer_norm(..., with_residual=True): x + f(norm(x)) gets the same gradients as the unfused graph, and no
    separate add kernel is needed for the residual stream."""
    from maggy_b200.ops import functional as F

    torch.manual_seed(0)
    C = 768
    x = torch.randn(4, 256, C, device="cuda").to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(C, device="cuda")).to(torch.bfloat16).requires_grad_(True)
    b = (0.1 * torch.randn(C, device="cuda")).to(torch.bfloat16).requires_grad_(True)
    lw = (torch.randn(C, C, device="cuda") / 28).to(torch.bfloat16).requires_grad_(True)
    g = torch.randn(4, 256, C, device="cuda").to(torch.bfloat16)
    x1 = x.clone().requires_grad_(True)
    h, xr = F.layer_norm(x1, w, b, 1e-5, with_residual=True)
    y1 = F.linear(h, lw, None, residual=xr)
    y1.backward(g)
    gx1, gw1 = x1.grad.clone(), w.grad.clone()
    w.grad = b.grad = lw.grad = None
    x2 = x.clone().float().requires_grad_(True)
    ref = x2 + torch.nn.functional.linear(torch.nn.functional.layer_norm(x2, (C,), w.float(), b.float(), 1e-5), lw.float())
    ref.backward(g.float())
    assert ((y1.float() - ref).norm() / ref.norm()).item() < 1e-2
    assert ((gx1.float() - x2.grad).norm() / x2.grad.norm()).item() < 1.5e-2


@pytest.mark.gpu
@pytest.mark.parametrize("rows,C", [(5000, 768), (4099, 1024), (300, 256), (513, 2048)])
@pytest.mark.parametrize("rms", [False, True])
@pytest.mark.parametrize("res", [False, True])
def test_norm_bwd_staged_with_residual_and_column_sums(rows, C, rms, res):
    """The bulk-copy staged norm backward (csrc/ops/norm.cu norm_bwd_kernel_v4): many rows per warp so the slot ring
    wraps several times, optional residual gradient, and the column sums of dx (the producer's bias gradient) from
    the same pass; C = 2048 takes the register-staged path plus a separate column-sum pass."""
    from maggy_b200.ops import functional as F

    torch.manual_seed(11)
    x = torch.randn(rows, C, device="cuda", dtype=torch.bfloat16) * 1.5 + 0.3
    w = (torch.randn(C, device="cuda") * 0.5 + 1).to(torch.bfloat16)
    b = (torch.randn(C, device="cuda") * 0.1).to(torch.bfloat16)
    dy = torch.randn(rows, C, device="cuda", dtype=torch.bfloat16)
    rg = torch.randn(rows, C, device="cuda", dtype=torch.bfloat16) if res else None
    xr, wr, br = x.float().requires_grad_(True), w.float().requires_grad_(True), b.float().requires_grad_(True)
    if rms:
        ref = xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-5) * wr
    else:
        ref = torch.nn.functional.layer_norm(xr, (C,), wr, br, 1e-5)
    ref.backward(dy.float())
    dx_ref = xr.grad + (rg.float() if res else 0)
    _, mean, rstd = F.layer_norm_fwd(x, w, None if rms else b, 1e-5, rms)
    dg, db = torch.empty_like(w), torch.empty_like(b)
    dcol = torch.full((C,), 3.0, device="cuda", dtype=torch.bfloat16)
    dx = F.layer_norm_bwd(dy, x, w, mean, rstd, dg, None if rms else db, rms=rms, res_grad=rg, dcol=dcol,
                          accumulate_col=True)
    rel = lambda a, r: ((a.float() - r.float()).norm() / r.float().norm()).item()
    assert rel(dx, dx_ref) < 1e-2
    assert rel(dg, wr.grad) < 1.5e-2
    if not rms:
        assert rel(db, br.grad) < 1.5e-2
    assert rel(dcol, dx.float().sum(0) + 3.0) < 1e-2
    # without the column sums / accumulate flag: same dx, dgamma accumulates
    dg2 = dg.clone()
    dx2 = F.layer_norm_bwd(dy, x, w, mean, rstd, dg2, None if rms else db, rms=rms, res_grad=rg, accumulate=True)
    assert torch.equal(dx2, dx)
    assert rel(dg2, 2 * wr.grad) < 2e-2


@pytest.mark.gpu
def test_bias_gradient_comes_out_of_the_following_norm():
    """layer_norm(..., col_bias=b): the producing linear / mlp skips its column-sum pass and the bias gradient equals
    the unfused one, with and without the main_grad arena convention."""
    from maggy_b200.ops import functional as F

    torch.manual_seed(5)
    C, M = 768, 2048
    mk = lambda *s, sc=1.0: (torch.randn(*s, device="cuda") * sc).to(torch.bfloat16)
    x0, g = mk(M, C), mk(M, C)
    params = [mk(C, C, sc=0.04), mk(C, sc=0.1), (1 + mk(C, sc=0.1)), mk(C, sc=0.1), mk(4 * C, C, sc=0.04), mk(4 * C, sc=0.1),
              mk(C, 4 * C, sc=0.02), mk(C, sc=0.1)]

    def run(fused: bool, main_grad: bool):
        ps = [p.clone().requires_grad_(True) for p in params]
        lw, lb, nw, nb, w1, b1, w2, b2 = ps
        if main_grad:
            for p in ps:
                p.main_grad = torch.zeros_like(p)
                p._mg_written = False
        x = x0.clone().requires_grad_(True)
        y = F.linear(x, lw, lb, residual=x)                                 # producer 1: linear with bias
        h, y_res = F.layer_norm(y, nw, nb, 1e-5, with_residual=True, col_bias=lb if fused else None)
        z = F.mlp(h, w1, b1, w2, b2, residual=y_res)                        # producer 2: mlp (fc2 bias)
        out = F.layer_norm(z, nw, nb, 1e-5, col_bias=b2 if fused else None)
        before = F.launch_count()
        out.backward(g)
        launches = F.launch_count() - before
        grads = [(p.main_grad if main_grad else p.grad).float().clone() for p in ps]
        return grads, x.grad.float().clone(), launches

    for main_grad in (False, True):
        (ref, gx_ref, n_ref), (got, gx, n_got) = run(False, main_grad), run(True, main_grad)
        assert n_got == n_ref - 4, (n_got, n_ref)   # two colsum passes (partial + finalize each) are gone
        assert torch.equal(gx, gx_ref)
        for name, a, r in zip(["lw", "lb", "nw", "nb", "w1", "b1", "w2", "b2"], got, ref):
            assert ((a - r).norm() / (r.norm() + 1e-6)).item() < 1e-2, (name, main_grad)


@pytest.mark.gpu
def test_pair_gemm_is_deterministic():
    """The CTA-pair (cta_group::2) GEMM path that serves the big training shapes: 20 launches on the same operands are
    bit-identical, for the K-major forward and the MN-major dgrad / wgrad forms (compute-sanitizer racecheck cannot follow
    the cluster-scope mbarrier that orders the peer CTA's multicast against the MMA reads: profiles/sanitizer_r2.md)."""
    from maggy_b200.ops import functional as F

    torch.manual_seed(9)
    M, N, K = 8192, 3072, 768
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    dy = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    forms = [lambda: F.gemm(a, b), lambda: F.gemm(dy, b, b_major=F.MN_MAJOR),
             lambda: F.gemm(dy, a, a_major=F.MN_MAJOR, b_major=F.MN_MAJOR, splits=1)]
    for f in forms:
        first = f().clone()
        for _ in range(20):
            assert torch.equal(f(), first)


@pytest.mark.gpu
@pytest.mark.parametrize("M,C", [(1000, 256), (4096, 768), (130, 64), (8192, 192)])
def test_dgelu_epilogue_leaves_the_bias_gradient_behind(M, C):
    """gemm(..., epi=EPI_DGELU, bias=<fp32 partial buffer>): the column sums of d_pre come out of the dgrad epilogue
    (rows past M are masked, N tails are clipped) and mg_colsum_finalize folds them into the bias gradient — compared
    with the separate column-sum pass over the same d_pre."""
    from maggy_b200.ops import _lib
    from maggy_b200.ops import functional as F

    torch.manual_seed(13)
    N = 4 * C
    dy = torch.randn(M, C, device="cuda", dtype=torch.bfloat16)
    w2 = (torch.randn(C, N, device="cuda") * 0.05).to(torch.bfloat16)
    pre = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    rows = 4 * ((M + 127) // 128)
    part = torch.full((rows * N,), float("nan"), device="cuda", dtype=torch.float32)   # every entry must be written
    dpre = F.gemm(dy, w2, b_major=F.MN_MAJOR, aux=pre, epi=F.EPI_DGELU, bias=part)
    plain = F.gemm(dy, w2, b_major=F.MN_MAJOR, aux=pre, epi=F.EPI_DGELU)
    assert torch.equal(dpre, plain)
    assert torch.isfinite(part).all()
    got = torch.full((N,), 2.0, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.cuda().mg_colsum_finalize(part.data_ptr(), rows, N, got.data_ptr(), 1, F._stream()), "finalize")
    ref = dpre.float().sum(0) + 2.0
    assert ((got.float() - ref).norm() / ref.norm()).item() < 5e-3
    # through autograd (the fused path is opt-in: it measured slower end to end): same fc bias gradient as the reference math
    import maggy_b200.ops.functional as Fmod
    saved, Fmod._DGELU_COLSUM = Fmod._DGELU_COLSUM, True
    x = torch.randn(M, C, device="cuda", dtype=torch.bfloat16)
    w1 = (torch.randn(N, C, device="cuda") * 0.05).to(torch.bfloat16).requires_grad_(True)
    b1 = (torch.randn(N, device="cuda") * 0.05).to(torch.bfloat16).requires_grad_(True)
    w2t = (torch.randn(C, N, device="cuda") * 0.05).to(torch.bfloat16).requires_grad_(True)
    b2 = torch.zeros(C, device="cuda", dtype=torch.bfloat16).requires_grad_(True)
    F.mlp(x, w1, b1, w2t, b2).backward(dy)
    tf = torch.nn.functional
    r = [t.detach().float().requires_grad_(True) for t in (w1, b1, w2t, b2)]
    tf.linear(tf.gelu(tf.linear(x.float(), r[0], r[1]), approximate="tanh"), r[2], r[3]).backward(dy.float())
    Fmod._DGELU_COLSUM = saved
    assert ((b1.grad.float() - r[1].grad).norm() / r[1].grad.norm()).item() < 3e-2
