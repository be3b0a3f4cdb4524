"""One eager GPT-2 small training step (B=64 x 1024) for `ncu --metrics gpu__time_duration.sum`: the launch list of the
flagship step. Warm-up happens before cudaProfilerStart so only ONE steady-state step is captured."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggy_b200.models.gpt2 import GPT2Config  # noqa: E402
from maggy_b200.parallel.trainer import build_gpt2_trainer  # noqa: E402

B = int(os.environ.get("PROF_B", 64))
cfg = GPT2Config.named("gpt2-small")
dev = torch.device("cuda:0")
model, flat, opt, step = build_gpt2_trainer(cfg, B, 1024, dev, use_cuda_graph=False)
idx = torch.randint(0, cfg.vocab_size, (B, 1024)).pin_memory()
tgt = torch.randint(0, cfg.vocab_size, (B, 1024)).pin_memory()
for _ in range(3):
    step(idx, tgt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loss = step(idx, tgt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss))
