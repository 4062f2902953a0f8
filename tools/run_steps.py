"""Runs N full-size training steps (LTX-2B LoRA r=64, 49x512x768, B=1) eagerly — one Python-launched kernel at a time —
for `ncu` launch lists.  Usage: python tools/run_steps.py [n_steps] [--graph]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finetrainers_b200 import ops  # noqa: E402
from finetrainers_b200.model import B200LTXTransformer, LTXConfig  # noqa: E402
from finetrainers_b200.trainer import SFTTrainStep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
graph = "--graph" in sys.argv
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = B200LTXTransformer(LTXConfig(), torch.bfloat16, dev)
with torch.no_grad():
    for name, p in model.named_parameters():
        if "scale_shift_table" in name:
            p.normal_(0, 1.0 / p.shape[-1] ** 0.5)
        elif "norm_q" in name or "norm_k" in name:
            p.fill_(1.0)
        else:
            p.normal_(0, 0.02)
model.add_adapter(64, 64)
with torch.no_grad():
    for name, p in model.named_parameters():
        if "lora_B" in name:
            p.normal_(0, 0.01)
model.prepare()
st = SFTTrainStep(model, use_cuda_graph=graph)
lat = torch.randn(1, 128, 7, 16, 24, device=dev).bfloat16()
ehs = (torch.randn(1, 128, 4096, device=dev) * 0.1).bfloat16()
mask = torch.arange(128, device=dev)[None] < 77
mean, std = torch.zeros(1, 128, device=dev), torch.ones(1, 128, device=dev)
torch.cuda.synchronize()
print("LAUNCHES_BEFORE", ops.LAUNCH_COUNT, flush=True)
for i in range(n):
    if i == n - 1:
        torch.cuda.profiler.start()   # ncu --profile-from-start off: the launch list holds exactly ONE step
    st.train_step({"encoder_hidden_states": ehs, "encoder_attention_mask": mask},
                  {"latents": lat, "latents_mean": mean, "latents_std": std})
    torch.cuda.synchronize()
    if i == n - 1:
        torch.cuda.profiler.stop()
    print("STEP", i, "launches so far", ops.LAUNCH_COUNT, flush=True)
