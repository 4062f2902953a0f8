"""Soak of the full-size graph-replayed step: N optimizer steps twice from the same seed; every 50th step's loss / grad norm
must be finite and the two runs must agree (the only non-bitwise-deterministic kernels are the fp32-atomic cross-attention
dK/dV accumulations, so agreement is held to 1e-3 relative).  A race between overlapped kernels (dependent launch) shows up
here as NaNs or as runs that drift apart.   python tools/soak.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finetrainers_b200.model import B200LTXTransformer, LTXConfig  # noqa: E402
from finetrainers_b200.trainer import SFTTrainStep  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
dev = torch.device("cuda", 0)


def run():
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
    model.prepare()
    st = SFTTrainStep(model, use_cuda_graph=True, lr=1e-4, seed=7)
    g = torch.Generator(device="cpu").manual_seed(1)
    lat = torch.randn(4, 1, 128, 7, 16, 24, generator=g).bfloat16().to(dev)
    ehs = (torch.randn(4, 1, 128, 4096, generator=g) * 0.1).bfloat16().to(dev)
    mask = torch.arange(128, device=dev)[None] < 77
    mean, std = torch.zeros(1, 128, device=dev), torch.ones(1, 128, device=dev)
    out = []
    for i in range(N):
        m = st.train_step({"encoder_hidden_states": ehs[i % 4], "encoder_attention_mask": mask},
                          {"latents": lat[i % 4], "latents_mean": mean, "latents_std": std}, sync_metrics=(i % 50 == 49))
        if m is not None:
            out.append((m["train/global_avg_loss"], m["train/grad_norm"]))
    torch.cuda.synchronize()
    return out, model.lora_flat.detach().clone()


a, pa = run()
b, pb = run()
print("run A:", " ".join(f"{l:.5f}/{g:.4f}" for l, g in a))
print("run B:", " ".join(f"{l:.5f}/{g:.4f}" for l, g in b))
ok = all(map(lambda t: t[0] == t[0] and abs(t[0]) < 1e4 and t[1] == t[1], a + b))
rel = max(abs(x[0] - y[0]) / abs(x[0]) for x, y in zip(a, b))
drift = ((pa - pb).norm() / pa.norm()).item()
print(f"finite={ok} max loss rel diff between runs={rel:.2e} adapter rel diff={drift:.2e}")
print("SOAK_OK" if (ok and rel < 1e-3 and drift < 1e-2) else "SOAK_DIFF")
