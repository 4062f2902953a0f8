"""Finds the first full-size training step whose LoRA gradient is not finite and reports where in backward it started."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finetrainers_b200 import lib
if "--lib" in sys.argv:
    lib.LIB_PATH = sys.argv[sys.argv.index("--lib") + 1]
from finetrainers_b200.model import B200LTXTransformer, LTXConfig
from finetrainers_b200.trainer import SFTTrainStep
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = B200LTXTransformer(LTXConfig(), torch.bfloat16, dev)
with torch.no_grad():
    for name, p in model.named_parameters():
        if "scale_shift_table" in name: p.normal_(0, 1.0 / p.shape[-1] ** 0.5)
        elif "norm_q" in name or "norm_k" in name: p.fill_(1.0)
        else: p.normal_(0, 0.02)
model.add_adapter(64, 64)
if "--randb" in sys.argv:
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "lora_B" in name: p.normal_(0, 0.01)
model.prepare()
st = SFTTrainStep(model, use_cuda_graph=False, lr=1e-4, seed=7)
g = torch.Generator(device="cpu").manual_seed(1)
lat = torch.randn(4, 1, 128, 7, 16, 24, generator=g).bfloat16().to(dev)
ehs = (torch.randn(4, 1, 128, 4096, generator=g) * 0.1).bfloat16().to(dev)
mask = torch.arange(128, device=dev)[None] < 77
mean, std = torch.zeros(1, 128, device=dev), torch.ones(1, 128, device=dev)
N = 80
for i in range(N):
    st.micro_step({"encoder_hidden_states": ehs[i % 4], "encoder_attention_mask": mask},
                  {"latents": lat[i % 4], "latents_mean": mean, "latents_std": std})
    torch.cuda.synchronize()
    gflat = model.lora_grad_flat
    if not torch.isfinite(gflat).all():
        print("step", i, "loss", st.loss_buf.item(), "non-finite grad elements:", (~torch.isfinite(gflat)).sum().item(), "of", gflat.numel())
        ws = model._workspace(1, 2688, 128)
        for k in ("dy_o2", "dy_q2", "dy_kv2", "dy_o", "dy_qkv", "du_o2", "du_q2", "du_kv2", "du_o", "du_qkv", "dk2h", "dv2h"):
            t = ws[k].float()
            bad = ~torch.isfinite(t.reshape(t.shape[0], -1))
            per = bad.any(1)
            print(f"{k:8s} blocks with non-finite: {''.join('X' if b else '.' for b in per.tolist())}")
        for k in ("dy_o2", "dy_q2", "dy_o", "dy_qkv", "dy_kv2"):
            t = ws[k].float()
            bad = ~torch.isfinite(t)
            ls = bad.reshape(t.shape[0], -1).any(1).nonzero().flatten().tolist()
            if ls:
                l = max(ls)
                b2 = bad[l]
                rows = b2.any(1).nonzero().flatten(); cols = b2.any(0).nonzero().flatten()
                print(f"  {k}[{l}]: {b2.sum().item()} bad; rows {rows.min().item()}..{rows.max().item()} ({rows.numel()} rows), cols {cols.min().item()}..{cols.max().item()} ({cols.numel()} cols); nan={torch.isnan(t[l]).sum().item()} inf={torch.isinf(t[l]).sum().item()}")
                print("   first bad rows:", rows[:12].tolist(), "first bad cols:", cols[:12].tolist())
        for k in ("dh", "g", "dn", "da", "dqh", "dkh", "dvh", "dwide", "delta"):
            t = ws[k].float()
            print(f"  scratch {k}: finite={torch.isfinite(t).all().item()}")
        for k in ("h", "qh", "kh", "vh", "ao", "lse", "h1", "h2", "ffpre", "ao2", "lse2", "q2h", "k2h", "v2h"):
            t = ws[k].float()
            if not torch.isfinite(t).all():
                print("  FORWARD tensor non-finite:", k)
        break
    st.optimizer_step()
else:
    print("no non-finite gradient in", N, "steps")
