"""Catches the first attention-backward call of a full-size training run whose dQ/dK/dV is not finite, re-runs it on the same
inputs (deterministic or not?) and describes the bad elements and the inputs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finetrainers_b200 import ops
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
with torch.no_grad():
    for name, p in model.named_parameters():
        if "lora_B" in name: p.normal_(0, 0.01)
model.prepare()
st = SFTTrainStep(model, use_cuda_graph=False, lr=1e-4, seed=7)
orig = ops.attn_bwd
state = {"calls": 0, "hit": False}
def wrapped(q, k, v, kb, out, dout, lse, ws, dq, dk, dv, B, H, Sq, Sk, scale):
    r = orig(q, k, v, kb, out, dout, lse, ws, dq, dk, dv, B, H, Sq, Sk, scale)
    state["calls"] += 1
    if state["hit"]:
        return r
    fin = [torch.isfinite(t.float()).all().item() for t in (dq, dk, dv)]
    if not all(fin):
        state["hit"] = True
        print(f"call {state['calls']} Sq={Sq} Sk={Sk}: finite dq/dk/dv = {fin}")
        print("  inputs finite:", [torch.isfinite(t.float()).all().item() for t in (q, k, v, out, dout, lse)])
        bad = (~torch.isfinite(dq.float())).nonzero()
        print("  bad dq elements:", bad.shape[0], "first:", bad[:8].tolist())
        bh = bad[0, 1].item(); row = bad[0, 2].item()
        print("  values at first bad row:", dq[0, bh, row].float().tolist()[:16])
        # deterministic?  same inputs again, three times
        for t in range(3):
            dq2, dk2, dv2 = torch.zeros_like(dq), torch.zeros_like(dk), torch.zeros_like(dv)
            orig(q, k, v, kb, out, dout, lse, ws, dq2, dk2, dv2, B, H, Sq, Sk, scale)
            torch.cuda.synchronize()
            print(f"  re-run {t}: dq finite {torch.isfinite(dq2.float()).all().item()}, bad elems {(~torch.isfinite(dq2.float())).sum().item()},"
                  f" max|dq| {dq2.float().abs().nan_to_num(0, 0, 0).max().item():.3e}")
        # reference for the bad head
        qf, kf, vf = q[0, bh].float(), k[0, bh].float(), v[0, bh].float()
        S = qf @ kf.t() * scale
        print(f"  head {bh}: max|S| {S.abs().max().item():.2f}, lse[row] {lse[0, bh, row].item():.3f} vs logsumexp {torch.logsumexp(S[row], 0).item():.3f},"
              f" max|dout| {dout.float().abs().max().item():.3e} max|out| {out.float().abs().max().item():.3e}")
        do = dout.view(B, Sq, H, 64)[0, :, bh].float(); o = out.view(B, Sq, H, 64)[0, :, bh].float()
        P = torch.softmax(S, -1); dP = do @ vf.t(); delta = (do * o).sum(-1, keepdim=True)
        dS = P * (dP - delta); dq_ref = dS @ kf * scale
        print(f"  reference dq row: max {dq_ref[row].abs().max().item():.3e}; whole head max {dq_ref.abs().max().item():.3e}; ws delta[row] {ws[bh * Sq + row].item():.4e} ref {delta[row].item():.4e}")
    return r
ops.attn_bwd = wrapped
import finetrainers_b200.model as M
g = torch.Generator(device="cpu").manual_seed(1)
lat = torch.randn(4, 1, 128, 7, 16, 24, generator=g).bfloat16().to(dev)
ehs = (torch.randn(4, 1, 128, 4096, generator=g) * 0.1).bfloat16().to(dev)
mask = torch.arange(128, device=dev)[None] < 77
mean, std = torch.zeros(1, 128, device=dev), torch.ones(1, 128, device=dev)
for i in range(150):
    st.micro_step({"encoder_hidden_states": ehs[i % 4], "encoder_attention_mask": mask},
                  {"latents": lat[i % 4], "latents_mean": mean, "latents_std": std})
    if state["hit"]:
        print("step", i)
        break
    st.optimizer_step()
else:
    print("no non-finite attention gradient in 150 steps,", state["calls"], "calls")
