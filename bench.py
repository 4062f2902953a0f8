"""bench.py — the DiT training-step benchmark (BASELINE.json: LTX-Video-2B T2V LoRA SFT, 49x512x768, bf16).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

One "step" = one full SFT step of the hot path on one batch of synthetic latents per GPU: noising/packing, DiT forward
(28 blocks), flow-match MSE loss, hand-written backward, gradient all-reduce (N>1), clip + AdamW.
`value`  = latent tokens/s over the whole job with the batch already resident in HBM.
`e2e`    = the same step through the public API (SFTTrainStep.train_step) with the batch in pinned HOST memory: per
           step H2D of latents + text embeddings + mask, and a D2H read of the step's loss/grad-norm metrics.
`--impl reference` times the CPU restatement of the reference step (oracle/ltx_oracle.py; the reference itself cannot
be installed here: diffusers/peft are absent, no network) on the host cores, a bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "latent_tokens_per_sec"
UNIT = "tokens/s"
F_LAT, H_LAT, W_LAT = 7, 16, 24          # 49x512x768 -> (49-1)/8+1, 512/32, 768/32
S_TOK = F_LAT * H_LAT * W_LAT            # 2688 latent tokens per sample
TEXT_LEN = 128
RANK_LORA = 64
FLOP_PER_TOKEN_ALG = 8.88e9              # 2G + 3.5A, no recompute counted


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1444.3), d.get("hbm_gbs", 6577.4), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    def __init__(self, idx):
        self.rows, self.stop = [], False
        self.idx = idx
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.t.start()

    def finish(self):
        self.stop = True
        self.t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nme in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference step on the host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(layers=4, threads=None):
    """Times forward+loss+backward of the oracle at full width (D=2048, S=2688, r=64, B=1) on `layers` of the 28
    blocks and scales by 28/layers (embeds/head are negligible).  Returns (tokens/s, seconds_per_full_step, cores)."""
    import torch
    from oracle import ltx_oracle as O
    if threads:
        torch.set_num_threads(threads)
    cores = torch.get_num_threads()  # torch's default honours the cgroup/affinity limits of the box
    cfg = O.LTXConfig(num_layers=layers)
    m = O.LTXTransformerOracle(cfg)
    O.add_lora(m, RANK_LORA, RANK_LORA)
    O.synthetic_init_(m, seed=0, lora_b_std=0.02)
    batch = O.make_synthetic_batch(cfg, 1, F_LAT, H_LAT, W_LAT, TEXT_LEN, seed=1234, dtype=torch.float32)
    t0 = time.time()
    O.oracle_step(m, batch)
    dt = time.time() - t0
    full = dt * 28.0 / layers
    return S_TOK / full, full, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        v, full, cores = cpu_reference_sample(layers=2)
        if i >= args.warmup:
            vals.append((v, full))
    v = sum(x[0] for x in vals) / len(vals)
    full = sum(x[1] for x in vals) / len(vals)
    sample = "oracle (CPU restatement of the reference step) fwd+loss+bwd fp32, 2 of 28 blocks at full width, scaled x14"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": full * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LTX-Video-2B T2V LoRA r=64 SFT step, 49x512x768 (2688 latent tokens), B=1", "device": "cpu"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# b200 arm
# ----------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from finetrainers_b200 import ops
    from finetrainers_b200.model import B200LTXTransformer, LTXConfig
    from finetrainers_b200.trainer import SFTTrainStep
    from finetrainers_b200.parallel import B200ParallelBackend

    os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    be = B200ParallelBackend(backend="nccl") if world > 1 else None
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch

    # ---- model: LTX-2B architecture, random init (no checkpoints offline), LoRA r=64 on to_q|to_k|to_v|to_out.0
    torch.manual_seed(0)
    model = B200LTXTransformer(LTXConfig(), torch.bfloat16, dev)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "scale_shift_table" in n:
                p.normal_(0, 1.0 / p.shape[-1] ** 0.5)
            elif "norm_q" in n or "norm_k" in n:
                p.fill_(1.0)
            else:
                p.normal_(0, 0.02)
    model.add_adapter(RANK_LORA, RANK_LORA)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "lora_B" in n:
                p.normal_(0, 0.01)
    model.prepare()
    if be is not None:
        be.apply_ddp(model)
    # optimiser settings of the reference example (examples/training/sft/ltx_video/crush_smol_lora/train.sh:88-98)
    step = SFTTrainStep(model, flow_weighting_scheme="logit_normal", seed=42 + rank, use_cuda_graph=not args.no_graph,
                        lr=5e-5, beta1=0.9, beta2=0.99, weight_decay=1e-4, eps=1e-8, max_grad_norm=1.0,
                        lr_scheduler="constant_with_warmup", lr_warmup_steps=1000)

    # ---- synthetic data: a small pool of pinned host batches, plus one device-resident copy
    g = torch.Generator().manual_seed(1234 + rank)
    pool = []
    for _ in range(4):
        lat = torch.randn(B, 128, F_LAT, H_LAT, W_LAT, generator=g).bfloat16().pin_memory()
        ehs = (torch.randn(B, TEXT_LEN, 4096, generator=g) * 0.1).bfloat16().pin_memory()
        lens = torch.randint(16, TEXT_LEN + 1, (B,), generator=g)
        mask = (torch.arange(TEXT_LEN)[None] < lens[:, None]).pin_memory()
        pool.append((lat, ehs, mask))
    mean = torch.zeros(B, 128, device=dev)
    std = torch.ones(B, 128, device=dev)
    dev_pool = [(a.to(dev), b.to(dev), c.to(dev)) for a, b, c in pool]
    h2d_bytes = sum(t.numel() * t.element_size() for t in pool[0])

    def step_resident(i):
        lat, ehs, mask = dev_pool[i % len(dev_pool)]
        step.train_step({"encoder_hidden_states": ehs, "encoder_attention_mask": mask},
                        {"latents": lat, "latents_mean": mean, "latents_std": std})

    def step_e2e(i):
        lat, ehs, mask = pool[i % len(pool)]
        lat_d = lat.to(dev, non_blocking=True)
        ehs_d = ehs.to(dev, non_blocking=True)
        mask_d = mask.to(dev, non_blocking=True)
        return step.train_step({"encoder_hidden_states": ehs_d, "encoder_attention_mask": mask_d},
                               {"latents": lat_d, "latents_mean": mean, "latents_std": std}, sync_metrics=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    n_before = ops.LAUNCH_COUNT
    step_resident(0)                       # eager: also counts the kernels one step launches
    launches_per_step = ops.LAUNCH_COUNT - n_before
    for i in range(max(args.warmup, 3) + 2):   # +2: eager warm-ups before the CUDA graph is captured
        step_resident(i)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = timed(step_resident, args.steps)
    launches = launches_per_step * args.steps
    # dominant kernel measured live: the FFN up-projection GEMM launch of the step (2688 x 8192 x 2048, GELU epilogue,
    # two outputs), CUDA events on the launching stream, operands rotated over 3 buffer sets (> 126 MB L2 in total)
    R_, D_ = B * S_TOK, 2048
    sets = [(torch.randn(R_, D_, device=dev).bfloat16(), torch.empty(R_, 4 * D_, device=dev, dtype=torch.bfloat16),
             torch.empty(R_, 4 * D_, device=dev, dtype=torch.bfloat16)) for _ in range(3)]
    e_blk = model._blk[0]

    def ffn_up(i):
        x_, f_, pre_ = sets[i % 3]
        ops.gemm(x_, e_blk["W1"], f_, M=R_, N=4 * D_, K=D_, bias=e_blk["b1"], epi=ops.EPI_GELU, out2=pre_)

    for i in range(3):
        ffn_up(i)
    n_k = 30
    ms_k = timed(ffn_up, n_k)
    ktimes = {"ffn_up": (ms_k, n_k)}
    del sets
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.finish() if sampler else None

    ms_step = ms_total / args.steps
    tokens_per_step = B * S_TOK * world
    value = tokens_per_step / (ms_step * 1e-3)
    e2e_value = tokens_per_step / (ms_e2e / args.steps * 1e-3)
    if rank != 0:
        if be is not None:
            be.destroy()
        return
    peak_tf, peak_hbm, peak_src = read_peaks()
    # dominant kernel: the tcgen05 GEMM; representative launch = FFN up-projection (2688 x 8192 x 2048) measured live
    roof = None
    if "ffn_up" in ktimes and ktimes["ffn_up"][1] > 0:
        tot_ms, cnt = ktimes["ffn_up"]
        avg_ms = tot_ms / cnt
        flops = 2.0 * (B * S_TOK) * 8192 * 2048
        ach = flops / (avg_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "b2d gemm_kernel<256,0,0> (FFN up-proj 2688x8192x2048 + GELU epilogue)",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                "peak_source": peak_src + " bf16_tflops_sustained", "avg_launch_us": avg_ms * 1e3, "launches_timed": cnt,
                "step_frac_of_alg_roofline": (value / world) * FLOP_PER_TOKEN_ALG / (peak_tf * 1e12)}
        tr = os.path.join(ROOT, "profiles", "traffic_ffn_up.json")
        if os.path.exists(tr):
            roof["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, full, cores = cpu_reference_sample(layers=2)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle fwd+loss+bwd fp32, 2 of 28 blocks at full width (D=2048,S=2688,r=64), scaled x14; "
                         f"{full:.1f} s per full step"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "tokens_per_sec_per_gpu": value / world,
        "config": {"workload": f"LTX-Video-2B T2V LoRA r={RANK_LORA} SFT step, 49x512x768 (2688 latent tokens/sample), "
                               f"B={B}/GPU, AdamW+clip, logit_normal sigmas", "global_batch": B * world,
                   "parallelism": f"ddp{world}", "l2": "working set (3.8 GB weights + 5.5 GB activations per step) >> 126 MB L2; no flush needed",
                   "random_init": True},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 12,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "cuda_graph": not args.no_graph,
    }
    print(json.dumps(line))
    if be is not None:
        be.destroy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 5:
            args.steps = min(args.steps, 5)
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
