#!/usr/bin/env python
"""Benchmark of the SpeechMix speech-to-text training step on B200 -> train audio-seconds / second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4|cfg5|gan]
                    [--dropout P]

Default workload (the headline, BASELINE.json configs[1]): SpeechMixEED wav2vec2-base + bart-base, down_scale=2,
batch 32 x 15 s per GPU, bf16 forward + loss + backward + optimizer step.  --config selects the other BASELINE
configurations (cfg3 Adapter hubert-large + bart-large with frozen backbones, cfg4 Self wav2vec2-large + t5-base
share_layer_ratio 0.5, cfg5 EED hubert-large + mbart-large-50 at 8 x 30 s per GPU).
N > 1 is launched by torchrun (one rank per GPU, NCCL); data parallel, weak scaling.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for what every key means.
"""
import argparse
import contextlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")

RATE = 16000
# name -> model class suffix, speech preset (kind, model_type), text preset, ctor kwargs, per-GPU batch, seconds, T_dec,
#         CPU-baseline sample batch, description.  Shapes: SURVEY.md section 8(d).
WORKLOADS = {
    "cfg2": dict(cls="EED", speech=("base", "wav2vec2"), text="bart-base", kwargs=dict(down_scale=2), batch=32,
                 seconds=15.0, t_dec=64, cpu_batch=2,
                 desc="SpeechMixEED wav2vec2-base + bart-base down_scale=2 (BASELINE.json configs[1])"),
    "cfg3": dict(cls="Adapter", speech=("large", "hubert"), text="bart-large",
                 kwargs=dict(down_scale=8, fixed_parameters=True, fixed_except=[]), batch=32, seconds=15.0, t_dec=64,
                 cpu_batch=1,
                 desc="SpeechMixAdapter hubert-large + bart-large down_scale=8, frozen backbones, adapter tuning "
                      "(BASELINE.json configs[2])"),
    "cfg4": dict(cls="Self", speech=("large", "wav2vec2"), text="t5-base",
                 kwargs=dict(down_scale=8, share_layer_ratio=0.5), batch=32, seconds=15.0, t_dec=64, cpu_batch=1,
                 desc="SpeechMixSelf wav2vec2-large + t5-base share_layer_ratio=0.5 down_scale=8, CE + KL + MSE "
                      "(BASELINE.json configs[3])"),
    "cfg5": dict(cls="EED", speech=("large", "hubert"), text="mbart-large-50", kwargs=dict(down_scale=8), batch=8,
                 seconds=30.0, t_dec=128, cpu_batch=1,
                 desc="SpeechMixEED hubert-large + mbart-large-50 (V=250054) down_scale=8, 8 x 30 s per GPU = 64 x 30 s "
                      "at 8 GPUs (BASELINE.json configs[4])"),
    # not a BASELINE configuration: the discriminator variant at the headline shapes (run with --no-graph: its update-phase
    # counters are host state)
    "gan": dict(cls="GAN", speech=("base", "wav2vec2"), text="bart-base", kwargs=dict(down_scale=2), batch=32,
                seconds=15.0, t_dec=64, cpu_batch=2,
                desc="SpeechMixGAN wav2vec2-base + bart-base down_scale=2, four BCE discriminator terms (cfg2 shapes)"),
}
SHAPES = {   # speech: (H, FF, L); text: (D, DFF, Le, Ld, V)
    ("base", "wav2vec2"): (768, 3072, 12), ("large", "hubert"): (1024, 4096, 24), ("large", "wav2vec2"): (1024, 4096, 24),
    "bart-base": (768, 3072, 6, 6, 50265), "bart-large": (1024, 4096, 12, 12, 50265),
    "t5-base": (768, 3072, 12, 12, 32128), "mbart-large-50": (1024, 4096, 12, 12, 250054),
}


def frames(secs):
    T = int(secs * RATE)
    for k, s in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
        T = (T - k) // s + 1
    return T


def fwd_flops_parts(w):
    """BASELINE.md section 3 formulas (2*MAC, dense attention), per sample, split by stage."""
    H, FF, L = SHAPES[w["speech"]]
    D, DFF, Le, Ld, V = SHAPES[w["text"]]
    L = L - int(L * w["kwargs"].get("share_layer_ratio", 0))
    ds, t_dec = w["kwargs"]["down_scale"], w["t_dec"]
    T = int(w["seconds"] * RATE)
    conv, cin = 0.0, 1
    for k, s in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
        T = (T - k) // s + 1
        conv += 2.0 * T * 512 * cin * k
        cin = 512
    speech = conv + 2.0 * T * 512 * H + 2.0 * T * H * (H // 16) * 128
    speech += L * (8.0 * T * H * H + 4.0 * T * T * H + 4.0 * T * H * FF)
    bridge, Tds = 0.0, T
    for _ in range(int(round(math.log2(ds)))):
        Tds = (Tds - 2) // 2 + 1
        bridge += 4.0 * Tds * H * H
    bridge += 2.0 * Tds * H * D
    tenc = Le * (8.0 * Tds * D * D + 4.0 * Tds * Tds * D + 4.0 * Tds * D * DFF)
    tdec = Ld * (8.0 * t_dec * D * D + 4.0 * t_dec * t_dec * D + 4.0 * t_dec * D * D + 4.0 * Tds * D * D +
                 4.0 * t_dec * Tds * D + 4.0 * t_dec * D * DFF)
    head = 2.0 * t_dec * D * V
    adapters = (Le * Tds + Ld * t_dec) * 2.0 * D * D if w["cls"] == "Adapter" else 0.0   # two D x D/2 GEMMs per layer
    return dict(speech=speech, bridge=bridge, text_enc=tenc, text_dec=tdec, head=head, adapters=adapters,
                frames=T, frames_ds=Tds)


def fwd_flops_per_sample(w=None):
    p = fwd_flops_parts(w or WORKLOADS["cfg2"])
    return p["speech"] + p["bridge"] + p["text_enc"] + p["text_dec"] + p["head"]


def step_flops_per_sample(w):
    """Algorithmic FLOPs of one training step per sample: 3 x forward where everything
    trains; forward + data gradient (2 x) for frozen layers above a trainable tensor; forward only below the lowest
    trainable tensor."""
    p = fwd_flops_parts(w)
    if w["cls"] == "EED":
        return 3.0 * (p["speech"] + p["bridge"] + p["text_enc"] + p["text_dec"] + p["head"])
    if w["cls"] == "Adapter":   # frozen speech (forward only), frozen text layers (fwd + dgrad), trainable bridge / adapters
        return p["speech"] + 3.0 * (p["bridge"] + p["adapters"]) + 2.0 * (p["text_enc"] + p["text_dec"] + p["head"])
    if w["cls"] == "Self":      # trainable speech + bridge, frozen text model run twice (student: fwd + dgrad; teacher: fwd)
        D, DFF, Le, Ld, V = SHAPES[w["text"]]
        t_txt = w["t_dec"]
        teacher = Le * (8.0 * t_txt * D * D + 4.0 * t_txt * t_txt * D + 4.0 * t_txt * D * DFF) + p["text_dec"] + p["head"]
        return 3.0 * (p["speech"] + p["bridge"]) + 2.0 * (p["text_enc"] + p["text_dec"] + p["head"]) + teacher
    if w["cls"] == "GAN":       # everything trains; text model run twice (speech embeddings, label ids); the LM head only
        D, DFF, Le, Ld, V = SHAPES[w["text"]]                                  # feeds the argmax ids (forward only)
        t_txt, Tds = w["t_dec"], p["frames_ds"]
        text_path = (Le * (8.0 * t_txt * D * D + 4.0 * t_txt * t_txt * D + 4.0 * t_txt * D * DFF) +
                     Ld * (8.0 * t_txt * D * D + 4.0 * t_txt * t_txt * D + 8.0 * t_txt * D * D + 4.0 * t_txt * t_txt * D +
                           4.0 * t_txt * D * DFF))
        disc = 2.0 * D * D * (Tds + 3 * t_txt)                                  # Z = X W^T for the four state tensors
        return 3.0 * (p["speech"] + p["bridge"] + p["text_enc"] + p["text_dec"] + text_path + disc) + p["head"]
    raise ValueError(w["cls"])


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _ncu_traffic():
    """dram__bytes_read + write per launch of the dominant kernel, from the committed ncu --set full capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        return None


def set_dropout(spc, txc, p):
    """switch the train-mode dropout of the HF backbones on at probability p (every site the stock configs name)"""
    if p > 0.0:
        for cfg, keys in ((spc, ("hidden_dropout", "attention_dropout", "activation_dropout", "feat_proj_dropout")),
                          (txc, ("dropout", "attention_dropout", "activation_dropout", "dropout_rate"))):
            for k in keys:
                if hasattr(cfg, k):
                    setattr(cfg, k, float(p))


def build_cpu_reference(w, batch, dropout=0.0):
    """The reference's CPU path, restated (oracle/hf_oracle.py; /root/reference does not exist on the
    GPU box): HFSpeechMix{EED,Adapter,Self} glue over the transformers backbones, fp32."""
    import torch
    from oracle import hf_oracle as O
    spc, txc = O.speech_config(w["speech"][0], model_type=w["speech"][1]), O.text_config(w["text"])
    if w["text"] == "t5-base":
        txc.decoder_start_token_id = 0
    set_dropout(spc, txc, dropout)
    s, t = O.build_backbones(spc, txc, seed=0)
    with contextlib.redirect_stdout(sys.stderr):
        model = getattr(O, "Oracle" + w["cls"])(s, t, **w["kwargs"]).train()
    x, labels = O.synthetic_batch(batch, w["seconds"], w["t_dec"], txc.vocab_size, seed=0)
    extra = {}
    if w["cls"] == "Self":
        extra["text_input_ids"] = torch.randint(4, txc.vocab_size, (batch, w["t_dec"]))
    return model, x, labels, extra


def time_cpu_reference(w, steps, warmup, batch=None, optimizer="adafactor", dropout=0.0):
    import torch
    batch = batch or w["cpu_batch"]
    torch.set_num_threads(os.cpu_count())
    model, x, labels, extra = build_cpu_reference(w, batch, dropout)
    params = [p for p in model.parameters() if p.requires_grad]
    if optimizer == "adafactor":   # the recipe's optimizer exactly as the HF Trainer builds it (ref:train.py:298)
        from transformers.optimization import Adafactor
        opt = Adafactor(params, lr=1e-5, scale_parameter=False, relative_step=False)
    else:
        opt = torch.optim.AdamW(params, lr=1e-5)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = model(x, labels=labels, **extra)
        out["loss"].backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return batch * w["seconds"] / sec, sec, torch.get_num_threads()


def run_reference_arm(args, rank):
    if rank != 0:
        return
    w = WORKLOADS[args.config]
    val, sec, threads = time_cpu_reference(w, args.steps, args.warmup, optimizer=args.optimizer, dropout=args.dropout)
    oname = "Adafactor" if args.optimizer == "adafactor" else "AdamW"
    line = {"impl": "reference", "metric": "train audio-sec/s", "value": val, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"] + ", fwd+bwd+" + oname, "name": args.config, "dropout": args.dropout,
                       "sample": "batch %d x %g s per step on host CPU" % (w["cpu_batch"], w["seconds"])},
            "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": threads, "kind": "port",
                             "sample": "oracle/hf_oracle.py Oracle%s (restated reference glue over transformers), "
                                       "batch %d x %g s, %d timed steps" % (w["cls"], w["cpu_batch"], w["seconds"], args.steps)},
            "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--optimizer", default="adafactor", choices=["adamw", "adafactor"],
                    help="adafactor = the recipe's optimizer (ref:train.py:298 optim=adafactor) -- ours: speechmix_b200.optim."
                         "FusedAdafactor, reference arm: transformers' Adafactor; adamw = torch AdamW (fused on the GPU)")
    ap.add_argument("--grad-payload", default="bf16", choices=["fp32", "bf16"],
                    help="wire format of the data-parallel gradient all-reduce (N > 1)")
    ap.add_argument("--dropout", type=float, default=0.0,
                    help="dropout probability at every site of both backbones (stock checkpoints: 0.1).  Default 0 = the "
                         "measurement plan of BASELINE.md section 4 (dropout / LayerDrop / SpecAugment zeroed on both arms)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="eager launches instead of a whole-step CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    import speechmix_b200
    from speechmix_b200 import kernels, parallel, presets

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = WORKLOADS[args.config]
    SECONDS, T_DEC = w["seconds"], w["t_dec"]
    spc = presets.speech_config(w["speech"][0], model_type=w["speech"][1])
    txc = presets.text_config(w["text"])
    set_dropout(spc, txc, args.dropout)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(sys.stderr):   # the reference-compatible ctor prints its layer-sharing summary
        model = getattr(speechmix_b200, "SpeechMix" + w["cls"])(spc, txc, **w["kwargs"])
    parallel.init_like_reference(model, seed=0)
    model = model.to(dev).train()
    B = args.batch or w["batch"]
    n_samples = int(SECONDS * RATE)
    g = torch.Generator().manual_seed(1234 + rank)
    host_x = [torch.randn(B, n_samples, generator=g).pin_memory() for _ in range(2)]
    host_y = [torch.randint(4, txc.vocab_size, (B, T_DEC), generator=g).pin_memory() for _ in range(2)]
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]
    fkw = {}
    if w["cls"] == "Self":      # text-teacher input (ref:speechmix/hf_model.py:541-546); resident on the device
        fkw["text_input_ids"] = torch.randint(4, txc.vocab_size, (B, T_DEC), generator=g).to(dev)

    params = [p for p in model.parameters() if p.requires_grad]
    n_train = sum(p.numel() for p in params)
    if args.optimizer == "adafactor":
        from speechmix_b200.optim import FusedAdafactor
        opt = FusedAdafactor(params, lr=1e-5, capturable=bool(args.graph and world == 1))
    else:
        opt = torch.optim.AdamW(params, lr=1e-5, fused=True, capturable=bool(args.graph and world == 1))
    dp = parallel.GradientAllReducer(model, world, payload=args.grad_payload) if world > 1 else None

    def step_eager(x, y):
        opt.zero_grad(set_to_none=True)
        out = model(x, labels=y, return_model_detail=False, **fkw)
        out["loss"].backward()
        if dp is not None:
            dp.finish()
        opt.step()
        return out["loss"]

    # CUDA graph of the step (N = 1: forward + backward + optimizer in one graph; N > 1: forward + backward in the
    # graph, bucketed NCCL all-reduce and the optimizer step eagerly after the replay) -- disable with --no-graph
    graphed = None
    if args.graph:
        from speechmix_b200.graph import GraphedTrainStep
        try:
            graphed = GraphedTrainStep(model, opt, dev_x[0], dev_y[0], warmup=max(args.warmup, 3), reducer=dp,
                                       forward_kwargs=fkw)
        except Exception as e:   # a failed capture must not cost the measurement: fall back to eager launches
            sys.stderr.write("CUDA graph capture failed (%r); running eager\n" % (e,))
            graphed = None
            kernels._ARENA.reset()
            if dp is not None:
                dp.graph_mode, dp.enabled = False, True
                dp.pending = [len(b) for b in dp.buckets]
                dp.works = [None] * len(dp.buckets)
                dp.ready, dp.next_bucket = [False] * len(dp.buckets), 0
            opt.zero_grad(set_to_none=True)

    def step(x, y):
        return graphed(x, y) if graphed is not None else step_eager(x, y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    copy_stream = torch.cuda.Stream()
    stage_x = [torch.empty_like(dev_x[0]) for _ in range(2)]
    stage_y = [torch.empty_like(dev_y[0]) for _ in range(2)]
    staged = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            stage_x[i & 1].copy_(host_x[i & 1], non_blocking=True)
            stage_y[i & 1].copy_(host_y[i & 1], non_blocking=True)
            staged[i & 1].record(copy_stream)

    host_loss = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    step_done = [torch.cuda.Event() for _ in range(2)]

    def timed(n, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        if e2e:
            prefetch(0)
        for i in range(n):
            if e2e:
                # every step: H2D copy of ITS batch from pinned host memory (issued on a copy stream while the
                # previous step computes), the step, and a D2H read of ITS loss.  The loss travels through a pinned
                # slot and is read on the host one step later (after step i + 1 has been queued), the way a training
                # loop logs: the GPU does not idle while the host reads a number.
                torch.cuda.current_stream().wait_event(staged[i & 1])
                loss = step(stage_x[i & 1], stage_y[i & 1])
                host_loss[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
                step_done[i & 1].record()
                if i + 1 < n:
                    if i >= 1:   # the staging buffer of batch i + 1 was last read by step i - 1
                        copy_stream.wait_event(step_done[(i - 1) & 1])
                    prefetch(i + 1)
                if i >= 1:
                    step_done[(i - 1) & 1].synchronize()
                    last = float(host_loss[(i - 1) & 1])
            else:
                last = step(dev_x[i & 1], dev_y[i & 1])
        if e2e:
            step_done[(n - 1) & 1].synchronize()
            last = float(host_loss[(n - 1) & 1])
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, last

    for _ in range(args.warmup):
        step(dev_x[0], dev_y[0])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = kernels.LAUNCHES[0]
    kernels.TIMING = [] if (rank == 0 and graphed is None) else None
    prof_range = os.environ.get("SMX_PROFILE_RANGE") == "1"   # `ncu --profile-from-start off` captures only the timed steps
    if prof_range:
        torch.cuda.profiler.start()
    ms, _ = timed(args.steps, e2e=False)
    if prof_range:
        torch.cuda.profiler.stop()
    gemm_events = kernels.TIMING
    kernels.TIMING = None
    launches = (kernels.LAUNCHES[0] - launches0) // args.steps
    ms_e2e, _ = timed(args.steps, e2e=True)
    if graphed is not None:
        # kernels inside a replayed graph cannot be bracketed by events: the dominant kernel is timed in two
        # extra EAGER steps (same process, same stream, same in-step cache / clock state) right after the timed
        # region -- on every rank (the steps contain the gradient all-reduce), recorded on rank 0
        kernels.TIMING = [] if rank == 0 else None
        for i in range(2):
            step_eager(dev_x[i & 1], dev_y[i & 1])
        torch.cuda.synchronize()
        gemm_events = kernels.TIMING
        kernels.TIMING = None
    sampler.stop_flag = True

    if rank == 0:
        audio_s = B * world * SECONDS
        fl_step = step_flops_per_sample(w) * B
        H_sp, FF_sp, _ = SHAPES[w["speech"]]
        M_dom = B * frames(SECONDS)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        # dominant kernel: the speech FFN up-projection GEMM (M = B*frames, N = FF, K = H), timed live with CUDA events
        dom = [(a.elapsed_time(b), fl) for (tag, fl, a, b) in (gemm_events or [])
               if tag == "nt_%d_%d_%d" % (M_dom, FF_sp, H_sp)]
        roof = None
        if dom:
            avg_ms = sum(d for d, _ in dom) / len(dom)
            ach = dom[0][1] / avg_ms / 1e9
            roof = {"bound": "tensor", "kernel": "gemm_kernel<NT> M=%d N=%d K=%d (FFN up-projection + bias + GELU)" % (M_dom, FF_sp, H_sp),
                    "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                    "launches_timed": len(dom), "traffic": _ncu_traffic() if args.config == "cfg2" and B == 32 else None,
                    "step_frac_of_peak": fl_step / (ms * 1e-3) / 1e12 / peak_tf}
        line = {"metric": "train audio-sec/s", "value": audio_s / (ms * 1e-3), "unit": "audio-s/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "%s, batch %d x %g s per GPU, T_dec=%d, fwd+loss+bwd+%s"
                                       % (w["desc"], B, SECONDS, T_DEC, "AdamW" if args.optimizer == "adamw" else "Adafactor"),
                           "name": args.config, "global_batch": B * world, "parallelism": "dp%d" % world,
                           "dropout": args.dropout, "dropout_sites": len(model.dropout_sites),
                           "trainable_parameters": n_train,
                           "allreduce_bytes_per_step": dp.payload_bytes() if dp is not None else 0,
                           "grad_payload": args.grad_payload if dp is not None else None,
                           "l2": "per-step activations (>3 GB) exceed the 126 MB L2",
                           "launch": ("eager" if graphed is None else "whole-step CUDA graph" if world == 1 else
                                      "CUDA graph (fwd+bwd) + eager NCCL all-reduce + optimizer")},
                "e2e": {"value": audio_s / (ms_e2e * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": B * n_samples * 4 + B * T_DEC * 8, "d2h_bytes_per_step": 4},
                "gpu_launches": int(launches) * args.steps,
                "gpu_launches_per_step": int(launches),
                "step_tflops": fl_step / (ms * 1e-3) / 1e12,
                "clocks": sampler.summary(),
                "roofline": roof}
        if not args.no_cpu_baseline and world == 1:
            val, sec, threads = time_cpu_reference(w, steps=2, warmup=1, optimizer=args.optimizer, dropout=args.dropout)
            line["cpu_baseline"] = {"value": val, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                    "sample": "oracle Oracle%s fp32 fwd+bwd+%s, batch %d x %g s, 2 timed steps (%.1f s/step)"
                                              % (w["cls"], "Adafactor" if args.optimizer == "adafactor" else "AdamW",
                                                 w["cpu_batch"], SECONDS, sec)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
