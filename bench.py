#!/usr/bin/env python
"""Headline benchmark: utterances/s of the TitaNet hot path, forward + backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic utterances:
raw waveform -> mel (CUDA) -> TitaNet-S/17 encoder + decoder -> CE loss -> backward
(gradients for every parameter; N>1: + one NCCL all-reduce of the gradients).  The
workload is BASELINE.json configs[1] (TitaNet-S, CE loss, batch 64 per GPU, 3 s @ 16 kHz
synthetic waveforms, dropout 0.1 as parameters.yml:57).

One JSON line on stdout (rank 0).  ``value`` is device-resident throughput (inputs already
in HBM), ``e2e`` the same metric through the public module API with pinned host buffers
(H2D of the step's waveforms + labels and a D2H read of the loss inside the timed region).
``--impl reference`` times the reference's own CPU implementation of the same path (the unmodified
modules from baseline/_ref, all host threads, plus the reference's 2-thread default) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

SAMPLE_RATE, N_CLASSES = 16000, 251


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU per step")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--model", default="s")
    ap.add_argument("--blocks", type=int, default=17)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--cpu-batch", type=int, default=0, help="utterances per step of the CPU arm / CPU baseline (0: the stated batch, capped at 64 / 16 / 8 for S / M / L)")
    ap.add_argument("--loss", default="ce", choices=["ce", "arc"], help="ce: CELoss (configs[1]); arc: ArcFaceLoss s=30 m=0.2 (configs[2], [3])")
    ap.add_argument("--ragged", action="store_true",
                    help="configs[3]: utterance lengths 1..--seconds s (whole seconds), each mel on its own length, zero padded")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly instead of replaying a CUDA graph")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}
    if os.path.exists(path):
        d = json.load(open(path))
        out = {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    # dense TF32 peak measured on this pool's B200 the way MEASURED_PEAKS.json measures bf16 (tools/measure_tf32_peak.py, burst:
    # the roofline kernel is timed alone); bf16 / 2 only when that file is missing
    out["tf32_tflops"], out["tf32_source"] = out["bf16_tflops"] / 2.0, "bf16 sustained / 2 (not measured)"
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_tf32_peak.json")))
        out["tf32_tflops"], out["tf32_source"] = float(t["tf32_tflops"]), "measured (tools/measure_tf32_peak.py, torch.matmul TF32 8192^3, burst)"
    except (OSError, ValueError, KeyError):
        pass
    return out


# ----------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference modules from baseline/_ref (installed by baseline/install_ref.py; git-ignored, shipped to the
# GPU box by gpurun), driven exactly as the reference drives them: per-utterance transforms.MelSpectrogram (datasets.py:292-293),
# zero-pad collate (datasets.py:48-73), model(spectrograms, speakers=...) and loss.backward() (learn.py:95-117).  Falls back to
# the oracle port (oracle/titanet_oracle.py) only when baseline/_ref is absent, and says which in `kind`.
# ----------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
CPU_BATCH_CAP = {"s": 64, "m": 16, "l": 8}          # utterances per CPU step: S runs the stated batch of 64


def reference_available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("modules.py", "models.py", "losses.py", "transforms.py"))


def cpu_batch(args) -> int:
    cap = CPU_BATCH_CAP.get(args.model.lower(), 8)
    if args.seconds > 3.0:
        cap = max(2, int(cap * 3.0 / args.seconds))
    return min(args.batch, args.cpu_batch if args.cpu_batch > 0 else cap)


def cpu_step_factory(args, batch):
    import titanet_oracle as O
    wave, labels = O.synthetic_batch(batch, seconds=args.seconds, n_classes=N_CLASSES, seed=42)
    lens = ragged_lengths(args, batch, 42).tolist() if args.ragged else [wave.shape[1]] * batch
    if reference_available():
        import warnings
        warnings.filterwarnings("ignore")
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import losses as ref_losses, models as ref_models, transforms as ref_transforms      # the reference's own modules
        torch.manual_seed(42)                                                                 # utils.set_seed (utils.py:281-291)
        head = (ref_losses.CELoss(192, N_CLASSES) if args.loss == "ce"
                else ref_losses.ArcFaceLoss(192, N_CLASSES, scale=ARC_SCALE, margin=ARC_MARGIN))
        model = ref_models.TitaNet.get_titanet(embedding_size=192, n_mels=80, n_mega_blocks=args.blocks, model_size=args.model,
                                               loss_function=head, dropout=args.dropout, device="cpu").train()
        mel = ref_transforms.MelSpectrogram(SAMPLE_RATE, n_fft=512, win_length=400, hop_length=160, n_mels=80,
                                            specaugment_probability=0.0)

        def step():
            mels = [mel({"waveform": w[:n].view(1, -1), "sample_rate": SAMPLE_RATE})["spectrogram"] for w, n in zip(wave, lens)]
            tmax = max(m.shape[-1] for m in mels)
            x = torch.zeros(batch, 80, tmax)                                                  # collate_fn: zero padding
            for i, m in enumerate(mels):
                x[i, :, :m.shape[-1]] = m[0]
            model.zero_grad()
            _, _, loss = model(x, speakers=labels)
            loss.backward()
            return float(loss)

        return step, "reference"
    spec = O.TitaNetSpec.named(args.model, args.blocks, dropout=args.dropout)
    sd = O.synth_state_dict(spec, args.loss, N_CLASSES)
    kw = dict(scale=ARC_SCALE, margin=ARC_MARGIN) if args.loss == "arc" else {}

    def step():
        x, _ = O.collate_pad([O.mel_spectrogram(w[:n].view(1, -1)) for w, n in zip(wave, lens)])
        out = O.titanet_step(sd, spec, x, labels, args.loss, training=True, **kw)
        return float(out[2])

    return step, "port"


def time_cpu(args, batch, steps, warmup, threads=None):
    """(utterances/s, seconds per step, kind, threads) of the CPU implementation; best of `steps` after `warmup`."""
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    step, kind = cpu_step_factory(args, batch)
    for _ in range(warmup):
        step()
    best = float("inf")
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        best = min(best, time.perf_counter() - t0)
    return batch / best, best, kind, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = cpu_batch(args)
    steps = max(1, min(args.steps, 3))
    warmup = 1
    value, dt, kind, threads = time_cpu(args, batch, steps, warmup)
    v2, dt2, _, _ = time_cpu(args, batch, 1, 0, threads=2)            # the reference's own default: workers = 2 (parameters.yml:74, train.py:21-22)
    impl = "unmodified reference modules (baseline/_ref)" if kind == "reference" else "oracle port of the reference modules"
    sample = (f"best of {steps} steps of {batch} utterances x {args.seconds:g} s (per-utterance mel + collate + fwd + bwd), {warmup} warm-up, "
              f"{impl}, torch CPU, {threads} threads")
    line = {
        "impl": "reference", "metric": metric_name(args), "value": round(value, 3),
        "unit": "utterances/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": round(dt * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, batch, device=False),
        "cpu_baseline": {"value": round(value, 3), "unit": "utterances/s", "cores": threads, "kind": kind, "sample": sample},
        "reference_default_threads": {"value": round(v2, 3), "unit": "utterances/s", "cores": 2, "ms_per_step": round(dt2 * 1e3, 2),
                                      "note": "torch.set_num_threads(2), the reference's generic.workers default; 1 step, no warm-up"},
        "e2e": {"value": round(value, 3), "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


ARC_SCALE, ARC_MARGIN = 30, 0.2          # parameters.yml:42-44 / BASELINE.json configs[2]


def ragged_lengths(args, batch, seed):
    """configs[3]: L_i = 16 000 * U{1..seconds} samples."""
    g = torch.Generator().manual_seed(seed + 7)
    return SAMPLE_RATE * torch.randint(1, int(args.seconds) + 1, (batch,), generator=g, dtype=torch.int32)


def workload_config(args, batch, device=True):
    """The workload both arms run; identical dicts when the CPU arm runs the stated batch (TitaNet-S: 64)."""
    loss = f"CE loss ({N_CLASSES} classes)" if args.loss == "ce" else f"ArcFace loss (s={ARC_SCALE}, m={ARC_MARGIN}, {N_CLASSES} classes)"
    dur = f"variable 1-{args.seconds:g}s padded to {args.seconds:g}s" if args.ragged else f"{args.seconds:g}s"
    return {"workload": f"TitaNet-{args.model.upper()}/{args.blocks} fwd+bwd, {loss}, batch {batch} per device, {dur}@16kHz synthetic "
                        f"waveform, mel + encoder + decoder + loss, dropout {args.dropout}",
            "batch_per_gpu": batch, "seconds": args.seconds, "frames": 1 + int(args.seconds * SAMPLE_RATE) // 160}


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
def gemm_work(tag: str):
    """(flops, bytes) of one conv-GEMM launch from its profile tag 'kind R.. Ci.. Co.. K..'."""
    f = dict((t[0] if t[0] != "C" else t[:2], int(t[1:] if t[0] != "C" else t[2:])) for t in tag.split()[1:])
    R, Ci, Co, K = f["R"], f["Ci"], f["Co"], f["K"]
    return 2.0 * R * Ci * Co * K, 4.0 * (R * Ci + R * Co + Co * Ci * K)


def graph_time_kernel(kind: str, R: int, Ci: int, Co: int, B: int, dropout: float, dev, reps: int = 20) -> float:
    """Microseconds per launch of one tensor-core kernel shape, replayed back to back from a CUDA graph."""
    import math
    from titanet_b200._lib import call, ptr
    from titanet_b200._ops import gemm_tc_raw
    T = R // B
    rnd = lambda *s: torch.randn(*s, device=dev)
    x, out = rnd(R, Ci), torch.empty(R, Co, device=dev)
    w = rnd(Co, Ci) / math.sqrt(Ci)
    ws = torch.empty(3, Co, Ci, device=dev)
    import ctypes
    from titanet_b200 import _ops as ops
    from titanet_b200._lib import TnBnBwd

    def bn_bwd_desc(C):          # BatchNorm-backward operand producer over the GEMM's input (C channels)
        zo, g = rnd(R, C), torch.empty(R, C, device=dev)
        v = [0.01 * rnd(C), 0.01 * rnd(C), 0.1 * rnd(C), torch.rand(C, device=dev) + 0.5, torch.rand(C, device=dev) + 0.5,
             torch.empty(C, device=dev), torch.empty(C, device=dev), torch.zeros(C, device=dev)]
        return TnBnBwd(ptr(zo), ptr(v[0]), ptr(v[1]), ptr(v[2]), ptr(v[3]), ptr(v[4]), float(R), ptr(g), ptr(v[7]), ptr(v[5]), ptr(v[6])), (zo, g, v)

    if kind == "tn_wgrad_tc":
        dz, dw = rnd(R, Co), torch.zeros(Co, Ci, device=dev)
        fn = lambda: call("tn_wgrad_tc", ptr(dz), ptr(x), ptr(dw), R, Ci, Co)
    elif kind == "tn_gemm_tc_dwbwd_bn":
        # tag convention: Ci = channels of dZ (reduction), Co = channels of the depthwise input
        call("tn_split_tf32", ptr(rnd(Ci, Co) / math.sqrt(Ci)), ptr(ws), Co, Ci, 1)
        zp, dzp = rnd(R, Co), torch.empty(R, Co, device=dev)
        dww, ddw = rnd(Co, 1, 3), torch.zeros(Co, 3, device=dev)
        acc = torch.zeros(3, Co, device=dev)
        sc, sh = torch.rand(Co, device=dev) + 0.5, 0.1 * rnd(Co)
        seed = torch.tensor([1], dtype=torch.int64, device=dev)
        bnb, keep = bn_bwd_desc(Ci)
        fn = lambda: call("tn_gemm_tc_dwbwd_bn", ptr(x), ptr(ws), ctypes.byref(bnb), ptr(zp), ptr(dzp), ptr(dww), ptr(ddw), acc[0].data_ptr(),
                          acc[1].data_ptr(), acc[2].data_ptr(), ptr(sc), ptr(sh), 1, float(dropout), ptr(seed) if dropout > 0 else None,
                          3, B, T, Ci, Co, 3)
    elif kind == "tn_gemm_tc_bnbwd":
        call("tn_split_tf32", ptr(rnd(Ci, Co) / math.sqrt(Ci)), ptr(ws), Co, Ci, 1)
        bnb, keep = bn_bwd_desc(Ci)
        fn = lambda: call("tn_gemm_tc_bnbwd", ptr(x), ptr(ws), ctypes.byref(bnb), ptr(out), R, Ci, Co, 0)
    elif kind == "tn_gemm_tc_bn":
        call("tn_split_tf32", ptr(w), ptr(ws), Co, Ci, 0)
        bias, st = rnd(Co), torch.empty(2 * Co, dtype=torch.float64, device=dev)
        fold = torch.empty(4, Co, device=dev)
        bnp = [torch.ones(Co, device=dev), torch.zeros(Co, device=dev), torch.zeros(Co, device=dev), torch.ones(Co, device=dev),
               torch.zeros((), dtype=torch.int64, device=dev)]
        bn = ops.make_bn_fold(bnp[0], bnp[1], bnp[2], bnp[3], bnp[4], 0.1, 1e-5, float(R), fold[0], fold[1], fold[2], fold[3])
        scr, keep = ops.scratch(x)
        fn = lambda: call("tn_gemm_tc_bn", ptr(x), ptr(ws), ptr(bias), ptr(out), ptr(st), ctypes.byref(bn), R, Ci, Co, 0, 3, ctypes.byref(scr))
    elif kind == "tn_gemm_tc_dwbwd":
        # tag convention: Ci = channels of dZ (reduction), Co = channels of the depthwise input
        call("tn_split_tf32", ptr(rnd(Ci, Co) / math.sqrt(Ci)), ptr(ws), Co, Ci, 1)
        zp, dzp = rnd(R, Co), torch.empty(R, Co, device=dev)
        dww, ddw = rnd(Co, 1, 3), torch.zeros(Co, 3, device=dev)
        acc = torch.zeros(3, Co, device=dev)
        sc, sh = torch.rand(Co, device=dev) + 0.5, 0.1 * rnd(Co)
        seed = torch.tensor([1], dtype=torch.int64, device=dev)
        fn = lambda: call("tn_gemm_tc_dwbwd", ptr(x), ptr(ws), ptr(zp), ptr(dzp), ptr(dww), ptr(ddw), acc[0].data_ptr(),
                          acc[1].data_ptr(), acc[2].data_ptr(), ptr(sc), ptr(sh), 1, float(dropout),
                          ptr(seed) if dropout > 0 else None, 3, B, T, Ci, Co, 3, 3)
    elif kind == "tn_gemm_tc_dwfwd":
        # depthwise conv (+ BN/ReLU/dropout on load) as the GEMM's operand producer: reads z, writes u (side output) and Z
        call("tn_split_tf32", ptr(w), ptr(ws), Co, Ci, 0)
        u = torch.empty(R, Ci, device=dev)
        dww, dwb, bias = rnd(Ci, 1, 3), rnd(Ci), rnd(Co)
        sc, sh = torch.rand(Ci, device=dev) + 0.5, 0.1 * rnd(Ci)
        seed = torch.tensor([1], dtype=torch.int64, device=dev)
        fn = lambda: call("tn_gemm_tc_dwfwd", ptr(x), ptr(ws), ptr(dww), ptr(dwb), ptr(sc), ptr(sh), 1, float(dropout),
                          ptr(seed) if dropout > 0 else None, 3, ptr(bias), ptr(u), ptr(out), None, None, B, T, Ci, Co, 3, None)
    else:
        call("tn_split_tf32", ptr(w), ptr(ws), Co, Ci, 0)
        bias, st = rnd(Co), torch.zeros(2 * Co, dtype=torch.float64, device=dev)
        fn = lambda: gemm_tc_raw(x, ws, bias, out, st, R, Ci, Co, 0, 3)
    fn()
    torch.cuda.synchronize()
    g, side = torch.cuda.CUDAGraph(), torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def step_traffic_model(args, B):
    """Bytes the step's kernels move by design (sum of every kernel's unique inputs + outputs), against SURVEY section 8d's
    compulsory figure.  Per mega-block, in units of one [B, H, T] fp32 tensor A_H:
      forward 17 = skip GEMM 2 + 3 x (depthwise 2 + pointwise GEMM 2) + squeeze / excitation / tail in one kernel 3
                   (18 = squeeze 1 + tail 3 where the fused kernel's shared-memory tile does not fit: TitaNet-L at 8 s)
      backward 39 = tail pass 1 3 + tail pass 2 6 + 3 x (BatchNorm-backward / dgrad / depthwise-backward kernel 5 + wgrad 2)
                    + skip (BatchNorm-backward / dgrad 4 + wgrad 2) + gradient sum of the block input 3
    (round 1: 18 + 43).  Prolog / epilog / pooling / decoder: epilog GEMM and its backward, the materialised epilog activation
    and the attentive pooling move ~13 A_1536 fwd+bwd (16 before the two gradients of the epilog activation were summed on
    load by tn_act_bwd2); the mel front end 4 L + A_80."""
    T = 1 + int(args.seconds * SAMPLE_RATE) // 160
    H = {"s": 256, "m": 512, "l": 1024}[args.model.lower()]
    a_h, a_e, a_m = 4.0 * B * T * H, 4.0 * B * T * 1536, 4.0 * B * T * 80
    cpc = 64 if (H + 63) // 64 <= 8 else 128                                   # tn_se_tail_fwd's plan (se_tail.cu: se_fused_plan)
    parts = max(1, min(4, 8 // ((H + cpc - 1) // cpc)))
    fused_tail = ((T + parts - 1) // parts) * cpc * 4 <= 160 * 1024
    moved = args.blocks * (56 if fused_tail else 57) * a_h + 13 * a_e + 6 * a_h + 8 * a_m + 4.0 * B * args.seconds * SAMPLE_RATE
    compulsory = 3.0 * (args.blocks * 11 * a_h + (a_m + a_h) + (a_h + a_e) + 2 * a_e + 4.0 * B * args.seconds * SAMPLE_RATE + a_m)
    return {"step_traffic_bytes": moved, "compulsory_bytes": compulsory, "ratio": round(moved / compulsory, 3)}


def metric_name(args) -> str:
    """BASELINE.json's metric for the default workload; other --model / --seconds runs (parity-test configurations) say so."""
    return f"utterances/sec (TitaNet-{args.model.upper()} fwd+bwd, {args.seconds:g}s@16kHz)"


def run_ours(args):
    import torch.distributed as dist
    from titanet_b200 import _lib, losses, models, transforms

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL writes its NCCL_DEBUG output to stdout (at WARN / VERSION level: "NCCL version 2.28.9+cuda12.9"); stdout carries ONE
        # JSON line, so the log goes to a per-process file unless the caller chose one
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/titanet_b200_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(42)
    B, L = args.batch, int(args.seconds * SAMPLE_RATE)
    head = losses.CELoss(192, N_CLASSES) if args.loss == "ce" else losses.ArcFaceLoss(192, N_CLASSES, scale=ARC_SCALE, margin=ARC_MARGIN)
    model = models.TitaNet.get_titanet(192, 80, args.blocks, args.model, loss_function=head,
                                       dropout=args.dropout, device=dev).train()
    mel = transforms.MelSpectrogram(SAMPLE_RATE, n_fft=512, win_length=400, hop_length=160, n_mels=80, specaugment_probability=0.0)
    params = [p for p in model.parameters()]
    g = torch.Generator().manual_seed(42 + rank)
    wave_h = (0.1 * torch.randn(B, L, generator=g)).pin_memory()
    labels_h = torch.randint(0, N_CLASSES, (B,), generator=g).pin_memory()
    wave_d, labels_d = wave_h.to(dev), labels_h.to(dev)
    lens_h = ragged_lengths(args, B, 42 + rank).pin_memory() if args.ragged else None
    lens_d = lens_h.to(dev) if args.ragged else None

    from titanet_b200.engine import GradAllReduce, GraphedTrainStep
    gts = GraphedTrainStep(model, mel, B, L, dev, use_graph=not args.no_graph, warmup=max(3, args.warmup), lengths=lens_d)
    allreduce = GradAllReduce(params, world, arena=gts.arena)     # every gradient lives in the step's arena: one in-place NCCL call

    def step_resident():
        gts.load(wave_d, labels_d, lens_d)        # device -> device: inputs are already in HBM
        loss = gts.run()
        allreduce()
        return loss

    def step_e2e():
        # every step: one pinned host -> device copy of a step's inputs (issued on the copy stream for the NEXT step while this
        # one computes: double-buffered staging, engine.GraphedTrainStep.prefetch) and a D2H read of this step's loss
        loss = gts.run_prefetched()
        allreduce()
        gts.prefetch(wave_h, labels_h, lens_h)
        return float(loss.detach())       # D2H read of the loss (synchronises)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    ms_step = timed(step_resident, args.steps)
    launches = gts.launches_per_step
    clocks = sampler.stop() if sampler else None
    gts.prefetch(wave_h, labels_h, lens_h)        # inputs of the first e2e step (each timed step issues the copy for the next)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    def step(wave, labels):               # eager body for the per-kernel profile below
        gts.wave.copy_(wave)
        gts.labels.copy_(labels)
        gts._body()

    # per-kernel device time: a separate pass with CUDA events around every launch
    roof = None
    if rank == 0:
        pk = peaks()
        _lib.profile_start()
        prof_steps = min(args.steps, 5)
        for _ in range(prof_steps):
            step(wave_d, labels_d)
        prof = _lib.profile_stop()
        total_ms = sum(ms for _, ms in prof.values())
        groups = {}
        for key, (n, ms) in prof.items():
            name = key.split("[")[0]
            gn, gms = groups.get(name, (0, 0.0))
            groups[name] = (gn + n, gms + ms)
        # The eager pass above ranks the kernels (its per-launch times include host launch gaps).  The dominant
        # tensor-core kernel shape is then timed live WITHOUT host gaps: 20 back-to-back launches replayed from
        # a CUDA graph, CUDA events around the replay, on the launching stream.
        gemm_keys = [(k, v) for k, v in prof.items() if k.startswith(("tn_gemm_tc", "tn_wgrad_tc"))]
        if gemm_keys:
            key, (n, _) = max(gemm_keys, key=lambda kv: kv[1][1])
            kind, tag = key.split("[")[0], key[key.index("[") + 1:-1]
            flops, byts = gemm_work(tag)
            f = dict((t[0] if t[0] != "C" else t[:2], int(t[1:] if t[0] != "C" else t[2:])) for t in tag.split()[1:])
            per_launch_s = graph_time_kernel(kind, f["R"], f["Ci"], f["Co"], B, args.dropout, dev) * 1e-6
            if kind == "tn_gemm_tc_dwbwd":          # reads dZ and z_prev, writes dz_prev (+ weights); du never leaves the SM
                byts = 4.0 * (f["R"] * f["Ci"] + 2 * f["R"] * f["Co"] + f["Ci"] * f["Co"])
            if kind == "tn_gemm_tc_dwbwd_bn":       # + the BatchNorm backward: reads dZ, z (of this conv), z_prev; writes g, dz_prev
                byts = 4.0 * (3 * f["R"] * f["Ci"] + 2 * f["R"] * f["Co"] + f["Ci"] * f["Co"])
            if kind == "tn_gemm_tc_bnbwd":          # reads dZ, z; writes g, dX
                byts = 4.0 * (3 * f["R"] * f["Ci"] + f["R"] * f["Co"] + f["Ci"] * f["Co"])
            if kind == "tn_gemm_tc_dwfwd":          # reads z, writes u (kept for the backward wgrad) and Z (+ weights)
                byts = 4.0 * (2 * f["R"] * f["Ci"] + f["R"] * f["Co"] + f["Ci"] * f["Co"])
            traffic = None                          # DRAM bytes per launch of this kernel from the committed ncu --set full capture
            try:
                with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as fh:
                    traffic = json.load(fh).get(kind)
            except (OSError, ValueError):
                pass
            tf32_peak = pk["tf32_tflops"]
            ach_tf = flops / per_launch_s / 1e12
            ach_gb = byts / per_launch_s / 1e9
            launches_per_step = n // prof_steps
            roof = {"kernel": key, "launches_per_step": launches_per_step, "us_per_launch": round(per_launch_s * 1e6, 2),
                    "share_of_step": round(launches_per_step * per_launch_s * 1e3 / ms_step, 4),
                    # TitaNet-S: AI = 2*R*Ci*Co / bytes ~ 64 FLOP/B < the TF32 ridge (~104 FLOP/B) -> HBM is the roofline
                    # of the algorithm; the fp32-equivalent split arithmetic (TF32 main product + one 16-bit correction MMA over a doubled K:
                    # scaled fp16 in forward GEMMs, bf16 in gradient GEMMs) issues 2 MMAs per algorithmic MAC.
                    "bound": "hbm", "achieved": round(ach_gb, 1), "peak": round(pk["hbm_gbs"], 1), "unit": "GB/s",
                    "frac": round(ach_gb / pk["hbm_gbs"], 4), "traffic": traffic, "peak_source": pk["source"],
                    "algorithmic_bytes_per_launch": byts, "algorithmic_flops_per_launch": flops,
                    "algorithmic_tflop_s": round(ach_tf, 2), "tf32_peak_tflop_s": round(tf32_peak, 1),
                    "tf32_peak_source": pk["tf32_source"],
                    "tensor_frac_of_tf32_peak": round(ach_tf / tf32_peak, 4),
                    "mma_issue_factor": 1 if kind == "tn_wgrad_tc" else 2,
                    "timing": "CUDA events around a CUDA graph of 20 back-to-back launches of this shape",
                    "eager_ms_per_step_by_entry_point": {k: round(v[1] / prof_steps, 3) for k, v in
                                                         sorted(groups.items(), key=lambda kv: -kv[1][1])}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * B / (ms_step * 1e-3)
    e2e = world * B / (ms_e2e * 1e-3)
    line = {
        "metric": metric_name(args), "value": round(value, 1), "unit": "utterances/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms_step, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, B),
        "l2": "per-step working set (~2 GB of activations) >> 126 MB L2; no explicit flush",
        "e2e": {"value": round(e2e, 1), "unit": "utterances/s", "h2d_bytes_per_step": B * L * 4 + B * 8 + (B * 4 if args.ragged else 0), "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e, 3)},
        "gpu_launches": launches, "cuda_graph": not args.no_graph, "traffic_model": step_traffic_model(args, B),
        "hbm_peak_gb": round(torch.cuda.max_memory_reserved(dev) / 2**30, 2), "clocks": clocks, "roofline": roof,
    }
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_batch(args)
        v, dt, kind, threads = time_cpu(args, cb, 2, 1)
        impl = "unmodified reference modules (baseline/_ref)" if kind == "reference" else "oracle port of the reference modules"
        line["cpu_baseline"] = {"value": round(v, 3), "unit": "utterances/s", "cores": threads, "kind": kind,
                                "sample": f"best of 2 steps of {cb} utterances x {args.seconds:g} s (per-utterance mel + collate + fwd + bwd) after "
                                          f"1 warm-up, {impl}, torch CPU, {threads} threads"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
