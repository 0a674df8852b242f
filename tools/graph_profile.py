"""Per-kernel device time INSIDE the CUDA-graph replay of the benchmark step (torch.profiler / CUPTI): the real step, warm
caches, no host gaps -- unlike the ncu launch list (cold, serialised).  Prints a table sorted by total time."""
import argparse, collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="s"); ap.add_argument("--blocks", type=int, default=17); ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--seconds", type=float, default=3.0); ap.add_argument("--loss", default="ce"); ap.add_argument("--dropout", type=float, default=0.1)
ap.add_argument("--replays", type=int, default=5); ap.add_argument("--out", default=""); ap.add_argument("--timeline", default="", help="write the launch-by-launch timeline of the last replay (start us, duration us, gap to the previous kernel, name)")
a = ap.parse_args()
from titanet_b200 import losses, models, transforms
from titanet_b200.engine import GraphedTrainStep
dev = torch.device("cuda", 0)
torch.manual_seed(42)
head = losses.CELoss(192, 251) if a.loss == "ce" else losses.ArcFaceLoss(192, 251, scale=30, margin=0.2)
model = models.TitaNet.get_titanet(192, 80, a.blocks, a.model, loss_function=head, dropout=a.dropout, device=dev).train()
mel = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80, specaugment_probability=0.0)
L = int(a.seconds * 16000)
gts = GraphedTrainStep(model, mel, a.batch, L, dev, use_graph=True, warmup=3)
gts.load(0.1 * torch.randn(a.batch, L, device=dev), torch.randint(0, 251, (a.batch,), device=dev))
for _ in range(3):
    gts.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(a.replays):
        gts.run()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time > 0:
        n, t = agg.get(ev.name, (0, 0.0))
        agg[ev.name] = (n + 1, t + ev.device_time)
tot = sum(t for _, t in agg.values())
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
print(f"| kernel | launches/step | us/launch | us/step | share |\n|---|---|---|---|---|")
for name, (n, t) in rows:
    short = name.split("(")[0][:70]
    print(f"| `{short}` | {n / a.replays:.0f} | {t / n:.2f} | {t / a.replays:.1f} | {100 * t / tot:.1f}% |")
print(f"\nsum of kernel time per step: {tot / a.replays / 1e3:.3f} ms over {sum(n for n, _ in agg.values()) / a.replays:.0f} launches")
if a.timeline:
    evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time > 0), key=lambda e: e.time_range.start)
    per = len(evs) // a.replays
    last = evs[-per:]
    t0 = last[0].time_range.start
    with open(a.timeline, "w") as f:
        prev_end = t0
        for ev in last:
            st = ev.time_range.start
            f.write(f"{st - t0:9.1f} {ev.device_time:8.2f} {st - prev_end:6.2f}  {ev.name.split('(')[0][:80]}\n")
            prev_end = st + ev.device_time
if a.out:
    json.dump({k: {"launches_per_step": n / a.replays, "us_per_launch": t / n} for k, (n, t) in rows}, open(a.out, "w"), indent=1)
