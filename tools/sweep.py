"""BASELINE.json configs[4]: mel + encoder (+ decoder, loss, backward) throughput over the batch size, one GPU.

``python tools/sweep.py [--model s] [--batches 32,64,...] [--out gpurun_out/sweep.jsonl]`` runs ``bench.py`` once per batch
size in a fresh process (no CPU baseline; 5 timed steps after 3 warm-up steps) and collects the JSON lines.  A batch is
skipped when the previous one's peak reservation says it would not fit in HBM (activations grow linearly with the batch;
TitaNet-S at batch 2048 holds ~90 GB of saved activations; TN_RECOMPUTE_U=1 drops a third of them).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="s")
    ap.add_argument("--blocks", type=int, default=17)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--batches", default="32,64,128,256,512,1024,2048")
    ap.add_argument("--hbm-gb", type=float, default=150.0, help="skip a batch predicted to need more than this")
    ap.add_argument("--gpus", type=int, default=1, help="ranks (torchrun) per run: the batch is per GPU (weak scaling)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    last = None           # (batch, peak GB)
    with open(args.out, "w") as f:
        for b in [int(v) for v in args.batches.split(",")]:
            if last and last[1] * b / last[0] > args.hbm_gb:
                print(f"batch {b}: skipped (predicted {last[1] * b / last[0]:.0f} GB)", flush=True)
                continue
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--batch", str(b), "--model", args.model, "--blocks",
                   str(args.blocks), "--seconds", str(args.seconds), "--steps", "5", "--warmup", "3", "--no-cpu-baseline"]
            if args.gpus > 1:
                cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr",
                       "127.0.0.1", "--master-port", "29671"] + cmd[1:] + ["--gpus", str(args.gpus)]
            try:
                res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            except subprocess.TimeoutExpired:
                print(f"batch {b}: timed out", flush=True)
                break
            lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
            if res.returncode != 0 or not lines:
                print(f"batch {b}: failed\n{res.stderr[-2000:]}", flush=True)
                break
            line = json.loads(lines[-1])
            last = (b, line.get("hbm_peak_gb", 0.0))
            r = line["roofline"] or {}
            print(f"batch {b:5d}: {line['value']:9.1f} utt/s  {line['ms_per_step']:8.3f} ms/step  e2e {line['e2e']['value']:9.1f}"
                  f"  peak {last[1]:6.2f} GB  top kernel {r.get('kernel')} frac {r.get('frac')}", flush=True)
            f.write(json.dumps(line) + "\n")
            f.flush()


if __name__ == "__main__":
    main()
