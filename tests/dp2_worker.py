"""Worker of tests/test_gpu_dp2.py: one process per GPU (torchrun env), NCCL.  Each rank runs the captured training step of the
bench path (waveform -> mel -> TitaNet -> CE -> backward, engine.GraphedTrainStep) on ITS shard of the utterance batch, the
gradient arena is all-reduced in place (engine.GradAllReduce), and every rank checks:
  * its shard's embeddings / loss against the CPU oracle evaluated on that shard alone (BatchNorm statistics are per replica,
    SURVEY.md section 8e) and its BatchNorm running statistics against the oracle's for that shard;
  * the all-reduced gradients against the oracle gradients of all shards, averaged (fp64 oracle as truth, fp32 oracle's own
    error as the yard-stick).
Prints one JSON line per rank; exit code 0 = all checks passed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import titanet_oracle as O  # noqa: E402  (checker only)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/titanet_b200_dp2_%h_%p.log")
    dist.init_process_group("nccl", device_id=dev)
    from titanet_b200 import losses, models, transforms
    from titanet_b200.engine import GradAllReduce, GraphedTrainStep

    spec = O.TitaNetSpec.named("s", 2)
    sd = O.synth_state_dict(spec, "ce", 251)
    B, seconds = 4, 1.0
    shards = [O.synthetic_batch(B, seconds=seconds, n_classes=251, seed=100 + r) for r in range(world)]
    model = models.TitaNet.get_titanet(192, 80, 2, "s", loss_function=losses.CELoss(192, 251), dropout=0.0)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).train()
    mel = transforms.MelSpectrogram(16000, n_fft=512, win_length=400, hop_length=160, n_mels=80, specaugment_probability=0.0)
    wave, labels = shards[rank]
    gts = GraphedTrainStep(model, mel, B, wave.shape[1], dev, use_graph=True, warmup=2)
    model.load_state_dict(sd, strict=True)            # the warm-up steps moved the BatchNorm running statistics
    allreduce = GradAllReduce([p for p in model.parameters()], world, arena=gts.arena)
    gts.load(wave.to(dev), labels.to(dev))
    loss = gts.run()
    allreduce()
    torch.cuda.synchronize()

    def rel(a, b, floor=1e-30):
        a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
        return float((a - b).abs().max() / b.abs().max().clamp_min(floor))

    outs32, outs64 = [], []
    for w, y in shards:
        x = torch.cat([O.mel_spectrogram(u.view(1, -1)) for u in w])
        outs32.append(O.titanet_step({k: v.clone() for k, v in sd.items()}, spec, x, y, "ce"))
        outs64.append(O.titanet_step(O.synth_state_dict(spec, "ce", 251, dtype=torch.float64), spec, x.double(), y, "ce"))
    mine = outs32[rank]
    e_emb = rel(gts.emb, mine[0])
    e_loss = abs(float(loss) - float(mine[2])) / abs(float(mine[2]))
    keys = list(outs64[0][3])
    avg64 = {k: sum(o[3][k] for o in outs64) / world for k in keys}
    avg32 = {k: sum(o[3][k].double() for o in outs32) / world for k in keys}
    gmax = max(float(v.abs().max()) for v in avg64.values())
    worst = max(rel(p.grad, avg64[k], floor=1e-3 * gmax) for k, p in model.named_parameters())
    worst32 = max(rel(avg32[k], avg64[k], floor=1e-3 * gmax) for k in keys)
    # per-replica BatchNorm buffers: this rank's running mean follows ITS shard
    rm_key = "encoder.prolog.conv_block.1.running_mean"
    e_rm = rel(model.state_dict()[rm_key], mine[4][rm_key])
    other = outs32[(rank + 1) % world][4][rm_key]
    differs = rel(model.state_dict()[rm_key], other) > 10 * max(e_rm, 1e-7)
    ok = e_emb < 1e-3 and e_loss < 1e-3 and worst <= max(3.0 * worst32, 2e-3) and e_rm < 1e-4 and differs
    # every rank holds the same averaged gradients after the exchange
    probe = torch.stack([p.grad.double().sum() for p in model.parameters()]).sum().reshape(1)
    gathered = [torch.zeros_like(probe) for _ in range(world)]
    dist.all_gather(gathered, probe)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    print(json.dumps({"rank": rank, "world": world, "emb": e_emb, "loss": e_loss, "grad_worst": worst, "grad_worst_fp32_oracle": worst32,
                      "running_mean": e_rm, "replica_buffers_differ": bool(differs), "ranks_agree": bool(same), "ok": bool(ok and same)}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if (ok and same) else 1)


if __name__ == "__main__":
    main()
