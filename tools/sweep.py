"""BASELINE.json configs[3]/[4]: full `VISinger.forward(infer=True)` over N synthetic mixed-length utterances, sharded by
utterance across the ranks of one box (no collective on the data path), from "token tensors on the host" to "all
waveforms on the host".

    python tools/sweep.py --utterances 512 [--precision bf16]                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py --utterances 512
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from model_inputs import full_hparams, synth_utterances
from visinger_b200.models.visinger import VISinger
from visinger_b200.sharding import bucket_by_length, max_over_ranks, shard_utterances


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utterances", type=int, default=512)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--max-frames-per-batch", type=int, default=16000)
    ap.add_argument("--max-batch", type=int, default=64)
    ap.add_argument("--eager", action="store_true", help="plain forward() per batch instead of one CUDA graph per batch shape")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    rng = np.random.default_rng(1234)   # CSD-like length distribution
    lengths = np.clip(np.round(80 * rng.lognormal(np.log(5.6), 0.45, args.utterances)), 120, 1280).astype(int).tolist()
    mine = shard_utterances(lengths, world)[rank]
    batches = bucket_by_length(mine, lengths, args.max_frames_per_batch, args.max_batch)

    torch.manual_seed(1234)
    model = VISinger(73, 117, 132, full_hparams(), precision=args.precision).eval()
    gen = torch.Generator().manual_seed(7)
    for f in range(4):                  # the reference zero-initialises `post`; make the flow non-trivial
        post = model.flow.flows[2 * f].post
        post.weight.data.copy_(0.05 * torch.randn(post.weight.shape, generator=gen))
    model = model.to(dev)

    host_batches = [synth_utterances(seed=1000 + i, n=len(b), lengths=[lengths[j] for j in b]) for i, b in enumerate(batches)]
    host_batches = [{k: v.pin_memory() for k, v in hb.items()} for hb in host_batches]

    def run_all():
        outs = []
        for hb in host_batches:
            d = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
            if args.eager:
                out = model(d["text_tokens"], d["note_pitch"], d["note_dur"], d["mel2ph"], spk_id=d["spk_ids"], infer=True)
            else:   # one CUDA graph per batch shape (captured in the warm-up pass)
                out = model.forward_graphed(d["text_tokens"], d["note_pitch"], d["note_dur"], d["mel2ph"], spk_id=d["spk_ids"])
            outs.append(out["wav_out"].to("cpu", non_blocking=True))
        torch.cuda.synchronize()
        return outs

    run_all()                           # warm-up (weight pack, lazy module load)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = run_all()
    dt = max_over_ranks(time.perf_counter() - t0, dev)
    audio = sum(lengths) * 300 / 24000.0
    padded = sum(len(b) * max(lengths[j] for j in b) for b in batches)
    if rank == 0:
        print(json.dumps({"workload": f"full VISinger.forward(infer=True), {args.utterances} mixed-length utterances "
                                      f"({audio:.0f} s audio), host tokens -> host waveforms",
                          "n_gpus": world, "precision": args.precision, "seconds": dt, "audio_s_per_s": audio / dt,
                          "batches_rank0": len(batches), "padding_overhead_rank0": padded / max(1, sum(lengths[j] for j in mine)),
                          "wav_finite": bool(all(torch.isfinite(o).all() for o in outs))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
