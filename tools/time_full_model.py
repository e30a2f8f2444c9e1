"""Where the time of a full `VISinger.forward(infer=True)` goes at bench size (B=16 x T=1000): prior network (PyTorch) vs the
native hot path, CUDA-event timed.   python tools/time_full_model.py [precision]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from model_inputs import full_hparams, synth_utterances
from visinger_b200.models.visinger import VISinger

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = VISinger(73, 117, 132, full_hparams(), precision=prec).eval().to(dev)
hb = synth_utterances(seed=1, n=16, lengths=[1000] * 16)
d = {k: v.to(dev) for k, v in hb.items()}


def ev():
    return torch.cuda.Event(enable_timing=True)


def timeit(fn, n=5):
    fn(); fn()
    a, b = ev(), ev()
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


with torch.no_grad():
    full = timeit(lambda: m(d["text_tokens"], d["note_pitch"], d["note_dur"], d["mel2ph"], spk_id=d["spk_ids"], infer=True))
    prior = timeit(lambda: m.prior(d["text_tokens"], d["note_pitch"], d["note_dur"], d["mel2ph"], None, d["spk_ids"]))
    import time
    t0 = time.perf_counter()
    for _ in range(5):
        m.prior(d["text_tokens"], d["note_pitch"], d["note_dur"], d["mel2ph"], None, d["spk_ids"])
    cpu_issue = (time.perf_counter() - t0) / 5 * 1e3
    torch.cuda.synchronize()
# the same prior network replayed from a CUDA graph: its pure GPU time, without the eager launch overhead
with torch.no_grad():
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            m.prior(d["text_tokens"], d["note_pitch"], d["note_dur"], d["mel2ph"], None, d["spk_ids"])
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = m.prior(d["text_tokens"], d["note_pitch"], d["note_dur"], d["mel2ph"], None, d["spk_ids"])
    graphed = timeit(g.replay)
print(f"prior network from a CUDA graph {graphed:.2f} ms")
with torch.no_grad():
    whole = timeit(lambda: m.forward_graphed(d["text_tokens"], d["note_pitch"], d["note_dur"], d["mel2ph"], spk_id=d["spk_ids"]))
print(f"forward_graphed (whole forward, one graph) {whole:.2f} ms = {16 * 12.5 / whole * 1e3:.0f} audio-s/s")
print(f"full forward {full:.2f} ms | prior network {prior:.2f} ms (CPU issue time {cpu_issue:.2f} ms) | hot path {full - prior:.2f} ms")
