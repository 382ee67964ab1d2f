"""How much of the GT decoder's time can hide under kernel 4b?  Two parses of the same text on two streams:
thread A re-runs the parse (tokenise + sites + decode), thread B re-runs the frames of the other parse; the
wall time of both together is compared with each alone.  python tools/overlap_probe.py [variants] [samples]"""
import os, sys, time, threading, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from haplohyped_varawareml_b200 import capi

V = int(sys.argv[1]) if len(sys.argv) > 1 else 1_100_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2504
spec = capi.synth_spec(V, S, seed=42)
T = int(capi.lib().hb_synth_body_bytes(spec))
text = torch.empty(T + 256, dtype=torch.uint8, device="cuda")
text[T:].zero_()
capi.check(capi.lib().hb_synth_device(spec, text.data_ptr(), T, 0, None))
torch.cuda.synchronize()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(priority=-1 if len(sys.argv) > 3 and sys.argv[3] == "hi" else 0)
p1 = capi.Parse.from_device(text.data_ptr(), T, S, region="chr22", stream=s1.cuda_stream)
p2 = capi.Parse.from_device(text.data_ptr(), T, S, region="chr22", stream=s2.cuda_stream)
fr2 = p2.compress(0)
for _ in range(2):
    p1.rerun(); fr2.rerun(p2)

def timed(fn, reps=4):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), sum(ts) / len(ts)

a = timed(p1.rerun)
b = timed(lambda: fr2.rerun(p2))

def both():
    bar = threading.Barrier(3)
    def wa(): bar.wait(); p1.rerun()
    def wb(): bar.wait(); fr2.rerun(p2)
    ta, tb = threading.Thread(target=wa), threading.Thread(target=wb)
    ta.start(); tb.start()
    bar.wait()
    t0 = time.perf_counter()
    ta.join(); tb.join()
    return (time.perf_counter() - t0) * 1e3
for _ in range(2): both()
cs = [both() for _ in range(6)]
print(json.dumps({"parse_alone_ms": a, "frames_alone_ms": b, "both_ms_min": min(cs), "both_ms_avg": sum(cs) / len(cs),
                  "sum_alone_min": a[0] + b[0], "split": os.environ.get("HB_DONOR_SPLIT", "lane")}))
