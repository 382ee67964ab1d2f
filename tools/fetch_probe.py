"""hb_frames_fetch_packed against a raw device->pinned copy of the same size, bench shape.  python tools/fetch_probe.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from haplohyped_varawareml_b200 import capi
V, S = 1_100_000, 2504
spec = capi.synth_spec(V, S, seed=42, mix=1 << 8)
T = int(capi.lib().hb_synth_body_bytes(spec))
text = torch.empty(T + 256, dtype=torch.uint8, device="cuda"); text[T:].zero_()
capi.check(capi.lib().hb_synth_device(spec, text.data_ptr(), T, 0, None))
p = capi.Parse.from_device(text.data_ptr(), T, S, region="chr22")
fr = p.compress(0)
tot = int(fr.info.total_bytes)
cap = int(tot * 1.02) + (64 << 20)
pin = torch.empty(cap, dtype=torch.uint8).pin_memory()
dev = torch.empty(tot, dtype=torch.uint8, device="cuda")
out = {"total_bytes": tot}
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(round((time.perf_counter() - t0) * 1e3, 1))
    return ts
out["raw_d2h_ms"] = t(lambda: pin[:tot].copy_(dev, non_blocking=True))
out["raw_d2h_GBs"] = tot / min(out["raw_d2h_ms"]) / 1e6
for mode, name in ((1, "device_gather"), (2, "host_assembly")):
    capi.lib().hb_set_fetch_mode(mode)
    out["fetch_packed_%s_ms" % name] = t(lambda: fr.fetch_packed(out=(pin.data_ptr(), cap)))
    out["fetch_packed_%s_GBs" % name] = tot / min(out["fetch_packed_%s_ms" % name]) / 1e6
for th in (4, 8, 16):
    capi.lib().hb_set_host_threads(th)
    out["host_assembly_%d_threads_ms" % th] = t(lambda: fr.fetch_packed(out=(pin.data_ptr(), cap)))
capi.lib().hb_set_host_threads(0); capi.lib().hb_set_fetch_mode(0)
out["layout_only_ms"] = t(lambda: fr.layout())
print(json.dumps(out))
