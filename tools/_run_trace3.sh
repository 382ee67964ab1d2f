HB_TRACE=1 timeout 500 python - <<'PY' 2>&1 | grep "stream_bgzf_host\|^{" | head -40
import os, sys, time, json, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from haplohyped_varawareml_b200 import capi
V, S = 1_100_000, 2504
spec = capi.synth_spec(V, S, seed=42)
T = int(capi.lib().hb_synth_body_bytes(spec))
text = torch.empty(T + 256, dtype=torch.uint8, device="cuda")
capi.check(capi.lib().hb_synth_device(spec, text.data_ptr(), T, 0, None))
hdr = capi.synth_header(spec)
full = np.empty(len(hdr) + T, np.uint8); full[:len(hdr)] = np.frombuffer(hdr, np.uint8); full[len(hdr):] = text[:T].cpu().numpy()
del text
bg = capi.bgzf_compress_host(full, 6); del full
bgp = torch.empty(bg.size, dtype=torch.uint8).pin_memory(); bgp.numpy()[:] = bg; del bg
out0 = torch.empty((S, V), dtype=torch.int8).pin_memory(); out1 = torch.empty((S, V), dtype=torch.int8).pin_memory()
sites = [torch.empty(V, dtype=torch.int32).pin_memory(), torch.empty(V, dtype=torch.int32).pin_memory(), torch.empty(V, dtype=torch.uint8).pin_memory(), torch.empty(V, dtype=torch.uint8).pin_memory()]
nrec = C.c_uint64()
for slab, outs in [(1 << 30, True)] * 4 + [(1 << 30, False)] * 3 + [(256 << 20, True)] * 4:
    a = [out0.data_ptr(), out1.data_ptr()] if outs else [None, None]
    s = [x.data_ptr() for x in sites] if outs else [None] * 4
    t0 = time.perf_counter()
    capi.check(capi.lib().hb_parse_stream_bgzf_host(bgp.data_ptr(), bgp.numel(), b"chr22", 1, 0, slab, a[0], a[1], V, s[0], s[1], s[2], s[3], None, None, C.byref(nrec), None))
    print(json.dumps({"slab_MiB": slab >> 20, "outs": outs, "ms": (time.perf_counter() - t0) * 1e3}), flush=True)
PY
