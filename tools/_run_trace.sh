HB_TRACE=1 timeout 500 python - <<'PY' 2>&1 | grep "hb_parse_stream_host\|^{" | head -40
import os, sys, time, json, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from haplohyped_varawareml_b200 import capi
V, S = 1_100_000, 2504
spec = capi.synth_spec(V, S, seed=42)
T = int(capi.lib().hb_synth_body_bytes(spec))
text = torch.empty(T + 256, dtype=torch.uint8, device="cuda")
capi.check(capi.lib().hb_synth_device(spec, text.data_ptr(), T, 0, None))
host = torch.empty(T, dtype=torch.uint8).pin_memory(); host.copy_(text[:T]); torch.cuda.synchronize()
del text
out0 = torch.empty((S, V), dtype=torch.int8).pin_memory(); out1 = torch.empty((S, V), dtype=torch.int8).pin_memory()
sites = [torch.empty(V, dtype=torch.int32).pin_memory(), torch.empty(V, dtype=torch.int32).pin_memory(), torch.empty(V, dtype=torch.uint8).pin_memory(), torch.empty(V, dtype=torch.uint8).pin_memory()]
nrec = C.c_uint64()
for slab in (256 << 20, 256 << 20, 256 << 20, 256 << 20, 1 << 30, 1 << 30, 1 << 30):
    o = capi.Parse._opts(S, "chr22", False, True, 0, 0, None)
    t0 = time.perf_counter()
    capi.check(capi.lib().hb_parse_stream_host(host.data_ptr(), T, C.byref(o), slab, out0.data_ptr(), out1.data_ptr(), V, sites[0].data_ptr(), sites[1].data_ptr(), sites[2].data_ptr(), sites[3].data_ptr(), None, None, C.byref(nrec), None))
    print(json.dumps({"slab_MiB": slab >> 20, "ms": (time.perf_counter() - t0) * 1e3}), flush=True)
PY
