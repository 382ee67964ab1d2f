"""Phase timing of the e2e conversion step (BGZF bytes in pinned host memory -> packed frame image in pinned host memory):
python tools/convert_trace.py [variants] [samples] [reps]   (HB_TRACE=1 adds the library's own phase lines on stderr)"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from haplohyped_varawareml_b200 import capi
V = int(sys.argv[1]) if len(sys.argv) > 1 else 1_100_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2504
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
spec = capi.synth_spec(V, S, seed=42, mix=1 << 8)
T = int(capi.lib().hb_synth_body_bytes(spec))
text = torch.empty(T + 256, dtype=torch.uint8, device="cuda")
capi.check(capi.lib().hb_synth_device(spec, text.data_ptr(), T, 0, None))
torch.cuda.synchronize()
hdr = capi.synth_header(spec)
full = np.empty(len(hdr) + T, np.uint8)
full[:len(hdr)] = np.frombuffer(hdr, np.uint8)
full[len(hdr):] = text[:T].cpu().numpy()
del text
torch.cuda.empty_cache()
bg = capi.bgzf_compress_host(full, 6)
del full
bgp = torch.empty(bg.size, dtype=torch.uint8).pin_memory()
bgp.numpy()[:] = bg
pin = torch.empty(int(17e9 * V / 1.1e6 * S / 2504) + (64 << 20), dtype=torch.uint8).pin_memory()
for r in range(reps):
    t0 = time.perf_counter()
    q = capi.Parse.from_vcf_bytes(bgp.data_ptr(), region="chr22", device=0, nbytes=bgp.numel())
    t1 = time.perf_counter()
    fr = q.compress(0)
    t2 = time.perf_counter()
    tot, offs, sizes = fr.fetch_packed(out=(pin.data_ptr(), pin.numel()))
    t3 = time.perf_counter()
    fr.close(); q.close()
    t4 = time.perf_counter()
    print(json.dumps({"rep": r, "parse_bytes_ms": 1e3 * (t1 - t0), "compress_ms": 1e3 * (t2 - t1), "fetch_packed_ms": 1e3 * (t3 - t2),
                      "close_ms": 1e3 * (t4 - t3), "total_ms": 1e3 * (t4 - t0), "packed_GB": tot / 1e9,
                      "d2h_GBs": tot / 1e9 / (t3 - t2)}), flush=True)
