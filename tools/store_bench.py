"""Kernel-4 timing at a given shape: python tools/store_bench.py [variants] [samples] [reps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from haplohyped_varawareml_b200 import capi

V = int(sys.argv[1]) if len(sys.argv) > 1 else 1_100_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2504
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mix = int(sys.argv[4]) if len(sys.argv) > 4 else 0            # bits 8-15: ALT-frequency skew (see hb_synth_spec)
spec = capi.synth_spec(V, S, seed=42, mix=mix)
T = int(capi.lib().hb_synth_body_bytes(spec))
text = torch.empty(T + 256, dtype=torch.uint8, device="cuda")
text[T:].zero_()
capi.check(capi.lib().hb_synth_device(spec, text.data_ptr(), T, 0, None))
torch.cuda.synchronize()
p = capi.Parse.from_device(text.data_ptr(), T, S, region="chr22")
fr = p.compress(0)
for r in range(reps):
    fr.rerun(p)
    i = fr.info
    pi = p.info
    alg = 2.0 * i.n_records * S + 33.0 * i.n_records + i.total_bytes
    ms = i.ms_site + i.ms_frames
    print(json.dumps({"records": i.n_records, "chunks": i.n_chunks, "cr": i.chunk_records, "C_out": i.total_bytes,
                      "padded": i.padded_bytes, "ratio": i.raw_bytes / max(1, i.total_bytes),
                      "ms_site": i.ms_site, "ms_frames": i.ms_frames, 
                      "ms_total": ms, "alg_GBs": alg / ms / 1e6, "frames_alg_GBs": (2.0 * i.n_records * S + i.total_bytes) / max(1e-9, i.ms_frames) / 1e6,
                      "bytes_per_frame": i.total_bytes / max(1, i.n_chunks * S), "mix": mix, "site_lz4_per_chunk": i.site_lz4_bytes / max(1, i.n_chunks)}))
