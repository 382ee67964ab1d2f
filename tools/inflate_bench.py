"""GPU BGZF inflate at bench shape: python tools/inflate_bench.py [variants] [samples]
text -> BGZF (zlib, host threads) -> hb_bgzf_inflate (kernel time) ; file -> hb_parse_file with GPU vs CPU inflate"""
import os, sys, time, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from haplohyped_varawareml_b200 import capi

V = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2504
spec = capi.synth_spec(V, S, seed=42)
t0 = time.time()
text = np.frombuffer(capi.synth_header(spec) + capi.synth_host(spec), np.uint8)
t1 = time.time()
bg = capi.bgzf_compress_host(text, 6)
t2 = time.time()
print(json.dumps({"text_bytes": int(text.size), "bgzf_bytes": int(bg.size), "ratio": text.size / bg.size,
                  "synth_s": t1 - t0, "compress_s": t2 - t1, "host_threads": os.cpu_count()}))
for _ in range(2):
    out, ms = capi.bgzf_inflate(bg, with_ms=True)
    print(json.dumps({"inflate_kernel_ms": ms, "out_GBs": text.size / ms / 1e6, "in_GBs": bg.size / ms / 1e6, "equal": out == text.tobytes()}))
del out
d = tempfile.mkdtemp()
path = os.path.join(d, "chr22.filtered.vcf.gz")
bg.tofile(path)
for mode in ("gpu", "cpu", "gpu"):
    os.environ["HB_CPU_INFLATE"] = "1" if mode == "cpu" else "0"
    t0 = time.time()
    p = capi.Parse.from_file(path, region="chr22")
    t1 = time.time()
    i = p.info
    print(json.dumps({"mode": mode, "parse_file_s": t1 - t0, "records": int(i.n_records), "ms_inflate": i.ms_inflate,
                      "compressed_bytes": int(i.compressed_bytes), "calls_per_s": V * S / (t1 - t0)}))
    p.close()
