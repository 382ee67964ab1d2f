"""Kernel 0 alone: BGZF members (stock zlib level 6, all host threads) of bench-shaped text -> text on the GPU.
python tools/inflate_bench.py [variants] [samples] [af_skew]   prints text GB/s of the kernel and checks the bytes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from haplohyped_varawareml_b200 import capi
V = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2504
skew = int(sys.argv[3]) if len(sys.argv) > 3 else 1
spec = capi.synth_spec(V, S, seed=42, mix=skew << 8)
text = capi.synth_header(spec) + capi.synth_host(spec)
bg = capi.bgzf_compress_host(text, 6)
out = []
for _ in range(4):
    got, ms = capi.bgzf_inflate(bg.tobytes(), with_ms=True)
    out.append(ms)
ok = got == text
ms = min(out[1:])
print(json.dumps({"variants": V, "samples": S, "text_bytes": len(text), "bgzf_bytes": int(bg.size), "ratio": len(text) / bg.size,
                  "kernel_ms": out, "text_gbs": len(text) / ms / 1e6, "bit_exact": bool(ok)}))
