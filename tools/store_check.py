"""Round-trip check of kernel 4 at a bench-like shape: every frame of a few samples is decoded by the
oracle (chunk -> LZ4 -> unshuffle) and compared with the oracle's records.
python tools/store_check.py [variants] [samples] [mix]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from haplohyped_varawareml_b200 import capi

V = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 300
mix = int(sys.argv[3]) if len(sys.argv) > 3 else 0
crs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
spec = capi.synth_spec(V, S, seed=7, mix=mix)
text = capi.synth_header(spec) + capi.synth_host(spec)
ora = oracle.parse_text(text, "*", "chr22")
body = text[text.index(b"\n", text.index(b"#CHROM")) + 1:]
p = capi.Parse.from_host(body, S, region="chr22")
for cr0 in crs:
    fr = p.compress(cr0)
    cr = int(fr.info.chunk_records)
    bad = 0
    tot = 0
    for s in sorted(set([0, 1, S // 2, S - 1] + list(range(0, S, max(1, S // 16))))):
        frames = fr.sample(s)
        rec = oracle.records_from_columns(ora["chrom"], ora["start"], ora["stop"], ora["ref"], ora["alt"], ora["gt0"][s], ora["gt1"][s])
        raw = rec.tobytes() + b"\0" * (len(frames) * cr * 35 - rec.nbytes)
        for k, f in enumerate(frames):
            tot += len(f)
            try:
                got = oracle.blosc_chunk_decode(f, cr * 35).tobytes()
            except ValueError:
                got = None
            if got != raw[k * cr * 35:(k + 1) * cr * 35]:
                bad += 1
    print("cr", cr, "chunks", fr.info.n_chunks, "bad frames", bad, "avg frame", fr.info.total_bytes / max(1, fr.info.n_chunks * S))
    fr.close()
    assert bad == 0
