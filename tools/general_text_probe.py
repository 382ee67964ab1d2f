"""Parse timings on variable-width text (FORMAT GT:GQ:DP, ~30 bytes per call, the shape of the reference's own fixture):
tokenizer with column checkpoints + general (tab-scanning) decode path.  python tools/general_text_probe.py [variants] [samples]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import synth, oracle
from haplohyped_varawareml_b200 import capi
V = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2504
text, samples = synth.random_vcf(V, S, seed=5, fmt="GT:GQ:DP", kinds="mixed", site_mix=False)
body = synth.body_of(text)
p = capi.Parse.from_host(body, S, region="chr22")
for _ in range(3):
    p.rerun()
i = p.info
ora = oracle.parse_text(text, samples[S // 2], "chr22")
g0, g1 = p.sample(S // 2)
ms = i.ms_tokenize + i.ms_sites + i.ms_decode
print(json.dumps({"variants": V, "samples": S, "text_bytes": len(body), "bytes_per_call": len(body) / (V * S), "tokenizer": i.tokenizer_used,
                  "ms_tokenize": i.ms_tokenize, "ms_sites": i.ms_sites, "ms_decode": i.ms_decode, "text_GBs": len(body) / ms / 1e6,
                  "calls_per_s": V * S / ms * 1e3, "parity_one_sample": bool(np.array_equal(g0, ora["gt0"]) and np.array_equal(g1, ora["gt1"]))}))
