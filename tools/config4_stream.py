"""BASELINE configs[3] at full width on ONE B200, tile-streamed: 500,000 variants x 200,000 samples (100 G genotype calls,
~400 GB of VCF text) with multiallelic / indel sites (dropped by the SNP filter), unphased and missing calls.
The text is generated on the device, slab by slab, into ONE buffer (counter-based generator: the CPU oracle regenerates
any slab byte for byte), each slab goes through the whole path -- locate records, sites, GT decode, Blosc2 frames of
every (donor, chunk) -- and device memory stays O(slab).  A slab is an independent run of variants (own chunks).
python tools/config4_stream.py [variants] [samples] [slab_variants]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from haplohyped_varawareml_b200 import capi

V = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
SLAB = int(sys.argv[3]) if len(sys.argv) > 3 else 12_500
n_slabs = (V + SLAB - 1) // SLAB

def slab_spec(k):
    nv = min(SLAB, V - k * SLAB)
    return capi.synth_spec(nv, S, seed=1000 + k, mix=1, first_pos=10_000_000 + k * SLAB * 35)

cap = max(int(capi.lib().hb_synth_body_bytes(slab_spec(k))) for k in (0, n_slabs - 1)) + (1 << 20)
text = torch.empty(cap + 256, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]

p = fr = None
tot = {"text": 0, "records": 0, "c_out": 0, "ms_parse": 0.0, "ms_store": 0.0, "ms_synth": 0.0, "alg": 0.0}
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
wall0 = time.time()
for k in range(n_slabs):
    sp = slab_spec(k)
    T = int(capi.lib().hb_synth_body_bytes(sp))
    assert T <= cap
    ev[0].record()
    capi.check(capi.lib().hb_synth_device(sp, text.data_ptr(), T, 0, None))
    text[T:T + 256].zero_()
    ev[1].record()
    if p is None:
        p = capi.Parse.from_device(text.data_ptr(), T, S, region="chr22", stream=stream)
    else:
        capi.check(capi.lib().hb_parse_rerun_bytes(p._h, T))
    ev[2].record()
    if fr is None:
        fr = p.compress(977)        # h5py's auto-chunk for a 500,000-record dataset of 35-byte items
    else:
        fr.rerun(p)
    ev[3].record()
    torch.cuda.synchronize()
    i, fi = p.info, fr.info
    tot["text"] += T; tot["records"] += int(i.n_records); tot["c_out"] += int(fi.total_bytes)
    tot["ms_synth"] += ev[0].elapsed_time(ev[1]); tot["ms_parse"] += ev[1].elapsed_time(ev[2]); tot["ms_store"] += ev[2].elapsed_time(ev[3])
    tot["alg"] += T + 2.0 * (2.0 * i.n_records * S + 33.0 * i.n_records) + fi.total_bytes
    if os.environ.get("C4_VERBOSE"):
        print(k, "parse %.2f store %.2f (site %.2f frames %.2f) n_rec %d tok %d" % (ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), fi.ms_site, fi.ms_frames, i.n_records, i.tokenizer_used), flush=True)
    if k == 0:      # parity spot check: the oracle regenerates the first records of this slab on the CPU
        nchk = 3
        head = capi.synth_header(sp) + capi.synth_host(sp, 0, nchk)
        ora = oracle.parse_text(head, "*", "chr22")
        g0, g1 = p.sample(S - 1)
        ok = bool(np.array_equal(g0[:ora["n"]], ora["gt0"][S - 1]) and np.array_equal(g1[:ora["n"]], ora["gt1"][S - 1]))
        start = p.sites()[0]
        ok = ok and bool(np.array_equal(start[:ora["n"]], ora["start"]))
wall = time.time() - wall0
ms = tot["ms_parse"] + tot["ms_store"]
print(json.dumps({"variants": V, "samples": S, "calls": V * S, "slabs": n_slabs, "slab_variants": SLAB, "text_bytes": tot["text"],
                  "records_kept": tot["records"], "c_out_bytes": tot["c_out"], "device_buffer_bytes": cap,
                  "ms_parse": tot["ms_parse"], "ms_store": tot["ms_store"], "ms_synth": tot["ms_synth"], "wall_s": wall,
                  "calls_per_s_path": V * S / (ms / 1e3), "calls_per_s_parse_only": V * S / (tot["ms_parse"] / 1e3),
                  "path_GBs": tot["alg"] / ms / 1e6, "path_frac_of_measured_peak": tot["alg"] / ms / 1e6 / peak,
                  "parity_spot_check": ok}))
