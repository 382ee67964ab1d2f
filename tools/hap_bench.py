"""Kernel 5 (batched haplotype one-hot) at BASELINE configs[4] shapes: python tools/hap_bench.py
B in {32, 1024} x L in {1000, 131072}, C = 5; algorithmic bytes = B*L (window) + 2*B*L*C*4 (two float32 one-hots)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from haplohyped_varawareml_b200 import haplotype_dataset as hd
from haplohyped_varawareml_b200.common_utils import parse_encode_dict

dev = torch.device("cuda:0")
rng = np.random.default_rng(1)
chrom_len = 64_000_000
seq = torch.from_numpy(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=chrom_len)).to(dev)
n_rec = chrom_len // 1000                                   # ~1 SNP / kb
start = torch.from_numpy(np.sort(rng.integers(0, chrom_len, n_rec)).astype(np.uint32).view(np.int32)).to(dev)
ref = torch.from_numpy(rng.choice(np.frombuffer(b"ACGT", np.uint8), n_rec)).to(dev)
alt = torch.from_numpy(rng.choice(np.frombuffer(b"ACGT", np.uint8), n_rec)).to(dev)
p1 = torch.from_numpy(rng.integers(0, 2, n_rec).astype(np.int8)).to(dev)
p2 = torch.from_numpy(rng.integers(0, 2, n_rec).astype(np.int8)).to(dev)
spec = parse_encode_dict(None)
lut = hd.build_lut(spec, True)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
for B, L in ((32, 1000), (1024, 1000), (32, 131072), (1024, 131072)):
    ws = rng.integers(0, chrom_len - L, B)
    cols = [(start.data_ptr(), ref.data_ptr(), alt.data_ptr(), p1.data_ptr(), p2.data_ptr(), n_rec)] * B
    args = (B, L, 5, [seq.data_ptr() + int(w) for w in ws], [L] * B, [int(w) for w in ws], cols, lut, dev)
    # kernel alone: outputs and per-item metadata resident, C ABI called directly
    import ctypes as C
    from haplohyped_varawareml_b200 import capi
    hap1 = torch.empty((B, L, 5), dtype=torch.float32, device=dev)
    hap2 = torch.empty((B, L, 5), dtype=torch.float32, device=dev)
    meta = np.zeros((9, B), dtype=np.uint64)
    meta[0] = args[3]; meta[1] = args[4]; meta[2] = args[5]
    for b_, c_ in enumerate(cols):
        meta[3:9, b_] = c_
    m = torch.from_numpy(meta.view(np.int64)).to(dev)
    lens32 = torch.from_numpy(np.asarray(args[4], np.uint32).view(np.int32)).to(dev)
    ws32 = torch.from_numpy(np.asarray(args[5], np.uint32).view(np.int32)).to(dev)
    lut_t = torch.from_numpy(lut.copy()).to(dev)
    hb = capi.HapBatch()
    hb.B, hb.L, hb.C = B, L, 5
    hb.item_seq, hb.item_len, hb.item_win_start = m[0].data_ptr(), lens32.data_ptr(), ws32.data_ptr()
    hb.item_start, hb.item_ref, hb.item_alt = m[3].data_ptr(), m[4].data_ptr(), m[5].data_ptr()
    hb.item_p1, hb.item_p2, hb.item_nrec = m[6].data_ptr(), m[7].data_ptr(), m[8].data_ptr()
    hb.lut, hb.hap1, hb.hap2 = lut_t.data_ptr(), hap1.data_ptr(), hap2.data_ptr()
    hb.stream = torch.cuda.current_stream(dev).cuda_stream
    for _ in range(3):
        capi.check(capi.lib().hb_encode_haplotypes(C.byref(hb)))
    torch.cuda.synchronize()
    reps = 5 if B * L > 1e8 else 50
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.check(capi.lib().hb_encode_haplotypes(C.byref(hb)))
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    ms = sorted(t)[len(t) // 2]
    alg = B * L + 2.0 * B * L * 5 * 4
    print(json.dumps({"B": B, "L": L, "C": 5, "kernel_ms": ms, "alg_bytes": alg, "GBs": alg / ms / 1e6,
                      "frac_of_measured_peak": alg / ms / 1e6 / peak, "bases_per_s": B * L / ms * 1e3}))
