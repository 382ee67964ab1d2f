"""vcf_to_h5 end to end at scale: BGZF file on disk -> GPU inflate -> GPU parse -> kernel 4 -> one bulk write into the
HDF5 container, then a read-back check of a few donors.  python tools/convert_bench.py [variants] [samples] [outdir]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from haplohyped_varawareml_b200 import capi, h5_reader, vcf_to_h5

V = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2504
root = sys.argv[3] if len(sys.argv) > 3 else tempfile.mkdtemp()
window = int(sys.argv[4]) if len(sys.argv) > 4 else None          # samples per frames pass (None: as many as HBM allows)
vdir = os.path.join(root, "vcf"); os.makedirs(vdir, exist_ok=True)
spec = capi.synth_spec(V, S, seed=42)
text = np.frombuffer(capi.synth_header(spec) + capi.synth_host(spec), np.uint8)
bg = capi.bgzf_compress_host(text, 6)
bg.tofile(os.path.join(vdir, "chr22.filtered.vcf.gz"))
names = capi.synth_sample_names(spec)
open(os.path.join(root, "samples.txt"), "w").write("\n".join(names))
capi.Parse.from_host(b"chr22\t5\t.\tA\tC\t.\t.\t.\tGT\t0|1\n", 1).close()      # CUDA context + module load outside the timing
t0 = time.time()
conv = vcf_to_h5.VCFtoHDF5Converter("cohort", vdir, os.path.join(root, "out"), os.path.join(root, "samples.txt"), os.cpu_count(), 4,
                                    chromosomes=[22], sample_window=window)
conv.run()
t1 = time.time()
out = os.path.join(root, "out", "cohort.h5")
res = {"variants": V, "samples": S, "sample_window": window, "text_bytes": int(text.size), "bgzf_bytes": int(bg.size), "h5_bytes": os.path.getsize(out),
       "convert_s": t1 - t0, "records_per_s": conv.stats["records"] / (t1 - t0), "variants_per_s": V / (t1 - t0),
       "datasets": conv.stats["datasets"], "stored_bytes": conv.stats["stored_bytes"],
       "ratio": 35.0 * conv.stats["records"] / max(1, conv.stats["stored_bytes"]),
       "reference_published": {"parse_variants_per_s": 559390, "hdf5_write_records_per_s": 256047, "ratio": 6.5}}
# read-back parity on a few donors (oracle on the first 3000 variants)
nchk = min(V, 3000)
head = capi.synth_header(spec) + capi.synth_host(spec, 0, nchk)
rd = h5_reader.VCFH5Reader(out)
t2 = time.time()
ok = True
for s in (0, S // 2, S - 1):
    got = rd.fetch_genotypes(names[s], 22)
    exp = oracle.records_from_tuples(oracle.load_vcf_text(head, names[s], "chr22")) if hasattr(oracle, "load_vcf_text") else None
    if exp is None:
        o = oracle.parse_text(head, names[s], "chr22")
        exp = oracle.records_from_columns(o["chrom"], o["start"], o["stop"], o["ref"], o["alt"], o["gt0"], o["gt1"])
    ok = ok and len(got) == conv.stats["records"] // S and got[:len(exp)].tobytes() == exp.tobytes()
res["read_back_ok"] = bool(ok)
res["read_3_donors_s"] = time.time() - t2
rd.close()
print(json.dumps(res))
