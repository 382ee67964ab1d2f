import gzip, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import synth
from haplohyped_varawareml_b200 import capi
text = gzip.open("tests/golden/chr22.filtered.vcf.gz").read()
p = capi.Parse.from_host(synth.body_of(text), 3, region="chr22")
fr = p.compress(0)
frames = fr.sample(0)
os.makedirs("gpurun_out", exist_ok=True)
for k, f in enumerate(frames):
    open(f"gpurun_out/frame{k}.bin", "wb").write(f)
print(fr.info.n_chunks, fr.info.chunk_records, [len(f) for f in frames])
