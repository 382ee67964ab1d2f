"""GPU parity tests for kernel 5 / RandomHaplotypeDataset against the numpy oracle (exact: the
one-hot is 0.0/1.0 float32, so equality is bit-exact, tolerance 0)."""
import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built):
    from haplohyped_varawareml_b200 import capi, haplotype_dataset, common_utils
    return capi, haplotype_dataset, common_utils


def _make_world(tmp_path, rng, n_donors=4, chrom_len=60_000, n_rec=700, L=1000):
    donors = ["d%02d" % i for i in range(n_donors)]
    (tmp_path / "samples.txt").write_text("\n".join(donors))
    seqs = {}
    for c in range(1, 23):
        a = rng.choice(np.frombuffer(b"ACGTNacgtnRY", np.uint8), size=chrom_len, p=[.22, .22, .22, .22, .02, .02, .02, .02, .02, .01, .005, .005])
        seqs[f"chr{c}"] = a
    records = {}
    for d in donors:
        for c in range(1, 23):
            start = np.sort(rng.integers(0, chrom_len, n_rec)).astype(np.uint32)   # sorted, with duplicates
            rec = np.zeros(n_rec, oracle.RECORD_DTYPE)
            rec["chrom"] = f"chr{c}".encode()[:5]
            rec["start"], rec["stop"] = start, start + 1
            rec["ref"] = rng.choice(np.array([b"A", b"C", b"G", b"T", b"N"]), n_rec)
            rec["alt"] = rng.choice(np.array([b"A", b"C", b"G", b"T"]), n_rec)
            rec["phase1"] = rng.choice(np.array([0, 1, -9, 2], np.int8), n_rec, p=[.5, .4, .05, .05])
            rec["phase2"] = rng.choice(np.array([0, 1, -9, 2], np.int8), n_rec, p=[.5, .4, .05, .05])
            records[(d, c)] = rec
    bed = tmp_path / "regions.bed"
    rows = []
    for _ in range(30):
        s = int(rng.integers(0, chrom_len - 200))
        rows.append(f"chr22\t{s}\t{s + int(rng.integers(50, 3000))}")
    rows.append("chr22\t10\t200")                       # window clamped at 0
    rows.append(f"chr22\t{chrom_len - 300}\t{chrom_len - 10}")   # window runs past the chromosome end
    bed.write_text("\n".join(rows) + "\n")
    return donors, seqs, records, bed


def _expected(items, seqs, records, L, spec):
    Lout = 2 * (L // 2)
    C = len(spec)
    h1 = np.zeros((len(items), Lout, C), np.float32)
    h2 = np.zeros_like(h1)
    for b, (chrom, donor, ns, ne) in enumerate(items):
        seq = seqs[f"chr{chrom}"]
        win = seq[ns:ne]
        rec = records[(donor, chrom)]
        a, bb = oracle.encode_haplotypes(win, rec["start"], rec["ref"], rec["alt"], rec["phase1"], rec["phase2"], ns, ne, spec)
        n = min(len(win), Lout)
        h1[b, :n] = oracle.onehot(a[:n], C)
        h2[b, :n] = oracle.onehot(bb[:n], C)
    return h1, h2


@pytest.mark.parametrize("L,B,spec", [(1000, 8, None), (1001, 5, None), (4096, 3, "ACGT"), (333, 6, ["T", "G", "C", "A", "N"]),
                                      (131072, 3, None)])             # BASELINE configs[4]: seq_length 1000 and 131072
def test_dataset_matches_oracle(mods, tmp_path, L, B, spec):
    capi, hd, cu = mods
    rng = np.random.default_rng(7)
    donors, seqs, records, bed = (_make_world(tmp_path, rng) if L < 100_000 else
                                  _make_world(tmp_path, rng, n_donors=2, chrom_len=260_000, n_rec=2500, L=L))
    ds = hd.RandomHaplotypeDataset(str(bed), None, None, str(tmp_path / "samples.txt"), encode_spec=spec, seed=42,
                                   batch_size=B, seq_length=L,
                                   genotype_store=hd.GenotypeStore.from_records(records),
                                   reference_genome=hd.ReferenceGenome(sequences=seqs, encode_spec=spec))
    assert len(ds) == 32
    # RNG contract: three draws per item, order region, donor, chromosome, global numpy RNG seeded once
    np.random.seed(42)
    exp_items = []
    for _ in range(B):
        r, d, c = np.random.randint(0, 32), np.random.randint(0, len(donors)), np.random.randint(0, 22)
        exp_items.append((int(np.arange(1, 23)[c]), donors[d], *oracle.calculate_midpoint_region(int(ds.bed_start[r]), int(ds.bed_end[r]), L)))
    np.random.seed(42)
    for it in range(3):
        items = ds.draw()
        if it == 0:
            assert items == exp_items
        hap1, hap2 = ds.encode_items(items)
        assert hap1.is_cuda and hap1.dtype.is_floating_point and tuple(hap1.shape) == (B, 2 * (L // 2), len(ds.encode_spec))
        e1, e2 = _expected(items, seqs, records, L, oracle.parse_encode_dict(spec))
        assert np.array_equal(hap1.cpu().numpy(), e1)
        assert np.array_equal(hap2.cpu().numpy(), e2)
    h1, h2 = ds[0]                                       # the public contract: idx ignored, tuple of tensors
    assert tuple(h1.shape) == tuple(h2.shape) == (B, 2 * (L // 2), len(ds.encode_spec))
    ds.close()


def test_dataset_on_parse_store(mods, tmp_path):
    """Dataset fed straight from a device-resident parse (no copies): parser -> dataset end to end."""
    capi, hd, cu = mods
    text, samples = synth.random_vcf(3000, 9, seed=4, fmt="GT", kinds="mixed", site_mix=False)
    p = capi.Parse.from_host(synth.body_of(text), len(samples), region="chr22")
    ora = oracle.parse_text(text, "*", "chr22")
    store = hd.GenotypeStore()
    for c in range(1, 23):
        store.add_parse(c, p, samples)
    lo, hi = int(ora["start"].min()), int(ora["start"].max())
    rng = np.random.default_rng(1)
    seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=hi + 5000)
    seqs = {f"chr{c}": seq for c in range(1, 23)}
    (tmp_path / "samples.txt").write_text("\n".join(samples))
    bed = tmp_path / "r.bed"
    bed.write_text("".join(f"chr22\t{s}\t{s + 1000}\n" for s in rng.integers(lo, hi, 16)))
    ds = hd.RandomHaplotypeDataset(str(bed), None, None, str(tmp_path / "samples.txt"), batch_size=16, seq_length=2000,
                                   genotype_store=store, reference_genome=hd.ReferenceGenome(sequences=seqs))
    items = ds.draw()
    hap1, hap2 = ds.encode_items(items)
    records = {}
    for s, name in enumerate(samples):
        rec = oracle.records_from_columns(ora["chrom"], ora["start"], ora["stop"], ora["ref"], ora["alt"], ora["gt0"][s], ora["gt1"][s])
        for c in range(1, 23):
            records[(name, c)] = rec
    e1, e2 = _expected(items, seqs, records, 2000, oracle.parse_encode_dict(None))
    assert np.array_equal(hap1.cpu().numpy(), e1) and np.array_equal(hap2.cpu().numpy(), e2)
    assert (hap1 != hap2).any()                           # heterozygous sites do differ between haplotypes


def test_encode_sequence_contract(mods):
    """The reference's own assertions for encode_sequence (tests/test_utils.py:38-70), minus the one that
    contradicts its code (column order): shape (L, 5), one 1 per row, case-insensitive, N for ambiguous."""
    capi, hd, cu = mods
    r = cu.encode_sequence("ACGT")
    assert r.shape == (4, 5) and r.sum() == 4
    assert np.array_equal(r, np.eye(5, dtype=np.uint8)[:4])
    assert np.array_equal(cu.encode_sequence("acgt"), r)
    assert np.array_equal(cu.encode_sequence(np.array([b"A", b"C", b"G", b"T"], dtype="|S1")), r)
    r5 = cu.encode_sequence("ACGTN")
    assert r5.shape == (5, 5) and r5[4, 4] == 1 and r5.sum() == 5
    assert cu.encode_sequence("AXGT")[1, 4] == 1
    with pytest.raises(TypeError):
        cu.encode_sequence([1, 2, 3, 4])


def test_fasta_encoder_to_dataset(mods, tmp_path):
    """fasta_encoder mirror end to end: FASTA -> reference_genome.h5 -> RandomHaplotypeDataset windows; and its
    encode_sequence (GPU) against the oracle one-hot."""
    capi, hd, cu = mods
    from haplohyped_varawareml_b200 import fasta_encoder as fe
    rng = np.random.default_rng(11)
    donors, seqs, records, bed = _make_world(tmp_path, rng, n_donors=2, chrom_len=20_000, n_rec=200)
    fa = tmp_path / "ref.fa"
    with open(fa, "wb") as f:
        for k, v in seqs.items():
            b = v.tobytes()
            f.write(b">" + k.encode() + b"\n" + b"\n".join(b[i:i + 70] for i in range(0, len(b), 70)) + b"\n")
    fe.main(["--fasta", str(fa), "--outdir", str(tmp_path / "refout")])
    ds = hd.RandomHaplotypeDataset(str(bed), None, str(tmp_path / "refout" / "reference_genome.h5"), str(tmp_path / "samples.txt"),
                                   batch_size=4, seq_length=600, genotype_store=hd.GenotypeStore.from_records(records))
    items = ds.draw()
    h1, h2 = ds.encode_items(items)
    e1, e2 = _expected(items, seqs, records, 600, oracle.parse_encode_dict(None))
    assert np.array_equal(h1.cpu().numpy(), e1) and np.array_equal(h2.cpu().numpy(), e2)
    ds.close()
    rg = fe.ReferenceGenome()
    s = "ACGTNacgtnRYxx"
    spec = oracle.parse_encode_dict(None)
    idx = np.array([spec.get(c.upper(), spec["N"]) for c in s], np.int8)
    assert np.array_equal(rg.encode_sequence(s), oracle.onehot(idx, 5))
    assert np.array_equal(rg.encode_sequence(np.frombuffer(s.encode(), "|S1")), oracle.onehot(idx, 5))
    with pytest.raises(TypeError):
        rg.encode_sequence([1, 2])


def test_dataset_reads_windows_from_compressed_resident_file(mods, tmp_path):
    """Row f4: the dataset on a cohort .h5 -- stored chunks stay compressed in HBM, a batch decodes ON THE DEVICE only the
    chunks its windows touch (chunk choice by binary search on the per-chunk first positions), no record crosses PCIe.
    Same tensors as the oracle built from the full record arrays."""
    capi, hd, cu = mods
    from haplohyped_varawareml_b200 import container, h5_reader
    text, samples = synth.random_vcf(5000, 6, seed=14, fmt="GT", kinds="mixed", site_mix=False)
    ora = oracle.parse_text(text, "*", "chr22")
    p = capi.Parse.from_host(synth.body_of(text), len(samples), region="chr22")
    cr = 64
    fr = p.compress(cr)
    n = int(p.info.n_records)
    path = str(tmp_path / "cohort.h5")
    records = {}
    with container.open_h5(path, "w", backend="minih5") as f:
        for s, name in enumerate(samples):
            frames = fr.sample(s)
            rec = oracle.records_from_columns(ora["chrom"], ora["start"], ora["stop"], ora["ref"], ora["alt"], ora["gt0"][s], ora["gt1"][s])
            for c in range(1, 23):
                f.write_chunked(f"donor_{name}/chr_{c}/snp_data", oracle.RECORD_DTYPE, n, cr, frames)
                records[(name, c)] = rec
    lo, hi = int(ora["start"].min()), int(ora["start"].max())
    rng = np.random.default_rng(2)
    seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=hi + 30000)
    seqs = {f"chr{c}": seq for c in range(1, 23)}
    (tmp_path / "samples.txt").write_text("\n".join(samples))
    bed = tmp_path / "r.bed"
    bed.write_text("".join(f"chr22\t{s}\t{s + 1000}\n" for s in list(rng.integers(lo, hi, 14)) + [lo - 400, hi - 100]))
    for L, B in ((1000, 32), (20000, 8)):
        ds = hd.RandomHaplotypeDataset(str(bed), path, None, str(tmp_path / "samples.txt"), batch_size=B, seq_length=L,
                                       reference_genome=hd.ReferenceGenome(sequences=seqs))
        store = ds.genotypes
        for it in range(2):
            items = ds.draw()
            hap1, hap2 = ds.encode_items(items)
            e1, e2 = _expected(items, seqs, records, L, oracle.parse_encode_dict(None))
            assert np.array_equal(hap1.cpu().numpy(), e1) and np.array_equal(hap2.cpu().numpy(), e2)
        store.check_last()
        assert store._stored and all(v is not None for v in store._stored.values())     # kept compressed, never read whole
        assert not store._cols
        # window-limited: an item's columns cover a few chunks, not the chromosome
        cols, keep = store.window_columns([(samples[0], 22, lo + 5000, lo + 5000 + L)])
        span = (ora["start"] >= lo + 5000) & (ora["start"] < lo + 5000 + L)
        assert cols[0][5] <= (int(span.sum()) // cr + 3) * cr and cols[0][5] < n
        ds.close()
    # a damaged stored chunk is reported, not decoded into garbage
    st = hd.GenotypeStore.from_reader(h5_reader.VCFH5Reader(path))
    sd = st._compressed(samples[1], 7)
    st._chunk_first(7, sd)
    sd.blob[int(sd.offs[3]) + 30:int(sd.offs[3]) + 60] = 0xFF
    first = st._first[(7, sd.n, sd.cr)]
    st.window_columns([(samples[1], 7, int(first[3]), int(first[3]) + 10)])
    with pytest.raises(OSError):
        st.check_last()
    st.close()
