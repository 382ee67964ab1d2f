"""GPU parity tests for the write side end to end (vcf_to_h5 mirror -> HDF5 container) and the read
side (VCFH5Reader -> GPU Blosc2 decode -> RandomHaplotypeDataset).

Bar (north_star): record arrays read back from the file are bit-exact with what the reference would
have stored -- np.array([tuple(row) ...], dtype=35-byte struct) of load_vcf's tuples
(vcf_to_h5.py:101-129), restated by the oracle."""
import gzip
import os

import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built):
    from haplohyped_varawareml_b200 import capi, container, h5_reader, haplotype_dataset, vcf_to_h5
    return capi, container, h5_reader, haplotype_dataset, vcf_to_h5


def _expected_records(path, donor, chrom):
    rows = oracle.load_vcf(path, donor, chrom)
    return oracle.records_from_tuples(rows)


def test_decode_frames_matches_oracle(mods):
    capi = mods[0]
    text, samples = synth.random_vcf(2600, 7, seed=5, fmt="GT", kinds="mixed")
    ora = oracle.parse_text(text, "*", "chr22")
    p = capi.Parse.from_host(synth.body_of(text), len(samples), region="chr22")
    fr = p.compress(0)
    cr = int(fr.info.chunk_records)
    for s in (0, 3, 6):
        frames = fr.sample(s)
        rec = oracle.records_from_columns(ora["chrom"], ora["start"], ora["stop"], ora["ref"], ora["alt"],
                                          ora["gt0"][s], ora["gt1"][s])
        raw = rec.tobytes() + b"\0" * (len(frames) * cr * 35 - rec.nbytes)
        got = capi.decode_frames(frames, cr * 35)
        assert got.tobytes() == raw
        planar = capi.decode_frames(frames, cr * 35, planar=True)
        for k in range(len(frames)):
            assert planar[k].tobytes() == oracle.shuffle(raw[k * cr * 35:(k + 1) * cr * 35], 35).tobytes()
            assert oracle.blosc_chunk_decode(frames[k], cr * 35).tobytes() == got[k].tobytes()
    # corrupt frames are rejected, not decoded into garbage
    good = fr.sample(0)[0]
    import struct
    def mutated(at, value):
        b = bytearray(good); b[at:at + len(value)] = value; return bytes(b)
    hostile = [
        mutated(0, b"\x09"),                                  # format version from the future
        mutated(2, bytes([good[2] | 0x08])),                  # reserved flag bit
        mutated(2, bytes([(good[2] & 0x1f) | (2 << 5)])),     # another codec
        mutated(4, struct.pack("<I", cr * 35 + 35)),          # nbytes != the dataset's chunk size
        mutated(8, struct.pack("<I", 0)),                     # blocksize 0
        mutated(8, struct.pack("<I", 1)),                     # blocksize 1: 37 K bstarts would lie outside the chunk
        mutated(12, struct.pack("<I", len(good) + 1000)),     # cbytes beyond the stored bytes
        mutated(16, struct.pack("<I", 0xfffffff0)),           # bstarts[0] far outside
        mutated(16, struct.pack("<I", 4)),                    # bstarts[0] inside the header
        mutated(20, struct.pack("<i", -5)),                   # negative stream size
        mutated(20, struct.pack("<I", len(good))),            # stream runs past the end of the chunk
        good[:len(good) // 2],                                # truncated
        mutated(24, b"\xff" * 8),                             # LZ4: literal run longer than the block
        mutated(len(good) - 40, b"\x00" * 8),                 # LZ4: damaged tail (offset 0 / wrong length)
        good[:12],                                            # shorter than a header
    ]
    for k, bad in enumerate(hostile):
        with pytest.raises(capi.HaploError):
            capi.decode_frames([bad], cr * 35)
    assert capi.decode_frames([good], cr * 35).tobytes() == oracle.blosc_chunk_decode(good, cr * 35).tobytes()   # the device is still healthy
    assert capi.decode_frames([], cr * 35).shape == (0, cr * 35)


def test_fixture_through_cli_and_reader(mods, tmp_path, golden_dir):
    """configs[0]: the reference's own fixture through the converter (chr22 only; chr1-21 absent -> skipped)."""
    capi, container, h5_reader, hd, v2h = mods
    vdir = tmp_path / "vcf"
    vdir.mkdir()
    src = os.path.join(golden_dir, "chr22.filtered.vcf.gz")
    os.symlink(src, vdir / "chr22.filtered.vcf.gz")
    out = tmp_path / "out"
    conv = v2h.VCFtoHDF5Converter("cohort", str(vdir), str(out), os.path.join(golden_dir, "ipscs_samples_test.txt"), 4, 4)
    assert len(conv.donor_ids) == 3 and os.path.isdir(out / "tmp_files")
    conv.run()
    assert not os.path.exists(out / "tmp_files") and os.path.exists(out / "cohort.h5")
    assert conv.stats["datasets"] == 3 and conv.stats["skipped_files"] == 21
    rd = h5_reader.VCFH5Reader(str(out / "cohort.h5"))
    for d in conv.donor_ids:
        got = rd.fetch_genotypes(d, 22)
        exp = _expected_records(src, d, "chr22")
        assert got.dtype == oracle.RECORD_DTYPE and len(got) == 1000
        assert got.tobytes() == exp.tobytes()
    with pytest.raises(KeyError):
        rd.fetch_genotypes(conv.donor_ids[0], 1)
    with pytest.raises(KeyError):
        rd.fetch_genotypes("nobody", 22)
    rd.close()


def test_multi_chromosome_cohort(mods, tmp_path):
    capi, container, h5_reader, hd, v2h = mods
    vdir = tmp_path / "vcf"
    vdir.mkdir()
    samples = None
    files = {}
    for c, (nv, fmt, kinds) in {1: (1500, "GT", "phased"), 7: (2300, "GT:GQ:DP", "mixed"), 22: (900, "GT", "mixed")}.items():
        text, samples = synth.random_vcf(nv, 11, seed=c, fmt=fmt, kinds=kinds, chrom=f"chr{c}", site_mix=True)
        path = vdir / f"chr{c}.filtered.vcf.gz"
        with gzip.open(path, "wb") as f:
            f.write(text)
        files[c] = str(path)
    (vdir / "chr5.filtered.vcf.gz").write_bytes(gzip.compress(synth.header(samples).encode()))    # no records at all
    donors = [samples[0], samples[4], "not-in-the-vcf", samples[10]]
    (tmp_path / "donors.txt").write_text("\n".join(donors))
    conv = v2h.VCFtoHDF5Converter("c2", str(vdir), str(tmp_path / "o"), str(tmp_path / "donors.txt"), 2, 4)
    conv.run()
    assert conv.stats["skipped_donors"] == 4 and conv.stats["datasets"] == 3 * 4
    f = container.open_h5(str(tmp_path / "o" / "c2.h5"))
    assert f.keys() == sorted(f"donor_{d}" for d in donors if d != "not-in-the-vcf")
    assert f.keys(f"donor_{samples[0]}") == sorted(["chr_1", "chr_5", "chr_7", "chr_22"])
    f.close()
    rd = h5_reader.VCFH5Reader(str(tmp_path / "o" / "c2.h5"))
    for d in (samples[0], samples[4], samples[10]):
        for c, path in files.items():
            got = rd.fetch_genotypes(d, c)
            exp = _expected_records(path, d, f"chr{c}")
            assert len(exp) > 100 and got.tobytes() == exp.tobytes()
        assert len(rd.fetch_genotypes(d, 5)) == 0
    rd.close()


def test_dataset_from_files(mods, tmp_path):
    """RandomHaplotypeDataset constructed exactly as the reference constructs it: four paths."""
    capi, container, h5_reader, hd, v2h = mods
    rng = np.random.default_rng(3)
    vdir = tmp_path / "vcf"
    vdir.mkdir()
    texts = {}
    for c in range(1, 23):
        text, samples = synth.random_vcf(400, 5, seed=100 + c, fmt="GT", kinds="mixed", chrom=f"chr{c}", site_mix=False)
        texts[c] = text
        with gzip.open(vdir / f"chr{c}.filtered.vcf.gz", "wb") as f:
            f.write(text)
    (tmp_path / "samples.txt").write_text("\n".join(samples))
    v2h.VCFtoHDF5Converter("g", str(vdir), str(tmp_path), str(tmp_path / "samples.txt"), 2, 4).run()
    # reference chromosomes long enough to cover the variant positions (10.0 Mb + ...)
    hi = max(int(oracle.parse_text(texts[c], "*", f"chr{c}")["start"].max()) for c in texts) + 3000
    lo = 10_000_000
    seqs = {f"chr{c}": rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), size=hi) for c in range(1, 23)}
    with container.open_h5(str(tmp_path / "ref.h5"), "w") as f:
        for k, v in seqs.items():
            f.write_array(k, v.view("S1"))
    bed = tmp_path / "r.bed"
    bed.write_text("".join(f"chr1\t{s}\t{s + 500}\n" for s in rng.integers(lo, hi - 3000, 12)))
    ds = hd.RandomHaplotypeDataset(str(bed), str(tmp_path / "g.h5"), str(tmp_path / "ref.h5"), str(tmp_path / "samples.txt"),
                                   batch_size=6, seq_length=1500)
    items = ds.draw()
    h1, h2 = ds.encode_items(items)
    spec = oracle.parse_encode_dict(None)
    for b, (chrom, donor, ns, ne) in enumerate(items):
        ora = oracle.parse_text(texts[chrom], donor, f"chr{chrom}")
        a, bb = oracle.encode_haplotypes(seqs[f"chr{chrom}"][ns:ne], ora["start"], ora["ref"], ora["alt"],
                                         ora["gt0"], ora["gt1"], ns, ne, spec)
        assert np.array_equal(h1[b].cpu().numpy(), oracle.onehot(a, 5))
        assert np.array_equal(h2[b].cpu().numpy(), oracle.onehot(bb, 5))
    ds.close()


def test_sample_windows_give_the_same_file_contents(mods, tmp_path):
    """Big chromosomes are converted a window of samples at a time (frames of all donors would not fit in HBM next to
    the genotype planes): any window size must store the same records; frames API: window-relative sample indices."""
    capi, container, h5_reader, hd, v2h = mods
    vdir = tmp_path / "vcf"
    vdir.mkdir()
    text, samples = synth.random_vcf(2100, 23, seed=8, fmt="GT", kinds="mixed", chrom="chr3", site_mix=True)
    with open(vdir / "chr3.filtered.vcf.gz", "wb") as f:
        f.write(synth.bgzf_compress(text, 6))
    (tmp_path / "donors.txt").write_text("\n".join(samples))
    got = {}
    for w in (None, 5, 1):
        out = tmp_path / f"o{w}"
        conv = v2h.VCFtoHDF5Converter("c", str(vdir), str(out), str(tmp_path / "donors.txt"), 2, 4, chromosomes=[3], sample_window=w)
        conv.run()
        assert conv.stats["datasets"] == 23
        rd = h5_reader.VCFH5Reader(str(out / "c.h5"))
        got[w] = {d: rd.fetch_genotypes(d, 3).tobytes() for d in samples}
        rd.close()
    ora = oracle.parse_text(text, "*", "chr3")
    for k, d in enumerate(samples):
        rec = oracle.records_from_columns(ora["chrom"], ora["start"], ora["stop"], ora["ref"], ora["alt"], ora["gt0"][k], ora["gt1"][k])
        assert got[None][d] == rec.tobytes() and got[5][d] == got[None][d] and got[1][d] == got[None][d]
    # the window API directly: frames of samples [7, 12) equal the same samples of an all-sample pass
    p = capi.Parse.from_host(synth.body_of(text), 23, region="chr3")
    allf = p.compress(0)
    win = p.compress(0, 7, 5)
    assert win.info.n_samples == 5
    for k in range(5):
        assert win.sample(k) == allf.sample(7 + k)
    win.set_window(18, 5)
    win.rerun(p)
    assert win.sample(4) == allf.sample(22)
    with pytest.raises(capi.HaploError):
        win.set_window(0, 6)
    p.release_text()
    with pytest.raises(capi.HaploError):
        p.rerun()
    win.set_window(2, 3)
    win.rerun(p)                                             # frames only need the planes, not the text
    assert win.sample(0) == allf.sample(2)


def test_gpu_decoder_reads_what_the_reference_writer_stores(mods):
    """Read side against the reference's WRITE side: chunks framed as c-blosc 1.x frames them (hdf5-blosc, clevel 5, shuffle,
    LZ4HC) around stock liblz4 LZ4HC streams -- one block, several blocks, a leftover block, incompressible data stored raw,
    tiny chunks -- decode on the GPU to the records."""
    capi = mods[0]
    if oracle.stock_lz4() is None:
        pytest.skip("no system liblz4")
    rng = np.random.default_rng(11)
    def records(n):
        rec = np.zeros(n, dtype=oracle.RECORD_DTYPE)
        rec["chrom"] = b"chr1"
        rec["start"] = np.sort(rng.integers(1, 200_000_000, n)); rec["stop"] = rec["start"] + 1
        rec["ref"] = rng.choice([b"A", b"C", b"G", b"T"], n); rec["alt"] = rng.choice([b"A", b"C", b"G", b"T"], n)
        rec["phase1"] = rng.random(n) < 0.2; rec["phase2"] = rng.random(n) < 0.2
        return rec
    for n in (1075, 8 * 1075 + 13, 5, 30000):
        datas = [records(n).tobytes() for _ in range(3)] + [rng.integers(0, 256, 35 * n).astype(np.uint8).tobytes()]
        chunks = [oracle.reference_like_chunk(d) for d in datas]
        got = capi.decode_frames(chunks, 35 * n)
        for k, d in enumerate(datas):
            assert got[k].tobytes() == d
        if oracle.cblosc1_blocksize(35 * n, 35) == 35 * n:              # single block: the planar view exists too
            planar = capi.decode_frames(chunks, 35 * n, planar=True)
            assert planar[0].tobytes() == oracle.shuffle(datas[0], 35).tobytes()


def test_full_shape_parity_60000_x_2504(mods):
    """The bench shape at a size the oracle finishes in seconds: 60,000 variants x 2,504 samples, chunk_records 1075 --
    the WHOLE genotype matrix and site columns, and every stored chunk of 8 donors, against the oracle."""
    capi = mods[0]
    spec = capi.synth_spec(60000, 2504, seed=77, mix=1 << 8)
    text = capi.synth_header(spec) + capi.synth_host(spec)
    ora = oracle.parse_text(text, "*", "chr22")
    p = capi.Parse.from_host(synth.body_of(text), 2504, region="chr22")
    assert p.info.n_records == ora["n"] == 60000
    g0, g1 = p.matrix()
    assert np.array_equal(g0, ora["gt0"]) and np.array_equal(g1, ora["gt1"])
    start, stop, ref, alt = p.sites()
    assert np.array_equal(start, ora["start"]) and np.array_equal(stop, ora["stop"])
    assert np.array_equal(ref, ora["ref"]) and np.array_equal(alt, ora["alt"])
    fr = p.compress(1075)
    assert fr.info.n_chunks == 56
    pk, offs, sizes = fr.fetch_packed()
    for s in (0, 1, 313, 1251, 1252, 2000, 2502, 2503):
        rec = oracle.records_from_columns(ora["chrom"], ora["start"], ora["stop"], ora["ref"], ora["alt"], ora["gt0"][s], ora["gt1"][s])
        raw = rec.tobytes() + b"\0" * (56 * 1075 * 35 - rec.nbytes)
        for c in range(56):
            f = pk[int(offs[s, c]):int(offs[s, c]) + int(sizes[s, c])].tobytes()
            assert oracle.blosc_chunk_decode(f, 1075 * 35).tobytes() == raw[c * 1075 * 35:(c + 1) * 1075 * 35], (s, c)
    # and the GPU read side gives the same records back from its own frames
    s = 1251
    frames = [pk[int(offs[s, c]):int(offs[s, c]) + int(sizes[s, c])].tobytes() for c in range(56)]
    rec = oracle.records_from_columns(ora["chrom"], ora["start"], ora["stop"], ora["ref"], ora["alt"], ora["gt0"][s], ora["gt1"][s])
    assert capi.decode_frames(frames, 1075 * 35).reshape(-1)[:rec.nbytes].tobytes() == rec.tobytes()


def test_file_larger_than_the_text_budget_converts_to_the_same_file(mods, tmp_path):
    """A .vcf.gz whose text does not fit HBM next to its planes takes the streamed-resident route inside hb_parse_file /
    hb_load_vcf (hb_parse_set_text_limit forces it here): the .h5 the converter writes is byte-identical to the whole-file
    one -- chunk geometry of the whole dataset (vcf_to_h5.py:135), no chunk boundary at a slab boundary -- and load_vcf
    returns the same tuples."""
    capi, container, h5_reader, hd, v2h = mods
    import sys
    sys.path.insert(0, os.path.dirname(capi.__file__))
    import parse_vcf
    vdir = tmp_path / "vcf"
    vdir.mkdir()
    text, samples = synth.random_vcf(5200, 19, seed=77, fmt="GT", kinds="mixed", chrom="chr9", site_mix=True)
    path = vdir / "chr9.filtered.vcf.gz"
    path.write_bytes(synth.bgzf_compress(text, 6))
    (tmp_path / "donors.txt").write_text("\n".join(samples))
    blobs, tuples = {}, {}
    try:
        for limit in (0, 64 << 10):
            capi.lib().hb_parse_set_text_limit(limit)
            capi.lib().hb_cache_clear()
            out = tmp_path / f"o{limit}"
            conv = v2h.VCFtoHDF5Converter("c", str(vdir), str(out), str(tmp_path / "donors.txt"), 2, 4, chromosomes=[9])
            conv.run()
            assert conv.stats["datasets"] == 19
            blobs[limit] = (out / "c.h5").read_bytes()
            tuples[limit] = parse_vcf.load_vcf(str(path), samples[3], "chr9")
            h = capi.Parse.from_file(str(path), region="chr9")
            if limit:
                assert h.info.text_bytes == 0                   # the streamed route was taken: no text is resident
                with pytest.raises(capi.HaploError):
                    h.rerun()
            else:
                h.rerun()
            h.close()
    finally:
        capi.lib().hb_parse_set_text_limit(0)
        capi.lib().hb_cache_clear()
    assert blobs[0] == blobs[64 << 10]
    assert tuples[0] == tuples[64 << 10] == oracle.load_vcf(str(path), samples[3], "chr9")


def test_two_gpus_write_the_same_datasets(mods, tmp_path):
    """vcf_to_h5 --devices 0,1: the chromosome files are converted on two GPUs at once (one host thread per GPU, one
    output file); every dataset holds the same stored chunks as the single-GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    capi, container, h5_reader, hd, v2h = mods
    vdir = tmp_path / "vcf"
    vdir.mkdir()
    samples = None
    for c, nv in {1: 2600, 4: 1900, 9: 1300, 17: 700, 22: 350}.items():
        text, samples = synth.random_vcf(nv, 13, seed=40 + c, fmt="GT", kinds="mixed", chrom=f"chr{c}", site_mix=True)
        (vdir / f"chr{c}.filtered.vcf.gz").write_bytes(synth.bgzf_compress(text, 6))
    (tmp_path / "donors.txt").write_text("\n".join(samples))
    got = {}
    for name, devs in (("one", [0]), ("two", [0, 1])):
        conv = v2h.VCFtoHDF5Converter("c", str(vdir), str(tmp_path / name), str(tmp_path / "donors.txt"), 2, 4, devices=devs)
        conv.run()
        assert conv.stats["datasets"] == 5 * 13 and conv.stats["skipped_files"] == 17
        rd = h5_reader.VCFH5Reader(str(tmp_path / name / "c.h5"))
        got[name] = {(d, c): rd.fetch_genotypes(d, c).tobytes() for d in samples for c in (1, 4, 9, 17, 22)}
        rd.close()
    assert got["one"] == got["two"]
    plan = v2h.VCFtoHDF5Converter("c", str(vdir), str(tmp_path / "p"), str(tmp_path / "donors.txt"), 2, 4, devices=[0, 1]).plan_devices([1, 4, 9, 17, 22])
    assert all(plan) and sorted(c for b in plan for c in b) == [1, 4, 9, 17, 22]


def test_general_text_full_width_parity(mods):
    """The general path at the cohort's width: FORMAT=GT:GQ:DP text of 2,504 samples (variable-width fields, unphased and
    missing calls, dropped sites), 4,000 records -- tokenizer with column checkpoints + tab-scanning decoder: the WHOLE
    matrix, the site columns and every stored chunk of 4 donors against the oracle."""
    capi = mods[0]
    block, samples = synth.random_vcf(160, 2504, seed=91, fmt="GT:GQ:DP", kinds="mixed", site_mix=True)
    body = synth.body_of(block)
    text = block[:len(block) - len(body)] + body * 25
    ora = oracle.parse_text(text, "*", "chr22")
    p = capi.Parse.from_host(synth.body_of(text), 2504, region="chr22")
    assert p.info.n_records == ora["n"] and ora["n"] > 3000 and p.info.tokenizer_used == 2
    g0, g1 = p.matrix()
    assert np.array_equal(g0, ora["gt0"]) and np.array_equal(g1, ora["gt1"])
    start, stop, ref, alt = p.sites()
    assert np.array_equal(start, ora["start"]) and np.array_equal(stop, ora["stop"])
    assert np.array_equal(ref, ora["ref"]) and np.array_equal(alt, ora["alt"])
    pe, be = p.sample_errors()
    assert not pe.any() and not be.any()
    fr = p.compress(0)
    cr, nc = int(fr.info.chunk_records), int(fr.info.n_chunks)
    pk, offs, sizes = fr.fetch_packed()
    for s in (0, 127, 128, 2503):                              # both sides of a decode-tile seam, first and last donor
        rec = oracle.records_from_columns(ora["chrom"], ora["start"], ora["stop"], ora["ref"], ora["alt"], ora["gt0"][s], ora["gt1"][s])
        raw = rec.tobytes() + b"\0" * (nc * cr * 35 - rec.nbytes)
        for c in range(nc):
            f = pk[int(offs[s, c]):int(offs[s, c]) + int(sizes[s, c])].tobytes()
            assert oracle.blosc_chunk_decode(f, cr * 35).tobytes() == raw[c * cr * 35:(c + 1) * cr * 35], (s, c)
