"""CPU tests (no GPU): the oracle against the golden vectors and against an independent pure-python
restatement on the edge cases; host-side pieces that need no device."""
import gzip
import json
import os

import numpy as np
import pytest

import oracle
import synth
from golden.make_golden import expected_tuples


def test_oracle_matches_fixture_golden(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "fixture_load_vcf.json")))
    vcf = os.path.join(golden_dir, "chr22.filtered.vcf.gz")
    for s in gold["samples"]:
        got = oracle.load_vcf(vcf, s, "chr22")
        assert got == [tuple(r) for r in gold["load_vcf"][s]]
    # SURVEY.md 8c known answers for sample 0
    s0 = oracle.load_vcf(vcf, gold["samples"][0], "chr22")
    assert s0[0] == ("chr22", 10012121, 10012122, "C", "G", 1, 1)
    assert s0[1] == ("chr22", 10026998, 10026999, "G", "A", 1, 0)
    assert s0[2] == ("chr22", 10044730, 10044731, "C", "G", 0, 0)
    assert s0[-1] == ("chr22", 19991258, 19991259, "G", "T", 0, 0)
    assert len(s0) == 1000
    sites = oracle.load_vcf_without_sample(vcf, "chr22")
    assert sites == [r[:5] for r in s0]


def test_oracle_unknown_sample_and_region(golden_dir):
    vcf = os.path.join(golden_dir, "chr22.filtered.vcf.gz")
    with pytest.raises(RuntimeError, match="Error parsing VCF file: the 1-th sample are not in the VCF"):
        oracle.load_vcf(vcf, "nobody", "chr22")
    gold = json.load(open(os.path.join(golden_dir, "fixture_load_vcf.json")))
    assert oracle.load_vcf(vcf, gold["samples"][0], "chr1") == []          # other contig: empty result
    assert len(oracle.load_vcf(vcf, gold["samples"][0], "")) == 1000       # no region: everything
    sub = oracle.load_vcf(vcf, gold["samples"][0], "chr22:10012122-10044731")
    assert [r[1] for r in sub] == [10012121, 10026998, 10044730]


@pytest.mark.parametrize("fmt,kinds", [("GT", "mixed"), ("GT:GQ:DP", "mixed"), ("GT", "phased"), ("DP:GT", "mixed")])
def test_oracle_vs_python_restatement(fmt, kinds):
    text, samples = synth.random_vcf(400, 7, seed=3, fmt=fmt, kinds=kinds, multidigit=(fmt != "GT"))
    for s in (samples[0], samples[3], samples[-1]):
        exp = [tuple(r) for r in expected_tuples(text.decode(), s, "chr22")]
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".vcf") as f:
            f.write(text)
            f.flush()
            got = oracle.load_vcf(f.name, s, "chr22")
        assert got == exp
    m = oracle.parse_text(text, "*", "chr22")
    one = oracle.parse_text(text, samples[3], "chr22")
    assert np.array_equal(m["gt0"][3], one["gt0"]) and np.array_equal(m["gt1"][3], one["gt1"])


def test_oracle_edge_semantics():
    S = ["a", "b"]
    h = synth.header(S)
    rows = [
        "chr22\t100\t.\tA\tC\t.\t.\t.\tGT\t0/1\t0|1",          # unphased == phased
        "chr22\t101\t.\tA\tC\t.\t.\t.\tGT\t./.\t.|.",          # missing -> -9
        "chr22\t102\t.\tA\tC\t.\t.\t.\tGT\t.|1\t0/.",
        "chr22\t103\t.\tA\tC\t.\t.\t.\tGT\t200|3\t12/0",       # multi-digit, int8 wrap (200 -> -56)
        "chr22\t104\t.\tA\tC,G\t.\t.\t.\tGT\t1|2\t0|0",        # multiallelic: dropped
        "chr22\t105\t.\tAT\tA\t.\t.\t.\tGT\t1|0\t0|0",         # indel: dropped
        "chr22\t106\t.\tA\t*\t.\t.\t.\tGT\t1|0\t0|0",          # ALT *: dropped
        "chr22\t107\t.\tA\t<DEL>\t.\t.\t.\tGT\t1|0\t0|0",      # symbolic: dropped
        "chr22\t108\t.\tA\tc\t.\t.\t.\tGT\t1|0\t0|0",          # lower-case ALT: dropped
        "chr22\t109\t.\tN\tT\t.\t.\t.\tGT\t1|0\t0|0",          # REF N: kept
        "chr21\t110\t.\tA\tT\t.\t.\t.\tGT\t1|0\t0|0",          # other contig: region-filtered
        "chr22\t111\t.\tA\tT\t.\t.\t.\tGT:GQ:DP\t1|0:9:3\t0|1:50:20",
    ]
    text = (h + "\n".join(rows) + "\n").encode()
    a = oracle.parse_text(text, "a", "chr22")
    assert list(a["start"]) == [99, 100, 101, 102, 108, 110]
    assert list(a["gt0"]) == [0, -9, -9, -56, 1, 1]
    assert list(a["gt1"]) == [1, -9, 1, 3, 0, 0]
    b = oracle.parse_text(text, "b", "chr22")
    assert list(b["gt0"]) == [0, -9, 0, 12, 0, 0] and list(b["gt1"]) == [1, -9, -9, 0, 0, 1]
    assert list(a["ref"]) == [b"A", b"A", b"A", b"A", b"N", b"A"]
    allc = oracle.parse_text(text, "a", "")
    assert allc["n"] == 7 and allc["chrom"][5] == "chr21"
    # haploid GT for the requested sample: the reference aborts (parse_vcf.cpp:46); here an error
    bad = (h + "chr22\t100\t.\tA\tC\t.\t.\t.\tGT\t0\t0|1\n").encode()
    with pytest.raises(RuntimeError, match="ploidy"):
        oracle.parse_text(bad, "a", "chr22")
    assert oracle.parse_text(bad, "b", "chr22")["n"] == 1       # ...but only for that sample
    # contig names longer than 5 chars are truncated by the S5 field of the record struct
    long = (h + "chr22_KI270731v1_random\t5\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1|1\n").encode()
    r = oracle.parse_text(long, "a", "")
    rec = oracle.records_from_columns(r["chrom"], r["start"], r["stop"], r["ref"], r["alt"], r["gt0"], r["gt1"])
    assert rec["chrom"][0] == b"chr22" and rec.dtype.itemsize == 35
    assert rec.tobytes()[:5] == b"chr22" and rec.tobytes()[13:23] == b"A" + b"\0" * 9


def test_oracle_info_end():
    S = ["a"]
    extra = '##INFO=<ID=END,Number=1,Type=Integer,Description="End">\n'
    rows = ["chr22\t100\t.\tA\tC\t.\t.\tEND=150\tGT\t0|1", "chr22\t200\t.\tA\tC\t.\t.\tAF=1;END=90\tGT\t0|1",
            "chr22\t300\t.\tA\tC\t.\t.\tSVEND=999\tGT\t0|1"]
    text = (synth.header(S, extra) + "\n".join(rows) + "\n").encode()
    r = oracle.parse_text(text, "a", "chr22")
    assert list(r["stop"]) == [150, 200, 300]
    r2 = oracle.parse_text((synth.header(S) + "\n".join(rows) + "\n").encode(), "a", "chr22")
    assert list(r2["stop"]) == [100, 200, 300]          # END not declared Integer: plain rlen


def test_shuffle_and_records_roundtrip():
    rng = np.random.default_rng(0)
    for n in (0, 1, 35, 36, 35 * 17 + 4, 37625):
        a = rng.integers(0, 256, n, dtype=np.uint8)
        sh = oracle.shuffle(a, 35)
        assert np.array_equal(oracle.unshuffle(sh, 35), a)
        ne = n // 35
        if ne:
            assert np.array_equal(sh[:ne], a[: ne * 35 : 35])             # plane 0 = byte 0 of every record
            assert np.array_equal(sh[34 * ne:35 * ne], a[34: ne * 35 : 35])


def test_encode_dict_golden(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "encode_dict.json")))["parse_encode_dict"]
    for case in gold:
        inp = case["input"]
        assert oracle.parse_encode_dict(inp) == case["output"]
    assert oracle.parse_encode_dict(None) == {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4}   # tests/test_utils.py:11-15
    with pytest.raises(TypeError):
        oracle.parse_encode_dict(123)                                                       # tests/test_utils.py:29-32


def test_guess_chunk_matches_h5py_table():
    # values h5py's guess_chunk gives for 1-D 35-byte items
    assert [oracle.guess_chunk_1d(n) for n in (1000, 500_000, 1_100_000, 3_000_000)] == [250, 977, 1075, 1465]


def test_dataset_oracle_semantics():
    ref = np.frombuffer(b"ACGTNacgtACGTACGTACGT", "S1")
    idx = oracle.base_to_index(ref, None)
    assert list(idx[:9]) == [0, 1, 2, 3, 4, 0, 1, 2, 3]
    start = np.array([102, 105, 105, 130], np.uint32)
    refs, alts = np.array([b"G", b"A", b"A", b"C"]), np.array([b"T", b"C", b"G", b"T"])
    p1, p2 = np.array([1, 1, 0, 1], np.int8), np.array([0, -9, 1, 1], np.int8)
    h1, h2 = oracle.encode_haplotypes(ref, start, refs, alts, p1, p2, 100, 121)
    assert h1[2] == 3 and h2[2] == 2                      # phase 1 -> ALT, else the VCF REF index
    assert h1[5] == 0 and h2[5] == 2                      # duplicate position: last record wins
    assert np.array_equal(np.delete(h1, [2, 5]), np.delete(idx, [2, 5]))   # 130 is outside the window
    oh = oracle.onehot(h1, 5)
    assert oh.shape == (21, 5) and oh.sum() == 21          # tests/test_utils.py:42-43 intent
    assert oracle.calculate_midpoint_region(10_000_000, 10_001_000, 1000) == (10_000_000, 10_001_000)
    assert oracle.calculate_midpoint_region(100, 300, 1001) == (0, 700)     # clamped at 0, 2*(L//2)


def test_blosc1_chunk_codec_against_stock_liblz4():
    """Filter 32001 stores bare Blosc1 chunks.  The oracle's chunk decoder reads chunks framed the way c-blosc 1.x frames
    them around STOCK LZ4HC streams (one block below 256 KB, several above), and the oracle's own greedy LZ4 encoder
    writes streams stock liblz4 accepts."""
    import ctypes
    L = oracle.stock_lz4()
    if L is None:
        pytest.skip("no system liblz4")
    rng = np.random.default_rng(3)
    rec = np.zeros(1075, dtype=oracle.RECORD_DTYPE)
    rec["chrom"] = b"chr22"
    rec["start"] = np.sort(rng.integers(10_000_000, 50_000_000, 1075)); rec["stop"] = rec["start"] + 1
    rec["ref"] = rng.choice([b"A", b"C", b"G", b"T"], 1075); rec["alt"] = rng.choice([b"A", b"C", b"G", b"T"], 1075)
    rec["phase1"] = rng.random(1075) < 0.1; rec["phase2"] = rng.random(1075) < 0.1
    for data in (rec.tobytes(), np.tile(rec, 8).tobytes(), rec[:3].tobytes(), rng.integers(0, 256, 35 * 400).astype(np.uint8).tobytes()):
        for hc in (True, False):
            c = oracle.reference_like_chunk(data, 35, 5, hc)
            assert c[:4] == bytes([2, 1, 0x31, 35]) and int.from_bytes(c[12:16], "little") == len(c)
            assert int.from_bytes(c[8:12], "little") == oracle.cblosc1_blocksize(len(data), 35, 5, hc)
            assert oracle.blosc_chunk_decode(c, len(data)).tobytes() == data
        own = oracle.blosc1_chunk_encode(data, 35)
        assert oracle.blosc_chunk_decode(own, len(data)).tobytes() == data
        csize = int.from_bytes(own[20:24], "little")
        if csize != len(data):                                       # an LZ4 stream (not stored raw): stock liblz4 reads it
            out = ctypes.create_string_buffer(len(data))
            assert L.LZ4_decompress_safe(own[24:24 + csize], out, csize, len(data)) == len(data)
            assert out.raw == oracle.shuffle(data, 35).tobytes()
    assert oracle.cblosc1_blocksize(37625, 35) == 37625 and oracle.cblosc1_blocksize(35 * 30000, 35) == 262115
    with pytest.raises(ValueError):
        oracle.blosc_chunk_decode(b"\x09" + oracle.reference_like_chunk(rec.tobytes())[1:], rec.nbytes)
