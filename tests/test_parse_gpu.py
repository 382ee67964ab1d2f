"""GPU parity tests: the CUDA path (through the C ABI / the pybind11 `parse_vcf` module) against
the CPU oracle on the same inputs.  Bit-exact: all arithmetic on this path is integer/byte."""
import json
import os
import sys
import tempfile

import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi(built):
    from haplohyped_varawareml_b200 import capi as c
    return c


@pytest.fixture(scope="module")
def parse_vcf(built):
    sys.path.insert(0, os.path.join(ROOT, "haplohyped-varawareml_b200"))
    import parse_vcf as m
    return m


def _check_matrix(capi, text, n_samples, region, tokenizer=0, end_is_int=False):
    ora = oracle.parse_text(text, "*", region)
    p = capi.Parse.from_host(synth.body_of(text), n_samples, region=region, tokenizer=tokenizer, end_is_int=end_is_int)
    info = p.info
    assert info.n_records == ora["n"]
    start, stop, ref, alt = p.sites()
    assert np.array_equal(start, ora["start"]) and np.array_equal(stop, ora["stop"])
    assert np.array_equal(ref, ora["ref"]) and np.array_equal(alt, ora["alt"])
    assert p.chrom_column() == ora["chrom"]
    g0, g1 = p.matrix()
    assert np.array_equal(g0, ora["gt0"]), "phase1 plane differs"
    assert np.array_equal(g1, ora["gt1"]), "phase2 plane differs"
    for s in (0, n_samples // 2, n_samples - 1):
        a0, a1 = p.sample(s)
        assert np.array_equal(a0, ora["gt0"][s]) and np.array_equal(a1, ora["gt1"][s])
    pl, bg = p.sample_errors()
    assert pl.sum() == 0 and bg.sum() == 0
    return p, info


def test_fixture_through_pybind_module(parse_vcf, golden_dir):
    """config 1: the reference's own fixture, through the reference's own API names."""
    gold = json.load(open(os.path.join(golden_dir, "fixture_load_vcf.json")))
    vcf = os.path.join(golden_dir, "chr22.filtered.vcf.gz")
    loader = parse_vcf.VCFLoader()
    for s in gold["samples"]:
        exp = [tuple(r) for r in gold["load_vcf"][s]]
        assert loader.load_vcf(vcf, s, "chr22") == exp               # class method (parse_vcf.cpp:120)
        assert parse_vcf.load_vcf(vcf, s, "chr22") == exp            # module-level (vcf_to_h5.py:101)
        assert loader.load_vcf(in_vcf=vcf, sample=s, chrom="chr22") == oracle.load_vcf(vcf, s, "chr22")
    sites = loader.load_vcf_without_sample(vcf, "chr22")
    assert sites == [tuple(r[:5]) for r in gold["load_vcf"][gold["samples"][0]]]
    assert loader.load_vcf(vcf, gold["samples"][0], "chr1") == []
    assert len(loader.load_vcf(vcf, gold["samples"][0])) == 1000    # chrom defaults to ""
    cols = loader.load_vcf_columns(vcf, gold["samples"][1], "chr22")
    exp = gold["load_vcf"][gold["samples"][1]]
    assert list(cols["phase1"]) == [r[5] for r in exp] and list(cols["start"]) == [r[1] for r in exp]


def test_errors_match_reference_contract(parse_vcf, golden_dir):
    vcf = os.path.join(golden_dir, "chr22.filtered.vcf.gz")
    with pytest.raises(RuntimeError, match="Error parsing VCF file: the 1-th sample are not in the VCF"):
        parse_vcf.load_vcf(vcf, "nobody", "chr22")
    with pytest.raises(RuntimeError, match="Error parsing VCF file"):
        parse_vcf.load_vcf("/nonexistent/file.vcf.gz", "x", "chr22")
    S = ["a", "b"]
    bad = (synth.header(S) + "chr22\t100\t.\tA\tC\t.\t.\t.\tGT\t0\t0|1\n").encode()
    with tempfile.NamedTemporaryFile(suffix=".vcf", delete=False) as f:
        f.write(bad)
    try:
        with pytest.raises(RuntimeError, match="ploidy"):
            parse_vcf.load_vcf(f.name, "a", "chr22")                  # reference: assert -> SIGABRT
        assert parse_vcf.load_vcf(f.name, "b", "chr22") == [("chr22", 99, 100, "A", "C", 0, 1)]
    finally:
        os.unlink(f.name)


@pytest.mark.parametrize("tokenizer", [0, 1, 2])
@pytest.mark.parametrize("fmt,kinds,multidigit", [("GT", "phased", False), ("GT", "mixed", False),
                                                   ("GT", "mixed", True), ("GT:GQ:DP", "mixed", True),
                                                   ("DP:GT", "mixed", False)])
def test_random_vcf_matrix_parity(capi, fmt, kinds, multidigit, tokenizer):
    text, samples = synth.random_vcf(700, 301, seed=5, fmt=fmt, kinds=kinds, multidigit=multidigit)
    _check_matrix(capi, text, len(samples), "chr22", tokenizer=tokenizer)
    _check_matrix(capi, text, len(samples), "", tokenizer=tokenizer)


@pytest.mark.parametrize("n_samples", [1, 3, 127, 128, 129, 1000])
def test_sample_count_edges(capi, n_samples):
    text, samples = synth.random_vcf(300, n_samples, seed=n_samples, fmt="GT", kinds="mixed")
    _check_matrix(capi, text, n_samples, "chr22")
    text, samples = synth.random_vcf(150, n_samples, seed=n_samples + 1, fmt="GT:GQ:DP", kinds="mixed")
    _check_matrix(capi, text, n_samples, "chr22")


def test_crlf_info_end_and_regions(capi):
    text, samples = synth.random_vcf(500, 40, seed=9, fmt="GT", kinds="mixed", crlf=True)
    _check_matrix(capi, text, len(samples), "chr22")
    text, samples = synth.random_vcf(500, 40, seed=10, fmt="GT", kinds="mixed", info_end=True)
    _check_matrix(capi, text, len(samples), "chr22", end_is_int=True)
    _check_matrix(capi, text, len(samples), "chr22:10020000-10060000", end_is_int=True)
    _check_matrix(capi, text, len(samples), "chr22_KI270731v1_random", end_is_int=True)
    p, info = _check_matrix(capi, text, len(samples), "chrNope", end_is_int=True)
    assert info.n_records == 0


def test_empty_and_tiny_inputs(capi):
    S = ["a", "b", "c"]
    one = (synth.header(S) + "chr22\t5\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1|1\t./.\n").encode()
    _check_matrix(capi, one, 3, "chr22")
    p = capi.Parse.from_host(b"", 3, region="chr22")
    assert p.info.n_records == 0 and p.info.n_lines == 0
    # no trailing newline on the last record
    _check_matrix(capi, one[:-1], 3, "chr22")


def test_per_sample_errors(capi):
    S = ["a", "b", "c"]
    rows = ["chr22\t5\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1\t0|1", "chr22\t6\t.\tA\tC\t.\t.\t.\tGT\t0|1\t0/1/1\tx|1",
            "chr22\t7\t.\tA\tC\t.\t.\t.\tGT\t1|1\t0|0\t0|1"]
    text = (synth.header(S) + "\n".join(rows) + "\n").encode()
    p = capi.Parse.from_host(synth.body_of(text), 3, region="chr22")
    pl, bg = p.sample_errors()
    assert list(pl) == [0, 2, 0] and list(bg) == [0, 0, 1]
    a0, a1 = p.sample(0)
    assert list(a0) == [0, 0, 1] and list(a1) == [1, 1, 1]
    ora = oracle.parse_text(text, "a", "chr22")
    assert np.array_equal(a0, ora["gt0"]) and np.array_equal(a1, ora["gt1"])
    with pytest.raises(RuntimeError):
        oracle.parse_text(text, "b", "chr22")


def test_synth_device_equals_host_and_parses(capi):
    import torch
    for mix in (0, 1):
        spec = capi.synth_spec(5000, 515, seed=42 + mix, mix=mix)
        host = capi.synth_host(spec)
        n = capi.lib().hb_synth_body_bytes(spec)
        assert n == len(host)
        buf = torch.zeros(n + 256, dtype=torch.uint8, device="cuda:0")
        capi.check(capi.lib().hb_synth_device(spec, buf.data_ptr(), n, 0, None))
        assert bytes(buf[:n].cpu().numpy().tobytes()) == host
        p = capi.Parse.from_device(buf.data_ptr(), n, spec.n_samples, region="chr22")
        ora = oracle.parse_text(capi.synth_header(spec) + host, "*", "chr22")
        g0, g1 = p.matrix()
        assert p.info.n_records == ora["n"] and p.info.n_nonuniform == 0
        assert np.array_equal(g0, ora["gt0"]) and np.array_equal(g1, ora["gt1"])
        p.rerun()                                   # steady-state path used by the bench
        h0, h1 = p.matrix()
        assert np.array_equal(h0, g0) and np.array_equal(h1, g1)


def test_many_tiles_lookback(capi):
    """> 2 x SM-count tokenizer tiles and > 1 site tile: exercises both decoupled look-backs."""
    spec = capi.synth_spec(40000, 257, seed=3)
    text = capi.synth_header(spec) + capi.synth_host(spec)
    assert len(text) > 40 * 1024 * 1024
    _check_matrix(capi, text, spec.n_samples, "chr22", tokenizer=1)
    _check_matrix(capi, text, spec.n_samples, "chr22", tokenizer=2)
