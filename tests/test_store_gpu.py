"""GPU parity tests for kernel 4: byte-shuffle + LZ4 + Blosc chunk framing (HDF5 filter 32001 = hdf5-blosc, c-blosc 1.x).

north_star bar: decompressed chunks bit-exact.  Each frame produced on the GPU is decoded by the
oracle's independent chunk -> LZ4 -> unshuffle path and compared with the 35-byte record
array the reference would have handed to h5py (vcf_to_h5.py:119-129); every header field a stock
blosc_decompress reads is pinned, and the LZ4 payload is also fed to the stock liblz4 (ctypes)."""
import ctypes
import ctypes.util
import struct

import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi(built):
    from haplohyped_varawareml_b200 import capi as c
    return c


def _stock_lz4():
    for name in ("liblz4.so.1", ctypes.util.find_library("lz4")):
        if not name:
            continue
        try:
            L = ctypes.CDLL(name)
            L.LZ4_decompress_safe.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
            return L
        except OSError:
            continue
    return None


def _expected_chunks(ora, s, cr):
    rec = oracle.records_from_columns(ora["chrom"], ora["start"], ora["stop"], ora["ref"], ora["alt"],
                                      ora["gt0"][s], ora["gt1"][s])
    raw = rec.tobytes()
    n_chunks = (ora["n"] + cr - 1) // cr
    raw += b"\0" * (n_chunks * cr * 35 - len(raw))          # HDF5 edge chunks are zero-filled to full size
    return [raw[k * cr * 35:(k + 1) * cr * 35] for k in range(n_chunks)]


def _check(capi, text, n_samples, region, chunk_records, samples_to_check):
    ora = oracle.parse_text(text, "*", region)
    p = capi.Parse.from_host(synth.body_of(text), n_samples, region=region)
    fr = p.compress(chunk_records)
    info = fr.info
    cr = int(info.chunk_records)
    if chunk_records == 0:
        assert cr == oracle.guess_chunk_1d(ora["n"])
    assert info.n_chunks == (ora["n"] + cr - 1) // cr
    lz4 = _stock_lz4()
    total = 0
    for s in samples_to_check:
        frames = fr.sample(s)
        exp = _expected_chunks(ora, s, cr)
        assert len(frames) == len(exp)
        for f, e in zip(frames, exp):
            got = oracle.blosc_chunk_decode(f, len(e)).tobytes()
            assert got == e, "decompressed chunk differs"
            total += len(f)
            # the 16 header bytes + bstarts[0] + csize a stock blosc_decompress (c-blosc 1.x blosc.c) reads: format
            # version 2, LZ4 format version 1, flags = byte-shuffle | don't-split | LZ4 << 5, typesize, nbytes,
            # blocksize (one block), cbytes = the whole stored chunk
            assert f[:4] == bytes([2, 1, 0x31, 35])
            assert struct.unpack("<IIII", f[4:20]) == (len(e), len(e), len(f), 20)
            csize = struct.unpack("<i", f[20:24])[0]
            assert csize == len(f) - 24
            if lz4 is not None:
                payload = f[24:24 + csize]
                if csize == len(e):                      # Blosc convention: csize == size means stored raw
                    assert payload == oracle.shuffle(e, 35).tobytes()
                    continue
                out = ctypes.create_string_buffer(len(e))
                n = lz4.LZ4_decompress_safe(payload, out, len(payload), len(e))
                assert n == len(e), "stock liblz4 rejected the block"
                assert out.raw == oracle.shuffle(e, 35).tobytes()
    return info, total


def test_fixture_chunks_roundtrip(capi, golden_dir):
    import gzip, os
    text = gzip.open(os.path.join(golden_dir, "chr22.filtered.vcf.gz")).read()
    info, total = _check(capi, text, 3, "chr22", 0, [0, 1, 2])
    assert info.chunk_records == 250 and info.n_chunks == 4       # h5py auto-chunk for 1000 x 35 B
    assert total < 3 * 35 * 1000                                    # it does compress


@pytest.mark.parametrize("chunk_records", [0, 64, 100, 1075])
def test_random_vcf_chunks_roundtrip(capi, chunk_records):
    text, samples = synth.random_vcf(900, 37, seed=21, fmt="GT", kinds="mixed", multidigit=False)
    _check(capi, text, len(samples), "chr22", chunk_records, [0, 5, 36])
    _check(capi, text, len(samples), "", chunk_records, [1])         # several CHROM values, long contig name truncated to S5


def test_synthetic_shapes_chunks_roundtrip(capi):
    for mix in (0, 1):
        spec = capi.synth_spec(6000, 130, seed=5 + mix, mix=mix)
        text = capi.synth_header(spec) + capi.synth_host(spec)
        info, total = _check(capi, text, spec.n_samples, "chr22", 0, [0, 64, 129])
        assert info.total_bytes < info.raw_bytes


@pytest.mark.parametrize("mix,chunk_records", [(0, 1075), (0, 1465), (1, 2730), (1, 777), (0, 33)])
def test_long_segments_and_layout(capi, mix, chunk_records):
    """Chunk sizes of the bench shapes (1075 = 1.1M variants, 1465 = 3M) and the largest one: the bit-parallel
    allele encoder then walks several mask words per lane.  Also pins the buffer layout: [sample][chunk],
    16-byte aligned, fetch_all + layout agree with the per-sample fetch."""
    spec = capi.synth_spec(7000, 96, seed=31 + mix, mix=mix)
    text = capi.synth_header(spec) + capi.synth_host(spec)
    _check(capi, text, spec.n_samples, "chr22", chunk_records, [0, 1, 47, 95])
    p = capi.Parse.from_host(synth.body_of(text), spec.n_samples, region="chr22")
    fr = p.compress(chunk_records)
    offs, sizes = fr.layout()
    buf = fr.fetch_all()
    assert offs.shape == (96, fr.info.n_chunks) and (offs % 16 == 0).all()
    flat = offs.reshape(-1)
    assert (np.diff(flat.astype(np.int64)) >= sizes.reshape(-1)[:-1]).all()          # [sample][chunk] order, no overlap
    assert int(sizes.sum()) == fr.info.total_bytes and len(buf) == fr.info.padded_bytes
    for s in (0, 95):
        for k, f in enumerate(fr.sample(s)):
            o = int(offs[s, k])
            assert buf[o:o + int(sizes[s, k])].tobytes() == f
    # packed image: the same frames back to back (16-byte aligned), pageable and pinned destinations, pieces of any size
    import torch
    pk, poffs, psizes = fr.fetch_packed()
    assert np.array_equal(psizes, sizes) and (poffs % 16 == 0).all() and len(pk) == int(((sizes.astype(np.int64) + 15) // 16 * 16).sum())
    assert np.array_equal(poffs.reshape(-1)[1:], np.cumsum((sizes.reshape(-1).astype(np.uint64) + 15) // 16 * 16)[:-1])
    pin = torch.empty(len(pk) + 64, dtype=torch.uint8).pin_memory()
    pin.fill_(0xEE)
    tot, _, _ = fr.fetch_packed(out=(pin.data_ptr(), pin.numel()))
    assert tot == len(pk) and np.array_equal(pin.numpy()[:tot], pk) and (pin.numpy()[tot:] == 0xEE).all()
    for s in (0, 47, 95):
        for k, f in enumerate(fr.sample(s)):
            o = int(poffs[s, k])
            assert pk[o:o + int(psizes[s, k])].tobytes() == f
    # the two ways the packed image is made (hb_set_fetch_mode): whole frames gathered on the device, or templates once +
    # headers and tails per donor put together by host threads: the same bytes, pad bytes included, with 1 or many threads
    try:
        for mode, threads in ((1, 0), (2, 1), (2, 3), (2, 0)):
            capi.lib().hb_set_fetch_mode(mode)
            capi.lib().hb_set_host_threads(threads)
            pin.fill_(0xEE)
            tot2, o2, z2 = fr.fetch_packed(out=(pin.data_ptr(), pin.numel()))
            assert tot2 == len(pk) and np.array_equal(pin.numpy()[:tot2], pk) and (pin.numpy()[tot2:] == 0xEE).all()
            assert np.array_equal(o2, poffs) and np.array_equal(z2, psizes)
        # a destination that is not 16-byte aligned (the streaming stores need alignment: plain copies then)
        odd = np.empty(len(pk) + 64, np.uint8)
        shift = (1 - odd.ctypes.data) % 16 or 16
        tot3, _, _ = fr.fetch_packed(out=(odd.ctypes.data + shift, len(pk)))
        assert tot3 == len(pk) and np.array_equal(odd[shift:shift + tot3], pk)
    finally:
        capi.lib().hb_set_fetch_mode(0)
        capi.lib().hb_set_host_threads(0)
    fr.rerun(p)                                                                       # deterministic: same bytes again
    assert np.array_equal(fr.fetch_all(), buf)


def test_deep_site_matcher_is_smaller_and_exact(capi):
    """hb_set_site_matcher(1): 4-way hash buckets in the site encoder -- same decoded bytes, smaller frames."""
    spec = capi.synth_spec(6000, 40, seed=9)
    text = capi.synth_header(spec) + capi.synth_host(spec)
    info_fast, total_fast = _check(capi, text, 40, "chr22", 1075, [0, 39])
    capi.lib().hb_set_site_matcher(1)
    try:
        info_deep, total_deep = _check(capi, text, 40, "chr22", 1075, [0, 39])
    finally:
        capi.lib().hb_set_site_matcher(0)
    assert total_deep < total_fast and info_deep.site_lz4_bytes < info_fast.site_lz4_bytes


def test_tiny_and_empty(capi):
    S = ["a", "b"]
    one = (synth.header(S) + "chr22\t5\t.\tA\tC\t.\t.\t.\tGT\t0|1\t1|1\n").encode()
    _check(capi, one, 2, "chr22", 0, [0, 1])
    p = capi.Parse.from_host(synth.body_of(one), 2, region="chrNope")
    fr = p.compress(0)
    assert fr.info.n_chunks == 0 and fr.info.total_bytes == 0


def test_slab_streaming_rerun_with_other_record_counts(capi):
    """The streaming form used for inputs larger than HBM (tools/config4_stream.py): ONE device text buffer and ONE
    frames handle are re-used for slabs whose kept-record counts differ (explicit chunk_records keeps the chunk
    geometry); every slab's frames still decode to that slab's records."""
    import torch
    S, cr = 70, 64
    specs = [capi.synth_spec(nv, S, seed=50 + k, mix=1, first_pos=10_000_000 + 100_000 * k) for k, nv in enumerate((900, 1000, 760, 1030))]
    cap = max(int(capi.lib().hb_synth_body_bytes(sp)) for sp in specs)
    text = torch.zeros(cap + 256, dtype=torch.uint8, device="cuda")
    p = fr = None
    seen = set()
    for sp in specs:
        T = int(capi.lib().hb_synth_body_bytes(sp))
        capi.check(capi.lib().hb_synth_device(sp, text.data_ptr(), T, 0, None))
        text[T:T + 256].zero_()
        torch.cuda.synchronize()
        if p is None:
            p = capi.Parse.from_device(text.data_ptr(), T, S, region="chr22")
            fr = p.compress(cr)
            p.attach(fr)
        else:
            capi.check(capi.lib().hb_parse_rerun_bytes(p._h, T))
            fr.rerun(p)
        ora = oracle.parse_text(capi.synth_header(sp) + capi.synth_host(sp), "*", "chr22")
        assert p.info.n_records == ora["n"] == fr.info.n_records and fr.info.chunk_records == cr
        seen.add(ora["n"])
        for s in (0, S - 1):
            exp = _expected_chunks(ora, s, cr)
            frames = fr.sample(s)
            assert len(frames) == len(exp) == fr.info.n_chunks
            for f, e in zip(frames, exp):
                assert oracle.blosc_chunk_decode(f, len(e)).tobytes() == e
    assert len(seen) > 1                                   # the record count really changed between slabs
    p.attach(None)
    with pytest.raises(capi.HaploError):                   # without explicit chunk_records the geometry follows n_records
        q = capi.Parse.from_host(synth.body_of(capi.synth_header(specs[0]) + capi.synth_host(specs[0])), S, region="chr22")
        f2 = q.compress(0)
        q2 = capi.Parse.from_host(synth.body_of(capi.synth_header(specs[2]) + capi.synth_host(specs[2])), S, region="chr22")
        f2.rerun(q2)


def _pattern_vcf(n_variants, columns):
    """VCF whose sample k has the GT sequence columns[k](i) for record i (all records biallelic SNPs)."""
    S = synth.sample_names(len(columns))
    out = [synth.header(S)]
    for i in range(n_variants):
        out.append("\t".join(["chr22", str(1000 + 3 * i), ".", "ACGT"[i % 4], "CGTA"[i % 4], ".", "PASS", ".", "GT"] +
                             [f(i) for f in columns]) + "\n")
    return "".join(out).encode(), S


def test_allele_plane_encoder_on_degenerate_and_hostile_planes(capi):
    """The bit-parallel allele encoder is exact for ANY plane bytes: all zero / all one / all missing planes (one long
    match, or no match at all), period-2 and period-5 patterns, zero runs of every length 1..40 around single ones,
    runs ending exactly at segment and plane ends, multi-digit alleles (bytes with other bits set, incl. negative int8),
    hom/het structure between the two planes."""
    import random
    rng = random.Random(4)
    cols = [
        lambda i: "0|0", lambda i: "1|1", lambda i: "./.", lambda i: "0|1", lambda i: "1|0",
        lambda i: "%d|%d" % (i & 1, (i >> 1) & 1), lambda i: "%d|%d" % (i % 5 == 0, i % 5 == 3),
        lambda i: "%d|0" % (1 if any(i == k * (k + 3) // 2 for k in range(60)) else 0),      # zero runs of growing length
        lambda i: "0|%d" % (i % 67 == 66), lambda i: "%d|%d" % (i % 68 == 0, i % 68 == 67),  # run ends at segment ends (cr 1075 -> 68)
        lambda i: "10|200", lambda i: "%s|%s" % (rng.choice(["0", "1", "12", "127", "255", "."]), rng.choice(["0", "1", "7", "128"])),
        lambda i: "1|1" if (i // 97) & 1 else "0|0", lambda i: ".|1" if i % 11 == 0 else "0|0",
        lambda i: "%d|%d" % (rng.random() < 0.02, rng.random() < 0.02), lambda i: "%d|%d" % (rng.random() < 0.5, rng.random() < 0.5),
    ]
    text, samples = _pattern_vcf(2300, cols)
    for cr in (1075, 0, 64, 6, 2300):
        _check(capi, text, len(samples), "chr22", cr, list(range(len(samples))))


def test_frames_launched_from_inside_the_parse_are_the_same_frames(capi):
    """hb_parse_attach_frames: on a re-run of the parse the frame kernel is queued right behind the GT decoder from inside
    hb_parse_rerun (the host waits for the templates only) and hb_frames_rerun just collects: byte-identical to frames made
    the plain way, run after run, and a moved sample window falls back to the plain path."""
    spec = capi.synth_spec(5000, 150, seed=12, mix=1 << 8)
    text = capi.synth_header(spec) + capi.synth_host(spec)
    body = synth.body_of(text)
    plain = capi.Parse.from_host(body, 150, region="chr22")
    fp = plain.compress(300)
    want, o_want, z_want = fp.fetch_packed()
    p = capi.Parse.from_host(body, 150, region="chr22")
    fr = p.compress(300)
    p.attach(fr)
    for _ in range(3):
        p.rerun()
        fr.rerun(p)
        got, o, z = fr.fetch_packed()
        assert np.array_equal(z, z_want) and np.array_equal(o, o_want) and np.array_equal(got, want)
        assert fr.info.total_bytes == fp.info.total_bytes == int(z_want.sum())
    p.rerun()                                              # frames launched for all samples ...
    fr.set_window(10, 40)                                  # ... but the caller wants a window now
    fr.rerun(p)
    win = plain.compress(300, 10, 40)
    a, _, za = fr.fetch_packed()
    b, _, zb = win.fetch_packed()
    assert np.array_equal(za, zb) and np.array_equal(a, b)
    p.attach(None)
