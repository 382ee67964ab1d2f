"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``haplohyped-varawareml_b200/``) never does: it fails loudly when its CUDA library is missing.

PARITY STATUS: "parity unpinned" for the parser (the reference parser needs htslib, absent here;
see ``vcf_oracle.c``).  Pinned against the fixture-derived known answers in ``tests/golden``.

Python side:
  * ctypes binding of ``liboracle.so`` (parser / shuffle / LZ4 / Blosc chunk restatement in C);
  * numpy restatement of the dataset leg (``haplotype_dataset.py:11-16,54-110`` and
    ``common_utils.py:62-103``) with the repairs R1-R4 listed in SURVEY.md section 8(a).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Result(C.Structure):
    _fields_ = [
        ("n", C.c_uint64), ("n_lines", C.c_uint64), ("n_samples", C.c_uint32),
        ("start", C.POINTER(C.c_uint32)), ("stop", C.POINTER(C.c_uint32)),
        ("ref", C.POINTER(C.c_char)), ("alt", C.POINTER(C.c_char)),
        ("chrom_off", C.POINTER(C.c_uint32)), ("chrom_pool", C.POINTER(C.c_char)),
        ("chrom_pool_len", C.c_uint64),
        ("gt0", C.POINTER(C.c_int8)), ("gt1", C.POINTER(C.c_int8)),
        ("err", C.c_char * 256),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "vcf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_parse_text.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_char_p, C.POINTER(_Result)]
        L.orc_load_vcf.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(_Result)]
        L.orc_free.argtypes = [C.POINTER(_Result)]
        L.orc_shuffle.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_unshuffle.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]
        for f in (L.orc_lz4_decode, L.orc_blosc_chunk_decode):
            f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
            f.restype = C.c_int64
        L.orc_blosc1_chunk_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64]
        L.orc_blosc1_chunk_encode.restype = C.c_int64
        L.orc_pack_records.argtypes = [C.POINTER(_Result), C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("Error parsing VCF file: " + msg)
        self.code = code


RECORD_DTYPE = np.dtype([("chrom", "S5"), ("start", np.uint32), ("stop", np.uint32),
                         ("ref", "S10"), ("alt", "S10"), ("phase1", np.int8), ("phase2", np.int8)])
assert RECORD_DTYPE.itemsize == 35


def _harvest(res: _Result, sample):
    n = int(res.n)
    out = {
        "n": n, "n_lines": int(res.n_lines), "n_samples": int(res.n_samples),
        "start": np.ctypeslib.as_array(res.start, (n,)).copy() if n else np.zeros(0, np.uint32),
        "stop": np.ctypeslib.as_array(res.stop, (n,)).copy() if n else np.zeros(0, np.uint32),
        "ref": np.frombuffer(C.string_at(res.ref, n), dtype="S1").copy() if n else np.zeros(0, "S1"),
        "alt": np.frombuffer(C.string_at(res.alt, n), dtype="S1").copy() if n else np.zeros(0, "S1"),
    }
    pool = C.string_at(res.chrom_pool, int(res.chrom_pool_len)) if res.chrom_pool_len else b""
    offs = np.ctypeslib.as_array(res.chrom_off, (n,)).copy() if n else np.zeros(0, np.uint32)
    names = {}
    for o in np.unique(offs):
        names[int(o)] = pool[int(o):pool.index(b"\0", int(o))].decode()
    out["chrom"] = [names[int(o)] for o in offs]
    if sample:
        rows = int(res.n_samples) if sample == "*" else 1
        shape = (rows, n) if sample == "*" else (n,)
        if n:
            out["gt0"] = np.ctypeslib.as_array(res.gt0, (rows * n,)).copy().reshape(shape)
            out["gt1"] = np.ctypeslib.as_array(res.gt1, (rows * n,)).copy().reshape(shape)
        else:
            out["gt0"] = np.zeros(shape, np.int8)
            out["gt1"] = np.zeros(shape, np.int8)
    return out


def parse_text(text: bytes, sample: str | None = None, region: str = ""):
    """Restates VCFLoader::load_vcf / load_vcf_without_sample on decompressed VCF text.

    sample=None/"" -> sites only; "*" -> every sample as a [S, n] matrix; name -> that sample."""
    res = _Result()
    rc = lib().orc_parse_text(text, len(text), (sample or "").encode(), region.encode(), C.byref(res))
    try:
        if rc != 0:
            raise OracleError(rc, res.err.decode(errors="replace"))
        return _harvest(res, sample)
    finally:
        lib().orc_free(C.byref(res))


def load_vcf(path: str, sample: str, chrom: str = ""):
    """list[tuple] exactly as the reference's pybind11 ``load_vcf`` returns (parse_vcf.cpp:118-121)."""
    res = _Result()
    rc = lib().orc_load_vcf(path.encode(), sample.encode(), chrom.encode(), C.byref(res))
    try:
        if rc != 0:
            raise OracleError(rc, res.err.decode(errors="replace"))
        d = _harvest(res, sample)
    finally:
        lib().orc_free(C.byref(res))
    return [(d["chrom"][i], int(d["start"][i]), int(d["stop"][i]), d["ref"][i].decode(), d["alt"][i].decode(),
             int(d["gt0"][i]), int(d["gt1"][i])) for i in range(d["n"])]


def load_vcf_without_sample(path: str, chrom: str = ""):
    res = _Result()
    rc = lib().orc_load_vcf(path.encode(), b"", chrom.encode(), C.byref(res))
    try:
        if rc != 0:
            raise OracleError(rc, res.err.decode(errors="replace"))
        d = _harvest(res, None)
    finally:
        lib().orc_free(C.byref(res))
    return [(d["chrom"][i], int(d["start"][i]), int(d["stop"][i]), d["ref"][i].decode(), d["alt"][i].decode())
            for i in range(d["n"])]


def records_from_tuples(rows) -> np.ndarray:
    """vcf_to_h5.py:119-129 -- tuples -> 35-byte struct array (NUL pad, silent truncation)."""
    return np.array([(r[0].encode()[:5], r[1], r[2], r[3].encode()[:10], r[4].encode()[:10], r[5], r[6])
                     for r in rows], dtype=RECORD_DTYPE)


def records_from_columns(chrom, start, stop, ref, alt, gt0, gt1) -> np.ndarray:
    n = len(start)
    rec = np.zeros(n, dtype=RECORD_DTYPE)
    rec["chrom"] = np.array([c.encode()[:5] for c in chrom], dtype="S5") if n else np.zeros(0, "S5")
    rec["start"], rec["stop"] = start, stop
    rec["ref"], rec["alt"] = ref, alt
    rec["phase1"], rec["phase2"] = gt0, gt1
    return rec


def shuffle(buf: bytes | np.ndarray, typesize: int) -> np.ndarray:
    a = np.ascontiguousarray(np.frombuffer(bytes(buf), np.uint8))
    out = np.empty_like(a)
    lib().orc_shuffle(typesize, a.size, a.ctypes.data, out.ctypes.data)
    return out


def unshuffle(buf, typesize: int) -> np.ndarray:
    a = np.ascontiguousarray(np.frombuffer(bytes(buf), np.uint8))
    out = np.empty_like(a)
    lib().orc_unshuffle(typesize, a.size, a.ctypes.data, out.ctypes.data)
    return out


def _decode(fn, buf, cap):
    a = np.ascontiguousarray(np.frombuffer(bytes(buf), np.uint8))
    out = np.empty(max(cap, 1), np.uint8)
    n = fn(a.ctypes.data, a.size, out.ctypes.data, cap)
    if n < 0:
        raise ValueError("oracle: corrupt stream")
    return out[:n]


def lz4_decode(buf, cap: int) -> np.ndarray:
    return _decode(lib().orc_lz4_decode, buf, cap)


def blosc_chunk_decode(buf, cap: int) -> np.ndarray:
    """One stored HDF5 chunk of filter 32001 (a bare Blosc1 chunk) -> its `cap` uncompressed bytes."""
    return _decode(lib().orc_blosc_chunk_decode, buf, cap)


def blosc1_chunk_encode(data, typesize: int, blocksize: int = 0, split: bool = False) -> bytes:
    """Scalar Blosc1 chunk writer (byte-shuffle + greedy LZ4): chunks for the decoders that the GPU encoder did not write."""
    a = np.ascontiguousarray(np.frombuffer(bytes(data), np.uint8))
    nblocks = 1 if not blocksize else (a.size + blocksize - 1) // blocksize + 1
    cap = 16 + 4 * nblocks + a.size + 4 * (nblocks * (typesize if split else 1)) + 64
    out = np.empty(cap, np.uint8)
    n = lib().orc_blosc1_chunk_encode(a.ctypes.data, a.size, typesize, blocksize, int(split), out.ctypes.data, cap)
    if n < 0:
        raise ValueError("oracle: chunk encode failed")
    return out[:n].tobytes()


# --------------------------------------------------------------------------------------------
# What the reference's writer stores: blosc_compress of c-blosc 1.x (hdf5-blosc, filter 32001) with
# the opts of vcf_to_h5.py:135 -- clevel 5, shuffle 1, compcode 2 = LZ4HC -- restated around the STOCK
# LZ4HC codec of the system liblz4 (the one third-party piece of that writer that is in the image).
# [third-party, not in /root/reference: c-blosc 1.21.x blosc.c compute_blocksize / split_block /
#  blosc_c / lz4hc_wrap_compress, published source; hdf5plugin >= 4 bundles it, requirements.txt]
# --------------------------------------------------------------------------------------------
def stock_lz4():
    """ctypes handle of the system liblz4 (LZ4_compress_HC, LZ4_compress_default, LZ4_decompress_safe) or None."""
    import ctypes.util
    for name in ("liblz4.so.1", ctypes.util.find_library("lz4")):
        if not name:
            continue
        try:
            L = C.CDLL(name)
            L.LZ4_compress_HC.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
            L.LZ4_compress_default.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
            L.LZ4_decompress_safe.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
            return L
        except (OSError, AttributeError):
            continue
    return None


def cblosc1_blocksize(nbytes: int, typesize: int, clevel: int = 5, hcr: bool = True) -> int:
    """blosc.c compute_blocksize for a codec that never splits this typesize (35 > MAX_SPLITS = 16)."""
    L1 = 32 * 1024
    if nbytes < typesize:
        return 1
    bs = nbytes
    if nbytes >= L1:
        bs = L1 * (2 if hcr else 1)
        bs = {0: bs // 4, 1: bs // 2, 2: bs, 3: bs * 2, 4: bs * 4, 5: bs * 4, 6: bs * 8, 7: bs * 8, 8: bs * 8,
              9: bs * 8 * (2 if hcr else 1)}[clevel]
    bs = min(bs, nbytes)
    if bs > typesize:
        bs = bs // typesize * typesize
    return bs


def reference_like_chunk(data, typesize: int = 35, clevel: int = 5, hc: bool = True) -> bytes:
    """The stored HDF5 chunk the reference's writer produces for `data` (one HDF5 chunk of 35-byte records): Blosc1
    header, bstarts, per block byte-shuffle + one stock LZ4HC(clevel) stream (stored raw where LZ4 does not gain)."""
    L = stock_lz4()
    if L is None:
        raise RuntimeError("no system liblz4")
    data = bytes(data)
    n = len(data)
    bs = cblosc1_blocksize(n, typesize, clevel, hc)
    nblocks = (n + bs - 1) // bs
    body, bstarts = b"", []
    pos = 16 + 4 * nblocks
    for b in range(nblocks):
        blk = data[b * bs:(b + 1) * bs]
        sh = shuffle(blk, typesize).tobytes()
        out = C.create_string_buffer(len(sh) + 64)
        # blosc_c: maxout = neblock (+ the codec's slack for LZ4); a stream that does not gain is stored raw
        cs = (L.LZ4_compress_HC(sh, out, len(sh), len(sh) - 1, clevel) if hc else L.LZ4_compress_default(sh, out, len(sh), len(sh) - 1))
        payload = out.raw[:cs] if cs > 0 else sh
        bstarts.append(pos)
        body += np.int32(len(payload)).tobytes() + payload
        pos += 4 + len(payload)
    hdr = bytes([2, 1, 0x01 | 0x10 | (1 << 5), typesize]) + np.array([n, bs, pos], "<u4").tobytes()
    return hdr + np.array(bstarts, "<u4").tobytes() + body


# --------------------------------------------------------------------------------------------
# h5py auto-chunk heuristic (h5py/_hl/filters.py guess_chunk), restated: vcf_to_h5.py:134-135
# passes chunks=True.  [third-party, not in /root/reference: h5py >= 3.0, requirements.txt]
# --------------------------------------------------------------------------------------------
def guess_chunk_1d(n: int, itemsize: int = 35) -> int:
    CHUNK_BASE, CHUNK_MIN, CHUNK_MAX = 16 * 1024, 8 * 1024, 1024 * 1024
    if n == 0:
        return 1
    chunk = float(n)
    dset_size = chunk * itemsize
    target = CHUNK_BASE * (2 ** np.log10(dset_size / (1024.0 * 1024)))
    if target > CHUNK_MAX:
        target = CHUNK_MAX
    elif target < CHUNK_MIN:
        target = CHUNK_MIN
    while True:
        chunk_bytes = chunk * itemsize
        if (chunk_bytes < target or abs(chunk_bytes - target) / target < 0.5) and chunk_bytes < CHUNK_MAX:
            break
        if chunk == 1:
            break
        chunk = np.ceil(chunk / 2.0)
    return int(chunk)


# --------------------------------------------------------------------------------------------
# Dataset leg (numpy restatement; integer arithmetic, exact)
# --------------------------------------------------------------------------------------------
def parse_encode_dict(encode_spec):
    """common_utils.py:62-79."""
    if not encode_spec:
        return {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4}
    elif isinstance(encode_spec, (list, tuple, str)):
        return {base: i for i, base in enumerate(encode_spec)}
    elif isinstance(encode_spec, dict):
        return encode_spec
    raise TypeError("Please input as dict, list or string!")


def calculate_midpoint_region(start, end, seq_length):
    """haplotype_dataset.py:11-16."""
    midpt = (start + end) // 2
    half = seq_length // 2
    return max(0, midpt - half), midpt + half


def base_to_index(seq: np.ndarray, encode_spec) -> np.ndarray:
    """Intent of common_utils.py:84-103: upper-case, anything outside ACGT -> N, index in
    encode_spec key order (repair R3; the reference code returns all zeros).  A spec
    without an "N" key gives -1 = an all-zero one-hot row, which is what reindex(columns=base_list)
    at :86 produces for a category that is not a column."""
    spec = parse_encode_dict(encode_spec)
    lut = np.full(256, spec.get("N", -1), dtype=np.int8)
    for b in "ACGT":
        if b in spec:
            lut[ord(b)] = spec[b]
            lut[ord(b.lower())] = spec[b]
    return lut[np.frombuffer(bytes(seq), np.uint8) if not isinstance(seq, np.ndarray) else seq.view(np.uint8)]


def encode_haplotypes(ref_window: np.ndarray, start, ref, alt, p1, p2, new_start, new_end, encode_spec=None):
    """haplotype_dataset.py:86-110 with repairs R1 (initialise from the reference window) and
    R2 (only variants with new_start <= start < new_end).  Kept literally: phase == 1 selects the
    ALT index, anything else the *VCF REF* index (:99-100); duplicates: last record wins
    (np.put_along_axis).  Returns two int8 index vectors of len(ref_window)."""
    spec = parse_encode_dict(encode_spec)
    idx = base_to_index(ref_window, spec)
    hap1, hap2 = idx.copy(), idx.copy()
    L = len(ref_window)
    # record REF/ALT go through the same byte -> class map as the window (the reference's
    # np.vectorize(dict.get) at :96-97 yields None for anything outside the spec and cannot run)
    ridx = base_to_index(np.asarray(ref, dtype="S1"), spec) if len(start) else []
    aidx = base_to_index(np.asarray(alt, dtype="S1"), spec) if len(start) else []
    for i in range(len(start)):
        s = int(start[i])
        if not (new_start <= s < new_end) or s - new_start >= L:
            continue
        r, a = ridx[i], aidx[i]
        hap1[s - new_start] = a if p1[i] == 1 else r
        hap2[s - new_start] = a if p2[i] == 1 else r
    return hap1, hap2


def onehot(idx: np.ndarray, n_classes: int) -> np.ndarray:
    """Repair R3: out[i, c] = float(c == idx[i])."""
    out = np.zeros((len(idx), n_classes), np.float32)
    ok = (idx >= 0) & (idx < n_classes)
    out[np.nonzero(ok)[0], idx[ok]] = 1.0
    return out
