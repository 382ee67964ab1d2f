"""Import-gated acceptance tests against the STOCK third-party codecs / containers the reference uses.  None of
python-blosc, blosc2, h5py or hdf5plugin is in this image or on the GPU boxes (profiles/r02_probe_box.txt), so every test
here skips today; they are what flips "parity unpinned" for the storage rows the day one of the packages appears.

Reference side: h5py + hdf5plugin write `snp_data` with compression=32001 (hdf5-blosc: c-blosc 1.x chunks), vcf_to_h5.py:
134-135; h5_reader.py:37-41 reads it back through the same filter; tests/test_compression.py:45-55,86-108 round-trips."""
import os

import numpy as np
import pytest

import oracle
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi(built):
    from haplohyped_varawareml_b200 import capi as c
    return c


def _frames_and_records(capi, n=2600, s=5):
    text, samples = synth.random_vcf(n, 9, seed=8, fmt="GT", kinds="mixed")
    ora = oracle.parse_text(text, "*", "chr22")
    p = capi.Parse.from_host(synth.body_of(text), len(samples), region="chr22")
    fr = p.compress(0)
    cr = int(fr.info.chunk_records)
    rec = oracle.records_from_columns(ora["chrom"], ora["start"], ora["stop"], ora["ref"], ora["alt"], ora["gt0"][s], ora["gt1"][s])
    return fr.sample(s), rec, cr


def test_stock_python_blosc_decompresses_gpu_chunks(capi):
    blosc = pytest.importorskip("blosc")                     # python-blosc (c-blosc 1.x): the library behind filter 32001
    frames, rec, cr = _frames_and_records(capi)
    raw = rec.tobytes() + b"\0" * (len(frames) * cr * 35 - rec.nbytes)
    for k, f in enumerate(frames):
        assert blosc.decompress(f) == raw[k * cr * 35:(k + 1) * cr * 35]


def test_stock_blosc2_decompresses_gpu_chunks(capi):
    blosc2 = pytest.importorskip("blosc2")                   # c-blosc2 reads Blosc1 chunks (north_star: "round-trips through stock Blosc2")
    frames, rec, cr = _frames_and_records(capi)
    raw = rec.tobytes() + b"\0" * (len(frames) * cr * 35 - rec.nbytes)
    for k, f in enumerate(frames):
        assert bytes(blosc2.decompress(f)) == raw[k * cr * 35:(k + 1) * cr * 35]


def test_gpu_decoder_reads_stock_blosc_chunks(capi):
    blosc = pytest.importorskip("blosc")
    frames, rec, cr = _frames_and_records(capi)
    data = rec[:cr].tobytes()
    for cname in ("lz4hc", "lz4"):
        c = blosc.compress(data, typesize=35, clevel=5, shuffle=blosc.SHUFFLE, cname=cname)
        assert capi.decode_frames([c], len(data))[0].tobytes() == data
        assert c[:4] == oracle.reference_like_chunk(data)[:4]            # the framing the oracle restates


def test_h5py_reads_a_file_this_repo_wrote(capi, tmp_path):
    h5py = pytest.importorskip("h5py")
    pytest.importorskip("hdf5plugin")
    import hdf5plugin  # noqa: F401
    from haplohyped_varawareml_b200 import container
    frames, rec, cr = _frames_and_records(capi)
    for backend in ("minih5", "h5py"):
        path = str(tmp_path / (backend + ".h5"))
        with container.open_h5(path, "w", backend=backend) as f:
            f.write_chunked("donor_x/chr_22/snp_data", oracle.RECORD_DTYPE, len(rec), cr, frames)
        with h5py.File(path, "r") as f:                      # the reference's reader: h5_reader.py:37-41
            d = f["donor_x/chr_22/snp_data"]
            assert d.chunks == (cr,) and d.dtype == oracle.RECORD_DTYPE
            assert np.array_equal(d[()], rec)


def test_this_repo_reads_a_file_h5py_wrote(capi, tmp_path):
    h5py = pytest.importorskip("h5py")
    pytest.importorskip("hdf5plugin")
    import hdf5plugin  # noqa: F401
    from haplohyped_varawareml_b200 import container
    frames, rec, cr = _frames_and_records(capi)
    path = str(tmp_path / "ref.h5")
    with h5py.File(path, "w") as f:                          # the reference's writer: vcf_to_h5.py:131-135
        f.create_group("donor_x/chr_22").create_dataset("snp_data", data=rec, compression=32001,
                                                        compression_opts=(2, 2, 0, 0, 5, 1, 2), chunks=True)
    for backend in ("minih5", "h5py"):
        with container.open_h5(path, "r", backend=backend) as f:
            assert np.array_equal(f.read_dataset("donor_x/chr_22/snp_data"), rec)
