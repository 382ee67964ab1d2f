"""CPU tests of the host-side logic: C-ABI exports, encode dict mirror, shard planning, and the
world_size-2 metadata exchange over gloo."""
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol(built):
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "haplo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(os.path.join(ROOT, "haplohyped-varawareml_b200", "libhaplo_b200.so"))
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, f"declared in include/haplo_b200.h but not exported: {missing}"
    from haplohyped_varawareml_b200 import capi
    assert sorted(capi.EXPORTS) == declared
    lib.hb_version.restype = ctypes.c_char_p
    assert b"haplo_b200" in lib.hb_version()


def test_no_gpu_means_loud_failure_not_fallback(built):
    """Without a device the product path must fail (HB_ERR_CUDA), never compute on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from haplohyped_varawareml_b200 import capi
    with pytest.raises(capi.HaploError) as e:
        capi.Parse.from_host(b"chr22\t1\t.\tA\tC\t.\t.\t.\tGT\t0|1\n", 1)
    assert e.value.code == 9 and "no CPU fallback" in str(e.value)
    import sys
    sys.path.insert(0, os.path.join(ROOT, "haplohyped-varawareml_b200"))
    import parse_vcf
    with pytest.raises(RuntimeError, match="Error parsing VCF file"):
        parse_vcf.load_vcf(os.path.join(ROOT, "tests/golden/chr22.filtered.vcf.gz"), "x", "chr22")


def test_pybind_module_surface(built):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "haplohyped-varawareml_b200"))
    import parse_vcf
    assert hasattr(parse_vcf, "VCFLoader") and hasattr(parse_vcf.VCFLoader, "load_vcf")
    assert hasattr(parse_vcf.VCFLoader, "load_vcf_without_sample")
    assert callable(parse_vcf.load_vcf) and callable(parse_vcf.load_vcf_without_sample)   # vcf_to_h5.py:101 call form
    assert "chrom" in parse_vcf.VCFLoader.load_vcf.__doc__ and "sample" in parse_vcf.VCFLoader.load_vcf.__doc__


def test_encode_dict_mirror(golden_dir):
    import json
    from haplohyped_varawareml_b200.common_utils import parse_encode_dict, build_lut
    for case in json.load(open(os.path.join(golden_dir, "encode_dict.json")))["parse_encode_dict"]:
        assert parse_encode_dict(case["input"]) == case["output"]       # outputs of the reference's own function
    with pytest.raises(TypeError):
        parse_encode_dict(123)
    lut = build_lut(None)
    assert lut[ord("A")] == 0 and lut[ord("t")] == 3 and lut[ord("N")] == 4 and lut[ord("X")] == 4
    assert build_lut("ACGT")[ord("N")] == -1


def test_synth_host_is_deterministic_and_well_formed(built):
    from haplohyped_varawareml_b200 import capi
    import oracle
    spec = capi.synth_spec(500, 33, seed=9, mix=1)
    a, b = capi.synth_host(spec), capi.synth_host(spec)
    assert a == b and len(a) == capi.lib().hb_synth_body_bytes(spec)
    lines = a.split(b"\n")[:-1]
    assert len(lines) == 500 and all(len(l.split(b"\t")) == 9 + 33 for l in lines)
    mid = b"\n".join(lines[100:150]) + b"\n"
    assert capi.synth_host(spec, 100, 50) == mid
    ora = oracle.parse_text(capi.synth_header(spec) + a, "*", "chr22")
    assert 0 < ora["n"] < 500                                     # multiallelic / indel sites are dropped
    assert set(np.unique(ora["gt0"])) <= {-9, 0, 1}
    pos = ora["start"]
    assert (np.diff(pos.astype(np.int64)) > 0).all()               # strictly increasing positions


def test_plan_shards_lpt():
    from haplohyped_varawareml_b200.shard import plan_shards, global_row_offsets
    grch38 = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50]
    bins = plan_shards(grch38, 8)
    assert sorted(i for b in bins for i in b) == list(range(22))
    loads = [sum(grch38[i] for i in b) for b in bins]
    assert max(loads) <= 1.15 * sum(grch38) / 8 + max(grch38) * 0.2   # LPT is near-balanced
    assert plan_shards([5, 3], 4) == [[0], [1], [], []]
    assert global_row_offsets([{"n_records": 3}, {"n_records": 0}, {"n_records": 5}]) == [0, 3, 3]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from haplohyped_varawareml_b200.shard import gather_metadata, global_row_offsets, plan_shards
    mine = plan_shards([30, 10, 20, 25], world)[rank]
    meta = {"n_records": 100 * (rank + 1), "n_lines": 110 * (rank + 1), "text_bytes": 1000 + rank,
            "first_pos": 10 * rank, "last_pos": 10 * rank + 9, "out_bytes": 7}
    g = gather_metadata(meta)
    q.put((rank, mine, g, global_row_offsets(g)))
    dist.destroy_process_group()


def test_metadata_allgather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in ps])
    [p.join(timeout=60) for p in ps]
    assert res[0][1] == [0] and res[1][1] == [1, 2, 3] or sorted(res[0][1] + res[1][1]) == [0, 1, 2, 3]
    for rank, mine, g, offs in res:
        assert [x["n_records"] for x in g] == [100, 200] and offs == [0, 100]
        assert g[1]["text_bytes"] == 1001 and g[0]["last_pos"] == 9


def test_fasta_encoder_writes_the_reference_file_the_dataset_reads(tmp_path):
    """fasta_encoder mirror: FASTA (plain and gzip, wrapped lines, extra contigs) -> reference_genome.h5 with one
    byte-per-base dataset per chr1..chr22 -- the file RandomHaplotypeDataset(hdf5_reference_file=...) opens."""
    import gzip
    from haplohyped_varawareml_b200 import fasta_encoder as fe
    rng = np.random.default_rng(2)
    seqs = {f"chr{c}": bytes(rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), size=int(rng.integers(50, 400)))) for c in (1, 2, 22)}
    seqs["chrUn_x"] = b"ACGT" * 5
    text = b""
    for k, v in seqs.items():
        text += b">" + k.encode() + b" some description\n" + b"\n".join(v[i:i + 60] for i in range(0, len(v), 60)) + b"\n"
    plain, gz = tmp_path / "ref.fa", tmp_path / "ref.fa.gz"
    plain.write_bytes(text)
    with gzip.open(gz, "wb") as f:
        f.write(text)
    for src in (plain, gz):
        out = tmp_path / ("o_" + src.name)
        fe.main(["--fasta", str(src), "--outdir", str(out), "--cores", "2"])
        assert not (out / "tmp_chrom_files").exists()
        got = fe.HDF5Handler.load_from_hdf5(str(out / "reference_genome.h5"))
        assert sorted(got) == ["chr1", "chr2", "chr22"]                 # chr1..22 only, as the reference's chrom list
        for k in got:
            assert got[k].tobytes() == seqs[k]
    assert fe.ReferenceGenome.parse_encode_list(None) == [b"A", b"C", b"G", b"T", b"N"]


def test_converter_spreads_chromosome_files_over_devices(tmp_path, built):
    """vcf_to_h5 --devices: chromosome files are bin-packed onto the GPUs by size, largest first on every GPU (no GPU
    work here: only the plan)."""
    from haplohyped_varawareml_b200 import vcf_to_h5 as v2h
    vdir = tmp_path / "vcf"
    vdir.mkdir()
    sizes = {1: 900, 2: 870, 3: 700, 7: 560, 12: 470, 19: 200, 21: 150, 22: 170}
    for c, n in sizes.items():
        (vdir / f"chr{c}.filtered.vcf.gz").write_bytes(b"x" * n)
    (tmp_path / "donors.txt").write_text("a\nb\n")
    conv = v2h.VCFtoHDF5Converter("c", str(vdir), str(tmp_path / "o"), str(tmp_path / "donors.txt"), 2, 4, devices=[0, 1, 2])
    plan = conv.plan_devices(sorted(sizes))
    assert sorted(c for b in plan for c in b) == sorted(sizes) and len(plan) == 3
    loads = [sum(sizes[c] for c in b) for b in plan]
    assert max(loads) - min(loads) <= max(sizes.values()) / 2            # LPT: bins within half the largest file of each other
    for b in plan:
        assert [sizes[c] for c in b] == sorted((sizes[c] for c in b), reverse=True)
    assert conv.devices == [0, 1, 2]
    assert v2h.VCFtoHDF5Converter("c", str(vdir), str(tmp_path / "o"), str(tmp_path / "donors.txt"), 2, 4, device=3).devices == [3]
