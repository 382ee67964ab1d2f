#!/usr/bin/env python
"""Regenerates tests/golden/* from the reference checkout (run in the build container only;
/root/reference does not exist on the GPU box, so tests read only the committed outputs).

Sources:
  * fixture files copied verbatim from /root/reference/tests/data (data, not source code);
  * expected load_vcf tuples derived by an INDEPENDENT pure-python field split (str.split) under
    the rules of parse_vcf.cpp:41-61 + vcfpp.h:990-1000,546-588 -- a third implementation beside
    the C oracle and the CUDA path;
  * parse_encode_dict outputs produced by IMPORTING the reference's own
    src/utils/common_utils.py (the one function on this path that runs as shipped).
"""
import gzip
import importlib.util
import json
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def expected_tuples(text: str, sample: str, chrom: str):
    out = []
    names = None
    for line in text.split("\n"):
        if line.startswith("##") or not line:
            continue
        f = line.split("\t")
        if line.startswith("#CHROM"):
            names = f[9:]
            continue
        col = 9 + names.index(sample)
        if chrom and f[0] != chrom:
            continue
        ref, alt = f[3], f[4]
        if len(ref) > 1 or len(alt.split(",")) > 1 or alt not in ("A", "C", "G", "T"):
            continue
        gt = f[col].split(":")[f[8].split(":").index("GT")]
        a = gt.replace("|", "/").split("/")
        assert len(a) == 2
        g = [(-9 if x == "." else ((int(x) + 128) % 256) - 128) for x in a]
        out.append([f[0], int(f[1]) - 1, int(f[1]) - 1 + len(ref), ref, alt, g[0], g[1]])
    return out


def main():
    for name in ("chr22.filtered.vcf.gz", "ipscs_samples_test.txt", "test_regions.bed"):
        shutil.copyfile(os.path.join(REF, "tests/data", name), os.path.join(HERE, name))
    text = gzip.open(os.path.join(HERE, "chr22.filtered.vcf.gz")).read().decode()
    samples = open(os.path.join(HERE, "ipscs_samples_test.txt")).read().split()
    gold = {"samples": samples, "chrom": "chr22",
            "load_vcf": {s: expected_tuples(text, s, "chr22") for s in samples}}
    json.dump(gold, open(os.path.join(HERE, "fixture_load_vcf.json"), "w"))

    spec = importlib.util.spec_from_file_location("ref_common_utils", os.path.join(REF, "src/utils/common_utils.py"))
    cu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cu)
    cases = [None, "", ["A", "C", "G", "T"], "ACGTN", {"A": 0, "C": 1, "G": 2, "T": 3}, ("T", "G", "C", "A", "N")]
    enc = [{"input": c, "output": cu.parse_encode_dict(c)} for c in cases]
    json.dump({"parse_encode_dict": enc}, open(os.path.join(HERE, "encode_dict.json"), "w"), indent=1)
    print("records per sample:", {s: len(v) for s, v in gold["load_vcf"].items()})


if __name__ == "__main__":
    main()
