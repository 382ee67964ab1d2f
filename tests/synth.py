"""Small hand-authored / randomised VCF texts covering the edge cases the reference pins nowhere
(SURVEY.md section 8c "edge vectors").  Pure python; used by both the CPU and the GPU tests."""
from __future__ import annotations

import random

HEADER = (
    "##fileformat=VCFv4.2\n"
    '##FILTER=<ID=PASS,Description="All filters passed">\n'
    "##contig=<ID=chr21>\n##contig=<ID=chr22>\n##contig=<ID=chr22_KI270731v1_random>\n"
    '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
    '##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype Quality">\n'
    '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Read Depth">\n'
    '##INFO=<ID=AF,Number=A,Type=Float,Description="Allele Frequency">\n'
    "{extra}"
)


def header(samples, extra=""):
    h = HEADER.format(extra=extra)
    return h + "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples) + "\n"


def sample_names(n):
    return ["s%04d" % i for i in range(n)]


def random_gt(rng, kinds="phased"):
    if kinds == "phased":
        return "%d|%d" % (rng.random() < 0.3, rng.random() < 0.3)
    r = rng.random()
    a, b = int(rng.random() < 0.3), int(rng.random() < 0.3)
    if r < 0.80:
        return "%d|%d" % (a, b)
    if r < 0.88:
        return "%d/%d" % (a, b)
    if r < 0.91:
        return "./."
    if r < 0.94:
        return ".|."
    if r < 0.97:
        return ".|%d" % a
    return "%d/." % a


def random_vcf(n_variants, n_samples, seed=0, fmt="GT", kinds="mixed", chrom="chr22", site_mix=True,
               multidigit=False, crlf=False, info_end=False):
    """Returns (text, samples).  fmt: "GT", "GT:GQ:DP" or "DP:GT" (GT not first)."""
    rng = random.Random(seed)
    samples = sample_names(n_samples)
    extra = '##INFO=<ID=END,Number=1,Type=Integer,Description="End">\n' if info_end else ""
    out = [header(samples, extra)]
    pos = 10_000_000
    keys = fmt.split(":")
    nl = "\r\n" if crlf else "\n"
    if crlf:
        out[0] = out[0].replace("\n", "\r\n")
    for i in range(n_variants):
        pos += rng.randint(1, 400)
        ref = rng.choice("ACGT")
        alt = rng.choice([b for b in "ACGT" if b != ref])
        c = chrom
        if site_mix:
            r = rng.random()
            if r < 0.04:
                alt = alt + "," + rng.choice([b for b in "ACGT" if b not in (ref, alt)])   # multiallelic
            elif r < 0.07:
                ref = ref + "TG"                                                           # deletion
            elif r < 0.10:
                alt = ref + "A"                                                            # insertion
            elif r < 0.12:
                alt = "*"
            elif r < 0.14:
                alt = "<DEL>"
            elif r < 0.16:
                alt = alt.lower()
            elif r < 0.18:
                ref = "N"
            elif r < 0.20:
                c = "chr21" if rng.random() < 0.5 else "chr22_KI270731v1_random"
        info = "AF=0.5"
        if info_end and rng.random() < 0.3:
            info = rng.choice(["END=%d" % (pos + rng.randint(0, 50)), "AF=0.1;END=%d;X" % (pos + 7), "ENDX=5",
                               "END=%d" % (pos - 5)])
        cols = []
        for _ in range(n_samples):
            gt = random_gt(rng, kinds)
            if multidigit and rng.random() < 0.02:
                gt = rng.choice(["10|1", "1|12", "200/3", "127|128", "255|256"])
            sub = []
            for k in keys:
                if k == "GT":
                    sub.append(gt)
                elif k == "GQ":
                    sub.append(str(rng.randint(0, 99)))
                else:
                    sub.append(str(rng.randint(1, 60)))
            cols.append(":".join(sub))
        out.append("\t".join([c, str(pos), "rs%d" % i, ref, alt, ".", "PASS", info, fmt] + cols) + nl)
    return "".join(out).encode(), samples


def body_of(text: bytes) -> bytes:
    """Strip the header lines (everything up to and including the #CHROM line)."""
    i = text.index(b"#CHROM")
    j = text.index(b"\n", i)
    return text[j + 1:]


def bgzf_compress(data: bytes, level: int = 6, block: int = 0xff00, strategy=None) -> bytes:
    """What `bgzip` writes: <= 64 KiB gzip members with a 'BC' extra subfield + the 28-byte EOF member
    (htslib bgzf.c).  Stock zlib does the DEFLATE, so this is an independent producer for the GPU inflater."""
    import struct
    import zlib
    out = []
    for i in range(0, len(data), block):
        chunk = data[i:i + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY if strategy is None else strategy)
        comp = c.compress(chunk) + c.flush()
        assert len(comp) + 26 <= 65536
        out.append(struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord("B"), ord("C"), 2, len(comp) + 25))
        out.append(comp)
        out.append(struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(out)
