/*
 * oracle/vcf_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or
 * executed from the product path (haplohyped-varawareml_b200/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * PARITY STATUS: "parity unpinned" for the parser arithmetic.  The reference parser
 * (/root/reference/cpp/parse_vcf.cpp + vendored cpp/vcfpp.h v0.3.8) needs htslib, which is
 * neither vendored in the reference nor installed in this image, so it cannot be compiled
 * (g++ stops at vcfpp.h:50 "htslib/kstring.h: No such file").  The reference's own tests
 * never execute the parser and hold no golden output for it.  This file is therefore a
 * scalar, obviously-correct C restatement of the reference call sites and of the published
 * htslib (>= 1.15, unpinned: environment.yml:16) text-VCF semantics they rely on.  It IS
 * pinned against known-answer vectors derived independently (pure-python field splitting)
 * from the reference's own fixture tests/data/chr22.filtered.vcf.gz: tests/golden/.
 *
 * What each function follows (reference file:line):
 *   orc_parse_text          cpp/parse_vcf.cpp:30-71 (load_vcf) and :80-113 (without sample)
 *   header handling         cpp/vcfpp.h:1378-1385 (open/bcf_hdr_read), :1413-1418 + :369-378
 *                           (setSamples -> bcf_hdr_set_samples, unknown sample -> runtime_error)
 *   region filter           cpp/vcfpp.h:1424-1451 (tbx_itr_querys on "chrN" or "chrN:b-e")
 *   is_snp                  cpp/vcfpp.h:990-1000
 *   decode_gt_field         cpp/vcfpp.h:546-588 over htslib vcf_parse_format's GT branch:
 *                           '.' -> missing, digits -> allele index, '|' '/' separators;
 *                           missing -> -9 (:567-572), allele -> bcf_gt_allele (:574)
 *   start/stop              cpp/vcfpp.h:1118-1127 (pos, pos + rlen; rlen = strlen(REF) unless a
 *                           header-declared Integer INFO/END > pos overrides it)
 *   int8 narrowing          cpp/parse_vcf.cpp:51-52
 *   ploidy == 2 assert      cpp/parse_vcf.cpp:46  (reported as an error instead of SIGABRT)
 *   shuffle / blosc / lz4   c-blosc 1.x (hdf5plugin filter 32001 = hdf5-blosc, vcf_to_h5.py:134-135):
 *                           published Blosc1 chunk format, LZ4 block format
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define ORC_OK 0
#define ORC_ERR_IO 1
#define ORC_ERR_HEADER 2
#define ORC_ERR_SAMPLE 3
#define ORC_ERR_PLOIDY 4
#define ORC_ERR_GT 5
#define ORC_ERR_FORMAT 6
#define ORC_ERR_NOGT 7
#define ORC_ERR_MEM 8
#define ORC_ERR_CODEC 9

typedef struct {
    uint64_t n;          /* records kept */
    uint64_t n_lines;    /* body lines seen */
    uint32_t n_samples;  /* samples in the header */
    uint32_t *start;     /* POS-1 */
    uint32_t *stop;      /* POS-1+rlen */
    char *ref;           /* 1 char per record (isSNP => strlen(REF) <= 1) */
    char *alt;           /* 1 char per record */
    uint32_t *chrom_off; /* offset into chrom_pool (NUL terminated strings) */
    char *chrom_pool;
    uint64_t chrom_pool_len;
    int8_t *gt0, *gt1;   /* one sample: n; matrix: [sample][n] */
    char err[256];
} orc_result;

static void set_err(orc_result *r, const char *msg) {
    snprintf(r->err, sizeof r->err, "%s", msg);
}

void orc_free(orc_result *r) {
    free(r->start); free(r->stop); free(r->ref); free(r->alt);
    free(r->chrom_off); free(r->chrom_pool); free(r->gt0); free(r->gt1);
    memset(r, 0, sizeof *r);
}

/* ---- region "chr", "chr:beg", "chr:beg-end" (1-based inclusive, commas ignored) ---- */
typedef struct { char chrom[256]; int has_chrom; int64_t beg0, end0; } region_t;

static void parse_region(const char *s, region_t *rg) {
    memset(rg, 0, sizeof *rg);
    rg->beg0 = 0; rg->end0 = INT64_MAX;
    if (!s || !*s) return;
    rg->has_chrom = 1;
    const char *colon = strrchr(s, ':');
    size_t nlen = colon ? (size_t)(colon - s) : strlen(s);
    if (nlen >= sizeof rg->chrom) nlen = sizeof rg->chrom - 1;
    memcpy(rg->chrom, s, nlen); rg->chrom[nlen] = 0;
    if (colon) {
        int64_t v = 0; const char *p = colon + 1; int any = 0;
        for (; *p && *p != '-'; ++p) if (*p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); any = 1; }
        if (any && v > 0) rg->beg0 = v - 1;
        if (*p == '-') {
            v = 0; any = 0;
            for (++p; *p; ++p) if (*p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); any = 1; }
            if (any) rg->end0 = v;
        }
    }
}

/* vcfpp.h:990-1000 */
static int is_snp(const char *ref, size_t ref_len, const char *alt, size_t alt_len) {
    if (ref_len > 1) return 0;
    size_t n_alt = 1;
    for (size_t i = 0; i < alt_len; ++i) if (alt[i] == ',') ++n_alt;
    if (alt_len == 1 && alt[0] == '.') return 0;      /* n_allele == 1: allele[1] is UB upstream */
    if (1 + n_alt > 2) return 0;
    if (alt_len != 1) return 0;
    return alt[0] == 'A' || alt[0] == 'C' || alt[0] == 'G' || alt[0] == 'T';
}

/*
 * GT sub-field decode.  p..end is ONE sample column (no tab); g = index of GT among the
 * ':'-separated FORMAT keys.  Returns ploidy (number of alleles parsed), or -1 if an allele is
 * neither digits nor '.'.  out[] gets up to 2 alleles, -9 for missing, int8-narrowed.
 */
static int decode_gt_field(const char *p, const char *end, int g, int8_t out[2]) {
    for (int k = 0; k < g; ++k) {
        while (p < end && *p != ':') ++p;
        if (p >= end) { out[0] = out[1] = -9; return 0; }   /* GT sub-field absent */
        ++p;
    }
    int l = 0;
    for (;;) {
        int32_t val;
        if (p < end && *p == '.') { ++p; val = -9; }
        else {
            const char *t = p; uint32_t v = 0;
            while (p < end && *p >= '0' && *p <= '9') { v = v * 10u + (uint32_t)(*p - '0'); ++p; }
            if (p == t) {
                if (l == 0 && (p >= end || *p == ':')) { out[0] = -9; out[1] = -9; return 1; } /* empty => one missing */
                return -1;
            }
            val = (int32_t)v;
        }
        if (l < 2) out[l] = (int8_t)val;
        ++l;
        if (p < end && (*p == '|' || *p == '/')) { ++p; continue; }
        break;
    }
    return l;
}

static const char *next_tab(const char *p, const char *end) {
    while (p < end && *p != '\t') ++p;
    return p;
}

typedef struct {
    const char *f[9]; size_t fl[9]; int nf; const char *samples;  /* samples -> first sample column start or NULL */
} line_fields;

static int split_head(const char *p, const char *end, line_fields *lf) {
    lf->nf = 0; lf->samples = NULL;
    while (lf->nf < 9) {
        const char *t = next_tab(p, end);
        lf->f[lf->nf] = p; lf->fl[lf->nf] = (size_t)(t - p); lf->nf++;
        if (t >= end) return 0;
        p = t + 1;
    }
    lf->samples = p;
    return 0;
}

static int64_t info_end(const char *info, size_t n) {
    /* key at start or after ';', exactly "END=" */
    size_t i = 0;
    while (i < n) {
        size_t j = i;
        while (j < n && info[j] != ';') ++j;
        if (j - i > 4 && memcmp(info + i, "END=", 4) == 0) {
            int64_t v = 0; size_t k = i + 4; int any = 0;
            for (; k < j && info[k] >= '0' && info[k] <= '9'; ++k) { v = v * 10 + (info[k] - '0'); any = 1; }
            if (any && k == j) return v;
            return -1;
        }
        i = j + 1;
    }
    return -1;
}

static int find_gt_index(const char *fmt, size_t n) {
    int idx = 0; size_t i = 0;
    while (i <= n) {
        size_t j = i;
        while (j < n && fmt[j] != ':') ++j;
        if (j - i == 2 && fmt[i] == 'G' && fmt[i + 1] == 'T') return idx;
        ++idx; i = j + 1;
    }
    return -1;
}

/*
 * Parse a whole decompressed VCF (header + body).
 *   sample  : NULL/""  -> no genotypes (load_vcf_without_sample)
 *             "*"      -> all samples, gt0/gt1 laid out [sample][n] (matrix form, test helper)
 *             name     -> that sample only (load_vcf)
 *   region  : "" -> every record; else contig[:beg-end]
 */
int orc_parse_text(const char *text, uint64_t len, const char *sample, const char *region, orc_result *r) {
    memset(r, 0, sizeof *r);
    const char *p = text, *tend = text + len;
    int end_is_int = 0, have_chromline = 0;
    const char *chromline = NULL, *chromline_end = NULL;
    while (p < tend && *p == '#') {
        const char *e = memchr(p, '\n', (size_t)(tend - p));
        if (!e) e = tend;
        if (p + 1 < tend && p[1] == '#') {
            if ((size_t)(e - p) > 11 && memcmp(p, "##INFO=<ID=", 11) == 0) {
                const char *q = p + 11;
                if ((size_t)(e - q) > 4 && memcmp(q, "END,", 4) == 0) {
                    /* Type=Integer anywhere in the line */
                    for (const char *s = q; s + 12 <= e; ++s)
                        if (memcmp(s, "Type=Integer", 12) == 0) { end_is_int = 1; break; }
                }
            }
        } else { chromline = p; chromline_end = e; have_chromline = 1; }
        p = (e < tend) ? e + 1 : tend;
        if (have_chromline) break;
    }
    if (!have_chromline) { set_err(r, "no #CHROM header line"); return ORC_ERR_HEADER; }
    if (chromline_end > chromline && chromline_end[-1] == '\r') --chromline_end;
    /* sample names */
    uint32_t ns = 0; int want = -1; int all = 0, with_gt = (sample && *sample);
    if (with_gt && strcmp(sample, "*") == 0) all = 1;
    {
        const char *q = chromline; int col = 0;
        while (q <= chromline_end) {
            const char *t = next_tab(q, chromline_end);
            if (col >= 9) {
                if (with_gt && !all && (size_t)(t - q) == strlen(sample) && memcmp(q, sample, (size_t)(t - q)) == 0 && want < 0)
                    want = (int)ns;
                ++ns;
            }
            ++col;
            if (t >= chromline_end) break;
            q = t + 1;
        }
    }
    r->n_samples = ns;
    if (with_gt && !all && want < 0) {
        snprintf(r->err, sizeof r->err, "the 1-th sample are not in the VCF.\nparameter samples:%s", sample);
        return ORC_ERR_SAMPLE;
    }
    region_t rg; parse_region(region, &rg);

    uint64_t cap = 1024, pool_cap = 64;
    r->start = malloc(cap * 4); r->stop = malloc(cap * 4); r->ref = malloc(cap); r->alt = malloc(cap);
    r->chrom_off = malloc(cap * 4); r->chrom_pool = malloc(pool_cap);
    uint64_t gcols = all ? ns : 1;
    /* matrix form is built row-major [record][sample] first, transposed at the end */
    int8_t *g0 = NULL, *g1 = NULL;
    if (with_gt) { g0 = malloc(cap * gcols); g1 = malloc(cap * gcols); }
    char last_chrom[256]; last_chrom[0] = 0; uint32_t last_off = 0; int have_last = 0;

    while (p < tend) {
        const char *e = memchr(p, '\n', (size_t)(tend - p));
        if (!e) e = tend;
        const char *le = e;
        if (le > p && le[-1] == '\r') --le;
        if (le == p || *p == '#') { p = (e < tend) ? e + 1 : tend; continue; }
        r->n_lines++;
        line_fields lf; split_head(p, le, &lf);
        p = (e < tend) ? e + 1 : tend;
        if (lf.nf < 8) { set_err(r, "truncated VCF record"); free(g0); free(g1); return ORC_ERR_FORMAT; }
        /* region (tabix semantics: same contig, overlap) */
        int64_t pos0 = 0;
        for (size_t i = 0; i < lf.fl[1]; ++i) {
            char c = lf.f[1][i];
            if (c < '0' || c > '9') break;
            pos0 = pos0 * 10 + (c - '0');
        }
        pos0 -= 1;
        int64_t rlen = (int64_t)lf.fl[3];
        if (end_is_int) {
            int64_t ev = info_end(lf.f[7], lf.fl[7]);
            if (ev > pos0) rlen = ev - pos0;
        }
        if (rg.has_chrom) {
            if (lf.fl[0] != strlen(rg.chrom) || memcmp(lf.f[0], rg.chrom, lf.fl[0]) != 0) continue;
            if (!(pos0 < rg.end0 && pos0 + rlen > rg.beg0)) continue;
        }
        if (!is_snp(lf.f[3], lf.fl[3], lf.f[4], lf.fl[4])) continue;
        if (r->n == cap) {
            cap *= 2;
            r->start = realloc(r->start, cap * 4); r->stop = realloc(r->stop, cap * 4);
            r->ref = realloc(r->ref, cap); r->alt = realloc(r->alt, cap);
            r->chrom_off = realloc(r->chrom_off, cap * 4);
            if (with_gt) { g0 = realloc(g0, cap * gcols); g1 = realloc(g1, cap * gcols); }
        }
        uint64_t i = r->n;
        if (with_gt) {
            if (lf.nf < 9 || !lf.samples) { set_err(r, "genotypes not present. make sure you initilized the variant object first\n"); free(g0); free(g1); return ORC_ERR_NOGT; }
            int g = find_gt_index(lf.f[8], lf.fl[8]);
            if (g < 0) { set_err(r, "genotypes not present. make sure you initilized the variant object first\n"); free(g0); free(g1); return ORC_ERR_NOGT; }
            const char *q = lf.samples; uint32_t s = 0;
            for (;;) {
                const char *t = next_tab(q, le);
                if (all || (int)s == want) {
                    int8_t a[2] = {0, 0};
                    int pl = decode_gt_field(q, t, g, a);
                    if (pl < 0) { set_err(r, "Couldn't read GT data: value not a number or '.'"); free(g0); free(g1); return ORC_ERR_GT; }
                    if (pl != 2) { set_err(r, "ploidy != 2 (reference: assert(var.ploidy() == 2), parse_vcf.cpp:46)"); free(g0); free(g1); return ORC_ERR_PLOIDY; }
                    uint64_t c = all ? s : 0;
                    g0[i * gcols + c] = a[0]; g1[i * gcols + c] = a[1];
                }
                ++s;
                if (t >= le) break;
                q = t + 1;
            }
            if (s != ns) { set_err(r, "Number of columns does not match the number of samples"); free(g0); free(g1); return ORC_ERR_FORMAT; }
        }
        r->start[i] = (uint32_t)pos0;
        r->stop[i] = (uint32_t)(pos0 + rlen);
        r->ref[i] = lf.fl[3] ? lf.f[3][0] : 0;
        r->alt[i] = lf.f[4][0];
        if (!have_last || strlen(last_chrom) != lf.fl[0] || memcmp(last_chrom, lf.f[0], lf.fl[0]) != 0) {
            size_t cl = lf.fl[0] < 255 ? lf.fl[0] : 255;
            while (r->chrom_pool_len + cl + 1 > pool_cap) { pool_cap *= 2; r->chrom_pool = realloc(r->chrom_pool, pool_cap); }
            memcpy(r->chrom_pool + r->chrom_pool_len, lf.f[0], cl);
            r->chrom_pool[r->chrom_pool_len + cl] = 0;
            last_off = (uint32_t)r->chrom_pool_len;
            r->chrom_pool_len += cl + 1;
            memcpy(last_chrom, lf.f[0], cl); last_chrom[cl] = 0; have_last = 1;
        }
        r->chrom_off[i] = last_off;
        r->n++;
    }
    if (with_gt) {
        if (all) {
            r->gt0 = malloc(r->n * gcols + 1); r->gt1 = malloc(r->n * gcols + 1);
            for (uint64_t i = 0; i < r->n; ++i)
                for (uint64_t s = 0; s < gcols; ++s) {
                    r->gt0[s * r->n + i] = g0[i * gcols + s];
                    r->gt1[s * r->n + i] = g1[i * gcols + s];
                }
            free(g0); free(g1);
        } else { r->gt0 = g0; r->gt1 = g1; }
    }
    return ORC_OK;
}

/* Read a .vcf / .vcf.gz (plain gzip or BGZF: zlib's gz* layer reads both) fully. */
int orc_read_file(const char *path, char **out, uint64_t *out_len) {
    gzFile f = gzopen(path, "rb");
    if (!f) return ORC_ERR_IO;
    gzbuffer(f, 1 << 20);
    uint64_t cap = 1 << 22, n = 0; char *buf = malloc(cap);
    for (;;) {
        if (cap - n < (1 << 20)) { cap *= 2; buf = realloc(buf, cap); }
        int k = gzread(f, buf + n, (unsigned)(1 << 20));
        if (k < 0) { gzclose(f); free(buf); return ORC_ERR_IO; }
        if (k == 0) break;
        n += (uint64_t)k;
    }
    gzclose(f);
    *out = buf; *out_len = n;
    return ORC_OK;
}

void orc_free_buf(char *p) { free(p); }

int orc_load_vcf(const char *path, const char *sample, const char *region, orc_result *r) {
    char *buf; uint64_t n;
    memset(r, 0, sizeof *r);
    if (orc_read_file(path, &buf, &n) != ORC_OK) { set_err(r, "cannot open VCF file"); return ORC_ERR_IO; }
    int rc = orc_parse_text(buf, n, sample, region, r);
    free(buf);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Storage side: 35-byte records (vcf_to_h5.py:119-127), Blosc byte-shuffle, LZ4 block decode,
 * Blosc chunk decode (+ a scalar chunk encoder, so that the decoders see chunks the GPU encoder did not write).
 * ------------------------------------------------------------------------------------------ */

/* Pack records exactly as np.array([...], dtype=[S5,u4,u4,S10,S10,i1,i1]) would: NUL pad,
 * silent truncation.  chrom is one NUL-terminated name per record via chrom_off/pool. */
void orc_pack_records(const orc_result *r, const int8_t *gt0, const int8_t *gt1, uint8_t *out) {
    for (uint64_t i = 0; i < r->n; ++i) {
        uint8_t *o = out + 35 * i;
        memset(o, 0, 35);
        const char *c = r->chrom_pool + r->chrom_off[i];
        size_t cl = strlen(c); if (cl > 5) cl = 5;
        memcpy(o, c, cl);
        memcpy(o + 5, &r->start[i], 4);
        memcpy(o + 9, &r->stop[i], 4);
        o[13] = (uint8_t)r->ref[i];
        o[23] = (uint8_t)r->alt[i];
        o[33] = (uint8_t)gt0[i];
        o[34] = (uint8_t)gt1[i];
    }
}

/* c-blosc shuffle_generic: plane j = byte j of every element; trailing n % typesize bytes copied */
void orc_shuffle(uint32_t typesize, uint64_t n, const uint8_t *src, uint8_t *dst) {
    uint64_t ne = n / typesize, rem = n % typesize;
    for (uint64_t j = 0; j < typesize; ++j)
        for (uint64_t i = 0; i < ne; ++i) dst[j * ne + i] = src[i * typesize + j];
    memcpy(dst + ne * typesize, src + ne * typesize, rem);
}

void orc_unshuffle(uint32_t typesize, uint64_t n, const uint8_t *src, uint8_t *dst) {
    uint64_t ne = n / typesize, rem = n % typesize;
    for (uint64_t i = 0; i < ne; ++i)
        for (uint64_t j = 0; j < typesize; ++j) dst[i * typesize + j] = src[j * ne + i];
    memcpy(dst + ne * typesize, src + ne * typesize, rem);
}

/* LZ4 block format decoder, bounds-checked, enforcing the end-of-block rules a stock decoder
 * relies on (last sequence is literals only). Returns decoded size or -1. */
int64_t orc_lz4_decode(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap) {
    uint64_t ip = 0, op = 0;
    if (n == 0) return -1;
    for (;;) {
        if (ip >= n) return -1;
        uint32_t tok = src[ip++];
        uint64_t ll = tok >> 4;
        if (ll == 15) { uint8_t b; do { if (ip >= n) return -1; b = src[ip++]; ll += b; } while (b == 255); }
        if (ip + ll > n || op + ll > cap) return -1;
        memcpy(dst + op, src + ip, ll); ip += ll; op += ll;
        if (ip == n) break;                     /* last sequence: literals only */
        if (ip + 2 > n) return -1;
        uint32_t off = src[ip] | ((uint32_t)src[ip + 1] << 8); ip += 2;
        if (off == 0 || off > op) return -1;
        uint64_t ml = tok & 15;
        if (ml == 15) { uint8_t b; do { if (ip >= n) return -1; b = src[ip++]; ml += b; } while (b == 255); }
        ml += 4;
        if (op + ml > cap) return -1;
        for (uint64_t k = 0; k < ml; ++k) { dst[op] = dst[op - off]; ++op; }
    }
    return (int64_t)op;
}

static uint32_t rd32le(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

/*
 * Blosc chunk decoder: what blosc_decompress of c-blosc 1.x (the library behind HDF5 filter 32001, hdf5-blosc)
 * does with one stored HDF5 chunk.  Published format (c-blosc README_HEADER.rst / blosc.c blosc_d):
 *   [0] version (BLOSC_VERSION_FORMAT = 2)   [1] versionlz (LZ4 family: 1)
 *   [2] flags: 0x01 byte-shuffle, 0x02 memcpyed, 0x04 bit-shuffle, 0x08 reserved, 0x10 don't split, codec format << 5
 *   [3] typesize   [4,8) nbytes   [8,12) blocksize   [12,16) cbytes (all LE32), then bstarts[nblocks] (LE32, from the
 *   start of the chunk), then per block: per stream an LE32 csize + csize bytes; csize == stream size means stored raw.
 *   A block is split into `typesize` streams iff !dont_split && typesize <= 16 && blocksize / typesize >= 128 && it is
 *   not the leftover block.  Shuffle is per block.
 * Also accepts the extended 32-byte header of a c-blosc2 chunk (both shuffle bits set; filters in [16,22)) with
 * byte-shuffle as the only filter: memcpyed chunks, zero-run streams (csize == 0).
 * Returns nbytes or -1.
 */
int64_t orc_blosc_chunk_decode(const uint8_t *c, uint64_t clen, uint8_t *dst, uint64_t cap) {
    if (clen < 16) return -1;
    uint8_t version = c[0], versionlz = c[1], flags = c[2], typesize = c[3];
    uint32_t nbytes = rd32le(c + 4), blocksize = rd32le(c + 8), cbytes = rd32le(c + 12);
    if (cbytes > clen || nbytes > cap) return -1;
    int extended = (flags & 0x01) && (flags & 0x04);
    uint32_t hdr = extended ? 32 : 16;
    if (cbytes < hdr) return -1;
    if (!extended && (version != 2 || (flags & 0x08))) return -1;   /* blosc.c: "version from future", reserved bit */
    if (extended && (version < 2 || version > 5)) return -1;
    int doshuffle = extended ? 0 : (flags & 0x01);
    if (extended) for (int i = 0; i < 6; ++i) {
        if (c[16 + i] == 1) doshuffle = 1;
        else if (c[16 + i] != 0) return -1;       /* other filters not used on this path */
    }
    if (nbytes == 0) return 0;
    if (flags & 0x02) {                           /* memcpyed */
        if ((uint64_t)hdr + nbytes > cbytes) return -1;
        memcpy(dst, c + hdr, nbytes);
        return nbytes;
    }
    if (extended && (c[31] >> 4) & 7) return -1;  /* special chunks not emitted on this path */
    uint32_t codec = flags >> 5;
    if (codec != 1) return -1;                    /* LZ4 / LZ4HC share format id 1 */
    if (!extended && versionlz != 1) return -1;   /* BLOSC_LZ4_VERSION_FORMAT */
    if (blocksize == 0 || blocksize > nbytes || typesize == 0) return -1;
    int dont_split = (flags & 0x10) != 0;
    uint32_t nblocks = (nbytes + blocksize - 1) / blocksize;
    uint64_t data0 = hdr + 4ull * nblocks;
    if (data0 > cbytes) return -1;
    uint8_t *tmp = malloc(blocksize);
    for (uint32_t b = 0; b < nblocks; ++b) {
        int leftover = (b == nblocks - 1) && (nbytes % blocksize);
        uint32_t bsize = leftover ? nbytes % blocksize : blocksize;
        uint32_t bstart = rd32le(c + hdr + 4 * b);
        int split = !dont_split && !leftover && (extended || (typesize <= 16 && blocksize / typesize >= 128));
        uint32_t nstreams = split ? typesize : 1;
        uint32_t neblock = bsize / nstreams;
        uint64_t ip = bstart; uint8_t *o = doshuffle ? tmp : dst + (uint64_t)b * blocksize;
        if (ip < data0 || ip > cbytes) { free(tmp); return -1; }
        for (uint32_t s = 0; s < nstreams; ++s) {
            if (ip + 4 > cbytes) { free(tmp); return -1; }
            int32_t cs = (int32_t)rd32le(c + ip); ip += 4;
            if (cs == 0 && extended) memset(o + (uint64_t)s * neblock, 0, neblock);
            else if (cs <= 0 || ip + (uint64_t)cs > cbytes) { free(tmp); return -1; }
            else if ((uint32_t)cs == neblock) { memcpy(o + (uint64_t)s * neblock, c + ip, neblock); ip += cs; }
            else {
                int64_t got = orc_lz4_decode(c + ip, (uint64_t)cs, o + (uint64_t)s * neblock, neblock);
                if (got != (int64_t)neblock) { free(tmp); return -1; }
                ip += cs;
            }
        }
        if (doshuffle) orc_unshuffle(typesize, bsize, tmp, dst + (uint64_t)b * blocksize);
    }
    free(tmp);
    return nbytes;
}

/*
 * Blosc1 chunk ENCODER, scalar: what blosc_compress(clevel, doshuffle = 1, typesize, ...) of c-blosc 1.x emits with
 * the LZ4 codec when the whole buffer is one block (blocksize = nbytes, no split: typesize > 16 or forced), using a
 * plain greedy single-entry-hash LZ4 matcher.  Its LZ4 stream is NOT byte-identical to stock LZ4/LZ4HC (those are
 * encoder choices, not format); it exists so that the CPU tests have chunks written by something other than the GPU
 * encoder to feed the decoders with, incl. several blocks per chunk and split streams.
 * blocksize: 0 = one block.  split != 0: split blocks into typesize streams where c-blosc's rule allows it.
 * Returns cbytes or -1 (dst too small).
 */
static int64_t lz4_greedy(const uint8_t *src, uint32_t n, uint8_t *dst, uint64_t cap) {
    uint32_t table[4096]; memset(table, 0xff, sizeof table);
    uint64_t op = 0; uint32_t anchor = 0, ip = 0;
#define PUT(b) do { if (op >= cap) return -1; dst[op++] = (uint8_t)(b); } while (0)
    if (n >= 13) {
        const uint32_t mflimit = n - 12;                       /* last match must start >= 12 bytes before the end */
        while (ip < mflimit) {
            uint32_t v; memcpy(&v, src + ip, 4);
            uint32_t h = (v * 2654435761u) >> 20;
            uint32_t cand = table[h]; table[h] = ip;
            uint32_t cv = 0; if (cand != 0xffffffffu) memcpy(&cv, src + cand, 4);
            if (cand == 0xffffffffu || ip - cand > 65535 || cv != v) { ++ip; continue; }
            uint32_t ml = 4; const uint32_t lim = n - 5;       /* last 5 bytes are literals */
            while (ip + ml < lim && src[ip + ml] == src[cand + ml]) ++ml;
            uint32_t lit = ip - anchor, m = ml - 4;
            PUT(((lit < 15 ? lit : 15) << 4) | (m < 15 ? m : 15));
            if (lit >= 15) { uint32_t r = lit - 15; while (r >= 255) { PUT(255); r -= 255; } PUT(r); }
            for (uint32_t k = 0; k < lit; ++k) PUT(src[anchor + k]);
            PUT((ip - cand) & 255); PUT((ip - cand) >> 8);
            if (m >= 15) { uint32_t r = m - 15; while (r >= 255) { PUT(255); r -= 255; } PUT(r); }
            ip += ml; anchor = ip;
        }
    }
    uint32_t lit = n - anchor;
    PUT((lit < 15 ? lit : 15) << 4);
    if (lit >= 15) { uint32_t r = lit - 15; while (r >= 255) { PUT(255); r -= 255; } PUT(r); }
    for (uint32_t k = 0; k < lit; ++k) PUT(src[anchor + k]);
#undef PUT
    return (int64_t)op;
}

int64_t orc_blosc1_chunk_encode(const uint8_t *src, uint32_t nbytes, uint32_t typesize, uint32_t blocksize, int split,
                                uint8_t *dst, uint64_t cap) {
    if (typesize == 0 || typesize > 255 || nbytes == 0) return -1;
    if (blocksize == 0 || blocksize > nbytes) blocksize = nbytes;
    if (blocksize > typesize) blocksize = blocksize / typesize * typesize;
    uint32_t nblocks = (nbytes + blocksize - 1) / blocksize;
    uint64_t op = 16 + 4ull * nblocks;
    if (cap < op) return -1;
    int can_split = split && typesize <= 16 && blocksize / typesize >= 128;
    dst[0] = 2; dst[1] = 1; dst[2] = (uint8_t)(0x01 | (can_split ? 0 : 0x10) | (1 << 5)); dst[3] = (uint8_t)typesize;
    uint8_t *tmp = malloc(blocksize);
    for (uint32_t b = 0; b < nblocks; ++b) {
        int leftover = (b == nblocks - 1) && (nbytes % blocksize);
        uint32_t bsize = leftover ? nbytes % blocksize : blocksize;
        orc_shuffle(typesize, bsize, src + (uint64_t)b * blocksize, tmp);
        uint32_t bstart = (uint32_t)op; memcpy(dst + 16 + 4 * b, &bstart, 4);
        uint32_t nstreams = (can_split && !leftover) ? typesize : 1, ne = bsize / nstreams;
        for (uint32_t s = 0; s < nstreams; ++s) {
            if (op + 4 + ne > cap) { free(tmp); return -1; }
            int64_t cs = lz4_greedy(tmp + (uint64_t)s * ne, ne, dst + op + 4, ne - 1);
            if (cs <= 0) { cs = ne; memcpy(dst + op + 4, tmp + (uint64_t)s * ne, ne); }     /* incompressible: stored raw */
            int32_t c32 = (int32_t)cs; memcpy(dst + op, &c32, 4);
            op += 4 + (uint64_t)cs;
        }
    }
    free(tmp);
    uint32_t cb = (uint32_t)op;
    memcpy(dst + 4, &nbytes, 4); memcpy(dst + 8, &blocksize, 4); memcpy(dst + 12, &cb, 4);
    return (int64_t)op;
}
