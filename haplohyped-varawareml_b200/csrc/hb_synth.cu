// hb_synth.cu -- seeded synthetic VCF text (bench + tests), identical bytes from the CUDA
// generator and from the host generator (configs 2, 3 and 4).
//
// Record i:  CHROM \t POS \t rs<i> \t REF \t ALT \t . \t PASS \t . \t GT { \t a sep b } x S \n
// Everything is a pure function of (seed, i, s) through a counter-based hash, so any tile of the
// text can be regenerated anywhere (GPU for the bench input, CPU for the oracle's sample).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/haplo_b200.h"
#include "hb_internal.h"

#define HD __host__ __device__ __forceinline__

namespace hb {

HD uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
HD uint64_t site_hash(uint64_t seed, uint64_t i) { return mix64(seed ^ mix64(i)); }
HD uint64_t call_hash(uint64_t sh, uint32_t s) { return mix64(sh ^ mix64(0x5851F42D4C957F2Dull * (s + 1))); }
// ALT-allele frequency of a site ~ u^4 (u uniform): most sites rare, a few common -- mean 0.2, harsher than 1000G.
// Bits 8-15 of spec.mix ask for more squarings (1: u^8, mean 0.11, about Beta(0.2, 2); 2: u^16, mean 0.06).
HD uint32_t af_threshold(uint64_t sh, uint32_t mix) {
    uint64_t u = sh >> 32;
    u = (u * u) >> 32;
    u = (u * u) >> 32;
    for (uint32_t k = (mix >> 8) & 0xffu; k; --k) u = (u * u) >> 32;
    return (uint32_t)u;
}
HD int ndigits(uint64_t v) {
    int n = 1;
    while (v >= 10) { v /= 10; ++n; }
    return n;
}

struct SiteDesc {
    uint32_t pos;
    char ref[2], alt[3];
    int ref_len, alt_len;
};

HD SiteDesc site_desc(const hb_synth_spec &sp, uint64_t i, uint64_t sh) {
    const char B[4] = {'A', 'C', 'G', 'T'};
    SiteDesc d;
    d.pos = sp.first_pos + (uint32_t)(i * sp.pos_step) + (uint32_t)((sh >> 40) % sp.pos_step);
    int r = (int)(sh & 3), a = (r + 1 + (int)((sh >> 2) % 3)) & 3;
    d.ref[0] = B[r]; d.ref_len = 1;
    d.alt[0] = B[a]; d.alt_len = 1;
    if ((sp.mix & 0xffu) == 1) {
        uint32_t sk = (uint32_t)((sh >> 8) % 100);
        if (sk >= 90 && sk < 95) {                 // multiallelic: dropped by the SNP filter
            int a2 = (a + 1) & 3;
            if (a2 == r) a2 = (a2 + 1) & 3;
            d.alt[1] = ','; d.alt[2] = B[a2]; d.alt_len = 3;
        } else if (sk >= 95) {                     // indel: dropped
            if (sk & 1) { d.ref[1] = 'T'; d.ref_len = 2; d.alt[0] = B[r]; }
            else { d.alt[0] = B[r]; d.alt[1] = 'T'; d.alt[2] = 'G'; d.alt_len = 3; }
        }
    }
    return d;
}

HD uint32_t head_len(const hb_synth_spec &sp, int chrom_len, uint64_t i, const SiteDesc &d) {
    // CHROM \t POS \t rs<i> \t REF \t ALT \t . \t PASS \t . \t GT
    return chrom_len + 1 + ndigits(d.pos) + 1 + 2 + ndigits(i) + 1 + d.ref_len + 1 + d.alt_len + 1 + 1 + 1 + 4 + 1 + 1 + 1 + 2;
}

HD int put_uint(char *o, uint64_t v) {
    int n = ndigits(v);
    for (int k = n - 1; k >= 0; --k) { o[k] = (char)('0' + v % 10); v /= 10; }
    return n;
}

HD uint32_t write_head(const hb_synth_spec &sp, int chrom_len, uint64_t i, const SiteDesc &d, char *o) {
    uint32_t n = 0;
    for (int k = 0; k < chrom_len; ++k) o[n++] = sp.chrom[k];
    o[n++] = '\t';
    n += put_uint(o + n, d.pos);
    o[n++] = '\t'; o[n++] = 'r'; o[n++] = 's';
    n += put_uint(o + n, i);
    o[n++] = '\t';
    for (int k = 0; k < d.ref_len; ++k) o[n++] = d.ref[k];
    o[n++] = '\t';
    for (int k = 0; k < d.alt_len; ++k) o[n++] = d.alt[k];
    o[n++] = '\t'; o[n++] = '.'; o[n++] = '\t';
    o[n++] = 'P'; o[n++] = 'A'; o[n++] = 'S'; o[n++] = 'S';
    o[n++] = '\t'; o[n++] = '.'; o[n++] = '\t'; o[n++] = 'G'; o[n++] = 'T';
    return n;
}

// the 4 bytes "\t a sep b" of sample s at site hash sh, little-endian packed
HD uint32_t call_word(const hb_synth_spec &sp, uint64_t sh, uint32_t thr, uint32_t s) {
    uint64_t ch = call_hash(sh, s);
    uint32_t a = ((uint32_t)ch < thr) ? '1' : '0';
    uint32_t b = ((uint32_t)(ch >> 32) < thr) ? '1' : '0';
    uint32_t sep = '|';
    if ((sp.mix & 0xffu) == 1) {
        uint64_t k = mix64(ch);
        uint32_t ck = (uint32_t)(k % 100);
        if (ck == 0) sep = '/';
        else if (ck == 1) {
            uint32_t sub = (uint32_t)((k >> 8) % 3);
            if (sub == 0) { a = '.'; b = '.'; sep = '/'; }
            else if (sub == 1) { a = '.'; b = '.'; }
            else a = '.';
        }
    }
    return (uint32_t)'\t' | (a << 8) | (sep << 16) | (b << 24);
}

static int chrom_len_of(const hb_synth_spec &sp) {
    int n = 0;
    while (n < 15 && sp.chrom[n]) ++n;
    return n;
}

// one CTA per record
__global__ void __launch_bounds__(256)
synth_kernel(const hb_synth_spec sp, int chrom_len, const uint64_t *__restrict__ line_off, uint8_t *__restrict__ out) {
    __shared__ char s_head[96];
    __shared__ uint32_t s_hl, s_thr;
    __shared__ uint64_t s_sh;
    const uint64_t i = blockIdx.x;
    if (threadIdx.x == 0) {
        uint64_t sh = site_hash(sp.seed, i);
        SiteDesc d = site_desc(sp, i, sh);
        s_hl = write_head(sp, chrom_len, i, d, s_head);
        s_thr = af_threshold(sh, sp.mix);
        s_sh = sh;
    }
    __syncthreads();
    const uint64_t base = line_off[i];
    const uint32_t hl = s_hl, thr = s_thr;
    const uint64_t sh = s_sh;
    const uint64_t total = hl + 4ull * sp.n_samples + 1;
    for (uint64_t k = threadIdx.x; k < hl; k += blockDim.x) out[base + k] = (uint8_t)s_head[k];
    uint8_t *q = out + base + hl;
    const bool aligned = (((uintptr_t)q) & 3) == 0;
    for (uint64_t s = threadIdx.x; s < sp.n_samples; s += blockDim.x) {
        uint32_t w = call_word(sp, sh, thr, (uint32_t)s);
        if (aligned) reinterpret_cast<uint32_t *>(q)[s] = w;
        else { q[4 * s] = (uint8_t)w; q[4 * s + 1] = (uint8_t)(w >> 8); q[4 * s + 2] = (uint8_t)(w >> 16); q[4 * s + 3] = (uint8_t)(w >> 24); }
    }
    if (threadIdx.x == 0) out[base + total - 1] = '\n';
}

static void line_offsets(const hb_synth_spec &sp, uint64_t first, uint64_t n, std::vector<uint64_t> &off) {
    int cl = chrom_len_of(sp);
    off.resize(n + 1);
    uint64_t acc = 0;
    for (uint64_t k = 0; k < n; ++k) {
        uint64_t i = first + k;
        SiteDesc d = site_desc(sp, i, site_hash(sp.seed, i));
        off[k] = acc;
        acc += head_len(sp, cl, i, d) + 4ull * sp.n_samples + 1;
    }
    off[n] = acc;
}

}  // namespace hb

using namespace hb;

extern "C" {

uint64_t hb_synth_body_bytes(const hb_synth_spec *s) {
    std::vector<uint64_t> off;
    line_offsets(*s, 0, s->n_variants, off);
    return off.back();
}

int hb_synth_header(const hb_synth_spec *s, char *buf, uint64_t cap, uint64_t *len) {
    std::string h = "##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n";
    h += "##contig=<ID=" + std::string(s->chrom, chrom_len_of(*s)) + ">\n";
    h += "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n";
    h += "##source=haplo_b200_synth\n";
    h += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
    char name[16];
    for (uint32_t k = 0; k < s->n_samples; ++k) { snprintf(name, sizeof name, "\tS%06u", k); h += name; }
    h += "\n";
    if (len) *len = h.size();
    if (buf) {
        if (h.size() > cap) return HB_ERR_ARG;
        memcpy(buf, h.data(), h.size());
    }
    return HB_OK;
}

int hb_synth_device(const hb_synth_spec *s, uint8_t *d_text, uint64_t cap, int device, void *stream) {
    if (cudaSetDevice(device) != cudaSuccess) return HB_ERR_CUDA;
    std::vector<uint64_t> off;
    line_offsets(*s, 0, s->n_variants, off);
    if (off.back() > cap) return HB_ERR_ARG;
    uint64_t *d_off = nullptr;
    if (cudaMalloc(&d_off, off.size() * 8) != cudaSuccess) return HB_ERR_MEM;
    cudaMemcpyAsync(d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (s->n_variants) {
        synth_kernel<<<(unsigned)s->n_variants, 256, 0, (cudaStream_t)stream>>>(*s, chrom_len_of(*s), d_off, d_text);
        count_launch();
    }
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(d_off);
    return e == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

int hb_synth_host(const hb_synth_spec *s, uint64_t first_variant, uint64_t n, uint8_t *buf, uint64_t cap,
                  uint64_t *len) {
    std::vector<uint64_t> off;
    line_offsets(*s, first_variant, n, off);
    if (len) *len = off.back();
    if (!buf) return HB_OK;
    if (off.back() > cap) return HB_ERR_ARG;
    int cl = chrom_len_of(*s);
    for (uint64_t k = 0; k < n; ++k) {
        uint64_t i = first_variant + k;
        uint64_t sh = site_hash(s->seed, i);
        SiteDesc d = site_desc(*s, i, sh);
        char *o = (char *)buf + off[k];
        uint32_t hl = write_head(*s, cl, i, d, o);
        uint32_t thr = af_threshold(sh, s->mix);
        for (uint32_t sm = 0; sm < s->n_samples; ++sm) {
            uint32_t w = call_word(*s, sh, thr, sm);
            memcpy(o + hl + 4ull * sm, &w, 4);
        }
        o[hl + 4ull * s->n_samples] = '\n';
    }
    return HB_OK;
}

}  // extern "C"
