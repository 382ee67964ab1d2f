// hb_head.cuh -- the nine fixed columns of one VCF record, as a byte-fed state machine.
//
// One definition of what the reference reaches through vcfpp accessors after vcf_parse1
//   isSNP  cpp/vcfpp.h:990-1000    CHROM :1076    Start/End :1118-1127    REF :1130    ALT :1142-1151
// shared by the two kernels that read record heads: sites_kernel (lines located by the tokenizer)
// and walk_kernel (lines located by chaining, hb_walk.cu), so both give the same rows by construction.
#pragma once
#include "hb_common.cuh"
#include "hb_internal.h"

namespace hb {

struct HeadState {
    uint32_t pos, ref_len, alt_len, n_comma, fmt_len, chrom_len;
    uint8_t ref0, alt0;
    int g;                       // index of the GT key inside FORMAT, -1 if absent
    long long endval;            // INFO END= value (first END key only), -1 if none
    int f;                       // completed columns (tabs seen)
    bool chrom_ok, pos_digits, has_samples;
    uint64_t samp_abs;           // offset of the 9th TAB (the one that precedes sample 0)
    // INFO END= : 0 matching the key, 1 in the value, 2 skip to ';'
    int ist, kpos;
    long long ev;
    bool ev_any, end_seen;
    // FORMAT keys
    int ki, kl;
    uint8_t k0, k1;

    __device__ __forceinline__ void init() {
        pos = ref_len = alt_len = n_comma = fmt_len = chrom_len = 0;
        ref0 = alt0 = 0;
        g = -1;
        endval = -1;
        f = 0;
        chrom_ok = pos_digits = true;
        has_samples = false;
        samp_abs = 0;
        ist = kpos = 0;
        ev = 0;
        ev_any = end_seen = false;
        ki = kl = 0;
        k0 = k1 = 0;
    }

    // Feed the byte at offset q (never '\n' / the "\r\n" pair: the caller stops there).
    // Returns true once the TAB before sample 0 has been consumed.
    __device__ __forceinline__ bool feed(uint8_t c, uint64_t q, const RegionArg &rg) {
        if (c == '\t') {
            if (f == 7 && ist == 1 && ev_any) endval = ev;
            if (f == 8) { if (kl == 2 && k0 == 'G' && k1 == 'T' && g < 0) g = ki; }
            ++f;
            if (f == 9) { samp_abs = q; has_samples = true; return true; }
            return false;
        }
        switch (f) {
            case 0:
                if (rg.has_region) {
                    if (chrom_len >= rg.chrom_len || (uint8_t)rg.chrom[chrom_len] != c) chrom_ok = false;
                }
                ++chrom_len;
                break;
            case 1:
                if (pos_digits && c >= '0' && c <= '9') pos = pos * 10u + (uint32_t)(c - '0');
                else pos_digits = false;
                break;
            case 3:
                if (ref_len == 0) ref0 = c;
                ++ref_len;
                break;
            case 4:
                if (alt_len == 0) alt0 = c;
                if (c == ',') ++n_comma;
                ++alt_len;
                break;
            case 7:
                if (c == ';') {
                    if (ist == 1 && ev_any) endval = ev;
                    ist = 0; kpos = 0; ev = 0; ev_any = false;
                } else if (ist == 0) {
                    const uint32_t key = 0x3D444E45u;           // "END="
                    if (c == (uint8_t)(key >> (8 * kpos))) {
                        if (++kpos == 4) { ist = end_seen ? 2 : 1; end_seen = true; }
                    } else ist = 2;
                } else if (ist == 1) {
                    if (c >= '0' && c <= '9') { ev = ev * 10 + (c - '0'); ev_any = true; }
                    else { ist = 2; ev_any = false; }
                }
                break;
            case 8:
                ++fmt_len;
                if (c == ':') {
                    if (kl == 2 && k0 == 'G' && k1 == 'T' && g < 0) g = ki;
                    ++ki; kl = 0;
                } else {
                    if (kl == 0) k0 = c; else if (kl == 1) k1 = c;
                    ++kl;
                }
                break;
            default: break;
        }
        return false;
    }

    // The line ended before a 9th TAB: close the column that was open.
    __device__ __forceinline__ void finish_short() {
        if (f == 7 && ist == 1 && ev_any) endval = ev;
        if (f == 8 && kl == 2 && k0 == 'G' && k1 == 'T' && g < 0) g = ki;
    }
};

// What a fully read head means for the record: kept or not, its coordinates, how its samples decode.
struct HeadVerdict {
    bool malformed, keep, uniform;
    uint32_t start, stop;
};

// le = offset of the end of the record's text (the '\n', or the '\r' of a "\r\n")
__device__ __forceinline__ HeadVerdict judge_head(const HeadState &h, uint64_t le, uint32_t n_samples,
                                                  const RegionArg &rg, int end_is_int) {
    HeadVerdict v;
    v.malformed = h.f < 7;
    v.keep = v.uniform = false;
    v.start = v.stop = 0;
    if (v.malformed) return v;
    const long long pos0 = (long long)h.pos - 1;
    long long rlen = h.ref_len;
    if (end_is_int && h.endval > pos0) rlen = h.endval - pos0;
    bool in_region = true;
    if (rg.has_region)
        in_region = h.chrom_ok && h.chrom_len == rg.chrom_len && pos0 < rg.end0 && pos0 + rlen > rg.beg0;
    const bool snp = h.ref_len <= 1 && h.n_comma == 0 && h.alt_len == 1 &&
                     (h.alt0 == 'A' || h.alt0 == 'C' || h.alt0 == 'G' || h.alt0 == 'T');
    v.keep = in_region && snp;
    v.start = (uint32_t)pos0;
    v.stop = (uint32_t)(pos0 + rlen);
    if (h.has_samples) v.uniform = h.fmt_len == 2 && h.g == 0 && (le - h.samp_abs) == 4ull * n_samples;
    return v;
}

// Outputs of the site stage (dense rows in file order) -- written by sites_kernel or walk_kernel.
struct SiteOut {
    uint32_t *start, *stop;
    uint8_t *ref, *alt;
    uint64_t *chrom_abs;
    uint8_t *chrom_len;
    uint64_t *chrom5;
    RowInfo *rowinfo;
    uint32_t *nu_rows;
    DevStatus *st;
};

// Write row `row`.  cp_row_by_line: checkpoint-table row when the tabs tokenizer indexed by line
// (kNoCpRow = allocate a slot from the non-uniform counter instead).
constexpr uint32_t kNoCpRow = 0xffffffffu;
__device__ __forceinline__ void write_site_row(const SiteOut &o, const uint8_t *__restrict__ text, uint64_t row,
                                               uint64_t ls, uint64_t le, const HeadState &h, const HeadVerdict &v,
                                               int want_gt, uint32_t cp_row_by_line) {
    o.start[row] = v.start;
    o.stop[row] = v.stop;
    o.ref[row] = h.ref0;
    o.alt[row] = h.alt0;
    o.chrom_abs[row] = ls;
    o.chrom_len[row] = (uint8_t)(h.chrom_len > 255 ? 255 : h.chrom_len);
    {   // S5 field of the 35-byte record: first 5 CHROM bytes, NUL padded (silent truncation, vcf_to_h5.py:120)
        uint64_t c5 = 0;
        for (uint32_t k = 0; k < 5 && k < h.chrom_len; ++k) c5 |= (uint64_t)text[ls + k] << (8 * k);
        o.chrom5[row] = c5;
    }
    RowInfo ri;
    ri.samp_abs = h.samp_abs;
    ri.samp_len = h.has_samples ? (uint32_t)(le - h.samp_abs) : 0u;
    ri.cp_row = 0;
    ri.pad = 0;
    const uint32_t gi = h.g < 0 ? 255u : (uint32_t)(h.g > 254 ? 254 : h.g);
    ri.misc = gi | (v.uniform ? kRowUniform : 0u) | (h.has_samples ? kRowHasSamples : 0u);
    if (want_gt) {
        if (!h.has_samples || h.g < 0) atomicAdd(&o.st->n_nogt, 1ull);
        else if (!v.uniform) {
            const unsigned long long slot = atomicAdd(&o.st->n_nonuniform, 1ull);
            if (cp_row_by_line != kNoCpRow) ri.cp_row = cp_row_by_line;
            else { ri.cp_row = (uint32_t)slot; if (o.nu_rows) o.nu_rows[slot] = (uint32_t)row; }
        }
    }
    o.rowinfo[row] = ri;
}

}  // namespace hb
