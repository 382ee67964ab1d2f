// hb_tokenize.cu -- kernel 1: newline / tab tokenisation of decompressed VCF body text.
//
// Replaces the byte-at-a-time line fetch + field split that htslib does inside
// BcfReader::getNextVariant (reference cpp/vcfpp.h:1455-1484: tbx_itr_next / vcf_parse1).
//
// The text is cut into one CONTIGUOUS byte range per persistent CTA (2 CTAs per SM).  Each CTA
// streams its range in 32 KB tiles through a 3-stage shared-memory ring filled by TMA bulk
// copies (cp.async.bulk + mbarrier); every thread owns a contiguous 128-byte span, classifies its
// bytes with SWAR masks, and the (lines, tabs-since-last-newline) pair is scanned thread -> warp
// -> CTA with popc + shuffles; the running pair is carried from tile to tile inside the CTA, so
// there is NO inter-CTA dependency and the text is read from HBM exactly once.
// Outputs, staged per CTA (the site kernel turns them into global record indices with a prefix
// sum over the per-CTA newline counts):
//   nl_after[b][j]   offset of the byte after the j-th newline of CTA b's range (record starts)
//   cp[(b,jl)][c]    "tabs" mode only: offset of the TAB before sample c*kCP of the line that
//                    starts after newline jl-1 of CTA b (column checkpoints, which let the GT
//                    decoder cut 2-D tiles out of variable-width text)
// The piece of a line that spills into the next CTA's range (one per CTA) gets its checkpoints
// from a tiny fix-up kernel once the carried-in column is known.
#include "hb_common.cuh"
#include "hb_internal.h"

namespace hb {

constexpr int TK_THREADS = 256;
constexpr int TK_SPAN = 128;                       // bytes per thread
constexpr int TK_TILE = TK_THREADS * TK_SPAN;      // 32 KB
constexpr int TK_STAGES = 3;
constexpr int TK_WARPS = TK_THREADS / 32;

uint64_t tokenize_tile_bytes() { return TK_TILE; }

struct LT {            // scan element: lines, tabs since the last newline
    uint32_t lines, tabs;
};
__device__ __forceinline__ LT lt_combine(LT a, LT b) {   // a then b
    LT r;
    r.lines = a.lines + b.lines;
    r.tabs = b.lines ? b.tabs : a.tabs + b.tabs;
    return r;
}

struct TkSmem {
    alignas(128) uint8_t stage[TK_STAGES][TK_TILE];
    alignas(8) uint64_t full[TK_STAGES];
    LT warp_agg[TK_WARPS];
};


__device__ __forceinline__ uint32_t tabs4(const uint4 &v) {
    return __popc(eq_mask(v.x, kTab4)) + __popc(eq_mask(v.y, kTab4)) + __popc(eq_mask(v.z, kTab4)) +
           __popc(eq_mask(v.w, kTab4));
}

// span byte position of the k-th (1-based) TAB of a 128-byte span; pre8 byte r = tabs in chunks 0..r
__device__ __forceinline__ uint32_t locate_tab(const uint8_t *span, uint64_t pre8, uint32_t k) {
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 7; ++j) r += (((uint32_t)(pre8 >> (8 * j)) & 0xffu) < k) ? 1u : 0u;
    uint32_t kk = k - (r ? (uint32_t)(pre8 >> (8 * (r - 1))) & 0xffu : 0u);
    const uint4 v = *reinterpret_cast<const uint4 *>(span + 16 * r);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t m = eq_mask(w[q], kTab4);
        const uint32_t c = __popc(m);
        if (kk <= c) return 16u * r + 4u * q + (uint32_t)nth_flag(m, (int)kk - 1);
        kk -= c;
    }
    return TK_SPAN - 1;
}

// word `wi` (0..31) of the calling thread's span, bytes at/after `valid` (tile-relative) zeroed
__device__ __forceinline__ uint32_t span_word(const uint8_t *tile, uint32_t span_off, int wi, uint32_t valid) {
    uint32_t off = span_off + 4u * wi;
    if (off >= valid) return 0;
    uint32_t w = *reinterpret_cast<const uint32_t *>(tile + off);
    if (off + 4 > valid) w &= (1u << (8 * (valid - off))) - 1u;
    return w;
}

template <bool kTabs>
__global__ void __launch_bounds__(TK_THREADS, 2)
tokenize_kernel(const uint8_t *__restrict__ text, uint64_t nbytes, uint64_t tiles_per_cta,
                uint64_t *__restrict__ nl_after, uint32_t stage_cap, uint64_t *__restrict__ cp, uint32_t ncp,
                CtaTok *__restrict__ cta_out) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TkSmem &sm = *reinterpret_cast<TkSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t b = blockIdx.x;
    const uint64_t range_begin = b * tiles_per_cta * (uint64_t)TK_TILE;
    const uint64_t range_end = min(nbytes, (b + 1) * tiles_per_cta * (uint64_t)TK_TILE);
    const uint64_t n_my_tiles = range_begin < range_end ? (range_end - range_begin + TK_TILE - 1) / TK_TILE : 0;
    uint64_t *my_nl = nl_after + b * (uint64_t)stage_cap;

    if (tid == 0) {
        for (int s = 0; s < TK_STAGES; ++s) mbar_init(&sm.full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](uint64_t it) {
        if (it >= n_my_tiles) return;
        uint64_t off = range_begin + it * (uint64_t)TK_TILE;
        uint64_t rem = range_end - off;
        uint32_t bytes = rem >= TK_TILE ? TK_TILE : (uint32_t)((rem + 15) & ~15ull);
        int s = (int)(it % TK_STAGES);
        mbar_expect_tx(&sm.full[s], bytes);
        tma_load_1d(sm.stage[s], text + off, bytes, &sm.full[s]);
    };
    if (tid == 0)
        for (int s = 0; s < TK_STAGES; ++s) issue(s);

    LT carry;           // CTA-local running state, identical in every thread
    carry.lines = 0;
    carry.tabs = 0;

    for (uint64_t it = 0; it < n_my_tiles; ++it) {
        const int s = (int)(it % TK_STAGES);
        const uint64_t tile_off = range_begin + it * (uint64_t)TK_TILE;
        const uint64_t rem = range_end - tile_off;
        const uint32_t valid = rem >= TK_TILE ? TK_TILE : (uint32_t)rem;
        const uint8_t *buf = sm.stage[s];
        mbar_wait(&sm.full[s], (uint32_t)((it / TK_STAGES) & 1));

        // ---- per-thread counts over its 128-byte span (rotated 16-byte chunks: conflict-free LDS.128)
        const uint32_t span_off = (uint32_t)tid * TK_SPAN;
        const uint8_t *span = buf + span_off;
        const bool full_tile = valid == TK_TILE;
        uint32_t n_tab = 0, nlchunks = 0;
        uint64_t packed = 0;             // byte r = number of tabs in logical chunk r
        if (full_tile) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = (lane + i) & 7;
                const uint4 v = *reinterpret_cast<const uint4 *>(span + 16 * r);
                if (has_byte(v.x, kNl4) | has_byte(v.y, kNl4) | has_byte(v.z, kNl4) | has_byte(v.w, kNl4))
                    nlchunks |= 1u << r;
                if (kTabs) {
                    const uint32_t c = tabs4(v);
                    n_tab += c;
                    packed |= (uint64_t)c << (8 * r);
                }
            }
        }
        const uint64_t pre8 = packed * 0x0101010101010101ull;   // byte r = tabs in chunks 0..r
        // Spans without a newline (almost all) are done.  Exactly one newline: located from its chunk.
        // Anything else (several newlines, the partial last tile): plain word walk.
        LT mine;
        mine.lines = 0;
        mine.tabs = n_tab;
        bool generic = !full_tile;
        uint32_t pn = 0, tA = 0;         // span position of the single newline, tabs before it
        if (full_tile && nlchunks) {
            if (nlchunks & (nlchunks - 1)) generic = true;
            else {
                const int r = __ffs(nlchunks) - 1;
                const uint4 v = *reinterpret_cast<const uint4 *>(span + 16 * r);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                uint32_t tot = 0, before = 0;
                int q1 = 0;
                uint32_t m1 = 0, w1 = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t mn = eq_mask(w[q], kNl4);
                    if (mn && !tot) { q1 = q; m1 = mn; w1 = w[q]; }
                    if (!tot && !mn && kTabs) before += __popc(eq_mask(w[q], kTab4));
                    tot += __popc(mn);
                }
                if (tot != 1) generic = tot != 0;      // tot == 0 cannot happen (has_byte never misses)
                else {
                    const int byte = (__ffs(m1) - 1) >> 3;
                    pn = 16u * r + 4u * q1 + byte;
                    mine.lines = 1;
                    if (kTabs) {
                        before += __popc(eq_mask(w1, kTab4) & ((1u << (8 * byte)) - 1u));
                        tA = (r ? (uint32_t)(pre8 >> (8 * (r - 1))) & 0xffu : 0u) + before;
                        mine.tabs = n_tab - tA;
                    }
                }
            }
        }
        if (generic) {
            uint32_t lines = 0, tabs = 0;
            n_tab = 0;
            for (int wi = 0; wi < TK_SPAN / 4; ++wi) {
                uint32_t w = span_word(buf, span_off, wi, valid);
                uint32_t mn = eq_mask(w, kNl4);
                uint32_t mt = kTabs ? eq_mask(w, kTab4) : 0u;
                n_tab += __popc(mt);
                if (mn) {
                    lines += __popc(mn);
                    uint32_t after = ~((2u << (31 - __clz(mn))) - 1u);   // bytes above the last newline
                    tabs = __popc(mt & after);
                } else {
                    tabs += __popc(mt);
                }
            }
            mine.lines = lines;
            mine.tabs = tabs;
        }

        // ---- warp inclusive scan; warp aggregates through shared memory
        LT inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            LT o;
            o.lines = __shfl_up_sync(0xffffffffu, inc.lines, d);
            o.tabs = __shfl_up_sync(0xffffffffu, inc.tabs, d);
            if (lane >= d) inc = lt_combine(o, inc);
        }
        if (lane == 31) sm.warp_agg[warp] = inc;
        __syncthreads();

        // ---- emission: record starts and column checkpoints (rare per thread)
        LT pre = carry, tile_tot = carry;
#pragma unroll
        for (int w = 0; w < TK_WARPS; ++w) {
            LT a = sm.warp_agg[w];
            if (w < warp) pre = lt_combine(pre, a);
            tile_tot = lt_combine(tile_tot, a);
        }
        {
            LT ex;
            ex.lines = __shfl_up_sync(0xffffffffu, inc.lines, 1);
            ex.tabs = __shfl_up_sync(0xffffffffu, inc.tabs, 1);
            if (lane == 0) { ex.lines = 0; ex.tabs = 0; }
            pre = lt_combine(pre, ex);
            uint32_t L = pre.lines;          // newlines of this CTA before this span
            uint32_t Ccol = pre.tabs;
            const uint64_t gbase = tile_off + span_off;
            // the head fragment of ranges b > 0 (L == 0) gets its checkpoints from the fix-up kernel
            if (!generic) {
                const uint32_t nA = mine.lines ? tA : n_tab;     // tabs of the segment that continues line L
                if (kTabs && nA && (L > 0 || b == 0)) {
                    // first checkpoint tab number T* >= max(9, Ccol+1) with (T*-9) % kCP == 0
                    const uint32_t lo = Ccol + 1;
                    const uint32_t tstar = lo <= 9 ? 9 : 9 + ((lo - 9 + kCP - 1) / kCP) * kCP;
                    if (tstar <= Ccol + nA) {
                        const uint32_t j = (tstar - 9) / kCP;
                        if (j < ncp && L < stage_cap)
                            cp[(b * stage_cap + L) * ncp + j] = gbase + locate_tab(span, pre8, tstar - Ccol);
                    }
                }
                if (mine.lines) {
                    if (L < stage_cap) my_nl[L] = gbase + pn + 1;
                    if (kTabs && n_tab - tA >= 9 && ncp && L + 1 < stage_cap)     // TAB before sample 0 of the new line
                        cp[(b * stage_cap + L + 1) * ncp] = gbase + locate_tab(span, pre8, tA + 9);
                }
            } else if (mine.lines || (kTabs && n_tab)) {
                for (int wi = 0; wi < TK_SPAN / 4; ++wi) {
                    uint32_t w = span_word(buf, span_off, wi, valid);
                    uint32_t mn = eq_mask(w, kNl4);
                    uint32_t mt = kTabs ? eq_mask(w, kTab4) : 0u;
                    uint32_t m = mn | mt;
                    while (m) {
                        int bit = __ffs(m) - 1;
                        m &= m - 1;
                        uint64_t off = gbase + 4u * wi + (bit >> 3);
                        if (mn & (1u << bit)) {
                            if (L < stage_cap) my_nl[L] = off + 1;
                            ++L;
                            Ccol = 0;
                        } else {
                            ++Ccol;
                            if ((L > 0 || b == 0) && Ccol >= 9 && ((Ccol - 9) % kCP) == 0) {
                                uint32_t j = (Ccol - 9) / kCP;
                                if (j < ncp && L < stage_cap) cp[(b * stage_cap + L) * ncp + j] = off;
                            }
                        }
                    }
                }
            }
        }
        carry = tile_tot;
        __syncthreads();   // everyone is done with stage s and with warp_agg
        if (tid == 0) issue(it + TK_STAGES);
    }
    if (tid == 0) {
        CtaTok o;
        o.n_newlines = carry.lines;
        o.tail_tabs = carry.tabs;
        cta_out[b] = o;
    }
}

// ------------------------------------------------------------------------------------------
// Fix-up ("tabs" mode): checkpoints of the line fragment at the head of each CTA range b >= 1.
// One warp per range: the carried-in column comes from the previous ranges' tails, the owner row
// is the last line started in the nearest previous range that holds a newline.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_fragment_kernel(const uint8_t *__restrict__ text, uint64_t nbytes, uint64_t range_bytes, uint32_t n_cta,
                     const uint64_t *__restrict__ nl_after, uint32_t stage_cap, const CtaTok *__restrict__ cta,
                     uint64_t *__restrict__ cp, uint32_t ncp) {
    const int lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b == 0 || b >= n_cta) return;
    const uint64_t beg = (uint64_t)b * range_bytes;
    if (beg >= nbytes) return;
    uint64_t carry = 0, owner = 0;
    for (int c = (int)b - 1; c >= 0; --c) {
        CtaTok t = cta[c];
        carry += t.tail_tabs;
        if (t.n_newlines) { owner = (uint64_t)c * stage_cap + t.n_newlines; break; }
    }
    const CtaTok me = cta[b];
    const uint64_t end = me.n_newlines ? nl_after[(uint64_t)b * stage_cap] - 1 : min(nbytes, beg + range_bytes);
    uint64_t *out = cp + owner * ncp;
    uint64_t seen = carry;     // tabs of this line before `beg`
    for (uint64_t base = beg & ~15ull; base < end; base += 512) {
        uint64_t o = base + 16ull * lane;
        uint32_t w[4] = {0, 0, 0, 0};
        if (o < end) {
            uint4 v = *reinterpret_cast<const uint4 *>(text + o);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        }
        uint32_t m[4], cnt = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            m[q] = eq_mask(w[q], kTab4);
            uint64_t wo = o + 4ull * q;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (wo + k < beg || wo + k >= end) m[q] &= ~(0x80u << (8 * k));
            cnt += __popc(m[q]);
        }
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        uint64_t T = seen + inc - cnt;     // tabs before this lane's first tab
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t mm = m[q];
            while (mm) {
                int bit = __ffs(mm) - 1;
                mm &= mm - 1;
                ++T;
                if (T >= 9 && ((T - 9) % kCP) == 0 && (T - 9) / kCP < ncp) out[(T - 9) / kCP] = o + 4ull * q + (bit >> 3);
            }
        }
        seen += __shfl_sync(0xffffffffu, inc, 31);
    }
}

void launch_tokenize(bool with_tabs, const uint8_t *d_text, uint64_t nbytes, uint32_t n_cta, uint64_t tiles_per_cta,
                     uint64_t *d_nl_after, uint32_t stage_cap, uint64_t *d_cp, uint32_t ncp, CtaTok *d_cta,
                     const Launch &L) {
    size_t smem = sizeof(TkSmem);
    if (!n_cta) return;
    // > 48 KB of dynamic shared memory is opt-in per function AND per device: set on every launch (a microsecond), so a
    // process that moves to another device is served too
    if (with_tabs) cudaFuncSetAttribute(tokenize_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else cudaFuncSetAttribute(tokenize_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (with_tabs)
        tokenize_kernel<true><<<n_cta, TK_THREADS, smem, L.stream>>>(d_text, nbytes, tiles_per_cta, d_nl_after,
                                                                     stage_cap, d_cp, ncp, d_cta);
    else
        tokenize_kernel<false><<<n_cta, TK_THREADS, smem, L.stream>>>(d_text, nbytes, tiles_per_cta, d_nl_after,
                                                                      stage_cap, d_cp, ncp, d_cta);
    count_launch();
    if (with_tabs && n_cta > 1) {
        head_fragment_kernel<<<(n_cta + 7) / 8, 256, 0, L.stream>>>(d_text, nbytes, tiles_per_cta * TK_TILE, n_cta,
                                                                    d_nl_after, stage_cap, d_cta, d_cp, ncp);
        count_launch();
    }
}

// ------------------------------------------------------------------------------------------
// Column index for the (few) non-uniform records when the newline-only tokenizer was used:
// one warp per record walks the sample region and writes the same checkpoints.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
index_columns_kernel(const uint8_t *__restrict__ text, const RowInfo *__restrict__ rowinfo,
                     const uint32_t *__restrict__ nu_rows, uint64_t n_nu, uint64_t *__restrict__ cp, uint32_t ncp) {
    const int lane = threadIdx.x & 31;
    uint64_t wid = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= n_nu) return;
    const RowInfo ri = rowinfo[nu_rows[wid]];
    const uint64_t b = ri.samp_abs, e = ri.samp_abs + ri.samp_len;
    uint64_t *out = cp + (uint64_t)ri.cp_row * ncp;
    uint32_t seen = 0;   // tabs seen so far; tab ordinal k (0-based) precedes sample k
    for (uint64_t base = b & ~15ull; base < e; base += 512) {
        uint64_t o = base + 16ull * lane;
        uint32_t w[4] = {0, 0, 0, 0};
        if (o < e) {
            uint4 v = *reinterpret_cast<const uint4 *>(text + o);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        }
        uint32_t m[4], cnt = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            m[q] = eq_mask(w[q], kTab4);
            // keep only bytes inside [b, e)
            uint64_t wo = o + 4ull * q;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint64_t p = wo + k;
                if (p < b || p >= e) m[q] &= ~(0x80u << (8 * k));
            }
            cnt += __popc(m[q]);
        }
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        uint32_t k = seen + inc - cnt;   // ordinal of this lane's first tab
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t mm = m[q];
            while (mm) {
                int bit = __ffs(mm) - 1;
                mm &= mm - 1;
                if ((k % kCP) == 0 && k / kCP < ncp) out[k / kCP] = o + 4ull * q + (bit >> 3);
                ++k;
            }
        }
        seen += __shfl_sync(0xffffffffu, inc, 31);
    }
}

void launch_index_columns(const uint8_t *d_text, const RowInfo *d_rowinfo, const uint32_t *d_nu_rows,
                          uint64_t n_nu, uint64_t *d_cp, uint32_t ncp, const Launch &L) {
    if (!n_nu) return;
    uint64_t blocks = (n_nu + 7) / 8;
    index_columns_kernel<<<(unsigned)blocks, 256, 0, L.stream>>>(d_text, d_rowinfo, d_nu_rows, n_nu, d_cp, ncp);
    count_launch();
}

}  // namespace hb
