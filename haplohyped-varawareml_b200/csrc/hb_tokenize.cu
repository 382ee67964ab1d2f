// hb_tokenize.cu -- kernel 1: newline / tab tokenisation of decompressed VCF body text.
//
// Replaces the byte-at-a-time line fetch + field split that htslib does inside
// BcfReader::getNextVariant (reference cpp/vcfpp.h:1455-1484: tbx_itr_next / vcf_parse1).
//
// One persistent CTA per SM slot streams 32 KB tiles of text through a 3-stage shared-memory
// ring filled by TMA bulk copies (cp.async.bulk + mbarrier).  Every thread owns a contiguous
// 128-byte span, classifies its bytes with SWAR masks, and the (lines, tabs-since-last-newline)
// pair is scanned thread -> warp -> CTA -> device with a single-pass decoupled look-back, so the
// text is read from HBM exactly once.  Outputs: line_start[] (record offsets) and, in "tabs"
// mode, a checkpoint (byte offset of the TAB before sample j*kCP) for every kCP-th sample column
// of every line, which is what lets the GT decoder cut 2-D tiles out of variable-width text.
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
__device__ __forceinline__ uint64_t lt_pack(LT v) { return ((uint64_t)v.lines << 32) | v.tabs; }
__device__ __forceinline__ LT lt_unpack(uint64_t w) {
    LT v;
    v.lines = (uint32_t)((w & kPayload) >> 32);
    v.tabs = (uint32_t)w;
    return v;
}

struct TkSmem {
    alignas(128) uint8_t stage[TK_STAGES][TK_TILE];
    alignas(8) uint64_t full[TK_STAGES];
    LT warp_agg[TK_WARPS];
    LT tile_excl;
};

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
tokenize_kernel(const uint8_t *__restrict__ text, uint64_t nbytes, uint64_t n_tiles,
                uint64_t *__restrict__ tile_state, uint64_t *__restrict__ line_start, uint64_t line_cap,
                uint64_t *__restrict__ cp, uint32_t ncp, DevStatus *__restrict__ st) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TkSmem &sm = *reinterpret_cast<TkSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t G = gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < TK_STAGES; ++s) mbar_init(&sm.full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](uint64_t it) {
        uint64_t tile = blockIdx.x + it * G;
        if (tile >= n_tiles) return;
        uint64_t off = tile * (uint64_t)TK_TILE;
        uint64_t rem = nbytes - off;
        uint32_t bytes = rem >= TK_TILE ? TK_TILE : (uint32_t)((rem + 15) & ~15ull);
        int s = (int)(it % TK_STAGES);
        mbar_expect_tx(&sm.full[s], bytes);
        tma_load_1d(sm.stage[s], text + off, bytes, &sm.full[s]);
    };
    if (tid == 0)
        for (int s = 0; s < TK_STAGES; ++s) issue(s);

    for (uint64_t it = 0;; ++it) {
        const uint64_t tile = blockIdx.x + it * G;
        if (tile >= n_tiles) break;
        const int s = (int)(it % TK_STAGES);
        const uint64_t tile_off = tile * (uint64_t)TK_TILE;
        const uint64_t rem = nbytes - tile_off;
        const uint32_t valid = rem >= TK_TILE ? TK_TILE : (uint32_t)rem;
        const uint8_t *buf = sm.stage[s];
        mbar_wait(&sm.full[s], (uint32_t)((it / TK_STAGES) & 1));

        // ---- per-thread counts over its 128-byte span (rotated 16-byte chunks: conflict-free LDS.128)
        const uint32_t span_off = (uint32_t)tid * TK_SPAN;
        uint32_t n_nl = 0, n_tab = 0;
        if (valid == TK_TILE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int r = (lane + i) & 7;
                uint4 v = *reinterpret_cast<const uint4 *>(buf + span_off + 16 * r);
                if (kTabs) {
                    n_tab += __popc(eq_mask(v.x, kTab4)) + __popc(eq_mask(v.y, kTab4)) +
                             __popc(eq_mask(v.z, kTab4)) + __popc(eq_mask(v.w, kTab4));
                    n_nl += __popc(eq_mask(v.x, kNl4)) + __popc(eq_mask(v.y, kNl4)) +
                            __popc(eq_mask(v.z, kNl4)) + __popc(eq_mask(v.w, kNl4));
                } else {
                    n_nl |= has_byte(v.x, kNl4) | has_byte(v.y, kNl4) | has_byte(v.z, kNl4) | has_byte(v.w, kNl4);
                }
            }
        } else {
            for (int wi = 0; wi < TK_SPAN / 4; ++wi) {
                uint32_t w = span_word(buf, span_off, wi, valid);
                if (kTabs) n_tab += __popc(eq_mask(w, kTab4));
                n_nl += __popc(eq_mask(w, kNl4));
            }
        }
        // exact (lines, tabs after last newline) for the rare spans that hold a newline
        LT mine;
        mine.lines = 0;
        mine.tabs = n_tab;
        if (n_nl) {
            uint32_t lines = 0, tabs = 0;
            for (int wi = 0; wi < TK_SPAN / 4; ++wi) {
                uint32_t w = span_word(buf, span_off, wi, valid);
                uint32_t mn = eq_mask(w, kNl4);
                uint32_t mt = kTabs ? eq_mask(w, kTab4) : 0u;
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

        // ---- warp inclusive scan, CTA scan
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
        if (warp == 0) {
            LT agg = sm.warp_agg[0];
#pragma unroll
            for (int w = 1; w < TK_WARPS; ++w) agg = lt_combine(agg, sm.warp_agg[w]);
            // ---- decoupled look-back across tiles (tile i depends on tiles < i only)
            LT excl;
            excl.lines = 0;
            excl.tabs = 0;
            if (tile == 0) {
                if (lane == 0) st_relaxed(&tile_state[0], kFlagPre | lt_pack(agg));
            } else {
                if (lane == 0) st_relaxed(&tile_state[tile], kFlagAgg | lt_pack(agg));
                int64_t base = (int64_t)tile - 1;
                for (;;) {
                    int64_t idx = base - lane;
                    uint64_t w = kFlagPre;   // identity prefix for idx < 0
                    if (idx >= 0) {
                        do { w = ld_relaxed(&tile_state[idx]); } while ((w >> 62) == 0);
                    }
                    unsigned pre = __ballot_sync(0xffffffffu, (w >> 62) == 2);
                    int P = pre ? (__ffs(pre) - 1) : 31;
                    LT v = lt_unpack(w);
                    if (lane > P) { v.lines = 0; v.tabs = 0; }
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        LT o;
                        o.lines = __shfl_down_sync(0xffffffffu, v.lines, d);
                        o.tabs = __shfl_down_sync(0xffffffffu, v.tabs, d);
                        if (lane + d < 32) v = lt_combine(o, v);   // farther tiles come first
                    }
                    LT win;
                    win.lines = __shfl_sync(0xffffffffu, v.lines, 0);
                    win.tabs = __shfl_sync(0xffffffffu, v.tabs, 0);
                    excl = lt_combine(win, excl);
                    if (pre) break;
                    base -= 32;
                }
                if (lane == 0) st_relaxed(&tile_state[tile], kFlagPre | lt_pack(lt_combine(excl, agg)));
            }
            if (lane == 0) {
                sm.tile_excl = excl;
                if (tile == n_tiles - 1) {
                    LT tot = lt_combine(excl, agg);
                    st->n_lines = tot.lines;
                    if ((uint64_t)tot.lines + 1 > line_cap) st->line_overflow = 1;
                }
            }
        }
        __syncthreads();

        // ---- emission: line starts and column checkpoints (rare per thread)
        {
            // exclusive prefix of this thread = tile_excl . warps before . lanes before
            LT pre = sm.tile_excl;
            for (int w = 0; w < warp; ++w) pre = lt_combine(pre, sm.warp_agg[w]);
            LT ex;
            ex.lines = __shfl_up_sync(0xffffffffu, inc.lines, 1);
            ex.tabs = __shfl_up_sync(0xffffffffu, inc.tabs, 1);
            if (lane == 0) { ex.lines = 0; ex.tabs = 0; }
            pre = lt_combine(pre, ex);
            uint64_t L = pre.lines;
            uint32_t Ccol = pre.tabs;
            bool walk = n_nl != 0;
            if (kTabs && !walk && n_tab) {
                // first checkpoint tab number T* >= max(9, Ccol+1) with (T*-9) % kCP == 0
                uint32_t lo = Ccol + 1;
                uint32_t tstar = lo <= 9 ? 9 : 9 + ((lo - 9 + kCP - 1) / kCP) * kCP;
                walk = tstar <= Ccol + n_tab;
            }
            if (tile == 0 && tid == 0 && line_cap) line_start[0] = 0;
            if (walk) {
                const uint64_t gbase = tile_off + span_off;
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
                            ++L;
                            Ccol = 0;
                            if (L < line_cap) line_start[L] = off + 1;
                        } else {
                            ++Ccol;
                            if (Ccol >= 9 && ((Ccol - 9) % kCP) == 0) {
                                uint32_t j = (Ccol - 9) / kCP;
                                if (j < ncp && L + 1 < line_cap) cp[L * ncp + j] = off;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();   // everyone is done with stage s and with warp_agg / tile_excl
        if (tid == 0) issue(it + TK_STAGES);
    }
}

void launch_tokenize(bool with_tabs, const uint8_t *d_text, uint64_t nbytes, uint64_t *d_tile_state,
                     uint64_t n_tiles, uint64_t *d_line_start, uint64_t line_cap, uint64_t *d_cp, uint32_t ncp,
                     DevStatus *d_st, const Launch &L) {
    static bool attr_set = false;
    size_t smem = sizeof(TkSmem);
    if (!attr_set) {
        cudaFuncSetAttribute(tokenize_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(tokenize_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    cudaMemsetAsync(d_tile_state, 0, n_tiles * sizeof(uint64_t), L.stream);
    // all CTAs must be co-resident (look-back waits on lower tiles): 2 per SM by launch bounds + smem
    uint64_t grid = (uint64_t)L.sm_count * 2;
    if (grid > n_tiles) grid = n_tiles;
    if (grid == 0) return;
    if (with_tabs)
        tokenize_kernel<true><<<(unsigned)grid, TK_THREADS, smem, L.stream>>>(d_text, nbytes, n_tiles, d_tile_state,
                                                                               d_line_start, line_cap, d_cp, ncp, d_st);
    else
        tokenize_kernel<false><<<(unsigned)grid, TK_THREADS, smem, L.stream>>>(d_text, nbytes, n_tiles, d_tile_state,
                                                                                d_line_start, line_cap, d_cp, ncp, d_st);
    count_launch();
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
