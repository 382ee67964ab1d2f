// hb_store.cu -- kernel 4: Blosc2 byte-shuffle + LZ4 block encoder + Blosc2 chunk / cframe framing.
//
// Replaces what h5py + hdf5plugin do for every HDF5 chunk of
//   create_dataset('snp_data', data=<35-byte records>, compression=32001,
//                  compression_opts=(2,2,0,0,5,1,2), chunks=True)        (vcf_to_h5.py:119-135)
// i.e. c-blosc2's shuffle(typesize 35) + LZ4-family codec, wrapped by the hdf5-blosc2 filter as a
// contiguous frame holding one Blosc2 chunk.  The reference asks for LZ4HC (compcode 2); LZ4 and
// LZ4HC share one block format and one Blosc codec-format id, so a stock decoder cannot tell.
//
// The byte-shuffled image of one (sample, chunk) is 35 planes of `cr` bytes.  Planes 0..32 hold
// site bytes and are IDENTICAL for every sample; only planes 33/34 (the two allele planes, which
// the GT decoder already wrote in exactly this planar layout) differ.  So:
//   site_prefix_kernel   one CTA per chunk: builds the 33 site planes in shared memory straight
//                        from the SoA columns (the 35-byte AoS records are never materialised) and
//                        LZ4-encodes them once, as the head of a no-split LZ4 block;
//   donor_frames_kernel  one warp per (chunk, sample): LZ4-encodes the 2*cr allele bytes as the
//                        continuation of that block (history = tail of the site planes), then
//                        writes cframe header + chunk header + shared prefix + own sequences +
//                        offsets chunk + trailer.
// The encoder is a warp-cooperative greedy matcher: 32 candidate positions per step are hashed in
// parallel, the first hit is extended with ballots, literals are copied 32 bytes per step.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/haplo_b200.h"
#include "hb_common.cuh"
#include "hb_internal.h"
#include "hb_parse_struct.h"

namespace hb {

constexpr int FRAME_HDR = 97, CHUNK_HDR = 32, OFFS_CHUNK = 40, FRAME_TRAILER = 35;

__device__ __forceinline__ uint32_t rd32(const uint8_t *p) {     // unaligned 4-byte read (shared memory)
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    return __funnelshift_r(w[0], w[1], (uint32_t)(a & 3) * 8u);
}

// Warp-cooperative LZ4 encoder of one SEGMENT src[0..n) of a block whose earlier bytes (`hist` of them) sit right
// in front of src.  32 positions are examined per step; each lane tests three candidates for its position -- `far`
// bytes back (0 = none), the last position with the same 4-byte hash (own segment only), one byte back -- the first
// lane with a hit wins and the match is extended with ballots.  Only complete sequences are emitted and no match
// crosses the end of the segment.  The first sequence assumes that its literal run starts at src[0]: its header
// (token + literal-length bytes) occupies dst[0..*first_hdr) and announces *first_lit literals, so that a caller
// which carries literals in from the previous segment can re-write just that header.  *pending = trailing bytes of
// the segment not covered by a sequence.  deep: the hash table holds buckets of 4 (the last position of each p mod 4)
// and the candidate with the longest match over the next 8 bytes is taken -- about 6 % smaller output on random
// A/C/G/T planes for about twice the time.  table: 1 << hashlog uint16 entries (position + 1), cleared here;
// n < 65535.  Returns the bytes written to dst; all lanes return the same values.
// The site planes pass far = 4*cr: the stop column is the start column + 1, so bytes 1..3 of `stop` (planes 10-12)
// repeat bytes 1..3 of `start` (planes 6-8) almost everywhere.
__device__ int warp_lz4_segment(const uint8_t *src, int n, int hist, uint8_t *dst, uint16_t *table, int hashlog,
                                int far, bool deep, int *first_lit, int *first_hdr, int *pending) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < (1 << hashlog); i += 32) table[i] = 0;
    __syncwarp();
    int op = 0, anchor = 0, pos = 0;
    const int mflimit = n - 4;                       // last position a match may start at
    *first_lit = 0; *first_hdr = 0;
    while (pos <= mflimit) {
        const int p = pos + lane;
        const bool valid = p <= mflimit;
        uint32_t v = 0, h = 0;
        int cand = -0x40000000;
        if (valid) {
            v = rd32(src + p);
            h = deep ? ((v * 2654435761u) >> (34 - hashlog)) << 2      // bucket of 4: the last position of each p mod 4
                     : (v * 2654435761u) >> (32 - hashlog);
            if (far && p + hist >= far && rd32(src + p - far) == v) cand = p - far;
            else if (!deep) {
                const int c = (int)table[h] - 1;
                if (c >= 0 && c < p && rd32(src + c) == v) cand = c;   // c == p: left by a re-examined window
                else if (p + hist >= 1 && rd32(src + p - 1) == v) cand = p - 1;      // run of one byte
            } else {
                // longest of the bucket's candidates, measured over the next 8 bytes (ties: the nearest)
                const uint2 b4 = *reinterpret_cast<const uint2 *>(table + h);
                const int cs[4] = {(int)(b4.x & 0xffffu) - 1, (int)(b4.x >> 16) - 1, (int)(b4.y & 0xffffu) - 1, (int)(b4.y >> 16) - 1};
                int best = -1;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = cs[k];
                    if (c >= 0 && c < p && rd32(src + c) == v) {                 // c == p: left by a re-examined window
                        const uint32_t x1 = rd32(src + p + 4) ^ rd32(src + c + 4);
                        int ex;
                        if (x1) ex = (__ffs(x1) - 1) >> 3;
                        else { const uint32_t x2 = rd32(src + p + 8) ^ rd32(src + c + 8); ex = 4 + (x2 ? (__ffs(x2) - 1) >> 3 : 4); }
                        const int key = (ex << 16) | c;                          // longer first, then nearer (larger c)
                        if (key > best) best = key;
                    }
                }
                if (best >= 0) cand = best & 0xffff;
                else if (p + hist >= 1 && rd32(src + p - 1) == v) cand = p - 1;  // run of one byte
            }
        }
        __syncwarp();
        if (valid) table[deep ? h + (p & 3) : h] = (uint16_t)(p + 1);
        __syncwarp();
        const unsigned hit = __ballot_sync(0xffffffffu, cand > -0x40000000);
        if (!hit) { pos += 32; continue; }
        const int f = __ffs(hit) - 1;
        const int m = pos + f;
        const int c = __shfl_sync(0xffffffffu, cand, f);
        int ml = 4;
        for (;;) {
            const int i = ml + lane;
            const bool same = (m + i < n) && src[m + i] == src[c + i];
            const unsigned bal = __ballot_sync(0xffffffffu, same);
            if (bal == 0xffffffffu) { ml += 32; continue; }
            ml += __ffs(~bal) - 1;
            break;
        }
        const int litlen = m - anchor;
        int o = op;
        if (lane == 0) dst[o] = (uint8_t)((min(litlen, 15) << 4) | min(ml - 4, 15));
        ++o;
        if (litlen >= 15) {
            int rem = litlen - 15;
            while (rem >= 255) { if (lane == 0) dst[o] = 255; ++o; rem -= 255; }
            if (lane == 0) dst[o] = (uint8_t)rem;
            ++o;
        }
        if (op == 0) { *first_lit = litlen; *first_hdr = o; }
        for (int i = lane; i < litlen; i += 32) dst[o + i] = src[anchor + i];
        o += litlen;
        const int off = m - c;
        if (lane == 0) { dst[o] = (uint8_t)off; dst[o + 1] = (uint8_t)(off >> 8); }
        o += 2;
        if (ml - 4 >= 15) {
            int rem = ml - 19;
            while (rem >= 255) { if (lane == 0) dst[o] = 255; ++o; rem -= 255; }
            if (lane == 0) dst[o] = (uint8_t)rem;
            ++o;
        }
        op = o;
        anchor = pos = m + ml;
    }
    *pending = n - anchor;
    __syncwarp();
    return op;
}

// ------------------------------------------------------------------------------------------
// Frame anatomy.  One HDF5 chunk of one donor = one Blosc2 contiguous frame holding one chunk:
//   [0,97)     cframe header        frame_len @16 (BE64) and cbytes @39 (BE64) depend on the donor
//   [97,129)   Blosc2 chunk header  cbytes @109 (LE32) depends on the donor
//   [129,133)  bstarts[0] = 36
//   [133,137)  csize of the single (no-split) stream, LE32: depends on the donor
//   [137,137+plen)   LZ4 sequences of the 33 site planes    -- identical for every donor
//   [.., +dlen)      LZ4 sequences of the 2 allele planes   -- the donor's own
//   [.., +40)        offsets chunk (one int64 0, memcpyed)  -- constant
//   [.., +35)        cframe trailer                          -- constant
// Kernels:
//   site_template_kernel   one CTA per chunk: site planes -> LZ4; writes the frame TEMPLATE (header with
//                          the donor-dependent fields left zero + shared LZ4 head), 16-byte aligned.
//   donor_frames_kernel    one warp per (sample, chunk) frame, fused: allele planes -> LZ4 tail of the
//                          block (bit-parallel matcher, below) -> template (from L2) + own tail (from
//                          shared memory) written straight to the frame's slot with 16-byte vector
//                          stores, the four size fields patched in registers.  C_out is written exactly
//                          once and nothing else of size leaves the SM; warps never wait on each other.
// Output: ONE buffer of slots in [sample][chunk] order.  The slot of (sample s, chunk c) starts at
// s * row_stride + slot_off[c]; slot_off is the running sum of round16(template_len[c] + worst-case
// allele tail), so every frame's address is known before it is encoded (no scan, no look-back) and the
// layout is deterministic.  A frame fills the front of its slot; its true length is recorded.
// ------------------------------------------------------------------------------------------
constexpr int TMPL_HDR = FRAME_HDR + CHUNK_HDR + 8;        // 137 bytes of a frame precede its LZ4 block
constexpr int FRAME_TAIL = OFFS_CHUNK + FRAME_TRAILER;     // 75 bytes follow it

__device__ __align__(16) const uint8_t kFrameTail[FRAME_TAIL + 1] = {      // + 1: read as 19 words by the lane encoder
    // offsets chunk: Blosc2 chunk header (version 5, LZ4 format 1, flags memcpyed|shuffle|bitshuffle(=extended),
    // typesize 8, nbytes 8, blocksize 8, cbytes 40, filters[5] = shuffle) + one int64 0
    5, 1, 0x17, 8, 8, 0, 0, 0, 8, 0, 0, 0, OFFS_CHUNK, 0, 0, 0,
    0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0, 0, 0, 0, 0, 0, 0, 0,
    // trailer: [version 1, vlmetalayers {index, map16 0, array16 0}, uint32 trailer_len, fixext16 fingerprint]
    0x94, 0x01, 0x93, 0xcd, 0, 5, 0xde, 0, 0, 0xdc, 0, 0, 0xce, 0, 0, 0, FRAME_TRAILER, 0xd8, 0,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

__device__ __forceinline__ void put_be(uint8_t *p, uint64_t v, int nb) {
    for (int i = 0; i < nb; ++i) p[i] = (uint8_t)(v >> (8 * (nb - 1 - i)));
}
__device__ __forceinline__ void put_le32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

// the 137 header bytes of a frame for chunks of `nbytes` uncompressed bytes; donor-dependent fields are zero
__device__ void write_frame_head(uint8_t *h, uint32_t nbytes) {
    for (int i = 0; i < TMPL_HDR; ++i) h[i] = 0;
    // ---- cframe header (c-blosc2 README_CFRAME_FORMAT; msgpack, big-endian)
    h[0] = 0x9e; h[1] = 0xa8;
    const char magic[8] = {'b', '2', 'f', 'r', 'a', 'm', 'e', 0};
    for (int i = 0; i < 8; ++i) h[2 + i] = (uint8_t)magic[i];
    h[10] = 0xd2; put_be(h + 11, FRAME_HDR, 4);
    h[15] = 0xcf;                                            // frame_len: patched
    h[24] = 0xa4; h[25] = 0x12; h[26] = 0x00; h[27] = 0x51; h[28] = 0x03;   // v2 | 64-bit offs, contiguous, LZ4 | clevel 5, split mode
    h[29] = 0xd3; put_be(h + 30, nbytes, 8);
    h[38] = 0xd3;                                            // cbytes: patched
    h[47] = 0xd2; put_be(h + 48, 35, 4);
    h[52] = 0xd2; put_be(h + 53, nbytes, 4);
    h[57] = 0xd2; put_be(h + 58, nbytes, 4);
    h[62] = 0xd1; put_be(h + 63, 1, 2);
    h[65] = 0xd1; put_be(h + 66, 1, 2);
    h[68] = 0xc2;
    h[69] = 0xd8; h[70] = 6;
    h[76] = 1;                                               // filters[5] = BLOSC_SHUFFLE
    h[87] = 0x93; h[88] = 0xcd; put_be(h + 89, 5, 2);
    h[91] = 0xde; h[94] = 0xdc;
    // ---- Blosc2 chunk header (extended, 32 bytes, little-endian)
    uint8_t *k = h + FRAME_HDR;
    k[0] = 5; k[1] = 1; k[2] = 0x35; k[3] = 35;              // format 5, LZ4 format 1, shuffle|bitshuffle(=extended)|dont-split|LZ4
    put_le32(k + 4, nbytes); put_le32(k + 8, nbytes);        // cbytes @12: patched
    k[21] = 1;                                               // filters[5] = BLOSC_SHUFFLE
    put_le32(k + 32, CHUNK_HDR + 4);                         // bstarts[0]
}                                                            // csize @36: patched

// one LZ4 sequence, written by the whole warp; returns the new output offset
__device__ int warp_emit_seq(uint8_t *dst, int o, const uint8_t *lit, int litlen, int off, int ml) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) dst[o] = (uint8_t)((min(litlen, 15) << 4) | min(ml - 4, 15));
    ++o;
    if (litlen >= 15) {
        int rem = litlen - 15;
        while (rem >= 255) { if (lane == 0) dst[o] = 255; ++o; rem -= 255; }
        if (lane == 0) dst[o] = (uint8_t)rem;
        ++o;
    }
    for (int i = lane; i < litlen; i += 32) dst[o + i] = lit[i];
    o += litlen;
    if (lane == 0) { dst[o] = (uint8_t)off; dst[o + 1] = (uint8_t)(off >> 8); }
    o += 2;
    if (ml - 4 >= 15) {
        int rem = ml - 19;
        while (rem >= 255) { if (lane == 0) dst[o] = 255; ++o; rem -= 255; }
        if (lane == 0) dst[o] = (uint8_t)rem;
        ++o;
    }
    return o;
}

struct SiteArgs4 {
    const uint64_t *chrom5;      // per record: first 5 CHROM bytes, NUL padded, in the low 40 bits
    const uint32_t *start, *stop;
    const uint8_t *ref, *alt;
    uint64_t n_records;
    uint32_t cr;                 // records per chunk
    uint8_t *tmpl;               // [n_chunks][tmpl_cap], 16-byte aligned rows
    uint32_t tmpl_cap;
    uint32_t *tmpl_len;          // [n_chunks] = TMPL_HDR + LZ4 bytes of the site planes
    int deep;                    // 4-way bucket matcher (HB_SITE_MATCHER=deep): smaller templates, slower kernel
};

constexpr int kSiteSegs = 8;                     // = warps of the CTA
constexpr int kSiteHashLog = 10;
__host__ __device__ inline uint32_t site_seg_cap(uint32_t len) { return (len + len / 255 + 24 + 15) & ~15u; }
// Segment s of the encoded part [0, 24*cr+1) of the site planes covers [site_seg_begin(s), site_seg_begin(s+1)).
// The cuts follow where the work is: the REF plane (13) and the ALT plane (23) are random A/C/G/T -- hundreds of
// short matches each -- and get three warps each; everything else (constant, zero, or literal-only planes) is cheap.
__host__ __device__ inline uint32_t site_seg_begin(int s, uint32_t cr) {
    switch (s) {
        case 0: return 0;
        case 1: return 9 * cr;
        case 2: return 13 * cr;
        case 3: return 13 * cr + cr / 3;
        case 4: return 13 * cr + 2 * cr / 3;
        case 5: return 23 * cr;
        case 6: return 23 * cr + cr / 3;
        case 7: return 23 * cr + 2 * cr / 3;
        default: return 24 * cr + 1;
    }
}

__global__ void __launch_bounds__(kSiteSegs * 32) site_template_kernel(const SiteArgs4 a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_len[kSiteSegs], s_pend[kSiteSegs], s_flit[kSiteSegs], s_fhdr[kSiteSegs];
    __shared__ int s_dst[kSiteSegs], s_carry[kSiteSegs], s_final[2];
    __shared__ uint8_t s_head[TMPL_HDR + 7];
    const uint32_t cr = a.cr;
    const uint32_t n = 33u * cr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *planes = smem + 16;                            // 16 zero bytes of history in front
    uint8_t *outs = planes + ((n + 19) & ~15u);             // per-segment output regions
    const uint64_t c = blockIdx.x;
    const uint64_t r0 = c * cr;
    for (uint32_t i = threadIdx.x; i < cr; i += blockDim.x) {
        const uint64_t r = r0 + i;
        uint64_t ch = 0;
        uint32_t st = 0, sp = 0;
        uint8_t rf = 0, al = 0;
        if (r < a.n_records) { ch = a.chrom5[r]; st = a.start[r]; sp = a.stop[r]; rf = a.ref[r]; al = a.alt[r]; }
#pragma unroll
        for (int k = 0; k < 5; ++k) planes[k * cr + i] = (uint8_t)(ch >> (8 * k));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            planes[(5 + k) * cr + i] = (uint8_t)(st >> (8 * k));
            planes[(9 + k) * cr + i] = (uint8_t)(sp >> (8 * k));
        }
        planes[13 * cr + i] = rf;
        planes[23 * cr + i] = al;
#pragma unroll
        for (int k = 0; k < 9; ++k) { planes[(14 + k) * cr + i] = 0; planes[(24 + k) * cr + i] = 0; }
    }
    if (threadIdx.x < 16) smem[threadIdx.x] = 0;
    if (threadIdx.x == 0) write_frame_head(s_head, 35u * cr);
    __syncthreads();
    uint8_t *g = a.tmpl + c * a.tmpl_cap;
    // Chunks of fewer than 6 records cannot honour LZ4's end-of-block rules (last match >= 12 bytes
    // before the end) once the 2*cr allele bytes follow: such blocks are stored raw (csize == size).
    if (cr < 6) {
        for (uint32_t i = threadIdx.x; i < TMPL_HDR; i += blockDim.x) g[i] = s_head[i];
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) g[TMPL_HDR + i] = planes[i];
        const uint32_t tl = TMPL_HDR + n;
        for (uint32_t i = tl + threadIdx.x; i < ((tl + 15) & ~15u); i += blockDim.x) g[i] = 0;
        if (threadIdx.x == 0) a.tmpl_len[c] = tl;
        return;
    }
    // Planes 24..32 (ALT bytes 1..9) are all zero: the head is encoded up to the first of those zeros, then ONE
    // offset-1 match covers the other 9*cr-1 -- so the site part always ends on a sequence boundary and every donor
    // continues the block with nothing pending and zeros behind it.  The head is cut into kSiteSegs segments, one
    // per warp, encoded independently (own hash table; the fixed-distance candidates reach back across segments).
    const int n1 = 24 * (int)cr + 1;
    const int sb = (int)site_seg_begin(warp, cr), se = (int)site_seg_begin(warp + 1, cr);
    uint32_t out_off = 0;
    for (int w = 0; w < warp; ++w) out_off += site_seg_cap(site_seg_begin(w + 1, cr) - site_seg_begin(w, cr));
    uint8_t *my_out = outs + out_off;
    uint16_t *table = reinterpret_cast<uint16_t *>(outs + ((site_seg_cap((uint32_t)n1) + kSiteSegs * 48 + 15) & ~15u)) + ((size_t)warp << kSiteHashLog);
    {
        int flit, fhdr, pend;
        const int len = warp_lz4_segment(planes + sb, se - sb, sb, my_out, table, kSiteHashLog, 4 * (int)cr, a.deep != 0, &flit, &fhdr, &pend);
        if (lane == 0) { s_len[warp] = len; s_pend[warp] = pend; s_flit[warp] = flit; s_fhdr[warp] = fhdr; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                       // stitch: literals left over by a segment open the next one's first sequence
        int carry = 0, off = 0;
        for (int w = 0; w < kSiteSegs; ++w) {
            s_dst[w] = off; s_carry[w] = carry;
            if (s_len[w] == 0) { carry += s_pend[w]; continue; }
            const int lit = s_flit[w] + carry;
            off += 1 + (lit >= 15 ? 1 + (lit - 15) / 255 : 0) + carry + (s_len[w] - s_fhdr[w]);
            carry = s_pend[w];
        }
        s_final[0] = off; s_final[1] = carry;
    }
    __syncthreads();
    uint8_t *lz = g + TMPL_HDR;
    if (s_len[warp] > 0) {
        const int carry = s_carry[warp], lit = s_flit[warp] + carry;
        int o = s_dst[warp];
        if (lane == 0) lz[o] = (uint8_t)((min(lit, 15) << 4) | (my_out[0] & 15));
        ++o;
        if (lit >= 15) {
            int rem = lit - 15;
            while (rem >= 255) { if (lane == 0) lz[o] = 255; ++o; rem -= 255; }
            if (lane == 0) lz[o] = (uint8_t)rem;
            ++o;
        }
        for (int i = lane; i < carry; i += 32) lz[o + i] = planes[sb - carry + i];
        o += carry;
        const int rest = s_len[warp] - s_fhdr[warp];
        for (int i = lane; i < rest; i += 32) lz[o + i] = my_out[s_fhdr[warp] + i];
    }
    int o = s_final[0];
    if (warp == 0) o = warp_emit_seq(lz, o, planes + n1 - s_final[1], s_final[1], 1, (int)n - n1);
    else o = o + 1 + (s_final[1] >= 15 ? 1 + (s_final[1] - 15) / 255 : 0) + s_final[1] + 2 + ((int)n - n1 >= 19 ? 1 + ((int)n - n1 - 19) / 255 : 0);
    const uint32_t tl = TMPL_HDR + (uint32_t)o;
    for (uint32_t i = threadIdx.x; i < TMPL_HDR; i += blockDim.x) g[i] = s_head[i];
    for (uint32_t i = tl + threadIdx.x; i < ((tl + 15) & ~15u); i += blockDim.x) g[i] = 0;
    if (threadIdx.x == 0) a.tmpl_len[c] = tl;
}

// ------------------------------------------------------------------------------------------
// Allele-plane encoder, bit-parallel.  After the SNP filter the allele bytes are 0 / 1 (rarely -9), so
// the two planes are packed to one bit per byte (B = bit 0, N = "any other bit set") and LZ4 matches
// are found with word-wide logic instead of byte-wise hashing.  Two match sources:
//   Z  a zero byte: matches the all-zero site plane two planes back      (offset 2*cr),  Z = ~B & ~N
//   C  plane 1 only: equal to the same record's byte in plane 0           (offset cr),    C = ~(B1^B0) & ~(N1|N0)
// A byte with N set never matches, so the stream is exact for ANY byte values; it just compresses
// best on genotype data.  Runs of >= 5 Z positions become matches first, then runs of >= 4 C positions
// in what is left; everything else is literal.  Lanes 0-15 own 16 segments of plane 0, lanes 16-31 the
// same segments of plane 1 (a match never crosses a segment); trailing literals of a segment are
// carried into the next lane's first sequence; warp scans give every lane its output offset, its first
// global sequence index and literal index; each lane writes the headers of its own sequences, then all
// literals of the frame are copied in one balanced pass (32 equal shares of the literal index space).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_nc(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

constexpr int kMinZ = 5, kMinC = 4;
constexpr int kWpcMax = 12;                   // warps (= frames) per CTA: 4, 8 or 12, whichever fits most warps on an SM

__device__ __forceinline__ uint32_t pack_lsb4(uint32_t w) { return ((w & 0x01010101u) * 0x01020408u) >> 24; }
__device__ __forceinline__ uint32_t pack_nz4(uint32_t w) {       // bit j = byte j has one of bits 1..7 set
    const uint32_t t = w & 0xFEFEFEFEu;
    const uint32_t nz = ((((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u) >> 7;
    return (nz * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t low_mask(int n) { return n <= 0 ? 0u : (n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u)); }
__device__ __forceinline__ int ctz32(uint32_t v) { return __clz(__brev(v)); }       // 32 for 0
__device__ __forceinline__ int lit_ext(int lit) { return lit >= 15 ? 1 + (lit - 15) / 255 : 0; }

// R = the positions of X that lie inside a run of at least MINRUN consecutive ones (multi-word, LSB first)
template <int NW, int MINRUN>
__device__ __forceinline__ void runs_cover(const uint32_t (&X)[NW], uint32_t (&R)[NW]) {
    uint32_t S[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const uint32_t nxt = k + 1 < NW ? X[k + 1] : 0u;
        uint32_t s = X[k];
#pragma unroll
        for (int d = 1; d < MINRUN; ++d) s &= __funnelshift_r(X[k], nxt, d);
        S[k] = s;
    }
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const uint32_t prv = k > 0 ? S[k - 1] : 0u;
        uint32_t r = S[k];
#pragma unroll
        for (int d = 1; d < MINRUN; ++d) r |= __funnelshift_l(prv, S[k], d);
        R[k] = r;
    }
}

struct FusedArgs {
    const int8_t *gt0, *gt1;
    uint64_t gt_stride, n_records;
    uint32_t cr, n_samples, s0;          // samples [s0, s0 + n_samples)
    uint64_t n_chunks;
    const uint8_t *tmpl; uint32_t tmpl_cap; const uint32_t *tmpl_len;
    uint8_t *frames;
    const uint64_t *slot_off;            // [n_chunks + 1] offset of chunk c's slot inside a sample row; [n_chunks] = row stride
    uint32_t *size;                      // [n_samples * n_chunks] true length of each frame
    unsigned long long *totals;          // [0] sum of frame lengths
    uint32_t bww;       // words of one packed bit array (word 0 is a leading zero word)
    uint32_t caps;      // sequence slots per lane
    uint32_t dcap;      // literal-run descriptors per frame (non-empty runs only)
    uint32_t outcap;    // bytes of the frame-tail buffer
    uint32_t pf_dist;   // frames between a warp and the one that will follow it in its SM slot (0 = no L2 prefetch)
    uint32_t pf_dq, pf_dr;   // pf_dist = pf_dq * n_chunks + pf_dr
    uint32_t wpc;       // warps per CTA
    uint32_t warp_smem;
};

template <int NW>
__global__ void __launch_bounds__(kWpcMax * 32) donor_frames_kernel(const FusedArgs A) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n_frames = A.n_chunks * A.n_samples;
    const uint64_t wid = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (wid >= n_frames) return;
    const int cr = (int)A.cr, n = 2 * cr;
    uint8_t *base = smem + (size_t)warp * A.warp_smem;
    uint32_t *bits = reinterpret_cast<uint32_t *>(base);                 // B string, N string (BWW words each)
    const int BWW = (int)A.bww;
    uint16_t *seqs = reinterpret_cast<uint16_t *>(bits + 2 * BWW);       // [caps][32]
    uint32_t *d_sd = reinterpret_cast<uint32_t *>(seqs + A.caps * 32);   // literal runs: source position | destination << 16
    uint16_t *d_cum = reinterpret_cast<uint16_t *>(d_sd + A.dcap);       //               literal index of the run's first byte
    uint8_t *outb = reinterpret_cast<uint8_t *>(d_cum + A.dcap);

    uint32_t tl = 0;
    int dlen = 0, alpha = 0;
    uint64_t c = 0;
    // state of the parse that the emission needs
    int m = 0, carry = 0, a0 = 0, p = 0, first_ml = 0;
    typename std::conditional<(NW <= 4), uint32_t, unsigned long long>::type kindmask = 0;    // <= 8 sequences per word of the segment
    int out_base = 0, run_base = 0, lit_base = 0, total = 0, totrun = 0, totlit = 0, final_lit = 0;
    uint32_t sidx = 0;
    bool any_n = false;                  // some allele of the frame is neither 0 nor 1

    {
        uint32_t s;
        if (n_frames <= 0xFFFFFFFFull) { s = (uint32_t)wid / (uint32_t)A.n_chunks; c = (uint32_t)wid - s * (uint32_t)A.n_chunks; }
        else { s = (uint32_t)(wid / A.n_chunks); c = wid % A.n_chunks; }
        sidx = s;
        tl = A.tmpl_len[c];
        // the frame whose warp will take this warp's place when it retires: pull its two allele-plane slices into L2
        if (A.pf_dist && lane < 2) {
            uint64_t cf = c + A.pf_dr, sf = (uint64_t)s + A.pf_dq;           // frame wid + pf_dist, without a division
            if (cf >= A.n_chunks) { cf -= A.n_chunks; ++sf; }
            if (sf < A.n_samples) {
            const uint64_t rf = cf * (uint64_t)cr;
            const uint64_t off = (uint64_t)(A.s0 + sf) * A.gt_stride + (rf & ~15ull);
            const uint32_t bytes = (uint32_t)((((rf & 15ull) + min((uint64_t)cr, A.n_records - rf) + 15ull) & ~15ull));
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((lane ? A.gt1 : A.gt0) + off), "r"(bytes) : "memory");
            }
        }
        // ---- 1. planes -> packed bits in shared memory (the literals are rebuilt from the bits: keeping the raw bytes
        //         too would cost 2 * cr bytes of shared memory per warp, i.e. a quarter of the resident warps)
        const uint64_t r0 = c * (uint64_t)cr;
        alpha = (int)(r0 & 15);
        const int valid = (int)min((uint64_t)cr, A.n_records - r0);       // rows past n_records read as zero (HDF5 edge chunk)
        const int nvec = (alpha + cr + 15) >> 4;
        // B and N are ONE bit string each over both planes: block position x (0 .. 2 * cr) is bit alpha + x, so the
        // parse and the literal copy walk from plane 0 into plane 1 without a seam
        uint32_t *Bw = bits + 1, *Nw = Bw + BWW;
        for (int i = lane; i < BWW / 2; i += 32) reinterpret_cast<uint4 *>(bits)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        const uint64_t rowbase = (uint64_t)(A.s0 + s) * A.gt_stride + (r0 & ~15ull);
        uint32_t nacc = 0;
        for (int v = lane; v < 2 * nvec; v += 32) {
            const int pl = v >= nvec, k = v - pl * nvec;
            // rows of this vector that belong to the chunk: [lo_i, hi_i); rows before r0 or past n_records read as zero
            const int lo_i = max(0, alpha - 16 * k), hi_i = alpha + valid - 16 * k;
            if (hi_i <= 0) continue;
            const uint4 x = ldg_stream(reinterpret_cast<const uint4 *>((pl ? A.gt1 : A.gt0) + rowbase + 16 * k));
            uint32_t vm16 = 0xFFFFu;                                           // only the chunk's first / last vector is partial
            if (lo_i > 0 || hi_i < 16) vm16 = low_mask(hi_i) & ~low_mask(lo_i) & 0xFFFFu;
            const uint32_t b16 = (pack_lsb4(x.x) | (pack_lsb4(x.y) << 4) | (pack_lsb4(x.z) << 8) | (pack_lsb4(x.w) << 12)) & vm16;
            uint32_t n16 = 0;
            if ((x.x | x.y | x.z | x.w) & 0xFEFEFEFEu)
                n16 = (pack_nz4(x.x) | (pack_nz4(x.y) << 4) | (pack_nz4(x.z) << 8) | (pack_nz4(x.w) << 12)) & vm16;
            nacc |= n16;
            const int bit = pl * cr + 16 * k, w = bit >> 5, sh = bit & 31;      // row i of the vector is bit 16 * k + i (+ cr)
            if (b16) {
                atomicOr(Bw + w, b16 << sh);
                if (sh > 16) atomicOr(Bw + w + 1, b16 >> (32 - sh));
            }
            if (n16) {
                atomicOr(Nw + w, n16 << sh);
                if (sh > 16) atomicOr(Nw + w + 1, n16 >> (32 - sh));
            }
        }
        __syncwarp();
        any_n = __any_sync(0xffffffffu, nacc != 0);
        const uint32_t sh16 = tl & 15u;      // the tail buffer lines up with the template's end modulo 16
        uint8_t *seq = outb + sh16;
        if (lane < 16) outb[lane] = 0;
        __syncwarp();

        if (cr < 6) {                        // raw block, see site_template_kernel
            for (int i = lane; i < n; i += 32) {
                const int x = i < cr ? i : i - cr;
                seq[i] = x < valid ? (uint8_t)(i < cr ? A.gt0 : A.gt1)[(uint64_t)(A.s0 + s) * A.gt_stride + r0 + x] : (uint8_t)0;
            }
            dlen = n;
        } else {
            // ---- 2. per-lane parse of one segment, position-parallel
            p = lane >> 4;
            const int q = lane & 15;
            const int seg = (cr + 15) >> 4;
            a0 = q * seg;
            const int seglen = max(0, min(seg, cr - a0));
            const int mlim = min(seglen, n - 11 - (p * cr + a0));         // the last 11 bytes of the block stay literals
            const int bi = alpha + p * cr + a0, j0 = bi >> 5, shb = bi & 31;
            const int bi0 = alpha + a0, j00 = bi0 >> 5, shb0 = bi0 & 31;     // the same rows in plane 0
            uint32_t Z[NW], C[NW], ZR[NW], CR[NW];
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                const uint32_t bw = __funnelshift_r(Bw[j0 + k], Bw[j0 + k + 1], shb);
                const uint32_t nw = __funnelshift_r(Nw[j0 + k], Nw[j0 + k + 1], shb);
                const uint32_t vm = low_mask(mlim - 32 * k);
                Z[k] = ~(bw | nw) & vm;
                C[k] = 0;
                if (p) {
                    const uint32_t b0w = __funnelshift_r(Bw[j00 + k], Bw[j00 + k + 1], shb0);
                    const uint32_t n0w = __funnelshift_r(Nw[j00 + k], Nw[j00 + k + 1], shb0);
                    C[k] = ~((bw ^ b0w) | nw | n0w) & vm;
                }
            }
            runs_cover<NW, kMinZ>(Z, ZR);
#pragma unroll
            for (int k = 0; k < NW; ++k) C[k] &= ~ZR[k];
            runs_cover<NW, kMinC>(C, CR);
            int prev_end = 0, first_lit = 0, size_rest = 0, lit_rest = 0, ne_rest = 0;
            int n15 = 0, n19 = 0, nnz = 0;
            {
                bool open = false;
                int ost = 0, okind = 0;
#pragma unroll
                for (int k = 0; k < NW; ++k) {
                    const uint32_t zp = k > 0 ? ZR[k - 1] : 0u, zn = k + 1 < NW ? ZR[k + 1] : 0u;
                    const uint32_t cp = k > 0 ? CR[k - 1] : 0u, cn = k + 1 < NW ? CR[k + 1] : 0u;
                    uint32_t st = (ZR[k] & ~__funnelshift_l(zp, ZR[k], 1)) | (CR[k] & ~__funnelshift_l(cp, CR[k], 1));
                    uint32_t en = (ZR[k] & ~__funnelshift_r(ZR[k], zn, 1)) | (CR[k] & ~__funnelshift_r(CR[k], cn, 1));
                    for (;;) {
                        if (!open) {
                            if (!st) break;
                            const int ts = ctz32(st);
                            st &= st - 1;
                            ost = 32 * k + ts; okind = (CR[k] >> ts) & 1u; open = true;
                        }
                        if (!en) break;
                        const int te = ctz32(en);
                        en &= en - 1;
                        const int ml = 32 * k + te - ost + 1, lit = ost - prev_end;
                        seqs[m * 32 + lane] = (uint16_t)((ost << 8) | ml);
                        kindmask |= (decltype(kindmask))okind << m;
                        if (m == 0) { first_lit = lit; first_ml = ml; }
                        n15 += lit >= 15; n19 += ml >= 19; nnz += lit > 0;
                        prev_end = ost + ml;
                        ++m;
                        open = false;
                    }
                }
            }
            if (m > 0) {                     // sizes of the sequences after the first, from totals
                int matched = 0;
#pragma unroll
                for (int k = 0; k < NW; ++k) matched += __popc(ZR[k] | CR[k]);
                lit_rest = prev_end - matched - first_lit;
                size_rest = 3 * (m - 1) + lit_rest + (n15 - (first_lit >= 15)) + (n19 - (first_ml >= 19));
                ne_rest = nnz - (first_lit > 0);
            }
            const int trail = seglen - prev_end;

            // ---- 3. carry trailing literals forward; scans: output offset, literal index, sequence index
            int val = trail, flag = m > 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v2 = __shfl_up_sync(0xffffffffu, val, d), f2 = __shfl_up_sync(0xffffffffu, flag, d);
                if (lane >= d && !flag) { val += v2; flag = f2; }
            }
            carry = __shfl_up_sync(0xffffffffu, val, 1);
            if (lane == 0) carry = 0;
            final_lit = __shfl_sync(0xffffffffu, val, 31);
            const int lit0 = first_lit + carry;
            const int mysize = m > 0 ? 3 + lit0 + lit_ext(lit0) + (first_ml >= 19) + size_rest : 0;
            const int mylit = m > 0 ? lit0 + lit_rest : 0;
            const int myrun = m > 0 ? (lit0 > 0) + ne_rest : 0;
            uint32_t inc1 = (uint32_t)mysize | ((uint32_t)mylit << 16), inc2 = (uint32_t)myrun;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t1 = __shfl_up_sync(0xffffffffu, inc1, d), t2 = __shfl_up_sync(0xffffffffu, inc2, d);
                if (lane >= d) { inc1 += t1; inc2 += t2; }
            }
            const uint32_t tot1 = __shfl_sync(0xffffffffu, inc1, 31);
            total = (int)(tot1 & 0xFFFFu); totlit = (int)(tot1 >> 16);
            totrun = (int)__shfl_sync(0xffffffffu, inc2, 31);
            out_base = (int)(inc1 & 0xFFFFu) - mysize;
            lit_base = (int)(inc1 >> 16) - mylit;
            run_base = (int)inc2 - myrun;
            dlen = total + 1 + lit_ext(final_lit) + final_lit;
        }
    }
    const uint32_t flen = tl + (uint32_t)dlen + FRAME_TAIL;

    // ---- 4. every lane writes the headers of its own sequences, then the literals are copied in 32 equal shares
    if (cr >= 6) {
        uint8_t *seq = outb + (tl & 15u);
        {
            int o = out_base, r = run_base, lc = lit_base;
            int prev_abs = p * cr + a0 - carry;
            for (int j = 0; j < m; ++j) {
                const uint32_t e = seqs[j * 32 + lane];
                const int st = p * cr + a0 + (int)(e >> 8), ml = (int)(e & 255u);
                const int off = ((kindmask >> j) & 1u) ? cr : 2 * cr;
                const int lit = st - prev_abs;
                seq[o++] = (uint8_t)((min(lit, 15) << 4) | min(ml - 4, 15));
                if (lit >= 15) { int rem = lit - 15; while (rem >= 255) { seq[o++] = 255; rem -= 255; } seq[o++] = (uint8_t)rem; }
                // literal run: literal index g of the frame sits at bit g + dG of the B / N strings and goes to byte g + dD
                if (lit > 0) { d_cum[r] = (uint16_t)lc; d_sd[r] = (uint32_t)(alpha + prev_abs - lc) | ((uint32_t)(o - lc) << 16); ++r; }
                lc += lit; o += lit;
                seq[o++] = (uint8_t)off; seq[o++] = (uint8_t)(off >> 8);
                if (ml >= 19) seq[o++] = (uint8_t)(ml - 19);
                prev_abs = st + ml;
            }
        }
        if (lane == 0) {                     // the last sequence of the block: literals only (>= 11 of them)
            int o = total;
            seq[o++] = (uint8_t)(min(final_lit, 15) << 4);
            if (final_lit >= 15) { int rem = final_lit - 15; while (rem >= 255) { seq[o++] = 255; rem -= 255; } seq[o++] = (uint8_t)rem; }
            d_cum[totrun] = (uint16_t)totlit; d_sd[totrun] = (uint32_t)(alpha + n - final_lit - totlit) | ((uint32_t)(o - totlit) << 16);
            d_cum[totrun + 1] = (uint16_t)(totlit + final_lit);
        }
        __syncwarp();
        {
            const int totL = totlit + final_lit;
            const uint64_t grow = (uint64_t)(A.s0 + sidx) * A.gt_stride + c * (uint64_t)cr;
            const int share = (totL + 31) >> 5;
            const int g0 = lane * share;
            int cnt = min(share, totL - g0);
            int lo = 0, hi = totrun + 1;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int)d_cum[mid] <= g0) lo = mid; else hi = mid; }
            if (cnt > 0) {
                // literal bytes come back from the bits (block position x is bit alpha + x of the B / N strings)
                const uint32_t *Bw = bits + 1, *Nw = Bw + BWW;
                const uint32_t *psd = d_sd + lo;
                const uint16_t *pcum = d_cum + lo + 1;
                uint32_t sd = *psd;
                int nextcum = *pcum, dG = (int)(sd & 0xFFFFu), dD = (int)(sd >> 16);
                int g = g0;
                const int gend = g0 + cnt;
                if (!any_n) {                                 // every allele of the frame is 0 or 1: the bit is the byte
                    do {
                        const int G = g + dG;
                        seq[g + dD] = (uint8_t)(__funnelshift_r(Bw[G >> 5], 0u, G) & 1u);
                        ++g;
                        if (g == nextcum) { sd = *++psd; nextcum = *++pcum; dG = (int)(sd & 0xFFFFu); dD = (int)(sd >> 16); }
                    } while (g < gend);
                } else {
                    do {
                        const int G = g + dG;
                        uint32_t v = __funnelshift_r(Bw[G >> 5], 0u, G) & 1u;
                        if (__funnelshift_r(Nw[G >> 5], 0u, G) & 1u) {        // an allele other than 0 / 1: the byte itself
                            const int x = G - alpha;
                            v = (uint8_t)(x < cr ? A.gt0 : A.gt1)[grow + (x < cr ? x : x - cr)];
                        }
                        seq[g + dD] = (uint8_t)v;
                        ++g;
                        if (g == nextcum) { sd = *++psd; nextcum = *++pcum; dG = (int)(sd & 0xFFFFu); dD = (int)(sd >> 16); }
                    } while (g < gend);
                }
            }
        }
    }
    {
        uint8_t *seq = outb + (tl & 15u);
        const int padded = (int)((((tl & 15u) + (uint32_t)dlen + FRAME_TAIL + 15u) & ~15u) - (tl & 15u)) - dlen;   // tail + zero pad
        for (int i = lane; i < padded; i += 32) seq[dlen + i] = i < FRAME_TAIL ? kFrameTail[i] : (uint8_t)0;
    }
    __syncwarp();

    // ---- 5. template (L2) + own tail (shared memory) -> the frame's slot
    const unsigned long long fbase = (unsigned long long)sidx * A.slot_off[A.n_chunks] + A.slot_off[c];
    if (lane == 0) A.size[wid] = flen;
    const uint32_t lz = tl - TMPL_HDR + (uint32_t)dlen, cb = CHUNK_HDR + 8 + lz;
    const uint32_t jb = tl >> 4, nv = (flen + 15) >> 4;
    const uint4 *T = reinterpret_cast<const uint4 *>(A.tmpl + c * A.tmpl_cap);
    const uint4 *G = reinterpret_cast<const uint4 *>(outb);
    uint4 *D = reinterpret_cast<uint4 *>(A.frames + fbase);
#pragma unroll 4
    for (uint32_t j = lane; j < nv; j += 32) {
        uint4 v;
        if (j < jb) v = ldg_nc(T + j);
        else {
            v = G[j - jb];
            if (j == jb && (tl & 15u)) {           // the template's last bytes and the tail's pad are zero where the other has data
                const uint4 t = ldg_nc(T + j);
                v.x |= t.x; v.y |= t.y; v.z |= t.z; v.w |= t.w;
            }
        }
        if (j <= 8) {                              // the four donor-dependent size fields (all zero in the template)
            if (j == 1) v.y |= __byte_perm(flen, 0, 0x0123);                               // frame_len, BE64 @16
            else if (j == 2) { v.z |= (cb >> 24) << 24; v.w |= __byte_perm(cb, 0, 0x0123) >> 8; }   // cbytes, BE64 @39
            else if (j == 6) v.w |= cb << 8;                                               // chunk cbytes, LE32 @109
            else if (j == 7) v.x |= cb >> 24;
            else if (j == 8) { v.y |= lz << 8; v.z |= lz >> 24; }                          // stream csize, LE32 @133
        }
        stg_stream(D + j, v);
    }
}

// ------------------------------------------------------------------------------------------
// Kernel 4b, lane-per-frame split (experimental, HB_DONOR_SPLIT=lane; NOT the default).  The warp-per-frame
// kernel above spends ~2500-3000 warp instructions on a frame that has ~150 LZ4 sequences: scans, ballots and
// divergent per-lane loops keep most lanes idle.  Here a LANE owns a frame and runs a plain sequential encoder, a
// warp owns the 32 frames (one chunk) x (32 consecutive samples):
//   pack_alleles_kernel       allele planes gt[plane][sample][row] -> bit arrays B (byte & 1) and N (any other
//                             bit) laid out [array][row / 32][sample]: the 32 lanes of a warp read the same
//                             row word of 32 neighbouring samples with one 128-byte request.
//   donor_frames_lane_kernel  phase 1, lock-step over the chunk's words: Z / C run covers with a 2-word
//                             look-ahead pipeline in registers, matches appended to a per-lane list in
//                             shared memory; phase 2 (when a list is nearly full, and at the end): every lane
//                             emits its sequences -- token, literals rebuilt from the bits, offset -- through
//                             an 8-byte register accumulator straight into its frame's slot; phase 3: the
//                             chunk's template is read ONCE per warp and stored to the 32 slots with the four
//                             size fields patched.
// Same parse rules as above (zero runs >= 5 -> offset 2*cr, plane-1 == plane-0 runs >= 4 -> offset cr, last
// 11 bytes literal) without the 32 segment breaks: streams are 0.6 % smaller (5.877x vs 5.84x overall).
// Measured (1.1M x 2504, profiles/r01e_lane_split.txt): pack 1.60 ms + encode 11.9 ms = 13.5 ms against 9.19 ms
// for the warp-per-frame kernel.  Phases 1 and 3 are cheap (~400 and ~100 warp instructions per frame) but phase 2
// costs ~1650: the 32 lanes walk 32 different sequence lists in lock step, so every step pays for the longest
// literal run among them (mean 5.7 literals per sequence, maximum over 32 lanes ~25).  The total, 2370 per frame,
// is no better than the 2460 of the warp-per-frame kernel.
// ------------------------------------------------------------------------------------------
struct PackArgs {
    const int8_t *gt0, *gt1;
    uint64_t gt_stride, n_records;
    uint32_t s0, n_samples;
    uint32_t sp;                 // samples per row word of the bit arrays (window rounded up to 32)
    uint32_t rw;                 // row words per array
    uint32_t *bits;              // [4][rw][sp]: B0, N0, B1, N1
};

__device__ __forceinline__ uint32_t pack_lsb16(const uint4 &x) {
    return pack_lsb4(x.x) | (pack_lsb4(x.y) << 4) | (pack_lsb4(x.z) << 8) | (pack_lsb4(x.w) << 12);
}
__device__ __forceinline__ uint32_t pack_nz16(const uint4 &x) {
    if (!((x.x | x.y | x.z | x.w) & 0xFEFEFEFEu)) return 0u;
    return pack_nz4(x.x) | (pack_nz4(x.y) << 4) | (pack_nz4(x.z) << 8) | (pack_nz4(x.w) << 12);
}

constexpr int kPackWords = 4;    // row words (128 rows) per thread

__global__ void __launch_bounds__(256) pack_alleles_kernel(const PackArgs A) {
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    const uint32_t s = blockIdx.x * 32 + lane;
    const uint32_t j0 = (blockIdx.y * 8 + wy) * kPackWords;
    if (j0 >= A.rw) return;
    const bool live = s < A.n_samples;
    const uint64_t row = (uint64_t)(A.s0 + (live ? s : 0)) * A.gt_stride;
    const size_t plane = (size_t)A.rw * A.sp;
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
        const int8_t *g = (pl ? A.gt1 : A.gt0) + row;
        uint4 x[2 * kPackWords];
#pragma unroll
        for (int i = 0; i < 2 * kPackWords; ++i) {
            const uint64_t r = (uint64_t)j0 * 32 + 16 * i;
            x[i] = make_uint4(0, 0, 0, 0);
            if (live && r < A.n_records) x[i] = ldg_stream(reinterpret_cast<const uint4 *>(g + r));
        }
#pragma unroll
        for (int w = 0; w < kPackWords; ++w) {
            if (j0 + w < A.rw) {
                const uint64_t r = (uint64_t)(j0 + w) * 32;
                const uint32_t vm = r >= A.n_records ? 0u : low_mask((int)min((uint64_t)32, A.n_records - r));
                const uint32_t b = pack_lsb16(x[2 * w]) | (pack_lsb16(x[2 * w + 1]) << 16);
                const uint32_t nn = pack_nz16(x[2 * w]) | (pack_nz16(x[2 * w + 1]) << 16);
                uint32_t *o = A.bits + (size_t)(2 * pl) * plane + (size_t)(j0 + w) * A.sp + s;
                o[0] = b & vm;
                o[plane] = nn & vm;
            }
        }
    }
}

struct LaneArgs {
    const uint32_t *bits;
    uint32_t sp, rw;
    const int8_t *gt0, *gt1;             // raw bytes: only read for alleles other than 0 / 1
    uint64_t gt_stride, n_records;
    uint32_t cr, n_samples, s0, groups;  // groups: 32-sample groups of the window
    uint64_t n_chunks;
    const uint8_t *tmpl;
    uint32_t tmpl_cap;
    const uint32_t *tmpl_len;
    uint8_t *frames;
    const uint64_t *slot_off;
    uint32_t *size;
};

constexpr int kLaneWpc = 4;              // warps per CTA
constexpr int kEntCap = 48;              // matches a lane collects before the warp emits

// the 8-byte accumulator a lane writes its frame through
struct LaneOut {
    uint64_t lo;
    int fill;
    uint64_t *p;
    __device__ __forceinline__ void put(uint32_t v, int k) {       // k in 1..4 bytes of v (the others are zero)
        lo |= (uint64_t)v << (8 * fill);
        fill += k;
        if (fill >= 8) {
            *p++ = lo;
            fill -= 8;
            lo = (uint64_t)v >> (8 * (k - fill));
        }
    }
    __device__ __forceinline__ void put_len(int rem) {             // LZ4 length extension bytes
        while (rem >= 255) { put(255u, 1); rem -= 255; }
        put((uint32_t)rem, 1);
    }
};

template <int R>
__device__ __forceinline__ uint32_t run_starts(uint32_t x, uint32_t nxt) {      // bit i: x has R ones from i on
    uint32_t s = x;
#pragma unroll
    for (int d = 1; d < R; ++d) s &= __funnelshift_r(x, nxt, d);
    return s;
}
template <int R>
__device__ __forceinline__ uint32_t run_cover(uint32_t sprev, uint32_t s) {     // positions inside such runs
    uint32_t r = s;
#pragma unroll
    for (int d = 1; d < R; ++d) r |= __funnelshift_l(sprev, s, d);
    return r;
}

__global__ void __launch_bounds__(kLaneWpc * 32) donor_frames_lane_kernel(const LaneArgs A) {
    __shared__ uint32_t s_ent[kLaneWpc][kEntCap * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t wid = (uint64_t)blockIdx.x * kLaneWpc + warp;
    if (wid >= A.n_chunks * A.groups) return;
    const uint32_t c = (uint32_t)(wid / A.groups), g = (uint32_t)(wid - (uint64_t)c * A.groups);
    const uint32_t s = g * 32 + lane;
    const bool live = s < A.n_samples;
    const int cr = (int)A.cr, n = 2 * cr;
    const uint32_t tl = A.tmpl_len[c];
    const uint64_t row_stride = A.slot_off[A.n_chunks], so = A.slot_off[c];
    uint8_t *frame = A.frames + (uint64_t)(live ? s : 0) * row_stride + so;
    const uint64_t r0 = (uint64_t)c * cr;
    const uint32_t j0 = (uint32_t)(r0 >> 5);
    const int sh = (int)(r0 & 31);
    const size_t plane = (size_t)A.rw * A.sp;
    const uint32_t *col = A.bits + (size_t)j0 * A.sp + s;      // this lane's column, first row word of the chunk
    uint32_t *ent = s_ent[warp] + lane;

    LaneOut o;
    o.p = reinterpret_cast<uint64_t *>(frame + (tl & ~7u));
    o.fill = (int)(tl & 7u);
    o.lo = 0;
    if (o.fill) o.lo = *reinterpret_cast<const uint64_t *>(A.tmpl + (size_t)c * A.tmpl_cap + (tl & ~7u)) & ((1ull << (8 * o.fill)) - 1ull);
    uint64_t *const p_first = o.p;
    int lit_start = 0, cnt = 0;

    // literals [a, b) of the block, rebuilt from the bit arrays (raw bytes only where N is set)
    auto put_literals = [&](int a, int b) {
        int x = a;
        while (x < b) {
            const int pl = x >= cr, i = x - pl * cr;
            const int rb = sh + i, bo = rb & 31;
            int take = min(b - x, 32 - bo);
            if (!pl) take = min(take, cr - x);
            const uint32_t *w = col + (size_t)(2 * pl) * plane + (size_t)(rb >> 5) * A.sp;
            const uint32_t m = low_mask(take);
            const uint32_t bw = (__ldg(w) >> bo) & m, nw = (__ldg(w + plane) >> bo) & m;
            for (int q = 0; q < take; q += 4) {
                const int k4 = min(4, take - q);
                uint32_t v = (((bw >> q) & 15u) * 0x00204081u) & 0x01010101u;
                const uint32_t nq = (nw >> q) & 15u;
                if (nq) {
                    const int8_t *raw = (pl ? A.gt1 : A.gt0) + (uint64_t)(A.s0 + s) * A.gt_stride + r0 + i + q;
                    for (int t = 0; t < k4; ++t)
                        if ((nq >> t) & 1u) v = (v & ~(0xFFu << (8 * t))) | ((uint32_t)(uint8_t)raw[t] << (8 * t));
                }
                o.put(v, k4);
            }
            x += take;
        }
    };
    auto emit = [&](uint32_t e) {
        const int ost = (int)(e & 0x1FFFu), ml = (int)((e >> 13) & 0xFFFu);
        const int off = (e >> 31) ? cr : 2 * cr;
        const int lit = ost - lit_start;
        o.put((uint32_t)((min(lit, 15) << 4) | min(ml - 4, 15)), 1);
        if (lit >= 15) o.put_len(lit - 15);
        put_literals(lit_start, ost);
        if (ml < 19) o.put((uint32_t)off, 2);
        else if (ml < 19 + 255) o.put((uint32_t)off | ((uint32_t)(ml - 19) << 16), 3);
        else { o.put((uint32_t)off | 0xFF0000u, 3); o.put_len(ml - 19 - 255); }
        lit_start = ost + ml;
    };
    auto drain = [&]() {
        const int maxc = __reduce_max_sync(0xffffffffu, cnt);
        for (int i = 0; i < maxc; ++i)
            if (i < cnt) emit(ent[i * 32]);
        cnt = 0;
    };

    // ---- 1. matches of both planes, word by word
    for (int p = 0; p < 2; ++p) {
        const int lim = live ? max(0, min(cr, n - 11 - p * cr)) : 0;           // the last 11 bytes of the block stay literals
        const int W = (__reduce_max_sync(0xffffffffu, lim) + 31) >> 5;
        const uint32_t *cb = col + (size_t)(2 * p) * plane;
        // row words: cur = word j0 + t, nxt = word j0 + t + 1 (loaded one step ahead of their use)
        uint32_t rb_ = __ldg(cb), rn_ = __ldg(cb + plane), r0b = 0, r0n = 0;
        uint32_t xb_ = __ldg(cb + A.sp), xn_ = __ldg(cb + A.sp + plane), x0b = 0, x0n = 0;
        if (p) { r0b = __ldg(col); r0n = __ldg(col + plane); x0b = __ldg(col + A.sp); x0n = __ldg(col + A.sp + plane); }
        uint32_t zA = 0, cxA = 0, zsA = 0, zrA = 0, ccA = 0, csA = 0, pz = 0, pc = 0;
        bool open = false;
        int ost = 0;
        uint32_t okind = 0;
        for (int t = 0; t <= W + 2; ++t) {
            // chunk-relative word t of this plane (zero past lim)
            const uint32_t *nx = cb + (size_t)(t + 2) * A.sp;
            const uint32_t ldb = __ldg(nx), ldn = __ldg(nx + plane);
            uint32_t ld0b = 0, ld0n = 0;
            if (p) { const uint32_t *n0 = col + (size_t)(t + 2) * A.sp; ld0b = __ldg(n0); ld0n = __ldg(n0 + plane); }
            const uint32_t vm = low_mask(lim - 32 * t);
            const uint32_t bw = __funnelshift_r(rb_, xb_, sh), nw = __funnelshift_r(rn_, xn_, sh);
            const uint32_t zN = ~(bw | nw) & vm;
            uint32_t cxN = 0;
            if (p) {
                const uint32_t b0w = __funnelshift_r(r0b, x0b, sh), n0w = __funnelshift_r(r0n, x0n, sh);
                cxN = ~((bw ^ b0w) | nw | n0w) & vm;
            }
            rb_ = xb_; rn_ = xn_; r0b = x0b; r0n = x0n;
            xb_ = ldb; xn_ = ldn; x0b = ld0b; x0n = ld0n;
            const uint32_t zsB = run_starts<kMinZ>(zA, zN);           // word t-1
            const uint32_t zrB = run_cover<kMinZ>(zsA, zsB);          // word t-1
            const uint32_t ccB = cxA & ~zrB;                          // word t-1
            const uint32_t csB = run_starts<kMinC>(ccA, ccB);         // word t-2
            const uint32_t crW = run_cover<kMinC>(csA, csB);          // word t-2
            if (t >= 2) {
                if (__any_sync(0xffffffffu, cnt > kEntCap - 9)) drain();
                const int k = t - 2;
                const uint32_t mz = zrA, mc = crW;
                const uint32_t mz1 = (mz << 1) | pz, mc1 = (mc << 1) | pc;
                uint32_t st = (mz & ~mz1) | (mc & ~mc1);
                uint32_t enx = (mz1 & ~mz) | (mc1 & ~mc);
                pz = mz >> 31; pc = mc >> 31;
                const int base = p * cr + 32 * k;
                for (;;) {
                    if (open) {
                        if (!enx) break;
                        const int te = ctz32(enx);
                        enx &= enx - 1;
                        ent[cnt * 32] = (uint32_t)ost | ((uint32_t)(base + te - ost) << 13) | (okind << 31);
                        ++cnt;
                        open = false;
                    }
                    if (!st) break;
                    const int ts = ctz32(st);
                    st &= st - 1;
                    ost = base + ts; okind = (mc >> ts) & 1u; open = true;
                }
            }
            zA = zN; cxA = cxN; zsA = zsB; zrA = zrB; ccA = ccB; csA = csB;
        }
    }
    drain();

    // ---- 2. the last sequence (literals only), the frame's tail, zero pad to 16 bytes
    uint32_t flen = 0;
    int dlen = 0;
    if (live) {
        const int L = n - lit_start;
        o.put((uint32_t)(min(L, 15) << 4), 1);
        if (L >= 15) o.put_len(L - 15);
        put_literals(lit_start, n);
        dlen = (int)(o.p - p_first) * 8 + o.fill - (int)(tl & 7u);
        const uint32_t *tw = reinterpret_cast<const uint32_t *>(kFrameTail);
#pragma unroll 1
        for (int i = 0; i < FRAME_TAIL / 4; ++i) o.put(tw[i], 4);
        o.put(tw[FRAME_TAIL / 4] & 0xFFFFFFu, FRAME_TAIL & 3);
        if (o.fill) *o.p++ = o.lo;
        if ((o.p - reinterpret_cast<uint64_t *>(frame)) & 1) *o.p++ = 0ull;
        flen = tl + (uint32_t)dlen + FRAME_TAIL;
        A.size[(uint64_t)s * A.n_chunks + c] = flen;
    }

    // ---- 3. the template, once per warp, into the 32 slots; the four donor-dependent size fields patched
    const int nact = (int)min(32u, A.n_samples - g * 32);
    const uint32_t ncopy = tl & ~7u, nv = ncopy >> 4;
    const uint4 *T = reinterpret_cast<const uint4 *>(A.tmpl + (size_t)c * A.tmpl_cap);
    uint8_t *slot0 = A.frames + (uint64_t)(g * 32) * row_stride + so;
    {
        // vectors 0..31 (the patched ones are 1, 2, 6, 7, 8)
        const uint32_t j = (uint32_t)lane;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (j < nv) v = ldg_nc(T + j);
        for (int f = 0; f < nact; ++f) {
            const uint32_t fl = __shfl_sync(0xffffffffu, flen, f);
            const uint32_t lz = fl - FRAME_TAIL - TMPL_HDR, cbz = CHUNK_HDR + 8 + lz;
            uint4 x = v;
            if (j == 1) x.y |= __byte_perm(fl, 0, 0x0123);                                   // frame_len, BE64 @16
            else if (j == 2) { x.z |= (cbz >> 24) << 24; x.w |= __byte_perm(cbz, 0, 0x0123) >> 8; }   // cbytes, BE64 @39
            else if (j == 6) x.w |= cbz << 8;                                                // chunk cbytes, LE32 @109
            else if (j == 7) x.x |= cbz >> 24;
            else if (j == 8) { x.y |= lz << 8; x.z |= lz >> 24; }                            // stream csize, LE32 @133
            if (j < nv) stg_stream(reinterpret_cast<uint4 *>(slot0 + (uint64_t)f * row_stride) + j, x);
        }
    }
    for (uint32_t j = 32 + lane; j < nv; j += 32) {
        const uint4 v = ldg_nc(T + j);
        uint8_t *d = slot0 + (size_t)j * 16;
#pragma unroll 4
        for (int f = 0; f < nact; ++f) { stg_stream(reinterpret_cast<uint4 *>(d), v); d += row_stride; }
    }
    if ((ncopy & 8u) && live) *reinterpret_cast<uint64_t *>(frame + (size_t)nv * 16) = *reinterpret_cast<const uint64_t *>(A.tmpl + (size_t)c * A.tmpl_cap + (size_t)nv * 16);
}

__global__ void __launch_bounds__(256) sum_sizes_kernel(const uint32_t *__restrict__ size, uint64_t n, unsigned long long *__restrict__ total) {
    unsigned long long acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc += size[i];
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(total, acc);
}

uint64_t guess_chunk_records(uint64_t n) {      // h5py/_hl/filters.py guess_chunk for shape (n,), 35-byte items
    const double CHUNK_BASE = 16 * 1024, CHUNK_MIN = 8 * 1024, CHUNK_MAX = 1024 * 1024;
    if (n == 0) return 1;
    double chunk = (double)n;
    const double dset_size = chunk * 35.0;
    double target = CHUNK_BASE * std::pow(2.0, std::log10(dset_size / (1024.0 * 1024.0)));
    if (target > CHUNK_MAX) target = CHUNK_MAX;
    else if (target < CHUNK_MIN) target = CHUNK_MIN;
    for (;;) {
        const double bytes = chunk * 35.0;
        if ((bytes < target || std::fabs(bytes - target) / target < 0.5) && bytes < CHUNK_MAX) break;
        if (chunk == 1) break;
        chunk = std::ceil(chunk / 2.0);
    }
    return (uint64_t)chunk;
}

}  // namespace hb

using namespace hb;

struct hb_frames {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t n_records = 0, n_chunks = 0, cr = 0;
    uint64_t chunk_cap = 0;                  // chunks the per-chunk arrays were allocated for
    bool cr_explicit = false;                // chunk_records was given by the caller (it does not follow n_records)
    uint32_t n_samples = 0, tmpl_cap = 0;    // n_samples: samples of the current window
    uint32_t s0 = 0, win_cap = 0;            // window = samples [s0, s0 + n_samples) of the parse; win_cap = allocated for
    size_t smem_site = 0;
    FusedArgs fa;                            // geometry of the fused kernel
    int nw = 1;
    uint64_t n_ctas = 0;
    uint8_t *d_tmpl = nullptr, *d_frames = nullptr;
    uint32_t *d_tmpl_len = nullptr, *d_size = nullptr;
    uint64_t *d_slot_off = nullptr;
    unsigned long long *d_totals = nullptr;
    std::vector<uint64_t> h_slot_off;            // [n_chunks + 1]
    uint64_t frames_cap = 0;
    uint64_t total_bytes = 0, padded_bytes = 0;
    std::vector<uint32_t> h_tmpl_len;
    // host copies of the layout, fetched on demand
    bool layout_valid = false;
    std::vector<uint32_t> h_size;                // [n_samples][n_chunks]
    std::vector<uint8_t> h_row;                  // scratch for hb_frames_fetch_sample
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    float ms_site = 0, ms_frames = 0;
    // the site templates only need the site columns, not the genotype planes: when the frames are attached to
    // their parse (hb_parse_attach_frames) the template kernel runs on `side`, concurrently with the GT decoder
    cudaStream_t side = nullptr;
    cudaEvent_t ev_sites = nullptr, ev_tmpl = nullptr, ev_side0 = nullptr;
    bool early_site = false;                     // the template pass of the current parse run is already in flight
    // lane-per-frame encoder: bit arrays of the window's allele planes
    uint32_t *d_bits = nullptr;
    uint64_t bits_cap = 0;                       // bytes
    cudaEvent_t ev_pack = nullptr;
    float ms_pack = 0;
    bool lane_split = false;                     // the last run used the lane-per-frame kernels
};

template <int NW>
static void launch_donor_frames(const FusedArgs &fa, uint64_t n_ctas, cudaStream_t st) {
    donor_frames_kernel<NW><<<(unsigned)n_ctas, fa.wpc * 32, (size_t)fa.wpc * fa.warp_smem, st>>>(fa);
}
template <int NW>
static cudaError_t attr_donor_frames(size_t smem) {
    return cudaFuncSetAttribute(donor_frames_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// The frame buffer is many GB and cudaMalloc / cudaFree of that size cost ~100 ms (and cudaFree synchronises the device):
// a released buffer is kept, one per device, for the next frames handle of the process (a converter walks 22 chromosomes).
// hb_cache_clear() gives it back.
namespace {
struct BufCache { uint8_t *p = nullptr; uint64_t cap = 0; };
std::mutex g_fb_mu;
BufCache g_fb[64];
}
namespace hb {
void frames_buffer_cache_clear() {
    std::lock_guard<std::mutex> lk(g_fb_mu);
    for (int d = 0; d < 64; ++d)
        if (g_fb[d].p) { cudaSetDevice(d); cudaFree(g_fb[d].p); g_fb[d] = BufCache(); }
}
}

// templates of all chunks.  early: launched from inside run_parse on the side stream, right after the site columns
// were written on the parse's stream; otherwise on the parse's stream itself.
static int frames_site_pass(hb_frames *f, hb_parse *p, bool early) {
    cudaError_t e;
#define CUF(x) do { e = (x); if (e != cudaSuccess) return api_fail(HB_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e) + " at " #x); } while (0)
    SiteArgs4 sa;
    sa.chrom5 = p->d_chrom5; sa.start = p->d_start; sa.stop = p->d_stop; sa.ref = p->d_ref; sa.alt = p->d_alt;
    sa.n_records = f->n_records; sa.cr = (uint32_t)f->cr; sa.tmpl = f->d_tmpl; sa.tmpl_cap = f->tmpl_cap; sa.tmpl_len = f->d_tmpl_len;
    {
        const char *e = getenv("HB_SITE_MATCHER");
        sa.deep = e && !strcmp(e, "deep");
    }
    cudaStream_t st = early ? f->side : f->stream;
    if (early) {
        CUF(cudaEventRecord(f->ev_sites, p->stream));
        CUF(cudaStreamWaitEvent(f->side, f->ev_sites, 0));
    }
    CUF(cudaEventRecord(early ? f->ev_side0 : f->ev[0], st));
    site_template_kernel<<<(unsigned)f->n_chunks, 256, f->smem_site, st>>>(sa);
    count_launch();
    CUF(cudaEventRecord(early ? f->ev_tmpl : f->ev[1], st));
    f->early_site = early;
#undef CUF
    return HB_OK;
}

namespace hb {
// called by run_parse (hb_api.cu) once the site columns of this run are on their way
void frames_early_site_pass(void *frames, hb_parse *p) {
    hb_frames *f = static_cast<hb_frames *>(frames);
    if (!f || !f->n_chunks || !f->n_samples || !f->side) return;
    if (p->h_st.n_records != f->n_records || p->device != f->device) return;     // shape changed: hb_frames_rerun will say so
    frames_site_pass(f, p, true);
}
}  // namespace hb

static int frames_run(hb_frames *f, hb_parse *p) {
    cudaError_t e;
#define CUF(x) do { e = (x); if (e != cudaSuccess) return api_fail(HB_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e) + " at " #x); } while (0)
    CUF(cudaSetDevice(f->device));
    const uint64_t n = f->n_records;
    const uint32_t cr = (uint32_t)f->cr;
    const uint64_t n_frames = f->n_chunks * f->n_samples;
    f->layout_valid = false;
    CUF(cudaMemsetAsync(f->d_totals, 0, 8, f->stream));
    const bool was_early = f->early_site;
    if (was_early) {
        CUF(cudaStreamWaitEvent(f->stream, f->ev_tmpl, 0));       // the templates were made while the decoder ran
        CUF(cudaEventRecord(f->ev[1], f->stream));
    } else {
        int rc = frames_site_pass(f, p, false);
        if (rc != HB_OK) return rc;
    }
    f->early_site = false;
    // the frame buffer is sized from the longest template: frame <= template + worst-case allele tail
    CUF(cudaMemcpyAsync(f->h_tmpl_len.data(), f->d_tmpl_len, f->n_chunks * 4, cudaMemcpyDeviceToHost, f->stream));
    CUF(cudaStreamSynchronize(f->stream));
    CUF(cudaGetLastError());
    // slots: frame <= template + worst-case allele tail, so every address is known before the encode
    uint64_t need = 0;
    {
        const uint64_t tail = 2ull * cr + 2ull * cr / 255 + 24 + FRAME_TAIL;
        uint64_t run = 0;
        for (uint64_t c = 0; c < f->n_chunks; ++c) { f->h_slot_off[c] = run; run += (f->h_tmpl_len[c] + tail + 15) & ~15ull; }
        f->h_slot_off[f->n_chunks] = run;
        need = run * f->n_samples;
    }
    bool regrow = false;
    if (f->frames_cap < need) {
        regrow = f->d_frames != nullptr;
        if (f->d_frames) { cudaFree(f->d_frames); f->d_frames = nullptr; f->frames_cap = 0; }
        if (f->device >= 0 && f->device < 64) {            // a buffer left behind by an earlier handle?
            std::lock_guard<std::mutex> lk(g_fb_mu);
            BufCache &c = g_fb[f->device];
            if (c.p && c.cap >= need) { f->d_frames = c.p; f->frames_cap = c.cap; c = BufCache(); }
        }
    }
    if (f->frames_cap < need) {
        // head-room: a re-run on other data (slab streaming) has slightly different template lengths, and growing
        // means cudaFree + cudaMalloc of many GB (~100 ms)
        const uint64_t cap = regrow ? need + need / 16 + (64ull << 20) : need + need / 64 + (16ull << 20);
        e = cudaMalloc(&f->d_frames, cap);
        if (e != cudaSuccess) { cudaGetLastError(); dev_pool_flush(); e = cudaMalloc(&f->d_frames, cap); }     // idle pooled buffers may hold what is missing
        if (e != cudaSuccess) return api_fail(HB_ERR_MEM, std::string("cudaMalloc of the frame buffer (") + std::to_string(cap) + " bytes): " + cudaGetErrorString(e));
        f->frames_cap = cap;
    }
    CUF(cudaMemcpyAsync(f->d_slot_off, f->h_slot_off.data(), (f->n_chunks + 1) * 8, cudaMemcpyHostToDevice, f->stream));
    FusedArgs fa = f->fa;
    fa.gt0 = p->d_gt[0]; fa.gt1 = p->d_gt[1]; fa.gt_stride = p->gt_stride; fa.n_records = n;
    fa.cr = cr; fa.n_samples = f->n_samples; fa.s0 = f->s0; fa.n_chunks = f->n_chunks;
    fa.tmpl = f->d_tmpl; fa.tmpl_cap = f->tmpl_cap; fa.tmpl_len = f->d_tmpl_len;
    fa.frames = f->d_frames; fa.slot_off = f->d_slot_off; fa.size = f->d_size; fa.totals = f->d_totals;
    {
        int sms = 148, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device);
        // resident warps of this kernel on the whole GPU (shared memory decides): the distance to prefetch at
        const size_t smem_cta = (size_t)fa.wpc * fa.warp_smem + 1024;
        per_sm = (int)std::min<size_t>(32, (227 * 1024) / smem_cta);
        fa.pf_dist = (uint32_t)(sms * std::min<uint32_t>(40, per_sm * fa.wpc));
        if (const char *e = getenv("HB_DF_PREFETCH")) fa.pf_dist = (uint32_t)atoi(e);
        fa.pf_dq = (uint32_t)(fa.pf_dist / f->n_chunks); fa.pf_dr = (uint32_t)(fa.pf_dist % f->n_chunks);
    }
    // HB_DONOR_SPLIT=lane asks for the lane-per-frame kernels (measured slower, see their header); they need chunks
    // that are not tiny: the four size fields must lie inside the template copy
    bool lane_split = false;
    if (const char *e = getenv("HB_DONOR_SPLIT")) lane_split = !strcmp(e, "lane") && cr >= 16;
    for (uint64_t c = 0; c < f->n_chunks && lane_split; ++c) lane_split = f->h_tmpl_len[c] >= 152;
    f->lane_split = lane_split;
    if (lane_split) {
        const uint32_t sp_cap = (f->win_cap + 31) & ~31u, sp = (f->n_samples + 31) & ~31u;
        const uint64_t rw_cap = (f->chunk_cap * (uint64_t)cr + 31) / 32 + 8, rw = (f->n_chunks * (uint64_t)cr + 31) / 32 + 8;
        const uint64_t need_bits = 16ull * rw_cap * sp_cap;
        if (f->bits_cap < need_bits) {
            if (f->d_bits) { dev_pool_free(f->d_bits); f->d_bits = nullptr; f->bits_cap = 0; }
            e = dev_pool_alloc((void **)&f->d_bits, need_bits);
            if (e != cudaSuccess) return api_fail(HB_ERR_MEM, std::string("cudaMalloc of the allele bit arrays (") + std::to_string(need_bits) + " bytes): " + cudaGetErrorString(e));
            f->bits_cap = need_bits;
        }
        PackArgs pa;
        pa.gt0 = fa.gt0; pa.gt1 = fa.gt1; pa.gt_stride = fa.gt_stride; pa.n_records = n;
        pa.s0 = f->s0; pa.n_samples = f->n_samples; pa.sp = sp; pa.rw = (uint32_t)rw; pa.bits = f->d_bits;
        pack_alleles_kernel<<<dim3(sp / 32, (unsigned)((rw + 8 * kPackWords - 1) / (8 * kPackWords))), 256, 0, f->stream>>>(pa);
        CUF(cudaEventRecord(f->ev_pack, f->stream));
        LaneArgs la;
        la.bits = f->d_bits; la.sp = sp; la.rw = (uint32_t)rw;
        la.gt0 = fa.gt0; la.gt1 = fa.gt1; la.gt_stride = fa.gt_stride; la.n_records = n;
        la.cr = cr; la.n_samples = f->n_samples; la.s0 = f->s0; la.groups = sp / 32; la.n_chunks = f->n_chunks;
        la.tmpl = f->d_tmpl; la.tmpl_cap = f->tmpl_cap; la.tmpl_len = f->d_tmpl_len;
        la.frames = f->d_frames; la.slot_off = f->d_slot_off; la.size = f->d_size;
        const uint64_t n_warps = f->n_chunks * la.groups;
        donor_frames_lane_kernel<<<(unsigned)((n_warps + kLaneWpc - 1) / kLaneWpc), kLaneWpc * 32, 0, f->stream>>>(la);
        count_launch(1);
    } else
    switch (f->nw) {
        case 1: launch_donor_frames<1>(fa, f->n_ctas, f->stream); break;
        case 2: launch_donor_frames<2>(fa, f->n_ctas, f->stream); break;
        case 3: launch_donor_frames<3>(fa, f->n_ctas, f->stream); break;
        case 4: launch_donor_frames<4>(fa, f->n_ctas, f->stream); break;
        case 5: launch_donor_frames<5>(fa, f->n_ctas, f->stream); break;
        default: launch_donor_frames<6>(fa, f->n_ctas, f->stream); break;
    }
    sum_sizes_kernel<<<296, 256, 0, f->stream>>>(f->d_size, n_frames, f->d_totals);
    count_launch(2);
    CUF(cudaEventRecord(f->ev[2], f->stream));
    unsigned long long tot = 0;
    CUF(cudaMemcpyAsync(&tot, f->d_totals, 8, cudaMemcpyDeviceToHost, f->stream));
    CUF(cudaStreamSynchronize(f->stream));
    CUF(cudaGetLastError());
    f->padded_bytes = need;
    f->total_bytes = tot;
    if (was_early) cudaEventElapsedTime(&f->ms_site, f->ev_side0, f->ev_tmpl);
    else cudaEventElapsedTime(&f->ms_site, f->ev[0], f->ev[1]);
    cudaEventElapsedTime(&f->ms_frames, f->ev[1], f->ev[2]);
    f->ms_pack = 0;
    if (lane_split) cudaEventElapsedTime(&f->ms_pack, f->ev[1], f->ev_pack);
#undef CUF
    return HB_OK;
}

static int frames_layout(hb_frames *f) {
    if (f->layout_valid) return HB_OK;
    const uint64_t n_frames = f->n_chunks * f->n_samples;
    f->h_size.resize(n_frames);
    if (n_frames) {
        if (cudaSetDevice(f->device) != cudaSuccess) return api_fail(HB_ERR_CUDA, "cudaSetDevice failed");
        cudaError_t e = d2h_copy(f->h_size.data(), f->d_size, n_frames * 4, f->stream);      // pageable destination: staged (hb_api.cu)
        if (e != cudaSuccess) return api_fail(HB_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e));
    }
    f->layout_valid = true;
    return HB_OK;
}

extern "C" {

uint64_t hb_guess_chunk_records(uint64_t n_records) { return guess_chunk_records(n_records); }

void hb_frames_free(hb_frames *f) {
    if (!f) return;
    if (f->side) cudaStreamSynchronize(f->side);
    cudaSetDevice(f->device);
    if (f->d_frames && f->device >= 0 && f->device < 64) {     // keep the big buffer for the next handle (see g_fb)
        std::lock_guard<std::mutex> lk(g_fb_mu);
        BufCache &c = g_fb[f->device];
        if (c.cap < f->frames_cap) { std::swap(c.p, f->d_frames); std::swap(c.cap, f->frames_cap); }
    }
    dev_pool_free(f->d_tmpl); cudaFree(f->d_frames); dev_pool_free(f->d_tmpl_len); dev_pool_free(f->d_size);
    dev_pool_free(f->d_slot_off); dev_pool_free(f->d_totals); dev_pool_free(f->d_bits);
    if (f->ev_pack) cudaEventDestroy(f->ev_pack);
    for (auto &x : f->ev) if (x) cudaEventDestroy(x);
    if (f->ev_sites) cudaEventDestroy(f->ev_sites);
    if (f->ev_tmpl) cudaEventDestroy(f->ev_tmpl);
    if (f->ev_side0) cudaEventDestroy(f->ev_side0);
    if (f->side) cudaStreamDestroy(f->side);
    delete f;
}

int hb_compress_records(hb_parse *p, uint64_t chunk_records, hb_frames **out) {
    if (!p) return api_fail(HB_ERR_ARG, "null argument");
    return hb_compress_sample_range(p, chunk_records, 0, p->n_samples, out);
}

int hb_compress_sample_range(hb_parse *p, uint64_t chunk_records, uint32_t s0, uint32_t ns, hb_frames **out) {
    if (!p || !out) return api_fail(HB_ERR_ARG, "null argument");
    *out = nullptr;
    if ((uint64_t)s0 + ns > p->n_samples) return api_fail(HB_ERR_ARG, "sample window outside the parse");
    if (cudaSetDevice(p->device) != cudaSuccess) return api_fail(HB_ERR_CUDA, "cudaSetDevice failed");
    const uint64_t n = p->h_st.n_records;
    if (!p->d_gt[0] && n) return api_fail(HB_ERR_NOGT, "parse was made without genotypes");
    hb_frames *f = new hb_frames();
    memset(&f->fa, 0, sizeof f->fa);
    f->device = p->device; f->stream = p->stream;
    f->n_records = n; f->n_samples = ns; f->s0 = s0; f->win_cap = ns;
    f->cr = chunk_records ? chunk_records : guess_chunk_records(n);
    f->cr_explicit = chunk_records != 0;
    f->n_chunks = n ? (n + f->cr - 1) / f->cr : 0;
    if (!f->n_chunks || !f->n_samples) { f->n_chunks = n ? f->n_chunks : 0; *out = f; return HB_OK; }
    if (f->cr > 2730) { hb_frames_free(f); return api_fail(HB_ERR_ARG, "chunk_records too large (at most 2730: the site encoder indexes 24*chunk_records+1 positions with 16 bits)"); }
    const uint32_t cr = (uint32_t)f->cr;
    const uint32_t n_site = 33u * cr, n_gt = 2u * cr;
    auto bound = [](uint32_t x) { return x + x / 255 + 64; };
    f->tmpl_cap = (TMPL_HDR + bound(n_site) + 15) & ~15u;
    f->smem_site = 16 + ((n_site + 19) & ~15u) + ((site_seg_cap(24 * cr + 1) + kSiteSegs * 48 + 15) & ~15u) +
                   ((size_t)kSiteSegs << kSiteHashLog) * 2;
    if (f->smem_site > 220 * 1024) { hb_frames_free(f); return api_fail(HB_ERR_ARG, "chunk too large for the site encoder (33*chunk_records must fit shared memory)"); }
    const uint32_t seg = (cr + 15) / 16;
    f->nw = (int)((seg + 31) / 32);
    FusedArgs &fa = f->fa;
    fa.bww = ((2 * cr + 15) / 32 + 14) & ~3u;        // words of the B (and of the N) bit string: 1 pad + both planes + look-ahead
    // most sequences in a segment: Z (>= 5) and C (>= 4) runs alternating without a gap, 2 per 9 positions
    fa.caps = 2 * (seg / 9) + (seg % 9 >= 4 ? 1 : 0) + 1;
    fa.dcap = (cr / 6 + cr / 5 + 2 + 7) & ~7u;           // a literal run + Z run take >= 6 positions (plane 0), + C run >= 5 (plane 1); + the last run
    fa.outcap = (16 + n_gt + n_gt / 255 + 24 + FRAME_TAIL + 15) & ~15u;
    fa.warp_smem = 8 * fa.bww + 64 * fa.caps + 6 * fa.dcap + fa.outcap;   // dcap is a multiple of 8: 16-byte alignment holds
    if (const char *e = getenv("HB_DF_PAD")) fa.warp_smem += (uint32_t)atoi(e) & ~15u;       // experiment: occupancy sensitivity
    if ((size_t)4 * fa.warp_smem > 220 * 1024) { hb_frames_free(f); return api_fail(HB_ERR_ARG, "chunk too large for the allele encoder"); }
    {   // warps per CTA: every CTA costs 1 KB of reserved shared memory on top of its warps' buffers
        uint32_t best = 4, best_warps = 0;
        for (uint32_t w : {4u, 8u, 12u}) {
            const size_t cta = (size_t)w * fa.warp_smem + 1024;
            const uint32_t warps = (uint32_t)std::min<size_t>(40, std::min<size_t>(32, (227 * 1024) / cta) * w);   // 49 registers per thread: 40 warps
            if (warps > best_warps) { best = w; best_warps = warps; }
        }
        if (const char *e = getenv("HB_DF_WPC")) { const int v = atoi(e); if (v >= 1 && v <= kWpcMax && (size_t)v * fa.warp_smem + 1024 <= 227 * 1024) best = (uint32_t)v; }
        fa.wpc = best;
    }
    f->chunk_cap = f->n_chunks + 4;          // a re-run on a slightly longer record set (streaming) still fits
    const uint64_t n_frames = f->chunk_cap * f->n_samples;
    f->n_ctas = (f->n_chunks * f->n_samples + fa.wpc - 1) / fa.wpc;
    f->h_tmpl_len.resize(f->n_chunks);
    f->h_slot_off.resize(f->n_chunks + 1);
    cudaError_t e = cudaSuccess;
    auto ck = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    ck(dev_pool_alloc((void **)&f->d_tmpl, f->chunk_cap * (uint64_t)f->tmpl_cap));       // pooled (hb_api.cu): cudaMalloc / cudaFree are slow here
    ck(dev_pool_alloc((void **)&f->d_tmpl_len, f->chunk_cap * 4));
    ck(dev_pool_alloc((void **)&f->d_size, n_frames * 4));
    ck(dev_pool_alloc((void **)&f->d_slot_off, (f->chunk_cap + 1) * 8));
    ck(dev_pool_alloc((void **)&f->d_totals, 8));
    for (auto &x : f->ev) ck(cudaEventCreate(&x));
    {   // highest priority: its few long CTAs must get SM slots while the decoder's many short ones stream through
        int lo = 0, hi = 0;
        ck(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        ck(cudaStreamCreateWithPriority(&f->side, cudaStreamNonBlocking, hi));
    }
    ck(cudaEventCreateWithFlags(&f->ev_sites, cudaEventDisableTiming));
    ck(cudaEventCreate(&f->ev_tmpl)); ck(cudaEventCreate(&f->ev_side0)); ck(cudaEventCreate(&f->ev_pack));
    // the opt-in shared-memory ceiling is a per-function, process-wide attribute: always raise it to the device
    // maximum, so that concurrent callers (the converter parses two files at once) cannot shrink each other's limit
    int smem_max = 0;
    ck(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, f->device));
    smem_max -= 1024;                            // room for the kernels' few static __shared__ words
    ck(cudaFuncSetAttribute(site_template_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    ck(attr_donor_frames<1>(smem_max)); ck(attr_donor_frames<2>(smem_max)); ck(attr_donor_frames<3>(smem_max));
    ck(attr_donor_frames<4>(smem_max)); ck(attr_donor_frames<5>(smem_max)); ck(attr_donor_frames<6>(smem_max));
    if (e != cudaSuccess) { hb_frames_free(f); return api_fail(HB_ERR_MEM, std::string("CUDA: ") + cudaGetErrorString(e)); }
    int rc = frames_run(f, p);
    if (rc != HB_OK) { hb_frames_free(f); return rc; }
    *out = f;
    return HB_OK;
}

int hb_parse_attach_frames(hb_parse *p, hb_frames *f) {
    if (!p) return api_fail(HB_ERR_ARG, "null handle");
    p->attached_frames = f;
    return HB_OK;
}

int hb_frames_rerun(hb_frames *f, hb_parse *p) {
    if (!f || !p) return api_fail(HB_ERR_ARG, "null argument");
    if ((uint64_t)f->s0 + f->n_samples > p->n_samples || p->device != f->device)
        return api_fail(HB_ERR_ARG, "hb_frames_rerun: the parse no longer has the shape these frames were made for");
    if (p->h_st.n_records != f->n_records) {
        // another record count is fine when the chunk size was fixed by the caller and the chunk arrays are large enough
        const uint64_t nc = p->h_st.n_records ? (p->h_st.n_records + f->cr - 1) / f->cr : 0;
        if (!f->cr_explicit || nc > f->chunk_cap || nc == 0)
            return api_fail(HB_ERR_ARG, "hb_frames_rerun: the parse no longer has the shape these frames were made for");
        f->n_records = p->h_st.n_records;
        f->n_chunks = nc;
        f->n_ctas = (nc * f->n_samples + f->fa.wpc - 1) / f->fa.wpc;
        f->h_tmpl_len.resize(nc);
        f->h_slot_off.resize(nc + 1);
        f->early_site = false;                   // an early template pass (if any) was made for the old shape
    }
    if (!f->n_chunks || !f->n_samples) return HB_OK;
    return frames_run(f, p);
}

int hb_frames_set_window(hb_frames *f, uint32_t s0, uint32_t ns) {
    if (!f) return api_fail(HB_ERR_ARG, "null handle");
    if (ns == 0 || ns > f->win_cap) return api_fail(HB_ERR_ARG, "window larger than the one the frames were made for");
    f->s0 = s0;
    f->n_samples = ns;
    f->n_ctas = (f->n_chunks * (uint64_t)ns + f->fa.wpc - 1) / f->fa.wpc;
    f->layout_valid = false;
    return HB_OK;
}

int hb_frames_get_info(const hb_frames *f, hb_frames_info *info) {
    if (!f || !info) return api_fail(HB_ERR_ARG, "null argument");
    memset(info, 0, sizeof *info);
    info->n_records = f->n_records; info->n_chunks = f->n_chunks; info->chunk_records = f->cr;
    info->n_samples = f->n_samples; info->total_bytes = f->total_bytes;
    info->raw_bytes = 35ull * f->n_records * f->n_samples;
    info->ms_site = f->ms_site; info->ms_frames = f->ms_frames; info->ms_pack = f->ms_pack;
    info->padded_bytes = f->padded_bytes;
    info->d_frames = f->d_frames;
    uint64_t st = 0;
    for (uint32_t t : f->h_tmpl_len) st += t - TMPL_HDR;
    info->site_lz4_bytes = st;
    return HB_OK;
}

int hb_frames_layout(hb_frames *f, uint64_t *offsets, uint32_t *sizes) {
    if (!f) return api_fail(HB_ERR_ARG, "null handle");
    int rc = frames_layout(f);
    if (rc != HB_OK) return rc;
    const uint64_t n_frames = f->n_chunks * f->n_samples, nc = f->n_chunks;
    if (offsets)
        for (uint32_t s = 0; s < f->n_samples; ++s)
            for (uint64_t c = 0; c < nc; ++c) offsets[s * nc + c] = s * f->h_slot_off[nc] + f->h_slot_off[c];
    if (sizes && n_frames) memcpy(sizes, f->h_size.data(), n_frames * 4);
    return HB_OK;
}

int hb_frames_fetch_all(hb_frames *f, uint8_t *buf, uint64_t cap) {
    if (!f || !buf) return api_fail(HB_ERR_ARG, "null argument");
    if (cap < f->padded_bytes) return api_fail(HB_ERR_ARG, "buffer too small");
    if (!f->padded_bytes) return HB_OK;
    if (cudaSetDevice(f->device) != cudaSuccess) return api_fail(HB_ERR_CUDA, "cudaSetDevice failed");
    cudaError_t e = d2h_copy(buf, f->d_frames, f->padded_bytes, f->stream);      // pinned: one copy; pageable: staged
    if (e != cudaSuccess) return api_fail(HB_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e));
    return HB_OK;
}

int hb_frames_fetch_sample(hb_frames *f, uint32_t s, uint64_t *sizes, uint8_t *buf, uint64_t cap, uint64_t *total) {
    if (!f) return api_fail(HB_ERR_ARG, "null handle");
    if (s >= f->n_samples) return api_fail(HB_ERR_SAMPLE, "sample index out of range");
    int rc = frames_layout(f);
    if (rc != HB_OK) return rc;
    const uint64_t nc = f->n_chunks;
    uint64_t tot = 0;
    for (uint64_t c = 0; c < nc; ++c) {
        const uint32_t sz = f->h_size[s * nc + c];
        if (sizes) sizes[c] = sz;
        tot += sz;
    }
    if (total) *total = tot;
    if (!buf || !nc) return HB_OK;
    if (cap < tot) return api_fail(HB_ERR_ARG, "buffer too small");
    // one contiguous D2H of the sample's row of slots, then the frames are packed together on the host
    const uint64_t row0 = s * f->h_slot_off[nc];
    const uint64_t row = f->h_slot_off[nc - 1] + f->h_size[s * nc + nc - 1];
    f->h_row.resize(row);
    if (cudaSetDevice(f->device) != cudaSuccess) return api_fail(HB_ERR_CUDA, "cudaSetDevice failed");
    cudaError_t e = d2h_copy(f->h_row.data(), f->d_frames + row0, row, f->stream);
    if (e != cudaSuccess) return api_fail(HB_ERR_CUDA, std::string("D2H of frames failed: ") + cudaGetErrorString(e));
    uint64_t o = 0;
    for (uint64_t c = 0; c < nc; ++c) {
        const uint32_t sz = f->h_size[s * nc + c];
        memcpy(buf + o, f->h_row.data() + f->h_slot_off[c], sz);
        o += sz;
    }
    return HB_OK;
}

}  // extern "C"

// =============================================================================================
// Read side: Blosc2 cframe -> chunk -> LZ4 -> un-shuffle, one warp per HDF5 chunk.
// Replaces, for VCFH5Reader.fetch_genotypes (src/utils/h5_reader.py:37-41), what h5py + the Blosc2
// filter do when a `snp_data` dataset is read.  Accepts what stock c-blosc2 writes for this path
// too (several blocks per chunk, LZ4/LZ4HC streams, raw streams, zero-run streams, memcpyed chunks);
// anything else sets the frame's status to non-zero.
// =============================================================================================
namespace hb {

__device__ __forceinline__ uint32_t ld_le32(const uint8_t *p) {
    return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint64_t ld_be(const uint8_t *p, int nb) {
    uint64_t v = 0;
    for (int i = 0; i < nb; ++i) v = (v << 8) | p[i];
    return v;
}

// LZ4 block decode by one warp: sequences are walked in lock-step, bytes are copied 32 per step.
// dst is global memory written and re-read by different lanes: reads go through L2 (__ldcg).
__device__ bool warp_lz4_decode(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t cap) {
    const int lane = threadIdx.x & 31;
    uint32_t ip = 0, op = 0;
    if (n == 0) return false;
    for (;;) {
        if (ip >= n) return false;
        const uint32_t tok = src[ip++];
        uint32_t ll = tok >> 4;
        if (ll == 15) { uint32_t b; do { if (ip >= n) return false; b = src[ip++]; ll += b; } while (b == 255); }
        if (ip + ll > n || op + ll > cap) return false;
        for (uint32_t i = lane; i < ll; i += 32) dst[op + i] = src[ip + i];
        ip += ll; op += ll;
        if (ip == n) break;
        if (ip + 2 > n) return false;
        const uint32_t off = src[ip] | ((uint32_t)src[ip + 1] << 8);
        ip += 2;
        if (off == 0 || off > op) return false;
        uint32_t ml = tok & 15;
        if (ml == 15) { uint32_t b; do { if (ip >= n) return false; b = src[ip++]; ml += b; } while (b == 255); }
        ml += 4;
        if (op + ml > cap) return false;
        __syncwarp();
        // periodic copy: dst[op+i] = dst[op-off + i % off] only reads bytes written before this match
        for (uint32_t i = lane; i < ml; i += 32) {
            const uint32_t k = off >= ml ? i : i % off;
            dst[op + i] = __ldcg(dst + op - off + k);
        }
        op += ml;
        __syncwarp();
    }
    __syncwarp();
    return op == cap;
}

__global__ void __launch_bounds__(256)
decode_frames_kernel(const uint8_t *__restrict__ frames, const uint64_t *__restrict__ offsets, uint64_t n_frames,
                     uint32_t chunk_nbytes, uint8_t *__restrict__ tmp, uint8_t *__restrict__ out, int planar,
                     int *__restrict__ status) {
    const int lane = threadIdx.x & 31;
    const uint64_t wid = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= n_frames) return;
    const uint8_t *f = frames + offsets[wid];
    const uint64_t flen = offsets[wid + 1] - offsets[wid];
    uint8_t *t = tmp + wid * (uint64_t)chunk_nbytes;
    uint8_t *o = out + wid * (uint64_t)chunk_nbytes;
    int err = 0;
    // ---- cframe header
    if (flen < 87 + 35 || f[1] != 0xa8 || f[2] != 'b' || f[3] != '2' || f[4] != 'f' || f[10] != 0xd2) err = 1;
    uint64_t hlen = 0, nbytes = 0, cbytes = 0;
    if (!err) {
        hlen = ld_be(f + 11, 4); nbytes = ld_be(f + 30, 8); cbytes = ld_be(f + 39, 8);
        if (hlen < 87 || hlen + cbytes > flen || nbytes != chunk_nbytes) err = 2;
    }
    uint32_t typesize = 1;
    if (!err) {
        const uint8_t *c = f + hlen;                         // the (single) Blosc2 chunk
        const uint32_t flags = c[2];
        typesize = c[3];
        const uint32_t cn = ld_le32(c + 4), bs = ld_le32(c + 8), ccb = ld_le32(c + 12);
        const bool ext = (flags & 1) && (flags & 4);
        const uint32_t hdr = ext ? 32 : 16;
        bool shuffle = !ext && (flags & 1);
        if (ext) for (int i = 0; i < 6; ++i) { if (c[16 + i] == 1) shuffle = true; else if (c[16 + i] != 0) err = 3; }
        if (cn != chunk_nbytes || ccb > cbytes || bs == 0 || typesize == 0) err = 4;
        if (!err && (flags & 2)) {                           // memcpyed
            for (uint32_t i = lane; i < cn; i += 32) t[i] = c[hdr + i];
            shuffle = false;
        } else if (!err) {
            if ((flags >> 5) != 1) err = 5;                  // LZ4 / LZ4HC codec format
            if (ext && ((c[31] >> 4) & 7)) err = 6;          // special chunks
            const bool dont_split = flags & 0x10;
            const uint32_t nblocks = (cn + bs - 1) / bs;
            for (uint32_t b = 0; b < nblocks && !err; ++b) {
                const uint32_t bsize = (b == nblocks - 1 && cn % bs) ? cn % bs : bs;
                const bool leftover = (b == nblocks - 1) && (cn % bs);
                const uint32_t nstreams = (!dont_split && !leftover) ? typesize : 1;
                const uint32_t ne = bsize / nstreams;
                uint32_t ip = ld_le32(c + hdr + 4 * b);
                for (uint32_t s = 0; s < nstreams && !err; ++s) {
                    if (ip + 4 > ccb) { err = 7; break; }
                    const int32_t cs = (int32_t)ld_le32(c + ip);
                    ip += 4;
                    uint8_t *d = t + (uint64_t)b * bs + (uint64_t)s * ne;
                    if (cs == 0) { for (uint32_t i = lane; i < ne; i += 32) d[i] = 0; }
                    else if (cs < 0 || ip + (uint32_t)cs > ccb) err = 8;
                    else if ((uint32_t)cs == ne) { for (uint32_t i = lane; i < ne; i += 32) d[i] = c[ip + i]; ip += cs; }
                    else { if (!warp_lz4_decode(c + ip, (uint32_t)cs, d, ne)) err = 9; ip += cs; }
                    __syncwarp();
                }
            }
            // Blosc shuffles per block; this path only un-shuffles the single-block layout planar -> AoS
            if (!err && shuffle && nblocks != 1 && !planar) {
                // several blocks: un-shuffle each block on its own
                for (uint32_t b = 0; b < nblocks; ++b) {
                    const uint32_t bsize = (b == nblocks - 1 && cn % bs) ? cn % bs : bs;
                    const uint32_t ne = bsize / typesize;
                    const uint8_t *sb = t + (uint64_t)b * bs;
                    uint8_t *ob = o + (uint64_t)b * bs;
                    for (uint32_t i = lane; i < ne * typesize; i += 32) ob[i] = __ldcg(sb + (i % typesize) * ne + i / typesize);
                    for (uint32_t i = ne * typesize + lane; i < bsize; i += 32) ob[i] = __ldcg(sb + i);
                }
                if (lane == 0) status[wid] = 0;
                return;
            }
            if (!err && !shuffle) planar = 1;                // nothing to undo
        }
    }
    __syncwarp();
    if (!err) {
        if (planar) { for (uint32_t i = lane; i < chunk_nbytes; i += 32) o[i] = __ldcg(t + i); }
        else {
            const uint32_t ne = chunk_nbytes / typesize;
            for (uint32_t i = lane; i < ne * typesize; i += 32) o[i] = __ldcg(t + (i % typesize) * ne + i / typesize);
            for (uint32_t i = ne * typesize + lane; i < chunk_nbytes; i += 32) o[i] = __ldcg(t + i);
        }
    }
    if (lane == 0) status[wid] = err;
}

}  // namespace hb

extern "C" int hb_decode_frames(const uint8_t *frames, const uint64_t *offsets, uint64_t n_frames,
                                uint64_t chunk_nbytes, uint8_t *out, int planar, int device) {
    if (!n_frames) return HB_OK;
    if (!frames || !offsets || !out || chunk_nbytes == 0 || chunk_nbytes > 0x7fffffffull) return api_fail(HB_ERR_ARG, "bad argument");
    int nd = 0;
    if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) return api_fail(HB_ERR_CUDA, "no CUDA device: libhaplo_b200 has no CPU fallback");
    if (cudaSetDevice(device) != cudaSuccess) return api_fail(HB_ERR_CUDA, "cudaSetDevice failed");
    const uint64_t total = offsets[n_frames];
    uint8_t *d_frames = nullptr, *d_tmp = nullptr, *d_out = nullptr;
    uint64_t *d_off = nullptr;
    int *d_status = nullptr;
    std::vector<int> status(n_frames);
    int rc = HB_OK;
    cudaError_t e = cudaSuccess;
    auto ck = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    ck(cudaMalloc(&d_frames, total + 64));
    ck(cudaMalloc(&d_off, (n_frames + 1) * 8));
    ck(cudaMalloc(&d_tmp, n_frames * chunk_nbytes));
    ck(cudaMalloc(&d_out, n_frames * chunk_nbytes));
    ck(cudaMalloc(&d_status, n_frames * sizeof(int)));
    if (e == cudaSuccess) {
        ck(cudaMemcpy(d_frames, frames, total, cudaMemcpyHostToDevice));
        ck(cudaMemcpy(d_off, offsets, (n_frames + 1) * 8, cudaMemcpyHostToDevice));
        decode_frames_kernel<<<(unsigned)((n_frames + 7) / 8), 256>>>(d_frames, d_off, n_frames, (uint32_t)chunk_nbytes,
                                                                     d_tmp, d_out, planar, d_status);
        count_launch();
        ck(cudaGetLastError());
        ck(cudaMemcpy(out, d_out, n_frames * chunk_nbytes, cudaMemcpyDeviceToHost));
        ck(cudaMemcpy(status.data(), d_status, n_frames * sizeof(int), cudaMemcpyDeviceToHost));
    }
    cudaFree(d_frames); cudaFree(d_off); cudaFree(d_tmp); cudaFree(d_out); cudaFree(d_status);
    if (e != cudaSuccess) return api_fail(HB_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e));
    for (uint64_t i = 0; i < n_frames; ++i)
        if (status[i]) { rc = api_fail(HB_ERR_IO, "corrupt or unsupported Blosc2 frame (chunk " + std::to_string(i) + ", code " + std::to_string(status[i]) + ")"); break; }
    return rc;
}
