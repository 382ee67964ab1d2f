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
#include <string>
#include <vector>

#include "../../include/haplo_b200.h"
#include "hb_common.cuh"
#include "hb_internal.h"
#include "hb_parse_struct.h"

namespace hb {

constexpr int kHist = 64;                 // bytes of history kept in front of the donor planes
constexpr int FRAME_HDR = 97, CHUNK_HDR = 32, OFFS_CHUNK = 40, FRAME_TRAILER = 35;

__device__ __forceinline__ uint32_t rd32(const uint8_t *p) {     // unaligned 4-byte read (shared memory)
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    return __funnelshift_r(w[0], w[1], (uint32_t)(a & 3) * 8u);
}

// Warp-cooperative LZ4 block encoder over src[0..n) with `hist` readable bytes before src.
//   anchor0 <= 0: literals still pending from the previous part of the block (src[anchor0..0)).
//   final: obey the end-of-block rules (last 5 bytes literal, last match starts <= n-12) and flush
//          the trailing literals; otherwise stop after the last match and report the pending
//          literal count through *pending.
// table: 1 << hashlog uint16 entries (position + 1; 0 = empty), cleared here.  n < 65535.
// Returns the number of bytes written to dst.  All lanes return the same value.
__device__ int warp_lz4(const uint8_t *src, int n, int hist, int anchor0, bool final, uint8_t *dst,
                        uint16_t *table, int hashlog, int *pending) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < (1 << hashlog); i += 32) table[i] = 0;
    __syncwarp();
    int op = 0, anchor = anchor0, pos = 0;
    const int mflimit = final ? n - 12 : n - 4;      // last position a match may start at
    const int mend = final ? n - 5 : n;              // matches end at or before this
    while (pos <= mflimit) {
        const int p = pos + lane;
        const bool valid = p <= mflimit;
        uint32_t v = 0, h = 0;
        int cand = -0x40000000;
        if (valid) {
            v = rd32(src + p);
            h = (v * 2654435761u) >> (32 - hashlog);
            const int c = (int)table[h] - 1;
            if (c >= 0 && c < p && rd32(src + c) == v) cand = c;   // c == p: left by a re-examined window
            else if (p - 1 >= -hist && rd32(src + p - 1) == v) cand = p - 1;     // run of one byte
        }
        __syncwarp();
        if (valid) table[h] = (uint16_t)(p + 1);
        __syncwarp();
        const unsigned hit = __ballot_sync(0xffffffffu, cand > -0x40000000);
        if (!hit) { pos += 32; continue; }
        const int f = __ffs(hit) - 1;
        const int m = pos + f;
        const int c = __shfl_sync(0xffffffffu, cand, f);
        int ml = 4;
        for (;;) {
            const int i = ml + lane;
            const bool same = (m + i < mend) && src[m + i] == src[c + i];
            const unsigned bal = __ballot_sync(0xffffffffu, same);
            if (bal == 0xffffffffu) { ml += 32; continue; }
            ml += __ffs(~bal) - 1;
            break;
        }
        if (m + ml > mend) ml = mend - m;             // (cannot happen: guarded above)
        if (ml < 4) { pos = m + 1; continue; }        // match would cross the end-of-block limit
        const int litlen = m - anchor;
        int o = op;
        if (lane == 0) {
            dst[o] = (uint8_t)((min(litlen, 15) << 4) | min(ml - 4, 15));
        }
        ++o;
        if (litlen >= 15) {
            int rem = litlen - 15;
            while (rem >= 255) { if (lane == 0) dst[o] = 255; ++o; rem -= 255; }
            if (lane == 0) dst[o] = (uint8_t)rem;
            ++o;
        }
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
    if (final) {
        const int litlen = n - anchor;
        int o = op;
        if (lane == 0) dst[o] = (uint8_t)(min(litlen, 15) << 4);
        ++o;
        if (litlen >= 15) {
            int rem = litlen - 15;
            while (rem >= 255) { if (lane == 0) dst[o] = 255; ++o; rem -= 255; }
            if (lane == 0) dst[o] = (uint8_t)rem;
            ++o;
        }
        for (int i = lane; i < litlen; i += 32) dst[o + i] = src[anchor + i];
        op = o + litlen;
        if (pending) *pending = 0;
    } else if (pending) {
        *pending = n - anchor;
    }
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
//   donor_encode_kernel    one warp per (sample, chunk): allele planes -> LZ4 tail of the same block;
//                          writes [pad][sequences][offsets chunk][trailer] into a staging slot, shifted
//                          so that it lines up with the template's end modulo 16.
//   frame_offsets_kernel / row_base_kernel   frame sizes -> 16-byte aligned offsets, [sample][chunk] order.
//   assemble_kernel        one warp per frame: 16-byte vector copy template + staged tail -> final
//                          position, patching the four size fields in registers.  This is where the
//                          bytes go: C_out is written exactly once, the templates stay in L2.
// ------------------------------------------------------------------------------------------
constexpr int TMPL_HDR = FRAME_HDR + CHUNK_HDR + 8;        // 137 bytes of a frame precede its LZ4 block
constexpr int FRAME_TAIL = OFFS_CHUNK + FRAME_TRAILER;     // 75 bytes follow it

__device__ const uint8_t kFrameTail[FRAME_TAIL] = {
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
};

__global__ void __launch_bounds__(256) site_template_kernel(const SiteArgs4 a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_len;
    const uint32_t cr = a.cr;
    const uint32_t n = 33u * cr;
    uint8_t *planes = smem + 16;                            // 16 bytes of (unused) history in front
    uint8_t *outb = planes + ((n + 19) & ~15u);             // the template: header, then LZ4 bytes
    uint16_t *table = reinterpret_cast<uint16_t *>(outb + a.tmpl_cap);
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
    for (uint32_t i = threadIdx.x; i < a.tmpl_cap / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(outb)[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) write_frame_head(outb, 35u * cr);
    // Chunks of fewer than 6 records cannot honour LZ4's end-of-block rules (last match >= 12 bytes
    // before the end) once the 2*cr allele bytes follow: such blocks are stored raw (csize == size).
    const bool raw = cr < 6;
    uint8_t *lz = outb + TMPL_HDR;
    if (raw) {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) lz[i] = planes[i];
        if (threadIdx.x == 0) s_len = n;
    } else if (threadIdx.x < 32) {
        // Planes 24..32 (ALT bytes 1..9) are all zero: the head is encoded up to the first of those
        // zeros, then ONE offset-1 match covers the other 9*cr-1 -- so the site part always ends on a
        // sequence boundary and every donor continues the block with nothing pending and zeros behind it.
        const int n1 = 24 * (int)cr + 1;
        int pending = 0;
        int len = warp_lz4(planes, n1, 0, 0, false, lz, table, 12, &pending);
        len = warp_emit_seq(lz, len, planes + n1 - pending, pending, 1, (int)n - n1);
        if (threadIdx.x == 0) s_len = (uint32_t)len;
    }
    __syncthreads();
    const uint32_t tl = TMPL_HDR + s_len;
    if (threadIdx.x == 0) a.tmpl_len[c] = tl;
    uint4 *dstp = reinterpret_cast<uint4 *>(a.tmpl + c * a.tmpl_cap);
    const uint4 *srcp = reinterpret_cast<const uint4 *>(outb);
    for (uint32_t i = threadIdx.x; i < (tl + 15) / 16; i += blockDim.x) dstp[i] = srcp[i];
}

// ------------------------------------------------------------------------------------------
// one warp per (sample, chunk): allele planes -> LZ4 tail of the block -> staging slot
// ------------------------------------------------------------------------------------------
struct DonorArgs {
    const int8_t *gt0, *gt1;
    uint64_t gt_stride, n_records;
    uint32_t cr, n_samples, s0;  // samples [s0, s0 + n_samples)
    uint64_t n_chunks;
    const uint32_t *tmpl_len;
    uint8_t *stage;              // [n_samples * n_chunks][dslot]
    uint32_t dslot;
    uint32_t *dlen;              // [n_samples * n_chunks] LZ4 bytes of the allele planes
    uint32_t warp_smem;          // bytes of shared memory per warp
};

__global__ void __launch_bounds__(256) donor_encode_kernel(const DonorArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t wid = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (wid >= a.n_chunks * a.n_samples) return;
    const uint32_t s = (uint32_t)(wid / a.n_chunks);
    const uint64_t c = wid % a.n_chunks;
    const uint32_t cr = a.cr, n = 2u * cr;
    uint8_t *base = smem + (size_t)warp * a.warp_smem;
    uint8_t *src = base + kHist;                                   // history sits right in front
    uint8_t *outb = src + ((n + 19) & ~15u);
    uint16_t *table = reinterpret_cast<uint16_t *>(outb + a.dslot);
    // history (the zero tail of the site planes) + the two allele planes (zero-padded past n_records,
    // like an HDF5 edge chunk)
    for (int i = lane; i < kHist; i += 32) base[i] = 0;
    const uint64_t r0 = c * cr;
    const int8_t *g0 = a.gt0 + (uint64_t)(a.s0 + s) * a.gt_stride + r0, *g1 = a.gt1 + (uint64_t)(a.s0 + s) * a.gt_stride + r0;
    for (uint32_t i = lane; i < cr; i += 32) {
        const bool in = r0 + i < a.n_records;
        src[i] = in ? (uint8_t)g0[i] : 0;
        src[cr + i] = in ? (uint8_t)g1[i] : 0;
    }
    const uint32_t sh = a.tmpl_len[c] & 15u;                       // lines the slot up with the template's end
    if (lane < 16) outb[lane] = 0;
    __syncwarp();
    uint8_t *seq = outb + sh;
    int dlen;
    if (cr < 6) {                                  // raw block, see site_template_kernel
        for (uint32_t i = lane; i < n; i += 32) seq[i] = src[i];
        dlen = (int)n;
    } else {
        dlen = warp_lz4(src, (int)n, min(kHist, 9 * (int)cr - 1), 0, true, seq, table, 10, nullptr);
    }
    for (int i = lane; i < FRAME_TAIL; i += 32) seq[dlen + i] = kFrameTail[i];
    const uint32_t used = sh + (uint32_t)dlen + FRAME_TAIL;
    for (uint32_t i = used + lane; i < ((used + 15) & ~15u); i += 32) outb[i] = 0;
    __syncwarp();
    uint4 *dstp = reinterpret_cast<uint4 *>(a.stage + wid * (uint64_t)a.dslot);
    const uint4 *srcp = reinterpret_cast<const uint4 *>(outb);
    for (uint32_t i = lane; i < (used + 15) / 16; i += 32) dstp[i] = srcp[i];
    if (lane == 0) a.dlen[wid] = (uint32_t)dlen;
}

// ------------------------------------------------------------------------------------------
// Allele-plane encoder, bit-parallel (the default).  After the SNP filter the allele bytes are 0 / 1
// (rarely -9), so the two planes are packed to one bit per byte (B = bit 0, N = "any other bit set")
// and LZ4 matches are found with word-wide logic instead of byte-wise hashing:
//   offset 1   (run of equal bytes)                 m1 = ~(B ^ B<<1) & ~(N | N<<1)
//   offset cr  (same record, the other haplotype;   mc = ~(B1 ^ B0) & ~(N1 | N0)
//               for plane 0 the zero plane 32 of the site part: mc = ~B0 & ~N0)
// A byte with N set never matches, so the stream is exact for ANY byte values; it just compresses
// best on genotype data.  Lanes 0-15 parse 16 segments of plane 0, lanes 16-31 the same segments of
// plane 1, greedily and independently (a match never crosses a segment); trailing literals of a
// segment are carried into the next lane's first sequence, an exclusive scan of the encoded sizes
// gives every lane its output offset, and each lane writes its own sequences.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_lsb4(uint32_t w) { return ((w & 0x01010101u) * 0x01020408u) >> 24; }
__device__ __forceinline__ uint32_t pack_nz4(uint32_t w) {       // bit j = byte j has one of bits 1..7 set
    const uint32_t t = w & 0xFEFEFEFEu;
    const uint32_t nz = ((((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u) >> 7;
    return (nz * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t low_mask(int n) { return n <= 0 ? 0u : (n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u)); }
__device__ __forceinline__ int ctz32(uint32_t v) { return __clz(__brev(v)); }       // 32 for 0
__device__ __forceinline__ int seq_bytes(int lit, int ml) {
    return 3 + lit + (lit >= 15 ? 1 + (lit - 15) / 255 : 0) + (ml >= 19 ? 1 + (ml - 19) / 255 : 0);
}

struct DonorBitsArgs {
    DonorArgs d;
    uint32_t rb;        // bytes of one raw plane row in shared memory (multiple of 16)
    uint32_t bww;       // words of one packed bit array (word 0 is a leading zero word)
    uint32_t caps;      // sequence slots per lane
};

__global__ void __launch_bounds__(256) donor_encode_bits_kernel(const DonorBitsArgs A) {
    extern __shared__ __align__(16) uint8_t smem[];
    const DonorArgs &a = A.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t wid = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (wid >= a.n_chunks * a.n_samples) return;
    const uint32_t s = (uint32_t)(wid / a.n_chunks);
    const uint64_t c = wid % a.n_chunks;
    const int cr = (int)a.cr, n = 2 * cr;
    uint8_t *base = smem + (size_t)warp * a.warp_smem;
    uint8_t *raw0 = base, *raw1 = base + A.rb;
    uint32_t *bits = reinterpret_cast<uint32_t *>(base + 2 * A.rb);      // B0, N0, B1, N1
    const int BWW = (int)A.bww;
    uint16_t *seqs = reinterpret_cast<uint16_t *>(bits + 4 * BWW);       // [caps][32]
    uint8_t *outb = reinterpret_cast<uint8_t *>(seqs + A.caps * 32);

    // ---- 1. planes -> shared memory (raw bytes for the literals) + packed bits
    const uint64_t r0 = c * (uint64_t)cr;
    const int alpha = (int)(r0 & 15);
    const int valid = (int)min((uint64_t)cr, a.n_records - r0);          // rows past n_records read as zero (HDF5 edge chunk)
    const int nvec = (alpha + cr + 15) >> 4;
    for (int i = lane; i < 4 * BWW; i += 32) bits[i] = 0;
    if (lane < 16) outb[lane] = 0;
    __syncwarp();
    const uint64_t rowbase = (uint64_t)(a.s0 + s) * a.gt_stride + (r0 & ~15ull);
    for (int v = lane; v < 2 * nvec; v += 32) {
        const int p = v >= nvec, k = v - p * nvec;
        const int lo = alpha - 16 * k, hi = alpha + valid - 16 * k;      // bytes [lo, hi) of this vector belong to the chunk
        uint4 x = make_uint4(0, 0, 0, 0);
        if (hi > 0) x = ldg_stream(reinterpret_cast<const uint4 *>((p ? a.gt1 : a.gt0) + rowbase + 16 * k));
        reinterpret_cast<uint4 *>(p ? raw1 : raw0)[k] = x;
        uint32_t b16 = pack_lsb4(x.x) | (pack_lsb4(x.y) << 4) | (pack_lsb4(x.z) << 8) | (pack_lsb4(x.w) << 12);
        uint32_t n16 = 0;
        if ((x.x | x.y | x.z | x.w) & 0xFEFEFEFEu)
            n16 = pack_nz4(x.x) | (pack_nz4(x.y) << 4) | (pack_nz4(x.z) << 8) | (pack_nz4(x.w) << 12);
        const uint32_t m = low_mask(hi) & ~low_mask(lo) & 0xFFFFu;
        reinterpret_cast<uint16_t *>(bits + (2 * p) * BWW + 1)[k] = (uint16_t)(b16 & m);
        reinterpret_cast<uint16_t *>(bits + (2 * p + 1) * BWW + 1)[k] = (uint16_t)(n16 & m);
    }
    __syncwarp();
    if (valid < cr && lane < 16) {               // the vector that straddles n_records: bytes past it must read 0
        const int idx = alpha + valid + lane;
        if (idx < ((alpha + valid + 15) & ~15)) { raw0[idx] = 0; raw1[idx] = 0; }
    }
    const uint32_t sh16 = a.tmpl_len[c] & 15u;   // lines the slot up with the template's end
    uint8_t *seq = outb + sh16;

    // ---- 2. per-lane greedy parse of one segment
    const int p = lane >> 4, q = lane & 15;
    const int seg = (cr + 15) >> 4;
    const int a0 = q * seg;
    const int seglen = max(0, min(seg, cr - a0));
    const int NW = (seg + 31) >> 5;
    const uint32_t *B = bits + (2 * p) * BWW + 1, *N = B + BWW;
    const uint32_t *B0 = bits + 1, *N0 = bits + BWW + 1;
    const int bi = alpha + a0, j0 = bi >> 5, shb = bi & 31;
    const int mlimit = (p ? min(seglen, cr - 5 - a0) : seglen);      // positions < mlimit may lie inside a match
    const int slimit = p ? cr - 11 - a0 : 0x7fffffff;                // positions < slimit may start one
    uint32_t cb = 0, cn = 0;                                         // the byte in front of the segment
    if (a0 > 0 && seglen > 0) { const int i = bi - 1; cb = (B[i >> 5] >> (i & 31)) & 1u; cn = (N[i >> 5] >> (i & 31)) & 1u; }
    else if (p && seglen > 0) { const int i = alpha + cr - 1; cb = (B0[i >> 5] >> (i & 31)) & 1u; cn = (N0[i >> 5] >> (i & 31)) & 1u; }

    int m = 0, prev_end = 0, first_lit = 0, first_ml = 0, size_rest = 0;
    unsigned long long offmask = 0;
    bool open = false;
    int open_off = 0, open_start = 0, open_len = 0;
    auto record = [&](int st, int ml, int off) {
        seqs[m * 32 + lane] = (uint16_t)((st << 8) | ml);
        offmask |= (unsigned long long)off << m;
        const int lit = st - prev_end;
        if (m == 0) { first_lit = lit; first_ml = ml; }
        else size_rest += seq_bytes(lit, ml);
        prev_end = st + ml;
        ++m;
    };
    uint32_t m1c = 0, mcc = 0, m1n, mcn;
    auto masks = [&](int k, uint32_t &m1, uint32_t &mc) {
        if (k >= NW) { m1 = mc = 0; return; }
        const uint32_t bw = __funnelshift_r(B[j0 + k], B[j0 + k + 1], shb);
        const uint32_t nw = __funnelshift_r(N[j0 + k], N[j0 + k + 1], shb);
        const uint32_t bprev = (bw << 1) | cb, nprev = (nw << 1) | cn;
        cb = bw >> 31; cn = nw >> 31;
        const uint32_t vm = low_mask(mlimit - 32 * k);
        m1 = ~(bw ^ bprev) & ~(nw | nprev) & vm;
        if (p) {
            const uint32_t b0w = __funnelshift_r(B0[j0 + k], B0[j0 + k + 1], shb);
            const uint32_t n0w = __funnelshift_r(N0[j0 + k], N0[j0 + k + 1], shb);
            mc = ~(bw ^ b0w) & ~(nw | n0w) & vm;
        } else mc = ~bw & ~nw & vm;
    };
    masks(0, m1c, mcc);
    for (int k = 0; k < NW; ++k) {
        masks(k + 1, m1n, mcn);
        const uint32_t r1 = m1c & __funnelshift_r(m1c, m1n, 1) & __funnelshift_r(m1c, m1n, 2) & __funnelshift_r(m1c, m1n, 3);
        const uint32_t rc = mcc & __funnelshift_r(mcc, mcn, 1) & __funnelshift_r(mcc, mcn, 2) & __funnelshift_r(mcc, mcn, 3);
        const uint32_t r = (r1 | rc) & low_mask(slimit - 32 * k);
        int pos = 0;
        if (open) {
            const int cont = ctz32(~(open_off ? mcc : m1c));
            open_len += cont;
            pos = cont;
            if (cont < 32) { record(open_start, open_len, open_off); open = false; }
        }
        while (pos < 32) {
            const uint32_t x = r & (0xFFFFFFFFu << pos);
            if (!x) break;
            const int st = ctz32(x);
            const int l1 = ctz32(~(m1c >> st)), lc = ctz32(~(mcc >> st));
            // a run that reaches the end of this word is measured on into the next one, so that the
            // offset chosen is the one whose run is really the longer (and >= 4, as r promises)
            const int l1x = st + l1 == 32 ? l1 + ctz32(~m1n) : l1, lcx = st + lc == 32 ? lc + ctz32(~mcn) : lc;
            const int off = lcx > l1x;
            const int best = off ? lc : l1;
            if (st + best >= 32) { open = true; open_off = off; open_start = 32 * k + st; open_len = 32 - st; break; }
            record(32 * k + st, best, off);
            pos = st + best;
        }
        m1c = m1n; mcc = mcn;
    }
    if (open) record(open_start, open_len, open_off);
    const int trail = seglen - prev_end;

    // ---- 3. carry trailing literals forward, size scan
    int val = trail, flag = m > 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v2 = __shfl_up_sync(0xffffffffu, val, d), f2 = __shfl_up_sync(0xffffffffu, flag, d);
        if (lane >= d && !flag) { val += v2; flag = f2; }
    }
    int carry = __shfl_up_sync(0xffffffffu, val, 1);
    if (lane == 0) carry = 0;
    const int final_lit = __shfl_sync(0xffffffffu, val, 31);
    const int mysize = m > 0 ? seq_bytes(first_lit + carry, first_ml) + size_rest : 0;
    int inc = mysize;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    __syncwarp();

    // ---- 4. every lane writes its own sequences
    {
        int o = inc - mysize;
        int prev_abs = p * cr + a0 - carry;
        for (int j = 0; j < m; ++j) {
            const uint32_t e = seqs[j * 32 + lane];
            const int st = p * cr + a0 + (int)(e >> 8), ml = (int)(e & 255u);
            const int off = ((offmask >> j) & 1ull) ? cr : 1;
            const int lit = st - prev_abs;
            seq[o++] = (uint8_t)((min(lit, 15) << 4) | min(ml - 4, 15));
            if (lit >= 15) { int rem = lit - 15; while (rem >= 255) { seq[o++] = 255; rem -= 255; } seq[o++] = (uint8_t)rem; }
            for (int i = 0; i < lit; ++i) { const int P = prev_abs + i; seq[o++] = P < cr ? raw0[alpha + P] : raw1[alpha + P - cr]; }
            seq[o++] = (uint8_t)off; seq[o++] = (uint8_t)(off >> 8);
            if (ml >= 19) { int rem = ml - 19; while (rem >= 255) { seq[o++] = 255; rem -= 255; } seq[o++] = (uint8_t)rem; }
            prev_abs = st + ml;
        }
    }
    // the last sequence of the block: literals only (at least the 5 bytes the format demands)
    int o = total;
    if (lane == 0) seq[o] = (uint8_t)(min(final_lit, 15) << 4);
    ++o;
    if (final_lit >= 15) {
        int rem = final_lit - 15;
        while (rem >= 255) { if (lane == 0) seq[o] = 255; ++o; rem -= 255; }
        if (lane == 0) seq[o] = (uint8_t)rem;
        ++o;
    }
    for (int i = lane; i < final_lit; i += 32) { const int P = n - final_lit + i; seq[o + i] = P < cr ? raw0[alpha + P] : raw1[alpha + P - cr]; }
    const int dlen = o + final_lit;

    // ---- 5. constant frame tail, pad, copy out
    for (int i = lane; i < FRAME_TAIL; i += 32) seq[dlen + i] = kFrameTail[i];
    const uint32_t used = sh16 + (uint32_t)dlen + FRAME_TAIL;
    for (uint32_t i = used + lane; i < ((used + 15) & ~15u); i += 32) outb[i] = 0;
    __syncwarp();
    uint4 *dstp = reinterpret_cast<uint4 *>(a.stage + wid * (uint64_t)a.dslot);
    const uint4 *srcp = reinterpret_cast<const uint4 *>(outb);
    for (uint32_t i = lane; i < (used + 15) / 16; i += 32) dstp[i] = srcp[i];
    if (lane == 0) a.dlen[wid] = (uint32_t)dlen;
}

// ------------------------------------------------------------------------------------------
// frame sizes -> offsets.  Frames are laid out [sample][chunk], each starting on a 16-byte boundary.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
frame_offsets_kernel(const uint32_t *__restrict__ tmpl_len, const uint32_t *__restrict__ dlen, uint32_t n_chunks,
                     uint32_t *__restrict__ size, uint32_t *__restrict__ rowoff, uint64_t *__restrict__ rowtot,
                     unsigned long long *__restrict__ sum_sizes) {
    __shared__ uint32_t wsum[8];
    __shared__ uint64_t carry;
    const uint32_t s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    unsigned long long mine = 0;
    for (uint32_t base = 0; base < n_chunks; base += 256) {
        const uint32_t c = base + threadIdx.x;
        const uint64_t wid = (uint64_t)s * n_chunks + c;
        uint32_t sz = 0;
        if (c < n_chunks) { sz = FRAME_TAIL + tmpl_len[c] + dlen[wid]; size[wid] = sz; mine += sz; }
        const uint32_t pad = (sz + 15u) & ~15u;
        uint32_t inc = pad;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { if (w < (int)warp) wbase += wsum[w]; total += wsum[w]; }
        if (c < n_chunks) rowoff[wid] = (uint32_t)(carry + wbase + inc - pad);
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
    if (lane == 0 && mine) atomicAdd(sum_sizes, mine);
    if (threadIdx.x == 0) rowtot[s] = carry;
}

__global__ void __launch_bounds__(1024) row_base_kernel(const uint64_t *__restrict__ rowtot, uint32_t n, uint64_t *__restrict__ rowbase) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < n ? rowtot[i] : 0;
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t t = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint64_t wbase = 0, total = 0;
        for (int w = 0; w < 32; ++w) { if (w < (int)warp) wbase += wsum[w]; total += wsum[w]; }
        if (i < n) rowbase[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) rowbase[n] = carry;
}

// ------------------------------------------------------------------------------------------
// one warp per frame: template + staged tail -> final position
// ------------------------------------------------------------------------------------------
struct AsmArgs {
    const uint8_t *tmpl; uint32_t tmpl_cap; const uint32_t *tmpl_len;
    const uint8_t *stage; uint32_t dslot; const uint32_t *dlen;
    const uint32_t *rowoff; const uint64_t *rowbase;
    uint8_t *frames;
    uint32_t n_chunks, n_samples;
};

__device__ __forceinline__ uint4 ldg_nc(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__global__ void __launch_bounds__(256) assemble_kernel(const AsmArgs a) {
    const int lane = threadIdx.x & 31;
    const uint64_t wid = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (uint64_t)a.n_chunks * a.n_samples) return;
    const uint32_t s = (uint32_t)(wid / a.n_chunks), c = (uint32_t)(wid % a.n_chunks);
    const uint32_t tl = a.tmpl_len[c], dl = a.dlen[wid];
    const uint32_t lz = tl - TMPL_HDR + dl, cb = CHUNK_HDR + 8 + lz, flen = tl + dl + FRAME_TAIL;
    const uint32_t jb = tl >> 4, nvec = (flen + 15) >> 4;
    const uint4 *T = reinterpret_cast<const uint4 *>(a.tmpl + (uint64_t)c * a.tmpl_cap);
    const uint4 *G = reinterpret_cast<const uint4 *>(a.stage + wid * (uint64_t)a.dslot);
    uint4 *D = reinterpret_cast<uint4 *>(a.frames + a.rowbase[s] + a.rowoff[wid]);
#pragma unroll 4
    for (uint32_t j = lane; j < nvec; j += 32) {
        uint4 v;
        if (j < jb) v = ldg_nc(T + j);
        else {
            v = ldg_stream(G + (j - jb));
            if (j == jb && (tl & 15u)) {           // the template's last bytes and the slot's pad are zero where the other has data
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
    uint32_t n_samples = 0, tmpl_cap = 0, dslot = 0, warp_smem = 0;
    uint32_t rb = 0, bww = 0, caps = 0;      // bit-parallel allele encoder geometry
    bool bits_encoder = true;                // false: the byte-wise warp LZ4 matcher (HB_DONOR_ENCODER=lz4, and for chunks of < 6 records)
    int warps_per_cta = 8;
    size_t smem_site = 0;
    uint8_t *d_tmpl = nullptr, *d_stage = nullptr, *d_frames = nullptr;
    uint32_t *d_tmpl_len = nullptr, *d_dlen = nullptr, *d_size = nullptr, *d_rowoff = nullptr;
    uint64_t *d_rowtot = nullptr, *d_rowbase = nullptr;
    unsigned long long *d_sum = nullptr;
    uint64_t frames_cap = 0;
    uint64_t total_bytes = 0, padded_bytes = 0;
    // host copies of the layout, fetched on demand
    bool layout_valid = false;
    std::vector<uint32_t> h_size, h_rowoff;      // [n_samples][n_chunks]
    std::vector<uint64_t> h_rowbase;             // [n_samples + 1]
    std::vector<uint8_t> h_row;                  // scratch for hb_frames_fetch_sample
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float ms_site = 0, ms_gt = 0, ms_offsets = 0, ms_assemble = 0;
};

static int frames_run(hb_frames *f, hb_parse *p) {
    cudaError_t e;
#define CUF(x) do { e = (x); if (e != cudaSuccess) return api_fail(HB_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e) + " at " #x); } while (0)
    CUF(cudaSetDevice(f->device));
    const uint64_t n = f->n_records;
    const uint32_t cr = (uint32_t)f->cr;
    const uint64_t n_frames = f->n_chunks * f->n_samples;
    f->layout_valid = false;
    SiteArgs4 sa;
    sa.chrom5 = p->d_chrom5; sa.start = p->d_start; sa.stop = p->d_stop; sa.ref = p->d_ref; sa.alt = p->d_alt;
    sa.n_records = n; sa.cr = cr; sa.tmpl = f->d_tmpl; sa.tmpl_cap = f->tmpl_cap; sa.tmpl_len = f->d_tmpl_len;
    CUF(cudaMemsetAsync(f->d_sum, 0, 8, f->stream));
    CUF(cudaEventRecord(f->ev[0], f->stream));
    site_template_kernel<<<(unsigned)f->n_chunks, 256, f->smem_site, f->stream>>>(sa);
    count_launch();
    CUF(cudaEventRecord(f->ev[1], f->stream));
    DonorArgs da;
    da.gt0 = p->d_gt[0]; da.gt1 = p->d_gt[1]; da.gt_stride = p->gt_stride; da.n_records = n;
    da.cr = cr; da.n_samples = f->n_samples; da.s0 = 0; da.n_chunks = f->n_chunks;
    da.tmpl_len = f->d_tmpl_len; da.stage = f->d_stage; da.dslot = f->dslot; da.dlen = f->d_dlen;
    da.warp_smem = f->warp_smem;
    const int wpc = f->warps_per_cta;
    if (f->bits_encoder) {
        DonorBitsArgs ba;
        ba.d = da; ba.rb = f->rb; ba.bww = f->bww; ba.caps = f->caps;
        donor_encode_bits_kernel<<<(unsigned)((n_frames + wpc - 1) / wpc), wpc * 32, (size_t)wpc * f->warp_smem, f->stream>>>(ba);
    } else {
        donor_encode_kernel<<<(unsigned)((n_frames + wpc - 1) / wpc), wpc * 32, (size_t)wpc * f->warp_smem, f->stream>>>(da);
    }
    count_launch();
    CUF(cudaEventRecord(f->ev[2], f->stream));
    frame_offsets_kernel<<<f->n_samples, 256, 0, f->stream>>>(f->d_tmpl_len, f->d_dlen, (uint32_t)f->n_chunks, f->d_size,
                                                              f->d_rowoff, f->d_rowtot, f->d_sum);
    row_base_kernel<<<1, 1024, 0, f->stream>>>(f->d_rowtot, f->n_samples, f->d_rowbase);
    count_launch(2);
    CUF(cudaEventRecord(f->ev[3], f->stream));
    uint64_t tot[2] = {0, 0};
    CUF(cudaMemcpyAsync(&tot[0], f->d_rowbase + f->n_samples, 8, cudaMemcpyDeviceToHost, f->stream));
    CUF(cudaMemcpyAsync(&tot[1], f->d_sum, 8, cudaMemcpyDeviceToHost, f->stream));
    CUF(cudaStreamSynchronize(f->stream));
    CUF(cudaGetLastError());
    f->padded_bytes = tot[0];
    f->total_bytes = tot[1];
    if (f->frames_cap < f->padded_bytes) {
        if (f->d_frames) { cudaFree(f->d_frames); f->d_frames = nullptr; f->frames_cap = 0; }
        const uint64_t cap = f->padded_bytes + f->padded_bytes / 64 + 4096;     // head-room for re-runs on new data
        e = cudaMalloc(&f->d_frames, cap);
        if (e != cudaSuccess) return api_fail(HB_ERR_MEM, std::string("cudaMalloc of the frame buffer (") + std::to_string(cap) + " bytes): " + cudaGetErrorString(e));
        f->frames_cap = cap;
    }
    AsmArgs aa;
    aa.tmpl = f->d_tmpl; aa.tmpl_cap = f->tmpl_cap; aa.tmpl_len = f->d_tmpl_len;
    aa.stage = f->d_stage; aa.dslot = f->dslot; aa.dlen = f->d_dlen;
    aa.rowoff = f->d_rowoff; aa.rowbase = f->d_rowbase; aa.frames = f->d_frames;
    aa.n_chunks = (uint32_t)f->n_chunks; aa.n_samples = f->n_samples;
    assemble_kernel<<<(unsigned)((n_frames + 7) / 8), 256, 0, f->stream>>>(aa);
    count_launch();
    CUF(cudaEventRecord(f->ev[4], f->stream));
    CUF(cudaStreamSynchronize(f->stream));
    CUF(cudaGetLastError());
    cudaEventElapsedTime(&f->ms_site, f->ev[0], f->ev[1]);
    cudaEventElapsedTime(&f->ms_gt, f->ev[1], f->ev[2]);
    cudaEventElapsedTime(&f->ms_offsets, f->ev[2], f->ev[3]);
    cudaEventElapsedTime(&f->ms_assemble, f->ev[3], f->ev[4]);
#undef CUF
    return HB_OK;
}

static int frames_layout(hb_frames *f) {
    if (f->layout_valid) return HB_OK;
    const uint64_t n_frames = f->n_chunks * f->n_samples;
    f->h_size.resize(n_frames); f->h_rowoff.resize(n_frames); f->h_rowbase.resize((size_t)f->n_samples + 1);
    if (n_frames) {
        if (cudaSetDevice(f->device) != cudaSuccess) return api_fail(HB_ERR_CUDA, "cudaSetDevice failed");
        cudaError_t e = cudaMemcpyAsync(f->h_size.data(), f->d_size, n_frames * 4, cudaMemcpyDeviceToHost, f->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(f->h_rowoff.data(), f->d_rowoff, n_frames * 4, cudaMemcpyDeviceToHost, f->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(f->h_rowbase.data(), f->d_rowbase, ((size_t)f->n_samples + 1) * 8, cudaMemcpyDeviceToHost, f->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(f->stream);
        if (e != cudaSuccess) return api_fail(HB_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(e));
    }
    f->layout_valid = true;
    return HB_OK;
}

extern "C" {

uint64_t hb_guess_chunk_records(uint64_t n_records) { return guess_chunk_records(n_records); }

void hb_frames_free(hb_frames *f) {
    if (!f) return;
    cudaSetDevice(f->device);
    cudaFree(f->d_tmpl); cudaFree(f->d_stage); cudaFree(f->d_frames);
    cudaFree(f->d_tmpl_len); cudaFree(f->d_dlen); cudaFree(f->d_size); cudaFree(f->d_rowoff);
    cudaFree(f->d_rowtot); cudaFree(f->d_rowbase); cudaFree(f->d_sum);
    for (auto &x : f->ev) if (x) cudaEventDestroy(x);
    delete f;
}

int hb_compress_records(hb_parse *p, uint64_t chunk_records, hb_frames **out) {
    if (!p || !out) return api_fail(HB_ERR_ARG, "null argument");
    *out = nullptr;
    if (cudaSetDevice(p->device) != cudaSuccess) return api_fail(HB_ERR_CUDA, "cudaSetDevice failed");
    const uint64_t n = p->h_st.n_records;
    if (!p->d_gt[0] && n) return api_fail(HB_ERR_NOGT, "parse was made without genotypes");
    hb_frames *f = new hb_frames();
    f->device = p->device; f->stream = p->stream;
    f->n_records = n; f->n_samples = p->n_samples;
    f->cr = chunk_records ? chunk_records : guess_chunk_records(n);
    f->n_chunks = n ? (n + f->cr - 1) / f->cr : 0;
    if (!f->n_chunks || !f->n_samples) { f->n_chunks = n ? f->n_chunks : 0; *out = f; f->layout_valid = false; return HB_OK; }
    if (f->cr > 2730) { hb_frames_free(f); return api_fail(HB_ERR_ARG, "chunk_records too large (at most 2730: the site encoder indexes 24*chunk_records+1 positions with 16 bits)"); }
    const uint32_t cr = (uint32_t)f->cr;
    const uint32_t n_site = 33u * cr, n_gt = 2u * cr;
    auto bound = [](uint32_t x) { return x + x / 255 + 64; };
    f->tmpl_cap = (TMPL_HDR + bound(n_site) + 15) & ~15u;
    f->smem_site = 16 + ((n_site + 19) & ~15u) + f->tmpl_cap + (2u << 12);
    if (f->smem_site > 220 * 1024) { hb_frames_free(f); return api_fail(HB_ERR_ARG, "chunk too large for the site encoder (33*chunk_records must fit shared memory)"); }
    f->dslot = (16 + bound(n_gt) + FRAME_TAIL + 15) & ~15u;
    const char *enc = getenv("HB_DONOR_ENCODER");
    f->bits_encoder = cr >= 6 && !(enc && !strcmp(enc, "lz4"));
    if (f->bits_encoder) {
        f->rb = ((15 + cr + 15) & ~15u) + 16;
        f->bww = ((cr + 30) / 32 + 4 + 3) & ~3u;
        f->caps = ((cr + 15) / 16) / 4 + 2;
        f->warp_smem = 2 * f->rb + 16 * f->bww + 64 * f->caps + f->dslot;
    } else {
        f->warp_smem = (kHist + ((n_gt + 19) & ~15u) + f->dslot + (2u << 10) + 15) & ~15u;
    }
    f->warps_per_cta = 8;
    while (f->warps_per_cta > 1 && (size_t)f->warps_per_cta * f->warp_smem > 200 * 1024) f->warps_per_cta >>= 1;
    if ((size_t)f->warps_per_cta * f->warp_smem > 220 * 1024) { hb_frames_free(f); return api_fail(HB_ERR_ARG, "chunk too large for the allele encoder"); }
    const uint64_t n_frames = f->n_chunks * f->n_samples;
    cudaError_t e = cudaSuccess;
    auto ck = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    ck(cudaMalloc(&f->d_tmpl, f->n_chunks * (uint64_t)f->tmpl_cap));
    ck(cudaMalloc(&f->d_tmpl_len, f->n_chunks * 4));
    ck(cudaMalloc(&f->d_stage, n_frames * (uint64_t)f->dslot));
    ck(cudaMalloc(&f->d_dlen, n_frames * 4));
    ck(cudaMalloc(&f->d_size, n_frames * 4));
    ck(cudaMalloc(&f->d_rowoff, n_frames * 4));
    ck(cudaMalloc(&f->d_rowtot, (uint64_t)f->n_samples * 8));
    ck(cudaMalloc(&f->d_rowbase, ((uint64_t)f->n_samples + 1) * 8));
    ck(cudaMalloc(&f->d_sum, 8));
    for (auto &x : f->ev) ck(cudaEventCreate(&x));
    ck(cudaFuncSetAttribute(site_template_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_site));
    if (f->bits_encoder) ck(cudaFuncSetAttribute(donor_encode_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)f->warps_per_cta * f->warp_smem)));
    else ck(cudaFuncSetAttribute(donor_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)f->warps_per_cta * f->warp_smem)));
    if (e != cudaSuccess) { hb_frames_free(f); return api_fail(HB_ERR_MEM, std::string("CUDA: ") + cudaGetErrorString(e)); }
    int rc = frames_run(f, p);
    if (rc != HB_OK) { hb_frames_free(f); return rc; }
    *out = f;
    return HB_OK;
}

int hb_frames_rerun(hb_frames *f, hb_parse *p) {
    if (!f || !p) return api_fail(HB_ERR_ARG, "null argument");
    if (p->h_st.n_records != f->n_records || p->n_samples != f->n_samples || p->device != f->device)
        return api_fail(HB_ERR_ARG, "hb_frames_rerun: the parse no longer has the shape these frames were made for");
    if (!f->n_chunks || !f->n_samples) return HB_OK;
    return frames_run(f, p);
}

int hb_frames_get_info(const hb_frames *f, hb_frames_info *info) {
    if (!f || !info) return api_fail(HB_ERR_ARG, "null argument");
    memset(info, 0, sizeof *info);
    info->n_records = f->n_records; info->n_chunks = f->n_chunks; info->chunk_records = f->cr;
    info->n_samples = f->n_samples; info->total_bytes = f->total_bytes;
    info->raw_bytes = 35ull * f->n_records * f->n_samples;
    info->ms_site = f->ms_site; info->ms_gt = f->ms_gt;
    info->padded_bytes = f->padded_bytes;
    info->d_frames = f->d_frames;
    info->ms_offsets = f->ms_offsets; info->ms_assemble = f->ms_assemble;
    return HB_OK;
}

int hb_frames_layout(hb_frames *f, uint64_t *offsets, uint32_t *sizes) {
    if (!f) return api_fail(HB_ERR_ARG, "null handle");
    int rc = frames_layout(f);
    if (rc != HB_OK) return rc;
    const uint64_t nc = f->n_chunks;
    for (uint32_t s = 0; s < f->n_samples; ++s)
        for (uint64_t c = 0; c < nc; ++c) {
            if (offsets) offsets[s * nc + c] = f->h_rowbase[s] + f->h_rowoff[s * nc + c];
            if (sizes) sizes[s * nc + c] = f->h_size[s * nc + c];
        }
    return HB_OK;
}

int hb_frames_fetch_all(hb_frames *f, uint8_t *buf, uint64_t cap) {
    if (!f || !buf) return api_fail(HB_ERR_ARG, "null argument");
    if (cap < f->padded_bytes) return api_fail(HB_ERR_ARG, "buffer too small");
    if (!f->padded_bytes) return HB_OK;
    if (cudaSetDevice(f->device) != cudaSuccess) return api_fail(HB_ERR_CUDA, "cudaSetDevice failed");
    cudaError_t e = cudaMemcpyAsync(buf, f->d_frames, f->padded_bytes, cudaMemcpyDeviceToHost, f->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(f->stream);
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
    // one contiguous D2H of the sample's (16-byte padded) row, then the pads are squeezed out on the host
    const uint64_t row = f->h_rowbase[s + 1] - f->h_rowbase[s];
    f->h_row.resize(row);
    if (cudaSetDevice(f->device) != cudaSuccess) return api_fail(HB_ERR_CUDA, "cudaSetDevice failed");
    cudaError_t e = cudaMemcpyAsync(f->h_row.data(), f->d_frames + f->h_rowbase[s], row, cudaMemcpyDeviceToHost, f->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(f->stream);
    if (e != cudaSuccess) return api_fail(HB_ERR_CUDA, std::string("D2H of frames failed: ") + cudaGetErrorString(e));
    uint64_t o = 0;
    for (uint64_t c = 0; c < nc; ++c) {
        const uint32_t sz = f->h_size[s * nc + c];
        memcpy(buf + o, f->h_row.data() + f->h_rowoff[s * nc + c], sz);
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
