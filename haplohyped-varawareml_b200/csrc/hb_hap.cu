// hb_hap.cu -- kernel 5: batched on-the-fly haplotype construction for the dataset.
//
// Replaces, for a whole batch in one launch, RandomHaplotypeDataset.encode_haplotypes
// (reference src/datasets/haplotype_dataset.py:86-110) and encode_sequence / array_to_onehot
// (src/utils/common_utils.py:84-103), with the documented repairs R1-R3:
//   R1 both haplotypes start from the index-encoded reference window;
//   R2 only records with window_start <= start < window_start + len are applied;
//   R3 one-hot out[i, c] = float(c == idx[i]), columns in encode_spec order.
// Kept literally: phase == 1 selects the ALT index, anything else the record's REF index
// (:99-100); duplicate positions -> the last record in file order wins (np.put_along_axis).
//
// One CTA builds 1024 window positions of one batch item: index bytes for both haplotypes are
// composed in shared memory (window gather through a 256-entry LUT, then the item's records in
// range -- found by binary search in the sorted start column -- scattered on top), and the
// float32 one-hot rows are streamed out as 16-byte vectors.  HBM-write bound: 2*B*L*C*4 bytes.
#include <map>
#include <mutex>
#include <utility>

#include "../../include/haplo_b200.h"
#include "hb_common.cuh"
#include "hb_internal.h"

namespace hb {

constexpr int HP_THREADS = 256;
constexpr int HP_TILE = 1024;

__device__ __forceinline__ uint64_t lower_bound_u32(const uint32_t *a, uint64_t lo, uint64_t hi, uint32_t key) {
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Record range of every tile, found once per tile boundary by one thread each (a binary search is ~16 dependent
// loads: done inside hap_kernel it was most of a CTA's lifetime).  rng[b * (ntiles + 1) + t] = first record of item b
// at or after window position min(t * HP_TILE, len).
__global__ void __launch_bounds__(256) hap_ranges_kernel(const hb_hap_batch a, uint32_t ntiles, uint64_t *__restrict__ rng) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)a.B * (ntiles + 1)) return;
    const uint32_t b = (uint32_t)(i / (ntiles + 1)), t = (uint32_t)(i % (ntiles + 1));
    const uint64_t nrec = a.item_nrec[b];
    const uint64_t g = (uint64_t)a.item_win_start[b] + min((uint64_t)t * HP_TILE, (uint64_t)a.item_len[b]);
    const uint32_t *start = reinterpret_cast<const uint32_t *>(a.item_start[b]);
    rng[i] = (nrec == 0 || g > 0xffffffffull) ? nrec : lower_bound_u32(start, 0, nrec, (uint32_t)g);
}

template <int CC>   // CC > 0: compile-time class count; 0: runtime
__global__ void __launch_bounds__(HP_THREADS) hap_kernel(const hb_hap_batch a, const uint64_t *__restrict__ rng, uint32_t ntiles) {
    __shared__ int8_t s_idx[2][HP_TILE];
    __shared__ int8_t s_lut[256];
    __shared__ uint64_t s_rng[2];
    const int tid = threadIdx.x;
    const uint32_t b = blockIdx.y;
    const uint32_t tile0 = blockIdx.x * HP_TILE;
    const uint32_t C = CC > 0 ? (uint32_t)CC : a.C;
    const uint32_t L = a.L;
    const uint32_t tn = min((uint32_t)HP_TILE, L - tile0);       // positions of this tile inside [0, L)
    const uint32_t len = a.item_len[b];
    const uint32_t ws = a.item_win_start[b];
    s_lut[tid] = a.lut[tid];
    __syncthreads();

    // ---- reference window -> class index (R1); beyond the window: -1 => all-zero row
    const uint8_t *seq = reinterpret_cast<const uint8_t *>(a.item_seq[b]);
    for (uint32_t k = tid; k < tn; k += HP_THREADS) {
        uint32_t p = tile0 + k;
        int8_t v = -1;
        if (p < len) v = s_lut[seq[p]];
        s_idx[0][k] = v;
        s_idx[1][k] = v;
    }
    __syncthreads();

    // ---- records inside this tile (R2), last duplicate wins
    {
        const uint64_t nrec = a.item_nrec[b];
        const uint32_t wend = min(len, tile0 + tn);              // exclusive, window-relative
        if (nrec && tile0 < wend) {              // (uniform over the CTA: the barrier below is reached by all or none)
            const uint32_t *start = reinterpret_cast<const uint32_t *>(a.item_start[b]);
            const uint8_t *ref = reinterpret_cast<const uint8_t *>(a.item_ref[b]);
            const uint8_t *alt = reinterpret_cast<const uint8_t *>(a.item_alt[b]);
            const int8_t *p1 = reinterpret_cast<const int8_t *>(a.item_p1[b]);
            const int8_t *p2 = reinterpret_cast<const int8_t *>(a.item_p2[b]);
            // genomic range [ws + tile0, ws + wend): from the pre-kernel's table, or -- short windows, one launch -- searched
            // here by two threads while the others wait (the table only pays when many tiles share the searches)
            uint64_t rb, re;
            if (rng) { rb = rng[(uint64_t)b * (ntiles + 1) + blockIdx.x]; re = rng[(uint64_t)b * (ntiles + 1) + blockIdx.x + 1]; }
            else {
                if (tid < 2) {
                    const uint64_t g = (uint64_t)ws + (tid ? wend : tile0);
                    s_rng[tid] = g > 0xffffffffull ? nrec : lower_bound_u32(start, 0, nrec, (uint32_t)g);
                }
                __syncthreads();
                rb = s_rng[0]; re = s_rng[1];
            }
            for (uint64_t r = rb + tid; r < re; r += HP_THREADS) {
                const uint32_t st = start[r];
                if (r + 1 < nrec && start[r + 1] == st) continue;   // a later record at the same position wins
                const uint32_t k = st - ws - tile0;
                const int8_t ri = s_lut[ref[r]], ai = s_lut[alt[r]];
                s_idx[0][k] = p1[r] == 1 ? ai : ri;
                s_idx[1][k] = p2[r] == 1 ? ai : ri;
            }
        }
    }
    __syncthreads();

    // ---- one-hot rows (R3), 16-byte stores
    const uint64_t base_el = ((uint64_t)b * L + tile0) * C;      // first float of this tile
    const uint32_t n_el = tn * C;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float *out = (h ? a.hap2 : a.hap1) + base_el;
        const int8_t *idx = s_idx[h];
        if ((base_el & 3) == 0) {
            const uint32_t n4 = n_el >> 2;
            for (uint32_t q = tid; q < n4; q += HP_THREADS) {
                uint32_t e = 4 * q;
                uint32_t p = e / C, c = e - p * C;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = (idx[p] == (int)c) ? 1.0f : 0.0f;
                    if (++c == C) { c = 0; ++p; }
                }
                uint4 w;
                w.x = __float_as_uint(v[0]); w.y = __float_as_uint(v[1]);
                w.z = __float_as_uint(v[2]); w.w = __float_as_uint(v[3]);
                stg_stream(reinterpret_cast<uint4 *>(out) + q, w);
            }
            for (uint32_t e = (n4 << 2) + tid; e < n_el; e += HP_THREADS) {
                uint32_t p = e / C, c = e - p * C;
                out[e] = (idx[p] == (int)c) ? 1.0f : 0.0f;
            }
        } else {
            for (uint32_t e = tid; e < n_el; e += HP_THREADS) {
                uint32_t p = e / C, c = e - p * C;
                out[e] = (idx[p] == (int)c) ? 1.0f : 0.0f;
            }
        }
    }
}

}  // namespace hb

using namespace hb;

namespace {
// scratch for the tile ranges: one buffer per (device, stream), grown on demand and kept (at most ~1 MB each).  Work on
// one stream is ordered, so a launch never overwrites ranges that an earlier launch on the same stream still reads;
// cudaMallocAsync would do too but costs ~200 us per call, ten times the kernel at small batches.
struct RangeScratch { uint64_t *p = nullptr; uint64_t n = 0; };
std::mutex g_rng_mu;
std::map<std::pair<int, cudaStream_t>, RangeScratch> g_rng;
}

extern "C" int hb_encode_haplotypes(const hb_hap_batch *batch) {
    if (!batch || !batch->B || !batch->L || !batch->C) return HB_OK;
    const uint32_t ntiles = (batch->L + HP_TILE - 1) / HP_TILE;
    dim3 grid(ntiles, batch->B);
    cudaStream_t st = (cudaStream_t)batch->stream;
    const uint64_t n_rng = (uint64_t)batch->B * (ntiles + 1);
    uint64_t *rng = nullptr;
    const bool fused = ntiles <= 2;              // short windows (the reference's default seq_length 1000): ONE launch
    if (!fused) {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(g_rng_mu);
        RangeScratch &sc = g_rng[{dev, st}];
        if (sc.n < n_rng) {
            if (sc.p) { cudaStreamSynchronize(st); dev_pool_free(sc.p); sc.p = nullptr; sc.n = 0; }
            if (dev_pool_alloc((void **)&sc.p, (n_rng + n_rng / 4 + 64) * 8) != cudaSuccess) return HB_ERR_MEM;
            sc.n = n_rng + n_rng / 4 + 64;
        }
        rng = sc.p;
    }
    if (!fused) hap_ranges_kernel<<<(unsigned)((n_rng + 255) / 256), 256, 0, st>>>(*batch, ntiles, rng);
    if (batch->C == 5) hap_kernel<5><<<grid, HP_THREADS, 0, st>>>(*batch, rng, ntiles);
    else if (batch->C == 4) hap_kernel<4><<<grid, HP_THREADS, 0, st>>>(*batch, rng, ntiles);
    else hap_kernel<0><<<grid, HP_THREADS, 0, st>>>(*batch, rng, ntiles);
    count_launch(fused ? 1 : 2);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}
