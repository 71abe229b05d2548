// (b) Tile intersection: count -> scan -> emit (key,value) -> stable radix sort -> per-tile ranges.
//
// Replaces gsplat's isect_tiles (count pass, torch.cumsum, fill pass, cub::DeviceRadixSort::SortPairs) and
// isect_offset_encode behind qed_splatter/model.py:267-288 (tile_size=16, model.py:243,277).
// Semantics: SURVEY.md Appendix A.3/A.4 == oracle/torch_impl.py::isect_tiles / isect_offset_encode.
// Everything here is integer work and must be bit-exact: same keys, same stable order (ties keep emission
// order = ascending flat index), same ranges.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace qed {

// ------------------------------------------------------------------------------------------------
// count (only for callers that did not project through qed_project_fwd)
// ------------------------------------------------------------------------------------------------
__global__ void isect_count_kernel(int64_t CN, const float2* __restrict__ means2d, const int32_t* __restrict__ radii,
                                   float tile_size, int tile_w, int tile_h, int32_t* __restrict__ tiles) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CN) return;
    int r = radii[i];
    int cnt = 0;
    if (r > 0) {
        float2 m = means2d[i];
        TileBox tb = tile_box(m.x, m.y, r, tile_size, tile_w, tile_h);
        cnt = (tb.x1 - tb.x0) * (tb.y1 - tb.y0);
    }
    tiles[i] = cnt;
}

// ------------------------------------------------------------------------------------------------
// inclusive scan int32 -> int64, three phases (reduce / scan of block sums / scan + add)
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t* smem_warp, int64_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < (kScanThreads / 32) ? smem_warp[lane] : 0;
        int64_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (kScanThreads / 32)) smem_warp[lane] = winc - w;
        if (lane == (kScanThreads / 32) - 1) smem_warp[kScanThreads / 32] = winc;
    }
    __syncthreads();
    total = smem_warp[kScanThreads / 32];
    int64_t res = smem_warp[warp] + inc - v;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(int64_t n, const int32_t* __restrict__ in, int64_t* __restrict__ block_sums) {
    __shared__ int64_t sw[kScanThreads / 32 + 1];
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k * kScanThreads + threadIdx.x;
        if (i < n) s += in[i];
    }
    int64_t total;
    block_exclusive_scan(s, sw, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block sums in place; writes grand total
__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(int64_t nb, int64_t* __restrict__ block_sums, int64_t* __restrict__ total_dev,
                                                                int64_t* __restrict__ total_host) {
    __shared__ int64_t sw[kScanThreads / 32 + 1];
    int64_t carry = 0;
    for (int64_t base = 0; base < nb; base += kScanThreads) {
        int64_t i = base + threadIdx.x;
        int64_t v = i < nb ? block_sums[i] : 0;
        int64_t total;
        int64_t ex = block_exclusive_scan(v, sw, total);
        if (i < nb) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) {
        *total_dev = carry;
        if (total_host) *total_host = carry;
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_final_kernel(int64_t n, const int32_t* __restrict__ in, const int64_t* __restrict__ block_sums,
                                                                 int64_t* __restrict__ out) {
    __shared__ int64_t sw[kScanThreads / 32 + 1];
    // blocked arrangement: thread t owns items [t*kScanItems, (t+1)*kScanItems) of the tile
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int32_t v[kScanItems];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int64_t total;
    int64_t ex = block_exclusive_scan(s, sw, total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        ex += v[k];
        if (base + k < n) out[base + k] = ex;
    }
}

// ------------------------------------------------------------------------------------------------
// emit
// ------------------------------------------------------------------------------------------------
__global__ void isect_emit_kernel(int C, int N, const float2* __restrict__ means2d, const int32_t* __restrict__ radii,
                                  const float* __restrict__ depths, const int64_t* __restrict__ cum, float tile_size, int tile_w,
                                  int tile_h, int tile_n_bits, int64_t* __restrict__ isect_ids, int32_t* __restrict__ flatten_ids) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)C * N) return;
    int r = radii[idx];
    if (r <= 0) return;
    float2 m = means2d[idx];
    TileBox tb = tile_box(m.x, m.y, r, tile_size, tile_w, tile_h);
    int64_t cur = idx > 0 ? cum[idx - 1] : 0;
    const int64_t cam = idx / N;
    const int64_t hi = cam << tile_n_bits;
    const int64_t depth_bits = (int64_t)(uint32_t)__float_as_int(depths[idx]);
    for (int y = tb.y0; y < tb.y1; ++y) {
        for (int x = tb.x0; x < tb.x1; ++x) {
            int64_t tile = (int64_t)y * tile_w + x;
            isect_ids[cur] = ((hi | tile) << 32) | depth_bits;
            flatten_ids[cur] = (int32_t)idx;
            ++cur;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ranges
// ------------------------------------------------------------------------------------------------
__global__ void tile_ranges_kernel(int64_t n_isects, const int64_t* __restrict__ ids, int C, int n_tiles, int tile_n_bits,
                                   int32_t* __restrict__ offsets) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_isects) return;
    const int64_t mask = ((int64_t)1 << tile_n_bits) - 1;
    int64_t hi = ids[i] >> 32;
    int64_t cur = (hi >> tile_n_bits) * n_tiles + (hi & mask);
    if (i == 0) {
        for (int64_t t = 0; t <= cur; ++t) offsets[t] = 0;
    } else {
        int64_t hp = ids[i - 1] >> 32;
        int64_t prev = (hp >> tile_n_bits) * n_tiles + (hp & mask);
        for (int64_t t = prev + 1; t <= cur; ++t) offsets[t] = (int32_t)i;
    }
    if (i == n_isects - 1) {
        for (int64_t t = cur + 1; t < (int64_t)C * n_tiles; ++t) offsets[t] = (int32_t)n_isects;
    }
}

// ------------------------------------------------------------------------------------------------
// own LSD radix sort (8-bit digits): per pass  upsweep histogram -> digit-major scan -> downsweep
// (block-local stable rank via __match_any_sync, block-sorted staging in smem, coalesced run writes)
// ------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 pairs per block
constexpr int kRadix = 256;

__global__ void __launch_bounds__(kSortThreads) sort_upsweep_kernel(int64_t n, const uint64_t* __restrict__ keys, int shift, int nblocks,
                                                                   uint32_t* __restrict__ hist /* [kRadix][nblocks] */) {
    __shared__ uint32_t sh[kRadix];
    sh[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int k = 0; k < kSortItems; ++k) {
        int64_t i = base + k * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&sh[(keys[i] >> shift) & (kRadix - 1)], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// exclusive scan over the digit-major [kRadix*nblocks] histogram, single block
__global__ void __launch_bounds__(kScanThreads) sort_scan_kernel(int64_t total, uint32_t* __restrict__ hist) {
    __shared__ int64_t sw[kScanThreads / 32 + 1];
    int64_t carry = 0;
    for (int64_t base = 0; base < total; base += (int64_t)kScanThreads * 4) {
        int64_t i0 = base + (int64_t)threadIdx.x * 4;
        uint32_t v[4];
        int64_t s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = (i0 + k < total) ? hist[i0 + k] : 0u;
            s += v[k];
        }
        int64_t tot;
        int64_t ex = block_exclusive_scan(s, sw, tot) + carry;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < total) hist[i0 + k] = (uint32_t)ex;
            ex += v[k];
        }
        carry += tot;
    }
}

__global__ void __launch_bounds__(kSortThreads) sort_downsweep_kernel(int64_t n, const uint64_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                                                                     uint64_t* __restrict__ keys_out, int32_t* __restrict__ vals_out, int shift,
                                                                     int nblocks, const uint32_t* __restrict__ hist) {
    constexpr int kWarps = kSortThreads / 32;
    constexpr int kPerWarp = kSortTile / kWarps;  // 512 consecutive pairs per warp
    constexpr int kRounds = kPerWarp / 32;        // 16
    extern __shared__ __align__(16) unsigned char sort_smem[];
    uint64_t* skeys = reinterpret_cast<uint64_t*>(sort_smem);                      // [kSortTile]
    int32_t* svals = reinterpret_cast<int32_t*>(skeys + kSortTile);                // [kSortTile]
    uint32_t(*warp_hist)[kRadix] = reinterpret_cast<uint32_t(*)[kRadix]>(svals + kSortTile);  // [kWarps][kRadix]
    uint32_t* digit_start = &warp_hist[0][0] + kWarps * kRadix;  // start of each digit's run inside the block-sorted tile
    uint32_t* global_base = digit_start + kRadix;                // global output index of the run's first element
    uint32_t* sscan = global_base + kRadix;                      // [kWarps + 1]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    const int tile_n = (n - base) < kSortTile ? (int)(n - base) : kSortTile;

    for (int i = threadIdx.x; i < kWarps * kRadix; i += kSortThreads) (&warp_hist[0][0])[i] = 0;
    __syncthreads();

    uint64_t key[kRounds];
    uint16_t rank[kRounds];
    // phase A: stable rank inside the warp's 512-pair sub-chunk
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int local = warp * kPerWarp + r * 32 + lane;
        const bool valid = local < tile_n;
        key[r] = valid ? keys_in[base + local] : ~0ull;
        const uint32_t d = valid ? (uint32_t)((key[r] >> shift) & (kRadix - 1)) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = warp_hist[warp][d];
            warp_hist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = (uint16_t)(old + __popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
    }
    __syncthreads();
    // phase B: thread d: exclusive prefix over warps, block count; then exclusive scan over digits
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            uint32_t c = warp_hist[w][d];
            warp_hist[w][d] = run;
            run += c;
        }
        // exclusive scan of `run` over the 256 digits
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) sscan[warp] = inc;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (int w = 0; w < kWarps; ++w) {
                uint32_t t = sscan[w];
                sscan[w] = acc;
                acc += t;
            }
        }
        __syncthreads();
        digit_start[d] = sscan[warp] + inc - run;
        global_base[d] = hist[(int64_t)d * nblocks + blockIdx.x];
    }
    __syncthreads();
    // phase C: place into block-sorted order in smem
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int local = warp * kPerWarp + r * 32 + lane;
        if (local < tile_n) {
            const uint32_t d = (uint32_t)((key[r] >> shift) & (kRadix - 1));
            const uint32_t pos = digit_start[d] + warp_hist[warp][d] + rank[r];
            skeys[pos] = key[r];
            svals[pos] = vals_in[base + local];
        }
    }
    __syncthreads();
    // phase D: coalesced run writes
    for (int i = threadIdx.x; i < tile_n; i += kSortThreads) {
        const uint64_t k = skeys[i];
        const uint32_t d = (uint32_t)((k >> shift) & (kRadix - 1));
        const int64_t dst = (int64_t)global_base[d] + (i - digit_start[d]);
        keys_out[dst] = k;
        vals_out[dst] = svals[i];
    }
}

constexpr size_t kSortSmemBytes = (size_t)kSortTile * 12 + (size_t)(kSortThreads / 32) * kRadix * 4 + 2 * kRadix * 4 + (kSortThreads / 32 + 1) * 4;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace qed

using namespace qed;

extern "C" int qed_isect_count(int C, int N, const float* means2d, const int32_t* radii, int tile_size, int tile_width,
                               int tile_height, int32_t* tiles_per_gauss, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (C < 0 || N < 0 || tile_size <= 0) return QED_ERR_BAD_ARG;
    int64_t CN = (int64_t)C * N;
    if (CN == 0) return QED_OK;
    if (!means2d || !radii || !tiles_per_gauss) return QED_ERR_BAD_ARG;
    isect_count_kernel<<<(unsigned)((CN + 255) / 256), 256, 0, stream>>>(CN, reinterpret_cast<const float2*>(means2d), radii, (float)tile_size,
                                                                         tile_width, tile_height, tiles_per_gauss);
    QED_LAUNCH_CHECK();
    return QED_OK;
}

extern "C" size_t qed_isect_scan_workspace_bytes(int64_t n) {
    int64_t nb = (n + kScanTile - 1) / kScanTile;
    return (size_t)(nb > 0 ? nb : 1) * sizeof(int64_t);
}

extern "C" int qed_isect_scan(int64_t n, const int32_t* tiles_per_gauss, int64_t* cum, int64_t* n_isects_dev,
                              int64_t* n_isects_host_pinned, void* workspace, size_t workspace_bytes, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || !n_isects_dev) return QED_ERR_BAD_ARG;
    if (n == 0) {
        QED_CUDA_TRY(cudaMemsetAsync(n_isects_dev, 0, sizeof(int64_t), stream));
        if (n_isects_host_pinned) QED_CUDA_TRY(cudaMemcpyAsync(n_isects_host_pinned, n_isects_dev, 8, cudaMemcpyDeviceToHost, stream));
        return QED_OK;
    }
    if (!tiles_per_gauss || !cum || !workspace) return QED_ERR_BAD_ARG;
    if (workspace_bytes < qed_isect_scan_workspace_bytes(n)) return QED_ERR_WORKSPACE;
    int64_t nb = (n + kScanTile - 1) / kScanTile;
    int64_t* sums = reinterpret_cast<int64_t*>(workspace);
    scan_reduce_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(n, tiles_per_gauss, sums);
    QED_LAUNCH_CHECK();
    scan_sums_kernel<<<1, kScanThreads, 0, stream>>>(nb, sums, n_isects_dev, nullptr);
    QED_LAUNCH_CHECK();
    scan_final_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(n, tiles_per_gauss, sums, cum);
    QED_LAUNCH_CHECK();
    if (n_isects_host_pinned) QED_CUDA_TRY(cudaMemcpyAsync(n_isects_host_pinned, n_isects_dev, 8, cudaMemcpyDeviceToHost, stream));
    return QED_OK;
}

extern "C" int qed_isect_emit(int C, int N, const float* means2d, const int32_t* radii, const float* depths,
                              const int64_t* cum, int tile_size, int tile_width, int tile_height, int64_t* isect_ids,
                              int32_t* flatten_ids, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (C < 0 || N < 0 || tile_size <= 0) return QED_ERR_BAD_ARG;
    int64_t CN = (int64_t)C * N;
    if (CN == 0) return QED_OK;
    if (CN > 0x7fffffffLL) return QED_ERR_UNSUPPORTED;  // flatten_ids are int32 (as in gsplat)
    if (!means2d || !radii || !depths || !cum || !isect_ids || !flatten_ids) return QED_ERR_BAD_ARG;
    int n_tiles = tile_width * tile_height;
    int tile_n_bits = 0;
    while ((1 << tile_n_bits) <= n_tiles) ++tile_n_bits;  // == int.bit_length()
    isect_emit_kernel<<<(unsigned)((CN + 255) / 256), 256, 0, stream>>>(C, N, reinterpret_cast<const float2*>(means2d), radii, depths, cum,
                                                                        (float)tile_size, tile_width, tile_height, tile_n_bits, isect_ids,
                                                                        flatten_ids);
    QED_LAUNCH_CHECK();
    return QED_OK;
}

extern "C" int qed_tile_ranges(int64_t n_isects, const int64_t* isect_ids_sorted, int C, int tile_width, int tile_height,
                               int32_t* isect_offsets, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_isects < 0 || C < 0) return QED_ERR_BAD_ARG;
    int n_tiles = tile_width * tile_height;
    if ((int64_t)C * n_tiles == 0) return QED_OK;
    if (!isect_offsets) return QED_ERR_BAD_ARG;
    if (n_isects == 0) {
        QED_CUDA_TRY(cudaMemsetAsync(isect_offsets, 0, (size_t)C * n_tiles * 4, stream));
        return QED_OK;
    }
    if (n_isects > 0x7fffffffLL) return QED_ERR_UNSUPPORTED;  // offsets are int32 (as in gsplat)
    if (!isect_ids_sorted) return QED_ERR_BAD_ARG;
    int tile_n_bits = 0;
    while ((1 << tile_n_bits) <= n_tiles) ++tile_n_bits;
    tile_ranges_kernel<<<(unsigned)((n_isects + 255) / 256), 256, 0, stream>>>(n_isects, isect_ids_sorted, C, n_tiles, tile_n_bits, isect_offsets);
    QED_LAUNCH_CHECK();
    return QED_OK;
}

// ---- library baseline (what gsplat calls) ----
extern "C" size_t qed_sort_pairs_cub_workspace_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    n > 0 ? n : 1, 0, 64, (cudaStream_t)0);
    return bytes;
}

extern "C" int qed_sort_pairs_cub(int64_t n, int64_t* keys_in, int32_t* vals_in, int64_t* keys_out, int32_t* vals_out,
                                  int end_bit, void* workspace, size_t workspace_bytes, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || end_bit < 0 || end_bit > 64) return QED_ERR_BAD_ARG;
    if (n == 0) return QED_OK;
    if (!keys_in || !vals_in || !keys_out || !vals_out || !workspace) return QED_ERR_BAD_ARG;
    size_t need = qed_sort_pairs_cub_workspace_bytes(n);
    if (workspace_bytes < need) return QED_ERR_WORKSPACE;
    QED_CUDA_TRY(cub::DeviceRadixSort::SortPairs(workspace, workspace_bytes, reinterpret_cast<const uint64_t*>(keys_in),
                                                 reinterpret_cast<uint64_t*>(keys_out), vals_in, vals_out, n, 0, end_bit, stream));
    return QED_OK;
}

// ---- own radix sort ----
extern "C" size_t qed_sort_pairs_workspace_bytes(int64_t n) {
    if (n <= 0) return 256;
    int64_t nb = (n + kSortTile - 1) / kSortTile;
    return align_up((size_t)n * 8, 256) + align_up((size_t)n * 4, 256) + align_up((size_t)kRadix * nb * 4, 256);
}

extern "C" int qed_sort_pairs(int64_t n, int64_t* keys_in, int32_t* vals_in, int64_t* keys_out, int32_t* vals_out,
                              int end_bit, void* workspace, size_t workspace_bytes, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || end_bit < 0 || end_bit > 64) return QED_ERR_BAD_ARG;
    if (n == 0) return QED_OK;
    if (n > 0x7fffffffLL) return QED_ERR_UNSUPPORTED;
    if (!keys_in || !vals_in || !keys_out || !vals_out || !workspace) return QED_ERR_BAD_ARG;
    if (workspace_bytes < qed_sort_pairs_workspace_bytes(n)) return QED_ERR_WORKSPACE;
    const int nb = (int)((n + kSortTile - 1) / kSortTile);
    char* ws = reinterpret_cast<char*>(workspace);
    uint64_t* tmp_keys = reinterpret_cast<uint64_t*>(ws);
    int32_t* tmp_vals = reinterpret_cast<int32_t*>(ws + align_up((size_t)n * 8, 256));
    uint32_t* hist = reinterpret_cast<uint32_t*>(ws + align_up((size_t)n * 8, 256) + align_up((size_t)n * 4, 256));
    const int passes = (end_bit + 7) / 8;
    if (passes == 0) {
        QED_CUDA_TRY(cudaMemcpyAsync(keys_out, keys_in, (size_t)n * 8, cudaMemcpyDeviceToDevice, stream));
        QED_CUDA_TRY(cudaMemcpyAsync(vals_out, vals_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
        return QED_OK;
    }
    QED_CUDA_TRY(cudaFuncSetAttribute(sort_downsweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmemBytes));
    const uint64_t* src_k = reinterpret_cast<const uint64_t*>(keys_in);
    const int32_t* src_v = vals_in;
    for (int pass = 0; pass < passes; ++pass) {
        // last pass must land in *_out: destinations alternate out/tmp ending on out
        const bool to_out = ((passes - 1 - pass) % 2) == 0;
        uint64_t* dst_k = to_out ? reinterpret_cast<uint64_t*>(keys_out) : tmp_keys;
        int32_t* dst_v = to_out ? vals_out : tmp_vals;
        const int shift = pass * 8;
        sort_upsweep_kernel<<<nb, kSortThreads, 0, stream>>>(n, src_k, shift, nb, hist);
        QED_LAUNCH_CHECK();
        sort_scan_kernel<<<1, kScanThreads, 0, stream>>>((int64_t)kRadix * nb, hist);
        QED_LAUNCH_CHECK();
        sort_downsweep_kernel<<<nb, kSortThreads, kSortSmemBytes, stream>>>(n, src_k, src_v, dst_k, dst_v, shift, nb, hist);
        QED_LAUNCH_CHECK();
        src_k = dst_k;
        src_v = dst_v;
    }
    return QED_OK;
}
