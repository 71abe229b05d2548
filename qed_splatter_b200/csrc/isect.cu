// (b) Tile intersection: count -> scan -> emit (key,value) -> stable radix sort -> per-tile ranges.
//
// Replaces gsplat's isect_tiles (count pass, torch.cumsum, fill pass, cub::DeviceRadixSort::SortPairs) and
// isect_offset_encode behind qed_splatter/model.py:267-288 (tile_size=16, model.py:243,277).
// Semantics: SURVEY.md Appendix A.3/A.4 == oracle/torch_impl.py::isect_tiles / isect_offset_encode.
// Everything here is integer work and must be bit-exact: same keys, same stable order (ties keep emission
// order = ascending flat index), same ranges.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "radix.cuh"
#include "scan.cuh"

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

extern "C" size_t qed_isect_scan_workspace_bytes(int64_t n) { return scan_workspace_bytes(n); }

extern "C" int qed_isect_scan(int64_t n, const int32_t* tiles_per_gauss, int64_t* cum, int64_t* n_isects_dev,
                              int64_t* n_isects_host_pinned, void* workspace, size_t workspace_bytes, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || !n_isects_dev) return QED_ERR_BAD_ARG;
    if (n == 0) {
        QED_CUDA_TRY(cudaMemsetAsync(n_isects_dev, 0, sizeof(int64_t), stream));
    } else {
        if (!tiles_per_gauss || !cum || !workspace) return QED_ERR_BAD_ARG;
        if (workspace_bytes < scan_workspace_bytes(n)) return QED_ERR_WORKSPACE;
        int rc = scan_inclusive(n, nullptr, ScanIdentity{tiles_per_gauss}, cum, n_isects_dev, workspace, stream);
        if (rc != QED_OK) return rc;
    }
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

// test hook (not part of the reference surface): 1 = single-kernel radix passes with decoupled look-back (default),
// 0 = upsweep / scan / downsweep per pass.  Identical output.
extern "C" int qed_debug_set_radix_onesweep(int enabled) {
    int old = g_radix_onesweep;
    g_radix_onesweep = enabled;  // 0 = three kernels per pass, 1 = look-back passes for small sorts, 2 = for every size
    return old;
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
    return align_up((size_t)n * 8, 256) + align_up((size_t)n * 4, 256) + radix_hist_bytes(n);
}

extern "C" int qed_sort_pairs(int64_t n, int64_t* keys_in, int32_t* vals_in, int64_t* keys_out, int32_t* vals_out,
                              int end_bit, void* workspace, size_t workspace_bytes, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || end_bit < 0 || end_bit > 64) return QED_ERR_BAD_ARG;
    if (n == 0) return QED_OK;
    if (n > 0x7fffffffLL) return QED_ERR_UNSUPPORTED;
    if (!keys_in || !vals_in || !keys_out || !vals_out || !workspace) return QED_ERR_BAD_ARG;
    if (workspace_bytes < qed_sort_pairs_workspace_bytes(n)) return QED_ERR_WORKSPACE;
    char* ws = reinterpret_cast<char*>(workspace);
    uint64_t* tmp_keys = reinterpret_cast<uint64_t*>(ws);
    int32_t* tmp_vals = reinterpret_cast<int32_t*>(ws + align_up((size_t)n * 8, 256));
    void* hist = ws + align_up((size_t)n * 8, 256) + align_up((size_t)n * 4, 256);
    return radix_sort_pairs<uint64_t>(n, nullptr, reinterpret_cast<const uint64_t*>(keys_in), vals_in, reinterpret_cast<uint64_t*>(keys_out),
                                      vals_out, tmp_keys, tmp_vals, hist, end_bit, stream);
}

// ------------------------------------------------------------------------------------------------
// Two-level intersection build (the product path).  Same output as emit + 64-bit sort + ranges, with ~5x
// less memory traffic:
//   prepare: compact the visible (camera, Gaussian) entries in flat-index order, radix sort them by
//            (camera, depth bits) (stable, so equal depths keep ascending flat index), scan their tile
//            counts in that order -> write offsets + n_isects.
//   fill:    emit (camera|tile, flat index) in depth order, warp-cooperatively and coalesced, then a stable
//            radix sort on the camera|tile bits only (2 passes for <= 16 bits).  Within a tile the depth
//            order survives, which is exactly the order of the reference's 64-bit key sort.
// ------------------------------------------------------------------------------------------------
namespace qed {

struct PrepareLayout {
    size_t scan_ws, keys[3], vals[3], hist, cum2, total;
};

static PrepareLayout prepare_layout(int64_t CN) {
    PrepareLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += align_up(bytes, 256);
        return r;
    };
    L.scan_ws = take(scan_workspace_bytes(CN));
    for (int i = 0; i < 3; ++i) L.keys[i] = take((size_t)CN * 8);
    for (int i = 0; i < 3; ++i) L.vals[i] = take((size_t)CN * 4);
    L.hist = take(radix_hist_bytes(CN));
    L.cum2 = take((size_t)CN * 8);
    L.total = o;
    return L;
}

// Sink of the visible-flag scan: the compaction IS the final scan phase (entry idx with flag 1 and inclusive count c
// goes to slot c - 1), so no cumulative-flag array is written and no separate compaction kernel runs.
template <typename KeyT>
struct CompactVisibleSink {
    int N;
    const float* depths;
    KeyT* keys;
    int32_t* vals;
    __device__ __forceinline__ void operator()(int64_t idx, int32_t flag, int64_t inc) const {
        if (!flag) return;
        const uint32_t db = (uint32_t)__float_as_int(depths[idx]);
        keys[inc - 1] = sizeof(KeyT) == 8 ? (KeyT)(((uint64_t)(idx / N) << 32) | db) : (KeyT)db;
        vals[inc - 1] = (int32_t)idx;
    }
};

// Entry-parallel emission in depth order.  Every block owns kEmitTile consecutive OUTPUT entries, so the
// work is balanced no matter how the tile counts are distributed (the nearest Gaussians are adjacent in
// depth order and can cover thousands of tiles each).  The block finds the Gaussians overlapping its entry
// range by binary search in cum2, stages their (first tile, box width, flat index, start offset) in shared
// memory, and every entry then finds its Gaussian by a binary search in that shared slice.
constexpr int kEmitThreads = 256;
constexpr int kEmitTile = 1024;

// first_j[b] = index (in depth order) of the Gaussian that owns output entry b * kEmitTile
// counts_dev != NULL (sizes only known on the device: no host sync before this launch): n_vis / n_isects are the
// CAPACITIES the buffers were sized for; the real counts are read here and published, clamped, in `clamped[2]` for the
// kernels that follow.  If the real entry count exceeds the capacity, an EMPTY list is built (clamped[1] = 0: every
// later kernel is then trivially memory-safe) and the caller, who reads the counts after the fact, redoes the step with
// larger buffers.
__global__ void emit_boundaries_kernel(int64_t n_vis, int64_t n_isects, const int64_t* __restrict__ counts_dev, int64_t* __restrict__ clamped,
                                       const int64_t* __restrict__ cum2, int32_t* __restrict__ first_j) {
    pdl_enter();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (counts_dev) {
        const int64_t real_vis = counts_dev[0], real_isects = counts_dev[1];
        n_vis = real_vis < n_vis ? real_vis : n_vis;
        n_isects = real_isects <= n_isects ? real_isects : 0;
        if (n_isects == 0) n_vis = 0;
        if (j == 0) {
            clamped[0] = n_vis;
            clamped[1] = n_isects;
        }
    }
    if (j >= n_vis) return;
    const int64_t start = j > 0 ? cum2[j - 1] : 0, end = cum2[j];
    for (int64_t b = (start + kEmitTile - 1) / kEmitTile; b * kEmitTile < end; ++b) first_j[b] = (int32_t)j;
}

// EXACT = exact tile lists for the compositor (gsplat's own lists hold every tile of the 3-sigma bounding box):
// every candidate (Gaussian, tile) entry is tested with the compositor's own conservative alpha >= 1/255
// ellipse test against the tile's pixel-centre rectangle and only survivors are written -- compacted in
// order at the START of the block's own kEmitTile-slot segment, with the count in seg_counts[block].  There
// is no global compaction (a decoupled look-back across the ~900 resident blocks cost more than the emit
// itself): the first pass of the tile sort reads the segments (SegCounts) and scatters densely.  About half
// of the candidates go (S1: 47 %), which halves the tile sort and the compositor's gather work and changes
// no pixel: a removed entry cannot pass the alpha test anywhere in its tile.
struct ExactEmit {
    const float4* geom;    // [C*N][2]: {mx, my, opacity, depth | conic a, b, c, -}
    int32_t* seg_counts;   // survivors per block
    unsigned long long* n_out;  // device scalar: total number of survivors (zeroed by the caller)
    int width, height;
};

template <bool EXACT>
__global__ void __launch_bounds__(kEmitThreads) emit_sorted_kernel(int64_t n_vis, int64_t n_isects, const int64_t* __restrict__ clamped, int N,
                                                                  const int32_t* __restrict__ sorted_vals,
                                                                  const int64_t* __restrict__ cum2, const int32_t* __restrict__ first_j,
                                                                  const float2* __restrict__ means2d, const int32_t* __restrict__ radii,
                                                                  float tile_size, int tile_w, int tile_h, int tile_n_bits,
                                                                  uint32_t* __restrict__ tkeys, int32_t* __restrict__ tvals, ExactEmit ex) {
    // per staged Gaussian: the part [lo, hi) of its entries that falls into this block (block-relative), the
    // offset k0 of entry `lo` inside the Gaussian's own tile list, its box and its flat index
    __shared__ int16_t s_lo[kEmitTile + 1], s_hi[kEmitTile + 1];
    __shared__ int32_t s_k0[kEmitTile + 1], s_first[kEmitTile + 1], s_idx[kEmitTile + 1];
    __shared__ int16_t s_bw[kEmitTile + 1], s_large[kEmitTile + 1];  // box width in tiles (<= 32767), staged index
    __shared__ __align__(16) uint32_t s_okey[kEmitTile];  // the block's output, staged so the global writes are fully coalesced
    __shared__ __align__(16) int32_t s_oval[kEmitTile];
    __shared__ int s_nlarge;
    __shared__ int s_wsum[kEmitThreads / 32];
    pdl_enter();
    const uint32_t vb = blockIdx.x;
    const int64_t e0 = (int64_t)vb * kEmitTile;
    if (clamped) {  // device-side counts: the grid covers the capacity, blocks behind the real count have nothing to do
        n_vis = clamped[0];
        n_isects = clamped[1];
        if (e0 >= n_isects) {
            if (EXACT && threadIdx.x == 0) ex.seg_counts[vb] = 0;
            return;
        }
    }
    const int64_t e1 = min(e0 + kEmitTile, n_isects);
    const int64_t j0 = first_j[vb];
    int64_t j1;  // Gaussian that owns entry e1 - 1
    if (e1 >= n_isects) {
        j1 = n_vis - 1;
    } else {
        const int64_t jn = first_j[vb + 1];  // owns entry e1
        j1 = (jn > 0 && cum2[jn - 1] == e1) ? jn - 1 : jn;  // jn starts exactly at e1 -> the previous one owns e1-1
    }
    const int G = (int)(j1 - j0) + 1;  // every Gaussian has >= 1 entry, so G <= kEmitTile
    if (threadIdx.x == 0) s_nlarge = 0;
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += kEmitThreads) {
        const int64_t j = j0 + g;
        const int idx = sorted_vals[j];
        const float2 m = means2d[idx];
        const TileBox tb = tile_box(m.x, m.y, radii[idx], tile_size, tile_w, tile_h);
        const int64_t gs = j > 0 ? cum2[j - 1] : 0, ge = cum2[j];
        const int64_t lo = max(gs, e0), hi = min(ge, e1);
        s_lo[g] = (int16_t)(lo - e0);
        s_hi[g] = (int16_t)(hi - e0);
        s_k0[g] = (int32_t)(lo - gs);
        s_first[g] = (int32_t)(((uint32_t)(idx / N) << tile_n_bits) | (uint32_t)(tb.y0 * tile_w + tb.x0));
        s_bw[g] = (int16_t)(tb.x1 - tb.x0);
        s_idx[g] = idx;
        if (hi - lo > 32) s_large[atomicAdd(&s_nlarge, 1)] = (int16_t)g;
    }
    __syncthreads();
    // small pieces (<= 32 entries, the common case): one lane per Gaussian
    for (int g = threadIdx.x; g < G; g += kEmitThreads) {
        const int lo = s_lo[g], n = s_hi[g] - lo;
        if (n > 32) continue;
        const int bw = s_bw[g], idx = s_idx[g], first = s_first[g];
        int ry = s_k0[g] / bw, rx = s_k0[g] - ry * bw;
        for (int t = 0; t < n; ++t) {
            s_okey[lo + t] = (uint32_t)(first + ry * tile_w + rx);
            s_oval[lo + t] = idx;
            if (++rx == bw) {
                rx = 0;
                ++ry;
            }
        }
    }
    // large pieces: one warp per Gaussian, coalesced
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nlarge = s_nlarge;
    for (int i = warp; i < nlarge; i += kEmitThreads / 32) {
        const int g = s_large[i];
        const int lo = s_lo[g], n = s_hi[g] - lo, k0 = s_k0[g];
        const int bw = s_bw[g], idx = s_idx[g], first = s_first[g];
        const float inv_bw = 1.0f / (float)bw;
        for (int t = lane; t < n; t += 32) {
            const int k = k0 + t;
            const int ry = (int)(((float)k + 0.5f) * inv_bw);  // exact: k < 2^16, margins 0.5/bw >> float error
            const int rx = k - ry * bw;
            s_okey[lo + t] = (uint32_t)(first + ry * tile_w + rx);
            s_oval[lo + t] = idx;
        }
    }
    __syncthreads();
    if (!EXACT) {
        for (int t = threadIdx.x; t < (int)(e1 - e0); t += kEmitThreads) {
            tkeys[e0 + t] = s_okey[t];
            tvals[e0 + t] = s_oval[t];
        }
        return;
    }
    // ---- exact lists: test, compact in order, look back for the block offset, write ----
    constexpr int kPer = kEmitTile / kEmitThreads;  // consecutive entries per thread
    const int cnt = (int)(e1 - e0);
    const uint32_t tile_mask = (1u << tile_n_bits) - 1u;
    static_assert(kPer == 4, "vector loads of the staged entries");
    uint32_t key[kPer];
    int32_t val[kPer];
    {
        const uint4 kk = reinterpret_cast<const uint4*>(s_okey)[threadIdx.x];
        const int4 vv = reinterpret_cast<const int4*>(s_oval)[threadIdx.x];
        key[0] = kk.x, key[1] = kk.y, key[2] = kk.z, key[3] = kk.w;
        val[0] = vv.x, val[1] = vv.y, val[2] = vv.z, val[3] = vv.w;
    }
    const float inv_tile_w = 1.0f / (float)tile_w;
    int keep = 0, mine = 0;
    // One strip evaluation per entry, straight-line (all four geometry records are requested before any is used); the
    // extent of a (Gaussian, tile row) pair is the same for every tile of the row, but a branch that reuses it across a
    // thread's four entries serialises the loads and measured slower.
    float4 ga[kPer], gb[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int64_t v = threadIdx.x * kPer + k < cnt ? val[k] : 0;  // (slots behind the block's count hold stale shared memory)
        ga[k] = ex.geom[v * 2];
        gb[k] = ex.geom[v * 2 + 1];
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int t = threadIdx.x * kPer + k;
        const int tile = (int)(key[k] & tile_mask);
        int ty = (int)(((float)tile + 0.5f) * inv_tile_w);  // exact below 2^22 tiles; corrected below anyway
        int tx = tile - ty * tile_w;
        if (tx < 0) {
            --ty;
            tx += tile_w;
        } else if (tx >= tile_w) {
            ++ty;
            tx -= tile_w;
        }
        const float ya = (float)(ty * (int)tile_size) + 0.5f;
        const float yb = (float)min(ty * (int)tile_size + (int)tile_size - 1, ex.height - 1) + 0.5f;
        float xlo, xhi;
        const bool strip_hit = ellipse_strip_x_extent(ga[k].x, ga[k].y, 0.5f * kLog2e * gb[k].x, kLog2e * gb[k].y, 0.5f * kLog2e * gb[k].z,
                                                      __log2f(ga[k].z) + kLog2_255, ya, yb, xlo, xhi);
        const float xa = (float)(tx * (int)tile_size) + 0.5f;
        const float xb = (float)min(tx * (int)tile_size + (int)tile_size - 1, ex.width - 1) + 0.5f;
        if (t < cnt && strip_hit && xa <= xhi && xb >= xlo) {
            keep |= 1 << k;
            ++mine;
        }
    }
    // exclusive scan of the per-thread counts over the block
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();  // also: everyone has read its entries from s_okey / s_oval
    int before = inc - mine, total = 0;
#pragma unroll
    for (int w = 0; w < kEmitThreads / 32; ++w) {
        if (w < warp) before += s_wsum[w];
        total += s_wsum[w];
    }
    if (threadIdx.x == 0) {
        ex.seg_counts[vb] = total;
        if (total) atomicAdd(ex.n_out, (unsigned long long)total);
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        if ((keep >> k) & 1) {
            s_okey[before] = key[k];
            s_oval[before] = val[k];
            ++before;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < total; t += kEmitThreads) {
        tkeys[e0 + t] = s_okey[t];
        tvals[e0 + t] = s_oval[t];
    }
}

// The exact emit leaves its survivors at the start of every block's own kEmitTile-slot segment.  Two small kernels make
// the list dense before the tile sort (ncu, 8 M Gaussians at 4K: a first radix pass that walks 154 M half-empty slots took
// 1.48 ms, the dense pass over the 69.5 M survivors 0.63 ms; the copy costs 1.1 GB of traffic):
//   seg_offsets_kernel : exclusive scan of the per-segment counts (one block, 8 counts per thread and round)
//   seg_compact_kernel : one block per segment copies its survivors to their dense position (coalesced both ways)
constexpr int kSegScanThreads = 1024, kSegScanPer = 8;
__global__ void __launch_bounds__(kSegScanThreads) seg_offsets_kernel(int64_t nb_cap, const int64_t* __restrict__ clamped, const int32_t* __restrict__ counts,
                                                                     int32_t* __restrict__ offsets) {
    __shared__ int s_warp[kSegScanThreads / 32];
    __shared__ int s_carry;
    pdl_enter();
    const int64_t nb = clamped ? min(nb_cap, (clamped[1] + kEmitTile - 1) / kEmitTile) : nb_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += kSegScanThreads * kSegScanPer) {
        const int64_t i0 = base + (int64_t)threadIdx.x * kSegScanPer;
        int v[kSegScanPer], sum = 0;
#pragma unroll
        for (int k = 0; k < kSegScanPer; ++k) {
            v[k] = i0 + k < nb ? counts[i0 + k] : 0;
            sum += v[k];
        }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int w = s_warp[lane];
            int winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        int run = s_carry + s_warp[warp] + inc - sum;
#pragma unroll
        for (int k = 0; k < kSegScanPer; ++k) {
            if (i0 + k < nb) offsets[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == kSegScanThreads - 1) s_carry = run;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) seg_compact_kernel(const int64_t* __restrict__ clamped, const int32_t* __restrict__ counts,
                                                         const int32_t* __restrict__ offsets, const uint32_t* __restrict__ keys_in,
                                                         const int32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                                                         int32_t* __restrict__ vals_out) {
    pdl_enter();
    const int64_t seg = blockIdx.x;
    if (clamped && seg * kEmitTile >= clamped[1]) return;
    const int n = counts[seg];
    const int64_t src = seg * kEmitTile, dst = offsets[seg];
    for (int t = threadIdx.x; t < n; t += 256) {
        keys_out[dst + t] = keys_in[src + t];
        vals_out[dst + t] = vals_in[src + t];
    }
}

// isect_ids = key << 32 | bits(depth), fused with the per-tile ranges (same rule as tile_ranges_kernel)
// n_dev (exact lists): the number of entries lives on the device and offsets gets one more element, the end
// of the last range, so that the compositor needs no host-side count.  Four consecutive entries per thread (one
// 16-byte key load): the kernel is a 28 MB stream and was latency-bound at one entry per thread.
constexpr int kComposePer = 4;
__global__ void __launch_bounds__(256) compose_ids_ranges_kernel(int64_t n_host, const int64_t* __restrict__ n_dev, const uint32_t* __restrict__ tkeys,
                                                                 const int32_t* __restrict__ flat, const float* __restrict__ depths, int C, int n_tiles,
                                                                 int tile_n_bits, int64_t* __restrict__ isect_ids, int32_t* __restrict__ offsets) {
    pdl_enter();
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kComposePer;
    const int64_t n = n_dev ? *n_dev : n_host;
    if (n_dev && offsets && i0 == 0) {
        offsets[(int64_t)C * n_tiles] = (int32_t)n;
        if (n == 0)
            for (int64_t t = 0; t < (int64_t)C * n_tiles; ++t) offsets[t] = 0;
    }
    if (i0 >= n) return;
    uint32_t key[kComposePer];
    if (i0 + kComposePer <= n) {
        const uint4 k4 = *reinterpret_cast<const uint4*>(tkeys + i0);  // i0 is a multiple of 4, the buffer 256-byte aligned
        key[0] = k4.x, key[1] = k4.y, key[2] = k4.z, key[3] = k4.w;
    } else {
#pragma unroll
        for (int k = 0; k < kComposePer; ++k) key[k] = i0 + k < n ? tkeys[i0 + k] : 0u;
    }
    const uint32_t mask = (1u << tile_n_bits) - 1u;
    uint32_t kp = i0 > 0 ? tkeys[i0 - 1] : 0u;
#pragma unroll
    for (int k = 0; k < kComposePer; ++k) {
        const int64_t i = i0 + k;
        if (i >= n) break;
        if (isect_ids) {
            const uint32_t db = (uint32_t)__float_as_int(depths[flat[i]]);
            isect_ids[i] = (int64_t)(((uint64_t)key[k] << 32) | db);
        }
        if (offsets) {
            const int64_t cur = (int64_t)(key[k] >> tile_n_bits) * n_tiles + (key[k] & mask);
            if (i == 0) {
                for (int64_t t = 0; t <= cur; ++t) offsets[t] = 0;
            } else if (key[k] != kp) {
                const int64_t prev = (int64_t)(kp >> tile_n_bits) * n_tiles + (kp & mask);
                for (int64_t t = prev + 1; t <= cur; ++t) offsets[t] = (int32_t)i;
            }
            if (i == n - 1) {
                for (int64_t t = cur + 1; t < (int64_t)C * n_tiles; ++t) offsets[t] = (int32_t)n;
            }
        }
        kp = key[k];
    }
}

static int bit_length(int64_t v) {
    int b = 0;
    while (((int64_t)1 << b) <= v) ++b;
    return b;
}

}  // namespace qed

extern "C" size_t qed_isect_prepare_workspace_bytes(int64_t CN) { return prepare_layout(CN > 0 ? CN : 1).total; }

extern "C" int qed_isect_prepare(int C, int N, const float* depths, const int32_t* tiles_per_gauss, void* workspace,
                                 size_t workspace_bytes, int64_t* counts_dev, int64_t* counts_host_pinned, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (C < 0 || N < 0 || !counts_dev) return QED_ERR_BAD_ARG;
    const int64_t CN = (int64_t)C * N;
    if (CN == 0) {
        QED_CUDA_TRY(cudaMemsetAsync(counts_dev, 0, 16, stream));
    } else {
        if (CN > 0x7fffffffLL) return QED_ERR_UNSUPPORTED;
        if (!depths || !tiles_per_gauss || !workspace) return QED_ERR_BAD_ARG;
        const PrepareLayout L = prepare_layout(CN);
        if (workspace_bytes < L.total) return QED_ERR_WORKSPACE;
        char* ws = reinterpret_cast<char*>(workspace);
        int32_t* vals0 = reinterpret_cast<int32_t*>(ws + L.vals[0]);
        int32_t* vals1 = reinterpret_cast<int32_t*>(ws + L.vals[1]);
        int32_t* vals2 = reinterpret_cast<int32_t*>(ws + L.vals[2]);
        int64_t* cum2 = reinterpret_cast<int64_t*>(ws + L.cum2);
        // 1. ordered compaction of the visible entries (scan of the flags, the compaction being its final phase;
        //    n_visible -> counts_dev[0]);  2. stable sort by (camera, depth bits): sorted flat indices land in vals1
        int rc;
        if (C == 1) {
            uint32_t* k0 = reinterpret_cast<uint32_t*>(ws + L.keys[0]);
            uint32_t* k1 = reinterpret_cast<uint32_t*>(ws + L.keys[1]);
            uint32_t* k2 = reinterpret_cast<uint32_t*>(ws + L.keys[2]);
            rc = scan_inclusive_to(CN, nullptr, ScanFlagPositive{tiles_per_gauss}, CompactVisibleSink<uint32_t>{N, depths, k0, vals0}, counts_dev,
                                   ws + L.scan_ws, stream);
            if (rc != QED_OK) return rc;
            rc = radix_sort_pairs<uint32_t>(CN, counts_dev, k0, vals0, k1, vals1, k2, vals2, ws + L.hist, 32, stream);
        } else {
            uint64_t* k0 = reinterpret_cast<uint64_t*>(ws + L.keys[0]);
            uint64_t* k1 = reinterpret_cast<uint64_t*>(ws + L.keys[1]);
            uint64_t* k2 = reinterpret_cast<uint64_t*>(ws + L.keys[2]);
            rc = scan_inclusive_to(CN, nullptr, ScanFlagPositive{tiles_per_gauss}, CompactVisibleSink<uint64_t>{N, depths, k0, vals0}, counts_dev,
                                   ws + L.scan_ws, stream);
            if (rc != QED_OK) return rc;
            rc = radix_sort_pairs<uint64_t>(CN, counts_dev, k0, vals0, k1, vals1, k2, vals2, ws + L.hist, 32 + bit_length(C - 1), stream);
        }
        if (rc != QED_OK) return rc;
        // 3. tile counts in depth order -> write offsets, n_isects -> counts_dev[1]
        rc = scan_inclusive(CN, counts_dev, ScanGather{tiles_per_gauss, vals1}, cum2, counts_dev + 1, ws + L.scan_ws, stream);
        if (rc != QED_OK) return rc;
    }
    if (counts_host_pinned) QED_CUDA_TRY(cudaMemcpyAsync(counts_host_pinned, counts_dev, 16, cudaMemcpyDeviceToHost, stream));
    return QED_OK;
}

extern "C" size_t qed_isect_fill_workspace_bytes(int64_t n_isects) {
    if (n_isects <= 0) return 256;
    const size_t nb = (size_t)((n_isects + kEmitTile - 1) / kEmitTile + 1);
    return 5 * align_up((size_t)n_isects * 4, 256) + radix_hist_bytes(n_isects) + align_up(nb * 4, 256) + align_up(nb * 8 + 8, 256) + 256;
}

extern "C" int qed_isect_fill(int C, int N, int64_t n_visible, int64_t n_isects, const float* means2d, const int32_t* radii,
                              const float* depths, const float* geom, int image_width, int image_height, int tile_size, int tile_width,
                              int tile_height, const void* prepare_workspace, void* workspace, size_t workspace_bytes,
                              const int64_t* counts_dev, int64_t* isect_ids, int32_t* flatten_ids, int32_t* isect_offsets,
                              int64_t* n_exact_dev, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (C < 0 || N < 0 || n_visible < 0 || n_isects < 0 || tile_size <= 0) return QED_ERR_BAD_ARG;
    // counts_dev: no host sync happened; n_visible / n_isects are capacities and the offsets always carry their end
    if (counts_dev && !isect_offsets) return QED_ERR_BAD_ARG;
    const int64_t CN = (int64_t)C * N;
    const int n_tiles = tile_width * tile_height;
    // exact tile lists are requested by passing n_exact_dev (geom may legitimately be NULL for an empty scene):
    // offsets has C * n_tiles + 1 elements, the count stays on the device
    const bool exact = n_exact_dev != nullptr;
    if (exact && (!isect_offsets || image_width <= 0 || image_height <= 0)) return QED_ERR_BAD_ARG;
    if (exact && n_isects > 0 && CN > 0 && (!geom || (reinterpret_cast<uintptr_t>(geom) & 15))) return QED_ERR_BAD_ARG;
    if (!exact && geom) return QED_ERR_BAD_ARG;  // geom without n_exact_dev: the caller would not learn the count
    if (n_isects == 0 || CN == 0) {
        if (isect_offsets && (int64_t)C * n_tiles > 0)
            QED_CUDA_TRY(cudaMemsetAsync(isect_offsets, 0, ((size_t)C * n_tiles + ((exact || counts_dev) ? 1 : 0)) * 4, stream));
        if (exact) QED_CUDA_TRY(cudaMemsetAsync(n_exact_dev, 0, 8, stream));
        return QED_OK;
    }
    if (n_isects > 0x7fffffffLL) return QED_ERR_UNSUPPORTED;
    if (!means2d || !radii || !depths || !prepare_workspace || !workspace || !flatten_ids) return QED_ERR_BAD_ARG;
    if (workspace_bytes < qed_isect_fill_workspace_bytes(n_isects)) return QED_ERR_WORKSPACE;
    const int tile_n_bits = bit_length(n_tiles);
    const int cam_bits = bit_length(C - 1);
    if (tile_n_bits + cam_bits > 32) return QED_ERR_UNSUPPORTED;
    const PrepareLayout L = prepare_layout(CN);
    const char* pws = reinterpret_cast<const char*>(prepare_workspace);
    const int32_t* sorted_vals = reinterpret_cast<const int32_t*>(pws + L.vals[1]);
    const int64_t* cum2 = reinterpret_cast<const int64_t*>(pws + L.cum2);
    char* ws = reinterpret_cast<char*>(workspace);
    const size_t seg = align_up((size_t)n_isects * 4, 256);
    uint32_t* k0 = reinterpret_cast<uint32_t*>(ws);
    uint32_t* k1 = reinterpret_cast<uint32_t*>(ws + seg);
    uint32_t* k2 = reinterpret_cast<uint32_t*>(ws + 2 * seg);
    int32_t* v0 = reinterpret_cast<int32_t*>(ws + 3 * seg);
    int32_t* v2 = reinterpret_cast<int32_t*>(ws + 4 * seg);
    void* hist = ws + 5 * seg;
    const size_t nb = (size_t)((n_isects + kEmitTile - 1) / kEmitTile);
    int32_t* first_j = reinterpret_cast<int32_t*>(ws + 5 * seg + radix_hist_bytes(n_isects));
    int32_t* seg_counts = reinterpret_cast<int32_t*>(ws + 5 * seg + radix_hist_bytes(n_isects) + align_up((nb + 1) * 4, 256));
    int64_t* clamped = counts_dev ? reinterpret_cast<int64_t*>(ws + 5 * seg + radix_hist_bytes(n_isects) + align_up((nb + 1) * 4, 256) +
                                                               align_up((nb + 1) * 8 + 8, 256))
                                  : nullptr;
    if (n_visible == 0) return QED_ERR_BAD_ARG;  // n_isects > 0 needs at least one visible entry (capacity when counts_dev)
    QED_CUDA_TRY(launch_pdl(emit_boundaries_kernel, dim3((unsigned)((n_visible + 255) / 256)), dim3(256), 0, stream, n_visible, n_isects, counts_dev,
                            clamped, cum2, first_j));
    ExactEmit ex{};
    if (exact) {
        QED_CUDA_TRY(cudaMemsetAsync(n_exact_dev, 0, 8, stream));
        ex.geom = reinterpret_cast<const float4*>(geom);
        ex.seg_counts = seg_counts;
        ex.n_out = reinterpret_cast<unsigned long long*>(n_exact_dev);
        ex.width = image_width;
        ex.height = image_height;
        QED_CUDA_TRY(launch_pdl(emit_sorted_kernel<true>, dim3((unsigned)nb), dim3(kEmitThreads), 0, stream, n_visible, n_isects, clamped, N, sorted_vals,
                                cum2, first_j, reinterpret_cast<const float2*>(means2d), radii, (float)tile_size, tile_width, tile_height,
                                tile_n_bits, k0, v0, ex));
    } else {
        QED_CUDA_TRY(launch_pdl(emit_sorted_kernel<false>, dim3((unsigned)nb), dim3(kEmitThreads), 0, stream, n_visible, n_isects, clamped, N, sorted_vals,
                                cum2, first_j, reinterpret_cast<const float2*>(means2d), radii, (float)tile_size, tile_width, tile_height,
                                tile_n_bits, k0, v0, ex));
    }
    const int64_t* n_dev = exact ? n_exact_dev : (clamped ? clamped + 1 : nullptr);
    int rc;
    if (exact) {
        // survivors sit at the start of every emit block's segment: make the list dense (k2 / v2), then sort it
        int32_t* seg_offsets = seg_counts + (nb + 1);
        QED_CUDA_TRY(launch_pdl(seg_offsets_kernel, dim3(1), dim3(kSegScanThreads), 0, stream, (int64_t)nb, (const int64_t*)clamped,
                                (const int32_t*)seg_counts, seg_offsets));
        QED_CUDA_TRY(launch_pdl(seg_compact_kernel, dim3((unsigned)nb), dim3(256), 0, stream, (const int64_t*)clamped, (const int32_t*)seg_counts,
                                (const int32_t*)seg_offsets, (const uint32_t*)k0, (const int32_t*)v0, k2, v2));
        rc = radix_sort_pairs<uint32_t>(n_isects, n_dev, k2, v2, k1, flatten_ids, k0, v0, hist, tile_n_bits + cam_bits, stream);
    } else {
        rc = radix_sort_pairs<uint32_t>(n_isects, n_dev, k0, v0, k1, flatten_ids, k2, v2, hist, tile_n_bits + cam_bits, stream);
    }
    if (rc != QED_OK) return rc;
    QED_CUDA_TRY(launch_pdl(compose_ids_ranges_kernel, dim3((unsigned)((n_isects + 256 * kComposePer - 1) / (256 * kComposePer))), dim3(256), 0, stream,
                            n_isects, n_dev, k1, flatten_ids, depths, C, n_tiles, tile_n_bits, isect_ids, isect_offsets));
    return QED_OK;
}
