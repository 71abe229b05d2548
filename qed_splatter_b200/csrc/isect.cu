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

extern "C" int qed_debug_set_radix_small_tiles(int enabled) {
    int old = g_radix_small_tiles;
    g_radix_small_tiles = enabled ? 1 : 0;
    return old;
}

// 1 (default): the two scans of qed_isect_prepare run as single launches (scan_flat_to); 0: three launches each.
static thread_local int g_flat_scan = 1;
extern "C" int qed_debug_set_flat_scan(int enabled) {
    int old = g_flat_scan;
    g_flat_scan = enabled ? 1 : 0;
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
    size_t scan_ws, scan_state, keys[3], vals[3], hist, cum2, total;
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
    L.scan_state = take(2 * kFlatScanStateBytes);  // one per scan of qed_isect_prepare, zeroed by one memset
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
    }
    if (j == 0) {  // the counts the following kernels work with, on the device (also when the host knows them)
        clamped[0] = n_vis;
        clamped[1] = n_isects;
    }
    if (j >= n_vis) return;
    const int64_t start = j > 0 ? cum2[j - 1] : 0, end = cum2[j];
    for (int64_t b = (start + kEmitTile - 1) / kEmitTile; b * kEmitTile < end; ++b) first_j[b] = (int32_t)j;
}

// gsplat's lists: every tile of each Gaussian's 3-sigma bounding box, emitted in depth order.
__global__ void __launch_bounds__(kEmitThreads) emit_sorted_kernel(int64_t n_vis, int64_t n_isects, const int64_t* __restrict__ clamped, int N,
                                                                  const int32_t* __restrict__ sorted_vals,
                                                                  const int64_t* __restrict__ cum2, const int32_t* __restrict__ first_j,
                                                                  const float2* __restrict__ means2d, const int32_t* __restrict__ radii,
                                                                  float tile_size, int tile_w, int tile_h, int tile_n_bits,
                                                                  uint32_t* __restrict__ tkeys, int32_t* __restrict__ tvals) {
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
        if (e0 >= n_isects) return;
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
    for (int t = threadIdx.x; t < (int)(e1 - e0); t += kEmitThreads) {
        tkeys[e0 + t] = s_okey[t];
        tvals[e0 + t] = s_oval[t];
    }
}

// EXACT tile lists for the compositor: only the tiles a Gaussian can reach with alpha >= 1/255 at some pixel centre.
// The projection kernel already COUNTED them (exact_row_span summed over the rows of the bounding box; that count was
// scanned into cum2), so this kernel only ENUMERATES the same spans: no per-tile test, no compaction, dense output.
// About half of gsplat's entries (S1: 53 %) never exist, which halves the tile sort, the range pass and the compositors'
// gathers and changes no pixel.  Same entry-parallel layout as emit_sorted_kernel: a block owns kEmitTile consecutive
// output entries and stages the Gaussians that overlap them;
//   small pieces (<= 32 entries, bounding box <= 8 tile rows): one lane walks the rows of its Gaussian;
//   large pieces: one warp per Gaussian, 32 rows at a time -- every lane evaluates one row's span, a warp prefix sum
//   gives the rows' positions in the Gaussian's list and every lane writes its own row's part of the piece.
__global__ void __launch_bounds__(kEmitThreads) emit_exact_kernel(int64_t n_vis, int64_t n_isects, const int64_t* __restrict__ clamped, int N,
                                                                 const int32_t* __restrict__ sorted_vals, const int64_t* __restrict__ cum2,
                                                                 const int32_t* __restrict__ first_j, const float2* __restrict__ means2d,
                                                                 const int32_t* __restrict__ radii, const float4* __restrict__ geom, int height,
                                                                 float tile_size, int tile_w, int tile_h, int tile_n_bits,
                                                                 uint32_t* __restrict__ tkeys, int32_t* __restrict__ tvals) {
    __shared__ int16_t s_lo[kEmitTile + 1], s_hi[kEmitTile + 1], s_large[kEmitTile + 1];
    __shared__ int32_t s_k0[kEmitTile + 1], s_idx[kEmitTile + 1];
    __shared__ __align__(16) uint32_t s_okey[kEmitTile];
    __shared__ __align__(16) int32_t s_oval[kEmitTile];
    __shared__ int s_nlarge;
    pdl_enter();
    const uint32_t vb = blockIdx.x;
    const int64_t e0 = (int64_t)vb * kEmitTile;
    if (clamped) {
        n_vis = clamped[0];
        n_isects = clamped[1];
    }
    if (e0 >= n_isects) return;
    const int64_t e1 = min(e0 + kEmitTile, n_isects);
    const int64_t j0 = first_j[vb];
    int64_t j1;  // Gaussian that owns entry e1 - 1
    if (e1 >= n_isects) {
        j1 = n_vis - 1;
    } else {
        const int64_t jn = first_j[vb + 1];
        j1 = (jn > 0 && cum2[jn - 1] == e1) ? jn - 1 : jn;
    }
    const int G = (int)(j1 - j0) + 1;  // every listed Gaussian has >= 1 entry, so G <= kEmitTile
    if (threadIdx.x == 0) s_nlarge = 0;
    // (a count / enumeration mismatch must never leave stale shared memory in the output: keys index the range table)
    for (int t = threadIdx.x; t < kEmitTile; t += kEmitThreads) {
        s_okey[t] = 0u;
        s_oval[t] = 0;
    }
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
        s_idx[g] = idx;
        if (hi - lo > 32 || tb.y1 - tb.y0 > 8) s_large[atomicAdd(&s_nlarge, 1)] = (int16_t)g;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // small pieces: one lane per Gaussian walks the rows of its bounding box
    for (int g = threadIdx.x; g < G; g += kEmitThreads) {
        const int idx = s_idx[g];
        int pos = s_lo[g], remaining = s_hi[g] - pos, skip = s_k0[g];
        const float2 m = means2d[idx];
        const TileBox tb = tile_box(m.x, m.y, radii[idx], tile_size, tile_w, tile_h);
        if (remaining > 32 || tb.y1 - tb.y0 > 8) continue;
        const float4 ga = geom[(int64_t)idx * 2], gb = geom[(int64_t)idx * 2 + 1];
        const ExactCtx ec = exact_ctx(ga.x, ga.y, ga.z, gb.x, gb.y, gb.z);
        const uint32_t cam_bits = (uint32_t)(idx / N) << tile_n_bits;
        for (int ty = tb.y0; ty < tb.y1 && remaining > 0; ++ty) {
            int c0, c1;
            exact_row_span(ec, ty, height, tb.x0, tb.x1, c0, c1);
            const int w = c1 - c0;
            if (skip >= w) {
                skip -= w;
                continue;
            }
            for (int cx = c0 + skip; cx < c1 && remaining > 0; ++cx) {
                s_okey[pos] = cam_bits | (uint32_t)(ty * tile_w + cx);
                s_oval[pos] = idx;
                ++pos;
                --remaining;
            }
            skip = 0;
        }
    }
    // large pieces: one warp per Gaussian, 32 rows per round
    const int nlarge = s_nlarge;
    for (int i = warp; i < nlarge; i += kEmitThreads / 32) {
        const int g = s_large[i];
        const int idx = s_idx[g];
        const int lo = s_lo[g], n = s_hi[g] - lo, k0 = s_k0[g];
        const float2 m = means2d[idx];
        const TileBox tb = tile_box(m.x, m.y, radii[idx], tile_size, tile_w, tile_h);
        const float4 ga = geom[(int64_t)idx * 2], gb = geom[(int64_t)idx * 2 + 1];
        const ExactCtx ec = exact_ctx(ga.x, ga.y, ga.z, gb.x, gb.y, gb.z);
        const uint32_t cam_bits = (uint32_t)(idx / N) << tile_n_bits;
        int base = 0;  // entries of the rows before this round
        for (int r0 = tb.y0; r0 < tb.y1 && base < k0 + n; r0 += 32) {
            const int ty = r0 + lane;
            int c0 = 0, c1 = 0;
            if (ty < tb.y1) exact_row_span(ec, ty, height, tb.x0, tb.x1, c0, c1);
            const int w = c1 - c0;
            int inc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            const int row_start = base + inc - w;  // position of this lane's row in the Gaussian's list
            const int a = max(row_start, k0), b = min(row_start + w, k0 + n);
            for (int q = a; q < b; ++q) {
                s_okey[lo + (q - k0)] = cam_bits | (uint32_t)(ty * tile_w + c0 + (q - row_start));
                s_oval[lo + (q - k0)] = idx;
            }
            base += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < (int)(e1 - e0); t += kEmitThreads) {
        tkeys[e0 + t] = s_okey[t];
        tvals[e0 + t] = s_oval[t];
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
        char* st0 = ws + L.scan_state;
        char* st1 = st0 + kFlatScanStateBytes;
        const bool flat = g_flat_scan != 0;
        if (flat) QED_CUDA_TRY(cudaMemsetAsync(st0, 0, 2 * kFlatScanStateBytes, stream));
        if (C == 1) {
            uint32_t* k0 = reinterpret_cast<uint32_t*>(ws + L.keys[0]);
            uint32_t* k1 = reinterpret_cast<uint32_t*>(ws + L.keys[1]);
            uint32_t* k2 = reinterpret_cast<uint32_t*>(ws + L.keys[2]);
            rc = flat ? scan_flat_to(CN, nullptr, ScanFlagPositive{tiles_per_gauss}, CompactVisibleSink<uint32_t>{N, depths, k0, vals0}, counts_dev, st0, stream)
                      : scan_inclusive_to(CN, nullptr, ScanFlagPositive{tiles_per_gauss}, CompactVisibleSink<uint32_t>{N, depths, k0, vals0}, counts_dev,
                                          ws + L.scan_ws, stream);
            if (rc != QED_OK) return rc;
            rc = radix_sort_pairs<uint32_t>(CN, counts_dev, k0, vals0, k1, vals1, k2, vals2, ws + L.hist, 32, stream);
        } else {
            uint64_t* k0 = reinterpret_cast<uint64_t*>(ws + L.keys[0]);
            uint64_t* k1 = reinterpret_cast<uint64_t*>(ws + L.keys[1]);
            uint64_t* k2 = reinterpret_cast<uint64_t*>(ws + L.keys[2]);
            rc = flat ? scan_flat_to(CN, nullptr, ScanFlagPositive{tiles_per_gauss}, CompactVisibleSink<uint64_t>{N, depths, k0, vals0}, counts_dev, st0, stream)
                      : scan_inclusive_to(CN, nullptr, ScanFlagPositive{tiles_per_gauss}, CompactVisibleSink<uint64_t>{N, depths, k0, vals0}, counts_dev,
                                          ws + L.scan_ws, stream);
            if (rc != QED_OK) return rc;
            rc = radix_sort_pairs<uint64_t>(CN, counts_dev, k0, vals0, k1, vals1, k2, vals2, ws + L.hist, 32 + bit_length(C - 1), stream);
        }
        if (rc != QED_OK) return rc;
        // 3. tile counts in depth order -> write offsets, n_isects -> counts_dev[1]
        rc = flat ? scan_flat_to(CN, counts_dev, ScanGather{tiles_per_gauss, vals1}, ScanStore{cum2}, counts_dev + 1, st1, stream)
                  : scan_inclusive(CN, counts_dev, ScanGather{tiles_per_gauss, vals1}, cum2, counts_dev + 1, ws + L.scan_ws, stream);
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
    // exact tile lists (geom != NULL: qed_isect_prepare ran on the EXACT tile counts of qed_project_fwd): offsets has
    // C * n_tiles + 1 elements; n_exact_dev (optional) receives the entry count on the device
    const bool exact = geom != nullptr;
    // (an empty scene has no geom buffer: a caller that passes n_exact_dev still expects the end element)
    const bool has_end = exact || counts_dev != nullptr || n_exact_dev != nullptr;
    if (exact && (!isect_offsets || image_width <= 0 || image_height <= 0 || (reinterpret_cast<uintptr_t>(geom) & 15) || tile_size != 16))
        return QED_ERR_BAD_ARG;
    if (n_isects == 0 || CN == 0) {
        if (isect_offsets && (int64_t)C * n_tiles > 0)
            QED_CUDA_TRY(cudaMemsetAsync(isect_offsets, 0, ((size_t)C * n_tiles + (has_end ? 1 : 0)) * 4, stream));
        if (n_exact_dev) QED_CUDA_TRY(cudaMemsetAsync(n_exact_dev, 0, 8, stream));
        return QED_OK;
    }
    if (n_isects > 0x7fffffffLL) return QED_ERR_UNSUPPORTED;
    if (n_exact_dev && !exact && !counts_dev && !isect_offsets) return QED_ERR_BAD_ARG;
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
    int64_t* clamped = reinterpret_cast<int64_t*>(ws + 5 * seg + radix_hist_bytes(n_isects) + align_up((nb + 1) * 4, 256));
    if (n_visible == 0) return QED_ERR_BAD_ARG;  // n_isects > 0 needs at least one visible entry (capacity when counts_dev)
    QED_CUDA_TRY(launch_pdl(emit_boundaries_kernel, dim3((unsigned)((n_visible + 255) / 256)), dim3(256), 0, stream, n_visible, n_isects, counts_dev,
                            clamped, cum2, first_j));
    const int64_t* cl = counts_dev ? clamped : nullptr;  // kernels take the counts from the device only when the host does not know them
    if (exact) {
        QED_CUDA_TRY(launch_pdl(emit_exact_kernel, dim3((unsigned)nb), dim3(kEmitThreads), 0, stream, n_visible, n_isects, cl, N, sorted_vals, cum2,
                                first_j, reinterpret_cast<const float2*>(means2d), radii, reinterpret_cast<const float4*>(geom), image_height,
                                (float)tile_size, tile_width, tile_height, tile_n_bits, k0, v0));
    } else {
        QED_CUDA_TRY(launch_pdl(emit_sorted_kernel, dim3((unsigned)nb), dim3(kEmitThreads), 0, stream, n_visible, n_isects, cl, N, sorted_vals, cum2,
                                first_j, reinterpret_cast<const float2*>(means2d), radii, (float)tile_size, tile_width, tile_height, tile_n_bits,
                                k0, v0));
    }
    const int64_t* n_dev = has_end ? clamped + 1 : nullptr;  // (always written by emit_boundaries_kernel)
    int rc = radix_sort_pairs<uint32_t>(n_isects, n_dev, k0, v0, k1, flatten_ids, k2, v2, hist, tile_n_bits + cam_bits, stream);
    if (rc != QED_OK) return rc;
    QED_CUDA_TRY(launch_pdl(compose_ids_ranges_kernel, dim3((unsigned)((n_isects + 256 * kComposePer - 1) / (256 * kComposePer))), dim3(256), 0, stream,
                            n_isects, n_dev, k1, flatten_ids, depths, C, n_tiles, tile_n_bits, isect_ids, isect_offsets));
    if (n_exact_dev) QED_CUDA_TRY(cudaMemcpyAsync(n_exact_dev, clamped + 1, 8, cudaMemcpyDeviceToDevice, stream));
    return QED_OK;
}
