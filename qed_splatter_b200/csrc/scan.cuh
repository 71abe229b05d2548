// Three-phase inclusive scan (reduce / scan of block sums / scan + add) of a transformed int32 input to
// int64, with the element count optionally read from device memory (so upstream counts need no host sync).
#pragma once
#include "common.cuh"

namespace qed {

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

// input transforms
struct ScanIdentity {
    const int32_t* in;
    __device__ __forceinline__ int32_t operator()(int64_t i) const { return in[i]; }
};
struct ScanFlagPositive {  // 1 where in[i] > 0
    const int32_t* in;
    __device__ __forceinline__ int32_t operator()(int64_t i) const { return in[i] > 0 ? 1 : 0; }
};
struct ScanGather {  // in[index[i]]
    const int32_t* in;
    const int32_t* index;
    __device__ __forceinline__ int32_t operator()(int64_t i) const { return in[index[i]]; }
};

template <typename F>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(int64_t n_host, const int64_t* n_dev, F f, int64_t* __restrict__ block_sums) {
    __shared__ int64_t sw[kScanThreads / 32 + 1];
    pdl_enter();
    const int64_t n = n_dev ? *n_dev : n_host;
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    int64_t s = 0;
    if (base < n) {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            int64_t i = base + k * kScanThreads + threadIdx.x;
            if (i < n) s += f(i);
        }
    }
    int64_t total;
    block_exclusive_scan(s, sw, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of block sums in place; writes the grand total
static __global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(int64_t nb, int64_t* __restrict__ block_sums, int64_t* __restrict__ total_dev) {
    __shared__ int64_t sw[kScanThreads / 32 + 1];
    pdl_enter();
    int64_t carry = 0;
    for (int64_t base = 0; base < nb; base += kScanThreads) {
        int64_t i = base + threadIdx.x;
        int64_t v = i < nb ? block_sums[i] : 0;
        int64_t total;
        int64_t ex = block_exclusive_scan(v, sw, total);
        if (i < nb) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_dev) *total_dev = carry;
}

// output sinks of the final phase: store the inclusive scan, or consume (index, value, inclusive scan) directly
struct ScanStore {
    int64_t* out;
    __device__ __forceinline__ void operator()(int64_t i, int32_t, int64_t inc) const { out[i] = inc; }
};

template <typename F, typename Sink>
__global__ void __launch_bounds__(kScanThreads) scan_final_kernel(int64_t n_host, const int64_t* n_dev, F f, const int64_t* __restrict__ block_sums,
                                                                 Sink sink) {
    __shared__ int64_t sw[kScanThreads / 32 + 1];
    pdl_enter();
    const int64_t n = n_dev ? *n_dev : n_host;
    // blocked arrangement: thread t owns items [t*kScanItems, (t+1)*kScanItems) of the tile
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    if ((int64_t)blockIdx.x * kScanTile >= n) return;
    int32_t v[kScanItems];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? f(base + k) : 0;
        s += v[k];
    }
    int64_t total;
    int64_t ex = block_exclusive_scan(s, sw, total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        ex += v[k];
        if (base + k < n) sink(base + k, v[k], ex);
    }
}

inline size_t scan_workspace_bytes(int64_t capacity) {
    int64_t nb = (capacity + kScanTile - 1) / kScanTile;
    return (size_t)(nb > 0 ? nb : 1) * sizeof(int64_t);
}

// Inclusive scan of f(0..n) into out (int64); total -> *total_dev.  `capacity` >= n sizes the grid when n
// is only known on the device (n_dev != nullptr).
// `sink(i, f(i), inclusive_scan(i))` is called for every element in the final phase (ScanStore: write it out).
template <typename F, typename Sink>
inline int scan_inclusive_to(int64_t capacity, const int64_t* n_dev, F f, Sink sink, int64_t* total_dev, void* workspace, cudaStream_t stream) {
    int64_t nb = (capacity + kScanTile - 1) / kScanTile;
    if (nb == 0) nb = 1;
    int64_t* sums = reinterpret_cast<int64_t*>(workspace);
    QED_CUDA_TRY(launch_pdl(scan_reduce_kernel<F>, dim3((unsigned)nb), dim3(kScanThreads), 0, stream, capacity, n_dev, f, sums));
    QED_CUDA_TRY(launch_pdl(scan_sums_kernel, dim3(1), dim3(kScanThreads), 0, stream, nb, sums, total_dev));
    QED_CUDA_TRY(launch_pdl(scan_final_kernel<F, Sink>, dim3((unsigned)nb), dim3(kScanThreads), 0, stream, capacity, n_dev, f, sums, sink));
    return QED_OK;
}

// ---- single-launch variant ("flat" scan) ---------------------------------------------------------------------------------
// The three launches above cost ~23 us for 1M flags (489 blocks of 2048: launch + drain three times, 12 us of it in kernels
// that are 10-25 % issue-active).  Here at most kFlatScanBlocks blocks (<= what is co-resident: 148 SMs x 8 blocks of 256
// threads) each own a contiguous range of tiles -- ONE tile up to 2M elements, kept in registers: the range is summed, the
// sum is published as ONE 64-bit word (bit 63 = ready), every block then adds up the words of ALL its predecessors in
// parallel (a few per thread -- no look-back chain, which at this size costs a full walk because every block publishes at
// the same moment), and the scan is emitted.  Blocks take their range from a ticket, so a block only ever waits on blocks
// that are already running.  `state` must be zero when the kernel starts (FlatScanState, kFlatScanStateBytes).
constexpr int kFlatScanBlocks = 1024;
struct FlatScanState {
    unsigned long long agg[kFlatScanBlocks];
    unsigned int ticket, pad;
};
constexpr size_t kFlatScanStateBytes = sizeof(FlatScanState);

template <typename F, typename Sink>
__global__ void __launch_bounds__(kScanThreads) scan_flat_kernel(int64_t n_host, const int64_t* n_dev, F f, Sink sink, FlatScanState* __restrict__ st,
                                                                int64_t* __restrict__ total_dev, int tiles_per_block) {
    __shared__ int64_t sw[kScanThreads / 32 + 1];
    __shared__ unsigned int s_vb;
    pdl_enter();
    const int64_t n = n_dev ? min(*n_dev, n_host) : n_host;
    if (threadIdx.x == 0) s_vb = atomicAdd(&st->ticket, 1u);
    __syncthreads();
    const unsigned int vb = s_vb;
    const int64_t tile0 = (int64_t)vb * tiles_per_block;
    // blocked arrangement: thread t owns kScanItems consecutive items of a tile
    int32_t v[kScanItems];
    int64_t s = 0, ex0 = 0, total;
    if (tiles_per_block == 1) {  // the tile stays in registers
        const int64_t base = tile0 * kScanTile + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = (base + k < n) ? f(base + k) : 0;
            s += v[k];
        }
        ex0 = block_exclusive_scan(s, sw, total);
    } else {  // pass 1: sum of the range (strided: coalesced)
        for (int t = 0; t < tiles_per_block; ++t) {
            const int64_t base = (tile0 + t) * kScanTile;
            if (base >= n) break;
#pragma unroll
            for (int k = 0; k < kScanItems; ++k) {
                const int64_t i = base + k * kScanThreads + threadIdx.x;
                if (i < n) s += f(i);
            }
        }
        block_exclusive_scan(s, sw, total);
    }
    if (threadIdx.x == 0) {
        volatile unsigned long long* slot = st->agg + vb;
        *slot = (1ull << 63) | (unsigned long long)total;
    }
    // sum of all predecessors' totals: thread t waits for the words of blocks t, t + 256, ...
    int64_t mine = 0;
    for (unsigned int b = threadIdx.x; b < vb; b += kScanThreads) {
        volatile const unsigned long long* slot = st->agg + b;
        unsigned long long w;
        do {
            w = *slot;
        } while (!(w >> 63));
        mine += (int64_t)(w & ~(1ull << 63));
    }
    int64_t prefix;
    block_exclusive_scan(mine, sw, prefix);  // `prefix` = block-wide total of `mine`
    if (vb == gridDim.x - 1 && threadIdx.x == 0 && total_dev) *total_dev = prefix + total;
    if (tiles_per_block == 1) {
        const int64_t base = tile0 * kScanTile + (int64_t)threadIdx.x * kScanItems;
        int64_t ex = ex0 + prefix;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            ex += v[k];
            if (base + k < n) sink(base + k, v[k], ex);
        }
        return;
    }
    // pass 2: re-read (L2 hits), running carry over the tiles
    int64_t carry = prefix;
    for (int t = 0; t < tiles_per_block; ++t) {
        const int64_t tbase = (tile0 + t) * kScanTile;
        if (tbase >= n) break;
        const int64_t base = tbase + (int64_t)threadIdx.x * kScanItems;
        int64_t ts = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = (base + k < n) ? f(base + k) : 0;
            ts += v[k];
        }
        int64_t tile_total;
        int64_t ex = block_exclusive_scan(ts, sw, tile_total) + carry;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            ex += v[k];
            if (base + k < n) sink(base + k, v[k], ex);
        }
        carry += tile_total;
    }
}

// Same contract as scan_inclusive_to; `state` = kFlatScanStateBytes of ZEROED device memory (consumed by this launch).
template <typename F, typename Sink>
inline int scan_flat_to(int64_t capacity, const int64_t* n_dev, F f, Sink sink, int64_t* total_dev, void* state, cudaStream_t stream) {
    int64_t tiles = (capacity + kScanTile - 1) / kScanTile;
    if (tiles == 0) tiles = 1;
    const int blocks = (int)(tiles < kFlatScanBlocks ? tiles : kFlatScanBlocks);
    const int tiles_per_block = (int)((tiles + blocks - 1) / blocks);
    const int used = (int)((tiles + tiles_per_block - 1) / tiles_per_block);  // no trailing block without a tile
    QED_CUDA_TRY(launch_pdl(scan_flat_kernel<F, Sink>, dim3((unsigned)used), dim3(kScanThreads), 0, stream, capacity, n_dev, f, sink,
                            reinterpret_cast<FlatScanState*>(state), total_dev, tiles_per_block));
    return QED_OK;
}

template <typename F>
inline int scan_inclusive(int64_t capacity, const int64_t* n_dev, F f, int64_t* out, int64_t* total_dev, void* workspace, cudaStream_t stream) {
    return scan_inclusive_to(capacity, n_dev, f, ScanStore{out}, total_dev, workspace, stream);
}

}  // namespace qed
