// Own LSD radix sort of (key, int32 value) pairs, 8-bit digits, stable.
//
// Per pass: upsweep (per-block digit histogram) -> scan (one block per digit over the per-block counts) ->
// downsweep (block-local stable rank via __match_any_sync, block-sorted staging in shared memory, then
// coalesced run writes).  KeyT is uint32_t or uint64_t.  The element count may live in device memory
// (n_dev), so a sort can follow a device-side compaction without a host sync.
#pragma once
#include "common.cuh"

namespace qed {

static thread_local int g_radix_small_tiles = 0;  // test hook (qed_debug_set_radix_small_tiles): half-size tiles for small look-back sorts (measured SLOWER, off)
static thread_local int g_radix_onesweep = 1;  // test hook (qed_debug_set_radix_onesweep), thread-local: 0 = three kernels per pass

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 pairs per block
constexpr int kRadix = 256;

// Segmented input of the FIRST pass (exact tile lists): the input consists of segments of (1 << shift) slots of
// which only the first counts[s] hold pairs (a block-local compaction upstream, no global one).  The pass
// scatters by global digit offsets, so its output -- and every later pass -- is dense.
struct SegCounts {
    const int32_t* counts = nullptr;
    int shift = 0;
    __device__ __forceinline__ bool ok(int64_t i) const { return !counts || (int)(i & ((1 << shift) - 1)) < counts[i >> shift]; }
};

template <typename KeyT, bool SEG>
__global__ void __launch_bounds__(kSortThreads) radix_upsweep_kernel(int64_t n_host, const int64_t* n_dev, const KeyT* __restrict__ keys, int shift,
                                                                    uint32_t mask, int nblocks, uint32_t* __restrict__ hist /* [kRadix][nblocks] */, SegCounts seg) {
    __shared__ uint32_t sh[kRadix];
    pdl_enter();
    const int64_t n = n_dev ? *n_dev : n_host;
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    if (base < n) {
        // all of the tile's keys are requested before the first shared-memory atomic (ncu at 70 M pairs: 23 % issue-active,
        // long_scoreboard 44 cycles per issue with the loads interleaved four at a time)
        KeyT key[kSortItems];
        bool ok[kSortItems];
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            const int64_t i = base + k * kSortThreads + threadIdx.x;
            ok[k] = i < n && (!SEG || seg.ok(i));
            key[k] = ok[k] ? keys[i] : (KeyT)0;
        }
#pragma unroll
        for (int k = 0; k < kSortItems; ++k)
            if (ok[k]) atomicAdd(&sh[(uint32_t)(key[k] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = sh[threadIdx.x];
}

// grid = kRadix blocks: block d scans hist[d][0..nblocks) exclusively in place, totals[d] = digit count.
// Four consecutive counters per thread and round (the row of a 150 M-pair sort has 37 k entries: at one per thread and
// round this kernel took 133 us per pass).
constexpr int kRadixScanPer = 4;
__global__ void __launch_bounds__(256) radix_scan_kernel(int nblocks, uint32_t* __restrict__ hist, uint32_t* __restrict__ totals) {
    __shared__ uint32_t sw[9];
    pdl_enter();
    uint32_t* row = hist + (int64_t)blockIdx.x * nblocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0;
    for (int base = 0; base < nblocks; base += 256 * kRadixScanPer) {
        const int i0 = base + threadIdx.x * kRadixScanPer;
        uint32_t v[kRadixScanPer], sum = 0;
#pragma unroll
        for (int k = 0; k < kRadixScanPer; ++k) {
            v[k] = i0 + k < nblocks ? row[i0 + k] : 0u;
            sum += v[k];
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) sw[warp] = inc;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (int w = 0; w < 8; ++w) {
                uint32_t t = sw[w];
                sw[w] = acc;
                acc += t;
            }
            sw[8] = acc;
        }
        __syncthreads();
        uint32_t run = carry + sw[warp] + inc - sum;
#pragma unroll
        for (int k = 0; k < kRadixScanPer; ++k) {
            if (i0 + k < nblocks) row[i0 + k] = run;
            run += v[k];
        }
        carry += sw[8];
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// Lanes holding the same 8-bit digit, from 8 ballots.  ncu showed MATCH.ANY dominating the downsweep's stall
// samples (~55 %: its latency grows with the number of distinct values in the warp); ballots are cheap and pipeline.
__device__ __forceinline__ uint32_t peers_by_ballot(uint32_t d, bool valid) {
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return valid ? peers : 0u;
}

template <typename KeyT>
struct RadixSmem {
    static constexpr size_t kBytes = (size_t)kSortTile * (sizeof(KeyT) + 4) + (size_t)(kSortThreads / 32) * kRadix * 4 + 2 * kRadix * 4 + 16 * 4;
};

template <typename KeyT, bool SEG>
__global__ void __launch_bounds__(kSortThreads, 3) radix_downsweep_kernel(int64_t n_host, const int64_t* n_dev, const KeyT* __restrict__ keys_in,
                                                                      const int32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
                                                                      int32_t* __restrict__ vals_out, int shift, uint32_t mask, int nblocks,
                                                                      const uint32_t* __restrict__ hist, const uint32_t* __restrict__ totals,
                                                                      uint32_t* __restrict__ next_hist, int next_shift, uint32_t next_mask, SegCounts seg) {
    constexpr int kWarps = kSortThreads / 32;
    constexpr int kPerWarp = kSortTile / kWarps;  // 512 consecutive pairs per warp
    __shared__ int s_tile_pairs;
    constexpr int kRounds = kPerWarp / 32;        // 16
    extern __shared__ __align__(16) unsigned char sort_smem[];
    KeyT* skeys = reinterpret_cast<KeyT*>(sort_smem);                                         // [kSortTile]
    int32_t* svals = reinterpret_cast<int32_t*>(skeys + kSortTile);                           // [kSortTile]
    uint32_t(*warp_hist)[kRadix] = reinterpret_cast<uint32_t(*)[kRadix]>(svals + kSortTile);  // [kWarps][kRadix]
    uint32_t* digit_start = &warp_hist[0][0] + kWarps * kRadix;  // start of each digit's run inside the block-sorted tile
    uint32_t* global_base = digit_start + kRadix;                // global output index of the run's first element
    uint32_t* sscan = global_base + kRadix;                      // [16]

    pdl_enter();
    const int64_t n = n_dev ? *n_dev : n_host;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    if (base >= n) return;
    const int tile_n = (n - base) < kSortTile ? (int)(n - base) : kSortTile;
    // issued early: this block's scanned histogram column and the digit totals (thread d <-> digit d)
    const uint32_t my_hist = hist[(int64_t)threadIdx.x * nblocks + blockIdx.x];
    const uint32_t my_total = totals[threadIdx.x];

    for (int i = threadIdx.x; i < kWarps * kRadix; i += kSortThreads) (&warp_hist[0][0])[i] = 0;
    __syncthreads();

    // segmented input: the present pairs of a segment sit at its start and a warp's 512 slots lie inside one
    // segment, so the rounds behind the segment's count hold nothing and are skipped (warp-uniform)
    int rounds_w = kRounds;
    if (SEG) {
        const int64_t w0 = base + warp * kPerWarp;  // first slot of this warp
        const int left = w0 < n ? seg.counts[w0 >> seg.shift] - (int)(w0 & ((1 << seg.shift) - 1)) : 0;
        rounds_w = left <= 0 ? 0 : (left + 31) / 32 < kRounds ? (left + 31) / 32 : kRounds;
    }
    // phase 0: all of the tile's loads are issued before anything depends on them (one memory latency)
    KeyT key[kRounds];
    int32_t val[kRounds];
    uint16_t rank[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int local = warp * kPerWarp + r * 32 + lane;
        key[r] = (local < tile_n && (!SEG || seg.ok(base + local))) ? keys_in[base + local] : (KeyT)0;
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int local = warp * kPerWarp + r * 32 + lane;
        val[r] = (local < tile_n && (!SEG || seg.ok(base + local))) ? vals_in[base + local] : 0;
    }
    // phase A: stable rank inside the warp's 512-pair sub-chunk
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int local = warp * kPerWarp + r * 32 + lane;
        if (SEG && r >= rounds_w) {
            rank[r] = 0;
            continue;
        }
        const bool valid = local < tile_n && (!SEG || seg.ok(base + local));
        const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
        const uint32_t peers = peers_by_ballot(d, valid);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = warp_hist[warp][d];
            warp_hist[warp][d] = old + __popc(peers);
        }
        // every lane fetches `old` from its own leader (lanes of different digits have different leaders)
        old = __shfl_sync(0xffffffffu, old, valid ? leader : lane);
        rank[r] = (uint16_t)(old + __popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
    }
    __syncthreads();
    // phase B: thread d: exclusive prefix over warps + block count; exclusive scans over the digits
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            uint32_t c = warp_hist[w][d];
            warp_hist[w][d] = run;
            run += c;
        }
        const uint32_t tot = my_total;
        uint32_t inc = run, ginc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            uint32_t g = __shfl_up_sync(0xffffffffu, ginc, o);
            if (lane >= o) {
                inc += t;
                ginc += g;
            }
        }
        if (lane == 31) {
            sscan[warp] = inc;
            sscan[8 + warp] = ginc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0, gacc = 0;
            for (int w = 0; w < kWarps; ++w) {
                uint32_t t = sscan[w], g = sscan[8 + w];
                sscan[w] = acc;
                sscan[8 + w] = gacc;
                acc += t;
                gacc += g;
            }
        }
        __syncthreads();
        digit_start[d] = sscan[warp] + inc - run;
        global_base[d] = (sscan[8 + warp] + ginc - tot) + my_hist;
        if (d == kRadix - 1) s_tile_pairs = (int)(sscan[warp] + inc);  // pairs actually present in this tile
    }
    __syncthreads();
    // phase C: place into block-sorted order in smem
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int local = warp * kPerWarp + r * 32 + lane;
        if (local < tile_n && (!SEG || seg.ok(base + local))) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
            const uint32_t pos = digit_start[d] + warp_hist[warp][d] + rank[r];
            skeys[pos] = key[r];
            svals[pos] = val[r];
        }
    }
    __syncthreads();
    // phase D: coalesced run writes (+ the histogram of the NEXT pass: the destination index decides which block
    // of the next pass the pair lands in, so the next pass needs no upsweep)
    const int tile_pairs = s_tile_pairs;
    for (int i0 = 0; i0 < tile_pairs; i0 += kSortThreads) {
        const int i = i0 + threadIdx.x;
        const bool valid = i < tile_pairs;
        uint32_t bin = 0xffffffffu;
        if (valid) {
            const KeyT k = skeys[i];
            const uint32_t d = (uint32_t)(k >> shift) & mask;
            const int64_t dst = (int64_t)global_base[d] + (i - digit_start[d]);
            keys_out[dst] = k;
            vals_out[dst] = svals[i];
            if (next_hist) bin = (((uint32_t)(k >> next_shift) & next_mask) * (uint32_t)nblocks) + (uint32_t)(dst / kSortTile);
        }
        if (next_hist) {
            const uint32_t same = __match_any_sync(0xffffffffu, bin);
            if (valid && lane == __ffs(same) - 1) atomicAdd(next_hist + bin, (uint32_t)__popc(same));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Single-kernel passes ("onesweep"): one histogram kernel counts the digits of ALL passes up front (digit counts do
// not depend on the order of the pairs); each pass is then ONE kernel in which a block ranks its tile locally,
// publishes its per-digit counts and obtains its global offsets by a decoupled look-back over the blocks before it.
// Blocks take their tile from an atomic ticket, so a block only ever waits on blocks that are already running.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPasses = 8;
constexpr int kOnesweepMaxBlocks = 3 * 148;  // 3 resident blocks per SM x 148 SMs

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads) radix_histogram_kernel(int64_t n_host, const int64_t* n_dev, const KeyT* __restrict__ keys, int passes,
                                                                      int end_bit, uint32_t* __restrict__ hist_all /* [kMaxPasses][kRadix] */) {
    __shared__ uint32_t sh[kMaxPasses][kRadix];
    pdl_enter();
    const int64_t n = n_dev ? *n_dev : n_host;
    for (int i = threadIdx.x; i < passes * kRadix; i += kSortThreads) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    if (base < n) {
        for (int k = 0; k < kSortItems; ++k) {
            const int64_t i = base + k * kSortThreads + threadIdx.x;
            if (i < n) {
                const KeyT key = keys[i];
                for (int p = 0; p < passes; ++p) {
                    const int bits = (end_bit - p * 8) < 8 ? (end_bit - p * 8) : 8;
                    atomicAdd(&sh[p][(uint32_t)(key >> (p * 8)) & ((1u << bits) - 1u)], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += kSortThreads) {
        const uint32_t c = (&sh[0][0])[i];
        if (c) atomicAdd(hist_all + i, c);
    }
}

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ITEMS pairs per thread: 16 (the tile of the three-kernel passes) or 8 -- a sort of a few hundred thousand pairs has ~1 block
// of 4096 per SM, so a pass costs one block's serial work (16 ranking rounds per warp) + the look-back; half-size tiles double
// the parallelism.  MEASURED at S1 (668 k keys, 4 passes): isect_prepare 0.122 -> 0.133 ms with ITEMS = 8 -- twice the blocks make
// the look-back walk longer than the ranking gets shorter -- so 16 stays the default and 8 is only reachable through the hook.
template <typename KeyT, int ITEMS>
__global__ void __launch_bounds__(kSortThreads, ITEMS == 16 ? 3 : 4) radix_onesweep_kernel(int64_t n_host, const int64_t* n_dev, const KeyT* __restrict__ keys_in,
                                                                        const int32_t* __restrict__ vals_in, KeyT* __restrict__ keys_out,
                                                                        int32_t* __restrict__ vals_out, int shift, uint32_t mask, int pass,
                                                                        const uint32_t* __restrict__ hist_all, uint64_t* __restrict__ status,
                                                                        uint32_t* __restrict__ tickets) {
    constexpr int kTile = kSortThreads * ITEMS;
    constexpr int kWarps = kSortThreads / 32;
    constexpr int kPerWarp = kTile / kWarps;
    constexpr int kRounds = kPerWarp / 32;
    extern __shared__ __align__(16) unsigned char sort_smem[];
    KeyT* skeys = reinterpret_cast<KeyT*>(sort_smem);
    int32_t* svals = reinterpret_cast<int32_t*>(skeys + kTile);
    uint32_t(*warp_hist)[kRadix] = reinterpret_cast<uint32_t(*)[kRadix]>(svals + kTile);
    uint32_t* digit_start = &warp_hist[0][0] + kWarps * kRadix;
    uint32_t* global_base = digit_start + kRadix;
    uint32_t* sscan = global_base + kRadix;  // [16]
    __shared__ uint32_t s_ticket;

    pdl_enter();
    const int64_t n = n_dev ? *n_dev : n_host;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_ticket = atomicAdd(tickets + pass, 1u);
    // digit totals of this pass (thread d <-> digit d), issued early
    const uint32_t my_total = hist_all[pass * kRadix + threadIdx.x];
    for (int i = threadIdx.x; i < kWarps * kRadix; i += kSortThreads) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t vb = s_ticket;  // virtual block id: tiles are handed out in ticket order
    const int64_t base = (int64_t)vb * kTile;
    if (base >= n) return;
    const int tile_n = (n - base) < kTile ? (int)(n - base) : kTile;

    KeyT key[kRounds];
    int32_t val[kRounds];
    uint16_t rank[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int local = warp * kPerWarp + r * 32 + lane;
        key[r] = (local < tile_n) ? keys_in[base + local] : (KeyT)0;
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int local = warp * kPerWarp + r * 32 + lane;
        val[r] = (local < tile_n) ? vals_in[base + local] : 0;
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int local = warp * kPerWarp + r * 32 + lane;
        const bool valid = local < tile_n;
        const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
        const uint32_t peers = peers_by_ballot(d, valid);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = warp_hist[warp][d];
            warp_hist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, valid ? leader : lane);
        rank[r] = (uint16_t)(old + __popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
    }
    __syncthreads();
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            uint32_t c = warp_hist[w][d];
            warp_hist[w][d] = run;
            run += c;
        }
        // publish this block's count for digit d, then look back for the exclusive prefix over earlier blocks
        const uint32_t FLAG_AGG = 2u * (uint32_t)pass + 1u, FLAG_PFX = 2u * (uint32_t)pass + 2u;
        uint64_t* my_status = status + (size_t)vb * kRadix + d;
        st_relaxed_u64(my_status, ((uint64_t)(vb == 0 ? FLAG_PFX : FLAG_AGG) << 32) | run);
        uint32_t excl = 0;
        if (vb > 0) {
            // All blocks of a wave publish their aggregate at about the same time, so the walk back to the last
            // published prefix can be as long as the number of resident blocks: kLook independent loads are kept
            // in flight per step instead of one dependent L2 round trip per predecessor.
            constexpr int kLook = 8;  // (16 in flight measured slower at S1: isect_prepare 0.122 -> 0.134 ms, 80 bytes spilled)
            int64_t b = (int64_t)vb - 1;
            bool done = false;
            while (!done) {
                uint64_t v[kLook];
#pragma unroll
                for (int j = 0; j < kLook; ++j) v[j] = (b - j >= 0) ? ld_relaxed_u64(status + (size_t)(b - j) * kRadix + d) : 0ull;
                int consumed = 0;
                bool stop = false;
#pragma unroll
                for (int j = 0; j < kLook; ++j) {
                    if (stop || done || b - j < 0) continue;
                    const uint32_t f = (uint32_t)(v[j] >> 32);
                    if (f == FLAG_PFX) {
                        excl += (uint32_t)v[j];
                        done = true;
                    } else if (f == FLAG_AGG) {
                        excl += (uint32_t)v[j];
                        consumed = j + 1;
                    } else {
                        stop = true;  // predecessor b-j has not published yet: retry from there
                    }
                }
                b -= consumed;
                // block 0 always publishes a prefix, so the walk ends before b < 0
            }
            st_relaxed_u64(my_status, ((uint64_t)FLAG_PFX << 32) | (uint64_t)(excl + run));
        }
        // exclusive scans over the digits: block-local run starts and global digit bases
        const uint32_t tot = my_total;
        uint32_t inc = run, ginc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            uint32_t g = __shfl_up_sync(0xffffffffu, ginc, o);
            if (lane >= o) {
                inc += t;
                ginc += g;
            }
        }
        if (lane == 31) {
            sscan[warp] = inc;
            sscan[8 + warp] = ginc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0, gacc = 0;
            for (int w = 0; w < kWarps; ++w) {
                uint32_t t = sscan[w], g = sscan[8 + w];
                sscan[w] = acc;
                sscan[8 + w] = gacc;
                acc += t;
                gacc += g;
            }
        }
        __syncthreads();
        digit_start[d] = sscan[warp] + inc - run;
        global_base[d] = (sscan[8 + warp] + ginc - tot) + excl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int local = warp * kPerWarp + r * 32 + lane;
        if (local < tile_n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
            const uint32_t pos = digit_start[d] + warp_hist[warp][d] + rank[r];
            skeys[pos] = key[r];
            svals[pos] = val[r];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tile_n; i += kSortThreads) {
        const KeyT k = skeys[i];
        const uint32_t d = (uint32_t)(k >> shift) & mask;
        const int64_t dst = (int64_t)global_base[d] + (i - digit_start[d]);
        keys_out[dst] = k;
        vals_out[dst] = svals[i];
    }
}

inline size_t radix_hist_bytes(int64_t capacity) {
    int64_t nb = (capacity + kSortTile - 1) / kSortTile;
    if (nb < 1) nb = 1;
    // [two per-block histograms + digit totals] for the three-kernel passes, [status words + all-pass histogram +
    // tickets] for the single-kernel passes
    // (status words: one per (block, digit); the half-size-tile passes have 2 nb blocks)
    return 2 * (((size_t)kRadix * nb * 4 + 255) / 256 * 256) + ((kRadix * 4 + 255) / 256 * 256) +
           (((size_t)kRadix * 2 * nb * 8 + 255) / 256 * 256) + ((size_t)(kMaxPasses * kRadix + kMaxPasses) * 4 + 255) / 256 * 256;
}

// Sort on key bits [0, end_bit).  Result lands in (keys_out, vals_out); (tmp_keys, tmp_vals) is the
// ping-pong buffer; `hist` = radix_hist_bytes(capacity) scratch.  keys_in/vals_in are not modified.
template <typename KeyT>
inline int radix_sort_pairs(int64_t capacity, const int64_t* n_dev, const KeyT* keys_in, const int32_t* vals_in, KeyT* keys_out,
                            int32_t* vals_out, KeyT* tmp_keys, int32_t* tmp_vals, void* hist_ws, int end_bit, cudaStream_t stream,
                            SegCounts seg = SegCounts()) {
    if (capacity <= 0) return QED_OK;
    const int nb = (int)((capacity + kSortTile - 1) / kSortTile);
    const size_t hist_bytes = ((size_t)kRadix * nb * 4 + 255) / 256 * 256;
    uint32_t* hist_a = reinterpret_cast<uint32_t*>(hist_ws);
    uint32_t* hist_b = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(hist_ws) + hist_bytes);
    uint32_t* totals = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(hist_ws) + 2 * hist_bytes);
    const int passes = (end_bit + 7) / 8;
    if (passes == 0) {
        if (n_dev) return QED_ERR_UNSUPPORTED;
        QED_CUDA_TRY(cudaMemcpyAsync(keys_out, keys_in, (size_t)capacity * sizeof(KeyT), cudaMemcpyDeviceToDevice, stream));
        QED_CUDA_TRY(cudaMemcpyAsync(vals_out, vals_in, (size_t)capacity * 4, cudaMemcpyDeviceToDevice, stream));
        return QED_OK;
    }
    if (!seg.counts && g_radix_onesweep && passes <= kMaxPasses && (nb <= kOnesweepMaxBlocks || g_radix_onesweep > 1)) {
        char* ows = reinterpret_cast<char*>(hist_ws) + 2 * hist_bytes + ((kRadix * 4 + 255) / 256 * 256);
        // half-size tiles while the whole sort is <= 2 blocks of 4096 per SM (g_radix_small_tiles: test hook)
        const bool small = g_radix_small_tiles && nb <= 2 * 148;
        const int nbk = small ? (int)((capacity + kSortTile / 2 - 1) / (kSortTile / 2)) : nb;
        uint64_t* status = reinterpret_cast<uint64_t*>(ows);
        const size_t status_bytes = ((size_t)kRadix * 2 * nb * 8 + 255) / 256 * 256;
        uint32_t* hist_all = reinterpret_cast<uint32_t*>(ows + status_bytes);
        uint32_t* tickets = hist_all + kMaxPasses * kRadix;
        QED_CUDA_TRY(cudaMemsetAsync(ows, 0, status_bytes + (size_t)(kMaxPasses * kRadix + kMaxPasses) * 4, stream));
        QED_CUDA_TRY(launch_pdl(radix_histogram_kernel<KeyT>, dim3(nb), dim3(kSortThreads), 0, stream, capacity, n_dev, keys_in, passes, end_bit, hist_all));
        auto one = small ? radix_onesweep_kernel<KeyT, 8> : radix_onesweep_kernel<KeyT, 16>;
        const size_t one_smem = small ? RadixSmem<KeyT>::kBytes - (size_t)(kSortTile / 2) * (sizeof(KeyT) + 4) : RadixSmem<KeyT>::kBytes;
        QED_CUDA_TRY(cudaFuncSetAttribute(one, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)one_smem));
        const KeyT* sk = keys_in;
        const int32_t* sv = vals_in;
        for (int pass = 0; pass < passes; ++pass) {
            const bool to_out = ((passes - 1 - pass) % 2) == 0;
            KeyT* dk = to_out ? keys_out : tmp_keys;
            int32_t* dv = to_out ? vals_out : tmp_vals;
            const int bits = (end_bit - pass * 8) < 8 ? (end_bit - pass * 8) : 8;
            QED_CUDA_TRY(launch_pdl(one, dim3(nbk), dim3(kSortThreads), one_smem, stream, capacity, n_dev, sk, sv, dk, dv, pass * 8,
                                    (1u << bits) - 1u, pass, hist_all, status, tickets));
            sk = dk;
            sv = dv;
        }
        return QED_OK;
    }
    auto down = radix_downsweep_kernel<KeyT, false>;
    auto down_seg = radix_downsweep_kernel<KeyT, true>;
    QED_CUDA_TRY(cudaFuncSetAttribute(down, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RadixSmem<KeyT>::kBytes));
    if (seg.counts) QED_CUDA_TRY(cudaFuncSetAttribute(down_seg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RadixSmem<KeyT>::kBytes));
    auto pass_mask = [&](int pass) {
        const int bits = (end_bit - pass * 8) < 8 ? (end_bit - pass * 8) : 8;
        return (uint32_t)((1u << bits) - 1u);
    };
    const KeyT* src_k = keys_in;
    const int32_t* src_v = vals_in;
    uint32_t* hist = hist_a;
    uint32_t* hist_next = hist_b;
    // (Accumulating the next pass's histogram inside the downsweep was measured SLOWER than a separate upsweep:
    //  115 us vs 73 + 20 us for 6.5 M pairs — the per-pair global atomics do not aggregate because consecutive
    //  block-sorted pairs differ in their next digit.  The kernel keeps the hook; the host does not use it.)
    for (int pass = 0; pass < passes; ++pass) {
        const bool to_out = ((passes - 1 - pass) % 2) == 0;  // destinations alternate, ending on *_out
        KeyT* dst_k = to_out ? keys_out : tmp_keys;
        int32_t* dst_v = to_out ? vals_out : tmp_vals;
        // segmented input: the first pass walks all `capacity` slots and keeps the present pairs; from then on the
        // data is dense and n_dev (the number of present pairs) applies
        const SegCounts sg = pass == 0 ? seg : SegCounts();
        const int64_t* nd = (pass == 0 && seg.counts) ? nullptr : n_dev;
        if (sg.counts)
            QED_CUDA_TRY(launch_pdl(radix_upsweep_kernel<KeyT, true>, dim3(nb), dim3(kSortThreads), 0, stream, capacity, nd, src_k, pass * 8,
                                    pass_mask(pass), nb, hist, sg));
        else
            QED_CUDA_TRY(launch_pdl(radix_upsweep_kernel<KeyT, false>, dim3(nb), dim3(kSortThreads), 0, stream, capacity, nd, src_k, pass * 8,
                                    pass_mask(pass), nb, hist, sg));
        QED_CUDA_TRY(launch_pdl(radix_scan_kernel, dim3(kRadix), dim3(256), 0, stream, nb, hist, totals));
        QED_CUDA_TRY(launch_pdl(sg.counts ? down_seg : down, dim3(nb), dim3(kSortThreads), RadixSmem<KeyT>::kBytes, stream, capacity, nd, src_k, src_v,
                                dst_k, dst_v, pass * 8, pass_mask(pass), nb, hist, totals, (uint32_t*)nullptr, 0, 0u, sg));
        src_k = dst_k;
        src_v = dst_v;
    }
    (void)hist_next;
    return QED_OK;
}

}  // namespace qed
