// View-sharded training: sum of the flat gradient arena over the ranks, written for NVLink 5 / NVSwitch.
//
// The reference has no collective at all; the view-sharded step of north_star needs ONE: every rank
// holds the gradient of its own views in a flat arena (59 floats per Gaussian, trainer.GaussianArena) and all ranks
// need the sum.  The arena lives in SYMMETRIC memory (same allocation on every rank, mapped into every peer and into
// one NVSwitch multicast address range -- the host side obtains the mappings from torch.distributed._symmetric_memory,
// which is plumbing only: no torch collective runs on the data path).
//
//   qed_comm_allreduce_f32: two-shot all-reduce in ONE kernel.  Rank r owns the r-th slice of the arena:
//       NVLS path   : multimem.ld_reduce.add.v4.f32 pulls the slice from all ranks and sums it INSIDE the switch (each
//                     GPU sends every byte of its arena once, receives only its slice), multimem.st.v4.f32 broadcasts the
//                     sums through the switch.  Per GPU and direction: S (1 + 1/G) bytes instead of 2 S (G-1)/G of a ring.
//       peer path   : the same with plain peer loads / stores (no multicast object on the box): sums in rank order.
//     Either way one rank computes each sum and everyone receives those bits: replicas stay bit-identical.
//   cross-GPU ordering: epoch flags in symmetric memory (release / acquire at system scope), no host involvement, no reset:
//     every call uses two fresh epochs (entry: all ranks' gradients are written; exit: all slices are broadcast).
//   Any [begin, end) element range can be reduced: the step only all-reduces the 11 non-SH floats per Gaussian; the SH
//   coefficient gradient (81 % of the arena) is rebuilt on every rank from exchanged per-view colour gradients
//   (project_bwd.cu: qed_project_bwd_exchange / qed_sh_grad_from_view_colors).
#include "common.cuh"

namespace qed {

constexpr int kCommMaxWorld = 16;
constexpr int kCommMaxBlocks = 256;
constexpr int kCommThreads = 512;
constexpr int kCommUnroll = 4;

struct CommPeers {
    float* base[kCommMaxWorld];      // unicast address of every rank's arena in THIS process (peer path)
    uint32_t* flags[kCommMaxWorld];  // every rank's flag block [kCommMaxBlocks][kCommMaxWorld]
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Block b of every rank meets block b of every other rank: thread t < world raises flag (b, my rank) on rank t and
// waits for rank t's flag (b, t) on this rank.  Epochs only grow, so a peer that is already one barrier ahead
// still satisfies the wait and nothing is ever reset.
__device__ __forceinline__ void barrier_all_ranks(const CommPeers& peers, int rank, int world, uint32_t epoch) {
    __syncthreads();
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + (size_t)blockIdx.x * kCommMaxWorld + rank, epoch);
        const uint32_t* mine = peers.flags[rank] + (size_t)blockIdx.x * kCommMaxWorld + threadIdx.x;
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// elements (float4 units) [lo4, hi4) of the arena; rank r reduces the r-th part of that range
template <bool NVLS>
__global__ void __launch_bounds__(kCommThreads) comm_allreduce_kernel(float* __restrict__ mc, CommPeers peers, int64_t lo4, int64_t hi4, int rank,
                                                                     int world, uint32_t epoch) {
    barrier_all_ranks(peers, rank, world, epoch);  // every rank's gradients are written
    const int64_t n4 = hi4 - lo4;
    const int64_t per = (n4 + world - 1) / world;
    const int64_t s = lo4 + min((int64_t)rank * per, n4), e = lo4 + min((int64_t)(rank + 1) * per, n4);
    const int64_t stride = (int64_t)gridDim.x * kCommThreads;
    for (int64_t i0 = s + (int64_t)blockIdx.x * kCommThreads + threadIdx.x; i0 < e; i0 += stride * kCommUnroll) {
        float4 v[kCommUnroll];
#pragma unroll
        for (int u = 0; u < kCommUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < e) {
                if (NVLS) {
                    v[u] = multimem_ld_reduce_add(mc + i * 4);
                } else {
                    float4 a = *reinterpret_cast<const float4*>(peers.base[0] + i * 4);
                    for (int r = 1; r < world; ++r) {
                        const float4 b = *reinterpret_cast<const float4*>(peers.base[r] + i * 4);
                        a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
                    }
                    v[u] = a;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kCommUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < e) {
                if (NVLS) {
                    multimem_st(mc + i * 4, v[u]);
                } else {
                    for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(peers.base[r] + i * 4) = v[u];
                }
            }
        }
    }
    barrier_all_ranks(peers, rank, world, epoch + 1);  // every slice has been broadcast
}

// every rank's earlier work on its stream (e.g. the exchange stores of the projection backward) is complete and visible
__global__ void comm_barrier_kernel(CommPeers peers, int rank, int world, uint32_t epoch) { barrier_all_ranks(peers, rank, world, epoch); }

}  // namespace qed

using namespace qed;

extern "C" int qed_comm_flag_words(void) { return kCommMaxBlocks * kCommMaxWorld; }

extern "C" int qed_comm_barrier(uint32_t* const* peer_flags, int rank, int world, uint32_t epoch, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world || !peer_flags) return QED_ERR_BAD_ARG;
    if (world == 1) return QED_OK;
    CommPeers peers{};
    for (int r = 0; r < world; ++r) {
        peers.flags[r] = peer_flags[r];
        if (!peers.flags[r]) return QED_ERR_BAD_ARG;
    }
    comm_barrier_kernel<<<1, 32, 0, stream>>>(peers, rank, world, epoch);
    QED_LAUNCH_CHECK();
    return QED_OK;
}

extern "C" int qed_comm_allreduce_f32(float* multicast_base, float* const* peer_bases, uint32_t* const* peer_flags, int rank, int world,
                                      int64_t begin, int64_t end, uint32_t epoch, int blocks, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world || begin < 0 || end < begin) return QED_ERR_BAD_ARG;
    if ((begin & 3) || (end & 3)) return QED_ERR_BAD_ARG;  // float4 units: the arena groups are 16-byte aligned
    if (!peer_flags || (!multicast_base && !peer_bases)) return QED_ERR_BAD_ARG;
    if (world == 1 || begin == end) return QED_OK;
    if (blocks <= 0) blocks = 32;  // NVLS saturates with few SMs (8 B200: 16 blocks 0.54 ms, 256 blocks 0.59 ms for 236 MB)
    if (blocks > kCommMaxBlocks) blocks = kCommMaxBlocks;
    CommPeers peers{};
    for (int r = 0; r < world; ++r) {
        peers.flags[r] = peer_flags[r];
        peers.base[r] = peer_bases ? peer_bases[r] : nullptr;
        if (!peers.flags[r] || (!multicast_base && !peers.base[r])) return QED_ERR_BAD_ARG;
    }
    if (multicast_base) {
        if (reinterpret_cast<uintptr_t>(multicast_base) & 15) return QED_ERR_BAD_ARG;
        comm_allreduce_kernel<true><<<blocks, kCommThreads, 0, stream>>>(multicast_base, peers, begin / 4, end / 4, rank, world, epoch);
    } else {
        comm_allreduce_kernel<false><<<blocks, kCommThreads, 0, stream>>>(nullptr, peers, begin / 4, end / 4, rank, world, epoch);
    }
    QED_LAUNCH_CHECK();
    return QED_OK;
}
