// (c)/(d) Per-tile front-to-back alpha compositing of D channels (RGB + depth), forward and backward.
//
// Replaces gsplat's rasterize_to_pixels_fwd/_bwd (+ the ED normalisation tail of rasterization()) behind
// qed_splatter/model.py:267-288; absgrad (model.py:284) is produced by the backward.
// Semantics: SURVEY.md Appendix A.5 / A.6 == oracle/torch_impl.py::rasterize_to_pixels(_bwd).
//
// Design (B200-first, not gsplat's).  Both kernels are bound by instruction issue and the fp32 pipes (ncu:
// 72-82 % issue-active, <10 % L2), so the design minimises instructions per evaluated (pixel, Gaussian) pair:
//   * one CTA of two warps per 16x16 tile (tile size is fixed by the bit-exact intersection contract); a warp
//     owns a 16x8 footprint = two 8x8 blocks, a lane owns 4 pixels (columns j0 + 8a, rows i0 + 4b).
//   * packed fp32: the two pixels of a lane in one 8x8 block share dx and sit in the two halves of an f32x2;
//     sm_100's FFMA2 / FMUL2 / FADD2 retire both per issue slot (scalar-broadcast, negate, |.| operands free).
//   * log2 domain: log2(alpha) = lo + qa dx^2 + qb dx dy + qc dy^2 with lo = log2(opacity),
//     (qa,qb,qc) = -log2(e) * (a/2, b, c/2); the exponential is a bare ex2.approx (one MUFU).
//   * warp streams: each warp walks the tile's sorted range on its own, 32 entries per batch; the gather
//     (id -> 32-B geom + colour) of the next batch is in flight as cp.async while the current one is
//     composited; no block barrier anywhere.
//   * EXACT culling: a Gaussian changes a pixel only if alpha >= 1/255, i.e. the pixel centre is inside the
//     ellipse Q(d) <= lo + log2(255).  Every lane tests its entry against the warp's two 8x8 blocks
//     (conservative: minimum of Q over the rectangle plus a rounding margin) and survivors are compacted, in
//     order, into a per-warp shared-memory queue with their block mask.  The inner loop is a counted loop over
//     that queue (3 broadcast LDS.128 per Gaussian).  Results are identical to the un-culled kernel (test hook).
//   * a pixel that fails the alpha test simply gets alpha = 0 (T * 1, + 0): no selects behind the packed ops.
//   * backward: back-to-front replay from the stored last index with a scalar running
//     bsum = sum_k buffer[k] v_out[k]; per Gaussian the 12 gradient values of a lane's 4 pixels are pre-added,
//     transposed through shared memory and summed by 24 lanes, one red.global.add.v4.f32 each, into the
//     packed [C*N,12] record.
// The scalar kernels (raster_fwd_kernel / raster_bwd_kernel: CTA-staged, 1/2/4 pixels per lane, shuffle
// butterfly) are kept as the cross-check the packed ones are tested against (qed_debug_set_raster_packed).
#include <cstddef>

#include "common.cuh"

namespace qed {

constexpr int kTile = 16;

struct RasterParams {
    int C, N, D, width, height, tile_w, tile_h, normalize_last, offsets_has_end;
    int64_t n_isects;
    const float4* geom;
    const float* colors;
    const float* backgrounds;
    const int32_t* offsets;
    const int32_t* flatten_ids;
    // fwd outputs / bwd inputs
    float* render;
    float* alphas;
    int32_t* last_ids;
    const float* v_render;
    const float* v_alphas;
    float* packed_grads;
    unsigned long long* counters;  // STATS instantiations only (bench/test instrumentation)
};

// counters layout: [0] entries loaded, [1] entries staged after the tile-level cull, [2] (warp, Gaussian)
// candidates after the warp-level cull, [3] (pixel, Gaussian) evaluations by live pixels, [4] pairs that
// passed the alpha test (composited / differentiated), [5] (warp, Gaussian) gradient reductions
template <bool STATS>
struct StatCounters {
    unsigned long long c[6] = {0, 0, 0, 0, 0, 0};
    __device__ __forceinline__ void add(int i, unsigned long long v) {
        if (STATS) c[i] += v;
    }
    __device__ __forceinline__ void flush(unsigned long long* out) {
        if (STATS) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                unsigned long long v = c[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((threadIdx.x & 31) == 0 && v) atomicAdd(out + i, v);
            }
        }
    }
};

// Geometry of the thread block: PX pixels per lane -> warps per tile and warp footprint.
template <int PX>
struct Shape {
    static constexpr int kWarps = 8 / PX;
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kFootW = PX >= 2 ? 16 : 8;
    static constexpr int kFootH = PX == 4 ? 8 : 4;
    static constexpr int kNX = PX >= 2 ? 2 : 1;  // distinct pixel columns per lane (8 apart)
    static constexpr int kNY = PX == 4 ? 2 : 1;  // distinct pixel rows per lane (4 apart)
    __device__ static __forceinline__ int origin_x(int warp) { return PX == 1 ? (warp & 1) * 8 : 0; }
    __device__ static __forceinline__ int origin_y(int warp) { return PX == 1 ? (warp >> 1) * 4 : warp * kFootH; }
};

// Shared staging: `sa/sb/sc` = one batch (one entry per thread, tile-level survivors compacted in order),
// `qa/qb/qc` = per-warp queue of the (<= 32) survivors of the warp-level cull for the current chunk.
//   a = {mx, my, lo, sorted index (bits)}   b = {qa, qb, qc, flat gaussian id (bits)}   c = colour[0..3]
template <int PX>
struct Staging {
    float4 sa[Shape<PX>::kThreads], sb[Shape<PX>::kThreads], sc[Shape<PX>::kThreads];
    float4 qa[Shape<PX>::kWarps][32], qb[Shape<PX>::kWarps][32], qc[Shape<PX>::kWarps][32];
    int qm[Shape<PX>::kWarps][32];  // which of the lane's PX 8x4 sub-rectangles the Gaussian can touch
    int warp_count[Shape<PX>::kWarps];
    int max_last[Shape<PX>::kWarps];
};

template <int D>
__device__ __forceinline__ float4 load_color4(const float* __restrict__ colors, int64_t g) {
    if (D == 4) return reinterpret_cast<const float4*>(colors)[g];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = colors[g * D];
    if (D >= 3) {
        v.y = colors[g * D + 1];
        v.z = colors[g * D + 2];
    }
    return v;
}

// Loads entry `e` of the sorted list for this thread (or nothing if !have), converts it to the log2 domain,
// runs the tile-level cull and writes survivors compacted in thread order.  Returns the staged count.
template <int D, int PX, bool CULL, bool STATS>
__device__ __forceinline__ int stage_batch(const RasterParams& p, Staging<PX>& sm, bool have, int64_t e, float tx0, float ty0, float tx1,
                                           float ty1, StatCounters<STATS>& st) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 A = make_float4(0, 0, 0, 0), B = make_float4(0, 0, 0, 0);
    int64_t g = 0;
    bool keep = false;
    if (have) {
        g = p.flatten_ids[e];
        const float4 ga = p.geom[g * 2];      // mx, my, opacity, depth
        const float4 gb = p.geom[g * 2 + 1];  // conic a, b, c
        const float lo = __log2f(ga.z);
        A = make_float4(ga.x, ga.y, lo, __int_as_float((int)e));
        B = make_float4(-0.5f * kLog2e * gb.x, -kLog2e * gb.y, -0.5f * kLog2e * gb.z, __int_as_float((int)g));
        keep = CULL ? ellipse_hits_rect(A.x, A.y, -B.x, -B.y, -B.z, lo + kLog2_255, tx0, ty0, tx1, ty1) : true;
    }
    st.add(0, have ? 1 : 0);
    st.add(1, keep ? 1 : 0);
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) sm.warp_count[warp] = __popc(m);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < Shape<PX>::kWarps; ++w) {
        const int cnt = sm.warp_count[w];
        if (w < warp) base += cnt;
        total += cnt;
    }
    if (keep) {
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        sm.sa[pos] = A;
        sm.sb[pos] = B;
        sm.sc[pos] = load_color4<D>(p.colors, g);
    }
    __syncthreads();
    return total;
}

// Pixel-centre rectangles of the PX 8x4 sub-rectangles of a warp footprint, clipped to the image.
template <int PX>
struct SubRects {
    float x0[PX], y0[PX], x1[PX], y1[PX];
    int present;  // bit k set if sub-rectangle k has at least one pixel inside the image
};

template <int PX>
__device__ __forceinline__ SubRects<PX> make_subrects(int ox, int oy, int width, int height) {
    SubRects<PX> r;
    r.present = 0;
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const int sx = ox + (k & 1) * 8, sy = oy + (k >> 1) * 4;
        r.x0[k] = (float)sx + 0.5f;
        r.y0[k] = (float)sy + 0.5f;
        r.x1[k] = (float)min(sx + 7, width - 1) + 0.5f;
        r.y1[k] = (float)min(sy + 3, height - 1) + 0.5f;
        if (sx < width && sy < height) r.present |= 1 << k;
    }
    return r;
}

// The two 8x8 pixel blocks of a 16x8 warp footprint (packed kernels: a lane's two pixels of a block, 4 rows
// apart, share one f32x2 register pair), clipped to the image.
__device__ __forceinline__ SubRects<2> make_blocks8(int ox, int oy, int width, int height) {
    SubRects<2> r;
    r.present = 0;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int sx = ox + a * 8;
        r.x0[a] = (float)sx + 0.5f;
        r.y0[a] = (float)oy + 0.5f;
        r.x1[a] = (float)min(sx + 7, width - 1) + 0.5f;
        r.y1[a] = (float)min(oy + 7, height - 1) + 0.5f;
        if (sx < width && oy < height) r.present |= 1 << a;
    }
    return r;
}

// Warp-level cull of staged entries [c0, c0+32): every lane tests one entry against the PX sub-rectangles
// still of interest (`want` bits; backward: additionally sorted index <= max_sid[k]); entries that can
// touch at least one are copied, compacted and in order, into the warp's queue together with their
// sub-rectangle mask.  Returns the number of queued entries (warp-uniform).
template <int PX, bool CULL, bool BWD, int NR = PX>
__device__ __forceinline__ int fill_queue(Staging<PX>& sm, int c0, int count, const SubRects<NR>& sr, int want, const int* max_sid) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = c0 + lane;
    int mask = 0;
    float4 A = make_float4(0, 0, 0, 0), B = make_float4(0, 0, 0, 0);
    if (q < count && want) {
        A = sm.sa[q];
        B = sm.sb[q];
        const int sid = __float_as_int(A.w);
        const float tau2 = A.z + kLog2_255;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            bool hit = (want >> k) & 1;
            if (BWD) hit = hit && (sid <= max_sid[k]);
            if (CULL && hit) hit = ellipse_hits_rect(A.x, A.y, -B.x, -B.y, -B.z, tau2, sr.x0[k], sr.y0[k], sr.x1[k], sr.y1[k]);
            if (hit) mask |= 1 << k;
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, mask != 0);
    __syncwarp();  // every lane is done reading the previous queue contents
    if (mask) {
        const int pos = __popc(m & ((1u << lane) - 1u));
        sm.qa[warp][pos] = A;
        sm.qb[warp][pos] = B;
        sm.qc[warp][pos] = sm.sc[q];
        if (NR > 1) sm.qm[warp][pos] = mask;
    }
    __syncwarp();
    return __popc(m);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int D, int PX, bool CULL, bool STATS>
__global__ void __launch_bounds__(Shape<PX>::kThreads) raster_fwd_kernel(const RasterParams p) {
    using S = Shape<PX>;
    __shared__ Staging<PX> sm;
    StatCounters<STATS> st;
    const int cam = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ox = tx * kTile + S::origin_x(warp), oy = ty * kTile + S::origin_y(warp);
    const int j0 = ox + (lane & 7), i0 = oy + (lane >> 3);  // first pixel of this lane; others at +8 / +4
    const float px0 = (float)j0 + 0.5f, py0 = (float)i0 + 0.5f;

    const int64_t tile_id = ((int64_t)cam * p.tile_h + ty) * p.tile_w + tx;
    const int64_t range_start = p.offsets[tile_id];
    const int64_t range_end = (!p.offsets_has_end && tile_id == (int64_t)p.C * p.tile_h * p.tile_w - 1) ? p.n_isects : (int64_t)p.offsets[tile_id + 1];

    // pixel-centre rectangles for the culling tests (clipped to the image)
    const float tx0 = (float)(tx * kTile) + 0.5f, ty0 = (float)(ty * kTile) + 0.5f;
    const float tx1 = (float)min(tx * kTile + kTile - 1, p.width - 1) + 0.5f, ty1 = (float)min(ty * kTile + kTile - 1, p.height - 1) + 0.5f;
    const SubRects<PX> sr = make_subrects<PX>(ox, oy, p.width, p.height);

    float T[PX], acc[PX][D];
    int32_t last[PX];
    bool live[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        T[k] = 1.0f;
        last[k] = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) acc[k][d] = 0.0f;
        live[k] = (i0 + (k >> 1) * 4 < p.height) && (j0 + (k & 1) * 8 < p.width);
    }
    int alive = sr.present;  // warp-uniform: sub-rectangles that still have a live pixel

    for (int64_t b0 = range_start; b0 < range_end; b0 += S::kThreads) {
        if (__syncthreads_count(alive != 0) == 0) break;
        const int64_t e = b0 + threadIdx.x;
        const int count = stage_batch<D, PX, CULL, STATS>(p, sm, e < range_end, e, tx0, ty0, tx1, ty1, st);
        for (int c0 = 0; c0 < count && alive; c0 += 32) {
            const int nq = fill_queue<PX, CULL, false>(sm, c0, count, sr, alive, nullptr);
            for (int q = 0; q < nq; ++q) {
                const float4 A = sm.qa[warp][q], B = sm.qb[warp][q], Cc = sm.qc[warp][q];
                const int mask = PX > 1 ? sm.qm[warp][q] : 1;
                if (lane == 0) st.add(2, __popc(mask));
                float dx[S::kNX], ax[S::kNX], dy[S::kNY], cy[S::kNY];
#pragma unroll
                for (int a = 0; a < S::kNX; ++a) {
                    dx[a] = A.x - (px0 + 8.0f * a);
                    ax[a] = B.x * dx[a];
                }
#pragma unroll
                for (int b = 0; b < S::kNY; ++b) {
                    dy[b] = A.y - (py0 + 4.0f * b);
                    cy[b] = fmaf(B.z * dy[b], dy[b], A.z);
                }
#pragma unroll
                for (int k = 0; k < PX; ++k) {
                    if (PX > 1 && !((mask >> k) & 1)) continue;  // warp-uniform
                    const int a = k & 1, b = k >> 1;
                    const float pw = fmaf(fmaf(B.y, dy[b], ax[a]), dx[a], cy[b]);  // log2(opacity * exp(-sigma))
                    const float alpha = fminf(kMaxAlpha, ex2_approx(pw));
                    const bool ok = live[k] && (pw <= A.z) && (alpha >= kAlphaThreshold);  // pw <= lo  <=>  sigma >= 0
                    const float next_T = T[k] * (1.0f - alpha);
                    const bool stop = ok && (next_T <= kTransmittanceThreshold);
                    const bool upd = ok && !stop;
                    st.add(3, live[k] ? 1 : 0);
                    st.add(4, upd ? 1 : 0);
                    const float w = upd ? alpha * T[k] : 0.0f;
                    acc[k][0] = fmaf(Cc.x, w, acc[k][0]);
                    if (D >= 3) {
                        acc[k][1] = fmaf(Cc.y, w, acc[k][1]);
                        acc[k][2] = fmaf(Cc.z, w, acc[k][2]);
                    }
                    if (D == 4) acc[k][3] = fmaf(Cc.w, w, acc[k][3]);
                    T[k] = upd ? next_T : T[k];
                    last[k] = upd ? __float_as_int(A.w) : last[k];
                    live[k] = live[k] && !stop;
                }
            }
            alive = 0;
#pragma unroll
            for (int k = 0; k < PX; ++k) alive |= __any_sync(0xffffffffu, live[k]) ? (1 << k) : 0;
        }
    }

#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const int i = i0 + (k >> 1) * 4, j = j0 + (k & 1) * 8;
        if (i < p.height && j < p.width) {
            const int64_t pix = ((int64_t)cam * p.height + i) * p.width + j;
            const float alpha_out = 1.0f - T[k];
            float out[D];
#pragma unroll
            for (int d = 0; d < D; ++d) out[d] = acc[k][d] + (p.backgrounds ? T[k] * p.backgrounds[cam * D + d] : 0.0f);
            if (p.normalize_last) out[D - 1] = out[D - 1] / fmaxf(alpha_out, 1e-10f);
            if (D == 4) {
                reinterpret_cast<float4*>(p.render)[pix] = make_float4(out[0], out[1], out[2], out[3]);
            } else {
#pragma unroll
                for (int d = 0; d < D; ++d) p.render[pix * D + d] = out[d];
            }
            p.alphas[pix] = alpha_out;
            p.last_ids[pix] = last[k];
        }
    }
    st.flush(p.counters);
}

// Warp streams.  Each warp walks the tile's sorted range on its own, 32 entries per batch, from ITS OWN
// start (backward: its last contributing index, walking back to front) to its own end (forward: until all its
// pixels are opaque).  The gather (id -> 32-B geom + colour) of batch i+1 is in flight as
// cp.async into a double-buffered, lane-private landing zone while batch i is composited, and the ids of
// batch i+2 are loaded one batch ahead of that, so no global-load latency and no block barrier sits on the
// critical path (ncu on a CTA-staged version of this kernel: long_scoreboard + barrier = 20 % of the stall
// cycles).  Every lane culls its own entry directly against the warp's two 8x8 blocks.
struct WarpStream {
    float4 ra[2][32], rb[2][32], rc[2][32];  // landing zone: geom (2 x 16 B) + colour of the lane's entry, two batches
    float4 qa[32], qb[32], qc[32];           // queue of the batch being composited (see Staging)
    int qm[32];
};
// A loop-invariant per-lane value that ptxas would otherwise re-derive (S2R / I2FP / address arithmetic) inside the
// inner loop when registers are tight: routed through a self-shuffle it cannot be rematerialised, so it stays in a register.
__device__ __forceinline__ float pin_reg(float x) { return __shfl_sync(0xffffffffu, x, threadIdx.x & 31); }
__device__ __forceinline__ unsigned pin_reg(unsigned x) { return __shfl_sync(0xffffffffu, x, threadIdx.x & 31); }

template <int D>
__device__ __forceinline__ void stream_fetch(const RasterParams& p, WarpStream& ws, int buf, int lane, bool have, int g) {
    if (have) {
        cp_async16(&ws.ra[buf][lane], p.geom + (int64_t)g * 2);
        cp_async16(&ws.rb[buf][lane], p.geom + (int64_t)g * 2 + 1);
        if (D == 4) {
            cp_async16(&ws.rc[buf][lane], p.colors + (int64_t)g * 4);
        } else {
#pragma unroll
            for (int d = 0; d < D; ++d) cp_async4(reinterpret_cast<float*>(&ws.rc[buf][lane]) + d, p.colors + (int64_t)g * D + d);
        }
    }
    cp_async_commit();
}

// ------------------------------------------------------------------------------------------------
// forward, packed + warp streams (the default).  Same layout as the packed backward: 2 warps per tile, lane
// pixels at columns j0 + 8a, rows i0 + 4b, the two rows of an 8x8 block in the two halves of an f32x2.
// ------------------------------------------------------------------------------------------------
// `thr` is the per-pixel alpha threshold: 1/255 while the pixel is live, 2 (never reached) once it is opaque
// or outside the image -- "live" costs no instruction in the test.
// One pixel's accept / stop decision of the forward, as a chain of predicated compares (the C++ formulation compiles to
// ~8.5 compare / select / predicate-logic instructions per pixel, this is 6):
//   ok = (am >= thr) && (pw <= lo);  upd = ok && (nT > 1e-4);  al = upd ? am : 0;  last = upd ? sid : last;
//   thr = (ok && !upd) ? 2 : thr     (stop: the pixel is opaque, no later Gaussian passes `am >= 2`)
__device__ __forceinline__ float fwd_accept(float am, float pw, float lo, float nT, int sid, float& thr, int32_t& last) {
    float al;
    asm("{\n\t.reg .pred ok, upd;\n\t"
        "setp.ge.f32 ok, %3, %1;\n\t"
        "setp.le.and.f32 ok, %4, %5, ok;\n\t"
        "setp.gt.and.f32 upd, %6, 0f38D1B717, ok;\n\t"  // kTransmittanceThreshold
        "selp.f32 %0, %3, 0f00000000, upd;\n\t"
        "selp.b32 %2, %7, %2, upd;\n\t"
        "@ok selp.f32 %1, %1, 0f40000000, upd;\n\t}"
        : "=f"(al), "+f"(thr), "+r"(last)
        : "f"(am), "f"(pw), "f"(lo), "f"(nT), "r"(sid));
    return al;
}

template <int D, bool STATS>
__device__ __forceinline__ void fwd_pk_block(int a, const float4& A, const float4& B, const float4& Cc, float px0, f32x2 dy2, f32x2 t2,
                                             f32x2 cy2, f32x2& T2, f32x2 (&acc2)[D], int32_t (&last)[2], float (&thr)[2],
                                             StatCounters<STATS>& st) {
    const float dx = A.x - (px0 + 8.0f * a);
    const float ax = B.x * dx;
    const f32x2 pw2 = fma2(add2(t2, bc2(ax)), bc2(dx), cy2);  // log2(opacity * exp(-sigma))
    const float pw0 = lo2(pw2), pw1 = hi2(pw2);
    const float am0 = fminf(kMaxAlpha, ex2_approx(pw0)), am1 = fminf(kMaxAlpha, ex2_approx(pw1));
    const f32x2 nT2 = mul2(T2, sub2(bc2(1.0f), pk2(am0, am1)));
    const int sid = __float_as_int(A.w);
    float al0, al1;
    if (STATS) {
        const bool ok0 = (pw0 <= A.z) && (am0 >= thr[0]);  // pw <= lo  <=>  sigma >= 0
        const bool ok1 = (pw1 <= A.z) && (am1 >= thr[1]);
        const bool stop0 = ok0 && (lo2(nT2) <= kTransmittanceThreshold), stop1 = ok1 && (hi2(nT2) <= kTransmittanceThreshold);
        const bool upd0 = ok0 && !stop0, upd1 = ok1 && !stop1;
        st.add(3, (thr[0] < 1.0f ? 1 : 0) + (thr[1] < 1.0f ? 1 : 0));
        st.add(4, (upd0 ? 1 : 0) + (upd1 ? 1 : 0));
        al0 = upd0 ? am0 : 0.0f;
        al1 = upd1 ? am1 : 0.0f;
        last[0] = upd0 ? sid : last[0];
        last[1] = upd1 ? sid : last[1];
        thr[0] = stop0 ? 2.0f : thr[0];
        thr[1] = stop1 ? 2.0f : thr[1];
    } else {
        al0 = fwd_accept(am0, pw0, A.z, lo2(nT2), sid, thr[0], last[0]);
        al1 = fwd_accept(am1, pw1, A.z, hi2(nT2), sid, thr[1], last[1]);
    }
    const f32x2 al2 = pk2(al0, al1);  // alpha = 0: the pixel skips this Gaussian
    const f32x2 w2 = mul2(al2, T2);
#pragma unroll
    for (int d = 0; d < D; ++d) acc2[d] = fma2(bc2(d == 0 ? Cc.x : d == 1 ? Cc.y : d == 2 ? Cc.z : Cc.w), w2, acc2[d]);
    T2 = mul2(T2, sub2(bc2(1.0f), al2));  // == nT2 where updated (same two roundings), T * 1 elsewhere
}

// MINB resident CTAs per SM asked of ptxas: 12 -> 78 registers (default), 14 -> 72; the pixel coordinate stays in a register
// through pin_reg.  More resident warps do not help (the kernel is bound by issue slots and the fp32 / ALU pipes, not by
// latency): 16 CTAs (64 registers) measured 1 % slower, 12 1 % faster than 14.  qed_debug_set_raster_fwd_minb.
template <int D, bool CULL, bool STATS, int MINB>
__global__ void __launch_bounds__(Shape<4>::kThreads, MINB) raster_fwd_ws_kernel(const RasterParams p) {
    using S = Shape<4>;
    __shared__ WarpStream wss[S::kWarps];
    StatCounters<STATS> st;
    pdl_enter();
    const int cam = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpStream& ws = wss[warp];
    const int ox = tx * kTile, oy = ty * kTile + warp * S::kFootH;
    const int j0 = ox + (lane & 7), i0 = oy + (lane >> 3);
    const float px0 = pin_reg((float)j0 + 0.5f), py0 = (float)i0 + 0.5f;
    const f32x2 npy2 = pk2(-py0, -(py0 + 4.0f));

    const int64_t tile_id = ((int64_t)cam * p.tile_h + ty) * p.tile_w + tx;
    const int range_start = p.offsets[tile_id];
    const int range_end = (!p.offsets_has_end && tile_id == (int64_t)p.C * p.tile_h * p.tile_w - 1) ? (int)p.n_isects : p.offsets[tile_id + 1];
    const SubRects<2> sr = make_blocks8(ox, oy, p.width, p.height);

    f32x2 T2[2], acc2[2][D];
    int32_t last[2][2];
    float thr[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        T2[a] = pk2(1.0f, 1.0f);
#pragma unroll
        for (int d = 0; d < D; ++d) acc2[a][d] = pk2(0.0f, 0.0f);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            last[a][b] = 0;
            thr[a][b] = ((i0 + b * 4 < p.height) && (j0 + a * 8 < p.width)) ? kAlphaThreshold : 2.0f;
        }
    }
    int alive = sr.present;  // warp-uniform: 8x8 blocks that still have a live pixel

    // batch i covers sorted indices [range_start + 32 i, + 32), lane l owns range_start + 32 i + l
    const int n_batches = (range_end - range_start + 31) / 32;
    int e_cur = range_start + lane;
    int g_cur = e_cur < range_end ? p.flatten_ids[e_cur] : 0;
    stream_fetch<D>(p, ws, 0, lane, e_cur < range_end, g_cur);
    int g_nxt = (e_cur + 32 < range_end) ? p.flatten_ids[e_cur + 32] : 0;

    for (int i = 0; i < n_batches && alive; ++i, e_cur += 32) {
        const int buf = i & 1;
        stream_fetch<D>(p, ws, buf ^ 1, lane, e_cur + 32 < range_end, g_nxt);       // batch i+1: geom + colour
        const int g_n2 = (e_cur + 64 < range_end) ? p.flatten_ids[e_cur + 64] : 0;  // batch i+2: id
        cp_async_wait<1>();                                                         // batch i has landed
        const bool have = e_cur < range_end;
        int mask = 0;
        float4 A = make_float4(0, 0, 0, 0), B = make_float4(0, 0, 0, 0);
        st.add(0, have ? 1 : 0);
        if (have) {
            const float4 ga = ws.ra[buf][lane];  // mx, my, opacity, depth
            const float4 gb = ws.rb[buf][lane];  // conic a, b, c
            const float lo = __log2f(ga.z);
            A = make_float4(ga.x, ga.y, lo, __int_as_float(e_cur));
            B = make_float4(-0.5f * kLog2e * gb.x, -kLog2e * gb.y, -0.5f * kLog2e * gb.z, __int_as_float(g_cur));
            const float tau2 = lo + kLog2_255;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                bool hit = (alive >> k) & 1;
                if (CULL && hit) hit = ellipse_hits_rect(A.x, A.y, -B.x, -B.y, -B.z, tau2, sr.x0[k], sr.y0[k], sr.x1[k], sr.y1[k]);
                if (hit) mask |= 1 << k;
            }
        }
        st.add(1, mask ? 1 : 0);
        const uint32_t m = __ballot_sync(0xffffffffu, mask != 0);  // also: every lane is done with the previous queue
        if (mask) {
            const int pos = __popc(m & ((1u << lane) - 1u));
            float4 c = ws.rc[buf][lane];
            if (D < 4) c.w = 0.0f;
            if (D < 3) c.y = c.z = 0.0f;
            ws.qa[pos] = A;
            ws.qb[pos] = B;
            ws.qc[pos] = c;
            ws.qm[pos] = mask;
        }
        __syncwarp();
        const int nq = __popc(m);
        for (int q = 0; q < nq; ++q) {
            const float4 A = ws.qa[q], B = ws.qb[q], Cc = ws.qc[q];
            const int mask = ws.qm[q];
            if (lane == 0) st.add(2, __popc(mask));
            const f32x2 dy2 = add2(bc2(A.y), npy2);
            const f32x2 t2 = mul2(bc2(B.y), dy2);
            const f32x2 cy2 = fma2(mul2(bc2(B.z), dy2), dy2, bc2(A.z));
            // two separate basic blocks: with a fused both-blocks path ptxas no longer updates the accumulators in
            // place (18 register moves per Gaussian), and the forward has the occupancy to hide the latency instead
            if (mask & 1) fwd_pk_block<D, STATS>(0, A, B, Cc, px0, dy2, t2, cy2, T2[0], acc2[0], last[0], thr[0], st);
            if (mask & 2) fwd_pk_block<D, STATS>(1, A, B, Cc, px0, dy2, t2, cy2, T2[1], acc2[1], last[1], thr[1], st);
        }
        alive = (__any_sync(0xffffffffu, fminf(thr[0][0], thr[0][1]) < 1.0f) ? 1 : 0) | (__any_sync(0xffffffffu, fminf(thr[1][0], thr[1][1]) < 1.0f) ? 2 : 0);
        g_cur = g_nxt;
        g_nxt = g_n2;
    }
    cp_async_wait<0>();

#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int i = i0 + b * 4, j = j0 + a * 8;
            if (i < p.height && j < p.width) {
                const int64_t pix = ((int64_t)cam * p.height + i) * p.width + j;
                const float T = b ? hi2(T2[a]) : lo2(T2[a]);
                const float alpha_out = 1.0f - T;
                float out[D];
#pragma unroll
                for (int d = 0; d < D; ++d)
                    out[d] = (b ? hi2(acc2[a][d]) : lo2(acc2[a][d])) + (p.backgrounds ? T * p.backgrounds[cam * D + d] : 0.0f);
                if (p.normalize_last) out[D - 1] = out[D - 1] / fmaxf(alpha_out, 1e-10f);
                if (D == 4) {
                    reinterpret_cast<float4*>(p.render)[pix] = make_float4(out[0], out[1], out[2], out[3]);
                } else {
#pragma unroll
                    for (int d = 0; d < D; ++d) p.render[pix * D + d] = out[d];
                }
                p.alphas[pix] = alpha_out;
                p.last_ids[pix] = last[a][b];
            }
        }
    }
    st.flush(p.counters);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Sum the 12 per-lane slots v[0..11] over the 32 lanes with 16 shuffles (a transposed butterfly: each
// stage halves the number of slots a lane still carries).  Slots 4..7 take a plain butterfly in the first
// stage, so no lane ever carries padding.  Returns the total of slot `reduce_slot(lane)`; lanes for
// which reduce_lane_active(lane) is false hold a duplicate.
__device__ __forceinline__ int reduce_slot(int lane) {
    return ((lane & 8) ? 4 : ((lane & 16) ? 8 : 0)) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
}
__device__ __forceinline__ bool reduce_lane_active(int lane) { return !(lane & 1) && !((lane & 8) && (lane & 16)); }

__device__ __forceinline__ float warp_reduce_transpose12(const float (&v)[12], int lane) {
    float r8[8], r4[4], r2[2];
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // slots i <-> i+8 swap halves
        const float send = h4 ? v[i] : v[i + 8];
        const float keep = h4 ? v[i + 8] : v[i];
        r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 4; i < 8; ++i) r8[i] = v[i] + __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h3 ? r8[i] : r8[i + 4];
        const float keep = h3 ? r8[i + 4] : r8[i];
        r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h2 ? r4[i] : r4[i + 2];
        const float keep = h2 ? r4[i + 2] : r4[i];
        r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float send = h1 ? r2[0] : r2[1];
    const float keep = h1 ? r2[1] : r2[0];
    float r1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
    return r1;
}

template <int D, int PX, bool CULL, bool STATS>
__global__ void __launch_bounds__(Shape<PX>::kThreads) raster_bwd_kernel(const RasterParams p) {
    using S = Shape<PX>;
    __shared__ Staging<PX> sm;
    StatCounters<STATS> st;
    const int cam = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ox = tx * kTile + S::origin_x(warp), oy = ty * kTile + S::origin_y(warp);
    const int j0 = ox + (lane & 7), i0 = oy + (lane >> 3);
    const float px0 = (float)j0 + 0.5f, py0 = (float)i0 + 0.5f;

    const int64_t tile_id = ((int64_t)cam * p.tile_h + ty) * p.tile_w + tx;
    const int64_t range_start = p.offsets[tile_id];

    const float tx0 = (float)(tx * kTile) + 0.5f, ty0 = (float)(ty * kTile) + 0.5f;
    const float tx1 = (float)min(tx * kTile + kTile - 1, p.width - 1) + 0.5f, ty1 = (float)min(ty * kTile + kTile - 1, p.height - 1) + 0.5f;
    const SubRects<PX> sr = make_subrects<PX>(ox, oy, p.width, p.height);

    // per-lane role in the gradient reduction: which slot this lane ends up holding, and its scale back to
    // the (mean2d, conic, opacity, colour) parametrisation:
    //   v_mean = g u / log2e ; v_conic = -(g dx^2 / 2, g dx dy, g dy^2 / 2) ; v_opacity = g / opacity
    const int slot = reduce_slot(lane);
    const bool slot_active = reduce_lane_active(lane) && slot < 8 + D;
    const float slot_scale = slot < 4 ? (1.0f / kLog2e) : ((slot == 4 || slot == 6) ? -0.5f : (slot == 5 ? -1.0f : 1.0f));

    // per-pixel state: T (transmittance after the current Gaussian), bsum = sum_k buffer[k] v_out[k]
    // - T_final * v_alpha_out, the output gradients and the last composited index (-1 = nothing composited)
    float T[PX], bsum[PX], v_out[PX][D];
    int32_t bin_final[PX];
    int wmax = -1;
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        const int i = i0 + (k >> 1) * 4, j = j0 + (k & 1) * 8;
        T[k] = 1.0f;
        bsum[k] = 0.0f;
        bin_final[k] = -1;
#pragma unroll
        for (int d = 0; d < D; ++d) v_out[k][d] = 0.0f;
        if (i < p.height && j < p.width) {
            const int64_t pix = ((int64_t)cam * p.height + i) * p.width + j;
            const float alpha_out = p.alphas[pix];
            const float T_final = 1.0f - alpha_out;
            T[k] = T_final;
            if (T_final < 1.0f) bin_final[k] = p.last_ids[pix];
            if (D == 4) {
                const float4 v = reinterpret_cast<const float4*>(p.v_render)[pix];
                v_out[k][0] = v.x;
                v_out[k][1] = v.y;
                v_out[k][2] = v.z;
                v_out[k][3] = v.w;
            } else {
#pragma unroll
                for (int d = 0; d < D; ++d) v_out[k][d] = p.v_render[pix * D + d];
            }
            float v_alpha_out = p.v_alphas ? p.v_alphas[pix] : 0.0f;
            if (p.normalize_last) {
                // out_last = raw_last / max(alpha,1e-10)  (raw includes the background term)
                const float denom = fmaxf(alpha_out, 1e-10f);
                const float out_last = p.render[pix * D + D - 1];
                const float g = v_out[k][D - 1];
                v_out[k][D - 1] = g / denom;
                if (alpha_out > 1e-10f) v_alpha_out += -g * out_last / denom;
            }
            if (p.backgrounds) {
                // render = acc + T_final * bg  ->  d/dalpha_out of the bg term is -bg
                float s = 0.0f;
#pragma unroll
                for (int d = 0; d < D; ++d) s += p.backgrounds[cam * D + d] * v_out[k][d];
                v_alpha_out -= s;
            }
            bsum[k] = -T_final * v_alpha_out;
        }
        wmax = max(wmax, bin_final[k]);
    }

    // per-sub-rectangle, warp-wide and CTA-wide last contributing index
    int sub_max[PX];
    int want = 0;
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        int m = bin_final[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        sub_max[k] = m;
        if (m >= 0) want |= 1 << k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) sm.max_last[warp] = wmax;
    __syncthreads();
    int cta_max = -1;
#pragma unroll
    for (int w = 0; w < S::kWarps; ++w) cta_max = max(cta_max, sm.max_last[w]);
    if (cta_max < 0) return;

    // walk the range back to front: batch entries in descending sorted index
    for (int64_t b1 = (int64_t)cta_max + 1; b1 > range_start; b1 -= S::kThreads) {
        const int64_t e = b1 - 1 - threadIdx.x;
        const int count = stage_batch<D, PX, CULL, STATS>(p, sm, e >= range_start, e, tx0, ty0, tx1, ty1, st);
        for (int c0 = 0; c0 < count; c0 += 32) {
            const int nq = fill_queue<PX, CULL, true>(sm, c0, count, sr, want, sub_max);
            for (int q = 0; q < nq; ++q) {
                const float4 A = sm.qa[warp][q], B = sm.qb[warp][q], Cc = sm.qc[warp][q];
                const int mask = PX > 1 ? sm.qm[warp][q] : 1;
                if (lane == 0) st.add(2, __popc(mask));
                const int sid = __float_as_int(A.w);
                float dx[S::kNX], ax[S::kNX], dy[S::kNY], cy[S::kNY];
#pragma unroll
                for (int a = 0; a < S::kNX; ++a) {
                    dx[a] = A.x - (px0 + 8.0f * a);
                    ax[a] = B.x * dx[a];
                }
#pragma unroll
                for (int b = 0; b < S::kNY; ++b) {
                    dy[b] = A.y - (py0 + 4.0f * b);
                    cy[b] = fmaf(B.z * dy[b], dy[b], A.z);
                }
                // lane-local sums over this lane's PX pixels.  Slots (before the per-slot scale applied
                // after the reduction): 0,1 g*ux, g*uy | 2,3 |.| | 4,5,6 g dx^2, g dx dy, g dy^2 | 7 g |
                // 8.. fac * v_out, with g = alpha_raw * v_alpha, ux = 2 qa dx + qb dy, uy = qb dx + 2 qc dy
                float v[12];
#pragma unroll
                for (int s = 0; s < 12; ++s) v[s] = 0.0f;
                bool any_valid = false;
#pragma unroll
                for (int k = 0; k < PX; ++k) {
                    if (PX > 1 && !((mask >> k) & 1)) continue;  // warp-uniform
                    const int a = k & 1, b = k >> 1;
                    const float pw = fmaf(fmaf(B.y, dy[b], ax[a]), dx[a], cy[b]);
                    const float alpha_raw = ex2_approx(pw);
                    const float alpha = fminf(kMaxAlpha, alpha_raw);
                    const bool valid = (sid <= bin_final[k]) && (pw <= A.z) && (alpha >= kAlphaThreshold);
                    st.add(3, (sid <= bin_final[k]) ? 1 : 0);
                    st.add(4, valid ? 1 : 0);
                    any_valid |= valid;
                    const float ra = rcp_approx(1.0f - alpha);
                    const float Tn = valid ? T[k] * ra : T[k];  // transmittance in front of this Gaussian
                    const float fac = valid ? alpha * Tn : 0.0f;
                    float s1 = Cc.x * v_out[k][0];
                    v[8] = fmaf(fac, v_out[k][0], v[8]);
                    if (D >= 3) {
                        s1 = fmaf(Cc.y, v_out[k][1], s1);
                        s1 = fmaf(Cc.z, v_out[k][2], s1);
                        v[9] = fmaf(fac, v_out[k][1], v[9]);
                        v[10] = fmaf(fac, v_out[k][2], v[10]);
                    }
                    if (D == 4) {
                        s1 = fmaf(Cc.w, v_out[k][3], s1);
                        v[11] = fmaf(fac, v_out[k][3], v[11]);
                    }
                    const float v_alpha = fmaf(Tn, s1, -ra * bsum[k]);
                    bsum[k] = fmaf(fac, s1, bsum[k]);
                    T[k] = Tn;
                    const float g = (valid && alpha_raw <= kMaxAlpha) ? alpha_raw * v_alpha : 0.0f;
                    const float ux = fmaf(2.0f * B.x, dx[a], B.y * dy[b]);
                    const float uy = fmaf(2.0f * B.z, dy[b], B.y * dx[a]);
                    const float gx = g * ux, gy = g * uy;
                    v[0] += gx;
                    v[1] += gy;
                    v[2] += fabsf(gx);
                    v[3] += fabsf(gy);
                    const float gdx = g * dx[a], gdy = g * dy[b];
                    v[4] = fmaf(gdx, dx[a], v[4]);
                    v[5] = fmaf(gdx, dy[b], v[5]);
                    v[6] = fmaf(gdy, dy[b], v[6]);
                    v[7] += g;
                }
                if (!__any_sync(0xffffffffu, any_valid)) continue;
                if (lane == 0) st.add(5, 1);
                const float r = warp_reduce_transpose12(v, lane);
                const float inv_opac = ex2_approx(-A.z);
                if (slot_active) {
                    const float scale = (slot == 7) ? inv_opac : slot_scale;
                    atomicAdd(p.packed_grads + (int64_t)__float_as_int(B.w) * kGradFloats + slot, r * scale);
                }
            }
        }
        // no trailing barrier needed: stage_batch() only overwrites the staging buffers after its first
        // __syncthreads, which every warp reaches only after finishing this batch
    }
    st.flush(p.counters);
}

// ------------------------------------------------------------------------------------------------
// backward, packed (the default): the algorithm of raster_bwd_kernel re-cut for Blackwell's issue limits.
//   * per-pixel arithmetic on sm_100's two-wide fp32 instructions: FFMA2 / FMUL2 / FADD2 retire two IEEE
//     fp32 results per issue slot, with free scalar-broadcast, negate and |.| operand forms.  Layout =
//     Shape<4> (2 warps per tile, 16x8 footprint, lane pixels at columns j0 + 8a, rows i0 + 4b); a pair is
//     (a, b=0) / (a, b=1): the two pixels of one 8x8 block share dx and differ in dy.  Culling masks are per
//     8x8 block; when both blocks of a Gaussian are hit they run as one basic block (ILP).
//   * warp streams: the two warps of a tile never meet (no block barrier), see raster_bwd_ws_kernel.
//   * the cross-lane gradient sums go through a shared-memory transpose + red.global.add.v4.f32 instead of a
//     shuffle butterfly (reduce_entries).
// ------------------------------------------------------------------------------------------------
// One pixel's accept decision of the backward as a chain of predicated compares (4 compares + 2 selects; the C++
// formulation compiles to one compare AND one select per condition):
//   valid = (sid <= bin_final) && (pw <= lo) && (am >= 1/255);  al = valid ? am : 0;  ag = (valid && ar <= 0.999) ? am : 0
__device__ __forceinline__ void bwd_accept(int sid, int bin_final, float pw, float lo, float am, float ar, float& al, float& ag) {
    asm("{\n\t.reg .pred v, c;\n\t"
        "setp.le.s32 v, %2, %3;\n\t"
        "setp.le.and.f32 v, %4, %5, v;\n\t"
        "setp.ge.and.f32 v, %6, 0f3B808081, v;\n\t"  // kAlphaThreshold
        "selp.f32 %0, %6, 0f00000000, v;\n\t"
        "setp.le.and.f32 c, %7, 0f3F7FBE77, v;\n\t"  // kMaxAlpha
        "selp.f32 %1, %6, 0f00000000, c;\n\t}"
        : "=f"(al), "=f"(ag)
        : "r"(sid), "r"(bin_final), "f"(pw), "f"(lo), "f"(am), "f"(ar));
}

// One 8x8 block (column half `a`): the lane's two pixels (rows i0, i0 + 4) in the two halves of every
// f32x2.  FIRST: v[] is written, else accumulated (no zero-fill, no register shuffling where the paths
// join).  Returns nonzero if a pixel passed the alpha test.
template <int D, bool FIRST, bool STATS>
__device__ __forceinline__ int bwd_pk_block(int a, const float4& A, const float4& B, const float4& Cc, int sid, float px0, f32x2 dy2,
                                            f32x2 t2, f32x2 cy2, float qc2, f32x2& T2, f32x2& bsum2, const f32x2 (&vout2)[D],
                                            const int32_t (&bin_final)[2], f32x2 (&v)[12], StatCounters<STATS>& st) {
    const float dx = A.x - (px0 + 8.0f * a);
    const float ax = B.x * dx;
    const f32x2 pw2 = fma2(add2(t2, bc2(ax)), bc2(dx), cy2);  // log2(opacity * exp(-sigma))
    const float pw0 = lo2(pw2), pw1 = hi2(pw2);
    const float ar0 = ex2_approx(pw0), ar1 = ex2_approx(pw1);
    const float am0 = fminf(kMaxAlpha, ar0), am1 = fminf(kMaxAlpha, ar1);
    // a pixel that fails the test gets alpha = 0: then 1/(1-alpha) = 1 exactly (rcp.approx is exact at 1),
    // T and bsum pass through unchanged and every gradient term is 0 -- no selects after the packed ops
    float al0, al1, ag0, ag1;  // ag: alpha where it carries a gradient (not clamped to kMaxAlpha)
    bool valid0 = false, valid1 = false;
    if (STATS) {
        valid0 = (sid <= bin_final[0]) && (pw0 <= A.z) && (am0 >= kAlphaThreshold);
        valid1 = (sid <= bin_final[1]) && (pw1 <= A.z) && (am1 >= kAlphaThreshold);
        st.add(3, (sid <= bin_final[0] ? 1 : 0) + (sid <= bin_final[1] ? 1 : 0));
        st.add(4, (valid0 ? 1 : 0) + (valid1 ? 1 : 0));
        al0 = valid0 ? am0 : 0.0f;
        al1 = valid1 ? am1 : 0.0f;
        ag0 = ar0 <= kMaxAlpha ? al0 : 0.0f;
        ag1 = ar1 <= kMaxAlpha ? al1 : 0.0f;
    } else {
        bwd_accept(sid, bin_final[0], pw0, A.z, am0, ar0, al0, ag0);
        bwd_accept(sid, bin_final[1], pw1, A.z, am1, ar1, al1, ag1);
    }
    const f32x2 al2 = pk2(al0, al1);
    const f32x2 ag2 = pk2(ag0, ag1);
    const f32x2 oma2 = sub2(bc2(1.0f), al2);
    const f32x2 ra2 = pk2(rcp_approx(lo2(oma2)), rcp_approx(hi2(oma2)));
    const f32x2 Tn2 = mul2(T2, ra2);  // transmittance in front of this Gaussian
    const f32x2 fac2 = mul2(al2, Tn2);
    f32x2 s1 = mul2(bc2(Cc.x), vout2[0]);
    if (D >= 3) {
        s1 = fma2(bc2(Cc.y), vout2[1], s1);
        s1 = fma2(bc2(Cc.z), vout2[2], s1);
    }
    if (D == 4) s1 = fma2(bc2(Cc.w), vout2[3], s1);
    const f32x2 va2 = fma2(Tn2, s1, neg2(mul2(ra2, bsum2)));
    bsum2 = fma2(fac2, s1, bsum2);
    T2 = Tn2;
    const f32x2 g2 = mul2(ag2, va2);
    const f32x2 ux2 = add2(t2, bc2(ax + ax));               // 2 qa dx + qb dy
    const f32x2 uy2 = fma2(dy2, bc2(qc2), bc2(B.y * dx));  // qb dx + 2 qc dy
    const f32x2 gx2 = mul2(g2, ux2), gy2 = mul2(g2, uy2);
    const f32x2 gdx2 = mul2(g2, bc2(dx)), gdy2 = mul2(g2, dy2);
    if (FIRST) {
        v[0] = gx2;
        v[1] = gy2;
        v[2] = abs2(gx2);
        v[3] = abs2(gy2);
        v[4] = mul2(gdx2, bc2(dx));
        v[5] = mul2(gdx2, dy2);
        v[6] = mul2(gdy2, dy2);
        v[7] = g2;
#pragma unroll
        for (int d = 0; d < 4; ++d) v[8 + d] = d < D ? mul2(fac2, vout2[d < D ? d : 0]) : pk2(0.0f, 0.0f);
    } else {
        v[0] = add2(v[0], gx2);
        v[1] = add2(v[1], gy2);
        v[2] = add2(v[2], abs2(gx2));
        v[3] = add2(v[3], abs2(gy2));
        v[4] = fma2(gdx2, bc2(dx), v[4]);
        v[5] = fma2(gdx2, dy2, v[5]);
        v[6] = fma2(gdy2, dy2, v[6]);
        v[7] = add2(v[7], g2);
#pragma unroll
        for (int d = 0; d < D; ++d) v[8 + d] = fma2(fac2, vout2[d], v[8 + d]);
    }
    return (valid0 || valid1) ? 1 : 0;  // STATS only
}

// Per-pixel start state of the packed backward (warp footprint 16x8 at (j0 - lane%8, i0 - lane/8)): final
// transmittance, bsum = -T_final * dL/dalpha_out (ED normalisation and background folded in), output
// gradients, last composited index; per 8x8 block the warp-wide last index (sub_max) and `want` bits.
template <int D>
__device__ __forceinline__ void bwd_pk_prologue(const RasterParams& p, int cam, int i0, int j0, f32x2 (&T2)[2], f32x2 (&bsum2)[2],
                                                f32x2 (&vout2)[2][D], int32_t (&bin_final)[2][2], int (&sub_max)[2], int& want, int& wmax) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        float Ts[2], bs[2], vo[2][D];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int i = i0 + b * 4, j = j0 + a * 8;
            Ts[b] = 1.0f;
            bs[b] = 0.0f;
            bin_final[a][b] = -1;
#pragma unroll
            for (int d = 0; d < D; ++d) vo[b][d] = 0.0f;
            if (i < p.height && j < p.width) {
                const int64_t pix = ((int64_t)cam * p.height + i) * p.width + j;
                const float alpha_out = p.alphas[pix];
                const float T_final = 1.0f - alpha_out;
                Ts[b] = T_final;
                if (T_final < 1.0f) bin_final[a][b] = p.last_ids[pix];
                if (D == 4) {
                    const float4 v4 = reinterpret_cast<const float4*>(p.v_render)[pix];
                    vo[b][0] = v4.x;
                    vo[b][1] = v4.y;
                    vo[b][2] = v4.z;
                    vo[b][3] = v4.w;
                } else {
#pragma unroll
                    for (int d = 0; d < D; ++d) vo[b][d] = p.v_render[pix * D + d];
                }
                float v_alpha_out = p.v_alphas ? p.v_alphas[pix] : 0.0f;
                if (p.normalize_last) {
                    const float denom = fmaxf(alpha_out, 1e-10f);
                    const float out_last = p.render[pix * D + D - 1];
                    const float g = vo[b][D - 1];
                    vo[b][D - 1] = g / denom;
                    if (alpha_out > 1e-10f) v_alpha_out += -g * out_last / denom;
                }
                if (p.backgrounds) {
                    float sbg = 0.0f;
#pragma unroll
                    for (int d = 0; d < D; ++d) sbg += p.backgrounds[cam * D + d] * vo[b][d];
                    v_alpha_out -= sbg;
                }
                bs[b] = -T_final * v_alpha_out;
            }
        }
        T2[a] = pk2(Ts[0], Ts[1]);
        bsum2[a] = pk2(bs[0], bs[1]);
#pragma unroll
        for (int d = 0; d < D; ++d) vout2[a][d] = pk2(vo[0][d], vo[1][d]);
        int m = max(bin_final[a][0], bin_final[a][1]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        sub_max[a] = m;
        if (m >= 0) want |= 1 << a;
        wmax = max(wmax, m);
    }
}

// backward only -- gradient transpose: [entry parity * 3 + slot group][lane, 8 + 1 pad]
struct GradTranspose {
    float4 v[6][36];
};
constexpr int kRedRowBytes = 36 * 16;               // one [slot group] row
constexpr int kRedParityBytes = 3 * kRedRowBytes;  // the three rows of one queue-entry parity

// Position of lane l in a `red` row: every 8 lanes are followed by one float4 of padding, so that the eight
// lanes of a shared-memory phase hit eight different 16-B columns on the write AND on the transposed read.
__device__ __forceinline__ int red_pos(int l) { return (l >> 3) * 9 + (l & 7); }

// Shared-memory accesses by 32-bit shared-window address + immediate offset: the per-lane base addresses are computed
// once (pin_reg) instead of being re-derived from %tid / %ctaid inside the inner loop.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int OFF>
__device__ __forceinline__ float4 lds_v4(unsigned addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr), "n"(OFF) : "memory");
    return r;
}
template <int OFF>
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(r) : "r"(addr), "n"(OFF) : "memory");
    return r;
}
template <int OFF>
__device__ __forceinline__ int lds_b32(unsigned addr) {
    int r;
    asm volatile("ld.shared.b32 %0, [%1+%2];" : "=r"(r) : "r"(addr), "n"(OFF) : "memory");
    return r;
}

// The lane's 12 gradient values of one queue entry (two pixels per f32x2 -> one float) into its column of the transpose
// rows.  One copy per control-flow path of the caller (PATH only makes the three asm strings differ): sunk into a common
// tail, the compiler has to shuffle 12 registers into a common layout at the join.
#define QED_STS_V4(PATH)                                                                                                        \
    asm volatile("st.shared.v4.f32 [%0+%5], {%1,%2,%3,%4};  // path " #PATH ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "n"(OFF) \
                 : "memory")
template <int PATH, int OFF>
__device__ __forceinline__ void sts_v4(unsigned addr, float a, float b, float c, float d) {
    if (PATH == 0) QED_STS_V4(0);
    else if (PATH == 1) QED_STS_V4(1);
    else QED_STS_V4(2);
}
template <int PATH>
__device__ __forceinline__ void store_grad_slots(unsigned addr, const f32x2 (&v)[12]) {
    float vs[12];
#pragma unroll
    for (int s = 0; s < 12; ++s) vs[s] = lo2(v[s]) + hi2(v[s]);
    sts_v4<PATH, 0>(addr, vs[0], vs[1], vs[2], vs[3]);
    sts_v4<PATH, kRedRowBytes>(addr, vs[4], vs[5], vs[6], vs[7]);
    sts_v4<PATH, 2 * kRedRowBytes>(addr, vs[8], vs[9], vs[10], vs[11]);
}

// Role of a lane in the cross-lane sums: lane = (entry parity qq, slot group sg, quarter) sums 8 lanes' float4 of slot
// group sg of queue entry q0 + qq with FADD2 and adds its quarter-sum with ONE vector red.global.add.v4.f32 -- instead
// of a shuffle butterfly per Gaussian (16 SHFL + 22 FSEL + 16 FADD).  All per-lane constants of the role live in registers.
struct ReduceRole {
    unsigned src;    // shared address of red.v[qq * 3 + sg][quarter * 9]
    unsigned entry;  // shared address of ws.qa[qq]
    int need;        // the lane works iff (entries in this round) > need
    unsigned sg16;   // 16 * slot group (0 = means2d + absgrad, 1 = conic + opacity, 2 = colour): byte offset of the group
};                   // inside the packed record and of its row in kSlotScale
// back to the (mean2d, conic, opacity, colour) parametrisation, per slot group:
//   v_mean = g u / log2e ; v_conic = -(g dx^2 / 2, g dx dy, g dy^2 / 2) ; v_opacity = g / opacity (.w of group 1, patched per Gaussian)
__constant__ float4 kSlotScale[3] = {{1.0f / kLog2e, 1.0f / kLog2e, 1.0f / kLog2e, 1.0f / kLog2e}, {-0.5f, -1.0f, -0.5f, 1.0f}, {1.0f, 1.0f, 1.0f, 1.0f}};

__device__ __forceinline__ ReduceRole make_reduce_role(const WarpStream& ws, const GradTranspose& red, int lane) {
    ReduceRole r;
    const int item = min(lane >> 2, 5), quarter = lane & 3;  // item = entry parity * 3 + slot group
    const int qq = item >= 3 ? 1 : 0;
    r.sg16 = (unsigned)(item - qq * 3) * 16u;
    r.src = pin_reg(smem_u32(&red.v[item][quarter * 9]));
    r.entry = pin_reg(smem_u32(&ws.qa[qq]));
    r.need = lane < 24 ? qq : 2;
    return r;
}

// Cross-lane sums of the gradient slots of queue entries q0 (even) .. q0+count-1 (count <= 2), whose per-lane values
// sit in `red`.
__device__ __forceinline__ void reduce_entries(const RasterParams& p, const ReduceRole& r, int q0, int count) {
    if (count > r.need) {
        float4 x = lds_v4<0>(r.src);
        f32x2 s01 = pk2(x.x, x.y), s23 = pk2(x.z, x.w);
#define QED_RED_STEP(J)                 \
        x = lds_v4<16 * (J)>(r.src);    \
        s01 = add2(s01, pk2(x.x, x.y)); \
        s23 = add2(s23, pk2(x.z, x.w));
        QED_RED_STEP(1) QED_RED_STEP(2) QED_RED_STEP(3) QED_RED_STEP(4) QED_RED_STEP(5) QED_RED_STEP(6) QED_RED_STEP(7)
#undef QED_RED_STEP
        const unsigned e = r.entry + q0 * 16;
        const int g = lds_b32<512 + 12>(e);  // WarpStream::qb[q].w
        const char* ks = reinterpret_cast<const char*>(kSlotScale) + r.sg16;
        const float2 k01 = *reinterpret_cast<const float2*>(ks);
        float k3 = *reinterpret_cast<const float*>(ks + 12);
        if (r.sg16 == 16u) k3 = ex2_approx(-lds_f32<8>(e));  // WarpStream::qa[q].z = log2(opacity)
        s01 = mul2(s01, pk2(k01.x, k01.y));
        s23 = mul2(s23, pk2(k01.x, k3));
        char* dst = reinterpret_cast<char*>(p.packed_grads) + ((uint64_t)(unsigned)g * (kGradFloats * 4) + r.sg16);
        red_add_v4(reinterpret_cast<float*>(dst), lo2(s01), hi2(s01), lo2(s23), hi2(s23));
    }
}

// MINB resident CTAs per SM asked of ptxas: 10 -> 96 registers, no spills (default); 8 -> 102 registers, 3 % slower; 12 ->
// 80 registers with spills, 1 % slower (qed_debug_set_raster_bwd_minb).
template <int D, bool CULL, bool STATS, int MINB>
__global__ void __launch_bounds__(Shape<4>::kThreads, MINB) raster_bwd_ws_kernel(const RasterParams p) {
    using S = Shape<4>;
    __shared__ WarpStream wss[S::kWarps];
    __shared__ GradTranspose reds[S::kWarps];
    static_assert(offsetof(WarpStream, qb) - offsetof(WarpStream, qa) == 512, "reduce_entries addresses qb relative to qa");
    StatCounters<STATS> st;
    pdl_enter();
    const int cam = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpStream& ws = wss[warp];
    GradTranspose& red = reds[warp];
    const int ox = tx * kTile, oy = ty * kTile + warp * S::kFootH;
    const int j0 = ox + (lane & 7), i0 = oy + (lane >> 3);
    const float px0 = pin_reg((float)j0 + 0.5f), py0 = (float)i0 + 0.5f;
    const f32x2 npy2 = pk2(-py0, -(py0 + 4.0f));

    const int64_t tile_id = ((int64_t)cam * p.tile_h + ty) * p.tile_w + tx;
    const int range_start = p.offsets[tile_id];
    const SubRects<2> sr = make_blocks8(ox, oy, p.width, p.height);

    f32x2 T2[2], bsum2[2], vout2[2][D];
    int32_t bin_final[2][2];
    int sub_max[2];
    int want = 0, wmax = -1;
    bwd_pk_prologue<D>(p, cam, i0, j0, T2, bsum2, vout2, bin_final, sub_max, want, wmax);
    if (wmax < 0) return;  // nothing composited under this warp's footprint (no block-wide barrier below)

    const unsigned red_dst = pin_reg(smem_u32(&red.v[0][red_pos(lane)]));  // this lane's column of the transpose rows
    const ReduceRole role = make_reduce_role(ws, red, lane);

    // batch i covers sorted indices (top - 32 i - 31 .. top - 32 i], lane l owns top - 32 i - l
    const int top = wmax;
    const int n_batches = (top - range_start) / 32 + 1;
    int e_cur = top - lane;
    int g_cur = e_cur >= range_start ? p.flatten_ids[e_cur] : 0;
    stream_fetch<D>(p, ws, 0, lane, e_cur >= range_start, g_cur);
    int g_nxt = (e_cur - 32 >= range_start) ? p.flatten_ids[e_cur - 32] : 0;

    for (int i = 0; i < n_batches; ++i, e_cur -= 32) {
        const int buf = i & 1;
        stream_fetch<D>(p, ws, buf ^ 1, lane, e_cur - 32 >= range_start, g_nxt);       // batch i+1: geom + colour
        const int g_n2 = (e_cur - 64 >= range_start) ? p.flatten_ids[e_cur - 64] : 0;  // batch i+2: id
        cp_async_wait<1>();                                                            // batch i has landed
        const bool have = e_cur >= range_start;
        int mask = 0;
        float4 A = make_float4(0, 0, 0, 0), B = make_float4(0, 0, 0, 0);
        st.add(0, have ? 1 : 0);
        if (have && want) {
            const float4 ga = ws.ra[buf][lane];  // mx, my, opacity, depth
            const float4 gb = ws.rb[buf][lane];  // conic a, b, c
            const float lo = __log2f(ga.z);
            A = make_float4(ga.x, ga.y, lo, __int_as_float(e_cur));
            B = make_float4(-0.5f * kLog2e * gb.x, -kLog2e * gb.y, -0.5f * kLog2e * gb.z, __int_as_float(g_cur));
            const float tau2 = lo + kLog2_255;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                bool hit = ((want >> k) & 1) && (e_cur <= sub_max[k]);
                if (CULL && hit) hit = ellipse_hits_rect(A.x, A.y, -B.x, -B.y, -B.z, tau2, sr.x0[k], sr.y0[k], sr.x1[k], sr.y1[k]);
                if (hit) mask |= 1 << k;
            }
        }
        st.add(1, mask ? 1 : 0);
        const uint32_t m = __ballot_sync(0xffffffffu, mask != 0);  // also: every lane is done with the previous queue
        if (mask) {
            const int pos = __popc(m & ((1u << lane) - 1u));
            float4 c = ws.rc[buf][lane];
            if (D < 4) c.w = 0.0f;
            if (D < 3) c.y = c.z = 0.0f;
            ws.qa[pos] = A;
            ws.qb[pos] = B;
            ws.qc[pos] = c;
            ws.qm[pos] = mask;
        }
        __syncwarp();
        const int nq = __popc(m);
        unsigned par = 0;  // byte offset of the transpose rows of this entry's parity
        for (int q = 0; q < nq; ++q, par ^= kRedParityBytes) {
            const float4 A = ws.qa[q], B = ws.qb[q], Cc = ws.qc[q];
            const int mask = ws.qm[q];
            if (lane == 0) st.add(2, __popc(mask));
            const int sid = __float_as_int(A.w);
            const f32x2 dy2 = add2(bc2(A.y), npy2);
            const f32x2 t2 = mul2(bc2(B.y), dy2);
            const f32x2 cy2 = fma2(mul2(bc2(B.z), dy2), dy2, bc2(A.z));
            const float qc2 = B.z + B.z;
            const unsigned dst = red_dst + par;
            f32x2 v[12];
            int any_valid;
            if (mask == 3) {  // one basic block: the two independent blocks interleave (ILP)
                any_valid = bwd_pk_block<D, true, STATS>(0, A, B, Cc, sid, px0, dy2, t2, cy2, qc2, T2[0], bsum2[0], vout2[0], bin_final[0], v, st);
                any_valid |= bwd_pk_block<D, false, STATS>(1, A, B, Cc, sid, px0, dy2, t2, cy2, qc2, T2[1], bsum2[1], vout2[1], bin_final[1], v, st);
                store_grad_slots<0>(dst, v);
            } else if (mask == 1) {
                any_valid = bwd_pk_block<D, true, STATS>(0, A, B, Cc, sid, px0, dy2, t2, cy2, qc2, T2[0], bsum2[0], vout2[0], bin_final[0], v, st);
                store_grad_slots<1>(dst, v);
            } else {
                any_valid = bwd_pk_block<D, true, STATS>(1, A, B, Cc, sid, px0, dy2, t2, cy2, qc2, T2[1], bsum2[1], vout2[1], bin_final[1], v, st);
                store_grad_slots<2>(dst, v);
            }
            if (STATS) {  // every lane votes
                const bool any = __any_sync(0xffffffffu, any_valid);
                if (lane == 0 && any) st.add(5, 1);
            }
            if ((q & 1) || q == nq - 1) {
                __syncwarp();
                reduce_entries(p, role, q & ~1, (q & 1) + 1);
                __syncwarp();
            }
        }
        g_cur = g_nxt;
        g_nxt = g_n2;
    }
    cp_async_wait<0>();
    st.flush(p.counters);
}

// ------------------------------------------------------------------------------------------------
// pack / unpack helpers
// ------------------------------------------------------------------------------------------------
__global__ void pack_geom_kernel(int64_t CN, const float2* __restrict__ means2d, const float* __restrict__ conics,
                                 const float* __restrict__ opacities, const float* __restrict__ depths, float4* __restrict__ geom) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CN) return;
    float2 m = means2d[i];
    geom[i * 2] = make_float4(m.x, m.y, opacities[i], depths ? depths[i] : 0.0f);
    geom[i * 2 + 1] = make_float4(conics[i * 3], conics[i * 3 + 1], conics[i * 3 + 2], 0.0f);
}

__global__ void unpack_grads_kernel(int64_t CN, int D, const float4* __restrict__ packed, float2* __restrict__ v_means2d,
                                    float2* __restrict__ v_abs, float* __restrict__ v_conics, float* __restrict__ v_colors,
                                    float* __restrict__ v_opacities) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CN) return;
    float4 r0 = packed[i * 3], r1 = packed[i * 3 + 1], r2 = packed[i * 3 + 2];
    if (v_means2d) v_means2d[i] = make_float2(r0.x, r0.y);
    if (v_abs) v_abs[i] = make_float2(r0.z, r0.w);
    if (v_conics) {
        v_conics[i * 3] = r1.x;
        v_conics[i * 3 + 1] = r1.y;
        v_conics[i * 3 + 2] = r1.z;
    }
    if (v_opacities) v_opacities[i] = r1.w;
    if (v_colors) {
        float c[4] = {r2.x, r2.y, r2.z, r2.w};
        for (int k = 0; k < D; ++k) v_colors[i * D + k] = c[k];
    }
}

// test / instrumentation hooks (not part of the reference surface).  THREAD-LOCAL: a test or the benchmark's pair
// counters change them for the calling host thread only, so another thread (nerfstudio's viewer rendering through
// the same library) never sees them -- the library holds no cross-thread mutable state.
static thread_local int g_raster_cull = 1;                           // 0 disables the culling (identical results, slower)
static thread_local unsigned long long* g_raster_counters = nullptr;  // device uint64[6] -> STATS kernels
static thread_local int g_px_fwd = 4, g_px_bwd = 4;                   // pixels per lane
static thread_local int g_raster_packed = 1;                          // f32x2 kernels where they exist (backward, 4 px/lane)
constexpr int kBwdMinBlocks = 10, kBwdMinBlocksAlt = 8;               // resident CTAs per SM asked of ptxas (96 / 102 registers; 12 = 80 registers was 1 % slower)
static thread_local int g_bwd_minb = kBwdMinBlocks;
constexpr int kFwdMinBlocks = 12, kFwdMinBlocksAlt = 14;              // forward: 78 / 72 registers (measured at S1: 12 is 1 % faster than 14, 16 = 64 registers 1 % slower)
static thread_local int g_fwd_minb = kFwdMinBlocks;

template <int D, int PX, bool BWD>
static void launch_raster_px(const RasterParams& p, cudaStream_t stream) {
    dim3 grid(p.tile_w, p.tile_h, p.C);
    constexpr int T = Shape<PX>::kThreads;
    const bool cull = g_raster_cull != 0, stats = p.counters != nullptr;
    if (!BWD && PX == 4 && g_raster_packed) {
        if (stats) {
            if (cull) (void)launch_pdl(raster_fwd_ws_kernel<D, true, true, kFwdMinBlocks>, grid, dim3(T), 0, stream, p);
            else (void)launch_pdl(raster_fwd_ws_kernel<D, false, true, kFwdMinBlocks>, grid, dim3(T), 0, stream, p);
        } else if (!cull) {
            (void)launch_pdl(raster_fwd_ws_kernel<D, false, false, kFwdMinBlocks>, grid, dim3(T), 0, stream, p);
        } else if (g_fwd_minb == kFwdMinBlocksAlt) {
            (void)launch_pdl(raster_fwd_ws_kernel<D, true, false, kFwdMinBlocksAlt>, grid, dim3(T), 0, stream, p);
        } else {
            (void)launch_pdl(raster_fwd_ws_kernel<D, true, false, kFwdMinBlocks>, grid, dim3(T), 0, stream, p);
        }
    } else if (!BWD) {
        if (stats) {
            if (cull) raster_fwd_kernel<D, PX, true, true><<<grid, T, 0, stream>>>(p);
            else raster_fwd_kernel<D, PX, false, true><<<grid, T, 0, stream>>>(p);
        } else {
            if (cull) raster_fwd_kernel<D, PX, true, false><<<grid, T, 0, stream>>>(p);
            else raster_fwd_kernel<D, PX, false, false><<<grid, T, 0, stream>>>(p);
        }
    } else if (PX == 4 && g_raster_packed) {
        if (stats) {
            if (cull) (void)launch_pdl(raster_bwd_ws_kernel<D, true, true, kBwdMinBlocks>, grid, dim3(T), 0, stream, p);
            else (void)launch_pdl(raster_bwd_ws_kernel<D, false, true, kBwdMinBlocks>, grid, dim3(T), 0, stream, p);
        } else if (!cull) {
            (void)launch_pdl(raster_bwd_ws_kernel<D, false, false, kBwdMinBlocks>, grid, dim3(T), 0, stream, p);
        } else if (g_bwd_minb == kBwdMinBlocksAlt) {
            (void)launch_pdl(raster_bwd_ws_kernel<D, true, false, kBwdMinBlocksAlt>, grid, dim3(T), 0, stream, p);
        } else {
            (void)launch_pdl(raster_bwd_ws_kernel<D, true, false, kBwdMinBlocks>, grid, dim3(T), 0, stream, p);
        }
    } else {
        if (stats) {
            if (cull) raster_bwd_kernel<D, PX, true, true><<<grid, T, 0, stream>>>(p);
            else raster_bwd_kernel<D, PX, false, true><<<grid, T, 0, stream>>>(p);
        } else {
            if (cull) raster_bwd_kernel<D, PX, true, false><<<grid, T, 0, stream>>>(p);
            else raster_bwd_kernel<D, PX, false, false><<<grid, T, 0, stream>>>(p);
        }
    }
}

template <int D, bool BWD>
static int launch_raster(RasterParams p, cudaStream_t stream) {
    p.counters = g_raster_counters;
    const int px = BWD ? g_px_bwd : g_px_fwd;
    switch (px) {
        case 1: launch_raster_px<D, 1, BWD>(p, stream); break;
        case 2: launch_raster_px<D, 2, BWD>(p, stream); break;
        default: launch_raster_px<D, 4, BWD>(p, stream); break;
    }
    QED_LAUNCH_CHECK();
    return QED_OK;
}

static inline bool misaligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

static int check_raster_args(int C, int N, int64_t n_isects, int D, int width, int height, int tile_size, int tile_width, int tile_height) {
    if (C < 0 || N < 0 || n_isects < 0 || width <= 0 || height <= 0) return QED_ERR_BAD_ARG;
    if (!(D == 1 || D == 3 || D == 4)) return QED_ERR_UNSUPPORTED;
    if (tile_size != kTile) return QED_ERR_UNSUPPORTED;
    if (tile_width != (width + kTile - 1) / kTile || tile_height != (height + kTile - 1) / kTile) return QED_ERR_BAD_ARG;
    if (tile_height > 65535 || C > 65535) return QED_ERR_UNSUPPORTED;
    if (n_isects > 0x7fffffffLL) return QED_ERR_UNSUPPORTED;
    return QED_OK;
}

static RasterParams make_params(int C, int N, int64_t n_isects, int D, const float* geom, const float* colors, const float* backgrounds,
                                int width, int height, int tile_width, int tile_height, const int32_t* isect_offsets,
                                const int32_t* flatten_ids, int normalize_last) {
    RasterParams p{};
    p.C = C;
    p.N = N;
    p.D = D;
    p.width = width;
    p.height = height;
    p.tile_w = tile_width;
    p.tile_h = tile_height;
    p.normalize_last = normalize_last;
    p.n_isects = n_isects;
    p.geom = reinterpret_cast<const float4*>(geom);
    p.colors = colors;
    p.backgrounds = backgrounds;
    p.offsets = isect_offsets;
    p.flatten_ids = flatten_ids;
    return p;
}

}  // namespace qed

using namespace qed;

extern "C" int qed_debug_set_raster_cull(int enabled) {
    int old = g_raster_cull;
    g_raster_cull = enabled ? 1 : 0;
    return old;
}

extern "C" int qed_debug_set_raster_counters(void* counters) {
    g_raster_counters = reinterpret_cast<unsigned long long*>(counters);
    return QED_OK;
}

// nonzero: two-wide fp32 (f32x2) kernels where they exist; 0: scalar kernels only.  Returns the previous value.
extern "C" int qed_debug_set_raster_packed(int enabled) {
    int old = g_raster_packed;
    g_raster_packed = enabled ? 1 : 0;
    return old;
}

// occupancy / register trade-off of the packed backward (10 or 8 resident CTAs per SM).  Returns the previous value.
extern "C" int qed_debug_set_raster_bwd_minb(int minb) {
    int old = g_bwd_minb;
    if (minb == kBwdMinBlocks || minb == kBwdMinBlocksAlt) g_bwd_minb = minb;
    return old;
}

extern "C" int qed_debug_set_raster_fwd_minb(int minb) {
    int old = g_fwd_minb;
    if (minb == kFwdMinBlocks || minb == kFwdMinBlocksAlt) g_fwd_minb = minb;
    return old;
}

// pixels per lane of the forward / backward compositor (1, 2 or 4; 0 keeps the current value)
extern "C" int qed_debug_set_raster_px(int px_fwd, int px_bwd) {
    if (px_fwd == 1 || px_fwd == 2 || px_fwd == 4) g_px_fwd = px_fwd;
    if (px_bwd == 1 || px_bwd == 2 || px_bwd == 4) g_px_bwd = px_bwd;
    return g_px_fwd * 10 + g_px_bwd;
}

extern "C" int qed_raster_fwd(int C, int N, int64_t n_isects, int D, const float* geom, const float* colors,
                              const float* backgrounds, int width, int height, int tile_size, int tile_width,
                              int tile_height, const int32_t* isect_offsets, int offsets_has_end, const int32_t* flatten_ids,
                              int normalize_last, float* render, float* alphas, int32_t* last_ids, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_raster_args(C, N, n_isects, D, width, height, tile_size, tile_width, tile_height);
    if (rc != QED_OK) return rc;
    if (C == 0) return QED_OK;
    if (!isect_offsets || !render || !alphas || !last_ids) return QED_ERR_BAD_ARG;
    if (n_isects > 0 && (!geom || !colors || !flatten_ids)) return QED_ERR_BAD_ARG;
    if (misaligned16(geom) || (D == 4 && (misaligned16(colors) || misaligned16(render)))) return QED_ERR_BAD_ARG;  // float4 access
    RasterParams p = make_params(C, N, n_isects, D, geom, colors, backgrounds, width, height, tile_width, tile_height, isect_offsets,
                                 flatten_ids, normalize_last);
    p.render = render;
    p.alphas = alphas;
    p.last_ids = last_ids;
    p.offsets_has_end = offsets_has_end ? 1 : 0;
    switch (D) {
        case 1: return launch_raster<1, false>(p, stream);
        case 3: return launch_raster<3, false>(p, stream);
        default: return launch_raster<4, false>(p, stream);
    }
}

extern "C" int qed_raster_bwd(int C, int N, int64_t n_isects, int D, const float* geom, const float* colors,
                              const float* backgrounds, int width, int height, int tile_size, int tile_width,
                              int tile_height, const int32_t* isect_offsets, const int32_t* flatten_ids,
                              int normalize_last, const float* render, const float* alphas, const int32_t* last_ids,
                              const float* v_render, const float* v_alphas, float* packed_grads, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_raster_args(C, N, n_isects, D, width, height, tile_size, tile_width, tile_height);
    if (rc != QED_OK) return rc;
    if (C == 0 || n_isects == 0) return QED_OK;
    if (!isect_offsets || !render || !alphas || !last_ids || !v_render || !packed_grads || !geom || !colors || !flatten_ids)
        return QED_ERR_BAD_ARG;
    if (misaligned16(geom) || (D == 4 && (misaligned16(colors) || misaligned16(v_render)))) return QED_ERR_BAD_ARG;  // float4 access
    RasterParams p = make_params(C, N, n_isects, D, geom, colors, backgrounds, width, height, tile_width, tile_height, isect_offsets,
                                 flatten_ids, normalize_last);
    p.render = const_cast<float*>(render);  // read-only in the backward kernel
    p.alphas = const_cast<float*>(alphas);
    p.last_ids = const_cast<int32_t*>(last_ids);
    p.v_render = v_render;
    p.v_alphas = v_alphas;
    p.packed_grads = packed_grads;
    switch (D) {
        case 1: return launch_raster<1, true>(p, stream);
        case 3: return launch_raster<3, true>(p, stream);
        default: return launch_raster<4, true>(p, stream);
    }
}

extern "C" int qed_pack_geom(int CN, const float* means2d, const float* conics, const float* opacities,
                             const float* depths, float* geom, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (CN < 0) return QED_ERR_BAD_ARG;
    if (CN == 0) return QED_OK;
    if (!means2d || !conics || !opacities || !geom) return QED_ERR_BAD_ARG;
    pack_geom_kernel<<<(CN + 255) / 256, 256, 0, stream>>>(CN, reinterpret_cast<const float2*>(means2d), conics, opacities, depths,
                                                          reinterpret_cast<float4*>(geom));
    QED_LAUNCH_CHECK();
    return QED_OK;
}

extern "C" int qed_unpack_grads(int CN, int D, const float* packed_grads, float* v_means2d, float* v_means2d_abs,
                                float* v_conics, float* v_colors, float* v_opacities, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (CN < 0 || !(D == 1 || D == 3 || D == 4)) return QED_ERR_BAD_ARG;
    if (CN == 0) return QED_OK;
    if (!packed_grads) return QED_ERR_BAD_ARG;
    unpack_grads_kernel<<<(CN + 255) / 256, 256, 0, stream>>>(CN, D, reinterpret_cast<const float4*>(packed_grads),
                                                             reinterpret_cast<float2*>(v_means2d), reinterpret_cast<float2*>(v_means2d_abs),
                                                             v_conics, v_colors, v_opacities);
    QED_LAUNCH_CHECK();
    return QED_OK;
}
