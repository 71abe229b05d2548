// (c)/(d) Per-tile front-to-back alpha compositing of D channels (RGB + depth), forward and backward.
//
// Replaces gsplat's rasterize_to_pixels_fwd/_bwd (+ the ED normalisation tail of rasterization()) behind
// qed_splatter/model.py:267-288; absgrad (model.py:284) is produced by the backward.
// Semantics: SURVEY.md Appendix A.5 / A.6 == oracle/torch_impl.py::rasterize_to_pixels(_bwd).
//
// Design (B200-first, not gsplat's):
//   * one CTA per 16x16 tile (the tile size is fixed by the bit-exact intersection contract), 8 warps,
//     each warp owns an 8x4 pixel sub-rectangle (compact footprint -> warp-level culling pays).
//   * Gaussians of the tile's sorted range are gathered through flatten_ids in batches of 256 as 32-byte
//     geometry records + D-float colour records (L2-resident) into shared memory.
//   * EXACT two-level culling: a Gaussian only changes a pixel if opacity*exp(-sigma) >= 1/255, i.e. the
//     pixel centre lies inside the ellipse sigma <= ln(255*opacity).  The loader thread tests that ellipse
//     against the whole tile (survivors are compacted in order with ballots), then every warp tests the
//     survivors against its own 8x4 rectangle, 32 candidates at a time with one ballot, and only iterates
//     over the set bits.  The test is conservative (minimum of sigma over the rectangle, plus a margin for
//     rounding), so results are bit-identical to the un-culled kernel (template CULL=false, kept for tests).
//   * early termination: per-lane `done`, warp exit on __all_sync(done), CTA exit on __syncthreads_count.
//   * backward: back-to-front replay from the stored last index; the 12-float per-Gaussian gradient
//     record is reduced across the warp with a transposed butterfly (16 shuffles instead of 60), then one
//     red.global.add per value from 12 lanes into the packed [C*N,12] record.
// Bound: FP32 issue + MUFU (ex2) pipes, not HBM (~30 FLOP + 1 ex2 per evaluated pixel-Gaussian pair).
#include "common.cuh"

namespace qed {

constexpr int kTile = 16;
constexpr int kRasterThreads = kTile * kTile;
constexpr int kBatch = kRasterThreads;

struct RasterParams {
    int C, N, D, width, height, tile_w, tile_h, normalize_last;
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
// candidates after the warp-level cull, [3] lane evaluations by live lanes, [4] pairs that passed the
// alpha test (composited / differentiated), [5] (warp, Gaussian) groups that ran the gradient reduction
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

// Conservative test: can a Gaussian reach alpha >= 1/255 at any pixel centre in [x0,x1]x[y0,y1]?
// sigma(d) = 0.5*(a dx^2 + c dy^2) + b dx dy with d = mean - pixel.  Returns false only if the minimum of
// sigma over the rectangle exceeds tau = ln(255*opacity) by a safety margin.
__device__ __forceinline__ bool ellipse_hits_rect(float mx, float my, float a, float b, float c, float tau, float x0, float y0,
                                                  float x1, float y1) {
    if (!(tau >= 0.0f)) return false;  // opacity < 1/255 can never pass the alpha test
    // d ranges
    const float dxl = mx - x1, dxh = mx - x0;  // dx in [dxl, dxh]
    const float dyl = my - y1, dyh = my - y0;
    if (dxl <= 0.0f && dxh >= 0.0f && dyl <= 0.0f && dyh >= 0.0f) return true;  // centre inside
    // the quadratic is convex: its minimum over the rectangle (centre outside) lies on an edge
    float best = 3.0e38f;
    float mag = 0.0f;
    const float inv_c = 1.0f / c, inv_a = 1.0f / a;
    {  // edges dx = dxl / dxh : minimise over dy
        float dxs[2] = {dxl, dxh};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float dx = dxs[e];
            float dy = fminf(fmaxf(-b * dx * inv_c, dyl), dyh);
            float t0 = 0.5f * a * dx * dx, t1 = 0.5f * c * dy * dy, t2 = b * dx * dy;
            float s = t0 + t1 + t2;
            if (s < best) {
                best = s;
                mag = t0 + t1 + fabsf(t2);
            }
        }
    }
    {  // edges dy = dyl / dyh : minimise over dx
        float dys[2] = {dyl, dyh};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float dy = dys[e];
            float dx = fminf(fmaxf(-b * dy * inv_a, dxl), dxh);
            float t0 = 0.5f * a * dx * dx, t1 = 0.5f * c * dy * dy, t2 = b * dx * dy;
            float s = t0 + t1 + t2;
            if (s < best) {
                best = s;
                mag = t0 + t1 + fabsf(t2);
            }
        }
    }
    // margin: 1% + absolute + rounding of the three terms (cancellation for thin, tilted ellipses)
    return !(best > tau * 1.01f + 0.02f + 1e-5f * mag);
}

template <int D>
__device__ __forceinline__ void load_color(const float* __restrict__ colors, int64_t g, float* out) {
    if (D == 4) {
        float4 v = reinterpret_cast<const float4*>(colors)[g];
        out[0] = v.x;
        out[1] = v.y;
        out[2] = v.z;
        out[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) out[k] = colors[g * D + k];
    }
}

// Shared staging of one batch.  Entries are stored in processing order; with CULL only tile-level
// survivors are kept (order preserved), `sid` is the entry's index in the sorted intersection list.
template <int D>
struct Batch {
    float4 ga[kBatch];  // mx, my, opacity, tau
    float4 gb[kBatch];  // conic a, b, c, (bits) flat gaussian id
    float col[kBatch][D];
    int32_t sid[kBatch];
    int warp_count[kRasterThreads / 32];
};

// Loads entry `e` (sorted index) for this thread (or nothing if !have), runs the tile-level cull and
// writes survivors compacted in thread order.  Returns the number of entries staged.  Needs all threads.
template <int D, bool CULL, bool STATS>
__device__ __forceinline__ int stage_batch(const RasterParams& p, Batch<D>& sb, bool have, int64_t e, float tx0, float ty0, float tx1,
                                           float ty1, StatCounters<STATS>& st) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 A = make_float4(0, 0, 0, -1.0f), B = make_float4(0, 0, 0, 0);
    int64_t g = 0;
    bool keep = false;
    if (have) {
        g = p.flatten_ids[e];
        A = p.geom[g * 2];
        B = p.geom[g * 2 + 1];
        float tau = __logf(255.0f * A.z);
        A.w = tau;
        keep = CULL ? ellipse_hits_rect(A.x, A.y, B.x, B.y, B.z, tau, tx0, ty0, tx1, ty1) : true;
    }
    st.add(0, have ? 1 : 0);
    st.add(1, keep ? 1 : 0);
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) sb.warp_count[warp] = __popc(m);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kRasterThreads / 32; ++w) {
        int cnt = sb.warp_count[w];
        if (w < warp) base += cnt;
        total += cnt;
    }
    if (keep) {
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        B.w = __int_as_float((int)g);
        sb.ga[pos] = A;
        sb.gb[pos] = B;
        load_color<D>(p.colors, g, sb.col[pos]);
        sb.sid[pos] = (int32_t)e;
    }
    __syncthreads();
    return total;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int D, bool CULL, bool STATS>
__global__ void __launch_bounds__(kRasterThreads) raster_fwd_kernel(const RasterParams p) {
    __shared__ Batch<D> sb;
    StatCounters<STATS> st;
    const int cam = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub_x = (warp & 1) * 8, sub_y = (warp >> 1) * 4;
    const int j = tx * kTile + sub_x + (lane & 7);   // pixel column
    const int i = ty * kTile + sub_y + (lane >> 3);  // pixel row
    const bool inside = (i < p.height) && (j < p.width);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;

    const int64_t tile_id = ((int64_t)cam * p.tile_h + ty) * p.tile_w + tx;
    const int64_t range_start = p.offsets[tile_id];
    const int64_t range_end = (tile_id == (int64_t)p.C * p.tile_h * p.tile_w - 1) ? p.n_isects : (int64_t)p.offsets[tile_id + 1];

    // pixel-centre rectangles for the culling tests (clipped to the image)
    const float tx0 = (float)(tx * kTile) + 0.5f, ty0 = (float)(ty * kTile) + 0.5f;
    const float tx1 = (float)min(tx * kTile + kTile - 1, p.width - 1) + 0.5f, ty1 = (float)min(ty * kTile + kTile - 1, p.height - 1) + 0.5f;
    const float wx0 = (float)(tx * kTile + sub_x) + 0.5f, wy0 = (float)(ty * kTile + sub_y) + 0.5f;
    const float wx1 = fminf((float)(tx * kTile + sub_x + 7) + 0.5f, tx1), wy1 = fminf((float)(ty * kTile + sub_y + 3) + 0.5f, ty1);
    const bool warp_has_pixels = (wx0 <= tx1) && (wy0 <= ty1);

    float T = 1.0f;
    float acc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] = 0.0f;
    int32_t last = 0;
    bool done = !inside;

    for (int64_t b0 = range_start; b0 < range_end; b0 += kBatch) {
        if (__syncthreads_count(done) == kRasterThreads) break;
        const int64_t e = b0 + threadIdx.x;
        const int count = stage_batch<D, CULL, STATS>(p, sb, e < range_end, e, tx0, ty0, tx1, ty1, st);
        if (!__all_sync(0xffffffffu, done)) {
            for (int c0 = 0; c0 < count; c0 += 32) {
                uint32_t cand;
                if (CULL) {
                    const int q = c0 + lane;
                    bool hit = false;
                    if (q < count && warp_has_pixels) {
                        const float4 A = sb.ga[q], B = sb.gb[q];
                        hit = ellipse_hits_rect(A.x, A.y, B.x, B.y, B.z, A.w, wx0, wy0, wx1, wy1);
                    }
                    cand = __ballot_sync(0xffffffffu, hit);
                } else {
                    const int rem = count - c0;
                    cand = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
                }
                if (lane == 0) st.add(2, __popc(cand));
                while (cand) {
                    const int q = c0 + __ffs(cand) - 1;
                    cand &= cand - 1;
                    const float4 A = sb.ga[q], B = sb.gb[q];
                    const float dx = A.x - px, dy = A.y - py;
                    const float sigma = 0.5f * (B.x * dx * dx + B.z * dy * dy) + B.y * dx * dy;
                    const float vis = __expf(-sigma);
                    const float alpha = fminf(kMaxAlpha, A.z * vis);
                    st.add(3, done ? 0 : 1);
                    st.add(4, (!done && sigma >= 0.0f && alpha >= kAlphaThreshold) ? 1 : 0);
                    if (!done && sigma >= 0.0f && alpha >= kAlphaThreshold) {
                        const float next_T = T * (1.0f - alpha);
                        if (next_T <= kTransmittanceThreshold) {
                            done = true;
                        } else {
                            const float w = alpha * T;
#pragma unroll
                            for (int k = 0; k < D; ++k) acc[k] += sb.col[q][k] * w;
                            last = sb.sid[q];
                            T = next_T;
                        }
                    }
                }
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
    }

    if (inside) {
        const int64_t pix = ((int64_t)cam * p.height + i) * p.width + j;
        const float alpha_out = 1.0f - T;
        float out[D];
#pragma unroll
        for (int k = 0; k < D; ++k) out[k] = acc[k] + (p.backgrounds ? T * p.backgrounds[cam * D + k] : 0.0f);
        if (p.normalize_last) out[D - 1] = out[D - 1] / fmaxf(alpha_out, 1e-10f);
        if (D == 4) {
            reinterpret_cast<float4*>(p.render)[pix] = make_float4(out[0], out[1], out[2], out[3]);
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) p.render[pix * D + k] = out[k];
        }
        p.alphas[pix] = alpha_out;
        p.last_ids[pix] = last;
    }
    st.flush(p.counters);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Sum 16 per-lane slots over the 32 lanes with 16 shuffles; lane l returns the total of slot l>>1.
__device__ __forceinline__ float warp_reduce_transpose16(float (&v)[16], int lane) {
    float r8[8], r4[4], r2[2];
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = h4 ? v[i] : v[i + 8];
        const float keep = h4 ? v[i + 8] : v[i];
        r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
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

template <int D, bool CULL, bool STATS>
__global__ void __launch_bounds__(kRasterThreads) raster_bwd_kernel(const RasterParams p) {
    __shared__ Batch<D> sb;
    StatCounters<STATS> st;
    __shared__ int s_max_last[kRasterThreads / 32];
    const int cam = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub_x = (warp & 1) * 8, sub_y = (warp >> 1) * 4;
    const int j = tx * kTile + sub_x + (lane & 7);
    const int i = ty * kTile + sub_y + (lane >> 3);
    const bool inside = (i < p.height) && (j < p.width);
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;

    const int64_t tile_id = ((int64_t)cam * p.tile_h + ty) * p.tile_w + tx;
    const int64_t range_start = p.offsets[tile_id];
    const int64_t range_end = (tile_id == (int64_t)p.C * p.tile_h * p.tile_w - 1) ? p.n_isects : (int64_t)p.offsets[tile_id + 1];

    const float tx0 = (float)(tx * kTile) + 0.5f, ty0 = (float)(ty * kTile) + 0.5f;
    const float tx1 = (float)min(tx * kTile + kTile - 1, p.width - 1) + 0.5f, ty1 = (float)min(ty * kTile + kTile - 1, p.height - 1) + 0.5f;
    const float wx0 = (float)(tx * kTile + sub_x) + 0.5f, wy0 = (float)(ty * kTile + sub_y) + 0.5f;
    const float wx1 = fminf((float)(tx * kTile + sub_x + 7) + 0.5f, tx1), wy1 = fminf((float)(ty * kTile + sub_y + 3) + 0.5f, ty1);
    const bool warp_has_pixels = (wx0 <= tx1) && (wy0 <= ty1);

    // per-pixel state
    float T_final = 1.0f, v_alpha_out = 0.0f;
    float v_out[D], buffer[D];
    int32_t bin_final = -1;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        v_out[k] = 0.0f;
        buffer[k] = 0.0f;
    }
    if (inside) {
        const int64_t pix = ((int64_t)cam * p.height + i) * p.width + j;
        const float alpha_out = p.alphas[pix];
        T_final = 1.0f - alpha_out;
        if (T_final < 1.0f) bin_final = p.last_ids[pix];  // pixels nothing was composited into stay at -1
        if (D == 4) {
            float4 v = reinterpret_cast<const float4*>(p.v_render)[pix];
            v_out[0] = v.x;
            v_out[1] = v.y;
            v_out[2] = v.z;
            v_out[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) v_out[k] = p.v_render[pix * D + k];
        }
        v_alpha_out = p.v_alphas ? p.v_alphas[pix] : 0.0f;
        if (p.normalize_last) {
            // out_last = raw_last / max(alpha,1e-10)  (raw includes the background term)
            const float denom = fmaxf(alpha_out, 1e-10f);
            const float out_last = p.render[pix * D + D - 1];
            const float g = v_out[D - 1];
            v_out[D - 1] = g / denom;
            if (alpha_out > 1e-10f) v_alpha_out += -g * out_last / denom;
        }
        if (p.backgrounds) {
            // render = acc + T_final * bg  ->  d/dalpha_out of the bg term is -bg
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < D; ++k) s += p.backgrounds[cam * D + k] * v_out[k];
            v_alpha_out -= s;
        }
    }
    float T = T_final;

    // CTA-wide last contributing index
    int wmax = bin_final;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) s_max_last[warp] = wmax;
    __syncthreads();
    int cta_max = -1;
#pragma unroll
    for (int w = 0; w < kRasterThreads / 32; ++w) cta_max = max(cta_max, s_max_last[w]);
    if (cta_max < 0) return;
    const int warp_max = wmax;

    // walk the range back to front: batch entries in descending sorted index
    for (int64_t b1 = (int64_t)cta_max + 1; b1 > range_start; b1 -= kBatch) {
        const int64_t e = b1 - 1 - threadIdx.x;
        const int count = stage_batch<D, CULL, STATS>(p, sb, e >= range_start, e, tx0, ty0, tx1, ty1, st);
        for (int c0 = 0; c0 < count; c0 += 32) {
            uint32_t cand;
            {
                const int q = c0 + lane;
                bool hit = false;
                if (q < count && warp_has_pixels && sb.sid[q] <= warp_max) {
                    if (CULL) {
                        const float4 A = sb.ga[q], B = sb.gb[q];
                        hit = ellipse_hits_rect(A.x, A.y, B.x, B.y, B.z, A.w, wx0, wy0, wx1, wy1);
                    } else {
                        hit = true;
                    }
                }
                cand = __ballot_sync(0xffffffffu, hit);
            }
            if (lane == 0) st.add(2, __popc(cand));
            while (cand) {
                const int q = c0 + __ffs(cand) - 1;
                cand &= cand - 1;
                const float4 A = sb.ga[q], B = sb.gb[q];
                const float dx = A.x - px, dy = A.y - py;
                const float sigma = 0.5f * (B.x * dx * dx + B.z * dy * dy) + B.y * dx * dy;
                const float vis = __expf(-sigma);
                const float opac_vis = A.z * vis;
                const float alpha = fminf(kMaxAlpha, opac_vis);
                const bool valid = (sb.sid[q] <= bin_final) && (sigma >= 0.0f) && (alpha >= kAlphaThreshold);
                st.add(3, (sb.sid[q] <= bin_final) ? 1 : 0);
                st.add(4, valid ? 1 : 0);
                if (!__any_sync(0xffffffffu, valid)) continue;
                if (lane == 0) st.add(5, 1);
                float v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = 0.0f;
                if (valid) {
                    const float ra = 1.0f / (1.0f - alpha);
                    T *= ra;
                    const float fac = alpha * T;
                    float v_alpha = 0.0f;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const float ck = sb.col[q][k];
                        v[8 + k] = fac * v_out[k];
                        v_alpha += (ck * T - buffer[k] * ra) * v_out[k];
                        buffer[k] += ck * fac;
                    }
                    v_alpha += T_final * ra * v_alpha_out;
                    if (opac_vis <= kMaxAlpha) {
                        const float v_sigma = -opac_vis * v_alpha;
                        v[4] = 0.5f * v_sigma * dx * dx;
                        v[5] = v_sigma * dx * dy;
                        v[6] = 0.5f * v_sigma * dy * dy;
                        const float gx = v_sigma * (B.x * dx + B.y * dy);
                        const float gy = v_sigma * (B.y * dx + B.z * dy);
                        v[0] = gx;
                        v[1] = gy;
                        v[2] = fabsf(gx);
                        v[3] = fabsf(gy);
                        v[7] = vis * v_alpha;
                    }
                }
                const float r = warp_reduce_transpose16(v, lane);
                const int slot = lane >> 1;
                if ((lane & 1) == 0 && slot < 8 + D) {
                    const int64_t g = __float_as_int(B.w);
                    atomicAdd(p.packed_grads + g * kGradFloats + slot, r);
                }
            }
        }
        // no trailing barrier needed: stage_batch() only overwrites the staging buffers after its first
        // __syncthreads, which every warp reaches only after finishing this batch
    }
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

static int g_raster_cull = 1;  // test hook: 0 disables the culling (bit-identical results, slower)
static unsigned long long* g_raster_counters = nullptr;  // instrumentation hook: device uint64[6] or null

template <int D>
static int launch_raster(RasterParams p, bool backward, cudaStream_t stream) {
    dim3 grid(p.tile_w, p.tile_h, p.C);
    p.counters = g_raster_counters;
    if (g_raster_counters) {  // instrumented variant (never used in timed regions)
        if (!backward) {
            if (g_raster_cull)
                raster_fwd_kernel<D, true, true><<<grid, kRasterThreads, 0, stream>>>(p);
            else
                raster_fwd_kernel<D, false, true><<<grid, kRasterThreads, 0, stream>>>(p);
        } else {
            if (g_raster_cull)
                raster_bwd_kernel<D, true, true><<<grid, kRasterThreads, 0, stream>>>(p);
            else
                raster_bwd_kernel<D, false, true><<<grid, kRasterThreads, 0, stream>>>(p);
        }
    } else if (!backward) {
        if (g_raster_cull)
            raster_fwd_kernel<D, true, false><<<grid, kRasterThreads, 0, stream>>>(p);
        else
            raster_fwd_kernel<D, false, false><<<grid, kRasterThreads, 0, stream>>>(p);
    } else {
        if (g_raster_cull)
            raster_bwd_kernel<D, true, false><<<grid, kRasterThreads, 0, stream>>>(p);
        else
            raster_bwd_kernel<D, false, false><<<grid, kRasterThreads, 0, stream>>>(p);
    }
    QED_LAUNCH_CHECK();
    return QED_OK;
}

static int check_raster_args(int C, int N, int64_t n_isects, int D, int width, int height, int tile_size, int tile_width, int tile_height) {
    if (C < 0 || N < 0 || n_isects < 0 || width <= 0 || height <= 0) return QED_ERR_BAD_ARG;
    if (!(D == 1 || D == 3 || D == 4)) return QED_ERR_UNSUPPORTED;
    if (tile_size != kTile) return QED_ERR_UNSUPPORTED;
    if (tile_width != (width + kTile - 1) / kTile || tile_height != (height + kTile - 1) / kTile) return QED_ERR_BAD_ARG;
    if (tile_height > 65535 || C > 65535) return QED_ERR_UNSUPPORTED;
    if (n_isects > 0x7fffffffLL) return QED_ERR_UNSUPPORTED;
    return QED_OK;
}

}  // namespace qed

using namespace qed;

// test hook (not part of the reference surface): toggles the exact culling
extern "C" int qed_debug_set_raster_cull(int enabled) {
    int old = g_raster_cull;
    g_raster_cull = enabled ? 1 : 0;
    return old;
}

// instrumentation hook (not part of the reference surface): counters = device uint64[6] or NULL
extern "C" int qed_debug_set_raster_counters(void* counters) {
    g_raster_counters = reinterpret_cast<unsigned long long*>(counters);
    return QED_OK;
}

extern "C" int qed_raster_fwd(int C, int N, int64_t n_isects, int D, const float* geom, const float* colors,
                              const float* backgrounds, int width, int height, int tile_size, int tile_width,
                              int tile_height, const int32_t* isect_offsets, const int32_t* flatten_ids,
                              int normalize_last, float* render, float* alphas, int32_t* last_ids, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_raster_args(C, N, n_isects, D, width, height, tile_size, tile_width, tile_height);
    if (rc != QED_OK) return rc;
    if (C == 0) return QED_OK;
    if (!isect_offsets || !render || !alphas || !last_ids) return QED_ERR_BAD_ARG;
    if (n_isects > 0 && (!geom || !colors || !flatten_ids)) return QED_ERR_BAD_ARG;
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
    p.render = render;
    p.alphas = alphas;
    p.last_ids = last_ids;
    switch (D) {
        case 1: return launch_raster<1>(p, false, stream);
        case 3: return launch_raster<3>(p, false, stream);
        default: return launch_raster<4>(p, false, stream);
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
    p.render = const_cast<float*>(render);  // read-only in the backward kernel
    p.alphas = const_cast<float*>(alphas);
    p.last_ids = const_cast<int32_t*>(last_ids);
    p.v_render = v_render;
    p.v_alphas = v_alphas;
    p.packed_grads = packed_grads;
    switch (D) {
        case 1: return launch_raster<1>(p, true, stream);
        case 3: return launch_raster<3>(p, true, stream);
        default: return launch_raster<4>(p, true, stream);
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
