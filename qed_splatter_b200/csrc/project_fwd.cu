// (a) Fused projection + EWA 2-D covariance + SH colour + tile count, forward.
//
// Replaces gsplat's fully_fused_projection_fwd + compute_sh_fwd + the rasterization() glue
// (clamp_min(c+0.5), depth concat, opacity*compensation) + the count pass of isect_tiles, i.e. the first
// stages behind the call at qed_splatter/model.py:267-288.  Semantics: SURVEY.md Appendix A.1-A.3;
// float op order mirrors oracle/torch_impl.py::fully_fused_projection exactly (individually rounded
// IEEE ops via __f*_rn), so radii / tile counts / sort keys are bit-identical to the oracle.
//
// Mapping: one thread per Gaussian, looping over the cameras of the launch (Sigma is built once per
// Gaussian).  SH coefficients (192 B per Gaussian at K=16) are staged per warp into shared memory with
// coalesced 16-byte cp.async copies, only for Gaussians visible in at least one camera, into rows padded
// to an odd number of float4 so the per-lane LDS.128 reads are bank-conflict free.
// HBM-bound: ~44 B read per (camera, Gaussian), +192 B SH and +~100 B written per visible one.
#include "common.cuh"

namespace qed {

constexpr int kProjThreads = 256;
constexpr int kMaxCamsPerLaunch = 32;
constexpr int kCamFloats = 32;

struct Cam {
    float W[9];
    float t[3];
    float fx, fy, cx, cy;
    float lim_xp, lim_xn, lim_yp, lim_yn;
    float campos[3];
    float pad[9];
};
static_assert(sizeof(Cam) == kCamFloats * 4, "Cam layout");

struct ProjFwdParams {
    int C, c0, Cc, N, K, sh_degree, colors_per_camera, width, height;
    float eps2d, near_plane, far_plane, radius_clip;
    int calc_comp, activations;
    float tile_size;
    int tile_w, tile_h, n_color, append_depth;
    const float *means, *quats, *scales, *opacities, *colors_in, *viewmats, *Ks;
    int32_t* radii;
    float *means2d, *depths, *conics, *comps, *colors_out, *opac_out;
    int32_t* tiles;
    int32_t* tiles_exact;  // optional: tiles that can be reached with alpha >= 1/255 (sum of exact_row_span over the bounding box rows)
    float* geom;
};

// Per-camera constants, same op order as the oracle (camera_positions + the limits in A.1).
__device__ void load_camera(const float* __restrict__ V, const float* __restrict__ Kc, int width, int height, Cam& cam) {
    float a[3][3];
    float t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            a[i][j] = V[i * 4 + j];
            cam.W[i * 3 + j] = a[i][j];
        }
        t[i] = V[i * 4 + 3];
        cam.t[i] = t[i];
    }
    float fx = Kc[0], fy = Kc[4], cx = Kc[2], cy = Kc[5];
    cam.fx = fx;
    cam.fy = fy;
    cam.cx = cx;
    cam.cy = cy;
    float tanx = dvd(0.5f * (float)width, fx);
    float tany = dvd(0.5f * (float)height, fy);
    cam.lim_xp = add(dvd(sub((float)width, cx), fx), mul(0.3f, tanx));
    cam.lim_xn = add(dvd(cx, fx), mul(0.3f, tanx));
    cam.lim_yp = add(dvd(sub((float)height, cy), fy), mul(0.3f, tany));
    cam.lim_yn = add(dvd(cy, fy), mul(0.3f, tany));
    // campos = -(adj(A) t) / det(A)
    float c00 = msb2(a[1][1], a[2][2], a[1][2], a[2][1]);
    float c01 = msb2(a[0][2], a[2][1], a[0][1], a[2][2]);
    float c02 = msb2(a[0][1], a[1][2], a[0][2], a[1][1]);
    float c10 = msb2(a[1][2], a[2][0], a[1][0], a[2][2]);
    float c11 = msb2(a[0][0], a[2][2], a[0][2], a[2][0]);
    float c12 = msb2(a[0][2], a[1][0], a[0][0], a[1][2]);
    float c20 = msb2(a[1][0], a[2][1], a[1][1], a[2][0]);
    float c21 = msb2(a[0][1], a[2][0], a[0][0], a[2][1]);
    float c22 = msb2(a[0][0], a[1][1], a[0][1], a[1][0]);
    float det = mad3(a[0][0], c00, a[0][1], c10, a[0][2], c20);
    cam.campos[0] = -dvd(mad3(c00, t[0], c01, t[1], c02, t[2]), det);
    cam.campos[1] = -dvd(mad3(c10, t[0], c11, t[1], c12, t[2]), det);
    cam.campos[2] = -dvd(mad3(c20, t[0], c21, t[1], c22, t[2]), det);
}

// Sloan fast SH bases, same op order as oracle sh_bases().
template <int DEG>
__device__ __forceinline__ void sh_bases(float x, float y, float z, float* b) {
    b[0] = 0.2820947917738781f;
    if (DEG >= 1) {
        b[1] = mul(-0.48860251190292f, y);
        b[2] = mul(0.48860251190292f, z);
        b[3] = mul(-0.48860251190292f, x);
    }
    if (DEG >= 2) {
        float z2 = mul(z, z);
        float fTmp0B = mul(-1.092548430592079f, z);
        float fC1 = sub(mul(x, x), mul(y, y));
        float fS1 = mul(2.0f, mul(x, y));
        b[4] = mul(0.5462742152960395f, fS1);
        b[5] = mul(fTmp0B, y);
        b[6] = sub(mul(0.9461746957575601f, z2), 0.3153915652525201f);
        b[7] = mul(fTmp0B, x);
        b[8] = mul(0.5462742152960395f, fC1);
        if (DEG >= 3) {
            float fTmp0C = add(mul(-2.285228997322329f, z2), 0.4570457994644658f);
            float fTmp1B = mul(1.445305721320277f, z);
            float fC2 = sub(mul(x, fC1), mul(y, fS1));
            float fS2 = add(mul(x, fS1), mul(y, fC1));
            b[9] = mul(-0.5900435899266435f, fS2);
            b[10] = mul(fTmp1B, fS1);
            b[11] = mul(fTmp0C, y);
            b[12] = mul(z, sub(mul(1.865881662950577f, z2), 1.119528997770346f));
            b[13] = mul(fTmp0C, x);
            b[14] = mul(fTmp1B, fC1);
            b[15] = mul(-0.5900435899266435f, fC2);
        }
    }
}

template <int DEG>
struct ShShape {
    static constexpr int kBases = (DEG + 1) * (DEG + 1);
    static constexpr int kFloats = 3 * kBases;
    static constexpr int kVec = (kFloats + 3) / 4;            // float4 per row actually read
    static constexpr int kStrideVec = (kVec % 2) ? kVec : kVec + 1;  // odd -> conflict-free LDS.128
    static constexpr int kStrideScalar = (kFloats % 2) ? kFloats : kFloats + 1;
};

// DEG = -1: colours pass through (or no colour channels at all).  VEC: 16-byte staging path.
// ONE != 0: a single view (the reference's case, model.py:211) -- the view loops and the visibility mask collapse at compile
// time (78 -> 68 registers); ONE = resident blocks per SM asked of ptxas (4: 64 registers, 32 warps per SM).
template <int DEG, bool VEC, int ONE = 0>
__global__ void __launch_bounds__(kProjThreads, ONE ? ONE : 1) project_fwd_kernel(const ProjFwdParams p) {
    extern __shared__ float4 smem4[];
    pdl_enter();
    const int nC = ONE ? 1 : p.Cc, c0 = ONE ? 0 : p.c0;
    Cam* cams = reinterpret_cast<Cam*>(smem4);
    float* shbuf = reinterpret_cast<float*>(smem4 + nC * (kCamFloats / 4));

    if (threadIdx.x < nC) {
        int c = c0 + threadIdx.x;
        load_camera(p.viewmats + c * 16, p.Ks + c * 9, p.width, p.height, cams[threadIdx.x]);
    }
    __syncthreads();

    const int n = blockIdx.x * kProjThreads + threadIdx.x;
    const bool in_range = n < p.N;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int D = p.n_color + p.append_depth;

    float S00 = 0, S01 = 0, S02 = 0, S11 = 0, S12 = 0, S22 = 0;
    float m0 = 0, m1 = 0, m2 = 0, opac = 0;
    if (in_range) {
        m0 = p.means[n * 3 + 0];
        m1 = p.means[n * 3 + 1];
        m2 = p.means[n * 3 + 2];
        float w = p.quats[n * 4 + 0], x = p.quats[n * 4 + 1], y = p.quats[n * 4 + 2], z = p.quats[n * 4 + 3];
        float s0 = p.scales[n * 3 + 0], s1 = p.scales[n * 3 + 1], s2 = p.scales[n * 3 + 2];
        opac = p.opacities ? p.opacities[n] : 0.0f;
        // parameters as the model stores them (qed_splatter/model.py:269-271): exp / sigmoid folded in
        if (p.activations & QED_ACT_LOG_SCALES) {
            s0 = expf(s0);
            s1 = expf(s1);
            s2 = expf(s2);
        }
        if ((p.activations & QED_ACT_LOGIT_OPACITIES) && p.opacities) opac = dvd(1.0f, add(1.0f, expf(-opac)));
        float nrm = fmaxf(sqr(add(add(add(mul(w, w), mul(x, x)), mul(y, y)), mul(z, z))), 1e-12f);
        w = dvd(w, nrm);
        x = dvd(x, nrm);
        y = dvd(y, nrm);
        z = dvd(z, nrm);
        float xx = mul(x, x), yy = mul(y, y), zz = mul(z, z);
        float xy = mul(x, y), xz = mul(x, z), yz = mul(y, z);
        float wx = mul(w, x), wy = mul(w, y), wz = mul(w, z);
        float R00 = sub(1.0f, mul(2.0f, add(yy, zz))), R01 = mul(2.0f, sub(xy, wz)), R02 = mul(2.0f, add(xz, wy));
        float R10 = mul(2.0f, add(xy, wz)), R11 = sub(1.0f, mul(2.0f, add(xx, zz))), R12 = mul(2.0f, sub(yz, wx));
        float R20 = mul(2.0f, sub(xz, wy)), R21 = mul(2.0f, add(yz, wx)), R22 = sub(1.0f, mul(2.0f, add(xx, yy)));
        float M00 = mul(R00, s0), M01 = mul(R01, s1), M02 = mul(R02, s2);
        float M10 = mul(R10, s0), M11 = mul(R11, s1), M12 = mul(R12, s2);
        float M20 = mul(R20, s0), M21 = mul(R21, s1), M22 = mul(R22, s2);
        S00 = mad3(M00, M00, M01, M01, M02, M02);
        S01 = mad3(M00, M10, M01, M11, M02, M12);
        S02 = mad3(M00, M20, M01, M21, M02, M22);
        S11 = mad3(M10, M10, M11, M11, M12, M12);
        S12 = mad3(M10, M20, M11, M21, M12, M22);
        S22 = mad3(M20, M20, M21, M21, M22, M22);
    }

    uint32_t vismask = 0;
    for (int ci = 0; ci < nC; ++ci) {
        if (!in_range) break;
        const Cam& cam = cams[ci];
        const float* W = cam.W;
        const int64_t idx = (int64_t)(c0 + ci) * p.N + n;
        float px = add(mad3(W[0], m0, W[1], m1, W[2], m2), cam.t[0]);
        float py = add(mad3(W[3], m0, W[4], m1, W[5], m2), cam.t[1]);
        float pz = add(mad3(W[6], m0, W[7], m1, W[8], m2), cam.t[2]);
        // A = W Sigma ; Sc = A W^T (upper triangle)
        float A00 = mad3(W[0], S00, W[1], S01, W[2], S02), A01 = mad3(W[0], S01, W[1], S11, W[2], S12), A02 = mad3(W[0], S02, W[1], S12, W[2], S22);
        float A10 = mad3(W[3], S00, W[4], S01, W[5], S02), A11 = mad3(W[3], S01, W[4], S11, W[5], S12), A12 = mad3(W[3], S02, W[4], S12, W[5], S22);
        float A20 = mad3(W[6], S00, W[7], S01, W[8], S02), A21 = mad3(W[6], S01, W[7], S11, W[8], S12), A22 = mad3(W[6], S02, W[7], S12, W[8], S22);
        float Sc00 = mad3(A00, W[0], A01, W[1], A02, W[2]);
        float Sc01 = mad3(A00, W[3], A01, W[4], A02, W[5]);
        float Sc02 = mad3(A00, W[6], A01, W[7], A02, W[8]);
        float Sc11 = mad3(A10, W[3], A11, W[4], A12, W[5]);
        float Sc12 = mad3(A10, W[6], A11, W[7], A12, W[8]);
        float Sc22 = mad3(A20, W[6], A21, W[7], A22, W[8]);
        float tx = mul(pz, fmaxf(fminf(dvd(px, pz), cam.lim_xp), -cam.lim_xn));
        float ty = mul(pz, fmaxf(fminf(dvd(py, pz), cam.lim_yp), -cam.lim_yn));
        float z2 = mul(pz, pz);
        float J00 = dvd(cam.fx, pz);
        float J02 = dvd(-mul(cam.fx, tx), z2);
        float J11 = dvd(cam.fy, pz);
        float J12 = dvd(-mul(cam.fy, ty), z2);
        float B00 = mad2(J00, Sc00, J02, Sc02), B01 = mad2(J00, Sc01, J02, Sc12), B02 = mad2(J00, Sc02, J02, Sc22);
        float B11 = mad2(J11, Sc11, J12, Sc12), B12 = mad2(J11, Sc12, J12, Sc22);
        float c00 = mad2(B00, J00, B02, J02);
        float c01 = mad2(B01, J11, B02, J12);
        float c11 = mad2(B11, J11, B12, J12);
        float mx = add(dvd(mul(cam.fx, px), pz), cam.cx);
        float my = add(dvd(mul(cam.fy, py), pz), cam.cy);
        float det0 = msb2(c00, c11, c01, c01);
        float c00b = add(c00, p.eps2d);
        float c11b = add(c11, p.eps2d);
        float det = msb2(c00b, c11b, c01, c01);
        float bb = mul(0.5f, add(c00b, c11b));
        float v1 = add(bb, sqr(fmaxf(sub(mul(bb, bb), det), 0.01f)));
        float radius = ceilf(mul(3.0f, sqr(v1)));
        bool keep = (det > 0.0f) && (pz > p.near_plane) && (pz < p.far_plane) && (radius > p.radius_clip);
        keep = keep && (add(mx, radius) > 0.0f) && (sub(mx, radius) < (float)p.width) &&
               (add(my, radius) > 0.0f) && (sub(my, radius) < (float)p.height);
        float ca = 0, cb = 0, cc = 0, comp = 0, o = 0;
        int ri = 0, ntiles = 0, nexact = 0;
        if (keep) {
            ca = dvd(c11b, det);
            cb = dvd(-c01, det);
            cc = dvd(c00b, det);
            comp = sqr(fmaxf(dvd(det0, det), 0.0f));
            o = p.calc_comp ? mul(opac, comp) : opac;
            ri = (int)radius;
            TileBox tb = tile_box(mx, my, ri, p.tile_size, p.tile_w, p.tile_h);
            ntiles = (tb.x1 - tb.x0) * (tb.y1 - tb.y0);
            if (p.tiles_exact) {
                // the tiles this Gaussian can reach with alpha >= 1/255: one column interval per tile row of the bounding box
                // (common.cuh, exact_row_span); the intersection stage enumerates exactly these spans from the same floats
                const ExactCtx ec = exact_ctx(mx, my, o, ca, cb, cc);
                for (int ty_ = tb.y0; ty_ < tb.y1; ++ty_) {
                    int c0, c1;
                    exact_row_span(ec, ty_, p.height, tb.x0, tb.x1, c0, c1);
                    nexact += c1 - c0;
                }
            }
            vismask |= (1u << ci);
        } else {
            mx = 0;
            my = 0;
            pz = 0;
        }
        p.radii[idx] = ri;
        p.tiles[idx] = ntiles;
        if (p.tiles_exact) p.tiles_exact[idx] = nexact;
        reinterpret_cast<float2*>(p.means2d)[idx] = make_float2(mx, my);
        p.depths[idx] = pz;
        p.conics[idx * 3 + 0] = ca;
        p.conics[idx * 3 + 1] = cb;
        p.conics[idx * 3 + 2] = cc;
        if (p.comps) p.comps[idx] = comp;
        p.opac_out[idx] = o;
        float4* g = reinterpret_cast<float4*>(p.geom) + idx * 2;
        g[0] = make_float4(mx, my, o, pz);
        g[1] = make_float4(ca, cb, cc, 0.0f);
        if (p.append_depth) p.colors_out[idx * D + p.n_color] = pz;
    }

    if (p.n_color == 0) return;

    if (DEG < 0) {
        // colours pass through ([N,3] or [C,N,3]); culled entries get zeros
        if (!in_range) return;
        for (int ci = 0; ci < nC; ++ci) {
            const int64_t idx = (int64_t)(c0 + ci) * p.N + n;
            const float* src = p.colors_in + (p.colors_per_camera ? idx * 3 : (int64_t)n * 3);
            bool vis = (vismask >> ci) & 1u;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) p.colors_out[idx * D + ch] = vis ? src[ch] : 0.0f;
        }
        return;
    } else {
        using Sh = ShShape<(DEG < 0 ? 0 : DEG)>;
        const uint32_t warp_vis = __ballot_sync(0xffffffffu, vismask != 0);
        const int64_t row0 = (int64_t)blockIdx.x * kProjThreads + warp * 32;  // first Gaussian of this warp
        const int row_floats = p.K * 3;
        if (VEC) {
            float4* wbuf = reinterpret_cast<float4*>(shbuf) + warp * 32 * Sh::kStrideVec;
            const float4* src = reinterpret_cast<const float4*>(p.colors_in) + row0 * (row_floats / 4);
            const int row_vec = row_floats / 4;
            if (Sh::kVec == 12 && row_vec == 12) {
                // degree 3, K = 16: 4 lanes per row, 8 rows per step -- row / column are bit operations of the lane plus
                // compile-time constants (the generic loop below spends ~30 instructions per float4 on q / kVec and
                // 64-bit address arithmetic)
                const int r0 = lane >> 2, j0 = lane & 3;
                const float4* s0 = src + r0 * 12 + j0;
                float4* d0 = wbuf + r0 * Sh::kStrideVec + j0;
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    if ((warp_vis >> (r0 + 8 * rr)) & 1u) {
#pragma unroll
                        for (int jj = 0; jj < 3; ++jj) cp_async16(d0 + rr * 8 * Sh::kStrideVec + 4 * jj, s0 + rr * 8 * 12 + 4 * jj);
                    }
                }
            } else {
#pragma unroll 4
                for (int q = lane; q < 32 * Sh::kVec; q += 32) {
                    int r = q / Sh::kVec, j = q - r * Sh::kVec;
                    if ((warp_vis >> r) & 1u) cp_async16(wbuf + r * Sh::kStrideVec + j, src + (int64_t)r * row_vec + j);
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
        } else {
            float* wbuf = shbuf + warp * 32 * Sh::kStrideScalar;
            const float* src = p.colors_in + row0 * row_floats;
            for (int q = lane; q < 32 * Sh::kFloats; q += 32) {
                int r = q / Sh::kFloats, j = q - r * Sh::kFloats;
                if ((warp_vis >> r) & 1u) wbuf[r * Sh::kStrideScalar + j] = src[(int64_t)r * row_floats + j];
            }
            __syncwarp();
        }
        if (!in_range || vismask == 0) {
            // still have to zero the colour channels of culled entries
            if (in_range) {
                for (int ci = 0; ci < nC; ++ci) {
                    const int64_t idx = (int64_t)(c0 + ci) * p.N + n;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) p.colors_out[idx * D + ch] = 0.0f;
                }
            }
            return;
        }
        float coef[Sh::kVec * 4];
        if (VEC) {
            const float4* row = reinterpret_cast<const float4*>(shbuf) + (warp * 32 + lane) * Sh::kStrideVec;
#pragma unroll
            for (int j = 0; j < Sh::kVec; ++j) {
                float4 v = row[j];
                coef[4 * j + 0] = v.x;
                coef[4 * j + 1] = v.y;
                coef[4 * j + 2] = v.z;
                coef[4 * j + 3] = v.w;
            }
        } else {
            const float* row = shbuf + (warp * 32 + lane) * Sh::kStrideScalar;
#pragma unroll
            for (int j = 0; j < Sh::kFloats; ++j) coef[j] = row[j];
        }
        for (int ci = 0; ci < nC; ++ci) {
            const int64_t idx = (int64_t)(c0 + ci) * p.N + n;
            float r = 0, g = 0, b = 0;
            if ((vismask >> ci) & 1u) {
                const Cam& cam = cams[ci];
                float dx = sub(m0, cam.campos[0]), dy = sub(m1, cam.campos[1]), dz = sub(m2, cam.campos[2]);
                float nrm = fmaxf(sqr(add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz))), 1e-12f);
                dx = dvd(dx, nrm);
                dy = dvd(dy, nrm);
                dz = dvd(dz, nrm);
                float bs[16];
                sh_bases<(DEG < 0 ? 0 : DEG)>(dx, dy, dz, bs);
                r = mul(bs[0], coef[0]);
                g = mul(bs[0], coef[1]);
                b = mul(bs[0], coef[2]);
#pragma unroll
                for (int k = 1; k < Sh::kBases; ++k) {
                    r = add(r, mul(bs[k], coef[3 * k + 0]));
                    g = add(g, mul(bs[k], coef[3 * k + 1]));
                    b = add(b, mul(bs[k], coef[3 * k + 2]));
                }
                r = fmaxf(add(r, 0.5f), 0.0f);
                g = fmaxf(add(g, 0.5f), 0.0f);
                b = fmaxf(add(b, 0.5f), 0.0f);
            }
            if (D == 4) {
                // depth channel was written above; rewrite the whole float4 in one store
                float d = p.depths[idx];
                reinterpret_cast<float4*>(p.colors_out)[idx] = make_float4(r, g, b, d);
            } else {
                p.colors_out[idx * D + 0] = r;
                p.colors_out[idx * D + 1] = g;
                p.colors_out[idx * D + 2] = b;
            }
        }
    }
}

template <int DEG, bool VEC, int ONE = 0>
static int launch_project_fwd(const ProjFwdParams& p, cudaStream_t stream) {
    using Sh = ShShape<(DEG < 0 ? 0 : DEG)>;
    size_t smem = (size_t)p.Cc * kCamFloats * 4;
    if (DEG >= 0 && p.n_color > 0) smem += (size_t)kProjThreads * (VEC ? Sh::kStrideVec * 16 : Sh::kStrideScalar * 4);
    auto kern = project_fwd_kernel<DEG, VEC, ONE>;
    if (smem > 48 * 1024) QED_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = (p.N + kProjThreads - 1) / kProjThreads;
    QED_CUDA_TRY(launch_pdl(kern, dim3(blocks), dim3(kProjThreads), smem, stream, p));
    return QED_OK;
}

}  // namespace qed

using namespace qed;

// test hook (thread-local): 0 = generic kernel, 3 / 4 = single-view specialisation with that many resident blocks per SM
static thread_local int g_projfwd_one = 4;
extern "C" int qed_debug_set_project_fwd_one(int blocks) {
    int old = g_projfwd_one;
    g_projfwd_one = (blocks == 4) ? 4 : (blocks ? 3 : 0);
    return old;
}

extern "C" int qed_project_fwd(int C, int N, const float* means, const float* quats, const float* scales,
                               const float* opacities, int activations, const float* colors_in, int K, int sh_degree,
                               int colors_per_camera, const float* viewmats, const float* Ks, int width, int height,
                               float eps2d, float near_plane, float far_plane, float radius_clip,
                               int calc_compensations, int tile_size, int n_color, int append_depth,
                               int32_t* radii, float* means2d, float* depths, float* conics, float* compensations,
                               float* colors_out, float* opacities_out, int32_t* tiles_per_gauss, int32_t* tiles_exact, float* geom,
                               qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (C < 0 || N < 0 || width <= 0 || height <= 0 || tile_size <= 0) return QED_ERR_BAD_ARG;
    if (!(n_color == 0 || n_color == 3) || !(append_depth == 0 || append_depth == 1)) return QED_ERR_BAD_ARG;
    const int D = n_color + append_depth;
    if (!(D == 1 || D == 3 || D == 4)) return QED_ERR_BAD_ARG;
    if (C == 0 || N == 0) return QED_OK;
    if (!means || !quats || !scales || !viewmats || !Ks || !radii || !means2d || !depths || !conics ||
        !colors_out || !opacities_out || !tiles_per_gauss || !geom)
        return QED_ERR_BAD_ARG;
    if (calc_compensations && (!compensations || !opacities)) return QED_ERR_BAD_ARG;
    if (n_color > 0 && !colors_in) return QED_ERR_BAD_ARG;
    if (n_color == 0) sh_degree = -1;
    if (sh_degree > 3) return QED_ERR_UNSUPPORTED;
    if (sh_degree >= 0 && (sh_degree + 1) * (sh_degree + 1) > K) return QED_ERR_BAD_ARG;

    ProjFwdParams p;
    p.C = C;
    p.N = N;
    p.K = K;
    p.sh_degree = sh_degree;
    p.colors_per_camera = colors_per_camera;
    p.width = width;
    p.height = height;
    p.eps2d = eps2d;
    p.near_plane = near_plane;
    p.far_plane = far_plane;
    p.radius_clip = radius_clip;
    p.calc_comp = calc_compensations;
    p.activations = activations;
    p.tile_size = (float)tile_size;
    p.tile_w = (width + tile_size - 1) / tile_size;
    p.tile_h = (height + tile_size - 1) / tile_size;
    p.n_color = n_color;
    p.append_depth = append_depth;
    p.means = means;
    p.quats = quats;
    p.scales = scales;
    p.opacities = opacities;
    p.colors_in = colors_in;
    p.viewmats = viewmats;
    p.Ks = Ks;
    p.radii = radii;
    p.means2d = means2d;
    p.depths = depths;
    p.conics = conics;
    p.comps = calc_compensations ? compensations : nullptr;
    p.colors_out = colors_out;
    p.opac_out = opacities_out;
    p.tiles = tiles_per_gauss;
    p.tiles_exact = tiles_exact;
    if (tiles_exact && tile_size != 16) return QED_ERR_UNSUPPORTED;
    p.geom = geom;

    const bool vec_ok = sh_degree >= 0 && ((K * 3) % 4 == 0) && ((reinterpret_cast<uintptr_t>(colors_in) & 15) == 0);
    for (int c0 = 0; c0 < C; c0 += kMaxCamsPerLaunch) {
        p.c0 = c0;
        p.Cc = (C - c0 < kMaxCamsPerLaunch) ? (C - c0) : kMaxCamsPerLaunch;
        int rc;
        switch (sh_degree) {
            case 0: rc = vec_ok ? launch_project_fwd<0, true>(p, stream) : launch_project_fwd<0, false>(p, stream); break;
            case 1: rc = vec_ok ? launch_project_fwd<1, true>(p, stream) : launch_project_fwd<1, false>(p, stream); break;
            case 2: rc = vec_ok ? launch_project_fwd<2, true>(p, stream) : launch_project_fwd<2, false>(p, stream); break;
            case 3:
                if (vec_ok && p.Cc == 1 && p.c0 == 0 && g_projfwd_one == 4) rc = launch_project_fwd<3, true, 4>(p, stream);
                else if (vec_ok && p.Cc == 1 && p.c0 == 0 && g_projfwd_one) rc = launch_project_fwd<3, true, 3>(p, stream);
                else rc = vec_ok ? launch_project_fwd<3, true>(p, stream) : launch_project_fwd<3, false>(p, stream);
                break;
            default: rc = launch_project_fwd<-1, false>(p, stream); break;
        }
        if (rc != QED_OK) return rc;
    }
    return QED_OK;
}
