// SSIM (11x11 Gaussian window, sigma 1.5, "valid" convolution) forward + backward, fused.
//
// This is the `1 - SSIM` term of splatfacto's RGB loss, reached from qed_splatter/model.py:83-85
// (nerfstudio: pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3)); SURVEY.md §8f#1.
// Semantics == oracle/torch_impl.py::ssim.  X = ground truth, Y = prediction, per channel:
//   mu = w*X, w*Y ; e = w*X^2, w*Y^2, w*XY ; s1 = e1 - mu1^2, s2 = e2 - mu2^2, s12 = e12 - mu1 mu2
//   map = (2 mu1 mu2 + C1)(2 s12 + C2) / ((mu1^2 + mu2^2 + C1)(s1 + s2 + C2)),  C1 = 0.01^2, C2 = 0.03^2
// forward kernel : one CTA per 16x16 tile of the valid output, separable 11-tap filter in shared memory, writes
//                  the three partial-derivative maps d map/d(mu2, e2, e12) and accumulates sum(map).
// backward kernel: dL/dY(p) = sum_q w(q-p) [Dmu(q) + 2 Y(p) De2(q) + X(p) De12(q)] * scale  (zero outside the
//                  valid region), again separable.  HBM-bound: ~15 floats per pixel-channel moved in total.
#include "common.cuh"

namespace qed {

constexpr int kWin = 11;
constexpr int kHalo = kWin - 1;
constexpr int kSsimTile = 32;               // outputs per CTA: 32 x 32
constexpr int kSsimIn = kSsimTile + kHalo;  // 42
constexpr int kBlk = 4;                     // outputs per thread along the filtered axis (register sliding window)
constexpr int kSsimThreads = 256;

struct GaussWin {
    float w[kWin];
};

static GaussWin make_window() {
    GaussWin g;
    double s = 0.0, v[kWin];
    for (int i = 0; i < kWin; ++i) {
        const double c = i - kWin / 2;
        v[i] = exp(-(c * c) / (2.0 * 1.5 * 1.5));
        s += v[i];
    }
    for (int i = 0; i < kWin; ++i) g.w[i] = (float)(v[i] / s);
    return g;
}

// grid (tiles_x, tiles_y, C*3); 256 threads; every thread produces kBlk adjacent outputs per pass from a register
// sliding window, so each input is read from shared memory once per kBlk outputs instead of once per output
__global__ void __launch_bounds__(kSsimThreads) ssim_fwd_kernel(int W, int H, const float* __restrict__ pred /*[C,H,W,3]*/,
                                                                const GtImage gt, const PixelMask mask, GaussWin win,
                                                                float* __restrict__ dmaps /*[C*3][3][OH][OW]*/, double* __restrict__ stats) {
    extern __shared__ float ssim_smem[];
    float(*sx)[kSsimIn + 1] = reinterpret_cast<float(*)[kSsimIn + 1]>(ssim_smem);            // [42][43]
    float(*sy)[kSsimIn + 1] = sx + kSsimIn;                                                   // [42][43]
    float(*h)[kSsimIn][kSsimTile + 1] = reinterpret_cast<float(*)[kSsimIn][kSsimTile + 1]>(&sy[kSsimIn][0]);  // [5][42][33]
    __shared__ double red[kSsimThreads / 32];
    pdl_enter();
    const int OW = W - kHalo, OH = H - kHalo;
    const int cam = blockIdx.z / 3, ch = blockIdx.z % 3;
    const int ox0 = blockIdx.x * kSsimTile, oy0 = blockIdx.y * kSsimTile;
    const int64_t img = (int64_t)cam * H * W;
    for (int i = threadIdx.x; i < kSsimIn * kSsimIn; i += kSsimThreads) {
        const int r = i / kSsimIn, c = i - r * kSsimIn;
        const int y = oy0 + r, x = ox0 + c;
        float a = 0.f, b = 0.f;
        if (y < H && x < W) {
            const int64_t pix = img + (int64_t)y * W + x, o = pix * 3 + ch;
            a = gt.at(o) * mask.at(pix);  // pred already carries the mask (loss_grad_kernel<WRITE_PRED>)
            b = pred[o];
        }
        sx[r][c] = a;
        sy[r][c] = b;
    }
    __syncthreads();
    // horizontal pass: 42 rows x (32 / kBlk) column groups
    for (int i = threadIdx.x; i < kSsimIn * (kSsimTile / kBlk); i += kSsimThreads) {
        const int r = i / (kSsimTile / kBlk), c0 = (i - r * (kSsimTile / kBlk)) * kBlk;
        float acc[5][kBlk];
#pragma unroll
        for (int q = 0; q < 5; ++q)
#pragma unroll
            for (int o = 0; o < kBlk; ++o) acc[q][o] = 0.f;
#pragma unroll
        for (int t = 0; t < kWin + kBlk - 1; ++t) {
            const float a = sx[r][c0 + t], b = sy[r][c0 + t];
            const float aa = a * a, bb = b * b, ab = a * b;
#pragma unroll
            for (int o = 0; o < kBlk; ++o) {
                const int k = t - o;
                if (k >= 0 && k < kWin) {
                    const float w = win.w[k];
                    acc[0][o] += w * a;
                    acc[1][o] += w * b;
                    acc[2][o] += w * aa;
                    acc[3][o] += w * bb;
                    acc[4][o] += w * ab;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 5; ++q)
#pragma unroll
            for (int o = 0; o < kBlk; ++o) h[q][r][c0 + o] = acc[q][o];
    }
    __syncthreads();
    // vertical pass: 32 columns x (32 / kBlk) row groups = 256 thread items
    double local = 0.0;
    {
        const int tx = threadIdx.x % kSsimTile, y0 = (threadIdx.x / kSsimTile) * kBlk;
        float acc[5][kBlk];
#pragma unroll
        for (int q = 0; q < 5; ++q)
#pragma unroll
            for (int o = 0; o < kBlk; ++o) acc[q][o] = 0.f;
#pragma unroll
        for (int t = 0; t < kWin + kBlk - 1; ++t) {
            float v[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) v[q] = h[q][y0 + t][tx];
#pragma unroll
            for (int o = 0; o < kBlk; ++o) {
                const int k = t - o;
                if (k >= 0 && k < kWin) {
                    const float w = win.w[k];
#pragma unroll
                    for (int q = 0; q < 5; ++q) acc[q][o] += w * v[q];
                }
            }
        }
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const int64_t plane = (int64_t)OH * OW;
        const int ox = ox0 + tx;
#pragma unroll
        for (int o = 0; o < kBlk; ++o) {
            const int oy = oy0 + y0 + o;
            if (oy < OH && ox < OW) {
                const float m1 = acc[0][o], m2 = acc[1][o], e1 = acc[2][o], e2 = acc[3][o], e12 = acc[4][o];
                const float s1 = e1 - m1 * m1, s2 = e2 - m2 * m2, s12 = e12 - m1 * m2;
                const float A1 = 2.f * m1 * m2 + C1, A2 = 2.f * s12 + C2, B1 = m1 * m1 + m2 * m2 + C1, B2 = s1 + s2 + C2;
                const float inv = 1.0f / (B1 * B2);
                const float map = A1 * A2 * inv;
                local += (double)map;
                float* d = dmaps + ((int64_t)blockIdx.z * 3) * plane + (int64_t)oy * OW + ox;
                d[0] = (2.f * m1 * (A2 - A1) - 2.f * m2 * map * (B2 - B1)) * inv;  // d map / d mu2 (e2, e12 fixed)
                d[plane] = -map / B2;                                              // d map / d e2
                d[2 * plane] = 2.f * A1 * inv;                                     // d map / d e12
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int i = 0; i < kSsimThreads / 32; ++i) s += red[i];
        atomicAdd(stats + cam * 8 + 5, s);
    }
}

// grid (ceil(W/32), ceil(H/32), C*3): dL/dpred for a 32x32 tile of INPUT pixels
__global__ void __launch_bounds__(kSsimThreads) ssim_bwd_kernel(int W, int H, const float* __restrict__ pred, const GtImage gt, const PixelMask mask,
                                                                GaussWin win, const float* __restrict__ dmaps, float scale,
                                                                float* __restrict__ v_pred /*[C,H,W,3]*/) {
    extern __shared__ float ssim_smem[];
    float(*sd)[kSsimIn][kSsimIn + 1] = reinterpret_cast<float(*)[kSsimIn][kSsimIn + 1]>(ssim_smem);                      // [3][42][43]
    float(*h)[kSsimIn][kSsimTile + 1] = reinterpret_cast<float(*)[kSsimIn][kSsimTile + 1]>(&sd[3][0][0]);               // [3][42][33]
    pdl_enter();
    const int OW = W - kHalo, OH = H - kHalo;
    const int cam = blockIdx.z / 3, ch = blockIdx.z % 3;
    const int x0 = blockIdx.x * kSsimTile, y0 = blockIdx.y * kSsimTile;
    const int64_t plane = (int64_t)OH * OW;
    const float* d = dmaps + ((int64_t)blockIdx.z * 3) * plane;
    // pixel p gets contributions from q in [p-10, p] (valid coords): load the q tile starting at (y0-10, x0-10)
    for (int i = threadIdx.x; i < kSsimIn * kSsimIn; i += kSsimThreads) {
        const int r = i / kSsimIn, c = i - r * kSsimIn;
        const int qy = y0 - kHalo + r, qx = x0 - kHalo + c;
        float a = 0.f, b = 0.f, e = 0.f;
        if (qy >= 0 && qy < OH && qx >= 0 && qx < OW) {
            const int64_t o = (int64_t)qy * OW + qx;
            a = d[o];
            b = d[plane + o];
            e = d[2 * plane + o];
        }
        sd[0][r][c] = a;
        sd[1][r][c] = b;
        sd[2][r][c] = e;
    }
    __syncthreads();
    // pixel p = x0 + c receives q = p - 10 + k with weight w[10 - k]  (w is symmetric)
    for (int i = threadIdx.x; i < kSsimIn * (kSsimTile / kBlk); i += kSsimThreads) {
        const int r = i / (kSsimTile / kBlk), c0 = (i - r * (kSsimTile / kBlk)) * kBlk;
        float acc[3][kBlk];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int o = 0; o < kBlk; ++o) acc[q][o] = 0.f;
#pragma unroll
        for (int t = 0; t < kWin + kBlk - 1; ++t) {
            const float a = sd[0][r][c0 + t], b = sd[1][r][c0 + t], e = sd[2][r][c0 + t];
#pragma unroll
            for (int o = 0; o < kBlk; ++o) {
                const int k = t - o;
                if (k >= 0 && k < kWin) {
                    const float w = win.w[kWin - 1 - k];
                    acc[0][o] += w * a;
                    acc[1][o] += w * b;
                    acc[2][o] += w * e;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int o = 0; o < kBlk; ++o) h[q][r][c0 + o] = acc[q][o];
    }
    __syncthreads();
    {
        const int tx = threadIdx.x % kSsimTile, yb = (threadIdx.x / kSsimTile) * kBlk;
        float acc[3][kBlk];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int o = 0; o < kBlk; ++o) acc[q][o] = 0.f;
#pragma unroll
        for (int t = 0; t < kWin + kBlk - 1; ++t) {
            const float a = h[0][yb + t][tx], b = h[1][yb + t][tx], e = h[2][yb + t][tx];
#pragma unroll
            for (int o = 0; o < kBlk; ++o) {
                const int k = t - o;
                if (k >= 0 && k < kWin) {
                    const float w = win.w[kWin - 1 - k];
                    acc[0][o] += w * a;
                    acc[1][o] += w * b;
                    acc[2][o] += w * e;
                }
            }
        }
        const int x = x0 + tx;
#pragma unroll
        for (int o = 0; o < kBlk; ++o) {
            const int y = y0 + yb + o;
            if (y < H && x < W) {
                const int64_t pix = ((int64_t)cam * H + y) * W + x, idx = pix * 3 + ch;
                v_pred[idx] = scale * (acc[0][o] + 2.f * pred[idx] * acc[1][o] + gt.at(idx) * mask.at(pix) * acc[2][o]);
            }
        }
    }
}

constexpr size_t kSsimFwdSmem = (size_t)(2 * kSsimIn * (kSsimIn + 1) + 5 * kSsimIn * (kSsimTile + 1)) * 4;
constexpr size_t kSsimBwdSmem = (size_t)(3 * kSsimIn * (kSsimIn + 1) + 3 * kSsimIn * (kSsimTile + 1)) * 4;

}  // namespace qed

using namespace qed;

// pred (already multiplied by the mask) / gt [C,H,W,3], mask [C,H,W] or NULL; dmaps scratch [C*3*3*(H-10)*(W-10)]; stats[c*8+5] += sum of the SSIM map of camera c;
// v_pred = scale * d(sum map)/d pred.  Internal to qed_loss_fwd_bwd (train.cu), declared there.
int qed_ssim_launch(int C, int W, int H, const float* pred, qed::GtImage gt, qed::PixelMask mask, float* dmaps, double* stats, float scale,
                    float* v_pred, cudaStream_t stream) {
    if (W <= kHalo || H <= kHalo) return QED_ERR_UNSUPPORTED;
    static const GaussWin win = make_window();
    const int OW = W - kHalo, OH = H - kHalo;
    QED_CUDA_TRY(cudaFuncSetAttribute(ssim_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSsimFwdSmem));
    QED_CUDA_TRY(cudaFuncSetAttribute(ssim_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSsimBwdSmem));
    dim3 g1((OW + kSsimTile - 1) / kSsimTile, (OH + kSsimTile - 1) / kSsimTile, C * 3);
    QED_CUDA_TRY(launch_pdl(ssim_fwd_kernel, g1, dim3(kSsimThreads), kSsimFwdSmem, stream, W, H, pred, gt, mask, win, dmaps, stats));
    dim3 g2((W + kSsimTile - 1) / kSsimTile, (H + kSsimTile - 1) / kSsimTile, C * 3);
    QED_CUDA_TRY(launch_pdl(ssim_bwd_kernel, g2, dim3(kSsimThreads), kSsimBwdSmem, stream, W, H, pred, gt, mask, win, dmaps, scale, v_pred));
    return QED_OK;
}
