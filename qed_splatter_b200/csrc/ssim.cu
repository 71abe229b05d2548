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
constexpr int kSsimTile = 16;
constexpr int kSsimIn = kSsimTile + kHalo;  // 26

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

// grid (tiles_x, tiles_y, C*3); block 256 threads (16x16)
__global__ void __launch_bounds__(256) ssim_fwd_kernel(int W, int H, const float* __restrict__ pred /*[C,H,W,3]*/, const float* __restrict__ gt,
                                                       GaussWin win, float* __restrict__ dmaps /*[C*3][3][OH][OW]*/, double* __restrict__ stats) {
    __shared__ float sx[kSsimIn][kSsimIn + 1], sy[kSsimIn][kSsimIn + 1];
    __shared__ float h[5][kSsimIn][kSsimTile + 1];
    __shared__ double red[8];
    const int OW = W - kHalo, OH = H - kHalo;
    const int cam = blockIdx.z / 3, ch = blockIdx.z % 3;
    const int ox0 = blockIdx.x * kSsimTile, oy0 = blockIdx.y * kSsimTile;
    const int64_t img = (int64_t)cam * H * W;
    for (int i = threadIdx.x; i < kSsimIn * kSsimIn; i += 256) {
        const int r = i / kSsimIn, c = i - r * kSsimIn;
        const int y = oy0 + r, x = ox0 + c;
        float a = 0.f, b = 0.f;
        if (y < H && x < W) {
            const int64_t o = (img + (int64_t)y * W + x) * 3 + ch;
            a = gt[o];
            b = pred[o];
        }
        sx[r][c] = a;
        sy[r][c] = b;
    }
    __syncthreads();
    // horizontal pass: 26 rows x 16 columns
    for (int i = threadIdx.x; i < kSsimIn * kSsimTile; i += 256) {
        const int r = i / kSsimTile, c = i - r * kSsimTile;
        float m1 = 0, m2 = 0, e1 = 0, e2 = 0, e12 = 0;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float a = sx[r][c + k], b = sy[r][c + k], w = win.w[k];
            m1 += w * a;
            m2 += w * b;
            e1 += w * a * a;
            e2 += w * b * b;
            e12 += w * a * b;
        }
        h[0][r][c] = m1;
        h[1][r][c] = m2;
        h[2][r][c] = e1;
        h[3][r][c] = e2;
        h[4][r][c] = e12;
    }
    __syncthreads();
    const int ty = threadIdx.x / kSsimTile, tx = threadIdx.x % kSsimTile;
    const int oy = oy0 + ty, ox = ox0 + tx;
    double local = 0.0;
    if (oy < OH && ox < OW) {
        float m1 = 0, m2 = 0, e1 = 0, e2 = 0, e12 = 0;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float w = win.w[k];
            m1 += w * h[0][ty + k][tx];
            m2 += w * h[1][ty + k][tx];
            e1 += w * h[2][ty + k][tx];
            e2 += w * h[3][ty + k][tx];
            e12 += w * h[4][ty + k][tx];
        }
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float s1 = e1 - m1 * m1, s2 = e2 - m2 * m2, s12 = e12 - m1 * m2;
        const float A1 = 2.f * m1 * m2 + C1, A2 = 2.f * s12 + C2, B1 = m1 * m1 + m2 * m2 + C1, B2 = s1 + s2 + C2;
        const float inv = 1.0f / (B1 * B2);
        const float map = A1 * A2 * inv;
        local = (double)map;
        const int64_t plane = (int64_t)OH * OW;
        float* d = dmaps + ((int64_t)blockIdx.z * 3) * plane + (int64_t)oy * OW + ox;
        d[0] = (2.f * m1 * (A2 - A1) - 2.f * m2 * map * (B2 - B1)) * inv;  // d map / d mu2 (e2, e12 fixed)
        d[plane] = -map / B2;                                              // d map / d e2
        d[2 * plane] = 2.f * A1 * inv;                                     // d map / d e12
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(stats + cam * 8 + 5, s);
    }
}

// grid (ceil(W/16), ceil(H/16), C*3): dL/dpred for a 16x16 tile of INPUT pixels
__global__ void __launch_bounds__(256) ssim_bwd_kernel(int W, int H, const float* __restrict__ pred, const float* __restrict__ gt, GaussWin win,
                                                       const float* __restrict__ dmaps, float scale, float* __restrict__ v_pred /*[C,H,W,3]*/) {
    __shared__ float sd[3][kSsimIn][kSsimIn + 1];
    __shared__ float h[3][kSsimIn][kSsimTile + 1];
    const int OW = W - kHalo, OH = H - kHalo;
    const int cam = blockIdx.z / 3, ch = blockIdx.z % 3;
    const int x0 = blockIdx.x * kSsimTile, y0 = blockIdx.y * kSsimTile;
    const int64_t plane = (int64_t)OH * OW;
    const float* d = dmaps + ((int64_t)blockIdx.z * 3) * plane;
    // output pixel p gets contributions from q in [p-10, p] (valid coords): load q tile starting at (y0-10, x0-10)
    for (int i = threadIdx.x; i < kSsimIn * kSsimIn; i += 256) {
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
    // pixel p = x0 + c receives q = p - 10 + k with weight w[10 - k]  (w symmetric)
    for (int i = threadIdx.x; i < kSsimIn * kSsimTile; i += 256) {
        const int r = i / kSsimTile, c = i - r * kSsimTile;
        float a = 0, b = 0, e = 0;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float w = win.w[kWin - 1 - k];
            a += w * sd[0][r][c + k];
            b += w * sd[1][r][c + k];
            e += w * sd[2][r][c + k];
        }
        h[0][r][c] = a;
        h[1][r][c] = b;
        h[2][r][c] = e;
    }
    __syncthreads();
    const int ty = threadIdx.x / kSsimTile, tx = threadIdx.x % kSsimTile;
    const int y = y0 + ty, x = x0 + tx;
    if (y < H && x < W) {
        float a = 0, b = 0, e = 0;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float w = win.w[kWin - 1 - k];
            a += w * h[0][ty + k][tx];
            b += w * h[1][ty + k][tx];
            e += w * h[2][ty + k][tx];
        }
        const int64_t o = (((int64_t)cam * H + y) * W + x) * 3 + ch;
        v_pred[o] = scale * (a + 2.f * pred[o] * b + gt[o] * e);
    }
}

}  // namespace qed

using namespace qed;

// pred/gt [C,H,W,3]; dmaps scratch [C*3*3*(H-10)*(W-10)]; stats[c*8+5] += sum of the SSIM map of camera c;
// v_pred = scale * d(sum map)/d pred.  Internal to qed_loss_fwd_bwd (train.cu), declared there.
int qed_ssim_launch(int C, int W, int H, const float* pred, const float* gt, float* dmaps, double* stats, float scale, float* v_pred,
                    cudaStream_t stream) {
    if (W <= kHalo || H <= kHalo) return QED_ERR_UNSUPPORTED;
    static const GaussWin win = make_window();
    const int OW = W - kHalo, OH = H - kHalo;
    dim3 g1((OW + kSsimTile - 1) / kSsimTile, (OH + kSsimTile - 1) / kSsimTile, C * 3);
    ssim_fwd_kernel<<<g1, 256, 0, stream>>>(W, H, pred, gt, win, dmaps, stats);
    QED_LAUNCH_CHECK();
    dim3 g2((W + kSsimTile - 1) / kSsimTile, (H + kSsimTile - 1) / kSsimTile, C * 3);
    ssim_bwd_kernel<<<g2, 256, 0, stream>>>(W, H, pred, gt, win, dmaps, scale, v_pred);
    QED_LAUNCH_CHECK();
    return QED_OK;
}
