// (f4) data side of the initial point cloud: depth back-projection and voxel merge of `qed-init-pc`
// (qed_splatter/create_init_pointcloud.py:148-261), which the reference runs on the host through Open3D
// (o3d.t.geometry.PointCloud.create_from_depth_image, :176-185, and .voxel_down_sample, :194 / :89 / :260).
//
// Open3D is not in this image, so the semantics follow its published behaviour as restated in oracle/pointcloud.py
// (PARITY UNPINNED, see there): a pixel (u, v) of the strided grid is kept iff 0 < d < depth_max, d = depth * unit scale
// (the reference zeroes non-finite / non-positive depths first, :165-168); camera point ((u - cx) d / fx, (v - cy) d / fy, d);
// world point = inverse(extrinsic) * camera point.  voxel_down_sample: points are binned by floor(p / voxel_size) and every
// occupied voxel yields the MEAN of its points.
//
// B200 formulation (HBM-bound, no atomics, deterministic):
//   back-projection = one flag scan over the strided pixel grid whose final phase writes the surviving points compacted in
//     pixel order (Open3D's order is an atomic counter's, i.e. unspecified);
//   voxel merge = 63-bit voxel key (3 x 21 bits) per point -> this library's stable radix sort of (key, index) pairs ->
//     run heads flagged + scanned -> one thread per occupied voxel averages its run (float64 accumulation, index order), so
//     the output is sorted by voxel key and bit-reproducible.
#include "common.cuh"
#include "scan.cuh"

namespace qed {

static inline size_t pc_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct DepthImage {
    const void* p;
    int u16;       // raw uint16 sensor image (else float32)
    float scale;   // depth_unit_scale_factor (qed_splatter/dataparser.py:15, create_init_pointcloud.py:201)
    // create_init_pointcloud.py:165 multiplies the float32 image (`_load_depth` converts to float32, :30-41) by the python
    // scalar: a float32 product
    __device__ __forceinline__ float at(int64_t pix) const {
        const float raw = u16 ? (float)reinterpret_cast<const uint16_t*>(p)[pix] : reinterpret_cast<const float*>(p)[pix];
        return mul(raw, scale);
    }
};

struct BackprojectCtx {
    DepthImage depth;
    int width, height, stride, gw;  // gw = strided grid width
    float depth_max;
    float fx, fy, cx, cy;
    float P[12];  // inverse(extrinsic): rows of [R | t], camera -> world
};

// flag of strided pixel i (row-major over the strided grid)
struct BackprojectFlag {
    BackprojectCtx c;
    __device__ __forceinline__ int32_t operator()(int64_t i) const {
        const int gy = (int)(i / c.gw), gx = (int)(i - (int64_t)gy * c.gw);
        const float d = c.depth.at((int64_t)gy * c.stride * c.width + (int64_t)gx * c.stride);
        return (d > 0.0f && d < c.depth_max) ? 1 : 0;  // NaN fails both compares, +inf the second
    }
};

struct BackprojectSink {
    BackprojectCtx c;
    float* points;
    __device__ __forceinline__ void operator()(int64_t i, int32_t flag, int64_t inc) const {
        if (!flag) return;
        const int gy = (int)(i / c.gw), gx = (int)(i - (int64_t)gy * c.gw);
        const int u = gx * c.stride, v = gy * c.stride;
        const float d = c.depth.at((int64_t)v * c.width + u);
        const float xc = dvd(mul(sub((float)u, c.cx), d), c.fx);
        const float yc = dvd(mul(sub((float)v, c.cy), d), c.fy);
        float* o = points + (inc - 1) * 3;
#pragma unroll
        for (int r = 0; r < 3; ++r) o[r] = add(add(add(mul(c.P[r * 4 + 0], xc), mul(c.P[r * 4 + 1], yc)), mul(c.P[r * 4 + 2], d)), c.P[r * 4 + 3]);
    }
};

// ---- voxel merge ----
constexpr int kVoxelBits = 21;
constexpr int kVoxelBias = 1 << (kVoxelBits - 1);

constexpr int64_t kVoxelPadKey = 0x7fffffffffffffffLL;  // all 63 sorted bits set: real keys stop one below in every axis

__global__ void voxel_keys_kernel(int64_t n, const int64_t* n_dev, const float* __restrict__ points, float voxel, int64_t* __restrict__ keys,
                                  int32_t* __restrict__ vals) {
    pdl_enter();
    const int64_t m = n_dev ? min(*n_dev, n) : n;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i >= m) {  // capacity padding behind a device-side count: sorts to the end, never a run head that is counted
        keys[i] = kVoxelPadKey;
        vals[i] = (int32_t)i;
        return;
    }
    int64_t key = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // floor(p / voxel) as Open3D computes it: a float32 division, then floor
        const float q = floorf(dvd(points[i * 3 + a], voxel));
        const float qc = fminf(fmaxf(q, (float)-kVoxelBias), (float)(kVoxelBias - 2));
        key = (key << kVoxelBits) | (int64_t)((int)qc + kVoxelBias);
    }
    keys[i] = key;
    vals[i] = (int32_t)i;
}

// 1 where sorted element i starts a run of equal keys (and is a real point)
struct VoxelHeadFlag {
    const int64_t* keys;
    __device__ __forceinline__ int32_t operator()(int64_t i) const {
        const int64_t k = keys[i];
        if (k == kVoxelPadKey) return 0;
        return (i == 0 || keys[i - 1] != k) ? 1 : 0;
    }
};

// final phase of the head scan: the head of voxel (inc - 1) averages its run
struct VoxelMeanSink {
    const int64_t* keys;
    const int32_t* order;
    const float* points;
    int64_t n;
    float* out;
    __device__ __forceinline__ void operator()(int64_t i, int32_t flag, int64_t inc) const {
        if (!flag) return;
        const int64_t k = keys[i];
        double sx = 0.0, sy = 0.0, sz = 0.0;
        int64_t j = i;
        for (; j < n && keys[j] == k; ++j) {
            const float* p = points + (int64_t)order[j] * 3;
            sx += (double)p[0];
            sy += (double)p[1];
            sz += (double)p[2];
        }
        const double cnt = (double)(j - i);
        float* o = out + (inc - 1) * 3;
        o[0] = (float)(sx / cnt);
        o[1] = (float)(sy / cnt);
        o[2] = (float)(sz / cnt);
    }
};

struct VoxelLayout {
    size_t keys_in, keys_out, vals_in, vals_out, sort_ws, scan_ws, total;
};
static VoxelLayout voxel_layout(int64_t n) {
    VoxelLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += pc_align_up(bytes, 256);
        return r;
    };
    L.keys_in = take((size_t)n * 8);
    L.keys_out = take((size_t)n * 8);
    L.vals_in = take((size_t)n * 4);
    L.vals_out = take((size_t)n * 4);
    L.sort_ws = take(qed_sort_pairs_workspace_bytes(n));
    L.scan_ws = take(scan_workspace_bytes(n));
    L.total = o;
    return L;
}

// general 3x4 inverse of [R | t] (R need not be orthonormal: the reference inverts numerically, create_init_pointcloud.py:70)
static bool invert_rigid(const float* E, float* P) {
    const double a = E[0], b = E[1], c = E[2], d = E[4], e = E[5], f = E[6], g = E[8], h = E[9], i = E[10];
    const double c00 = e * i - f * h, c01 = c * h - b * i, c02 = b * f - c * e;
    const double c10 = f * g - d * i, c11 = a * i - c * g, c12 = c * d - a * f;
    const double c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
    const double det = a * c00 + b * c10 + c * c20;
    if (!(det != 0.0)) return false;
    const double Ri[9] = {c00 / det, c01 / det, c02 / det, c10 / det, c11 / det, c12 / det, c20 / det, c21 / det, c22 / det};
    const double t[3] = {E[3], E[7], E[11]};
    for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 3; ++k) P[r * 4 + k] = (float)Ri[r * 3 + k];
        P[r * 4 + 3] = (float)(-(Ri[r * 3 + 0] * t[0] + Ri[r * 3 + 1] * t[1] + Ri[r * 3 + 2] * t[2]));
    }
    return true;
}

}  // namespace qed

using namespace qed;

extern "C" size_t qed_backproject_workspace_bytes(int width, int height, int stride) {
    if (width <= 0 || height <= 0 || stride <= 0) return 0;
    const int64_t g = (int64_t)((width + stride - 1) / stride) * ((height + stride - 1) / stride);
    return scan_workspace_bytes(g);
}

extern "C" int qed_backproject_depth(int width, int height, const void* depth, int depth_is_u16, double depth_unit_scale, float depth_max,
                                     int stride, const float* intrinsic_host, const float* extrinsic_host, float* points,
                                     int64_t* n_points_dev, void* workspace, size_t workspace_bytes, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (width <= 0 || height <= 0 || stride <= 0 || !(depth_max > 0.0f)) return QED_ERR_BAD_ARG;
    if (!depth || !intrinsic_host || !extrinsic_host || !points || !n_points_dev || !workspace) return QED_ERR_BAD_ARG;
    if (workspace_bytes < qed_backproject_workspace_bytes(width, height, stride)) return QED_ERR_WORKSPACE;
    BackprojectCtx c;
    c.depth = DepthImage{depth, depth_is_u16 ? 1 : 0, (float)depth_unit_scale};
    c.width = width;
    c.height = height;
    c.stride = stride;
    c.gw = (width + stride - 1) / stride;
    c.depth_max = depth_max;
    c.fx = intrinsic_host[0];
    c.fy = intrinsic_host[4];
    c.cx = intrinsic_host[2];
    c.cy = intrinsic_host[5];
    if (!(c.fx != 0.0f) || !(c.fy != 0.0f)) return QED_ERR_BAD_ARG;
    if (!invert_rigid(extrinsic_host, c.P)) return QED_ERR_BAD_ARG;
    const int64_t g = (int64_t)c.gw * ((height + stride - 1) / stride);
    return scan_inclusive_to(g, nullptr, BackprojectFlag{c}, BackprojectSink{c, points}, n_points_dev, workspace, stream);
}

extern "C" size_t qed_voxel_downsample_workspace_bytes(int64_t n) { return n > 0 ? voxel_layout(n).total : 0; }

extern "C" int qed_voxel_downsample(int64_t n, const int64_t* n_dev, const float* points, float voxel_size, float* points_out,
                                    int64_t* n_out_dev, void* workspace, size_t workspace_bytes, qed_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || !(voxel_size > 0.0f)) return QED_ERR_BAD_ARG;
    if (!n_out_dev) return QED_ERR_BAD_ARG;
    if (n == 0) {
        QED_CUDA_TRY(cudaMemsetAsync(n_out_dev, 0, sizeof(int64_t), stream));
        return QED_OK;
    }
    if (n > 0x7fffffffLL) return QED_ERR_UNSUPPORTED;
    if (!points || !points_out || !workspace) return QED_ERR_BAD_ARG;
    if (workspace_bytes < qed_voxel_downsample_workspace_bytes(n)) return QED_ERR_WORKSPACE;
    const VoxelLayout L = voxel_layout(n);
    char* ws = reinterpret_cast<char*>(workspace);
    int64_t* keys_in = reinterpret_cast<int64_t*>(ws + L.keys_in);
    int64_t* keys_out = reinterpret_cast<int64_t*>(ws + L.keys_out);
    int32_t* vals_in = reinterpret_cast<int32_t*>(ws + L.vals_in);
    int32_t* vals_out = reinterpret_cast<int32_t*>(ws + L.vals_out);
    QED_CUDA_TRY(launch_pdl(voxel_keys_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, n, n_dev, points, voxel_size, keys_in, vals_in));
    int rc = qed_sort_pairs(n, keys_in, vals_in, keys_out, vals_out, 3 * kVoxelBits, ws + L.sort_ws, L.total - L.sort_ws, stream_);
    if (rc != QED_OK) return rc;
    return scan_inclusive_to(n, nullptr, VoxelHeadFlag{keys_out}, VoxelMeanSink{keys_out, vals_out, points, n, points_out}, n_out_dev,
                             ws + L.scan_ws, stream);
}
