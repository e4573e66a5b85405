// b3d_deproject.cu -- K1: depth / disparity rasters -> point clouds.
//   z16   : rs.pointcloud().calculate() semantics, float32, every pixel            (pointcloud_capture.py:35,38)
//   rgbd  : Open3D create_from_rgbd_image semantics, float64, valid pixels only     (test/check84.py:155-178)
//   disp  : cv2.reprojectImageTo3D(disp/16, Q) semantics, float32                   (Q of Calib_depth/depth4.py:98)
// All three are pure streaming kernels: 14 B/pixel (2 in, 12 out) for z16 / disp; the roofline is HBM bandwidth.
#include "b3d_common.cuh"
#include "b3d_scan.cuh"

namespace b3d {
namespace {

// 4 pixels per thread: one 8-byte depth load, three 16-byte stores (requires w % 4 == 0 so a group never straddles rows)
__global__ void __launch_bounds__(256) deproject_z16_vec4_kernel(const uint16_t* __restrict__ depth, const uint8_t* __restrict__ bgr, int w, int h,
                                                                 int64_t n_groups, float fx, float fy, float ppx, float ppy, float scale,
                                                                 float* __restrict__ xyz, float* __restrict__ rgb) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = g * 4;
        const int64_t row = i0 / w;
        const int col = (int)(i0 - row * w);
        const int y = (int)(row % h);
        const ushort4 d = __ldg(reinterpret_cast<const ushort4*>(depth) + g);
        const float yy = ((float)y - ppy) / fy;
        float v[12];
        const unsigned short dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float z = scale * (float)dd[k];
            const float xx = ((float)(col + k) - ppx) / fx;
            v[3 * k] = z * xx;
            v[3 * k + 1] = z * yy;
            v[3 * k + 2] = z;
        }
        float4* o = reinterpret_cast<float4*>(xyz) + 3 * g;
        __stcs(o, make_float4(v[0], v[1], v[2], v[3]));
        __stcs(o + 1, make_float4(v[4], v[5], v[6], v[7]));
        __stcs(o + 2, make_float4(v[8], v[9], v[10], v[11]));
        if (bgr != nullptr) {
            const uint32_t* c = reinterpret_cast<const uint32_t*>(bgr) + 3 * g;
            const uint32_t c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2);
            float cv[12];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                cv[k] = (float)((double)((c0 >> (8 * k)) & 0xffu) / 255.0);
                cv[4 + k] = (float)((double)((c1 >> (8 * k)) & 0xffu) / 255.0);
                cv[8 + k] = (float)((double)((c2 >> (8 * k)) & 0xffu) / 255.0);
            }
            float4* oc = reinterpret_cast<float4*>(rgb) + 3 * g;
            __stcs(oc, make_float4(cv[0], cv[1], cv[2], cv[3]));
            __stcs(oc + 1, make_float4(cv[4], cv[5], cv[6], cv[7]));
            __stcs(oc + 2, make_float4(cv[8], cv[9], cv[10], cv[11]));
        }
    }
}

// scalar fallback for widths that are not a multiple of 4
__global__ void __launch_bounds__(256) deproject_z16_scalar_kernel(const uint16_t* __restrict__ depth, const uint8_t* __restrict__ bgr, int w, int h,
                                                                   int64_t n, float fx, float fy, float ppx, float ppy, float scale,
                                                                   float* __restrict__ xyz, float* __restrict__ rgb) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / w;
        const int col = (int)(i - row * w);
        const int y = (int)(row % h);
        const float z = scale * (float)depth[i];
        const float xx = ((float)col - ppx) / fx;
        const float yy = ((float)y - ppy) / fy;
        xyz[3 * i] = z * xx;
        xyz[3 * i + 1] = z * yy;
        xyz[3 * i + 2] = z;
        if (bgr != nullptr) {
#pragma unroll
            for (int k = 0; k < 3; ++k) rgb[3 * i + k] = (float)((double)bgr[3 * i + k] / 255.0);
        }
    }
}

struct RgbdPred {
    const uint16_t* depth;
    float scale, trunc;
    __device__ __forceinline__ bool operator()(int64_t i) const {
        float p = (float)depth[i];
        p /= scale;
        if (p >= trunc) p = 0.0f;
        return p > 0.0f;
    }
};
struct RgbdEmit {
    const uint16_t* depth;
    const uint8_t* color;
    int w;
    double fx, fy, cx, cy;
    float scale;
    int flip;
    double* xyz;
    double* rgb;
    __device__ __forceinline__ void operator()(int64_t i, int64_t slot) const {
        float p = (float)depth[i];
        p /= scale;
        const int row = (int)(i / w), col = (int)(i - (int64_t)row * w);
        double z = (double)p;
        double x = ((double)col - cx) * z / fx;
        double y = ((double)row - cy) * z / fy;
        if (flip) { y = -y; z = -z; }
        xyz[3 * slot] = x;
        xyz[3 * slot + 1] = y;
        xyz[3 * slot + 2] = z;
        if (color != nullptr && rgb != nullptr) {
            rgb[3 * slot] = (double)color[3 * i] / 255.0;
            rgb[3 * slot + 1] = (double)color[3 * i + 1] / 255.0;
            rgb[3 * slot + 2] = (double)color[3 * i + 2] / 255.0;
        }
    }
};

struct QMat {
    double q[16];
};

__global__ void __launch_bounds__(256) reproject_disparity_kernel(const int16_t* __restrict__ disp, int w, int64_t n, QMat Q, float* __restrict__ xyz) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / w;
        const double x = (double)(int)(i - row * w), y = (double)row;
        const double d = (double)((float)disp[i] / 16.0f);
        double v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = Q.q[4 * r] * x + Q.q[4 * r + 1] * y + Q.q[4 * r + 2] * d + Q.q[4 * r + 3] * 1.0;
        const double iw = 1.0 / v[3];
        // cv2: XYZ rounded to float first, then scaled by 1/W in double and rounded again
        __stcs(xyz + 3 * i, (float)((double)(float)v[0] * iw));
        __stcs(xyz + 3 * i + 1, (float)((double)(float)v[1] * iw));
        __stcs(xyz + 3 * i + 2, (float)((double)(float)v[2] * iw));
    }
}

// valid-only variant: pixels with disparity >= min16 (fixed point x16), raster order, frames stacked back to back
struct DispPred {
    const int16_t* disp;
    int min16;
    __device__ __forceinline__ bool operator()(int64_t i) const { return (int)disp[i] >= min16; }
};
struct DispEmit {
    const int16_t* disp;
    int w;
    int64_t frame_px;
    QMat Q;
    float* xyz;
    int32_t* src_px;  // optional: batch-global pixel index of every emitted point
    __device__ __forceinline__ void operator()(int64_t i, int64_t slot) const {
        // 32-bit divisions: a batch holds fewer than 2^31 pixels (checked by the callers), and a 64-bit division costs ~4x as much
        const uint32_t px = (uint32_t)i % (uint32_t)frame_px;
        const uint32_t row = px / (uint32_t)w;
        const double x = (double)(int)(px - row * (uint32_t)w), y = (double)(int)row;
        const double d = (double)((float)disp[i] / 16.0f);
        double v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = Q.q[4 * r] * x + Q.q[4 * r + 1] * y + Q.q[4 * r + 2] * d + Q.q[4 * r + 3] * 1.0;
        const double iw = 1.0 / v[3];
        xyz[3 * slot] = (float)((double)(float)v[0] * iw);
        xyz[3 * slot + 1] = (float)((double)(float)v[1] * iw);
        xyz[3 * slot + 2] = (float)((double)(float)v[2] * iw);
        if (src_px != nullptr) src_px[slot] = (int32_t)i;
    }
};

// off[f] = number of emitted points whose pixel index is below f * frame_px (f = 0..frames)
__global__ void frame_offsets_kernel(const int32_t* __restrict__ src_px, const int64_t* __restrict__ total_d, int64_t frame_px, int frames,
                                     int32_t* __restrict__ off) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f > frames) return;
    const int64_t key = (int64_t)f * frame_px;
    int64_t lo = 0, hi = *total_d;
    while (lo < hi) {
        const int64_t m = (lo + hi) >> 1;
        if ((int64_t)src_px[m] < key) lo = m + 1; else hi = m;
    }
    off[f] = (int32_t)lo;
}

}  // namespace

// Valid pixels of `frames` stacked disparity rasters -> xyz (capacity frames*h*w rows); off_h receives the per-frame offsets.
int reproject_disparity_valid_batch(b3d_ctx* ctx, const int16_t* disp, int w, int h, int frames, const double* Q_h, int min_disp16, float* xyz,
                                    std::vector<int32_t>* off_h) {
    const int64_t frame_px = (int64_t)w * h, n = frame_px * frames;
    off_h->assign((size_t)frames + 1, 0);
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(n < (int64_t)INT32_MAX, "%d frames x %lld pixels exceed 2^31-1 points", frames, (long long)frame_px);  // pixel indices are 32-bit below
    QMat Q;
    for (int i = 0; i < 16; ++i) Q.q[i] = Q_h[i];
    DevBuf<int64_t> total;
    DevBuf<int32_t> src_px, off;
    B3D_TRY(total.alloc(ctx, 1));
    B3D_TRY(src_px.alloc(ctx, (size_t)n));
    B3D_TRY(off.alloc(ctx, (size_t)frames + 1));
    B3D_TRY(compact(ctx, DispPred{disp, min_disp16}, DispEmit{disp, w, frame_px, Q, xyz, src_px.p}, n, total.p));
    B3D_LAUNCH(ctx, frame_offsets_kernel, (frames + 1 + 127) / 128, 128, 0, src_px.p, total.p, frame_px, frames, off.p);
    return ctx->download(off_h->data(), off.p, ((size_t)frames + 1) * sizeof(int32_t));
}

// frames: number of h x w rasters stacked back to back (batch); xyz [frames*h*w, 3]
int deproject_z16_batch(b3d_ctx* ctx, const uint16_t* depth, const uint8_t* bgr, int w, int h, int frames, float fx, float fy, float ppx,
                        float ppy, float scale, float* xyz, float* rgb) {
    const int64_t n = (int64_t)w * h * frames;
    if (n == 0) return B3D_OK;
    const bool vec = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(depth) & 7) == 0) && ((reinterpret_cast<uintptr_t>(xyz) & 15) == 0) &&
                     (bgr == nullptr || (((reinterpret_cast<uintptr_t>(bgr) & 3) == 0) && ((reinterpret_cast<uintptr_t>(rgb) & 15) == 0)));
    if (vec) {
        const int64_t g = n / 4;
        B3D_LAUNCH(ctx, deproject_z16_vec4_kernel, ctx->grid_for(g, 256, 1, 8), 256, 0, depth, bgr, w, h, g, fx, fy, ppx, ppy, scale, xyz, rgb);
    } else {
        B3D_LAUNCH(ctx, deproject_z16_scalar_kernel, ctx->grid_for(n, 256, 1, 8), 256, 0, depth, bgr, w, h, n, fx, fy, ppx, ppy, scale, xyz, rgb);
    }
    return B3D_OK;
}

}  // namespace b3d

using namespace b3d;

extern "C" {

int b3d_deproject_z16(b3d_ctx* ctx, const uint16_t* depth, int w, int h, float fx, float fy, float ppx, float ppy, float depth_scale, float* xyz) {
    return b3d_deproject_z16_color(ctx, depth, nullptr, w, h, fx, fy, ppx, ppy, depth_scale, xyz, nullptr);
}

int b3d_deproject_z16_color(b3d_ctx* ctx, const uint16_t* depth, const uint8_t* bgr, int w, int h, float fx, float fy, float ppx, float ppy,
                            float depth_scale, float* xyz, float* rgb) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(w >= 0 && h >= 0, "b3d_deproject_z16: negative image size");
    if ((int64_t)w * h == 0) return B3D_OK;
    B3D_REQUIRE(depth && xyz, "b3d_deproject_z16: NULL buffer");
    B3D_REQUIRE((bgr == nullptr) == (rgb == nullptr), "b3d_deproject_z16_color: bgr and rgb must both be given or both be NULL");
    B3D_REQUIRE(fx != 0.0f && fy != 0.0f, "b3d_deproject_z16: zero focal length");
    B3D_TRY(ctx->bind());
    return deproject_z16_batch(ctx, depth, bgr, w, h, 1, fx, fy, ppx, ppy, depth_scale, xyz, rgb);
}

int b3d_deproject_rgbd(b3d_ctx* ctx, const uint16_t* depth, const uint8_t* color, int w, int h, double fx, double fy, double cx, double cy,
                       float depth_scale, float depth_trunc, int flip_yz, double* xyz, double* rgb, int64_t* n_valid_h) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(n_valid_h != nullptr, "b3d_deproject_rgbd: n_valid_h is NULL");
    B3D_REQUIRE(w >= 0 && h >= 0, "b3d_deproject_rgbd: negative image size");
    *n_valid_h = 0;
    const int64_t n = (int64_t)w * h;
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(depth && xyz, "b3d_deproject_rgbd: NULL buffer");
    B3D_REQUIRE(fx != 0.0 && fy != 0.0, "b3d_deproject_rgbd: zero focal length");
    B3D_TRY(ctx->bind());
    DevBuf<int64_t> total;
    B3D_TRY(total.alloc(ctx, 1));
    RgbdPred pred{depth, depth_scale, depth_trunc};
    RgbdEmit emit{depth, color, w, fx, fy, cx, cy, depth_scale, flip_yz, xyz, rgb};
    B3D_TRY(compact(ctx, pred, emit, n, total.p));
    return ctx->download(n_valid_h, total.p, sizeof(int64_t));
}

int b3d_reproject_disparity_valid(b3d_ctx* ctx, const int16_t* disp, int w, int h, const double* Q_h, int min_disp16, float* xyz,
                                  int64_t* n_valid_h) {
    B3D_REQUIRE(ctx != nullptr && n_valid_h != nullptr, "b3d_reproject_disparity_valid: NULL argument");
    B3D_REQUIRE(w >= 0 && h >= 0, "b3d_reproject_disparity_valid: negative image size");
    *n_valid_h = 0;
    if ((int64_t)w * h == 0) return B3D_OK;
    B3D_REQUIRE(disp && xyz && Q_h, "b3d_reproject_disparity_valid: NULL buffer");
    B3D_TRY(ctx->bind());
    std::vector<int32_t> off;
    B3D_TRY(reproject_disparity_valid_batch(ctx, disp, w, h, 1, Q_h, min_disp16, xyz, &off));
    *n_valid_h = off[1];
    return B3D_OK;
}

int b3d_reproject_disparity(b3d_ctx* ctx, const int16_t* disp, int w, int h, const double* Q_h, float* xyz) {
    B3D_REQUIRE(ctx != nullptr, "ctx is NULL");
    B3D_REQUIRE(w >= 0 && h >= 0, "b3d_reproject_disparity: negative image size");
    const int64_t n = (int64_t)w * h;
    if (n == 0) return B3D_OK;
    B3D_REQUIRE(disp && xyz && Q_h, "b3d_reproject_disparity: NULL buffer");
    B3D_TRY(ctx->bind());
    QMat Q;
    for (int i = 0; i < 16; ++i) Q.q[i] = Q_h[i];
    B3D_LAUNCH(ctx, reproject_disparity_kernel, ctx->grid_for(n, 256, 1, 8), 256, 0, disp, w, n, Q, xyz);
    return B3D_OK;
}

}  // extern "C"
