// klt.cu — next row N1: the KLT front end that associates features frame to frame in this fork:
// cv::buildOpticalFlowPyramid (src/FrameKTL.cc:76) and cv::calcOpticalFlowPyrLK (src/Tracking.cc:1044-1047; 21x21 window,
// 5 levels, 30 iterations, eps 0.01, USE_INITIAL_FLOW | LK_GET_MIN_EIGENVALS).  Semantics restated from OpenCV
// video/lkpyramid.cpp and imgproc/pyramids.cpp (never its code).  Integer stages (pyrDown, Scharr, fixed-point window
// interpolation) are bit-exact; the float reductions run in warp-shuffle order, so positions agree with OpenCV to ~1e-3 px.
// C-ABI: include/uvip_orb.h.
#include "common.cuh"
#include <float.h>
#include <math.h>
#include <vector>

namespace uvip {

constexpr int KLT_MAXLEV = 10;

struct KltLevel {
    int w, h;            // interior size
    int B;               // border (win + 2): image planes hold reflect-101 pixels there, derivative planes zeros
    int istride;         // padded image row stride (bytes)
    int dstride;         // padded derivative row stride (short2 elements)
    size_t ioff, doff;   // offsets of the padded origins inside a slot (bytes / short2 elements)
};
struct KltPlan { int nlevels, win; KltLevel lv[KLT_MAXLEV]; };

__device__ __forceinline__ int refl(int p, int n) { return p < 0 ? -p : (p >= n ? 2 * (n - 1) - p : p); }

// level 0: padded copy of the input frame (reflect-101 border of win + 2 pixels)
__global__ void __launch_bounds__(256)
k_klt_import(const uint8_t* __restrict__ src, int stride, uint8_t* __restrict__ img, size_t src_pitch, size_t slot_img, const __grid_constant__ KltPlan P)
{
    src += (size_t)blockIdx.y * src_pitch; img += (size_t)blockIdx.y * slot_img;        // batched form: frame blockIdx.y -> slot blockIdx.y
    const KltLevel& L = P.lv[0];
    const int pw = L.w + 2 * L.B, ph = L.h + 2 * L.B;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= pw * ph) return;
    const int y = i / pw, x = i - y * pw;
    img[L.ioff + (size_t)y * L.istride + x] = __ldg(src + (size_t)refl(y - L.B, L.h) * stride + refl(x - L.B, L.w));
}

// level l from level l-1: cv::pyrDown (5x5 binomial, (sum + 128) >> 8, REFLECT_101), written for every padded position
__global__ void __launch_bounds__(256)
k_klt_pyrdown(uint8_t* __restrict__ img, int level, size_t slot_img, const __grid_constant__ KltPlan P)
{
    img += (size_t)blockIdx.y * slot_img;
    const KltLevel& D = P.lv[level]; const KltLevel& S = P.lv[level - 1];
    const int pw = D.w + 2 * D.B, ph = D.h + 2 * D.B;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= pw * ph) return;
    const int py = i / pw, px = i - py * pw;
    const int x = refl(px - D.B, D.w), y = refl(py - D.B, D.h);          // interior pixel this padded position mirrors
    const uint8_t* s = img + S.ioff + (size_t)(2 * y - 2 + S.B) * S.istride + (2 * x - 2 + S.B);   // source border >= 2 is reflect-101
    int acc = 0;
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const int kr = (r == 0 || r == 4) ? 1 : (r == 2 ? 6 : 4);
        const uint8_t* q = s + (size_t)r * S.istride;
        acc += kr * (q[2] * 6 + (q[1] + q[3]) * 4 + q[0] + q[4]);
    }
    img[D.ioff + (size_t)py * D.istride + px] = (uint8_t)((acc + 128) >> 8);
}

// Scharr derivatives (calcSharrDeriv): Ix = [3 10 3]^T x [-1 0 1], Iy = [-1 0 1]^T x [3 10 3], int16, interior only
__global__ void __launch_bounds__(256)
k_klt_scharr(const uint8_t* __restrict__ img, short2* __restrict__ der, int level, size_t slot_img, size_t slot_der, const __grid_constant__ KltPlan P)
{
    img += (size_t)blockIdx.y * slot_img; der += (size_t)blockIdx.y * slot_der;
    const KltLevel& L = P.lv[level];
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= L.w * L.h) return;
    const int y = i / L.w, x = i - y * L.w;
    const uint8_t* p = img + L.ioff + (size_t)(y + L.B) * L.istride + (x + L.B);
    const int s = L.istride;
    const int a00 = p[-s - 1], a01 = p[-s], a02 = p[-s + 1], a10 = p[-1], a12 = p[1], a20 = p[s - 1], a21 = p[s], a22 = p[s + 1];
    const int ix = ((a02 + a22) * 3 + a12 * 10) - ((a00 + a20) * 3 + a10 * 10);
    const int iy = ((a20 - a00) + (a22 - a02)) * 3 + (a21 - a01) * 10;
    der[L.doff + (size_t)(y + L.B) * L.dstride + (x + L.B)] = make_short2((short)ix, (short)iy);
}

#define KLT_DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

// LKTrackerInvoker: one warp per point, coarse to fine inside the kernel.  The window of the first image (interpolated
// intensities and derivatives, int16) lives in shared memory; every iteration the lanes interpolate the second image over
// the window, accumulate the mismatch vector in float and reduce it with shuffles.
__global__ void __launch_bounds__(256)
k_klt_track(const uint8_t* __restrict__ img0, const short2* __restrict__ der0, const uint8_t* __restrict__ img1,
            const float2* __restrict__ prev_pts, float2* __restrict__ next_pts, int n, int max_level, int max_iter, double eps2,
            int flags, double min_eig_thr, uint8_t* __restrict__ status, float* __restrict__ err, size_t slot_img, size_t slot_der,
            const __grid_constant__ KltPlan P)
{
    extern __shared__ short s_win[];                   // per warp: win*win intensities, then win*win (Ix, Iy)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pt = blockIdx.x * 8 + wib;
    if (pt >= n) return;
    {   // batched form: pair blockIdx.y tracks slot y -> slot y + 1 with its own n points (img1 is img0's base: one slot further)
        const size_t pr = blockIdx.y;
        img0 += pr * slot_img; der0 += pr * slot_der; img1 += pr * slot_img;
        prev_pts += pr * n; next_pts += pr * n; status += pr * n; if (err) err += pr * n;
    }
    const int win = P.win, npx = win * win;
    const int npx_pad = (npx + 1) & ~1;                   // keeps the short2 part 4-byte aligned
    short* Iw = s_win + (size_t)wib * 3 * npx_pad;
    short2* dIw = reinterpret_cast<short2*>(Iw + npx_pad);
    const unsigned rcpw = 0xFFFFFFFFu / (unsigned)win + 1u;
    const float half = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const float2 pp = prev_pts[pt];
    float2 np = next_pts[pt];
    bool ok = true;
    float errv = 0.f;
    for (int level = max_level; level >= 0; level--) {
        const KltLevel& L = P.lv[level];
        const float sc = (float)(1. / (1 << level));
        float px = pp.x * sc, py = pp.y * sc, nx, ny;
        if (level == max_level) {
            if (flags & 4) { nx = np.x * sc; ny = np.y * sc; } else { nx = px; ny = py; }
        } else { nx = np.x * 2.f; ny = np.y * 2.f; }
        np = make_float2(nx, ny);
        px -= half; py -= half;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -win || ipx >= L.w || ipy < -win || ipy >= L.h) { if (level == 0) { ok = false; errv = 0.f; } continue; }
        float a = px - ipx, b = py - ipy;
        int iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f), iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
        int iw10 = __float2int_rn((1.f - a) * b * 16384.f), iw11 = 16384 - iw00 - iw01 - iw10;
        const uint8_t* I0 = img0 + L.ioff + (size_t)(ipy + L.B) * L.istride + (ipx + L.B);
        const short2* D0 = der0 + L.doff + (size_t)(ipy + L.B) * L.dstride + (ipx + L.B);
        float A11 = 0.f, A12 = 0.f, A22 = 0.f;
        __syncwarp();
        for (int i = lane; i < npx; i += 32) {
            const int y = __umulhi((unsigned)i, rcpw), x = i - y * win;
            const uint8_t* p = I0 + (size_t)y * L.istride + x;
            const short2* q = D0 + (size_t)y * L.dstride + x;
            const int ival = KLT_DESCALE(p[0] * iw00 + p[1] * iw01 + p[L.istride] * iw10 + p[L.istride + 1] * iw11, 9);
            const short2 d00 = q[0], d01 = q[1], d10 = q[L.dstride], d11 = q[L.dstride + 1];
            const int ixval = KLT_DESCALE(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, 14);
            const int iyval = KLT_DESCALE(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, 14);
            Iw[i] = (short)ival; dIw[i] = make_short2((short)ixval, (short)iyval);
            A11 += (float)(ixval * ixval); A12 += (float)(ixval * iyval); A22 += (float)(iyval * iyval);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            A11 += __shfl_xor_sync(0xFFFFFFFFu, A11, o); A12 += __shfl_xor_sync(0xFFFFFFFFu, A12, o); A22 += __shfl_xor_sync(0xFFFFFFFFu, A22, o);
        }
        __syncwarp();
        A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
        if (flags & 8) errv = minEig;
        if (minEig < min_eig_thr || D < FLT_EPSILON) { if (level == 0) ok = false; continue; }
        D = 1.f / D;
        nx -= half; ny -= half;
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < max_iter; j++) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -win || inx >= L.w || iny < -win || iny >= L.h) { if (level == 0) ok = false; break; }
            a = nx - inx; b = ny - iny;
            iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f); iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
            iw10 = __float2int_rn((1.f - a) * b * 16384.f); iw11 = 16384 - iw00 - iw01 - iw10;
            const uint8_t* J = img1 + L.ioff + (size_t)(iny + L.B) * L.istride + (inx + L.B);
            float b1 = 0.f, b2 = 0.f;
            for (int i = lane; i < npx; i += 32) {
                const int y = __umulhi((unsigned)i, rcpw), x = i - y * win;
                const uint8_t* p = J + (size_t)y * L.istride + x;
                const int diff = KLT_DESCALE(p[0] * iw00 + p[1] * iw01 + p[L.istride] * iw10 + p[L.istride + 1] * iw11, 9) - Iw[i];
                const short2 d = dIw[i];
                b1 += (float)(diff * d.x); b2 += (float)(diff * d.y);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { b1 += __shfl_xor_sync(0xFFFFFFFFu, b1, o); b2 += __shfl_xor_sync(0xFFFFFFFFu, b2, o); }
            b1 *= FLT_SCALE; b2 *= FLT_SCALE;
            const float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
            nx += dx; ny += dy;
            np = make_float2(nx + half, ny + half);
            if ((double)dx * dx + (double)dy * dy <= eps2) break;
            if (j > 0 && fabsf(dx + pdx) < 0.01 && fabsf(dy + pdy) < 0.01) { np.x -= dx * 0.5f; np.y -= dy * 0.5f; break; }
            pdx = dx; pdy = dy;
        }
    }
    if (lane == 0) { next_pts[pt] = np; status[pt] = ok ? 1 : 0; if (err) err[pt] = errv; }
}


// The same tracker for the window the reference uses (21 x 21, src/Tracking.cc:1044), organised around what the first kernel's profile
// showed: it is bound by instruction issue, and two thirds of the instructions of an iteration were address arithmetic (a division and
// 64-bit pointers per pixel) and four byte loads from global memory per pixel.  Here
//   * the pixel offsets of a lane's 14 window pixels are computed once per kernel and live in registers;
//   * the part of the second image the window can reach is staged in shared memory once per level (32 rows x 36 bytes around the start
//     position, restaged only if the window moves more than 5 px), so an iteration reads it with 32-bit addresses and immediates;
//   * the derivative window is stored as float2: (float)(diff * d) == (float)diff * (float)d for integers (one rounding of the exact
//     product either way), which replaces two integer multiplies and two conversions per pixel by one conversion.
// Arithmetic and summation order are those of k_klt_track: the two kernels return identical bits.
constexpr int KLT_W21 = 21, KLT_NPX21 = KLT_W21 * KLT_W21, KLT_K21 = (KLT_NPX21 + 31) / 32;
constexpr int KLT_RS = 36, KLT_RR = 32, KLT_M = 5;       // staged region: row stride, rows, margin around the start position
#ifndef KLT_BUILD_UNROLL
#define KLT_BUILD_UNROLL 7
#endif
constexpr int KLT_BUILD_UNROLL_N = KLT_BUILD_UNROLL;
#ifndef KLT_WPB21_
#define KLT_WPB21_ 1
#endif
constexpr int KLT_WPB21 = KLT_WPB21_;                     // warps (= points) per CTA: a CTA lives as long as its slowest point, small CTAs waste fewer warp slots
constexpr int KLT_WARP_BYTES21 = 896 + 3536 + KLT_RS * KLT_RR;   // intensities (short), derivatives (float2), region of the second image

__global__ void __launch_bounds__(KLT_WPB21 * 32)
k_klt_track21(const uint8_t* __restrict__ img0, const short2* __restrict__ der0, const uint8_t* __restrict__ img1,
              const float2* __restrict__ prev_pts, float2* __restrict__ next_pts, int n, int max_level, int max_iter, double eps2,
              int flags, double min_eig_thr, uint8_t* __restrict__ status, float* __restrict__ err, size_t slot_img, size_t slot_der,
              const __grid_constant__ KltPlan P)
{
    extern __shared__ __align__(16) unsigned char s_klt[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pt = blockIdx.x * KLT_WPB21 + wib;
    if (pt >= n) return;
    {
        const size_t pr = blockIdx.y;
        img0 += pr * slot_img; der0 += pr * slot_der; img1 += pr * slot_img;
        prev_pts += pr * n; next_pts += pr * n; status += pr * n; if (err) err += pr * n;
    }
    constexpr int win = KLT_W21, npx = KLT_NPX21;
    short* Iw = reinterpret_cast<short*>(s_klt + (size_t)wib * KLT_WARP_BYTES21);
    float2* dF = reinterpret_cast<float2*>(s_klt + (size_t)wib * KLT_WARP_BYTES21 + 896);
    unsigned char* reg = s_klt + (size_t)wib * KLT_WARP_BYTES21 + 896 + 3536;
    int off[KLT_K21];                                     // offset of the lane's k-th window pixel inside the staged region
#pragma unroll
    for (int k = 0; k < KLT_K21; k++) { const int i = lane + 32 * k, y = i / win; off[k] = y * KLT_RS + (i - y * win); }
    const float half = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const float2 pp = prev_pts[pt];
    float2 np = next_pts[pt];
    bool ok = true;
    float errv = 0.f;
    for (int level = max_level; level >= 0; level--) {
        const KltLevel& L = P.lv[level];
        const float sc = (float)(1. / (1 << level));
        float px = pp.x * sc, py = pp.y * sc, nx, ny;
        if (level == max_level) {
            if (flags & 4) { nx = np.x * sc; ny = np.y * sc; } else { nx = px; ny = py; }
        } else { nx = np.x * 2.f; ny = np.y * 2.f; }
        np = make_float2(nx, ny);
        px -= half; py -= half;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -win || ipx >= L.w || ipy < -win || ipy >= L.h) { if (level == 0) { ok = false; errv = 0.f; } continue; }
        float a = px - ipx, b = py - ipy;
        int iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f), iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
        int iw10 = __float2int_rn((1.f - a) * b * 16384.f), iw11 = 16384 - iw00 - iw01 - iw10;
        const uint8_t* I0 = img0 + L.ioff + (size_t)(ipy + L.B) * L.istride + (ipx + L.B);
        const short2* D0 = der0 + L.doff + (size_t)(ipy + L.B) * L.dstride + (ipx + L.B);
        float A11 = 0.f, A12 = 0.f, A22 = 0.f;
        __syncwarp();
#pragma unroll KLT_BUILD_UNROLL_N
        for (int k = 0; k < KLT_K21; k++) {
            // (the last step covers 25 pixels: the other lanes repeat pixel 440 and drop the result, so that no load sits behind a branch
            //  and the loads of several steps can be in flight together)
            const int i = lane + 32 * k, ic = min(i, npx - 1);
            const int y = ic / win, x = ic - y * win;
            const uint8_t* p = I0 + y * L.istride + x;
            const short2* q = D0 + y * L.dstride + x;
            const int ival = KLT_DESCALE(p[0] * iw00 + p[1] * iw01 + p[L.istride] * iw10 + p[L.istride + 1] * iw11, 9);
            const short2 d00 = q[0], d01 = q[1], d10 = q[L.dstride], d11 = q[L.dstride + 1];
            const int ixval = KLT_DESCALE(d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11, 14);
            const int iyval = KLT_DESCALE(d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11, 14);
            if (i < npx) {
                Iw[i] = (short)ival; dF[i] = make_float2((float)(short)ixval, (float)(short)iyval);
                A11 += (float)(ixval * ixval); A12 += (float)(ixval * iyval); A22 += (float)(iyval * iyval);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            A11 += __shfl_xor_sync(0xFFFFFFFFu, A11, o); A12 += __shfl_xor_sync(0xFFFFFFFFu, A12, o); A22 += __shfl_xor_sync(0xFFFFFFFFu, A22, o);
        }
        __syncwarp();
        A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
        if (flags & 8) errv = minEig;
        if (minEig < min_eig_thr || D < FLT_EPSILON) { if (level == 0) ok = false; continue; }
        D = 1.f / D;
        nx -= half; ny -= half;
        float pdx = 0.f, pdy = 0.f;
        int ox = 0, oy = 0; bool staged = false;          // origin of the staged region in image coordinates
        for (int j = 0; j < max_iter; j++) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -win || inx >= L.w || iny < -win || iny >= L.h) { if (level == 0) ok = false; break; }
            if (!staged || inx < ox || inx - ox > KLT_RS - (win + 1) || iny < oy || iny - oy > KLT_RR - (win + 1)) {
                // (re)stage: rows oy .. oy + 31, 36 bytes from a 4-byte aligned column at or below inx - 5, clamped to the padded plane
                const int gx = min(max((inx - KLT_M + L.B) & ~3, 0), L.istride - KLT_RS);
                const int gy = min(max(iny - KLT_M + L.B, 0), L.h + 2 * L.B - KLT_RR);
                ox = gx - L.B; oy = gy - L.B; staged = true;
                const unsigned* src = reinterpret_cast<const unsigned*>(img1 + L.ioff + (size_t)(gy + lane) * L.istride + gx);
                unsigned* dst = reinterpret_cast<unsigned*>(reg + lane * KLT_RS);
                __syncwarp();
#pragma unroll
                for (int w = 0; w < KLT_RS / 4; w++) dst[w] = __ldg(src + w);
                __syncwarp();
            }
            a = nx - inx; b = ny - iny;
            iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f); iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
            iw10 = __float2int_rn((1.f - a) * b * 16384.f); iw11 = 16384 - iw00 - iw01 - iw10;
            const unsigned char* J = reg + (iny - oy) * KLT_RS + (inx - ox);
            float b1 = 0.f, b2 = 0.f;
#pragma unroll
            for (int k = 0; k < KLT_K21; k++) {
                if (k < KLT_K21 - 1 || lane + 32 * k < npx) {
                    const unsigned char* p = J + off[k];
                    const int diff = KLT_DESCALE(p[0] * iw00 + p[1] * iw01 + p[KLT_RS] * iw10 + p[KLT_RS + 1] * iw11, 9) - Iw[lane + 32 * k];
                    const float fd = (float)diff;
                    const float2 d = dF[lane + 32 * k];
                    b1 += fd * d.x; b2 += fd * d.y;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { b1 += __shfl_xor_sync(0xFFFFFFFFu, b1, o); b2 += __shfl_xor_sync(0xFFFFFFFFu, b2, o); }
            b1 *= FLT_SCALE; b2 *= FLT_SCALE;
            const float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
            nx += dx; ny += dy;
            np = make_float2(nx + half, ny + half);
            if ((double)dx * dx + (double)dy * dy <= eps2) break;
            if (j > 0 && fabsf(dx + pdx) < 0.01 && fabsf(dy + pdy) < 0.01) { np.x -= dx * 0.5f; np.y -= dy * 0.5f; break; }
            pdx = dx; pdy = dy;
        }
    }
    if (lane == 0) { next_pts[pt] = np; status[pt] = ok ? 1 : 0; if (err) err[pt] = errv; }
}

}  // namespace uvip

using namespace uvip;

struct uvip_klt {
    int device = 0, max_w = 0, max_h = 0, max_level = 0, win = 0, nslots = 0;
    int w = 0, h = 0;                                   // geometry of the current plan
    std::vector<int> slot_w, slot_h;                    // geometry every slot's pyramid was built for (0 = empty / invalidated)
    KltPlan plan;
    size_t img_bytes = 0, der_elems = 0;
    cudaStream_t stream = nullptr;
    DevBuf img, der, stage, pts;
    long long launches = 0;
    std::mutex mu;
};

static void klt_make_plan(uvip_klt* k, int w, int h)
{
    KltPlan& P = k->plan; memset(&P, 0, sizeof(P));
    P.win = k->win;
    size_t ioff = 0, doff = 0;
    int lw = w, lh = h, n = 0;
    for (int l = 0; l <= k->max_level && l < KLT_MAXLEV; l++) {
        if (l > 0) {
            lw = (lw + 1) / 2; lh = (lh + 1) / 2;
            if (lw <= k->win || lh <= k->win) break;          // lkpyramid.cpp: stop when a level is not larger than the window
        }
        KltLevel& L = P.lv[l];
        L.w = lw; L.h = lh; L.B = k->win + 2;
        L.istride = (int)align_up((size_t)lw + 2 * L.B, 16); L.dstride = (int)align_up((size_t)lw + 2 * L.B, 4);
        L.ioff = ioff; L.doff = doff;
        ioff += align_up((size_t)L.istride * (lh + 2 * L.B), 256);
        doff += align_up((size_t)L.dstride * (lh + 2 * L.B), 64);
        n = l + 1;
    }
    P.nlevels = n;
    k->img_bytes = ioff; k->der_elems = doff; k->w = w; k->h = h;
}

// --------------------------------------------------------------------------------------------------------
// N1, last third: cv::findFundamentalMat(pts0, pts1, FM_RANSAC, 1, 0.999, mask) of src/Tracking.cc:1062 (only the inlier mask is used
// there).  OpenCV's algorithm (calib3d/fundam.cpp, ptsetreg.cpp): 7-point minimal solver in a RANSAC loop, residual = max of the two
// squared point-to-epipolar-line distances in double, cast to float, compared with (float)(threshold^2); the mask of the best
// hypothesis is returned.  cv::RNG's samples cannot be reproduced, so the sampler is a counter-based SplitMix64 and a FIXED number of
// hypotheses is evaluated (no early stop: more than OpenCV tries) — one warp per hypothesis: lane 0 solves the 7-point problem
// (Gauss-Jordan null space, cubic by bisection: only + - * / sqrt, so the CPU oracle is reproduced bit for bit), all lanes count
// inliers.  The winner is the (count, lowest hypothesis, lowest root) maximum; a second kernel writes its mask.
// --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long rs_sm64(unsigned long long seed, unsigned long long k)
{
    unsigned long long z = seed + k * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double rs_det3(const double* F)
{
    return F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
}
__device__ __forceinline__ double rs_ceval(const double* c, double x) { return ((c[3] * x + c[2]) * x + c[1]) * x + c[0]; }
__device__ double rs_bisect(const double* c, double lo, double hi)
{
    double flo = rs_ceval(c, lo);
    for (int it = 0; it < 200; it++) {
        const double mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) break;
        const double fm = rs_ceval(c, mid);
        if ((fm < 0) == (flo < 0)) { lo = mid; flo = fm; } else hi = mid;
    }
    return 0.5 * (lo + hi);
}
__device__ int rs_cubic_roots(const double* c, double* r)
{
    double m = fabs(c[0]); if (fabs(c[1]) > m) m = fabs(c[1]); if (fabs(c[2]) > m) m = fabs(c[2]);
    if (fabs(c[3]) <= 1e-12 * m || c[3] == 0) {
        if (fabs(c[2]) <= 1e-12 * m || c[2] == 0) { if (c[1] == 0) return 0; r[0] = -c[0] / c[1]; return 1; }
        const double disc = c[1] * c[1] - 4 * c[2] * c[0];
        if (disc < 0) return 0;
        const double s = sqrt(disc);
        double a = (-c[1] - s) / (2 * c[2]), b = (-c[1] + s) / (2 * c[2]);
        if (a > b) { const double t = a; a = b; b = t; }
        r[0] = a; r[1] = b; return 2;
    }
    const double B = 1.0 + m / fabs(c[3]);
    const double disc = c[2] * c[2] - 3 * c[3] * c[1];
    int n = 0;
    if (disc <= 0) { r[n++] = rs_bisect(c, -B, B); return n; }
    const double s = sqrt(disc);
    double x1 = (-c[2] - s) / (3 * c[3]), x2 = (-c[2] + s) / (3 * c[3]);
    if (x1 > x2) { const double t = x1; x1 = x2; x2 = t; }
    const double fB0 = rs_ceval(c, -B), f1 = rs_ceval(c, x1), f2 = rs_ceval(c, x2), fB1 = rs_ceval(c, B);
    if ((fB0 < 0) != (f1 < 0)) r[n++] = rs_bisect(c, -B, x1);
    if ((f1 < 0) != (f2 < 0)) r[n++] = rs_bisect(c, x1, x2);
    if ((f2 < 0) != (fB1 < 0)) r[n++] = rs_bisect(c, x2, B);
    return n;
}
__device__ int rs_seven_point(const double* x0, const double* y0, const double* x1, const double* y1, double* Fs)
{
    double A[7][9];
    for (int i = 0; i < 7; i++) {
        A[i][0] = x1[i] * x0[i]; A[i][1] = x1[i] * y0[i]; A[i][2] = x1[i];
        A[i][3] = y1[i] * x0[i]; A[i][4] = y1[i] * y0[i]; A[i][5] = y1[i];
        A[i][6] = x0[i]; A[i][7] = y0[i]; A[i][8] = 1.0;
    }
    int pivcol[7]; int used[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 7; k++) {
        int br = k, bc = -1; double best = 0;
        for (int i = k; i < 7; i++)
            for (int j = 0; j < 9; j++)
                if (!used[j] && fabs(A[i][j]) > best) { best = fabs(A[i][j]); br = i; bc = j; }
        if (bc < 0 || best < 1e-300) return 0;
        if (br != k) for (int j = 0; j < 9; j++) { const double t = A[k][j]; A[k][j] = A[br][j]; A[br][j] = t; }
        used[bc] = 1; pivcol[k] = bc;
        const double inv = 1.0 / A[k][bc];
        for (int j = 0; j < 9; j++) A[k][j] *= inv;
        for (int i = 0; i < 7; i++)
            if (i != k) {
                const double f = A[i][bc];
                if (f != 0) for (int j = 0; j < 9; j++) A[i][j] -= f * A[k][j];
            }
    }
    int fre[2], nf = 0;
    for (int j = 0; j < 9; j++) if (!used[j]) fre[nf++] = j;
    double F1[9], F2[9];
    for (int j = 0; j < 9; j++) { F1[j] = 0; F2[j] = 0; }
    F1[fre[0]] = 1.0; F2[fre[1]] = 1.0;
    for (int k = 0; k < 7; k++) { F1[pivcol[k]] = -A[k][fre[0]]; F2[pivcol[k]] = -A[k][fre[1]]; }
    double G[9], c[4];
    const double d0 = rs_det3(F2), d1 = rs_det3(F1);
    for (int j = 0; j < 9; j++) G[j] = 2.0 * F2[j] - F1[j];
    const double dm = rs_det3(G);
    for (int j = 0; j < 9; j++) G[j] = 2.0 * F1[j] - F2[j];
    const double d2 = rs_det3(G);
    c[0] = d0;
    c[3] = (d2 - 3.0 * d1 + 3.0 * d0 - dm) / 6.0;
    c[2] = 0.5 * (d1 + dm) - d0;
    c[1] = d1 - d0 - c[2] - c[3];
    double r[3];
    const int nr = rs_cubic_roots(c, r);
    for (int k = 0; k < nr; k++) {
        double nrm = 0;
        for (int j = 0; j < 9; j++) { Fs[9 * k + j] = r[k] * F1[j] + (1.0 - r[k]) * F2[j]; nrm += Fs[9 * k + j] * Fs[9 * k + j]; }
        nrm = sqrt(nrm);
        if (nrm > 0) for (int j = 0; j < 9; j++) Fs[9 * k + j] /= nrm;
    }
    return nr;
}
__device__ __forceinline__ bool rs_inlier(const double* F, double x0, double y0, double x1, double y1, float t2)
{
    double a = F[0] * x0 + F[1] * y0 + F[2], b = F[3] * x0 + F[4] * y0 + F[5], c = F[6] * x0 + F[7] * y0 + F[8];
    const double s2 = 1.0 / (a * a + b * b), d2 = x1 * a + y1 * b + c;
    a = F[0] * x1 + F[3] * y1 + F[6]; b = F[1] * x1 + F[4] * y1 + F[7]; c = F[2] * x1 + F[5] * y1 + F[8];
    const double s1 = 1.0 / (a * a + b * b), d1 = x0 * a + y0 * b + c;
    const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
    return (float)(e1 > e2 ? e1 : e2) <= t2;
}

// best[0]: (count << 32 | ~(hypothesis * 4 + root)) maximum; Fbuf: 27 doubles per hypothesis
__global__ void __launch_bounds__(256)
k_ransac_fm(const float2* __restrict__ p0, const float2* __restrict__ p1, int n, float t2, int nhyp, unsigned long long seed,
            double* __restrict__ Fbuf, unsigned long long* __restrict__ best)
{
    __shared__ double s_F[8][27];
    __shared__ int s_nm[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x * 8 + warp;
    if (h >= nhyp) return;
    if (lane == 0) {
        int idx[7], ns = 0;
        for (int k = 1; k <= 16 && ns < 7; k++) {
            const int cand = (int)(rs_sm64(seed, (unsigned long long)h * 16ULL + (unsigned long long)k) % (unsigned long long)n);
            bool dup = false;
            for (int j = 0; j < ns; j++) dup |= idx[j] == cand;
            if (!dup) idx[ns++] = cand;
        }
        int nm = 0;
        if (ns == 7) {
            double x0[7], y0[7], x1[7], y1[7];
            for (int j = 0; j < 7; j++) { const float2 a = p0[idx[j]], b = p1[idx[j]]; x0[j] = a.x; y0[j] = a.y; x1[j] = b.x; y1[j] = b.y; }
            nm = rs_seven_point(x0, y0, x1, y1, s_F[warp]);
        }
        s_nm[warp] = nm;
        for (int j = 0; j < 9 * nm; j++) Fbuf[(size_t)h * 27 + j] = s_F[warp][j];
    }
    __syncwarp();
    const int nm = s_nm[warp];
    for (int k = 0; k < nm; k++) {
        double F[9];
#pragma unroll
        for (int j = 0; j < 9; j++) F[j] = s_F[warp][9 * k + j];
        int cnt = 0;
        for (int i = lane; i < n; i += 32) { const float2 a = p0[i], b = p1[i]; cnt += rs_inlier(F, a.x, a.y, b.x, b.y, t2) ? 1 : 0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        if (lane == 0 && cnt > 6)
            atomicMax(best, ((unsigned long long)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(h * 4 + k)));
    }
}

__global__ void __launch_bounds__(256)
k_ransac_mask(const float2* __restrict__ p0, const float2* __restrict__ p1, int n, float t2, const double* __restrict__ Fbuf,
              const unsigned long long* __restrict__ best, uint8_t* __restrict__ mask, double* __restrict__ Fout)
{
    const unsigned long long b = best[0];
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (b == 0) { if (i < n) mask[i] = 0; return; }
    const unsigned id = 0xFFFFFFFFu - (unsigned)(b & 0xFFFFFFFFu);
    const double* F = Fbuf + (size_t)(id >> 2) * 27 + 9 * (id & 3);
    if (i < 9) Fout[i] = F[i];
    if (i < n) { const float2 a = p0[i], c = p1[i]; mask[i] = rs_inlier(F, a.x, a.y, c.x, c.y, t2) ? 1 : 0; }
}

extern "C" {

int uvip_klt_create(int device, int max_width, int max_height, int win, int max_level, int nslots, uvip_klt** out)
{
    UVIP_CHECK_ARG(out && max_width > 0 && max_height > 0 && win >= 3 && win <= 63 && max_level >= 0 && nslots >= 2 && nslots <= 64);
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_last_error("no CUDA device available; libuvip_orb has no CPU fallback"); return UVIP_ERR_NO_DEVICE; }
    UVIP_CHECK_ARG(device >= 0 && device < ndev);
    DeviceGuard g(device);
    uvip_klt* k = new uvip_klt();
    k->device = device; k->max_w = max_width; k->max_h = max_height; k->win = win; k->max_level = max_level; k->nslots = nslots;
    klt_make_plan(k, max_width, max_height);
    const size_t ib = k->img_bytes, de = k->der_elems;
    int rc = k->img.reserve(ib * nslots) | k->der.reserve(de * sizeof(short2) * nslots) | k->stage.reserve((size_t)max_width * max_height);
    if (rc || cudaStreamCreateWithFlags(&k->stream, cudaStreamNonBlocking) != cudaSuccess) { uvip_klt_destroy(k); return UVIP_ERR_CUDA; }
    cudaMemset(k->der.p, 0, k->der.cap);                 // derivative borders are BORDER_CONSTANT 0 and are never written
    k->w = k->h = 0;
    k->slot_w.assign((size_t)nslots, 0); k->slot_h.assign((size_t)nslots, 0);
    *out = k;
    return UVIP_OK;
}

int uvip_klt_destroy(uvip_klt* k)
{
    if (!k) return UVIP_OK;
    DeviceGuard g(k->device);
    if (k->stream) { cudaStreamSynchronize(k->stream); cudaStreamDestroy(k->stream); }
    k->img.release(); k->der.release(); k->stage.release(); k->pts.release();
    delete k;
    return UVIP_OK;
}

int uvip_klt_build_pyramid(uvip_klt* k, int slot, const uint8_t* image, int w, int h, int stride, int* nlevels)
{
    UVIP_CHECK_ARG(k && image && slot >= 0 && slot < k->nslots && w > 0 && h > 0 && stride >= w && w <= k->max_w && h <= k->max_h);
    std::lock_guard<std::mutex> lk(k->mu);
    DeviceGuard g(k->device);
    cudaStream_t st = k->stream;
    if (w != k->w || h != k->h) {
        const size_t ib = k->img_bytes, de = k->der_elems;     // slot strides stay those of the maximal frame
        klt_make_plan(k, w, h);
        k->img_bytes = ib; k->der_elems = de;
        UVIP_CUDA(cudaMemsetAsync(k->der.p, 0, k->der.cap, st));
        // the other slots still hold pyramids laid out for the old geometry: they are unusable with the new plan
        for (int s = 0; s < k->nslots; s++) k->slot_w[(size_t)s] = k->slot_h[(size_t)s] = 0;
    }
    k->slot_w[(size_t)slot] = w; k->slot_h[(size_t)slot] = h;
    const KltPlan& P = k->plan;
    uint8_t* img = k->img.as<uint8_t>() + (size_t)slot * k->img_bytes;
    short2* der = k->der.as<short2>() + (size_t)slot * k->der_elems;
    UVIP_CUDA(cudaMemcpy2DAsync(k->stage.p, w, image, stride, w, h, cudaMemcpyHostToDevice, st));
    { const KltLevel& L = P.lv[0]; k_klt_import<<<div_up((L.w + 2 * L.B) * (L.h + 2 * L.B), 256), 256, 0, st>>>(k->stage.as<uint8_t>(), w, img, 0, 0, P); k->launches++; }
    for (int l = 1; l < P.nlevels; l++) {
        const KltLevel& L = P.lv[l];
        k_klt_pyrdown<<<div_up((L.w + 2 * L.B) * (L.h + 2 * L.B), 256), 256, 0, st>>>(img, l, 0, P); k->launches++;
    }
    for (int l = 0; l < P.nlevels; l++) {
        const KltLevel& L = P.lv[l];
        k_klt_scharr<<<div_up(L.w * L.h, 256), 256, 0, st>>>(img, der, l, 0, 0, P); k->launches++;
    }
    UVIP_CUDA(cudaGetLastError());
    UVIP_CUDA(cudaStreamSynchronize(st));
    if (nlevels) *nlevels = P.nlevels;
    return UVIP_OK;
}

int uvip_klt_get_level(uvip_klt* k, int slot, int level, uint8_t* img, int16_t* der, int* w, int* h)
{
    UVIP_CHECK_ARG(k && slot >= 0 && slot < k->nslots && k->w > 0 && level >= 0 && level < k->plan.nlevels);
    const KltLevel& L = k->plan.lv[level];
    if (w) *w = L.w; if (h) *h = L.h;
    DeviceGuard g(k->device);
    UVIP_CUDA(cudaDeviceSynchronize());
    if (img) UVIP_CUDA(cudaMemcpy2D(img, L.w, k->img.as<uint8_t>() + (size_t)slot * k->img_bytes + L.ioff + (size_t)L.B * L.istride + L.B, L.istride,
                                    L.w, L.h, cudaMemcpyDeviceToHost));
    if (der) UVIP_CUDA(cudaMemcpy2D(der, (size_t)L.w * 4, k->der.as<short2>() + (size_t)slot * k->der_elems + L.doff + (size_t)L.B * L.dstride + L.B,
                                    (size_t)L.dstride * 4, (size_t)L.w * 4, L.h, cudaMemcpyDeviceToHost));
    return UVIP_OK;
}

int uvip_klt_track(uvip_klt* k, int slot_prev, int slot_next, const float* prev_pts, float* next_pts, int n, int max_level,
                   int max_iter, double epsilon, int flags, double min_eig_threshold, uint8_t* status, float* err)
{
    UVIP_CHECK_ARG(k && n >= 0 && slot_prev >= 0 && slot_prev < k->nslots && slot_next >= 0 && slot_next < k->nslots && k->w > 0);
    if (n == 0) return UVIP_OK;
    UVIP_CHECK_ARG(prev_pts && next_pts && status);
    std::lock_guard<std::mutex> lk(k->mu);
    if (k->slot_w[(size_t)slot_prev] != k->w || k->slot_h[(size_t)slot_prev] != k->h || k->slot_w[(size_t)slot_next] != k->w || k->slot_h[(size_t)slot_next] != k->h) {
        set_last_error("KLT slots %d / %d do not hold pyramids of the current %dx%d geometry (a build with another frame size invalidates the other slots)",
                       slot_prev, slot_next, k->w, k->h);
        return UVIP_ERR_ARG;
    }
    DeviceGuard g(k->device);
    if (max_iter < 0) max_iter = 0; if (max_iter > 100) max_iter = 100;            // calcOpticalFlowPyrLK clamps the criteria
    if (epsilon < 0) epsilon = 0; if (epsilon > 10) epsilon = 10;
    const double eps2 = epsilon * epsilon;
    const KltPlan& P = k->plan;
    if (max_level > P.nlevels - 1) max_level = P.nlevels - 1;
    if (max_level < 0) max_level = 0;
    int rc;
    const size_t o_next = align_up((size_t)n * 8, 256), o_st = 2 * o_next, o_err = o_st + align_up((size_t)n, 256);
    if ((rc = k->pts.reserve(o_err + (size_t)n * 4))) return rc;
    uint8_t* base = k->pts.as<uint8_t>();
    cudaStream_t st = k->stream;
    UVIP_CUDA(cudaMemcpyAsync(base, prev_pts, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    UVIP_CUDA(cudaMemcpyAsync(base + o_next, next_pts, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    const int npx = P.win * P.win;
    const bool w21 = P.win == KLT_W21 && !getenv("UVIP_KLT_GENERIC");
    const size_t smem = w21 ? (size_t)KLT_WPB21 * KLT_WARP_BYTES21 : (size_t)8 * 3 * ((npx + 1) & ~1) * sizeof(short);
    auto kern = w21 ? k_klt_track21 : k_klt_track;
    const int wpb = w21 ? KLT_WPB21 : 8;
    UVIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<div_up(n, wpb), wpb * 32, smem, st>>>(k->img.as<uint8_t>() + (size_t)slot_prev * k->img_bytes, k->der.as<short2>() + (size_t)slot_prev * k->der_elems,
                                                k->img.as<uint8_t>() + (size_t)slot_next * k->img_bytes, (const float2*)base, (float2*)(base + o_next), n,
                                                max_level, max_iter, eps2, flags, min_eig_threshold, base + o_st, (float*)(base + o_err), 0, 0, P);
    k->launches++;
    UVIP_CUDA(cudaGetLastError());
    UVIP_CUDA(cudaMemcpyAsync(next_pts, base + o_next, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaMemcpyAsync(status, base + o_st, (size_t)n, cudaMemcpyDeviceToHost, st));
    if (err) UVIP_CUDA(cudaMemcpyAsync(err, base + o_err, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    return UVIP_OK;
}

// Batched, device-resident form of the two calls above for a SEQUENCE: the pyramids of nframes frames are built in slots 0 .. nframes-1
// (one launch per pyramid level over all frames), then frame f -> f + 1 is tracked for f = 0 .. nframes-2 in ONE launch (one warp per
// point and pair).  All pointers are device memory; asynchronous on `stream` (NULL = the handle's stream).
int uvip_klt_track_sequence_device(uvip_klt* k, const uint8_t* d_frames, int nframes, int w, int h, int stride, size_t frame_pitch,
                                   const float* d_prev_pts, float* d_next_pts, int npts, int max_level, int max_iter, double epsilon, int flags,
                                   double min_eig_threshold, uint8_t* d_status, float* d_err, void* stream)
{
    UVIP_CHECK_ARG(k && d_frames && d_prev_pts && d_next_pts && d_status && nframes >= 2 && nframes <= k->nslots && npts >= 1);
    UVIP_CHECK_ARG(w > 0 && h > 0 && stride >= w && w <= k->max_w && h <= k->max_h && frame_pitch >= (size_t)stride * (h - 1) + w);
    std::lock_guard<std::mutex> lk(k->mu);
    DeviceGuard g(k->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : k->stream;
    if (w != k->w || h != k->h) {
        const size_t ib = k->img_bytes, de = k->der_elems;
        klt_make_plan(k, w, h);
        k->img_bytes = ib; k->der_elems = de;
        UVIP_CUDA(cudaMemsetAsync(k->der.p, 0, k->der.cap, st));
        for (int s2 = 0; s2 < k->nslots; s2++) k->slot_w[(size_t)s2] = k->slot_h[(size_t)s2] = 0;
    }
    for (int s2 = 0; s2 < nframes; s2++) { k->slot_w[(size_t)s2] = w; k->slot_h[(size_t)s2] = h; }
    const KltPlan& P = k->plan;
    uint8_t* img = k->img.as<uint8_t>(); short2* der = k->der.as<short2>();
    const size_t si = k->img_bytes, sd = k->der_elems;
    { const KltLevel& L = P.lv[0]; k_klt_import<<<dim3(div_up((L.w + 2 * L.B) * (L.h + 2 * L.B), 256), nframes), 256, 0, st>>>(d_frames, stride, img, frame_pitch, si, P); }
    for (int l = 1; l < P.nlevels; l++) { const KltLevel& L = P.lv[l]; k_klt_pyrdown<<<dim3(div_up((L.w + 2 * L.B) * (L.h + 2 * L.B), 256), nframes), 256, 0, st>>>(img, l, si, P); }
    for (int l = 0; l < P.nlevels; l++) { const KltLevel& L = P.lv[l]; k_klt_scharr<<<dim3(div_up(L.w * L.h, 256), nframes), 256, 0, st>>>(img, der, l, si, sd, P); }
    k->launches += 2 * P.nlevels;
    if (max_iter < 0) max_iter = 0; if (max_iter > 100) max_iter = 100;
    if (epsilon < 0) epsilon = 0; if (epsilon > 10) epsilon = 10;
    if (max_level > P.nlevels - 1) max_level = P.nlevels - 1;
    if (max_level < 0) max_level = 0;
    const int npx = P.win * P.win;
    const bool w21 = P.win == KLT_W21 && !getenv("UVIP_KLT_GENERIC");
    const size_t smem = w21 ? (size_t)KLT_WPB21 * KLT_WARP_BYTES21 : (size_t)8 * 3 * ((npx + 1) & ~1) * sizeof(short);
    auto kern = w21 ? k_klt_track21 : k_klt_track;
    const int wpb = w21 ? KLT_WPB21 : 8;
    UVIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(div_up(npts, wpb), nframes - 1), wpb * 32, smem, st>>>(img, der, img + si, (const float2*)d_prev_pts, (float2*)d_next_pts, npts, max_level, max_iter,
                                                                      epsilon * epsilon, flags, min_eig_threshold, d_status, d_err, si, sd, P);
    k->launches++;
    UVIP_CUDA(cudaGetLastError());
    return UVIP_OK;
}

int uvip_klt_ransac_fundamental(uvip_klt* k, const float* pts0, const float* pts1, int n, double threshold, int nhyp, uint8_t* mask, double* F,
                                int* ninliers)
{
    UVIP_CHECK_ARG(k && pts0 && pts1 && mask && ninliers && n >= 0 && threshold > 0 && nhyp >= 1 && nhyp <= (1 << 20));
    if (n < 15) { set_last_error("findFundamentalMat(FM_RANSAC) needs at least 15 points (OpenCV switches to LMedS below that; not built)"); return UVIP_ERR_UNSUPPORTED; }
    std::lock_guard<std::mutex> lk(k->mu);
    DeviceGuard g(k->device);
    int rc;
    const size_t o_p1 = align_up((size_t)n * 8, 256), o_F = 2 * o_p1, o_best = o_F + align_up((size_t)nhyp * 27 * 8, 256), o_mask = o_best + 256,
                 o_Fout = o_mask + align_up((size_t)n, 256);
    if ((rc = k->pts.reserve(o_Fout + 128))) return rc;
    uint8_t* base = k->pts.as<uint8_t>();
    cudaStream_t st = k->stream;
    UVIP_CUDA(cudaMemcpyAsync(base, pts0, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    UVIP_CUDA(cudaMemcpyAsync(base + o_p1, pts1, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    UVIP_CUDA(cudaMemsetAsync(base + o_best, 0, 8, st));
    UVIP_CUDA(cudaMemsetAsync(base + o_Fout, 0, 72, st));
    const float t2 = (float)(threshold * threshold);
    k_ransac_fm<<<div_up(nhyp, 8), 256, 0, st>>>((const float2*)base, (const float2*)(base + o_p1), n, t2, nhyp, 0x5EEDULL, (double*)(base + o_F),
                                                 (unsigned long long*)(base + o_best));
    k_ransac_mask<<<div_up(n > 9 ? n : 9, 256), 256, 0, st>>>((const float2*)base, (const float2*)(base + o_p1), n, t2, (const double*)(base + o_F),
                                                              (const unsigned long long*)(base + o_best), base + o_mask, (double*)(base + o_Fout));
    k->launches += 2;
    UVIP_CUDA(cudaGetLastError());
    unsigned long long best = 0; double Fh[9];
    UVIP_CUDA(cudaMemcpyAsync(mask, base + o_mask, (size_t)n, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaMemcpyAsync(&best, base + o_best, 8, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaMemcpyAsync(Fh, base + o_Fout, 72, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    *ninliers = (int)(best >> 32);
    if (F) memcpy(F, Fh, 72);
    return UVIP_OK;
}

long long uvip_klt_launch_count(const uvip_klt* k) { return k ? k->launches : 0; }

}  // extern "C"
