// extractor.cu — USLAM::ORBextractor on sm_100a.  C-ABI: include/uvip_orb.h.
//
// Pipeline per launch group (a batch of equally sized frames resident in HBM), one kernel per stage over the
// whole batch:  K1 import + cascaded fixed-point bilinear pyramid with reflect-101 border  ->  K2/K3 tiled FAST-9/16
// score map, cell-local NMS, warp-aggregated candidate emission + per-cell maxima (the empty-cell retry rule)  ->
// K4 quadtree distribution, one CTA per (frame, level), nodes kept in list order in shared memory  ->
// K6 7x7 fixed-point Gaussian  ->  K8 selection (FullDetect / occupancy grid)  ->  K5+K7 one warp per keypoint:
// IC_Angle on the unblurred level and rotated BRIEF on the blurred one.
//
// Reference semantics restated here (never its code): src/ORBextractor.cc:74-78,125-195,458-512,748-1004,1006-1287.
#include "common.cuh"
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace uvip {

constexpr int MAXLEV = 16;
constexpr int EDGE = 16;            // EDGE_THRESHOLD: layout padding of every pyramid plane
constexpr int BORDER_W = 4;         // border pixels actually materialised (only <=3 are ever read: SURVEY A.2)
constexpr int HALF_PATCH = 15;

static const int8_t h_pattern[1024] = {
#include "orb_pattern.inc"
};
__constant__ int8_t c_pattern[1024];
__constant__ int c_umax[HALF_PATCH + 1];

struct LevelInfo {
    int w, h;                 // interior size
    int pstride;              // padded row stride in bytes (multiple of 32)
    unsigned poff;            // byte offset of the padded origin inside a frame block
    int ncols, nrows, wcell, hcell, cell_off;     // FAST cells (src/ORBextractor.cc:767-770)
    int quota;                // mnFeaturesPerLevel
    int raw_cap, raw_off;     // candidate list (u32 entries) inside the frame's candidate block
    int kp_cap, kp_off;       // quadtree winners
    int ftile_off, fntx, fnty;        // FAST tiles of FAST_CW x FAST_CH cells
    int btile_off, bntx, bnty;    // blur tiles
    int tab_off;              // offset of this level's resize tables
    int nini;                 // quadtree roots
    float hx;
    float scale;              // mvScaleFactor
    float size;               // keypoint size (float)(int)(31*scale)
};

struct Plan {
    int nlevels, W, H;
    int fast_th, retry_th, t1, t2, tmin;   // effective thresholds (>=1)
    int cells_per_frame, raw_per_frame, kp_per_frame;
    int ftiles, btiles, tile_tab_off;
    int node_cap;             // quadtree node capacity (power of two)
    unsigned rcp_cpr;         // ceil(2^32 / chunks-per-row) for the import kernel
    unsigned f_rcp_srow;      // ceil(2^32 / f_srow)
    int rs_boxw, rs_boxh;     // resize: TMA box of the source level
    int f_irow, f_irows, f_srow, f_srows, f_gw;   // FAST shared-memory carve-up (largest tile over all levels)
    // k_fast2 (one warp per cell): per-warp score map (f2_srow x f2_srows bytes, padded to f2_score_bytes), group-queue capacity,
    // bytes of one warp's private area, ceil(2^32 / ftiles)
    int f2_srow, f2_srows, f2_score_bytes, f2_gcap, f2_wbytes;
    int f2_imgbytes, f2_imgstride;               // bytes of a staged tile (row stride x rows) and the 128-byte aligned distance between stages
    int f2_tiles, f2_tab_off, f2_irows;          // tiles of F2_CW x 1 cells per frame, their table inside tabs, rows of a staged tile
    unsigned f2_rcp_tiles;
    unsigned long long frame_bytes;
    LevelInfo lv[MAXLEV];
    // host-side only, kept behind lv[] so that no kernel-visible offset depends on it: bit l = the byte pairs of columns 0..2 of
    // every 4-column group of level l lie inside the first two window words of the resize (k_resize<NARROW>)
    unsigned rs_narrow_mask;
};

// --------------------------------------------------------------------------------------------------------
// K1a: import level 0 into the padded plane (+ reflect-101 border), 16 pixels per thread
// --------------------------------------------------------------------------------------------------------
// reflect-101 source coordinate of ring coordinate r in [-BORDER_W, n + BORDER_W)
__device__ __forceinline__ int reflect101(int r, int n) { return r < 0 ? -r : (r >= n ? 2 * (n - 1) - r : r); }

// Border ring of a tile of a freshly written plane: every ring pixel whose reflect-101 SOURCE pixel lies inside [X0,X1) x [Y0,Y1) is
// owned by that tile.  All 256 threads of the CTA, no lists and no divisions: interior
// tiles leave after four uniform compares; edge tiles copy (A) the mirrored columns of their interior rows and (B) the mirrored rows
// over their columns and mirrored columns.  Sources are pixels this CTA wrote before the barrier that precedes the call.
__device__ __forceinline__ void tile_ring(uint8_t* inner, int ps, int w, int h, int X0, int X1, int Y0, int Y1, int tid)
{
    const bool hl = X0 <= BORDER_W && X1 > 1, hr = X1 > w - 2 - BORDER_W && X0 < w - 1;
    const bool vt = Y0 <= BORDER_W && Y1 > 1, vb = Y1 > h - 2 - BORDER_W && Y0 < h - 1;
    if (!(hl | hr | vt | vb)) return;                         // uniform
    const int th = Y1 - Y0;
    // (A) mirrored columns of the tile's rows: one thread per (row, side), its four ring pixels in sequence
    if (hl | hr)
        for (int i = tid; i < 2 * th; i += 256) {
            const int side = i & 1;
            if (side ? hr : hl) {
                uint8_t* row = inner + (ptrdiff_t)(Y0 + (i >> 1)) * ps;
#pragma unroll
                for (int k = 1; k <= BORDER_W; k++) {
                    const int c = side ? w - 1 - k : k, p = side ? w - 1 + k : -k;
                    if (c >= X0 && c < X1) row[p] = row[c];
                }
            }
        }
    // (B) mirrored rows: threads 0..63 copy the tile's columns in 16-byte pieces (tiles start at multiples of 128, plane rows are
    //     16-byte aligned), threads 64..127 the 4 + 4 ring columns beside the image; mirror row m = t >> 3 of 2 * BORDER_W
    if ((vt | vb) && tid < 128) {
        const int m = (tid >> 3) & 7, k = (m % BORDER_W) + 1;
        const int sr = m < BORDER_W ? k : h - 1 - k, pr = m < BORDER_W ? -k : h - 1 + k;
        if (sr >= Y0 && sr < Y1) {
            const uint8_t* src = inner + (ptrdiff_t)sr * ps;
            uint8_t* dst = inner + (ptrdiff_t)pr * ps;
            if (tid < 64) {
                for (int x = X0 + 16 * (tid & 7); x < X1; x += 128) {
                    if (x + 16 <= X1) *reinterpret_cast<uint4*>(dst + x) = *reinterpret_cast<const uint4*>(src + x);
                    else for (int c = x; c < X1; c++) dst[c] = src[c];
                }
            } else {
                const int e = tid & 7;
                const int p = e < BORDER_W ? -1 - e : w + e - BORDER_W;
                const int c = reflect101(p, w);
                if (c >= X0 && c < X1) dst[p] = src[c];
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_import(const uint8_t* __restrict__ frames, int stride, size_t frame_pitch, int vec_ok, int copy_blocks, uint8_t* __restrict__ pyr,
         const __grid_constant__ Plan P)
{
    const LevelInfo& L = P.lv[0];
    const int f = blockIdx.y;
    const uint8_t* fsrc = frames + (size_t)f * frame_pitch;
    uint8_t* inner = pyr + (size_t)f * P.frame_bytes + L.poff + (size_t)EDGE * L.pstride + EDGE;
    if ((int)blockIdx.x >= copy_blocks) {                       // ring blocks: reflect-101 border straight from the input frame
        // 2 * BORDER_W ring blocks: block m copies mirror row m (over the image columns and the mirrored columns) and an eighth of the
        // mirrored columns of the interior rows; no lists, no divisions
        const int m = blockIdx.x - copy_blocks, w = L.w, h = L.h, ps = L.pstride;
        {
            const int k = (m % BORDER_W) + 1;
            const int sr = m < BORDER_W ? k : h - 1 - k, pr = m < BORDER_W ? -k : h - 1 + k;
            for (int t = threadIdx.x; t < w + 2 * BORDER_W; t += 256) {
                const int p = t - BORDER_W;
                inner[(ptrdiff_t)pr * ps + p] = __ldg(fsrc + (size_t)sr * stride + reflect101(p, w));
            }
        }
        for (int i = m * 256 + threadIdx.x; i < h * 2 * BORDER_W; i += 2 * BORDER_W * 256) {
            const int r = i / (2 * BORDER_W), c = i % (2 * BORDER_W), k = (c % BORDER_W) + 1;      // powers of two
            inner[(ptrdiff_t)r * ps + (c < BORDER_W ? -k : w - 1 + k)] = __ldg(fsrc + (size_t)r * stride + (c < BORDER_W ? k : w - 1 - k));
        }
        return;
    }
    const int cpr = (L.w + 15) >> 4;                        // 16-px chunks per row
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= cpr * L.h) return;
    const int y = __umulhi((unsigned)idx, P.rcp_cpr), x0 = (idx - y * cpr) << 4;
    const uint8_t* src = fsrc + (size_t)y * stride + x0;
    uint8_t* dst = inner + (size_t)y * L.pstride + x0;
    if (x0 + 16 <= L.w && vec_ok) *reinterpret_cast<uint4*>(dst) = __ldg(reinterpret_cast<const uint4*>(src));
    else {
#pragma unroll 1
        for (int k = 0; k < 16 && x0 + k < L.w; k++) dst[k] = __ldg(src + k);      // never touches the border ring
    }
}

// --------------------------------------------------------------------------------------------------------
// K1b: level l from level l-1, cv::resize INTER_LINEAR 8-bit fixed point (11-bit coefficients), cascaded
// (src/ORBextractor.cc:982).  One CTA per 128x96 output tile; the source region arrives as one TMA box.  A thread
// owns 4 output columns (source offsets and coefficient pairs live in registers) and walks down 12 rows; the
// horizontal pass of a source row is one IDP.2A per pixel on a funnel-shifted word and is reused by the next output
// row whenever that row's upper source row is this row's lower one (5 rows out of 6 at scale 1.2).
// tables (host-built, per level): xofs[w] int32, xcoef[w] {a0,a1} int16x2, yofs[h], ycoef[h]
// --------------------------------------------------------------------------------------------------------
#ifndef RS_R_
#define RS_R_ 12
#endif
constexpr int RS_W = 128, RS_RMAX = RS_R_, RS_H = 8 * RS_RMAX;     // widest tile; launches that would not fill the GPU use half-height tiles
constexpr int RS_RSMALL = RS_RMAX / 2;

template <bool NARROW, int RS_R>
__global__ void __launch_bounds__(256)
k_resize(const CUtensorMap* __restrict__ tmaps, uint8_t* __restrict__ pyr, const int* __restrict__ tabs, int level, const __grid_constant__ Plan P)
{
    extern __shared__ __align__(128) unsigned char s_rs[];
    __shared__ __align__(8) uint64_t s_mbar;
    const LevelInfo& D = P.lv[level];
    const int f = blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = blockIdx.y, tx = blockIdx.x;
    const int X0 = tx * RS_W, Y0 = ty * (8 * RS_R);
    const int* xofs = tabs + D.tab_off;
    const int* xcoef = xofs + D.w;
    const int* yofs = xcoef + D.w;
    const int* ycoef = yofs + D.h;
    const int bx = (__ldg(xofs + X0) + EDGE) & ~15;        // padded source coords of the box origin (16 B aligned for TMA)
    const int by = __ldg(yofs + Y0) + EDGE;
    if (tid == 0) { mbar_init(&s_mbar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&s_mbar, (unsigned)(P.rs_boxw * P.rs_boxh));
        tma_load_3d(s_rs, tmaps + 2 * MAXLEV + (level - 1), bx, by, f, &s_mbar);
    }
    const int x = X0 + 4 * lane;
    const int yw = Y0 + RS_R * warp;
    const int Dw = D.w, Dh = D.h, dps = D.pstride;         // level constants in registers (P.lv[level] is an indexed constant load)
    const bool live = yw < Dh;                             // warp-uniform (the row table is exchanged by shuffles); columns >= w are computed on clamped sources and never stored
    uint8_t* inner = pyr + (size_t)f * P.frame_bytes + D.poff + (size_t)EDGE * dps + EDGE;
    if (live) {
        // A thread's 4 columns read source bytes d_k, d_k + 1 of a 12-byte window that starts at the aligned word of column 0
        // (d_3 + 1 <= 11 is checked at create time).  The window is 3 LDS per source row; PRMT picks the byte pair of each
        // column: one PRMT over words 0-1, a second one bringing in word 2 where the pair can reach it (NARROW: only column 3;
        // 8 LDS + 4 funnel shifts per row made the kernel shared-memory bound: 2-way bank conflicts at a lane stride of 1.2 words).
        unsigned selA[4], selB[4], ck[4];
        const unsigned char* a0;
        {
            const int s0 = __ldg(xofs + min(x, Dw - 1)) + EDGE - bx;    // byte offset inside a box row
            a0 = s_rs + (s0 & ~3);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int xk = min(x + k, Dw - 1);
                const int d = __ldg(xofs + xk) + EDGE - bx - (s0 & ~3);
                ck[k] = (unsigned)__ldg(xcoef + xk);
                selA[k] = (unsigned)((d <= 7 ? d : 0) | ((d + 1 <= 7 ? d + 1 : 0) << 4));
                selB[k] = (unsigned)((d <= 7 ? 0 : d - 4) | ((d + 1 <= 7 ? 1 : d - 3) << 4));
            }
        }
        // the warp's row table: lane j holds source row and coefficient pair of output row yw + j (one load per warp, not per row)
        static_assert(RS_R <= 32, "row table lives in one warp");
        const int nrow = min(RS_R, Dh - yw);
        int my_sy = 0, my_yc = 0;
        if (lane < nrow) { my_sy = (__ldg(yofs + yw + lane) + EDGE - by) * P.rs_boxw; my_yc = __ldg(ycoef + yw + lane); }
        const int rowb = P.rs_boxw;
        uint8_t* dst = inner + (size_t)yw * dps + x;       // one row down per iteration
        const bool fullw = x + 3 < Dw;
        mbar_wait(&s_mbar, 0);
        // horizontal pass of one source row (byte offset ro): (a0 * s[x] + a1 * s[x+1]) >> 4 for the thread's 4 columns
        auto hrow = [&](int ro, int* h) {
            const unsigned* q = reinterpret_cast<const unsigned*>(a0 + ro);
            const unsigned w0 = q[0], w1 = q[1], w2 = q[2];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                unsigned t = __byte_perm(w0, w1, selA[k]);
                if (!NARROW || k == 3) t = __byte_perm(t, w2, selB[k]);
                h[k] = (int)__dp2a_lo(ck[k], t, 0u) >> 4;
            }
        };
        // one output row: `top` holds source row sy if the previous output row's lower row was sy (5 rows out of 6 at scale
        // 1.2), otherwise it is recomputed; the lower row goes to `bot`, which is the next row's `top` (roles swap, no moves)
        int prev_ro = -0x40000000;
        auto step = [&](int j, int* top, int* bot) {
            const int ro = __shfl_sync(0xFFFFFFFFu, my_sy, j), yc = __shfl_sync(0xFFFFFFFFu, my_yc, j);
            if (ro != prev_ro + rowb) hrow(ro, top);       // warp-uniform
            hrow(ro + rowb, bot);                          // row sy+1 is valid (border row) when sy is the last row, and b1 == 0 there
            prev_ro = ro;
            const int b0 = (short)(yc & 0xFFFF), b1 = yc >> 16;
            unsigned o = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // <= 255 by construction: coefficient pairs sum to at most 2049 on each axis
                const int v = (((b0 * top[k]) >> 16) + ((b1 * bot[k]) >> 16) + 2) >> 2;
                o |= (unsigned)v << (8 * k);
            }
            if (fullw) *reinterpret_cast<unsigned*>(dst) = o;
            else for (int k = 0; k < 4; k++) if (x + k < Dw) dst[k] = (uint8_t)(o >> (8 * k));
            dst += dps;
            asm volatile("" : "+l"(dst));                 // keep the running pointer: the unrolled rows otherwise recompute a 64-bit address each
        };
        int hA[4] = {0, 0, 0, 0}, hB[4] = {0, 0, 0, 0};
        static_assert(RS_R % 2 == 0, "rows are processed in pairs");
#pragma unroll
        for (int j = 0; j < RS_R; j += 2) {
            if (j >= nrow) break;
            step(j, hA, hB);
            if (j + 1 >= nrow) break;
            step(j + 1, hB, hA);
        }
    } else mbar_wait(&s_mbar, 0);
    // reflect-101 ring: every ring pixel whose source lies in this tile, after the tile is complete
    __syncthreads();
    tile_ring(inner, dps, Dw, Dh, X0, min(X0 + RS_W, Dw), Y0, min(Y0 + 8 * RS_R, Dh), tid);
}

// --------------------------------------------------------------------------------------------------------
// K2+K3: FAST-9/16 exactly as the reference runs it: cv::FAST(th, nms=true) per 30-px cell ROI, and a second run at
// the retry threshold when a cell returned nothing (src/ORBextractor.cc:772-812).  One CTA per tile of up to 4x2
// cells (cell-aligned, so the empty-cell decision never leaves the CTA):
//   A   stage the tile + 3 px halo in shared memory with aligned 32-bit loads
//   B1  SWAR pre-test, 4 pixels per thread per step, all groups: |ring - v| > t on the compass pairs (0,8),(4,12)
//       via VABSDIFF4 + a per-byte carry trick; a 9-arc contains one pixel of every opposite pair, so groups where
//       no pixel passes are dropped (85 % on the benchmark frame).  Survivors are compacted (warp ballot) so that
//   B2  the groups are tested on dense warps at the 8 even ring positions (ring pixel k of 4 neighbouring pixels is one
//       funnel-shifted 32-bit word; |ring - v| > t is one VABSDIFF4 + 3 ALU ops per word): a 9-arc contains 4 consecutive
//       even positions, so pixels without such a run are dropped; the rest are queued
//   C   exact score s = max_arc min_k |ring_k - v| - 1 (== OpenCV cornerScore, DPX min3/max3) of every queued pixel;
//       s >= t  <=>  the pixel is a FAST-9 corner at t, so this is also the exact corner decision; corners are
//       compacted once more
//   D   3x3 strict NMS inside the cell, survivors appended to the (frame, level) raw-corner list (warp-aggregated)
// Pass 2 repeats B..D at the retry threshold for the cells of the tile that produced no survivor.
// --------------------------------------------------------------------------------------------------------
constexpr int FAST_CW = 4, FAST_CH = 2;      // cells per tile
constexpr int FAST_WARPS = 8;

// bit 7 of every byte: a > t for a constant threshold; k7 = (0x7f - (t & 0x7f)) * 0x01010101
__device__ __forceinline__ unsigned swar_gt_const(unsigned a, unsigned k7, bool t_low)
{
    const unsigned s = (a & 0x7f7f7f7fu) + k7;
    return t_low ? (s | a) : (s & a);          // t < 128: high bit alone decides; t >= 128: need both
}

// B1: does any of the 4 pixels of this group pass the compass pre-test?  (bit 7 per byte)
__device__ __forceinline__ unsigned fast_pre4(const unsigned* __restrict__ W, int rs, unsigned k7, bool t_low)
{
    const unsigned v = W[0];
    const unsigned g0 = swar_gt_const(__vabsdiffu4(W[3 * rs], v), k7, t_low);
    const unsigned g8 = swar_gt_const(__vabsdiffu4(W[-3 * rs], v), k7, t_low);
    const unsigned p08 = g0 | g8;
    if ((p08 & 0x80808080u) == 0) return 0;
    const unsigned g4 = swar_gt_const(__vabsdiffu4(__funnelshift_r(v, W[1], 24), v), k7, t_low);
    const unsigned g12 = swar_gt_const(__vabsdiffu4(__funnelshift_r(W[-1], v, 8), v), k7, t_low);
    return p08 & (g4 | g12) & 0x80808080u;
}

// B2: candidate flags (bit 7 per byte) of the 4 pixels whose centre word is W[0]; rs = row stride in words.
// Only the 8 even ring positions are examined, and without polarity: a 9-arc of the 16-ring always contains 4 consecutive
// even positions, all on one side of v, so "4 consecutive even positions all differ from v by more than t" is necessary for a
// corner.  One VABSDIFF4 + 3 ops per ring word; 13.5 % of the benchmark frame's pixels pass (a polarity-aware version of the
// same test passes 11.9 % at twice the ALU cost: measured 0.846 -> 0.820 ms per 256 frames in favour of this one).  The exact
// decision is left to the score stage (score >= t  <=>  corner at t), which runs on dense warps of single pixels.
__device__ __forceinline__ unsigned fast_even8(const unsigned* __restrict__ W, int rs, unsigned k7, bool t_low, unsigned valid)
{
    const unsigned v = W[0];
    unsigned f[8];
#define RINGA(k, word) f[k] = swar_gt_const(__vabsdiffu4((word), v), k7, t_low)
    RINGA(0, W[3 * rs]);
    RINGA(4, W[-3 * rs]);
    {
        const unsigned l = W[-1], r = W[1];
        RINGA(2, __funnelshift_r(v, r, 24));
        RINGA(6, __funnelshift_r(l, v, 8));
    }
    {
        const unsigned* p = W + 2 * rs; const unsigned* q = W - 2 * rs;
        RINGA(1, __funnelshift_r(p[0], p[1], 16));
        RINGA(7, __funnelshift_r(p[-1], p[0], 16));
        RINGA(3, __funnelshift_r(q[0], q[1], 16));
        RINGA(5, __funnelshift_r(q[-1], q[0], 16));
    }
#undef RINGA
    unsigned p2[8], out = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) p2[k] = f[k] & f[(k + 1) & 7];
#pragma unroll
    for (int k = 0; k < 8; k++) out |= p2[k] & p2[(k + 2) & 7];
    return out & valid & 0x80808080u;
}

__global__ void __launch_bounds__(256)
k_fast(const CUtensorMap* __restrict__ tmaps, const unsigned* __restrict__ tile_tab, unsigned* __restrict__ cand,
       int* __restrict__ cand_count, int* __restrict__ status, const __grid_constant__ Plan P)
{
    extern __shared__ __align__(128) unsigned char s_fast[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ unsigned s_colvalid[FAST_CH][64];
    __shared__ int s_surv[FAST_CW * FAST_CH];
    __shared__ unsigned s_rcpg;

    const unsigned te = __ldg(tile_tab + blockIdx.x);     // level | cy0 << 4 | cx0 << 16
    const int level = te & 15, cy0 = (te >> 4) & 0xFFF, cx0 = te >> 16;
    const LevelInfo& L = P.lv[level];
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncx = min(FAST_CW, L.ncols - cx0), ncy = min(FAST_CH, L.nrows - cy0);
    const int xend = L.w - EDGE, yend = L.h - EDGE;
    const int X0 = EDGE + cx0 * L.wcell, X1 = min(EDGE + (cx0 + ncx) * L.wcell, xend);
    const int Y0 = EDGE + cy0 * L.hcell, Y1 = min(EDGE + (cy0 + ncy) * L.hcell, yend);
    if (X0 >= X1 || Y0 >= Y1) return;
    const int irow = P.f_irow, srow = P.f_srow;           // shared-memory strides (bytes) from the plan
    unsigned* s_img = reinterpret_cast<unsigned*>(s_fast);
    uint8_t* s_score = s_fast + (size_t)irow * P.f_irows;
    // per-warp queues: every warp filters its own share of the tile through B1 -> B2 -> C without a CTA barrier
    unsigned short* gq = reinterpret_cast<unsigned short*>(s_score + (size_t)srow * P.f_srows) + warp * P.f_gw;   // groups, later corners
    unsigned short* pq = reinterpret_cast<unsigned short*>(s_score + (size_t)srow * P.f_srows) + FAST_WARPS * P.f_gw + warp * 4 * P.f_gw;   // candidate pixels

    const int gx0 = X0 & ~3, gxe = (X1 - 1) & ~3;         // first / last 4-pixel group (image coords)
    // at least 2 groups per row: with one, the reciprocal below would be 2^32 and wrap to 0 (a sliver tile whose detection range lies
    // inside one 4-pixel group: found by the 1026-px-wide parity case); the extra group's pixels are masked out by s_colvalid
    const int ngx = max(((gxe - gx0) >> 2) + 1, 2);
    const int ax0 = ((gx0 - 4 + EDGE) & ~15) - EDGE;      // image coords of s_img[0][0]: TMA needs a 16-byte aligned start
    const int ay0 = Y0 - 3;
    const int cofs = (gx0 - 4 - ax0) >> 2;                // word column of the first group's left neighbour
    const int rsw = irow >> 2;
    const int wc = L.wcell, hc = L.hcell;
    if (tid < FAST_CW * FAST_CH) s_surv[tid] = 0;
    static_assert(FAST_CW == 4 && FAST_CH == 2, "the cell arithmetic below is written for 4 x 2 cells");
    // A: stage the tile with one TMA box load (zero-filled outside the padded plane); the box is the plan's
    //    largest tile, so every CTA issues the same shape.  The issuing thread initialises the barrier itself; the other
    //    threads only touch it after the CTA barrier below.
    if (tid == 0) {
        mbar_init(&s_mbar, 1); mbar_fence_init();
        mbar_arrive_expect_tx(&s_mbar, (unsigned)(irow * P.f_irows));
        tma_load_3d(s_img, tmaps + level, ax0 + EDGE, ay0 + EDGE, f, &s_mbar);
    }
    if (tid == 32) s_rcpg = 0xFFFFFFFFu / (unsigned)ngx + 1u;     // one division per CTA, not per thread
    // byte masks per (cell row, group column): pixel inside [X0, X1) and its cell active in this pass (pass 0: every cell)
    auto col_masks = [&](int pass) {
        if (tid < FAST_CH * 64) {
            const int ci = tid >> 6, c = tid & 63;
            unsigned m = 0;
            if (c < ngx && ci < ncy)
                for (int bb = 0; bb < 4; bb++) {
                    const int x = gx0 + 4 * c + bb;
                    if (x >= X0 && x < X1) {
                        const int xr = x - X0, cj = (xr >= wc) + (xr >= 2 * wc) + (xr >= 3 * wc);
                        if (pass == 0 || s_surv[ci * FAST_CW + cj] == 0) m |= 0x80u << (8 * bb);
                    }
                }
            s_colvalid[ci][c] = m;
        }
    };
    col_masks(0);
    // score map: pixel (x, y) of cell (ci, cj) lives at row (y - Y0) + 1 + ci, column (x - X0) + 1 + cj, i.e. cells are
    // separated by one row / column that stays 0, so the cell-local NMS reads its 8 neighbours unconditionally
    // (128-bit stores; the last one may run a few bytes into the queues behind the map, which nobody has written yet)
    for (int i = tid; i < (srow * P.f_srows + 15) >> 4; i += 256) reinterpret_cast<uint4*>(s_score)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();                                       // s_surv, s_rcpg, masks, zeroed map, initialised barrier
    mbar_wait(&s_mbar, 0);
    const int nrows = Y1 - Y0, ngroups = ngx * nrows;
    const unsigned rcpg = s_rcpg;
    const unsigned rcps = P.f_rcp_srow;
    const unsigned ltmask = (1u << lane) - 1u;
    int* gcount = cand_count + (size_t)f * P.nlevels + level;
    unsigned* gdst = cand + (size_t)f * P.raw_per_frame + L.raw_off;
    const uint8_t* img8 = reinterpret_cast<const uint8_t*>(s_img) + 3 * irow + (X0 - ax0);    // pixel (X0, Y0)

    for (int pass = 0; pass < 2; pass++) {
        const int t = pass ? P.t2 : P.t1;
        if (pass) {
            if (P.t2 >= P.t1) break;                       // the retry cannot add anything (uniform)
            // every thread reads the 8 survivor flags itself (complete: barrier after stage D), so the decision needs no vote
            bool retry = false;
#pragma unroll
            for (int cell = 0; cell < FAST_CW * FAST_CH; cell++) retry |= (cell / FAST_CW) < ncy && (cell % FAST_CW) < ncx && s_surv[cell] == 0;
            if (!retry) break;                             // uniform
            col_masks(1);
            __syncthreads();
        }
        // B1: compass pre-test over this warp's groups, survivors compacted into gq
        const bool t_low = t < 128;
        const unsigned k7 = (unsigned)(0x7f - (t & 0x7f)) * 0x01010101u;
        int ngw = 0;
        for (int g0 = 0; g0 < ngroups; g0 += 256) {
            const int g = g0 + tid;
            unsigned pf = 0;
            if (g < ngroups) {
                const int r = __umulhi((unsigned)g, rcpg), c = g - r * ngx;
                pf = fast_pre4(s_img + (r + 3) * rsw + (cofs + c + 1), rsw, k7, t_low);
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, pf != 0);
            if (pf) gq[ngw + __popc(bal & ltmask)] = (unsigned short)g;
            ngw += __popc(bal);
        }
        __syncwarp();
        // B2: even-position candidate test on the compacted groups; candidate pixels go to pq as (y - Y0) * srow + (x - X0)
        int npw = 0;
        for (int i0 = 0; i0 < ngw; i0 += 32) {
            const int gi = i0 + lane;
            const int g = gi < ngw ? gq[gi] : 0;
            const int r = __umulhi((unsigned)g, rcpg), c = g - r * ngx;
            const unsigned valid = gi < ngw ? s_colvalid[(r >= hc) ? 1 : 0][c] : 0u;
            unsigned cf = valid ? fast_even8(s_img + (r + 3) * rsw + (cofs + c + 1), rsw, k7, t_low, valid) : 0u;
            // warp-aggregated append of the candidate pixels (<= 4 per lane), one ballot per byte position: no scan, no
            // data-dependent loop (the order inside pq does not matter)
            const int q0 = r * srow + (gx0 + 4 * c - X0);
            int total = 0;
#pragma unroll
            for (int bb = 0; bb < 4; bb++) {
                const bool on = (cf >> (8 * bb + 7)) & 1u;
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, on);
                if (on) pq[npw + total + __popc(bal & ltmask)] = (unsigned short)(q0 + bb);
                total += __popc(bal);
            }
            npw += total;
        }
        __syncwarp();
        // C: exact score s = max_arc min_k |ring_k - v| - 1 (== OpenCV cornerScore); the pixel is a corner at t  <=>
        //    s >= t.  Two queued pixels per lane: their ring values are packed as 16x2 and the arc min / max trees
        //    run on the DPX three-input SIMD min / max.  Corners are compacted into gq (score-map positions) while they
        //    fit; pq is rewritten in place (position or 0xFFFF) as the fallback list for a denser warp.
        int ncw = 0;
        for (int i0 = 0; i0 < npw; i0 += 64) {
            const int i = i0 + 2 * lane;
            const bool h0 = i < npw, h1 = i + 1 < npw;
            const int qa = h0 ? pq[i] : 0, qb = h1 ? pq[i + 1] : qa;
            const int rya = __umulhi((unsigned)qa, rcps), rxa = qa - rya * srow;
            const int ryb = __umulhi((unsigned)qb, rcps), rxb = qb - ryb * srow;
            const uint8_t* pa = img8 + rya * irow + rxa;
            const uint8_t* pb = img8 + ryb * irow + rxb;
            const int va = pa[0], vb = pb[0];
            unsigned w[16];
#define PK(k, o) w[k] = (unsigned)pa[o] | ((unsigned)pb[o] << 16)
            PK(0, 3 * irow);      PK(1, 3 * irow + 1);   PK(2, 2 * irow + 2);   PK(3, irow + 3);
            PK(4, 3);             PK(5, -irow + 3);      PK(6, -2 * irow + 2);  PK(7, -3 * irow + 1);
            PK(8, -3 * irow);     PK(9, -3 * irow - 1);  PK(10, -2 * irow - 2); PK(11, -irow - 3);
            PK(12, -3);           PK(13, irow - 3);      PK(14, 2 * irow - 2);  PK(15, 3 * irow - 1);
#undef PK
            unsigned A, Bm;
            {
                unsigned m3[16];
#pragma unroll
                for (int k = 0; k < 16; k++) m3[k] = __vimin3_u16x2(w[k], w[(k + 1) & 15], w[(k + 2) & 15]);
                unsigned m9[16];
#pragma unroll
                for (int k = 0; k < 16; k++) m9[k] = __vimin3_u16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
                A = __vimax3_u16x2(m9[0], m9[1], m9[2]);
#pragma unroll
                for (int k = 3; k < 15; k += 2) A = __vimax3_u16x2(A, m9[k], m9[k + 1]);
                A = __vmaxu2(A, m9[15]);
            }
            {
                unsigned m3[16];
#pragma unroll
                for (int k = 0; k < 16; k++) m3[k] = __vimax3_u16x2(w[k], w[(k + 1) & 15], w[(k + 2) & 15]);
                unsigned m9[16];
#pragma unroll
                for (int k = 0; k < 16; k++) m9[k] = __vimax3_u16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
                Bm = __vimin3_u16x2(m9[0], m9[1], m9[2]);
#pragma unroll
                for (int k = 3; k < 15; k += 2) Bm = __vimin3_u16x2(Bm, m9[k], m9[k + 1]);
                Bm = __vminu2(Bm, m9[15]);
            }
            const int sca = max((int)(A & 0xFFFFu) - va, va - (int)(Bm & 0xFFFFu)) - 1;
            const int scb = max((int)(A >> 16) - vb, vb - (int)(Bm >> 16)) - 1;
            const bool ca = h0 && sca >= t, cb = h1 && scb >= t;
            // score-map positions (cells separated by the zero row / column)
            const int posa = qa + (1 + (rya >= hc)) * srow + 1 + (rxa >= wc) + (rxa >= 2 * wc) + (rxa >= 3 * wc);
            const int posb = qb + (1 + (ryb >= hc)) * srow + 1 + (rxb >= wc) + (rxb >= 2 * wc) + (rxb >= 3 * wc);
            if (h0) { s_score[posa] = (uint8_t)(ca ? sca : 0); pq[i] = (unsigned short)(ca ? posa : 0xFFFF); }
            if (h1) { s_score[posb] = (uint8_t)(cb ? scb : 0); pq[i + 1] = (unsigned short)(cb ? posb : 0xFFFF); }
            const unsigned bala = __ballot_sync(0xFFFFFFFFu, ca), balb = __ballot_sync(0xFFFFFFFFu, cb);
            const int ia = ncw + __popc(bala & ltmask), ib = ncw + __popc(bala) + __popc(balb & ltmask);
            if (ca && ia < P.f_gw) gq[ia] = (unsigned short)posa;
            if (cb && ib < P.f_gw) gq[ib] = (unsigned short)posb;
            ncw += __popc(bala) + __popc(balb);
        }
        __syncthreads();                                   // every warp's scores are in the map
        // D: strict 3x3 NMS inside the cell; append survivors to the (frame, level) raw-corner list
        //    Two sweeps over the warp's corner list so that the warp makes ONE reservation in the global list (a global atomic
        //    per 32 corners kept the warps waiting on its latency): sweep 1 decides NMS and strikes the losers out of the list
        //    in place, sweep 2 writes the survivors.
        const bool fits = ncw <= P.f_gw;                   // warp-uniform
        unsigned short* dq = fits ? gq : pq;
        const int nd = fits ? ncw : npw;
        int nkeep = 0;
        for (int i0 = 0; i0 < nd; i0 += 32) {
            const int i = i0 + lane;
            const int pos = i < nd ? dq[i] : 0xFFFF;
            bool keep = false;
            if (pos != 0xFFFF) {
                const uint8_t* sp = s_score + pos;
                const int s = sp[0];
                const int m = max(max(max((int)sp[-1], (int)sp[1]), max((int)sp[-srow - 1], (int)sp[-srow])),
                                  max(max((int)sp[-srow + 1], (int)sp[srow - 1]), max((int)sp[srow], (int)sp[srow + 1])));
                keep = s > m;                              // strict maximum of its 8 neighbours, no short-circuit branches
                if (!keep) dq[i] = 0xFFFF;
            }
            nkeep += __popc(__ballot_sync(0xFFFFFFFFu, keep));
        }
        if (nkeep) {                                       // warp-uniform
            int base = 0;
            if (lane == 0) base = atomicAdd(gcount, nkeep);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            __syncwarp();
            for (int i0 = 0; i0 < nd; i0 += 32) {
                const int i = i0 + lane;
                const int pos = i < nd ? dq[i] : 0xFFFF;
                const bool keep = pos != 0xFFFF;
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
                if (keep) {
                    const int sy = __umulhi((unsigned)pos, rcps), sx = pos - sy * srow;
                    const int ci = sy > hc + 1, cj = (sx > wc + 1) + (sx > 2 * wc + 2) + (sx > 3 * wc + 3);
                    const unsigned rec = (unsigned)(X0 + sx - 1 - cj) | ((unsigned)(Y0 + sy - 1 - ci) << 12) | ((unsigned)s_score[pos] << 24);
                    const int o = base + __popc(bal & ltmask);
                    if (o < L.raw_cap) gdst[o] = rec; else atomicOr(status, 1);
                    s_surv[ci * FAST_CW + cj] = 1;         // only "any survivor" matters: plain store, every writer stores 1
                }
                base += __popc(bal);
            }
        }
        __syncthreads();                                   // s_surv complete before the retry pass reads it
    }
}

// --------------------------------------------------------------------------------------------------------
// K2+K3, second design: the same computation (cv::FAST(th, nms=true) per 30-px cell ROI + the retry at the second threshold when a
// cell returned nothing, src/ORBextractor.cc:772-812) organised around what the first kernel's profile showed: it was bound by the ALU
// pipe (70 % busy at 74 % issue) and lost the rest to CTA barriers (1.4 stalled warps per issue), to a 300-instruction prologue per
// 7.7 k-pixel CTA (11 % of all instructions) and to cell bookkeeping inside a 4 x 2-cell tile (which cell does this pixel belong to).
//   * ONE WARP PER CELL.  NMS, the survivor count and the retry decision are cell-local, so a warp that owns a cell needs no CTA
//     barrier at all: its queues, its score map and its retry loop are private.  Cell geometry is a handful of warp-uniform values.
//   * PERSISTENT CTAs over the (tile, frame) space with a two-stage TMA pipeline: the box of tile i + 2 is requested by the warp that
//     finishes the last cell of tile i; the eight warps meet only at the mbarrier of each tile's box.  Cell <-> warp assignment
//     rotates from tile to tile, so edge tiles with fewer than eight cells do not idle the same warps.
//   * ALU diet of the two filter stages: |ring - v| > t is tested as bit 7 of (a + K) | a with K = (127 - t) * 0x01010101 — without
//     the & 0x7f7f7f7f of the exact SWAR compare.  A byte carry can only ADD a false positive (a == t beside a neighbour >= 129 + t),
//     which is all a filter needs: the exact decision is the score stage's (score >= t  <=>  FAST-9 corner at t).  No early-out
//     branch inside the compass test (it diverged on every warp), cell-edge masks from an 8-word table instead of compares.
//   * queue entries are (row << 8 | column) relative to the cell: image offset, score-map position and output coordinates are one
//     IMAD each; no divisions anywhere behind the first stage.
// Stages per cell: B1 compass pre-test on 4-pixel groups -> B2 8 even ring positions (4 consecutive ones must differ) -> C exact
// score, two pixels per lane packed 16x2 on VIMNMX3 -> D strict 3x3 NMS inside the cell, one global reservation per warp and cell.
// --------------------------------------------------------------------------------------------------------
constexpr int F2_IROW = 160;                 // row stride of the staged tile in the common case (4 cells of <= 33 px + halo + alignment)
#ifndef F2_PQ_
#define F2_PQ_ 1024
#endif
constexpr int F2_PQ = F2_PQ_;                  // candidate-pixel queue of a warp (flushed through the score stage when nearly full)
__constant__ unsigned c_rcp32[33];           // ceil(2^32 / n), n = 1..32 (groups per cell row)

constexpr int F2_CW = 4;                     // cells per tile: one row of four (a buffer is recycled as soon as its four cells are done)
#ifndef F2_STAGES_
#define F2_STAGES_ 3
#endif
constexpr int F2_STAGES = F2_STAGES_, F2_RING = 8;    // staged tiles per CTA; ring of tile announcements (> F2_STAGES, power of two)

#ifndef F2_MINB
#define F2_MINB 4
#endif
#ifndef F2_PARK_NS
#define F2_PARK_NS 2000
#endif
template <int IROWT>                         // row stride of the staged tile in bytes; 0 = take it from the plan at run time
__global__ void __launch_bounds__(256, F2_MINB)
k_fast2(const CUtensorMap* __restrict__ tmaps, const unsigned* __restrict__ tile_tab, unsigned* __restrict__ cand,
        int* __restrict__ cand_count, int* __restrict__ status, int nframes, const __grid_constant__ Plan P)
{
    extern __shared__ __align__(128) unsigned char s_f2[];
    __shared__ __align__(8) uint64_t s_full[F2_STAGES], s_ann[F2_RING];
    __shared__ int s_loads[F2_STAGES], s_next, s_lock, s_issued;
    __shared__ int s_woff[FAST_WARPS];                    // offset of each warp's private area (read back per cell: one LDS instead of the
                                                          // special-register and constant-bank arithmetic the compiler re-materialises)
    // announcement of the CTA's tile number q (slot q & 7): which launch tile it is (-1: the launch has no more), where it lands and the
    // parity of that buffer's mbarrier phase.  The issuing thread writes the fields and arrives on s_ann[slot] (release); the four warps
    // that take the tile's cells wait on it (acquire) with the parity of the slot's use number q / F2_RING.  s_rdone counts the cells
    // of the slot's tile that are done: at F2_CW the tile's buffer is requested again, and the slot may be announced again.
    __shared__ int s_rdone[F2_RING], s_rtile[F2_RING], s_rbuf[F2_RING];
    __shared__ unsigned s_rte[F2_RING];                   // the tile's table entry (level | cell row << 4 | first cell column << 16)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int irow = IROWT ? IROWT : P.f_irow;
    const int imgbytes = P.f2_imgbytes, imgstride = P.f2_imgstride;
    const int T = P.f2_tiles * nframes;
    int* tile_ctr = status + 2;                            // [2] next (tile, frame) of the launch, [3] CTAs that have left; both 0 between launches
    // Request the CTA's next tile into buffer `buf` (free: every cell of its previous tile is done).  Tiles come from a launch-wide
    // counter, so CTAs that meet dense tiles simply draw fewer; buffers are recycled in COMPLETION order, so a slow cell holds back its
    // own buffer only.  One thread at a time (lock): tile numbers and launch tiles stay in the same order, which the hand-out of cells
    // in tile order and the end-of-launch test rely on.
    auto issue = [&](int buf) {
        while (atomicCAS(&s_lock, 0, 1) != 0) __nanosleep(20);
        const int q = atomicAdd(&s_issued, 1);
        const int t = atomicAdd(tile_ctr, 1);
        atomicExch(&s_lock, 0);
        const int slot = q & (F2_RING - 1);
        // the slot's previous tile (q - F2_RING) must be complete, i.e. its four warps have read the announcement.  With three buffers
        // a tile that far behind is still running only if one cell outlasts seven whole tiles: the wait is there for the proof.
        while (atomicCAS(&s_rdone[slot], F2_CW, 0) != F2_CW) __nanosleep(20);
        if (t >= T) { s_rtile[slot] = -1; mbar_arrive(&s_ann[slot]); return; }
        const int f = P.f2_tiles > 1 ? (int)__umulhi((unsigned)t, P.f2_rcp_tiles) : t, tl = t - f * P.f2_tiles;
        const unsigned te = __ldg(tile_tab + tl);
        const LevelInfo& L = P.lv[te & 15];
        const int TX0 = EDGE + (int)(te >> 16) * L.wcell, TY0 = EDGE + (int)((te >> 4) & 0xFFF) * L.hcell;
        const int ax0 = ((((TX0 & ~3) - 4 + EDGE) & ~15) - EDGE), ay0 = TY0 - 3;
        const int par = atomicAdd(&s_loads[buf], 1) & 1;
        s_rtile[slot] = f; s_rte[slot] = te; s_rbuf[slot] = buf | (par << 8);
        mbar_arrive(&s_ann[slot]);
        mbar_arrive_expect_tx(&s_full[buf], (unsigned)imgbytes);
        tma_load_3d(s_f2 + (size_t)buf * imgstride, tmaps + 3 * MAXLEV + (te & 15), ax0 + EDGE, ay0 + EDGE, f, &s_full[buf]);
    };
    if (lane == 0) s_woff[warp] = F2_STAGES * imgstride + warp * P.f2_wbytes;
    if (tid == 0) {
        for (int k = 0; k < F2_STAGES; k++) { mbar_init(&s_full[k], 1); s_loads[k] = 0; }
        for (int k = 0; k < F2_RING; k++) { mbar_init(&s_ann[k], 1); s_rdone[k] = F2_CW; }
        mbar_fence_init();
        s_next = 0; s_lock = 0; s_issued = 0;
        for (int k = 0; k < F2_STAGES; k++) issue(k);
    }
    __syncthreads();                                       // the only CTA barrier: the mbarriers and the bookkeeping words exist

    // the warp's private area
    unsigned char* wb = s_f2 + *reinterpret_cast<volatile int*>(&s_woff[warp]);
    uint8_t* s_score = wb;
    unsigned short* gq = reinterpret_cast<unsigned short*>(wb + P.f2_score_bytes);
    unsigned short* pq = gq + P.f2_gcap;
    // The corner list is the pixel queue compacted IN PLACE by the score stage (corner k of a round lands at pq[k], k <= the entries the
    // round has read so far), so it needs no storage of its own.  That only works while one round covers the cell: a cell with more
    // than F2_PQ candidate pixels (noise images) is scored in several rounds and takes the score-map sweep of stage D instead of a list.
    unsigned short* cq = pq;
    unsigned* vtab = reinterpret_cast<unsigned*>(pq + F2_PQ + 2);
    const int srow = P.f2_srow;
    const unsigned ltmask = (1u << lane) - 1u;
    const int rsw = irow >> 2;

    // Cells are handed out in order (cell n = cell n & 3 of the CTA's tile n >> 2): a warp that finishes early takes the next cell, of
    // whichever staged tile comes next, instead of waiting for its tile's slowest cell.
    for (;;) {
        int n = 0;
        if (lane == 0) n = atom_add_shared(&s_next, 1);         // (plain PTX: atomicAdd() on shared memory expands to the warp-aggregation idiom)
        n = __shfl_sync(0xFFFFFFFFu, n, 0);
        const int q = n >> 2, cj = n & 3, slot = q & (F2_RING - 1);
        static_assert(F2_CW == 4, "four cells per tile");
        static_assert((F2_RING & (F2_RING - 1)) == 0 && F2_RING > F2_STAGES + 2, "ring of announcements");
        mbar_wait_parked(&s_ann[slot], (unsigned)(q / F2_RING) & 1u, F2_PARK_NS);   // until the tile is announced (its buffer may still be busy)
        const int f = s_rtile[slot];                       // the tile's frame
        if (f < 0) break;                                  // the launch has no more tiles (cells are handed out in tile order)
        const int bp = s_rbuf[slot], b = bp & 255;
        const unsigned te = s_rte[slot];
        const int level = te & 15, tcy0 = (te >> 4) & 0xFFF, tcx0 = te >> 16;
        const LevelInfo& L = P.lv[level];
        const int wc = L.wcell, hc = L.hcell;
        const int ci = 0;
        const int TX0 = EDGE + tcx0 * wc, TY0 = EDGE + tcy0 * hc;
        const int ax0 = ((((TX0 & ~3) - 4 + EDGE) & ~15) - EDGE), ay0 = TY0 - 3;
        const int cx0 = TX0 + cj * wc, cx1 = min(cx0 + wc, L.w - EDGE);
        const int cy0 = TY0 + ci * hc, cy1 = min(cy0 + hc, L.h - EDGE);
        // the announcement names the phase: a buffer is only requested again once all four cells of its tile are done, this one
        // included, so the parity cannot refer to a phase two steps away
        mbar_wait(&s_full[b], (unsigned)(bp >> 8));
        if (tcx0 + cj < L.ncols && tcy0 + ci < L.nrows && cx0 < cx1 && cy0 < cy1) {
            const unsigned char* img = s_f2 + (size_t)b * imgstride;
            const int gx0 = cx0 & ~3;
            const int ngx_real = ((((cx1 - 1) & ~3) - gx0) >> 2) + 1;      // 4-pixel groups per row (<= 17 for cells <= 64 px)
            const int ngx = max(ngx_real, 2);                              // c_rcp32[1] would be 2^32: a one-group sliver gets a masked second group
            const int nrows = cy1 - cy0, ng = ngx * nrows;
            const unsigned rcp = c_rcp32[ngx];
            const unsigned* Wc = reinterpret_cast<const unsigned*>(img + (cy0 - ay0) * irow + (gx0 - ax0));   // word of group (0, 0)
            const int rskip = rsw - ngx;
            const uint8_t* cimg = img + (cy0 - 1 - ay0) * irow + (cx0 - 1 - ax0);   // pixel (xrel, yrel) = (0, 0): entries count from 1
            const int xoff0 = gx0 - cx0 + 1;
            // cell-edge masks per group column (bit 7 of the bytes inside [cx0, cx1)), and a clean score map
            if (lane < ngx) {
                unsigned m = lane < ngx_real ? 0x80808080u : 0u;
                if (lane == 0) m &= 0x80808080u << (8 * (cx0 & 3));
                if (lane == ngx_real - 1) m &= 0x80808080u >> (8 * (3 - ((cx1 - 1) & 3)));
                vtab[lane] = m;
            }
            {   // the whole map of the largest cell (75 x 16 bytes for 30-px cells): three predicated stores, no loop in the common case
                const int n16 = P.f2_score_bytes >> 4;
#pragma unroll
                for (int it = 0; it < 3; it++) if (lane + 32 * it < n16) reinterpret_cast<uint4*>(s_score)[lane + 32 * it] = make_uint4(0u, 0u, 0u, 0u);
                for (int k = lane + 96; k < n16; k += 32) reinterpret_cast<uint4*>(s_score)[k] = make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
            int* gcount = cand_count + (size_t)f * P.nlevels + level;
            unsigned* gdst = cand + (size_t)f * P.raw_per_frame + L.raw_off;
            const unsigned recbase = (unsigned)(cx0 - 1) + ((unsigned)(cy0 - 1) << 12);

            for (int pass = 0; pass < 2; pass++) {
                const int t_exact = pass ? P.t2 : P.t1;
                if (pass && P.t2 >= P.t1) break;                           // the retry cannot add anything
                const unsigned K = (unsigned)(127 - min(t_exact, 127)) * 0x01010101u;
                // ---- B1: compass pre-test on every group of the cell
                int ngw = 0;
                for (int g0 = 0; g0 < ng; g0 += 32) {
                    // (lanes behind the last group repeat it and drop the result: no branch around the loads)
                    const int g = g0 + lane, gc = min(g, ng - 1);
                    unsigned pf;
                    {
                        const int r = __umulhi((unsigned)gc, rcp);
                        const unsigned* W = Wc + gc + r * rskip;
                        const unsigned v = W[0];
                        const unsigned a0 = __vabsdiffu4(W[3 * rsw], v), a8 = __vabsdiffu4(W[-3 * rsw], v);
                        const unsigned a4 = __vabsdiffu4(__funnelshift_r(v, W[1], 24), v), a12 = __vabsdiffu4(__funnelshift_r(W[-1], v, 8), v);
                        pf = ((a0 + K) | a0 | (a8 + K) | a8) & ((a4 + K) | a4 | (a12 + K) | a12) & 0x80808080u;
                    }
                    if (g >= ng) pf = 0;
                    const unsigned bal = __ballot_sync(0xFFFFFFFFu, pf != 0);
                    if (pf) gq[ngw + __popc(bal & ltmask)] = (unsigned short)g;
                    ngw += __popc(bal);
                }
                __syncwarp();
                // ---- B2 -> C, in rounds: B2 fills the pixel queue from the group queue until it is nearly full, C empties it
                int ncw = 0, i0 = 0;
                bool listok = true;
                while (i0 < ngw) {
                    int npw = 0;
                    if (i0) { listok = false; ncw = 0; }                   // a second round: the first round's list is being overwritten
                    for (; i0 < ngw && npw + 128 <= F2_PQ; i0 += 32) {
                        const int gi = i0 + lane;
                        unsigned cf; int e0;
                        {   // (lanes behind the last queued group repeat it and drop the result)
                            const int g = gq[min(gi, ngw - 1)];
                            const int r = __umulhi((unsigned)g, rcp), c = g - r * ngx;
                            const unsigned* W = Wc + g + r * rskip;
                            const unsigned v = W[0];
                            unsigned fl[8];
#define RINGF(k, word) { const unsigned a_ = __vabsdiffu4((word), v); fl[k] = (a_ + K) | a_; }
                            RINGF(0, W[3 * rsw]);
                            RINGF(4, W[-3 * rsw]);
                            { const unsigned l = W[-1], rr = W[1]; RINGF(2, __funnelshift_r(v, rr, 24)); RINGF(6, __funnelshift_r(l, v, 8)); }
                            { const unsigned* p = W + 2 * rsw; const unsigned* q = W - 2 * rsw;
                              RINGF(1, __funnelshift_r(p[0], p[1], 16)); RINGF(7, __funnelshift_r(p[-1], p[0], 16));
                              RINGF(3, __funnelshift_r(q[0], q[1], 16)); RINGF(5, __funnelshift_r(q[-1], q[0], 16)); }
#undef RINGF
                            unsigned out = 0;
#pragma unroll
                            for (int k = 0; k < 8; k++) out |= (fl[k] & fl[(k + 1) & 7] & fl[(k + 2) & 7]) & fl[(k + 3) & 7];
                            cf = gi < ngw ? out & vtab[c] : 0u;
                            e0 = ((r + 1) << 8) + xoff0 + 4 * c;
                        }
                        int total = 0;
#pragma unroll
                        for (int bb = 0; bb < 4; bb++) {
                            const bool on = (cf >> (8 * bb + 7)) & 1u;
                            const unsigned bal = __ballot_sync(0xFFFFFFFFu, on);
                            if (on) pq[npw + total + __popc(bal & ltmask)] = (unsigned short)(e0 + bb);
                            total += __popc(bal);
                        }
                        npw += total;
                    }
                    __syncwarp();
                    // ---- C: exact score of the queued pixels, two per lane packed 16x2
                    const unsigned efirst = npw ? pq[0] : 0x0101u;
                    for (int j0 = 0; j0 < npw; j0 += 64) {
                        const int j = j0 + 2 * lane;
                        const bool h0 = j < npw, h1 = j + 1 < npw;
                        const unsigned ee = *reinterpret_cast<const unsigned*>(pq + j);
                        const unsigned ea = h0 ? (ee & 0xFFFFu) : efirst, eb = h1 ? (ee >> 16) : ea;
                        const int ya = ea >> 8, xa = ea & 255, yb = eb >> 8, xb = eb & 255;
                        const uint8_t* pa = cimg + ya * irow + xa;
                        const uint8_t* pb = cimg + yb * irow + xb;
                        const int va = pa[0], vb = pb[0];
                        unsigned w[16];
#define PK(k, o) w[k] = (unsigned)pa[o] | ((unsigned)pb[o] << 16)
                        PK(0, 3 * irow);      PK(1, 3 * irow + 1);   PK(2, 2 * irow + 2);   PK(3, irow + 3);
                        PK(4, 3);             PK(5, -irow + 3);      PK(6, -2 * irow + 2);  PK(7, -3 * irow + 1);
                        PK(8, -3 * irow);     PK(9, -3 * irow - 1);  PK(10, -2 * irow - 2); PK(11, -irow - 3);
                        PK(12, -3);           PK(13, irow - 3);      PK(14, 2 * irow - 2);  PK(15, 3 * irow - 1);
#undef PK
                        unsigned A, Bm;
                        {
                            unsigned m3[16], m9[16];
#pragma unroll
                            for (int k = 0; k < 16; k++) m3[k] = __vimin3_u16x2(w[k], w[(k + 1) & 15], w[(k + 2) & 15]);
#pragma unroll
                            for (int k = 0; k < 16; k++) m9[k] = __vimin3_u16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
                            A = __vimax3_u16x2(m9[0], m9[1], m9[2]);
#pragma unroll
                            for (int k = 3; k < 15; k += 2) A = __vimax3_u16x2(A, m9[k], m9[k + 1]);
                            A = __vmaxu2(A, m9[15]);
                        }
                        {
                            unsigned m3[16], m9[16];
#pragma unroll
                            for (int k = 0; k < 16; k++) m3[k] = __vimax3_u16x2(w[k], w[(k + 1) & 15], w[(k + 2) & 15]);
#pragma unroll
                            for (int k = 0; k < 16; k++) m9[k] = __vimax3_u16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
                            Bm = __vimin3_u16x2(m9[0], m9[1], m9[2]);
#pragma unroll
                            for (int k = 3; k < 15; k += 2) Bm = __vimin3_u16x2(Bm, m9[k], m9[k + 1]);
                            Bm = __vminu2(Bm, m9[15]);
                        }
                        const int sca = max((int)(A & 0xFFFFu) - va, va - (int)(Bm & 0xFFFFu)) - 1;
                        const int scb = max((int)(A >> 16) - vb, vb - (int)(Bm >> 16)) - 1;
                        const bool ca = h0 && sca >= t_exact, cb = h1 && scb >= t_exact;
                        if (ca) s_score[ya * srow + xa] = (uint8_t)sca;
                        if (cb) s_score[yb * srow + xb] = (uint8_t)scb;
                        const unsigned bala = __ballot_sync(0xFFFFFFFFu, ca), balb = __ballot_sync(0xFFFFFFFFu, cb);
                        const int ia = ncw + __popc(bala & ltmask), ib = ncw + __popc(bala) + __popc(balb & ltmask);
                        __syncwarp();                      // every lane has read its two entries of this step: slots below j0 + 64 are free
                        if (ca) cq[ia] = (unsigned short)ea;
                        if (cb) cq[ib] = (unsigned short)eb;
                        ncw += __popc(bala) + __popc(balb);
                    }
                    __syncwarp();
                }
                // ---- D: strict 3x3 NMS inside the cell (the map's frame of zeros stands for "outside the ROI interior"), survivors to
                //      the (frame, level) list: sweep 1 decides and counts, ONE reservation per warp and cell, sweep 2 writes
                int nkeep = 0;
                if (listok) {
                    // sweep 1: decide, and compact the survivors to the front of the list (a survivor moves to a slot at or below its own;
                    // the barrier separates the step's reads from its writes)
                    for (int k0 = 0; k0 < ncw; k0 += 32) {
                        const int k = k0 + lane;
                        const unsigned e = k < ncw ? cq[k] : 0u;
                        bool keep = false;
                        if (e) {
                            const uint8_t* sp = s_score + (e >> 8) * srow + (e & 255);
                            const int sc = sp[0];
                            const int m = max(max(max((int)sp[-1], (int)sp[1]), max((int)sp[-srow - 1], (int)sp[-srow])),
                                              max(max((int)sp[-srow + 1], (int)sp[srow - 1]), max((int)sp[srow], (int)sp[srow + 1])));
                            keep = sc > m;
                        }
                        const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
                        __syncwarp();
                        if (keep) cq[nkeep + __popc(bal & ltmask)] = (unsigned short)e;
                        nkeep += __popc(bal);
                    }
                    if (nkeep) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(gcount, nkeep);
                        base = __shfl_sync(0xFFFFFFFFu, base, 0);
                        __syncwarp();
                        // sweep 2: the survivors, in list order
                        for (int k = lane; k < nkeep; k += 32) {
                            const unsigned e = cq[k];
                            const unsigned rec = recbase + (e & 255u) + ((e >> 8) << 12) + ((unsigned)s_score[(e >> 8) * srow + (e & 255)] << 24);
                            const int o = base + k;
                            if (o < L.raw_cap) gdst[o] = rec; else atomicOr(status, 1);
                        }
                    }
                } else {
                    // a cell with more corners than the list holds (noise images): sweep its score map instead, twice
                    const int ncols = cx1 - cx0;
                    int base = 0;
                    for (int sweep = 0; sweep < 2; sweep++) {
                        if (sweep) {
                            if (!nkeep) break;
                            if (lane == 0) base = atomicAdd(gcount, nkeep);
                            base = __shfl_sync(0xFFFFFFFFu, base, 0);
                        }
                        for (int yr = 1; yr <= nrows; yr++)
                            for (int x0 = 1; x0 <= ncols; x0 += 32) {
                                const int xr = x0 + lane;
                                bool keep = false; int sc = 0;
                                if (xr <= ncols) {
                                    const uint8_t* sp = s_score + yr * srow + xr;
                                    sc = sp[0];
                                    const int m = max(max(max((int)sp[-1], (int)sp[1]), max((int)sp[-srow - 1], (int)sp[-srow])),
                                                      max(max((int)sp[-srow + 1], (int)sp[srow - 1]), max((int)sp[srow], (int)sp[srow + 1])));
                                    keep = sc > m;
                                }
                                const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
                                if (!sweep) nkeep += __popc(bal);
                                else {
                                    if (keep) {
                                        const int o = base + __popc(bal & ltmask);
                                        if (o < L.raw_cap) gdst[o] = recbase + (unsigned)xr + ((unsigned)yr << 12) + ((unsigned)sc << 24); else atomicOr(status, 1);
                                    }
                                    base += __popc(bal);
                                }
                            }
                    }
                }
                __syncwarp();
                if (nkeep) break;                                          // the cell returned keypoints: no retry (warp-uniform)
            }
        }
        // this cell is done with its tile's box; the warp that completes the tile requests the CTA's next tile into the freed buffer
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atom_add_shared(&s_rdone[slot], 1) == F2_CW - 1) {
                fence_proxy_async_smem();                  // generic-proxy reads of the buffer before the async-proxy overwrite
                issue(b);
            }
        }
    }
    // the last CTA to leave puts the launch-wide counters back to zero (every CTA has stopped drawing by then)
    __syncthreads();
    if (tid == 0 && atomicAdd(tile_ctr + 1, 1) == (int)gridDim.x - 1) { tile_ctr[0] = 0; tile_ctr[1] = 0; }
}

// --------------------------------------------------------------------------------------------------------
// K4: DistributeOctTree (src/ORBextractor.cc:1006-1287) — one CTA per (frame, level).
// Array formulation (tools/quadtree_proto.py): nodes live in LIST ORDER in shared memory (slot == position in the
// reference's std::list), keys stay where they are and carry a node label.  A round divides a ranked set of nodes
// and rebuilds the list as  reverse(children in creation order) ++ surviving nodes in old order — exactly what
// push_front + erase do.  Near the quota the reference expands the largest nodes first and stops at the first
// moment |list| >= N (:1140-1204); here every candidate is divided speculatively in parallel and a prefix sum over
// the sorted order finds that stopping point.  The sort tie on node POINTER (:1151) is pinned to creation order
// (later-created = nearer the list front = first), same as the oracle.
// --------------------------------------------------------------------------------------------------------
#ifndef QT_THREADS_
#define QT_THREADS_ 256
#endif
constexpr int QT_THREADS = QT_THREADS_;

struct QtShared {          // laid out in dynamic shared memory, all arrays node_cap long unless noted
    int* box[2];           // 4 ints per node: x0,y0,x1,y1
    int* cnt[2];
    int* cc;               // 4 per node: child key counts
    int* rank;             // processing rank of a slot, -1 = not divided
    int* slot_of_rank;
    int* scan;             // node_cap + 1
    int* newpos;
    int* childpos;         // 4 per node
    unsigned* sortk;       // node_cap
    int* best_score;
    unsigned* best_okey;
};

__device__ int block_excl_scan(int* a, int n, int* s_warp)     // in-place exclusive scan of a[0..n), returns total (all threads)
{
    const int tid = threadIdx.x;
    const int per = (n + QT_THREADS - 1) / QT_THREADS;
    const int b = tid * per, e = min(b + per, n);
    int sum = 0;
    for (int i = b; i < e; i++) sum += a[i];
    // warp scan of sums
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((tid & 31) >= o) incl += v; }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < QT_THREADS / 32; w++) { const int v = s_warp[w]; if (w < (tid >> 5)) woff += v; total += v; }
    int run = woff + incl - sum;
    for (int i = b; i < e; i++) { const int v = a[i]; a[i] = run; run += v; }
    __syncthreads();
    return total;
}

__device__ __forceinline__ void child_box(const int* pb, int q, int* cb)
{
    const int x0 = pb[0], y0 = pb[1], x1 = pb[2], y1 = pb[3];
    const int hx = (x1 - x0 + 1) >> 1, hy = (y1 - y0 + 1) >> 1;     // ceil(float(d)/2), d >= 0
    cb[0] = (q & 1) ? x0 + hx : x0;  cb[2] = (q & 1) ? x1 : x0 + hx;
    cb[1] = (q & 2) ? y0 + hy : y0;  cb[3] = (q & 2) ? y1 : y0 + hy;
}
__device__ __forceinline__ int quadrant(const int* pb, int x, int y)
{
    const int hx = (pb[2] - pb[0] + 1) >> 1, hy = (pb[3] - pb[1] + 1) >> 1;
    return (x >= pb[0] + hx ? 1 : 0) | (y >= pb[1] + hy ? 2 : 0);   // 0:n1(UL) 1:n2(UR) 2:n3(BL) 3:n4(BR)
}

__global__ void __launch_bounds__(QT_THREADS)
k_quadtree(const unsigned* __restrict__ cand, const int* __restrict__ cand_count, unsigned short* __restrict__ labels,
           unsigned* __restrict__ winners, int* __restrict__ win_count, int* __restrict__ status,
           const __grid_constant__ Plan P)
{
    extern __shared__ int s_dyn[];
    __shared__ int s_warp[QT_THREADS / 32];
    __shared__ int s_nexp, s_ncand, s_ndiv;

    const int level = blockIdx.y, f = blockIdx.x;      // level-major dispatch: the large levels of every frame start first, the small ones fill the tail
    const LevelInfo& L = P.lv[level];
    const int cap = P.node_cap;
    const int tid = threadIdx.x;
    QtShared S;
    {
        int* p = s_dyn;
        S.box[0] = p; p += 4 * cap; S.box[1] = p; p += 4 * cap;
        S.cnt[0] = p; p += cap; S.cnt[1] = p; p += cap;
        S.cc = p; p += 4 * cap; S.rank = p; p += cap; S.slot_of_rank = p; p += cap;
        S.scan = p; p += cap + 1; S.newpos = p; p += cap; S.childpos = p; p += 4 * cap;
        S.sortk = reinterpret_cast<unsigned*>(p); p += cap;
        S.best_score = p; p += cap; S.best_okey = reinterpret_cast<unsigned*>(p); p += cap;
    }
    const unsigned* kk = cand + (size_t)f * P.raw_per_frame + L.raw_off;     // raw corners of this (frame, level), any order
    unsigned short* lab = labels + (size_t)f * P.raw_per_frame + L.raw_off;
    const int n = min(cand_count[(size_t)f * P.nlevels + level], L.raw_cap);
    const int N = L.quota;
    unsigned* wout = winners + (size_t)f * P.kp_per_frame + L.kp_off;
    if (n == 0) { if (tid == 0) win_count[(size_t)f * P.nlevels + level] = 0; return; }

    // ---- b. roots (:1010-1053).  window coords = level coords - 13
    const int minB = EDGE - 3;
    const int nini = L.nini;
    const float hX = L.hx;
    for (int i = tid; i < nini; i += QT_THREADS) {
        int* b = S.box[0] + 4 * i;
        b[0] = (int)__fmul_rn(hX, (float)i); b[1] = 0; b[2] = (int)__fmul_rn(hX, (float)(i + 1)); b[3] = L.h - 2 * minB;
        S.cnt[0][i] = 0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += QT_THREADS) {
        const int x = (int)(kk[i] & 0xFFF) - minB;
        int r = (int)__fdiv_rn((float)x, hX);
        r = min(max(r, 0), nini - 1);
        lab[i] = (unsigned short)r;
        atomicAdd(&S.cnt[0][r], 1);
    }
    __syncthreads();
    // drop empty roots (keeps order)
    for (int i = tid; i < nini; i += QT_THREADS) S.scan[i] = S.cnt[0][i] > 0 ? 1 : 0;
    __syncthreads();
    int size = block_excl_scan(S.scan, nini, s_warp);
    for (int i = tid; i < nini; i += QT_THREADS)
        if (S.cnt[0][i] > 0) {
            const int p = S.scan[i];
            S.newpos[i] = p;
            for (int k = 0; k < 4; k++) S.box[1][4 * p + k] = S.box[0][4 * i + k];
            S.cnt[1][p] = S.cnt[0][i];
        }
    __syncthreads();
    for (int i = tid; i < n; i += QT_THREADS) lab[i] = (unsigned short)S.newpos[lab[i]];
    int cur = 1;                      // current node buffer
    int cprev = size;
    bool phase2 = false, finish = false;
    __syncthreads();

    // ---- c. rounds
    while (!finish) {
        const int prev_size = size;
        int* box = S.box[cur]; int* cnt = S.cnt[cur];
        int* nbox = S.box[cur ^ 1]; int* ncnt = S.cnt[cur ^ 1];
        int ncandidates;
        // 1. candidates and their processing order
        if (!phase2) {
            for (int s = tid; s < size; s += QT_THREADS) S.scan[s] = cnt[s] > 1 ? 1 : 0;
            __syncthreads();
            ncandidates = block_excl_scan(S.scan, size, s_warp);
            for (int s = tid; s < size; s += QT_THREADS) {
                if (cnt[s] > 1) { S.rank[s] = S.scan[s]; S.slot_of_rank[S.scan[s]] = s; } else S.rank[s] = -1;
            }
        } else {
            int m = 1; while (m < cprev) m <<= 1;
            for (int s = tid; s < m; s += QT_THREADS)
                S.sortk[s] = (s < cprev && cnt[s] > 1) ? (((unsigned)(0xFFFFF - min(cnt[s], 0xFFFFF)) << 12) | (unsigned)s) : 0xFFFFFFFFu;
            for (int s = tid; s < size; s += QT_THREADS) S.rank[s] = -1;
            if (tid == 0) s_ncand = 0;
            __syncthreads();
            for (int k2 = 2; k2 <= m; k2 <<= 1)                      // bitonic sort ascending
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < m; i += QT_THREADS) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const unsigned a = S.sortk[i], b = S.sortk[ixj];
                            const bool up = (i & k2) == 0;
                            if ((a > b) == up) { S.sortk[i] = b; S.sortk[ixj] = a; }
                        }
                    }
                    __syncthreads();
                }
            int local = 0;
            for (int r = tid; r < m; r += QT_THREADS)
                if (S.sortk[r] != 0xFFFFFFFFu) { const int s = S.sortk[r] & 0xFFF; S.rank[s] = r; S.slot_of_rank[r] = s; local++; }
            if (local) atomicAdd(&s_ncand, local);
            __syncthreads();
            ncandidates = s_ncand;
        }
        __syncthreads();
        // 2. child key counts of every candidate
        for (int i = tid; i < 4 * size; i += QT_THREADS) S.cc[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += QT_THREADS) {
            const int s = lab[i];
            if (S.rank[s] >= 0) {
                const unsigned c = kk[i];
                const int q = quadrant(box + 4 * s, (int)(c & 0xFFF) - minB, (int)((c >> 12) & 0xFFF) - minB);
                atomicAdd(&S.cc[4 * s + q], 1);
            }
        }
        __syncthreads();
        // 3. children per candidate in rank order; stopping point in phase 2
        for (int r = tid; r < ncandidates; r += QT_THREADS) {
            const int* c4 = S.cc + 4 * S.slot_of_rank[r];
            S.scan[r] = (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0);
        }
        if (tid == 0) { s_ndiv = ncandidates; s_nexp = 0; }
        __syncthreads();
        block_excl_scan(S.scan, ncandidates, s_warp);                 // scan[r] = children created before rank r
        if (phase2) {
            for (int r = tid; r < ncandidates; r += QT_THREADS) {
                const int* c4 = S.cc + 4 * S.slot_of_rank[r];
                const int nch = (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0);
                const int after = size + S.scan[r] + nch - (r + 1);   // |list| after dividing ranks 0..r
                if (after >= N) atomicMin(&s_ndiv, r + 1);
            }
            __syncthreads();
        }
        const int ndiv = s_ndiv;
        for (int r = ndiv + tid; r < ncandidates; r += QT_THREADS) S.rank[S.slot_of_rank[r]] = -1;
        __syncthreads();
        // 4. totals
        int C = 0;
        if (ndiv > 0) {
            const int sl = S.slot_of_rank[ndiv - 1]; const int* c4 = S.cc + 4 * sl;
            C = S.scan[ndiv - 1] + (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0);
        }
        __syncthreads();
        // children (scan[] is consumed here before it is reused for the survivors)
        for (int r = tid; r < ndiv; r += QT_THREADS) {
            const int s = S.slot_of_rank[r];
            int t = S.scan[r];
            int nexp = 0;
            for (int q = 0; q < 4; q++) {
                const int c = S.cc[4 * s + q];
                if (c > 0) {
                    const int pos = C - 1 - t; t++;
                    child_box(box + 4 * s, q, nbox + 4 * pos);
                    ncnt[pos] = c;
                    S.childpos[4 * s + q] = pos;
                    nexp += c > 1;
                }
            }
            if (nexp) atomicAdd(&s_nexp, nexp);
        }
        __syncthreads();
        for (int s = tid; s < size; s += QT_THREADS) S.scan[s] = S.rank[s] < 0 ? 1 : 0;
        __syncthreads();
        const int nsurv = block_excl_scan(S.scan, size, s_warp);
        if (C + nsurv > cap) { if (tid == 0) { atomicOr(status, 2); win_count[(size_t)f * P.nlevels + level] = 0; } return; }
        for (int s = tid; s < size; s += QT_THREADS)
            if (S.rank[s] < 0) {
                const int pos = C + S.scan[s];
                S.newpos[s] = pos;
                for (int k = 0; k < 4; k++) nbox[4 * pos + k] = box[4 * s + k];
                ncnt[pos] = cnt[s];
            }
        __syncthreads();
        // 6. relabel keys
        for (int i = tid; i < n; i += QT_THREADS) {
            const int s = lab[i];
            if (S.rank[s] >= 0) {
                const unsigned c = kk[i];
                const int q = quadrant(box + 4 * s, (int)(c & 0xFFF) - minB, (int)((c >> 12) & 0xFFF) - minB);
                lab[i] = (unsigned short)S.childpos[4 * s + q];
            } else lab[i] = (unsigned short)S.newpos[s];
        }
        __syncthreads();
        size = C + nsurv; cprev = C; cur ^= 1;
        const int nexp = s_nexp;
        // 7. control flow of :1141-1204
        if (size >= N || size == prev_size) finish = true;
        else if (!phase2 && size + 3 * nexp > N) phase2 = true;
        __syncthreads();
    }

    // ---- d. best key per node: max response, first in the reference's raw order on ties (:1208-1227)
    for (int s = tid; s < size; s += QT_THREADS) { S.best_score[s] = -1; S.best_okey[s] = 0xFFFFFFFFu; }
    __syncthreads();
    for (int i = tid; i < n; i += QT_THREADS) atomicMax(&S.best_score[lab[i]], (int)(kk[i] >> 24));
    __syncthreads();
    auto okey = [&](unsigned c) -> unsigned {
        const int x = (int)(c & 0xFFF) - EDGE, y = (int)((c >> 12) & 0xFFF) - EDGE;
        const int cx = x / L.wcell, cy = y / L.hcell;
        return ((unsigned)(cy * L.ncols + cx) << 14) | ((unsigned)(y - cy * L.hcell) << 7) | (unsigned)(x - cx * L.wcell);
    };
    for (int i = tid; i < n; i += QT_THREADS) {
        const unsigned c = kk[i];
        if ((int)(c >> 24) == S.best_score[lab[i]]) atomicMin(&S.best_okey[lab[i]], okey(c));
    }
    __syncthreads();
    if (size > L.kp_cap) { if (tid == 0) { atomicOr(status, 4); win_count[(size_t)f * P.nlevels + level] = 0; } return; }
    for (int i = tid; i < n; i += QT_THREADS) {
        const unsigned c = kk[i];
        const int s = lab[i];
        if ((int)(c >> 24) == S.best_score[s] && okey(c) == S.best_okey[s]) wout[s] = c;
    }
    if (tid == 0) win_count[(size_t)f * P.nlevels + level] = size;
}

// --------------------------------------------------------------------------------------------------------
// K6: GaussianBlur 7x7 sigma 2, fixed point [18,34,48,56,48,34,18]/256 per axis, out = (sum + 2^15) >> 16
// (SURVEY A.6 variant A).  The blurred plane keeps the layout of the pyramid plane; its 4-px border ring holds the
// UNBLURRED reflect-101 border, which is what the reference's in-place ROI blur leaves there.
// One CTA per 128x192 tile (host-built tile table: no level search, no division): the tile + halo arrives as one TMA box; each
// thread owns 4 adjacent columns and walks down 24 rows (measured per 256 frames: 8 rows 0.290 ms, 16 rows 0.256 ms with the first
// vertical pass; 16 rows 0.202 ms, 24 rows 0.194 ms with this one — the 6 halo rows of the horizontal pass are amortised over more
// outputs; 32 rows would need a TMA box taller than 256).  Horizontal pass: 2 IDP.4A per pixel.  Vertical pass: the horizontal sums
// of consecutive rows are kept packed 16x2 in a 6-row register window, so the 7 taps are 3 IDP.2A + 1 IMAD per pixel (was 3 IADD +
// 4 IMAD), and the four output bytes are gathered by 3 PRMT (byte 2 of each sum).  The last tile entry of every level copies the
// border ring.
// --------------------------------------------------------------------------------------------------------
#ifndef BL_R_
#define BL_R_ 24
#endif
constexpr int BL_W = 128, BL_R = BL_R_, BL_H = 8 * BL_R;   // tile, rows per warp
constexpr int BL_BOXW = BL_W + 32, BL_BOXH = BL_H + 6;  // TMA box: 16 B aligned start, 16 px slack left and right

__global__ void __launch_bounds__(256)
k_blur(const CUtensorMap* __restrict__ tmaps, const unsigned* __restrict__ btab, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur,
       const __grid_constant__ Plan P)
{
    __shared__ __align__(128) unsigned s_in[BL_BOXH * (BL_BOXW / 4)];
    __shared__ __align__(8) uint64_t s_mbar;
    const unsigned te = __ldg(btab + blockIdx.x);         // level | ty << 4 | tx << 16, bit 31 = ring tile (host-built: no search, no division)
    const int level = te & 15;
    const LevelInfo& L = P.lv[level];
    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Lw = L.w, Lh = L.h, ps = L.pstride;
    const size_t plane = (size_t)f * P.frame_bytes + L.poff + (size_t)EDGE * ps + EDGE;
    uint8_t* out = blur + plane;
    if (te >> 31) {                                       // ring tile: unblurred border copy
        const uint8_t* in = pyr + plane;
        const int rw = Lw + 2 * BORDER_W;
        for (int r = 0; r < 2 * BORDER_W; r++) {                       // rows -4..-1 and h..h+3
            const int y = r < BORDER_W ? r - BORDER_W : Lh + (r - BORDER_W);
            for (int i = tid; i < rw; i += 256) out[(ptrdiff_t)y * ps + i - BORDER_W] = in[(ptrdiff_t)y * ps + i - BORDER_W];
        }
        for (int i = tid; i < 2 * BORDER_W * Lh; i += 256) {          // columns -4..-1 and w..w+3
            const int y = i / (2 * BORDER_W), k = i - y * (2 * BORDER_W);
            const int x = k < BORDER_W ? k - BORDER_W : Lw + (k - BORDER_W);
            out[(ptrdiff_t)y * ps + x] = in[(ptrdiff_t)y * ps + x];
        }
        return;
    }
    const int x0 = (int)(te >> 16) * BL_W, y0 = (int)((te >> 4) & 0xFFF) * BL_H;
    if (tid == 0) { mbar_init(&s_mbar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&s_mbar, BL_BOXW * BL_BOXH);
        tma_load_3d(s_in, tmaps + MAXLEV + level, x0, y0 + EDGE - 3, f, &s_mbar);   // padded coords: image (x0-16, y0-3)
    }
    const int x = x0 + 4 * lane;                          // first of this thread's 4 columns
    const int yw = y0 + BL_R * warp;                      // first output row of this warp
    const unsigned K0 = 18u | (34u << 8) | (48u << 16) | (56u << 24);     // taps for bytes x-3..x
    const unsigned K1 = 48u | (34u << 8) | (18u << 16);                   // taps for bytes x+1..x+3
    // vertical taps over row pairs packed 16x2 (a horizontal sum is at most 255 * 256): rows (j, j+1), (j+2, j+3), (j+4, j+5) by
    // one IDP.2A each with these coefficient pairs, row j+6 by an IMAD that also adds the rounding constant
    const unsigned V01 = 18u | (34u << 8), V23 = 48u | (56u << 8), V45 = 48u | (34u << 8);
    const unsigned* S = s_in + (BL_R * warp) * (BL_BOXW / 4) + 4 + lane;  // row yw-3, centre word
    uint8_t* dst = out + (ptrdiff_t)yw * ps + x;          // walks down one row per output: no per-row 64-bit address arithmetic
    const int nrow = min(BL_R, Lh - yw);                   // rows of this warp inside the image (>= 1 for live warps)
    const bool full = x + 3 < Lw;
    mbar_wait(&s_mbar, 0);
    if (x >= Lw || yw >= Lh) return;
    unsigned pr[6][4];                                     // pr[r % 6][k] = h(row r) | h(row r + 1) << 16
    int hprev[4];
#pragma unroll
    for (int r = 0; r < BL_R + 6; r++) {
        const unsigned l = S[r * (BL_BOXW / 4) - 1], m = S[r * (BL_BOXW / 4)], n = S[r * (BL_BOXW / 4) + 1];
        int hr[4];
        hr[0] = __dp4a(__funnelshift_r(l, m, 8), K0, __dp4a(__funnelshift_r(m, n, 8), K1, 0u));
        hr[1] = __dp4a(__funnelshift_r(l, m, 16), K0, __dp4a(__funnelshift_r(m, n, 16), K1, 0u));
        hr[2] = __dp4a(__funnelshift_r(l, m, 24), K0, __dp4a(__funnelshift_r(m, n, 24), K1, 0u));
        hr[3] = __dp4a(m, K0, __dp4a(n, K1, 0u));
        if (r >= 1) {
#pragma unroll
            for (int k = 0; k < 4; k++) pr[(r - 1) % 6][k] = __byte_perm((unsigned)hprev[k], (unsigned)hr[k], 0x5410);
        }
        if (r >= 6) {
            const int j = r - 6;                          // output row yw + j uses window rows j..j+6
            if (j < nrow) {
                unsigned a[4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                    a[k] = __dp2a_lo(pr[j % 6][k], V01, __dp2a_lo(pr[(j + 2) % 6][k], V23, __dp2a_lo(pr[(j + 4) % 6][k], V45, 18u * (unsigned)hr[k] + 32768u)));
                // a < 2^24: the output pixel (a >> 16) is byte 2 of a
                const unsigned o = __byte_perm(__byte_perm(a[0], a[1], 0x0062), __byte_perm(a[2], a[3], 0x0062), 0x5410);
                if (full) *reinterpret_cast<unsigned*>(dst) = o;
                else for (int k = 0; k < 4; k++) if (x + k < Lw) dst[k] = (uint8_t)(o >> (8 * k));
                dst += ps;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) hprev[k] = hr[k];
    }
}

// --------------------------------------------------------------------------------------------------------
// K8: selection — which (level, winner) pairs become output rows, in output order (src/ORBextractor.cc:872-915).
// full_detect: all winners, level-major.  Otherwise the greedy occupancy-grid filter (inherently sequential: one
// thread per frame replays it).  sel entry = level << 16 | index; 0xFFFF0000 | i marks incoming keypoint i.
// --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_select(const unsigned* __restrict__ winners, const int* __restrict__ win_count, unsigned* __restrict__ sel, int* __restrict__ nsel,
         int sel_cap, int full_detect, int n_incoming, int32_t* __restrict__ grid, int grid_rows, int grid_cols, int min_px_dist, int num_needed,
         const int* __restrict__ dyn, int* __restrict__ status, const __grid_constant__ Plan P)
{
    const int f = blockIdx.x, tid = threadIdx.x;
    // single-frame call: the two arguments that change from call to call on the live path (src/Tracking.cc:946: num_featsneeded;
    // the number of incoming keypoints) are read from device memory, so that the captured CUDA graph of the call stays valid
    if (dyn) { n_incoming = dyn[0]; num_needed = dyn[1]; }
    const int* wc = win_count + (size_t)f * P.nlevels;
    unsigned* out = sel + (size_t)f * sel_cap;
    if (full_detect) {
        int off = 0;
        for (int l = 0; l < P.nlevels; l++) {
            const int c = wc[l];
            if (off + c <= sel_cap) for (int i = tid; i < c; i += 256) out[off + i] = ((unsigned)l << 16) | (unsigned)i;
            off += c;
        }
        if (tid == 0) { if (off > sel_cap) { atomicOr(status, 8); off = 0; } nsel[f] = off; }
        return;
    }
    if (tid != 0) return;
    int n = 0;
    bool over = false;
    for (int i = 0; i < n_incoming; i++) { if (n < sel_cap) out[n] = 0xFFFF0000u | (unsigned)i; else over = true; n++; }
    int total = 0, kp = 0; bool brk = false;
    for (int l = 0; l < P.nlevels && !brk; l++) {
        const int c = wc[l];
        if (c == 0) continue;
        const int quota = num_needed * (8 - l) / 30;
        const float scale = P.lv[l].scale;
        const unsigned* w = winners + (size_t)f * P.kp_per_frame + P.lv[l].kp_off;
        for (int i = 0; i < c; i++) {
            const unsigned e = w[i];
            const float tx = __fmul_rn((float)(e & 0xFFF), scale), ty = __fmul_rn((float)((e >> 12) & 0xFFF), scale);
            const int r = (int)__fdiv_rn(ty, (float)min_px_dist), cc = (int)__fdiv_rn(tx, (float)min_px_dist);
            if (r < 0 || r >= grid_rows || cc < 0 || cc >= grid_cols) continue;   // the reference would index out of bounds
            int32_t* cell = grid + (size_t)cc * grid_rows + r;
            if (*cell > 0) continue;
            if (n < sel_cap) out[n] = ((unsigned)l << 16) | (unsigned)i; else over = true;
            n++;
            (*cell)++;
            kp++; total++;
            if (kp == quota) { kp = 0; break; }
            if (total == num_needed) { brk = true; break; }
        }
    }
    if (over) { atomicOr(status, 8); n = 0; }
    nsel[f] = n;
}

// Level-0 border ring beyond BORDER_W, out to the full EDGE_THRESHOLD (16 px), in BOTH planes (the blurred plane's ring holds the
// unblurred border).  Only the incoming level-0 keypoints of the occupancy-grid path (ComputeKeyPointsCopy,
// src/ORBextractor.cc:523-534) can lie closer than 16 px to the image border: their IC_Angle disc reaches 15 px and their
// descriptor pattern 18 px beyond the keypoint, i.e. into the reference's 16-px reflect-101 border.  Frame 0 of a single-frame call.
__global__ void __launch_bounds__(256)
k_ring16(uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur, const __grid_constant__ Plan P)
{
    const LevelInfo& L = P.lv[0];
    const int w = L.w, h = L.h, ps = L.pstride;
    const size_t plane = L.poff + (size_t)EDGE * ps + EDGE;
    uint8_t* a = pyr + plane; uint8_t* b = blur + plane;
    const int RW = w + 2 * EDGE, NB = EDGE - BORDER_W;                   // bands of NB rows / columns outside the materialised ring
    const int nrowpix = 2 * NB * RW, ncolpix = 2 * NB * (h + 2 * BORDER_W);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nrowpix + ncolpix; i += gridDim.x * 256) {
        int x, y;
        if (i < nrowpix) { const int r = i / RW; x = i - r * RW - EDGE; y = r < NB ? r - EDGE : h + BORDER_W + (r - NB); }
        else { const int j = i - nrowpix, r = j / (2 * NB), c = j - r * (2 * NB); y = r - BORDER_W; x = c < NB ? c - EDGE : w + BORDER_W + (c - NB); }
        const uint8_t v = a[(ptrdiff_t)reflect101(y, h) * ps + reflect101(x, w)];
        a[(ptrdiff_t)y * ps + x] = v; b[(ptrdiff_t)y * ps + x] = v;
    }
}

// --------------------------------------------------------------------------------------------------------
// K5+K7: one warp per output keypoint.  IC_Angle (src/ORBextractor.cc:125-152) on the unblurred plane: lane u
// accumulates column u-15 of the radius-15 disc, warp-shuffle reduction of the two moments, cv::fastAtan2
// polynomial with every float op rounded separately.  computeOrbDescriptor (:155-195) on the blurred plane: lane i
// produces descriptor byte i from its 16 pattern points (pattern staged transposed in shared memory).
// --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float sc = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * sc, p3 = -0.3258083974640975f * sc;
    const float p5 = 0.1555786518463281f * sc, p7 = -0.04432655554792128f * sc;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// sin / cos of a float angle (radians, |x| < ~7) evaluated in double and rounded once to float.  glibc's sinf/cosf
// (what the reference's `cos(angle)` / `sin(angle)` resolve to) are correctly rounded except in vanishingly rare
// cases, so a double evaluation with ~1e-16 error reproduces their bits.  Cody-Waite reduction by pi/2 + the classic
// fdlibm kernel polynomials on [-pi/4, pi/4].
__device__ __forceinline__ void sincos_as_float(float xf, float* s_out, float* c_out)
{
    const double x = (double)xf;
    const double q = rint(x * 0.63661977236758134308);            // 2/pi
    double r = fma(-q, 1.57079632673412561417e+00, x);            // pi/2 split in three parts (fdlibm pio2_1, pio2_1t hi/lo)
    r = fma(-q, 6.07710050650619224932e-11, r);
    r = fma(-q, 2.02226624879595063154e-21, r);
    const double z = r * r;
    const double ps = fma(z, fma(z, fma(z, fma(z, fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08), 2.75573137070700676789e-06),
                                         -1.98412698298579493134e-04), 8.33333333332248946124e-03), -1.66666666666666324348e-01);
    const double sn = fma(z * r, ps, r);
    const double pc = fma(z, fma(z, fma(z, fma(z, fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09), -2.75573143513906633035e-07),
                                         2.48015872894767294178e-05), -1.38888888888741095749e-03), 4.16666666666666019037e-02);
    const double cs = fma(z * z, pc, fma(z, -0.5, 1.0));
    const int n = (int)q & 3;
    const double s = (n & 1) ? cs : sn, c = (n & 1) ? sn : cs;
    *s_out = (float)((n & 2) ? -s : s);
    *c_out = (float)(((n + 1) & 2) ? -c : c);
}

#ifndef DESC_KPW_
#define DESC_KPW_ 4
#endif
constexpr int DESC_WARPS = 8, DESC_KPW = DESC_KPW_;           // warps per CTA, keypoints per warp
constexpr int DESC_AROW = 48;                                  // bytes per staged row of the IC_Angle patch (31 columns from a 16-byte aligned start)
constexpr int DESC_PR = 18, DESC_PROW = 80;                   // staged patch: rows / columns -18..18 (|pattern coordinate| <= 13, rotated): 64 bytes per row from a
                                                              // 16-byte aligned start, rows 80 bytes apart (20 words: the random byte gathers spread over all banks)
constexpr int DESC_PATCH_BYTES = (2 * DESC_PR + 1) * DESC_PROW, DESC_APATCH_BYTES = (2 * HALF_PATCH + 1) * DESC_AROW;
constexpr int DESC_SM_WU = 512 * 8, DESC_SM_PATCH = DESC_SM_WU + 8 * 32 * 4, DESC_SM_APATCH = DESC_SM_PATCH + DESC_WARPS * DESC_PATCH_BYTES,
              DESC_SMEM = DESC_SM_APATCH + DESC_WARPS * 2 * DESC_APATCH_BYTES;      // dynamic shared memory of k_describe
#ifndef DESC_MINB
#define DESC_MINB 4          // 64 registers: four CTAs per SM (the kernel is latency-bound; measured 0.37 -> 0.27 ms at batch 256)
#endif
__global__ void __launch_bounds__(DESC_WARPS * 32, DESC_MINB)
k_describe(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur, const unsigned* __restrict__ winners,
           const unsigned* __restrict__ sel, const int* __restrict__ nsel, int sel_cap,
           const uvip_keypoint* __restrict__ incoming, const float2* __restrict__ pat_t,
           uvip_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, int32_t* __restrict__ n_out, int out_cap,
           int* __restrict__ status, const __grid_constant__ Plan P)
{
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int n = nsel[f];
    if (blockIdx.x == 0 && threadIdx.x == 0) { n_out[f] = n; if (n > out_cap) atomicOr(status, 8); }
    const int slot0 = (blockIdx.x * DESC_WARPS + (threadIdx.x >> 5)) * DESC_KPW;
    if ((int)(blockIdx.x * DESC_WARPS * DESC_KPW) >= n) return;          // whole CTA idle (uniform)
    // pattern in shared memory, transposed [k][lane]: lane reads its 16 points (descriptor byte `lane`) conflict-free,
    // which keeps the kernel at 64 registers (occupancy matters more than the 16 LDS: the kernel is latency-bound)
    extern __shared__ __align__(16) unsigned char s_desc[];
    float2* s_pat = reinterpret_cast<float2*>(s_desc);                                           // [512]
    unsigned (*s_wu)[32] = reinterpret_cast<unsigned (*)[32]>(s_desc + DESC_SM_WU);              // [8][32]
    unsigned char* s_patch_w = s_desc + DESC_SM_PATCH + (threadIdx.x >> 5) * DESC_PATCH_BYTES;   // the warp's blurred patch
    unsigned char* s_apatch_w = s_desc + DESC_SM_APATCH + (threadIdx.x >> 5) * 2 * DESC_APATCH_BYTES;   // its two IC_Angle patches
    // IC_Angle weights of disc row v = lane - 15: byte b of word j stands for column u = 4 j + b - 15 and holds u + 16 (1..31) where
    // |u| <= umax[|v|], 0 elsewhere (and everywhere for lane 31).  sum (u + 16) I - 16 sum I = sum u I; the 0/1 weights of the row
    // sum are the non-zero bytes of the same word
    for (int i = threadIdx.x; i < 512; i += DESC_WARPS * 32) s_pat[i] = __ldg(pat_t + i);
    {
        const int j = threadIdx.x >> 5, ln = threadIdx.x & 31, v = ln - HALF_PATCH;
        const int vmr = ln < 31 ? c_umax[v < 0 ? -v : v] : -1;
        unsigned wu = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int uu = 4 * j + b - HALF_PATCH;
            if ((uu < 0 ? -uu : uu) <= vmr) wu |= (unsigned)(uu + 16) << (8 * b);
        }
        s_wu[j][ln] = wu;
    }
    static_assert(DESC_WARPS == 8, "one warp of the CTA per weight word");
    __syncthreads();
    if (slot0 >= n) return;
    // the warp's keypoints: lane i fetches selection entry and winner word of keypoint i, so that the two dependent global loads in
    // front of every keypoint's pixel loads are paid once per warp
    static_assert(DESC_KPW <= 32, "one lane per keypoint of the warp");
    unsigned my_e = 0, my_w = 0;
    if (lane < DESC_KPW && slot0 + lane < n && slot0 + lane < out_cap) {
        my_e = sel[(size_t)f * sel_cap + slot0 + lane];
        if ((my_e >> 16) != 0xFFFFu) my_w = winners[(size_t)f * P.kp_per_frame + P.lv[my_e >> 16].kp_off + (my_e & 0xFFFF)];
    }
    // where keypoint i of the warp is (its level and integer position); false behind the warp's last keypoint
    auto locate = [&](int i, unsigned& e, unsigned& wv, int& level, int& cx, int& cy) -> bool {
        const int slot = slot0 + i;
        if (i >= DESC_KPW || slot >= n || slot >= out_cap) return false;
        e = __shfl_sync(0xFFFFFFFFu, my_e, i); wv = __shfl_sync(0xFFFFFFFFu, my_w, i);
        if ((e >> 16) == 0xFFFFu) {                // incoming level-0 keypoint (ComputeKeyPointsCopy, :523-534)
            const uvip_keypoint k = incoming[e & 0xFFFF];
            level = 0; cx = __float2int_rn(k.x); cy = __float2int_rn(k.y);
        } else { level = e >> 16; cx = wv & 0xFFF; cy = (wv >> 12) & 0xFFF; }
        return true;
    };
    // IC_Angle patch of a keypoint: rows cy-15..cy+15 of the UNBLURRED level, 48 bytes each from the 16-byte aligned column at or below
    // cx-15, staged by three 16-byte asynchronous copies per lane into one of the warp's two buffers (committed as one cp.async group)
    auto stage_angle_patch = [&](int level, int cx, int cy, int buf) {
        const LevelInfo& L = P.lv[level];
        const int ps = L.pstride;
        const unsigned char* src = pyr + (size_t)f * P.frame_bytes + L.poff + (size_t)EDGE * ps + EDGE + (ptrdiff_t)(cy - HALF_PATCH) * ps + ((cx - HALF_PATCH) & ~15);
        const unsigned dst = smem_u32(s_apatch_w + buf * DESC_APATCH_BYTES);
#pragma unroll
        for (int it = 0; it < 3; it++) {
            const int idx = it * 32 + lane, row = idx / 3, ch = idx - 3 * row;
            if (idx < (2 * HALF_PATCH + 1) * 3)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + row * DESC_AROW + ch * 16), "l"(src + (ptrdiff_t)row * ps + ch * 16) : "memory");
        }
    };
    unsigned e = 0, wv = 0; int level = 0, cx = 0, cy = 0;
    bool have = locate(0, e, wv, level, cx, cy);
    if (have) stage_angle_patch(level, cx, cy, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll 1
    for (int i = 0; i < DESC_KPW && have; i++) {
        const int slot = slot0 + i;
        float response, size, ox, oy; int octave, class_id;
        if ((e >> 16) == 0xFFFFu) {
            const uvip_keypoint k = incoming[e & 0xFFFF];
            response = k.response; size = k.size; ox = k.x; oy = k.y; octave = k.octave; class_id = k.class_id;
        } else {
            response = (float)(wv >> 24); size = P.lv[level].size; octave = level; class_id = -1;
            ox = (float)cx; oy = (float)cy;
            if (level != 0) { ox = __fmul_rn(ox, P.lv[level].scale); oy = __fmul_rn(oy, P.lv[level].scale); }   // :951-957
        }
        const LevelInfo& L = P.lv[level];
        const int ps = L.pstride;
        const size_t plane = (size_t)f * P.frame_bytes + L.poff + (size_t)EDGE * ps + EDGE;
        // The 512 rotated pattern points lie within +-18 px of the keypoint (|pattern| <= 13 per axis).  The 37 x 64-byte blurred
        // patch is staged in shared memory by 5 asynchronous 16-byte copies per lane (cp.async: no registers, in flight during the
        // whole orientation phase); the 32 byte gathers per lane then cost a few bank-conflict cycles each instead of a fully
        // divergent trip through the L1 tag stage and a 64-bit address each.
        // (detected keypoints are >= 16 px inside the image.  An incoming level-0 keypoint may lie anywhere inside it: the full 16-px
        // ring of level 0 is materialised for such calls (k_ring16), which covers every point at least 2 px inside the image exactly
        // as the reference's padded buffer does; closer than 2 px the reference's own pattern reads leave its buffer row — undefined
        // there — and the staged centre is moved to 2 px here.  A staged row is 64 bytes from an aligned start >= cx - 33: it may
        // run past column w + 16 into the row padding / the next row, which is allocated memory that no gather touches; the buffers
        // carry 4 KB of slack behind the last plane.)
        const int cxs = min(max(cx, 2), L.w - 3), cys = min(max(cy, 2), L.h - 3);
        {
            const int xs = (cxs - DESC_PR) & ~15;                                 // 16-byte aligned first column (planes are 16 B aligned)
            const unsigned char* src = blur + plane + (ptrdiff_t)(cys - DESC_PR + (lane >> 2)) * ps + xs + (lane & 3) * 16;
            const unsigned dsts = smem_u32(s_patch_w) + (lane >> 2) * DESC_PROW + (lane & 3) * 16;
            __syncwarp();                                  // the previous keypoint's gathers are done
#pragma unroll
            for (int it = 0; it < 5; it++)                 // 37 rows x 4 chunks of 16 bytes, 32 chunks (8 rows) per step
                if (it < 4 || lane < (2 * DESC_PR + 1 - 32) * 4)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsts + it * 8 * DESC_PROW), "l"(src + (ptrdiff_t)it * 8 * ps) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // the next keypoint's IC_Angle patch goes into the other buffer while this keypoint is worked on
        unsigned ne = 0, nwv = 0; int nlevel = 0, ncx = 0, ncy = 0;
        const bool nhave = locate(i + 1, ne, nwv, nlevel, ncx, ncy);
        if (nhave) stage_angle_patch(nlevel, ncx, ncy, (i + 1) & 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");       // everything but this keypoint's blurred patch and the next angle patch
        __syncwarp();
        // ---- IC_Angle: lane r sums row v = r - 15 of the radius-15 disc.  The row's 31 pixels lie in 9 aligned words of the staged
        //      patch; funnel shifts bring column -15 to byte 0, then sum u I and the row sum are one IDP.4A each per word
        int m10 = 0, m01 = 0;
        {
            const int x0 = cx - HALF_PATCH;
            // the lane's row as three 16-byte loads (rows are 48 bytes apart: conflict-free per quarter warp); which 9 of the 12 words
            // hold the 31 columns depends on the keypoint only, so the choice is a warp-uniform branch
            const uint4* rp = reinterpret_cast<const uint4*>(s_apatch_w + (i & 1) * DESC_APATCH_BYTES + min(lane, 2 * HALF_PATCH) * DESC_AROW);
            const uint4 q0 = rp[0], q1 = rp[1], q2 = rp[2];
            const unsigned w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            const int sh = (x0 & 3) * 8;
            unsigned wsum = 0, rowsum = 0;
#define IC_ROW(WO) { _Pragma("unroll") for (int j = 0; j < 8; j++) { \
                const unsigned x = __funnelshift_r(w[WO + j], w[WO + j + 1], sh), wt = s_wu[j][lane]; \
                wsum = __dp4a(wt, x, wsum); rowsum = __dp4a(((wt + 0x7F7F7F7Fu) >> 7) & 0x01010101u, x, rowsum); } }
            switch ((x0 & 15) >> 2) { case 0: IC_ROW(0) break; case 1: IC_ROW(1) break; case 2: IC_ROW(2) break; default: IC_ROW(3) break; }
#undef IC_ROW
            m10 = (int)wsum - 16 * (int)rowsum;
            m01 = (lane - HALF_PATCH) * (int)rowsum;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o); m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o); }
        const float angle = fast_atan2_deg((float)m01, (float)m10);
        // ---- rotated BRIEF (:155-195): every float op rounded separately, cvRound = round-half-even
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        float a, b;
        sincos_as_float(__fmul_rn(angle, factorPI), &b, &a);
        asm volatile("cp.async.wait_group 1;" ::: "memory");       // this keypoint's blurred patch has landed
        __syncwarp();                                      // ... and is visible to every lane
        const uint8_t* cb = s_patch_w + DESC_PR * DESC_PROW + DESC_PR + ((cxs - DESC_PR) & 15);
        unsigned val = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float2 p0 = s_pat[(2 * k) * 32 + lane], p1 = s_pat[(2 * k + 1) * 32 + lane];
            const float x0 = p0.x, y0 = p0.y, x1 = p1.x, y1 = p1.y;
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a))), q0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a))), q1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
            const int t0 = cb[r0 * DESC_PROW + q0], t1 = cb[r1 * DESC_PROW + q1];
            val |= (unsigned)(t0 < t1) << k;
        }
        desc[((size_t)f * out_cap + slot) * 32 + lane] = (uint8_t)val;
        if (lane == 0) {
            uvip_keypoint o;
            o.x = ox; o.y = oy; o.size = size; o.angle = angle; o.response = response; o.octave = octave; o.class_id = class_id;
            kps[(size_t)f * out_cap + slot] = o;
        }
        have = nhave; e = ne; wv = nwv; level = nlevel; cx = ncx; cy = ncy;
    }
}

// --------------------------------------------------------------------------------------------------------
// E8 (scoring half): HarrisResponses (src/ORBextractor.cc:80-121) on the unblurred level for a list of points: integer
// Sobel-like gradients over a blockSize x blockSize window, a = sum Ix^2, b = sum Iy^2, c = sum IxIy, response =
// ((float)a*b - (float)c*c - k*((float)a+b)*((float)a+b)) * scale^4 with every float operation rounded separately.
// One thread per point (lists are short); the quota-cell distribution around it (:536-746) is dead code in the reference.
// --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_harris(const uint8_t* __restrict__ pyr, int frame, int level, const float* __restrict__ xs, const float* __restrict__ ys, int n,
         int block_size, float harris_k, float* __restrict__ out, const __grid_constant__ Plan P)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    const LevelInfo& L = P.lv[level];
    const int ps = L.pstride;
    const uint8_t* img = pyr + (size_t)frame * P.frame_bytes + L.poff + (size_t)EDGE * ps + EDGE;
    const int r = block_size / 2;
    const int x0 = __float2int_rn(__fsub_rn(xs[i], (float)r)), y0 = __float2int_rn(__fsub_rn(ys[i], (float)r));      // cvRound(pt - r)
    int a = 0, b = 0, c = 0;
    for (int dy = 0; dy < block_size; dy++)
        for (int dx = 0; dx < block_size; dx++) {
            const uint8_t* p = img + (ptrdiff_t)(y0 + dy) * ps + (x0 + dx);
            const int Ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-ps + 1] - (int)p[-ps - 1]) + ((int)p[ps + 1] - (int)p[ps - 1]);
            const int Iy = ((int)p[ps] - (int)p[-ps]) * 2 + ((int)p[ps - 1] - (int)p[-ps - 1]) + ((int)p[ps + 1] - (int)p[-ps + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
    float scale = __fmul_rn((float)((1 << 2) * block_size), 255.0f);
    scale = __fdiv_rn(1.0f, scale);
    const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float fa = (float)a, fb = (float)b, fc = (float)c, ab = __fadd_rn(fa, fb);
    const float v = __fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(harris_k, ab), ab));
    out[i] = __fmul_rn(v, s4);
}

// --------------------------------------------------------------------------------------------------------
// E8 (detector half, optional mode): the reference's dead ComputeKeyPoints path (src/ORBextractor.cc:536-746).  Its cells are
// large quota cells (e.g. 144 x 64 px on level 0), each run through cv::FAST(fastTh, nms) on the cell + 3 px and again at
// threshold 5 when it yields <= 3 corners.  k_fast_roi does exactly that for one cell per CTA, unoptimised on purpose (no
// caller reaches this path in the reference): exact score of every pixel (s >= t <=> FAST-9 corner at t), strict 3x3 NMS inside
// the ROI, survivors written in cv::FAST's row-major order.  The quota logic and KeyPointsFilter::retainBest stay on the host.
// --------------------------------------------------------------------------------------------------------
struct RoiCell { int level, x0, y0, w, h; };          // ROI origin in level coordinates (inside the padded plane)

__device__ __forceinline__ int fast_score_exact(const uint8_t* p, int st)
{
    const int v = p[0];
    int r[16] = {p[3 * st], p[3 * st + 1], p[2 * st + 2], p[st + 3], p[3], p[-st + 3], p[-2 * st + 2], p[-3 * st + 1],
                 p[-3 * st], p[-3 * st - 1], p[-2 * st - 2], p[-st - 3], p[-3], p[st - 3], p[2 * st - 2], p[3 * st - 1]};
    int A = 0, B = 255;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int mn = r[k], mx = r[k];
#pragma unroll
        for (int j = 1; j < 9; j++) { mn = min(mn, r[(k + j) & 15]); mx = max(mx, r[(k + j) & 15]); }
        A = max(A, mn); B = min(B, mx);
    }
    return max(A - v, v - B) - 1;
}

__global__ void __launch_bounds__(256)
k_fast_roi(const uint8_t* __restrict__ pyr, int frame, const RoiCell* __restrict__ cells, int t1, int t2, int cap_cell,
           unsigned* __restrict__ lists, int* __restrict__ counts, int* __restrict__ status, const __grid_constant__ Plan P)
{
    extern __shared__ __align__(16) unsigned char s_roi[];
    __shared__ int s_rowcnt[512];
    __shared__ int s_total;
    const RoiCell c = cells[blockIdx.x];
    const LevelInfo& L = P.lv[c.level];
    const int ps = L.pstride, w = c.w, h = c.h, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* img = pyr + (size_t)frame * P.frame_bytes + L.poff + (size_t)(EDGE + c.y0) * ps + (EDGE + c.x0);
    uint8_t* s_img = s_roi; uint8_t* s_sc = s_roi + (size_t)w * h; uint8_t* s_keep = s_sc + (size_t)w * h;
    for (int i = tid; i < w * h; i += 256) s_img[i] = img[(size_t)(i / w) * ps + (i % w)];
    __syncthreads();
    unsigned* out = lists + (size_t)blockIdx.x * cap_cell;
    for (int pass = 0; pass < 2; pass++) {
        const int t = pass ? t2 : t1;
        for (int i = tid; i < w * h; i += 256) {
            const int y = i / w, x = i - y * w;
            int s = 0;
            if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) { s = fast_score_exact(s_img + i, w); s = s >= t ? s : 0; }
            s_sc[i] = (uint8_t)s;
        }
        if (tid == 0) s_total = 0;
        for (int i = tid; i < h; i += 256) s_rowcnt[i] = 0;
        __syncthreads();
        for (int i = tid; i < w * h; i += 256) {
            const int y = i / w, x = i - y * w;
            const int s = s_sc[i];
            bool keep = false;
            if (s > 0 && x >= 3 && x < w - 3 && y >= 3 && y < h - 3)
                keep = s > s_sc[i - 1] && s > s_sc[i + 1] && s > s_sc[i - w - 1] && s > s_sc[i - w] && s > s_sc[i - w + 1] &&
                       s > s_sc[i + w - 1] && s > s_sc[i + w] && s > s_sc[i + w + 1];
            s_keep[i] = keep ? 1 : 0;
            if (keep) atomicAdd(&s_rowcnt[y], 1);
        }
        __syncthreads();
        if (tid == 0) { int acc = 0; for (int y = 0; y < h; y++) { const int n = s_rowcnt[y]; s_rowcnt[y] = acc; acc += n; } s_total = acc; }
        __syncthreads();
        const int total = s_total;
        if (pass == 0 && total <= 3) { __syncthreads(); continue; }            // cv::FAST again at threshold 5 (:646-652), uniform
        if (total > cap_cell) { if (tid == 0) { atomicOr(status, 1); counts[blockIdx.x] = 0; } return; }
        for (int y = warp; y < h; y += 8) {                                    // one warp writes one row, in x order
            int base = s_rowcnt[y];
            for (int x0 = 0; x0 < w; x0 += 32) {
                const int x = x0 + lane;
                const bool k = x < w && s_keep[y * w + x];
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, k);
                if (k) out[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned)x | ((unsigned)y << 12) | ((unsigned)s_sc[y * w + x] << 24);
                base += __popc(bal);
            }
        }
        if (tid == 0) counts[blockIdx.x] = total;
        return;
    }
}

// HarrisResponses (:80-121) for every entry of the cell lists (cell-relative coordinates), blockSize 7
__global__ void __launch_bounds__(128)
k_harris_cells(const uint8_t* __restrict__ pyr, int frame, const RoiCell* __restrict__ cells, int cap_cell, const unsigned* __restrict__ lists,
               const int* __restrict__ counts, float harris_k, float* __restrict__ resp, const __grid_constant__ Plan P)
{
    const RoiCell c = cells[blockIdx.x];
    const LevelInfo& L = P.lv[c.level];
    const int ps = L.pstride;
    const uint8_t* img = pyr + (size_t)frame * P.frame_bytes + L.poff + (size_t)(EDGE + c.y0) * ps + (EDGE + c.x0);
    const int n = counts[blockIdx.x];
    float scale = __fmul_rn((float)((1 << 2) * 7), 255.0f);
    scale = __fdiv_rn(1.0f, scale);
    const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    for (int i = threadIdx.x; i < n; i += 128) {
        const unsigned e = lists[(size_t)blockIdx.x * cap_cell + i];
        const int x0 = (int)(e & 0xFFF) - 3, y0 = (int)((e >> 12) & 0xFFF) - 3;
        int a = 0, b = 0, cc = 0;
        for (int dy = 0; dy < 7; dy++)
            for (int dx = 0; dx < 7; dx++) {
                const uint8_t* p = img + (ptrdiff_t)(y0 + dy) * ps + (x0 + dx);
                const int Ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-ps + 1] - (int)p[-ps - 1]) + ((int)p[ps + 1] - (int)p[ps - 1]);
                const int Iy = ((int)p[ps] - (int)p[-ps]) * 2 + ((int)p[ps - 1] - (int)p[-ps - 1]) + ((int)p[ps + 1] - (int)p[-ps + 1]);
                a += Ix * Ix; b += Iy * Iy; cc += Ix * Iy;
            }
        const float fa = (float)a, fb = (float)b, fc = (float)cc, ab = __fadd_rn(fa, fb);
        const float v = __fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(harris_k, ab), ab));
        resp[(size_t)blockIdx.x * cap_cell + i] = __fmul_rn(v, s4);
    }
}

// IC_Angle (:125-152) for a list of (level, x, y): one warp per keypoint, as in k_describe
__global__ void __launch_bounds__(256)
k_angle_list(const uint8_t* __restrict__ pyr, int frame, const int* __restrict__ lxy, int n, float* __restrict__ angle, const __grid_constant__ Plan P)
{
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const int level = lxy[3 * i], cx = lxy[3 * i + 1], cy = lxy[3 * i + 2];
    const LevelInfo& L = P.lv[level];
    const int ps = L.pstride;
    const int u = lane - HALF_PATCH;
    const int vm = lane < 31 ? c_umax[u < 0 ? -u : u] : -1;
    const uint8_t* p = pyr + (size_t)frame * P.frame_bytes + L.poff + (size_t)(EDGE + cy - HALF_PATCH) * ps + (EDGE + cx + u);
    int colsum = 0, m01 = 0;
    for (int v = -HALF_PATCH; v <= HALF_PATCH; v++) { const int val = ((v < 0 ? -v : v) <= vm) ? (int)p[0] : 0; colsum += val; m01 += v * val; p += ps; }
    int m10 = u * colsum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o); m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o); }
    if (lane == 0) angle[i] = fast_atan2_deg((float)m01, (float)m10);
}

// --------------------------------------------------------------------------------------------------------
// N3: CLAHE pre-processing, cv::createCLAHE(4, Size(12,12))->apply(im, im) at src/Tracking.cc:425-431,
// restated from OpenCV imgproc/clahe.cpp (8-bit path).  k_clahe_lut: one CTA per (tile, frame) — shared-memory histogram
// of the tile (image padded to a tile multiple by reflect-101), clip + redistribute, cumulative LUT.  k_clahe_apply: four
// LUT gathers per pixel and the float bilinear blend, every float op rounded separately like the x86 reference build.
// --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_clahe_lut(const uint8_t* __restrict__ src, int w, int h, int stride, size_t pitch, int tiles_x, int tw, int th, int clip_limit,
            float lut_scale, uint8_t* __restrict__ lut)
{
    __shared__ int s_hist[256];
    __shared__ int s_wh[8][256];                              // one histogram per warp: atomics only collide inside a warp
    __shared__ int s_part[8];
    __shared__ int s_clipped;
    const int tile = blockIdx.x, f = blockIdx.y, tid = threadIdx.x, warp = tid >> 5;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const uint8_t* img = src + (size_t)f * pitch;
#pragma unroll
    for (int k = 0; k < 8; k++) s_wh[k][tid] = 0;
    if (tid == 0) s_clipped = 0;
    __syncthreads();
    for (int yy = warp; yy < th; yy += 8) {
        int y = ty * th + yy;
        if (y >= h) y = 2 * (h - 1) - y;                      // copyMakeBorder(..., BORDER_REFLECT_101) padding to a tile multiple
        const uint8_t* row = img + (size_t)y * stride;
        for (int xx = tid & 31; xx < tw; xx += 32) {
            int x = tx * tw + xx;
            if (x >= w) x = 2 * (w - 1) - x;
            atomicAdd(&s_wh[warp][__ldg(row + x)], 1);
        }
    }
    __syncthreads();
    {
        int t = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) t += s_wh[k][tid];
        s_hist[tid] = t;
    }
    int v = s_hist[tid];
    if (clip_limit > 0) {
        if (v > clip_limit) { atomicAdd(&s_clipped, v - clip_limit); v = clip_limit; }
        __syncthreads();
        const int clipped = s_clipped;
        const int batch = clipped / 256; int residual = clipped - batch * 256;
        v += batch;
        if (residual != 0) {
            const int step = max(256 / residual, 1);
            // bins 0, step, 2*step, ... get one more, at most `residual` of them
            if (tid % step == 0 && tid / step < residual) v++;
        }
    }
    // inclusive prefix sum over the 256 bins
    int incl = v;
    const int lane = tid & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_part[tid >> 5] = incl;
    __syncthreads();
    int off = 0;
    for (int k = 0; k < (tid >> 5); k++) off += s_part[k];
    const int sum = incl + off;
    const int o = __float2int_rn(__fmul_rn((float)sum, lut_scale));
    lut[((size_t)f * gridDim.x + tile) * 256 + tid] = (uint8_t)min(max(o, 0), 255);
}

constexpr int CL_W = 64, CL_H = 32, CL_MAXT = 4;           // pixel rectangle of one CTA; tables it may touch per axis

// generic path (any tile size): one pixel per thread, tables read through L1
__global__ void __launch_bounds__(256)
k_clahe_apply_generic(const uint8_t* __restrict__ src, int w, int h, int stride, size_t pitch, int tiles_x, int tiles_y, float inv_tw, float inv_th,
                      const uint8_t* __restrict__ lut, uint8_t* __restrict__ dst, int dstride, size_t dpitch)
{
    const int f = blockIdx.z;
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= w || y >= h) return;
    const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f), tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
    int tx1 = (int)floorf(txf), ty1 = (int)floorf(tyf);
    const float xa = __fsub_rn(txf, (float)tx1), ya = __fsub_rn(tyf, (float)ty1);
    const float xa1 = __fsub_rn(1.0f, xa), ya1 = __fsub_rn(1.0f, ya);
    int tx2 = min(tx1 + 1, tiles_x - 1), ty2 = min(ty1 + 1, tiles_y - 1);
    tx1 = max(tx1, 0); ty1 = max(ty1, 0);
    const int v = __ldg(src + (size_t)f * pitch + (size_t)y * stride + x);
    const uint8_t* L = lut + (size_t)f * tiles_x * tiles_y * 256 + v;
    const float l11 = (float)__ldg(L + (ty1 * tiles_x + tx1) * 256), l12 = (float)__ldg(L + (ty1 * tiles_x + tx2) * 256);
    const float l21 = (float)__ldg(L + (ty2 * tiles_x + tx1) * 256), l22 = (float)__ldg(L + (ty2 * tiles_x + tx2) * 256);
    const float res = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa)), ya1),
                                __fmul_rn(__fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa)), ya));
    dst[(size_t)f * dpitch + (size_t)y * dstride + x] = (uint8_t)min(max(__float2int_rn(res), 0), 255);
}


__global__ void __launch_bounds__(256)
k_clahe_apply(const uint8_t* __restrict__ src, int w, int h, int stride, size_t pitch, int tiles_x, int tiles_y, float inv_tw, float inv_th,
              const uint8_t* __restrict__ lut, uint8_t* __restrict__ dst, int dstride, size_t dpitch)
{
    // one CTA per CL_W x CL_H pixel rectangle; the look-up tables of every tile its pixels interpolate between are staged in
    // shared memory first (at most CL_MAXT x CL_MAXT of them, checked by the host), so the 4 look-ups per pixel are LDS.
    // The interpolation arithmetic is OpenCV's CLAHE_Interpolation_Body float sequence, one rounding per operation.
    __shared__ __align__(16) uint8_t s_lut[CL_MAXT * CL_MAXT][256];
    const int f = blockIdx.z, tid = threadIdx.x;
    const int x0 = blockIdx.x * CL_W, y0 = blockIdx.y * CL_H;
    const int x1 = min(x0 + CL_W, w) - 1, y1 = min(y0 + CL_H, h) - 1;
    const int txb = max((int)floorf(__fsub_rn(__fmul_rn((float)x0, inv_tw), 0.5f)), 0);
    const int tyb = max((int)floorf(__fsub_rn(__fmul_rn((float)y0, inv_th), 0.5f)), 0);
    const int txe = min((int)floorf(__fsub_rn(__fmul_rn((float)x1, inv_tw), 0.5f)) + 1, tiles_x - 1);
    const int tye = min((int)floorf(__fsub_rn(__fmul_rn((float)y1, inv_th), 0.5f)) + 1, tiles_y - 1);
    const int ntx = txe - txb + 1, nty = tye - tyb + 1;
    const uint8_t* Lf = lut + (size_t)f * tiles_x * tiles_y * 256;
    for (int i = tid; i < ntx * nty * 64; i += 256) {          // 64 words per table
        const int t = i >> 6, wd = i & 63;
        const int tyy = t / ntx, txx = t - tyy * ntx;
        reinterpret_cast<unsigned*>(s_lut[tyy * CL_MAXT + txx])[wd] =
            __ldg(reinterpret_cast<const unsigned*>(Lf + ((size_t)(tyb + tyy) * tiles_x + (txb + txx)) * 256) + wd);
    }
    __syncthreads();
    // thread = 4 consecutive columns x (CL_H / 16) rows; the column terms are computed once
    const int cx = x0 + 4 * (tid & (CL_W / 4 - 1)), ry = tid / (CL_W / 4);
    if (cx >= w) return;
    int t1[4], t2[4]; float xa[4], xa1[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float txf = __fsub_rn(__fmul_rn((float)(cx + k), inv_tw), 0.5f);
        const int a1 = (int)floorf(txf);
        xa[k] = __fsub_rn(txf, (float)a1); xa1[k] = __fsub_rn(1.0f, xa[k]);
        t2[k] = min(a1 + 1, tiles_x - 1) - txb; t1[k] = max(a1, 0) - txb;
    }
    const bool vec = ((cx + 3) < w) && (((size_t)src | (size_t)dst | (size_t)stride | (size_t)dstride | pitch | dpitch) & 3) == 0;
    for (int y = y0 + ry; y <= y1; y += 256 / (CL_W / 4)) {
        const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
        const int b1 = (int)floorf(tyf);
        const float ya = __fsub_rn(tyf, (float)b1), ya1 = __fsub_rn(1.0f, ya);
        const int r2 = (min(b1 + 1, tiles_y - 1) - tyb) * CL_MAXT, r1 = (max(b1, 0) - tyb) * CL_MAXT;
        const uint8_t* sp = src + (size_t)f * pitch + (size_t)y * stride + cx;
        uint8_t* dp = dst + (size_t)f * dpitch + (size_t)y * dstride + cx;
        unsigned in4 = 0;
        if (vec) in4 = __ldg(reinterpret_cast<const unsigned*>(sp));
        else
            for (int k = 0; k < 4; k++) if (cx + k < w) in4 |= (unsigned)__ldg(sp + k) << (8 * k);
        unsigned out4 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int v = (in4 >> (8 * k)) & 255;
            const float l11 = (float)s_lut[r1 + t1[k]][v], l12 = (float)s_lut[r1 + t2[k]][v];
            const float l21 = (float)s_lut[r2 + t1[k]][v], l22 = (float)s_lut[r2 + t2[k]][v];
            const float res = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(l11, xa1[k]), __fmul_rn(l12, xa[k])), ya1),
                                        __fmul_rn(__fadd_rn(__fmul_rn(l21, xa1[k]), __fmul_rn(l22, xa[k])), ya));
            out4 |= (unsigned)min(max(__float2int_rn(res), 0), 255) << (8 * k);
        }
        if (vec) *reinterpret_cast<unsigned*>(dp) = out4;
        else
            for (int k = 0; k < 4; k++) if (cx + k < w) dp[k] = (uint8_t)(out4 >> (8 * k));
    }
}

}  // namespace uvip

// =====================================================================================================
// host side
// =====================================================================================================
using namespace uvip;

static inline int cv_round_f(float v) { return (int)lrintf(v); }

struct uvip_extractor {
    uvip_extractor_params prm;
    int device = 0;
    int num_sms = 148;         // B200; read from the device at the first plan
    int fast2_ctas_per_sm = 0; // resident CTAs of k_fast2 per SM for the current plan (persistent grid = this x num_sms)
    bool use_fast1 = false;    // UVIP_FAST1=1: the first FAST kernel (A/B measurements)
    int subbatch = 0;          // > 0: a launch group is enqueued as sub-batches of this many frames through ALL stages (L2-resident pyramid)
    cudaStream_t stream = nullptr;
    float scale[MAXLEV], inv_scale[MAXLEV];
    int quota[MAXLEV];
    int umax[HALF_PATCH + 1];
    Plan plan;                 // for (plan.W, plan.H); W == 0 -> none yet
    // working set sized for (max_width, max_height, max_batch)
    size_t cap_frame_bytes = 0; int cap_cells = 0, cap_raw = 0, cap_kp = 0, cap_tab = 0;
    DevBuf pyr, blur, cand, labels, winners, counters, sel, nsel, tabs, status, grid, incoming, tmaps, pat_t;
    DevBuf in_frames, out_kps, out_desc, out_n;      // staging for the host-buffer entry points
    DevBuf clahe_lut, clahe_io;                      // CLAHE LUTs and host-call staging
    DevBuf dyn;                                      // single-frame call: {n_incoming, num_needed} read by k_select (outside the graph key)
    int* h_dyn = nullptr;                            // pinned source of the dyn upload
    DevBuf in2, kps2, desc2, n2;                     // second staging set: uvip_extract_batch double-buffers its chunks
    DevBuf knn_i[2], knn_d[2];                       // uvip_extract_match_batch_submit: kNN2 results of a chunk's frame pairs
    int prev_nb = 0;                                 // frames of the previously submitted chunk (the boundary pair reads its last frame)
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr};       // one per ticket of uvip_extract_batch_submit
    int* h_status = nullptr;                           // pinned, one word per ticket
    long long chunk_seq = 0, ticket_seq = 0;
    int inflight = 0;
    bool ticket_busy[2] = {false, false};
    int sel_cap = 0;
    int last_frames = 0;
    long long launches = 0;
    // optional per-stage timing (bench.py roofline): a ring of event sets, one per launch group
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;      // PROF_RING * (UVIP_NUM_STAGES + 1)
    long long prof_groups = 0;
    // single-frame call (uvip_extract = operator()): the memsets + 13 kernels of one frame are captured once per call shape
    // into a CUDA graph and replayed (the call is launch-bound: a frame's kernels are tiny), results come back through one
    // pinned staging block with a single synchronisation
    struct GraphKey { int w, h, stride, cap, full, has_in, gr, gc, mpd; const void *in, *kps, *desc, *n, *grid, *incoming, *pyr; };
    GraphKey gkey; bool ghave = false; cudaGraphExec_t gexec = nullptr; long long glaunches = 0, gcaptures = 0;
    uint8_t* h_stage = nullptr; size_t h_stage_bytes = 0;      // pinned: [n, status | keypoints cap x 28 | descriptors cap x 32]
    std::mutex mu;
};
constexpr int PROF_RING = 256;

// per-axis resize table of cv::resize INTER_LINEAR (8-bit fixed point), see SURVEY A.2
static void build_axis_table(int src_n, int dst_n, int* ofs, int* coef)
{
    const double inv_scale = (double)dst_n / src_n;
    const double scale = 1. / inv_scale;
    for (int d = 0; d < dst_n; d++) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(fx);
        fx -= s;
        if (s < 0) { s = 0; fx = 0; }
        if (s >= src_n - 1) { s = src_n - 1; fx = 0; }
        ofs[d] = s;
        int a0 = cv_round_f((1.f - fx) * 2048.f), a1 = cv_round_f(fx * 2048.f);
        a0 = a0 > 32767 ? 32767 : (a0 < -32768 ? -32768 : a0);
        a1 = a1 > 32767 ? 32767 : (a1 < -32768 ? -32768 : a1);
        coef[d] = (int)(((unsigned)a0 & 0xFFFFu) | ((unsigned)a1 << 16));
    }
}

// geometry of a launch group for frames of w x h; returns UVIP_ERR_UNSUPPORTED if outside the envelope
static int make_plan(const uvip_extractor* ex, int w, int h, Plan* out, std::vector<int>* tabs)
{
    Plan P; memset(&P, 0, sizeof(P));
    const uvip_extractor_params& p = ex->prm;
    P.nlevels = p.nlevels; P.W = w; P.H = h;
    P.fast_th = p.fast_th; P.retry_th = p.retry_th;
    P.t1 = p.fast_th > 1 ? p.fast_th : 1; P.t2 = p.retry_th > 1 ? p.retry_th : 1;
    P.tmin = P.t1 < P.t2 ? P.t1 : P.t2;
    if (P.t1 > 254 || P.t2 > 254) { set_last_error("FAST thresholds above 254 are unsupported"); return UVIP_ERR_UNSUPPORTED; }
    size_t off = 0; int cells = 0, raw = 0, kp = 0, ft = 0, bt = 0, tab = 0, maxN = 0, max_tw = 0, max_th = 0;
    for (int l = 0; l < p.nlevels; l++) {
        LevelInfo& L = P.lv[l];
        L.w = cv_round_f((float)w * ex->inv_scale[l]);          // src/ORBextractor.cc:968
        L.h = cv_round_f((float)h * ex->inv_scale[l]);
        if (L.w < 64 || L.h < 64 || L.w > 4095 || L.h > 4095) {
            set_last_error("level %d is %dx%d: supported level sizes are 64..4095", l, L.w, L.h);
            return UVIP_ERR_UNSUPPORTED;
        }
        L.pstride = (int)align_up((size_t)L.w + 2 * EDGE, 32);
        L.poff = (unsigned)off;
        off += align_up((size_t)L.pstride * (L.h + 2 * EDGE) + 64, 256);
        // cells, src/ORBextractor.cc:756-770
        const int minB = EDGE - 3, maxBX = L.w - EDGE + 3, maxBY = L.h - EDGE + 3;
        const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
        const float Wc = (float)p.cell;
        L.ncols = (int)(width / Wc); L.nrows = (int)(height / Wc);
        if (L.ncols < 1 || L.nrows < 1) { set_last_error("level %d has no FAST cell", l); return UVIP_ERR_UNSUPPORTED; }
        L.wcell = (int)ceilf(width / L.ncols); L.hcell = (int)ceilf(height / L.nrows);
        if (L.wcell > 64 || L.hcell > 64) { set_last_error("FAST cell larger than 64 px"); return UVIP_ERR_UNSUPPORTED; }
        L.cell_off = cells; cells += L.ncols * L.nrows;
        L.quota = ex->quota[l]; if (L.quota > maxN) maxN = L.quota;
        const int area = (L.w - 2 * EDGE) * (L.h - 2 * EDGE);
        L.raw_cap = area / 4 + area / 32 + 256;   // bound on cell-local maxima of the score map; overflow is reported, never truncated
        L.raw_off = raw; raw += (L.raw_cap + 3) & ~3;
        // quadtree roots, :1010-1012
        L.nini = (int)roundf((float)(maxBX - minB) / (float)(maxBY - minB));
        if (L.nini < 1) { set_last_error("portrait frames with aspect < 0.5 are unsupported (reference divides by zero)"); return UVIP_ERR_UNSUPPORTED; }
        L.hx = (float)(maxBX - minB) / (float)L.nini;
        L.kp_cap = (L.quota + 4 > 4 * L.nini ? L.quota + 4 : 4 * L.nini);
        L.kp_off = kp; kp += L.kp_cap;
        L.fntx = div_up(L.ncols, FAST_CW); L.fnty = div_up(L.nrows, FAST_CH);
        L.ftile_off = ft; ft += L.fntx * L.fnty;
        { const int tw = (L.ncols < FAST_CW ? L.ncols : FAST_CW) * L.wcell, th = (L.nrows < FAST_CH ? L.nrows : FAST_CH) * L.hcell;
          if (tw > max_tw) max_tw = tw; if (th > max_th) max_th = th; }
        L.bntx = div_up(L.w, BL_W); L.bnty = div_up(L.h, BL_H);
        L.btile_off = bt; bt += L.bntx * L.bnty + 1;            // + 1 ring tile
        L.tab_off = tab; if (l > 0) tab += 2 * L.w + 2 * L.h;
        L.scale = ex->scale[l];
        L.size = (float)(int)(31 * ex->scale[l]);               // :820
    }
    P.frame_bytes = off; P.cells_per_frame = cells; P.raw_per_frame = raw; P.kp_per_frame = kp;
    P.ftiles = ft; P.btiles = bt;
    int nc = 64; while (nc < maxN + 8 || nc < 16) nc <<= 1;
    for (int l = 0; l < p.nlevels; l++) while (nc < 4 * P.lv[l].nini + 4) nc <<= 1;
    if (nc > 4096) { set_last_error("per-level quota %d exceeds the quadtree node capacity", maxN); return UVIP_ERR_UNSUPPORTED; }
    P.node_cap = nc;
    {
        const int ngx_max = (max_tw + 2) / 4 + 1;
        P.f_irow = (int)align_up((size_t)4 * (ngx_max + 2) + 12, 16); P.f_irows = max_th + 6;   // TMA box: 16-byte aligned start and extent
        if (P.f_irow <= F2_IROW) P.f_irow = F2_IROW;           // the common case (cells up to 33 px wide) runs k_fast2 with a compile-time row stride
        P.f_srow = (int)align_up((size_t)max_tw + 2 + FAST_CW, 4); P.f_srows = max_th + 2 + FAST_CH;   // one zero row / column between cells
        P.f_rcp_srow = 0xFFFFFFFFu / (unsigned)P.f_srow + 1u;
        P.f_gw = 32 * div_up(ngx_max * max_th, 32 * FAST_WARPS);           // groups one warp can meet
        if ((size_t)P.f_srow * P.f_srows >= 0xFFFFu) { set_last_error("FAST tile does not fit 16-bit positions"); return UVIP_ERR_UNSUPPORTED; }
        // k_fast2: one warp per cell
        int max_wc = 0, max_hc = 0, gcap = 0;
        for (int l = 0; l < p.nlevels; l++) {
            const LevelInfo& L = P.lv[l];
            if (L.wcell > max_wc) max_wc = L.wcell;
            if (L.hcell > max_hc) max_hc = L.hcell;
            int ngx = (L.wcell + 6) / 4 + 1; if (ngx < 2) ngx = 2;       // any alignment of the cell inside its first / last group
            if (ngx * L.hcell > gcap) gcap = ngx * L.hcell;
        }
        P.f2_srow = (int)align_up((size_t)max_wc + 2, 4); P.f2_srows = max_hc + 2;
        P.f2_score_bytes = (int)align_up((size_t)P.f2_srow * P.f2_srows, 16);
        P.f2_gcap = (int)align_up((size_t)gcap + 2, 8);
        P.f2_wbytes = (int)align_up((size_t)P.f2_score_bytes + 2 * ((size_t)P.f2_gcap + F2_PQ + 2) + 4 * 20, 16);
        P.f2_irows = max_hc + 6;
        P.f2_imgbytes = P.f_irow * P.f2_irows; P.f2_imgstride = (int)align_up((size_t)P.f2_imgbytes, 128);
        int f2t = 0;
        for (int l = 0; l < p.nlevels; l++) f2t += div_up(P.lv[l].ncols, F2_CW) * P.lv[l].nrows;
        P.f2_tiles = f2t;
        P.f2_rcp_tiles = f2t > 1 ? 0xFFFFFFFFu / (unsigned)f2t + 1u : 0u;
        if ((long long)f2t * p.max_batch >= (1LL << 30)) { set_last_error("too many FAST tiles for one launch group"); return UVIP_ERR_UNSUPPORTED; }
    }
    P.rcp_cpr = 0xFFFFFFFFu / (unsigned)((P.lv[0].w + 15) >> 4) + 1u;
    {
        // source extent of a 128x64 output tile: exact maximum over all tiles of all levels (tables are exact)
        int bw = 16, bh = 2;
        std::vector<int> ofs;
        for (int l = 1; l < p.nlevels; l++) {
            const LevelInfo& D = P.lv[l]; const LevelInfo& S = P.lv[l - 1];
            ofs.assign(2 * (D.w > D.h ? D.w : D.h), 0);
            build_axis_table(S.w, D.w, ofs.data(), ofs.data() + D.w);
            for (int X0 = 0; X0 < D.w; X0 += RS_W) {
                const int X1 = (X0 + RS_W < D.w ? X0 + RS_W : D.w) - 1;
                const int bx = (ofs[X0] + EDGE) & ~15;
                const int need = ofs[X1] + EDGE + 12 - bx;          // a thread reads the 12-byte window from its column 0's aligned word
                if (need > bw) bw = need;
            }
            // window offsets of the 4 columns of every group (independent of the tile: the box origin is 16-byte aligned)
            int narrow = 1;
            for (int x = 0; x < D.w; x += 4) {
                const int base = ofs[x] - ((ofs[x] + EDGE) & 3);
                for (int k = 0; k < 4; k++) {
                    const int d = ofs[x + k < D.w ? x + k : D.w - 1] - base;
                    if (d + 1 > 11) { set_last_error("scale factor too large for the resize window"); return UVIP_ERR_UNSUPPORTED; }
                    if (k < 3 && d + 1 > 7) narrow = 0;
                }
            }
            if (narrow) P.rs_narrow_mask |= 1u << l;
            build_axis_table(S.h, D.h, ofs.data(), ofs.data() + D.h);
            for (int Y0 = 0; Y0 < D.h; Y0 += RS_H) {
                const int Y1 = (Y0 + RS_H < D.h ? Y0 + RS_H : D.h) - 1;
                const int need = ofs[Y1] + 2 - ofs[Y0];
                if (need > bh) bh = need;
            }
        }
        P.rs_boxw = (int)align_up((size_t)bw + 4, 16); P.rs_boxh = bh;
        if (P.rs_boxw > 256 || P.rs_boxh > 256) { set_last_error("scale factor too large for the resize tile"); return UVIP_ERR_UNSUPPORTED; }
    }
    P.tile_tab_off = tab;
    P.f2_tab_off = tab + ft + 1 + bt;
    if (tabs) {
        tabs->assign(tab + ft + 1 + bt + P.f2_tiles + 1, 0);
        {   // k_fast2 tiles: level | cell row << 4 | first cell column << 16, level-major (the large levels of a frame first)
            int* t = tabs->data() + P.f2_tab_off;
            for (int l = 0; l < p.nlevels; l++)
                for (int cy = 0; cy < P.lv[l].nrows; cy++)
                    for (int tx = 0; tx < div_up(P.lv[l].ncols, F2_CW); tx++) *t++ = (int)((unsigned)l | ((unsigned)cy << 4) | ((unsigned)(tx * F2_CW) << 16));
        }
        for (int l = 0; l < p.nlevels; l++) {
            const LevelInfo& L = P.lv[l];
            for (int ty = 0; ty < L.fnty; ty++)
                for (int tx = 0; tx < L.fntx; tx++)
                    (*tabs)[tab + L.ftile_off + ty * L.fntx + tx] = (int)((unsigned)l | ((unsigned)(ty * FAST_CH) << 4) | ((unsigned)(tx * FAST_CW) << 16));
        }
        for (int l = 0; l < p.nlevels; l++) {                   // blur tiles: level | ty << 4 | tx << 16, bit 31 = the level's ring tile
            const LevelInfo& L = P.lv[l];
            int* t = tabs->data() + tab + ft + 1 + L.btile_off;
            for (int ty = 0; ty < L.bnty; ty++)
                for (int tx = 0; tx < L.bntx; tx++) t[ty * L.bntx + tx] = (int)((unsigned)l | ((unsigned)ty << 4) | ((unsigned)tx << 16));
            t[L.bntx * L.bnty] = (int)((unsigned)l | 0x80000000u);
        }
        for (int l = 1; l < p.nlevels; l++) {
            const LevelInfo& D = P.lv[l]; const LevelInfo& S = P.lv[l - 1];
            int* t = tabs->data() + D.tab_off;
            build_axis_table(S.w, D.w, t, t + D.w);
            build_axis_table(S.h, D.h, t + 2 * D.w, t + 2 * D.w + D.h);
        }
    }
    *out = P;
    return UVIP_OK;
}

static size_t fast2_smem_bytes(const Plan& P, int irow)
{
    const size_t imgstride = align_up((size_t)irow * P.f2_irows, 128);
    return F2_STAGES * imgstride + (size_t)FAST_WARPS * P.f2_wbytes;
}
static size_t fast_smem_bytes(const Plan& P) { return (size_t)P.f_irow * P.f_irows + (size_t)P.f_srow * P.f_srows + 2 * (size_t)FAST_WARPS * 5 * P.f_gw + 128; }
static size_t qt_smem_bytes(int cap) { return (size_t)(4 * cap * 2 + cap * 2 + 4 * cap + cap + cap + (cap + 1) + cap + 4 * cap + cap + cap + cap) * 4; }

static int ensure_plan(uvip_extractor* ex, int w, int h)
{
    if (ex->plan.W == w && ex->plan.H == h) return UVIP_OK;
    cudaDeviceGetAttribute(&ex->num_sms, cudaDevAttrMultiProcessorCount, ex->device);
    UVIP_CHECK_ARG(w <= ex->prm.max_width && h <= ex->prm.max_height);
    Plan P; std::vector<int> tabs;
    int rc = make_plan(ex, w, h, &P, &tabs);
    if (rc) return rc;
    if (P.frame_bytes > ex->cap_frame_bytes || P.cells_per_frame > ex->cap_cells || P.raw_per_frame > ex->cap_raw ||
        P.kp_per_frame > ex->cap_kp || (int)tabs.size() > ex->cap_tab) {
        set_last_error("frame %dx%d needs a larger working set than max_width x max_height provides", w, h);
        return UVIP_ERR_UNSUPPORTED;
    }
    // a new geometry rewrites the tables and tensor maps every queued launch group reads: nothing of this handle may still be running,
    // on its own streams or on a caller's (device entry point)
    if (ex->inflight) { set_last_error("frame geometry changes to %dx%d while a submitted batch is in flight: wait for its ticket first", w, h); return UVIP_ERR_ARG; }
    UVIP_CUDA(cudaDeviceSynchronize());
    UVIP_CUDA(cudaFuncSetAttribute(k_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qt_smem_bytes(P.node_cap)));
    UVIP_CUDA(cudaFuncSetAttribute(k_describe, cudaFuncAttributeMaxDynamicSharedMemorySize, DESC_SMEM));
    UVIP_CUDA(cudaFuncSetAttribute(k_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes(P)));
    {
        const size_t sm2 = fast2_smem_bytes(P, P.f_irow);
        int nb = 0;
        if (P.f_irow == F2_IROW) {
            UVIP_CUDA(cudaFuncSetAttribute(k_fast2<F2_IROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            UVIP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_fast2<F2_IROW>, 256, sm2));
        } else {
            UVIP_CUDA(cudaFuncSetAttribute(k_fast2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            UVIP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_fast2<0>, 256, sm2));
        }
        if (nb < 1) { set_last_error("FAST kernel does not fit one SM (%zu bytes of shared memory)", sm2); return UVIP_ERR_UNSUPPORTED; }
        if (const char* e = getenv("UVIP_FAST2_CTAS")) { const int v = atoi(e); if (v >= 1 && v < nb) nb = v; }   // tuning: leave room for co-running kernels
        ex->fast2_ctas_per_sm = nb;
    }
    UVIP_CUDA(cudaFuncSetAttribute(k_resize<true, RS_RMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.rs_boxw * P.rs_boxh + 128));
    UVIP_CUDA(cudaFuncSetAttribute(k_resize<false, RS_RMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.rs_boxw * P.rs_boxh + 128));
    UVIP_CUDA(cudaFuncSetAttribute(k_resize<true, RS_RSMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.rs_boxw * P.rs_boxh + 128));
    UVIP_CUDA(cudaFuncSetAttribute(k_resize<false, RS_RSMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.rs_boxw * P.rs_boxh + 128));
    // TMA descriptors of the pyramid planes: (x bytes, rows, frame) with a box of one FAST tile.  Everything is built in host
    // temporaries first; tables, tensor maps and the plan are committed together once every encode has succeeded.
    {
        typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fp = nullptr; cudaDriverEntryPointQueryResult qres;
        UVIP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
        if (!fp || qres != cudaDriverEntryPointSuccess) { set_last_error("cuTensorMapEncodeTiled is not available in this driver"); return UVIP_ERR_CUDA; }
        alignas(64) CUtensorMap maps[4 * MAXLEV];         // FAST (first kernel) tile boxes | blur tile boxes | resize source boxes | k_fast2 tile boxes
        memset(maps, 0, sizeof(maps));
        for (int l = 0; l < P.nlevels; l++) {
            const LevelInfo& L = P.lv[l];
            const cuuint64_t gdim[3] = {(cuuint64_t)L.pstride, (cuuint64_t)(L.h + 2 * EDGE), (cuuint64_t)ex->prm.max_batch};
            const cuuint64_t gstr[2] = {(cuuint64_t)L.pstride, (cuuint64_t)P.frame_bytes};
            const cuuint32_t box[3] = {(cuuint32_t)P.f_irow, (cuuint32_t)P.f_irows, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = ((encode_fn)fp)(&maps[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ex->pyr.as<uint8_t>() + L.poff, gdim, gstr, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled(level %d) failed: %d", l, (int)r); return UVIP_ERR_CUDA; }
            const cuuint32_t bbox[3] = {(cuuint32_t)BL_BOXW, (cuuint32_t)BL_BOXH, 1};
            r = ((encode_fn)fp)(&maps[MAXLEV + l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ex->pyr.as<uint8_t>() + L.poff, gdim, gstr, bbox, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled(blur, level %d) failed: %d", l, (int)r); return UVIP_ERR_CUDA; }
            const cuuint32_t rbox[3] = {(cuuint32_t)P.rs_boxw, (cuuint32_t)P.rs_boxh, 1};
            r = ((encode_fn)fp)(&maps[2 * MAXLEV + l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ex->pyr.as<uint8_t>() + L.poff, gdim, gstr, rbox, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled(resize, level %d) failed: %d", l, (int)r); return UVIP_ERR_CUDA; }
            const cuuint32_t f2box[3] = {(cuuint32_t)P.f_irow, (cuuint32_t)P.f2_irows, 1};
            r = ((encode_fn)fp)(&maps[3 * MAXLEV + l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ex->pyr.as<uint8_t>() + L.poff, gdim, gstr, f2box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled(FAST, level %d) failed: %d", l, (int)r); return UVIP_ERR_CUDA; }
        }
        ex->plan.W = 0;                                        // from here on the old plan is gone; a failed copy leaves "no plan"
        UVIP_CUDA(cudaMemcpy(ex->tabs.p, tabs.data(), tabs.size() * sizeof(int), cudaMemcpyHostToDevice));
        UVIP_CUDA(cudaMemcpy(ex->tmaps.p, maps, sizeof(maps), cudaMemcpyHostToDevice));
    }
    ex->plan = P;
    ex->ghave = false;                                         // a captured single-frame graph embeds the old plan
    return UVIP_OK;
}

// enqueue the whole pipeline for nframes frames on st
static int enqueue_group(uvip_extractor* ex, const uint8_t* d_frames, int nframes, int stride, size_t frame_pitch,
                         uvip_keypoint* d_kps, int32_t* d_n_out, int out_cap, uint8_t* d_desc,
                         int full_detect, int n_incoming, int grid_rows, int grid_cols, int min_px_dist, int num_needed, cudaStream_t st,
                         bool reset_status = true, const int* d_dyn = nullptr)
{
    const Plan& P = ex->plan;
    uint8_t* pyr = ex->pyr.as<uint8_t>(); uint8_t* blur = ex->blur.as<uint8_t>();
    int* counters = ex->counters.as<int>();
    const size_t cstride = (size_t)ex->prm.max_batch * P.nlevels;
    int* cand_count = counters; int* win_count = counters + cstride;
    UVIP_CUDA(cudaMemsetAsync(counters, 0, 2 * cstride * sizeof(int), st));
    if (reset_status) UVIP_CUDA(cudaMemsetAsync(ex->status.p, 0, sizeof(int), st));
    cudaEvent_t* pe = ex->prof ? ex->prof_ev.data() + (size_t)(ex->prof_groups % PROF_RING) * (UVIP_NUM_STAGES + 1) : nullptr;
#define PROF_MARK(i) do { if (pe) UVIP_CUDA(cudaEventRecord(pe[i], st)); } while (0)
    NvtxRange nv_group("uvip_extract_group");
    PROF_MARK(0);
    {
        NvtxRange nv("uvip_pyramid");
        const LevelInfo& L = P.lv[0];
        const int chunks = ((L.w + 15) >> 4) * L.h;
        const int vec_ok = (((uintptr_t)d_frames & 15) == 0 && (stride & 15) == 0 && (frame_pitch & 15) == 0) ? 1 : 0;
        const int copy_blocks = div_up(chunks, 256);
        k_import<<<dim3(copy_blocks + 2 * BORDER_W, nframes), 256, 0, st>>>(d_frames, stride, frame_pitch, vec_ok, copy_blocks, pyr, P);
        ex->launches++;
    }
    for (int l = 1; l < P.nlevels; l++) {
        NvtxRange nv("uvip_pyramid");
        const LevelInfo& L = P.lv[l];
        // half-height tiles when full ones would not even fill one wave (6 CTAs per SM): twice the CTAs for a lone frame or a small
        // batch (single-frame pyramid 0.063 -> 0.054 ms).  At batch 256 every level stays on full tiles: half tiles for the 1.2-2.6-wave
        // levels 4-7 were measured slower (pyramid 0.308 -> 0.316 ms, and FAST behind them 0.759 -> 0.794 ms with identical code: every
        // CTA loads the full-height box, and the extra L2 traffic evicts pyramid planes FAST is about to read)
        const bool narrow = (P.rs_narrow_mask >> l) & 1u;
        const bool small = (long long)div_up(L.w, RS_W) * div_up(L.h, RS_H) * nframes < 6LL * ex->num_sms;
        const int th = 8 * (small ? RS_RSMALL : RS_RMAX);
        auto kern = small ? (narrow ? k_resize<true, RS_RSMALL> : k_resize<false, RS_RSMALL>) : (narrow ? k_resize<true, RS_RMAX> : k_resize<false, RS_RMAX>);
        kern<<<dim3(div_up(L.w, RS_W), div_up(L.h, th), nframes), 256, (size_t)P.rs_boxw * P.rs_boxh + 128, st>>>(
            ex->tmaps.as<CUtensorMap>(), pyr, ex->tabs.as<int>(), l, P);
        ex->launches++;
    }
    PROF_MARK(1);
    {
    NvtxRange nv("uvip_fast");
    if (ex->use_fast1)
        k_fast<<<dim3(P.ftiles, nframes), 256, fast_smem_bytes(P), st>>>(ex->tmaps.as<CUtensorMap>(), ex->tabs.as<unsigned>() + P.tile_tab_off, ex->cand.as<unsigned>(),
                                                                         cand_count, ex->status.as<int>(), P);
    else {
        // persistent CTAs: as many as stay resident, each walking the (tile, frame) space with stride gridDim.x
        // (every CTA stages F2_STAGES tiles at once: a grid larger than T / F2_STAGES would only draw empty tickets)
        const long long T = (long long)P.f2_tiles * nframes, res = (long long)ex->fast2_ctas_per_sm * ex->num_sms;
        const long long want = (T + F2_STAGES - 1) / F2_STAGES;
        const unsigned grid = (unsigned)(want < res ? want : res);
        if (P.f_irow == F2_IROW)
            k_fast2<F2_IROW><<<grid, 256, fast2_smem_bytes(P, P.f_irow), st>>>(ex->tmaps.as<CUtensorMap>(), ex->tabs.as<unsigned>() + P.f2_tab_off,
                                                                               ex->cand.as<unsigned>(), cand_count, ex->status.as<int>(), nframes, P);
        else
            k_fast2<0><<<grid, 256, fast2_smem_bytes(P, P.f_irow), st>>>(ex->tmaps.as<CUtensorMap>(), ex->tabs.as<unsigned>() + P.f2_tab_off,
                                                                         ex->cand.as<unsigned>(), cand_count, ex->status.as<int>(), nframes, P);
    }
    ex->launches++;
    }
    PROF_MARK(2);
    {
    NvtxRange nv("uvip_quadtree");
    k_quadtree<<<dim3(nframes, P.nlevels), QT_THREADS, qt_smem_bytes(P.node_cap), st>>>(
        ex->cand.as<unsigned>(), cand_count, ex->labels.as<unsigned short>(), ex->winners.as<unsigned>(), win_count,
        ex->status.as<int>(), P);
    ex->launches++;
    }
    PROF_MARK(3);
    {
    NvtxRange nv("uvip_blur");
    k_blur<<<dim3(P.btiles, nframes), 256, 0, st>>>(ex->tmaps.as<CUtensorMap>(), ex->tabs.as<unsigned>() + P.tile_tab_off + P.ftiles + 1, pyr, blur, P);
    ex->launches++;
    if (!full_detect && n_incoming > 0) {                      // incoming level-0 keypoints may sit inside the 16-px border zone
        k_ring16<<<8, 256, 0, st>>>(pyr, blur, P);
        ex->launches++;
    }
    }
    PROF_MARK(4);
    {
    NvtxRange nv("uvip_select");
    k_select<<<nframes, 256, 0, st>>>(ex->winners.as<unsigned>(), win_count, ex->sel.as<unsigned>(), ex->nsel.as<int>(), ex->sel_cap,
                                       full_detect, n_incoming, ex->grid.as<int32_t>(), grid_rows, grid_cols, min_px_dist, num_needed,
                                       d_dyn, ex->status.as<int>(), P);
    ex->launches++;
    }
    const int slots = out_cap < ex->sel_cap ? out_cap : ex->sel_cap;
    PROF_MARK(5);
    {
    NvtxRange nv("uvip_describe");
    k_describe<<<dim3(div_up(slots, DESC_WARPS * DESC_KPW), nframes), DESC_WARPS * 32, DESC_SMEM, st>>>(
        pyr, blur, ex->winners.as<unsigned>(), ex->sel.as<unsigned>(), ex->nsel.as<int>(), ex->sel_cap,
        ex->incoming.as<uvip_keypoint>(), ex->pat_t.as<float2>(), d_kps, d_desc, d_n_out, out_cap, ex->status.as<int>(), P);
    ex->launches++;
    }
    PROF_MARK(6);
#undef PROF_MARK
    if (pe) ex->prof_groups++;
    UVIP_CUDA(cudaGetLastError());
    ex->last_frames = nframes;
    return UVIP_OK;
}

static int read_status(uvip_extractor* ex, cudaStream_t st)
{
    int s = 0;
    UVIP_CUDA(cudaMemcpyAsync(&s, ex->status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    if (s) {
        set_last_error("device capacity overflow, flags 0x%x (1 FAST candidates, 2 quadtree nodes, 4 winners, 8 output rows)", s);
        return UVIP_ERR_CAPACITY;
    }
    return UVIP_OK;
}

extern "C" {

int uvip_extractor_create(const uvip_extractor_params* params, uvip_extractor** out)
{
    UVIP_CHECK_ARG(params && out);
    *out = nullptr;
    UVIP_CHECK_ARG(params->nlevels >= 1 && params->nlevels <= MAXLEV && params->nfeatures >= 1);
    UVIP_CHECK_ARG(params->scale_factor > 1.0f && params->max_batch >= 1 && params->max_width >= 64 && params->max_height >= 64);
    UVIP_CHECK_ARG(params->fast_th >= 0);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_last_error("no CUDA device available; libuvip_orb has no CPU fallback");
        return UVIP_ERR_NO_DEVICE;
    }
    UVIP_CHECK_ARG(params->device >= 0 && params->device < ndev);
    DeviceGuard g(params->device);
    uvip_extractor* ex = new uvip_extractor();
    ex->prm = *params;
    if (ex->prm.retry_th <= 0) ex->prm.retry_th = 7;
    if (ex->prm.cell <= 0) ex->prm.cell = 30;
    ex->device = params->device;
    memset(&ex->plan, 0, sizeof(ex->plan));
    // constructor tables, src/ORBextractor.cc:458-512
    const double scaleFactor = (double)params->scale_factor;
    const int nl = params->nlevels;
    ex->scale[0] = 1.0f;
    for (int i = 1; i < nl; i++) ex->scale[i] = (float)((double)ex->scale[i - 1] * scaleFactor);
    const float invScaleFactor = (float)(1.0f / scaleFactor);
    ex->inv_scale[0] = 1.0f;
    for (int i = 1; i < nl; i++) ex->inv_scale[i] = ex->inv_scale[i - 1] * invScaleFactor;
    const float factor = (float)(1.0 / scaleFactor);
    float nDesired = params->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) { ex->quota[l] = cv_round_f(nDesired); sum += ex->quota[l]; nDesired *= factor; }
    ex->quota[nl - 1] = params->nfeatures - sum > 0 ? params->nfeatures - sum : 0;
    {
        int v, v0;
        const int vmax = (int)floor(HALF_PATCH * sqrtf(2.f) / 2 + 1), vmin = (int)ceil(HALF_PATCH * sqrtf(2.f) / 2);
        const double hp2 = HALF_PATCH * HALF_PATCH;
        for (v = 0; v <= vmax; ++v) ex->umax[v] = (int)lrint(sqrt(hp2 - v * v));
        for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) { while (ex->umax[v0] == ex->umax[v0 + 1]) ++v0; ex->umax[v] = v0; ++v0; }
    }
    // capacity from the maximal frame
    Plan P;
    int rc = make_plan(ex, params->max_width, params->max_height, &P, nullptr);
    if (rc) { delete ex; return rc; }
    const int B = params->max_batch;
    ex->cap_frame_bytes = P.frame_bytes; ex->cap_cells = P.cells_per_frame; ex->cap_raw = P.raw_per_frame; ex->cap_kp = P.kp_per_frame;
    int tab = 0; for (int l = 1; l < nl; l++) tab += 2 * P.lv[l].w + 2 * P.lv[l].h;
    ex->cap_tab = tab + 2 * (P.ftiles + P.btiles + P.f2_tiles) + 4096;      // resize tables + tile tables (+ slack for other aspect ratios)
    ex->sel_cap = P.kp_per_frame + 4096;     // output rows per frame: quadtree winners + up to 4096 incoming keypoints
    cudaError_t e = cudaStreamCreateWithFlags(&ex->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_last_error("cudaStreamCreate -> %s", cudaGetErrorString(e)); delete ex; return UVIP_ERR_CUDA; }
    rc = 0;
    rc |= ex->pyr.reserve((size_t)B * P.frame_bytes + 4096);
    rc |= ex->blur.reserve((size_t)B * P.frame_bytes + 4096);
    rc |= ex->cand.reserve((size_t)B * P.raw_per_frame * 4);
    rc |= ex->labels.reserve((size_t)B * P.raw_per_frame * 2);
    rc |= ex->winners.reserve((size_t)B * P.kp_per_frame * 4);
    rc |= ex->counters.reserve((size_t)2 * B * nl * 4);
    rc |= ex->sel.reserve((size_t)B * ex->sel_cap * 4);
    rc |= ex->nsel.reserve((size_t)B * 4);
    rc |= ex->tabs.reserve((size_t)ex->cap_tab * 4);
    rc |= ex->status.reserve(16);
    rc |= ex->grid.reserve(16);
    rc |= ex->incoming.reserve(sizeof(uvip_keypoint));
    rc |= ex->tmaps.reserve(sizeof(CUtensorMap) * 4 * MAXLEV);
    rc |= ex->pat_t.reserve(sizeof(float) * 2 * 512);
    if (rc) { uvip_extractor_destroy(ex); return UVIP_ERR_CUDA; }
    // zero the planes once so halo loads never see uninitialised memory
    cudaMemset(ex->pyr.p, 0, ex->pyr.cap); cudaMemset(ex->blur.p, 0, ex->blur.cap);
    cudaMemset(ex->status.p, 0, 16);
    {   // pattern transposed for the descriptor kernel: entry [k][lane] = point 16*lane + k as (x, y) floats
        std::vector<float> pt(2 * 512);
        for (int lane = 0; lane < 32; lane++)
            for (int k = 0; k < 16; k++) {
                pt[2 * (k * 32 + lane)] = (float)h_pattern[2 * (16 * lane + k)];
                pt[2 * (k * 32 + lane) + 1] = (float)h_pattern[2 * (16 * lane + k) + 1];
            }
        cudaMemcpy(ex->pat_t.p, pt.data(), pt.size() * sizeof(float), cudaMemcpyHostToDevice);
        for (int uu = 0; uu <= HALF_PATCH; uu++)               // the descriptor kernel relies on the disc table being symmetric
            for (int vv = 0; vv <= HALF_PATCH; vv++)
                if ((uu <= ex->umax[vv]) != (vv <= ex->umax[uu])) { set_last_error("umax table is not symmetric"); uvip_extractor_destroy(ex); return UVIP_ERR_UNSUPPORTED; }
    }
    {
        unsigned rcp[33]; rcp[0] = rcp[1] = 0;
        for (int n = 2; n <= 32; n++) rcp[n] = 0xFFFFFFFFu / (unsigned)n + 1u;
        if (cudaMemcpyToSymbol(c_rcp32, rcp, sizeof(rcp)) != cudaSuccess) { set_last_error("cudaMemcpyToSymbol(c_rcp32) failed"); uvip_extractor_destroy(ex); return UVIP_ERR_CUDA; }
    }
    { const char* e1 = getenv("UVIP_FAST1"); ex->use_fast1 = e1 && e1[0] == '1'; }
    { const char* e2 = getenv("UVIP_SUBBATCH"); if (e2) ex->subbatch = atoi(e2) > 0 ? atoi(e2) : 0; }
    if (cudaMemcpyToSymbol(c_pattern, h_pattern, sizeof(h_pattern)) != cudaSuccess ||
        cudaMemcpyToSymbol(c_umax, ex->umax, sizeof(ex->umax)) != cudaSuccess) {
        set_last_error("cudaMemcpyToSymbol failed: %s", cudaGetErrorString(cudaGetLastError()));
        uvip_extractor_destroy(ex); return UVIP_ERR_CUDA;
    }
    const size_t qsm = qt_smem_bytes(P.node_cap);
    if (cudaFuncSetAttribute(k_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsm) != cudaSuccess) {
        set_last_error("quadtree kernel needs %zu bytes of shared memory: %s", qsm, cudaGetErrorString(cudaGetLastError()));
        uvip_extractor_destroy(ex); return UVIP_ERR_UNSUPPORTED;
    }
    if (cudaFuncSetAttribute(k_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes(P)) != cudaSuccess) {
        set_last_error("FAST kernel needs %zu bytes of shared memory: %s", fast_smem_bytes(P), cudaGetErrorString(cudaGetLastError()));
        uvip_extractor_destroy(ex); return UVIP_ERR_UNSUPPORTED;
    }
    *out = ex;
    return UVIP_OK;
}

int uvip_extractor_destroy(uvip_extractor* ex)
{
    if (!ex) return UVIP_OK;
    DeviceGuard g(ex->device);
    if (ex->stream) cudaStreamSynchronize(ex->stream);
    DevBuf* bufs[] = {&ex->pyr, &ex->blur, &ex->cand, &ex->labels, &ex->winners, &ex->counters, &ex->sel,
                      &ex->nsel, &ex->tabs, &ex->status, &ex->grid, &ex->incoming, &ex->tmaps, &ex->pat_t, &ex->in_frames, &ex->out_kps, &ex->out_desc, &ex->out_n,
                      &ex->in2, &ex->kps2, &ex->desc2, &ex->n2, &ex->clahe_lut, &ex->clahe_io, &ex->dyn,
                      &ex->knn_i[0], &ex->knn_i[1], &ex->knn_d[0], &ex->knn_d[1]};
    for (DevBuf* b : bufs) b->release();
    for (cudaEvent_t e : ex->prof_ev) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) { if (ex->ev_h2d[i]) cudaEventDestroy(ex->ev_h2d[i]); if (ex->ev_comp[i]) cudaEventDestroy(ex->ev_comp[i]); if (ex->ev_d2h[i]) cudaEventDestroy(ex->ev_d2h[i]); }
    for (int i = 0; i < 2; i++) if (ex->ev_done[i]) cudaEventDestroy(ex->ev_done[i]);
    if (ex->h_status) cudaFreeHost(ex->h_status);
    if (ex->h_dyn) cudaFreeHost(ex->h_dyn);
    if (ex->gexec) cudaGraphExecDestroy(ex->gexec);
    if (ex->h_stage) cudaFreeHost(ex->h_stage);
    if (ex->h2d_stream) cudaStreamDestroy(ex->h2d_stream);
    if (ex->d2h_stream) cudaStreamDestroy(ex->d2h_stream);
    if (ex->stream) cudaStreamDestroy(ex->stream);
    delete ex;
    return UVIP_OK;
}

int uvip_extractor_levels(const uvip_extractor* ex) { return ex ? ex->prm.nlevels : 0; }
float uvip_extractor_scale_factor(const uvip_extractor* ex) { return ex ? (float)(double)ex->prm.scale_factor : 0.f; }
long long uvip_extractor_launch_count(const uvip_extractor* ex) { return ex ? ex->launches : 0; }
long long uvip_extractor_graph_captures(const uvip_extractor* ex) { return ex ? ex->gcaptures : 0; }
void* uvip_extractor_stream(uvip_extractor* ex) { return ex ? (void*)ex->stream : nullptr; }

int uvip_extractor_tables(const uvip_extractor* ex, float* scale, float* inv_scale, int32_t* quota, int32_t* umax)
{
    UVIP_CHECK_ARG(ex);
    for (int l = 0; l < ex->prm.nlevels; l++) {
        if (scale) scale[l] = ex->scale[l];
        if (inv_scale) inv_scale[l] = ex->inv_scale[l];
        if (quota) quota[l] = ex->quota[l];
    }
    if (umax) for (int v = 0; v <= HALF_PATCH; v++) umax[v] = ex->umax[v];
    return UVIP_OK;
}

int uvip_extract_batch_device(uvip_extractor* ex, const uint8_t* d_frames, int nframes, int w, int h, int stride,
                              size_t frame_pitch, uvip_keypoint* d_kps, int32_t* d_n_out, int cap, uint8_t* d_desc, void* stream)
{
    UVIP_CHECK_ARG(ex && d_frames && d_kps && d_n_out && d_desc);
    UVIP_CHECK_ARG(nframes >= 1 && nframes <= ex->prm.max_batch && w > 0 && h > 0 && stride >= w && cap >= 1);
    std::lock_guard<std::mutex> lk(ex->mu);
    // the launch group uses the handle's pyramid / candidate / selection scratch, which batches queued by uvip_extract_batch_submit
    // are still working on
    if (ex->inflight) { set_last_error("a submitted batch is in flight on this handle: wait for its ticket first"); return UVIP_ERR_ARG; }
    DeviceGuard g(ex->device);
    int rc = ensure_plan(ex, w, h);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : ex->stream;
    if (ex->subbatch <= 0 || ex->subbatch >= nframes)
        return enqueue_group(ex, d_frames, nframes, stride, frame_pitch, d_kps, d_n_out, cap, d_desc, 1, 0, 1, 1, 1, 0, st);
    // Sub-batches run through every stage before the next one starts and reuse the same scratch slots, so that a sub-batch's pyramid
    // and blurred planes are still in the 126 MB L2 when FAST, the blur and the descriptor stage read them (uvip_extractor_set_subbatch)
    for (int f0 = 0; f0 < nframes; f0 += ex->subbatch) {
        const int nb = nframes - f0 < ex->subbatch ? nframes - f0 : ex->subbatch;
        rc = enqueue_group(ex, d_frames + (size_t)f0 * frame_pitch, nb, stride, frame_pitch, d_kps + (size_t)f0 * cap, d_n_out + f0, cap,
                           d_desc + (size_t)f0 * cap * 32, 1, 0, 1, 1, 1, 0, st, f0 == 0);
        if (rc) return rc;
    }
    return UVIP_OK;
}

int uvip_extractor_set_subbatch(uvip_extractor* ex, int frames)
{
    UVIP_CHECK_ARG(ex && frames >= 0);
    std::lock_guard<std::mutex> lk(ex->mu);
    ex->subbatch = frames;
    return UVIP_OK;
}

int uvip_extractor_status(uvip_extractor* ex)
{
    UVIP_CHECK_ARG(ex);
    DeviceGuard g(ex->device);
    UVIP_CUDA(cudaDeviceSynchronize());
    return read_status(ex, ex->stream);
}

// Chunks of max_batch frames flow through a 3-stream pipeline: H2D of chunk c+1 and D2H of chunk c-1 overlap the kernels of
// chunk c (two staging sets; pinned host memory is needed for the copies to be truly asynchronous).  The chunk counter
// and the guard events live in the handle, so that two submitted batches overlap in the same way across calls.
static int submit_batch(uvip_extractor* ex, const uint8_t* frames, int nframes, int w, int h, int stride,
                        size_t frame_pitch, uvip_keypoint* kps, int32_t* n_out, int cap, uint8_t* desc, int* ticket_out,
                        uvip_matcher* mt = nullptr, int32_t* knn_idx = nullptr, int32_t* knn_dist = nullptr)
{
    if (ex->inflight >= 2) { set_last_error("two batches are already in flight: wait for a ticket first"); return UVIP_ERR_ARG; }
    DeviceGuard g(ex->device);
    int rc = ensure_plan(ex, w, h);
    if (rc) return rc;
    const int B = ex->prm.max_batch;
    const size_t fbytes = (size_t)stride * h;
    DevBuf* in[2] = {&ex->in_frames, &ex->in2}; DevBuf* ok[2] = {&ex->out_kps, &ex->kps2};
    DevBuf* od[2] = {&ex->out_desc, &ex->desc2}; DevBuf* on[2] = {&ex->out_n, &ex->n2};
    const int nchunks = (nframes + B - 1) / B;
    cudaStream_t st = ex->stream;
    bool grew = false;
    for (int s = 0; s < 2; s++) {
        const size_t need[4] = {(size_t)B * fbytes, (size_t)B * cap * sizeof(uvip_keypoint), (size_t)B * cap * 32, (size_t)B * 4};
        DevBuf* b[4] = {in[s], ok[s], od[s], on[s]};
        for (int i = 0; i < 4; i++) if (b[i]->cap < need[i]) grew = true;
    }
    if (grew) {                                                  // reallocation: nothing may still be using the old staging sets
        if (ex->inflight) { set_last_error("staging buffers must grow while a batch is in flight: wait first"); return UVIP_ERR_ARG; }
        UVIP_CUDA(cudaDeviceSynchronize());
        for (int s = 0; s < 2; s++) {
            if ((rc = in[s]->reserve((size_t)B * fbytes))) return rc;
            if ((rc = ok[s]->reserve((size_t)B * cap * sizeof(uvip_keypoint)))) return rc;
            if ((rc = od[s]->reserve((size_t)B * cap * 32))) return rc;
            if ((rc = on[s]->reserve((size_t)B * 4))) return rc;
        }
    }
    if (mt)
        for (int s = 0; s < 2; s++) {
            if (ex->knn_i[s].cap < (size_t)B * cap * 8 && ex->inflight) { set_last_error("kNN staging must grow while a batch is in flight: wait first"); return UVIP_ERR_ARG; }
            if ((rc = ex->knn_i[s].reserve((size_t)B * cap * 8))) return rc;
            if ((rc = ex->knn_d[s].reserve((size_t)B * cap * 8))) return rc;
        }
    if (!ex->h2d_stream) {
        UVIP_CUDA(cudaStreamCreateWithFlags(&ex->h2d_stream, cudaStreamNonBlocking));
        UVIP_CUDA(cudaStreamCreateWithFlags(&ex->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            UVIP_CUDA(cudaEventCreateWithFlags(&ex->ev_h2d[i], cudaEventDisableTiming));
            UVIP_CUDA(cudaEventCreateWithFlags(&ex->ev_comp[i], cudaEventDisableTiming));
            UVIP_CUDA(cudaEventCreateWithFlags(&ex->ev_d2h[i], cudaEventDisableTiming));
            UVIP_CUDA(cudaEventCreateWithFlags(&ex->ev_done[i], cudaEventDisableTiming));
        }
        UVIP_CUDA(cudaHostAlloc((void**)&ex->h_status, 2 * sizeof(int), cudaHostAllocDefault));
        ex->chunk_seq = 0;
    }
    cudaStream_t sh = ex->h2d_stream, sd = ex->d2h_stream;
    if (ex->inflight == 0) {
        UVIP_CUDA(cudaMemsetAsync(ex->status.p, 0, sizeof(int), st));
        // the single-frame entry point shares in_frames / out_* on the compute stream: order the copy streams behind it
        UVIP_CUDA(cudaEventRecord(ex->ev_done[0], st));
        UVIP_CUDA(cudaStreamWaitEvent(sh, ex->ev_done[0], 0));
        UVIP_CUDA(cudaStreamWaitEvent(sd, ex->ev_done[0], 0));
    }
    int s = 0;
    for (int c = 0; c < nchunks; c++, ex->chunk_seq++) {
        s = (int)(ex->chunk_seq & 1);
        const int f0 = c * B;
        const int nb = nframes - f0 < B ? nframes - f0 : B;
        if (ex->chunk_seq >= 2) UVIP_CUDA(cudaStreamWaitEvent(sh, ex->ev_comp[s], 0));   // kernels of chunk c-2 are done with in[s]
        if (frame_pitch == fbytes) UVIP_CUDA(cudaMemcpyAsync(in[s]->p, frames + (size_t)f0 * frame_pitch, (size_t)nb * fbytes, cudaMemcpyHostToDevice, sh));
        else UVIP_CUDA(cudaMemcpy2DAsync(in[s]->p, fbytes, frames + (size_t)f0 * frame_pitch, frame_pitch, fbytes - (stride - w), nb, cudaMemcpyHostToDevice, sh));
        UVIP_CUDA(cudaEventRecord(ex->ev_h2d[s], sh));
        UVIP_CUDA(cudaStreamWaitEvent(st, ex->ev_h2d[s], 0));
        if (ex->chunk_seq >= 2) UVIP_CUDA(cudaStreamWaitEvent(st, ex->ev_d2h[s], 0));    // results of chunk c-2 have left out[s]
        rc = enqueue_group(ex, in[s]->as<uint8_t>(), nb, stride, fbytes, ok[s]->as<uvip_keypoint>(), on[s]->as<int32_t>(),
                           cap, od[s]->as<uint8_t>(), 1, 0, 1, 1, 1, 0, st, false);
        if (rc) return rc;
        if (mt) {
            // consecutive-frame kNN2 on the descriptors where they lie (frame f = queries, frame f + 1 = train): the pairs inside this
            // chunk, and the pair across the chunk boundary, whose query frame is the last frame of the previous chunk — still intact in
            // the other staging set, which the compute stream only overwrites two chunks later.  Device rows: 0 = boundary pair,
            // 1 + j = pair (f0 + j, f0 + j + 1).
            int32_t* ki = ex->knn_i[s].as<int32_t>(); int32_t* kd = ex->knn_d[s].as<int32_t>();
            const size_t dp = (size_t)cap * 32, rp = (size_t)cap * 2;
            if (c > 0 && (rc = uvip_knn2_batch_device(mt, od[s ^ 1]->as<uint8_t>() + (size_t)(ex->prev_nb - 1) * dp, on[s ^ 1]->as<int32_t>() + (ex->prev_nb - 1), dp,
                                                      od[s]->as<uint8_t>(), on[s]->as<int32_t>(), dp, 1, cap, ki, kd, (size_t)cap, st))) return rc;
            if (nb > 1 && (rc = uvip_knn2_batch_device(mt, od[s]->as<uint8_t>(), on[s]->as<int32_t>(), dp, od[s]->as<uint8_t>() + dp, on[s]->as<int32_t>() + 1, dp,
                                                       nb - 1, cap, ki + rp, kd + rp, (size_t)cap, st))) return rc;
            const int r0 = c > 0 ? 0 : 1, nr = nb - r0;                  // device rows r0 .. nb-1 -> host pairs f0 - 1 + r0 ..
            UVIP_CUDA(cudaEventRecord(ex->ev_comp[s], st));
            UVIP_CUDA(cudaStreamWaitEvent(sd, ex->ev_comp[s], 0));
            if (nr > 0) {
                UVIP_CUDA(cudaMemcpyAsync(knn_idx + (size_t)(f0 - 1 + r0) * rp, ki + (size_t)r0 * rp, (size_t)nr * rp * 4, cudaMemcpyDeviceToHost, sd));
                UVIP_CUDA(cudaMemcpyAsync(knn_dist + (size_t)(f0 - 1 + r0) * rp, kd + (size_t)r0 * rp, (size_t)nr * rp * 4, cudaMemcpyDeviceToHost, sd));
            }
            ex->prev_nb = nb;
        } else {
            UVIP_CUDA(cudaEventRecord(ex->ev_comp[s], st));
            UVIP_CUDA(cudaStreamWaitEvent(sd, ex->ev_comp[s], 0));
        }
        UVIP_CUDA(cudaMemcpyAsync(n_out + f0, on[s]->p, (size_t)nb * 4, cudaMemcpyDeviceToHost, sd));
        UVIP_CUDA(cudaMemcpyAsync(kps + (size_t)f0 * cap, ok[s]->p, (size_t)nb * cap * sizeof(uvip_keypoint), cudaMemcpyDeviceToHost, sd));
        UVIP_CUDA(cudaMemcpyAsync(desc + (size_t)f0 * cap * 32, od[s]->p, (size_t)nb * cap * 32, cudaMemcpyDeviceToHost, sd));
        UVIP_CUDA(cudaEventRecord(ex->ev_d2h[s], sd));
    }
    const int t = ex->ticket_busy[0] ? 1 : 0;
    UVIP_CUDA(cudaMemcpyAsync(ex->h_status + t, ex->status.p, sizeof(int), cudaMemcpyDeviceToHost, sd));   // sd is behind the last chunk's kernels
    UVIP_CUDA(cudaEventRecord(ex->ev_done[t], sd));
    ex->ticket_busy[t] = true; ex->inflight++;
    *ticket_out = t;
    return UVIP_OK;
}

static int wait_batch(uvip_extractor* ex, int ticket)
{
    if (ticket < 0 || ticket > 1 || !ex->ticket_busy[ticket]) { set_last_error("ticket %d is not in flight", ticket); return UVIP_ERR_ARG; }
    DeviceGuard g(ex->device);
    UVIP_CUDA(cudaEventSynchronize(ex->ev_done[ticket]));
    ex->ticket_busy[ticket] = false; ex->inflight--;
    const int s = ex->h_status[ticket];
    if (s) {       // flags are sticky on the device until every batch in flight has reported them
        set_last_error("device capacity overflow, flags 0x%x (1 FAST candidates, 2 quadtree nodes, 4 winners, 8 output rows)", s);
        return UVIP_ERR_CAPACITY;
    }
    return UVIP_OK;
}

int uvip_extract_batch_submit(uvip_extractor* ex, const uint8_t* frames, int nframes, int w, int h, int stride,
                              size_t frame_pitch, uvip_keypoint* kps, int32_t* n_out, int cap, uint8_t* desc, int* ticket)
{
    UVIP_CHECK_ARG(ex && frames && kps && n_out && desc && ticket && nframes >= 1 && w > 0 && h > 0 && stride >= w && cap >= 1);
    UVIP_CHECK_ARG(frame_pitch >= (size_t)stride * (h - 1) + w);
    std::lock_guard<std::mutex> lk(ex->mu);
    return submit_batch(ex, frames, nframes, w, h, stride, frame_pitch, kps, n_out, cap, desc, ticket);
}

int uvip_extract_match_batch_submit(uvip_extractor* ex, uvip_matcher* m, const uint8_t* frames, int nframes, int w, int h, int stride,
                                    size_t frame_pitch, uvip_keypoint* kps, int32_t* n_out, int cap, uint8_t* desc,
                                    int32_t* knn_idx, int32_t* knn_dist, int* ticket)
{
    UVIP_CHECK_ARG(ex && m && frames && kps && n_out && desc && knn_idx && knn_dist && ticket && nframes >= 1 && w > 0 && h > 0 && stride >= w && cap >= 1);
    UVIP_CHECK_ARG(frame_pitch >= (size_t)stride * (h - 1) + w);
    std::lock_guard<std::mutex> lk(ex->mu);
    return submit_batch(ex, frames, nframes, w, h, stride, frame_pitch, kps, n_out, cap, desc, ticket, m, knn_idx, knn_dist);
}

int uvip_extract_batch_wait(uvip_extractor* ex, int ticket)
{
    UVIP_CHECK_ARG(ex);
    std::lock_guard<std::mutex> lk(ex->mu);
    return wait_batch(ex, ticket);
}

int uvip_extract_batch(uvip_extractor* ex, const uint8_t* frames, int nframes, int w, int h, int stride,
                       size_t frame_pitch, uvip_keypoint* kps, int32_t* n_out, int cap, uint8_t* desc)
{
    UVIP_CHECK_ARG(ex && frames && kps && n_out && desc && nframes >= 1 && w > 0 && h > 0 && stride >= w && cap >= 1);
    UVIP_CHECK_ARG(frame_pitch >= (size_t)stride * (h - 1) + w);
    std::lock_guard<std::mutex> lk(ex->mu);
    int t = 0;
    int rc = submit_batch(ex, frames, nframes, w, h, stride, frame_pitch, kps, n_out, cap, desc, &t);
    if (rc) return rc;
    return wait_batch(ex, t);
}

int uvip_extract(uvip_extractor* ex, const uint8_t* image, int w, int h, int stride,
                 uvip_keypoint* kps, int* n_inout, int cap, uint8_t* desc,
                 int32_t* grid, int grid_rows, int grid_cols, int min_px_dist, int full_detect, int num_needed)
{
    UVIP_CHECK_ARG(ex && n_inout);
    if (!image || w <= 0 || h <= 0) return UVIP_OK;               // empty image: outputs untouched (src/ORBextractor.cc:852-853)
    UVIP_CHECK_ARG(kps && desc && stride >= w && cap >= 1);
    const int n_in = (!full_detect && *n_inout > 0) ? *n_inout : 0;
    UVIP_CHECK_ARG(n_in <= 4096 && n_in <= cap);
    if (!full_detect) UVIP_CHECK_ARG(grid && grid_rows > 0 && grid_cols > 0 && min_px_dist > 0);
    // incoming keypoints become level-0 points (ComputeKeyPointsCopy, src/ORBextractor.cc:523-534): the reference indexes its level-0
    // buffer at (cvRound(y), cvRound(x)) unchecked — outside the image that is a read out of bounds there and a refused call here
    for (int i = 0; i < n_in; i++) {
        const float x = kps[i].x, y = kps[i].y;
        if (!(x == x) || !(y == y) || !(fabsf(x) < 1e9f) || !(fabsf(y) < 1e9f) || lrintf(x) < 0 || lrintf(x) >= w || lrintf(y) < 0 || lrintf(y) >= h) {
            set_last_error("incoming keypoint %d at (%g, %g) lies outside the %dx%d image", i, (double)x, (double)y, w, h);
            return UVIP_ERR_ARG;
        }
    }
    std::lock_guard<std::mutex> lk(ex->mu);
    if (ex->inflight) { set_last_error("a submitted batch is in flight on this handle: wait for its ticket first"); return UVIP_ERR_ARG; }
    DeviceGuard g(ex->device);
    int rc = ensure_plan(ex, w, h);
    if (rc) return rc;
    const size_t fbytes = (size_t)stride * h;
    if ((rc = ex->in_frames.reserve(fbytes))) return rc;
    if ((rc = ex->out_kps.reserve((size_t)cap * sizeof(uvip_keypoint)))) return rc;
    if ((rc = ex->out_desc.reserve((size_t)cap * 32))) return rc;
    if ((rc = ex->out_n.reserve(4))) return rc;
    cudaStream_t st = ex->stream;
    UVIP_CUDA(cudaMemcpyAsync(ex->in_frames.p, image, fbytes - (stride - w), cudaMemcpyHostToDevice, st));
    if (!full_detect) {
        if ((rc = ex->grid.reserve((size_t)grid_rows * grid_cols * 4))) return rc;
        UVIP_CUDA(cudaMemcpyAsync(ex->grid.p, grid, (size_t)grid_rows * grid_cols * 4, cudaMemcpyHostToDevice, st));
        if (n_in) {
            if ((rc = ex->incoming.reserve(align_up((size_t)n_in, 512) * sizeof(uvip_keypoint)))) return rc;   // coarse steps: the pointer is part of the graph key
            UVIP_CUDA(cudaMemcpyAsync(ex->incoming.p, kps, (size_t)n_in * sizeof(uvip_keypoint), cudaMemcpyHostToDevice, st));
        }
    }
    // the arguments that change from frame to frame on the live path reach k_select through device memory
    if ((rc = ex->dyn.reserve(16))) return rc;
    if (!ex->h_dyn) UVIP_CUDA(cudaHostAlloc((void**)&ex->h_dyn, 16, cudaHostAllocDefault));
    ex->h_dyn[0] = n_in; ex->h_dyn[1] = num_needed;
    UVIP_CUDA(cudaMemcpyAsync(ex->dyn.p, ex->h_dyn, 8, cudaMemcpyHostToDevice, st));
    if (ex->prof) {                                            // per-stage event timing wants plain launches
        rc = enqueue_group(ex, ex->in_frames.as<uint8_t>(), 1, stride, fbytes, ex->out_kps.as<uvip_keypoint>(), ex->out_n.as<int32_t>(),
                           cap, ex->out_desc.as<uint8_t>(), full_detect ? 1 : 0, n_in, grid_rows, grid_cols, min_px_dist, num_needed, st, true, ex->dyn.as<int>());
        if (rc) return rc;
    } else {
        uvip_extractor::GraphKey key;
        memset(&key, 0, sizeof(key));
        // num_needed and the NUMBER of incoming keypoints are not part of the key (they change almost every frame while the system is
        // WORKING, src/Tracking.cc:946: a key that contained them re-captured and re-instantiated the graph per call)
        key.w = w; key.h = h; key.stride = stride; key.cap = cap; key.full = full_detect ? 1 : 0; key.has_in = n_in > 0 ? 1 : 0;
        key.gr = full_detect ? 0 : grid_rows; key.gc = full_detect ? 0 : grid_cols; key.mpd = full_detect ? 0 : min_px_dist;
        key.in = ex->in_frames.p; key.kps = ex->out_kps.p; key.desc = ex->out_desc.p; key.n = ex->out_n.p; key.grid = ex->grid.p;
        key.incoming = ex->incoming.p; key.pyr = ex->pyr.p;
        if (!ex->ghave || memcmp(&key, &ex->gkey, sizeof(key)) != 0) {
            if (ex->gexec) { cudaGraphExecDestroy(ex->gexec); ex->gexec = nullptr; }
            ex->ghave = false;
            cudaGraph_t graph = nullptr;
            const long long l0 = ex->launches;
            UVIP_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            rc = enqueue_group(ex, ex->in_frames.as<uint8_t>(), 1, stride, fbytes, ex->out_kps.as<uvip_keypoint>(), ex->out_n.as<int32_t>(),
                               cap, ex->out_desc.as<uint8_t>(), full_detect ? 1 : 0, n_in, grid_rows, grid_cols, min_px_dist, num_needed, st, true, ex->dyn.as<int>());
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            ex->glaunches = ex->launches - l0; ex->launches = l0;
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            UVIP_CUDA(ce);
            const cudaError_t ci = cudaGraphInstantiate(&ex->gexec, graph, 0);
            cudaGraphDestroy(graph);
            UVIP_CUDA(ci);
            ex->gkey = key; ex->ghave = true; ex->gcaptures++;
        }
        UVIP_CUDA(cudaGraphLaunch(ex->gexec, st));
        ex->launches += ex->glaunches;
    }
    // one pinned staging block, one synchronisation: [n, status | cap keypoints | cap descriptors]
    const size_t o_kps = 64, o_desc = o_kps + align_up((size_t)cap * sizeof(uvip_keypoint), 64), need_bytes = o_desc + (size_t)cap * 32;
    if (ex->h_stage_bytes < need_bytes) {
        if (ex->h_stage) { cudaFreeHost(ex->h_stage); ex->h_stage = nullptr; ex->h_stage_bytes = 0; }
        UVIP_CUDA(cudaMallocHost((void**)&ex->h_stage, need_bytes));
        ex->h_stage_bytes = need_bytes;
    }
    UVIP_CUDA(cudaMemcpyAsync(ex->h_stage, ex->out_n.p, 4, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaMemcpyAsync(ex->h_stage + 4, ex->status.p, 4, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaMemcpyAsync(ex->h_stage + o_kps, ex->out_kps.p, (size_t)cap * sizeof(uvip_keypoint), cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaMemcpyAsync(ex->h_stage + o_desc, ex->out_desc.p, (size_t)cap * 32, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    int n = 0, flags = 0;
    memcpy(&n, ex->h_stage, 4); memcpy(&flags, ex->h_stage + 4, 4);
    if (flags) {
        set_last_error("device capacity overflow, flags 0x%x (1 FAST candidates, 2 quadtree nodes, 4 winners, 8 output rows)", flags);
        return UVIP_ERR_CAPACITY;
    }
    if (n > cap) { set_last_error("%d keypoints do not fit cap %d", n, cap); return UVIP_ERR_CAPACITY; }
    if (n) {
        memcpy(kps, ex->h_stage + o_kps, (size_t)n * sizeof(uvip_keypoint));
        memcpy(desc, ex->h_stage + o_desc, (size_t)n * 32);
    }
    if (!full_detect) {                                        // the caller's occupancy grid comes back only on success
        UVIP_CUDA(cudaMemcpyAsync(grid, ex->grid.p, (size_t)grid_rows * grid_cols * 4, cudaMemcpyDeviceToHost, st));
        UVIP_CUDA(cudaStreamSynchronize(st));
    }
    *n_inout = n;
    return UVIP_OK;
}

// ---- CLAHE pre-processing (next row N3) ---------------------------------------------------------------
static int enqueue_clahe(uvip_extractor* ex, const uint8_t* d_src, int nframes, int w, int h, int stride, size_t pitch, double clip, int tiles_x,
                         int tiles_y, uint8_t* d_dst, int dstride, size_t dpitch, cudaStream_t st)
{
    int ew = w, eh = h;
    if (w % tiles_x != 0 || h % tiles_y != 0) { ew = w + (tiles_x - (w % tiles_x)); eh = h + (tiles_y - (h % tiles_y)); }
    const int tw = ew / tiles_x, th = eh / tiles_y;
    UVIP_CHECK_ARG(tw >= 1 && th >= 1 && ew - w < w && eh - h < h);
    const int total = tw * th;
    const float lut_scale = (float)255 / total;
    int clip_limit = 0;
    if (clip > 0.0) { clip_limit = (int)(clip * total / 256); if (clip_limit < 1) clip_limit = 1; }
    int rc;
    if ((rc = ex->clahe_lut.reserve((size_t)nframes * tiles_x * tiles_y * 256))) return rc;
    k_clahe_lut<<<dim3(tiles_x * tiles_y, nframes), 256, 0, st>>>(d_src, w, h, stride, pitch, tiles_x, tw, th, clip_limit, lut_scale, ex->clahe_lut.as<uint8_t>());
    ex->launches++;
    // a CL_W x CL_H rectangle touches at most (CL_W-1)/tw + 3 tile columns; smaller tiles take the generic kernel
    if ((CL_W - 1) / tw + 3 > CL_MAXT || (CL_H - 1) / th + 3 > CL_MAXT)
        k_clahe_apply_generic<<<dim3(div_up(w, 64), div_up(h, 4), nframes), 256, 0, st>>>(d_src, w, h, stride, pitch, tiles_x, tiles_y, 1.0f / tw, 1.0f / th,
                                                                                          ex->clahe_lut.as<uint8_t>(), d_dst, dstride, dpitch);
    else
    k_clahe_apply<<<dim3(div_up(w, CL_W), div_up(h, CL_H), nframes), 256, 0, st>>>(d_src, w, h, stride, pitch, tiles_x, tiles_y, 1.0f / tw, 1.0f / th,
                                                                              ex->clahe_lut.as<uint8_t>(), d_dst, dstride, dpitch);
    ex->launches++;
    UVIP_CUDA(cudaGetLastError());
    return UVIP_OK;
}

int uvip_clahe_batch_device(uvip_extractor* ex, const uint8_t* d_src, int nframes, int w, int h, int stride, size_t frame_pitch,
                            double clip_limit, int tiles_x, int tiles_y, uint8_t* d_dst, int dst_stride, size_t dst_pitch, void* stream)
{
    UVIP_CHECK_ARG(ex && d_src && d_dst && nframes >= 1 && w > 0 && h > 0 && stride >= w && dst_stride >= w && tiles_x >= 1 && tiles_y >= 1);
    DeviceGuard g(ex->device);
    return enqueue_clahe(ex, d_src, nframes, w, h, stride, frame_pitch, clip_limit, tiles_x, tiles_y, d_dst, dst_stride, dst_pitch,
                         stream ? (cudaStream_t)stream : ex->stream);
}

int uvip_clahe(uvip_extractor* ex, const uint8_t* src, int w, int h, int stride, double clip_limit, int tiles_x, int tiles_y,
               uint8_t* dst, int dst_stride)
{
    UVIP_CHECK_ARG(ex && tiles_x >= 1 && tiles_y >= 1);
    if (!src || w <= 0 || h <= 0) return UVIP_OK;
    UVIP_CHECK_ARG(dst && stride >= w && dst_stride >= w);
    std::lock_guard<std::mutex> lk(ex->mu);
    DeviceGuard g(ex->device);
    int rc;
    const size_t bytes = (size_t)w * h;
    if ((rc = ex->clahe_io.reserve(2 * bytes))) return rc;
    uint8_t* d_in = ex->clahe_io.as<uint8_t>(); uint8_t* d_out = d_in + bytes;
    cudaStream_t st = ex->stream;
    UVIP_CUDA(cudaMemcpy2DAsync(d_in, w, src, stride, w, h, cudaMemcpyHostToDevice, st));
    if ((rc = enqueue_clahe(ex, d_in, 1, w, h, w, bytes, clip_limit, tiles_x, tiles_y, d_out, w, bytes, st))) return rc;
    UVIP_CUDA(cudaMemcpy2DAsync(dst, dst_stride, d_out, w, w, h, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    return UVIP_OK;
}

// ---- per-stage timing --------------------------------------------------------------------------------
int uvip_extractor_profile(uvip_extractor* ex, int enable)
{
    UVIP_CHECK_ARG(ex);
    DeviceGuard g(ex->device);
    UVIP_CUDA(cudaDeviceSynchronize());
    if (enable && ex->prof_ev.empty()) {
        ex->prof_ev.resize((size_t)PROF_RING * (UVIP_NUM_STAGES + 1));
        for (auto& e : ex->prof_ev) UVIP_CUDA(cudaEventCreate(&e));
    }
    ex->prof = enable != 0;
    ex->prof_groups = 0;
    return UVIP_OK;
}

int uvip_extractor_stage_ms(uvip_extractor* ex, float* ms, int* ngroups)
{
    UVIP_CHECK_ARG(ex && ms);
    DeviceGuard g(ex->device);
    UVIP_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < UVIP_NUM_STAGES; i++) ms[i] = 0.f;
    const long long n = ex->prof_groups < PROF_RING ? ex->prof_groups : PROF_RING;
    for (long long gI = 0; gI < n; gI++) {
        cudaEvent_t* pe = ex->prof_ev.data() + (size_t)gI * (UVIP_NUM_STAGES + 1);
        for (int i = 0; i < UVIP_NUM_STAGES; i++) { float t = 0; UVIP_CUDA(cudaEventElapsedTime(&t, pe[i], pe[i + 1])); ms[i] += t; }
    }
    if (ngroups) *ngroups = (int)n;
    return UVIP_OK;
}

// ---- debug taps ------------------------------------------------------------------------------------
int uvip_get_pyramid_level(uvip_extractor* ex, int frame, int level, int blurred, uint8_t* dst, int dstride, int* w, int* h)
{
    UVIP_CHECK_ARG(ex && ex->plan.W > 0 && frame >= 0 && frame < ex->last_frames && level >= 0 && level < ex->prm.nlevels);
    const LevelInfo& L = ex->plan.lv[level];
    if (w) *w = L.w; if (h) *h = L.h;
    if (!dst) return UVIP_OK;
    UVIP_CHECK_ARG(dstride >= L.w);
    DeviceGuard g(ex->device);
    UVIP_CUDA(cudaDeviceSynchronize());
    const uint8_t* base = (blurred ? ex->blur.as<uint8_t>() : ex->pyr.as<uint8_t>()) + (size_t)frame * ex->plan.frame_bytes + L.poff +
                          (size_t)EDGE * L.pstride + EDGE;
    UVIP_CUDA(cudaMemcpy2D(dst, dstride, base, L.pstride, L.w, L.h, cudaMemcpyDeviceToHost));
    return UVIP_OK;
}

static int fetch_list(uvip_extractor* ex, const DevBuf& buf, size_t elem_off, int n, std::vector<unsigned>* out)
{
    out->resize(n > 0 ? n : 0);
    if (n > 0) UVIP_CUDA(cudaMemcpy(out->data(), buf.as<unsigned>() + elem_off, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return UVIP_OK;
}

int uvip_get_raw_corners(uvip_extractor* ex, int frame, int level, int32_t* xs, int32_t* ys, int32_t* scores, int cap, int* n)
{
    UVIP_CHECK_ARG(ex && n && ex->plan.W > 0 && frame >= 0 && frame < ex->last_frames && level >= 0 && level < ex->prm.nlevels);
    DeviceGuard g(ex->device);
    UVIP_CUDA(cudaDeviceSynchronize());
    const Plan& P = ex->plan; const LevelInfo& L = P.lv[level];
    int cnt = 0;
    const size_t cstride = (size_t)ex->prm.max_batch * P.nlevels;
    UVIP_CUDA(cudaMemcpy(&cnt, ex->counters.as<int>() + (size_t)frame * P.nlevels + level, 4, cudaMemcpyDeviceToHost));
    (void)cstride;
    std::vector<unsigned> v;
    int rc = fetch_list(ex, ex->cand, (size_t)frame * P.raw_per_frame + L.raw_off, cnt, &v);
    if (rc) return rc;
    // reference order: cell row, cell column, then y, x inside the cell (src/ORBextractor.cc:772-812)
    auto okey = [&](unsigned c) {
        const int x = (int)(c & 0xFFF) - EDGE, y = (int)((c >> 12) & 0xFFF) - EDGE;
        const int cx = x / L.wcell, cy = y / L.hcell;
        return ((unsigned long long)(cy * L.ncols + cx) << 24) | ((unsigned long long)y << 12) | (unsigned long long)x;
    };
    std::vector<std::pair<unsigned long long, unsigned>> s; s.reserve(v.size());
    for (unsigned c : v) s.emplace_back(okey(c), c);
    std::sort(s.begin(), s.end());
    *n = cnt;
    for (int i = 0; i < cnt && i < cap; i++) {
        const unsigned c = s[i].second;
        if (xs) xs[i] = (int)(c & 0xFFF) - (EDGE - 3);          // window coords (relative to minBorder = 13)
        if (ys) ys[i] = (int)((c >> 12) & 0xFFF) - (EDGE - 3);
        if (scores) scores[i] = (int)(c >> 24);
    }
    return UVIP_OK;
}

int uvip_get_level_keypoints(uvip_extractor* ex, int frame, int level, int32_t* xs, int32_t* ys, int32_t* scores, int cap, int* n)
{
    UVIP_CHECK_ARG(ex && n && ex->plan.W > 0 && frame >= 0 && frame < ex->last_frames && level >= 0 && level < ex->prm.nlevels);
    DeviceGuard g(ex->device);
    UVIP_CUDA(cudaDeviceSynchronize());
    const Plan& P = ex->plan; const LevelInfo& L = P.lv[level];
    int cnt = 0;
    const size_t cstride = (size_t)ex->prm.max_batch * P.nlevels;
    UVIP_CUDA(cudaMemcpy(&cnt, ex->counters.as<int>() + cstride + (size_t)frame * P.nlevels + level, 4, cudaMemcpyDeviceToHost));
    std::vector<unsigned> v;
    int rc = fetch_list(ex, ex->winners, (size_t)frame * P.kp_per_frame + L.kp_off, cnt, &v);
    if (rc) return rc;
    *n = cnt;
    for (int i = 0; i < cnt && i < cap; i++) {
        if (xs) xs[i] = (int)(v[i] & 0xFFF);
        if (ys) ys[i] = (int)((v[i] >> 12) & 0xFFF);
        if (scores) scores[i] = (int)(v[i] >> 24);
    }
    return UVIP_OK;
}

// HarrisResponses (src/ORBextractor.cc:80-121) at n points of pyramid level `level` of frame `frame` of the last extract call
int uvip_harris_responses(uvip_extractor* ex, int frame, int level, const float* xs, const float* ys, int n, int block_size, float harris_k,
                          float* out)
{
    UVIP_CHECK_ARG(ex && ex->plan.W > 0 && frame >= 0 && frame < ex->last_frames && level >= 0 && level < ex->prm.nlevels && n >= 0);
    UVIP_CHECK_ARG(block_size >= 1 && block_size * block_size <= 2048);
    if (n == 0) return UVIP_OK;
    UVIP_CHECK_ARG(xs && ys && out);
    const LevelInfo& L = ex->plan.lv[level];
    const int r = block_size / 2, reach = block_size - r;        // the window plus its gradient neighbours must stay inside the materialised ring
    for (int i = 0; i < n; i++)
        UVIP_CHECK_ARG(xs[i] - r - 1 >= -BORDER_W && xs[i] + reach + 1 <= L.w + BORDER_W && ys[i] - r - 1 >= -BORDER_W && ys[i] + reach + 1 <= L.h + BORDER_W);
    std::lock_guard<std::mutex> lk(ex->mu);
    DeviceGuard g(ex->device);
    int rc;
    if ((rc = ex->clahe_io.reserve((size_t)n * 12))) return rc;
    float* d = ex->clahe_io.as<float>();
    cudaStream_t st = ex->stream;
    UVIP_CUDA(cudaMemcpyAsync(d, xs, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    UVIP_CUDA(cudaMemcpyAsync(d + n, ys, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    k_harris<<<div_up(n, 128), 128, 0, st>>>(ex->pyr.as<uint8_t>(), frame, level, d, d + n, n, block_size, harris_k, d + 2 * (size_t)n, ex->plan);
    ex->launches++;
    UVIP_CUDA(cudaGetLastError());
    UVIP_CUDA(cudaMemcpyAsync(out, d + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    return UVIP_OK;
}

// The reference's dead detector path ComputeKeyPoints (src/ORBextractor.cc:536-746) as an OPTIONAL mode, on the pyramid of a
// frame of the last extract call: quota cells, cv::FAST per cell (+ retry at 5 when <= 3 corners), HarrisResponses when the
// extractor was created with HARRIS_SCORE, quota redistribution, KeyPointsFilter::retainBest per cell and per level,
// orientation.  kps = [nlevels][cap_per_level], level coordinates like allKeypoints[level] of the reference.
// Detection, scoring and orientation run on the GPU; the quota bookkeeping and retainBest (std::nth_element + std::partition,
// exactly as OpenCV's KeyPointsFilter) on the host — which of several keypoints with EQUAL response survive a cut is decided
// by std::nth_element in the reference too, i.e. by the C++ library.
namespace {
struct QuotaKP { float x, y, response; };
void retain_best(std::vector<QuotaKP>& k, int n)            // cv::KeyPointsFilter::retainBest (features2d/keypoint.cpp)
{
    if (n >= 0 && k.size() > (size_t)n) {
        if (n == 0) { k.clear(); return; }
        std::nth_element(k.begin(), k.begin() + n - 1, k.end(), [](const QuotaKP& a, const QuotaKP& b) { return a.response > b.response; });
        const float amb = k[(size_t)n - 1].response;
        auto e = std::partition(k.begin() + n, k.end(), [amb](const QuotaKP& q) { return q.response >= amb; });
        k.resize((size_t)(e - k.begin()));
    }
}
}

int uvip_compute_keypoints_quota(uvip_extractor* ex, int frame, uvip_keypoint* kps, int32_t* n_per_level, int cap_per_level)
{
    UVIP_CHECK_ARG(ex && kps && n_per_level && cap_per_level > 0 && ex->plan.W > 0 && frame >= 0 && frame < ex->last_frames);
    std::lock_guard<std::mutex> lk(ex->mu);
    DeviceGuard g(ex->device);
    const Plan& P = ex->plan;
    const int nlevels = P.nlevels;
    const float imageRatio = (float)P.lv[0].w / P.lv[0].h;                             // :541
    struct LevelCells { int cols, rows, cellW, cellH, nfeaturesCell, first; std::vector<int> iniX, iniY; std::vector<char> skipped; };
    std::vector<LevelCells> LC((size_t)nlevels);
    std::vector<RoiCell> cells;
    size_t max_area = 1;
    for (int level = 0; level < nlevels; level++) {
        const LevelInfo& L = P.lv[level];
        LevelCells& C = LC[(size_t)level];
        const int nDesired = ex->quota[level];
        C.cols = (int)sqrtf((float)nDesired / (5 * imageRatio));                     // :547-548
        C.rows = (int)(imageRatio * C.cols);
        if (C.cols < 1 || C.rows < 1) { set_last_error("quota detector: level %d has no cell grid (the reference divides by zero here)", level); return UVIP_ERR_UNSUPPORTED; }
        const int minB = EDGE, maxBX = L.w - EDGE, maxBY = L.h - EDGE, W = maxBX - minB, H = maxBY - minB;
        C.cellW = (int)ceilf((float)W / C.cols); C.cellH = (int)ceilf((float)H / C.rows);
        C.nfeaturesCell = (int)ceilf((float)nDesired / (C.rows * C.cols));
        C.first = (int)cells.size();
        C.iniX.assign((size_t)C.cols, 0); C.iniY.assign((size_t)C.rows, 0); C.skipped.assign((size_t)C.rows * C.cols, 1);
        float hY = (float)(C.cellH + 6);
        for (int i = 0; i < C.rows; i++) {                                           // :574-611
            const float iniY = (float)(minB + i * C.cellH - 3);
            C.iniY[(size_t)i] = (int)iniY;
            if (i == C.rows - 1) { hY = maxBY + 3 - iniY; if (hY <= 0) { for (int j = 0; j < C.cols; j++) cells.push_back(RoiCell{level, 0, 0, 0, 0}); continue; } }
            float hX = (float)(C.cellW + 6);
            for (int j = 0; j < C.cols; j++) {
                float iniX;
                if (i == 0) { iniX = (float)(minB + j * C.cellW - 3); C.iniX[(size_t)j] = (int)iniX; } else iniX = (float)C.iniX[(size_t)j];
                if (j == C.cols - 1) { hX = maxBX + 3 - iniX; if (hX <= 0) { cells.push_back(RoiCell{level, 0, 0, 0, 0}); continue; } }
                RoiCell rc; rc.level = level; rc.x0 = (int)iniX; rc.y0 = (int)iniY; rc.w = (int)(iniX + hX) - (int)iniX; rc.h = (int)(iniY + hY) - (int)iniY;
                UVIP_CHECK_ARG(rc.w >= 0 && rc.h >= 0 && rc.h <= 512 && rc.w < 4096 && rc.x0 + rc.w <= L.w + 3 && rc.y0 + rc.h <= L.h + 3);
                C.skipped[(size_t)i * C.cols + j] = 0;
                cells.push_back(rc);
                if ((size_t)rc.w * rc.h > max_area) max_area = (size_t)rc.w * rc.h;
            }
        }
    }
    const int ncells = (int)cells.size(), cap_cell = 4096;
    const size_t smem = 3 * max_area + 16;
    if (smem > 200 * 1024) { set_last_error("quota detector: a %zu-pixel cell does not fit shared memory", max_area); return UVIP_ERR_UNSUPPORTED; }
    int rc;
    size_t off = 0;
    auto sect = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_cells = sect((size_t)ncells * sizeof(RoiCell)), o_cnt = sect((size_t)ncells * 4), o_list = sect((size_t)ncells * cap_cell * 4),
                 o_resp = sect((size_t)ncells * cap_cell * 4);
    if ((rc = ex->clahe_io.reserve(off))) return rc;
    uint8_t* base = ex->clahe_io.as<uint8_t>();
    cudaStream_t st = ex->stream;
    UVIP_CUDA(cudaStreamSynchronize(st));
    UVIP_CUDA(cudaMemcpyAsync(base + o_cells, cells.data(), (size_t)ncells * sizeof(RoiCell), cudaMemcpyHostToDevice, st));
    UVIP_CUDA(cudaMemsetAsync(base + o_cnt, 0, (size_t)ncells * 4, st));
    UVIP_CUDA(cudaMemsetAsync(ex->status.p, 0, sizeof(int), st));
    UVIP_CUDA(cudaFuncSetAttribute(k_fast_roi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fast_roi<<<ncells, 256, smem, st>>>(ex->pyr.as<uint8_t>(), frame, (const RoiCell*)(base + o_cells), ex->prm.fast_th, 5, cap_cell,
                                          (unsigned*)(base + o_list), (int*)(base + o_cnt), ex->status.as<int>(), P);
    const bool harris = ex->prm.score_type == 0;                                    // ORBextractor::HARRIS_SCORE (include/ORBextractor.h:49)
    if (harris)
        k_harris_cells<<<ncells, 128, 0, st>>>(ex->pyr.as<uint8_t>(), frame, (const RoiCell*)(base + o_cells), cap_cell, (const unsigned*)(base + o_list),
                                               (const int*)(base + o_cnt), 0.04f, (float*)(base + o_resp), P);
    ex->launches += harris ? 2 : 1;
    UVIP_CUDA(cudaGetLastError());
    std::vector<int> counts((size_t)ncells);
    UVIP_CUDA(cudaMemcpyAsync(counts.data(), base + o_cnt, (size_t)ncells * 4, cudaMemcpyDeviceToHost, st));
    if ((rc = read_status(ex, st))) return rc;
    std::vector<unsigned> lists((size_t)ncells * cap_cell); std::vector<float> resp;
    UVIP_CUDA(cudaMemcpyAsync(lists.data(), base + o_list, lists.size() * 4, cudaMemcpyDeviceToHost, st));
    if (harris) { resp.resize(lists.size()); UVIP_CUDA(cudaMemcpyAsync(resp.data(), base + o_resp, resp.size() * 4, cudaMemcpyDeviceToHost, st)); }
    UVIP_CUDA(cudaStreamSynchronize(st));
    // ---- host: quotas and retainBest (:653-745)
    std::vector<std::vector<QuotaKP> > all((size_t)nlevels);
    for (int level = 0; level < nlevels; level++) {
        const LevelCells& C = LC[(size_t)level];
        const int nDesired = ex->quota[level], nCells = C.rows * C.cols;
        std::vector<std::vector<QuotaKP> > cellK((size_t)nCells);
        std::vector<int> nToRetain((size_t)nCells, 0), nTotal((size_t)nCells, 0); std::vector<char> bNoMore((size_t)nCells, 0);
        int nNoMore = 0, nToDistribute = 0;
        for (int ci = 0; ci < nCells; ci++) {
            if (C.skipped[(size_t)ci]) continue;                                      // `continue` of :584 / :604: the cell keeps its zeros
            const int gi = C.first + ci, n = counts[(size_t)gi];
            std::vector<QuotaKP>& v = cellK[(size_t)ci];
            v.resize((size_t)n);
            for (int k = 0; k < n; k++) {
                const unsigned e = lists[(size_t)gi * cap_cell + k];
                v[(size_t)k].x = (float)(e & 0xFFF); v[(size_t)k].y = (float)((e >> 12) & 0xFFF);
                v[(size_t)k].response = harris ? resp[(size_t)gi * cap_cell + k] : (float)(e >> 24);
            }
            nTotal[(size_t)ci] = n;
            if (n > C.nfeaturesCell) { nToRetain[(size_t)ci] = C.nfeaturesCell; bNoMore[(size_t)ci] = 0; }
            else { nToRetain[(size_t)ci] = n; nToDistribute += C.nfeaturesCell - n; bNoMore[(size_t)ci] = 1; nNoMore++; }
        }
        while (nToDistribute > 0 && nNoMore < nCells) {                               // :685-710
            const int nNew = C.nfeaturesCell + (int)ceilf((float)nToDistribute / (nCells - nNoMore));
            nToDistribute = 0;
            for (int ci = 0; ci < nCells; ci++) {
                if (bNoMore[(size_t)ci]) continue;
                if (nTotal[(size_t)ci] > nNew) { nToRetain[(size_t)ci] = nNew; bNoMore[(size_t)ci] = 0; }
                else { nToRetain[(size_t)ci] = nTotal[(size_t)ci]; nToDistribute += nNew - nTotal[(size_t)ci]; bNoMore[(size_t)ci] = 1; nNoMore++; }
            }
        }
        std::vector<QuotaKP>& keypoints = all[(size_t)level];
        for (int i = 0; i < C.rows; i++)
            for (int j = 0; j < C.cols; j++) {                                        // :718-735
                std::vector<QuotaKP>& v = cellK[(size_t)i * C.cols + j];
                retain_best(v, nToRetain[(size_t)i * C.cols + j]);
                if ((int)v.size() > nToRetain[(size_t)i * C.cols + j]) v.resize((size_t)nToRetain[(size_t)i * C.cols + j]);
                for (size_t k = 0; k < v.size(); k++) { v[k].x += (float)C.iniX[(size_t)j]; v[k].y += (float)C.iniY[(size_t)i]; keypoints.push_back(v[k]); }
            }
        if ((int)keypoints.size() > nDesired) { retain_best(keypoints, nDesired); keypoints.resize((size_t)nDesired); }     // :737-741
        if ((int)keypoints.size() > cap_per_level) { set_last_error("level %d: %zu keypoints do not fit cap %d", level, keypoints.size(), cap_per_level); return UVIP_ERR_CAPACITY; }
    }
    // ---- orientation (:744-745)
    std::vector<int> lxy; std::vector<float> angle;
    for (int level = 0; level < nlevels; level++)
        for (const QuotaKP& k : all[(size_t)level]) { lxy.push_back(level); lxy.push_back((int)lrintf(k.x)); lxy.push_back((int)lrintf(k.y)); }
    const int ntot = (int)(lxy.size() / 3);
    angle.assign((size_t)ntot, 0.f);
    if (ntot) {
        if ((rc = ex->clahe_io.reserve((size_t)ntot * 16))) return rc;
        int* d_lxy = ex->clahe_io.as<int>(); float* d_ang = (float*)(d_lxy + 3 * (size_t)ntot);
        UVIP_CUDA(cudaMemcpyAsync(d_lxy, lxy.data(), (size_t)ntot * 12, cudaMemcpyHostToDevice, st));
        k_angle_list<<<div_up(ntot, 8), 256, 0, st>>>(ex->pyr.as<uint8_t>(), frame, d_lxy, ntot, d_ang, P);
        ex->launches++;
        UVIP_CUDA(cudaGetLastError());
        UVIP_CUDA(cudaMemcpyAsync(angle.data(), d_ang, (size_t)ntot * 4, cudaMemcpyDeviceToHost, st));
        UVIP_CUDA(cudaStreamSynchronize(st));
    }
    size_t a = 0;
    for (int level = 0; level < nlevels; level++) {
        const float size = (float)(int)(31 * ex->scale[level]);                      // scaledPatchSize (:713)
        n_per_level[level] = (int)all[(size_t)level].size();
        for (size_t k = 0; k < all[(size_t)level].size(); k++, a++) {
            uvip_keypoint& o = kps[(size_t)level * cap_per_level + k];
            o.x = all[(size_t)level][k].x; o.y = all[(size_t)level][k].y; o.size = size; o.angle = angle[a];
            o.response = all[(size_t)level][k].response; o.octave = level; o.class_id = -1;
        }
    }
    return UVIP_OK;
}

}  // extern "C"
