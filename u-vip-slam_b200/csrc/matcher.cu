// matcher.cu — descriptor path of USLAM::ORBmatcher on sm_100a: 256-bit Hamming distance, brute-force
// k=2 nearest neighbours (popc-pipe bound), shard merge, ratio test, rotation histogram, frame grid and
// the grid-windowed search with sequentially-consistent claims.  C-ABI: include/uvip_orb.h.
//
// Reference semantics restated here (never its code): src/ORBmatcher.cc:40-133,1748-1810,
// src/FrameKTL.cc:250-264,359-436, include/utils.h:81-111.
#include "common.cuh"
#include <cooperative_groups.h>
#include <stdarg.h>
#include <limits.h>

namespace cg = cooperative_groups;

namespace uvip {

static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// =====================================================================================================
// K9: tiled Hamming top-2.  One thread per query; train rows stream through shared memory in 8 KB stages filled by
// the TMA engine (cp.async.bulk + mbarrier); every lane reads the same train row (LDS.128 broadcast).
//
// Per descriptor pair the naive count is 8 XOR + 8 POPC.  POPC issues at 16 lanes/clk/SM, LOP3 on the ALU pipe at 64.
// A carry-save (Harley-Seal) reduction of the 8 XOR words x_i = a_i ^ b_i trades POPCs for LOP3s:
//      (s1, c1) = CSA(x0, x1, x2)   (s2, c2) = CSA(x3, x4, x5)   (s3, c3) = CSA(s1, s2, x6)   (ts, tc) = CSA(c1, c2, c3)
//      distance = popc(s3) + popc(x7) + 2 popc(ts) + 4 popc(tc)                                        -> 4 POPC
// The SUM outputs of a CSA are linear in the inputs: s1 = (a0^a1^a2) ^ (b0^b1^b2), and the CARRY of a CSA is a function of
// two inputs and the sum (the third input is their XOR with the sum).  So with the query-side XORs kept in registers and
// the train-side XORs B012 = b0^b1^b2, B345, B06 = b0^..^b6 stored WITH THE ROW, the words x2, x5, x6 are never formed:
//      x0, x1, s1 = A012^B012, c1 = f(x0, x1, s1),  x3, x4, s2, c2,  s3 = A06^B06, c3 = f(s1, s2, s3),  x7,  ts, tc
// = 13 LOP3 + 4 POPC per pair (was 8 + 14 LOP3), which makes the kernel POPC-bound at 4 POPC per pair instead of ALU-bound.
// Each stage is rewritten in place after it lands: row = {b0, b1, b3, b4 | b7, B012, B345, B06} (one thread per row,
// b2, b5, b6 are not needed any more), so the inner loop still reads two LDS.128 per pair.
// Running top-2 is kept as packed 32-bit keys (dist << 16 | row-in-supertile): unique keys give the
// (distance, index) order, i.e. the reference's strict '<' scan where the first candidate wins ties.
// =====================================================================================================
constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 256;                      // train rows per stage
constexpr int KNN_STAGE_BYTES = KNN_TILE * 32;
constexpr int KNN_SUPER = 256;                     // tiles per 16-bit index window (65536 rows)

__device__ __forceinline__ void top2_push(int d, int gi, int& d1, int& i1, int& d2, int& i2) {
    if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = gi; }
    else if (d < d2) { d2 = d; i2 = gi; }
}

// carry of CSA(x, y, z) given x, y and the sum s = x^y^z: one LOP3
__device__ __forceinline__ unsigned csa_carry(unsigned x, unsigned y, unsigned s) { const unsigned xy = x ^ y; return (x & y) | (xy & (xy ^ s)); }

struct KnnQuery { unsigned a0, a1, a3, a4, a7, A012, A345, A06; };

// u = {b0, b1, b3, b4}, v = {b7, B012, B345, B06} (the in-place rewritten train row)
__device__ __forceinline__ unsigned hamming8(const KnnQuery& q, const uint4 u, const uint4 v)
{
    const unsigned x0 = q.a0 ^ u.x, x1 = q.a1 ^ u.y, s1 = q.A012 ^ v.y, c1 = csa_carry(x0, x1, s1);
    const unsigned x3 = q.a3 ^ u.z, x4 = q.a4 ^ u.w, s2 = q.A345 ^ v.z, c2 = csa_carry(x3, x4, s2);
    const unsigned s3 = q.A06 ^ v.w, c3 = csa_carry(s1, s2, s3);
    const unsigned x7 = q.a7 ^ v.x;
    const unsigned ts = c1 ^ c2 ^ c3, tc = (c1 & c2) | ((c1 ^ c2) & c3);
    return (unsigned)__popc(s3) + (unsigned)__popc(x7) + 2u * (unsigned)__popc(ts) + 4u * (unsigned)__popc(tc);
}

// rewrite the rows [first, first + count) of a landed stage in place (see above); one thread per row
__device__ __forceinline__ void knn_prepare_rows(uint4* T, int rows, int tid)
{
    for (int r = tid; r < rows; r += KNN_THREADS) {
        const uint4 u = T[2 * r], v = T[2 * r + 1];
        const unsigned B012 = u.x ^ u.y ^ u.z, B345 = u.w ^ v.x ^ v.y;
        T[2 * r] = make_uint4(u.x, u.y, u.w, v.x);
        T[2 * r + 1] = make_uint4(v.w, B012, B345, B012 ^ B345 ^ v.z);
    }
}

__global__ void __launch_bounds__(KNN_THREADS)
k_knn2(const uint8_t* __restrict__ q_base, const int32_t* __restrict__ d_nq, size_t q_pitch,
       const uint8_t* __restrict__ t_base, const int32_t* __restrict__ d_nt, size_t t_pitch,
       int nq_fixed, int nt_fixed, int idx_base,
       int32_t* __restrict__ idx2, int32_t* __restrict__ dist2, size_t res_pitch, int split_rows, size_t split_stride)
{
    __shared__ __align__(128) uint8_t s_tile[2][KNN_STAGE_BYTES];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int pair = blockIdx.y;
    const int nq = d_nq ? d_nq[pair] : nq_fixed;
    int nt = d_nt ? d_nt[pair] : nt_fixed;
    if ((int)(blockIdx.x * blockDim.x) >= nq) return;           // uniform per block
    const uint8_t* q = q_base + (size_t)pair * q_pitch;
    const uint8_t* t = t_base + (size_t)pair * t_pitch;
    if (split_rows > 0) {                                       // blockIdx.z owns train rows [z*split_rows, (z+1)*split_rows)
        const int z0 = blockIdx.z * split_rows;
        t += (size_t)z0 * 32; idx_base += z0; nt = max(0, min(split_rows, nt - z0));
        idx2 += (size_t)blockIdx.z * split_stride; dist2 += (size_t)blockIdx.z * split_stride;
    }
    const int tid = threadIdx.x;
    const int qi = blockIdx.x * blockDim.x + tid;
    const bool live = qi < nq;

    KnnQuery Q = {0, 0, 0, 0, 0, 0, 0, 0};
    if (live) {
        const uint4* qp = reinterpret_cast<const uint4*>(q + (size_t)qi * 32);
        const uint4 u = __ldg(qp), v = __ldg(qp + 1);
        Q.a0 = u.x; Q.a1 = u.y; Q.a3 = u.w; Q.a4 = v.x; Q.a7 = v.w;
        Q.A012 = u.x ^ u.y ^ u.z; Q.A345 = u.w ^ v.x ^ v.y; Q.A06 = Q.A012 ^ Q.A345 ^ v.z;
    }

    const int ntiles = (nt + KNN_TILE - 1) / KNN_TILE;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < 2 && s < ntiles; s++) {
            const int rows = min(KNN_TILE, nt - s * KNN_TILE);
            mbar_arrive_expect_tx(&s_bar[s], rows * 32);
            bulk_g2s(s_tile[s], t + (size_t)s * KNN_STAGE_BYTES, rows * 32, &s_bar[s]);
        }
    }

    int g1d = 257, g2d = 257, g1i = -1, g2i = -1;
    uint32_t b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;

    if (ntiles > 0) {                                              // stage 0: wait, rewrite in place, publish
        mbar_wait(&s_bar[0], 0);
        knn_prepare_rows(reinterpret_cast<uint4*>(s_tile[0]), min(KNN_TILE, nt), tid);
        __syncthreads();
    }
    for (int i = 0; i < ntiles; i++) {
        const int s = i & 1;
        const int rows = min(KNN_TILE, nt - i * KNN_TILE);
        const uint4* T = reinterpret_cast<const uint4*>(s_tile[s]);
        uint32_t key = (uint32_t)((i & (KNN_SUPER - 1)) * KNN_TILE);
        // keys of 4 train rows at a time; the running top-2 is touched only when one of them beats the current second best
        // (rare: a query's top-2 changes O(log n) times over a scan), which takes the 3 VIMNMX per pair off the ALU pipe
        int j = 0;
        for (; j + 4 <= rows; j += 4) {
            uint32_t k4[4];
#pragma unroll
            for (int r = 0; r < 4; r++) k4[r] = (hamming8(Q, T[2 * (j + r)], T[2 * (j + r) + 1]) << 16) + key + (uint32_t)(j + r);
            const uint32_t mn = min(__vimin3_u32(k4[0], k4[1], k4[2]), k4[3]);
            if (mn < b2) {
#pragma unroll
                for (int r = 0; r < 4; r++) { const uint32_t lo = min(b1, k4[r]), hi = max(b1, k4[r]); b2 = min(b2, hi); b1 = lo; }
            }
        }
        for (; j < rows; j++) {
            const uint32_t k = (hamming8(Q, T[2 * j], T[2 * j + 1]) << 16) + key + (uint32_t)j;
            const uint32_t lo = min(b1, k), hi = max(b1, k);
            b2 = min(b2, hi);
            b1 = lo;
        }
        if (i + 1 < ntiles) {                                      // the next stage landed long ago: rewrite it while this one drains
            mbar_wait(&s_bar[s ^ 1], ((i + 1) >> 1) & 1);
            knn_prepare_rows(reinterpret_cast<uint4*>(s_tile[s ^ 1]), min(KNN_TILE, nt - (i + 1) * KNN_TILE), tid);
        }
        __syncthreads();                                           // stage s fully consumed, stage s^1 rewritten
        if (tid == 0 && i + 2 < ntiles) {
            const int r2 = min(KNN_TILE, nt - (i + 2) * KNN_TILE);
            fence_proxy_async_smem();                              // the stage was rewritten through the generic proxy
            mbar_arrive_expect_tx(&s_bar[s], r2 * 32);
            bulk_g2s(s_tile[s], t + (size_t)(i + 2) * KNN_STAGE_BYTES, r2 * 32, &s_bar[s]);
        }
        if ((i & (KNN_SUPER - 1)) == KNN_SUPER - 1 || i == ntiles - 1) {
            const int sbase = idx_base + (i & ~(KNN_SUPER - 1)) * KNN_TILE;
            if (b1 != 0xFFFFFFFFu) top2_push((int)(b1 >> 16), sbase + (int)(b1 & 0xFFFFu), g1d, g1i, g2d, g2i);
            if (b2 != 0xFFFFFFFFu) top2_push((int)(b2 >> 16), sbase + (int)(b2 & 0xFFFFu), g1d, g1i, g2d, g2i);
            b1 = b2 = 0xFFFFFFFFu;
        }
    }
    if (live) {
        int32_t* oi = idx2 + (size_t)pair * 2 * res_pitch + 2 * (size_t)qi;
        int32_t* od = dist2 + (size_t)pair * 2 * res_pitch + 2 * (size_t)qi;
        *reinterpret_cast<int2*>(oi) = make_int2(g1i, g2i);
        *reinterpret_cast<int2*>(od) = make_int2(g1d, g2d);
    }
}

// K12 merge: lexicographic min-2 over (dist, global index) across partial lists — shard-count invariant
__global__ void k_knn2_merge(const int32_t* __restrict__ idx_parts, const int32_t* __restrict__ dist_parts, int nparts,
                             size_t part_stride, int nq, int32_t* __restrict__ idx2, int32_t* __restrict__ dist2)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    int d1 = 257, d2 = 257, i1 = -1, i2 = -1;
    for (int p = 0; p < nparts; p++) {
        const int2 ii = *reinterpret_cast<const int2*>(idx_parts + p * part_stride + 2 * (size_t)qi);
        const int2 dd = *reinterpret_cast<const int2*>(dist_parts + p * part_stride + 2 * (size_t)qi);
        const int ci[2] = {ii.x, ii.y}, cd[2] = {dd.x, dd.y};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int d = cd[k], gi = ci[k];
            if (gi < 0) continue;
            if (d < d1 || (d == d1 && gi < i1)) { d2 = d1; i2 = i1; d1 = d; i1 = gi; }
            else if (d < d2 || (d == d2 && gi < i2)) { d2 = d; i2 = gi; }
        }
    }
    *reinterpret_cast<int2*>(idx2 + 2 * (size_t)qi) = make_int2(i1, i2);
    *reinterpret_cast<int2*>(dist2 + 2 * (size_t)qi) = make_int2(d1, d2);
}

// N4, second half: haloc::Hash::getHash (src/hash.cpp:57-85) for a batch of descriptor sets and haloc::Hash::match (:190-206)
// of one query hash against a table.  hash[i*32 + n] = (sum_m r_i[m] * (float)desc[m][n]) / rows with the products and
// the running sum in float, in row order (one thread per output element walks its column); match = sum |a - b| in order.
__global__ void __launch_bounds__(128)
k_haloc_hash(const uint8_t* __restrict__ desc, const int32_t* __restrict__ start, const float* __restrict__ proj, int num_proj, int proj_len,
             float* __restrict__ hash)
{
    const int s = blockIdx.x, b = start[s], rows = start[s + 1] - b;
    for (int o = threadIdx.x; o < num_proj * 32; o += blockDim.x) {
        const int i = o >> 5, n = o & 31;
        const float* r = proj + (size_t)i * proj_len;
        const uint8_t* d = desc + (size_t)b * 32 + n;
        float sum = 0.0f;
        for (int m = 0; m < rows; m++) sum = __fadd_rn(sum, __fmul_rn(__ldg(r + m), (float)d[(size_t)m * 32]));
        hash[(size_t)s * num_proj * 32 + o] = rows > 0 ? __fdiv_rn(sum, (float)rows) : 0.0f;
    }
}
__global__ void k_haloc_match(const float* __restrict__ query, const float* __restrict__ table, int n, int len, float* __restrict__ score)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float* h = table + (size_t)t * len;
    float sum = 0.0f;
    for (int i = 0; i < len; i++) sum = __fadd_rn(sum, fabsf(__fsub_rn(query[i], h[i])));
    score[t] = sum;
}

// M1: DescriptorDistance for n row pairs
__global__ void k_descriptor_distance(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n, int32_t* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* pa = reinterpret_cast<const uint4*>(a + (size_t)i * 32);
    const uint4* pb = reinterpret_cast<const uint4*>(b + (size_t)i * 32);
    const uint4 u = pa[0], v = pa[1], x = pb[0], y = pb[1];
    out[i] = __popc(u.x ^ x.x) + __popc(u.y ^ x.y) + __popc(u.z ^ x.z) + __popc(u.w ^ x.w) +
             __popc(v.x ^ y.x) + __popc(v.y ^ y.y) + __popc(v.z ^ y.z) + __popc(v.w ^ y.w);
}

// N4: MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:236-264) for a batch of map points, one warp per point.
// Lane i owns row i of the N x N distance matrix; its median (sorted row, element floor((N-1)/2), the diagonal 0
// included) is found by a 9-step bisection on the distance value, every step recounting the row (row j is a broadcast
// load for the whole warp), so no N-sized storage exists and N is unbounded.  Winner = least median, first on ties.
__global__ void k_distinctive(const uint8_t* __restrict__ desc, const int32_t* __restrict__ start, int npoints,
                              int32_t* __restrict__ best_idx, int32_t* __restrict__ best_median)
{
    const int p = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (p >= npoints) return;
    const int s0 = start[p], N = start[p + 1] - s0;
    if (N <= 0) { if (lane == 0) { best_idx[p] = -1; best_median[p] = -1; } return; }
    const int k1 = ((N - 1) >> 1) + 1;                     // the median has at least k1 row elements <= it
    const uint4* rows = reinterpret_cast<const uint4*>(desc + (size_t)s0 * 32);
    unsigned bestkey = 0xFFFFFFFFu;
    for (int i0 = 0; i0 < N; i0 += 32) {
        const int i = i0 + lane;
        const bool act = i < N;
        const uint4 a0 = __ldg(rows + 2 * (act ? i : 0)), a1 = __ldg(rows + 2 * (act ? i : 0) + 1);
        int lo = 0, hi = 256;
#pragma unroll 1
        for (int it = 0; it < 9; it++) {
            const int mid = (lo + hi) >> 1;
            int cnt = 0;
            for (int j = 0; j < N; j++) {
                const uint4 b0 = __ldg(rows + 2 * j), b1 = __ldg(rows + 2 * j + 1);
                const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                              __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                cnt += d <= mid;
            }
            if (cnt >= k1) hi = mid; else lo = mid + 1;
        }
        const unsigned key = act ? ((unsigned)lo << 16 | (unsigned)i) : 0xFFFFFFFFu;
        bestkey = min(bestkey, key);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bestkey = min(bestkey, __shfl_xor_sync(0xFFFFFFFFu, bestkey, o));
    if (lane == 0) { best_idx[p] = (int)(bestkey & 0xFFFFu); best_median[p] = (int)(bestkey >> 16); }
}

// ratio test, include/utils.h:104-108 (float distances, double ratio)
__global__ void k_ratio_filter(const int32_t* __restrict__ idx2, const int32_t* __restrict__ dist2, int nq, double ratio,
                               int32_t* __restrict__ match, int* __restrict__ nmatches)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int m = -1;
    if (idx2[2 * i] >= 0 && idx2[2 * i + 1] >= 0) {
        const float d0 = (float)dist2[2 * i], d1 = (float)dist2[2 * i + 1];
        if ((double)d0 <= __dmul_rn((double)d1, ratio)) m = idx2[2 * i];
    }
    match[i] = m;
    if (m >= 0) atomicAdd(nmatches, 1);
}

// K11 rotation histogram: single CTA
__global__ void k_rot_hist(int32_t* __restrict__ match, int n, const float* __restrict__ angle_a, const float* __restrict__ angle_b,
                           int* __restrict__ nkept)
{
    __shared__ int hist[UVIP_HISTO_LENGTH];
    __shared__ int keep[3];
    __shared__ int kept;
    const float factor = 1.0f / UVIP_HISTO_LENGTH;
    for (int i = threadIdx.x; i < UVIP_HISTO_LENGTH; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x == 0) kept = 0;
    __syncthreads();
    auto bin_of = [&](int i) -> int {
        float rot = __fsub_rn(angle_a[i], angle_b[match[i]]);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, factor));
        if (bin == UVIP_HISTO_LENGTH) bin = 0;
        return bin;
    };
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (match[i] < 0) continue;
        const int b = bin_of(i);
        if (b >= 0 && b < UVIP_HISTO_LENGTH) atomicAdd(&hist[b], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {     // ComputeThreeMaxima, src/ORBmatcher.cc:1748-1789
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < UVIP_HISTO_LENGTH; i++) {
            const int s = hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
        keep[0] = ind1; keep[1] = ind2; keep[2] = ind3;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (match[i] < 0) continue;
        const int b = bin_of(i);
        if (b >= 0 && b < UVIP_HISTO_LENGTH && b != keep[0] && b != keep[1] && b != keep[2]) match[i] = -1;
        else atomicAdd(&kept, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) *nkept = kept;
}

// frame grid: FrameKTL.cc:250-264 + PosInGrid :426-436.  Single CTA; counts/cursors in dynamic smem.
__global__ void k_grid_build(const float* __restrict__ kx, const float* __restrict__ ky, int n,
                             float min_x, float min_y, float inv_w, float inv_h, int cols, int rows,
                             int32_t* __restrict__ cell_start, int32_t* __restrict__ cell_items, int32_t* __restrict__ cell_of,
                             const int32_t* __restrict__ d_n, int k_stride)
{
    extern __shared__ int s_cnt[];      // ncell + 1 counts, then ncell cursors
    const int ncell = cols * rows;
    {   // batched form: blockIdx.x = frame, per-frame arrays k_stride apart, per-frame sizes in d_n
        const size_t f = blockIdx.x;
        kx += f * k_stride; ky += f * k_stride; cell_items += f * k_stride; cell_of += f * k_stride; cell_start += f * (ncell + 1);
        if (d_n) n = d_n[f];
    }
    int* s_cur = s_cnt + ncell + 1;
    for (int c = threadIdx.x; c <= ncell; c += blockDim.x) s_cnt[c] = 0;
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) s_cur[c] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kx[i], min_x), inv_w));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(ky[i], min_y), inv_h));
        int c = -1;
        if (px >= 0 && px < cols && py >= 0 && py < rows) { c = px * rows + py; atomicAdd(&s_cnt[c + 1], 1); }
        cell_of[i] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) for (int c = 0; c < ncell; c++) s_cnt[c + 1] += s_cnt[c];
    __syncthreads();
    for (int c = threadIdx.x; c <= ncell; c += blockDim.x) cell_start[c] = s_cnt[c];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = cell_of[i];
        if (c >= 0) cell_items[s_cnt[c] + atomicAdd(&s_cur[c], 1)] = i;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) {       // ascending keypoint index inside a cell
        const int b = s_cnt[c], e = s_cnt[c + 1];
        for (int i = b + 1; i < e; i++) {
            const int v = cell_items[i];
            int j = i - 1;
            while (j >= b && cell_items[j] > v) { cell_items[j + 1] = cell_items[j]; j--; }
            cell_items[j + 1] = v;
        }
    }
}

// K10: grid-windowed search (GetFeaturesInArea + best/second-best scan + accept rule) for one query under a
// given view of the claims: keypoint idx is unavailable iff owner[idx] < q.
struct SearchCtx {
    uvip_search_params sp;
    const float *qu, *qv, *qr; const int32_t *qminL, *qmaxL; const uint8_t* qdesc;
    const float *kx, *ky; const int32_t* octave; const uint8_t* kdesc;
    const int32_t *cell_start, *cell_items;
    const int32_t *cand_start, *cand_idx;          // explicit candidate lists (node-restricted searches); NULL = grid window
    const float* qline;                            // mode 5: epipolar line (a, b, c, a*a+b*b) per query
    const double* kthr;                            // mode 5: 3.84 * sigma2(octave) per keypoint
    // optional per-query candidate cache of the grid-window search (NULL = none): the keypoints that pass the level and radius tests of
    // a query do not depend on the claims, and neither do their distances, so the first round of the claim iteration records them
    // (in scan order: id | octave << 19 | distance << 23) and later rounds only replay the list under the new claims.
    // qcache_n[q] = entries, or -1 when the query has more than SW_CACHE_K (it is searched in full every round).
    unsigned* qcache; int* qcache_n;
};
constexpr int SW_CACHE_K = 6;
static bool env_set(const char* name) { const char* e = getenv(name); return e && *e && *e != '0'; }

__device__ __forceinline__ int hamming256(const uint4& u, const uint4& v, const uint8_t* __restrict__ row)
{
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(row)), y = __ldg(reinterpret_cast<const uint4*>(row) + 1);
    return __popc(u.x ^ x.x) + __popc(u.y ^ x.y) + __popc(u.z ^ x.z) + __popc(u.w ^ x.w) +
           __popc(v.x ^ y.x) + __popc(v.y ^ y.y) + __popc(v.z ^ y.z) + __popc(v.w ^ y.w);
}

// the accept rule at the end of a scan (src/ORBmatcher.cc:113-117, :470, :1703)
__device__ __forceinline__ int search_accept(const uvip_search_params& sp, int bestIdx, int bestDist, int bestDist2, int bestLevel, int bestLevel2)
{
    if (bestIdx < 0 || bestDist > sp.th_dist) return -1;
    if (sp.mode == 0 && bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(sp.ratio, (float)bestDist2)) return -1;   // :116-117
    if (sp.mode == 6 && !((float)bestDist <= __fmul_rn((float)bestDist2, sp.ratio))) return -1;                         // :470, :585
    return bestIdx;
}

// round >= 1 of the claim iteration: replay the query's recorded candidates under the current claims
__device__ __forceinline__ int search_replay(const SearchCtx& c, int q, const int* owner, int n)
{
    const uvip_search_params& sp = c.sp;
    int bestDist = sp.mode == 0 ? 256 : INT_MAX, bestDist2 = sp.mode == 6 ? INT_MAX : 256, bestLevel = -1, bestLevel2 = -1, bestIdx = -1;
    const unsigned* e = c.qcache + (size_t)q * SW_CACHE_K;
    for (int j = 0; j < n; j++) {
        const unsigned w = e[j];
        const int id = (int)(w & 0x7FFFFu), oct = (int)((w >> 19) & 15u), dist = (int)(w >> 23);
        if (owner[id] < q) continue;
        if (sp.mode == 0) {
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oct; bestIdx = id; }
            else if (dist < bestDist2) { bestLevel2 = oct; bestDist2 = dist; }
        } else if (sp.mode == 6) {
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx = id; }
            else if (dist < bestDist2) bestDist2 = dist;
        } else if (dist < bestDist) { bestDist = dist; bestIdx = id; }
    }
    return search_accept(sp, bestIdx, bestDist, bestDist2, bestLevel, bestLevel2);
}

__device__ int search_one(const SearchCtx& c, int q, const int* owner, bool record)
{
    const uvip_search_params& sp = c.sp;
    int nrec = 0;                                      // recorded candidates (record == true); > SW_CACHE_K = overflow
    const float x = c.qu[q], y = c.qv[q], r = c.qr[q];
    const int minL = c.qminL[q], maxL = c.qmaxL[q];
    // FrameKTL::GetFeaturesInArea, src/FrameKTL.cc:359-386 (float ops rounded one by one, no FMA)
    int cx0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, sp.min_x), r), sp.inv_w)); cx0 = max(0, cx0);
    if (cx0 >= sp.cols) { if (record) c.qcache_n[q] = 0; return -1; }
    int cx1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, sp.min_x), r), sp.inv_w)); cx1 = min(sp.cols - 1, cx1);
    if (cx1 < 0) { if (record) c.qcache_n[q] = 0; return -1; }
    int cy0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, sp.min_y), r), sp.inv_h)); cy0 = max(0, cy0);
    if (cy0 >= sp.rows) { if (record) c.qcache_n[q] = 0; return -1; }
    int cy1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, sp.min_y), r), sp.inv_h)); cy1 = min(sp.rows - 1, cy1);
    if (cy1 < 0) { if (record) c.qcache_n[q] = 0; return -1; }
    const bool check = !(minL == -1 && maxL == -1);
    const bool same = check && (minL == maxL);
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(c.qdesc + (size_t)q * 32));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(c.qdesc + (size_t)q * 32) + 1);
    int bestDist = sp.mode == 0 ? 256 : INT_MAX, bestDist2 = sp.mode == 6 ? INT_MAX : 256, bestLevel = -1, bestLevel2 = -1, bestIdx = -1;
    for (int ix = cx0; ix <= cx1; ix++)
        for (int iy = cy0; iy <= cy1; iy++) {
            const int cell = ix * sp.rows + iy;
            const int e = c.cell_start[cell + 1];
            for (int j = c.cell_start[cell]; j < e; j++) {
                const int id = c.cell_items[j];
                const int oct = c.octave[id];
                if (check) {
                    if (same) { if (oct != minL) continue; }
                    else if (oct < minL || oct > maxL) continue;
                }
                if (fabsf(__fsub_rn(c.kx[id], x)) > r || fabsf(__fsub_rn(c.ky[id], y)) > r) continue;
                const int own = owner[id];
                if (own == -2) continue;                           // taken before the call: never a candidate of any round
                if (!record && own < q) continue;                  // F.mvpMapPoints[idx] already set (:91 / :1689)
                const int dist = hamming256(u, v, c.kdesc + (size_t)id * 32);
                if (record) {                                      // (round 0: nothing is claimed yet, so every candidate is scored anyway)
                    if (nrec < SW_CACHE_K) c.qcache[(size_t)q * SW_CACHE_K + nrec] = (unsigned)id | ((unsigned)oct << 19) | ((unsigned)dist << 23);
                    nrec += ((unsigned)oct > 15u || id >= (1 << 19)) ? SW_CACHE_K + 1 : 1;       // what the entry cannot hold: search in full every round
                    if (own < q) continue;
                }
                if (sp.mode == 0) {                                // src/ORBmatcher.cc:98-111
                    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oct; bestIdx = id; }
                    else if (dist < bestDist2) { bestLevel2 = oct; bestDist2 = dist; }
                } else if (sp.mode == 6) {                         // WindowSearch / SearchByProjection(F1, F2, ...): top-2 without levels (:454-465)
                    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx = id; }
                    else if (dist < bestDist2) bestDist2 = dist;
                } else if (dist < bestDist) { bestDist = dist; bestIdx = id; }   // :1697-1701
            }
        }
    if (record) c.qcache_n[q] = nrec <= SW_CACHE_K ? nrec : -1;
    return search_accept(sp, bestIdx, bestDist, bestDist2, bestLevel, bestLevel2);
}

// node-restricted search (SearchByBoW inner loops, src/ORBmatcher.cc:186-245 and :751-811; mode 1 = best-only lists)
__device__ int search_one_list(const SearchCtx& c, int q, const int* owner)
{
    const uvip_search_params& sp = c.sp;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(c.qdesc + (size_t)q * 32));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(c.qdesc + (size_t)q * 32) + 1);
    int best1 = INT_MAX, best2 = INT_MAX, bestIdx = -1;
    const int e = c.cand_start[q + 1];
    for (int j = c.cand_start[q]; j < e; j++) {
        const int id = c.cand_idx[j];
        if (owner[id] < q) continue;                                   // vpMapPointMatches[realIdxF] / vbMatched2[idx2] already set
        const int dist = hamming256(u, v, c.kdesc + (size_t)id * 32);
        if (dist < best1) { best2 = best1; best1 = dist; bestIdx = id; }
        else if (dist < best2) best2 = dist;
    }
    if (sp.mode == 5) return -1;                                       // handled by search_one_epipolar
    if (bestIdx < 0) return -1;
    bool ok;
    if (sp.mode == 1) ok = best1 <= sp.th_dist;
    else if (sp.mode == 2) ok = best1 <= sp.th_dist && (float)best1 < __fmul_rn(sp.ratio, (float)best2);
    else ok = best1 < sp.th_dist && (float)best1 < __fmul_rn(sp.ratio, (float)best2);
    return ok ? bestIdx : -1;
}

// SearchForTriangulation's inner loops (src/ORBmatcher.cc:893-952 + CheckDistEpipolarLine :136-153): candidates of the same
// vocabulary node with dist <= TH_LOW, sorted by (dist, index); among those with dist <= round(2 * best dist) the first one whose
// distance to the epipolar line passes the chi-square test is taken and claimed.  One pass: the best (dist, index) key over all
// candidates gives the bound, the best key over the epipolar-consistent ones the match.
__device__ int search_one_epipolar(const SearchCtx& c, int q, const int* owner)
{
    const uvip_search_params& sp = c.sp;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(c.qdesc + (size_t)q * 32));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(c.qdesc + (size_t)q * 32) + 1);
    const float4 L = __ldg(reinterpret_cast<const float4*>(c.qline) + q);       // a, b, c, den
    unsigned bestAll = 0xFFFFFFFFu, bestEpi = 0xFFFFFFFFu;
    const int e = c.cand_start[q + 1];
    for (int j = c.cand_start[q]; j < e; j++) {
        const int id = c.cand_idx[j];
        if (owner[id] < q) continue;                                            // vbMatched2[idx2]
        const int dist = hamming256(u, v, c.kdesc + (size_t)id * 32);
        if (dist > sp.th_dist) continue;
        const unsigned key = ((unsigned)dist << 20) | (unsigned)id;
        bestAll = min(bestAll, key);
        if (L.w == 0.f) continue;                                               // den == 0 -> false
        const float num = __fadd_rn(__fadd_rn(__fmul_rn(L.x, c.kx[id]), __fmul_rn(L.y, c.ky[id])), L.z);
        const float dsqr = __fdiv_rn(__fmul_rn(num, num), L.w);
        if ((double)dsqr < c.kthr[id]) bestEpi = min(bestEpi, key);
    }
    if (bestEpi == 0xFFFFFFFFu) return -1;
    const int distTh = (int)round(2.0 * (double)(bestAll >> 20));               // int DistTh = round(2*BestDist)
    return (int)(bestEpi >> 20) <= distTh ? (int)(bestEpi & 0xFFFFFu) : -1;
}

// The reference loop is sequential: a query skips keypoints claimed by EARLIER queries.  Here every query
// runs in parallel against the claims of the previous round (owner[idx] = lowest query index that claimed idx)
// and rounds repeat until the claim table is a fixed point; by induction over the query index the fixed point
// is exactly the sequential result ("claims").
// One thread-block CLUSTER per frame (hardware cluster barrier between the phases of
// a round), so that a single frame's 10 000 queries spread over several SMs instead of one.  Claim tables live in global
// memory (L2), the "anything changed" flag of a round in a small per-frame rotating triple.  The cluster barrier has
// release / acquire semantics at cluster scope (the acquire side invalidates L1), so the tables are read with plain cached loads:
// within a round `prev` is read-only and `cur` only receives atomics.
// Launch shapes (measured on B200, config 3): a lone frame wants its queries on many SMs (8 CTAs x 512 threads: 0.26 ms per
// host call against 0.60 ms with one 1024-thread CTA); a large batch already fills the GPU with frames and prefers few fat
// CTAs (2 x 1024: 1.06 ms per 256 frames against 1.44 ms).  The kernel reads its shape at run time.
constexpr int SW_THREADS = 1024;
__global__ void __launch_bounds__(SW_THREADS)
k_search_window(SearchCtx c, int nq, int nk, int32_t* __restrict__ taken, int32_t* __restrict__ match,
                int* ownerA, int* ownerB, int* __restrict__ out_counts,
                const int32_t* __restrict__ d_nq, const int32_t* __restrict__ d_nk, int q_stride, int k_stride, int* flags)
{
    cg::cluster_group cl = cg::this_cluster();
    const int CL = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int T = CL * (int)blockDim.x, t0 = rank * (int)blockDim.x + (int)threadIdx.x;
    {   // batched form (frames shard trivially: claims never cross frames): one cluster per frame
        const size_t f = blockIdx.x / CL;
        const size_t qo = f * q_stride, ko = f * k_stride;
        c.qu += qo; c.qv += qo; c.qr += qo; c.qminL += qo; c.qmaxL += qo; c.qdesc += qo * 32;
        c.kx += ko; c.ky += ko; c.octave += ko; c.kdesc += ko * 32;
        c.cell_start += f * (size_t)(c.sp.cols * c.sp.rows + 1); c.cell_items += ko;
        if (c.qcache) { c.qcache += qo * SW_CACHE_K; c.qcache_n += qo; }
        taken += ko; match += qo; ownerA += ko; ownerB += ko; out_counts += 2 * f; flags += 4 * f;
        if (d_nq) nq = d_nq[f];
        if (d_nk) nk = d_nk[f];
    }
    int* prev = ownerA; int* cur = ownerB;
    const bool claims = c.sp.mode != 4;                    // mode 4 (Fuse): every query is independent, taken[] is not consulted
    for (int i = t0; i < nk; i += T) prev[i] = (claims && taken[i] != -1) ? -2 : INT_MAX;
    if (t0 == 0) { flags[0] = flags[1] = flags[2] = 0; out_counts[0] = 0; }
    cl.sync();
    int rounds = 0;
    for (;;) {
        int* flag = flags + rounds % 3;
        for (int i = t0; i < nk; i += T) cur[i] = (claims && taken[i] != -1) ? -2 : INT_MAX;
        if (t0 == 0) flags[(rounds + 1) % 3] = 0;          // next round's flag: nobody reads or writes it during this round
        cl.sync();
        for (int q = t0; q < nq; q += T) {
            int r;
            if (c.cand_start) r = c.sp.mode == 5 ? search_one_epipolar(c, q, prev) : search_one_list(c, q, prev);
            else if (c.qcache && rounds > 0 && c.qcache_n[q] >= 0) r = search_replay(c, q, prev, c.qcache_n[q]);
            else r = search_one(c, q, prev, c.qcache != nullptr && rounds == 0);
            match[q] = r;
            if (r >= 0 && claims) atomicMin(&cur[r], q);
        }
        cl.sync();
        int ch = 0;
        for (int i = t0; i < nk; i += T) ch |= (cur[i] != prev[i]);
        if (__any_sync(0xFFFFFFFFu, ch) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
        cl.sync();
        rounds++;
        const int changed = __ldcg(flag);
        int* tmp = prev; prev = cur; cur = tmp;
        if (!changed) break;                               // uniform over the cluster: every CTA read the same flag after the barrier
    }
    int n = 0;
    if (claims) {
        for (int i = t0; i < nk; i += T) {
            const int o = prev[i];
            if (o >= 0 && o != INT_MAX) { taken[i] = o; n++; }
        }
    } else {
        for (int q = t0; q < nq; q += T) n += match[q] >= 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&out_counts[0], n);
    if (t0 == 0) out_counts[1] = rounds;
}

// launches the search for nframes frames: one cluster per frame
static cudaError_t launch_search(cudaStream_t st, int nframes, const SearchCtx& c, int nq, int nk, int32_t* taken, int32_t* match, int* oa, int* ob,
                                 int* counts, const int32_t* d_nq, const int32_t* d_nk, int q_stride, int k_stride, int* flags)
{
    NvtxRange nv("uvip_search_window");
    const int cluster = nframes <= 32 ? 8 : 2, threads = nframes <= 32 ? 512 : 1024;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nframes * cluster)); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_search_window, c, nq, nk, taken, match, oa, ob, counts, d_nq, d_nk, q_stride, k_stride, flags);
}

// N2: DBoW2 vocabulary-tree descent, TemplatedVocabulary::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1259)
// with FORB::distance = 256-bit Hamming.  One warp per descriptor: at every level lane c scores child c (k <= 32 per pass),
// the warp takes the minimum of (distance << 8 | child position) — strict '<', first child wins ties — and descends.
__global__ void __launch_bounds__(256)
k_bow_transform(const int32_t* __restrict__ child_start, const int32_t* __restrict__ child_ids, const uint8_t* __restrict__ node_desc,
                const double* __restrict__ node_weight, const int32_t* __restrict__ node_word, int L,
                const uint8_t* __restrict__ desc, int n, int levelsup,
                int32_t* __restrict__ word_id, int32_t* __restrict__ node_id, double* __restrict__ weight)
{
    const int f = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= n) return;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)f * 32)), v = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)f * 32) + 1);
    const int nid_level = L - levelsup;
    int final_id = 0, level = 0, nid = 0;
    for (;;) {
        const int b = __ldg(child_start + final_id), e = __ldg(child_start + final_id + 1);
        if (e <= b && level > 0) break;                                   // leaf reached
        if (e <= b) break;                                                 // malformed: root without children
        ++level;
        unsigned best = 0xFFFFFFFFu;
        for (int c0 = b; c0 < e; c0 += 32) {
            const int c = c0 + lane;
            unsigned key = 0xFFFFFFFFu;
            if (c < e) {
                const int id = __ldg(child_ids + c);
                key = ((unsigned)hamming256(u, v, node_desc + (size_t)id * 32) << 20) | (unsigned)(c - b);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xFFFFFFFFu, key, o));
            best = min(best, key);
        }
        final_id = __ldg(child_ids + b + (int)(best & 0xFFFFFu));
        if (level == nid_level) nid = final_id;
    }
    if (lane == 0) { word_id[f] = __ldg(node_word + final_id); weight[f] = __ldg(node_weight + final_id); node_id[f] = nid; }
}

// popc-pipe microbenchmark (roofline denominator for K9): 8 independent popc chains per thread
__global__ void k_popc_peak(int iters, uint32_t seed, uint32_t* __restrict__ out)
{
    uint32_t v0 = seed ^ threadIdx.x, v1 = v0 * 3u, v2 = v0 * 5u, v3 = v0 * 7u, v4 = v0 * 11u, v5 = v0 * 13u, v6 = v0 * 17u, v7 = v0 * 19u;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        s0 += __popc(v0 ^ i); s1 += __popc(v1 ^ i); s2 += __popc(v2 ^ i); s3 += __popc(v3 ^ i);
        s4 += __popc(v4 ^ i); s5 += __popc(v5 ^ i); s6 += __popc(v6 ^ i); s7 += __popc(v7 ^ i);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
}

}  // namespace uvip

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace uvip;

struct uvip_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevBuf q, t, idx, dist, misc, misc2, misc3, misc4, qcache;
    long long launches = 0;
    uint8_t* h_stage = nullptr; size_t h_stage_bytes = 0;      // pinned mirror of the upload / download block of uvip_search_frame
    std::mutex mu;
};

extern "C" {

int uvip_abi_version(void) { return UVIP_ABI_VERSION; }
const char* uvip_last_error(void) { return g_err; }
int uvip_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
float uvip_radius_by_viewing_cos(float view_cos) { return ((double)view_cos > 0.998) ? 2.5f : 4.0f; }

int uvip_matcher_create(int device, uvip_matcher** out)
{
    UVIP_CHECK_ARG(out != nullptr);
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_last_error("no CUDA device available; libuvip_orb has no CPU fallback");
        return UVIP_ERR_NO_DEVICE;
    }
    UVIP_CHECK_ARG(device >= 0 && device < ndev);
    DeviceGuard g(device);
    uvip_matcher* m = new uvip_matcher();
    m->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_last_error("cudaStreamCreate -> %s", cudaGetErrorString(e)); delete m; return UVIP_ERR_CUDA; }
    *out = m;
    return UVIP_OK;
}

int uvip_matcher_destroy(uvip_matcher* m)
{
    if (!m) return UVIP_OK;
    DeviceGuard g(m->device);
    cudaStreamSynchronize(m->stream);
    m->q.release(); m->t.release(); m->idx.release(); m->dist.release();
    m->misc.release(); m->misc2.release(); m->misc3.release(); m->misc4.release(); m->qcache.release();
    if (m->h_stage) cudaFreeHost(m->h_stage);
    cudaStreamDestroy(m->stream);
    delete m;
    return UVIP_OK;
}

long long uvip_matcher_launch_count(const uvip_matcher* m) { return m ? m->launches : 0; }
void* uvip_matcher_stream(uvip_matcher* m) { return m ? (void*)m->stream : nullptr; }
int uvip_matcher_sync(uvip_matcher* m)
{
    UVIP_CHECK_ARG(m);
    DeviceGuard g(m->device);
    UVIP_CUDA(cudaStreamSynchronize(m->stream));
    return UVIP_OK;
}

static int launch_knn2(uvip_matcher* m, const uint8_t* d_q, const int32_t* d_nq, size_t q_pitch,
                       const uint8_t* d_t, const int32_t* d_nt, size_t t_pitch, int npairs, int max_nq,
                       int nq_fixed, int nt_fixed, int idx_base, int32_t* d_idx2, int32_t* d_dist2, size_t res_pitch,
                       cudaStream_t st)
{
    NvtxRange nv("uvip_knn2");
    UVIP_CHECK_ARG(((uintptr_t)d_q & 15) == 0 && ((uintptr_t)d_t & 15) == 0 && (q_pitch & 15) == 0 && (t_pitch & 15) == 0);
    UVIP_CHECK_ARG(((uintptr_t)d_idx2 & 7) == 0 && ((uintptr_t)d_dist2 & 7) == 0);
    if (npairs <= 0 || max_nq <= 0) return UVIP_OK;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
    const long long ctas = (long long)div_up(max_nq, KNN_THREADS) * npairs;
    // mid-size single problems leave SMs short of warps: split the train rows over blockIdx.z and merge the partial top-2
    int nz = 1;
    if (npairs == 1 && !d_nt && ctas < 8LL * sms && nt_fixed >= 8192) {
        nz = (int)((8LL * sms + ctas - 1) / ctas); if (nz > 16) nz = 16;
        while (nz > 1 && nt_fixed / nz < 4096) nz--;
    }
    if (nz > 1) {
        const size_t part = (size_t)max_nq * 2;
        int rc;
        if ((rc = m->misc2.reserve(part * nz * 4))) return rc;
        if ((rc = m->misc3.reserve(part * nz * 4))) return rc;
        const int rows = (int)align_up((size_t)div_up(nt_fixed, nz), 256);
        dim3 grid(div_up(max_nq, KNN_THREADS), 1, div_up(nt_fixed, rows));
        k_knn2<<<grid, KNN_THREADS, 0, st>>>(d_q, d_nq, q_pitch, d_t, d_nt, t_pitch, nq_fixed, nt_fixed, idx_base,
                                              m->misc2.as<int32_t>(), m->misc3.as<int32_t>(), 0, rows, part);
        m->launches++;
        k_knn2_merge<<<div_up(max_nq, 256), 256, 0, st>>>(m->misc2.as<int32_t>(), m->misc3.as<int32_t>(), (int)grid.z, part, max_nq, d_idx2, d_dist2);
    } else {
        dim3 grid(div_up(max_nq, KNN_THREADS), npairs);
        k_knn2<<<grid, KNN_THREADS, 0, st>>>(d_q, d_nq, q_pitch, d_t, d_nt, t_pitch, nq_fixed, nt_fixed, idx_base,
                                              d_idx2, d_dist2, res_pitch, 0, 0);
    }
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    return UVIP_OK;
}

int uvip_knn2_device(uvip_matcher* m, const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int idx_base,
                     int32_t* d_idx2, int32_t* d_dist2, void* stream)
{
    UVIP_CHECK_ARG(m && nq >= 0 && nt >= 0);
    UVIP_CHECK_ARG(nq == 0 || (d_q && d_idx2 && d_dist2));
    UVIP_CHECK_ARG(nt == 0 || d_t);
    DeviceGuard g(m->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    return launch_knn2(m, d_q, nullptr, 0, d_t ? d_t : d_q, nullptr, 0, 1, nq, nq, nt, idx_base, d_idx2, d_dist2, 0, st);
}

int uvip_knn2_batch_device(uvip_matcher* m, const uint8_t* d_q, const int32_t* d_nq, size_t q_pitch,
                           const uint8_t* d_t, const int32_t* d_nt, size_t t_pitch, int npairs, int max_nq,
                           int32_t* d_idx2, int32_t* d_dist2, size_t res_pitch, void* stream)
{
    UVIP_CHECK_ARG(m && d_q && d_t && d_nq && d_nt && d_idx2 && d_dist2 && npairs >= 0 && max_nq >= 0);
    UVIP_CHECK_ARG(res_pitch >= (size_t)max_nq);
    DeviceGuard g(m->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    return launch_knn2(m, d_q, d_nq, q_pitch, d_t, d_nt, t_pitch, npairs, max_nq, 0, 0, 0, d_idx2, d_dist2, res_pitch, st);
}

int uvip_knn2(uvip_matcher* m, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx2, int32_t* dist2)
{
    UVIP_CHECK_ARG(m && nq >= 0 && nt >= 0);
    if (nq == 0) return UVIP_OK;
    UVIP_CHECK_ARG(q && idx2 && dist2 && (nt == 0 || t));
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    int rc;
    if ((rc = m->q.reserve((size_t)nq * 32))) return rc;
    if ((rc = m->t.reserve((size_t)(nt > 0 ? nt : 1) * 32))) return rc;
    if ((rc = m->idx.reserve((size_t)nq * 8))) return rc;
    if ((rc = m->dist.reserve((size_t)nq * 8))) return rc;
    UVIP_CUDA(cudaMemcpyAsync(m->q.p, q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
    if (nt) UVIP_CUDA(cudaMemcpyAsync(m->t.p, t, (size_t)nt * 32, cudaMemcpyHostToDevice, m->stream));
    rc = launch_knn2(m, m->q.as<uint8_t>(), nullptr, 0, m->t.as<uint8_t>(), nullptr, 0, 1, nq, nq, nt, 0,
                     m->idx.as<int32_t>(), m->dist.as<int32_t>(), 0, m->stream);
    if (rc) return rc;
    UVIP_CUDA(cudaMemcpyAsync(idx2, m->idx.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, m->stream));
    UVIP_CUDA(cudaMemcpyAsync(dist2, m->dist.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, m->stream));
    UVIP_CUDA(cudaStreamSynchronize(m->stream));
    return UVIP_OK;
}

int uvip_knn2_batch(uvip_matcher* m, const uint8_t* q, const int32_t* nq, size_t q_pitch,
                    const uint8_t* t, const int32_t* nt, size_t t_pitch, int npairs, int max_nq,
                    int32_t* idx2, int32_t* dist2, size_t res_pitch)
{
    UVIP_CHECK_ARG(m && q && t && nq && nt && idx2 && dist2 && npairs >= 0 && max_nq >= 0 && res_pitch >= (size_t)max_nq);
    UVIP_CHECK_ARG((q_pitch & 31) == 0 && (t_pitch & 31) == 0);
    if (npairs == 0 || max_nq == 0) return UVIP_OK;
    int max_nt = 0;
    for (int p = 0; p < npairs; p++) {
        UVIP_CHECK_ARG(nq[p] >= 0 && nq[p] <= max_nq && (size_t)nq[p] * 32 <= q_pitch && nt[p] >= 0 && (size_t)nt[p] * 32 <= t_pitch);
        if (nt[p] > max_nt) max_nt = nt[p];
    }
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    int rc;
    // frame-to-frame matching passes t = q + k*pitch (descriptors of the following frames): upload the union once
    const bool alias = t >= q && t_pitch == q_pitch && (size_t)(t - q) % q_pitch == 0 && (size_t)(t - q) <= (size_t)npairs * q_pitch;
    const size_t q_bytes = alias ? (size_t)(t - q) + (size_t)npairs * t_pitch : (size_t)npairs * q_pitch;
    if ((rc = m->q.reserve(q_bytes))) return rc;
    if (!alias && (rc = m->t.reserve((size_t)npairs * t_pitch))) return rc;
    if ((rc = m->idx.reserve((size_t)npairs * res_pitch * 8))) return rc;
    if ((rc = m->dist.reserve((size_t)npairs * res_pitch * 8))) return rc;
    if ((rc = m->misc.reserve((size_t)npairs * 8))) return rc;
    cudaStream_t st = m->stream;
    int32_t* d_nq = m->misc.as<int32_t>(); int32_t* d_nt = d_nq + npairs;
    UVIP_CUDA(cudaMemcpyAsync(m->q.p, q, q_bytes, cudaMemcpyHostToDevice, st));
    if (!alias) UVIP_CUDA(cudaMemcpyAsync(m->t.p, t, (size_t)npairs * t_pitch, cudaMemcpyHostToDevice, st));
    UVIP_CUDA(cudaMemcpyAsync(d_nq, nq, (size_t)npairs * 4, cudaMemcpyHostToDevice, st));
    UVIP_CUDA(cudaMemcpyAsync(d_nt, nt, (size_t)npairs * 4, cudaMemcpyHostToDevice, st));
    const uint8_t* d_t = alias ? m->q.as<uint8_t>() + (size_t)(t - q) : m->t.as<uint8_t>();
    rc = launch_knn2(m, m->q.as<uint8_t>(), d_nq, q_pitch, d_t, d_nt, t_pitch, npairs, max_nq, 0, 0, 0,
                     m->idx.as<int32_t>(), m->dist.as<int32_t>(), res_pitch, st);
    if (rc) return rc;
    UVIP_CUDA(cudaMemcpyAsync(idx2, m->idx.p, (size_t)npairs * res_pitch * 8, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaMemcpyAsync(dist2, m->dist.p, (size_t)npairs * res_pitch * 8, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    return UVIP_OK;
}

int uvip_knn2_merge_device(uvip_matcher* m, const int32_t* d_idx_parts, const int32_t* d_dist_parts, int nparts,
                           size_t part_stride, int nq, int32_t* d_idx2, int32_t* d_dist2, void* stream)
{
    UVIP_CHECK_ARG(m && d_idx_parts && d_dist_parts && d_idx2 && d_dist2 && nparts >= 1 && nq >= 0);
    UVIP_CHECK_ARG((part_stride & 1) == 0);
    if (nq == 0) return UVIP_OK;
    DeviceGuard g(m->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    k_knn2_merge<<<div_up(nq, 256), 256, 0, st>>>(d_idx_parts, d_dist_parts, nparts, part_stride, nq, d_idx2, d_dist2);
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    return UVIP_OK;
}

int uvip_descriptor_distance(uvip_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* out)
{
    UVIP_CHECK_ARG(m && n >= 0);
    if (n == 0) return UVIP_OK;
    UVIP_CHECK_ARG(a && b && out);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    int rc;
    if ((rc = m->q.reserve((size_t)n * 32))) return rc;
    if ((rc = m->t.reserve((size_t)n * 32))) return rc;
    if ((rc = m->idx.reserve((size_t)n * 4))) return rc;
    UVIP_CUDA(cudaMemcpyAsync(m->q.p, a, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream));
    UVIP_CUDA(cudaMemcpyAsync(m->t.p, b, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream));
    k_descriptor_distance<<<div_up(n, 256), 256, 0, m->stream>>>(m->q.as<uint8_t>(), m->t.as<uint8_t>(), n, m->idx.as<int32_t>());
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    UVIP_CUDA(cudaMemcpyAsync(out, m->idx.p, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    UVIP_CUDA(cudaStreamSynchronize(m->stream));
    return UVIP_OK;
}

int uvip_distinctive_descriptors(uvip_matcher* m, const uint8_t* desc, const int32_t* start, int npoints,
                                 int32_t* best_idx, int32_t* best_median)
{
    UVIP_CHECK_ARG(m && npoints >= 0);
    if (npoints == 0) return UVIP_OK;
    UVIP_CHECK_ARG(start && best_idx);
    const int total = start[npoints];
    UVIP_CHECK_ARG(start[0] == 0 && total >= 0 && (total == 0 || desc));
    for (int p = 0; p < npoints; p++) UVIP_CHECK_ARG(start[p + 1] >= start[p] && start[p + 1] - start[p] <= 65535);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    int rc;
    if ((rc = m->t.reserve((size_t)total * 32 + 32))) return rc;
    if ((rc = m->idx.reserve((size_t)(npoints + 1) * 4))) return rc;
    if ((rc = m->dist.reserve((size_t)npoints * 8))) return rc;
    if (total) UVIP_CUDA(cudaMemcpyAsync(m->t.p, desc, (size_t)total * 32, cudaMemcpyHostToDevice, m->stream));
    UVIP_CUDA(cudaMemcpyAsync(m->idx.p, start, (size_t)(npoints + 1) * 4, cudaMemcpyHostToDevice, m->stream));
    int32_t* d_best = m->dist.as<int32_t>();
    k_distinctive<<<div_up(npoints, 4), 128, 0, m->stream>>>(m->t.as<uint8_t>(), m->idx.as<int32_t>(), npoints, d_best, d_best + npoints);
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    UVIP_CUDA(cudaMemcpyAsync(best_idx, d_best, (size_t)npoints * 4, cudaMemcpyDeviceToHost, m->stream));
    if (best_median) UVIP_CUDA(cudaMemcpyAsync(best_median, d_best + npoints, (size_t)npoints * 4, cudaMemcpyDeviceToHost, m->stream));
    UVIP_CUDA(cudaStreamSynchronize(m->stream));
    return UVIP_OK;
}

int uvip_ratio_filter(uvip_matcher* m, const int32_t* idx2, const int32_t* dist2, int nq, double ratio,
                      int32_t* match, int* nmatches)
{
    UVIP_CHECK_ARG(m && nq >= 0);
    if (nmatches) *nmatches = 0;
    if (nq == 0) return UVIP_OK;
    UVIP_CHECK_ARG(idx2 && dist2 && match);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    int rc;
    if ((rc = m->idx.reserve((size_t)nq * 8))) return rc;
    if ((rc = m->dist.reserve((size_t)nq * 8))) return rc;
    if ((rc = m->misc.reserve((size_t)nq * 4 + 16))) return rc;
    int32_t* d_match = m->misc.as<int32_t>() + 4;
    int* d_n = m->misc.as<int>();
    UVIP_CUDA(cudaMemcpyAsync(m->idx.p, idx2, (size_t)nq * 8, cudaMemcpyHostToDevice, m->stream));
    UVIP_CUDA(cudaMemcpyAsync(m->dist.p, dist2, (size_t)nq * 8, cudaMemcpyHostToDevice, m->stream));
    UVIP_CUDA(cudaMemsetAsync(d_n, 0, 4, m->stream));
    k_ratio_filter<<<div_up(nq, 256), 256, 0, m->stream>>>(m->idx.as<int32_t>(), m->dist.as<int32_t>(), nq, ratio, d_match, d_n);
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    int n = 0;
    UVIP_CUDA(cudaMemcpyAsync(match, d_match, (size_t)nq * 4, cudaMemcpyDeviceToHost, m->stream));
    UVIP_CUDA(cudaMemcpyAsync(&n, d_n, 4, cudaMemcpyDeviceToHost, m->stream));
    UVIP_CUDA(cudaStreamSynchronize(m->stream));
    if (nmatches) *nmatches = n;
    return UVIP_OK;
}

int uvip_rot_hist_filter(uvip_matcher* m, int32_t* match, int n, const float* angle_a, const float* angle_b, int* nkept)
{
    UVIP_CHECK_ARG(m && n >= 0);
    if (nkept) *nkept = 0;
    if (n == 0) return UVIP_OK;
    UVIP_CHECK_ARG(match && angle_a && angle_b);
    int nb = 0;
    for (int i = 0; i < n; i++) if (match[i] >= nb) nb = match[i] + 1;
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    int rc;
    if ((rc = m->misc.reserve((size_t)n * 4 + 16))) return rc;
    if ((rc = m->misc2.reserve((size_t)n * 4))) return rc;
    if ((rc = m->misc3.reserve((size_t)(nb > 0 ? nb : 1) * 4))) return rc;
    int32_t* d_match = m->misc.as<int32_t>() + 4;
    int* d_n = m->misc.as<int>();
    UVIP_CUDA(cudaMemcpyAsync(d_match, match, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
    UVIP_CUDA(cudaMemcpyAsync(m->misc2.p, angle_a, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
    if (nb) UVIP_CUDA(cudaMemcpyAsync(m->misc3.p, angle_b, (size_t)nb * 4, cudaMemcpyHostToDevice, m->stream));
    k_rot_hist<<<1, 256, 0, m->stream>>>(d_match, n, m->misc2.as<float>(), m->misc3.as<float>(), d_n);
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    int kept = 0;
    UVIP_CUDA(cudaMemcpyAsync(match, d_match, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    UVIP_CUDA(cudaMemcpyAsync(&kept, d_n, 4, cudaMemcpyDeviceToHost, m->stream));
    UVIP_CUDA(cudaStreamSynchronize(m->stream));
    if (nkept) *nkept = kept;
    return UVIP_OK;
}

int uvip_grid_build(uvip_matcher* m, const float* kx, const float* ky, int n,
                    float min_x, float min_y, float inv_w, float inv_h, int cols, int rows,
                    int32_t* cell_start, int32_t* cell_items)
{
    UVIP_CHECK_ARG(m && n >= 0 && cols > 0 && rows > 0 && cell_start);
    const int ncell = cols * rows;
    UVIP_CHECK_ARG((size_t)(2 * ncell + 1) * 4 <= 200 * 1024);
    UVIP_CHECK_ARG(n == 0 || (kx && ky && cell_items));
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    int rc;
    const size_t nn = n > 0 ? n : 1;
    if ((rc = m->misc.reserve(nn * 4))) return rc;
    if ((rc = m->misc2.reserve(nn * 4))) return rc;
    if ((rc = m->misc3.reserve((size_t)(ncell + 1) * 4))) return rc;
    if ((rc = m->misc4.reserve(nn * 8))) return rc;
    if (n) {
        UVIP_CUDA(cudaMemcpyAsync(m->misc.p, kx, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
        UVIP_CUDA(cudaMemcpyAsync(m->misc2.p, ky, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
    }
    const size_t smem = (size_t)(2 * ncell + 1) * 4;
    UVIP_CUDA(cudaFuncSetAttribute(k_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int32_t* d_items = m->misc4.as<int32_t>();
    int32_t* d_cellof = d_items + nn;
    k_grid_build<<<1, 1024, smem, m->stream>>>(m->misc.as<float>(), m->misc2.as<float>(), n, min_x, min_y, inv_w, inv_h,
                                                cols, rows, m->misc3.as<int32_t>(), d_items, d_cellof, nullptr, 0);
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    UVIP_CUDA(cudaMemcpyAsync(cell_start, m->misc3.p, (size_t)(ncell + 1) * 4, cudaMemcpyDeviceToHost, m->stream));
    UVIP_CUDA(cudaStreamSynchronize(m->stream));
    const int nitems = cell_start[ncell];
    if (nitems) UVIP_CUDA(cudaMemcpy(cell_items, d_items, (size_t)nitems * 4, cudaMemcpyDeviceToHost));
    return UVIP_OK;
}

int uvip_search_window(uvip_matcher* m, const uvip_search_params* sp,
                       const float* qu, const float* qv, const float* qr, const int32_t* qmin_level,
                       const int32_t* qmax_level, const uint8_t* qdesc, int nq,
                       const float* kx, const float* ky, const int32_t* octave, const uint8_t* kdesc, int nk,
                       const int32_t* cell_start, const int32_t* cell_items,
                       int32_t* taken, int32_t* match, int* nmatches)
{
    UVIP_CHECK_ARG(m && sp && nq >= 0 && nk >= 0 && sp->cols > 0 && sp->rows > 0);
    if (nmatches) *nmatches = 0;
    if (nq == 0) return UVIP_OK;
    UVIP_CHECK_ARG(qu && qv && qr && qmin_level && qmax_level && qdesc && match && cell_start);
    UVIP_CHECK_ARG(nk == 0 || (kx && ky && octave && kdesc && taken && cell_items));
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    const int ncell = sp->cols * sp->rows;
    const int nitems = cell_start[ncell];
    // one staging buffer, 32-byte aligned sections
    size_t off = 0;
    auto sect = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 32); return o; };
    const size_t nkk = nk > 0 ? nk : 1;
    const size_t o_qu = sect((size_t)nq * 4), o_qv = sect((size_t)nq * 4), o_qr = sect((size_t)nq * 4);
    const size_t o_qmin = sect((size_t)nq * 4), o_qmax = sect((size_t)nq * 4), o_qd = sect((size_t)nq * 32);
    const size_t o_kx = sect(nkk * 4), o_ky = sect(nkk * 4), o_oct = sect(nkk * 4), o_kd = sect(nkk * 32);
    const size_t o_cs = sect((size_t)(ncell + 1) * 4), o_ci = sect((size_t)(nitems > 0 ? nitems : 1) * 4);
    const size_t o_taken = sect(nkk * 4), o_match = sect((size_t)nq * 4), o_oa = sect(nkk * 4), o_ob = sect(nkk * 4);
    const size_t o_cnt = sect(32);
    int rc;
    if ((rc = m->misc.reserve(off))) return rc;
    uint8_t* base = m->misc.as<uint8_t>();
    cudaStream_t st = m->stream;
#define UP(o, src, bytes) if ((bytes) > 0) UVIP_CUDA(cudaMemcpyAsync(base + (o), (src), (bytes), cudaMemcpyHostToDevice, st))
    UP(o_qu, qu, (size_t)nq * 4); UP(o_qv, qv, (size_t)nq * 4); UP(o_qr, qr, (size_t)nq * 4);
    UP(o_qmin, qmin_level, (size_t)nq * 4); UP(o_qmax, qmax_level, (size_t)nq * 4); UP(o_qd, qdesc, (size_t)nq * 32);
    UP(o_kx, kx, (size_t)nk * 4); UP(o_ky, ky, (size_t)nk * 4); UP(o_oct, octave, (size_t)nk * 4); UP(o_kd, kdesc, (size_t)nk * 32);
    UP(o_cs, cell_start, (size_t)(ncell + 1) * 4); UP(o_ci, cell_items, (size_t)nitems * 4);
    UP(o_taken, taken, (size_t)nk * 4);
#undef UP
    SearchCtx c; memset(&c, 0, sizeof(c));
    c.sp = *sp;
    c.qu = (const float*)(base + o_qu); c.qv = (const float*)(base + o_qv); c.qr = (const float*)(base + o_qr);
    c.qminL = (const int32_t*)(base + o_qmin); c.qmaxL = (const int32_t*)(base + o_qmax); c.qdesc = base + o_qd;
    c.kx = (const float*)(base + o_kx); c.ky = (const float*)(base + o_ky); c.octave = (const int32_t*)(base + o_oct);
    c.kdesc = base + o_kd; c.cell_start = (const int32_t*)(base + o_cs); c.cell_items = (const int32_t*)(base + o_ci);
    UVIP_CUDA(launch_search(st, 1, c, nq, nk, (int32_t*)(base + o_taken), (int32_t*)(base + o_match), (int*)(base + o_oa), (int*)(base + o_ob),
                            (int*)(base + o_cnt), nullptr, nullptr, 0, 0, (int*)(base + o_cnt) + 4));
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    int counts[2] = {0, 0};
    UVIP_CUDA(cudaMemcpyAsync(match, base + o_match, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (nk) UVIP_CUDA(cudaMemcpyAsync(taken, base + o_taken, (size_t)nk * 4, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaMemcpyAsync(counts, base + o_cnt, 8, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    if (nmatches) *nmatches = counts[0];
    return UVIP_OK;
}

int uvip_search_lists(uvip_matcher* m, int mode, int th_dist, float ratio, const uint8_t* qdesc, int nq,
                      const int32_t* cand_start, const int32_t* cand_idx, const uint8_t* kdesc, int nk,
                      int32_t* taken, int32_t* match, int* nmatches)
{
    UVIP_CHECK_ARG(m && nq >= 0 && nk >= 0 && mode >= 1 && mode <= 3);
    if (nmatches) *nmatches = 0;
    if (nq == 0) return UVIP_OK;
    UVIP_CHECK_ARG(qdesc && cand_start && match && (nk == 0 || (kdesc && taken)));
    const int ncand = cand_start[nq];
    UVIP_CHECK_ARG(ncand >= 0 && (ncand == 0 || cand_idx));
    for (int i = 0; i < ncand; i++) UVIP_CHECK_ARG(cand_idx[i] >= 0 && cand_idx[i] < nk);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    size_t off = 0;
    auto sect = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 32); return o; };
    const size_t nkk = nk > 0 ? nk : 1;
    const size_t o_qd = sect((size_t)nq * 32), o_kd = sect(nkk * 32), o_cs = sect((size_t)(nq + 1) * 4), o_ci = sect((size_t)(ncand > 0 ? ncand : 1) * 4);
    const size_t o_taken = sect(nkk * 4), o_match = sect((size_t)nq * 4), o_oa = sect(nkk * 4), o_ob = sect(nkk * 4), o_cnt = sect(32);
    int rc;
    if ((rc = m->misc.reserve(off))) return rc;
    uint8_t* base = m->misc.as<uint8_t>();
    cudaStream_t st = m->stream;
#define UP(o, src, bytes) if ((bytes) > 0) UVIP_CUDA(cudaMemcpyAsync(base + (o), (src), (bytes), cudaMemcpyHostToDevice, st))
    UP(o_qd, qdesc, (size_t)nq * 32); UP(o_kd, kdesc, (size_t)nk * 32); UP(o_cs, cand_start, (size_t)(nq + 1) * 4);
    UP(o_ci, cand_idx, (size_t)ncand * 4); UP(o_taken, taken, (size_t)nk * 4);
#undef UP
    SearchCtx c; memset(&c, 0, sizeof(c));
    c.sp.mode = mode; c.sp.th_dist = th_dist; c.sp.ratio = ratio;
    c.qdesc = base + o_qd; c.kdesc = base + o_kd;
    c.cand_start = (const int32_t*)(base + o_cs); c.cand_idx = (const int32_t*)(base + o_ci);
    UVIP_CUDA(launch_search(st, 1, c, nq, nk, (int32_t*)(base + o_taken), (int32_t*)(base + o_match), (int*)(base + o_oa), (int*)(base + o_ob),
                            (int*)(base + o_cnt), nullptr, nullptr, 0, 0, (int*)(base + o_cnt) + 4));
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    int counts[2] = {0, 0};
    UVIP_CUDA(cudaMemcpyAsync(match, base + o_match, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (nk) UVIP_CUDA(cudaMemcpyAsync(taken, base + o_taken, (size_t)nk * 4, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaMemcpyAsync(counts, base + o_cnt, 8, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    if (nmatches) *nmatches = counts[0];
    return UVIP_OK;
}

// uvip_search_window without the caller-side grid: one upload, grid build + search on the device, one download and a single
// synchronisation.  This is the per-frame call of the Tracking thread (SearchByProjection(F, local map points, th)).
int uvip_search_frame(uvip_matcher* m, const uvip_search_params* sp,
                      const float* qu, const float* qv, const float* qr, const int32_t* qmin_level,
                      const int32_t* qmax_level, const uint8_t* qdesc, int nq,
                      const float* kx, const float* ky, const int32_t* octave, const uint8_t* kdesc, int nk,
                      int32_t* taken, int32_t* match, int* nmatches)
{
    UVIP_CHECK_ARG(m && sp && nq >= 0 && nk >= 0 && sp->cols > 0 && sp->rows > 0);
    UVIP_CHECK_ARG(sp->mode == 0 || sp->mode == 1 || sp->mode == 4 || sp->mode == 6);
    if (nmatches) *nmatches = 0;
    if (nq == 0) return UVIP_OK;
    UVIP_CHECK_ARG(qu && qv && qr && qmin_level && qmax_level && qdesc && match);
    UVIP_CHECK_ARG(nk == 0 || (kx && ky && octave && kdesc && taken));
    const int ncell = sp->cols * sp->rows;
    UVIP_CHECK_ARG((size_t)(2 * ncell + 1) * 4 <= 200 * 1024);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    size_t off = 0;
    auto sect = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 32); return o; };
    const size_t nkk = nk > 0 ? nk : 1;
    // host-visible part first (one contiguous upload from a pinned mirror), device-only scratch after it
    const size_t o_qu = sect((size_t)nq * 4), o_qv = sect((size_t)nq * 4), o_qr = sect((size_t)nq * 4);
    const size_t o_qmin = sect((size_t)nq * 4), o_qmax = sect((size_t)nq * 4), o_qd = sect((size_t)nq * 32);
    const size_t o_kx = sect(nkk * 4), o_ky = sect(nkk * 4), o_oct = sect(nkk * 4), o_kd = sect(nkk * 32), o_taken = sect(nkk * 4);
    const size_t up_bytes = off;
    const size_t o_match = sect((size_t)nq * 4), o_cnt = sect(32);
    const size_t down_bytes = off - o_taken;
    const size_t o_cs = sect((size_t)(ncell + 1) * 4), o_ci = sect(nkk * 4), o_cof = sect(nkk * 4), o_oa = sect(nkk * 4), o_ob = sect(nkk * 4);
    const size_t o_qc = sect((size_t)nq * SW_CACHE_K * 4), o_qcn = sect((size_t)nq * 4);
    int rc;
    if ((rc = m->misc.reserve(off))) return rc;
    if (m->h_stage_bytes < o_cs) {
        if (m->h_stage) { cudaFreeHost(m->h_stage); m->h_stage = nullptr; m->h_stage_bytes = 0; }
        UVIP_CUDA(cudaMallocHost((void**)&m->h_stage, o_cs));
        m->h_stage_bytes = o_cs;
    }
    uint8_t* hs = m->h_stage; uint8_t* base = m->misc.as<uint8_t>();
    memcpy(hs + o_qu, qu, (size_t)nq * 4); memcpy(hs + o_qv, qv, (size_t)nq * 4); memcpy(hs + o_qr, qr, (size_t)nq * 4);
    memcpy(hs + o_qmin, qmin_level, (size_t)nq * 4); memcpy(hs + o_qmax, qmax_level, (size_t)nq * 4); memcpy(hs + o_qd, qdesc, (size_t)nq * 32);
    if (nk) {
        memcpy(hs + o_kx, kx, (size_t)nk * 4); memcpy(hs + o_ky, ky, (size_t)nk * 4); memcpy(hs + o_oct, octave, (size_t)nk * 4);
        memcpy(hs + o_kd, kdesc, (size_t)nk * 32); memcpy(hs + o_taken, taken, (size_t)nk * 4);
    }
    cudaStream_t st = m->stream;
    UVIP_CUDA(cudaMemcpyAsync(base, hs, up_bytes, cudaMemcpyHostToDevice, st));
    const size_t smem = (size_t)(2 * ncell + 1) * 4;
    UVIP_CUDA(cudaFuncSetAttribute(k_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_grid_build<<<1, 1024, smem, st>>>((const float*)(base + o_kx), (const float*)(base + o_ky), nk, sp->min_x, sp->min_y, sp->inv_w, sp->inv_h,
                                         sp->cols, sp->rows, (int32_t*)(base + o_cs), (int32_t*)(base + o_ci), (int32_t*)(base + o_cof), nullptr, 0);
    SearchCtx c; memset(&c, 0, sizeof(c));
    c.sp = *sp;
    c.qu = (const float*)(base + o_qu); c.qv = (const float*)(base + o_qv); c.qr = (const float*)(base + o_qr);
    c.qminL = (const int32_t*)(base + o_qmin); c.qmaxL = (const int32_t*)(base + o_qmax); c.qdesc = base + o_qd;
    c.kx = (const float*)(base + o_kx); c.ky = (const float*)(base + o_ky); c.octave = (const int32_t*)(base + o_oct);
    c.kdesc = base + o_kd; c.cell_start = (const int32_t*)(base + o_cs); c.cell_items = (const int32_t*)(base + o_ci);
    if (!env_set("UVIP_SEARCH_NOCACHE")) { c.qcache = (unsigned*)(base + o_qc); c.qcache_n = (int*)(base + o_qcn); }
    UVIP_CUDA(launch_search(st, 1, c, nq, nk, (int32_t*)(base + o_taken), (int32_t*)(base + o_match), (int*)(base + o_oa), (int*)(base + o_ob),
                            (int*)(base + o_cnt), nullptr, nullptr, 0, 0, (int*)(base + o_cnt) + 4));
    m->launches += 2;
    UVIP_CUDA(cudaGetLastError());
    UVIP_CUDA(cudaMemcpyAsync(hs + o_taken, base + o_taken, down_bytes, cudaMemcpyDeviceToHost, st));      // taken | match | counts
    UVIP_CUDA(cudaStreamSynchronize(st));
    memcpy(match, hs + o_match, (size_t)nq * 4);
    if (nk) memcpy(taken, hs + o_taken, (size_t)nk * 4);
    int counts[2]; memcpy(counts, hs + o_cnt, 8);
    if (nmatches) *nmatches = counts[0];
    return UVIP_OK;
}

// haloc::Hash::getHash for nsets descriptor sets (rows start[s]..start[s+1]-1 of desc); proj = the caller's num_proj random
// projection vectors of proj_len floats each (the reference draws them from rand() seeded with time(NULL), src/hash.cpp:95)
int uvip_haloc_hash(uvip_matcher* m, const uint8_t* desc, const int32_t* start, int nsets, const float* proj, int num_proj, int proj_len, float* hash)
{
    UVIP_CHECK_ARG(m && nsets >= 0 && num_proj > 0 && proj_len > 0);
    if (nsets == 0) return UVIP_OK;
    UVIP_CHECK_ARG(start && proj && hash);
    const int nrows = start[nsets];
    UVIP_CHECK_ARG(nrows >= 0 && (nrows == 0 || desc));
    for (int s2 = 0; s2 < nsets; s2++) UVIP_CHECK_ARG(start[s2 + 1] >= start[s2] && start[s2 + 1] - start[s2] <= proj_len);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    size_t off = 0;
    auto sect = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 32); return o; };
    const size_t o_d = sect((size_t)(nrows > 0 ? nrows : 1) * 32), o_s = sect((size_t)(nsets + 1) * 4), o_p = sect((size_t)num_proj * proj_len * 4);
    const size_t o_h = sect((size_t)nsets * num_proj * 32 * 4);
    int rc;
    if ((rc = m->misc.reserve(off))) return rc;
    uint8_t* base = m->misc.as<uint8_t>();
    cudaStream_t st = m->stream;
    if (nrows) UVIP_CUDA(cudaMemcpyAsync(base + o_d, desc, (size_t)nrows * 32, cudaMemcpyHostToDevice, st));
    UVIP_CUDA(cudaMemcpyAsync(base + o_s, start, (size_t)(nsets + 1) * 4, cudaMemcpyHostToDevice, st));
    UVIP_CUDA(cudaMemcpyAsync(base + o_p, proj, (size_t)num_proj * proj_len * 4, cudaMemcpyHostToDevice, st));
    k_haloc_hash<<<nsets, 128, 0, st>>>(base + o_d, (const int32_t*)(base + o_s), (const float*)(base + o_p), num_proj, proj_len, (float*)(base + o_h));
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    UVIP_CUDA(cudaMemcpyAsync(hash, base + o_h, (size_t)nsets * num_proj * 32 * 4, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    return UVIP_OK;
}

// haloc::Hash::match of one query hash against n stored hashes of `len` floats (KeyFrameDatabase::DetectLoopCandidatesHaloc)
int uvip_haloc_match(uvip_matcher* m, const float* query, const float* table, int n, int len, float* score)
{
    UVIP_CHECK_ARG(m && n >= 0 && len > 0);
    if (n == 0) return UVIP_OK;
    UVIP_CHECK_ARG(query && table && score);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    size_t off = 0;
    auto sect = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 32); return o; };
    const size_t o_q = sect((size_t)len * 4), o_t = sect((size_t)n * len * 4), o_s = sect((size_t)n * 4);
    int rc;
    if ((rc = m->misc.reserve(off))) return rc;
    uint8_t* base = m->misc.as<uint8_t>();
    cudaStream_t st = m->stream;
    UVIP_CUDA(cudaMemcpyAsync(base + o_q, query, (size_t)len * 4, cudaMemcpyHostToDevice, st));
    UVIP_CUDA(cudaMemcpyAsync(base + o_t, table, (size_t)n * len * 4, cudaMemcpyHostToDevice, st));
    k_haloc_match<<<div_up(n, 128), 128, 0, st>>>((const float*)(base + o_q), (const float*)(base + o_t), n, len, (float*)(base + o_s));
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    UVIP_CUDA(cudaMemcpyAsync(score, base + o_s, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    return UVIP_OK;
}

// SearchForTriangulation's matching core (src/ORBmatcher.cc:893-952): see search_one_epipolar
int uvip_search_lists_epipolar(uvip_matcher* m, int th_dist, const uint8_t* qdesc, const float* qline, int nq,
                               const int32_t* cand_start, const int32_t* cand_idx, const uint8_t* kdesc, const float* kx, const float* ky,
                               const double* kthr, int nk, int32_t* taken, int32_t* match, int* nmatches)
{
    UVIP_CHECK_ARG(m && nq >= 0 && nk >= 0 && nk < (1 << 20));
    if (nmatches) *nmatches = 0;
    if (nq == 0) return UVIP_OK;
    UVIP_CHECK_ARG(qdesc && qline && cand_start && match && (nk == 0 || (kdesc && kx && ky && kthr && taken)));
    const int ncand = cand_start[nq];
    UVIP_CHECK_ARG(ncand >= 0 && (ncand == 0 || cand_idx));
    for (int i = 0; i < ncand; i++) UVIP_CHECK_ARG(cand_idx[i] >= 0 && cand_idx[i] < nk);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    size_t off = 0;
    auto sect = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 32); return o; };
    const size_t nkk = nk > 0 ? nk : 1;
    const size_t o_qd = sect((size_t)nq * 32), o_ql = sect((size_t)nq * 16), o_kd = sect(nkk * 32), o_kx = sect(nkk * 4), o_ky = sect(nkk * 4), o_kt = sect(nkk * 8);
    const size_t o_cs = sect((size_t)(nq + 1) * 4), o_ci = sect((size_t)(ncand > 0 ? ncand : 1) * 4);
    const size_t o_taken = sect(nkk * 4), o_match = sect((size_t)nq * 4), o_oa = sect(nkk * 4), o_ob = sect(nkk * 4), o_cnt = sect(32);
    int rc;
    if ((rc = m->misc.reserve(off))) return rc;
    uint8_t* base = m->misc.as<uint8_t>();
    cudaStream_t st = m->stream;
#define UP(o, src, bytes) if ((bytes) > 0) UVIP_CUDA(cudaMemcpyAsync(base + (o), (src), (bytes), cudaMemcpyHostToDevice, st))
    UP(o_qd, qdesc, (size_t)nq * 32); UP(o_ql, qline, (size_t)nq * 16); UP(o_kd, kdesc, (size_t)nk * 32); UP(o_kx, kx, (size_t)nk * 4); UP(o_ky, ky, (size_t)nk * 4);
    UP(o_kt, kthr, (size_t)nk * 8); UP(o_cs, cand_start, (size_t)(nq + 1) * 4); UP(o_ci, cand_idx, (size_t)ncand * 4); UP(o_taken, taken, (size_t)nk * 4);
#undef UP
    SearchCtx c; memset(&c, 0, sizeof(c));
    c.sp.mode = 5; c.sp.th_dist = th_dist;
    c.qdesc = base + o_qd; c.qline = (const float*)(base + o_ql); c.kdesc = base + o_kd; c.kx = (const float*)(base + o_kx); c.ky = (const float*)(base + o_ky);
    c.kthr = (const double*)(base + o_kt); c.cand_start = (const int32_t*)(base + o_cs); c.cand_idx = (const int32_t*)(base + o_ci);
    UVIP_CUDA(launch_search(st, 1, c, nq, nk, (int32_t*)(base + o_taken), (int32_t*)(base + o_match), (int*)(base + o_oa), (int*)(base + o_ob),
                            (int*)(base + o_cnt), nullptr, nullptr, 0, 0, (int*)(base + o_cnt) + 4));
    m->launches++;
    UVIP_CUDA(cudaGetLastError());
    int counts[2] = {0, 0};
    UVIP_CUDA(cudaMemcpyAsync(match, base + o_match, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (nk) UVIP_CUDA(cudaMemcpyAsync(taken, base + o_taken, (size_t)nk * 4, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaMemcpyAsync(counts, base + o_cnt, 8, cudaMemcpyDeviceToHost, st));
    UVIP_CUDA(cudaStreamSynchronize(st));
    if (nmatches) *nmatches = counts[0];
    return UVIP_OK;
}

// Batched grid-windowed search over device-resident frames (BASELINE config 3 as a throughput workload): frame f owns
// queries [f*q_stride, f*q_stride + nq[f]) and keypoints [f*k_stride, f*k_stride + nk[f]).  One launch builds all frame
// grids, one launch runs all searches (one CTA per frame: the claims of a frame are sequentially consistent inside its CTA
// and never cross frames, so frames shard with no exchange).  Nothing is copied or synchronised here.
int uvip_search_window_batch_device(uvip_matcher* m, const uvip_search_params* sp, int nframes,
                                    const float* d_qu, const float* d_qv, const float* d_qr, const int32_t* d_qmin_level,
                                    const int32_t* d_qmax_level, const uint8_t* d_qdesc, const int32_t* d_nq, int q_stride,
                                    const float* d_kx, const float* d_ky, const int32_t* d_octave, const uint8_t* d_kdesc,
                                    const int32_t* d_nk, int k_stride,
                                    int32_t* d_taken, int32_t* d_match, int32_t* d_counts, void* stream)
{
    UVIP_CHECK_ARG(m && sp && nframes >= 0 && q_stride > 0 && k_stride > 0 && sp->cols > 0 && sp->rows > 0);
    UVIP_CHECK_ARG(sp->mode == 0 || sp->mode == 1 || sp->mode == 4 || sp->mode == 6);
    if (nframes == 0) return UVIP_OK;
    UVIP_CHECK_ARG(d_qu && d_qv && d_qr && d_qmin_level && d_qmax_level && d_qdesc && d_nq && d_kx && d_ky && d_octave && d_kdesc && d_nk &&
                   d_taken && d_match && d_counts);
    const int ncell = sp->cols * sp->rows;
    UVIP_CHECK_ARG((size_t)(2 * ncell + 1) * 4 <= 200 * 1024);
    std::lock_guard<std::mutex> lk(m->mu);
    DeviceGuard g(m->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    int rc;
    const size_t nk_all = (size_t)nframes * k_stride;
    if ((rc = m->misc2.reserve((size_t)nframes * (ncell + 1) * 4))) return rc;      // cell_start
    if ((rc = m->misc3.reserve(nk_all * 4 * 2))) return rc;                         // cell_items, cell_of
    if ((rc = m->misc4.reserve(nk_all * 4 * 2 + (size_t)nframes * 16))) return rc;   // owner tables A, B, round flags
    int32_t* cs = m->misc2.as<int32_t>();
    int32_t* ci = m->misc3.as<int32_t>();
    int32_t* cof = ci + nk_all;
    const size_t smem = (size_t)(2 * ncell + 1) * 4;
    UVIP_CUDA(cudaFuncSetAttribute(k_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_grid_build<<<nframes, 1024, smem, st>>>(d_kx, d_ky, 0, sp->min_x, sp->min_y, sp->inv_w, sp->inv_h, sp->cols, sp->rows, cs, ci, cof, d_nk, k_stride);
    SearchCtx c; memset(&c, 0, sizeof(c));
    c.sp = *sp;
    c.qu = d_qu; c.qv = d_qv; c.qr = d_qr; c.qminL = d_qmin_level; c.qmaxL = d_qmax_level; c.qdesc = d_qdesc;
    c.kx = d_kx; c.ky = d_ky; c.octave = d_octave; c.kdesc = d_kdesc; c.cell_start = cs; c.cell_items = ci;
    if (!env_set("UVIP_SEARCH_NOCACHE")) {                                           // candidate cache of the claim rounds (SearchCtx)
        const size_t nq_all = (size_t)nframes * q_stride;
        if ((rc = m->qcache.reserve(nq_all * (SW_CACHE_K + 1) * 4))) return rc;
        c.qcache = m->qcache.as<unsigned>(); c.qcache_n = m->qcache.as<int>() + nq_all * SW_CACHE_K;
    }
    UVIP_CUDA(launch_search(st, nframes, c, 0, 0, d_taken, d_match, m->misc4.as<int>(), m->misc4.as<int>() + nk_all, d_counts, d_nq, d_nk, q_stride, k_stride,
                            m->misc4.as<int>() + 2 * nk_all));
    m->launches += 2;
    UVIP_CUDA(cudaGetLastError());
    return UVIP_OK;
}

/* measurement utility (not part of the reference surface): sustained __popc throughput of `device`, in popc/s */
int uvip_popc_peak(int device, int iters, double* popc_per_s)
{
    UVIP_CHECK_ARG(popc_per_s && iters > 0);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return UVIP_ERR_NO_DEVICE;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    UVIP_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    uint32_t* d_out = nullptr;
    UVIP_CUDA(cudaMalloc(&d_out, (size_t)blocks * threads * 4));
    cudaEvent_t e0, e1;
    UVIP_CUDA(cudaEventCreate(&e0)); UVIP_CUDA(cudaEventCreate(&e1));
    k_popc_peak<<<blocks, threads>>>(iters, 12345u, d_out);
    UVIP_CUDA(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        UVIP_CUDA(cudaEventRecord(e0));
        k_popc_peak<<<blocks, threads>>>(iters, 12345u + r, d_out);
        UVIP_CUDA(cudaEventRecord(e1));
        UVIP_CUDA(cudaEventSynchronize(e1));
        float ms = 0; UVIP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *popc_per_s = (double)blocks * threads * 8.0 * iters / (best * 1e-3);
    return UVIP_OK;
}

}  // extern "C"

// ---- DBoW2 vocabulary (next row N2) ----------------------------------------------------------------------
struct uvip_vocabulary {
    int device = 0, nnodes = 0, L = 0;
    cudaStream_t stream = nullptr;
    DevBuf child_start, child_ids, desc, weight, word, io;
    long long launches = 0;
    std::mutex mu;
};

extern "C" {

int uvip_vocabulary_create(int device, int nnodes, const int32_t* child_start, const int32_t* child_ids, const uint8_t* node_desc,
                           const double* node_weight, const int32_t* node_word, int L, uvip_vocabulary** out)
{
    UVIP_CHECK_ARG(out && nnodes >= 2 && child_start && child_ids && node_desc && node_weight && node_word && L >= 1);
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_last_error("no CUDA device available; libuvip_orb has no CPU fallback"); return UVIP_ERR_NO_DEVICE; }
    UVIP_CHECK_ARG(device >= 0 && device < ndev);
    UVIP_CHECK_ARG(child_start[0] == 0 && child_start[1] > 0);
    const int nchild = child_start[nnodes];
    for (int i = 0; i < nnodes; i++) UVIP_CHECK_ARG(child_start[i + 1] >= child_start[i]);
    for (int i = 0; i < nchild; i++) UVIP_CHECK_ARG(child_ids[i] > 0 && child_ids[i] < nnodes);
    DeviceGuard g(device);
    uvip_vocabulary* v = new uvip_vocabulary();
    v->device = device; v->nnodes = nnodes; v->L = L;
    int rc = 0;
    rc |= v->child_start.reserve((size_t)(nnodes + 1) * 4); rc |= v->child_ids.reserve((size_t)(nchild > 0 ? nchild : 1) * 4);
    rc |= v->desc.reserve((size_t)nnodes * 32); rc |= v->weight.reserve((size_t)nnodes * 8); rc |= v->word.reserve((size_t)nnodes * 4);
    if (rc || cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking) != cudaSuccess) { uvip_vocabulary_destroy(v); return UVIP_ERR_CUDA; }
    cudaMemcpy(v->child_start.p, child_start, (size_t)(nnodes + 1) * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(v->child_ids.p, child_ids, (size_t)nchild * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(v->desc.p, node_desc, (size_t)nnodes * 32, cudaMemcpyHostToDevice);
    cudaMemcpy(v->weight.p, node_weight, (size_t)nnodes * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(v->word.p, node_word, (size_t)nnodes * 4, cudaMemcpyHostToDevice);
    UVIP_CUDA(cudaGetLastError());
    *out = v;
    return UVIP_OK;
}

int uvip_vocabulary_destroy(uvip_vocabulary* v)
{
    if (!v) return UVIP_OK;
    DeviceGuard g(v->device);
    if (v->stream) { cudaStreamSynchronize(v->stream); cudaStreamDestroy(v->stream); }
    DevBuf* bufs[] = {&v->child_start, &v->child_ids, &v->desc, &v->weight, &v->word, &v->io};
    for (DevBuf* b : bufs) b->release();
    delete v;
    return UVIP_OK;
}

int uvip_bow_transform_device(uvip_vocabulary* v, const uint8_t* d_desc, int n, int levelsup, int32_t* d_word_id, int32_t* d_node_id,
                              double* d_weight, void* stream)
{
    UVIP_CHECK_ARG(v && n >= 0 && levelsup >= 0);
    if (n == 0) return UVIP_OK;
    UVIP_CHECK_ARG(d_desc && d_word_id && d_node_id && d_weight && ((uintptr_t)d_desc & 15) == 0);
    DeviceGuard g(v->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : v->stream;
    k_bow_transform<<<div_up(n, 8), 256, 0, st>>>(v->child_start.as<int32_t>(), v->child_ids.as<int32_t>(), v->desc.as<uint8_t>(),
                                                   v->weight.as<double>(), v->word.as<int32_t>(), v->L, d_desc, n, levelsup, d_word_id, d_node_id, d_weight);
    v->launches++;
    UVIP_CUDA(cudaGetLastError());
    return UVIP_OK;
}

int uvip_bow_transform(uvip_vocabulary* v, const uint8_t* desc, int n, int levelsup, int32_t* word_id, int32_t* node_id, double* weight)
{
    UVIP_CHECK_ARG(v && n >= 0);
    if (n == 0) return UVIP_OK;
    UVIP_CHECK_ARG(desc && word_id && node_id && weight);
    std::lock_guard<std::mutex> lk(v->mu);
    DeviceGuard g(v->device);
    int rc;
    const size_t o_w = align_up((size_t)n * 32, 256), o_n = o_w + align_up((size_t)n * 8, 256), o_i = o_n + align_up((size_t)n * 4, 256);
    if ((rc = v->io.reserve(o_i + (size_t)n * 4))) return rc;
    uint8_t* base = v->io.as<uint8_t>();
    UVIP_CUDA(cudaMemcpyAsync(base, desc, (size_t)n * 32, cudaMemcpyHostToDevice, v->stream));
    rc = uvip_bow_transform_device(v, base, n, levelsup, (int32_t*)(base + o_i), (int32_t*)(base + o_n), (double*)(base + o_w), v->stream);
    if (rc) return rc;
    UVIP_CUDA(cudaMemcpyAsync(word_id, base + o_i, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    UVIP_CUDA(cudaMemcpyAsync(node_id, base + o_n, (size_t)n * 4, cudaMemcpyDeviceToHost, v->stream));
    UVIP_CUDA(cudaMemcpyAsync(weight, base + o_w, (size_t)n * 8, cudaMemcpyDeviceToHost, v->stream));
    UVIP_CUDA(cudaStreamSynchronize(v->stream));
    return UVIP_OK;
}

}  // extern "C"
