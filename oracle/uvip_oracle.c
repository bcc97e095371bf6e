/* uvip_oracle.c — CPU ORACLE (test infrastructure, see uvip_oracle.h).
 *
 * Plain-C restatement of the reference's ORB front-end.  Every function cites the reference
 * lines it follows (paths relative to the U-VIP-SLAM checkout).  Arithmetic owned by OpenCV is
 * restated from OpenCV 3.4 semantics and pinned against cv2 4.13 by tests/golden/.
 * Build: -O2 -ffp-contract=off, no -march=native, no fast-math (reference CMakeLists.txt:19-22
 * is plain -O3 on x86-64 baseline: every float op rounds separately, no FMA).
 */
#include "uvip_oracle.h"
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EDGE 16          /* EDGE_THRESHOLD  ORBextractor.cc:78 */
#define HALF_PATCH 15    /* HALF_PATCH_SIZE ORBextractor.cc:77 */
#define PATCH 31         /* PATCH_SIZE      ORBextractor.cc:76 */
#define MAXLEV 16

static const int8_t k_pattern[1024] = {
#include "orb_pattern.inc"
};

/* cvRound: round-half-to-even (x86 cvtss2si under the default rounding mode) */
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }

struct uo_extractor {
    uo_params p;
    float scale[MAXLEV], inv_scale[MAXLEV];
    int   quota[MAXLEV];
    int   umax[HALF_PATCH + 1];
    /* per-call state */
    int   W, H;
    int   lw[MAXLEV], lh[MAXLEV];
    uint8_t* pad[MAXLEV];       /* padded (w+32)x(h+32), stride w+32; interior blurred in place later */
    uint8_t* plain[MAXLEV];     /* copy of the unblurred interior (debug tap) */
    int   blurred[MAXLEV];
    int   nraw[MAXLEV]; int *rx[MAXLEV], *ry[MAXLEV], *rs[MAXLEV];
    int   nkp[MAXLEV];  uo_keypoint* kp[MAXLEV];
};

/* ---------------------------------------------------------------- tables: ORBextractor.cc:458-512 */
uo_extractor* uo_create(const uo_params* p)
{
    if (!p || p->nlevels < 1 || p->nlevels > MAXLEV) return NULL;
    uo_extractor* ex = (uo_extractor*)calloc(1, sizeof(*ex));
    ex->p = *p;
    if (ex->p.retry_th <= 0) ex->p.retry_th = 7;
    if (ex->p.cell <= 0) ex->p.cell = 30;
    const double scaleFactor = (double)p->scale_factor;          /* member is double, ORBextractor.h:78 */
    const int nlevels = p->nlevels;
    ex->scale[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) ex->scale[i] = (float)((double)ex->scale[i - 1] * scaleFactor);   /* :463-466 */
    float invScaleFactor = (float)(1.0f / scaleFactor);                                                   /* :468 */
    ex->inv_scale[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) ex->inv_scale[i] = ex->inv_scale[i - 1] * invScaleFactor;          /* :471-472 */
    float factor = (float)(1.0 / scaleFactor);                                                            /* :478 */
    float nDesired = p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));    /* :479 */
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {                                                               /* :481-487 */
        ex->quota[l] = cv_round_f(nDesired);
        sum += ex->quota[l];
        nDesired *= factor;
    }
    ex->quota[nlevels - 1] = (p->nfeatures - sum) > 0 ? (p->nfeatures - sum) : 0;                         /* :488 */
    /* umax :496-511 */
    int v, v0, vmax = (int)floor(HALF_PATCH * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceil(HALF_PATCH * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH * HALF_PATCH;
    for (v = 0; v <= vmax; ++v) ex->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (ex->umax[v0] == ex->umax[v0 + 1]) ++v0;
        ex->umax[v] = v0;
        ++v0;
    }
    return ex;
}

static void free_call_state(uo_extractor* ex)
{
    for (int l = 0; l < MAXLEV; l++) {
        free(ex->pad[l]); free(ex->plain[l]); free(ex->rx[l]); free(ex->ry[l]); free(ex->rs[l]); free(ex->kp[l]);
        ex->pad[l] = ex->plain[l] = NULL; ex->rx[l] = ex->ry[l] = ex->rs[l] = NULL; ex->kp[l] = NULL;
        ex->nraw[l] = ex->nkp[l] = 0; ex->blurred[l] = 0;
    }
}
void uo_destroy(uo_extractor* ex) { if (ex) { free_call_state(ex); free(ex); } }

void uo_tables(const uo_extractor* ex, float* scale, float* inv_scale, int* quota, int* umax16)
{
    for (int l = 0; l < ex->p.nlevels; l++) {
        if (scale) scale[l] = ex->scale[l];
        if (inv_scale) inv_scale[l] = ex->inv_scale[l];
        if (quota) quota[l] = ex->quota[l];
    }
    if (umax16) for (int v = 0; v <= HALF_PATCH; v++) umax16[v] = ex->umax[v];
}

/* level size: ORBextractor.cc:967-968 (always from the level-0 size) */
void uo_level_size(const uo_extractor* ex, int W, int H, int level, int* w, int* h)
{
    float s = ex->inv_scale[level];
    *w = cv_round_f((float)W * s);
    *h = cv_round_f((float)H * s);
}

/* ---------------------------------------------------------------- cv::resize INTER_LINEAR, CV_8UC1
 * (called at ORBextractor.cc:982).  OpenCV's 11-bit fixed-point path: per-axis coefficient tables
 * (resizeGeneric_Invoker / HResizeLinear / VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>). */
static void resize_axis_table(int src_n, int dst_n, int* ofs, short* coef /*2 per dst*/)
{
    double inv_scale = (double)dst_n / src_n;
    double scale = 1. / inv_scale;
    for (int d = 0; d < dst_n; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (s < 0) { s = 0; f = 0; }
        if (s >= src_n - 1) { s = src_n - 1; f = 0; }
        ofs[d] = s;
        /* saturate_cast<short>(float) = cvRound + clamp */
        int a0 = cv_round_f((1.f - f) * 2048.f), a1 = cv_round_f(f * 2048.f);
        if (a0 > 32767) a0 = 32767; if (a0 < -32768) a0 = -32768;
        if (a1 > 32767) a1 = 32767; if (a1 < -32768) a1 = -32768;
        coef[2 * d] = (short)a0; coef[2 * d + 1] = (short)a1;
    }
}

void uo_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                         uint8_t* dst, int dw, int dh, int dstride)
{
    int* xofs = (int*)malloc(sizeof(int) * dw); short* alpha = (short*)malloc(sizeof(short) * 2 * dw);
    int* yofs = (int*)malloc(sizeof(int) * dh); short* beta = (short*)malloc(sizeof(short) * 2 * dh);
    int* r0 = (int*)malloc(sizeof(int) * dw); int* r1 = (int*)malloc(sizeof(int) * dw);
    resize_axis_table(sw, dw, xofs, alpha);
    resize_axis_table(sh, dh, yofs, beta);
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = yofs[dy], sy1 = sy0 + 1 < sh ? sy0 + 1 : sh - 1;
        const uint8_t* S0 = src + (size_t)sy0 * sstride; const uint8_t* S1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            r0[dx] = S0[sx] * alpha[2 * dx] + S0[sx1] * alpha[2 * dx + 1];
            r1[dx] = S1[sx] * alpha[2 * dx] + S1[sx1] * alpha[2 * dx + 1];
        }
        int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(xofs); free(alpha); free(yofs); free(beta); free(r0); free(r1);
}

/* cv::copyMakeBorder(BORDER_REFLECT_101) (ORBextractor.cc:988-989,996-997): gfedcb|abcdefgh|gfedcba */
static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * (n - 1) - p; }
    return p;
}
void uo_border_reflect101(uint8_t* buf, int w, int h, int stride, int pad)
{
    for (int y = 0; y < h + 2 * pad; y++) {
        int sy = reflect101(y - pad, h) + pad;
        uint8_t* D = buf + (size_t)y * stride; const uint8_t* S = buf + (size_t)sy * stride;
        if (y < pad || y >= h + pad) memcpy(D + pad, S + pad, (size_t)w); /* source interior rows are final */
        for (int x = 0; x < pad; x++) D[x] = D[reflect101(x - pad, w) + pad];
        for (int x = w + pad; x < w + 2 * pad; x++) D[x] = D[reflect101(x - pad, w) + pad];
    }
}

/* ---------------------------------------------------------------- cv::FAST(img, kps, th, nms) type 9_16
 * (called at ORBextractor.cc:792,797).  Corner iff >=9 contiguous ring pixels all > v+th or all < v-th;
 * score = largest t at which it is still a corner (OpenCV cornerScore<16>); NMS strict over 8 neighbours
 * with non-corners / ROI frame = 0; output order ascending y then x.  Scores live in a uchar buffer. */
static const int k_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int k_ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static int fast_corner_score(const uint8_t* p, int stride, int th)  /* returns -1 if not a corner at th */
{
    /* OpenCV-style early outs first (a 9-arc holds one pixel of every opposite ring pair), so that this CPU baseline
     * is not artificially slow; the decision and the score are those of the plain definition below. */
    const int v = p[0], lo = v - th, hi = v + th;
    int r, br, dk, b = 1, k_ = 1;
#define PAIR(o1, o2) r = p[o1]; br = r > hi; dk = r < lo; r = p[o2]; br |= r > hi; dk |= r < lo; b &= br; k_ &= dk; if (!(b | k_)) return -1;
    PAIR(3 * stride, -3 * stride) PAIR(3, -3) PAIR(2 * stride + 2, -2 * stride - 2) PAIR(-2 * stride + 2, 2 * stride - 2)
    PAIR(3 * stride + 1, -3 * stride - 1) PAIR(stride + 3, -stride - 3) PAIR(-stride + 3, stride - 3) PAIR(-3 * stride + 1, 3 * stride - 1)
#undef PAIR
    int d[25];
    for (int k = 0; k < 16; k++) d[k] = (int)p[k_ring_dy[k] * stride + k_ring_dx[k]] - v;
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int A = -256, B = -256;     /* A: best bright arc min(d); B: best dark arc min(-d) */
    for (int s = 0; s < 16; s++) {
        int mn = 256, mx = -256;
        for (int k = s; k < s + 9; k++) { if (d[k] < mn) mn = d[k]; if (d[k] > mx) mx = d[k]; }
        if (mn > A) A = mn;
        if (-mx > B) B = -mx;
    }
    int m = A > B ? A : B;
    if (m <= th) return -1;
    return m - 1;
}

int uo_fast9(const uint8_t* img, int stride, int w, int h, int th, int nms,
             int* xs, int* ys, int* scores, int cap)
{
    int n = 0;
    if (w < 7 || h < 7) return 0;
    static __thread uint8_t* scratch = NULL; static __thread size_t scratch_cap = 0;   /* per-thread, reused across the ~1000 cell calls of a frame */
    if (scratch_cap < 2 * (size_t)w * h) { free(scratch); scratch_cap = 2 * (size_t)w * h; scratch = (uint8_t*)malloc(scratch_cap); }
    uint8_t* sc = scratch; uint8_t* isc = scratch + (size_t)w * h;
    memset(scratch, 0, 2 * (size_t)w * h);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = fast_corner_score(img + (size_t)y * stride + x, stride, th);
            if (s >= 0) { isc[(size_t)y * w + x] = 1; sc[(size_t)y * w + x] = (uint8_t)s; }
        }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            if (!isc[(size_t)y * w + x]) continue;
            int s = sc[(size_t)y * w + x];
            int keep = 1;
            if (nms) {
                for (int dy = -1; dy <= 1 && keep; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (!dx && !dy) continue;
                        if (!(s > sc[(size_t)(y + dy) * w + (x + dx)])) { keep = 0; break; }
                    }
            }
            if (keep) {
                if (n < cap) { xs[n] = x; ys[n] = y; scores[n] = nms ? s : 0; }
                n++;
            }
        }
    return n;
}

/* ---------------------------------------------------------------- cv::GaussianBlur(7x7, sigma 2, REFLECT_101)
 * (called at ORBextractor.cc:942).  Canonical variant (A) of SURVEY A.6: fixed-point separable kernel
 * [18,34,48,56,48,34,18]/256, out = (sum_v k * (sum_h k*p) + 32768) >> 16.  src points at the interior
 * origin of a buffer that holds >= 3 valid border pixels on every side. */
static const int k_gauss7[7] = {18, 34, 48, 56, 48, 34, 18};
void uo_blur7(const uint8_t* src, int w, int h, int stride, uint8_t* dst, int dstride)
{
    int* hb = (int*)malloc(sizeof(int) * (size_t)w * (h + 6));
    for (int y = -3; y < h + 3; y++) {
        const uint8_t* S = src + (ptrdiff_t)y * stride;
        int* Hh = hb + (size_t)(y + 3) * w;
        for (int x = 0; x < w; x++) {
            int a = 0;
            for (int k = 0; k < 7; k++) a += k_gauss7[k] * S[x + k - 3];
            Hh[x] = a;
        }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int a = 0;
            for (int k = 0; k < 7; k++) a += k_gauss7[k] * hb[(size_t)(y + k) * w + x];
            dst[(size_t)y * dstride + x] = (uint8_t)((a + 32768) >> 16);
        }
    free(hb);
}

/* ---------------------------------------------------------------- cv::fastAtan2 (called at ORBextractor.cc:151) */
float uo_fast_atan2(float y, float x)
{
    const float sc = (float)(180 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * sc, p3 = -0.3258083974640975f * sc;
    const float p5 = 0.1555786518463281f * sc, p7 = -0.04432655554792128f * sc;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* IC_Angle: ORBextractor.cc:125-152 */
/* HarrisResponses, ORBextractor.cc:80-121 (only reachable from the dead ComputeKeyPoints path :536-746).  img = pixel (0,0)
 * of the level (unblurred), points in level coordinates. */
void uo_harris_responses(const uint8_t* img, int step, const float* xs, const float* ys, int n, int blockSize, float harris_k, float* out)
{
    const int r = blockSize / 2;
    float scale = (1 << 2) * blockSize * 255.0f;
    scale = 1.0f / scale;
    const float scale_sq_sq = scale * scale * scale * scale;
    for (int i = 0; i < n; i++) {
        const int x0 = cv_round_f(xs[i] - r), y0 = cv_round_f(ys[i] - r);
        const uint8_t* ptr0 = img + (ptrdiff_t)y0 * step + x0;
        int a = 0, b = 0, c = 0;
        for (int k = 0; k < blockSize * blockSize; k++) {
            const uint8_t* ptr = ptr0 + (k / blockSize) * step + (k % blockSize);
            const int Ix = (ptr[1] - ptr[-1]) * 2 + (ptr[-step + 1] - ptr[-step - 1]) + (ptr[step + 1] - ptr[step - 1]);
            const int Iy = (ptr[step] - ptr[-step]) * 2 + (ptr[step - 1] - ptr[-step - 1]) + (ptr[step + 1] - ptr[-step + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
        out[i] = ((float)a * b - (float)c * c - harris_k * ((float)a + b) * ((float)a + b)) * scale_sq_sq;
    }
}

float uo_ic_angle(const uint8_t* center, int step, const int* umax)
{
    int m_01 = 0, m_10 = 0;
    for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = center[u + v * step], vm = center[u - v * step];
            v_sum += (vp - vm);
            m_10 += u * (vp + vm);
        }
        m_01 += v * v_sum;
    }
    return uo_fast_atan2((float)m_01, (float)m_10);
}

/* computeOrbDescriptor: ORBextractor.cc:155-195 */
void uo_descriptor(const uint8_t* center, int step, float angle_deg, uint8_t* desc)
{
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    float angle = angle_deg * factorPI;
    float a = cosf(angle), b = sinf(angle);
    const int8_t* pat = k_pattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            float x0 = (float)pat[4 * k], y0 = (float)pat[4 * k + 1], x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
            float r0 = x0 * b, r0b = y0 * a, c0 = x0 * a, c0b = y0 * b;   /* separately rounded products */
            float r1 = x1 * b, r1b = y1 * a, c1 = x1 * a, c1b = y1 * b;
            int t0 = center[cv_round_f(r0 + r0b) * step + cv_round_f(c0 - c0b)];
            int t1 = center[cv_round_f(r1 + r1b) * step + cv_round_f(c1 - c1b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ---------------------------------------------------------------- DistributeOctTree: ORBextractor.cc:1006-1287
 * std::list is restated as an index-linked list; the sort tie-break on node POINTER (:1151, nondeterministic
 * in the reference) is pinned to the node creation sequence number. */
typedef struct {
    int ulx, uly, urx, ury, blx, bly, brx, bry;
    int* keys; int nkeys;
    int no_more;
    int prev, next;
    int seq;
} qnode;
typedef struct { qnode* a; int n, cap; int head, tail, size; int seq; } qlist;

static int ql_new(qlist* L)
{
    if (L->n == L->cap) { L->cap = L->cap ? L->cap * 2 : 64; L->a = (qnode*)realloc(L->a, sizeof(qnode) * L->cap); }
    qnode* q = &L->a[L->n]; memset(q, 0, sizeof(*q)); q->prev = q->next = -1; q->seq = L->seq++;
    return L->n++;
}
static void ql_push_front(qlist* L, int i) { qnode* q = &L->a[i]; q->prev = -1; q->next = L->head; if (L->head >= 0) L->a[L->head].prev = i; else L->tail = i; L->head = i; L->size++; }
static void ql_push_back(qlist* L, int i) { qnode* q = &L->a[i]; q->next = -1; q->prev = L->tail; if (L->tail >= 0) L->a[L->tail].next = i; else L->head = i; L->tail = i; L->size++; }
static int ql_erase(qlist* L, int i)  /* returns next */
{
    qnode* q = &L->a[i]; int nx = q->next;
    if (q->prev >= 0) L->a[q->prev].next = q->next; else L->head = q->next;
    if (q->next >= 0) L->a[q->next].prev = q->prev; else L->tail = q->prev;
    free(q->keys); q->keys = NULL; L->size--;
    return nx;
}

/* ExtractorNode::DivideNode :1231-1287.  Children are created (not yet linked); c[4] = node ids. */
static void divide_node(qlist* L, int ni, const float* x, const float* y, int c[4])
{
    for (int k = 0; k < 4; k++) c[k] = ql_new(L);       /* may realloc: re-fetch parent afterwards */
    qnode* P = &L->a[ni];
    const int halfX = (int)ceilf((float)(P->urx - P->ulx) / 2);
    const int halfY = (int)ceilf((float)(P->bry - P->uly) / 2);
    qnode *n1 = &L->a[c[0]], *n2 = &L->a[c[1]], *n3 = &L->a[c[2]], *n4 = &L->a[c[3]];
    n1->ulx = P->ulx; n1->uly = P->uly; n1->urx = P->ulx + halfX; n1->ury = P->uly;
    n1->blx = P->ulx; n1->bly = P->uly + halfY; n1->brx = P->ulx + halfX; n1->bry = P->uly + halfY;
    n2->ulx = n1->urx; n2->uly = n1->ury; n2->urx = P->urx; n2->ury = P->ury;
    n2->blx = n1->brx; n2->bly = n1->bry; n2->brx = P->urx; n2->bry = P->uly + halfY;
    n3->ulx = n1->blx; n3->uly = n1->bly; n3->urx = n1->brx; n3->ury = n1->bry;
    n3->blx = P->blx; n3->bly = P->bly; n3->brx = n1->brx; n3->bry = P->bly;
    n4->ulx = n3->urx; n4->uly = n3->ury; n4->urx = n2->brx; n4->ury = n2->bry;
    n4->blx = n3->brx; n4->bly = n3->bry; n4->brx = P->brx; n4->bry = P->bry;
    for (int k = 0; k < 4; k++) { L->a[c[k]].keys = (int*)malloc(sizeof(int) * (P->nkeys > 0 ? P->nkeys : 1)); L->a[c[k]].nkeys = 0; }
    for (int i = 0; i < P->nkeys; i++) {
        int id = P->keys[i];
        qnode* t;
        if (x[id] < (float)n1->urx) t = (y[id] < (float)n1->bry) ? n1 : n3;
        else                        t = (y[id] < (float)n1->bry) ? n2 : n4;
        t->keys[t->nkeys++] = id;
    }
    for (int k = 0; k < 4; k++) if (L->a[c[k]].nkeys == 1) L->a[c[k]].no_more = 1;
}

typedef struct { int size; int seq; int node; } qsp;
static int qsp_cmp(const void* a, const void* b)
{
    const qsp* A = (const qsp*)a; const qsp* B = (const qsp*)b;
    if (A->size != B->size) return A->size < B->size ? -1 : 1;
    return A->seq < B->seq ? -1 : (A->seq > B->seq);
}

int uo_distribute_octtree(const float* x, const float* y, const float* resp, int n,
                          int minX, int maxX, int minY, int maxY, int N, int* out_idx, int cap)
{
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));            /* :1010 */
    if (nIni < 1) return 0;   /* reference would divide by zero; portrait images with ratio < 0.5 are unsupported */
    const float hX = (float)(maxX - minX) / nIni;                                   /* :1012 */
    qlist L; memset(&L, 0, sizeof(L)); L.head = L.tail = -1;
    int* ini = (int*)malloc(sizeof(int) * nIni);
    for (int i = 0; i < nIni; i++) {                                                /* :1019-1031 */
        int id = ql_new(&L); qnode* q = &L.a[id];
        q->ulx = (int)(hX * (float)i); q->uly = 0;
        q->urx = (int)(hX * (float)(i + 1)); q->ury = 0;
        q->blx = q->ulx; q->bly = maxY - minY;
        q->brx = q->urx; q->bry = maxY - minY;
        q->keys = (int*)malloc(sizeof(int) * (n > 0 ? n : 1)); q->nkeys = 0;
        ql_push_back(&L, id); ini[i] = id;
    }
    for (int i = 0; i < n; i++) {                                                   /* :1034-1038 */
        int r = (int)(x[i] / hX);
        if (r < 0) r = 0; if (r >= nIni) r = nIni - 1;   /* never taken for in-window keys */
        qnode* q = &L.a[ini[r]]; q->keys[q->nkeys++] = i;
    }
    free(ini);
    for (int it = L.head; it >= 0;) {                                               /* :1042-1053 */
        if (L.a[it].nkeys == 1) { L.a[it].no_more = 1; it = L.a[it].next; }
        else if (L.a[it].nkeys == 0) it = ql_erase(&L, it);
        else it = L.a[it].next;
    }
    int finish = 0;
    qsp* V = NULL; int nV = 0, capV = 0;
#define V_PUSH(sz, nd) do { if (nV == capV) { capV = capV ? capV * 2 : 256; V = (qsp*)realloc(V, sizeof(qsp) * capV); } \
        V[nV].size = (sz); V[nV].seq = L.a[nd].seq; V[nV].node = (nd); nV++; } while (0)
    while (!finish) {                                                               /* :1062 */
        int prevSize = L.size, nToExpand = 0;
        nV = 0;
        for (int it = L.head; it >= 0;) {                                           /* :1074-1137 */
            if (L.a[it].no_more) { it = L.a[it].next; continue; }
            int c[4]; divide_node(&L, it, x, y, c);
            for (int k = 0; k < 4; k++) {
                if (L.a[c[k]].nkeys > 0) {
                    ql_push_front(&L, c[k]);
                    if (L.a[c[k]].nkeys > 1) { nToExpand++; V_PUSH(L.a[c[k]].nkeys, c[k]); }
                } else { free(L.a[c[k]].keys); L.a[c[k]].keys = NULL; }
            }
            it = ql_erase(&L, it);
        }
        if (L.size >= N || L.size == prevSize) finish = 1;                          /* :1141-1144 */
        else if (L.size + nToExpand * 3 > N) {                                      /* :1145 */
            while (!finish) {
                prevSize = L.size;
                int nP = nV; qsp* P = (qsp*)malloc(sizeof(qsp) * (nP > 0 ? nP : 1));
                memcpy(P, V, sizeof(qsp) * nP); nV = 0;
                qsort(P, nP, sizeof(qsp), qsp_cmp);                                 /* :1155 (pointer -> seq) */
                for (int j = nP - 1; j >= 0; j--) {                                 /* :1156-1199 */
                    int c[4]; divide_node(&L, P[j].node, x, y, c);
                    for (int k = 0; k < 4; k++) {
                        if (L.a[c[k]].nkeys > 0) {
                            ql_push_front(&L, c[k]);
                            if (L.a[c[k]].nkeys > 1) V_PUSH(L.a[c[k]].nkeys, c[k]);
                        } else { free(L.a[c[k]].keys); L.a[c[k]].keys = NULL; }
                    }
                    ql_erase(&L, P[j].node);
                    if (L.size >= N) break;
                }
                free(P);
                if (L.size >= N || L.size == prevSize) finish = 1;                  /* :1201-1202 */
            }
        }
    }
    int nout = 0;
    for (int it = L.head; it >= 0; it = L.a[it].next) {                             /* :1208-1227 */
        qnode* q = &L.a[it];
        int best = q->keys[0]; float mr = resp[best];
        for (int k = 1; k < q->nkeys; k++) if (resp[q->keys[k]] > mr) { best = q->keys[k]; mr = resp[best]; }
        if (nout < cap) out_idx[nout] = best;
        nout++;
    }
    for (int i = 0; i < L.n; i++) free(L.a[i].keys);
    free(L.a); free(V);
    return nout;
}

/* ---------------------------------------------------------------- ComputePyramid: ORBextractor.cc:963-1004 */
static void compute_pyramid(uo_extractor* ex, const uint8_t* img, int w, int h, int stride)
{
    free_call_state(ex);
    ex->W = w; ex->H = h;
    for (int l = 0; l < ex->p.nlevels; l++) {
        int lw, lh; uo_level_size(ex, w, h, l, &lw, &lh);
        ex->lw[l] = lw; ex->lh[l] = lh;
        int ps = lw + 2 * EDGE;
        ex->pad[l] = (uint8_t*)calloc((size_t)ps * (lh + 2 * EDGE), 1);
        uint8_t* inner = ex->pad[l] + (size_t)EDGE * ps + EDGE;
        if (l == 0) for (int y = 0; y < lh; y++) memcpy(inner + (size_t)y * ps, img + (size_t)y * stride, (size_t)lw);
        else uo_resize_linear_u8(ex->pad[l - 1] + (size_t)EDGE * (ex->lw[l - 1] + 2 * EDGE) + EDGE, ex->lw[l - 1], ex->lh[l - 1],
                                 ex->lw[l - 1] + 2 * EDGE, inner, lw, lh, ps);
        uo_border_reflect101(ex->pad[l], lw, lh, ps, EDGE);
        ex->plain[l] = (uint8_t*)malloc((size_t)lw * lh);
        for (int y = 0; y < lh; y++) memcpy(ex->plain[l] + (size_t)y * lw, inner + (size_t)y * ps, (size_t)lw);
    }
}

/* ComputeKeyPointsOctTree: ORBextractor.cc:748-836 */
static void detect_level(uo_extractor* ex, int level)
{
    const int cols = ex->lw[level], rows = ex->lh[level], ps = cols + 2 * EDGE;
    const uint8_t* inner = ex->pad[level] + (size_t)EDGE * ps + EDGE;
    const float Wc = (float)ex->p.cell;
    const int minBX = EDGE - 3, minBY = minBX, maxBX = cols - EDGE + 3, maxBY = rows - EDGE + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / Wc), nRows = (int)(height / Wc);
    int capraw = 1024, nraw = 0;
    int* rx = (int*)malloc(sizeof(int) * capraw); int* ry = (int*)malloc(sizeof(int) * capraw); int* rs = (int*)malloc(sizeof(int) * capraw);
    if (nCols > 0 && nRows > 0 && width > 0 && height > 0) {
        const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
        const int ccap = (wCell + 6) * (hCell + 6);
        int* cx = (int*)malloc(sizeof(int) * ccap); int* cy = (int*)malloc(sizeof(int) * ccap); int* cs = (int*)malloc(sizeof(int) * ccap);
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = (float)maxBY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = (float)maxBX;
                const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
                const uint8_t* roi = inner + (ptrdiff_t)y0 * ps + x0;
                int nc = uo_fast9(roi, ps, x1 - x0, y1 - y0, ex->p.fast_th, 1, cx, cy, cs, ccap);
                if (nc == 0) nc = uo_fast9(roi, ps, x1 - x0, y1 - y0, ex->p.retry_th, 1, cx, cy, cs, ccap);
                for (int k = 0; k < nc; k++) {
                    if (nraw == capraw) { capraw *= 2; rx = (int*)realloc(rx, sizeof(int) * capraw); ry = (int*)realloc(ry, sizeof(int) * capraw); rs = (int*)realloc(rs, sizeof(int) * capraw); }
                    rx[nraw] = cx[k] + j * wCell; ry[nraw] = cy[k] + i * hCell; rs[nraw] = cs[k]; nraw++;
                }
            }
        }
        free(cx); free(cy); free(cs);
    }
    ex->rx[level] = rx; ex->ry[level] = ry; ex->rs[level] = rs; ex->nraw[level] = nraw;
    /* quadtree + bookkeeping :814-831 */
    float* fx = (float*)malloc(sizeof(float) * (nraw + 1)); float* fy = (float*)malloc(sizeof(float) * (nraw + 1)); float* fr = (float*)malloc(sizeof(float) * (nraw + 1));
    for (int k = 0; k < nraw; k++) { fx[k] = (float)rx[k]; fy[k] = (float)ry[k]; fr[k] = (float)rs[k]; }
    int capw = nraw + 1; int* win = (int*)malloc(sizeof(int) * capw);
    int nw = nraw ? uo_distribute_octtree(fx, fy, fr, nraw, minBX, maxBX, minBY, maxBY, ex->quota[level], win, capw) : 0;
    ex->kp[level] = (uo_keypoint*)malloc(sizeof(uo_keypoint) * (nw + 1)); ex->nkp[level] = nw;
    const int scaledPatchSize = (int)(PATCH * ex->scale[level]);
    for (int k = 0; k < nw; k++) {
        uo_keypoint* q = &ex->kp[level][k];
        q->x = fx[win[k]] + (float)minBX; q->y = fy[win[k]] + (float)minBY;
        q->size = (float)scaledPatchSize; q->angle = -1.f; q->response = fr[win[k]];
        q->octave = level; q->class_id = -1;
    }
    free(fx); free(fy); free(fr); free(win);
    /* computeOrientation :833-835, on the unblurred level */
    for (int k = 0; k < nw; k++) {
        uo_keypoint* q = &ex->kp[level][k];
        const uint8_t* c = inner + (ptrdiff_t)cv_round_f(q->y) * ps + cv_round_f(q->x);
        q->angle = uo_ic_angle(c, ps, ex->umax);
    }
}

/* operator(): ORBextractor.cc:849-961 */
int uo_extract(uo_extractor* ex, const uint8_t* img, int w, int h, int stride,
               uo_keypoint* kps, int* n_inout, int cap, uint8_t* desc,
               int32_t* grid, int grid_rows, int grid_cols, int min_px_dist,
               int full_detect, int num_needed)
{
    if (!ex || !n_inout) return -1;
    if (!img || w <= 0 || h <= 0) return 0;                       /* :852-853 empty image: outputs untouched */
    const int nlevels = ex->p.nlevels;
    compute_pyramid(ex, img, w, h, stride);                        /* :859 */
    /* ComputeKeyPointsCopy :523-534 — incoming keypoints become level-0 points with fresh angles */
    int n_in = *n_inout; if (n_in < 0) n_in = 0;
    uo_keypoint* incoming = (uo_keypoint*)malloc(sizeof(uo_keypoint) * (n_in + 1));
    {
        const int ps = ex->lw[0] + 2 * EDGE; const uint8_t* inner = ex->pad[0] + (size_t)EDGE * ps + EDGE;
        for (int i = 0; i < n_in; i++) {
            incoming[i] = kps[i];
            incoming[i].angle = uo_ic_angle(inner + (ptrdiff_t)cv_round_f(kps[i].y) * ps + cv_round_f(kps[i].x), ps, ex->umax);
        }
    }
    for (int l = 0; l < nlevels; l++) detect_level(ex, l);         /* :865 */
    /* selection :872-915 */
    uo_keypoint* all[MAXLEV]; int nall[MAXLEV];
    for (int l = 0; l < nlevels; l++) { all[l] = (uo_keypoint*)malloc(sizeof(uo_keypoint) * (ex->nkp[l] + n_in + 1)); nall[l] = 0; }
    if (!full_detect) {
        for (int i = 0; i < n_in; i++) all[0][nall[0]++] = incoming[i];
        int Total = 0, KP = 0, brk = 0;
        (void)grid_cols;
        for (int l = 0; l < nlevels; l++) {
            if (ex->nkp[l] == 0) continue;
            int quota = num_needed * (8 - l) / 30;
            float scale = ex->scale[l];
            for (int k = 0; k < ex->nkp[l]; k++) {
                const uo_keypoint* q = &ex->kp[l][k];
                float tx = q->x * scale, ty = q->y * scale;
                int r = (int)(ty / min_px_dist), c = (int)(tx / min_px_dist);
                int32_t* cellp = &grid[(size_t)c * grid_rows + r];
                if (*cellp > 0) continue;
                all[l][nall[l]++] = *q;
                (*cellp)++;
                KP++; Total++;
                if (KP == quota) { KP = 0; break; }
                if (Total == num_needed) { brk = 1; break; }
            }
            if (brk) break;
        }
    } else {
        for (int l = 0; l < nlevels; l++) { memcpy(all[l], ex->kp[l], sizeof(uo_keypoint) * ex->nkp[l]); nall[l] = ex->nkp[l]; }
    }
    int total = 0; for (int l = 0; l < nlevels; l++) total += nall[l];
    int rc = 0;
    if (total > cap) rc = -2;
    else {
        int off = 0;
        for (int l = 0; l < nlevels; l++) {
            if (nall[l] == 0) continue;
            const int lw = ex->lw[l], lh = ex->lh[l], ps = lw + 2 * EDGE;
            uint8_t* inner = ex->pad[l] + (size_t)EDGE * ps + EDGE;
            uint8_t* tmp = (uint8_t*)malloc((size_t)lw * lh);                  /* :941-942 in-place blur of the ROI */
            uo_blur7(inner, lw, lh, ps, tmp, lw);
            for (int y = 0; y < lh; y++) memcpy(inner + (size_t)y * ps, tmp + (size_t)y * lw, (size_t)lw);
            free(tmp); ex->blurred[l] = 1;
            for (int k = 0; k < nall[l]; k++) {                                   /* :944-946 */
                uo_keypoint* q = &all[l][k];
                uo_descriptor(inner + (ptrdiff_t)cv_round_f(q->y) * ps + cv_round_f(q->x), ps, q->angle, desc + (size_t)(off + k) * 32);
            }
            if (l != 0) { float s = ex->scale[l]; for (int k = 0; k < nall[l]; k++) { all[l][k].x *= s; all[l][k].y *= s; } }  /* :951-957 */
            memcpy(kps + off, all[l], sizeof(uo_keypoint) * nall[l]);
            off += nall[l];
        }
        *n_inout = total;
    }
    for (int l = 0; l < nlevels; l++) free(all[l]);
    free(incoming);
    return rc;
}

int uo_get_level(const uo_extractor* ex, int level, int blurred, uint8_t* dst, int dstride)
{
    if (level < 0 || level >= ex->p.nlevels || !ex->pad[level]) return -1;
    const int lw = ex->lw[level], lh = ex->lh[level], ps = lw + 2 * EDGE;
    if (blurred) {
        if (ex->blurred[level]) { const uint8_t* inner = ex->pad[level] + (size_t)EDGE * ps + EDGE; for (int y = 0; y < lh; y++) memcpy(dst + (size_t)y * dstride, inner + (size_t)y * ps, (size_t)lw); }
        else {  /* level had no keypoints: compute what the blur would give */
            uint8_t* padc = (uint8_t*)malloc((size_t)ps * (lh + 2 * EDGE)); memcpy(padc, ex->pad[level], (size_t)ps * (lh + 2 * EDGE));
            uo_blur7(padc + (size_t)EDGE * ps + EDGE, lw, lh, ps, dst, dstride); free(padc);
        }
    } else for (int y = 0; y < lh; y++) memcpy(dst + (size_t)y * dstride, ex->plain[level] + (size_t)y * lw, (size_t)lw);
    return 0;
}
int uo_get_padded_level(const uo_extractor* ex, int level, uint8_t* dst)
{
    if (level < 0 || level >= ex->p.nlevels || !ex->pad[level]) return -1;
    memcpy(dst, ex->pad[level], (size_t)(ex->lw[level] + 2 * EDGE) * (ex->lh[level] + 2 * EDGE));
    return 0;
}
int uo_get_raw_corners(const uo_extractor* ex, int level, int* xs, int* ys, int* scores, int cap)
{
    if (level < 0 || level >= ex->p.nlevels) return -1;
    int n = ex->nraw[level];
    for (int k = 0; k < n && k < cap; k++) { xs[k] = ex->rx[level][k]; ys[k] = ex->ry[level][k]; scores[k] = ex->rs[level][k]; }
    return n;
}
int uo_get_level_keypoints(const uo_extractor* ex, int level, uo_keypoint* out, int cap)
{
    if (level < 0 || level >= ex->p.nlevels) return -1;
    int n = ex->nkp[level];
    for (int k = 0; k < n && k < cap; k++) out[k] = ex->kp[level][k];
    return n;
}

int uo_extract_batch(const uo_params* p, const uint8_t* frames, int nframes, int w, int h,
                     uo_keypoint* kps, int* n_out, int cap, uint8_t* desc, int threads)
{
    int err = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel
    {
        uo_extractor* ex = uo_create(p);
#pragma omp for schedule(dynamic, 1)
        for (int f = 0; f < nframes; f++) {
            int n = 0;
            int rc = uo_extract(ex, frames + (size_t)f * w * h, w, h, w, kps + (size_t)f * cap, &n, cap,
                                desc + (size_t)f * cap * 32, NULL, 0, 0, 1, 1, 0);
            n_out[f] = n;
            if (rc) {
#pragma omp atomic write
                err = rc;
            }
        }
        uo_destroy(ex);
    }
    return err;
}

/* ================================================================ matcher */
/* DescriptorDistance: ORBmatcher.cc:1794-1810 (parallel bit count over 8 x int32) */
int uo_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t wa, wb; memcpy(&wa, a + 4 * i, 4); memcpy(&wb, b + 4 * i, 4);
        uint32_t v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

/* brute-force k=2: the inner scan every ORBmatcher search shares (e.g. :201-226) == OpenCV
 * BFMatcher(NORM_HAMMING).knnMatch(k=2) used by include/utils.h:81-111 — strict '<' running top-2,
 * first (lowest) train index wins ties.  idx2/dist2 = [best, second] per query; -1 / INT_MAX-free: missing
 * neighbours get idx -1 and dist 257 (> any real distance). */
void uo_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx2, int32_t* dist2, int threads)
{
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; i++) {
        const uint64_t* a = (const uint64_t*)(q + (size_t)i * 32);
        uint64_t a0, a1, a2, a3; memcpy(&a0, a, 8); memcpy(&a1, a + 1, 8); memcpy(&a2, a + 2, 8); memcpy(&a3, a + 3, 8);
        int b1 = 257, b2 = 257, i1 = -1, i2 = -1;
        for (int j = 0; j < nt; j++) {
            uint64_t c0, c1, c2, c3; const uint8_t* tp = t + (size_t)j * 32;
            memcpy(&c0, tp, 8); memcpy(&c1, tp + 8, 8); memcpy(&c2, tp + 16, 8); memcpy(&c3, tp + 24, 8);
            int d = __builtin_popcountll(a0 ^ c0) + __builtin_popcountll(a1 ^ c1) + __builtin_popcountll(a2 ^ c2) + __builtin_popcountll(a3 ^ c3);
            if (d < b1) { b2 = b1; i2 = i1; b1 = d; i1 = j; }
            else if (d < b2) { b2 = d; i2 = j; }
        }
        idx2[2 * i] = i1; idx2[2 * i + 1] = i2; dist2[2 * i] = b1; dist2[2 * i + 1] = b2;
    }
}

/* ratio test of haloc::Utils::ratioMatching: include/utils.h:104-108 (float distances, double ratio) */
int uo_ratio_filter(const int32_t* idx2, const int32_t* dist2, int nq, double ratio, int32_t* match_train)
{
    int n = 0;
    for (int i = 0; i < nq; i++) {
        match_train[i] = -1;
        if (idx2[2 * i] < 0 || idx2[2 * i + 1] < 0) continue;          /* knn_matches[m].size() < 2 */
        float d0 = (float)dist2[2 * i], d1 = (float)dist2[2 * i + 1];
        if ((double)d0 <= (double)d1 * ratio) { match_train[i] = idx2[2 * i]; n++; }
    }
    return n;
}

/* rotation histogram: bin code ORBmatcher.cc:232-241, ComputeThreeMaxima :1748-1789, rollback :263-281 */
int uo_rot_hist_filter(int32_t* match, int n, const float* angle_a, const float* angle_b)
{
    enum { HL = 30 };
    const float factor = 1.0f / HL;
    int cnt[HL]; memset(cnt, 0, sizeof(cnt));
    int* bin_of = (int*)malloc(sizeof(int) * (n + 1));
    for (int i = 0; i < n; i++) {
        bin_of[i] = -1;
        if (match[i] < 0) continue;
        float rot = angle_a[i] - angle_b[match[i]];
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)roundf(rot * factor);
        if (bin == HL) bin = 0;
        if (bin < 0 || bin >= HL) { bin_of[i] = -2; continue; }   /* ROS_ASSERT in the reference; unreachable for angles in [0,360) */
        bin_of[i] = bin; cnt[bin]++;
    }
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
    for (int i = 0; i < HL; i++) {
        const int s = cnt[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
    int kept = 0;
    for (int i = 0; i < n; i++) {
        if (match[i] < 0) continue;
        int b = bin_of[i];
        if (b >= 0 && b != ind1 && b != ind2 && b != ind3) match[i] = -1; else kept++;
    }
    free(bin_of);
    return kept;
}

/* frame grid fill: FrameKTL.cc:250-264 + PosInGrid :426-436 (round, not floor) */
void uo_grid_build(const float* kx, const float* ky, int n, float minX, float minY, float inv_w, float inv_h,
                   int cols, int rows, int32_t* cell_start, int32_t* cell_items)
{
    int nc = cols * rows;
    int* cell = (int*)malloc(sizeof(int) * (n + 1));
    memset(cell_start, 0, sizeof(int32_t) * (nc + 1));
    for (int i = 0; i < n; i++) {
        int px = (int)roundf((kx[i] - minX) * inv_w), py = (int)roundf((ky[i] - minY) * inv_h);
        if (px < 0 || px >= cols || py < 0 || py >= rows) { cell[i] = -1; continue; }
        cell[i] = px * rows + py; cell_start[cell[i] + 1]++;
    }
    for (int c = 0; c < nc; c++) cell_start[c + 1] += cell_start[c];
    int* fill = (int*)calloc(nc + 1, sizeof(int));
    for (int i = 0; i < n; i++) if (cell[i] >= 0) cell_items[cell_start[cell[i]] + fill[cell[i]]++] = i;
    free(fill); free(cell);
}

/* FrameKTL::GetFeaturesInArea: FrameKTL.cc:359-424 */
int uo_features_in_area(const float* kx, const float* ky, const int32_t* octave,
                        const int32_t* cell_start, const int32_t* cell_items,
                        float minX, float minY, float inv_w, float inv_h, int cols, int rows,
                        float x, float y, float r, int minLevel, int maxLevel, int32_t* out, int cap)
{
    int n = 0;
    int cx0 = (int)floorf((x - minX - r) * inv_w); if (cx0 < 0) cx0 = 0; if (cx0 >= cols) return 0;
    int cx1 = (int)ceilf((x - minX + r) * inv_w); if (cx1 > cols - 1) cx1 = cols - 1; if (cx1 < 0) return 0;
    int cy0 = (int)floorf((y - minY - r) * inv_h); if (cy0 < 0) cy0 = 0; if (cy0 >= rows) return 0;
    int cy1 = (int)ceilf((y - minY + r) * inv_h); if (cy1 > rows - 1) cy1 = rows - 1; if (cy1 < 0) return 0;
    int check = 1, same = 0;
    if (minLevel == -1 && maxLevel == -1) check = 0; else if (minLevel == maxLevel) same = 1;
    for (int ix = cx0; ix <= cx1; ix++)
        for (int iy = cy0; iy <= cy1; iy++) {
            int c = ix * rows + iy;
            for (int j = cell_start[c]; j < cell_start[c + 1]; j++) {
                int id = cell_items[j];
                if (check && !same) { if (octave[id] < minLevel || octave[id] > maxLevel) continue; }
                else if (same) { if (octave[id] != minLevel) continue; }
                if (fabsf(kx[id] - x) > r || fabsf(ky[id] - y) > r) continue;
                if (n < cap) out[n] = id;
                n++;
            }
        }
    return n;
}

float uo_radius_by_viewing_cos(float view_cos) { return ((double)view_cos > 0.998) ? 2.5f : 4.0f; }

/* windowed search with claims.  mode 0 = SearchByProjection(F, MPs, th) ORBmatcher.cc:49-125;
 * mode 1 = inner loop of SearchByProjection(F, KF, found, th, ORBdist) :1683-1715. */
int uo_search_window(const uo_search_params* sp,
                     const float* qu, const float* qv, const float* qr, const int32_t* qminL, const int32_t* qmaxL,
                     const uint8_t* qdesc, int nq,
                     const float* kx, const float* ky, const int32_t* octave, const uint8_t* kdesc, int nk,
                     const int32_t* cell_start, const int32_t* cell_items,
                     float minX, float minY, float inv_w, float inv_h, int cols, int rows,
                     int32_t* taken, int32_t* match_of_query)
{
    int nmatches = 0;
    int32_t* cand = (int32_t*)malloc(sizeof(int32_t) * (nk + 1));
    for (int q = 0; q < nq; q++) {
        match_of_query[q] = -1;
        int nc = uo_features_in_area(kx, ky, octave, cell_start, cell_items, minX, minY, inv_w, inv_h, cols, rows,
                                     qu[q], qv[q], qr[q], qminL[q], qmaxL[q], cand, nk);
        if (nc == 0) continue;
        const uint8_t* d = qdesc + (size_t)q * 32;
        if (sp->mode == 0) {
            int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
            for (int c = 0; c < nc; c++) {
                int idx = cand[c];
                if (taken[idx] != -1) continue;
                int dist = uo_descriptor_distance(d, kdesc + (size_t)idx * 32);
                if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = octave[idx]; bestIdx = idx; }
                else if (dist < bestDist2) { bestLevel2 = octave[idx]; bestDist2 = dist; }
            }
            if (bestDist <= sp->th_dist) {
                if (bestLevel == bestLevel2 && (float)bestDist > sp->ratio * (float)bestDist2) continue;
                if (bestIdx < 0) continue;   /* th_dist >= 256 with every candidate taken: reference would index -1 */
                taken[bestIdx] = q; match_of_query[q] = bestIdx; nmatches++;
            }
        } else {
            /* mode 1: src/ORBmatcher.cc:1683-1715 / :372-399 (claims);  mode 4: Fuse :1075-1100 (no claims: taken[] untouched) */
            int bestDist = 2147483647, bestIdx = -1;
            for (int c = 0; c < nc; c++) {
                int idx = cand[c];
                if (sp->mode != 4 && taken[idx] != -1) continue;
                int dist = uo_descriptor_distance(d, kdesc + (size_t)idx * 32);
                if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
            }
            if (bestDist <= sp->th_dist) { if (sp->mode != 4) taken[bestIdx] = q; match_of_query[q] = bestIdx; nmatches++; }
        }
    }
    free(cand);
    return nmatches;
}

/* candidate-list search with claims: the node-restricted loops of SearchByBoW.
 * mode 2 = SearchByBoW(KeyFrame*, FrameKTL&, ...) ORBmatcher.cc:186-245: top-2 from INT_MAX, accept best <= th && (float)best < ratio*(float)best2
 * mode 3 = SearchByBoW(KeyFrame*, KeyFrame*, ...) ORBmatcher.cc:751-811: same scan, accept best < th (strict) && ratio
 * mode 1 = best only, accept best <= th (Fuse / SearchByProjection(KF,Scw) with host-built candidate lists, :1075-1100)
 * queries are visited in array order (the caller lists them in the reference's node-major order). */
int uo_search_lists(int mode, int th_dist, float ratio, const uint8_t* qdesc, int nq,
                    const int32_t* cand_start, const int32_t* cand_idx, const uint8_t* kdesc, int nk,
                    int32_t* taken, int32_t* match_of_query)
{
    int nmatches = 0;
    (void)nk;
    for (int q = 0; q < nq; q++) {
        match_of_query[q] = -1;
        const uint8_t* d = qdesc + (size_t)q * 32;
        int best1 = 2147483647, best2 = 2147483647, bestIdx = -1;
        for (int c = cand_start[q]; c < cand_start[q + 1]; c++) {
            const int idx = cand_idx[c];
            if (taken[idx] != -1) continue;
            const int dist = uo_descriptor_distance(d, kdesc + (size_t)idx * 32);
            if (dist < best1) { best2 = best1; best1 = dist; bestIdx = idx; }
            else if (dist < best2) best2 = dist;
        }
        int ok;
        if (mode == 1) ok = best1 <= th_dist;
        else if (mode == 2) ok = best1 <= th_dist && (float)best1 < ratio * (float)best2;
        else ok = best1 < th_dist && (float)best1 < ratio * (float)best2;
        if (ok && bestIdx >= 0) { taken[bestIdx] = q; match_of_query[q] = bestIdx; nmatches++; }
    }
    return nmatches;
}

/* SearchForTriangulation's matching core, ORBmatcher.cc:893-952 + CheckDistEpipolarLine :136-153.  qline = (a, b, c, den) per
 * query in float; kthr = 3.84 * sigma2(octave) per keypoint.  Sequential, claims through taken[]. */
typedef struct { int dist; int idx; } uo_di;
static int uo_di_cmp(const void* a, const void* b)
{
    const uo_di* x = (const uo_di*)a; const uo_di* y = (const uo_di*)b;
    if (x->dist != y->dist) return x->dist < y->dist ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
int uo_search_lists_epipolar(int th_dist, const uint8_t* qdesc, const float* qline, int nq, const int32_t* cand_start, const int32_t* cand_idx,
                             const uint8_t* kdesc, const float* kx, const float* ky, const double* kthr, int nk, int32_t* taken, int32_t* match_of_query)
{
    int nmatches = 0;
    uo_di* v = (uo_di*)malloc(sizeof(uo_di) * (size_t)(nk + 1));
    for (int q = 0; q < nq; q++) {
        match_of_query[q] = -1;
        int n = 0;
        for (int c = cand_start[q]; c < cand_start[q + 1]; c++) {
            const int idx = cand_idx[c];
            if (taken[idx] != -1) continue;
            const int dist = uo_descriptor_distance(qdesc + (size_t)q * 32, kdesc + (size_t)idx * 32);
            if (dist > th_dist) continue;
            v[n].dist = dist; v[n].idx = idx; n++;
        }
        if (n == 0) continue;
        qsort(v, (size_t)n, sizeof(uo_di), uo_di_cmp);
        const int distTh = (int)round(2 * v[0].dist);
        const float a = qline[4 * q], b = qline[4 * q + 1], cc = qline[4 * q + 2], den = qline[4 * q + 3];
        for (int i = 0; i < n; i++) {
            if (v[i].dist > distTh) break;
            if (den == 0) continue;
            const float num = a * kx[v[i].idx] + b * ky[v[i].idx] + cc;
            const float dsqr = num * num / den;
            if ((double)dsqr < kthr[v[i].idx]) { taken[v[i].idx] = q; match_of_query[q] = v[i].idx; nmatches++; break; }
        }
    }
    free(v);
    return nmatches;
}

/* haloc::Hash::getHash (hash.cpp:57-85) on one descriptor set and haloc::Hash::match (:190-206) */
void uo_haloc_hash(const uint8_t* desc, int rows, const float* proj, int num_proj, int proj_len, float* hash)
{
    int k = 0;
    for (int i = 0; i < num_proj; i++)
        for (int n = 0; n < 32; n++) {
            float desc_sum = 0.0f;
            for (int m = 0; m < rows; m++) desc_sum += proj[(size_t)i * proj_len + m] * (float)desc[(size_t)m * 32 + n];
            hash[k++] = rows > 0 ? desc_sum / (float)rows : 0.0f;
        }
}
float uo_haloc_match(const float* a, const float* b, int n)
{
    float sum = 0.0f;
    for (int i = 0; i < n; i++) sum += fabsf(a[i] - b[i]);
    return sum;
}

/* ================================================================ CLAHE (next row N3)
 * cv::createCLAHE(clip, Size(tx,ty))->apply(im, im) as called at Tracking.cc:425-431 (clip 4, 12x12 tiles), restated from
 * OpenCV imgproc/clahe.cpp (8-bit path): pad to a tile multiple with REFLECT_101, per-tile clipped + redistributed
 * histogram -> LUT = saturate(cumsum * 255/tileArea), bilinear blend of the four neighbouring tile LUTs in float. */
void uo_clahe(const uint8_t* src, int w, int h, int stride, double clip_limit, int tiles_x, int tiles_y, uint8_t* dst, int dstride)
{
    const int histSize = 256;
    int ew = w, eh = h;
    if (w % tiles_x != 0 || h % tiles_y != 0) { ew = w + (tiles_x - (w % tiles_x)); eh = h + (tiles_y - (h % tiles_y)); }
    uint8_t* ext = (uint8_t*)malloc((size_t)ew * eh);
    for (int y = 0; y < eh; y++) {
        const int sy = y < h ? y : 2 * (h - 1) - y;
        for (int x = 0; x < ew; x++) { const int sx = x < w ? x : 2 * (w - 1) - x; ext[(size_t)y * ew + x] = src[(size_t)sy * stride + sx]; }
    }
    const int tw = ew / tiles_x, th = eh / tiles_y;
    const int tileSizeTotal = tw * th;
    const float lutScale = (float)(histSize - 1) / tileSizeTotal;
    int clipLimit = 0;
    if (clip_limit > 0.0) { clipLimit = (int)(clip_limit * tileSizeTotal / histSize); if (clipLimit < 1) clipLimit = 1; }
    uint8_t* lut = (uint8_t*)malloc((size_t)tiles_x * tiles_y * histSize);
    for (int ty = 0; ty < tiles_y; ty++)
        for (int tx = 0; tx < tiles_x; tx++) {
            int hist[256]; memset(hist, 0, sizeof(hist));
            for (int y = 0; y < th; y++) for (int x = 0; x < tw; x++) hist[ext[(size_t)(ty * th + y) * ew + tx * tw + x]]++;
            if (clipLimit > 0) {
                int clipped = 0;
                for (int i = 0; i < histSize; i++) if (hist[i] > clipLimit) { clipped += hist[i] - clipLimit; hist[i] = clipLimit; }
                int redistBatch = clipped / histSize, residual = clipped - redistBatch * histSize;
                for (int i = 0; i < histSize; i++) hist[i] += redistBatch;
                if (residual != 0) {
                    int step = histSize / residual; if (step < 1) step = 1;
                    for (int i = 0; i < histSize && residual > 0; i += step, residual--) hist[i]++;
                }
            }
            uint8_t* tl = lut + (size_t)(ty * tiles_x + tx) * histSize;
            int sum = 0;
            for (int i = 0; i < histSize; i++) { sum += hist[i]; int v = cv_round_f((float)sum * lutScale); tl[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
        }
    const float inv_tw = 1.0f / tw, inv_th = 1.0f / th;
    for (int y = 0; y < h; y++) {
        const float tyf = y * inv_th - 0.5f;
        int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
        const float ya = tyf - ty1, ya1 = 1.0f - ya;
        if (ty1 < 0) ty1 = 0; if (ty2 > tiles_y - 1) ty2 = tiles_y - 1;
        for (int x = 0; x < w; x++) {
            const float txf = x * inv_tw - 0.5f;
            int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
            const float xa = txf - tx1, xa1 = 1.0f - xa;
            if (tx1 < 0) tx1 = 0; if (tx2 > tiles_x - 1) tx2 = tiles_x - 1;
            const int v = src[(size_t)y * stride + x];
            const float l11 = lut[(size_t)(ty1 * tiles_x + tx1) * histSize + v], l12 = lut[(size_t)(ty1 * tiles_x + tx2) * histSize + v];
            const float l21 = lut[(size_t)(ty2 * tiles_x + tx1) * histSize + v], l22 = lut[(size_t)(ty2 * tiles_x + tx2) * histSize + v];
            const float res = (l11 * xa1 + l12 * xa) * ya1 + (l21 * xa1 + l22 * xa) * ya;
            int o = cv_round_f(res);
            dst[(size_t)y * dstride + x] = (uint8_t)(o < 0 ? 0 : o > 255 ? 255 : o);
        }
    }
    free(ext); free(lut);
}

/* ================================================================ DBoW2 tree descent (next row N2)
 * TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup), Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1259,
 * with FORB::distance (FORB.cpp:80-101) = 256-bit Hamming.  The tree is flat: children of node i are
 * child_ids[child_start[i] .. child_start[i+1]) in file order; a node without children is a leaf (word).  Strict '<':
 * the first child wins ties.  node_id = the node on the path at level L - levelsup (0 = root if that level is <= 0, and
 * also if the leaf is reached earlier - the reference leaves *nid unset there). */
void uo_bow_transform(const int32_t* child_start, const int32_t* child_ids, const uint8_t* node_desc, const double* node_weight,
                      const int32_t* node_word, int L, const uint8_t* desc, int n, int levelsup,
                      int32_t* word_id, int32_t* node_id, double* weight)
{
    const int nid_level = L - levelsup;
    for (int f = 0; f < n; f++) {
        const uint8_t* d = desc + (size_t)f * 32;
        int final_id = 0, level = 0, nid = 0;
        do {
            ++level;
            const int b = child_start[final_id], e = child_start[final_id + 1];
            final_id = child_ids[b];
            int best = uo_descriptor_distance(d, node_desc + (size_t)final_id * 32);
            for (int c = b + 1; c < e; c++) {
                const int id = child_ids[c];
                const int dd = uo_descriptor_distance(d, node_desc + (size_t)id * 32);
                if (dd < best) { best = dd; final_id = id; }
            }
            if (level == nid_level) nid = final_id;
        } while (child_start[final_id + 1] > child_start[final_id]);
        word_id[f] = node_word[final_id]; weight[f] = node_weight[final_id]; node_id[f] = nid;
    }
}

/* ================================================================ KLT front end (next row N1)
 * cv::buildOpticalFlowPyramid (FrameKTL.cc:76) + cv::calcOpticalFlowPyrLK (Tracking.cc:1044-1047: win 21x21, 5 levels, 30 it,
 * eps 0.01, USE_INITIAL_FLOW | LK_GET_MIN_EIGENVALS), restated from OpenCV video/lkpyramid.cpp + imgproc/pyramids.cpp.
 * Integer stages (pyrDown, Scharr derivatives, fixed-point window interpolation) are exact; the float accumulations use the
 * scalar (row-major) order, which differs from OpenCV's SIMD lane order in the last bits -> positions agree to ~1e-3 px. */
void uo_pyr_down(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dstride)
{
    const int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    int* rows = (int*)malloc(sizeof(int) * 5 * dw);
    for (int y = 0; y < dh; y++) {
        for (int k = 0; k < 5; k++) {
            const int sy = reflect101(2 * y - 2 + k, sh);
            const uint8_t* S = src + (size_t)sy * sstride;
            int* R = rows + k * dw;
            for (int x = 0; x < dw; x++) {
                const int x0 = reflect101(2 * x - 2, sw), x1 = reflect101(2 * x - 1, sw), x2 = reflect101(2 * x, sw), x3 = reflect101(2 * x + 1, sw), x4 = reflect101(2 * x + 2, sw);
                R[x] = S[x2] * 6 + (S[x1] + S[x3]) * 4 + S[x0] + S[x4];
            }
        }
        for (int x = 0; x < dw; x++) {
            const int v = rows[2 * dw + x] * 6 + (rows[dw + x] + rows[3 * dw + x]) * 4 + rows[x] + rows[4 * dw + x];
            dst[(size_t)y * dstride + x] = (uint8_t)((v + 128) >> 8);
        }
    }
    free(rows);
}

void uo_scharr(const uint8_t* src, int w, int h, int stride, int16_t* dxy /* w*h*2: Ix, Iy interleaved */)
{
    int* t0 = (int*)malloc(sizeof(int) * (w + 2)); int* t1 = (int*)malloc(sizeof(int) * (w + 2));
    for (int y = 0; y < h; y++) {
        const uint8_t* r0 = src + (size_t)(y > 0 ? y - 1 : (h > 1 ? 1 : 0)) * stride;
        const uint8_t* r1 = src + (size_t)y * stride;
        const uint8_t* r2 = src + (size_t)(y < h - 1 ? y + 1 : (h > 1 ? h - 2 : 0)) * stride;
        for (int x = 0; x < w; x++) { t0[x + 1] = (r0[x] + r2[x]) * 3 + r1[x] * 10; t1[x + 1] = r2[x] - r0[x]; }
        const int xl = w > 1 ? 1 : 0, xr = w > 1 ? w - 2 : 0;
        t0[0] = t0[xl + 1]; t0[w + 1] = t0[xr + 1]; t1[0] = t1[xl + 1]; t1[w + 1] = t1[xr + 1];
        for (int x = 0; x < w; x++) {
            dxy[((size_t)y * w + x) * 2] = (int16_t)(t0[x + 2] - t0[x]);
            dxy[((size_t)y * w + x) * 2 + 1] = (int16_t)((t1[x + 2] + t1[x]) * 3 + t1[x + 1] * 10);
        }
    }
    free(t0); free(t1);
}

#define LK_MAXLEV 10
typedef struct { int nlevels; int w[LK_MAXLEV], h[LK_MAXLEV]; uint8_t* img[LK_MAXLEV]; int16_t* der[LK_MAXLEV]; } uo_lkpyr;

void* uo_lk_pyramid_build(const uint8_t* img, int w, int h, int stride, int win, int max_level)
{
    uo_lkpyr* P = (uo_lkpyr*)calloc(1, sizeof(*P));
    if (max_level > LK_MAXLEV - 1) max_level = LK_MAXLEV - 1;
    P->w[0] = w; P->h[0] = h; P->img[0] = (uint8_t*)malloc((size_t)w * h);
    for (int y = 0; y < h; y++) memcpy(P->img[0] + (size_t)y * w, img + (size_t)y * stride, (size_t)w);
    P->nlevels = 1;
    for (int l = 1; l <= max_level; l++) {
        const int dw = (P->w[l - 1] + 1) / 2, dh = (P->h[l - 1] + 1) / 2;
        if (dw <= win || dh <= win) break;                       /* lkpyramid.cpp: stop when the level is not larger than the window */
        P->w[l] = dw; P->h[l] = dh; P->img[l] = (uint8_t*)malloc((size_t)dw * dh);
        uo_pyr_down(P->img[l - 1], P->w[l - 1], P->h[l - 1], P->w[l - 1], P->img[l], dw);
        P->nlevels = l + 1;
    }
    for (int l = 0; l < P->nlevels; l++) { P->der[l] = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)P->w[l] * P->h[l]); uo_scharr(P->img[l], P->w[l], P->h[l], P->w[l], P->der[l]); }
    return P;
}
void uo_lk_pyramid_free(void* P_) { uo_lkpyr* P = (uo_lkpyr*)P_; if (!P) return; for (int l = 0; l < P->nlevels; l++) { free(P->img[l]); free(P->der[l]); } free(P); }
int  uo_lk_pyramid_levels(const void* P) { return ((const uo_lkpyr*)P)->nlevels; }
void uo_lk_pyramid_get(const void* P_, int l, int* w, int* h, uint8_t* img, int16_t* der)
{
    const uo_lkpyr* P = (const uo_lkpyr*)P_;
    *w = P->w[l]; *h = P->h[l];
    if (img) memcpy(img, P->img[l], (size_t)P->w[l] * P->h[l]);
    if (der) memcpy(der, P->der[l], sizeof(int16_t) * 2 * (size_t)P->w[l] * P->h[l]);
}

static inline int lk_img(const uo_lkpyr* P, int l, int x, int y) { return P->img[l][(size_t)reflect101(y, P->h[l]) * P->w[l] + reflect101(x, P->w[l])]; }
static inline int lk_der(const uo_lkpyr* P, int l, int x, int y, int c)
{ return (x < 0 || y < 0 || x >= P->w[l] || y >= P->h[l]) ? 0 : P->der[l][((size_t)y * P->w[l] + x) * 2 + c]; }
#define LK_DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

/* calcOpticalFlowPyrLK over two prebuilt pyramids.  next_pts is in/out (initial flow).  flags: 4 = USE_INITIAL_FLOW, 8 = GET_MIN_EIGENVALS */
void uo_lk_track(const void* P0_, const void* P1_, const float* prev_pts, float* next_pts, int n, int win, int max_level,
                 int max_iter, double eps, int flags, double min_eig_thr, uint8_t* status, float* err)
{
    const uo_lkpyr* P0 = (const uo_lkpyr*)P0_; const uo_lkpyr* P1 = (const uo_lkpyr*)P1_;
    if (max_iter < 0) max_iter = 0; if (max_iter > 100) max_iter = 100;
    if (eps < 0) eps = 0; if (eps > 10) eps = 10;
    eps *= eps;
    int levels = P0->nlevels < P1->nlevels ? P0->nlevels : P1->nlevels;
    if (max_level > levels - 1) max_level = levels - 1;
    const int W_BITS = 14;
    const float FLT_SCALE = 1.f / (1 << 20);
    const float half = (win - 1) * 0.5f;
    short* Iw = (short*)malloc(sizeof(short) * win * win); short* dIw = (short*)malloc(sizeof(short) * 2 * win * win);
    for (int i = 0; i < n; i++) { status[i] = 1; if (err) err[i] = 0; }
    for (int level = max_level; level >= 0; level--) {
        const int cols = P0->w[level], rows = P0->h[level];
        for (int i = 0; i < n; i++) {
            float px = prev_pts[2 * i] * (float)(1. / (1 << level)), py = prev_pts[2 * i + 1] * (float)(1. / (1 << level));
            float nx, ny;
            if (level == max_level) {
                if (flags & 4) { nx = next_pts[2 * i] * (float)(1. / (1 << level)); ny = next_pts[2 * i + 1] * (float)(1. / (1 << level)); }
                else { nx = px; ny = py; }
            } else { nx = next_pts[2 * i] * 2.f; ny = next_pts[2 * i + 1] * 2.f; }
            next_pts[2 * i] = nx; next_pts[2 * i + 1] = ny;
            px -= half; py -= half;
            int ipx = (int)floorf(px), ipy = (int)floorf(py);
            if (ipx < -win || ipx >= cols || ipy < -win || ipy >= rows) { if (level == 0) { status[i] = 0; if (err) err[i] = 0; } continue; }
            float a = px - ipx, b = py - ipy;
            int iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS)), iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
            int iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS)), iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            float iA11 = 0, iA12 = 0, iA22 = 0;
            for (int y = 0; y < win; y++)
                for (int x = 0; x < win; x++) {
                    const int X = ipx + x, Y = ipy + y;
                    const int ival = LK_DESCALE(lk_img(P0, level, X, Y) * iw00 + lk_img(P0, level, X + 1, Y) * iw01 + lk_img(P0, level, X, Y + 1) * iw10 + lk_img(P0, level, X + 1, Y + 1) * iw11, W_BITS - 5);
                    const int ixval = LK_DESCALE(lk_der(P0, level, X, Y, 0) * iw00 + lk_der(P0, level, X + 1, Y, 0) * iw01 + lk_der(P0, level, X, Y + 1, 0) * iw10 + lk_der(P0, level, X + 1, Y + 1, 0) * iw11, W_BITS);
                    const int iyval = LK_DESCALE(lk_der(P0, level, X, Y, 1) * iw00 + lk_der(P0, level, X + 1, Y, 1) * iw01 + lk_der(P0, level, X, Y + 1, 1) * iw10 + lk_der(P0, level, X + 1, Y + 1, 1) * iw11, W_BITS);
                    Iw[y * win + x] = (short)ival; dIw[2 * (y * win + x)] = (short)ixval; dIw[2 * (y * win + x) + 1] = (short)iyval;
                    iA11 += (float)(ixval * ixval); iA12 += (float)(ixval * iyval); iA22 += (float)(iyval * iyval);
                }
            const float A11 = iA11 * FLT_SCALE, A12 = iA12 * FLT_SCALE, A22 = iA22 * FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
            if (err && (flags & 8)) err[i] = minEig;
            if (minEig < min_eig_thr || D < FLT_EPSILON) { if (level == 0) status[i] = 0; continue; }
            D = 1.f / D;
            nx -= half; ny -= half;
            float pdx = 0, pdy = 0;
            for (int j = 0; j < max_iter; j++) {
                const int inx = (int)floorf(nx), iny = (int)floorf(ny);
                if (inx < -win || inx >= P1->w[level] || iny < -win || iny >= P1->h[level]) { if (level == 0) status[i] = 0; break; }
                a = nx - inx; b = ny - iny;
                iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS)); iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
                iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS)); iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                float ib1 = 0, ib2 = 0;
                for (int y = 0; y < win; y++)
                    for (int x = 0; x < win; x++) {
                        const int X = inx + x, Y = iny + y;
                        const int diff = LK_DESCALE(lk_img(P1, level, X, Y) * iw00 + lk_img(P1, level, X + 1, Y) * iw01 + lk_img(P1, level, X, Y + 1) * iw10 + lk_img(P1, level, X + 1, Y + 1) * iw11, W_BITS - 5) - Iw[y * win + x];
                        ib1 += (float)(diff * dIw[2 * (y * win + x)]); ib2 += (float)(diff * dIw[2 * (y * win + x) + 1]);
                    }
                const float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
                const float dx = (float)((A12 * b2 - A22 * b1) * D), dy = (float)((A12 * b1 - A11 * b2) * D);
                nx += dx; ny += dy;
                next_pts[2 * i] = nx + half; next_pts[2 * i + 1] = ny + half;
                if ((double)dx * dx + (double)dy * dy <= eps) break;
                if (j > 0 && fabsf(dx + pdx) < 0.01 && fabsf(dy + pdy) < 0.01) { next_pts[2 * i] -= dx * 0.5f; next_pts[2 * i + 1] -= dy * 0.5f; break; }
                pdx = dx; pdy = dy;
            }
        }
    }
    free(Iw); free(dIw);
}

/* MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:236-264: N x N distances, per row the sorted row's element
 * (size_t)(0.5*(N-1)), the row with the least median wins (strict <, so the first).  scratch-free, O(N^2 log N). */
static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }
void uo_distinctive_descriptors(const uint8_t* desc, const int32_t* start, int npoints, int32_t* best_idx, int32_t* best_median)
{
    for (int p = 0; p < npoints; p++) {
        const int N = start[p + 1] - start[p];
        const uint8_t* D = desc + (size_t)start[p] * 32;
        if (N <= 0) { best_idx[p] = -1; if (best_median) best_median[p] = -1; continue; }
        int* row = (int*)malloc(sizeof(int) * (size_t)N);
        int BestMedian = 0x7fffffff, BestIdx = 0;
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++) row[j] = (i == j) ? 0 : uo_descriptor_distance(D + (size_t)i * 32, D + (size_t)j * 32);
            qsort(row, (size_t)N, sizeof(int), cmp_int);
            const int median = row[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        free(row);
        best_idx[p] = BestIdx; if (best_median) best_median[p] = BestMedian;
    }
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * N1, last third: cv::findFundamentalMat(pts0, pts1, FM_RANSAC, 1, 0.999, mask) as called at src/Tracking.cc:1062 (only the
 * inlier MASK is used there).  OpenCV (calib3d/fundam.cpp, ptsetreg.cpp): 7-point minimal solver inside a RANSAC loop, at most
 * 1000 iterations with early stop by confidence, samples from cv::RNG; residual = max of the two squared point-to-epipolar-line
 * distances, computed in double, cast to float and compared with (float)(threshold^2); the mask of the BEST hypothesis is returned.
 * cv::RNG's sample sequence and OpenCV's SVD cannot be reproduced bit for bit, so the restatement pins the ALGORITHM with its own
 * deterministic sampler (counter-based SplitMix64) and evaluates a fixed number of hypotheses (no early stop: a superset of what
 * OpenCV would have tried); tests compare the inlier SET with cv2's on synthetic two-view data (tests/golden/cv2_ransac.npz).
 * The cubic is solved by bisection (only + - * / and sqrt), so the CUDA kernel reproduces this function bit for bit.
 * ------------------------------------------------------------------------------------------------------------------------------ */
static unsigned long long uo_sm64(unsigned long long seed, unsigned long long k)
{
    unsigned long long z = seed + k * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static double uo_det3(const double* F)
{
    return F[0] * (F[4] * F[8] - F[5] * F[7]) - F[1] * (F[3] * F[8] - F[5] * F[6]) + F[2] * (F[3] * F[7] - F[4] * F[6]);
}
static double uo_cubic_eval(const double* c, double x) { return ((c[3] * x + c[2]) * x + c[1]) * x + c[0]; }
/* real root of c(x) inside [lo, hi] where c(lo) and c(hi) differ in sign: 200 bisection steps (deterministic on any IEEE machine) */
static double uo_bisect(const double* c, double lo, double hi)
{
    double flo = uo_cubic_eval(c, lo);
    for (int it = 0; it < 200; it++) {
        const double mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) break;
        const double fm = uo_cubic_eval(c, mid);
        if ((fm < 0) == (flo < 0)) { lo = mid; flo = fm; } else hi = mid;
    }
    return 0.5 * (lo + hi);
}
/* real roots of c[3] x^3 + c[2] x^2 + c[1] x + c[0] (ascending); returns their number (0..3) */
static int uo_cubic_roots(const double* c, double* r)
{
    double m = fabs(c[0]); if (fabs(c[1]) > m) m = fabs(c[1]); if (fabs(c[2]) > m) m = fabs(c[2]);
    if (fabs(c[3]) <= 1e-12 * m || c[3] == 0) {                       /* (near) quadratic */
        if (fabs(c[2]) <= 1e-12 * m || c[2] == 0) { if (c[1] == 0) return 0; r[0] = -c[0] / c[1]; return 1; }
        const double disc = c[1] * c[1] - 4 * c[2] * c[0];
        if (disc < 0) return 0;
        const double s = sqrt(disc);
        double a = (-c[1] - s) / (2 * c[2]), b = (-c[1] + s) / (2 * c[2]);
        if (a > b) { const double t = a; a = b; b = t; }
        r[0] = a; r[1] = b; return 2;
    }
    const double B = 1.0 + m / fabs(c[3]);                           /* Cauchy bound on |root| */
    /* stationary points of the cubic: 3 c3 x^2 + 2 c2 x + c1 = 0 */
    const double disc = c[2] * c[2] - 3 * c[3] * c[1];
    int n = 0;
    if (disc <= 0) {                                                 /* monotone: one real root */
        r[n++] = uo_bisect(c, -B, B);
        return n;
    }
    const double s = sqrt(disc);
    double x1 = (-c[2] - s) / (3 * c[3]), x2 = (-c[2] + s) / (3 * c[3]);
    if (x1 > x2) { const double t = x1; x1 = x2; x2 = t; }
    const double fB0 = uo_cubic_eval(c, -B), f1 = uo_cubic_eval(c, x1), f2 = uo_cubic_eval(c, x2), fB1 = uo_cubic_eval(c, B);
    if ((fB0 < 0) != (f1 < 0)) r[n++] = uo_bisect(c, -B, x1);
    if ((f1 < 0) != (f2 < 0)) r[n++] = uo_bisect(c, x1, x2);
    if ((f2 < 0) != (fB1 < 0)) r[n++] = uo_bisect(c, x2, B);
    return n;
}
/* 7-point solver: up to three 3x3 matrices F (row-major, m1^T F m0 = 0 for the seven pairs); returns their number */
static int uo_seven_point(const double* x0, const double* y0, const double* x1, const double* y1, double* Fs)
{
    double A[7][9];
    for (int i = 0; i < 7; i++) {
        A[i][0] = x1[i] * x0[i]; A[i][1] = x1[i] * y0[i]; A[i][2] = x1[i];
        A[i][3] = y1[i] * x0[i]; A[i][4] = y1[i] * y0[i]; A[i][5] = y1[i];
        A[i][6] = x0[i]; A[i][7] = y0[i]; A[i][8] = 1.0;
    }
    /* Gauss-Jordan with full pivoting: 7 pivot columns, 2 free ones */
    int pivcol[7]; int used[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 7; k++) {
        int br = k, bc = -1; double best = 0;
        for (int i = k; i < 7; i++)
            for (int j = 0; j < 9; j++)
                if (!used[j] && fabs(A[i][j]) > best) { best = fabs(A[i][j]); br = i; bc = j; }
        if (bc < 0 || best < 1e-300) return 0;                           /* degenerate sample */
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
    /* det(l F1 + (1 - l) F2) is a cubic in l: interpolate it through l = 0, 1, -1, 2 */
    double G[9], c[4];
    const double d0 = uo_det3(F2), d1 = uo_det3(F1);
    for (int j = 0; j < 9; j++) G[j] = 2.0 * F2[j] - F1[j];
    const double dm = uo_det3(G);
    for (int j = 0; j < 9; j++) G[j] = 2.0 * F1[j] - F2[j];
    const double d2 = uo_det3(G);
    c[0] = d0;
    c[3] = (d2 - 3.0 * d1 + 3.0 * d0 - dm) / 6.0;
    c[2] = 0.5 * (d1 + dm) - d0;
    c[1] = d1 - d0 - c[2] - c[3];
    double r[3];
    const int nr = uo_cubic_roots(c, r);
    for (int k = 0; k < nr; k++) {
        double nrm = 0;
        for (int j = 0; j < 9; j++) { Fs[9 * k + j] = r[k] * F1[j] + (1.0 - r[k]) * F2[j]; nrm += Fs[9 * k + j] * Fs[9 * k + j]; }
        nrm = sqrt(nrm);
        if (nrm > 0) for (int j = 0; j < 9; j++) Fs[9 * k + j] /= nrm;
    }
    return nr;
}
static int uo_fm_inlier(const double* F, double x0, double y0, double x1, double y1, float t2)
{
    double a = F[0] * x0 + F[1] * y0 + F[2], b = F[3] * x0 + F[4] * y0 + F[5], c = F[6] * x0 + F[7] * y0 + F[8];
    const double s2 = 1.0 / (a * a + b * b), d2 = x1 * a + y1 * b + c;
    a = F[0] * x1 + F[3] * y1 + F[6]; b = F[1] * x1 + F[4] * y1 + F[7]; c = F[2] * x1 + F[5] * y1 + F[8];
    const double s1 = 1.0 / (a * a + b * b), d1 = x0 * a + y0 * b + c;
    const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
    const float err = (float)(e1 > e2 ? e1 : e2);
    return err <= t2;
}
/* pts: n x 2 float each; mask[n] out; F[9] out (best hypothesis, Frobenius-normalised; may be NULL).  Returns the inlier count,
 * 0 when no hypothesis had more than 6 inliers, -1 when n < 15 (OpenCV switches to LMedS below 15 points: not restated). */
int uo_ransac_fundamental(const float* pts0, const float* pts1, int n, double threshold, int nhyp, unsigned long long seed, uint8_t* mask, double* Fout)
{
    if (n < 15) return -1;
    const float t2 = (float)(threshold * threshold);
    int best = 6; double bestF[9]; int have = 0;
    for (int h = 0; h < nhyp; h++) {
        int idx[7], ns = 0;
        for (int k = 1; k <= 16 && ns < 7; k++) {
            const int cand = (int)(uo_sm64(seed, (unsigned long long)h * 16ULL + (unsigned long long)k) % (unsigned long long)n);
            int dup = 0;
            for (int j = 0; j < ns; j++) dup |= idx[j] == cand;
            if (!dup) idx[ns++] = cand;
        }
        if (ns < 7) continue;
        double x0[7], y0[7], x1[7], y1[7], Fs[27];
        for (int j = 0; j < 7; j++) { x0[j] = pts0[2 * idx[j]]; y0[j] = pts0[2 * idx[j] + 1]; x1[j] = pts1[2 * idx[j]]; y1[j] = pts1[2 * idx[j] + 1]; }
        const int nm = uo_seven_point(x0, y0, x1, y1, Fs);
        for (int k = 0; k < nm; k++) {
            int cnt = 0;
            for (int i = 0; i < n; i++) cnt += uo_fm_inlier(Fs + 9 * k, pts0[2 * i], pts0[2 * i + 1], pts1[2 * i], pts1[2 * i + 1], t2);
            if (cnt > best) { best = cnt; memcpy(bestF, Fs + 9 * k, sizeof(bestF)); have = 1; }     /* strict: the first best hypothesis wins */
        }
    }
    if (!have) { for (int i = 0; i < n; i++) mask[i] = 0; return 0; }
    for (int i = 0; i < n; i++) mask[i] = (uint8_t)uo_fm_inlier(bestF, pts0[2 * i], pts0[2 * i + 1], pts1[2 * i], pts1[2 * i + 1], t2);
    if (Fout) memcpy(Fout, bestF, sizeof(bestF));
    return best;
}
