/*
 * oracle/sgbm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the semi-global block matching disparity path that the
 * reference's MatcherOpenCVSGBM plugin delegates to cv::StereoSGBM::compute
 * (reference call site: src/stereoMatcher/matcherOpenCVSGBM.cpp:21, setters :53-110,
 * driven by src/generate_disparity.cpp:245-256).  The arithmetic lives in a third-party
 * dependency that is NOT in /root/reference (OpenCV calib3d; the reference pins it only as
 * `find_package(OpenCV 3)` else `4`, CMakeLists.txt:48-51).  This file restates the
 * published algorithm as specified in SURVEY.md Appendix A (A.1 - A.8) and is pinned
 * bit-for-bit against OpenCV 4.13.0 (python `cv2`, the same library through its own
 * binding) by tests/test_oracle.py and by the golden vectors under tests/golden/
 * (generator: tests/golden/make_golden.py).
 *
 * Also restated here:
 *   - AbstractStereoMatcher::match()/MatcherOpenCVSGBM::forwardMatch() int16 -> float32
 *     conversion (src/stereoMatcher/abstractStereoMatcher.cpp:44-53, matcherOpenCVSGBM.cpp:34)
 *   - processDisparity() /16 + depth-window thresholding (src/generate_disparity.cpp:436-452)
 *   - calc_q + dispInfoMsg2depthMsg reprojection (src/disparity_to_depth.cpp:62-85,136-205)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (libb200sgm.so) never links or calls it.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct {
    int minDisparity;
    int numDisparities;
    int blockSize;
    int P1;
    int P2;
    int disp12MaxDiff;
    int preFilterCap;
    int uniquenessRatio;
    int speckleWindowSize;
    int speckleRange;
    int mode; /* 0 = MODE_SGBM (5 paths), 1 = MODE_HH (8 paths) */
} sgbm_oracle_params;

/* Optional stage dumps (any pointer may be NULL).  Volumes are [H][W1][D] int16. */
typedef struct {
    int16_t *C;        /* block-summed matching cost (A.4)                      */
    int16_t *S;        /* aggregated cost seen by WTA (A.5)                     */
    int16_t *disp_wta; /* H*W, after WTA + LR check, before median (A.6, A.7)   */
    int16_t *disp_med; /* H*W, after 3x3 median, before speckle filter (A.8)    */
} sgbm_oracle_dumps;

#define MAX_COST 32767

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---- A.2 prefilter for one image row -------------------------------------------------- */
typedef struct {
    uint8_t *sob, *sob_lo, *sob_hi, *raw, *raw_lo, *raw_hi; /* each W */
} rowfeat;

static void rowfeat_alloc(rowfeat *f, int W)
{
    uint8_t *p = (uint8_t *)malloc((size_t)6 * W);
    f->sob = p; f->sob_lo = p + W; f->sob_hi = p + 2 * W;
    f->raw = p + 3 * W; f->raw_lo = p + 4 * W; f->raw_hi = p + 5 * W;
}
static void rowfeat_free(rowfeat *f) { free(f->sob); }

static void half_pixel_interval(const uint8_t *a, uint8_t *lo, uint8_t *hi, int W)
{
    for (int x = 0; x < W; x++) {
        int v = a[x];
        int vl = x > 0 ? (v + a[x - 1]) / 2 : v;
        int vr = x < W - 1 ? (v + a[x + 1]) / 2 : v;
        lo[x] = (uint8_t)imin(v, imin(vl, vr));
        hi[x] = (uint8_t)imax(v, imax(vl, vr));
    }
}

static void prefilter_row(const uint8_t *img, int W, int H, int y, int ftzero, rowfeat *f)
{
    const uint8_t *r = img + (size_t)y * W;
    const uint8_t *rn = img + (size_t)(y > 0 ? y - 1 : y) * W;
    const uint8_t *rs = img + (size_t)(y < H - 1 ? y + 1 : y) * W;
    f->sob[0] = f->sob[W - 1] = (uint8_t)ftzero;
    f->raw[0] = f->raw[W - 1] = (uint8_t)ftzero; /* quirk: raw border forced to ftzero */
    for (int x = 1; x < W - 1; x++) {
        int g = (r[x + 1] - r[x - 1]) * 2 + (rn[x + 1] - rn[x - 1]) + (rs[x + 1] - rs[x - 1]);
        f->sob[x] = (uint8_t)(iclamp(g, -ftzero, ftzero) + ftzero);
        f->raw[x] = r[x];
    }
    half_pixel_interval(f->sob, f->sob_lo, f->sob_hi, W);
    half_pixel_interval(f->raw, f->raw_lo, f->raw_hi, W);
}

/* ---- A.3 Birchfield-Tomasi pixel cost of one channel ----------------------------------- */
static inline int bt_cost(int u, int ulo, int uhi, int v, int vlo, int vhi)
{
    int c0 = imax(0, imax(u - vhi, vlo - u));
    int c1 = imax(0, imax(v - uhi, ulo - v));
    return imin(c0, c1);
}

typedef struct {
    int W, H, D, minD, minX1, W1, SW2, SH2, ftzero;
    const uint8_t *L, *R;
    rowfeat fl, fr;
    int16_t *pd;       /* W1*D scratch, pixDiff of one row          */
    int16_t **hs;      /* ring of hsum rows, each W1*D              */
    int *hs_tag;
    int hs_n;
} costctx;

/* A.3 + horizontal part of A.4 for image row y -> hsum row (ring cached). */
static const int16_t *get_hsum(costctx *c, int y)
{
    int slot = y % c->hs_n;
    if (c->hs_tag[slot] == y) return c->hs[slot];
    const int W1 = c->W1, D = c->D;
    prefilter_row(c->L, c->W, c->H, y, c->ftzero, &c->fl);
    prefilter_row(c->R, c->W, c->H, y, c->ftzero, &c->fr);
    for (int x1 = 0; x1 < W1; x1++) {
        int x = x1 + c->minX1;
        for (int k = 0; k < D; k++) {
            int xr = x - (k + c->minD);
            int cs = bt_cost(c->fl.sob[x], c->fl.sob_lo[x], c->fl.sob_hi[x],
                             c->fr.sob[xr], c->fr.sob_lo[xr], c->fr.sob_hi[xr]);
            int cr = bt_cost(c->fl.raw[x], c->fl.raw_lo[x], c->fl.raw_hi[x],
                             c->fr.raw[xr], c->fr.raw_lo[xr], c->fr.raw_hi[xr]);
            c->pd[(size_t)x1 * D + k] = (int16_t)(cs + (cr >> 2));
        }
    }
    int16_t *out = c->hs[slot];
    for (int x1 = 0; x1 < W1; x1++)
        for (int k = 0; k < D; k++) {
            int s = 0;
            for (int j = -c->SW2; j <= c->SW2; j++)
                s += c->pd[(size_t)iclamp(x1 + j, 0, W1 - 1) * D + k];
            out[(size_t)x1 * D + k] = (int16_t)s;
        }
    c->hs_tag[slot] = y;
    return out;
}

/* A.4: C[y] = sum of hsum over the clamped vertical window, stored as int16 (wraps). */
static void cost_row(costctx *c, int y, int16_t *Crow)
{
    const size_t n = (size_t)c->W1 * c->D;
    int32_t *acc = (int32_t *)calloc(n, sizeof(int32_t));
    for (int i = -c->SH2; i <= c->SH2; i++) {
        const int16_t *h = get_hsum(c, iclamp(y + i, 0, c->H - 1));
        for (size_t t = 0; t < n; t++) acc[t] += h[t];
    }
    for (size_t t = 0; t < n; t++) Crow[t] = (int16_t)acc[t];
    free(acc);
}

/* ---- A.5 one min-plus step ------------------------------------------------------------ */
/* Lp may be NULL (= predecessor outside the domain: all-zero vector, m = 0). Returns min(L). */
static inline int path_step(const int16_t *C, const int16_t *Lp, int mp, int D, int P1, int P2, int16_t *L)
{
    int mn = MAX_COST;
    if (!Lp) {
        for (int k = 0; k < D; k++) { L[k] = C[k]; if (L[k] < mn) mn = L[k]; }
        return mn;
    }
    for (int k = 0; k < D; k++) {
        int a = Lp[k];
        int b = (k > 0 ? Lp[k - 1] : MAX_COST) + P1;
        int c = (k < D - 1 ? Lp[k + 1] : MAX_COST) + P1;
        int d = mp + P2;
        int v = C[k] + imin(imin(a, b), imin(c, d)) - mp;
        L[k] = (int16_t)v;
        if (L[k] < mn) mn = L[k];
    }
    return mn;
}

static inline int16_t sat_add16(int a, int b)
{
    int s = a + b;
    return (int16_t)(s > 32767 ? 32767 : (s < -32768 ? -32768 : s));
}

/* ---- A.6 + A.7 for one row, given the final S row ---------------------------------------- */
typedef struct {
    int W, D, minD, minX1, W1, uniq, d12, INVALID;
    int16_t *disp2; int *disp2cost;
} wtactx;

static void wta_row(wtactx *w, const int16_t *Srow, int16_t *drow)
{
    const int W = w->W, D = w->D, minD = w->minD, W1 = w->W1, minX1 = w->minX1;
    for (int x = 0; x < W; x++) { drow[x] = (int16_t)w->INVALID; w->disp2[x] = (int16_t)w->INVALID; w->disp2cost[x] = MAX_COST; }
    for (int x1 = W1 - 1; x1 >= 0; x1--) {
        const int16_t *S = Srow + (size_t)x1 * D;
        int minS = MAX_COST, best = -1;
        for (int k = 0; k < D; k++) if (S[k] < minS) { minS = S[k]; best = k; }
        int k;
        for (k = 0; k < D; k++)
            if (S[k] * (100 - w->uniq) < minS * 100 && abs(best - k) > 1) break;
        if (k < D) continue;
        if (best < 0) continue; /* all saturated: result equals INVALID, disp2 untouched */
        int x2 = x1 + minX1 - best - minD;
        if (x2 >= 0 && x2 < W && w->disp2cost[x2] > minS) { w->disp2cost[x2] = minS; w->disp2[x2] = (int16_t)(best + minD); }
        int dfix;
        if (0 < best && best < D - 1) {
            int den = imax(S[best - 1] + S[best + 1] - 2 * S[best], 1);
            dfix = best * 16 + ((S[best - 1] - S[best + 1]) * 16 + den) / (den * 2); /* C trunc */
        } else dfix = best * 16;
        drow[x1 + minX1] = (int16_t)(dfix + minD * 16);
    }
    for (int x = minX1; x < minX1 + W1; x++) {
        int d1 = drow[x];
        if (d1 == w->INVALID) continue;
        int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
        int _x = x - _d, x_ = x - d_;
        if (0 <= _x && _x < W && w->disp2[_x] >= minD && abs(w->disp2[_x] - _d) > w->d12 &&
            0 <= x_ && x_ < W && w->disp2[x_] >= minD && abs(w->disp2[x_] - d_) > w->d12)
            drow[x] = (int16_t)w->INVALID;
    }
}

/* ---- A.8 post filters -------------------------------------------------------------------- */
static void sort3(int16_t *a, int16_t *b, int16_t *c)
{
    int16_t t;
    if (*a > *b) { t = *a; *a = *b; *b = t; }
    if (*b > *c) { t = *b; *b = *c; *c = t; }
    if (*a > *b) { t = *a; *a = *b; *b = t; }
}

void sgbm_oracle_median3x3(const int16_t *src, int16_t *dst, int W, int H)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int16_t v[9]; int n = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++)
                    v[n++] = src[(size_t)iclamp(y + dy, 0, H - 1) * W + iclamp(x + dx, 0, W - 1)];
            /* median of 9 by insertion sort */
            for (int i = 1; i < 9; i++) { int16_t t = v[i]; int j = i - 1; while (j >= 0 && v[j] > t) { v[j + 1] = v[j]; j--; } v[j + 1] = t; }
            dst[(size_t)y * W + x] = v[4];
        }
    (void)sort3;
}

/* 4-connected components, |a-b| <= maxDiff, pixels == newVal excluded; size <= maxSize -> newVal. */
void sgbm_oracle_filter_speckles(int16_t *img, int W, int H, int newVal, int maxSize, int maxDiff)
{
    size_t n = (size_t)W * H;
    int32_t *label = (int32_t *)calloc(n, sizeof(int32_t));
    int32_t *stack = (int32_t *)malloc(n * sizeof(int32_t));
    uint8_t *small = (uint8_t *)malloc(n + 1);
    int cur = 0;
    for (size_t p0 = 0; p0 < n; p0++) {
        if (img[p0] == newVal) continue;
        if (label[p0]) { if (small[label[p0]]) img[p0] = (int16_t)newVal; continue; }
        cur++;
        size_t sp = 0, count = 0;
        stack[sp++] = (int32_t)p0; label[p0] = cur;
        while (sp) {
            int32_t p = stack[--sp]; count++;
            int x = p % W, y = p / W; int v = img[p];
            const int nx[4] = { x - 1, x + 1, x, x }, ny[4] = { y, y, y - 1, y + 1 };
            for (int t = 0; t < 4; t++) {
                if (nx[t] < 0 || nx[t] >= W || ny[t] < 0 || ny[t] >= H) continue;
                size_t q = (size_t)ny[t] * W + nx[t];
                if (label[q] || img[q] == newVal || abs(img[q] - v) > maxDiff) continue;
                label[q] = cur; stack[sp++] = (int32_t)q;
            }
        }
        small[cur] = count <= (size_t)maxSize;
        if (small[cur]) img[p0] = (int16_t)newVal;
    }
    free(label); free(stack); free(small);
}

/* ---- full pipeline ----------------------------------------------------------------------- */
int sgbm_oracle_compute(const uint8_t *L, const uint8_t *R, int W, int H,
                        const sgbm_oracle_params *prm, int16_t *disp, const sgbm_oracle_dumps *dumps)
{
    /* A.1 parameter normalisation */
    const int minD = prm->minDisparity, D = prm->numDisparities;
    if (D <= 0 || W <= 0 || H <= 0) return -1;
    const int maxD = minD + D;
    const int uniq = prm->uniquenessRatio >= 0 ? prm->uniquenessRatio : 10;
    const int d12 = prm->disp12MaxDiff > 0 ? prm->disp12MaxDiff : 1;
    const int P1 = prm->P1 > 0 ? prm->P1 : 2;
    const int P2 = imax(prm->P2 > 0 ? prm->P2 : 5, P1 + 1);
    const int SW2 = (prm->blockSize > 0 ? prm->blockSize : 5) / 2, SH2 = SW2;
    const int ftzero = imax(prm->preFilterCap, 15) | 1;
    const int INVALID = (minD - 1) * 16;
    const int minX1 = imax(maxD, 0), maxX1 = W + imin(minD, 0), W1 = maxX1 - minX1;
    const int hh = prm->mode == 1;
    const size_t npix = (size_t)W * H;

    for (size_t i = 0; i < npix; i++) disp[i] = (int16_t)INVALID;
    int16_t *wta_out = (int16_t *)malloc(npix * sizeof(int16_t));
    for (size_t i = 0; i < npix; i++) wta_out[i] = (int16_t)INVALID;

    if (W1 > 0) {
        const size_t rowsz = (size_t)W1 * D;
        costctx cc; memset(&cc, 0, sizeof cc);
        cc.W = W; cc.H = H; cc.D = D; cc.minD = minD; cc.minX1 = minX1; cc.W1 = W1; cc.SW2 = SW2; cc.SH2 = SH2;
        cc.ftzero = ftzero; cc.L = L; cc.R = R;
        rowfeat_alloc(&cc.fl, W); rowfeat_alloc(&cc.fr, W);
        cc.pd = (int16_t *)malloc(rowsz * sizeof(int16_t));
        cc.hs_n = 2 * SH2 + 2;
        cc.hs = (int16_t **)malloc(cc.hs_n * sizeof(int16_t *));
        cc.hs_tag = (int *)malloc(cc.hs_n * sizeof(int));
        for (int i = 0; i < cc.hs_n; i++) { cc.hs[i] = (int16_t *)malloc(rowsz * sizeof(int16_t)); cc.hs_tag[i] = -1; }

        wtactx wc; wc.W = W; wc.D = D; wc.minD = minD; wc.minX1 = minX1; wc.W1 = W1; wc.uniq = uniq; wc.d12 = d12; wc.INVALID = INVALID;
        wc.disp2 = (int16_t *)malloc(W * sizeof(int16_t)); wc.disp2cost = (int *)malloc(W * sizeof(int));

        /* Lr rows: [2 rows][4 dirs][(W1+2) columns][D]; column index x1+1; NULL-like borders handled by flags. */
        const size_t lrrow = (size_t)(W1 + 2) * D;
        int16_t *Lr = (int16_t *)calloc(2 * 4 * lrrow, sizeof(int16_t));
        int *mLr = (int *)calloc((size_t)2 * 4 * (W1 + 2), sizeof(int));
        int16_t *Crow = (int16_t *)malloc(rowsz * sizeof(int16_t));
        int16_t *Srow = (int16_t *)malloc(rowsz * sizeof(int16_t));
        int16_t *Lh = (int16_t *)malloc(2 * (size_t)D * sizeof(int16_t));
        int16_t *Cfull = NULL, *Sfull = NULL;
        if (hh) { Cfull = (int16_t *)malloc(rowsz * H * sizeof(int16_t)); Sfull = (int16_t *)malloc(rowsz * H * sizeof(int16_t)); }

#define LR(row, dir, x1) (Lr + (((size_t)(row) * 4 + (dir)) * lrrow) + (size_t)((x1) + 1) * D)
#define MLR(row, dir, x1) (mLr[((size_t)(row) * 4 + (dir)) * (W1 + 2) + (x1) + 1])

        for (int pass = 0; pass < (hh ? 2 : 1); pass++) {
            const int y0 = pass ? H - 1 : 0, y1 = pass ? -1 : H, dy = pass ? -1 : 1;
            const int x0 = pass ? W1 - 1 : 0, xe = pass ? -1 : W1, dx = pass ? -1 : 1;
            for (int y = y0; y != y1; y += dy) {
                const int cur = (y - y0) * dy & 1, prv = cur ^ 1;
                const int first_row = (y == y0);
                int16_t *Cr = hh ? Cfull + rowsz * y : Crow;
                int16_t *Sr = hh ? Sfull + rowsz * y : Srow;
                if (pass == 0) cost_row(&cc, y, Cr);
                if (dumps && dumps->C && pass == 0) memcpy(dumps->C + rowsz * y, Cr, rowsz * sizeof(int16_t));
                /* four paths whose predecessors are (x-dx,y), (x-dx,y-dy), (x,y-dy), (x+dx,y-dy) */
                for (int x1 = x0; x1 != xe; x1 += dx) {
                    const int16_t *Cp = Cr + (size_t)x1 * D;
                    int xa = x1 - dx, xb = x1 + dx;
                    int in_a = (xa >= 0 && xa < W1), in_b = (xb >= 0 && xb < W1);
                    const int16_t *p0 = in_a ? LR(cur, 0, xa) : NULL;
                    const int16_t *p1 = (!first_row && in_a) ? LR(prv, 1, xa) : NULL;
                    const int16_t *p2 = (!first_row) ? LR(prv, 2, x1) : NULL;
                    const int16_t *p3 = (!first_row && in_b) ? LR(prv, 3, xb) : NULL;
                    MLR(cur, 0, x1) = path_step(Cp, p0, p0 ? MLR(cur, 0, xa) : 0, D, P1, P2, LR(cur, 0, x1));
                    MLR(cur, 1, x1) = path_step(Cp, p1, p1 ? MLR(prv, 1, xa) : 0, D, P1, P2, LR(cur, 1, x1));
                    MLR(cur, 2, x1) = path_step(Cp, p2, p2 ? MLR(prv, 2, x1) : 0, D, P1, P2, LR(cur, 2, x1));
                    MLR(cur, 3, x1) = path_step(Cp, p3, p3 ? MLR(prv, 3, xb) : 0, D, P1, P2, LR(cur, 3, x1));
                    int16_t *Sp = Sr + (size_t)x1 * D;
                    for (int k = 0; k < D; k++) {
                        int16_t s = sat_add16(LR(cur, 0, x1)[k], LR(cur, 1, x1)[k]);
                        s = sat_add16(s, LR(cur, 2, x1)[k]);
                        s = sat_add16(s, LR(cur, 3, x1)[k]);
                        Sp[k] = pass ? sat_add16(Sp[k], s) : s;
                    }
                }
                if (!hh) {
                    /* MODE_SGBM: 5th path, right-to-left within the row, before WTA */
                    int mprev = 0; int have = 0;
                    for (int x1 = W1 - 1; x1 >= 0; x1--) {
                        int16_t *Lc = Lh + (size_t)(x1 & 1) * D, *Lpv = Lh + (size_t)((x1 & 1) ^ 1) * D;
                        mprev = path_step(Cr + (size_t)x1 * D, have ? Lpv : NULL, mprev, D, P1, P2, Lc);
                        have = 1;
                        int16_t *Sp = Sr + (size_t)x1 * D;
                        for (int k = 0; k < D; k++) Sp[k] = sat_add16(Sp[k], Lc[k]);
                    }
                }
                if (!hh || pass == 1) {
                    if (dumps && dumps->S) memcpy(dumps->S + rowsz * y, Sr, rowsz * sizeof(int16_t));
                    wta_row(&wc, Sr, wta_out + (size_t)y * W);
                }
            }
        }
        free(Lr); free(mLr); free(Crow); free(Srow); free(Lh); free(Cfull); free(Sfull);
        free(wc.disp2); free(wc.disp2cost);
        for (int i = 0; i < cc.hs_n; i++) free(cc.hs[i]);
        free(cc.hs); free(cc.hs_tag); free(cc.pd); rowfeat_free(&cc.fl); rowfeat_free(&cc.fr);
    }
    if (dumps && dumps->disp_wta) memcpy(dumps->disp_wta, wta_out, npix * sizeof(int16_t));
    sgbm_oracle_median3x3(wta_out, disp, W, H);
    if (dumps && dumps->disp_med) memcpy(dumps->disp_med, disp, npix * sizeof(int16_t));
    if (prm->speckleWindowSize > 0)
        sgbm_oracle_filter_speckles(disp, W, H, INVALID, prm->speckleWindowSize, 16 * prm->speckleRange);
    free(wta_out);
    return 0;
}

/* ---- a10: int16 -> float32 (value unchanged, still x16); matcherOpenCVSGBM.cpp:34 ----------- */
void sgbm_oracle_to_float(const int16_t *disp16, float *out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = (float)disp16[i];
}

/* ---- a11: processDisparity thresholding; generate_disparity.cpp:436-452 --------------------- */
/* dmat = disp16 * (1/16) as float; dmat < min_disp -> 10000; dmat > max_disp -> 10000.          */
void sgbm_oracle_process_disparity(const int16_t *disp16, float *dmat, size_t n, float min_disp, float max_disp)
{
    for (size_t i = 0; i < n; i++) {
        float d = (float)((double)disp16[i] * (1.0 / 16.0)); /* convertTo(CV_32F, inv_dpp): exact */
        if (d < min_disp) d = 10000.0f;
        if (d > max_disp) d = 10000.0f;
        dmat[i] = d;
    }
}

/* ---- R: disparity_to_depth.cpp:136-205.  q = {q03, q13, wz, q32, q33} as float32.             */
/* color: MONO8 (channels 1) or BGR8 (channels 3), tight rows; depth window as the doubles the  */
/* reference compares the float Z with (:175).  depth: H*W float (0 where rejected); xyz: up to */
/* H*W records of 4 floats {X,Y,Z,rgb-packed}; returns number of points (row-major scan order). */
typedef struct { float x, y, z; uint32_t rgb; } sgbm_oracle_point;

uint32_t sgbm_oracle_reproject(const float *dmat, const uint8_t *color, int channels, int W, int H, const float q[5],
                               double depth_min, double depth_max, float *depth, sgbm_oracle_point *pts)
{
    const float q03 = q[0], q13 = q[1], wz = q[2], q32 = q[3], q33 = q[4];
    uint32_t n = 0;
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            size_t p = (size_t)i * W + j;
            float d = dmat[p];
            if (depth) depth[p] = 0.0f;
            if (d != 0 && d != 10000) {
                volatile float w = d * q32;   /* volatile: forbid fused multiply-add */
                w = w + q33;
                float X = ((float)j + q03) / w, Y = ((float)i + q13) / w, Z = wz / w;
                if (w > 0 && Z > 0 && (double)Z <= depth_max && (double)Z >= depth_min) {
                    if (depth) depth[p] = Z;
                    if (pts) {
                        uint32_t b = 0, g = 0, r = 0;
                        if (color && channels == 1) b = g = r = color[p];
                        else if (color && channels == 3) { b = color[3 * p]; g = color[3 * p + 1]; r = color[3 * p + 2]; }
                        pts[n].x = X; pts[n].y = Y; pts[n].z = Z; pts[n].rgb = (r << 16) | (g << 8) | b;
                    }
                    n++;
                }
            }
        }
    return n;
}
