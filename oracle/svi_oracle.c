/* svi_oracle.c -- plain-C restatement of the reference's CPU algorithm for the stereo front-end
 * hot path.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs load this library; libsvi_gpu never does.
 *
 * It follows the reference call-for-call (paths relative to the svi_mapper tree):
 *   CFundamentalMatcher::addNewLandmarks                  src/core/CFundamentalMatcher.cpp:83-193
 *     cv::GFTTDetector(1000,0.01,7,7,true)::detect        :18,:101   -> svo_gftt (cornerHarris, minMaxLoc,
 *                                                                        threshold, dilate, sort, bucket grid)
 *     BriefDescriptorExtractor::compute(imgL, kps)        :106       -> svo_brief32 (integral image of the
 *                                                                        whole input, border filter, 256 tests)
 *     per key-point getPointTriangulatedInRIGHTFull       src/core/CTriangulator.cpp:51-119
 *       one BRIEF::compute on the 57-row ROI (its own integral image), one BFMatcher::match,
 *       cut-off, getPointInLEFT :326-356
 *   CTriangulator::getPointTriangulatedInLEFT (7 args)    src/core/CTriangulator.cpp:255-324
 * OpenCV arithmetic (not in the reference tree, un-versioned "trunk") is restated from its
 * published algorithms in the operation order validated against cv2 4.13 with optimisations off
 * (SURVEY.md Appendix A); tests/test_oracle.py pins this file against cv2 and against the numpy
 * oracle.  BRIEF: PARITY UNPINNED at pair-table level (opencv_contrib generated_32.i is absent);
 * the table is svi_mapper_b200/csrc/brief_pattern_32.h, shared with the CUDA kernels.
 *
 * Per-frame work allocates and frees its buffers like the OpenCV calls it stands for, and one
 * frame is processed by one thread (cv::setNumThreads(1), src/core/CTrackerGT.cpp:48-49);
 * svo_stereo_frames_mt runs independent frames on several threads for the all-cores baseline. */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../svi_mapper_b200/csrc/brief_pattern_32.h"

static const signed char kPat[SVI_BRIEF_NTESTS][4] = SVI_BRIEF_PATTERN_INIT;

enum { ST_OK = 0, ST_TRI_RANGE = 1, ST_TRI_NO_DESC = 2, ST_TRI_NO_MATCH = 3, ST_TRI_DISTANCE = 4,
       ST_TRI_ZERO_DISP = 5, ST_TRI_BAD_ROI = 6 };

typedef struct svo_config {
    int32_t width, height;
    double P_left[12], P_right[12];
    double quality_level, min_distance, harris_k, min_disparity;
    int32_t max_corners;
    float keypoint_size, search_range, match_cutoff;
} svo_config;

static inline int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

/* ------------------------------------------------------------------ cv::cornerHarris(img, 7, 3, k) */
int svo_harris_response(const uint8_t* img, int w, int h, int pitch, double k, float* out) {
    const double scale = 1.0 / (4.0 * 7.0 * 255.0);
    const float f1 = (float)scale, f0 = (float)(2.0 * scale), kf = (float)k;
    const size_t n = (size_t)w * h;
    float* r = (float*)malloc(sizeof(float) * (size_t)w * (h + 2));  /* row-filtered planes, rows -1..h */
    float* q = (float*)malloc(sizeof(float) * (size_t)w * (h + 2));
    float* cov = (float*)malloc(sizeof(float) * 3 * n);
    double* hs = (double*)malloc(sizeof(double) * 3 * n);
    double* col = (double*)malloc(sizeof(double) * 3 * w);
    if (!r || !q || !cov || !hs || !col) { free(r); free(q); free(cov); free(hs); free(col); return -1; }
    /* Sobel row pass on every source row incl. the two REFLECT_101 border rows */
    for (int yy = -1; yy <= h; ++yy) {
        const uint8_t* p = img + (size_t)reflect101(yy, h) * pitch;
        float* rr = r + (size_t)(yy + 1) * w;
        float* qq = q + (size_t)(yy + 1) * w;
        for (int x = 0; x < w; ++x) {
            const float a = (float)p[reflect101(x - 1, w)], b = (float)p[x], c = (float)p[reflect101(x + 1, w)];
            rr[x] = c - a;
            float t0 = a * f1, t1 = b * f0, t2 = c * f1;
            float s = t0 + t1;
            qq[x] = s + t2;
        }
    }
    /* column pass + products */
    for (int y = 0; y < h; ++y) {
        const float *r0 = r + (size_t)y * w, *r1 = r0 + w, *r2 = r1 + w;
        const float *q0 = q + (size_t)y * w, *q2 = q0 + 2 * (size_t)w;
        float* c = cov + (size_t)y * w * 3;
        for (int x = 0; x < w; ++x) {
            float s = r0[x] + r2[x];
            float m0 = s * f1, m1 = r1[x] * f0;
            float dx = m0 + m1;
            float dy = q2[x] - q0[x];
            c[3 * x] = dx * dx;
            c[3 * x + 1] = dx * dy;
            c[3 * x + 2] = dy * dy;
        }
    }
    /* boxFilter 7x7, normalize=false: RowSum<float,double> then ColumnSum<double,float> */
    for (int y = 0; y < h; ++y) {
        const float* c = cov + (size_t)y * w * 3;
        double* d = hs + (size_t)y * w * 3;
        for (int ch = 0; ch < 3; ++ch) {
            double s = 0;
            for (int i = -3; i <= 3; ++i) s += (double)c[3 * reflect101(i, w) + ch];
            d[ch] = s;
            for (int x = 0; x < w - 1; ++x) {
                s += (double)c[3 * reflect101(x + 4, w) + ch] - (double)c[3 * reflect101(x - 3, w) + ch];
                d[3 * (x + 1) + ch] = s;
            }
        }
    }
    for (int i = 0; i < 3 * w; ++i) col[i] = 0;
    for (int yy = -3; yy < 3; ++yy) {
        const double* d = hs + (size_t)reflect101(yy, h) * w * 3;
        for (int i = 0; i < 3 * w; ++i) col[i] += d[i];
    }
    for (int y = 0; y < h; ++y) {
        const double* dp = hs + (size_t)reflect101(y + 3, h) * w * 3;
        const double* dm = hs + (size_t)reflect101(y - 3, h) * w * 3;
        float* o = out + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            double s0 = col[3 * x] + dp[3 * x], s1 = col[3 * x + 1] + dp[3 * x + 1], s2 = col[3 * x + 2] + dp[3 * x + 2];
            float a = (float)s0, b = (float)s1, c = (float)s2;
            col[3 * x] = s0 - dm[3 * x];
            col[3 * x + 1] = s1 - dm[3 * x + 1];
            col[3 * x + 2] = s2 - dm[3 * x + 2];
            float ac = a * c, bb = b * b;
            float det = ac - bb;
            float tr = a + c;
            float kt = kf * tr;
            float ktt = kt * tr;
            o[x] = det - ktt;
        }
    }
    free(r); free(q); free(cov); free(hs); free(col);
    return 0;
}

/* ------------------------------------------------------------------ cv::goodFeaturesToTrack (Harris) */
typedef struct { float v; int32_t addr; } cand_t;
static int cand_cmp(const void* pa, const void* pb) {
    const cand_t *a = (const cand_t*)pa, *b = (const cand_t*)pb;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return (a->addr > b->addr) ? -1 : (a->addr < b->addr);
}

int svo_gftt(const uint8_t* img, int w, int h, int pitch, const uint8_t* mask, int mask_pitch, int max_corners,
             double quality, double min_distance, double k, int32_t* xy, int cap) {
    const size_t n = (size_t)w * h;
    float* eig = (float*)malloc(sizeof(float) * n);
    float* dil = (float*)malloc(sizeof(float) * n);
    if (!eig || !dil || svo_harris_response(img, w, h, pitch, k, eig) != 0) { free(eig); free(dil); return -1; }
    /* minMaxLoc over the mask */
    double max_val = 0;
    int have = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            if (!mask || mask[(size_t)y * mask_pitch + x]) {
                float v = eig[(size_t)y * w + x];
                if (!have || v > max_val) { max_val = v; have = 1; }
            }
    /* threshold(eig, eig, maxVal*quality, 0, THRESH_TOZERO) */
    const float thr = (float)(max_val * quality);
    for (size_t i = 0; i < n; ++i) eig[i] = eig[i] > thr ? eig[i] : 0.f;
    /* dilate 3x3 */
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float m = eig[(size_t)y * w + x];
            for (int dy = -1; dy <= 1; ++dy) {
                int yy = y + dy;
                if (yy < 0 || yy >= h) continue;
                for (int dx = -1; dx <= 1; ++dx) {
                    int xx = x + dx;
                    if (xx < 0 || xx >= w) continue;
                    float v = eig[(size_t)yy * w + xx];
                    if (v > m) m = v;
                }
            }
            dil[(size_t)y * w + x] = m;
        }
    size_t nc = 0, ccap = 4096;
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * ccap);
    for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x) {
            float v = eig[(size_t)y * w + x];
            if (v != 0 && v == dil[(size_t)y * w + x] && (!mask || mask[(size_t)y * mask_pitch + x])) {
                if (nc == ccap) { ccap *= 2; c = (cand_t*)realloc(c, sizeof(cand_t) * ccap); }
                c[nc].v = v;
                c[nc].addr = y * w + x;
                ++nc;
            }
        }
    free(eig); free(dil);
    qsort(c, nc, sizeof(cand_t), cand_cmp);
    int count = 0;
    if (min_distance >= 1) {
        const int cell = (int)lround(min_distance);
        const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
        int32_t* head = (int32_t*)malloc(sizeof(int32_t) * gw * gh);
        int32_t* next = (int32_t*)malloc(sizeof(int32_t) * (nc ? nc : 1));
        for (int i = 0; i < gw * gh; ++i) head[i] = -1;
        const double md2 = min_distance * min_distance;
        for (size_t i = 0; i < nc; ++i) {
            const int y = c[i].addr / w, x = c[i].addr - y * w;
            const int cx = x / cell, cy = y / cell;
            int good = 1;
            for (int yy = (cy > 0 ? cy - 1 : 0); good && yy <= (cy + 1 < gh ? cy + 1 : gh - 1); ++yy)
                for (int xx = (cx > 0 ? cx - 1 : 0); good && xx <= (cx + 1 < gw ? cx + 1 : gw - 1); ++xx)
                    for (int32_t j = head[yy * gw + xx]; j >= 0; j = next[j]) {
                        float dx = (float)(x - xy[2 * j]), dy = (float)(y - xy[2 * j + 1]);
                        if ((double)(dx * dx + dy * dy) < md2) { good = 0; break; }
                    }
            if (good) {
                if (count >= cap) break;
                xy[2 * count] = x;
                xy[2 * count + 1] = y;
                next[count] = head[cy * gw + cx];
                head[cy * gw + cx] = count;
                ++count;
                if (max_corners > 0 && count == max_corners) break;
            }
        }
        free(head); free(next);
    } else {
        for (size_t i = 0; i < nc && count < cap; ++i) {
            xy[2 * count] = c[i].addr % w;
            xy[2 * count + 1] = c[i].addr / w;
            ++count;
            if (max_corners > 0 && count == max_corners) break;
        }
    }
    free(c);
    return count;
}

/* ------------------------------------------------------------------ BriefDescriptorExtractor(32)::compute */
static inline long lrint_half_even(float v) { return lrintf(v); } /* cvRound, default rounding mode */

/* `img` is the (ROI) image BRIEF is called on: its own integral image is built here, as OpenCV does.
 * Returns the number of kept key-points; kept[i] = index into pts of the i-th surviving one. */
int svo_brief32(const uint8_t* img, int w, int h, int pitch, const float* pts, int n, uint8_t* desc, int32_t* kept) {
    int nk = 0;
    if (w <= 56 || h <= 56) return 0;
    for (int i = 0; i < n; ++i) {
        long rx = lrint_half_even(pts[2 * i]), ry = lrint_half_even(pts[2 * i + 1]);
        /* x.5 coordinates: cvRound and the sampling centre (int)(pt+0.5) can differ by one and the reference then
         * reads outside its integral image (undefined); defined here: such a key-point is erased */
        const int sx = (int)(pts[2 * i] + 0.5), sy = (int)(pts[2 * i + 1] + 0.5);
        if (rx >= 28 && rx < w - 28 && ry >= 28 && ry < h - 28 && sx >= 28 && sx < w - 28 && sy >= 28 && sy < h - 28) kept[nk++] = i;
    }
    if (!nk) return 0;
    const int sw = w + 1;
    int32_t* sum = (int32_t*)malloc(sizeof(int32_t) * (size_t)sw * (h + 1));
    if (!sum) return -1;
    memset(sum, 0, sizeof(int32_t) * sw);
    for (int y = 0; y < h; ++y) {
        const uint8_t* p = img + (size_t)y * pitch;
        int32_t* s = sum + (size_t)(y + 1) * sw;
        const int32_t* sp = s - sw;
        int32_t run = 0;
        s[0] = 0;
        for (int x = 0; x < w; ++x) { run += p[x]; s[x + 1] = sp[x + 1] + run; }
    }
    for (int kk = 0; kk < nk; ++kk) {
        const float* pt = pts + 2 * kept[kk];
        const int cx = (int)(pt[0] + 0.5), cy = (int)(pt[1] + 0.5);
        uint8_t* d = desc + (size_t)kk * 32;
        for (int j = 0; j < 32; ++j) {
            unsigned byte = 0;
            for (int i = 0; i < 8; ++i) {
                const signed char* t = kPat[8 * j + i];
                int y1 = cy + t[0], x1 = cx + t[1], y2 = cy + t[2], x2 = cx + t[3];
                int32_t s1 = sum[(size_t)(y1 + 5) * sw + x1 + 5] - sum[(size_t)(y1 + 5) * sw + x1 - 4] -
                             sum[(size_t)(y1 - 4) * sw + x1 + 5] + sum[(size_t)(y1 - 4) * sw + x1 - 4];
                int32_t s2 = sum[(size_t)(y2 + 5) * sw + x2 + 5] - sum[(size_t)(y2 + 5) * sw + x2 - 4] -
                             sum[(size_t)(y2 - 4) * sw + x2 + 5] + sum[(size_t)(y2 - 4) * sw + x2 - 4];
                byte |= (unsigned)(s1 < s2) << (7 - i);
            }
            d[j] = (uint8_t)byte;
        }
    }
    free(sum);
    return nk;
}

/* ------------------------------------------------------------------ BFMatcher(NORM_HAMMING)::match, 1 x N */
int svo_match(const uint8_t* q, const uint8_t* t, int n, int32_t* dist_out) {
    int best = -1, bd = 1 << 30;
    uint64_t qa[4];
    memcpy(qa, q, 32);
    for (int i = 0; i < n; ++i) {
        uint64_t ta[4];
        memcpy(ta, t + (size_t)i * 32, 32);
        int d = __builtin_popcountll(qa[0] ^ ta[0]) + __builtin_popcountll(qa[1] ^ ta[1]) +
                __builtin_popcountll(qa[2] ^ ta[2]) + __builtin_popcountll(qa[3] ^ ta[3]);
        if (d < bd) { bd = d; best = i; }
    }
    *dist_out = best >= 0 ? bd : -1;
    return best;
}

/* ------------------------------------------------------------------ CTriangulator */
typedef struct { double f_inv, pu, pv, du_r, du_r_flipped, min_disp; float width; float cutoff; } tri_t;

static void tri_init(tri_t* t, const svo_config* c) {
    t->f_inv = 1.0 / c->P_left[0];
    t->pu = c->P_left[2];
    t->pv = c->P_left[6];
    t->du_r = c->P_right[3];
    t->du_r_flipped = -t->du_r;
    t->min_disp = c->min_disparity;
    t->width = (float)c->width;
    t->cutoff = c->match_cutoff;
}

int svo_point_in_left(const svo_config* c, const float* uvl, const float* uvr, double* xyz) {
    tri_t t;
    tri_init(&t, c);
    const float d = uvl[0] - uvr[0];
    if ((double)d < t.min_disp) return ST_TRI_ZERO_DISP;
    const double z = t.du_r_flipped / (double)d;
    const double fz = t.f_inv * z;
    xyz[0] = fz * ((double)uvl[0] - t.pu);
    xyz[1] = fz * ((double)uvl[1] - t.pv);
    xyz[2] = z;
    return ST_OK;
}

/* shared body of CTriangulator.cpp:59-117 / :264-322 */
static int search(const svo_config* c, const uint8_t* img, int pitch, float u_tl, float v_tl, float size, int first,
                  int n_pool, const uint8_t* ref, float* uv, uint8_t* desc_out, int32_t* dist, int32_t* idx) {
    const float border = 4 * size, full_h = 8 * size + 1;
    float wroi = (float)n_pool + full_h;
    const float wmax = (float)c->width - u_tl;
    if (wmax < wroi) wroi = wmax;
    const int rx = (int)u_tl, ry = (int)v_tl, rw = (int)wroi, rh = (int)full_h;
    *dist = -1;
    *idx = -1;
    if (rx < 0 || ry < 0 || rw < 0 || rh < 0 || rx + rw > c->width || ry + rh > c->height) return ST_TRI_BAD_ROI;
    if (n_pool <= 0) return ST_TRI_NO_DESC;
    float* pool = (float*)malloc(sizeof(float) * 2 * n_pool);
    uint8_t* pd = (uint8_t*)malloc((size_t)32 * n_pool);
    int32_t* kept = (int32_t*)malloc(sizeof(int32_t) * n_pool);
    for (int i = 0; i < n_pool; ++i) {
        pool[2 * i] = (border + (float)i) + (float)first;
        pool[2 * i + 1] = border;
    }
    const int nk = svo_brief32(img + (size_t)ry * pitch + rx, rw, rh, pitch, pool, n_pool, pd, kept);
    int st;
    if (nk <= 0) st = ST_TRI_NO_DESC;
    else {
        int32_t d;
        const int m = svo_match(ref, pd, nk, &d);
        if (m < 0) st = ST_TRI_NO_MATCH;
        else {
            *dist = d;
            *idx = m;
            if (c->match_cutoff > (float)d) {
                uv[0] = pool[2 * kept[m]] + u_tl;
                uv[1] = pool[2 * kept[m] + 1] + v_tl;
                memcpy(desc_out, pd + (size_t)m * 32, 32);
                st = ST_OK;
            } else st = ST_TRI_DISTANCE;
        }
    }
    free(pool); free(pd); free(kept);
    return st;
}

int svo_triangulate_right(const svo_config* c, const uint8_t* img_r, int pitch, float u_tl, float v_tl, float size,
                          const float* uvl, const uint8_t* desc_l, float* uv, double* xyz, uint8_t* desc, int32_t* dist,
                          int32_t* idx) {
    const float border = 4 * size;
    *dist = -1;
    *idx = -1;
    if (uvl[0] <= u_tl + border) return ST_TRI_RANGE;
    const int n_pool = (int)ceilf(uvl[0] - u_tl - border);
    int st = search(c, img_r, pitch, u_tl, v_tl, size, 0, n_pool, desc_l, uv, desc, dist, idx);
    if (st == ST_OK) st = svo_point_in_left(c, uvl, uv, xyz);
    return st;
}

int svo_triangulate_left(const svo_config* c, const uint8_t* img_l, int pitch, float search_range, float u_tl, float v_tl,
                         float size, const float* uvr, const uint8_t* desc_r, float* uv, double* xyz, uint8_t* desc,
                         int32_t* dist, int32_t* idx) {
    *dist = -1;
    *idx = -1;
    if (0 >= search_range) return ST_TRI_RANGE;
    float m = search_range;
    const float wl = (float)c->width - u_tl;
    if (wl < m) m = wl;
    const int n_pool = (int)ceilf(m) + 1;
    int st = search(c, img_l, pitch, u_tl, v_tl, size, 1, n_pool, desc_r, uv, desc, dist, idx);
    if (st == ST_OK) st = svo_point_in_left(c, uv, uvr, xyz);
    return st;
}

/* ------------------------------------------------------------------ addNewLandmarks, one pair */
typedef struct svo_result {
    int32_t capacity;
    float* uv_left; float* uv_right; double* xyz_left; uint8_t* desc_left; uint8_t* desc_right;
    int32_t* distance; int32_t* match_index; uint8_t* status;
} svo_result;

int svo_add_new_landmarks(const svo_config* c, const uint8_t* left, const uint8_t* right, int pitch, const uint8_t* mask,
                          const svo_result* out, size_t o0, int32_t* n_detected) {
    const int cap = c->max_corners > 0 ? c->max_corners : out->capacity;
    int32_t* xy = (int32_t*)malloc(sizeof(int32_t) * 2 * cap);
    const int nd = svo_gftt(left, c->width, c->height, pitch, mask, pitch, c->max_corners, c->quality_level, c->min_distance,
                            c->harris_k, xy, cap);
    if (nd < 0) { free(xy); return -1; }
    if (n_detected) *n_detected = nd;
    float* pts = (float*)malloc(sizeof(float) * 2 * (nd ? nd : 1));
    int32_t* kept = (int32_t*)malloc(sizeof(int32_t) * (nd ? nd : 1));
    uint8_t* dl = (uint8_t*)malloc((size_t)32 * (nd ? nd : 1));
    for (int i = 0; i < nd; ++i) { pts[2 * i] = (float)xy[2 * i]; pts[2 * i + 1] = (float)xy[2 * i + 1]; }
    const int nk = svo_brief32(left, c->width, c->height, pitch, pts, nd, dl, kept);
    const float size = c->keypoint_size;
    for (int u = 0; u < nk && u < out->capacity; ++u) {
        const size_t o = o0 + u;
        const float x = pts[2 * kept[u]], y = pts[2 * kept[u] + 1];
        float u_tl = x - c->search_range - 4 * size;
        if (u_tl < 0.0f) u_tl = 0.0f;
        const float v_tl = y - 4 * size;
        float uvl[2] = {x, y}, uvr[2] = {0, 0};
        double xyz[3] = {0, 0, 0};
        uint8_t dr[32];
        int32_t dist, idx;
        memset(dr, 0, 32);
        const int st = svo_triangulate_right(c, right, pitch, u_tl, v_tl, size, uvl, dl + (size_t)u * 32, uvr, xyz, dr, &dist, &idx);
        out->uv_left[2 * o] = x; out->uv_left[2 * o + 1] = y;
        memcpy(out->desc_left + o * 32, dl + (size_t)u * 32, 32);
        out->status[o] = (uint8_t)st;
        out->distance[o] = dist;
        out->match_index[o] = idx;
        if (st == ST_OK) {
            out->uv_right[2 * o] = uvr[0]; out->uv_right[2 * o + 1] = uvr[1];
            out->xyz_left[3 * o] = xyz[0]; out->xyz_left[3 * o + 1] = xyz[1]; out->xyz_left[3 * o + 2] = xyz[2];
            memcpy(out->desc_right + o * 32, dr, 32);
        }
    }
    free(xy); free(pts); free(kept); free(dl);
    return nk;
}

/* ------------------------------------------------------------------ batch driver (frame-parallel) */
typedef struct {
    const svo_config* c; const uint8_t* left; const uint8_t* right; const uint8_t* masks;
    int pitch; size_t stride; int n_frames; const svo_result* out; int32_t* n_kp; int32_t* n_det;
    int next; pthread_mutex_t mu; int error;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        const int f = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (f >= j->n_frames) break;
        int32_t nd = 0;
        const int nk = svo_add_new_landmarks(j->c, j->left + f * j->stride, j->right + f * j->stride, j->pitch,
                                             j->masks ? j->masks + f * j->stride : NULL, j->out,
                                             (size_t)f * j->out->capacity, &nd);
        if (nk < 0) j->error = 1;
        j->n_kp[f] = nk;
        if (j->n_det) j->n_det[f] = nd;
    }
    return NULL;
}

int svo_stereo_frames_mt(const svo_config* c, const uint8_t* left, const uint8_t* right, int pitch, size_t frame_stride,
                         int n_frames, const uint8_t* masks, const svo_result* out, int32_t* n_keypoints,
                         int32_t* n_detected, int n_threads) {
    job_t j;
    j.c = c; j.left = left; j.right = right; j.masks = masks; j.pitch = pitch; j.stride = frame_stride;
    j.n_frames = n_frames; j.out = out; j.n_kp = n_keypoints; j.n_det = n_detected; j.next = 0; j.error = 0;
    pthread_mutex_init(&j.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_frames) n_threads = n_frames > 0 ? n_frames : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    for (int i = 1; i < n_threads; ++i) pthread_create(&th[i], NULL, worker, &j);
    worker(&j);
    for (int i = 1; i < n_threads; ++i) pthread_join(th[i], NULL);
    free(th);
    pthread_mutex_destroy(&j.mu);
    return j.error ? -1 : 0;
}
