// brief_match.cuh -- BRIEF-32 on box-sum images, dense scan-line Hamming search, triangulation.
//
// Replaces the bodies of (reference paths):
//   cv::xfeatures2d::BriefDescriptorExtractor::compute        src/core/CTriangulator.cpp:83,217,288
//   cv::BFMatcher(NORM_HAMMING)::match (1 x N arg-min)         src/core/CTriangulator.cpp:93,227,298
//   CTriangulator::getPointTriangulatedInRIGHT[Full]          src/core/CTriangulator.cpp:51-119,185-253
//   CTriangulator::getPointTriangulatedInLEFT (7 args)        src/core/CTriangulator.cpp:255-324
//   CTriangulator::getPointInLEFT                             src/core/CTriangulator.cpp:326-356
//   the per-key-point loop of addNewLandmarks                  src/core/CFundamentalMatcher.cpp:109-175
//   trackManual stage 1                                        src/core/CFundamentalMatcher.cpp:1404-1538
//
// Work decomposition: one warp per query.  The candidates of a scan-line search are consecutive
// pixels of one image row, so lane l evaluates candidates 2l and 2l+1 of a 64-candidate chunk
// from ONE 32-bit shared-memory load per test point: the u16 box sums of two neighbouring
// candidates are neighbours in memory.  The 256 test pairs are template constants
// (brief_pattern_32.h), so every shared-memory offset is an instruction immediate.  Both
// comparisons of a pair are one 32-bit subtract (values <= 20655 < 2^15 leave the half-word sign
// bits free): d = b + 0x7FFF7FFF - a has bit 15 / bit 31 set iff a < b in the low / high half.
#pragma once
#include <utility>
#include "brief_pattern_32.h"
#include "common.cuh"

namespace svi {

constexpr signed char kPat[SVI_BRIEF_NTESTS][4] = SVI_BRIEF_PATTERN_INIT;
__device__ const signed char d_pat[SVI_BRIEF_NTESTS][4] = SVI_BRIEF_PATTERN_INIT;

constexpr int PATCH_ROWS = 2 * kBriefReach + 1;        // 49
constexpr int PATCH_CHUNK = 64;                        // candidates per pass
constexpr int PATCH_W = PATCH_CHUNK + 2 * kBriefReach; // 112 u16 per row
constexpr int PATCH_WORDS = PATCH_W / 2;               // 56
constexpr int PATCH_COPY_WORDS = PATCH_ROWS * PATCH_WORDS;
constexpr int MATCH_WARPS = 4;
constexpr int MATCH_SMEM_PER_WARP = 2 * PATCH_COPY_WORDS * 4;  // even- and odd-aligned copies
constexpr int MATCH_SMEM = MATCH_WARPS * MATCH_SMEM_PER_WARP;

// ------------------------------------------------------------------ descriptor at one point
// Lanes = tests (8 rounds of 32).  Word format used throughout the kernels: word j holds tests
// 32j..32j+31 with test 32j+t at bit 31-t, i.e. the big-endian read of descriptor bytes 4j..4j+3.
__device__ __forceinline__ void brief_at_point(const uint16_t* __restrict__ box, int box_pitch, int cx,
                                               int cy, int lane, uint32_t (&w)[kDescWords]) {
#pragma unroll
    for (int j = 0; j < kDescWords; ++j) {
        const signed char* p = d_pat[32 * j + lane];
        uint32_t s1 = box[(size_t)(cy + p[0]) * box_pitch + cx + p[1]];
        uint32_t s2 = box[(size_t)(cy + p[2]) * box_pitch + cx + p[3]];
        w[j] = __brev(__ballot_sync(0xFFFFFFFFu, s1 < s2));
    }
}

__device__ __forceinline__ uint32_t desc_word_to_bytes(uint32_t w) { return __byte_perm(w, 0, 0x0123); }

__device__ __forceinline__ void store_desc(uint8_t* dst, const uint32_t (&w)[kDescWords], int lane) {
    uint32_t v = 0;
#pragma unroll
    for (int j = 0; j < kDescWords; ++j) if (lane == j) v = w[j];
    if (lane < kDescWords) reinterpret_cast<uint32_t*>(dst)[lane] = desc_word_to_bytes(v);
}

__device__ __forceinline__ void load_desc(const uint8_t* src, uint32_t (&w)[kDescWords]) {
#pragma unroll
    for (int j = 0; j < kDescWords; ++j) w[j] = desc_word_to_bytes(reinterpret_cast<const uint32_t*>(src)[j]);
}

__device__ __forceinline__ int hamming_words(const uint32_t (&a)[kDescWords], const uint32_t (&b)[kDescWords]) {
    int d = 0;
#pragma unroll
    for (int j = 0; j < kDescWords; ++j) d += __popc(a[j] ^ b[j]);
    return d;
}

// ------------------------------------------------------------------ unrolled pair tests
template <int T>
__device__ __forceinline__ void brief_pair_test(const uint32_t* __restrict__ Al, const uint32_t* __restrict__ Bl,
                                                uint32_t& wlo, uint32_t& whi) {
    constexpr int c1 = kPat[T][1] + kBriefReach, c2 = kPat[T][3] + kBriefReach;
    constexpr int o1 = (kPat[T][0] + kBriefReach) * PATCH_WORDS + (c1 >> 1);
    constexpr int o2 = (kPat[T][2] + kBriefReach) * PATCH_WORDS + (c2 >> 1);
    const uint32_t a = (c1 & 1) ? Bl[o1] : Al[o1];
    const uint32_t b = (c2 & 1) ? Bl[o2] : Al[o2];
    const uint32_t d = b + 0x7FFF7FFFu - a;
    whi = __funnelshift_l(d, whi, 1);
    wlo = __funnelshift_l(d << 16, wlo, 1);
}

template <int J, int... I>
__device__ __forceinline__ void brief_pair_word(const uint32_t* __restrict__ Al, const uint32_t* __restrict__ Bl,
                                                uint32_t& wlo, uint32_t& whi, std::integer_sequence<int, I...>) {
    wlo = 0u;
    whi = 0u;
    (brief_pair_test<J * 32 + I>(Al, Bl, wlo, whi), ...);
}

template <int... J>
__device__ __forceinline__ void brief_pair_all(const uint32_t* __restrict__ Al, const uint32_t* __restrict__ Bl,
                                               uint32_t (&wlo)[kDescWords], uint32_t (&whi)[kDescWords],
                                               std::integer_sequence<int, J...>) {
    (brief_pair_word<J>(Al, Bl, wlo[J], whi[J], std::make_integer_sequence<int, 32>{}), ...);
}

struct SearchResult {
    int status, dist, idx;
    float u, v;
    uint32_t w[kDescWords];
};

// Stage the (49 x 112) u16 window of `box` whose top-left is (row0, col0) into the warp's two
// shared copies: A[r][k] = box[row0+r][col0+k], B[r][k] = A[r][k+1].
__device__ __forceinline__ void load_patch(const uint16_t* __restrict__ box, int box_pitch, int W, int H,
                                           int row0, int col0, uint16_t* __restrict__ A, uint16_t* __restrict__ B,
                                           int lane) {
    for (int idx = lane; idx < PATCH_ROWS * PATCH_W; idx += 32) {
        int r = idx / PATCH_W, k = idx - r * PATCH_W;
        int gy = row0 + r, gx = col0 + k;
        uint16_t v = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(box + (size_t)gy * box_pitch + gx);
        A[idx] = v;
        if (k > 0) B[idx - 1] = v;
    }
}

// Shared body of CTriangulator.cpp:59-117 (RIGHT, first = 0) and :264-322 (LEFT, first = 1):
// ROI by truncation, pool key-points (border + first + i, border), BRIEF's border filter inside
// the ROI, descriptors of the survivors, first arg-min, cut-off.  Warp-uniform arguments.
__device__ __forceinline__ void scanline_search(const uint16_t* __restrict__ box, int box_pitch, int W, int H,
                                                float img_width_f, float u_tl, float v_tl, float size,
                                                int first, int n_pool, const uint32_t (&ref)[kDescWords],
                                                float cutoff, uint32_t* __restrict__ smem, int lane,
                                                SearchResult& out) {
    const float border = 4.f * size, full_h = 8.f * size + 1.f;
    const float roi_w_f = fminf((float)n_pool + full_h, img_width_f - u_tl);
    const int rx = (int)u_tl, ry = (int)v_tl, rw = (int)roi_w_f, rh = (int)full_h;  // cv::Rect(float...) truncates
    out.dist = -1;
    out.idx = -1;
    if (rx < 0 || ry < 0 || rw < 0 || rh < 0 || rx + rw > W || ry + rh > H) { out.status = SVI_TRI_BAD_ROI; return; }
    // KeyPointsFilter::runByImageBorder inside the ROI: keep 28 <= cvRound(pt) < size - 28
    const int ky_r = cv_round_f(border);
    int i_lo = 0, i_hi = 0;
    if (ky_r >= kBriefBorder && ky_r < rh - kBriefBorder) {
        for (int i0 = 0; i0 < n_pool; i0 += 32) {
            int i = i0 + lane;
            int kr = cv_round_f((border + (float)i) + (float)first);
            bool in = i < n_pool;
            i_lo += __popc(__ballot_sync(0xFFFFFFFFu, in && kr < kBriefBorder));
            i_hi += __popc(__ballot_sync(0xFFFFFFFFu, in && kr < rw - kBriefBorder));
        }
    }
    const int n_valid = i_hi - i_lo;
    if (n_valid <= 0) { out.status = SVI_TRI_NO_DESC; return; }
    const float kx_lo = (border + (float)i_lo) + (float)first;
    const int gx_lo = rx + brief_centre(kx_lo), gy = ry + brief_centre(border);

    uint16_t* A = reinterpret_cast<uint16_t*>(smem);
    uint16_t* B = A + PATCH_ROWS * PATCH_W;
    const uint32_t* Al = smem + lane;
    const uint32_t* Bl = smem + PATCH_COPY_WORDS + lane;
    uint32_t best_key = 0xFFFFFFFFu;
    for (int cb = 0; cb < n_valid; cb += PATCH_CHUNK) {
        __syncwarp();
        load_patch(box, box_pitch, W, H, gy - kBriefReach, gx_lo + cb - kBriefReach, A, B, lane);
        __syncwarp();
        uint32_t wlo[kDescWords], whi[kDescWords];
        brief_pair_all(Al, Bl, wlo, whi, std::make_integer_sequence<int, kDescWords>{});
        const int c_lo = cb + 2 * lane, c_hi = c_lo + 1;
        uint32_t k_lo = (c_lo < n_valid) ? (((uint32_t)hamming_words(wlo, ref) << 16) | (uint32_t)(c_lo & 0xFFFF)) : 0xFFFFFFFFu;
        uint32_t k_hi = (c_hi < n_valid) ? (((uint32_t)hamming_words(whi, ref) << 16) | (uint32_t)(c_hi & 0xFFFF)) : 0xFFFFFFFFu;
        const uint32_t k_mine = min(k_lo, k_hi);
        const uint32_t k_min = warp_min_u32(k_mine);
        if (k_min < best_key) {  // strict: earlier chunks win ties (BFMatcher keeps the first minimum)
            best_key = k_min;
            const int owner = __ffs(__ballot_sync(0xFFFFFFFFu, k_mine == k_min)) - 1;
            const bool hi = (k_hi == k_min) && (k_lo != k_min);
#pragma unroll
            for (int j = 0; j < kDescWords; ++j) out.w[j] = __shfl_sync(0xFFFFFFFFu, hi ? whi[j] : wlo[j], owner);
        }
    }
    // the key carries the candidate index in 16 bits (n_valid < 65536: image widths are 16-bit)
    out.dist = (int)(best_key >> 16);
    out.idx = (int)(best_key & 0xFFFFu);
    if (!(cutoff > (float)out.dist)) { out.status = SVI_TRI_DISTANCE; return; }
    const float px = (border + (float)(i_lo + out.idx)) + (float)first;
    out.u = px + u_tl;
    out.v = border + v_tl;
    out.status = SVI_OK;
}

// CTriangulator::getPointInLEFT :326-356, fp64, evaluation order of the reference, no FMA.
__device__ __forceinline__ int point_in_left(const TriConst& tc, float xl, float yl, float xr, double* xyz) {
    const float d = __fsub_rn(xl, xr);
    if ((double)d < tc.min_disp) return SVI_TRI_ZERO_DISP;
    const double z = __ddiv_rn(tc.du_r_flipped, (double)d);
    const double fz = __dmul_rn(tc.f_inv, z);
    xyz[0] = __dmul_rn(fz, __dsub_rn((double)xl, tc.pu));
    xyz[1] = __dmul_rn(fz, __dsub_rn((double)yl, tc.pv));
    xyz[2] = z;
    return SVI_OK;
}

// getPointTriangulatedInRIGHT: range check + pool size (:63-67), then the shared body.
__device__ __forceinline__ void triangulate_right_dev(const uint16_t* __restrict__ box_r, const FrameGeom& g,
                                                      const TriConst& tc, float u_tl, float v_tl, float size,
                                                      float xl, float yl, const uint32_t (&ref)[kDescWords],
                                                      uint32_t* smem, int lane, SearchResult& r, double* xyz) {
    const float border = 4.f * size;
    r.dist = -1; r.idx = -1;
    if (xl <= u_tl + border) { r.status = SVI_TRI_RANGE; return; }
    const int n_pool = (int)ceilf((xl - u_tl) - border);
    scanline_search(box_r, g.box_pitch, g.W, g.H, tc.width_right, u_tl, v_tl, size, 0, n_pool, ref,
                    tc.match_cutoff, smem, lane, r);
    if (r.status == SVI_OK) r.status = point_in_left(tc, xl, yl, r.u, xyz);
}

// getPointTriangulatedInLEFT (7 args): :268-272 then the shared body; the LEFT point is the match.
__device__ __forceinline__ void triangulate_left_dev(const uint16_t* __restrict__ box_l, const FrameGeom& g,
                                                     const TriConst& tc, float search_range, float u_tl, float v_tl,
                                                     float size, float xr, float yr, const uint32_t (&ref)[kDescWords],
                                                     uint32_t* smem, int lane, SearchResult& r, double* xyz) {
    r.dist = -1; r.idx = -1;
    if (0.f >= search_range) { r.status = SVI_TRI_RANGE; return; }
    const int n_pool = (int)ceilf(fminf(search_range, tc.width_left - u_tl)) + 1;
    scanline_search(box_l, g.box_pitch, g.W, g.H, tc.width_left, u_tl, v_tl, size, 1, n_pool, ref,
                    tc.match_cutoff, smem, lane, r);
    if (r.status == SVI_OK) r.status = point_in_left(tc, r.u, r.v, xr, xyz);
    (void)yr;
}

struct StereoOutDev {
    int cap;
    float* uv_l; float* uv_r; double* xyz; uint8_t* desc_l; uint8_t* desc_r;
    int* dist; int* idx; uint8_t* status;
};

// K5: per key-point of addNewLandmarks (:109-175): LEFT descriptor, scan-line search in RIGHT
// with uTL = max(0, x - range - 4*size), vTL = y - 4*size (:120-121), triangulation, outputs.
__global__ void __launch_bounds__(MATCH_WARPS * 32)
stereo_match_kernel(const uint16_t* __restrict__ box_l, const uint16_t* __restrict__ box_r, FrameGeom g,
                    TriConst tc, float size, float range, const ushort2* __restrict__ kp_xy,
                    const int* __restrict__ n_kp, int max_corners, StereoOutDev out, int out_frame0) {
    extern __shared__ __align__(16) uint32_t match_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y, slot = blockIdx.x * MATCH_WARPS + warp;
    if (slot >= n_kp[f]) return;
    uint32_t* smem = match_smem + warp * (MATCH_SMEM_PER_WARP / 4);
    const ushort2 kp = kp_xy[(size_t)f * max_corners + slot];
    const uint16_t* bl = box_l + (size_t)f * g.H * g.box_pitch;
    const uint16_t* br = box_r + (size_t)f * g.H * g.box_pitch;
    const float x = (float)kp.x, y = (float)kp.y;

    uint32_t ref[kDescWords];
    brief_at_point(bl, g.box_pitch, kp.x, kp.y, lane, ref);

    const float u_tl = fmaxf(0.f, (x - range) - 4.f * size);
    const float v_tl = y - 4.f * size;
    SearchResult r;
    double xyz[3] = {0.0, 0.0, 0.0};
    triangulate_right_dev(br, g, tc, u_tl, v_tl, size, x, y, ref, smem, lane, r, xyz);

    const size_t o = (size_t)(out_frame0 + f) * out.cap + slot;
    store_desc(out.desc_l + o * 32, ref, lane);
    if (r.status == SVI_OK) store_desc(out.desc_r + o * 32, r.w, lane);
    if (lane == 0) {
        out.uv_l[o * 2] = x; out.uv_l[o * 2 + 1] = y;
        out.status[o] = (uint8_t)r.status;
        out.dist[o] = r.dist;
        out.idx[o] = r.idx;
        if (r.status == SVI_OK) {
            out.uv_r[o * 2] = r.u; out.uv_r[o * 2 + 1] = r.v;
            out.xyz[o * 3] = xyz[0]; out.xyz[o * 3 + 1] = xyz[1]; out.xyz[o * 3 + 2] = xyz[2];
        }
    }
}

// svi_describe: BRIEF-32 at n points of one image (full-image border filter).
__global__ void describe_kernel(const uint16_t* __restrict__ box, FrameGeom g, const float* __restrict__ xy, int n,
                                uint8_t* __restrict__ desc, uint8_t* __restrict__ kept) {
    const int lane = threadIdx.x & 31, q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= n) return;
    const float x = xy[2 * q], y = xy[2 * q + 1];
    const int rx = cv_round_f(x), ry = cv_round_f(y);
    const bool ok = rx >= kBriefBorder && rx < g.W - kBriefBorder && ry >= kBriefBorder && ry < g.H - kBriefBorder;
    uint32_t w[kDescWords] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (ok) brief_at_point(box, g.box_pitch, brief_centre(x), brief_centre(y), lane, w);
    store_desc(desc + (size_t)q * 32, w, lane);
    if (lane == 0) kept[q] = ok ? 1 : 0;
}

// svi_match_hamming: BFMatcher(NORM_HAMMING).match, one warp per query row.
__global__ void hamming_match_kernel(const uint8_t* __restrict__ q32, int nq, const uint8_t* __restrict__ t32, int nt,
                                     int* __restrict__ index, int* __restrict__ distance) {
    const int lane = threadIdx.x & 31, q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    uint32_t ref[kDescWords];
    load_desc(q32 + (size_t)q * 32, ref);
    unsigned long long best = ~0ull;
    for (int t = lane; t < nt; t += 32) {
        uint32_t w[kDescWords];
        load_desc(t32 + (size_t)t * 32, w);
        unsigned long long k = ((unsigned long long)hamming_words(ref, w) << 32) | (unsigned)t;
        best = min(best, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
    if (lane == 0) {
        index[q] = nt > 0 ? (int)(best & 0xFFFFFFFFu) : -1;
        distance[q] = nt > 0 ? (int)(best >> 32) : -1;
    }
}

struct TriOutDev {
    float* uv; double* xyz; uint8_t* desc; int* dist; int* idx; uint8_t* status;
};

// svi_triangulate_right / svi_triangulate_left: one warp per query against one image.
template <bool kLeft>
__global__ void __launch_bounds__(MATCH_WARPS * 32)
triangulate_kernel(const uint16_t* __restrict__ box, FrameGeom g, TriConst tc, int n, const float* __restrict__ search_range,
                   const float* __restrict__ top_left, const float* __restrict__ uv_ref,
                   const uint8_t* __restrict__ desc_ref, float size, TriOutDev out) {
    extern __shared__ __align__(16) uint32_t match_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * MATCH_WARPS + warp;
    if (q >= n) return;
    uint32_t* smem = match_smem + warp * (MATCH_SMEM_PER_WARP / 4);
    uint32_t ref[kDescWords];
    load_desc(desc_ref + (size_t)q * 32, ref);
    SearchResult r;
    double xyz[3] = {0.0, 0.0, 0.0};
    const float u_tl = top_left[2 * q], v_tl = top_left[2 * q + 1], xq = uv_ref[2 * q], yq = uv_ref[2 * q + 1];
    if (kLeft) triangulate_left_dev(box, g, tc, search_range[q], u_tl, v_tl, size, xq, yq, ref, smem, lane, r, xyz);
    else triangulate_right_dev(box, g, tc, u_tl, v_tl, size, xq, yq, ref, smem, lane, r, xyz);
    if (r.status == SVI_OK) store_desc(out.desc + (size_t)q * 32, r.w, lane);
    if (lane == 0) {
        out.status[q] = (uint8_t)r.status;
        out.dist[q] = r.dist;
        out.idx[q] = r.idx;
        if (r.status == SVI_OK) {
            out.uv[2 * q] = r.u; out.uv[2 * q + 1] = r.v;
            out.xyz[3 * q] = xyz[0]; out.xyz[3 * q + 1] = xyz[1]; out.xyz[3 * q + 2] = xyz[2];
        }
    }
}

__global__ void point_in_left_kernel(TriConst tc, int n, const float* __restrict__ uvl, const float* __restrict__ uvr,
                                     double* __restrict__ xyz, uint8_t* __restrict__ status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3] = {0.0, 0.0, 0.0};
    int st = point_in_left(tc, uvl[2 * i], uvl[2 * i + 1], uvr[2 * i], p);
    status[i] = (uint8_t)st;
    xyz[3 * i] = p[0]; xyz[3 * i + 1] = p[1]; xyz[3 * i + 2] = p[2];
}

// ------------------------------------------------------------------ tracking, stage 1
struct TrackConst {
    double T[12];     // rows 0..2 of WORLDtoLEFT
    double PL[12], PR[12];
    float tri_scale;  // 1 + motion scaling (CFundamentalMatcher.cpp:1363)
    float cutoff1;    // m_dMatchingDistanceCutoffTrackingStage1
};
struct LandmarksDev {
    const double* xyz_w; const uint8_t* desc_l; const uint8_t* desc_r; const float* disparity; const float* size;
};
struct TrackOutDev {
    uint8_t* status; uint8_t* stage; float* uv_l; float* uv_r; double* xyz; uint8_t* desc_l; uint8_t* desc_r;
};

// CPinholeCamera::getProjectionRounded (src/vision/CPinholeCamera.h:202-210)
__device__ __forceinline__ void projection_rounded(const double* P, const double* p, float& u, float& v) {
    double h0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[0], p[0]), __dmul_rn(P[1], p[1])), __dmul_rn(P[2], p[2])), P[3]);
    double h1 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[4], p[0]), __dmul_rn(P[5], p[1])), __dmul_rn(P[6], p[2])), P[7]);
    double h2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[8], p[0]), __dmul_rn(P[9], p[1])), __dmul_rn(P[10], p[2])), P[11]);
    u = round_half_away((float)__ddiv_rn(h0, h2));
    v = round_half_away((float)__ddiv_rn(h1, h2));
}

__global__ void __launch_bounds__(MATCH_WARPS * 32)
track_stage1_kernel(const uint16_t* __restrict__ box_l, const uint16_t* __restrict__ box_r, FrameGeom g, TriConst tc,
                    TrackConst k, LandmarksDev lm, int n, TrackOutDev out) {
    extern __shared__ __align__(16) uint32_t match_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * MATCH_WARPS + warp;
    if (q >= n) return;
    uint32_t* smem = match_smem + warp * (MATCH_SMEM_PER_WARP / 4);
    const double wx = lm.xyz_w[3 * q], wy = lm.xyz_w[3 * q + 1], wz = lm.xyz_w[3 * q + 2];
    double p[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)   // Isometry3d * Vector3d = linear * v + translation (:1404)
        p[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(k.T[4 * r], wx), __dmul_rn(k.T[4 * r + 1], wy)), __dmul_rn(k.T[4 * r + 2], wz)), k.T[4 * r + 3]);
    float ul, vl, ur, vr;
    projection_rounded(k.PL, p, ul, vl);
    projection_rounded(k.PR, p, ur, vr);
    const float size = lm.size[q], half = 4.f * size;
    const float search = k.tri_scale * lm.disparity[q];
    int status = SVI_TRK_OUT_OF_FOV, stage = 0;
    SearchResult r;
    double xyz[3] = {0.0, 0.0, 0.0};
    uint32_t mine[kDescWords], last_l[kDescWords], last_r[kDescWords];
    float o_ul = 0.f, o_vl = 0.f, o_ur = 0.f, o_vr = 0.f;
    const bool in_fov = ul >= 28.f && ul < (float)(g.W - 28) && vl >= 28.f && vl < (float)(g.H - 28) &&
                        ur >= 28.f && ur < (float)(g.W - 28) && vr >= 28.f && vr < (float)(g.H - 28);
    // the single key-point (half, half) of the (8*size+1)^2 ROI must survive BRIEF's border filter
    const int roi_len = (int)(8.f * size + 1.f);
    const bool kp_ok = cv_round_f(half) >= kBriefBorder && cv_round_f(half) < roi_len - kBriefBorder;
    if (in_fov) {
        load_desc(lm.desc_l + (size_t)q * 32, last_l);
        load_desc(lm.desc_r + (size_t)q * 32, last_r);
        // STAGE 1 LEFT :1419-1476 -- descriptor exactly at the projection, then search RIGHT
        {
            const float roi_x = ul - half, roi_y = vl - half;
            const bool roi_ok = (int)roi_x >= 0 && (int)roi_y >= 0 && (int)roi_x + roi_len <= g.W && (int)roi_y + roi_len <= g.H;
            if (kp_ok && roi_ok)
                brief_at_point(box_l, g.box_pitch, (int)roi_x + brief_centre(half), (int)roi_y + brief_centre(half), lane, mine);
            status = roi_ok ? SVI_TRK_STAGE1_DIST : SVI_TRI_BAD_ROI;
            if (kp_ok && roi_ok && k.cutoff1 > (float)hamming_words(last_l, mine)) {
                triangulate_right_dev(box_r, g, tc, fmaxf(0.f, roi_x - search), roi_y, size, roi_x + half, roi_y + half,
                                      mine, smem, lane, r, xyz);
                status = r.status;
                if (status == SVI_OK) {
                    if (tc.depth_min > xyz[2] || tc.depth_max < xyz[2]) status = SVI_TRK_DEPTH;
                    else if (k.cutoff1 < (float)hamming_words(last_r, r.w)) status = SVI_TRK_TRI_DESC;
                    else { stage = 1; o_ul = ul; o_vl = vl; o_ur = r.u; o_vr = r.v; }
                }
            }
        }
        // STAGE 1 RIGHT :1480-1538
        if (stage == 0) {
            const float roi_x = ur - half, roi_y = vr - half;
            const bool roi_ok = (int)roi_x >= 0 && (int)roi_y >= 0 && (int)roi_x + roi_len <= g.W && (int)roi_y + roi_len <= g.H;
            if (kp_ok && roi_ok)
                brief_at_point(box_r, g.box_pitch, (int)roi_x + brief_centre(half), (int)roi_y + brief_centre(half), lane, mine);
            status = roi_ok ? SVI_TRK_STAGE1_DIST : SVI_TRI_BAD_ROI;
            if (kp_ok && roi_ok && k.cutoff1 > (float)hamming_words(last_r, mine)) {
                triangulate_left_dev(box_l, g, tc, search, roi_x, roi_y, size, roi_x + half, roi_y + half, mine, smem,
                                     lane, r, xyz);
                status = r.status;
                if (status == SVI_OK) {
                    if (tc.depth_min > xyz[2] || tc.depth_max < xyz[2]) status = SVI_TRK_DEPTH;
                    else if (k.cutoff1 < (float)hamming_words(last_l, r.w)) status = SVI_TRK_TRI_DESC;
                    else { stage = 2; o_ul = r.u; o_vl = r.v; o_ur = ur; o_vr = vr; }
                }
            }
        }
    }
    if (stage == 1) { store_desc(out.desc_l + (size_t)q * 32, mine, lane); store_desc(out.desc_r + (size_t)q * 32, r.w, lane); }
    if (stage == 2) { store_desc(out.desc_l + (size_t)q * 32, r.w, lane); store_desc(out.desc_r + (size_t)q * 32, mine, lane); }
    if (lane == 0) {
        out.status[q] = (uint8_t)status;
        out.stage[q] = (uint8_t)stage;
        if (stage) {
            out.uv_l[2 * q] = o_ul; out.uv_l[2 * q + 1] = o_vl;
            out.uv_r[2 * q] = o_ur; out.uv_r[2 * q + 1] = o_vr;
            out.xyz[3 * q] = xyz[0]; out.xyz[3 * q + 1] = xyz[1]; out.xyz[3 * q + 2] = xyz[2];
        }
    }
}

}  // namespace svi
