// binned.cuh -- the batch form of the new-landmark matcher: key-points binned by image position, one TMA tile per bin.
//
// Same results as stereo_match_kernel / describe_left_kernel (CTriangulator::getPointTriangulatedInRIGHTFull,
// reference src/core/CTriangulator.cpp:51-119, called per key-point by addNewLandmarks, CFundamentalMatcher.cpp:109-175),
// different data movement.  The per-key-point kernels stage a 49 x 120 window of the RIGHT box-sum plane (twice: even and
// odd alignment, 23.5 KB) for EVERY key-point: 2.5 GB of L2 -> SM traffic per 64-frame launch (ncu: 10.6 TB/s of
// l1tex__m_xbar2l1tex_read_bytes), and -- what turned out to matter more -- 23 KB of shared memory per WARP, which holds
// an SM at 9 resident warps whose LDS -> compare -> LOP3 chains cannot fill the shared-memory pipe (72 %).
// Neighbouring key-points search overlapping windows, so here the key-points of a frame are first sorted into bins of
// BIN_W x BIN_H pixels (bin_keypoints_kernel, a counting sort in shared memory); a CTA then owns one bin, brings in ONE
// tile that covers the windows of all of the bin's key-points (two TMA loads: the plane and its one-element-shifted copy)
// and its warps walk the bin's key-points with the same unrolled tests as before -- every shared-memory offset is still an
// instruction immediate, only the row pitch changed.  Tile traffic per key-point drops to about a third, the shared
// memory per warp to 14 KB: 15 warps per SM (the registers are the limit now), shared-memory pipe 78 %.
// The LEFT descriptors are computed the same way from one LEFT tile per bin.
//
// Used when the scan line fits one pass (pool size <= PATCH_CHUNK candidates: the reference's 60-px range) and the ROI
// border 4 * size is a whole number of pixels (the reference's size 7); everything else keeps the per-key-point kernels.
#pragma once
#include "brief_match.cuh"
#include "select.cuh"

namespace svi {

// Bin and CTA shape.  Measured on C2 (1636 key-points per 1241 x 376 frame, device-resident, A/B builds on one box,
// frames/s; per-key-point kernels: 103.6 k):
//   64 x 32 bins, 4 warps x 3 CTAs/SM 102.8 k | 64 x 32, 5 x 3 105.2 k | 128 x 32, 8 x 2 106.4 k | 64 x 16, 4 x 4 102.7 k |
//   96 x 24, 5 x 3 107.2 k | 64 x 24, 4 x 4 108.1 k | 128 x 24, 5 x 3 108.9 k
// The matcher's time barely moves with the tile traffic; it follows the number of resident warps (9 per SM for the
// per-key-point kernels, 15-16 here at 127 registers) and the balance inside a CTA (key-points per bin over warps per CTA).
// Sparse frames (maxCorners 1000 at KITTI size: ~7 key-points per 128 x 24 bin, ~3.5 per 64 x 24 bin) are faster with the
// per-key-point kernels (C4: 136.0 k vs 132.2 k / 129.5 k; 128 x 48 bins with 7 warps x 2 CTAs: 132.4 k), so svi_create
// selects this path by expected density.
template <int BW, int BH, int NW, int NCTA, int DW>
struct BinGeom {
    static constexpr int BIN_W = BW, BIN_H = BH;
    static constexpr int ROWS = BH + 2 * kBriefReach;                                   // rows y - 24 .. y + 24 for every y of the bin
    static constexpr int REACH = PATCH_CHUNK;                                           // candidates start at most 62 px left of the key-point
    static constexpr int COLS = (BW + REACH + 2 * kBriefReach + 7 + 7) / 8 * 8;         // + the 8-element alignment of the tile start
    static constexpr int WORDS = COLS / 2;
    static constexpr int COPY_BYTES = ROWS * COLS * 2;                                  // one TMA box
    static constexpr int COPY_STRIDE = (COPY_BYTES + 127) / 128 * 128;
    static constexpr int M_WARPS = NW, M_CTAS = NCTA;                                   // matcher: warps per CTA, CTAs per SM
    static constexpr int M_SMEM = 2 * COPY_STRIDE + 16;                                 // two copies + the CTA's mbarrier
    static constexpr int D_COLS = BW + 2 * kBriefReach + 8;                             // LEFT tile of a bin
    static constexpr int D_BYTES = ROWS * D_COLS * 2;
    static constexpr int D_WARPS = DW;
    static constexpr int D_SMEM = (D_BYTES + 127) / 128 * 128 + 16;
    static_assert(COLS <= 256 && ROWS <= 256 && D_COLS <= 256, "TMA box dimensions");
    static_assert(COLS % 8 == 0 && D_COLS % 8 == 0, "TMA rows are multiples of 16 bytes");
    static_assert(M_SMEM * NCTA <= 227 * 1024 - 1024 * NCTA, "the CTAs of an SM fit its shared memory");
};
using BinWide = BinGeom<128, 24, 5, 3, 4>;
constexpr int BT_REACH = PATCH_CHUNK;
constexpr int BINK_THREADS = 256;

// Counting sort of a frame's key-points by bin: bin_start[f][b] .. bin_start[f][b + 1] delimit bin b's entries of
// bin_slot[f][] (the key-point's slot = its rank in the detector's output order, which the outputs keep) and bin_kp[f][]
// (its coordinates, so that the consumers need no dependent load).  The order inside a bin is whatever the atomics give:
// every key-point writes its own output slot, so the results do not depend on it.  One CTA per frame.
__global__ void __launch_bounds__(BINK_THREADS)
bin_keypoints_kernel(const ushort2* __restrict__ kp_xy, const int* __restrict__ n_kp, int max_corners, int bin_w, int bin_h,
                     int nbx, int n_bins, int* __restrict__ bin_start, uint32_t* __restrict__ bin_slot,
                     ushort2* __restrict__ bin_kp) {
    extern __shared__ __align__(16) unsigned char bk_smem[];
    __shared__ unsigned long long wsum[BINK_THREADS / 32];
    __shared__ unsigned long long total;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(bk_smem);
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = min(n_kp[f], max_corners);
    const ushort2* kps = kp_xy + (size_t)f * max_corners;
    int* bs = bin_start + (size_t)f * (n_bins + 1);
    for (int b = tid; b < n_bins; b += BINK_THREADS) cnt[b] = 0u;
    __syncthreads();
    for (int i = tid; i < n; i += BINK_THREADS) {
        const ushort2 kp = kps[i];
        SVI_CHECK(7, (kp.y / bin_h) * nbx + kp.x / bin_w < n_bins);
        atomicAdd(&cnt[(kp.y / bin_h) * nbx + kp.x / bin_w], 1u);
    }
    __syncthreads();
    {
        const int per = (n_bins + BINK_THREADS - 1) / BINK_THREADS, b0 = tid * per, b1 = min(b0 + per, n_bins);
        uint32_t sum = 0;
        for (int b = b0; b < b1; ++b) sum += cnt[b];
        uint32_t run = (uint32_t)block_exclusive_scan<BINK_THREADS>((unsigned long long)sum, wsum, &total, tid);
        for (int b = b0; b < b1; ++b) {
            const uint32_t c = cnt[b];
            bs[b] = (int)run;
            cnt[b] = run;   // becomes the fill cursor
            run += c;
        }
        if (tid == 0) bs[n_bins] = n;
    }
    __syncthreads();
    uint32_t* slots = bin_slot + (size_t)f * max_corners;
    ushort2* bkp = bin_kp + (size_t)f * max_corners;
    for (int i = tid; i < n; i += BINK_THREADS) {
        const ushort2 kp = kps[i];
        const uint32_t pos = atomicAdd(&cnt[(kp.y / bin_h) * nbx + kp.x / bin_w], 1u);
        SVI_CHECK(7, pos < (uint32_t)n);
        slots[pos] = (uint32_t)i;
        bkp[pos] = kp;
    }
}

// LEFT descriptors of a chunk, one CTA per (bin, frame): the bin's LEFT tile by one TMA load, then one warp per key-point
// (lanes = tests, eight rounds of two 16-bit shared loads and a ballot).
template <class G>
__global__ void __launch_bounds__(G::D_WARPS * 32)
describe_left_binned_kernel(const __grid_constant__ CUtensorMap map_l, FrameGeom g, const int* __restrict__ bin_start,
                            const uint32_t* __restrict__ bin_slot, const ushort2* __restrict__ bin_kp, int nbx, int n_bins,
                            int max_corners, uint8_t* __restrict__ desc_l, int cap, int out_frame0) {
    extern __shared__ __align__(128) unsigned char db_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y, b = blockIdx.x;
    const int* bs = bin_start + (size_t)f * (n_bins + 1);
    const int k0 = bs[b], k1 = bs[b + 1];
    if (k0 >= k1) return;   // CTA-uniform: an empty bin loads nothing
    constexpr int BIN_W = G::BIN_W, BIN_H = G::BIN_H, DB_COLS = G::D_COLS, DB_BYTES = G::D_BYTES, DB_SMEM = G::D_SMEM, DB_WARPS = G::D_WARPS;
    const int bx0 = (b % nbx) * BIN_W, by0 = (b / nbx) * BIN_H;
    const int tcol0 = ((bx0 - kBriefReach) >> 3) << 3;   // arithmetic shift: rounds down for the first bin's negative start
    const uint32_t bar = smem_u32(db_smem + DB_SMEM - 16);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        mbar_expect_tx(bar, DB_BYTES);
        tma_load_2d(smem_u32(db_smem), &map_l, tcol0, f * g.H + by0 - kBriefReach, bar);   // out-of-range elements arrive as zeros, and are never sampled
    }
    __syncthreads();   // the barrier is initialised before anyone polls it
    int o1[kDescWords], o2[kDescWords];   // this lane's test points relative to (row y - by0, column x - tcol0 - 24) of the tile
#pragma unroll
    for (int j = 0; j < kDescWords; ++j) {
        const char4 pt = brief_pattern(32 * j + lane);
        o1[j] = (pt.x + kBriefReach) * DB_COLS + pt.y + kBriefReach;
        o2[j] = (pt.z + kBriefReach) * DB_COLS + pt.w + kBriefReach;
    }
    const uint32_t* slots = bin_slot + (size_t)f * max_corners;
    const ushort2* bkp = bin_kp + (size_t)f * max_corners;
    int k = k0 + warp;
    if (k >= k1) return;
    ushort2 kp = bkp[k];
    uint32_t slot = slots[k];
    mbar_wait(bar, 0u);
    const uint16_t* tile = reinterpret_cast<const uint16_t*>(db_smem);
    for (; k < k1; k += DB_WARPS) {
        const ushort2 kp_c = kp;
        const uint32_t slot_c = slot;
        if (k + DB_WARPS < k1) { kp = bkp[k + DB_WARPS]; slot = slots[k + DB_WARPS]; }
        SVI_CHECK(7, kp_c.y >= by0 && kp_c.y < by0 + BIN_H && kp_c.x >= bx0 && kp_c.x < bx0 + BIN_W &&
                         kp_c.x - tcol0 - kBriefReach >= 0 && kp_c.x - tcol0 + kBriefReach < DB_COLS);
        const uint16_t* c = tile + (kp_c.y - by0) * DB_COLS + (kp_c.x - tcol0 - kBriefReach);
        uint32_t w[kDescWords];
#pragma unroll
        for (int j = 0; j < kDescWords; ++j) w[j] = __brev(__ballot_sync(0xFFFFFFFFu, c[o1[j]] < c[o2[j]]));
        SVI_CHECK(7, (int)slot_c < cap && (int)slot_c < max_corners);
        store_desc(desc_l + ((size_t)(out_frame0 + f) * cap + slot_c) * 32, w, lane);
    }
}

// K5, binned: one CTA per (bin, frame), BM_WARPS warps walking the bin's key-points over one shared RIGHT tile.
// The LEFT descriptors are already in out.desc_l (describe_left_binned_kernel).  `bad_tile` (mapped host memory, the
// ctx's overflow flag) receives 4 if a key-point's window ever leaves the tile: the host only selects this kernel for
// parameters where that cannot happen, so the flag is an internal-error report, not a data path.
template <class G>
__global__ void __launch_bounds__(G::M_WARPS * 32, G::M_CTAS)
stereo_match_binned_kernel(const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_rs, FrameGeom g,
                           TriConst tc, float size, float range, const int* __restrict__ bin_start,
                           const uint32_t* __restrict__ bin_slot, const ushort2* __restrict__ bin_kp, int nbx, int n_bins,
                           int max_corners, StereoOutDev out, int out_frame0, int* __restrict__ bad_tile) {
    extern __shared__ __align__(128) unsigned char bm_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y, b = blockIdx.x;
    const int* bs = bin_start + (size_t)f * (n_bins + 1);
    const int k0 = bs[b], k1 = bs[b + 1];
    if (k0 >= k1) return;   // CTA-uniform
    constexpr int BIN_W = G::BIN_W, BIN_H = G::BIN_H, BT_ROWS = G::ROWS, BT_COLS = G::COLS, BT_WORDS = G::WORDS, BM_WARPS = G::M_WARPS;
    constexpr int BT_COPY_BYTES = G::COPY_BYTES, BT_COPY_STRIDE = G::COPY_STRIDE;
    const int bx0 = (b % nbx) * BIN_W, by0 = (b / nbx) * BIN_H;
    const int tcol0 = ((bx0 - BT_REACH - kBriefReach) >> 3) << 3, trow0 = by0 - kBriefReach;
    const uint32_t bar = smem_u32(bm_smem + 2 * BT_COPY_STRIDE);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        mbar_expect_tx(bar, 2 * BT_COPY_BYTES);
        // A[r][k] = S[trow0 + r][tcol0 + k], B[r][k] = S[trow0 + r][tcol0 + 1 + k]
        tma_load_2d(smem_u32(bm_smem), &map_r, tcol0, f * g.H + trow0, bar);
        tma_load_2d(smem_u32(bm_smem) + BT_COPY_STRIDE, &map_rs, tcol0, f * g.H + trow0, bar);
    }
    __syncthreads();
    int k = k0 + warp;
    if (k >= k1) return;
    const uint32_t* slots = bin_slot + (size_t)f * max_corners;
    const ushort2* bkp = bin_kp + (size_t)f * max_corners;
    const uint32_t* dl = reinterpret_cast<const uint32_t*>(out.desc_l) + ((size_t)(out_frame0 + f) * out.cap) * kDescWords;
    const uint32_t* A = reinterpret_cast<const uint32_t*>(bm_smem);
    const uint32_t* B = reinterpret_cast<const uint32_t*>(bm_smem + BT_COPY_STRIDE);

    ushort2 kp_n = bkp[k];
    uint32_t slot_n = slots[k];
    uint32_t ref_raw[kDescWords];
#pragma unroll
    for (int j = 0; j < kDescWords; ++j) ref_raw[j] = __ldg(dl + (size_t)slot_n * kDescWords + j);
    mbar_wait(bar, 0u);
    for (; k < k1; k += BM_WARPS) {
        const ushort2 kp = kp_n;
        const uint32_t slot = slot_n;
        uint32_t ref[kDescWords];
#pragma unroll
        for (int j = 0; j < kDescWords; ++j) ref[j] = desc_word_to_bytes(ref_raw[j]);
        if (k + BM_WARPS < k1) {   // the next key-point's loads fly under this one's tests
            kp_n = bkp[k + BM_WARPS];
            slot_n = slots[k + BM_WARPS];
#pragma unroll
            for (int j = 0; j < kDescWords; ++j) ref_raw[j] = __ldg(dl + (size_t)slot_n * kDescWords + j);
        }
        const float x = (float)kp.x, y = (float)kp.y;
        SearchPlan p;
        plan_right(g, tc, fmaxf(0.f, (x - range) - 4.f * size), y - 4.f * size, size, x, lane, p);
        int status = p.status, dist = -1, idx = -1;
        float u = 0.f, v = 0.f;
        uint32_t wbest[kDescWords];
        if (status == SVI_OK) {
            const int rel = (p.gx_lo - kBriefReach) - tcol0, rrow = (p.gy - kBriefReach) - trow0;
            // the window of this search inside the tile: rows rrow .. rrow + 48, elements rel .. rel + n_valid + 48
            if (rel < 0 || rrow < 0 || rrow + PATCH_ROWS > BT_ROWS || p.n_valid > PATCH_CHUNK || rel + p.n_valid + 2 * kBriefReach >= BT_COLS) {
                if (lane == 0) *reinterpret_cast<volatile int*>(bad_tile) = 4;
                status = SVI_TRI_BAD_ROI;
            } else {
                const int par = rel & 1, woff = rel >> 1;   // slot s evaluates candidate s - par
                const int l_lo = 2 * lane - par, l_hi = l_lo + 1;
                // a lane without a live candidate reads what lane 0 reads: its words may lie beyond the tile
                const int lo = (l_lo < p.n_valid) ? lane : 0;
                const uint32_t* Al = A + rrow * BT_WORDS + woff + lo;
                const uint32_t* Bl = B + rrow * BT_WORDS + woff + lo;
                uint32_t wlo[kDescWords], whi[kDescWords];
                brief_pair_all_pw<BT_WORDS>(Al, Bl, wlo, whi, std::make_integer_sequence<int, kDescWords>{});
                const uint32_t k_lo = (l_lo >= 0 && l_lo < p.n_valid) ? (((uint32_t)hamming_words(wlo, ref) << 16) | (uint32_t)l_lo) : 0xFFFFFFFFu;
                const uint32_t k_hi = (l_hi < p.n_valid) ? (((uint32_t)hamming_words(whi, ref) << 16) | (uint32_t)l_hi) : 0xFFFFFFFFu;
                const uint32_t k_mine = min(k_lo, k_hi);
                const uint32_t k_min = warp_min_u32(k_mine);   // n_valid >= 1: some lane holds a live key
                const int owner = __ffs(__ballot_sync(0xFFFFFFFFu, k_mine == k_min)) - 1;
                const bool hi = (k_hi == k_min) && (k_lo != k_min);
#pragma unroll
                for (int j = 0; j < kDescWords; ++j) wbest[j] = __shfl_sync(0xFFFFFFFFu, hi ? whi[j] : wlo[j], owner);
                dist = (int)(k_min >> 16);
                idx = (int)(k_min & 0xFFFFu);
                if (!(tc.match_cutoff > (float)dist)) status = SVI_TRI_DISTANCE;
                else {
                    const float px = (p.border + (float)(p.i_lo + idx)) + (float)p.first;
                    u = px + p.u_tl;
                    v = p.border + p.v_tl;
                }
            }
        }
        double xyz[3] = {0.0, 0.0, 0.0};
        if (status == SVI_OK) status = point_in_left(tc, x, y, u, xyz);
        SVI_CHECK(7, (int)slot < out.cap && (int)slot < max_corners);
        const size_t o = (size_t)(out_frame0 + f) * out.cap + slot;
        if (status == SVI_OK) store_desc(out.desc_r + o * 32, wbest, lane);
        if (lane == 0) {
            out.uv_l[o * 2] = x; out.uv_l[o * 2 + 1] = y;
            out.status[o] = (uint8_t)status;
            out.dist[o] = dist;
            out.idx[o] = idx;
            if (status == SVI_OK) {
                out.uv_r[o * 2] = u; out.uv_r[o * 2 + 1] = v;
                out.xyz[o * 3] = xyz[0]; out.xyz[o * 3 + 1] = xyz[1]; out.xyz[o * 3 + 2] = xyz[2];
            }
        }
    }
}

}  // namespace svi
