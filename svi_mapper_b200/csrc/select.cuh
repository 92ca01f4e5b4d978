// select.cuh -- corner selection of cv::goodFeaturesToTrack after candidate collection:
// sort by (response desc, address desc), greedy minimum-distance filter, cut at maxCorners,
// then BRIEF's 28-px image-border filter (KeyPointsFilter::runByImageBorder) that
// BriefDescriptorExtractor::compute applies in place (reference
// src/core/CFundamentalMatcher.cpp:101,106).  One CTA per frame.
//
// The sequential greedy loop is evaluated in its exact parallel form:
// with candidates in priority order and hp(i) = higher-priority candidates closer than
// minDistance, i is REJECTED once some member of hp(i) is ACCEPTED and ACCEPTED once all of
// hp(i) are REJECTED; iterate to the fixed point.  Decisions are monotone, so reading a
// neighbour's state while another thread updates it is benign.
//
// Priority is the 64-bit key itself, so the peeling needs no sorted input: on the shared-memory path the
// candidates are first counting-sorted by grid cell (so the 3x3 cell neighbourhood of a candidate is three
// contiguous runs of keys, no pointer chasing), peeled, and only the ACCEPTED corners (typically under half
// of the candidates) are compacted and sorted: the bitonic network shrinks with them.
#pragma once
#include "common.cuh"

namespace svi {

constexpr int SEL_THREADS = 1024;
constexpr int SEL_SMEM_KEYS = 16384;   // keys that fit the shared-memory fast path
constexpr int SEL_SMEM_CELLS = 8192;
constexpr int SEL_SMALL_THREADS = 256, SEL_SMALL_KEYS = 2048, SEL_SMALL_CELLS = 2048;   // window-mode configuration
__host__ __device__ constexpr int select_smem_bytes(int keys, int cells) { return 11 * keys + 4 * cells; }
__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }
constexpr uint32_t SEL_NIL = 0xFFFFFFFFu;
#ifndef SEL_CUT_TENTHS
#define SEL_CUT_TENTHS 26   // the priority cut keeps about 2.6 x maxCorners candidates
#endif

struct SelectParams {
    int W, H;
    int raw_cap;         // per-frame stride and capacity of `cand`: the local maxima the Harris kernel emits
    int cand_cap;        // capacity for the candidates that pass the frame's quality threshold (power of two)
    double quality;      // qualityLevel of goodFeaturesToTrack
    int max_corners;
    int cell;            // grid cell edge >= ceil(minDistance)
    int gw, gh;          // grid size
    double min_dist_sq;  // minDistance^2 (cv compares dx*dx+dy*dy < minDistance*minDistance)
    int min_dist_sq_ceil; // ceil(minDistance^2): for integer d2, d2 < minDistance^2  <=>  d2 < this
    uint32_t cell_magic; // ceil(2^24 / cell) when cell < 256 (exact x / cell for x < 65536 by one multiply-high), else 0
    int filter;          // 0 when minDistance < 1 (cv skips the distance filter)
    int cap_is_error;    // FAST mode: OpenCV returns every corner, so exceeding max_corners must be reported
};

__device__ __forceinline__ int cell_of(int v, const SelectParams& sp) {
    return sp.cell_magic ? (int)__umulhi((uint32_t)v << 8, sp.cell_magic) : v / sp.cell;
}

__device__ __forceinline__ void key_xy(unsigned long long k, int& x, int& y) {
    x = (int)(k & 0xFFFFu);
    y = (int)((k >> 16) & 0xFFFFu);
}

// In-place bitonic sort, descending, of n_pad (power of two) keys by the whole CTA.  Thread t handles the pair
// (i, i + j) with i = 2t - (t & (j - 1)): for j <= 32 the 32 pairs of a warp lie inside one aligned block of 64 keys that
// no other warp touches, so those steps (six of every k-level, all of the levels k <= 64) only need __syncwarp();
// the CTA-wide barrier is kept for the steps that exchange keys between warps.
template <int T>
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int n_pad, int tid) {
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (n_pad >> 1); t += T) {
                int i = 2 * t - (t & (j - 1));
                int l = i + j;
                unsigned long long a = keys[i], b = keys[l];
                bool desc = ((i & k) == 0);
                if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[l] = a; }
            }
            if (j > 32) __syncthreads(); else __syncwarp();
        }
        if (k >= 64) __syncthreads();   // the next level starts with j = k >= 64: keys cross warps again
    }
    __syncthreads();
}

// Exclusive scan over the CTA of a packed pair of 32-bit counters; *total receives the CTA sum.
template <int T>
__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* wsum,
                                                                   unsigned long long* total, int tid) {
    unsigned long long incl = v;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += u;
    }
    __syncthreads();   // wsum may still be read from an earlier scan
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        unsigned long long w = lane < T / 32 ? wsum[lane] : 0ull, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= o) wi += u;
        }
        if (lane < T / 32) wsum[lane] = wi - w;  // exclusive
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    return wsum[wid] + incl - v;
}

// kSmem: keys/lists in shared memory (n <= SEL_SMEM_KEYS, cells <= SEL_SMEM_CELLS); otherwise
// they live in the per-frame global scratch (stress-size frames).
// The big configuration (1024 threads, 16384 keys, 8192 cells, 208 KB) owns an SM: right for whole frames.  The
// tracker's stage 2 runs the detector on thousands of small windows per frame; for those the small configuration
// (256 threads, 2048 keys, 2048 cells, 30 KB) keeps seven CTAs on an SM.  A window that does not fit the small
// configuration sets defer[f] and is picked up by a second launch of the big one (run_if).
template <bool kSmem, int SEL_THREADS = 1024, int SEL_SMEM_KEYS = 16384, int SEL_SMEM_CELLS = 8192>
__global__ void __launch_bounds__(SEL_THREADS, 1)
select_corners_kernel(unsigned long long* __restrict__ cand, const int* __restrict__ cand_count,
                      const uint32_t* __restrict__ frame_max,
                      SelectParams sp, uint32_t* __restrict__ g_head, uint32_t* __restrict__ g_next,
                      uint8_t* __restrict__ g_state, ushort2* __restrict__ det_xy,
                      int* __restrict__ n_detected, ushort2* __restrict__ kp_xy,
                      int* __restrict__ n_keypoints, int* __restrict__ overflow,
                      const RoiItem* __restrict__ rois, int* __restrict__ defer = nullptr,
                      const int* __restrict__ run_if = nullptr, const int* __restrict__ n_items = nullptr) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    __shared__ unsigned long long wsum[SEL_THREADS / 32];
    __shared__ unsigned long long scan_total;
    __shared__ uint32_t s_fill, s_min_bucket;
    const int f = blockIdx.x, tid = threadIdx.x;
    if (n_items && f >= *n_items) return;   // window mode with a device-side item count
    if (rois) {   // window mode: the bucket grid and the border filter follow the item's rectangle
        sp.W = rois[f].rw;
        sp.H = rois[f].rh;
        sp.gw = (sp.W + sp.cell - 1) / sp.cell;
        sp.gh = (sp.H + sp.cell - 1) / sp.cell;
    }
    if (run_if && !run_if[f]) return;
    // The list holds the local maxima above the threshold of their TILE's maximum; goodFeaturesToTrack keeps the
    // ones above the threshold of the FRAME's maximum (final by now): R > thr  <=>  ordered(R) > ordered(thr).
    // frame_max == nullptr (FAST mode): every key is a corner.
    int n_raw = cand_count[f];
    if (n_raw > sp.raw_cap) {
        if (tid == 0) *reinterpret_cast<volatile int*>(overflow) = 1;   // mapped host memory: a plain store
        n_raw = sp.raw_cap;
    }
    unsigned long long* gk = cand + (size_t)f * sp.raw_cap;
    const uint32_t thr_ord = frame_max ? float_to_ordered(gftt_threshold(frame_max[f], sp.quality)) : 0u;
    auto live = [thr_ord](unsigned long long k) { return (uint32_t)(k >> 32) > thr_ord; };
    // The cell-sorted path walks the raw list four times (count, priority histogram, cell histogram, scatter): stage it
    // once at the upper end of the shared key array -- every later pass then costs shared-memory, not L2, latency.
    const bool staged_ok = kSmem && sp.filter && n_raw <= SEL_SMEM_KEYS;
    const unsigned long long* raw = gk;
    int n;
    {
        int c = 0;
        if (staged_ok) {
            unsigned long long* raw_s = reinterpret_cast<unsigned long long*>(sel_smem) + (SEL_SMEM_KEYS - n_raw);
            for (int i = tid; i < n_raw; i += SEL_THREADS) {
                const unsigned long long k = gk[i];
                raw_s[i] = k;
                c += live(k) ? 1 : 0;
            }
            raw = raw_s;
        } else {
            for (int i = tid; i < n_raw; i += SEL_THREADS) c += live(gk[i]) ? 1 : 0;
        }
        block_exclusive_scan<SEL_THREADS>((unsigned long long)c, wsum, &scan_total, tid);   // its barriers publish the staged list
        n = (int)scan_total;
    }
    if (defer && (n > SEL_SMEM_KEYS || sp.gw * sp.gh > SEL_SMEM_CELLS)) {   // CTA-uniform
        if (tid == 0) defer[f] = 1;
        return;
    }
    if (n > sp.cand_cap || (kSmem && (n > SEL_SMEM_KEYS || sp.gw * sp.gh > SEL_SMEM_CELLS))) {
        // more candidates than the ctx was sized for: reported as SVI_ERR_CAPACITY by the host, never a silent cut
        if (tid == 0) { *reinterpret_cast<volatile int*>(overflow) = n > sp.cand_cap ? 1 : 2; n_detected[f] = 0; n_keypoints[f] = 0; }
        return;
    }
    if (sp.cap_is_error && n > sp.max_corners && tid == 0) *reinterpret_cast<volatile int*>(overflow) = 3;
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    int n_act = n;   // candidates that take part in the peeling (shared-memory path: the top-priority part)
    const int ncells = sp.gw * sp.gh;

    // shared layout: keys[16384] u64 | next16[16384] u16 (cell-sorted path: cell_start u16) | state[16384] u8 |
    //                head[8192] u32 (cell-sorted path: per-cell counters / cursors)
    unsigned long long* keys = kSmem ? reinterpret_cast<unsigned long long*>(sel_smem) : gk;
    uint16_t* next16 = reinterpret_cast<uint16_t*>(sel_smem + sizeof(unsigned long long) * SEL_SMEM_KEYS);
    uint8_t* state = kSmem ? reinterpret_cast<uint8_t*>(next16 + SEL_SMEM_KEYS) : g_state + (size_t)f * sp.cand_cap;
    uint32_t* head = kSmem ? reinterpret_cast<uint32_t*>(sel_smem + 11 * SEL_SMEM_KEYS)   // = select_smem_bytes() - 4 * cells
                           : g_head + (size_t)f * ncells;
    uint32_t* next = kSmem ? nullptr : g_next + (size_t)f * sp.cand_cap;
    const bool peel_first = kSmem && sp.filter;
    constexpr int kBucketShift = 64 - ilog2(SEL_SMEM_CELLS);   // top bits of the key: sign, exponent, leading mantissa bits
    const int target = (int)min((long long)n, (long long)sp.max_corners * SEL_CUT_TENTHS / 10 + 256);
    if (peel_first) {
        // ---- counting sort of the candidates by grid cell, straight from the global list into `keys`
        uint16_t* cell_start = next16;            // ncells + 1 <= 8193 entries
        uint32_t* cursor = head;
        // ---- priority cut.  Whether a candidate is accepted depends only on candidates of HIGHER priority, and the
        // output stops at max_corners accepted corners: peeling the top-priority part of the list gives the exact
        // head of the accepted sequence.  A 13-bit bucket of the key (sign, exponent, 4 mantissa bits of the response)
        // is monotone in priority; take the buckets that hold about 2.6 x max_corners candidates, and fall back to
        // the whole list in the rare case that they yield fewer than max_corners corners.
        if (tid == 0) s_min_bucket = 0u;
        if (n > target) {
            constexpr int NB = SEL_SMEM_CELLS;   // one bucket per entry of the cursor area
            for (int b = tid; b < NB; b += SEL_THREADS) cursor[b] = 0u;
            __syncthreads();
            for (int i = tid; i < n_raw; i += SEL_THREADS) {
                const unsigned long long k = raw[i];
                SVI_CHECK(2, (uint32_t)(k >> kBucketShift) < (uint32_t)SEL_SMEM_CELLS);
                if (live(k)) atomicAdd(&cursor[(uint32_t)(k >> kBucketShift)], 1u);
            }
            __syncthreads();
            constexpr int BPT = NB / SEL_THREADS;   // thread t owns buckets NB-1-BPT*t ... downwards
            uint32_t sum = 0;
#pragma unroll
            for (int k = 0; k < BPT; ++k) sum += cursor[NB - 1 - (tid * BPT + k)];
            uint32_t above = (uint32_t)block_exclusive_scan<SEL_THREADS>((unsigned long long)sum, wsum, &scan_total, tid);
#pragma unroll
            for (int k = 0; k < BPT; ++k) {
                const int b = NB - 1 - (tid * BPT + k);
                const uint32_t c = cursor[b];
                if (above < (uint32_t)target && above + c >= (uint32_t)target) s_min_bucket = (uint32_t)b;
                above += c;
            }
        }
        __syncthreads();
        for (;;) {   // at most two attempts: cut list, then (rarely) the whole list
        const uint32_t min_bucket = s_min_bucket;
        for (int c = tid; c < ncells; c += SEL_THREADS) cursor[c] = 0u;
        __syncthreads();
        for (int i = tid; i < n_raw; i += SEL_THREADS) {
            const unsigned long long k = raw[i];
            if ((uint32_t)(k >> kBucketShift) < min_bucket || !live(k)) continue;
            int x, y;
            key_xy(k, x, y);
            SVI_CHECK(2, (unsigned)(cell_of(y, sp) * sp.gw + cell_of(x, sp)) < (unsigned)ncells && ncells <= SEL_SMEM_CELLS);
            atomicAdd(&cursor[cell_of(y, sp) * sp.gw + cell_of(x, sp)], 1u);
        }
        __syncthreads();
        {
            const int cper = (ncells + SEL_THREADS - 1) / SEL_THREADS, c0 = tid * cper, c1 = min(c0 + cper, ncells);
            uint32_t sum = 0;
            for (int c = c0; c < c1; ++c) sum += cursor[c];
            uint32_t run = (uint32_t)block_exclusive_scan<SEL_THREADS>((unsigned long long)sum, wsum, &scan_total, tid);
            for (int c = c0; c < c1; ++c) {
                const uint32_t cnt = cursor[c];
                cell_start[c] = (uint16_t)run;
                cursor[c] = run;
                run += cnt;
            }
            n_act = (int)scan_total;                           // candidates taking part in this attempt
            if (tid == 0) cell_start[ncells] = (uint16_t)n_act;   // n_act <= 16384
        }
        __syncthreads();
        // the scatter writes keys[0 .. n_act): it may read the staged list only while the two do not overlap
        const unsigned long long* src = (raw != gk && n_act + n_raw <= SEL_SMEM_KEYS) ? raw : gk;
        for (int i = tid; i < n_raw; i += SEL_THREADS) {
            const unsigned long long k = src[i];
            if ((uint32_t)(k >> kBucketShift) < min_bucket || !live(k)) continue;
            int x, y;
            key_xy(k, x, y);
            const uint32_t pos = atomicAdd(&cursor[cell_of(y, sp) * sp.gw + cell_of(x, sp)], 1u);
            SVI_CHECK(2, pos < (uint32_t)n_act && n_act <= SEL_SMEM_KEYS);
            keys[pos] = k;
            state[pos] = 0;
        }
        __syncthreads();
        // ---- peel to the fixed point; neighbours = three contiguous key runs.
        // A candidate's decision depends only on its BLOCKERS (higher-priority candidates closer than minDistance,
        // rarely more than two): the first sweep scans the 3x3 cell neighbourhood once and remembers up to four of
        // them (8 B per candidate in the free upper half of the key array); every later sweep only re-reads their
        // states.  Candidates with more blockers (marker 0xFFFE) and lists too long for the cache rescan as before.
        auto scan_neighbours = [&](int i, unsigned long long ki, bool& any_acc, bool& any_und) {
            int x, y;
            key_xy(ki, x, y);
            const int cx = cell_of(x, sp), cy = cell_of(y, sp);
            const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, sp.gw - 1);
            for (int yy = max(cy - 1, 0); yy <= min(cy + 1, sp.gh - 1) && !any_acc; ++yy) {
                const int j1 = cell_start[yy * sp.gw + x_hi + 1];
                SVI_CHECK(2, yy * sp.gw + x_hi + 1 <= ncells && j1 <= n_act);
                for (int j = cell_start[yy * sp.gw + x_lo]; j < j1; ++j) {
                    const unsigned long long kj = keys[j];
                    if (kj > ki) {   // higher priority: larger response, then larger address (keys are unique)
                        int px, py;
                        key_xy(kj, px, py);
                        const int dx = x - px, dy = y - py;
                        if (dx * dx + dy * dy < sp.min_dist_sq_ceil) {
                            const uint8_t s = state[j];
                            any_acc |= (s == 1);
                            any_und |= (s == 0);
                        }
                    }
                }
            }
        };
        const bool cache_ok = n_act <= SEL_SMEM_KEYS / 2;
        ushort4* blockers = reinterpret_cast<ushort4*>(keys + SEL_SMEM_KEYS / 2);
        if (cache_ok) {
            for (int i = tid; i < n_act; i += SEL_THREADS) {
                int x, y;
                const unsigned long long ki = keys[i];
                key_xy(ki, x, y);
                const int cx = cell_of(x, sp), cy = cell_of(y, sp);
                const int x_lo = max(cx - 1, 0), x_hi = min(cx + 1, sp.gw - 1);
                unsigned short b[4] = {0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu};
                int nb = 0;
                for (int yy = max(cy - 1, 0); yy <= min(cy + 1, sp.gh - 1); ++yy) {
                    const int j1 = cell_start[yy * sp.gw + x_hi + 1];
                    SVI_CHECK(2, yy * sp.gw + x_hi + 1 <= ncells && j1 <= n_act);
                    for (int j = cell_start[yy * sp.gw + x_lo]; j < j1; ++j) {
                        const unsigned long long kj = keys[j];
                        if (kj > ki) {
                            int px, py;
                            key_xy(kj, px, py);
                            const int dx = x - px, dy = y - py;
                            if (dx * dx + dy * dy < sp.min_dist_sq_ceil) {
                                if (nb == 0) b[0] = (unsigned short)j;
                                else if (nb == 1) b[1] = (unsigned short)j;
                                else if (nb == 2) b[2] = (unsigned short)j;
                                else if (nb == 3) b[3] = (unsigned short)j;
                                ++nb;
                            }
                        }
                    }
                }
                if (nb > 4) b[0] = 0xFFFEu;
                blockers[i] = make_ushort4(b[0], b[1], b[2], b[3]);
                if (nb == 0) state[i] = 1;   // nothing of higher priority nearby: accepted at once
            }
            __syncthreads();
        }
        for (;;) {
            int changed = 0;
            for (int i = tid; i < n_act; i += SEL_THREADS) {
                if (state[i] != 0) continue;
                bool any_acc = false, any_und = false;
                const ushort4 b = cache_ok ? blockers[i] : make_ushort4(0xFFFEu, 0, 0, 0);
                if (b.x == 0xFFFEu) {
                    scan_neighbours(i, keys[i], any_acc, any_und);
                } else {
                    const unsigned short bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (bb[q] != 0xFFFFu) {
                            const uint8_t s = state[bb[q]];
                            any_acc |= (s == 1);
                            any_und |= (s == 0);
                        }
                }
                if (any_acc) { state[i] = 2; changed = 1; }
                else if (!any_und) { state[i] = 1; changed = 1; }
            }
            if (!__syncthreads_or(changed)) break;
        }
        if (min_bucket == 0u) break;   // the whole list was peeled
        int acc = 0;
        for (int i = tid; i < n_act; i += SEL_THREADS) acc += (state[i] == 1);
        block_exclusive_scan<SEL_THREADS>((unsigned long long)acc, wsum, &scan_total, tid);
        if ((int)scan_total >= sp.max_corners) break;   // enough corners from the cut list: exact
        if (tid == 0) s_min_bucket = 0u;
        raw = gk;                                      // the staged copy may have been overwritten by the scatter
        __syncthreads();
        }
    }

    // ---- (sorted path: FAST / no distance filter in shared memory, or the global variant for very large frames)
    //      fill the key array with the live candidates, sort, cell lists linked through `next`, peel.
    //      The global variant applies the same priority cut as the cell-sorted path (the cut list is compacted into the
    //      free upper half of the frame's raw list, so a 3840x1080 frame sorts ~26 k keys instead of ~130 k), with the
    //      same exact fallback to the whole list.
    if (!peel_first) {
        const bool can_cut = !kSmem && sp.filter && n_raw <= sp.raw_cap / 2 && n > target;
        if (tid == 0) s_min_bucket = 0u;
        if (can_cut) {   // histogram of the 13-bit priority buckets in (dynamic) shared memory
            uint32_t* hist = reinterpret_cast<uint32_t*>(sel_smem);
            constexpr int NB = SEL_SMEM_CELLS;
            for (int b = tid; b < NB; b += SEL_THREADS) hist[b] = 0u;
            __syncthreads();
            for (int i = tid; i < n_raw; i += SEL_THREADS) {
                const unsigned long long k = gk[i];
                if (live(k)) atomicAdd(&hist[(uint32_t)(k >> kBucketShift)], 1u);
            }
            __syncthreads();
            constexpr int BPT = NB / SEL_THREADS;
            uint32_t sum = 0;
#pragma unroll
            for (int k = 0; k < BPT; ++k) sum += hist[NB - 1 - (tid * BPT + k)];
            uint32_t above = (uint32_t)block_exclusive_scan<SEL_THREADS>((unsigned long long)sum, wsum, &scan_total, tid);
#pragma unroll
            for (int k = 0; k < BPT; ++k) {
                const int b = NB - 1 - (tid * BPT + k);
                const uint32_t c = hist[b];
                if (above < (uint32_t)target && above + c >= (uint32_t)target) s_min_bucket = (uint32_t)b;
                above += c;
            }
        }
        __syncthreads();
        const int n_live = n;
        for (;;) {   // at most two attempts: cut list, then (rarely) the whole list
            const uint32_t min_bucket = s_min_bucket;
            if (kSmem || n_raw <= sp.raw_cap / 2) {
                // compact the (cut) live keys -- into shared memory, or into the free upper half of the raw list
                if (!kSmem) keys = gk + sp.raw_cap / 2;
                if (tid == 0) s_fill = 0u;
                __syncthreads();
                for (int i = tid; i < n_raw; i += SEL_THREADS) {
                    const unsigned long long k = gk[i];
                    if (live(k) && (uint32_t)(k >> kBucketShift) >= min_bucket) {
                        const uint32_t pos = atomicAdd(&s_fill, 1u);
                        SVI_CHECK(2, kSmem ? pos < (uint32_t)SEL_SMEM_KEYS : pos < (uint32_t)(sp.raw_cap / 2));
                        keys[pos] = k;
                    }
                }
                __syncthreads();
                n = (int)s_fill;
                n_pad = 1;
                while (n_pad < n) n_pad <<= 1;
                for (int i = n + tid; i < n_pad; i += SEL_THREADS) keys[i] = 0ull;
            } else {
                // no room to compact: sort the raw list in place with the dead keys zeroed (they sink to the end)
                n_pad = 1;
                while (n_pad < n_raw) n_pad <<= 1;
                for (int i = tid; i < n_pad; i += SEL_THREADS)
                    if (i >= n_raw || !live(keys[i])) keys[i] = 0ull;
                n = n_live;
            }
            __syncthreads();
            bitonic_sort_desc<SEL_THREADS>(keys, n_pad, tid);
            for (int c = tid; c < ncells; c += SEL_THREADS) head[c] = SEL_NIL;
            for (int i = tid; i < n; i += SEL_THREADS) state[i] = sp.filter ? 0 : 1;
            __syncthreads();
            if (!sp.filter) break;   // every candidate is a corner
            for (int i = tid; i < n; i += SEL_THREADS) {
                int x, y;
                key_xy(keys[i], x, y);
                int c = cell_of(y, sp) * sp.gw + cell_of(x, sp);
                uint32_t prev = atomicExch(&head[c], (uint32_t)i);
                if (kSmem) next16[i] = (uint16_t)(prev == SEL_NIL ? 0xFFFFu : prev);
                else next[i] = prev;
            }
            __syncthreads();
            // ---- peel to the fixed point
            for (;;) {
                int changed = 0;
                for (int i = tid; i < n; i += SEL_THREADS) {
                    if (state[i] != 0) continue;
                    int x, y;
                    const unsigned long long ki = keys[i];
                    key_xy(ki, x, y);
                    const int cx = cell_of(x, sp), cy = cell_of(y, sp);
                    bool any_acc = false, any_und = false;
                    for (int yy = max(cy - 1, 0); yy <= min(cy + 1, sp.gh - 1); ++yy)
                        for (int xx = max(cx - 1, 0); xx <= min(cx + 1, sp.gw - 1); ++xx) {
                            uint32_t j = head[yy * sp.gw + xx];
                            while (j != SEL_NIL) {
                                const unsigned long long kj = keys[j];
                                if (kj > ki) {   // higher priority: larger response, then larger address (keys are unique)
                                    int px, py;
                                    key_xy(kj, px, py);
                                    int dx = x - px, dy = y - py;
                                    if (dx * dx + dy * dy < sp.min_dist_sq_ceil) {
                                        uint8_t s = state[j];
                                        any_acc |= (s == 1);
                                        any_und |= (s == 0);
                                    }
                                }
                                if (kSmem) { uint16_t nx = next16[j]; j = (nx == 0xFFFFu) ? SEL_NIL : nx; }
                                else j = next[j];
                            }
                        }
                    if (any_acc) { state[i] = 2; changed = 1; }
                    else if (!any_und) { state[i] = 1; changed = 1; }
                }
                if (!__syncthreads_or(changed)) break;
            }
            if (min_bucket == 0u) break;   // the whole list was peeled
            int acc = 0;
            for (int i = tid; i < n; i += SEL_THREADS) acc += (state[i] == 1);
            block_exclusive_scan<SEL_THREADS>((unsigned long long)acc, wsum, &scan_total, tid);
            if ((int)scan_total >= sp.max_corners) break;   // enough corners from the cut list: exact
            if (tid == 0) s_min_bucket = 0u;
            __syncthreads();
        }
    }

    if (peel_first && n_act <= SEL_SMEM_KEYS / 2) {
        // ---- order the accepted corners by priority without a sorting network: counting sort by a monotone bucket of
        // the key (sign, exponent, leading mantissa bits) into the free upper half of the key array, then every key
        // finds its place inside its bucket by counting the bucket's greater keys (a bucket holds a handful of keys).
        constexpr int NB2 = SEL_SMEM_CELLS < SEL_SMEM_KEYS / 2 ? SEL_SMEM_CELLS : SEL_SMEM_KEYS / 2;
        constexpr int kShift2 = 64 - ilog2(NB2);
        constexpr int BPT2 = NB2 / SEL_THREADS;
        static_assert(NB2 % SEL_THREADS == 0 && NB2 * 4 <= 2 * SEL_SMEM_KEYS, "bucket tables fit the cursor / cell_start areas");
        uint32_t* bcnt = head;                                    // counts, then fill cursors
        uint32_t* bstart = reinterpret_cast<uint32_t*>(next16);   // the cell_start table is dead after the peeling
        unsigned long long* tmp = keys + SEL_SMEM_KEYS / 2;       // so is the blocker cache
        for (int b = tid; b < NB2; b += SEL_THREADS) bcnt[b] = 0u;
        __syncthreads();
        for (int i = tid; i < n_act; i += SEL_THREADS)
            if (state[i] == 1) atomicAdd(&bcnt[(uint32_t)(keys[i] >> kShift2)], 1u);
        __syncthreads();
        {
            uint32_t sum = 0;
#pragma unroll
            for (int k = 0; k < BPT2; ++k) sum += bcnt[NB2 - 1 - (tid * BPT2 + k)];
            uint32_t above = (uint32_t)block_exclusive_scan<SEL_THREADS>((unsigned long long)sum, wsum, &scan_total, tid);
#pragma unroll
            for (int k = 0; k < BPT2; ++k) {
                const int b = NB2 - 1 - (tid * BPT2 + k);
                const uint32_t c = bcnt[b];
                bstart[b] = above;
                bcnt[b] = 0u;
                above += c;
            }
        }
        n = (int)scan_total;
        __syncthreads();
        for (int i = tid; i < n_act; i += SEL_THREADS)
            if (state[i] == 1) {
                const unsigned long long k = keys[i];
                const uint32_t b = (uint32_t)(k >> kShift2);
                const uint32_t pos = bstart[b] + atomicAdd(&bcnt[b], 1u);
                SVI_CHECK(2, pos < (uint32_t)n && n <= SEL_SMEM_KEYS / 2);
                tmp[pos] = k;
            }
        __syncthreads();
        for (int q = tid; q < n; q += SEL_THREADS) {
            const unsigned long long k = tmp[q];
            const uint32_t b = (uint32_t)(k >> kShift2);
            const uint32_t s0 = bstart[b], s1 = s0 + bcnt[b];
            uint32_t r = s0;
            for (uint32_t j = s0; j < s1; ++j) r += tmp[j] > k ? 1u : 0u;
            SVI_CHECK(2, r < (uint32_t)n);
            keys[r] = k;
            state[r] = 1;
        }
        __syncthreads();
    } else if (peel_first) {
        // ---- (lists longer than half the key array) compact the accepted corners (each thread parks its <= 16 keys in
        //      registers), sort only those
        constexpr int PER = SEL_SMEM_KEYS / SEL_THREADS;
        unsigned long long mine[PER];
        int cnt = 0;
        const int b0 = tid * PER;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = b0 + k;
            const bool acc = i < n_act && state[i] == 1;
            mine[k] = acc ? keys[i] : 0ull;     // accepted keys are never 0 (the ordered response of a corner is > 0)
            cnt += acc;
        }
        unsigned long long tot = 0;
        int rank = (int)block_exclusive_scan<SEL_THREADS>((unsigned long long)cnt, wsum, &scan_total, tid);   // ends with a barrier: all reads done
        tot = scan_total;
        n = (int)tot;
        n_pad = 1;
        while (n_pad < n) n_pad <<= 1;
#pragma unroll
        for (int k = 0; k < PER; ++k)
            if (mine[k] != 0ull) {
                SVI_CHECK(2, rank < SEL_SMEM_KEYS);
                keys[rank++] = mine[k];
            }
        for (int i = n + tid; i < n_pad; i += SEL_THREADS) keys[i] = 0ull;
        for (int i = tid; i < n; i += SEL_THREADS) state[i] = 1;
        __syncthreads();
        bitonic_sort_desc<SEL_THREADS>(keys, n_pad, tid);
    }

    // ---- ranks among accepted corners and among those passing BRIEF's border filter
    const int per = (n + SEL_THREADS - 1) / SEL_THREADS;
    const int i0 = tid * per, i1 = min(i0 + per, n);
    uint32_t acc_cnt = 0, kp_cnt = 0;
    for (int i = i0; i < i1; ++i) {
        if (state[i] == 1) {
            int x, y;
            key_xy(keys[i], x, y);
            ++acc_cnt;
            kp_cnt += (x >= kBriefBorder && x < sp.W - kBriefBorder && y >= kBriefBorder && y < sp.H - kBriefBorder);
        }
    }
    // block exclusive scan of (acc_cnt, kp_cnt) packed as 2 x 32 bit
    const unsigned long long excl = block_exclusive_scan<SEL_THREADS>(((unsigned long long)kp_cnt << 32) | acc_cnt, wsum, &scan_total, tid);
    uint32_t acc_rank = (uint32_t)(excl & 0xFFFFFFFFu), kp_rank = (uint32_t)(excl >> 32);
    ushort2* det = det_xy + (size_t)f * sp.max_corners;
    ushort2* kp = kp_xy + (size_t)f * sp.max_corners;
    for (int i = i0; i < i1; ++i) {
        if (state[i] == 1) {
            if ((int)acc_rank >= sp.max_corners) break;
            int x, y;
            key_xy(keys[i], x, y);
            SVI_CHECK(2, (int)acc_rank < sp.max_corners && (int)kp_rank < sp.max_corners);
            det[acc_rank] = make_ushort2((unsigned short)x, (unsigned short)y);
            ++acc_rank;
            if (x >= kBriefBorder && x < sp.W - kBriefBorder && y >= kBriefBorder && y < sp.H - kBriefBorder) {
                kp[kp_rank] = make_ushort2((unsigned short)x, (unsigned short)y);
                ++kp_rank;
            }
        }
    }
    // the key-point count is the kp_rank reached by the element with acc_rank == max_corners-1
    // (or the last accepted one): the thread that writes the last detected corner publishes it.
    {
        uint32_t a0 = (uint32_t)(excl & 0xFFFFFFFFu);
        uint32_t total = (uint32_t)(scan_total & 0xFFFFFFFFu);
        uint32_t limit = min(total, (uint32_t)sp.max_corners);
        if (limit == 0) {
            if (tid == 0) { n_detected[f] = 0; n_keypoints[f] = 0; }
        } else if (a0 < limit && a0 + acc_cnt >= limit && acc_cnt > 0) {
            n_detected[f] = (int)limit;
            n_keypoints[f] = (int)kp_rank;  // kp_rank after this thread's emission loop
        }
    }
}

}  // namespace svi
