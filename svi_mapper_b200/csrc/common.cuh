// common.cuh -- shared device helpers for libsvi_gpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svi {

constexpr int kBriefBorder = 28;  // PATCH_SIZE/2 + KERNEL_SIZE/2 of OpenCV BRIEF
constexpr int kBriefReach = 24;   // max |offset| of a test point
constexpr int kDescWords = 8;     // 256 bit

// Geometry + constants every kernel needs; lives in kernel parameter space.
struct FrameGeom {
    int W, H;           // image size
    int img_pitch;      // bytes per image row (caller's layout)
    size_t img_stride;  // bytes per frame
    int resp_pitch;     // floats per response row
    int box_pitch;      // uint16 per box-sum row
};

// One detection work item when the detector runs on windows of an image instead of whole frames
// (trackManual stage 2: GFTTDetector::detect(img(cSearchROI)), CFundamentalMatcher.cpp:1566,1690).
// plane selects the image of the staged pair (0 = LEFT, 1 = RIGHT); the rectangle is in image pixels.
struct RoiItem {
    int plane, rx, ry, rw, rh;
};

// CTriangulator members (src/core/CTriangulator.cpp:13-21)
struct TriConst {
    double f_inv, pu, pv, du_r_flipped, min_disp, depth_min, depth_max;
    float width_left, width_right;
    float match_cutoff;
};

// Checked build (-DSVI_BOUNDS_CHECK, build/libsvi_gpu_checked.so): every index that is computed at run time and goes
// into shared memory, a per-frame list or an output array is asserted in range; the first violation latches
// tag * 100000 + line in g_svi_check and the next host call reports it.  compute-sanitizer is closed on the GPU pool
// this library is developed on, so the parity tests are also run once against this build (tests: test_bounds_checked_build).
#ifdef SVI_BOUNDS_CHECK
__device__ int g_svi_check = 0;
#define SVI_CHECK(tag, cond)                                                        \
    do {                                                                            \
        if (!(cond)) atomicCAS(&g_svi_check, 0, (tag) * 100000 + __LINE__);         \
    } while (0)
#else
#define SVI_CHECK(tag, cond) ((void)0)
#endif

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

// order-preserving float <-> uint mapping (for atomicMax and for sort keys)
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __host__ __forceinline__ uint32_t ordered_to_float_bits(uint32_t e) {
    return (e & 0x80000000u) ? (e ^ 0x80000000u) : ~e;
}

// Threshold as cv::goodFeaturesToTrack computes it: minMaxLoc gives a double, threshold()
// converts (double)max * qualityLevel to float for the 32F image.
__device__ __forceinline__ float gftt_threshold(uint32_t ordered_max, double quality) {
    if (ordered_max == 0u) return 0.f;  // empty mask: minMaxLoc leaves maxVal = 0
    return (float)((double)__uint_as_float(ordered_to_float_bits(ordered_max)) * quality);
}

// u8 -> float in one ALU instruction: a 32-bit signed convert is I2FP (full rate), while the narrow
// unsigned forms the compiler would pick for a byte go through the slow I2F path.
__device__ __forceinline__ float u8_to_float(uint32_t v) {
    float f;
    asm("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(v));
    return f;
}

__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v) { return __reduce_min_sync(0xFFFFFFFFu, v); }
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) { return __reduce_max_sync(0xFFFFFFFFu, v); }

// cvRound (lrint, half to even) and OpenCV BRIEF's (int)(v + 0.5)
__device__ __forceinline__ int cv_round_f(float v) { return __float2int_rn(v); }
__device__ __forceinline__ int brief_centre(float v) { return (int)((double)v + 0.5); }

// std::round(float): half away from zero
__device__ __forceinline__ float round_half_away(float v) { return roundf(v); }

}  // namespace svi
