// landmark_opt.cuh -- CLandmark::optimize for every active landmark of a frame in one launch.
//
// Replaces the per-landmark CPU loop of CFundamentalMatcher::optimizeActiveLandmarks (reference
// src/core/CFundamentalMatcher.cpp:265-277) over CLandmark::optimize (src/types/CLandmark.cpp:281-296) and
// _getOptimizedLandmarkSTEREOUV (:447-581): Gauss-Newton on the stereo re-projection error of all measurements of a
// landmark, robust weights above dKernelMaximumErrorSquaredPixels, the homogeneous coordinate held fixed (the step is the
// least-squares solution of the 4 x 3 system H.block<4,3>(0,0) dx = -b by Householder QR).  On the tracker's frame loop
// this is the host CPU's largest item (11 ms per frame for 2100 landmarks with ~30 measurements each, against 1.4 ms for
// trackManual through the GPU); the landmarks are independent, so one WARP owns one landmark and evaluates its measurements
// with the reference's operation order and sums them in the reference's order (fp64, every operation rounded on its own:
// the library is built with -fmad=false) -- the result is bit-identical to the C++ host implementation of the same loop
// (svi_mapper_b200/host/CFundamentalMatcher.h, CLandmark::optimize), which the parity test uses as its checker.
// A measurement names the camera pose it was taken with by an index into a table of projection pairs
// (P_LEFT * T_WORLD->LEFT, P_RIGHT * T_WORLD->LEFT: one row per frame), instead of carrying two 3 x 4 matrices of its own.
#pragma once
#include "common.cuh"

namespace svi {

// CLandmark.h:90-98
constexpr uint32_t kOptCapIterations = 1000;
constexpr double kOptConvergenceDelta = 1e-5;
constexpr double kOptMinimumRatioInliers = 0.5;
constexpr double kOptKernelMaximumErrorSquaredPixels = 10.0;
constexpr double kOptMaximumErrorSquaredAveragePixels = 9.0;
constexpr int kOptMinimumMeasurements = 5;   // optimise when MORE than this many measurements exist

struct LandmarkOptIn {
    const double* xyz_guess;     // n x 3
    const int* first;            // n + 1
    const int* pose_index;       // m
    const float* uv_left;        // m x 2
    const float* uv_right;       // m x 2
    const double* proj_left;     // n_poses x 12, row-major 3 x 4
    const double* proj_right;    // n_poses x 12
    int n_poses;
};
struct LandmarkOptOut {
    double* xyz;                 // n x 3
    uint8_t* outcome;            // n: svi_optimize_outcome
    double* avg_sq_error;        // n
    int* iterations;             // n
};

// min || A x + b || for the 4 x 3 matrix A = first three columns of H, by Householder reflections (same sequence of
// operations as the host layer's solveLeastSquares4x3)
__device__ __forceinline__ void solve_least_squares_4x3(const double (&H)[4][4], const double (&b)[4], double (&x)[3]) {
    double A[4][3], y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) A[i][j] = H[i][j];
        y[i] = -b[i];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double norm = 0.0;
#pragma unroll
        for (int r = c; r < 4; ++r) norm += A[r][c] * A[r][c];
        norm = sqrt(norm);
        if (0.0 == norm) continue;
        const double alpha = A[c][c] > 0.0 ? -norm : norm;
        double v[4] = {0, 0, 0, 0};
        v[c] = A[c][c] - alpha;
#pragma unroll
        for (int r = c + 1; r < 4; ++r) v[r] = A[r][c];
        double vv = 0.0;
#pragma unroll
        for (int r = c; r < 4; ++r) vv += v[r] * v[r];
        if (0.0 == vv) continue;
#pragma unroll
        for (int j = c; j < 3; ++j) {
            double d = 0.0;
#pragma unroll
            for (int r = c; r < 4; ++r) d += v[r] * A[r][j];
#pragma unroll
            for (int r = c; r < 4; ++r) A[r][j] -= 2.0 * d / vv * v[r];
        }
        double d = 0.0;
#pragma unroll
        for (int r = c; r < 4; ++r) d += v[r] * y[r];
#pragma unroll
        for (int r = c; r < 4; ++r) y[r] -= 2.0 * d / vv * v[r];
    }
#pragma unroll
    for (int r = 2; r >= 0; --r) {
        double s2 = y[r];
#pragma unroll
        for (int j = r + 1; j < 3; ++j) s2 -= A[r][j] * x[j];
        x[r] = (0.0 != A[r][r]) ? s2 / A[r][r] : 0.0;
    }
}

// One WARP per landmark.  An iteration walks the measurements in chunks of 32: lane l evaluates measurement c0 + l
// (re-projection error, Jacobian, robust weight) and parks its fifteen contributions -- the ten distinct entries of the
// symmetric w * J^T J, the four of w * J^T e, and w * e^2 -- in shared memory; then lane j < 15 adds contribution j of the
// chunk's measurements to its accumulator IN MEASUREMENT ORDER, which is exactly the sequence of additions the reference's
// loop performs on that entry, so the sums are bit-identical to the sequential code.  The 4 x 3 solve and the convergence
// test run redundantly in every lane.
// Measured on the C3 sequence through the C++ tracker (2100 landmarks, 83 k measurements per call): 4.5 ms per call, against
// 13.8 ms with one thread per landmark and 11 ms for the host's CPU loop.  Nearly all of it is ONE landmark: a typical
// landmark converges in 4 iterations (p99 23), but the reference has no guard against a landmark whose position has run
// away (here to 1e83 m after an earlier "successful" optimisation): it walks chaotically, never meets the 1e-5 criterion and
// burns all 1000 iterations in every frame -- on the GPU 1000 x 4.4 us of dependent fp64 divisions (slow-path, the
// operands leave the fast path's exponent range) on a single warp.  No state repeats exactly, so there is no exact early
// exit; the launch takes 0.1 ms without such a landmark.  (Dividing on exponent-scaled operands -- bit-identical, checked on
// 4e7 random bit patterns -- to keep the compiler's division on its fast path made the iteration SLOWER, 8.1 vs 4.5 ms per
// call: the time is the length of the dependent chain, not slow-path calls.)
constexpr int kOptWarps = 4;          // landmarks per CTA
constexpr int kOptTerms = 15;
constexpr int kOptTermPitch = 17;     // doubles per measurement slot: odd pitch, parked rows do not collide on one bank

// outcome codes = svi_optimize_outcome of include/svi_gpu.h
__global__ void __launch_bounds__(kOptWarps * 32)
optimize_landmarks_kernel(LandmarkOptIn in, int n, LandmarkOptOut out) {
    __shared__ double s_term[kOptWarps][32][kOptTermPitch];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * kOptWarps + warp;
    if (i >= n) return;   // whole warps leave; the kernel has no CTA-wide barrier
    double (*term)[kOptTermPitch] = s_term[warp];
    const int m0 = in.first[i], m1 = in.first[i + 1], nm = m1 - m0;
    const double g0 = in.xyz_guess[3 * i], g1 = in.xyz_guess[3 * i + 1], g2 = in.xyz_guess[3 * i + 2];
    double X[4] = {g0, g1, g2, 1.0};
    int outcome = 4;   // SVI_OPT_NOT_CONVERGED
    double avg = 0.0;
    int iterations = 0;
    if (nm <= kOptMinimumMeasurements) {
        outcome = 0;   // SVI_OPT_SKIPPED: bIsOptimal = true, position untouched
    } else {
        double err_prev = 0.0;
        for (uint32_t it = 0; it < kOptCapIterations; ++it) {
            double acc = 0.0;      // lane j < 15: entry j of (H upper triangle row by row | b | total error)
            uint32_t inliers = 0;
            for (int c0 = m0; c0 < m1; c0 += 32) {
                const int m = c0 + lane;
                bool inlier = false;
                if (m < m1) {
                    const int pi = in.pose_index[m];
                    SVI_CHECK(8, pi >= 0 && pi < in.n_poses);
                    const double* P[2] = {in.proj_left + 12 * (size_t)pi, in.proj_right + 12 * (size_t)pi};
                    const float uv[2][2] = {{in.uv_left[2 * m], in.uv_left[2 * m + 1]}, {in.uv_right[2 * m], in.uv_right[2 * m + 1]}};
                    double J[4][4], e[4];
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        double p[12];
#pragma unroll
                        for (int k = 0; k < 12; ++k) p[k] = P[s][k];
                        double a[3];
#pragma unroll
                        for (int r = 0; r < 3; ++r) a[r] = p[4 * r] * X[0] + p[4 * r + 1] * X[1] + p[4 * r + 2] * X[2] + p[4 * r + 3] * X[3];
                        const double c = a[2];
                        e[2 * s] = a[0] / c - uv[s][0];
                        e[2 * s + 1] = a[1] / c - uv[s][1];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            J[2 * s][k] = p[k] / c - a[0] / (c * c) * p[8 + k];
                            J[2 * s + 1][k] = p[4 + k] / c - a[1] / (c * c) * p[8 + k];
                        }
                    }
                    const double e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3];
                    double w = 1.0;
                    if (kOptKernelMaximumErrorSquaredPixels < e2) w = kOptKernelMaximumErrorSquaredPixels / e2;
                    else inlier = true;
                    int t = 0;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = r; c < 4; ++c) term[lane][t++] = w * (J[0][r] * J[0][c] + J[1][r] * J[1][c] + J[2][r] * J[2][c] + J[3][r] * J[3][c]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) term[lane][10 + r] = w * (J[0][r] * e[0] + J[1][r] * e[1] + J[2][r] * e[2] + J[3][r] * e[3]);
                    term[lane][14] = w * e2;
                }
                inliers += __popc(__ballot_sync(0xFFFFFFFFu, inlier));
                __syncwarp();
                const int cnt = min(32, m1 - c0);
                if (lane < kOptTerms)
                    for (int k = 0; k < cnt; ++k) acc += term[k][lane];   // in measurement order
                __syncwarp();
            }
            double H[4][4], b[4];
            {
                double v[kOptTerms];
#pragma unroll
                for (int j = 0; j < kOptTerms; ++j) v[j] = __shfl_sync(0xFFFFFFFFu, acc, j);
                int t = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = r; c < 4; ++c) { H[r][c] = v[t]; H[c][r] = v[t]; ++t; }   // J_r * J_c == J_c * J_r: the sum is symmetric bit for bit
#pragma unroll
                for (int r = 0; r < 4; ++r) b[r] = v[10 + r];
                const double err_total = v[14];
                double dx[3];
                solve_least_squares_4x3(H, b, dx);
#pragma unroll
                for (int k = 0; k < 3; ++k) X[k] += dx[k];
                iterations = (int)it + 1;
                if (kOptConvergenceDelta > fabs(err_prev - err_total)) {
                    const double err_avg = err_total / (double)nm;
                    if (kOptMinimumRatioInliers < (double)inliers / (double)nm) {
                        avg = err_avg;
                        outcome = (kOptMaximumErrorSquaredAveragePixels > err_avg) ? 2 : 1;   // SVI_OPT_OPTIMAL / SVI_OPT_CONVERGED
                    } else {
                        outcome = 3;   // SVI_OPT_REJECTED
                    }
                    break;
                }
                err_prev = err_total;
            }
        }
    }
    if (lane == 0) {
        const bool moved = outcome == 1 || outcome == 2;
        out.xyz[3 * i] = moved ? X[0] : g0;
        out.xyz[3 * i + 1] = moved ? X[1] : g1;
        out.xyz[3 * i + 2] = moved ? X[2] : g2;
        out.outcome[i] = (uint8_t)outcome;
        out.avg_sq_error[i] = avg;
        out.iterations[i] = iterations;
    }
}

}  // namespace svi
