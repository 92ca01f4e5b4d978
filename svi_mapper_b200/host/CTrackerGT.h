// CTrackerGT -- the per-frame orchestration of the reference's ground-truth driven tracker
// (src/core/CTrackerGT.cpp:91-126 process, :137-332 _trackLandmarks) reduced to the calls that touch the
// stereo front-end: motion scaling, visibility reset, trackManual, the new-landmark trigger and
// addNewLandmarks and the per-frame landmark optimisation.  Display, key-framing, DBoW2 / BTree loop closing and
// g2o are the reference's CPU side and are not part of this host layer.
#ifndef SVI_HOST_CTRACKERGT_H
#define SVI_HOST_CTRACKERGT_H

#include <chrono>

#include "CFundamentalMatcher.h"

// Pinned arithmetic: every product and sum below rounds on its own, in the written order (the GPU kernels and the CPU
// oracle do the same), whatever flags the including project uses -- the reference builds with -O3 -march=native
// (CMakeLists.txt:51), where GCC's default -ffp-contract=fast would fuse a*b + c into an FMA.
#if defined(__clang__)
#pragma clang fp contract(off)
#elif defined(__GNUC__)
#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
#endif

class CTrackerGT {
public:
    CTrackerGT(const std::shared_ptr<CStereoCamera> p_pCameraSTEREO, const std::shared_ptr<CGpuContext> p_pGpu)
        : m_cMatcher(p_pCameraSTEREO, p_pGpu) {}

    // process(imgL, imgR, T_LEFTLAST->LEFTNOW) :91-126; the rotation magnitude is passed in by the caller (the
    // reference takes it from CMiniVisionToolbox::toOrientationRodrigues of the same transform)
    void process(const ImageView& p_matImageLEFT, const ImageView& p_matImageRIGHT, const Isometry3d& p_matTransformationLEFTLASTtoLEFTNOW,
                 const double p_dRotationNorm = 0.0) {
        const CPoint3D t(p_matTransformationLEFTLASTtoLEFTNOW(0, 3), p_matTransformationLEFTLASTtoLEFTNOW(1, 3), p_matTransformationLEFTLASTtoLEFTNOW(2, 3));
        const double dTranslationNorm = std::sqrt(t.x() * t.x() + t.y() * t.y() + t.z() * t.z());
        const Isometry3d matTransformationWORLDtoLEFT(p_matTransformationLEFTLASTtoLEFTNOW * m_matTransformationWORLDtoLEFTLAST);
        const Isometry3d matTransformationLEFTtoWORLD(inverseIsometry(matTransformationWORLDtoLEFT));
        // :157 motion scaling (capped)
        const double dMotionScaling = std::min(1.0 + (10.0 * p_dRotationNorm + 0.5 * dTranslationNorm), 5.0);
        const auto tStart = std::chrono::steady_clock::now();
        m_cMatcher.resetVisibilityActiveLandmarks();                                                               // :160
        m_cMatcher.trackManual(m_uFrameCount, p_matImageLEFT, p_matImageRIGHT, matTransformationWORLDtoLEFT, matTransformationLEFTtoWORLD,
                               dMotionScaling);                                                                    // :167-174
        m_uNumberofVisibleLandmarksLAST = m_cMatcher.getNumberOfVisibleLandmarks();                                // :176-193
        const auto tTracked = std::chrono::steady_clock::now();
        m_cMatcher.optimizeActiveLandmarks(m_uFrameCount);                                                         // :197
        const auto tOptimized = std::chrono::steady_clock::now();
        if (m_uVisibleLandmarksMinimum > m_uNumberofVisibleLandmarksLAST || m_uMaximumNumberOfFramesWithoutDetection < m_uNumberOfFramesWithoutDetection) {
            m_uNumberofVisibleLandmarksLAST = m_cMatcher.addNewLandmarks(p_matImageLEFT, p_matImageRIGHT, matTransformationWORLDtoLEFT,
                                                                         matTransformationLEFTtoWORLD, m_uFrameCount);   // :305-315
            m_uNumberOfFramesWithoutDetection = 0;
            ++m_uNumberOfDetections;
        } else {
            ++m_uNumberOfFramesWithoutDetection;
        }
        // the reference keeps such timers too (m_dDurationTotalSeconds* printed by tracker_gt.cpp:291-296)
        const auto tEnd = std::chrono::steady_clock::now();
        m_dDurationTrackingSeconds += std::chrono::duration<double>(tTracked - tStart).count();
        m_dDurationOptimizationSeconds += std::chrono::duration<double>(tOptimized - tTracked).count();
        m_dDurationDetectionSeconds += std::chrono::duration<double>(tEnd - tOptimized).count();
        m_matTransformationWORLDtoLEFTLAST = matTransformationWORLDtoLEFT;
        ++m_uFrameCount;
    }
    double getDurationTrackingSeconds() const { return m_dDurationTrackingSeconds; }
    double getDurationOptimizationSeconds() const { return m_dDurationOptimizationSeconds; }
    double getDurationDetectionSeconds() const { return m_dDurationDetectionSeconds; }

    CFundamentalMatcher& getMatcher() { return m_cMatcher; }
    const Isometry3d getTransformationLEFTtoWORLD() const { return inverseIsometry(m_matTransformationWORLDtoLEFTLAST); }
    UIDFrame getFrameCount() const { return m_uFrameCount; }
    uint64_t getNumberOfVisibleLandmarksLAST() const { return m_uNumberofVisibleLandmarksLAST; }
    uint64_t getNumberOfDetections() const { return m_uNumberOfDetections; }

private:
    CFundamentalMatcher m_cMatcher;
    Isometry3d m_matTransformationWORLDtoLEFTLAST;
    UIDFrame m_uFrameCount = 0;
    uint64_t m_uNumberofVisibleLandmarksLAST = 0, m_uNumberOfDetections = 0;
    const uint64_t m_uVisibleLandmarksMinimum = 100;            // CTrackerGT.cpp:29
    const uint8_t m_uMaximumNumberOfFramesWithoutDetection = 2; // CTrackerGT.h:56
    uint8_t m_uNumberOfFramesWithoutDetection = 0;
    double m_dDurationTrackingSeconds = 0.0, m_dDurationOptimizationSeconds = 0.0, m_dDurationDetectionSeconds = 0.0;
};

#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC pop_options
#endif
#endif
