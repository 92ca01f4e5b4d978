// facade_demo -- drives the C++ host facade the way CTrackerGT::_trackLandmarks drives the reference
// (src/core/CTrackerGT.cpp:160-174, 305-319): addNewLandmarks on frame 0, trackManual on frame 1.
// Usage: facade_demo <left_calib> <right_calib> <W> <H> <L0.raw> <R0.raw> <L1.raw> <R1.raw> <out.txt>
// Writes one line per landmark of frame 0 and one per tracked landmark of frame 1 (parsed by the tests).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "CTrackerGT.h"

static std::vector<uint8_t> readRaw(const char* path, size_t n) {
    std::vector<uint8_t> v(n);
    std::ifstream f(path, std::ios::binary);
    if (!f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)n)) { std::fprintf(stderr, "cannot read %s\n", path); std::exit(2); }
    return v;
}

int main(int argc, char** argv) {
    if (argc < 10) { std::fprintf(stderr, "usage: see source\n"); return 2; }
    try {
        CParameterBase::loadCameraLEFT(argv[1]);
        CParameterBase::loadCameraRIGHT(argv[2]);
        CParameterBase::constructCameraSTEREO(CPoint3D(-0.54, 0.0, 0.0));   // tracker_gt.cpp:123
        const int W = std::atoi(argv[3]), H = std::atoi(argv[4]);
        if (W != (int)CParameterBase::pCameraLEFT->m_uWidthPixel || H != (int)CParameterBase::pCameraLEFT->m_uHeightPixel) {
            std::fprintf(stderr, "image size does not match the calibration\n");
            return 2;
        }
        const std::vector<uint8_t> L0 = readRaw(argv[5], (size_t)W * H), R0 = readRaw(argv[6], (size_t)W * H);
        const std::vector<uint8_t> L1 = readRaw(argv[7], (size_t)W * H), R1 = readRaw(argv[8], (size_t)W * H);
        auto pGpu = std::make_shared<CGpuContext>(CParameterBase::pCameraSTEREO);
        CFundamentalMatcher cMatcher(CParameterBase::pCameraSTEREO, pGpu);
        const Isometry3d matIdentity;
        std::FILE* out = std::fopen(argv[9], "w");
        const size_t n = cMatcher.addNewLandmarks(ImageView(L0.data(), W, H), ImageView(R0.data(), W, H), matIdentity, matIdentity, 0);
        std::fprintf(out, "NEW %zu\n", n);
        for (const CLandmark* p : cMatcher.getLandmarksWINDOW()) {
            const CMeasurementLandmark* m = p->getLastMeasurement();
            std::fprintf(out, "L %lu %.1f %.1f %.1f %.1f %.17g %.17g %.17g\n", (unsigned long)p->uID, m->ptUVLEFT.x, m->ptUVLEFT.y, m->ptUVRIGHT.x,
                         m->ptUVRIGHT.y, m->vecPointXYZLEFT.x(), m->vecPointXYZLEFT.y(), m->vecPointXYZLEFT.z());
        }
        cMatcher.resetVisibilityActiveLandmarks();
        cMatcher.trackManual(1, ImageView(L1.data(), W, H), ImageView(R1.data(), W, H), matIdentity, matIdentity, 1.0);
        std::fprintf(out, "TRACKED %lu STAGE2 %lu STAGE3 %lu VISIBLE %zu\n", (unsigned long)cMatcher.getNumberOfTracksStage1(),
                     (unsigned long)cMatcher.getNumberOfTracksStage2_1(), (unsigned long)cMatcher.getNumberOfTracksStage3(),
                     cMatcher.getNumberOfVisibleLandmarks());
        for (const CMeasurementLandmark* m : cMatcher.getMeasurementsForVisibleLandmarks())
            std::fprintf(out, "T %lu %.1f %.1f %.1f %.1f %.17g\n", (unsigned long)m->uID, m->ptUVLEFT.x, m->ptUVLEFT.y, m->ptUVRIGHT.x, m->ptUVRIGHT.y,
                         m->vecPointXYZLEFT.z());
        // scalar CTriangulator calls with exception semantics
        const auto tri = cMatcher.getTriangulator();
        try {
            tri->getPointInLEFT(Point2f(100.f, 50.f), Point2f(100.f, 50.f));
            std::fprintf(out, "EXC none\n");
        } catch (const CExceptionNoMatchFound& e) {
            std::fprintf(out, "EXC %s\n", e.what());
        }
        // a short sequence through the CTrackerGT orchestration: the same pair shown 6 times while the camera is
        // reported to move 1 cm per frame along x (tracker_gt.cpp feeds ground-truth poses the same way)
        CTrackerGT cTracker(CParameterBase::pCameraSTEREO, pGpu);
        Isometry3d matStep;
        matStep(0, 3) = 0.01;
        for (int uFrame = 0; uFrame < 6; ++uFrame) {
            cTracker.process(ImageView(L0.data(), W, H), ImageView(R0.data(), W, H), uFrame ? matStep : matIdentity);
            std::fprintf(out, "SEQ %d VISIBLE %lu TOTAL %lu S1 %lu S2 %lu S3 %lu DETECTIONS %lu\n", uFrame,
                         (unsigned long)cTracker.getNumberOfVisibleLandmarksLAST(), (unsigned long)cTracker.getMatcher().getNumberOfLandmarksTotal(),
                         (unsigned long)cTracker.getMatcher().getNumberOfTracksStage1(), (unsigned long)cTracker.getMatcher().getNumberOfTracksStage2_1(),
                         (unsigned long)cTracker.getMatcher().getNumberOfTracksStage3(), (unsigned long)cTracker.getNumberOfDetections());
        }
        std::fclose(out);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "facade_demo failed: %s\n", e.what());
        return 1;
    }
    return 0;
}
