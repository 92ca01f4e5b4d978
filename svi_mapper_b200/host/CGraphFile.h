// g2o graph file of the front-end's hand-off: what Cg2oOptimizer::optimizeContinuous saves before
// and after a run (src/optimization/Cg2oOptimizer.cpp:495-514).  g2o itself is an un-vendored third-party library
// ("trunk"); tags and field orders are those of its published types_slam3d text format, EDGE_SE3_LINEAR_ACCELERATION is
// the reference's own type (src/optimization/edge_se3_linear_acceleration.cpp:35-112).  The graph is assembled exactly as
// the reference assembles it -- parameters :99-120, landmark vertices :1143-1152, pose vertices and pose edges :1229-1270
// (id shift 1000000, first pose fixed, information 100000 * I6 with the translation block scaled by 1 / (1 + |t|^2)),
// one gravity edge per key frame :982-997, and per measurement the XYZ / UV-depth / UV-disparity edge that
// _setLandmarkMeasurementsWORLD :1383-1466 selects.  Vertices in id order, FIX after a fixed vertex, edges in insertion
// order (g2o's save order); numbers with 17 significant digits.  The numpy twin is svi_mapper_b200/formats.py::g2o_lines;
// tests/test_host.py compares the two byte for byte.
#ifndef SVI_HOST_CGRAPHFILE_H
#define SVI_HOST_CGRAPHFILE_H

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "Types.h"

#if defined(__clang__)
#pragma clang fp contract(off)
#elif defined(__GNUC__)
#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
#endif

struct CGraphLandmark {               // CLandmark members the graph needs
    UIDLandmark uID = 0;
    CPoint3DWORLD vecPointXYZOptimized;
};
struct CGraphMeasurement {            // CMeasurementLandmark members the graph needs (Types.h:79-121)
    UIDLandmark uID = 0;
    Point2f ptUVLEFT, ptUVRIGHT;
    CPoint3DCAMERA vecPointXYZLEFT;
};
struct CGraphKeyFrame {               // CKeyFrame members the graph needs (src/types/CKeyFrame.h)
    UIDFrame uID = 0;
    Isometry3d matTransformationLEFTtoWORLD;
    CPoint3D vecLinearAccelerationNormalized;
    std::vector<CGraphMeasurement> vecMeasurements;
};
struct CGraphCameras {                // m_dFxP, m_dFyP, m_dCxP, m_dCyP of both cameras, CStereoCamera::m_dBaselineMeters
    double dLEFT[4] = {0, 0, 0, 0}, dRIGHT[4] = {0, 0, 0, 0};
    double dBaselineMeters = 0.0;
};

namespace graph_io {
constexpr int64_t uIDShift = 1000000;                           // Cg2oOptimizer.h:83
constexpr double dMaximumReliableDepthForPointXYZL2 = 10.0;      // :92-94
constexpr double dMaximumReliableDepthForUVDepthL2 = 50.0;
constexpr double dMaximumReliableDepthForUVDisparityL2 = 10000.0;

inline std::string num(double v) {
    char ch[40];
    std::snprintf(ch, sizeof(ch), "%.17g", v);
    return ch;
}
// Eigen::Quaterniond(R) followed by normalize() -- g2o::internal::toVectorQT; order x y z w
inline void quaternionXYZW(const Isometry3d& T, double q[4]) {
    auto R = [&](int r, int c) { return T(r, c); };
    double t = (R(0, 0) + R(1, 1)) + R(2, 2);
    q[0] = q[1] = q[2] = q[3] = 0.0;
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R(2, 1) - R(1, 2)) * t;
        q[1] = (R(0, 2) - R(2, 0)) * t;
        q[2] = (R(1, 0) - R(0, 1)) * t;
    } else {
        int i = 0;
        if (R(1, 1) > R(0, 0)) i = 1;
        if (R(2, 2) > R(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        t = std::sqrt(((R(i, i) - R(j, j)) - R(k, k)) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R(k, j) - R(j, k)) * t;
        q[j] = (R(j, i) + R(i, j)) * t;
        q[k] = (R(k, i) + R(i, k)) * t;
    }
    const double n = std::sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
    for (int a = 0; a < 4; ++a) q[a] = q[a] / n;
}
inline std::string se3(const Isometry3d& T) {
    double q[4];
    quaternionXYZW(T, q);
    std::string s = num(T(0, 3)) + " " + num(T(1, 3)) + " " + num(T(2, 3));
    for (int a = 0; a < 4; ++a) s += " " + num(q[a]);
    return s;
}
inline std::string diagonalUpper(const double* d, int n) {   // upper triangle of diag(d), row by row
    std::string s;
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) s += " " + num(i == j ? d[i] : 0.0);
    return s;
}
}   // namespace graph_io

// the lines Cg2oOptimizer's graph serialises to; p_vecKeyFrames in insertion order (the first one is the fixed pose)
inline std::vector<std::string> getGraphLines(const CGraphCameras& p_cCameras, const std::vector<CGraphKeyFrame>& p_vecKeyFrames,
                                              const std::vector<CGraphLandmark>& p_vecLandmarks, const CPoint3D& p_vecTranslationToG2o) {
    using namespace graph_io;
    std::vector<std::string> vecLines;
    const std::string strIdentity = "0 0 0 0 0 0 1";
    auto camera = [&](const double* d) { return " " + num(d[0]) + " " + num(d[1]) + " " + num(d[2]) + " " + num(d[3]); };
    vecLines.push_back("PARAMS_SE3OFFSET 0 " + strIdentity);
    vecLines.push_back("PARAMS_CAMERAPARAMETERS 1 " + strIdentity + camera(p_cCameras.dLEFT));
    vecLines.push_back("PARAMS_CAMERAPARAMETERS 2 " + strIdentity + camera(p_cCameras.dRIGHT));
    vecLines.push_back("PARAMS_SE3OFFSET 3 " + strIdentity);

    std::map<UIDLandmark, CPoint3DWORLD> mapLandmarks;
    for (const CGraphLandmark& cLandmark : p_vecLandmarks)
        mapLandmarks[cLandmark.uID] = CPoint3DWORLD(cLandmark.vecPointXYZOptimized.v[0] + p_vecTranslationToG2o.v[0],
                                                    cLandmark.vecPointXYZOptimized.v[1] + p_vecTranslationToG2o.v[1],
                                                    cLandmark.vecPointXYZOptimized.v[2] + p_vecTranslationToG2o.v[2]);
    for (const auto& it : mapLandmarks)
        vecLines.push_back("VERTEX_TRACKXYZ " + std::to_string(it.first) + " " + num(it.second.v[0]) + " " + num(it.second.v[1]) + " " + num(it.second.v[2]));

    std::vector<Isometry3d> vecPoses;
    for (const CGraphKeyFrame& cKeyFrame : p_vecKeyFrames) {
        Isometry3d T = cKeyFrame.matTransformationLEFTtoWORLD;
        for (int r = 0; r < 3; ++r) T(r, 3) += p_vecTranslationToG2o.v[r];
        vecPoses.push_back(T);
    }
    std::vector<size_t> vecOrder(p_vecKeyFrames.size());
    for (size_t n = 0; n < vecOrder.size(); ++n) vecOrder[n] = n;
    std::stable_sort(vecOrder.begin(), vecOrder.end(), [&](size_t a, size_t b) { return p_vecKeyFrames[a].uID < p_vecKeyFrames[b].uID; });
    for (size_t n : vecOrder) {
        const std::string strID = std::to_string((int64_t)p_vecKeyFrames[n].uID + uIDShift);
        vecLines.push_back("VERTEX_SE3:QUAT " + strID + " " + se3(vecPoses[n]));
        if (n == 0) vecLines.push_back("FIX " + strID);
    }
    for (size_t n = 0; n < p_vecKeyFrames.size(); ++n) {
        const CGraphKeyFrame& cKeyFrame = p_vecKeyFrames[n];
        const std::string strID = std::to_string((int64_t)cKeyFrame.uID + uIDShift);
        if (n > 0) {
            const Isometry3d M = inverseIsometry(vecPoses[n - 1]) * vecPoses[n];
            const double dInformationFactor = 1.0 / (1.0 + ((M(0, 3) * M(0, 3) + M(1, 3) * M(1, 3)) + M(2, 3) * M(2, 3)));
            double d[6];
            for (int i = 0; i < 6; ++i) d[i] = i < 3 ? 100000.0 * dInformationFactor : 100000.0;
            vecLines.push_back("EDGE_SE3:QUAT " + std::to_string((int64_t)p_vecKeyFrames[n - 1].uID + uIDShift) + " " + strID + " " + se3(M) + diagonalUpper(d, 6));
        }
        const double dOne[3] = {1.0, 1.0, 1.0};
        vecLines.push_back("EDGE_SE3_LINEAR_ACCELERATION " + strID + " 3 " + num(cKeyFrame.vecLinearAccelerationNormalized.v[0]) + " " +
                           num(cKeyFrame.vecLinearAccelerationNormalized.v[1]) + " " + num(cKeyFrame.vecLinearAccelerationNormalized.v[2]) + diagonalUpper(dOne, 3));
        const Isometry3d matWORLDtoLEFT = inverseIsometry(vecPoses[n]);
        for (const CGraphMeasurement& cMeasurement : cKeyFrame.vecMeasurements) {
            const auto itLandmark = mapLandmarks.find(cMeasurement.uID);
            if (itLandmark == mapLandmarks.end()) continue;
            const CPoint3D& p = itLandmark->second;
            double e[3];
            for (int r = 0; r < 3; ++r) e[r] = ((matWORLDtoLEFT(r, 0) * p.v[0] + matWORLDtoLEFT(r, 1) * p.v[1]) + matWORLDtoLEFT(r, 2) * p.v[2]) + matWORLDtoLEFT(r, 3);
            const double* x = cMeasurement.vecPointXYZLEFT.v;
            const double dDistanceL2Absolute = (x[0] * x[0] + x[1] * x[1]) + x[2] * x[2];
            const double dDistanceL2Relative = ((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]) / dDistanceL2Absolute;
            if (!(0.75 < dDistanceL2Relative && 1.25 > dDistanceL2Relative)) continue;
            const double f = 1.0 / x[2];
            const std::string strHead = strID + " " + std::to_string(cMeasurement.uID) + " ";
            if (dMaximumReliableDepthForPointXYZL2 > dDistanceL2Absolute) {
                const double d[3] = {f * 1000, f * 1000, f * 1000};
                vecLines.push_back("EDGE_SE3_TRACKXYZ " + strHead + "0 " + num(x[0]) + " " + num(x[1]) + " " + num(x[2]) + diagonalUpper(d, 3));
            } else if (dMaximumReliableDepthForUVDepthL2 > dDistanceL2Absolute) {
                const double d[3] = {f, f, f * 100};
                vecLines.push_back("EDGE_PROJECT_DEPTH " + strHead + "1 " + num(cMeasurement.ptUVLEFT.x) + " " + num(cMeasurement.ptUVLEFT.y) + " " + num(x[2]) + diagonalUpper(d, 3));
            } else if (dMaximumReliableDepthForUVDisparityL2 > dDistanceL2Absolute) {
                const double dDisparity = cMeasurement.ptUVLEFT.x - cMeasurement.ptUVRIGHT.x;   // float subtraction, then widened
                if (1.0 < dDisparity) {
                    const double d[3] = {f, f, f * 1000};
                    vecLines.push_back("EDGE_PROJECT_DISPARITY " + strHead + "1 " + num(cMeasurement.ptUVLEFT.x) + " " + num(cMeasurement.ptUVLEFT.y) + " " +
                                       num(dDisparity / (p_cCameras.dLEFT[0] * p_cCameras.dBaselineMeters)) + diagonalUpper(d, 3));
                }
            }
        }
    }
    return vecLines;
}

inline void saveGraphToFile(const std::string& p_strFile, const CGraphCameras& p_cCameras, const std::vector<CGraphKeyFrame>& p_vecKeyFrames,
                            const std::vector<CGraphLandmark>& p_vecLandmarks, const CPoint3D& p_vecTranslationToG2o = CPoint3D()) {
    std::ofstream ofGraph(p_strFile);
    if (!ofGraph.good()) throw std::invalid_argument("cannot write graph file");
    for (const std::string& strLine : getGraphLines(p_cCameras, p_vecKeyFrames, p_vecLandmarks, p_vecTranslationToG2o)) ofGraph << strLine << "\n";
}

#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC pop_options
#endif
#endif
