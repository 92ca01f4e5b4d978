// Types.h -- value types on the boundary between the GPU front-end and the CPU optimisation side.
// Same names and members as the reference's src/types/Types.h:79-148 and Typedefs.h:29-63, on small
// POD stand-ins for cv::Point2f / cv::KeyPoint / cv::Mat / Eigen::Vector3d / Isometry3d (neither
// OpenCV-C++ nor Eigen is a dependency of this host layer; a tracker that has them converts with a
// memcpy: the layouts are float[2], double[3], row-major double[12] and double[16]).
#ifndef SVI_HOST_TYPES_H
#define SVI_HOST_TYPES_H

#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

// Pinned arithmetic: every product and sum below rounds on its own, in the written order (the GPU kernels and the CPU
// oracle do the same), whatever flags the including project uses -- the reference builds with -O3 -march=native
// (CMakeLists.txt:51), where GCC's default -ffp-contract=fast would fuse a*b + c into an FMA.
#if defined(__clang__)
#pragma clang fp contract(off)
#elif defined(__GNUC__)
#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
#endif

#define DESCRIPTOR_SIZE_BITS 256
#define DESCRIPTOR_SIZE_BYTES (DESCRIPTOR_SIZE_BITS / 8)

typedef uint64_t UIDLandmark;
typedef uint64_t UIDFrame;
typedef uint64_t UIDDetectionPoint;

struct Point2f {                      // cv::Point2f
    float x = 0.f, y = 0.f;
    Point2f() {}
    Point2f(float p_x, float p_y) : x(p_x), y(p_y) {}
};

struct KeyPoint {                     // cv::KeyPoint (members the hot path touches)
    Point2f pt;
    float size = 0.f;
    KeyPoint() {}
    KeyPoint(float p_x, float p_y, float p_size) : pt(p_x, p_y), size(p_size) {}
};

struct CPoint3D {                     // Eigen::Vector3d stand-in (CPoint3DCAMERA / CPoint3DWORLD)
    double v[3] = {0.0, 0.0, 0.0};
    CPoint3D() {}
    CPoint3D(double p_x, double p_y, double p_z) { v[0] = p_x; v[1] = p_y; v[2] = p_z; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
    double operator()(int i) const { return v[i]; }
};
typedef CPoint3D CPoint3DCAMERA;
typedef CPoint3D CPoint3DWORLD;

struct MatrixProjection {             // Eigen::Matrix<double,3,4>, row-major here
    double m[12] = {0};
    double operator()(int r, int c) const { return m[4 * r + c]; }
    double& operator()(int r, int c) { return m[4 * r + c]; }
};

struct Matrix3d {                     // Eigen::Matrix3d, row-major here
    double m[9] = {0};
    double operator()(int r, int c) const { return m[3 * r + c]; }
    double& operator()(int r, int c) { return m[3 * r + c]; }
    Matrix3d transpose() const {
        Matrix3d t;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) t(r, c) = (*this)(c, r);
        return t;
    }
    Matrix3d inverse() const {        // adjugate / determinant (Eigen's closed form for fixed 3x3); singular -> inf / nan like Eigen
        auto cof = [&](int i, int j) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            return (*this)(i1, j1) * (*this)(i2, j2) - (*this)(i1, j2) * (*this)(i2, j1);
        };
        const double det = (cof(0, 0) * (*this)(0, 0) + cof(1, 0) * (*this)(1, 0)) + cof(2, 0) * (*this)(2, 0);
        const double inv_det = 1.0 / det;
        Matrix3d r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r(i, j) = cof(j, i) * inv_det;
        return r;
    }
};

struct Vector4d {                     // Eigen::Vector4d
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    double operator()(int i) const { return v[i]; }
    double& operator()(int i) { return v[i]; }
};

struct Isometry3d {                   // Eigen::Isometry3d as a row-major 4x4
    double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double operator()(int r, int c) const { return m[4 * r + c]; }
    double& operator()(int r, int c) { return m[4 * r + c]; }
    CPoint3D operator*(const CPoint3D& p) const {
        return CPoint3D(m[0] * p.v[0] + m[1] * p.v[1] + m[2] * p.v[2] + m[3],
                        m[4] * p.v[0] + m[5] * p.v[1] + m[6] * p.v[2] + m[7],
                        m[8] * p.v[0] + m[9] * p.v[1] + m[10] * p.v[2] + m[11]);
    }
};

inline MatrixProjection operator*(const MatrixProjection& P, const Isometry3d& T) {
    MatrixProjection R;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += P(r, k) * T(k, c);
            R(r, c) = s;
        }
    return R;
}

inline Isometry3d inverseIsometry(const Isometry3d& T) {   // Eigen::Isometry3d::inverse(): R^T, -R^T t
    Isometry3d I;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) I(r, c) = T(c, r);
    for (int r = 0; r < 3; ++r) I(r, 3) = -(I(r, 0) * T(0, 3) + I(r, 1) * T(1, 3) + I(r, 2) * T(2, 3));
    return I;
}

inline Isometry3d operator*(const Isometry3d& A, const Isometry3d& B) {
    Isometry3d C;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += A(r, k) * B(k, c);
            C(r, c) = s + (c == 3 ? A(r, 3) : 0.0);
        }
    }
    return C;
}

typedef std::array<uint8_t, DESCRIPTOR_SIZE_BYTES> CDescriptor;   // cv::Mat 1x32 CV_8U

struct ImageView {                    // const cv::Mat& (8-bit, single channel)
    const uint8_t* data = nullptr;
    int width = 0, height = 0;
    size_t pitch = 0;
    ImageView() {}
    ImageView(const uint8_t* p_data, int p_w, int p_h, size_t p_pitch = 0) : data(p_data), width(p_w), height(p_h), pitch(p_pitch ? p_pitch : (size_t)p_w) {}
};

// src/types/Types.h:134-148
struct CMatchTriangulation {
    const CPoint3DCAMERA vecPointXYZCAMERA;
    const Point2f ptUVCAMERA;
    const CDescriptor matDescriptorCAMERA;
    CMatchTriangulation(const CPoint3DCAMERA& p_vecPointXYZCAMERA, const Point2f& p_ptUVCAMERA, const CDescriptor& p_matDescriptorCAMERA)
        : vecPointXYZCAMERA(p_vecPointXYZCAMERA), ptUVCAMERA(p_ptUVCAMERA), matDescriptorCAMERA(p_matDescriptorCAMERA) {}
};

// src/types/Types.h:79-121 -- what CLandmark::addMeasurement stores and Cg2oOptimizer / CSolverStereoPosit read
struct CMeasurementLandmark {
    UIDLandmark uID;
    Point2f ptUVLEFT;
    Point2f ptUVRIGHT;
    float fDisparity;
    CPoint3DCAMERA vecPointXYZLEFT;
    CPoint3DWORLD vecPointXYZWORLD;
    CPoint3DWORLD vecPointXYZWORLDOptimized;
    Isometry3d matTransformationWORLDtoLEFT;
    MatrixProjection matProjectionWORLDtoLEFT;
    MatrixProjection matProjectionWORLDtoRIGHT;
    uint32_t uOptimizations;
    CMeasurementLandmark(const UIDLandmark& p_uID, const Point2f& p_ptUVLEFT, const Point2f& p_ptUVRIGHT, const CPoint3DCAMERA& p_vecPointXYZ,
                         const CPoint3DWORLD& p_vecPointXYZWORLD, const CPoint3DWORLD& p_vecPointXYZWORLDOptimized,
                         const Isometry3d& p_matTransformationWORLDtoLEFT, const MatrixProjection& p_matProjectionWORLDtoLEFT,
                         const MatrixProjection& p_matProjectionWORLDtoRIGHT, const uint32_t& p_uOptimizations)
        : uID(p_uID), ptUVLEFT(p_ptUVLEFT), ptUVRIGHT(p_ptUVRIGHT), fDisparity(p_ptUVLEFT.x - p_ptUVRIGHT.x), vecPointXYZLEFT(p_vecPointXYZ),
          vecPointXYZWORLD(p_vecPointXYZWORLD), vecPointXYZWORLDOptimized(p_vecPointXYZWORLDOptimized),
          matTransformationWORLDtoLEFT(p_matTransformationWORLDtoLEFT), matProjectionWORLDtoLEFT(p_matProjectionWORLDtoLEFT),
          matProjectionWORLDtoRIGHT(p_matProjectionWORLDtoRIGHT), uOptimizations(p_uOptimizations) {}
};

// src/exceptions/CExceptionNoMatchFound.h, CExceptionParameter.h
class CExceptionNoMatchFound : public std::exception {
public:
    explicit CExceptionNoMatchFound(const std::string& p_strExceptionDescription, int p_iStatus = -1)
        : m_strExceptionDescription(p_strExceptionDescription), iStatus(p_iStatus) {}
    const char* what() const noexcept override { return m_strExceptionDescription.c_str(); }
    const std::string m_strExceptionDescription;
    const int iStatus;   // svi_status of the failing item
};
class CExceptionPoseOptimization : public std::exception {   // src/exceptions/CExceptionPoseOptimization.h
public:
    explicit CExceptionPoseOptimization(const std::string& p_strExceptionDescription) : m_strExceptionDescription(p_strExceptionDescription) {}
    const char* what() const noexcept override { return m_strExceptionDescription.c_str(); }
    const std::string m_strExceptionDescription;
};
class CExceptionParameter : public std::exception {
public:
    explicit CExceptionParameter(const std::string& p_strExceptionDescription) : m_strExceptionDescription(p_strExceptionDescription) {}
    const char* what() const noexcept override { return m_strExceptionDescription.c_str(); }
    const std::string m_strExceptionDescription;
};

#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC pop_options
#endif
#endif
