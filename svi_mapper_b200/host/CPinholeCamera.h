// CPinholeCamera / CStereoCamera / CParameterBase -- host-side camera model and calibration loading,
// same public members as the reference (src/vision/CPinholeCamera.h:16-227, src/vision/CStereoCamera.h:14-35,
// src/utility/CParameterBase.h:21-226,312-318) without Eigen/OpenCV.
#ifndef SVI_HOST_CPINHOLECAMERA_H
#define SVI_HOST_CPINHOLECAMERA_H

#include <algorithm>
#include <cmath>
#include <fstream>
#include <memory>
#include <stdexcept>

#include "Types.h"

// Pinned arithmetic: every product and sum below rounds on its own, in the written order (the GPU kernels and the CPU
// oracle do the same), whatever flags the including project uses -- the reference builds with -O3 -march=native
// (CMakeLists.txt:51), where GCC's default -ffp-contract=fast would fuse a*b + c into an FMA.
#if defined(__clang__)
#pragma clang fp contract(off)
#elif defined(__GNUC__)
#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
#endif

struct CRect {   // cv::Rect
    int x, y, width, height;
    bool contains(const Point2f& p) const {   // cv::Rect::contains after Point2f -> Point2i (cvRound)
        const long px = std::lrint(p.x), py = std::lrint(p.y);
        return x <= px && px < x + width && y <= py && py < y + height;
    }
};

class CPinholeCamera {
public:
    // the reference's constructor (CPinholeCamera.h:16-61): eight arguments, the derived members in its order
    CPinholeCamera(const std::string& p_strLabel, const uint32_t& p_uWidthPixels, const uint32_t& p_uHeightPixels,
                   const MatrixProjection& p_matProjection, const Matrix3d& p_matIntrinsic, const double& p_dFocalLengthMeters,
                   const Vector4d& p_vecDistortionCoefficients, const Matrix3d& p_matRectification)
        : m_strCameraLabel(p_strLabel), m_uWidthPixel(p_uWidthPixels), m_uHeightPixel(p_uHeightPixels), m_matProjection(p_matProjection),
          m_matIntrinsic(p_matIntrinsic), m_matIntrinsicP(block33(p_matProjection)), m_matIntrinsicInverse(m_matIntrinsic.inverse()),
          m_matIntrinsicPInverse(m_matIntrinsicP.inverse()), m_matIntrinsicInverseTransposed(m_matIntrinsicInverse.transpose()),
          m_matIntrinsicPInverseTransposed(m_matIntrinsicPInverse.transpose()), m_matIntrinsicTransposed(m_matIntrinsic.transpose()),
          m_matIntrinsicPTransposed(m_matIntrinsicP.transpose()), m_dFx(m_matIntrinsic(0, 0)), m_dFy(m_matIntrinsic(1, 1)),
          m_dFxP(m_matIntrinsicP(0, 0)), m_dFyP(m_matIntrinsicP(1, 1)), m_dFxNormalized(m_dFx / m_uWidthPixel),
          m_dFyNormalized(m_dFy / m_uHeightPixel), m_dCx(m_matIntrinsic(0, 2)), m_dCy(m_matIntrinsic(1, 2)), m_dCxP(m_matIntrinsicP(0, 2)),
          m_dCyP(m_matIntrinsicP(1, 2)), m_dCxNormalized(m_dCx / m_uWidthPixel), m_dCyNormalized(m_dCy / m_uHeightPixel),
          m_dFocalLengthMeters(p_dFocalLengthMeters), m_vecDistortionCoefficients(p_vecDistortionCoefficients),
          m_matRectification(p_matRectification), m_iWidthPixel(p_uWidthPixels), m_iHeightPixel(p_uHeightPixels),
          m_dWidthPixels(p_uWidthPixels), m_dHeightPixels(p_uHeightPixels), m_fWidthPixels(p_uWidthPixels), m_fHeightPixels(p_uHeightPixels),
          m_cFieldOfView{28, 28, (int)p_uWidthPixels - 56, (int)p_uHeightPixels - 56} {}
    // convenience for synthetic cameras that only have a projection matrix (not a reference signature)
    CPinholeCamera(const std::string& p_strLabel, const uint32_t& p_uWidthPixels, const uint32_t& p_uHeightPixels,
                   const MatrixProjection& p_matProjection)
        : CPinholeCamera(p_strLabel, p_uWidthPixels, p_uHeightPixels, p_matProjection, Matrix3d(), 0.0, Vector4d(), Matrix3d()) {}

    const std::string m_strCameraLabel;
    const uint32_t m_uWidthPixel, m_uHeightPixel;
    const MatrixProjection m_matProjection;
    const Matrix3d m_matIntrinsic, m_matIntrinsicP, m_matIntrinsicInverse, m_matIntrinsicPInverse, m_matIntrinsicInverseTransposed,
        m_matIntrinsicPInverseTransposed, m_matIntrinsicTransposed, m_matIntrinsicPTransposed;
    const double m_dFx, m_dFy, m_dFxP, m_dFyP, m_dFxNormalized, m_dFyNormalized, m_dCx, m_dCy, m_dCxP, m_dCyP, m_dCxNormalized,
        m_dCyNormalized, m_dFocalLengthMeters;
    const Vector4d m_vecDistortionCoefficients;
    const Matrix3d m_matRectification;
    const int32_t m_iWidthPixel, m_iHeightPixel;
    const double m_dWidthPixels, m_dHeightPixels;
    const float m_fWidthPixels, m_fHeightPixels;
    const CRect m_cFieldOfView;

    static Matrix3d block33(const MatrixProjection& P) {   // m_matProjection.block<3,3>(0,0)
        Matrix3d k;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) k(r, c) = P(r, c);
        return k;
    }

    // CPinholeCamera.h:202-210
    const Point2f getProjectionRounded(const CPoint3DCAMERA& p) const {
        const double u = m_matProjection(0, 0) * p.x() + m_matProjection(0, 1) * p.y() + m_matProjection(0, 2) * p.z() + m_matProjection(0, 3);
        const double v = m_matProjection(1, 0) * p.x() + m_matProjection(1, 1) * p.y() + m_matProjection(1, 2) * p.z() + m_matProjection(1, 3);
        const double w = m_matProjection(2, 0) * p.x() + m_matProjection(2, 1) * p.y() + m_matProjection(2, 2) * p.z() + m_matProjection(2, 3);
        return Point2f(std::round(static_cast<float>(u / w)), std::round(static_cast<float>(v / w)));
    }
    // CPinholeCamera.h:220-227
    double getPrincipalWeightU(const Point2f& p_ptUV) const { return std::sqrt(std::fabs(p_ptUV.x - m_dCxP)) / 10.0; }
    double getPrincipalWeightV(const Point2f& p_ptUV) const { return std::sqrt(std::fabs(p_ptUV.y - m_dCyP)) / 10.0; }
};

class CStereoCamera {
public:
    CStereoCamera(const std::shared_ptr<CPinholeCamera> p_pCameraLEFT, const std::shared_ptr<CPinholeCamera> p_pCameraRIGHT,
                  const CPoint3D& p_vecTranslationToRIGHT)
        : m_pCameraLEFT(p_pCameraLEFT), m_pCameraRIGHT(p_pCameraRIGHT), m_uPixelWidth(p_pCameraLEFT->m_uWidthPixel),
          m_uPixelHeight(p_pCameraLEFT->m_uHeightPixel), m_fWidthPixels(m_uPixelWidth), m_fHeightPixels(m_uPixelHeight) {
        m_dBaselineMeters = std::sqrt(p_vecTranslationToRIGHT.x() * p_vecTranslationToRIGHT.x() + p_vecTranslationToRIGHT.y() * p_vecTranslationToRIGHT.y() +
                                      p_vecTranslationToRIGHT.z() * p_vecTranslationToRIGHT.z());
    }
    const std::shared_ptr<CPinholeCamera> m_pCameraLEFT, m_pCameraRIGHT;
    double m_dBaselineMeters;
    const uint32_t m_uPixelWidth, m_uPixelHeight;
    const float m_fWidthPixels, m_fHeightPixels;
};

class CParameterBase {
public:
    // CParameterBase.h:21-66 -- every non-empty line split on single spaces, flattened
    static const std::vector<std::string> getParametersFromFile(const std::string& p_strCameraConfigurationFile) {
        std::vector<std::string> vecParameters;
        std::ifstream ifConfiguration(p_strCameraConfigurationFile, std::ifstream::in);
        if (!ifConfiguration.is_open() || ifConfiguration.bad())
            throw CExceptionParameter("unable to open file: '" + p_strCameraConfigurationFile + "'");
        std::string strLineBuffer;
        while (std::getline(ifConfiguration, strLineBuffer)) {
            if (strLineBuffer.empty()) continue;
            std::string::size_type uLastStart = 0, uLastSeparator = strLineBuffer.find(' ', 0);
            while (std::string::npos != uLastSeparator) {
                vecParameters.push_back(strLineBuffer.substr(uLastStart, uLastSeparator - uLastStart));
                uLastStart = uLastSeparator + 1;
                uLastSeparator = strLineBuffer.find(' ', uLastStart);
            }
            vecParameters.push_back(strLineBuffer.substr(uLastStart));
        }
        return vecParameters;
    }
    static std::vector<std::string>::const_iterator find(const std::vector<std::string>& p, const std::string& name) {
        auto it = std::find(p.begin(), p.end(), name);
        if (p.end() == it) throw CExceptionParameter("cannot find parameter: " + name);
        return it;
    }
    static uint32_t getIntegerFromFile(const std::vector<std::string>& p, const std::string& name) { return std::stoul(*(find(p, name) + 1)); }
    static double getDoubleFromFile(const std::vector<std::string>& p, const std::string& name) { return std::stod(*(find(p, name) + 1)); }
    // getMatrixFromFile<uRows, uCols> :106-141: the uRows*uCols tokens after the key, row by row
    template <uint32_t uRows, uint32_t uCols>
    static std::array<double, uRows * uCols> getMatrixFromFile(const std::vector<std::string>& p, const std::string& name) {
        auto it = find(p, name);
        std::array<double, uRows * uCols> m{};
        for (uint32_t i = 0; i < uRows * uCols; ++i) {
            if (p.end() - it <= (std::ptrdiff_t)(i + 1)) throw std::out_of_range("not enough values for " + name);   // the reference reads past the end here
            m[i] = std::stod(*(it + 1 + i));
        }
        return m;
    }
    // loadCameraLEFT / loadCameraRIGHT :169-226 (THROWS CExceptionParameter, std::invalid_argument, std::out_of_range):
    // the same eight keys in the same order; a file without one of them is rejected like the reference rejects it
    static std::shared_ptr<CPinholeCamera> loadCamera(const std::string& p_strCameraConfigurationFile) {
        const std::vector<std::string> vecParameters(getParametersFromFile(p_strCameraConfigurationFile));
        if (vecParameters.empty()) throw CExceptionParameter("unable to open file: '" + p_strCameraConfigurationFile + "'");
        const std::string strCameraLabel(vecParameters.front());
        const uint32_t uWidthPixel = getIntegerFromFile(vecParameters, "uWidthPixels");
        const uint32_t uHeightPixel = getIntegerFromFile(vecParameters, "uHeightPixels");
        MatrixProjection matProjection;
        Matrix3d matIntrinsic, matRectification;
        Vector4d vecDistortionCoefficients;
        const auto aP = getMatrixFromFile<3, 4>(vecParameters, "matProjection");
        const auto aK = getMatrixFromFile<3, 3>(vecParameters, "matIntrinsic");
        const double dFocalLengthMeters = getDoubleFromFile(vecParameters, "dFocalLengthMeters");
        const auto aD = getMatrixFromFile<4, 1>(vecParameters, "vecDistortionCoefficients");
        const auto aR = getMatrixFromFile<3, 3>(vecParameters, "matRectification");
        std::copy(aP.begin(), aP.end(), matProjection.m);
        std::copy(aK.begin(), aK.end(), matIntrinsic.m);
        std::copy(aD.begin(), aD.end(), vecDistortionCoefficients.v);
        std::copy(aR.begin(), aR.end(), matRectification.m);
        return std::make_shared<CPinholeCamera>(strCameraLabel, uWidthPixel, uHeightPixel, matProjection, matIntrinsic, dFocalLengthMeters,
                                                vecDistortionCoefficients, matRectification);
    }
    static void loadCameraLEFT(const std::string& f) { pCameraLEFT = loadCamera(f); }
    static void loadCameraRIGHT(const std::string& f) { pCameraRIGHT = loadCamera(f); }
    static void constructCameraSTEREO(const CPoint3D& p_vecTranslationToRIGHT) {
        pCameraSTEREO = std::make_shared<CStereoCamera>(pCameraLEFT, pCameraRIGHT, p_vecTranslationToRIGHT);
    }
    static inline std::shared_ptr<CPinholeCamera> pCameraLEFT, pCameraRIGHT;
    static inline std::shared_ptr<CStereoCamera> pCameraSTEREO;
};

#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC pop_options
#endif
#endif
