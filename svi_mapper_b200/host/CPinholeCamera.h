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

struct CRect {   // cv::Rect
    int x, y, width, height;
    bool contains(const Point2f& p) const {   // cv::Rect::contains after Point2f -> Point2i (cvRound)
        const long px = std::lrint(p.x), py = std::lrint(p.y);
        return x <= px && px < x + width && y <= py && py < y + height;
    }
};

class CPinholeCamera {
public:
    CPinholeCamera(const std::string& p_strLabel, const uint32_t& p_uWidthPixels, const uint32_t& p_uHeightPixels,
                   const MatrixProjection& p_matProjection)
        : m_strCameraLabel(p_strLabel), m_uWidthPixel(p_uWidthPixels), m_uHeightPixel(p_uHeightPixels), m_matProjection(p_matProjection),
          m_dFxP(p_matProjection(0, 0)), m_dFyP(p_matProjection(1, 1)), m_dCxP(p_matProjection(0, 2)), m_dCyP(p_matProjection(1, 2)),
          m_iWidthPixel(p_uWidthPixels), m_iHeightPixel(p_uHeightPixels), m_dWidthPixels(p_uWidthPixels), m_dHeightPixels(p_uHeightPixels),
          m_fWidthPixels(p_uWidthPixels), m_fHeightPixels(p_uHeightPixels),
          m_cFieldOfView{28, 28, (int)p_uWidthPixels - 56, (int)p_uHeightPixels - 56} {}

    const std::string m_strCameraLabel;
    const uint32_t m_uWidthPixel, m_uHeightPixel;
    const MatrixProjection m_matProjection;
    const double m_dFxP, m_dFyP, m_dCxP, m_dCyP;
    const int32_t m_iWidthPixel, m_iHeightPixel;
    const double m_dWidthPixels, m_dHeightPixels;
    const float m_fWidthPixels, m_fHeightPixels;
    const CRect m_cFieldOfView;

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
    static MatrixProjection getMatrixProjectionFromFile(const std::vector<std::string>& p, const std::string& name) {
        auto it = find(p, name);
        MatrixProjection m;
        for (int i = 0; i < 12; ++i) m.m[i] = std::stod(*(it + 1 + i));
        return m;
    }
    // loadCameraLEFT / loadCameraRIGHT :169-226 (THROWS CExceptionParameter, std::invalid_argument, std::out_of_range)
    static std::shared_ptr<CPinholeCamera> loadCamera(const std::string& p_strCameraConfigurationFile) {
        const std::vector<std::string> vecParameters(getParametersFromFile(p_strCameraConfigurationFile));
        if (vecParameters.empty()) throw CExceptionParameter("unable to open file: '" + p_strCameraConfigurationFile + "'");
        return std::make_shared<CPinholeCamera>(vecParameters.front(), getIntegerFromFile(vecParameters, "uWidthPixels"),
                                                getIntegerFromFile(vecParameters, "uHeightPixels"),
                                                getMatrixProjectionFromFile(vecParameters, "matProjection"));
    }
    static void loadCameraLEFT(const std::string& f) { pCameraLEFT = loadCamera(f); }
    static void loadCameraRIGHT(const std::string& f) { pCameraRIGHT = loadCamera(f); }
    static void constructCameraSTEREO(const CPoint3D& p_vecTranslationToRIGHT) {
        pCameraSTEREO = std::make_shared<CStereoCamera>(pCameraLEFT, pCameraRIGHT, p_vecTranslationToRIGHT);
    }
    static inline std::shared_ptr<CPinholeCamera> pCameraLEFT, pCameraRIGHT;
    static inline std::shared_ptr<CStereoCamera> pCameraSTEREO;
};

#endif
