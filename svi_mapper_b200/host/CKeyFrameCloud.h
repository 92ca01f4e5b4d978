// Key-frame cloud files and KITTI ground-truth poses: the data formats on either side of the front-end
// (SURVEY.md 8f rank 4).
//   .cloud  binary, native endianness, written field by field (src/types/CKeyFrame.cpp:138-186, read back :186-270):
//           16 f64 LEFTtoWORLD (row by row) | u64 number of points | per point: 3 f64 world position, 3 f64 camera
//           position, 4 f64 (uL vL uR vR), u64 number of descriptors, that many 32-byte BRIEF descriptors.
//   poses   KITTI odometry format, one line of 12 numbers per frame = rows of the 3x4 LEFTtoWORLD
//           (src/runnable/tracker_gt.cpp:208-229); the tracker is fed inverse(T_now) * T_last per frame.
// The thin_drivers txt_io message log (republisher_kitti.cpp:28-192) is serialised by an un-vendored library and
// is not restated here; the front-end takes plain 8-bit buffers (ImageView).
#ifndef SVI_HOST_CKEYFRAMECLOUD_H
#define SVI_HOST_CKEYFRAMECLOUD_H

#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "CFundamentalMatcher.h"

struct CDescriptorVectorPoint3DWORLD {   // src/types/Types.h (cloud point with its descriptor history)
    UIDLandmark uID = 0;
    CPoint3DWORLD vecPointXYZWORLD;
    CPoint3DCAMERA vecPointXYZCAMERA;
    double ptUVLEFT[2] = {0, 0}, ptUVRIGHT[2] = {0, 0};   // cv::Point2d
    std::vector<CDescriptor> vecDescriptors;
};

namespace cloud_io {
template <class T> inline void writeDatum(std::ostream& os, const T& v) { os.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <class T> inline void readDatum(std::istream& is, T& v) {
    if (!is.read(reinterpret_cast<char*>(&v), sizeof(T))) throw std::invalid_argument("truncated cloud file");
}
}   // namespace cloud_io

// CKeyFrame::saveCloudToFile
inline void saveCloudToFile(const std::string& p_strFile, const Isometry3d& p_matTransformationLEFTtoWORLD,
                            const std::vector<CDescriptorVectorPoint3DWORLD>& p_vecCloud) {
    std::ofstream ofCloud(p_strFile, std::ofstream::out | std::ofstream::binary);
    if (!ofCloud.good()) throw std::invalid_argument("cannot write cloud file");
    for (int i = 0; i < 16; ++i) cloud_io::writeDatum(ofCloud, p_matTransformationLEFTtoWORLD.m[i]);
    cloud_io::writeDatum(ofCloud, (uint64_t)p_vecCloud.size());
    for (const CDescriptorVectorPoint3DWORLD& cPoint : p_vecCloud) {
        for (int k = 0; k < 3; ++k) cloud_io::writeDatum(ofCloud, cPoint.vecPointXYZWORLD.v[k]);
        for (int k = 0; k < 3; ++k) cloud_io::writeDatum(ofCloud, cPoint.vecPointXYZCAMERA.v[k]);
        cloud_io::writeDatum(ofCloud, cPoint.ptUVLEFT[0]);
        cloud_io::writeDatum(ofCloud, cPoint.ptUVLEFT[1]);
        cloud_io::writeDatum(ofCloud, cPoint.ptUVRIGHT[0]);
        cloud_io::writeDatum(ofCloud, cPoint.ptUVRIGHT[1]);
        cloud_io::writeDatum(ofCloud, (uint64_t)cPoint.vecDescriptors.size());
        for (const CDescriptor& d : cPoint.vecDescriptors) ofCloud.write(reinterpret_cast<const char*>(d.data()), DESCRIPTOR_SIZE_BYTES);
    }
}

// CKeyFrame::getCloudFromFile
inline std::vector<CDescriptorVectorPoint3DWORLD> getCloudFromFile(const std::string& p_strFile, Isometry3d& p_matTransformationLEFTtoWORLD) {
    std::ifstream ifCloud(p_strFile, std::ifstream::in | std::ifstream::binary);
    if (!ifCloud.is_open()) throw std::invalid_argument("invalid cloud file");
    for (int i = 0; i < 16; ++i) cloud_io::readDatum(ifCloud, p_matTransformationLEFTtoWORLD.m[i]);
    uint64_t uNumberOfPoints = 0;
    cloud_io::readDatum(ifCloud, uNumberOfPoints);
    std::vector<CDescriptorVectorPoint3DWORLD> vecPoints;
    for (uint64_t u = 0; u < uNumberOfPoints; ++u) {
        CDescriptorVectorPoint3DWORLD cPoint;
        cPoint.uID = u;   // ids are positional in the file (:262)
        for (int k = 0; k < 3; ++k) cloud_io::readDatum(ifCloud, cPoint.vecPointXYZWORLD.v[k]);
        for (int k = 0; k < 3; ++k) cloud_io::readDatum(ifCloud, cPoint.vecPointXYZCAMERA.v[k]);
        cloud_io::readDatum(ifCloud, cPoint.ptUVLEFT[0]);
        cloud_io::readDatum(ifCloud, cPoint.ptUVLEFT[1]);
        cloud_io::readDatum(ifCloud, cPoint.ptUVRIGHT[0]);
        cloud_io::readDatum(ifCloud, cPoint.ptUVRIGHT[1]);
        uint64_t uNumberOfDescriptors = 0;
        cloud_io::readDatum(ifCloud, uNumberOfDescriptors);
        if (uNumberOfDescriptors > (1u << 24)) throw std::invalid_argument("corrupt cloud file");
        cPoint.vecDescriptors.resize(uNumberOfDescriptors);
        for (CDescriptor& d : cPoint.vecDescriptors)
            if (!ifCloud.read(reinterpret_cast<char*>(d.data()), DESCRIPTOR_SIZE_BYTES)) throw std::invalid_argument("truncated cloud file");
        vecPoints.push_back(cPoint);
    }
    return vecPoints;
}

// CFundamentalMatcher::getCloudForVisibleOptimizedLandmarks :293-319 (free function over the matcher's visible landmarks)
inline std::vector<CDescriptorVectorPoint3DWORLD> getCloudForVisibleOptimizedLandmarks(const std::vector<CLandmark*>& p_vecVisibleLandmarks) {
    std::vector<CDescriptorVectorPoint3DWORLD> vecCloud;
    for (const CLandmark* pLandmark : p_vecVisibleLandmarks) {
        if (!pLandmark->bIsOptimal) continue;
        CDescriptorVectorPoint3DWORLD cPoint;
        cPoint.uID = pLandmark->uID;
        cPoint.vecPointXYZWORLD = pLandmark->vecPointXYZOptimized;
        cPoint.vecPointXYZCAMERA = pLandmark->getLastMeasurement()->vecPointXYZLEFT;
        cPoint.ptUVLEFT[0] = pLandmark->getLastDetectionLEFT().x; cPoint.ptUVLEFT[1] = pLandmark->getLastDetectionLEFT().y;
        cPoint.ptUVRIGHT[0] = pLandmark->getLastDetectionRIGHT().x; cPoint.ptUVRIGHT[1] = pLandmark->getLastDetectionRIGHT().y;
        cPoint.vecDescriptors = pLandmark->vecDescriptorsLEFT;
        vecCloud.push_back(cPoint);
    }
    return vecCloud;
}

// KITTI odometry ground truth: LEFTtoWORLD per frame
inline std::vector<Isometry3d> readPosesKITTI(const std::string& p_strFile) {
    std::ifstream ifGroundTruth(p_strFile);
    if (!ifGroundTruth.is_open()) throw std::invalid_argument("invalid pose file");
    std::vector<Isometry3d> vecPoses;
    std::string strLineBuffer;
    while (std::getline(ifGroundTruth, strLineBuffer)) {
        if (strLineBuffer.empty()) continue;
        std::istringstream issLine(strLineBuffer);
        Isometry3d T;
        for (int i = 0; i < 12; ++i)
            if (!(issLine >> T.m[i])) throw std::invalid_argument("malformed pose line");
        vecPoses.push_back(T);
    }
    return vecPoses;
}
// what tracker_gt.cpp:229 hands to CTrackerGT::process for frame i: LEFT(i-1) -> LEFT(i)
inline Isometry3d getTransformationLEFTLASTtoLEFTNOW(const Isometry3d& p_matLEFTtoWORLDNOW, const Isometry3d& p_matLEFTtoWORLDLAST) {
    return inverseIsometry(p_matLEFTtoWORLDNOW) * p_matLEFTtoWORLDLAST;
}

#endif
