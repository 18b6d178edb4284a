// Host mirror of class ATANCamera (reference include/ATANCamera.h:66-135, src/ATANCamera.cc:27-140,
// 179-209): the FOV / arctan radial-distortion pinhole model.  It exists at the boundary because
// Bundle and Tracker are constructed from one (Bundle.h:110, Tracker.h:158); the device code receives
// only its five parameters and image size and evaluates the model itself (csrc/common.cuh).
// GVars3 is absent: the parameter vector ("Camera.Parameters", ATANCamera.cc:15) is passed in.
#pragma once
#include <algorithm>
#include <string>
#include "toon_cvd_shim.h"

namespace ptam_b200 {

class ATANCamera {
 public:
  // reference: ATANCamera(std::string sName) + gvar lookup; here the 5-vector is explicit.
  // Defaults: the reference's shipped calibration (config/camera.cfg:7) at 640x480.
  explicit ATANCamera(const std::string& sName = "Camera",
                      const TooN::Vector<5>& params = TooN::makeVector(1.0803, 1.43987, 0.519983, 0.548655, 0.244943),
                      CVD::ImageRef irSize = CVD::ImageRef(640, 480))
      : msName(sName), mvParams(params) {
    mvImageSize[0] = irSize.x; mvImageSize[1] = irSize.y;
    RefreshParams();
  }
  void SetImageSize(TooN::Vector<2> v2) { mvImageSize = v2; RefreshParams(); }
  void SetImageSize(CVD::ImageRef ir) { mvImageSize[0] = ir.x; mvImageSize[1] = ir.y; RefreshParams(); }
  TooN::Vector<2> GetImageSize() const { return mvImageSize; }
  const TooN::Vector<5>& GetParams() const { return mvParams; }

  void RefreshParams() {  // ATANCamera.cc:27-105 (the part the projection functions need)
    mvFocal[0] = mvImageSize[0] * mvParams[0]; mvFocal[1] = mvImageSize[1] * mvParams[1];
    mvCenter[0] = mvImageSize[0] * mvParams[2] - 0.5; mvCenter[1] = mvImageSize[1] * mvParams[3] - 0.5;
    mvInvFocal[0] = 1.0 / mvFocal[0]; mvInvFocal[1] = 1.0 / mvFocal[1];
    mdW = mvParams[4];
    if (mdW != 0.0) { md2Tan = 2.0 * std::tan(mdW / 2.0); mdOneOver2Tan = 1.0 / md2Tan; mdWinv = 1.0 / mdW; mdDistortionEnabled = 1.0; }
    else { mdWinv = 0.0; md2Tan = 0.0; mdOneOver2Tan = 0.0; mdDistortionEnabled = 0.0; }
    const double v0 = std::max(mvParams[2], 1.0 - mvParams[2]) / mvParams[0];
    const double v1 = std::max(mvParams[3], 1.0 - mvParams[3]) / mvParams[1];
    mdLargestRadius = invrtrans(std::sqrt(v0 * v0 + v1 * v1));
    mdMaxR = 1.5 * mdLargestRadius;
  }

  TooN::Vector<2> Project(const TooN::Vector<2>& camframe) {  // ATANCamera.cc:109-121
    mvLastCam = camframe;
    mdLastR = std::sqrt(camframe * camframe);
    mbInvalid = mdLastR > mdMaxR;
    mdLastFactor = rtrans_factor(mdLastR);
    mvLastIm[0] = mvCenter[0] + mvFocal[0] * (mdLastFactor * camframe[0]);
    mvLastIm[1] = mvCenter[1] + mvFocal[1] * (mdLastFactor * camframe[1]);
    return mvLastIm;
  }
  TooN::Vector<2> UnProject(const TooN::Vector<2>& imframe) {  // ATANCamera.cc:125-140
    mvLastIm = imframe;
    TooN::Vector<2> dc;
    dc[0] = (imframe[0] - mvCenter[0]) * mvInvFocal[0];
    dc[1] = (imframe[1] - mvCenter[1]) * mvInvFocal[1];
    const double dr = std::sqrt(dc * dc);
    mdLastR = invrtrans(dr);
    const double f = dr > 0.01 ? mdLastR / dr : 1.0;
    mdLastFactor = 1.0 / f;
    mvLastCam = dc * f;
    return mvLastCam;
  }
  TooN::Matrix<2, 2> GetProjectionDerivs() {  // ATANCamera.cc:179-209; uses the state of the last Project
    const double k = md2Tan, x = mvLastCam[0], y = mvLastCam[1], r = mdLastR * mdDistortionEnabled;
    double fx = 0, fy = 0;
    if (r >= 0.01) {
      fx = mdWinv * (k * x) / (r * r * (1 + k * k * r * r)) - x * mdLastFactor / (r * r);
      fy = mdWinv * (k * y) / (r * r * (1 + k * k * r * r)) - y * mdLastFactor / (r * r);
    }
    TooN::Matrix<2, 2> m;
    m[0][0] = mvFocal[0] * (fx * x + mdLastFactor); m[0][1] = mvFocal[0] * (fy * x);
    m[1][0] = mvFocal[1] * (fx * y);                m[1][1] = mvFocal[1] * (fy * y + mdLastFactor);
    return m;
  }
  bool Invalid() const { return mbInvalid; }
  double LargestRadiusInImage() const { return mdLargestRadius; }

 private:
  double rtrans_factor(double r) const { return (r < 0.001 || mdW == 0.0) ? 1.0 : mdWinv * std::atan(r * md2Tan) / r; }
  double invrtrans(double r) const { return mdW == 0.0 ? r : std::tan(r * mdW) * mdOneOver2Tan; }

  std::string msName;
  TooN::Vector<5> mvParams;
  TooN::Vector<2> mvImageSize, mvFocal, mvCenter, mvInvFocal, mvLastCam, mvLastIm;
  double mdW = 0, mdWinv = 0, md2Tan = 0, mdOneOver2Tan = 0, mdDistortionEnabled = 0;
  double mdLargestRadius = 0, mdMaxR = 0, mdLastR = 0, mdLastFactor = 1;
  bool mbInvalid = false;
};

}  // namespace ptam_b200
