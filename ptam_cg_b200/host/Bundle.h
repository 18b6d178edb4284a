// Host mirror of class Bundle (reference include/Bundle.h:105-156): the same public methods with
// the same argument meaning and return conventions, forwarding to the C-ABI of
// include/ptam_b200.h.  All arithmetic of Bundle::Compute (src/Bundle.cc:116-551) runs in the CUDA
// library; this class only marshals TooN values.  MapMaker::BundleAdjust (src/MapMaker.cc:838-933)
// compiles against it unchanged.
#pragma once
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "../../include/ptam_b200.h"
#include "ATANCamera.h"

namespace ptam_b200 {

class Bundle {
 public:
  // Bundle::Bundle(const ATANCamera&) — Bundle.cc:35-43.  `device` / `params` are the only additions
  // (GVars3 keys Bundle.MaxIterations / UpdateSquaredConvergenceLimit / MEstimator / MinTukeySigma).
  explicit Bundle(const ATANCamera& TCam, int device = 0, const ptam_bundle_params* params = nullptr) {
    double p[5];
    for (int i = 0; i < 5; i++) p[i] = TCam.GetParams()[i];
    h = ptam_bundle_create(device, p, (int)TCam.GetImageSize()[0], (int)TCam.GetImageSize()[1], params);
    if (!h) throw std::runtime_error(std::string("ptam_bundle_create: ") + ptam_global_last_error());
  }
  ~Bundle() { if (h) ptam_bundle_destroy(h); }
  Bundle(const Bundle&) = delete;
  Bundle& operator=(const Bundle&) = delete;

  int AddCamera(TooN::SE3<> se3CamFromWorld, bool bFixed) {  // Bundle.cc:46-63
    double a[12];
    se3_to_array(se3CamFromWorld, a);
    return ptam_bundle_add_camera(h, a, bFixed ? 1 : 0);
  }
  int AddPoint(TooN::Vector<3> v3Pos) {  // Bundle.cc:66-78
    const double a[3] = {v3Pos[0], v3Pos[1], v3Pos[2]};
    return ptam_bundle_add_point(h, a);
  }
  void AddMeas(int nCam, int nPoint, TooN::Vector<2> v2Pos, double dSigmaSquared) {  // Bundle.cc:81-93
    const double a[2] = {v2Pos[0], v2Pos[1]};
    // the reference asserts on bad ids (Bundle.cc:83-84); here it is an exception
    if (ptam_bundle_add_meas(h, nCam, nPoint, a, dSigmaSquared) != PTAM_OK) throw std::out_of_range(ptam_bundle_last_error(h));
    mnMeas++;
  }
  int NumMeasurements() const { return mnMeas; }                      // not in the reference: lets a caller skip an empty adjustment
  const char* LastError() const { return ptam_bundle_last_error(h); } // not in the reference: text behind a negative Compute()
  // Returns the number of accepted update iterations, or negative on error (Bundle.h:114).
  int Compute(bool* pbAbortSignal) {
    static_assert(sizeof(bool) == 1, "abort flag is polled as one byte");
    return ptam_bundle_compute(h, reinterpret_cast<const volatile unsigned char*>(pbAbortSignal));
  }
  // Not in the reference: Compute again on the device-resident graph — what
  // MapMaker::BundleAdjustAll does by rebuilding an identical Bundle until it converges (MapMaker.cc:67-77).
  int Recompute(bool* pbAbortSignal) {
    return ptam_bundle_recompute(h, reinterpret_cast<const volatile unsigned char*>(pbAbortSignal));
  }
  void UpdateCamera(int n, TooN::SE3<> se3CamFromWorld) { double a[12]; se3_to_array(se3CamFromWorld, a); ptam_bundle_update_camera(h, n, a); }
  void UpdatePoint(int n, TooN::Vector<3> v3Pos) { const double a[3] = {v3Pos[0], v3Pos[1], v3Pos[2]}; ptam_bundle_update_point(h, n, a); }
  bool Converged() { return ptam_bundle_converged(h) != 0; }
  TooN::Vector<3> GetPoint(int n) {
    double a[3] = {0, 0, 0};
    ptam_bundle_get_point(h, n, a);
    return TooN::makeVector(a[0], a[1], a[2]);
  }
  TooN::SE3<> GetCamera(int n) {
    double a[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    ptam_bundle_get_camera(h, n, a);
    return se3_from_array(a);
  }
  std::vector<std::pair<int, int> > GetOutlierMeasurements() {  // (point, camera) pairs
    const int n = ptam_bundle_get_outliers(h, nullptr, 0);
    std::vector<int32_t> raw(2 * (size_t)std::max(n, 0));
    if (n > 0) ptam_bundle_get_outliers(h, raw.data(), n);
    std::vector<std::pair<int, int> > out;
    for (int i = 0; i < n; i++) out.emplace_back(raw[2 * i], raw[2 * i + 1]);
    return out;
  }
  ptam_bundle* handle() { return h; }

 private:
  ptam_bundle* h = nullptr;
  int mnMeas = 0;
};

}  // namespace ptam_b200
