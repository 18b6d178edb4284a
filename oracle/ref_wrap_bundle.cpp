// TEST INFRASTRUCTURE — C ABI around the reference's OWN Bundle class (src/Bundle.cc +
// src/ATANCamera.cc, compiled in place from /root/reference against the header stand-ins in
// oracle/shim/), with the same signatures as the oracle (orc_bundle_*) and the product
// (ptam_bundle_*), prefix ref_.  Used to pin the oracle: tests/test_ref_pin.py feeds the same graphs
// to both and compares.  Output goes to oracle/_ref/libref_ptam.so only; nothing here ships.
#include "Bundle.h"
#include "../include/ptam_b200.h"
#include <gvars3/instances.h>
#include <cstring>
#include <string>

namespace {
struct RefBundle : public Bundle {
  explicit RefBundle(const ATANCamera& c) : Bundle(c) {}
  using Bundle::mbHitMaxIterations;
  using Bundle::mdLambda;
  using Bundle::mdSigmaSquared;
  using Bundle::mnAccepted;
  using Bundle::mnCounter;
  using Bundle::mvCameras;
  using Bundle::mvPoints;
  using Bundle::mMeasList;
};
struct Handle {
  RefBundle* b = nullptr;
  int n_computes = 0;
  size_t outlier_base = 0;  // mvOutlierMeasurementIdx is never cleared: outliers of earlier Compute calls
  std::string err;
  ~Handle() { delete b; }
};
const char* kEstimators[3] = {"Tukey", "Cauchy", "Huber"};
}  // namespace

extern "C" {

void ref_bundle_default_params(ptam_bundle_params* p) {
  p->max_iterations = 20; p->mestimator = 0; p->update_squared_convergence = 1e-6; p->min_tukey_sigma = 0.4;
}
void* ref_bundle_create(int, const double* cam_params, int w, int h, const ptam_bundle_params* prm) {
  ptam_bundle_params p;
  if (prm) p = *prm; else ref_bundle_default_params(&p);
  // the reference reads its settings from GVars3 (Bundle.cc:40-42,126,233; ATANCamera.cc:17)
  TooN::Vector<5> cp;
  for (int i = 0; i < 5; i++) cp[i] = cam_params[i];
  GVars3::GV3::set<TooN::Vector<5>>("Camera.Parameters", cp);
  GVars3::GV3::set<int>("Bundle.MaxIterations", p.max_iterations);
  GVars3::GV3::set<double>("Bundle.UpdateSquaredConvergenceLimit", p.update_squared_convergence);
  GVars3::GV3::set<double>("Bundle.MinTukeySigma", p.min_tukey_sigma);
  GVars3::GV3::set<std::string>("Bundle.MEstimator", kEstimators[p.mestimator >= 0 && p.mestimator < 3 ? p.mestimator : 0]);
  GVars3::GV3::set<int>("Bundle.Cout", 0);
  ATANCamera cam("Camera");
  cam.SetImageSize(TooN::makeVector((double)w, (double)h));
  Handle* hd = new Handle;
  hd->b = new RefBundle(cam);
  return hd;
}
void ref_bundle_destroy(void* h) { delete (Handle*)h; }
const char* ref_bundle_last_error(const void* h) { return ((const Handle*)h)->err.c_str(); }

static TooN::SE3<> se3_from12(const double* p) {
  TooN::Vector<3> t = TooN::makeVector(p[9], p[10], p[11]);
  TooN::Matrix<3> R;
  for (int i = 0; i < 9; i++) R(i / 3, i % 3) = p[i];
  return TooN::SE3<>(TooN::SE3<>::raw(R), t);  // bits preserved: no re-orthonormalisation
}
int ref_bundle_add_camera(void* h, const double* se3, int fixed) { return ((Handle*)h)->b->AddCamera(se3_from12(se3), fixed != 0); }
int ref_bundle_add_point(void* h, const double* xyz) { return ((Handle*)h)->b->AddPoint(TooN::makeVector(xyz[0], xyz[1], xyz[2])); }
int ref_bundle_add_meas(void* h, int cam, int point, const double* uv, double sigma_sq) {
  Handle* hd = (Handle*)h;
  if (cam < 0 || cam >= (int)hd->b->mvCameras.size() || point < 0 || point >= (int)hd->b->mvPoints.size()) return PTAM_ERR_INVALID;
  hd->b->AddMeas(cam, point, TooN::makeVector(uv[0], uv[1]), sigma_sq);
  return 0;
}
int ref_bundle_add_cameras(void* h, int n, const double* se3, const int32_t* fixed) {
  for (int i = 0; i < n; i++) ref_bundle_add_camera(h, se3 + 12 * i, fixed[i]);
  return 0;
}
int ref_bundle_add_points(void* h, int n, const double* xyz) {
  for (int i = 0; i < n; i++) ref_bundle_add_point(h, xyz + 3 * i);
  return 0;
}
int ref_bundle_add_measurements(void* h, int n, const int32_t* cam, const int32_t* point, const double* uv, const double* s2) {
  for (int i = 0; i < n; i++) {
    const int rc = ref_bundle_add_meas(h, cam[i], point[i], uv + 2 * i, s2[i]);
    if (rc) return rc;
  }
  return 0;
}
int ref_bundle_set_shard(void*, int, int world, void*) { return world == 1 ? 0 : PTAM_ERR_INVALID; }
int ref_bundle_compute(void* h, const volatile unsigned char* abort_flag) {
  Handle* hd = (Handle*)h;
  if (hd->b->mMeasList.empty()) { hd->err = "no measurements (the reference asserts, Tools.h:155)"; return PTAM_ERR_INVALID; }
  bool never = false;
  hd->n_computes++;
  // Bundle::Compute polls a bool (Bundle.cc:134,338); the C ABI's flag is a byte of the same meaning
  return hd->b->Compute(abort_flag ? (bool*)const_cast<unsigned char*>(abort_flag) : &never);
}
// Do_LM_Step is a protected template defined in Bundle.cc: the reference cannot be stepped from outside
// the reference object can simply be computed again: Compute resets the LM control (Bundle.cc:121-126) and
// the erased measurements are gone from mMeasList
int ref_bundle_recompute(void* h, const volatile unsigned char* abort_flag) {
  Handle* hd = (Handle*)h;
  hd->outlier_base = hd->b->GetOutlierMeasurements().size();
  return ref_bundle_compute(h, abort_flag);
}
int ref_bundle_update_camera(void* h, int n, const double* se3) {
  Handle* hd = (Handle*)h;
  if (n < 0 || n >= (int)hd->b->mvCameras.size()) return PTAM_ERR_INVALID;
  hd->b->mvCameras[n].se3CfW = se3_from12(se3); return 0;
}
int ref_bundle_update_point(void* h, int n, const double* xyz) {
  Handle* hd = (Handle*)h;
  if (n < 0 || n >= (int)hd->b->mvPoints.size()) return PTAM_ERR_INVALID;
  hd->b->mvPoints[n].v3Pos = TooN::makeVector(xyz[0], xyz[1], xyz[2]); return 0;
}
int ref_bundle_begin(void* h) { ((Handle*)h)->err = "the reference has no step-wise interface"; return PTAM_ERR_INVALID; }
int ref_bundle_lm_step(void* h, const volatile unsigned char*) { ((Handle*)h)->err = "the reference has no step-wise interface"; return PTAM_ERR_INVALID; }
int ref_bundle_converged(const void* h) { return ((const Handle*)h)->b->Converged(); }
int ref_bundle_get_point(void* h, int n, double* xyz) {
  Handle* hd = (Handle*)h;
  if (n < 0 || n >= (int)hd->b->mvPoints.size()) return PTAM_ERR_INVALID;
  const TooN::Vector<3> v = hd->b->GetPoint(n);
  for (int i = 0; i < 3; i++) xyz[i] = v[i];
  return 0;
}
int ref_bundle_get_camera(void* h, int n, double* se3) {
  Handle* hd = (Handle*)h;
  if (n < 0 || n >= (int)hd->b->mvCameras.size()) return PTAM_ERR_INVALID;
  const TooN::SE3<> s = hd->b->GetCamera(n);
  for (int i = 0; i < 9; i++) se3[i] = s.get_rotation().get_matrix()(i / 3, i % 3);
  for (int i = 0; i < 3; i++) se3[9 + i] = s.get_translation()[i];
  return 0;
}
int ref_bundle_get_points(void* h, double* xyz) {
  Handle* hd = (Handle*)h;
  for (size_t i = 0; i < hd->b->mvPoints.size(); i++) ref_bundle_get_point(h, (int)i, xyz + 3 * i);
  return 0;
}
int ref_bundle_get_cameras(void* h, double* se3) {
  Handle* hd = (Handle*)h;
  for (size_t i = 0; i < hd->b->mvCameras.size(); i++) ref_bundle_get_camera(h, (int)i, se3 + 12 * i);
  return 0;
}
int ref_bundle_get_outliers(void* h, int32_t* pairs, int cap) {
  const std::vector<std::pair<int, int>> o = ((Handle*)h)->b->GetOutlierMeasurements();
  const size_t base = ((Handle*)h)->outlier_base;
  for (size_t i = base; i < o.size() && (int)(i - base) < cap; i++) { pairs[2 * (i - base)] = o[i].first; pairs[2 * (i - base) + 1] = o[i].second; }
  return (int)(o.size() - base);
}
int ref_bundle_get_stats(void* h, ptam_bundle_stats* s) {
  Handle* hd = (Handle*)h;
  std::memset(s, 0, sizeof *s);
  s->accepted = hd->b->mnAccepted; s->lambda_trials = hd->b->mnCounter; s->lm_steps = -1;  // not observable
  s->converged = hd->b->Converged(); s->hit_max_iterations = hd->b->mbHitMaxIterations;
  s->n_outliers = (int)(hd->b->GetOutlierMeasurements().size() - hd->outlier_base);
  s->sigma_squared = hd->b->mdSigmaSquared; s->lambda = hd->b->mdLambda;
  return 0;
}
int ref_bundle_get_reduced_system(void*, double*, double*, int) { return 0; }  // mS is a local of Do_LM_Step
int ref_bundle_synchronize(void*) { return 0; }
}  // extern "C"
