// TEST INFRASTRUCTURE — CPU oracle for PTAM path B (bundle adjuster).  NOT part of the product.
// PINNED against the reference's own src/Bundle.cc + src/ATANCamera.cc compiled in place (oracle/_ref,
// Makefile.ref): bit-identical runs (tests/test_ref_pin_bundle.py).  Restated, not pinned: TooN's own
// arithmetic (LDL^T Cholesky, SE3::exp) in oracle_math.h — the library is absent from this image.
// Faithful single-threaded restatement of class Bundle — src/Bundle.cc (all), include/Bundle.h:38-103
// — including its data structures' cost: the dense [camera][point] measurement LUT
// (Bundle.cc:558-567) and the O(cameras x points) scans (Bundle.cc:396,466) ARE the reference's CPU
// cost and are kept, because this oracle is also the timed CPU baseline.
// The exported orc_bundle_* functions have the signatures of ptam_bundle_* (include/ptam_b200.h).
#include "oracle_math.h"
#include "../include/ptam_b200.h"
#include <list>
#include <set>
#include <string>
#include <utility>

namespace orc {

struct BCamera {
  bool fixed;
  SE3 cfw, cfw_new;
  double U[36];
  double epsA[6];
  int start_row;
};
struct OffDiag { int j, k; };
struct BPoint {
  double pos[3] = {0, 0, 0}, pos_new[3] = {0, 0, 0};
  double V[9] = {0}, epsB[3] = {0}, VStarInv[9] = {0};
  int n_meas = 0, n_outliers = 0;
  std::set<int> cams;
  std::vector<OffDiag> script;
};
struct BMeas {
  int p, c;
  bool bad = false;
  double found[2], eps[2];
  double A[12];  // 2x6
  double B[6];   // 2x3
  double W[18];  // 6x3
  double sqrt_inv_noise;
  double v3cam[3];
  double err_sq;
  double derivs[4];
};

struct Bundle {
  Camera cam;
  ptam_bundle_params prm;
  std::vector<BPoint> pts;
  std::vector<BCamera> cams;
  std::list<BMeas> meas;
  std::vector<std::pair<int, int>> outliers;
  std::vector<std::vector<BMeas*>> lut;
  int n_cams_to_update = 0, start_row = 0;
  double sigma_sq = 0, lambda = 0, lambda_factor = 0;
  bool converged = false, hit_max = false;
  int counter = 0, accepted = 0, lm_steps = 0;
  double last_error = 0, last_new_error = 0;
  std::vector<double> S, vE;  // last reduced system
  std::string err;

  void begin() {  // Compute() up to its loop (Bundle.cc:116-131)
    lut.clear();
    for (size_t c = 0; c < cams.size(); c++) lut.emplace_back(pts.size(), nullptr);
    for (auto& m : meas) lut[m.c][m.p] = &m;
    for (size_t i = 0; i < pts.size(); i++) {  // GenerateOffDiagScripts (Bundle.cc:572-599)
      BPoint& p = pts[i];
      p.script.clear();
      for (auto itj = p.cams.begin(); itj != p.cams.end(); ++itj) {
        const int j = *itj;
        if (cams[j].fixed) continue;
        for (auto itk = p.cams.begin(); itk != itj; ++itk) {
          const int k = *itk;
          if (cams[k].fixed) continue;
          p.script.push_back({j, k});
        }
      }
    }
    lambda = 0.0001; lambda_factor = 2.0;
    converged = false; hit_max = false;
    counter = 0; accepted = 0; lm_steps = 0;
  }

  void project_and_error(BMeas& m) {  // Bundle.cc:164-180
    const BCamera& c = cams[m.c];
    const BPoint& p = pts[m.p];
    c.cfw.apply(p.pos, m.v3cam);
    if (m.v3cam[2] <= 0) { m.bad = true; return; }
    m.bad = false;
    double ip[2] = {m.v3cam[0] / m.v3cam[2], m.v3cam[1] / m.v3cam[2]};
    Camera::Proj q = cam.project(ip);
    cam.derivs(q, m.derivs);
    m.eps[0] = m.sqrt_inv_noise * (m.found[0] - q.im[0]);
    m.eps[1] = m.sqrt_inv_noise * (m.found[1] - q.im[1]);
    m.err_sq = m.eps[0] * m.eps[0] + m.eps[1] * m.eps[1];
  }

  double find_new_error() {  // Bundle.cc:188-207
    double e = 0;
    for (auto& m : meas) {
      double v3[3];
      cams[m.c].cfw_new.apply(pts[m.p].pos_new, v3);
      if (v3[2] <= 0) { e += 1.0; continue; }
      double ip[2] = {v3[0] / v3[2], v3[1] / v3[2]};
      Camera::Proj q = cam.project(ip);
      const double e0 = m.sqrt_inv_noise * (m.found[0] - q.im[0]), e1 = m.sqrt_inv_noise * (m.found[1] - q.im[1]);
      e += mest_objective(e0 * e0 + e1 * e1, sigma_sq, prm.mestimator);
    }
    return e;
  }

  bool lm_step(const volatile unsigned char* abort_flag) {  // Do_LM_Step (Bundle.cc:209-551)
    auto aborted = [&]() { return abort_flag && *abort_flag; };
    const int est = prm.mestimator;
    lm_steps++;
    for (auto& p : pts) { std::memset(p.V, 0, sizeof p.V); std::memset(p.epsB, 0, sizeof p.epsB); }
    for (auto& c : cams) { std::memset(c.U, 0, sizeof c.U); std::memset(c.epsA, 0, sizeof c.epsA); }
    std::vector<double> e2;
    for (auto& m : meas) {
      project_and_error(m);
      if (!m.bad) e2.push_back(m.err_sq);
    }
    sigma_sq = mest_find_sigma_squared(e2, est);
    const double min_s2 = prm.min_tukey_sigma * prm.min_tukey_sigma;
    if (sigma_sq < min_s2) sigma_sq = min_s2;

    double cur_err = 0.0;
    for (auto& m : meas) {
      BCamera& c = cams[m.c];
      BPoint& p = pts[m.p];
      if (m.bad) { cur_err += 1.0; continue; }
      const double w = mest_sqrt_weight(m.err_sq, sigma_sq, est);
      m.eps[0] = w * m.eps[0]; m.eps[1] = w * m.eps[1];
      if (w == 0) { m.bad = true; cur_err += 1.0; continue; }
      cur_err += mest_objective(m.err_sq, sigma_sq, est);
      double d[4];
      for (int k = 0; k < 4; k++) d[k] = w * m.derivs[k];
      const double ooz = 1.0 / m.v3cam[2];
      const double X = m.v3cam[0], Y = m.v3cam[1], Z = m.v3cam[2];
      if (c.fixed) std::memset(m.A, 0, sizeof m.A);
      else {
        const double g[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, -Z, Y}, {Z, 0, -X}, {-Y, X, 0}};
        for (int mm = 0; mm < 6; mm++) {
          const double a0 = (g[mm][0] - X * g[mm][2] * ooz) * ooz;
          const double a1 = (g[mm][1] - Y * g[mm][2] * ooz) * ooz;
          // meas.dSqrtInvNoise * m2CamDerivs * v2CamFrameMotion: (s * M) * v
          m.A[mm] = (m.sqrt_inv_noise * d[0]) * a0 + (m.sqrt_inv_noise * d[1]) * a1;
          m.A[6 + mm] = (m.sqrt_inv_noise * d[2]) * a0 + (m.sqrt_inv_noise * d[3]) * a1;
        }
      }
      for (int mm = 0; mm < 3; mm++) {
        const double mo[3] = {c.cfw.R[mm], c.cfw.R[3 + mm], c.cfw.R[6 + mm]};  // column mm of R
        const double a0 = (mo[0] - X * mo[2] * ooz) * ooz;
        const double a1 = (mo[1] - Y * mo[2] * ooz) * ooz;
        m.B[mm] = (m.sqrt_inv_noise * d[0]) * a0 + (m.sqrt_inv_noise * d[1]) * a1;
        m.B[3 + mm] = (m.sqrt_inv_noise * d[2]) * a0 + (m.sqrt_inv_noise * d[3]) * a1;
      }
      if (!c.fixed) {
        for (int r = 0; r < 6; r++)
          for (int cc = 0; cc <= r; cc++) c.U[6 * r + cc] += m.A[r] * m.A[cc] + m.A[6 + r] * m.A[6 + cc];
        for (int r = 0; r < 6; r++) c.epsA[r] += m.A[r] * m.eps[0] + m.A[6 + r] * m.eps[1];
      }
      for (int r = 0; r < 3; r++)
        for (int cc = 0; cc <= r; cc++) p.V[3 * r + cc] += m.B[r] * m.B[cc] + m.B[3 + r] * m.B[3 + cc];
      for (int r = 0; r < 3; r++) p.epsB[r] += m.B[r] * m.eps[0] + m.B[3 + r] * m.eps[1];
      if (c.fixed) std::memset(m.W, 0, sizeof m.W);
      else
        for (int r = 0; r < 6; r++)
          for (int cc = 0; cc < 3; cc++) m.W[3 * r + cc] = m.A[r] * m.B[cc] + m.A[6 + r] * m.B[3 + cc];
    }
    last_error = cur_err;

    const int n = n_cams_to_update * 6;
    double new_err = cur_err + 9999;
    while (new_err > cur_err && !converged && !hit_max && !aborted()) {
      for (auto& p : pts) {  // V* inverse (Bundle.cc:341-359)
        double Vs[9];
        std::memcpy(Vs, p.V, sizeof Vs);
        if (Vs[0] * Vs[4] * Vs[8] == 0) std::memset(p.VStarInv, 0, sizeof p.VStarInv);
        else {
          Vs[1] = Vs[3]; Vs[2] = Vs[6]; Vs[5] = Vs[7];
          for (int i = 0; i < 3; i++) Vs[4 * i] *= (1.0 + lambda);
          ldlt_inverse(Vs, 3, p.VStarInv);
        }
      }
      S.assign((size_t)n * n, 0.0);
      vE.assign(n, 0.0);
      for (size_t j = 0; j < cams.size(); j++) {  // diagonal blocks (Bundle.cc:374-406)
        BCamera& cj = cams[j];
        if (cj.fixed) continue;
        double m6[36], v6[6];
        for (int r = 0; r < 6; r++) {
          for (int c = 0; c < r; c++) m6[6 * r + c] = m6[6 * c + r] = cj.U[6 * r + c];
          m6[7 * r] = cj.U[7 * r];
        }
        for (int q = 0; q < 6; q++) m6[7 * q] *= (1.0 + lambda);
        for (int q = 0; q < 6; q++) v6[q] = cj.epsA[q];
        std::vector<BMeas*>& lj = lut[j];
        for (size_t i = 0; i < pts.size(); i++) {
          BMeas* pm = lj[i];
          if (pm == nullptr || pm->bad) continue;
          const double* Vi = pts[i].VStarInv;
          double WV[18];  // W * V*inv (6x3)
          for (int r = 0; r < 6; r++)
            for (int c = 0; c < 3; c++) WV[3 * r + c] = pm->W[3 * r] * Vi[c] + pm->W[3 * r + 1] * Vi[3 + c] + pm->W[3 * r + 2] * Vi[6 + c];
          for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++) m6[6 * r + c] -= WV[3 * r] * pm->W[3 * c] + WV[3 * r + 1] * pm->W[3 * c + 1] + WV[3 * r + 2] * pm->W[3 * c + 2];
          double Ve[3];
          for (int r = 0; r < 3; r++) Ve[r] = Vi[3 * r] * pts[i].epsB[0] + Vi[3 * r + 1] * pts[i].epsB[1] + Vi[3 * r + 2] * pts[i].epsB[2];
          for (int r = 0; r < 6; r++) v6[r] -= pm->W[3 * r] * Ve[0] + pm->W[3 * r + 1] * Ve[1] + pm->W[3 * r + 2] * Ve[2];
        }
        for (int r = 0; r < 6; r++)
          for (int c = 0; c < 6; c++) S[(size_t)(cj.start_row + r) * n + cj.start_row + c] = m6[6 * r + c];
        for (int r = 0; r < 6; r++) vE[cj.start_row + r] = v6[r];
      }
      for (size_t i = 0; i < pts.size(); i++) {  // off-diagonal blocks (Bundle.cc:410-446)
        BPoint& p = pts[i];
        int cur_j = -1, jrow = -1;
        BMeas* mij = nullptr;
        double WV[18];
        for (auto& e : p.script) {
          BMeas* mik = lut[e.k][i];
          if (mik == nullptr || mik->bad) continue;
          if (e.j != cur_j) {
            mij = lut[e.j][i];
            if (mij == nullptr || mij->bad) continue;
            cur_j = e.j;
            jrow = cams[e.j].start_row;
            for (int r = 0; r < 6; r++)
              for (int c = 0; c < 3; c++)
                WV[3 * r + c] = mij->W[3 * r] * p.VStarInv[c] + mij->W[3 * r + 1] * p.VStarInv[3 + c] + mij->W[3 * r + 2] * p.VStarInv[6 + c];
          }
          const int krow = cams[mik->c].start_row;
          for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++)
              S[(size_t)(jrow + r) * n + krow + c] -= WV[3 * r] * mik->W[3 * c] + WV[3 * r + 1] * mik->W[3 * c + 1] + WV[3 * r + 2] * mik->W[3 * c + 2];
        }
      }
      for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) S[(size_t)j * n + i] = S[(size_t)i * n + j];
      // Cholesky<>(mS).backsub(vE)  (Bundle.cc:457-458)
      std::vector<double> L(S), upd(n, 0.0);
      if (n > 0) { ldlt_factor(L.data(), n, n); ldlt_backsub(L.data(), n, n, vE.data(), upd.data()); }
      std::vector<double> mapupd(pts.size() * 3);
      for (size_t i = 0; i < pts.size(); i++) {  // Bundle.cc:461-483
        double sum[3] = {0, 0, 0};
        for (size_t j = 0; j < cams.size(); j++) {
          BCamera& c = cams[j];
          if (c.fixed) continue;
          BMeas* pm = lut[j][i];
          if (pm == nullptr || pm->bad) continue;
          for (int r = 0; r < 3; r++) {
            double a = 0;
            for (int q = 0; q < 6; q++) a += pm->W[3 * q + r] * upd[c.start_row + q];
            sum[r] += a;
          }
        }
        double v3[3];
        for (int r = 0; r < 3; r++) v3[r] = pts[i].epsB[r] - sum[r];
        const double* Vi = pts[i].VStarInv;
        for (int r = 0; r < 3; r++) mapupd[3 * i + r] = Vi[3 * r] * v3[0] + Vi[3 * r + 1] * v3[1] + Vi[3 * r + 2] * v3[2];
      }
      double ssu = 0;
      for (int i = 0; i < n; i++) ssu += upd[i] * upd[i];
      double ssm = 0;
      for (double v : mapupd) ssm += v * v;
      if (ssu + ssm < prm.update_squared_convergence) converged = true;
      for (auto& c : cams) {
        if (c.fixed) c.cfw_new = c.cfw;
        else c.cfw_new = se3_mul(se3_exp(&upd[c.start_row]), c.cfw);
      }
      for (size_t i = 0; i < pts.size(); i++)
        for (int r = 0; r < 3; r++) pts[i].pos_new[r] = pts[i].pos[r] + mapupd[3 * i + r];
      new_err = find_new_error();
      last_new_error = new_err;
      if (new_err > cur_err) { lambda = lambda * lambda_factor; lambda_factor = lambda_factor * 2; }
      counter++;
      if (counter >= prm.max_iterations) hit_max = true;
    }
    if (new_err < cur_err) {
      lambda_factor = 2.0; lambda *= 0.3;
      for (auto& c : cams) c.cfw = c.cfw_new;
      for (auto& p : pts) std::memcpy(p.pos, p.pos_new, sizeof p.pos);
      accepted++;
    }
    for (auto it = meas.begin(); it != meas.end();) {  // ditch the outliers (Bundle.cc:536-547)
      if (it->bad) {
        outliers.push_back({it->p, it->c});
        pts[it->p].n_outliers++;
        lut[it->c][it->p] = nullptr;
        it = meas.erase(it);
      } else ++it;
    }
    return true;
  }

  int compute(const volatile unsigned char* abort_flag) {
    begin();
    while (!converged && !hit_max && !(abort_flag && *abort_flag))
      if (!lm_step(abort_flag)) return -1;
    return accepted;
  }
};

}  // namespace orc

using orc::Bundle;

extern "C" {

void orc_bundle_default_params(ptam_bundle_params* p) {
  p->max_iterations = 20; p->mestimator = 0; p->update_squared_convergence = 1e-6; p->min_tukey_sigma = 0.4;
}
void* orc_bundle_create(int, const double* cam_params, int w, int h, const ptam_bundle_params* prm) {
  Bundle* b = new Bundle;
  b->cam.init(cam_params, w, h);
  if (prm) b->prm = *prm; else orc_bundle_default_params(&b->prm);
  return b;
}
void orc_bundle_destroy(void* b) { delete (Bundle*)b; }
const char* orc_bundle_last_error(const void* b) { return ((const Bundle*)b)->err.c_str(); }

int orc_bundle_add_camera(void* bp, const double* se3, int fixed) {  // Bundle.cc:46-63
  Bundle* b = (Bundle*)bp;
  orc::BCamera c{};
  c.fixed = fixed != 0;
  c.cfw = orc::SE3::from12(se3);
  if (!c.fixed) { c.start_row = b->start_row; b->start_row += 6; b->n_cams_to_update++; }
  else c.start_row = -999999999;
  b->cams.push_back(c);
  return (int)b->cams.size() - 1;
}
int orc_bundle_add_point(void* bp, const double* xyz) {  // Bundle.cc:66-78
  Bundle* b = (Bundle*)bp;
  orc::BPoint p;
  double v[3] = {xyz[0], xyz[1], xyz[2]};
  if (std::isnan(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])) v[0] = v[1] = v[2] = 0;
  std::memcpy(p.pos, v, sizeof v);
  b->pts.push_back(p);
  return (int)b->pts.size() - 1;
}
int orc_bundle_add_meas(void* bp, int cam, int point, const double* uv, double sigma_sq) {  // Bundle.cc:81-93
  Bundle* b = (Bundle*)bp;
  if (cam < 0 || cam >= (int)b->cams.size() || point < 0 || point >= (int)b->pts.size()) return PTAM_ERR_INVALID;
  b->pts[point].n_meas++;
  b->pts[point].cams.insert(cam);
  orc::BMeas m;
  m.p = point; m.c = cam;
  m.found[0] = uv[0]; m.found[1] = uv[1];
  m.sqrt_inv_noise = std::sqrt(1.0 / sigma_sq);
  b->meas.push_back(m);
  return 0;
}
int orc_bundle_add_cameras(void* b, int n, const double* se3, const int32_t* fixed) {
  for (int i = 0; i < n; i++) orc_bundle_add_camera(b, se3 + 12 * i, fixed[i]);
  return 0;
}
int orc_bundle_add_points(void* b, int n, const double* xyz) {
  for (int i = 0; i < n; i++) orc_bundle_add_point(b, xyz + 3 * i);
  return 0;
}
int orc_bundle_add_measurements(void* b, int n, const int32_t* cam, const int32_t* point, const double* uv, const double* s2) {
  for (int i = 0; i < n; i++) {
    int rc = orc_bundle_add_meas(b, cam[i], point[i], uv + 2 * i, s2[i]);
    if (rc) return rc;
  }
  return 0;
}
int orc_bundle_set_shard(void*, int, int world, void*) { return world == 1 ? 0 : PTAM_ERR_INVALID; }
// the reference asserts on an empty measurement list (Tools.h:155); both libraries report it as an error
int orc_bundle_compute(void* b, const volatile unsigned char* abort_flag) {
  if (((Bundle*)b)->meas.empty()) { ((Bundle*)b)->err = "no measurements"; return PTAM_ERR_INVALID; }
  return ((Bundle*)b)->compute(abort_flag);
}
int orc_bundle_begin(void* b) {
  if (((Bundle*)b)->meas.empty()) { ((Bundle*)b)->err = "no measurements"; return PTAM_ERR_INVALID; }
  ((Bundle*)b)->begin(); return 0;
}
int orc_bundle_lm_step(void* b, const volatile unsigned char* abort_flag) { return ((Bundle*)b)->lm_step(abort_flag) ? 0 : -1; }
// persistent graph: Compute again on the state the previous Compute left (erased measurements stay erased)
int orc_bundle_recompute(void* bp, const volatile unsigned char* abort_flag) {
  Bundle* b = (Bundle*)bp;
  if (b->meas.empty()) { b->err = "no measurements"; return PTAM_ERR_INVALID; }
  b->outliers.clear();
  return b->compute(abort_flag);
}
int orc_bundle_update_camera(void* bp, int n, const double* se3) {
  Bundle* b = (Bundle*)bp;
  if (n < 0 || n >= (int)b->cams.size()) return PTAM_ERR_INVALID;
  b->cams[n].cfw = orc::SE3::from12(se3); return 0;
}
int orc_bundle_update_point(void* bp, int n, const double* xyz) {
  Bundle* b = (Bundle*)bp;
  if (n < 0 || n >= (int)b->pts.size()) return PTAM_ERR_INVALID;
  double v[3] = {xyz[0], xyz[1], xyz[2]};
  if (std::isnan(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])) v[0] = v[1] = v[2] = 0;
  std::memcpy(b->pts[n].pos, v, sizeof v); return 0;
}
int orc_bundle_converged(const void* b) { return ((const Bundle*)b)->converged; }
int orc_bundle_get_point(void* bp, int n, double* xyz) {
  Bundle* b = (Bundle*)bp;
  if (n < 0 || n >= (int)b->pts.size()) return PTAM_ERR_INVALID;
  std::memcpy(xyz, b->pts[n].pos, 24); return 0;
}
int orc_bundle_get_camera(void* bp, int n, double* se3) {
  Bundle* b = (Bundle*)bp;
  if (n < 0 || n >= (int)b->cams.size()) return PTAM_ERR_INVALID;
  b->cams[n].cfw.to12(se3); return 0;
}
int orc_bundle_get_points(void* bp, double* xyz) {
  Bundle* b = (Bundle*)bp;
  for (size_t i = 0; i < b->pts.size(); i++) std::memcpy(xyz + 3 * i, b->pts[i].pos, 24);
  return 0;
}
int orc_bundle_get_cameras(void* bp, double* se3) {
  Bundle* b = (Bundle*)bp;
  for (size_t i = 0; i < b->cams.size(); i++) b->cams[i].cfw.to12(se3 + 12 * i);
  return 0;
}
int orc_bundle_get_outliers(void* bp, int32_t* pairs, int cap) {
  Bundle* b = (Bundle*)bp;
  for (size_t i = 0; i < b->outliers.size() && (int)i < cap; i++) { pairs[2 * i] = b->outliers[i].first; pairs[2 * i + 1] = b->outliers[i].second; }
  return (int)b->outliers.size();
}
int orc_bundle_get_stats(void* bp, ptam_bundle_stats* s) {
  Bundle* b = (Bundle*)bp;
  s->accepted = b->accepted; s->lambda_trials = b->counter; s->lm_steps = b->lm_steps;
  s->converged = b->converged; s->hit_max_iterations = b->hit_max; s->n_outliers = (int)b->outliers.size();
  s->sigma_squared = b->sigma_sq; s->lambda = b->lambda; s->last_error = b->last_error; s->last_new_error = b->last_new_error;
  return 0;
}
int orc_bundle_get_reduced_system(void* bp, double* S, double* vE, int cap_n) {
  Bundle* b = (Bundle*)bp;
  const int n = b->n_cams_to_update * 6;
  if (cap_n < n || b->S.size() != (size_t)n * n) return n;
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) S[(size_t)i * cap_n + j] = b->S[(size_t)i * n + j];
    vE[i] = b->vE[i];
  }
  return n;
}
int orc_bundle_synchronize(void*) { return 0; }
}  // extern "C"
