// Host side of path B behind the C-ABI: class-Bundle-shaped ingest (AddCamera/AddPoint/AddMeas),
// device graph construction (CSR by point), and the LM control loop of Bundle::Compute /
// Do_LM_Step (reference src/Bundle.cc:116-158,209-551).  The host reads back two scalars per
// lambda trial (new error, squared update) to take the accept / reject / converge decisions exactly
// where the reference takes them; everything else stays on the device.
#include "bundle_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

using namespace ptam;

ptam::CamModel ptam_make_cam_model(const double* p, double W, double H);
void ptam_set_global_error(const std::string& e);

namespace {
template <class T>
struct Buf {
  T* p = nullptr;
  cudaError_t alloc(size_t n) {
    release();
    if (!n) n = 1;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(p, 0, n * sizeof(T));
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; }
};
}  // namespace

struct ptam_bundle {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  ptam_bundle_params prm{};
  CamModel cam{};
  // host-side graph (insertion order, like the reference's containers)
  std::vector<double> h_cam_se3, h_pts, h_found, h_sin;
  std::vector<int> h_cam_fixed, h_cam_row, h_mcam, h_mpt;
  int start_row = 0, n_free = 0;
  // LM state (Bundle.h:130-139)
  double sigma_sq = 0, lambda = 0, lambda_factor = 0, trial_lambda = 0;
  bool converged = false, hit_max = false, begun = false;
  int counter = 0, accepted = 0, lm_steps = 0, n_outliers = 0;
  double last_error = 0, last_new_error = 0;
  bool s_mirrored = false;
  // device
  BundleDev d{};
  Buf<double> cam_se3, cam_se3_new, U, epsA, pt_pos, pt_pos_new, V, epsB, Vinv, Ve, m_found, m_sin, m_v3cam, m_derivs,
      m_eps, m_e2, m_W, e2c, S, vE, upd, scal, Wp;
  Buf<int> cam_fixed, cam_row, pt_off, pt_meas, m_cam, m_pt, m_state, counters, outliers;
  double* h_scal = nullptr;  // pinned
  int* h_cnt = nullptr;      // pinned
  // multi-GPU shard
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;

  void set_error(const std::string& e) { err = e; }

  ~ptam_bundle() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (auto* b : {&cam_se3, &cam_se3_new, &U, &epsA, &pt_pos, &pt_pos_new, &V, &epsB, &Vinv, &Ve, &m_found, &m_sin,
                    &m_v3cam, &m_derivs, &m_eps, &m_e2, &m_W, &e2c, &S, &vE, &upd, &scal, &Wp})
      b->release();
    for (auto* b : {&cam_fixed, &cam_row, &pt_off, &pt_meas, &m_cam, &m_pt, &m_state, &counters, &outliers}) b->release();
    if (h_scal) cudaFreeHost(h_scal);
    if (h_cnt) cudaFreeHost(h_cnt);
    if (stream) cudaStreamDestroy(stream);
  }

  int init(int dev, const double* cam_params, int w, int h, const ptam_bundle_params* p) {
    device = dev;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: the B200 path has no CPU fallback"); return PTAM_ERR_NO_DEVICE; }
    if (dev < 0 || dev >= ndev) { set_error("bad device index"); return PTAM_ERR_INVALID; }
    PTAM_CUDA_TRY(this, cudaSetDevice(dev));
    PTAM_CUDA_TRY(this, cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    PTAM_CUDA_TRY(this, cudaMallocHost(&h_scal, 8 * sizeof(double)));
    PTAM_CUDA_TRY(this, cudaMallocHost(&h_cnt, 4 * sizeof(int)));
    if (p) prm = *p; else ptam_bundle_default_params(&prm);
    cam = ptam_make_cam_model(cam_params, w, h);
    return PTAM_OK;
  }

  int n_cams() const { return (int)h_cam_fixed.size(); }
  int n_pts() const { return (int)h_pts.size() / 3; }
  int n_meas() const { return (int)h_mcam.size(); }

  // Compute() before its loop: GenerateMeasLUTs / GenerateOffDiagScripts become a CSR by point with
  // each point's measurements sorted by camera id (std::set<int> order, Bundle.h:69).
  int begin() {
    cudaSetDevice(device);
    const int C = n_cams(), P = n_pts(), M = n_meas();
    const int n = 6 * n_free;
    std::vector<int> off(P + 1, 0), idx(M);
    for (int m = 0; m < M; m++) off[h_mpt[m] + 1]++;
    for (int i = 0; i < P; i++) off[i + 1] += off[i];
    {
      std::vector<int> cur(off.begin(), off.end() - 1);
      for (int m = 0; m < M; m++) idx[cur[h_mpt[m]]++] = m;
      for (int i = 0; i < P; i++) {
        std::sort(idx.begin() + off[i], idx.begin() + off[i + 1], [&](int a, int b) { return h_mcam[a] < h_mcam[b]; });
        for (int o = off[i] + 1; o < off[i + 1]; o++)
          if (h_mcam[idx[o]] == h_mcam[idx[o - 1]]) { set_error("duplicate (camera, point) measurement"); return PTAM_ERR_INVALID; }
      }
    }
#define AL(buf, cnt) PTAM_CUDA_TRY(this, buf.alloc(cnt))
    AL(cam_se3, 12 * (size_t)C); AL(cam_se3_new, 12 * (size_t)C); AL(U, 21 * (size_t)C); AL(epsA, 6 * (size_t)C);
    AL(cam_fixed, C); AL(cam_row, C);
    AL(pt_pos, 3 * (size_t)P); AL(pt_pos_new, 3 * (size_t)P); AL(V, 6 * (size_t)P); AL(epsB, 3 * (size_t)P);
    AL(Vinv, 9 * (size_t)P); AL(Ve, 3 * (size_t)P); AL(pt_off, P + 1); AL(pt_meas, M);
    AL(m_cam, M); AL(m_pt, M); AL(m_found, 2 * (size_t)M); AL(m_sin, M); AL(m_state, M); AL(m_v3cam, 3 * (size_t)M);
    AL(m_derivs, 4 * (size_t)M); AL(m_eps, 2 * (size_t)M); AL(m_e2, M); AL(m_W, 18 * (size_t)M); AL(e2c, M);
    AL(S, (size_t)n * n); AL(vE, n); AL(upd, n); AL(scal, 8); AL(counters, 4); AL(outliers, 2 * (size_t)M);
    AL(Wp, (size_t)n * kNB);
#undef AL
#define UP(buf, vec) if (!vec.empty()) PTAM_CUDA_TRY(this, cudaMemcpy(buf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice))
    UP(cam_se3, h_cam_se3); UP(cam_fixed, h_cam_fixed); UP(cam_row, h_cam_row); UP(pt_pos, h_pts);
    UP(pt_off, off); UP(pt_meas, idx); UP(m_cam, h_mcam); UP(m_pt, h_mpt); UP(m_found, h_found); UP(m_sin, h_sin);
#undef UP
    d.cam = cam; d.n_cams = C; d.n_pts = P; d.n_meas = M; d.n = n; d.est = prm.mestimator;
    d.cam_se3 = cam_se3.p; d.cam_se3_new = cam_se3_new.p; d.cam_fixed = cam_fixed.p; d.cam_row = cam_row.p;
    d.U = U.p; d.epsA = epsA.p; d.pt_pos = pt_pos.p; d.pt_pos_new = pt_pos_new.p; d.V = V.p; d.epsB = epsB.p;
    d.Vinv = Vinv.p; d.Ve = Ve.p; d.pt_off = pt_off.p; d.pt_meas = pt_meas.p; d.m_cam = m_cam.p; d.m_pt = m_pt.p;
    d.m_found = m_found.p; d.m_sin = m_sin.p; d.m_state = m_state.p; d.m_v3cam = m_v3cam.p; d.m_derivs = m_derivs.p;
    d.m_eps = m_eps.p; d.m_e2 = m_e2.p; d.m_W = m_W.p; d.e2_compact = e2c.p; d.S = S.p; d.vE = vE.p; d.upd = upd.p;
    d.scal = scal.p; d.counters = counters.p; d.outliers = outliers.p;
    lambda = 0.0001; lambda_factor = 2.0;
    converged = false; hit_max = false;
    counter = 0; accepted = 0; lm_steps = 0; n_outliers = 0;
    begun = true;
    return PTAM_OK;
  }

  int solve_reduced() {
    const int n = d.n;
    if (n == 0) return PTAM_OK;
    for (int k0 = 0; k0 < n; k0 += kNB) {
      const int nb = std::min(kNB, n - k0);
      k_ldlt_diag<<<1, 32, 0, stream>>>(d.S, n, k0);
      launches++;
      const int rem = n - k0 - nb;
      if (rem > 0) {
        k_ldlt_panel<<<(rem + 127) / 128, 128, 0, stream>>>(d.S, Wp.p, n, k0);
        const int nt = (rem + kUT - 1) / kUT;
        k_ldlt_update<<<nt * (nt + 1) / 2, 256, 0, stream>>>(d.S, Wp.p, n, k0);
        launches += 2;
      }
    }
    k_ldlt_solve<<<1, 1024, n * sizeof(double), stream>>>(d.S, d.vE, d.upd, n);
    launches++;
    PTAM_CUDA_TRY(this, cudaGetLastError());
    return PTAM_OK;
  }

  // Do_LM_Step (Bundle.cc:209-551)
  int lm_step(const volatile unsigned char* abort_flag) {
    cudaSetDevice(device);
    if (!begun) { set_error("ptam_bundle_begin() has not been called"); return PTAM_ERR_INVALID; }
    auto aborted = [&]() { return abort_flag && *abort_flag; };
    const int C = d.n_cams, P = d.n_pts, M = d.n_meas, n = d.n;
    lm_steps++;
    // ClearAccumulators + error / counter scalars
    PTAM_CUDA_TRY(this, cudaMemsetAsync(U.p, 0, sizeof(double) * 21 * C, stream));
    PTAM_CUDA_TRY(this, cudaMemsetAsync(epsA.p, 0, sizeof(double) * 6 * C, stream));
    PTAM_CUDA_TRY(this, cudaMemsetAsync(V.p, 0, sizeof(double) * 6 * P, stream));
    PTAM_CUDA_TRY(this, cudaMemsetAsync(epsB.p, 0, sizeof(double) * 3 * P, stream));
    PTAM_CUDA_TRY(this, cudaMemsetAsync(scal.p, 0, sizeof(double) * 8, stream));
    PTAM_CUDA_TRY(this, cudaMemsetAsync(counters.p, 0, sizeof(int), stream));
    if (M > 0) {
      const int gm = (M + 255) / 256;
      k_ba_project<<<gm, 256, 0, stream>>>(d);
      k_ba_gather_e2<<<gm, 256, 0, stream>>>(d);
      k_ba_select<<<1, 1024, 0, stream>>>(d, prm.min_tukey_sigma * prm.min_tukey_sigma);
      k_ba_jacobian<<<(M + 127) / 128, 128, 0, stream>>>(d);
      launches += 4;
    }
    PTAM_CUDA_TRY(this, cudaGetLastError());
    PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_scal, scal.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, stream));
    PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
    sigma_sq = h_scal[1];
    const double cur_err = h_scal[2];
    last_error = cur_err;
    double new_err = cur_err + 9999;
    while (new_err > cur_err && !converged && !hit_max && !aborted()) {
      // scal[3] new error, scal[4] squared update, scal[5] lambda
      trial_lambda = lambda;
      double init[3] = {0.0, 0.0, lambda};
      h_scal[3] = init[0]; h_scal[4] = init[1]; h_scal[5] = init[2];
      PTAM_CUDA_TRY(this, cudaMemcpyAsync(scal.p + 3, h_scal + 3, sizeof(double) * 3, cudaMemcpyHostToDevice, stream));
      if (P > 0) k_ba_vinv<<<(P + 255) / 256, 256, 0, stream>>>(d);
      if (n > 0) {
        k_ba_init_s<<<std::min<size_t>(((size_t)n * n + 255) / 256, 148 * 8), 256, 0, stream>>>(d);
        k_ba_init_diag<<<C, 64, 0, stream>>>(d);
        launches += 2;
      }
      if (P > 0) k_ba_schur<<<(P + 7) / 8, 256, 0, stream>>>(d);
      launches += 2;
      s_mirrored = false;
      int rc = solve_reduced();
      if (rc) return rc;
      if (P > 0) k_ba_point_update<<<(P + 255) / 256, 256, 0, stream>>>(d);
      if (C > 0) k_ba_cam_update<<<(C + 127) / 128, 128, 0, stream>>>(d);
      if (M > 0) k_ba_new_error<<<(M + 255) / 256, 256, 0, stream>>>(d);
      launches += 3;
      PTAM_CUDA_TRY(this, cudaGetLastError());
      PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_scal + 3, scal.p + 3, sizeof(double) * 2, cudaMemcpyDeviceToHost, stream));
      PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
      new_err = h_scal[3];
      last_new_error = new_err;
      if (h_scal[4] < prm.update_squared_convergence) converged = true;
      if (new_err > cur_err) { lambda = lambda * lambda_factor; lambda_factor = lambda_factor * 2; }  // ModifyLambda_BadStep
      counter++;
      if (counter >= prm.max_iterations) hit_max = true;
    }
    if (new_err < cur_err) {  // ModifyLambda_GoodStep + commit
      lambda_factor = 2.0; lambda *= 0.3;
      PTAM_CUDA_TRY(this, cudaMemcpyAsync(cam_se3.p, cam_se3_new.p, sizeof(double) * 12 * C, cudaMemcpyDeviceToDevice, stream));
      PTAM_CUDA_TRY(this, cudaMemcpyAsync(pt_pos.p, pt_pos_new.p, sizeof(double) * 3 * P, cudaMemcpyDeviceToDevice, stream));
      accepted++;
    }
    if (M > 0) { k_ba_erase<<<1, 1024, 0, stream>>>(d); launches++; }
    PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_cnt, counters.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, stream));
    PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
    n_outliers = h_cnt[1];
    return PTAM_OK;
  }
};

extern "C" {

void ptam_bundle_default_params(ptam_bundle_params* p) {
  p->max_iterations = 20; p->mestimator = 0; p->update_squared_convergence = 1e-6; p->min_tukey_sigma = 0.4;
}

ptam_bundle* ptam_bundle_create(int device, const double* cam_params, int width, int height, const ptam_bundle_params* params) {
  ptam_bundle* b = new ptam_bundle;
  if (b->init(device, cam_params, width, height, params) != PTAM_OK) { ptam_set_global_error(b->err); delete b; return nullptr; }
  return b;
}
void ptam_bundle_destroy(ptam_bundle* b) { delete b; }
const char* ptam_bundle_last_error(const ptam_bundle* b) { return b->err.c_str(); }

int ptam_bundle_add_camera(ptam_bundle* b, const double* se3, int fixed) {
  b->h_cam_se3.insert(b->h_cam_se3.end(), se3, se3 + 12);
  b->h_cam_fixed.push_back(fixed ? 1 : 0);
  if (!fixed) { b->h_cam_row.push_back(b->start_row); b->start_row += 6; b->n_free++; }
  else b->h_cam_row.push_back(-1);
  b->begun = false;
  return b->n_cams() - 1;
}
int ptam_bundle_add_point(ptam_bundle* b, const double* xyz) {
  double v[3] = {xyz[0], xyz[1], xyz[2]};
  if (std::isnan(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])) v[0] = v[1] = v[2] = 0;
  b->h_pts.insert(b->h_pts.end(), v, v + 3);
  b->begun = false;
  return b->n_pts() - 1;
}
int ptam_bundle_add_meas(ptam_bundle* b, int cam, int point, const double* uv, double sigma_sq) {
  if (cam < 0 || cam >= b->n_cams() || point < 0 || point >= b->n_pts()) { b->set_error("measurement references unknown camera/point"); return PTAM_ERR_INVALID; }
  b->h_mcam.push_back(cam); b->h_mpt.push_back(point);
  b->h_found.push_back(uv[0]); b->h_found.push_back(uv[1]);
  b->h_sin.push_back(std::sqrt(1.0 / sigma_sq));
  b->begun = false;
  return PTAM_OK;
}
int ptam_bundle_add_cameras(ptam_bundle* b, int n, const double* se3, const int32_t* fixed) {
  for (int i = 0; i < n; i++) ptam_bundle_add_camera(b, se3 + 12 * i, fixed[i]);
  return PTAM_OK;
}
int ptam_bundle_add_points(ptam_bundle* b, int n, const double* xyz) {
  for (int i = 0; i < n; i++) ptam_bundle_add_point(b, xyz + 3 * i);
  return PTAM_OK;
}
int ptam_bundle_add_measurements(ptam_bundle* b, int n, const int32_t* cam, const int32_t* point, const double* uv, const double* s2) {
  for (int i = 0; i < n; i++) {
    int rc = ptam_bundle_add_meas(b, cam[i], point[i], uv + 2 * i, s2[i]);
    if (rc) return rc;
  }
  return PTAM_OK;
}

int ptam_bundle_set_shard(ptam_bundle* b, int rank, int world, void* comm) {
  if (world < 1 || rank < 0 || rank >= world) { b->set_error("bad shard description"); return PTAM_ERR_INVALID; }
  if (world > 1) { b->set_error("sharded bundle adjustment is not available in this build"); return PTAM_ERR_NCCL; }
  b->rank = rank; b->world = world; b->nccl_comm = comm;
  return PTAM_OK;
}

int ptam_bundle_begin(ptam_bundle* b) { return b->begin(); }
int ptam_bundle_lm_step(ptam_bundle* b, const volatile unsigned char* abort_flag) { return b->lm_step(abort_flag); }

int ptam_bundle_compute(ptam_bundle* b, const volatile unsigned char* abort_flag) {  // Bundle.cc:116-158
  int rc = b->begin();
  if (rc) return rc;
  while (!b->converged && !b->hit_max && !(abort_flag && *abort_flag)) {
    rc = b->lm_step(abort_flag);
    if (rc) return rc;
  }
  return b->accepted;
}

int ptam_bundle_converged(const ptam_bundle* b) { return b->converged ? 1 : 0; }

int ptam_bundle_get_points(ptam_bundle* b, double* xyz) {
  cudaSetDevice(b->device);
  if (!b->begun) { std::memcpy(xyz, b->h_pts.data(), b->h_pts.size() * sizeof(double)); return PTAM_OK; }
  PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  PTAM_CUDA_TRY(b, cudaMemcpy(xyz, b->pt_pos.p, sizeof(double) * 3 * b->d.n_pts, cudaMemcpyDeviceToHost));
  return PTAM_OK;
}
int ptam_bundle_get_cameras(ptam_bundle* b, double* se3) {
  cudaSetDevice(b->device);
  if (!b->begun) { std::memcpy(se3, b->h_cam_se3.data(), b->h_cam_se3.size() * sizeof(double)); return PTAM_OK; }
  PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  PTAM_CUDA_TRY(b, cudaMemcpy(se3, b->cam_se3.p, sizeof(double) * 12 * b->d.n_cams, cudaMemcpyDeviceToHost));
  return PTAM_OK;
}
int ptam_bundle_get_point(ptam_bundle* b, int n, double* xyz) {
  cudaSetDevice(b->device);
  if (n < 0 || n >= b->n_pts()) { b->set_error("bad point id"); return PTAM_ERR_INVALID; }
  if (!b->begun) { std::memcpy(xyz, b->h_pts.data() + 3 * n, 24); return PTAM_OK; }
  PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  PTAM_CUDA_TRY(b, cudaMemcpy(xyz, b->pt_pos.p + 3 * n, 24, cudaMemcpyDeviceToHost));
  return PTAM_OK;
}
int ptam_bundle_get_camera(ptam_bundle* b, int n, double* se3) {
  cudaSetDevice(b->device);
  if (n < 0 || n >= b->n_cams()) { b->set_error("bad camera id"); return PTAM_ERR_INVALID; }
  if (!b->begun) { std::memcpy(se3, b->h_cam_se3.data() + 12 * n, 96); return PTAM_OK; }
  PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  PTAM_CUDA_TRY(b, cudaMemcpy(se3, b->cam_se3.p + 12 * n, 96, cudaMemcpyDeviceToHost));
  return PTAM_OK;
}
int ptam_bundle_get_outliers(ptam_bundle* b, int32_t* pairs, int cap) {
  cudaSetDevice(b->device);
  if (!b->begun) return 0;
  PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  const int n = b->n_outliers;
  if (pairs && cap > 0 && n > 0)
    PTAM_CUDA_TRY(b, cudaMemcpy(pairs, b->outliers.p, sizeof(int) * 2 * std::min(n, cap), cudaMemcpyDeviceToHost));
  return n;
}
int ptam_bundle_get_stats(ptam_bundle* b, ptam_bundle_stats* s) {
  s->accepted = b->accepted; s->lambda_trials = b->counter; s->lm_steps = b->lm_steps;
  s->converged = b->converged; s->hit_max_iterations = b->hit_max; s->n_outliers = b->n_outliers;
  s->sigma_squared = b->sigma_sq; s->lambda = b->lambda; s->last_error = b->last_error; s->last_new_error = b->last_new_error;
  return PTAM_OK;
}
int ptam_bundle_get_reduced_system(ptam_bundle* b, double* S, double* vE, int cap_n) {
  // After a solve the lower triangle of S holds the LDL^T factors, so S and vE are re-assembled
  // here from the accumulators of the last LM step with the lambda of its last trial.
  cudaSetDevice(b->device);
  if (!b->begun) return 0;
  const int n = b->d.n;
  if (cap_n < n || n == 0) return n;
  // re-assemble S, vE for the current lambda / accumulators (same kernels as the LM loop)
  double l = b->trial_lambda;
  PTAM_CUDA_TRY(b, cudaMemcpyAsync(b->scal.p + 5, &l, sizeof(double), cudaMemcpyHostToDevice, b->stream));
  PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  if (b->d.n_pts > 0) k_ba_vinv<<<(b->d.n_pts + 255) / 256, 256, 0, b->stream>>>(b->d);
  k_ba_init_s<<<std::min<size_t>(((size_t)n * n + 255) / 256, 148 * 8), 256, 0, b->stream>>>(b->d);
  k_ba_init_diag<<<b->d.n_cams, 64, 0, b->stream>>>(b->d);
  if (b->d.n_pts > 0) k_ba_schur<<<(b->d.n_pts + 7) / 8, 256, 0, b->stream>>>(b->d);
  k_ba_mirror<<<std::min<size_t>(((size_t)n * n + 255) / 256, 148 * 8), 256, 0, b->stream>>>(b->d.S, n);
  b->launches += 5;
  PTAM_CUDA_TRY(b, cudaGetLastError());
  PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  PTAM_CUDA_TRY(b, cudaMemcpy2D(S, sizeof(double) * cap_n, b->S.p, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyDeviceToHost));
  PTAM_CUDA_TRY(b, cudaMemcpy(vE, b->vE.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return n;
}
int ptam_bundle_synchronize(ptam_bundle* b) {
  cudaSetDevice(b->device);
  PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  return PTAM_OK;
}
void* ptam_bundle_cuda_stream(ptam_bundle* b) { return (void*)b->stream; }
int64_t ptam_bundle_launch_count(const ptam_bundle* b) { return b->launches; }

}  // extern "C"
