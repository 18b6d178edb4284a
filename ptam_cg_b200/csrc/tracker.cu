// Host side of path T behind the C-ABI (include/ptam_b200.h): owns the device buffers, builds the
// level geometry, enqueues the kernels of tracker_kernels.cuh on the handle's stream.
// One TrackFrame for a whole batch of streams = 8 kernel launches, no host round trip in between;
// the only D2H is the per-stream result block at the end (skipped when results == NULL).
#include "tracker_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

using namespace ptam;

static thread_local std::string g_last_error;

namespace {
// k_pvs_select: 64 registers per thread, so 512 threads put two trackers on an SM (one wave for two streams per SM)
inline int pvs_threads() {
  static const int n = [] { const char* e = std::getenv("PTAM_B200_PVS_THREADS"); const int v = e ? std::atoi(e) : 512; return v >= 32 && v <= 1024 && v % 32 == 0 ? v : 512; }();
  return n;
}
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    free();
    n = count;
    if (!count) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(p, 0, count * sizeof(T));
    // cudaMemset runs on the legacy default stream, which the library's non-blocking streams do not
    // synchronise with: without this a kernel queued right after a late allocation (the epipolar / re-find /
    // MakeKeyFrame_Rest buffers) could see its output zeroed behind its back
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    return e;
  }
  void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

CamModel make_cam(const double* p, double W, double H) {  // ATANCamera::RefreshParams
  CamModel c;
  c.img_w = W; c.img_h = H;
  c.focal[0] = W * p[0]; c.focal[1] = H * p[1];
  c.center[0] = W * p[2] - 0.5; c.center[1] = H * p[3] - 0.5;
  c.inv_focal[0] = 1.0 / c.focal[0]; c.inv_focal[1] = 1.0 / c.focal[1];
  c.w = p[4];
  if (c.w != 0.0) {
    c.tan2 = 2.0 * std::tan(c.w / 2.0);
    c.one_over_tan2 = 1.0 / c.tan2;
    c.winv = 1.0 / c.w;
    c.dist_enabled = 1.0;
  } else {
    c.winv = 0.0; c.tan2 = 0.0; c.one_over_tan2 = 0.0; c.dist_enabled = 0.0;
  }
  const double v0 = std::max(p[2], 1.0 - p[2]) / p[0];
  const double v1 = std::max(p[3], 1.0 - p[3]) / p[1];
  const double r = std::sqrt(v0 * v0 + v1 * v1);
  c.largest_radius = c.w == 0.0 ? r : std::tan(r * c.w) * c.one_over_tan2;
  c.max_r = 1.5 * c.largest_radius;
  return c;
}
}  // namespace

ptam::CamModel ptam_make_cam_model(const double* p, double W, double H) { return make_cam(p, W, H); }
void ptam_set_global_error(const std::string& e) { g_last_error = e; }

struct ptam_tracker {
  int device = 0, W = 0, H = 0, S = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  TrackerDev dev{};
  // per-stream frame buffers
  DevBuf<uint8_t> pyr;
  DevBuf<int2> corners;
  DevBuf<int> lut;
  DevBuf<uint32_t> mask;
  DevBuf<int> ncorn;           // [S][kLevels] corners per level of the current frames (k_fast2)
  bool lists_stale = false;    // the corner lists / row LUTs lag behind the masks (TrackFrame does not need them)
  DevBuf<StreamCtl> ctl;
  DevBuf<int> pt_count;
  std::vector<int> h_pt_count;
  // keyframe store
  std::vector<uint8_t*> kf_bufs;
  DevBuf<const uint8_t*> kf_ptrs;
  size_t kf_ptr_cap = 0;
  // per-point arrays
  int cap = 0;
  DevBuf<double> world, right, down, last_warp, m2buf, v3cam, v2image, derivs, warp_inv, v2found, sin_, J, e2;
  DevBuf<int> src_kf, src_level, tsum, tsumsq, flags, level, search_level, outliers, inliers, pvs, iter_idx;
  DevBuf<int2> center;
  DevBuf<int4> geo;
  DevBuf<float> sbi_tmpl;
  DevBuf<double> refind_pose, unit_mu;
  DevBuf<int> unit_nfound;
  bool unit_searched = false;  // a ptam_patch_search_batch result is resident (what ptam_pose_update works on)
  // relocaliser: keyframe poses (host copy + device table) and which of them have been given
  DevBuf<double> kf_pose;
  size_t kf_pose_cap = 0;
  std::vector<double> h_kf_pose;
  std::vector<char> h_kf_has_pose;
  // MakeKeyFrame_Rest scratch (one stream at a time; allocated on first use)
  DevBuf<uint8_t> rest_smap;
  DevBuf<int2> rest_max, rest_cand;
  DevBuf<double> rest_cand_score;
  DevBuf<int> rest_counts;
  int rest_stream = -1;
  // AddPointEpipolar scratch (allocated / grown on first use)
  DevBuf<int2> epi_cand;
  DevBuf<double2> epi_implane;
  DevBuf<int> epi_found, epi_best;
  DevBuf<double> epi_sub;
  size_t sbi_smem = 0;
  DevBuf<uint8_t> tmpl;
  // pinned staging
  uint8_t* h_stage = nullptr;
  StreamCtl* h_ctl = nullptr;
  cudaEvent_t stage_ev = nullptr;  // staging buffer is free again once this has fired
  // optional per-kernel timing (CUDA events on the handle's stream around every launch)
  bool profiling = false;
  cudaEvent_t prof_ev[16] = {};
  double prof_ms[8] = {};
  int64_t prof_n[8] = {};
  // pipelined submit / collect: two frame landing buffers, a copy stream, per-slot events and results
  cudaStream_t cstream = nullptr;
  DevBuf<uint8_t> l0slot[2];
  StreamCtl* h_ctl_slot[2] = {nullptr, nullptr};
  cudaEvent_t ev_h2d[2] = {}, ev_done[2] = {};
  long long n_submit = 0, n_collect = 0;
  cudaEvent_t dbg_ev[2][4] = {};  // PTAM_B200_DEBUG_TIMES: H2D begin / end, image work end, chain end per slot
  bool dbg_times = false;

  void set_error(const std::string& e) { err = e; g_last_error = e; }

  ~ptam_tracker() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (auto p : kf_bufs) cudaFree(p);
    pyr.free(); corners.free(); lut.free(); mask.free(); ctl.free(); pt_count.free(); kf_ptrs.free();
    world.free(); right.free(); down.free(); last_warp.free(); m2buf.free(); geo.free(); sbi_tmpl.free(); refind_pose.free(); unit_mu.free(); unit_nfound.free(); rest_smap.free(); rest_max.free(); rest_cand.free(); rest_cand_score.free(); rest_counts.free(); v3cam.free(); v2image.free(); derivs.free();
    warp_inv.free(); v2found.free(); sin_.free(); J.free(); e2.free(); src_kf.free(); src_level.free();
    tsum.free(); tsumsq.free(); flags.free(); level.free(); search_level.free(); outliers.free(); inliers.free();
    pvs.free(); iter_idx.free(); center.free(); tmpl.free();
    epi_cand.free(); epi_implane.free(); epi_found.free(); epi_best.free(); epi_sub.free(); kf_pose.free();
    if (stage_ev) cudaEventDestroy(stage_ev);
    for (auto e : prof_ev) if (e) cudaEventDestroy(e);
    for (int k = 0; k < 2; k++) {
      l0slot[k].free();
      if (h_ctl_slot[k]) cudaFreeHost(h_ctl_slot[k]);
      if (ev_h2d[k]) cudaEventDestroy(ev_h2d[k]);
      if (ev_done[k]) cudaEventDestroy(ev_done[k]);
    }
    if (cstream) cudaStreamDestroy(cstream);
    if (istream) { cudaStreamSynchronize(istream); cudaStreamDestroy(istream); }
    for (auto e : {ev_fork, ev_pyr, ev_img, ev_imgfree}) if (e) cudaEventDestroy(e);
    if (h_stage) cudaFreeHost(h_stage);
    if (h_ctl) cudaFreeHost(h_ctl);
    if (stream) cudaStreamDestroy(stream);
  }

  int init(int dev_id, const double* cam_params, int w, int h, int n_streams, const ptam_tracker_params* prm) {
    device = dev_id; W = w; H = h; S = n_streams;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: the B200 path has no CPU fallback"); return PTAM_ERR_NO_DEVICE; }
    if (dev_id < 0 || dev_id >= ndev) { set_error("bad device index"); return PTAM_ERR_INVALID; }
    if (w < 64 || h < 64 || n_streams < 1) { set_error("image must be at least 64x64 and n_streams >= 1"); return PTAM_ERR_INVALID; }
    PTAM_CUDA_TRY(this, cudaSetDevice(dev_id));
    PTAM_CUDA_TRY(this, cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    PTAM_CUDA_TRY(this, cudaEventCreateWithFlags(&stage_ev, cudaEventDisableTiming));
    Geom& g = dev.g;
    make_geom(g, w, h);
    dev.cam = make_cam(cam_params, w, h);
    if (prm) dev.prm = *prm; else ptam_tracker_default_params(&dev.prm);
    {  // SmallBlurryImage geometry and Gaussian taps (ImageProcess.cc:279-304; libCVD convolveGaussian)
      SbiDev& sb = dev.sbi;
      sb.w = g.lev[3].w / 2; sb.h = g.lev[3].h / 2; sb.n = sb.w * sb.h;
      const double sigma = dev.prm.rotation_estimator_blur;
      sb.ks = (int)std::ceil(3.0 * sigma);
      if (sb.ks < 0 || sb.ks > 7 || !(sigma > 0)) { set_error("rotation_estimator_blur out of range (0, 7/3]"); return PTAM_ERR_INVALID; }
      float ksum = 0.f;
      for (int i = 0; i < 8; i++) sb.taps[i] = 0.f;
      for (int i = 1; i <= sb.ks; i++) ksum += (sb.taps[i] = (float)std::exp(-i * i / (2 * sigma * sigma)));
      sb.taps[0] = 1.f;
      ksum = ksum * 2 + sb.taps[0];
      const double factor = 1.0 / ksum;
      for (int i = 0; i <= sb.ks; i++) sb.taps[i] = (float)(sb.taps[i] * factor);
      sb.cam_small = make_cam(cam_params, sb.w, sb.h);
      if (sb.n < 9) { set_error("image too small for the rotation estimator"); return PTAM_ERR_INVALID; }
      sbi_smem = (size_t)sb.n * (7 * sizeof(float) + 1) + 16;
      // the relocaliser's small blurry images use SmallBlurryImage's default blur 2.5 (ImageProcess.h:54)
      const double s25 = 2.5;
      dev.ks25 = (int)std::ceil(3.0 * s25);
      float k25 = 0.f;
      for (int i = 0; i < 12; i++) dev.taps25[i] = 0.f;
      for (int i = 1; i <= dev.ks25; i++) k25 += (dev.taps25[i] = (float)std::exp(-i * i / (2 * s25 * s25)));
      dev.taps25[0] = 1.f;
      k25 = k25 * 2 + dev.taps25[0];
      const double f25 = 1.0 / k25;
      for (int i = 0; i <= dev.ks25; i++) dev.taps25[i] = (float)(dev.taps25[i] * f25);
      dev.kf_sbi_off = (g.pyr_bytes + 255) & ~(size_t)255;
      dev.reloc_on = 0; dev.kf_pose = nullptr;
    }
    dev.S = S;
    dev.pose_ws_smem = 1;  // 0: k_pose works on the global arrays (measured: 0.149 -> 0.178 ms per 296 frames)
    if (const char* e = std::getenv("PTAM_B200_POSE_SMEM")) dev.pose_ws_smem = std::atoi(e);
    PTAM_CUDA_TRY(this, pyr.alloc(g.pyr_bytes * S));
    PTAM_CUDA_TRY(this, corners.alloc(g.corner_stride * S));
    PTAM_CUDA_TRY(this, lut.alloc((size_t)g.lut_stride * S));
    PTAM_CUDA_TRY(this, mask.alloc(g.mask_stride * S));
    PTAM_CUDA_TRY(this, ncorn.alloc((size_t)kLevels * S));
    PTAM_CUDA_TRY(this, ctl.alloc(S));
    PTAM_CUDA_TRY(this, pt_count.alloc(S));
    h_pt_count.assign(S, 0);
    PTAM_CUDA_TRY(this, cudaMallocHost(&h_stage, g.pyr_bytes * S));
    PTAM_CUDA_TRY(this, cudaMallocHost(&h_ctl, sizeof(StreamCtl) * S));
    std::memset(h_ctl, 0, sizeof(StreamCtl) * S);
    for (int s = 0; s < S; s++) {
      ptam_tracker_state& st = h_ctl[s].st;
      st.se3_cam_from_world[0] = st.se3_cam_from_world[4] = st.se3_cam_from_world[8] = 1.0;
      st.tracking_quality = 2;
      st.scene_depth_mean = 1.0; st.scene_depth_sigma = 1.0;
    }
    PTAM_CUDA_TRY(this, cudaMemcpy(ctl.p, h_ctl, sizeof(StreamCtl) * S, cudaMemcpyHostToDevice));
    PTAM_CUDA_TRY(this, cudaStreamSynchronize(cudaStreamLegacy));
    dev.pyr = pyr.p; dev.corners = corners.p; dev.lut = lut.p; dev.mask = mask.p; dev.ncorn = ncorn.p; dev.ctl = ctl.p;
    dev.pt_count = pt_count.p;
    dev.kf_ptrs = nullptr; dev.n_kf = 0;
    PTAM_CUDA_TRY(this, ensure_points(1024));
    PTAM_CUDA_TRY(this, sbi_tmpl.alloc((size_t)2 * S * dev.sbi.n));
    dev.sbi.tmpl = sbi_tmpl.p;
    if (sbi_smem > 48 * 1024) {
      PTAM_CUDA_TRY(this, cudaFuncSetAttribute(k_sbi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbi_smem));
      PTAM_CUDA_TRY(this, cudaFuncSetAttribute(k_kf_sbi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbi_smem));
      PTAM_CUDA_TRY(this, cudaFuncSetAttribute(k_reloc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbi_smem));
    }
    PTAM_CUDA_TRY(this, cudaFuncSetAttribute(k_pose, cudaFuncAttributeMaxDynamicSharedMemorySize, kPoseSmemBytes));
    return PTAM_OK;
  }

  template <class T>
  cudaError_t regrow(DevBuf<T>& b, int per_point, int old_cap, int new_cap) {
    DevBuf<T> nb;
    cudaError_t e = nb.alloc((size_t)new_cap * per_point * S);
    if (e != cudaSuccess) return e;
    if (b.p && old_cap)
      e = cudaMemcpy2D(nb.p, (size_t)new_cap * per_point * sizeof(T), b.p, (size_t)old_cap * per_point * sizeof(T),
                       (size_t)old_cap * per_point * sizeof(T), S, cudaMemcpyDeviceToDevice);
    b.free();
    b = nb;
    return e;
  }

  cudaError_t ensure_points(int n) {
    if (n <= cap) return cudaSuccess;
    cudaStreamSynchronize(stream);
    const int nc = std::max(n, cap * 2);
    cudaError_t e = cudaSuccess;
#define RG(buf, k) if (e == cudaSuccess) e = regrow(buf, k, cap, nc)
    RG(world, 3); RG(right, 3); RG(down, 3); RG(last_warp, 4); RG(m2buf, 4); RG(geo, 2); RG(v3cam, 3); RG(v2image, 2); RG(derivs, 4);
    RG(warp_inv, 4); RG(v2found, 2); RG(sin_, 1); RG(J, 12); RG(e2, 1); RG(src_kf, 1); RG(src_level, 1);
    RG(tsum, 1); RG(tsumsq, 1); RG(flags, 1); RG(level, 1); RG(search_level, 1); RG(outliers, 1); RG(inliers, 1);
    RG(pvs, 4); RG(iter_idx, 1); RG(center, 1); RG(tmpl, 64);
#undef RG
    if (e != cudaSuccess) return e;
    cap = nc;
    PointArrays& p = dev.p;
    p.world = world.p; p.right = right.p; p.down = down.p; p.src_kf = src_kf.p; p.src_level = src_level.p; p.center = center.p;
    p.tmpl = tmpl.p; p.tsum = tsum.p; p.tsumsq = tsumsq.p; p.last_warp = last_warp.p; p.m2 = m2buf.p; p.geo = geo.p;
    p.flags = flags.p; p.level = level.p; p.search_level = search_level.p;
    p.v3cam = v3cam.p; p.v2image = v2image.p; p.derivs = derivs.p; p.warp_inv = warp_inv.p;
    p.v2found = v2found.p; p.sqrt_inv_noise = sin_.p; p.J = J.p; p.outliers = outliers.p; p.inliers = inliers.p;
    p.pvs = pvs.p; p.iter_idx = iter_idx.p; p.e2 = e2.p; p.cap = cap;
    return cudaSuccess;
  }

  // cudaPointerGetAttributes costs tens of microseconds per host pointer; frame buffers come from a
  // small ring in practice, so the answer is cached per pointer.  A stale "pinned" entry is harmless
  // (cudaMemcpyAsync from pageable memory is still correct, only synchronous).
  std::unordered_map<const void*, bool> pinned_cache;
  std::vector<void*> batch_dst, batch_src;
  std::vector<size_t> batch_size;
  bool is_pinned(const void* p) {
    auto it = pinned_cache.find(p);
    if (it != pinned_cache.end()) return it->second;
    cudaPointerAttributes a;
    bool pinned = false;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) cudaGetLastError();
    else pinned = a.type == cudaMemoryTypeHost;
    if (pinned_cache.size() > 8192) pinned_cache.clear();
    pinned_cache.emplace(p, pinned);
    return pinned;
  }

  int upload_images(const uint8_t* const* images, int stride, uint8_t* dst, size_t dst_stream_pitch, int n, cudaStream_t st = nullptr) {
    if (!st) st = stream;
    const LevelDesc& L0 = dev.g.lev[0];
    // page-locked caller buffers (cudaHostAlloc / cudaHostRegister): DMA straight from them
    bool pinned = true;
    for (int s = 0; s < n && pinned; s++) pinned = is_pinned(images[s]);
    if (pinned) {
      const bool dense = stride == W && L0.pitch == W;  // whole frame is one contiguous run on both sides
      if (dense && n > 1) {  // one driver call for the whole batch instead of n copy submissions
        batch_dst.resize(n); batch_src.resize(n); batch_size.assign(n, (size_t)W * H);
        for (int s = 0; s < n; s++) { batch_dst[s] = dst + (size_t)s * dst_stream_pitch; batch_src[s] = const_cast<uint8_t*>(images[s]); }
        cudaMemcpyAttributes at{};
        at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        size_t attr_idx = 0, fail_idx = 0;
        if (cudaMemcpyBatchAsync(batch_dst.data(), batch_src.data(), batch_size.data(), (size_t)n, &at, &attr_idx, 1, &fail_idx, st) == cudaSuccess)
          return PTAM_OK;
        cudaGetLastError();  // not available on this driver: fall back to one copy per frame
      }
      for (int s = 0; s < n; s++) {
        if (dense) PTAM_CUDA_TRY(this, cudaMemcpyAsync(dst + (size_t)s * dst_stream_pitch, images[s], (size_t)W * H, cudaMemcpyHostToDevice, st));
        else PTAM_CUDA_TRY(this, cudaMemcpy2DAsync(dst + (size_t)s * dst_stream_pitch, L0.pitch, images[s], stride, W, H,
                                                   cudaMemcpyHostToDevice, st));
      }
      return PTAM_OK;
    }
    // pageable memory: pack into the pinned staging buffer with the library pitch, then async H2D
    PTAM_CUDA_TRY(this, cudaEventSynchronize(stage_ev));
    for (int s = 0; s < n; s++) {
      uint8_t* o = h_stage + (size_t)s * dst_stream_pitch;
      for (int y = 0; y < H; y++) std::memcpy(o + (size_t)y * L0.pitch, images[s] + (size_t)y * stride, W);
    }
    for (int s = 0; s < n; s++)
      PTAM_CUDA_TRY(this, cudaMemcpyAsync(dst + (size_t)s * dst_stream_pitch, h_stage + (size_t)s * dst_stream_pitch,
                                          (size_t)L0.pitch * H, cudaMemcpyHostToDevice, st));
    PTAM_CUDA_TRY(this, cudaEventRecord(stage_ev, st));
    return PTAM_OK;
  }

  void pbegin(int k) { if (profiling) cudaEventRecord(prof_ev[2 * k], stream); }
  void pend(int k) { if (profiling) cudaEventRecord(prof_ev[2 * k + 1], stream); launches++; }
  int pcollect(unsigned used) {
    if (!profiling) return PTAM_OK;
    PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
    for (int k = 0; k < 8; k++)
      if (used & (1u << k)) {
        float ms = 0;
        PTAM_CUDA_TRY(this, cudaEventElapsedTime(&ms, prof_ev[2 * k], prof_ev[2 * k + 1]));
        prof_ms[k] += ms; prof_n[k]++;
      }
    return PTAM_OK;
  }

  // Image stream.  A frame batch is two chains: the image work (pyramid + FAST + corner lists) on `istream`, and
  // the tracking work on the handle's stream.  Within a batch, k_sbi / k_pvs_select (latency-bound, one CTA per
  // tracker) run beside the FAST of levels 1..3 and k_compact; across batches of the pipelined entry points
  // (ptam_tracker_submit_frames[_device]) the image work of batch i+1 runs beside the ten fine Gauss-Newton
  // iterations of batch i, which touch neither the frame nor the corner lists.  Ordering:
  //   istream:  [wait: frames ready, ev_imgfree = fine search of the previous batch done]  k_fast2<true>  (ev_pyr)
  //             k_fast2<false>  k_compact  (ev_img)
  //   stream:   [wait ev_pyr]  k_sbi  k_reloc  k_pvs_select  [wait ev_img]  coarse search + pose, fine search
  //             (ev_imgfree)  fine pose
  // The handle's stream always waits for ev_img of its own batch, so once it has drained, istream has too.
  cudaStream_t istream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_pyr = nullptr, ev_img = nullptr, ev_imgfree = nullptr;
  bool imgfree_pending = false;  // ev_imgfree has been recorded (a batch went through the pipelined path)

  int ensure_istream() {
    if (istream) return PTAM_OK;
    PTAM_CUDA_TRY(this, cudaStreamCreateWithFlags(&istream, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&ev_fork, &ev_pyr, &ev_img, &ev_imgfree}) PTAM_CUDA_TRY(this, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return PTAM_OK;
  }

  // pyramid + FAST (+ corner lists and row LUTs when `lists`: TrackFrame itself reads the corner masks only, the
  // lists are made on demand by ensure_lists())
  void queue_keyframe(const TrackerDev& d, cudaStream_t st, bool prof, bool lists, cudaEvent_t after_l0 = nullptr) {
    const LevelDesc& L0 = d.g.lev[0];
    cudaMemsetAsync(ncorn.p, 0, sizeof(int) * kLevels * S, st);
    if (prof) pbegin(0);
    k_fast2<true><<<dim3(L0.tiles_x, L0.tiles_y, S), 256, 0, st>>>(d);
    if (prof) pend(0); else launches++;
    if (after_l0) cudaEventRecord(after_l0, st);
    if (prof) pbegin(1);
    k_fast2<false><<<dim3(d.g.fast_tiles, S), 256, 0, st>>>(d);
    if (prof) pend(1); else launches++;
    if (lists) {
      if (prof) pbegin(2);
      k_compact<<<dim3(kLevels, S), 1024, 0, st>>>(d);
      if (prof) pend(2); else launches++;
    }
    lists_stale = !lists;
  }

  // corner lists, row LUTs and StreamCtl::n_corners of the frames last processed (everything but TrackFrame's own
  // search reads them: get_level, MakeKeyFrame_Rest, the epipolar search)
  int ensure_lists() {
    if (!lists_stale) return PTAM_OK;
    k_compact<<<dim3(kLevels, S), 1024, 0, stream>>>(dev);
    launches++;
    PTAM_CUDA_TRY(this, cudaGetLastError());
    lists_stale = false;
    return PTAM_OK;
  }

  int launch_keyframe(const TrackerDev& d, bool collect = true) {
    rest_stream = -1;  // MakeKeyFrame_Rest results belong to the previous frame
    queue_keyframe(d, stream, profiling, true);
    PTAM_CUDA_TRY(this, cudaGetLastError());
    return collect ? pcollect(7u) : PTAM_OK;
  }

  // frames_ready: nullptr = the frames are ordered by the handle's stream (everything queued on it so far);
  // otherwise an event the image stream waits for instead (pipelined entry points)
  int launch_track(const TrackerDev& d, cudaEvent_t frames_ready = nullptr, bool pipelined = false) {
    rest_stream = -1;
    const bool prof = profiling;  // per-kernel event timing needs one serial chain
    static const bool no_istream = std::getenv("PTAM_B200_ISTREAM") && std::atoi(std::getenv("PTAM_B200_ISTREAM")) == 0;
    int maxn = 0;
    for (int s = 0; s < S; s++) maxn = std::max(maxn, h_pt_count[s]);
    unsigned used = 3u | 8u | 32u | 128u;
    if (prof) queue_keyframe(d, stream, true, false);
    else if (no_istream) { queue_keyframe(d, stream, false, false); ensure_istream(); cudaEventRecord(ev_pyr, stream); cudaEventRecord(ev_img, stream); }
    else {
      int rc = ensure_istream();
      if (rc) return rc;
      if (!pipelined) {
        PTAM_CUDA_TRY(this, cudaEventRecord(ev_fork, stream));
        PTAM_CUDA_TRY(this, cudaStreamWaitEvent(istream, ev_fork, 0));
      } else {
        if (frames_ready) PTAM_CUDA_TRY(this, cudaStreamWaitEvent(istream, frames_ready, 0));
        if (imgfree_pending) PTAM_CUDA_TRY(this, cudaStreamWaitEvent(istream, ev_imgfree, 0));
      }
      queue_keyframe(d, istream, false, false, ev_pyr);
      PTAM_CUDA_TRY(this, cudaEventRecord(ev_img, istream));
      PTAM_CUDA_TRY(this, cudaStreamWaitEvent(stream, ev_pyr, 0));
    }
    if (prof) pbegin(3);
    k_sbi<<<S, 256, sbi_smem, stream>>>(d);
    if (d.reloc_on) { k_reloc<<<S, 256, sbi_smem, stream>>>(d); launches++; }
    k_pvs_select<<<S, pvs_threads(), 0, stream>>>(d);
    if (prof) pend(3); else launches++;
    launches++;
    if (!prof) PTAM_CUDA_TRY(this, cudaStreamWaitEvent(stream, ev_img, 0));
    const int coarse_items = std::min(maxn, 2 * std::max(0, d.prm.coarse_max));
    if (coarse_items > 0) {
      if (prof) pbegin(4);
      k_search_prep<<<dim3((coarse_items + 127) / 128, S), 128, 0, stream>>>(d, 0);
      k_search<<<dim3((coarse_items + 3) / 4, S), 128, 0, stream>>>(d, 0);
      if (prof) pend(4); else launches++;
      launches++; used |= 16u;
    }
    if (prof) pbegin(5);
    k_pose<<<S, kPoseThreads, d.pose_ws_smem ? kPoseSmemBytes : 0, stream>>>(d, 0);
    if (prof) pend(5); else launches++;
    if (maxn > 0) {
      if (prof) pbegin(6);
      k_search_prep<<<dim3((maxn + 127) / 128, S), 128, 0, stream>>>(d, 1);
      k_search<<<dim3((maxn + 3) / 4, S), 128, 0, stream>>>(d, 1);
      if (prof) pend(6); else launches++;
      launches++; used |= 64u;
    }
    if (!prof) { PTAM_CUDA_TRY(this, cudaEventRecord(ev_imgfree, stream)); imgfree_pending = true; }
    if (prof) pbegin(7);
    k_pose<<<S, kPoseThreads, d.pose_ws_smem ? kPoseSmemBytes : 0, stream>>>(d, 1);
    if (prof) pend(7); else launches++;
    PTAM_CUDA_TRY(this, cudaGetLastError());
    return pcollect(used);
  }

  int fetch_results(ptam_track_result* results) {
    PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_ctl, ctl.p, sizeof(StreamCtl) * S, cudaMemcpyDeviceToHost, stream));
    PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
    convert_results(h_ctl, results);
    return PTAM_OK;
  }

  int ensure_pipeline(bool landing_buffers) {
    if (landing_buffers && !l0slot[0].p)
      for (int k = 0; k < 2; k++) PTAM_CUDA_TRY(this, l0slot[k].alloc((size_t)dev.g.lev[0].pitch * H * S));
    if (cstream) return PTAM_OK;
    PTAM_CUDA_TRY(this, cudaStreamCreateWithFlags(&cstream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      PTAM_CUDA_TRY(this, cudaMallocHost(&h_ctl_slot[k], sizeof(StreamCtl) * S));
      PTAM_CUDA_TRY(this, cudaEventCreateWithFlags(&ev_h2d[k], cudaEventDisableTiming));
      PTAM_CUDA_TRY(this, cudaEventCreateWithFlags(&ev_done[k], cudaEventDisableTiming));
    }
    return PTAM_OK;
  }

  void convert_results(const StreamCtl* hc, ptam_track_result* results) {
    for (int s = 0; s < S; s++) {
      const StreamCtl& c = hc[s];
      ptam_track_result& r = results[s];
      std::memcpy(r.se3_cam_from_world, c.st.se3_cam_from_world, sizeof(double) * 12);
      r.scene_depth_mean = c.st.scene_depth_mean; r.scene_depth_sigma = c.st.scene_depth_sigma;
      for (int l = 0; l < kLevels; l++) {
        r.meas_attempted[l] = c.attempted[l]; r.meas_found[l] = c.found[l];
        r.n_corners[l] = c.res_n_corners[l]; r.n_pvs[l] = c.n_pvs[l];
      }
      r.did_coarse = c.did_coarse; r.n_coarse = c.n_coarse; r.n_level3 = c.n_l3; r.n_fine = c.n_fine;
      r.tracking_quality = c.st.tracking_quality; r.quality_needs_kf_distance = c.needs_kf_distance;
      r.n_candidates = c.n_cand;
      const int fm = dev.reloc_on ? c.frame_mode : 0;
      r.recovery = fm; r.reloc_keyframe = fm ? c.reloc_kf : -1; r.reserved1 = 0;
      r.reloc_score = fm ? c.reloc_score : 0.0;
    }
  }
};

extern "C" {

void ptam_tracker_default_params(ptam_tracker_params* p) {
  p->coarse_min = 20; p->coarse_max = 60; p->coarse_range = 30; p->coarse_subpix_its = 8;
  p->disable_coarse = 0; p->max_patches_per_frame = 1000; p->mestimator = 0; p->use_constant_velocity = 1;
  p->coarse_min_velocity = 0.006; p->quality_good = 0.3; p->quality_lost = 0.13;
  p->use_rotation_estimator = 1; p->reserved0 = 0; p->rotation_estimator_blur = 0.75;
}

const char* ptam_global_last_error(void) { return g_last_error.c_str(); }

ptam_tracker* ptam_tracker_create(int device, const double* cam_params, int width, int height, int n_streams,
                                  const ptam_tracker_params* params) {
  ptam_tracker* t = new ptam_tracker;
  if (t->init(device, cam_params, width, height, n_streams, params) != PTAM_OK) {
    g_last_error = t->err;
    delete t;
    return nullptr;
  }
  return t;
}
void ptam_tracker_destroy(ptam_tracker* t) { delete t; }
const char* ptam_tracker_last_error(const ptam_tracker* t) { return t->err.c_str(); }

int ptam_tracker_add_keyframe(ptam_tracker* t, const uint8_t* image, int stride) {
  cudaSetDevice(t->device);
  uint8_t* buf = nullptr;
  PTAM_CUDA_TRY(t, cudaMalloc(&buf, t->dev.kf_sbi_off + sizeof(float) * t->dev.sbi.n));  // pyramid + small blurry image
  t->kf_bufs.push_back(buf);
  t->h_kf_pose.resize(12 * t->kf_bufs.size(), 0.0);
  t->h_kf_has_pose.push_back(0);
  t->dev.reloc_on = 0;  // until the new keyframe has a pose too
  const uint8_t* imgs[1] = {image};
  int rc = t->upload_images(imgs, stride, buf, t->dev.g.pyr_bytes, 1);
  if (rc) return rc;
  TrackerDev d = t->dev;
  d.pyr = buf; d.src.l0 = buf; d.src.stream_pitch = 0; d.src.pitch = d.g.lev[0].pitch;
  const LevelDesc& L0 = d.g.lev[0];
  k_pyramid<<<dim3((L0.w + 63) / 64, (L0.h + 63) / 64, 1), 256, 0, t->stream>>>(d);
  t->launches++;
  PTAM_CUDA_TRY(t, cudaGetLastError());
  if (t->kf_bufs.size() > t->kf_ptr_cap) {
    PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
    t->kf_ptr_cap = std::max<size_t>(256, t->kf_ptr_cap * 2);
    PTAM_CUDA_TRY(t, t->kf_ptrs.alloc(t->kf_ptr_cap));
  }
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  PTAM_CUDA_TRY(t, cudaMemcpy(t->kf_ptrs.p, t->kf_bufs.data(), sizeof(uint8_t*) * t->kf_bufs.size(), cudaMemcpyHostToDevice));
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(cudaStreamLegacy));  // pageable H2D: the DMA may still be in flight on return
  t->dev.kf_ptrs = t->kf_ptrs.p;
  t->dev.n_kf = (int)t->kf_bufs.size();
  k_kf_sbi<<<1, 256, t->sbi_smem, t->stream>>>(t->dev, t->dev.n_kf - 1);  // KeyFrame::pSBI (KeyFrame.cc:80-81)
  t->launches++;
  PTAM_CUDA_TRY(t, cudaGetLastError());
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  return t->dev.n_kf - 1;
}

int ptam_tracker_set_keyframe_pose(ptam_tracker* t, int kf, const double* se3) {
  cudaSetDevice(t->device);
  if (kf < 0 || kf >= (int)t->kf_bufs.size()) { t->set_error("unknown keyframe"); return PTAM_ERR_INVALID; }
  std::memcpy(&t->h_kf_pose[12 * (size_t)kf], se3, sizeof(double) * 12);
  t->h_kf_has_pose[kf] = 1;
  if (t->kf_bufs.size() > t->kf_pose_cap) {
    PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
    t->kf_pose.free();
    t->kf_pose_cap = std::max<size_t>(256, 2 * t->kf_bufs.size());
    PTAM_CUDA_TRY(t, t->kf_pose.alloc(12 * t->kf_pose_cap));
  }
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  PTAM_CUDA_TRY(t, cudaMemcpy(t->kf_pose.p, t->h_kf_pose.data(), sizeof(double) * t->h_kf_pose.size(), cudaMemcpyHostToDevice));
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(cudaStreamLegacy));
  t->dev.kf_pose = t->kf_pose.p;
  bool all = true;
  for (char c : t->h_kf_has_pose) all = all && c;
  t->dev.reloc_on = all ? 1 : 0;
  return PTAM_OK;
}

int ptam_tracker_set_map(ptam_tracker* t, int stream, int n, const double* world, const double* right, const double* down,
                         const int32_t* src_kf, const int32_t* src_level, const int32_t* center) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S || n < 0) { t->set_error("bad stream / point count"); return PTAM_ERR_INVALID; }
  for (int i = 0; i < n; i++)
    if (src_kf[i] < 0 || src_kf[i] >= t->dev.n_kf || src_level[i] < 0 || src_level[i] >= kLevels) {
      t->set_error("map point references an unknown keyframe or level");
      return PTAM_ERR_INVALID;
    }
  PTAM_CUDA_TRY(t, t->ensure_points(n));
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  const size_t o = (size_t)stream * t->cap;
#define UP(buf, src, k, T) PTAM_CUDA_TRY(t, cudaMemcpy(buf.p + o * k, src, sizeof(T) * k * n, cudaMemcpyHostToDevice))
  if (n) {
    UP(t->world, world, 3, double); UP(t->right, right, 3, double); UP(t->down, down, 3, double);
    UP(t->src_kf, src_kf, 1, int); UP(t->src_level, src_level, 1, int);
    PTAM_CUDA_TRY(t, cudaMemcpy(t->center.p + o, center, sizeof(int2) * n, cudaMemcpyHostToDevice));
  }
#undef UP
  PTAM_CUDA_TRY(t, cudaMemset(t->flags.p + o, 0, sizeof(int) * t->cap));
  PTAM_CUDA_TRY(t, cudaMemset(t->outliers.p + o, 0, sizeof(int) * t->cap));
  PTAM_CUDA_TRY(t, cudaMemset(t->inliers.p + o, 0, sizeof(int) * t->cap));
  PTAM_CUDA_TRY(t, cudaMemset(t->level.p + o, 0xff, sizeof(int) * t->cap));
  t->h_pt_count[stream] = n;
  PTAM_CUDA_TRY(t, cudaMemcpy(t->pt_count.p, t->h_pt_count.data(), sizeof(int) * t->S, cudaMemcpyHostToDevice));
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(cudaStreamLegacy));  // the uploads and memsets above ran on the legacy stream
  return PTAM_OK;
}

int ptam_tracker_set_state(ptam_tracker* t, int stream, const ptam_tracker_state* s) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S) { t->set_error("bad stream"); return PTAM_ERR_INVALID; }
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  PTAM_CUDA_TRY(t, cudaMemcpy(&t->ctl.p[stream].st, s, sizeof(*s), cudaMemcpyHostToDevice));
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(cudaStreamLegacy));
  return PTAM_OK;
}
int ptam_tracker_get_state(ptam_tracker* t, int stream, ptam_tracker_state* s) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S) { t->set_error("bad stream"); return PTAM_ERR_INVALID; }
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  PTAM_CUDA_TRY(t, cudaMemcpy(s, &t->ctl.p[stream].st, sizeof(*s), cudaMemcpyDeviceToHost));
  return PTAM_OK;
}

static int use_host_frames(ptam_tracker* t, const uint8_t* const* images, int stride, TrackerDev& d) {
  int rc = t->upload_images(images, stride, t->pyr.p, t->dev.g.pyr_bytes, t->S);
  if (rc) return rc;
  d = t->dev;
  d.src.l0 = t->pyr.p; d.src.stream_pitch = d.g.pyr_bytes; d.src.pitch = d.g.lev[0].pitch;
  t->dev.src = d.src;
  return PTAM_OK;
}

int ptam_tracker_make_keyframes(ptam_tracker* t, const uint8_t* const* images, int stride) {
  cudaSetDevice(t->device);
  TrackerDev d;
  int rc = use_host_frames(t, images, stride, d);
  if (rc) return rc;
  rc = t->launch_keyframe(d);
  if (rc) return rc;
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  return PTAM_OK;
}

int ptam_tracker_track_frames(ptam_tracker* t, const uint8_t* const* images, int stride, ptam_track_result* results) {
  cudaSetDevice(t->device);
  TrackerDev d;
  int rc = use_host_frames(t, images, stride, d);
  if (rc) return rc;
  rc = t->launch_track(d);
  if (rc) return rc;
  if (results) return t->fetch_results(results);
  return PTAM_OK;
}

int ptam_tracker_track_frames_device(ptam_tracker* t, const uint8_t* d_images, size_t frame_pitch_bytes, int stride,
                                     ptam_track_result* results) {
  cudaSetDevice(t->device);
  if (!d_images || stride < t->W) { t->set_error("bad device frame description"); return PTAM_ERR_INVALID; }
  TrackerDev d = t->dev;
  d.src.l0 = d_images; d.src.stream_pitch = frame_pitch_bytes; d.src.pitch = stride;
  t->dev.src = d.src;
  int rc = t->launch_track(d);
  if (rc) return rc;
  if (results) return t->fetch_results(results);
  return PTAM_OK;
}

static int submit_common(ptam_tracker* t, const TrackerDev& d, cudaEvent_t frames_ready, int k) {
  int rc = t->launch_track(d, frames_ready, true);
  if (rc) return rc;
  if (t->dbg_times && t->dbg_ev[k][0]) { cudaEventRecord(t->dbg_ev[k][2], t->istream ? t->istream : t->stream); cudaEventRecord(t->dbg_ev[k][3], t->stream); }
  PTAM_CUDA_TRY(t, cudaMemcpyAsync(t->h_ctl_slot[k], t->ctl.p, sizeof(StreamCtl) * t->S, cudaMemcpyDeviceToHost, t->stream));
  PTAM_CUDA_TRY(t, cudaEventRecord(t->ev_done[k], t->stream));
  t->n_submit++;
  return PTAM_OK;
}

int ptam_tracker_submit_frames(ptam_tracker* t, const uint8_t* const* images, int stride) {
  cudaSetDevice(t->device);
  if (t->n_submit - t->n_collect >= 2) { t->set_error("two frame batches are already in flight: collect one first"); return PTAM_ERR_CAPACITY; }
  int rc = t->ensure_pipeline(true);
  if (rc) return rc;
  const int k = (int)(t->n_submit & 1);
  const size_t pitch = (size_t)t->dev.g.lev[0].pitch * t->H;
  // the slot's previous batch was collected (its ev_done has fired), so the landing buffer is free
  static const bool dbg = std::getenv("PTAM_B200_DEBUG_TIMES") != nullptr;
  t->dbg_times = dbg;
  if (dbg) {
    for (auto& e : t->dbg_ev[k]) if (!e) cudaEventCreate(&e);
    cudaEventRecord(t->dbg_ev[k][0], t->cstream);
  }
  rc = t->upload_images(images, stride, t->l0slot[k].p, pitch, t->S, t->cstream);
  if (rc) return rc;
  if (dbg) cudaEventRecord(t->dbg_ev[k][1], t->cstream);
  PTAM_CUDA_TRY(t, cudaEventRecord(t->ev_h2d[k], t->cstream));
  PTAM_CUDA_TRY(t, cudaStreamWaitEvent(t->stream, t->ev_h2d[k], 0));
  TrackerDev d = t->dev;
  d.src.l0 = t->l0slot[k].p; d.src.stream_pitch = pitch; d.src.pitch = d.g.lev[0].pitch;
  t->dev.src = d.src;
  return submit_common(t, d, t->ev_h2d[k], k);
}

int ptam_tracker_submit_frames_device(ptam_tracker* t, const uint8_t* d_images, size_t frame_pitch_bytes, int stride, void* ready_event) {
  cudaSetDevice(t->device);
  if (!d_images || stride < t->W) { t->set_error("bad device frame description"); return PTAM_ERR_INVALID; }
  if (t->n_submit - t->n_collect >= 2) { t->set_error("two frame batches are already in flight: collect one first"); return PTAM_ERR_CAPACITY; }
  int rc = t->ensure_pipeline(false);
  if (rc) return rc;
  const int k = (int)(t->n_submit & 1);
  TrackerDev d = t->dev;
  d.src.l0 = d_images; d.src.stream_pitch = frame_pitch_bytes; d.src.pitch = stride;
  t->dev.src = d.src;
  if (ready_event) PTAM_CUDA_TRY(t, cudaStreamWaitEvent(t->stream, (cudaEvent_t)ready_event, 0));
  return submit_common(t, d, (cudaEvent_t)ready_event, k);
}

int ptam_tracker_collect(ptam_tracker* t, ptam_track_result* results) {
  cudaSetDevice(t->device);
  if (t->n_collect >= t->n_submit) { t->set_error("nothing in flight"); return PTAM_ERR_INVALID; }
  const int k = (int)(t->n_collect & 1);
  PTAM_CUDA_TRY(t, cudaEventSynchronize(t->ev_done[k]));
  if (t->dbg_times && t->dbg_ev[k][3]) {
    float h2d = 0, img = 0, chain = 0, gap = 0;
    cudaEventElapsedTime(&h2d, t->dbg_ev[k][0], t->dbg_ev[k][1]);
    cudaEventElapsedTime(&img, t->dbg_ev[k][1], t->dbg_ev[k][2]);
    cudaEventElapsedTime(&chain, t->dbg_ev[k][1], t->dbg_ev[k][3]);
    if (t->dbg_ev[k ^ 1][1] && t->n_collect > 0) cudaEventElapsedTime(&gap, t->dbg_ev[k ^ 1][1], t->dbg_ev[k][0]);
    std::fprintf(stderr, "[ptam dbg] batch %lld: H2D %.3f ms, H2D end -> image work end %.3f, -> chain end %.3f, previous H2D end -> this H2D begin %.3f\n",
                 t->n_collect, h2d, img, chain, gap);
  }
  if (results) t->convert_results(t->h_ctl_slot[k], results);
  t->n_collect++;
  return PTAM_OK;
}

#ifdef PTAM_POSE_CLOCKS
extern "C" int ptam_debug_pose_clocks(long long* out, int reset) {
  if (out && cudaMemcpyFromSymbol(out, ptam::g_pose_clk, sizeof(long long) * 32) != cudaSuccess) return -1;
  if (reset) { long long z[32] = {}; if (cudaMemcpyToSymbol(ptam::g_pose_clk, z, sizeof(z)) != cudaSuccess) return -1; }
  return 0;
}
#endif

int ptam_tracker_synchronize(ptam_tracker* t) {
  cudaSetDevice(t->device);
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  return PTAM_OK;
}
void* ptam_tracker_cuda_stream(ptam_tracker* t) { return (void*)t->stream; }
int64_t ptam_tracker_launch_count(const ptam_tracker* t) { return t->launches; }

int ptam_tracker_set_profiling(ptam_tracker* t, int on) {
  cudaSetDevice(t->device);
  if (on && !t->prof_ev[0])
    for (auto& e : t->prof_ev) PTAM_CUDA_TRY(t, cudaEventCreate(&e));
  t->profiling = on != 0;
  for (int k = 0; k < 8; k++) { t->prof_ms[k] = 0; t->prof_n[k] = 0; }
  return PTAM_OK;
}
int ptam_tracker_get_kernel_times(ptam_tracker* t, double* ms_total, int64_t* launches) {
  for (int k = 0; k < 8; k++) { ms_total[k] = t->prof_ms[k]; launches[k] = t->prof_n[k]; }
  return PTAM_OK;
}

int ptam_tracker_level_size(const ptam_tracker* t, int level, int* w, int* h) {
  if (level < 0 || level >= kLevels) return PTAM_ERR_INVALID;
  *w = t->dev.g.lev[level].w; *h = t->dev.g.lev[level].h;
  return PTAM_OK;
}

int ptam_tracker_get_level(ptam_tracker* t, int stream, int level, uint8_t* pixels, int32_t* corners_xy, int cap, int32_t* row_lut) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S || level < 0 || level >= kLevels) { t->set_error("bad stream / level"); return PTAM_ERR_INVALID; }
  { const int rc = t->ensure_lists(); if (rc) return rc; }
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  const LevelDesc& L = t->dev.g.lev[level];
  if (pixels) {
    const uint8_t* src; size_t pitch;
    if (level == 0) { src = t->dev.src.l0 + (size_t)stream * t->dev.src.stream_pitch; pitch = t->dev.src.pitch; }
    else { src = t->pyr.p + (size_t)stream * t->dev.g.pyr_bytes + L.img_off; pitch = L.pitch; }
    if (!src) { t->set_error("no frame processed yet"); return PTAM_ERR_INVALID; }
    PTAM_CUDA_TRY(t, cudaMemcpy2D(pixels, L.w, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost));
  }
  int n = 0;
  PTAM_CUDA_TRY(t, cudaMemcpy(&n, &t->ctl.p[stream].n_corners[level], sizeof(int), cudaMemcpyDeviceToHost));
  if (corners_xy && cap > 0 && n > 0)
    PTAM_CUDA_TRY(t, cudaMemcpy(corners_xy, t->corners.p + (size_t)stream * t->dev.g.corner_stride + L.corner_off,
                                sizeof(int2) * std::min(n, cap), cudaMemcpyDeviceToHost));
  if (row_lut)
    PTAM_CUDA_TRY(t, cudaMemcpy(row_lut, t->lut.p + (size_t)stream * t->dev.g.lut_stride + L.lut_off, sizeof(int) * L.h, cudaMemcpyDeviceToHost));
  return n;
}

int ptam_tracker_get_points(ptam_tracker* t, int stream, int32_t* flags, int32_t* level, double* v2_found, double* v2_image,
                            int32_t* outl, int32_t* inl) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S) { t->set_error("bad stream"); return PTAM_ERR_INVALID; }
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  const int n = t->h_pt_count[stream];
  const size_t o = (size_t)stream * t->cap;
  if (!n) return 0;
  std::vector<int> fl(n);
  PTAM_CUDA_TRY(t, cudaMemcpy(fl.data(), t->flags.p + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (level) PTAM_CUDA_TRY(t, cudaMemcpy(level, t->level.p + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (v2_found) PTAM_CUDA_TRY(t, cudaMemcpy(v2_found, t->v2found.p + 2 * o, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
  if (v2_image) PTAM_CUDA_TRY(t, cudaMemcpy(v2_image, t->v2image.p + 2 * o, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
  if (outl) PTAM_CUDA_TRY(t, cudaMemcpy(outl, t->outliers.p + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (inl) PTAM_CUDA_TRY(t, cudaMemcpy(inl, t->inliers.p + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) {
    const int f = fl[i];
    int out = 0;
    if (f & F_IN_PVS) {
      out |= PTAM_PT_IN_PVS;
      if (f & F_IN_IMAGE) out |= PTAM_PT_IN_IMAGE;
      if (f & F_SEARCHED) out |= PTAM_PT_SEARCHED;
      if (f & F_FOUND) out |= PTAM_PT_FOUND;
      if ((f & F_FOUND) && (f & F_SUBPIX)) out |= PTAM_PT_SUBPIX;
    }
    if (f & F_TEMPLATE_BAD) out |= PTAM_PT_TEMPLATE_BAD;
    if (flags) flags[i] = out;
    const bool fnd = (f & F_IN_PVS) && (f & F_FOUND);
    if (v2_found && !fnd) v2_found[2 * i] = v2_found[2 * i + 1] = 0;
    if (v2_image && !(f & F_IN_PVS)) v2_image[2 * i] = v2_image[2 * i + 1] = 0;
  }
  return n;
}

int ptam_tracker_refind_in_keyframes(ptam_tracker* t, const uint8_t* const* images, int stride, const double* se3) {
  cudaSetDevice(t->device);
  if (!t->refind_pose.p) PTAM_CUDA_TRY(t, t->refind_pose.alloc((size_t)12 * t->S));
  TrackerDev d;
  int rc = use_host_frames(t, images, stride, d);
  if (rc) return rc;
  PTAM_CUDA_TRY(t, cudaMemcpyAsync(t->refind_pose.p, se3, sizeof(double) * 12 * t->S, cudaMemcpyHostToDevice, t->stream));
  rc = t->launch_keyframe(d, false);
  if (rc) return rc;
  d.mode = 1;
  d.refind_pose = t->refind_pose.p;
  int maxn = 0;
  for (int s = 0; s < t->S; s++) maxn = std::max(maxn, t->h_pt_count[s]);
  k_pvs_select<<<t->S, pvs_threads(), 0, t->stream>>>(d);
  t->launches++;
  if (maxn > 0) {
    k_search_prep<<<dim3((maxn + 127) / 128, t->S), 128, 0, t->stream>>>(d, 1);
    k_search<<<dim3((maxn + 3) / 4, t->S), 128, 0, t->stream>>>(d, 1);
    t->launches += 2;
  }
  PTAM_CUDA_TRY(t, cudaGetLastError());
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  return PTAM_OK;
}

int ptam_tracker_keyframe_rest(ptam_tracker* t, int stream, double min_shi_tomasi_score) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S) { t->set_error("bad stream"); return PTAM_ERR_INVALID; }
  if (!t->dev.src.l0) { t->set_error("no frame processed yet"); return PTAM_ERR_INVALID; }
  { const int rc = t->ensure_lists(); if (rc) return rc; }
  const Geom& g = t->dev.g;
  if (!t->rest_smap.p) {
    PTAM_CUDA_TRY(t, t->rest_smap.alloc(g.pyr_bytes));
    PTAM_CUDA_TRY(t, t->rest_max.alloc(g.corner_stride));
    PTAM_CUDA_TRY(t, t->rest_cand.alloc(g.corner_stride));
    PTAM_CUDA_TRY(t, t->rest_cand_score.alloc(g.corner_stride));
    PTAM_CUDA_TRY(t, t->rest_counts.alloc(8));
  }
  RestDev r{t->rest_smap.p, t->rest_max.p, t->rest_cand.p, t->rest_cand_score.p, t->rest_counts.p, min_shi_tomasi_score};
  PTAM_CUDA_TRY(t, cudaMemsetAsync(t->rest_smap.p, 0, g.pyr_bytes, t->stream));
  // the corner counts live on the device; this call is not on the per-frame path, so read them back
  StreamCtl c;
  PTAM_CUDA_TRY(t, cudaMemcpyAsync(&c, &t->ctl.p[stream], sizeof(c), cudaMemcpyDeviceToHost, t->stream));
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  for (int l = 0; l < kLevels; l++)
    if (c.n_corners[l] > 0) { k_rest_score<<<(c.n_corners[l] + 255) / 256, 256, 0, t->stream>>>(t->dev, r, stream, l); t->launches++; }
  k_rest_select<<<kLevels, 1024, 0, t->stream>>>(t->dev, r, stream);
  t->launches++;
  PTAM_CUDA_TRY(t, cudaGetLastError());
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  t->rest_stream = stream;
  return PTAM_OK;
}

int ptam_tracker_epipolar_search(ptam_tracker* t, int stream, int level, int src_kf, const double* src_se3, double src_depth_mean,
                                 double src_depth_sigma, const double* target_se3, double wiggle_scale, int n_cand,
                                 const int32_t* cand_xy, int32_t* found, int32_t* best_corner, double* sub_pos) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S || level < 0 || level >= kLevels) { t->set_error("bad stream / level"); return PTAM_ERR_INVALID; }
  if (src_kf < 0 || src_kf >= (int)t->kf_bufs.size()) { t->set_error("unknown source keyframe"); return PTAM_ERR_INVALID; }
  if (!t->dev.src.l0) { t->set_error("no current frame: run ptam_tracker_make_keyframes / track_frames first"); return PTAM_ERR_INVALID; }
  if (n_cand <= 0) return PTAM_OK;
  { const int rc = t->ensure_lists(); if (rc) return rc; }
  const Geom& g = t->dev.g;
  if (!t->epi_implane.p) PTAM_CUDA_TRY(t, t->epi_implane.alloc(g.corner_stride));
  if ((size_t)n_cand > t->epi_cand.n) {
    t->epi_cand.free(); t->epi_found.free(); t->epi_best.free(); t->epi_sub.free();
    PTAM_CUDA_TRY(t, t->epi_cand.alloc(n_cand)); PTAM_CUDA_TRY(t, t->epi_found.alloc(n_cand));
    PTAM_CUDA_TRY(t, t->epi_best.alloc(n_cand)); PTAM_CUDA_TRY(t, t->epi_sub.alloc((size_t)2 * n_cand));
  }
  EpiDev e{};
  e.stream = stream; e.level = level; e.src_kf = src_kf; e.n_cand = n_cand;
  for (int i = 0; i < 12; i++) { e.src[i] = src_se3[i]; e.tgt[i] = target_se3[i]; }
  e.d_start = std::max(wiggle_scale, src_depth_mean - src_depth_sigma);       // MapMaker.cc:552-556
  e.d_end = std::min(40 * wiggle_scale, src_depth_mean + src_depth_sigma);
  {  // mdOnePixelDist (ATANCamera.cc:59-64), on the host like the other derived camera parameters
    const CamModel& c = t->dev.cam;
    auto unproject = [&](double ix, double iy, double* o) {
      const double d0 = (ix - c.center[0]) * c.inv_focal[0], d1 = (iy - c.center[1]) * c.inv_focal[1];
      const double dr = std::sqrt(d0 * d0 + d1 * d1);
      const double rr = c.w == 0.0 ? dr : std::tan(dr * c.w) * c.one_over_tan2;
      const double f = dr > 0.01 ? rr / dr : 1.0;
      o[0] = f * d0; o[1] = f * d1;
    };
    double uc[2], ua[2];
    unproject(c.img_w / 2, c.img_h / 2, uc);
    unproject(c.img_w / 2 + 1.0, c.img_h / 2 + 1.0, ua);
    const double d0 = uc[0] - ua[0], d1 = uc[1] - ua[1];
    const double one_pixel_dist = std::sqrt(d0 * d0 + d1 * d1) / std::sqrt(2.0);
    const double md = one_pixel_dist * (4.0 + 1.0 * (1 << level));          // MapMaker.cc:618-619
    e.max_dist_sq = md * md;
  }
  e.cand = t->epi_cand.p; e.implane = t->epi_implane.p; e.found = t->epi_found.p; e.best = t->epi_best.p; e.sub = t->epi_sub.p;
  static_assert(sizeof(int2) == 2 * sizeof(int32_t), "candidate layout");
  PTAM_CUDA_TRY(t, cudaMemcpyAsync(t->epi_cand.p, cand_xy, sizeof(int2) * n_cand, cudaMemcpyHostToDevice, t->stream));
  k_epi_implane<<<(g.lev[level].corner_cap + 255) / 256, 256, 0, t->stream>>>(t->dev, e);
  k_epi_search<<<(n_cand + 3) / 4, 128, 0, t->stream>>>(t->dev, e);
  t->launches += 2;
  PTAM_CUDA_TRY(t, cudaGetLastError());
  PTAM_CUDA_TRY(t, cudaMemcpyAsync(found, t->epi_found.p, sizeof(int) * n_cand, cudaMemcpyDeviceToHost, t->stream));
  PTAM_CUDA_TRY(t, cudaMemcpyAsync(best_corner, t->epi_best.p, sizeof(int) * n_cand, cudaMemcpyDeviceToHost, t->stream));
  PTAM_CUDA_TRY(t, cudaMemcpyAsync(sub_pos, t->epi_sub.p, sizeof(double) * 2 * n_cand, cudaMemcpyDeviceToHost, t->stream));
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  return PTAM_OK;
}

int ptam_tracker_get_level_rest(ptam_tracker* t, int stream, int level, int32_t* max_xy, int max_cap, int32_t* cand_xy,
                                double* cand_score, int cand_cap, int* n_cand) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S || level < 0 || level >= kLevels) { t->set_error("bad stream / level"); return PTAM_ERR_INVALID; }
  if (t->rest_stream != stream) { t->set_error("ptam_tracker_keyframe_rest has not been run for this stream's current frame"); return PTAM_ERR_INVALID; }
  int counts[8];
  PTAM_CUDA_TRY(t, cudaMemcpy(counts, t->rest_counts.p, sizeof(counts), cudaMemcpyDeviceToHost));
  const LevelDesc& L = t->dev.g.lev[level];
  const int nm = counts[level], nc = counts[4 + level];
  if (max_xy && max_cap > 0 && nm > 0)
    PTAM_CUDA_TRY(t, cudaMemcpy(max_xy, t->rest_max.p + L.corner_off, sizeof(int2) * std::min(nm, max_cap), cudaMemcpyDeviceToHost));
  if (cand_xy && cand_cap > 0 && nc > 0)
    PTAM_CUDA_TRY(t, cudaMemcpy(cand_xy, t->rest_cand.p + L.corner_off, sizeof(int2) * std::min(nc, cand_cap), cudaMemcpyDeviceToHost));
  if (cand_score && cand_cap > 0 && nc > 0)
    PTAM_CUDA_TRY(t, cudaMemcpy(cand_score, t->rest_cand_score.p + L.corner_off, sizeof(double) * std::min(nc, cand_cap), cudaMemcpyDeviceToHost));
  if (n_cand) *n_cand = nc;
  return nm;
}

int ptam_tracker_get_sbi(ptam_tracker* t, int stream, float* tmpl, int cap, double* rot3, double* score) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S) { t->set_error("bad stream"); return PTAM_ERR_INVALID; }
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  StreamCtl c;
  PTAM_CUDA_TRY(t, cudaMemcpy(&c, &t->ctl.p[stream], sizeof(c), cudaMemcpyDeviceToHost));
  const int n = t->dev.sbi.n;
  if (tmpl && cap > 0) {
    if (!c.sbi_valid) std::memset(tmpl, 0, sizeof(float) * std::min(n, cap));
    else PTAM_CUDA_TRY(t, cudaMemcpy(tmpl, t->sbi_tmpl.p + ((size_t)c.sbi_idx * t->S + stream) * n, sizeof(float) * std::min(n, cap), cudaMemcpyDeviceToHost));
  }
  if (rot3) for (int k = 0; k < 3; k++) rot3[k] = c.sbi_rot[k];
  if (score) *score = c.sbi_score;
  return n;
}

int ptam_tracker_get_templates(ptam_tracker* t, int stream, uint8_t* tmpl, int32_t* sums) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S) { t->set_error("bad stream"); return PTAM_ERR_INVALID; }
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  const int n = t->h_pt_count[stream];
  const size_t o = (size_t)stream * t->cap;
  if (!n) return 0;
  std::vector<int> fl(n), a(n), b(n);
  PTAM_CUDA_TRY(t, cudaMemcpy(fl.data(), t->flags.p + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  PTAM_CUDA_TRY(t, cudaMemcpy(a.data(), t->tsum.p + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  PTAM_CUDA_TRY(t, cudaMemcpy(b.data(), t->tsumsq.p + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (tmpl) PTAM_CUDA_TRY(t, cudaMemcpy(tmpl, t->tmpl.p + 64 * o, (size_t)64 * n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) {
    const bool has = fl[i] & F_HAS_TEMPLATE;
    if (tmpl && !has) std::memset(tmpl + 64 * i, 0, 64);
    if (sums) { sums[2 * i] = has ? a[i] : 0; sums[2 * i + 1] = has ? b[i] : 0; }
  }
  return n;
}

int ptam_tracker_get_iteration_set(ptam_tracker* t, int stream, int32_t* idx, int cap) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S) { t->set_error("bad stream"); return PTAM_ERR_INVALID; }
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  StreamCtl c;
  PTAM_CUDA_TRY(t, cudaMemcpy(&c, &t->ctl.p[stream], sizeof(c), cudaMemcpyDeviceToHost));
  const int n = c.n_coarse + c.n_l3 + c.n_fine;
  if (idx && n && cap)
    PTAM_CUDA_TRY(t, cudaMemcpy(idx, t->iter_idx.p + (size_t)stream * t->cap, sizeof(int) * std::min(n, cap), cudaMemcpyDeviceToHost));
  return n;
}

// ---- PatchFinder / CalcPoseUpdate unit entry points (finer-grained entries for unit parity) ----
// PatchFinder steps 1-5 (reference include/PatchFinder.h:54-98) for every map point of every stream, against the
// streams' CURRENT frame (the pyramid + corners of the last make_keyframes / track call), at the given poses.
int ptam_patch_search_batch(ptam_tracker* t, const double* se3, unsigned range, int subpix_its) {
  cudaSetDevice(t->device);
  if (!t->dev.src.l0) { t->set_error("no current frame: call ptam_tracker_make_keyframes (or track) first"); return PTAM_ERR_INVALID; }
  if (!se3 || subpix_its < 0) { t->set_error("bad arguments"); return PTAM_ERR_INVALID; }
  if (!t->refind_pose.p) PTAM_CUDA_TRY(t, t->refind_pose.alloc((size_t)12 * t->S));
  PTAM_CUDA_TRY(t, cudaMemcpyAsync(t->refind_pose.p, se3, sizeof(double) * 12 * t->S, cudaMemcpyHostToDevice, t->stream));
  TrackerDev d = t->dev;
  d.mode = 2;
  d.refind_pose = t->refind_pose.p;
  d.unit_range = range; d.unit_subpix_its = subpix_its;
  int maxn = 0;
  for (int s = 0; s < t->S; s++) maxn = std::max(maxn, t->h_pt_count[s]);
  k_pvs_select<<<t->S, pvs_threads(), 0, t->stream>>>(d);
  t->launches++;
  if (maxn > 0) {
    k_search_prep<<<dim3((maxn + 127) / 128, t->S), 128, 0, t->stream>>>(d, 1);
    k_search<<<dim3((maxn + 3) / 4, t->S), 128, 0, t->stream>>>(d, 1);
    t->launches += 2;
  }
  PTAM_CUDA_TRY(t, cudaGetLastError());
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  t->unit_searched = true;
  return PTAM_OK;
}

// Per-point results of the last ptam_patch_search_batch; any output may be NULL.  Returns the point count.
int ptam_patch_get_results(ptam_tracker* t, int stream, int32_t* level, double* warp_inverse, int32_t* template_bad, int32_t* found,
                           double* pos, int32_t* subpix_converged) {
  cudaSetDevice(t->device);
  if (stream < 0 || stream >= t->S) { t->set_error("bad stream"); return PTAM_ERR_INVALID; }
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  const int n = t->h_pt_count[stream];
  const size_t o = (size_t)stream * t->cap;
  if (!n) return 0;
  std::vector<int> fl(n), lv(n);
  std::vector<double> wi(4 * (size_t)n), vf(2 * (size_t)n);
  PTAM_CUDA_TRY(t, cudaMemcpy(fl.data(), t->flags.p + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  PTAM_CUDA_TRY(t, cudaMemcpy(lv.data(), t->level.p + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  PTAM_CUDA_TRY(t, cudaMemcpy(wi.data(), t->warp_inv.p + 4 * o, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost));
  PTAM_CUDA_TRY(t, cudaMemcpy(vf.data(), t->v2found.p + 2 * o, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) {
    const int f = fl[i];
    // projected into the frame in this call: it entered the PVS, or its warp was judged inappropriate (level -1, never searched)
    const bool projected = lv[i] >= 0 || (f & F_IN_IMAGE);
    const bool fnd = (f & F_IN_PVS) && (f & F_FOUND);
    if (level) level[i] = lv[i];
    if (warp_inverse) for (int q = 0; q < 4; q++) warp_inverse[4 * i + q] = projected ? wi[4 * (size_t)i + q] : 0.0;
    if (template_bad) template_bad[i] = (projected && (f & F_TEMPLATE_BAD)) ? 1 : 0;
    if (found) found[i] = fnd ? 1 : 0;
    if (pos) { pos[2 * i] = fnd ? vf[2 * (size_t)i] : 0.0; pos[2 * i + 1] = fnd ? vf[2 * (size_t)i + 1] : 0.0; }
    if (subpix_converged) subpix_converged[i] = (fnd && (f & F_SUBPIX)) ? 1 : 0;
  }
  return n;
}

// Tracker::CalcPoseUpdate (Tracker.cc:928-1005) once per stream over the points FOUND by the last
// ptam_patch_search_batch, with the projections and Jacobians of that call's poses (Tracker.h:125-136).
int ptam_pose_update(ptam_tracker* t, double override_sigma_squared, int mark_outliers, double* mu6, int32_t* n_found) {
  cudaSetDevice(t->device);
  if (!t->unit_searched) { t->set_error("ptam_pose_update works on the result of ptam_patch_search_batch: call it first"); return PTAM_ERR_INVALID; }
  if (!t->unit_mu.p) {
    PTAM_CUDA_TRY(t, t->unit_mu.alloc((size_t)6 * t->S));
    PTAM_CUDA_TRY(t, t->unit_nfound.alloc((size_t)t->S));
  }
  TrackerDev d = t->dev;
  d.mode = 2;
  d.refind_pose = t->refind_pose.p;
  d.unit_override_sigma = override_sigma_squared; d.unit_mark = mark_outliers ? 1 : 0;
  d.unit_mu = t->unit_mu.p; d.unit_nfound = t->unit_nfound.p;
  k_pose<<<t->S, kPoseThreads, d.pose_ws_smem ? kPoseSmemBytes : 0, t->stream>>>(d, 1);
  t->launches++;
  PTAM_CUDA_TRY(t, cudaGetLastError());
  PTAM_CUDA_TRY(t, cudaStreamSynchronize(t->stream));
  if (mu6) PTAM_CUDA_TRY(t, cudaMemcpy(mu6, t->unit_mu.p, sizeof(double) * 6 * t->S, cudaMemcpyDeviceToHost));
  if (n_found) PTAM_CUDA_TRY(t, cudaMemcpy(n_found, t->unit_nfound.p, sizeof(int) * t->S, cudaMemcpyDeviceToHost));
  return PTAM_OK;
}

}  // extern "C"
