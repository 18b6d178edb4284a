// Host side of path B behind the C-ABI: class-Bundle-shaped ingest (AddCamera/AddPoint/AddMeas),
// device graph construction (CSR by point), and the LM control loop of Bundle::Compute /
// Do_LM_Step (reference src/Bundle.cc:116-158,209-551).  The host reads back two scalars per
// lambda trial (new error, squared update) to take the accept / reject / converge decisions exactly
// where the reference takes them; everything else stays on the device.
#include "bundle_kernels.cuh"
#include "ldlt.h"

#include <dlfcn.h>
#include <nccl.h>  // types and enums only: the symbols are resolved with dlopen (torch's bundled
                   // libnccl.so.2 when the process already loaded it, else the system library)

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <type_traits>
#include <unordered_map>
#include <vector>

using namespace ptam;

namespace {
struct NcclApi {
  void* lib = nullptr;
  bool tried = false;
  std::string err;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (tried) return lib != nullptr;
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
    AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(lib, "ncclAllGather"));
    Broadcast = reinterpret_cast<decltype(Broadcast)>(dlsym(lib, "ncclBroadcast"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !AllGather || !Broadcast || !GroupStart || !GroupEnd || !GetErrorString) { err = "NCCL library lacks required symbols"; lib = nullptr; return false; }
    return true;
  }
};
NcclApi& nccl_api() { static NcclApi a; return a; }

// One peer window per communicator (kept for the communicator's life, grown when a larger system comes): the memory
// the ranks of one node exchange S through (k_ba_peer_*).  Built collectively; any failure (no peer access, another
// node, IPC refused) leaves `ok` false on EVERY rank, and the handles keep using NCCL.
struct PeerWindow {
  bool tried = false, ok = false;
  int rank = 0, world = 0, device = 0;
  size_t doubles = 0;          // capacity of the data part
  double* local = nullptr;     // cudaMalloc: [doubles] + flag block
  double* peer[kPeerMax] = {};
  unsigned* flags[kPeerMax] = {};
  int* err = nullptr;          // device int, local
  unsigned epoch = 0;
  void release() {
    for (int p = 0; p < world; p++) if (p != rank && peer[p]) cudaIpcCloseMemHandle(peer[p]);
    if (local) cudaFree(local);
    if (err) cudaFree(err);
    *this = PeerWindow();
  }
};
std::mutex g_win_mutex;
std::unordered_map<void*, PeerWindow>& peer_windows() { static std::unordered_map<void*, PeerWindow> m; return m; }

// Contiguous point ranges balanced by measurement count: shard r owns points [begin[r], begin[r+1]).
void shard_plan(int n_points, int n_meas, const int32_t* meas_point, int world, int32_t* begin, int* per_shard = nullptr) {
  std::vector<int> cnt(n_points + 1, 0);
  for (int m = 0; m < n_meas; m++) cnt[meas_point[m] + 1]++;
  for (int i = 0; i < n_points; i++) cnt[i + 1] += cnt[i];
  begin[0] = 0;
  for (int r = 1; r < world; r++) {
    const int target = (int)((long long)n_meas * r / world);
    int b = (int)(std::lower_bound(cnt.begin(), cnt.end(), target) - cnt.begin());
    b = std::min(std::max(b, (int)begin[r - 1]), n_points);
    begin[r] = b;
  }
  begin[world] = n_points;
  if (per_shard) for (int r = 0; r < world; r++) per_shard[r] = cnt[begin[r + 1]] - cnt[begin[r]];
}
}  // namespace

ptam::CamModel ptam_make_cam_model(const double* p, double W, double H);
void ptam_set_global_error(const std::string& e);

namespace {
// A typed view into the handle's device arena (one cudaMalloc per graph size, reused by later
// Compute() calls on the same handle; MapMaker builds a Bundle per adjustment, Bundle.cc:35).
template <class T>
struct Buf {
  T* p = nullptr;
};
}  // namespace

static void release_window(void* comm) {
  std::lock_guard<std::mutex> lock(g_win_mutex);
  auto it = peer_windows().find(comm);
  if (it == peer_windows().end()) return;
  cudaSetDevice(it->second.device);
  cudaDeviceSynchronize();
  it->second.release();
  peer_windows().erase(it);
}

struct ptam_bundle {
  int device = 0;
  cudaStream_t stream = nullptr;
  LdltSolver ldlt;
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
  Buf<double> cam_se3, cam_se3_new, U, epsA, pt_pos, pt_pos_new, V, epsB, Vinv, m_found, m_sin, m_v3cam, m_derivs,
      m_eps, m_e2, m_W, m_B, e2c, S, vE, upd, scal, Wp, err_cam, partials;
  Buf<int> cam_fixed, cam_row, pt_off, pt_meas, pt_meas_ins, pt_cam, cam_off, cam_meas_ins, cam_meas_pt, blk_off, blk_cnt, nz_blocks, pair_info, free_cam,
      m_cam, m_pt, m_state, counters, outliers;
  Buf<unsigned> tickets;
  Buf<int> csr_cur, cam_order;
  PeerWindow* win = nullptr;       // sharded handles on one node: S / vE live in the communicator's peer window
  int peer_row_lo = 0, peer_row_hi = 0;
  Buf<double> s_pack, sel_gather;  // sharded handles: packed lower triangle of S + vE; every shard's squared errors
  int sel_slot = 0;                // measurements per slot of sel_gather (the largest shard's count)
  std::vector<int32_t> plan;       // points [plan[r], plan[r + 1]) belong to shard r
  Buf<long long> csr_pairs;
  int* pair_buf = nullptr;      // pr_mj | pr_mk: sized by the device's own count of co-visible triples
  size_t pair_cap = 0;
  double* h_scal = nullptr;  // pinned
  int* h_cnt = nullptr;      // pinned
  bool cnt_pending = false;  // h_cnt is on its way (lm_step queued the copy)
  // multi-GPU shard: points [p_lo, p_hi) and their measurements live here; cameras are replicated
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  bool own_comm = false;
  int p_lo = 0, p_hi = 0;
  std::vector<int> l_gid;            // local measurement -> insertion index
  Buf<int> m_gid, m_erase_step, g_steps, g_pairs, g_cnt, hist16, erase_cnt;
  Buf<unsigned long long> sel_state;
  bool shards_dirty = false, abort_seen = false;
  std::vector<int> h_outliers;       // merged (point, camera) pairs in the reference's erase order
  int n_meas_local = 0;
  unsigned char* arena = nullptr;
  size_t arena_cap = 0;
  // One lambda trial of a single-GPU handle is a fixed sequence of launches and two small copies: it is captured
  // once per handle into a CUDA graph and replayed (env PTAM_B200_NO_GRAPH=1: plain launches).
  cudaGraphExec_t trial_exec = nullptr;
  BundleDev trial_key{};
  int64_t trial_launches = 0;
  bool graph_off = false;
  // optional per-phase device timing (CUDA events on the handle's stream around each phase)
  bool profiling = false;
  cudaEvent_t prof_ev[2 * PTAM_BA_PHASES] = {};
  double prof_ms[PTAM_BA_PHASES] = {};
  int64_t prof_n[PTAM_BA_PHASES] = {};
  unsigned prof_pending = 0;
  double prof_begin_ms = 0, prof_wall_ms = 0;  // host wall clock: graph build + upload, whole Compute
  void pbegin(int k) { if (profiling) cudaEventRecord(prof_ev[2 * k], stream); }
  void pend(int k) { if (profiling) { cudaEventRecord(prof_ev[2 * k + 1], stream); prof_pending |= 1u << k; } }
  void pcollect() {  // call after a stream synchronisation
    if (!profiling) return;
    for (int k = 0; k < PTAM_BA_PHASES; k++)
      if (prof_pending & (1u << k)) { float ms = 0; if (cudaEventElapsedTime(&ms, prof_ev[2 * k], prof_ev[2 * k + 1]) == cudaSuccess) { prof_ms[k] += ms; prof_n[k]++; } }
    prof_pending = 0;
  }

  void set_error(const std::string& e) { err = e; }

  // Device memory comes from the device's stream-ordered pool, kept warm (release threshold = everything): MapMaker
  // builds a new Bundle per adjustment (Bundle.cc:35), and a cold cudaMalloc of the arena costs 0.6 - 16 ms.
  cudaError_t dev_alloc(void** p, size_t bytes) {
    static std::once_flag once[64];
    std::call_once(once[device & 63], [&] {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
    });
    return cudaMallocAsync(p, bytes, stream);
  }
  void dev_free(void* p) { if (p) cudaFreeAsync(p, stream); }

  ~ptam_bundle() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    dev_free(arena);
    dev_free(pair_buf);
    if (stream) cudaStreamSynchronize(stream);
    if (trial_exec) cudaGraphExecDestroy(trial_exec);
    for (auto e : prof_ev) if (e) cudaEventDestroy(e);
    if (comm && own_comm) { release_window((void*)comm); nccl_api().CommDestroy(comm); }
    if (h_scal) cudaFreeHost(h_scal);
    if (h_cnt) cudaFreeHost(h_cnt);
    ldlt.destroy();
    if (stream) cudaStreamDestroy(stream);
  }

  int nccl_try(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return PTAM_OK;
    set_error(std::string(what) + ": " + nccl_api().GetErrorString(r));
    return PTAM_ERR_NCCL;
  }
  int all_reduce(void* buf, size_t count, ncclDataType_t type, ncclRedOp_t op, const char* what) {
    if (world == 1) return PTAM_OK;
    return nccl_try(nccl_api().AllReduce(buf, buf, count, type, op, comm, stream), what);
  }

  int init(int dev, const double* cam_params, int w, int h, const ptam_bundle_params* p) {
    device = dev;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: the B200 path has no CPU fallback"); return PTAM_ERR_NO_DEVICE; }
    if (dev < 0 || dev >= ndev) { set_error("bad device index"); return PTAM_ERR_INVALID; }
    PTAM_CUDA_TRY(this, cudaSetDevice(dev));
    PTAM_CUDA_TRY(this, cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    PTAM_CUDA_TRY(this, cudaMallocHost(&h_scal, 16 * sizeof(double)));
    PTAM_CUDA_TRY(this, cudaMallocHost(&h_cnt, 8 * sizeof(int)));
    if (p) prm = *p; else ptam_bundle_default_params(&prm);
    cam = ptam_make_cam_model(cam_params, w, h);
    PTAM_CUDA_TRY(this, cudaFuncSetAttribute(k_ba_acc_cam, cudaFuncAttributeMaxDynamicSharedMemorySize, kAccCamSmem));
    PTAM_CUDA_TRY(this, cudaFuncSetAttribute(k_ba_schur_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kSchurDiagSmem));
    if (ldlt.init(stream) != cudaSuccess) { set_error("dense solver set-up failed: " + ldlt.err); return PTAM_ERR_CUDA; }
    return PTAM_OK;
  }

  int n_cams() const { return (int)h_cam_fixed.size(); }
  int n_pts() const { return (int)h_pts.size() / 3; }
  int n_meas() const { return (int)h_mcam.size(); }

  // Compute() before its loop: GenerateMeasLUTs / GenerateOffDiagScripts become a CSR by point with
  // each point's measurements sorted by camera id (std::set<int> order, Bundle.h:69).  A sharded
  // handle keeps only the measurements of the points it owns.
  int begin() {
    cudaSetDevice(device);
    static const bool dbg_t = std::getenv("PTAM_B200_DEBUG_TIMES") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!dbg_t) return;
      const auto now = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[ptam dbg] bundle begin: %-28s %.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
      t_last = now;
    };
    if (world > 1 && !comm) { set_error("sharded handle without a communicator: call ptam_bundle_init_shard first"); return PTAM_ERR_NCCL; }
    const int C = n_cams(), P = n_pts(), MG = n_meas();
    const int n = 6 * n_free;
    if (MG == 0) { set_error("no measurements (the reference asserts on this, Tools.h:155)"); return PTAM_ERR_INVALID; }
    p_lo = 0; p_hi = P;
    plan.assign(world + 1, 0);
    plan[world] = P;
    sel_slot = 0;
    if (world > 1) {
      std::vector<int> per_rank(world, 0);  // measurements per shard
      shard_plan(P, MG, h_mpt.data(), world, plan.data(), per_rank.data());
      p_lo = plan[rank]; p_hi = plan[rank + 1];
      sel_slot = *std::max_element(per_rank.begin(), per_rank.end());
      n_meas_local = per_rank[rank];
    }
    // a single-GPU handle uses the insertion-order arrays as they are; a shard copies out its own
    l_gid.clear();
    std::vector<int> l_mcam, l_mpt;
    std::vector<double> l_found, l_sin;
    const bool whole = world == 1;
    if (whole) {
      l_gid.resize(MG);
      std::iota(l_gid.begin(), l_gid.end(), 0);
    } else {
      const int ML = n_meas_local;
      l_gid.resize(ML); l_mcam.resize(ML); l_mpt.resize(ML); l_found.resize(2 * (size_t)ML); l_sin.resize(ML);
      int o = 0;
      for (int m = 0; m < MG; m++) {
        const int pt = h_mpt[m];
        if (pt < p_lo || pt >= p_hi) continue;
        l_gid[o] = m; l_mcam[o] = h_mcam[m]; l_mpt[o] = pt;
        l_found[2 * (size_t)o] = h_found[2 * (size_t)m]; l_found[2 * (size_t)o + 1] = h_found[2 * (size_t)m + 1]; l_sin[o] = h_sin[m];
        o++;
      }
    }
    const std::vector<int>& v_mcam = whole ? h_mcam : l_mcam;
    const std::vector<int>& v_mpt = whole ? h_mpt : l_mpt;
    const std::vector<double>& v_found = whole ? h_found : l_found;
    const std::vector<double>& v_sin = whole ? h_sin : l_sin;
    const int M = (int)l_gid.size();
    n_meas_local = M;
    // The usual list (MapMaker.cc:871-882 walks the keyframes, and a keyframe's measurements by map point) is sorted
    // by (camera, point): one sequential pass finds out, and counts the measurements per camera on the way.  Then
    // list order = ascending camera id inside a point and ascending point id inside a camera, the camera CSR is the
    // identity, and the point CSR is built on the device (k_ba_csr_*).  Any other list takes the host path below.
    std::vector<int> coff(C + 1, 0);
    bool fast = M > 0;
    for (int m = 0; m < M; m++) {
      coff[v_mcam[m] + 1]++;
      if (m && !(v_mcam[m] > v_mcam[m - 1] || (v_mcam[m] == v_mcam[m - 1] && v_mpt[m] > v_mpt[m - 1]))) fast = false;
    }
    for (int j = 0; j < C; j++) coff[j + 1] += coff[j];
    if (std::getenv("PTAM_B200_HOST_CSR")) fast = false;
    std::vector<int> off, idx, idx_ins, ptcam, cidx, cidx_pt;
    long long n_pairs_max = 0;
    if (!fast) {
    // CSR by point.  The bucket pass is stable, so `idx` first holds every point's measurements in LIST order
    // (the order V_i / epsB_i are summed in); the off-diagonal scripts want them by ascending camera id.
    off.assign(P + 1, 0); idx.resize(M);
    for (int m = 0; m < M; m++) off[v_mpt[m] + 1]++;
    for (int i = 0; i < P; i++) off[i + 1] += off[i];
    {
      std::vector<int> cur(off.begin(), off.end() - 1);
      for (int m = 0; m < M; m++) idx[cur[v_mpt[m]]++] = m;
      for (int i = 0; i < P; i++) {
        bool sorted = true;
        for (int o = off[i] + 1; o < off[i + 1]; o++) if (v_mcam[idx[o]] < v_mcam[idx[o - 1]]) { sorted = false; break; }
        if (!sorted) {
          if (idx_ins.empty()) idx_ins = idx;
          std::sort(idx.begin() + off[i], idx.begin() + off[i + 1], [&](int a, int b) { return v_mcam[a] < v_mcam[b]; });
        }
        for (int o = off[i] + 1; o < off[i + 1]; o++)
          if (v_mcam[idx[o]] == v_mcam[idx[o - 1]]) { set_error("duplicate (camera, point) measurement"); return PTAM_ERR_INVALID; }
        const long long k = off[i + 1] - off[i];
        n_pairs_max += k * (k - 1) / 2;
      }
    }
    ptcam.resize(M);
    for (int o = 0; o < M; o++) ptcam[o] = v_mcam[idx[o]];
    // CSR by camera: list order (U_j / epsA_j) and ascending point id (S_jj, the pair list); again one array
    // when the list is point-ordered inside every camera
    cidx.resize(M);
    {
      std::vector<int> cur(coff.begin(), coff.end() - 1);
      for (int m = 0; m < M; m++) cidx[cur[v_mcam[m]]++] = m;
      for (int j = 0; j < C; j++) {
        bool sorted = true;
        for (int o = coff[j] + 1; o < coff[j + 1]; o++) if (v_mpt[cidx[o]] < v_mpt[cidx[o - 1]]) { sorted = false; break; }
        if (!sorted) {
          if (cidx_pt.empty()) cidx_pt = cidx;
          std::sort(cidx_pt.begin() + coff[j], cidx_pt.begin() + coff[j + 1], [&](int a, int b) { return v_mpt[a] < v_mpt[b]; });
        }
      }
    }
    }
    lap("measurement lists (host)");
    std::vector<int> freecam;
    for (int j = 0; j < C; j++) if (!h_cam_fixed[j]) freecam.push_back(j);
    std::vector<int> camorder(C);  // per-camera kernels take the cameras with the most measurements first
    std::iota(camorder.begin(), camorder.end(), 0);
    std::stable_sort(camorder.begin(), camorder.end(), [&](int a, int b) { return coff[a + 1] - coff[a] > coff[b + 1] - coff[b]; });
    const long long n_blocks = (long long)n_free * (n_free - 1) / 2;
    if (n_pairs_max > 0x7fffffffLL || n_blocks > 0x7ffffff0LL) { set_error("graph too dense for the pair list (more than 2^31 camera pairs / co-visible triples)"); return PTAM_ERR_INVALID; }
    const int grid_max = std::max({(M + 255) / 256, (P + 255) / 256, 1});
    win = n > 0 ? peer_window(n) : nullptr;  // collective on a sharded handle
    if (win) {  // rows of the lower triangle per rank, equal element counts: row b_k = n sqrt(k / world)
      auto bound = [&](int k) { return k >= world ? n : (int)std::lround(n * std::sqrt((double)k / world)); };
      peer_row_lo = bound(rank); peer_row_hi = bound(rank + 1);
    }
    // one arena for everything: two passes over the same layout (measure, then assign)
    size_t need = 0;
    for (int pass = 0; pass < 2; pass++) {
      size_t off = 0;
      auto take = [&](auto& buf, size_t cnt) {
        using T = std::remove_pointer_t<decltype(buf.p)>;
        if (pass) buf.p = reinterpret_cast<T*>(arena + off);
        off += (std::max<size_t>(cnt, 1) * sizeof(T) + 255) & ~(size_t)255;
      };
#define AL(buf, cnt) take(buf, (size_t)(cnt))
      AL(cam_se3, 12 * (size_t)C); AL(cam_se3_new, 12 * (size_t)C); AL(U, 21 * (size_t)C); AL(epsA, 6 * (size_t)C);
      AL(cam_fixed, C); AL(cam_row, C);
      AL(pt_pos, 3 * (size_t)P); AL(pt_pos_new, 3 * (size_t)P); AL(V, 6 * (size_t)P); AL(epsB, 3 * (size_t)P);
      AL(Vinv, 12 * (size_t)P); AL(pt_off, P + 1); AL(pt_meas, M);
      AL(pt_meas_ins, idx_ins.empty() ? 0 : M); AL(pt_cam, M); AL(cam_off, C + 1); AL(cam_meas_ins, M);
      AL(cam_meas_pt, cidx_pt.empty() ? 0 : M); AL(blk_off, n_blocks + 1); AL(blk_cnt, n_blocks); AL(nz_blocks, n_blocks); AL(pair_info, 4); AL(free_cam, n_free); AL(cam_order, C); AL(csr_cur, P + 1); AL(csr_pairs, 1);
      AL(m_B, 6 * (size_t)M); AL(err_cam, C); AL(partials, grid_max); AL(tickets, 8);
      AL(m_cam, M); AL(m_pt, M); AL(m_found, 2 * (size_t)M); AL(m_sin, M); AL(m_state, M); AL(m_v3cam, 4 * (size_t)M);
      AL(m_derivs, 4 * (size_t)M); AL(m_eps, 2 * (size_t)M); AL(m_e2, M); AL(m_W, 18 * (size_t)M); AL(e2c, M);
      AL(S, win ? 0 : (size_t)n * n); AL(vE, win ? 0 : n); AL(upd, n); AL(scal, 8); AL(counters, 4); AL(outliers, 2 * (size_t)M);
      AL(Wp, ldlt_workspace_doubles(n));
      AL(m_gid, M); AL(m_erase_step, M); AL(hist16, kSelBins); AL(erase_cnt, (M + 1023) / 1024 + 1); AL(sel_state, 2);
      AL(g_steps, world > 1 ? MG : 0); AL(g_pairs, world > 1 ? 2 * (size_t)MG : 0); AL(g_cnt, 1);
      AL(s_pack, world > 1 ? (size_t)n * (n + 1) / 2 + n + 8 : 0); AL(sel_gather, world > 1 ? (size_t)world * sel_slot : 0);
#undef AL
      if (!pass) {
        need = off;
        if (need > arena_cap) {
          if (arena) { dev_free(arena); arena = nullptr; arena_cap = 0; }
          PTAM_CUDA_TRY(this, dev_alloc(reinterpret_cast<void**>(&arena), need));
          arena_cap = need;
        }
      }
    }
    if (win) { S.p = win->local; vE.p = win->local + (size_t)n * n; }
    lap("arena (cudaMalloc if grown)");
    PTAM_CUDA_TRY(this, cudaMemsetAsync(arena, 0, need, stream));
    PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
    lap("memset + sync");
#define UP(buf, vec) if (!vec.empty()) PTAM_CUDA_TRY(this, cudaMemcpy(buf.p, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice))
    UP(cam_se3, h_cam_se3); UP(cam_fixed, h_cam_fixed); UP(cam_row, h_cam_row); UP(pt_pos, h_pts);
    UP(pt_off, off); UP(pt_meas, idx); UP(m_cam, v_mcam); UP(m_pt, v_mpt); UP(m_found, v_found); UP(m_sin, v_sin);
    if (!whole) UP(m_gid, l_gid);
    UP(pt_meas_ins, idx_ins); UP(pt_cam, ptcam); UP(cam_off, coff); UP(cam_meas_ins, cidx); UP(cam_meas_pt, cidx_pt); UP(free_cam, freecam); UP(cam_order, camorder);
#undef UP
    // pageable H2D copies return once staged; the handle's streams are non-blocking (no implicit ordering with
    // the legacy stream the copies ran on), so finish them before the first kernel is queued
    PTAM_CUDA_TRY(this, cudaStreamSynchronize(cudaStreamLegacy));
    lap("uploads");
    d.cam = cam; d.n_cams = C; d.n_pts = P; d.n_meas = M; d.n = n; d.est = prm.mestimator;
    d.p_lo = p_lo; d.p_hi = p_hi; d.add_cam_update = rank == 0 ? 1 : 0;
    d.cam_se3 = cam_se3.p; d.cam_se3_new = cam_se3_new.p; d.cam_fixed = cam_fixed.p; d.cam_row = cam_row.p;
    d.U = U.p; d.epsA = epsA.p; d.pt_pos = pt_pos.p; d.pt_pos_new = pt_pos_new.p; d.V = V.p; d.epsB = epsB.p;
    d.Vinv = Vinv.p; d.pt_off = pt_off.p; d.pt_meas = pt_meas.p; d.m_cam = m_cam.p; d.m_pt = m_pt.p;
    d.m_found = m_found.p; d.m_sin = m_sin.p; d.m_state = m_state.p; d.m_v3cam = m_v3cam.p; d.m_derivs = m_derivs.p;
    d.m_eps = m_eps.p; d.m_e2 = m_e2.p; d.m_W = m_W.p; d.e2_compact = e2c.p; d.S = S.p; d.vE = vE.p; d.upd = upd.p;
    d.scal = scal.p; d.counters = counters.p; d.outliers = outliers.p;
    d.hist16 = hist16.p; d.sel_state = sel_state.p; d.m_erase_step = m_erase_step.p;
    d.sel_keys = world > 1 ? sel_gather.p : nullptr; d.sel_n = world * sel_slot;
    d.m_B = m_B.p; d.pt_meas_ins = idx_ins.empty() ? pt_meas.p : pt_meas_ins.p; d.pt_cam = pt_cam.p;
    d.cam_off = cam_off.p; d.cam_meas_ins = cam_meas_ins.p; d.cam_meas_pt = cidx_pt.empty() ? cam_meas_ins.p : cam_meas_pt.p;
    d.n_blocks = n_blocks; d.blk_off = blk_off.p;
    d.err_cam = err_cam.p; d.partials = partials.p; d.tickets = tickets.p;
    d.nz_blocks = nz_blocks.p; d.pair_info = pair_info.p; d.free_cam = free_cam.p; d.cam_order = cam_order.p;
    if (M > 0) {
      const int gm = (M + 255) / 256;
      if (whole) { k_ba_iota<<<gm, 256, 0, stream>>>(m_gid.p, M); launches++; }
      if (fast) {  // the point CSR on the device; the camera CSR is the identity
        k_ba_csr_count<<<gm, 256, 0, stream>>>(m_pt.p, M, csr_cur.p);
        k_ba_csr_scan<<<1, 1024, 0, stream>>>(csr_cur.p, pt_off.p, P, csr_pairs.p);
        k_ba_csr_fill<<<gm, 256, 0, stream>>>(m_pt.p, M, pt_off.p, csr_cur.p, pt_meas.p);
        k_ba_csr_sort<<<(P + 255) / 256, 256, 0, stream>>>(pt_off.p, P, pt_meas.p, m_cam.p, pt_cam.p);
        k_ba_iota<<<gm, 256, 0, stream>>>(cam_meas_ins.p, M);
        launches += 5;
      }
    }
    // the pair-major list of the off-diagonal blocks (GenerateOffDiagScripts, Bundle.cc:572-599), on the device
    d.pr_mj = d.pr_mk = nullptr;
    if (n_blocks > 0 && p_hi > p_lo) {
      k_ba_pair_count<<<(p_hi - p_lo + 255) / 256, 256, 0, stream>>>(d, blk_cnt.p);
      k_ba_pair_scan<<<1, 1024, 0, stream>>>(blk_cnt.p, blk_off.p, nz_blocks.p, pair_info.p, n_blocks);
      launches += 2;
      // the list itself is sized by the count just made (one scalar read-back)
      long long bound = n_pairs_max;
      int info[4] = {0, 0, 0, 0};
      if (fast) PTAM_CUDA_TRY(this, cudaMemcpyAsync(&bound, csr_pairs.p, sizeof(long long), cudaMemcpyDeviceToHost, stream));
      PTAM_CUDA_TRY(this, cudaMemcpyAsync(info, pair_info.p, sizeof(info), cudaMemcpyDeviceToHost, stream));
      PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
      if (bound > 0x7fffffffLL) { set_error("graph too dense for the pair list (more than 2^31 co-visible triples)"); return PTAM_ERR_INVALID; }
      const size_t triples = (size_t)std::max(info[2], 1);
      if (2 * triples > pair_cap) {
        dev_free(pair_buf);
        pair_buf = nullptr; pair_cap = 0;
        PTAM_CUDA_TRY(this, dev_alloc(reinterpret_cast<void**>(&pair_buf), 2 * triples * sizeof(int)));
        pair_cap = 2 * triples;
      }
      d.pr_mj = pair_buf; d.pr_mk = pair_buf + triples;
      k_ba_pair_fill<<<148 * 8, 128, 0, stream>>>(d, pair_buf, pair_buf + triples);
      launches++;
      PTAM_CUDA_TRY(this, cudaGetLastError());
    }
    lap("device lists");
    lambda = 0.0001; lambda_factor = 2.0;
    converged = false; hit_max = false; abort_seen = false;
    counter = 0; accepted = 0; lm_steps = 0; n_outliers = 0;
    h_outliers.clear();
    shards_dirty = false;
    begun = true;
    return PTAM_OK;
  }

  // sharded handles: sum of the shards' partial S and vE on every shard.  Only the lower triangle is exchanged (the
  // solver reads nothing else), packed row by row with vE behind it: one all-reduce of n (n + 1) / 2 + n doubles.
  int exchange_reduced(double* extra = nullptr, int n_extra = 0) {
    const int n = d.n;
    if (win) {  // one node: through NVLink peer memory (k_ba_peer_*), S and vE are reduced in place in every rank's window
      const PeerWin v = peer_view();
      const unsigned ea = ++win->epoch, eb = ++win->epoch;
      k_ba_peer_reduce<<<148 * 2, 256, 0, stream>>>(v, ea, eb, peer_row_lo, peer_row_hi, extra, extra, n_extra);
      k_ba_peer_wait<<<1, 32, 0, stream>>>(v, eb);
      launches += 2;
      PTAM_CUDA_TRY(this, cudaGetLastError());
      return PTAM_OK;
    }
    k_ba_pack_lower<<<148 * 4, 256, 0, stream>>>(d.S, d.vE, n, s_pack.p, extra, n_extra);
    int rc = all_reduce(s_pack.p, (size_t)n * (n + 1) / 2 + n + n_extra, ncclDouble, ncclSum, "all-reduce of the packed S and vE");
    if (rc) return rc;
    k_ba_unpack_lower<<<148 * 4, 256, 0, stream>>>(s_pack.p, n, d.S, d.vE, extra, n_extra);
    launches += 2;
    PTAM_CUDA_TRY(this, cudaGetLastError());
    return PTAM_OK;
  }

  // Collective.  Returns the communicator's peer window with room for an n x n system, or nullptr (NCCL path).
  PeerWindow* peer_window(int n) {
    // measured (C4, 36 MB triangle): 2 ranks 0.10 ms against 0.15 ms for pack + ncclAllReduce + unpack; from 4 ranks on the
    // owner-reduces scheme moves 2 (N - 1) / N of the triangle per rank and direction and NCCL (NVLS) is as fast or
    // faster (0.20 against 0.19 ms at 4), so larger worlds stay on NCCL unless PTAM_B200_PEER_MAX_WORLD says otherwise
    static const int max_world = [] { const char* e = std::getenv("PTAM_B200_PEER_MAX_WORLD"); return e ? std::atoi(e) : 2; }();
    if (world < 2 || world > kPeerMax || world > max_world || std::getenv("PTAM_B200_NO_PEER")) return nullptr;
    std::lock_guard<std::mutex> lock(g_win_mutex);
    PeerWindow& w = peer_windows()[(void*)comm];
    const size_t need = (size_t)n * n + n + (size_t)kPeerMax * kPeerExtra;
    if (w.tried && (!w.ok || w.doubles >= need)) return w.ok ? &w : nullptr;
    // (re)build: every rank takes this branch together, n is the same everywhere
    const unsigned keep_epoch = w.epoch;
    if (w.tried) { cudaStreamSynchronize(stream); w.release(); }
    w.tried = true; w.rank = rank; w.world = world; w.device = device; w.epoch = keep_epoch;
    const size_t flag_bytes = sizeof(unsigned) * 2 * kPeerMax;
    int good = 1;
    cudaIpcMemHandle_t mine{};
    if (cudaMalloc(&w.local, need * sizeof(double) + flag_bytes) != cudaSuccess) { good = 0; w.local = nullptr; cudaGetLastError(); }
    if (good && cudaMalloc(&w.err, 2 * sizeof(int)) != cudaSuccess) { good = 0; w.err = nullptr; cudaGetLastError(); }
    if (good && cudaIpcGetMemHandle(&mine, w.local) != cudaSuccess) { good = 0; cudaGetLastError(); }
    // the handles (and whether everybody got this far) travel through the communicator
    struct Slot { cudaIpcMemHandle_t h; int good; int pad[15]; };
    static_assert(sizeof(Slot) == 128, "slot size");
    Slot* d_slots = nullptr;
    std::vector<Slot> h_slots(world);
    if (cudaMalloc(&d_slots, sizeof(Slot) * world) != cudaSuccess) { cudaGetLastError(); w.ok = false; return nullptr; }
    Slot me{}; me.h = mine; me.good = good;
    cudaMemcpyAsync(d_slots + rank, &me, sizeof(Slot), cudaMemcpyHostToDevice, stream);
    if (w.local) cudaMemsetAsync(w.local, 0, need * sizeof(double) + flag_bytes, stream);
    if (w.err) cudaMemsetAsync(w.err, 0, 2 * sizeof(int), stream);
    nccl_api().AllGather(d_slots + rank, d_slots, sizeof(Slot), ncclChar, comm, stream);
    cudaMemcpyAsync(h_slots.data(), d_slots, sizeof(Slot) * world, cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    cudaFree(d_slots);
    for (int p = 0; p < world; p++) good &= h_slots[p].good;
    if (good) {
      for (int p = 0; p < world && good; p++) {
        if (p == rank) { w.peer[p] = w.local; continue; }
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, h_slots[p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { good = 0; cudaGetLastError(); }
        w.peer[p] = (double*)ptr;
      }
    }
    // second round: did everybody map everybody?  (also orders the zeroing above before anybody's first signal)
    int* d_good = nullptr;
    if (cudaMalloc(&d_good, sizeof(int)) == cudaSuccess) {
      cudaMemcpyAsync(d_good, &good, sizeof(int), cudaMemcpyHostToDevice, stream);
      nccl_api().AllReduce(d_good, d_good, 1, ncclInt32, ncclMin, comm, stream);
      cudaMemcpyAsync(&good, d_good, sizeof(int), cudaMemcpyDeviceToHost, stream);
      cudaStreamSynchronize(stream);
      cudaFree(d_good);
    } else { good = 0; cudaGetLastError(); }
    if (!good) { const unsigned e = w.epoch; w.release(); w.tried = true; w.epoch = e; return nullptr; }
    w.doubles = need;
    for (int p = 0; p < world; p++) w.flags[p] = reinterpret_cast<unsigned*>(w.peer[p] + need);
    w.ok = true;
    if (std::getenv("PTAM_B200_PEER_TEST") && rank == 0) {  // debug probe: what a plain kernel gets out of the peer mapping
      const size_t n2 = std::min<size_t>(need / 2, (size_t)2 << 20);  // 32 MB
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int mode = 0; mode < 3; mode++) {
        const double2* src = reinterpret_cast<const double2*>(mode == 0 ? w.peer[1] : w.local);
        double2* dst = reinterpret_cast<double2*>(mode == 1 ? w.peer[1] : w.local) + (mode == 2 ? n2 : 0);
        for (int rep = 0; rep < 3; rep++) {
          cudaEventRecord(e0, stream);
          k_ba_peer_copy<<<148 * 4, 256, 0, stream>>>(src, dst + (mode == 0 ? n2 : 0), n2);
          cudaEventRecord(e1, stream);
          cudaStreamSynchronize(stream);
          float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
          if (rep == 2) std::fprintf(stderr, "[ptam dbg] peer window probe: %s 32 MB in %.1f us = %.0f GB/s\n",
                                     mode == 0 ? "read from peer " : mode == 1 ? "write to peer  " : "local copy     ", 1e3 * ms, n2 * 16.0 / (ms * 1e-3) / 1e9);
        }
      }
      cudaMemsetAsync(w.local, 0, need * sizeof(double), stream);
      cudaStreamSynchronize(stream);
    }
    if (std::getenv("PTAM_B200_PEER_TEST")) {  // the exchange itself, back to back, on every rank
      win = &w;
      const int n_keep = d.n; d.n = n;
      auto bound = [&](int k) { return k >= world ? n : (int)std::lround(n * std::sqrt((double)k / world)); };
      peer_row_lo = bound(rank); peer_row_hi = bound(rank + 1);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0, stream);
        exchange_reduced();
        cudaEventRecord(e1, stream);
        cudaStreamSynchronize(stream);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        std::fprintf(stderr, "[ptam dbg] rank %d: exchange %d of the %d x %d lower triangle: %.1f us\n", rank, rep, n, n, 1e3 * ms);
      }
      d.n = n_keep; win = nullptr;
    }
    return &w;
  }

  PeerWin peer_view() const {
    PeerWin v{};
    for (int p = 0; p < world; p++) { v.S[p] = win->peer[p]; v.flags[p] = win->flags[p]; }
    v.rank = rank; v.world = world; v.n = d.n; v.err = win->err;
    return v;
  }

  int solve_reduced() {  // Cholesky<>(mS).backsub(vE), Bundle.cc:457-458
    const int64_t l0 = ldlt.launches;
    const cudaError_t e = ldlt.solve(d.S, d.vE, d.upd, Wp.p, d.n);
    launches += ldlt.launches - l0;
    if (e != cudaSuccess) { set_error("dense solve: " + ldlt.err + ": " + cudaGetErrorString(e)); return PTAM_ERR_CUDA; }
    return PTAM_OK;
  }

  // S and vE of the current lambda (Bundle.cc:365-453): every block of the lower triangle is written whole
  int build_reduced() {
    if (d.n == 0) return PTAM_OK;
    k_ba_zero_lower<<<148 * 4, 256, 0, stream>>>(d.S, d.n, d.tickets + 3);
    k_ba_schur_diag<<<d.n_cams, kSegThreads, kSchurDiagSmem, stream>>>(d);
    launches += 2;
    if (d.n_blocks > 0) {
      k_ba_schur_off<<<148 * 4, kOffThreads, 0, stream>>>(d);
      launches++;
    }
    PTAM_CUDA_TRY(this, cudaGetLastError());
    return PTAM_OK;
  }

  // sigma^2 (Bundle.cc:230-237).  Small single-GPU problems: one-CTA select on the compacted errors;
  // large or sharded ones: 4-pass 16-bit radix select whose histograms are summed across shards.
  int find_sigma_squared() {
    const int M = d.n_meas;
    const double min_s2 = prm.min_tukey_sigma * prm.min_tukey_sigma;
    if (world == 1 && M < 65536) {
      if (M > 0) {
        const int gm = (M + 255) / 256;
        k_ba_gather_e2<<<gm, 256, 0, stream>>>(d);
        k_ba_select<<<1, 1024, 0, stream>>>(d, min_s2);
        launches += 2;
      }
      return PTAM_OK;
    }
    // sharded: ONE exchange, every shard's squared errors to every shard (8 B per measurement), then the same
    // six local passes on the gathered keys: the median of the same multiset, bit for bit, on every shard
    const int n_keys = world > 1 ? world * sel_slot : M;
    if (world > 1 && sel_slot > 0) {
      k_ba_sel_keys<<<(sel_slot + 255) / 256, 256, 0, stream>>>(d, sel_gather.p + (size_t)rank * sel_slot, sel_slot);
      launches++;
      int rc = nccl_try(nccl_api().AllGather(sel_gather.p + (size_t)rank * sel_slot, sel_gather.p, (size_t)sel_slot, ncclDouble, comm, stream),
                        "all-gather of the squared errors");
      if (rc) return rc;
    }
    for (int pass = 0; pass < kSelPasses; pass++) {  // the histogram starts zeroed (begin) and every pass leaves it so
      k_ba_hist_pick<<<std::max(1, std::min((n_keys + 255) / 256, 148 * 8)), 256, 0, stream>>>(d, pass, min_s2);
      launches++;
    }
    return PTAM_OK;
  }

  // One lambda trial of a single-GPU handle, from the upload of its scalars (h_scal[3..6]) to the read-back of the
  // results (h_scal[0..7], h_cnt[5]): Bundle.cc:341-506.  No host decision in between.
  int enqueue_trial_single() {
    int rc;
    const int M = d.n_meas, PO = d.p_hi - d.p_lo;
    PTAM_CUDA_TRY(this, cudaMemcpyAsync(scal.p + 3, h_scal + 3, sizeof(double) * 4, cudaMemcpyHostToDevice, stream));
    pbegin(3);
    if (PO > 0) { k_ba_vinv<<<(PO + 255) / 256, 256, 0, stream>>>(d); launches++; }
    pend(3);
    pbegin(4);
    if ((rc = build_reduced())) return rc;
    pend(4);
    pbegin(6); if ((rc = solve_reduced())) return rc; pend(6);
    pbegin(7);
    k_ba_cam_update<<<1, 256, 0, stream>>>(d); launches++;  // first: it SETS the |delta|^2 slot, the points add to it
    if (PO > 0) { k_ba_point_update<<<(PO + 255) / 256, 256, 0, stream>>>(d); launches++; }
    if (M > 0) { k_ba_new_error<<<(M + 255) / 256, 256, 0, stream>>>(d); launches++; }
    pend(7);
    PTAM_CUDA_TRY(this, cudaGetLastError());
    PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_scal, scal.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, stream));
    if (d.n > 0 && ldlt.dag_err) PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_cnt + 5, ldlt.dag_err, sizeof(int), cudaMemcpyDeviceToHost, stream));
    return PTAM_OK;
  }

  // The same through a CUDA graph: captured on first use (and again when the device graph of the handle changed),
  // replayed afterwards.  Anything that cannot be captured switches the handle back to plain launches.
  int run_trial_single() {
    static const bool env_off = std::getenv("PTAM_B200_NO_GRAPH") != nullptr;
    // (the all-per-panel schedule of the solve, PTAM_B200_LDLT_STEPS=1, is a diagnostic: it is not captured)
    if (env_off || graph_off || profiling || !ldlt.use_dag) return enqueue_trial_single();
    if (!trial_exec || std::memcmp(&trial_key, &d, sizeof(BundleDev)) != 0) {
      if (trial_exec) { cudaGraphExecDestroy(trial_exec); trial_exec = nullptr; }
      if (d.n > 0 && ldlt.prepare(Wp.p, d.n) != cudaSuccess) { graph_off = true; cudaGetLastError(); return enqueue_trial_single(); }
      const int64_t l0 = launches;
      cudaGraph_t g = nullptr;
      std::string keep = err;
      bool ok = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        const int rc = enqueue_trial_single();
        const cudaError_t e = cudaStreamEndCapture(stream, &g);
        ok = rc == PTAM_OK && e == cudaSuccess && g != nullptr;
      }
      if (ok) ok = cudaGraphInstantiate(&trial_exec, g, 0) == cudaSuccess;
      if (g) cudaGraphDestroy(g);
      trial_launches = launches - l0;
      launches = l0;
      if (!ok) {
        cudaGetLastError();
        err = keep;
        if (trial_exec) { cudaGraphExecDestroy(trial_exec); trial_exec = nullptr; }
        graph_off = true;
        return enqueue_trial_single();
      }
      std::memcpy(&trial_key, &d, sizeof(BundleDev));
    }
    PTAM_CUDA_TRY(this, cudaGraphLaunch(trial_exec, stream));
    launches += trial_launches;
    return PTAM_OK;
  }

  // Do_LM_Step (Bundle.cc:209-551)
  int lm_step(const volatile unsigned char* abort_flag) {
    cudaSetDevice(device);
    if (!begun) { set_error("ptam_bundle_begin() has not been called"); return PTAM_ERR_INVALID; }
    // a sharded run must take the same branch on every rank: the local flags are summed with the
    // error scalars and only the reduced value is acted on
    auto local_abort = [&]() { return abort_flag && *abort_flag; };
    auto aborted = [&]() { return world > 1 ? abort_seen : local_abort(); };
    const int C = d.n_cams, M = d.n_meas, n = d.n;
    const int PO = d.p_hi - d.p_lo;  // points owned by this shard
    int rc;
    lm_steps++;
    shards_dirty = true;
    // error / counter scalars (the accumulators are written whole by k_ba_acc_cam / k_ba_acc_pt, not summed into)
    PTAM_CUDA_TRY(this, cudaMemsetAsync(scal.p, 0, sizeof(double) * 8, stream));
    PTAM_CUDA_TRY(this, cudaMemsetAsync(counters.p, 0, sizeof(int), stream));
    pbegin(0); if (M > 0) { k_ba_project<<<(M + 255) / 256, 256, 0, stream>>>(d); launches++; } pend(0);
    pbegin(1); if ((rc = find_sigma_squared())) return rc; pend(1);
    pbegin(2);
    if (M > 0) { k_ba_jacobian<<<(M + 127) / 128, 128, 0, stream>>>(d); launches++; }
    if (C > 0) { k_ba_acc_cam<<<C, kSegThreads, kAccCamSmem, stream>>>(d); launches++; }
    if (PO > 0) { k_ba_acc_pt<<<(PO + 255) / 256, 256, 0, stream>>>(d); launches++; }
    pend(2);
    PTAM_CUDA_TRY(this, cudaGetLastError());
    // sharded: the error sums and abort votes of this LM step (scal[2..5]) ride with the first lambda trial's
    // exchange of S instead of a collective and a host synchronisation of their own
    const bool merged = world > 1 && n > 0;
    // a single-GPU handle needs no round trip of its own for them either: they are read back with the first trial's scalars
    const bool defer = world == 1 || merged;
    if (world > 1) {
      h_scal[5] = local_abort() ? 1.0 : 0.0;
      PTAM_CUDA_TRY(this, cudaMemcpyAsync(scal.p + 5, h_scal + 5, sizeof(double), cudaMemcpyHostToDevice, stream));
      if (!merged && (rc = all_reduce(scal.p + 2, 4, ncclDouble, ncclSum, "all-reduce of the error sums"))) return rc;
    }
    double cur_err = 0.0;
    if (!defer) {
      PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_scal, scal.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, stream));
      PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
      pcollect();
      sigma_sq = h_scal[1];
      cur_err = h_scal[2];
      if (world > 1) abort_seen = h_scal[5] > 0.0;
      last_error = cur_err;
    }
    double new_err = cur_err + 9999;
    bool first = true;
    while ((new_err > cur_err || (defer && first)) && !converged && !hit_max && !aborted()) {
      // scal[3] new error, scal[4] squared update, scal[5] abort votes, scal[6] lambda
      trial_lambda = lambda;
      if (world == 1) {
        h_scal[3] = 0.0; h_scal[4] = 0.0; h_scal[5] = 0.0; h_scal[6] = lambda;
        if ((rc = run_trial_single())) return rc;
        s_mirrored = false;
        PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
        pcollect();
        if (n > 0 && ldlt.dag_err && h_cnt[5]) { set_error("dense solve: a dependency of the persistent factorisation timed out"); return PTAM_ERR_CUDA; }
      } else {
      if (merged && first) {  // scal[3..5] are still on their way: only the lambda goes up now
        h_scal[6] = lambda;
        PTAM_CUDA_TRY(this, cudaMemcpyAsync(scal.p + 6, h_scal + 6, sizeof(double), cudaMemcpyHostToDevice, stream));
      } else {
        h_scal[3] = 0.0; h_scal[4] = 0.0; h_scal[5] = (world > 1 && local_abort()) ? 1.0 : 0.0; h_scal[6] = lambda;
        PTAM_CUDA_TRY(this, cudaMemcpyAsync(scal.p + 3, h_scal + 3, sizeof(double) * 4, cudaMemcpyHostToDevice, stream));
      }
      pbegin(3);
      if (PO > 0) { k_ba_vinv<<<(PO + 255) / 256, 256, 0, stream>>>(d); launches++; }
      pend(3);
      pbegin(4);
      if ((rc = build_reduced())) return rc;
      pend(4);
      s_mirrored = false;
      pbegin(5);
      if (n > 0 && world > 1) {  // the cross-camera J^T J reduction: partial S, vE of every shard -> total on every shard
        if (merged && first) {
          if ((rc = exchange_reduced(scal.p + 2, 4))) return rc;
          // the reduced error sums and votes to the host (read at the trial's synchronisation below), then the
          // trial's own accumulators start from zero
          PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_scal + 8, scal.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, stream));
          PTAM_CUDA_TRY(this, cudaMemsetAsync(scal.p + 3, 0, sizeof(double) * 3, stream));
        } else if ((rc = exchange_reduced())) return rc;
      }
      pend(5);
      pbegin(6); if ((rc = solve_reduced())) return rc; pend(6);
      pbegin(7);
      k_ba_cam_update<<<1, 256, 0, stream>>>(d); launches++;  // first: it SETS the |delta|^2 slot, the points add to it
      if (PO > 0) { k_ba_point_update<<<(PO + 255) / 256, 256, 0, stream>>>(d); launches++; }
      if (M > 0) { k_ba_new_error<<<(M + 255) / 256, 256, 0, stream>>>(d); launches++; }
      pend(7);
      PTAM_CUDA_TRY(this, cudaGetLastError());
      if ((rc = all_reduce(scal.p + 3, 3, ncclDouble, ncclSum, "all-reduce of the trial scalars"))) return rc;
      if (world == 1) PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_scal, scal.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, stream));
      else PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_scal + 3, scal.p + 3, sizeof(double) * 3, cudaMemcpyDeviceToHost, stream));
      if (win) PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_cnt + 4, win->err, sizeof(int), cudaMemcpyDeviceToHost, stream));
      const bool solve_flag = n > 0 && ldlt.dag_err != nullptr;
      if (solve_flag) PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_cnt + 5, ldlt.dag_err, sizeof(int), cudaMemcpyDeviceToHost, stream));
      PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
      pcollect();
      if (win && h_cnt[4]) { set_error("peer exchange of the reduced system timed out (a rank is gone)"); return PTAM_ERR_NCCL; }
      if (solve_flag && h_cnt[5]) { set_error("dense solve: a dependency of the persistent factorisation timed out"); return PTAM_ERR_CUDA; }
      }
      bool step_vote = false;
      if (defer && first) {  // sigma^2 and the error sum of the LM step: what the exchange brought along / this read-back
        const double* hs = merged ? h_scal + 8 : h_scal;
        sigma_sq = hs[1];
        cur_err = hs[2];
        last_error = cur_err;
        step_vote = merged && hs[5] > 0.0;
      }
      first = false;
      new_err = h_scal[3];
      last_new_error = new_err;
      if (world > 1) abort_seen = step_vote || h_scal[5] > 0.0;
      if (h_scal[4] < prm.update_squared_convergence) converged = true;
      if (new_err > cur_err) { lambda = lambda * lambda_factor; lambda_factor = lambda_factor * 2; }  // ModifyLambda_BadStep
      counter++;
      if (counter >= prm.max_iterations) hit_max = true;
    }
    if (defer && first) {  // no trial ran (converged / iteration limit / abort on entry): fetch the sums on their own
      if ((rc = all_reduce(scal.p + 2, 4, ncclDouble, ncclSum, "all-reduce of the error sums"))) return rc;
      PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_scal, scal.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, stream));
      PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
      pcollect();
      sigma_sq = h_scal[1]; cur_err = h_scal[2]; last_error = cur_err;
      if (world > 1) abort_seen = h_scal[5] > 0.0;
      new_err = cur_err + 9999;
    }
    pbegin(9);
    if (new_err < cur_err) {  // ModifyLambda_GoodStep + commit
      lambda_factor = 2.0; lambda *= 0.3;
      PTAM_CUDA_TRY(this, cudaMemcpyAsync(cam_se3.p, cam_se3_new.p, sizeof(double) * 12 * C, cudaMemcpyDeviceToDevice, stream));
      if (PO > 0)
        PTAM_CUDA_TRY(this, cudaMemcpyAsync(pt_pos.p + 3 * (size_t)d.p_lo, pt_pos_new.p + 3 * (size_t)d.p_lo, sizeof(double) * 3 * PO,
                                            cudaMemcpyDeviceToDevice, stream));
      accepted++;
    }
    if (M > 0 && M < 65536) { k_ba_erase<<<1, 1024, 0, stream>>>(d, lm_steps); launches++; }
    else if (M > 0) {
      const int nblk = (M + 1023) / 1024;
      k_ba_erase_count<<<nblk, 1024, 0, stream>>>(d, erase_cnt.p);
      k_ba_erase_scan<<<1, 1024, 0, stream>>>(d, erase_cnt.p, nblk);
      k_ba_erase_write<<<nblk, 1024, 0, stream>>>(d, erase_cnt.p, lm_steps);
      launches += 3;
    }
    pend(9);
    PTAM_CUDA_TRY(this, cudaMemcpyAsync(h_cnt, counters.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, stream));
    cnt_pending = true;  // read at the next synchronisation that needs it (refresh_counts): no round trip of its own
    if (profiling) { PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream)); pcollect(); }
    return PTAM_OK;
  }

  int refresh_counts() {
    if (!cnt_pending) return PTAM_OK;
    PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
    n_outliers = h_cnt[1];
    cnt_pending = false;
    return PTAM_OK;
  }

  // Sharded handles: make every rank hold all adjusted points and the merged outlier list (in the
  // reference's erase order: LM step by LM step, measurement list order inside a step).  Collective.
  int sync_shards() {
    if (world == 1 || !begun || !shards_dirty) return PTAM_OK;
    cudaSetDevice(device);
    const int P = d.n_pts, MG = n_meas(), M = d.n_meas;
    int rc;
    if (P > 0) {  // every shard's own points to every shard: one broadcast per (unequal) range, in one group
      if ((rc = nccl_try(nccl_api().GroupStart(), "ncclGroupStart"))) return rc;
      for (int r = 0; r < world; r++) {
        const size_t lo = (size_t)plan[r], cnt = (size_t)(plan[r + 1] - plan[r]);
        if (cnt == 0) continue;
        if ((rc = nccl_try(nccl_api().Broadcast(pt_pos.p + 3 * lo, pt_pos.p + 3 * lo, 3 * cnt, ncclDouble, r, comm, stream), "all-gather of the points"))) return rc;
      }
      if ((rc = nccl_try(nccl_api().GroupEnd(), "ncclGroupEnd"))) return rc;
    }
    h_outliers.clear();
    if (MG > 0) {
      PTAM_CUDA_TRY(this, cudaMemsetAsync(g_steps.p, 0, sizeof(int) * (size_t)MG, stream));
      if (M > 0) { k_ba_scatter_steps<<<(M + 255) / 256, 256, 0, stream>>>(m_erase_step.p, m_gid.p, g_steps.p, M); launches++; }
      if ((rc = all_reduce(g_steps.p, MG, ncclInt32, ncclMax, "all-reduce of the outlier marks"))) return rc;
      // the marked ones as (index, step) pairs: a few per cent of the list, sorted here by (step, list order)
      PTAM_CUDA_TRY(this, cudaMemsetAsync(g_cnt.p, 0, sizeof(int), stream));
      k_ba_marks_compact<<<(MG + 255) / 256, 256, 0, stream>>>(g_steps.p, MG, g_pairs.p, g_cnt.p);
      launches++;
      int n_marked = 0;
      PTAM_CUDA_TRY(this, cudaMemcpyAsync(&n_marked, g_cnt.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
      PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
      std::vector<int> pairs(2 * (size_t)n_marked);
      if (n_marked) PTAM_CUDA_TRY(this, cudaMemcpy(pairs.data(), g_pairs.p, sizeof(int) * pairs.size(), cudaMemcpyDeviceToHost));
      std::vector<int> ord(n_marked);
      std::iota(ord.begin(), ord.end(), 0);
      std::sort(ord.begin(), ord.end(), [&](int a, int b) {
        return pairs[2 * a + 1] != pairs[2 * b + 1] ? pairs[2 * a + 1] < pairs[2 * b + 1] : pairs[2 * a] < pairs[2 * b]; });
      for (int q : ord) { const int m = pairs[2 * q]; h_outliers.push_back(h_mpt[m]); h_outliers.push_back(h_mcam[m]); }
    } else {
      PTAM_CUDA_TRY(this, cudaStreamSynchronize(stream));
    }
    n_outliers = (int)h_outliers.size() / 2;
    shards_dirty = false;
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
  if (world > 1 && !comm) { b->set_error("world > 1 needs an ncclComm_t (or use ptam_bundle_init_shard)"); return PTAM_ERR_NCCL; }
  if (world > 1 && !nccl_api().load()) { b->set_error(nccl_api().err); return PTAM_ERR_NCCL; }
  if (b->comm && b->own_comm) { release_window((void*)b->comm); nccl_api().CommDestroy(b->comm); }
  b->win = nullptr;
  b->rank = rank; b->world = world; b->comm = world > 1 ? (ncclComm_t)comm : nullptr; b->own_comm = false;
  b->begun = false;
  return PTAM_OK;
}

int ptam_nccl_unique_id(unsigned char id[PTAM_NCCL_UNIQUE_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) == PTAM_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  if (!nccl_api().load()) { ptam_set_global_error(nccl_api().err); return PTAM_ERR_NCCL; }
  ncclUniqueId u;
  const ncclResult_t r = nccl_api().GetUniqueId(&u);
  if (r != ncclSuccess) { ptam_set_global_error(std::string("ncclGetUniqueId: ") + nccl_api().GetErrorString(r)); return PTAM_ERR_NCCL; }
  std::memcpy(id, &u, sizeof(u));
  return PTAM_OK;
}

int ptam_bundle_init_shard(ptam_bundle* b, int rank, int world, const unsigned char id[PTAM_NCCL_UNIQUE_ID_BYTES]) {
  if (world < 1 || rank < 0 || rank >= world) { b->set_error("bad shard description"); return PTAM_ERR_INVALID; }
  if (b->comm && b->own_comm) { release_window((void*)b->comm); nccl_api().CommDestroy(b->comm); b->comm = nullptr; }
  b->win = nullptr;
  b->rank = rank; b->world = world; b->own_comm = false; b->begun = false;
  if (world == 1) return PTAM_OK;
  if (!nccl_api().load()) { b->set_error(nccl_api().err); return PTAM_ERR_NCCL; }
  cudaSetDevice(b->device);
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  int rc = b->nccl_try(nccl_api().CommInitRank(&b->comm, world, u, rank), "ncclCommInitRank");
  if (rc) { b->comm = nullptr; return rc; }
  b->own_comm = true;
  return PTAM_OK;
}

void* ptam_nccl_comm_create(int device, int rank, int world, const unsigned char id[PTAM_NCCL_UNIQUE_ID_BYTES]) {
  if (world < 1 || rank < 0 || rank >= world) { ptam_set_global_error("bad rank / world"); return nullptr; }
  if (!nccl_api().load()) { ptam_set_global_error(nccl_api().err); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { ptam_set_global_error("bad device index"); return nullptr; }
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  const ncclResult_t r = nccl_api().CommInitRank(&comm, world, u, rank);
  if (r != ncclSuccess) { ptam_set_global_error(std::string("ncclCommInitRank: ") + nccl_api().GetErrorString(r)); return nullptr; }
  return comm;
}
void ptam_nccl_comm_destroy(void* comm) {
  if (!comm) return;
  release_window(comm);
  if (nccl_api().load()) nccl_api().CommDestroy((ncclComm_t)comm);
}

int ptam_bundle_shard_plan(int n_points, int n_meas, const int32_t* meas_point, int world, int32_t* point_begin) {
  if (n_points < 0 || n_meas < 0 || world < 1 || !point_begin) return PTAM_ERR_INVALID;
  for (int m = 0; m < n_meas; m++) if (meas_point[m] < 0 || meas_point[m] >= n_points) return PTAM_ERR_INVALID;
  shard_plan(n_points, n_meas, meas_point, world, point_begin);
  return PTAM_OK;
}

int ptam_bundle_begin(ptam_bundle* b) { return b->begin(); }
int ptam_bundle_lm_step(ptam_bundle* b, const volatile unsigned char* abort_flag) {
  const int rc = b->lm_step(abort_flag);
  return rc ? rc : b->refresh_counts();  // the single-step entry returns with the handle's stream drained
}

int ptam_bundle_compute(ptam_bundle* b, const volatile unsigned char* abort_flag) {  // Bundle.cc:116-158
  const auto t0 = std::chrono::steady_clock::now();
  int rc = b->begin();
  if (rc) return rc;
  if (b->profiling) {
    cudaStreamSynchronize(b->stream);  // the pair-list kernels belong to the set-up
    b->prof_begin_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  // a sharded run acts on the reduced abort votes only, so that all ranks leave the loop together
  while (!b->converged && !b->hit_max && !(b->world > 1 ? b->abort_seen : (abort_flag && *abort_flag))) {
    rc = b->lm_step(abort_flag);
    if (rc) return rc;
  }
  if ((rc = b->refresh_counts())) return rc;
  rc = b->sync_shards();
  if (rc) return rc;
  if (b->profiling) b->prof_wall_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return b->accepted;
}

// Persistent graph.  MapMaker::BundleAdjust rebuilds a Bundle from the map for every call
// (MapMaker.cc:852-882) although consecutive calls on the same keyframe set (BundleAdjustAll until
// converged, MapMaker.cc:67-77) feed back exactly what the previous Compute left: the adjusted poses and
// points and the measurement list without the erased outliers.  That state is resident on the device:
// recompute restarts the LM control (Bundle.cc:121-126) on it without the host rebuild and upload.
int ptam_bundle_recompute(ptam_bundle* b, const volatile unsigned char* abort_flag) {
  cudaSetDevice(b->device);
  if (!b->begun) { b->set_error("ptam_bundle_recompute needs a graph that has been computed (or begun) before"); return PTAM_ERR_INVALID; }
  int rc = b->sync_shards();
  if (rc) return rc;
  b->lambda = 0.0001; b->lambda_factor = 2.0;
  b->converged = false; b->hit_max = false; b->abort_seen = false;
  b->counter = 0; b->accepted = 0; b->lm_steps = 0; b->n_outliers = 0;
  b->h_outliers.clear();
  PTAM_CUDA_TRY(b, cudaMemsetAsync(b->counters.p, 0, sizeof(int) * 4, b->stream));
  if (b->d.n_meas > 0) PTAM_CUDA_TRY(b, cudaMemsetAsync(b->m_erase_step.p, 0, sizeof(int) * (size_t)b->d.n_meas, b->stream));
  while (!b->converged && !b->hit_max && !(b->world > 1 ? b->abort_seen : (abort_flag && *abort_flag))) {
    rc = b->lm_step(abort_flag);
    if (rc) return rc;
  }
  if ((rc = b->refresh_counts())) return rc;
  rc = b->sync_shards();
  if (rc) return rc;
  return b->accepted;
}
// Overwrite one camera pose / point of the resident graph (e.g. the tracker's latest estimate) without a rebuild.
int ptam_bundle_update_camera(ptam_bundle* b, int n, const double* se3) {
  cudaSetDevice(b->device);
  if (n < 0 || n >= b->n_cams()) { b->set_error("bad camera index"); return PTAM_ERR_INVALID; }
  std::memcpy(&b->h_cam_se3[12 * (size_t)n], se3, sizeof(double) * 12);
  if (b->begun) {
    PTAM_CUDA_TRY(b, cudaMemcpyAsync(b->cam_se3.p + 12 * (size_t)n, &b->h_cam_se3[12 * (size_t)n], sizeof(double) * 12, cudaMemcpyHostToDevice, b->stream));
    PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  }
  return PTAM_OK;
}
int ptam_bundle_update_point(ptam_bundle* b, int n, const double* xyz) {
  cudaSetDevice(b->device);
  if (n < 0 || n >= b->n_pts()) { b->set_error("bad point index"); return PTAM_ERR_INVALID; }
  double v[3] = {xyz[0], xyz[1], xyz[2]};
  if (std::isnan(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])) v[0] = v[1] = v[2] = 0;
  std::memcpy(&b->h_pts[3 * (size_t)n], v, sizeof v);
  if (b->begun) {
    PTAM_CUDA_TRY(b, cudaMemcpyAsync(b->pt_pos.p + 3 * (size_t)n, &b->h_pts[3 * (size_t)n], sizeof v, cudaMemcpyHostToDevice, b->stream));
    PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  }
  return PTAM_OK;
}

int ptam_bundle_converged(const ptam_bundle* b) { return b->converged ? 1 : 0; }

int ptam_bundle_get_points(ptam_bundle* b, double* xyz) {
  cudaSetDevice(b->device);
  { const int rc = b->sync_shards(); if (rc) return rc; }
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
  { const int rc = b->sync_shards(); if (rc) return rc; }
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
  if (b->world > 1) {
    const int rc = b->sync_shards();
    if (rc) return rc;
    const int n = b->n_outliers;
    if (pairs && cap > 0 && n > 0) std::memcpy(pairs, b->h_outliers.data(), sizeof(int) * 2 * std::min(n, cap));
    return n;
  }
  PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  const int n = b->n_outliers;
  if (pairs && cap > 0 && n > 0)
    PTAM_CUDA_TRY(b, cudaMemcpy(pairs, b->outliers.p, sizeof(int) * 2 * std::min(n, cap), cudaMemcpyDeviceToHost));
  return n;
}
int ptam_bundle_get_stats(ptam_bundle* b, ptam_bundle_stats* s) {
  { const int rc = b->sync_shards(); if (rc) return rc; }
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
  PTAM_CUDA_TRY(b, cudaMemcpyAsync(b->scal.p + 6, &l, sizeof(double), cudaMemcpyHostToDevice, b->stream));
  PTAM_CUDA_TRY(b, cudaStreamSynchronize(b->stream));
  const int PO = b->d.p_hi - b->d.p_lo;
  if (PO > 0) k_ba_vinv<<<(PO + 255) / 256, 256, 0, b->stream>>>(b->d);
  { const int rc = b->build_reduced(); if (rc) return rc; }
  if (b->world > 1) { const int rc = b->exchange_reduced(); if (rc) return rc; }
  k_ba_mirror<<<std::min<size_t>(((size_t)n * n + 255) / 256, 148 * 8), 256, 0, b->stream>>>(b->d.S, n);
  b->launches += 2;
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
int ptam_bundle_set_profiling(ptam_bundle* b, int on) {
  cudaSetDevice(b->device);
  if (on && !b->prof_ev[0])
    for (auto& e : b->prof_ev) PTAM_CUDA_TRY(b, cudaEventCreate(&e));
  b->profiling = on != 0;
  b->prof_pending = 0;
  for (int k = 0; k < PTAM_BA_PHASES; k++) { b->prof_ms[k] = 0; b->prof_n[k] = 0; }
  b->prof_begin_ms = 0; b->prof_wall_ms = 0;
  return PTAM_OK;
}
int ptam_bundle_get_phase_times(ptam_bundle* b, double* ms_total, int64_t* count) {
  double dev = 0;
  for (int k = 0; k < PTAM_BA_PHASES; k++) { ms_total[k] = b->prof_ms[k]; count[k] = b->prof_n[k]; if (k < 10) dev += b->prof_ms[k]; }
  ms_total[10] = b->prof_begin_ms; count[10] = b->prof_begin_ms > 0 ? 1 : 0;
  // what is left of the wall clock of Compute: host LM control, the blocking read-backs, launch gaps
  ms_total[11] = b->prof_wall_ms > 0 ? b->prof_wall_ms - dev - b->prof_begin_ms : 0; count[11] = b->prof_wall_ms > 0 ? 1 : 0;
  return PTAM_OK;
}
#ifdef PTAM_PANEL_DEBUG
int ptam_debug_read(long long* out) { cudaDeviceSynchronize(); return (int)cudaMemcpyFromSymbol(out, g_dbg, sizeof(long long) * 8); }
#endif
void* ptam_bundle_cuda_stream(ptam_bundle* b) { return (void*)b->stream; }
int64_t ptam_bundle_launch_count(const ptam_bundle* b) { return b->launches; }

int ptam_bundle_solve_schedule(int n, int tail_tiles, int* k_start, int32_t* task_off, int cap) {
  if (n < 0 || !k_start || !task_off) return PTAM_ERR_INVALID;
  std::vector<int> off;
  const int ks = ldlt_dag_schedule(n, tail_tiles, off);
  if ((int)off.size() > cap) return PTAM_ERR_INVALID;
  *k_start = ks;
  for (size_t i = 0; i < off.size(); i++) task_off[i] = off[i];
  return (int)off.size() - 1;
}

}  // extern "C"
