// Exercises the MapMaker host mirror (host/MapMaker.h: BundleAdjustAll / BundleAdjustRecent / BundleAdjust over
// the Bundle mirror) on a map read from raw arrays, and writes the map back: adjusted points and keyframe
// poses, bad flags, surviving measurement counts, the failure queue and the never-retry sets.
// Driven by tests/test_zz_host_mapmaker_gpu.py (CUDA library) and tests/test_host_mapmaker_cpu.py (the same
// source compiled against the CPU oracle's identical ABI, to check the marshalling without a GPU).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include "MapMaker.h"

using namespace ptam_b200;
using namespace TooN;

template <class T>
static std::vector<T> rd(const std::string& dir, const char* name) {
  std::ifstream f(dir + "/" + name, std::ios::binary | std::ios::ate);
  if (!f) { std::cerr << "missing " << name << "\n"; std::exit(2); }
  const size_t bytes = (size_t)f.tellg();
  f.seekg(0);
  std::vector<T> v(bytes / sizeof(T));
  f.read(reinterpret_cast<char*>(v.data()), bytes);
  return v;
}
template <class T>
static void wr(const std::string& dir, const char* name, const std::vector<T>& v) {
  std::ofstream f(dir + "/" + name, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(T));
}

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: mapmaker_check <dir>\n"; return 2; }
  const std::string dir = argv[1];
  try {
    auto cams = rd<double>(dir, "mm_cams.f64");
    auto fixed = rd<int32_t>(dir, "mm_fixed.i32");
    auto pts = rd<double>(dir, "mm_pts.f64");
    auto mcam = rd<int32_t>(dir, "mm_mcam.i32");
    auto mpt = rd<int32_t>(dir, "mm_mpt.i32");
    auto uv = rd<double>(dir, "mm_uv.f64");
    auto lvl = rd<int32_t>(dir, "mm_level.i32");
    auto src = rd<int32_t>(dir, "mm_src.i32");
    auto mode = rd<int32_t>(dir, "mm_mode.i32");  // [0]: 0 = BundleAdjustAll, 1 = BundleAdjustRecent; [1]: max iterations
    const size_t C = fixed.size(), P = pts.size() / 3;
    // contiguous storage: pointer order = index order, which is the order std::set / std::map walk them in
    std::vector<KeyFrame> kfs(C);
    std::vector<MapPoint> points(P);
    Map map;
    for (size_t c = 0; c < C; c++) {
      kfs[c].se3CfromW = se3_from_array(&cams[12 * c]);
      kfs[c].bFixed = fixed[c] != 0;
      map.vpKeyFrames.push_back(&kfs[c]);
    }
    for (size_t p = 0; p < P; p++) {
      points[p].v3WorldPos = makeVector(pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]);
      map.vpPoints.push_back(&points[p]);
    }
    map.bGood = true;
    ATANCamera cam("Camera");
    ptam_bundle_params prm;
    ptam_bundle_default_params(&prm);
    prm.max_iterations = mode[1];
    MapMaker mm(map, cam, 0, &prm);
    for (size_t m = 0; m < mcam.size(); m++) {
      Measurement me;
      me.nLevel = lvl[m];
      me.bSubPix = false;
      me.v2RootPos = makeVector(uv[2 * m], uv[2 * m + 1]);
      me.Source = static_cast<decltype(me.Source)>(src[m]);
      kfs[mcam[m]].mMeasurements[&points[mpt[m]]] = me;
      mm.MMData(&points[mpt[m]]).sMeasurementKFs.insert(&kfs[mcam[m]]);
    }
    if (mode[0] == 0) mm.BundleAdjustAll(); else mm.BundleAdjustRecent();
    std::vector<double> opts, ocams(12 * C);
    std::vector<int32_t> bad, nmeas, queue, never;
    for (size_t p = 0; p < P; p++) {
      for (int k = 0; k < 3; k++) opts.push_back(points[p].v3WorldPos[k]);
      bad.push_back(points[p].bBad ? 1 : 0);
      for (KeyFrame* kf : mm.MMData(&points[p]).sNeverRetryKFs) { never.push_back((int32_t)(kf - kfs.data())); never.push_back((int32_t)p); }
    }
    for (size_t c = 0; c < C; c++) { se3_to_array(kfs[c].se3CfromW, &ocams[12 * c]); nmeas.push_back((int32_t)kfs[c].mMeasurements.size()); }
    for (auto& q : mm.mvFailureQueue) { queue.push_back((int32_t)(q.first - kfs.data())); queue.push_back((int32_t)(q.second - points.data())); }
    std::vector<int32_t> flags = {mm.mbBundleConverged_Full ? 1 : 0, mm.mbBundleConverged_Recent ? 1 : 0, mm.mbResetRequested ? 1 : 0,
                                  mm.mbBundleRunning ? 1 : 0};
    wr(dir, "mm_out_pts.f64", opts); wr(dir, "mm_out_cams.f64", ocams); wr(dir, "mm_out_bad.i32", bad); wr(dir, "mm_out_nmeas.i32", nmeas);
    wr(dir, "mm_out_queue.i32", queue); wr(dir, "mm_out_never.i32", never); wr(dir, "mm_out_flags.i32", flags);
    std::printf("mapmaker: mode %d, %zu keyframes, %zu points, failure queue %zu\n", mode[0], C, P, mm.mvFailureQueue.size());
  } catch (const std::exception& e) {
    std::cerr << "mapmaker_check failed: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
