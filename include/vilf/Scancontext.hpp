// vilf/Scancontext.hpp — host-side mirror of the reference's SCManager (src/global_fusion/include/Scancontext/Scancontext.h)
// over the C ABI (vilf_sc_*): same method names, argument meaning and return values as the reference's user-side API
// (makeAndSaveScancontextAndKeys SC:196-208, detectLoopClosureID SC:210-299, setSCdistThres SC:300, setMaximumRadius SC:305)
// and its public parameter members.  Call sites in the reference: poseGraphOptimization.cpp:553 (every key frame) and
// :600-603 (performSCLoopClosure); setters :642-643.
//
// Differences a maintainer should know about:
//  * descriptors live on the device; getScancontext(i) / distanceBtnScanContext(i, j) read them back on demand.  The Eigen
//    members polarcontexts_ / polarcontext_invkeys_ / polarcontext_vkeys_ are not mirrored as containers.
//  * the setters must run before the first key frame (they do in the reference: poseGraphOptimization.cpp main()), because the
//    device object is created lazily from the parameter members on first use.
//  * makeAndSaveScancontextAndKeys(session) takes the cloud the odometry would publish as /GlobalMap straight from device
//    memory (vilf_sc_make_and_save_resident) when the loop detector runs in the odometry's process.
//  * errors THROW std::runtime_error (no CPU fallback); the "[Loop found] / [Not loop]" line is printed like the reference's.
#pragma once

#include <iostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "cloud.hpp"
#include "session.hpp"

namespace vilf {

class SCManager {
 public:
  SCManager() : sc_(nullptr), device_(0), quiet_(false) {}
  ~SCManager() {
    if (sc_) vilf_sc_destroy(sc_);
  }
  SCManager(const SCManager&) = delete;
  SCManager& operator=(const SCManager&) = delete;

  // ---- user-side API (SC:196-309) ----
  template <class CloudT>
  void makeAndSaveScancontextAndKeys(CloudT& _scan_down) {
    std::vector<float> packed;
    pack_cloud(_scan_down, packed);
    float dummy[4] = {0, 0, 0, 0};
    check(vilf_sc_make_and_save(handle(), packed.empty() ? dummy : packed.data(), (int)_scan_down.points.size()), "makeAndSaveScancontextAndKeys");
  }
  void makeAndSaveScancontextAndKeys(Session& odometry) {
    check(vilf_sc_make_and_save_resident(handle(), odometry.handle()), "makeAndSaveScancontextAndKeys(resident)");
  }
  std::pair<int, float> detectLoopClosureID(void) {  // int: nearest node index, float: relative yaw
    int loop_id = -1, nn_idx = 0, n = 0;
    float yaw = 0.0f;
    double min_dist = 0;
    check(vilf_sc_size(handle(), &n), "detectLoopClosureID");
    check(vilf_sc_detect_loop_closure(handle(), &loop_id, &yaw, &min_dist, &nn_idx), "detectLoopClosureID");
    if (!quiet_ && n >= NUM_EXCLUDE_RECENT + 1) {  // SC:279-289 (the early return of SC:220-224 prints nothing)
      std::cout.precision(3);
      std::cout << (loop_id >= 0 ? "[Loop found]" : "[Not loop]") << " Nearest distance: " << min_dist << " btn " << n - 1 << " and " << nn_idx << "." << std::endl;
    }
    return std::pair<int, float>(loop_id, yaw);
  }
  void setSCdistThres(double _new_thres) { require_fresh("setSCdistThres"); SC_DIST_THRES = _new_thres; }
  void setMaximumRadius(double _max_r) { require_fresh("setMaximumRadius"); PC_MAX_RADIUS = _max_r; }

  // ---- the pieces (SC:42-193), row-major num_ring x num_sector ----
  std::vector<double> getScancontext(int index) {
    std::vector<double> d((std::size_t)PC_NUM_RING * PC_NUM_SECTOR);
    check(vilf_sc_get(handle(), index, d.data(), nullptr, nullptr), "getScancontext");
    return d;
  }
  std::vector<double> getRingkey(int index) {
    std::vector<double> k((std::size_t)PC_NUM_RING);
    check(vilf_sc_get(handle(), index, nullptr, k.data(), nullptr), "getRingkey");
    return k;
  }
  std::vector<double> getSectorkey(int index) {
    std::vector<double> k((std::size_t)PC_NUM_SECTOR);
    check(vilf_sc_get(handle(), index, nullptr, nullptr, k.data()), "getSectorkey");
    return k;
  }
  std::pair<double, int> distanceBtnScanContext(const std::vector<double>& _sc1, const std::vector<double>& _sc2) {
    if (_sc1.size() != (std::size_t)PC_NUM_RING * PC_NUM_SECTOR || _sc2.size() != _sc1.size()) throw std::runtime_error("distanceBtnScanContext: descriptor size");
    double d = 0;
    int s = 0;
    check(vilf_sc_distance(handle(), _sc1.data(), _sc2.data(), &d, &s), "distanceBtnScanContext");
    return std::make_pair(d, s);
  }
  std::pair<double, int> distanceBtnScanContext(int i, int j) {
    double d = 0;
    int s = 0;
    check(vilf_sc_distance_between(handle(), i, j, &d, &s), "distanceBtnScanContext");
    return std::make_pair(d, s);
  }
  int size() {
    int n = 0;
    check(vilf_sc_size(handle(), &n), "size");
    return n;
  }

  void set_device(int d) { require_fresh("set_device"); device_ = d; }
  void set_quiet(bool q) { quiet_ = q; }
  void set_capacity(int max_keyframes, int max_points) { require_fresh("set_capacity"); max_keyframes_ = max_keyframes; max_points_ = max_points; }

 public:
  // hyper parameters (SC:313-332), same names and defaults
  double LIDAR_HEIGHT = 2.0;
  int PC_NUM_RING = 20;
  int PC_NUM_SECTOR = 60;
  double PC_MAX_RADIUS = 80.0;
  int NUM_EXCLUDE_RECENT = 30;
  int NUM_CANDIDATES_FROM_TREE = 3;
  double SEARCH_RATIO = 0.1;
  double SC_DIST_THRES = 0.2;
  int TREE_MAKING_PERIOD_ = 30;

 private:
  vilf_sc* handle() {
    if (!sc_) {
      vilf_sc_params p;
      vilf_sc_default_params(&p);
      p.lidar_height = LIDAR_HEIGHT; p.num_ring = PC_NUM_RING; p.num_sector = PC_NUM_SECTOR; p.max_radius = PC_MAX_RADIUS;
      p.num_exclude_recent = NUM_EXCLUDE_RECENT; p.num_candidates = NUM_CANDIDATES_FROM_TREE; p.search_ratio = SEARCH_RATIO;
      p.dist_thres = SC_DIST_THRES; p.tree_making_period = TREE_MAKING_PERIOD_;
      if (max_keyframes_ > 0) p.max_keyframes = max_keyframes_;
      if (max_points_ > 0) p.max_points = max_points_;
      const int rc = vilf_sc_create(&p, device_, &sc_);
      if (rc != VILF_OK) {
        sc_ = nullptr;
        throw std::runtime_error("vilf_sc_create failed with status " + std::to_string(rc) + " (libvilf_cuda.so needs a CUDA device; there is no CPU fallback)");
      }
    }
    return sc_;
  }
  void require_fresh(const char* what) {
    if (sc_) throw std::runtime_error(std::string(what) + ": parameters are fixed once the first key frame has been stored");
  }
  void check(int rc, const char* what) {
    if (rc != VILF_OK) throw std::runtime_error(std::string(what) + ": status " + std::to_string(rc) + ": " + (sc_ ? vilf_sc_last_error(sc_) : "no handle"));
  }
  vilf_sc* sc_;
  int device_;
  bool quiet_;
  int max_keyframes_ = 0, max_points_ = 0;
};

}  // namespace vilf
