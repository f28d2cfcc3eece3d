// vilf/session.hpp — one device-resident lidar sequence (a vilf_handle) shared by the two host-side classes.
//
// In the reference, featureExtraction and EstimationMapping are independent objects that exchange host point clouds
// (feature_tracker_node.cpp:346, :373, :384).  Here both hold a shared_ptr<Session>; when they share ONE session
// (EstimationMapping::shareSession) the extracted features stay on the device between extractFeature and
// optimation_processing, and only the pose (7 doubles) comes back per frame.  Without sharing everything still runs on
// the GPU — the features just take one extra host round trip, exactly as the reference's signatures imply.
//
// The handle is created lazily (first use), after both initParam and initParameter had the chance to fill the config.
// Errors: the reference's methods are void and print; a CUDA / capacity / state error here THROWS std::runtime_error
// with vilf_last_error() — there is no CPU fallback to degrade to.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#include "../vilf.h"

namespace vilf {

class Session {
 public:
  Session() : h_(nullptr), device_(0), resident_tag_(0), next_tag_(1) { vilf_default_config(&cfg_); }
  ~Session() {
    if (h_) vilf_destroy(h_);
  }
  Session(const Session&) = delete;
  Session& operator=(const Session&) = delete;

  vilf_config& config() {  // editable until the handle exists
    return cfg_;
  }
  void set_device(int d) { device_ = d; }
  bool created() const { return h_ != nullptr; }

  vilf_handle* handle() {
    if (!h_) {
      const int rc = vilf_create(&cfg_, device_, &h_);
      if (rc != VILF_OK) {
        h_ = nullptr;
        throw std::runtime_error("vilf_create failed with status " + std::to_string(rc) +
                                 " (libvilf_cuda.so needs a CUDA device; there is no CPU fallback)");
      }
    }
    return h_;
  }
  void check(int rc, const char* what) {
    if (rc != VILF_OK) throw std::runtime_error(std::string(what) + ": status " + std::to_string(rc) + ": " + (h_ ? vilf_last_error(h_) : "no handle"));
  }

  // Which host clouds correspond to the features currently resident on the device: extractFeature records the sizes
  // and a hash of EVERY point it appended; optimation_processing / localMapInited use the resident copy only when the
  // clouds they are given still hash to the same value (a caller who edits any point in between gets an upload).
  struct Resident {
    std::uint64_t tag;
    std::size_t n_edge, n_surf;
    std::uint64_t hash_edge, hash_surf;
  };
  // 64-bit FNV-1a over the x, y, z, intensity bit patterns of a cloud (works for pcl::PointXYZI's padded layout too).
  template <class CloudT>
  static std::uint64_t content_hash(const CloudT& c) {
    std::uint64_t h = 1469598103934665603ull;
    for (std::size_t i = 0; i < c.points.size(); ++i) {
      const float v[4] = {c.points[i].x, c.points[i].y, c.points[i].z, c.points[i].intensity};
      std::uint32_t w[4];
      std::memcpy(w, v, 16);
      for (int k = 0; k < 4; ++k) { h ^= w[k]; h *= 1099511628211ull; }
    }
    return h;
  }
  Resident resident = Resident();
  std::uint64_t new_tag() { return resident_tag_ = next_tag_++; }
  std::uint64_t resident_tag() const { return resident_tag_; }
  void invalidate_resident() { resident_tag_ = 0; }

 private:
  vilf_config cfg_;
  vilf_handle* h_;
  int device_;
  std::uint64_t resident_tag_, next_tag_;
};

typedef std::shared_ptr<Session> SessionPtr;

}  // namespace vilf
