// Spatial partition of the GPU's SMs between the kernels of the encode pipeline (CUDA green contexts).
//
// k_model owns one CTA per SM (the context-state table fills the shared memory), k_range is a handful of latency-bound
// single-warp CTAs, k_emit / k_pack are throughput kernels. Left to the hardware scheduler they get in each other's way
// (DESIGN.md §4): k_range waits behind queued k_model CTAs, k_emit lands on the SMs k_range runs on. A green context pins
// the streams created from it to a fixed set of SMs, so every kernel class gets SMs of its own and the three run side by side.
//
// The driver entry points are looked up through cudaGetDriverEntryPoint: the library does not link libcuda.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace b200 {

class SmPartition {
  public:
    SmPartition() = default;
    ~SmPartition();
    SmPartition(const SmPartition&) = delete;
    SmPartition& operator=(const SmPartition&) = delete;

    // Splits the SMs of `device` into parts of (at least) want[i] SMs, i = 0 .. want.size()-2; the last part takes every SM that
    // is left (want.back() is ignored). Counts are rounded up to the device's granularity (8 SMs on sm_90+). Returns false
    // (and leaves the object unusable, why() says why) when green contexts are not available or the split does not fit.
    bool create(int device, const std::vector<int>& want);
    bool ok() const { return !parts_.empty(); }
    int parts() const { return (int)parts_.size(); }
    int sm_count(int part) const { return parts_[part].sms; }
    // a new non-blocking stream whose kernels run on the SMs of `part` only
    cudaStream_t stream(int part, int priority = 0);
    const std::string& why() const { return why_; }

  private:
    struct Part { void* gctx = nullptr; int sms = 0; };
    std::vector<Part> parts_;
    std::vector<cudaStream_t> streams_;
    std::string why_;
};

}  // namespace b200
