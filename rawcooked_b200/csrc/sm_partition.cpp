// Green-context SM partition (see sm_partition.h).
#include "sm_partition.h"

#include <cuda.h>

#include <cstdio>

namespace b200 {

namespace {
struct Driver {
    decltype(&cuDeviceGet) DeviceGet = nullptr;
    decltype(&cuDeviceGetDevResource) DeviceGetDevResource = nullptr;
    decltype(&cuDevSmResourceSplitByCount) DevSmResourceSplitByCount = nullptr;
    decltype(&cuDevResourceGenerateDesc) DevResourceGenerateDesc = nullptr;
    decltype(&cuGreenCtxCreate) GreenCtxCreate = nullptr;
    decltype(&cuGreenCtxStreamCreate) GreenCtxStreamCreate = nullptr;
    decltype(&cuGreenCtxDestroy) GreenCtxDestroy = nullptr;
    bool ok = false;
    Driver() {
        ok = get("cuDeviceGet", &DeviceGet) && get("cuDeviceGetDevResource", &DeviceGetDevResource) &&
             get("cuDevSmResourceSplitByCount", &DevSmResourceSplitByCount) && get("cuDevResourceGenerateDesc", &DevResourceGenerateDesc) &&
             get("cuGreenCtxCreate", &GreenCtxCreate) && get("cuGreenCtxStreamCreate", &GreenCtxStreamCreate) &&
             get("cuGreenCtxDestroy", &GreenCtxDestroy);
    }
    template <class F>
    static bool get(const char* name, F* out) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &qr) != cudaSuccess || !f || qr != cudaDriverEntryPointSuccess) return false;
        *out = reinterpret_cast<F>(f);
        return true;
    }
};
const Driver& driver() { static const Driver d; return d; }
}  // namespace

bool SmPartition::create(int device, const std::vector<int>& want) {
    parts_.clear();
    const Driver& D = driver();
    if (!D.ok) { why_ = "green-context entry points not found in this driver"; return false; }
    if (want.size() < 2) { why_ = "need at least two parts"; return false; }
    CUdevice dev;
    if (D.DeviceGet(&dev, device) != CUDA_SUCCESS) { why_ = "cuDeviceGet failed"; return false; }
    CUdevResource all;
    if (D.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) { why_ = "cuDeviceGetDevResource failed"; return false; }
    // the device's granularity: groups of the smallest size it hands out
    unsigned ngrp = 0;
    if (D.DevSmResourceSplitByCount(nullptr, &ngrp, &all, nullptr, 0, 1) != CUDA_SUCCESS || ngrp == 0) { why_ = "cuDevSmResourceSplitByCount (query) failed"; return false; }
    std::vector<CUdevResource> grp(ngrp);
    CUdevResource rem;
    rem.type = CU_DEV_RESOURCE_TYPE_INVALID;
    if (D.DevSmResourceSplitByCount(grp.data(), &ngrp, &all, &rem, 0, 1) != CUDA_SUCCESS || ngrp == 0) { why_ = "cuDevSmResourceSplitByCount failed"; return false; }
    grp.resize(ngrp);
    const unsigned gsz = grp[0].sm.smCount;
    size_t next = 0;
    std::vector<std::vector<CUdevResource>> sets;
    for (size_t i = 0; i + 1 < want.size(); i++) {
        const unsigned need = ((unsigned)(want[i] > 0 ? want[i] : 1) + gsz - 1) / gsz;
        if (next + need >= grp.size()) { why_ = "the parts asked for do not leave SMs for the last one"; return false; }
        sets.emplace_back(grp.begin() + next, grp.begin() + next + need);
        next += need;
    }
    std::vector<CUdevResource> last(grp.begin() + next, grp.end());
    if (rem.type == CU_DEV_RESOURCE_TYPE_SM && rem.sm.smCount > 0) last.push_back(rem);
    sets.push_back(last);
    for (auto& set : sets) {
        CUdevResourceDesc desc;
        if (D.DevResourceGenerateDesc(&desc, set.data(), (unsigned)set.size()) != CUDA_SUCCESS) { why_ = "cuDevResourceGenerateDesc failed"; parts_.clear(); return false; }
        CUgreenCtx g;
        if (D.GreenCtxCreate(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) { why_ = "cuGreenCtxCreate failed"; parts_.clear(); return false; }
        Part p;
        p.gctx = g;
        for (auto& r : set) p.sms += (int)r.sm.smCount;
        parts_.push_back(p);
    }
    return true;
}

cudaStream_t SmPartition::stream(int part, int priority) {
    if (part < 0 || part >= (int)parts_.size()) return nullptr;
    CUstream s = nullptr;
    if (driver().GreenCtxStreamCreate(&s, (CUgreenCtx)parts_[part].gctx, CU_STREAM_NON_BLOCKING, priority) != CUDA_SUCCESS) return nullptr;
    streams_.push_back((cudaStream_t)s);
    return (cudaStream_t)s;
}

SmPartition::~SmPartition() {
    for (cudaStream_t s : streams_) cudaStreamDestroy(s);
    for (auto& p : parts_) if (p.gctx) driver().GreenCtxDestroy((CUgreenCtx)p.gctx);
}

}  // namespace b200
