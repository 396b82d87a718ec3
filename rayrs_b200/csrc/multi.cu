// Multi-GPU render call: the samples of every pixel are split over the GPUs of one box, each GPU accumulates its
// slice into its own fp32 radiance buffer, ONE ncclReduce(sum) over NVLink / NVSwitch merges the buffers on global
// rank 0, which divides by spp.  The reference's seam stays one call (rayrs/src/main.rs:57-101).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a single-GPU host needs no NCCL, and a process that already
// holds an NCCL (torch's) shares it instead of loading a second copy.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <nccl.h>  // types and prototypes only; no NCCL symbol is linked

#include "wavefront.cuh"

using namespace rrs;


namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            const char* why = dlerror();  // one call: dlerror() clears the message it returns
            api.error = std::string("cannot load libnccl.so.2: ") + (why ? why : "not found");
            return;
        }
        bool ok = true;
        auto sym = [&](const char* n) {
            void* p = dlsym(api.lib, n);
            if (!p) {
                ok = false;
                api.error = std::string("libnccl.so.2 lacks ") + n;
            }
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        if (!ok) {
            dlclose(api.lib);
            api.lib = nullptr;
        }
    });
    return api;
}

}  // namespace

// api.cu owns the error channel of the library
extern "C" const char* rrs_last_error(void);
namespace rrs {
int api_fail(int code, const std::string& msg);  // api.cu
}

struct RrsComm {
    int world = 0;
    std::vector<int> devices;       // local devices
    std::vector<int> ranks;         // global rank of each local device
    std::vector<ncclComm_t> comms;  // one per local device
};

#define RRS_NCCL_CHECK(expr)                                                                              \
    do {                                                                                                  \
        ncclResult_t _r = (expr);                                                                         \
        if (_r != ncclSuccess) return api_fail(RRS_ERR_COMM, std::string(#expr) + ": " + nccl().GetErrorString(_r)); \
    } while (0)

extern "C" {

int rrs_sample_range(uint32_t rank, uint32_t world, uint32_t spp, uint32_t* first, uint32_t* count) {
    if (world == 0 || rank >= world || !first || !count) return api_fail(RRS_ERR_INVALID, "bad rank / world");
    const uint32_t base = spp / world, extra = spp % world;
    *count = base + (rank < extra ? 1u : 0u);
    *first = rank * base + std::min(rank, extra);
    return RRS_OK;
}

int rrs_comm_init_all(const int* devices, int n, RrsComm** out) {
    if (!devices || n < 1 || !out) return api_fail(RRS_ERR_INVALID, "null argument");
    *out = nullptr;
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < i; ++k)
            if (devices[i] == devices[k]) return api_fail(RRS_ERR_INVALID, "a device is listed twice");
    auto* c = new RrsComm();
    c->world = n;
    c->devices.assign(devices, devices + n);
    for (int i = 0; i < n; ++i) c->ranks.push_back(i);
    if (n > 1) {
        if (!nccl().lib) {
            delete c;
            return api_fail(RRS_ERR_COMM, nccl().error);
        }
        c->comms.resize(n);
        ncclResult_t r = nccl().CommInitAll(c->comms.data(), n, devices);
        if (r != ncclSuccess) {
            delete c;
            return api_fail(RRS_ERR_COMM, std::string("ncclCommInitAll: ") + nccl().GetErrorString(r));
        }
    }
    *out = c;
    return RRS_OK;
}

int rrs_comm_unique_id(RrsUniqueId* out) {
    if (!out) return api_fail(RRS_ERR_INVALID, "null argument");
    static_assert(sizeof(RrsUniqueId) == sizeof(ncclUniqueId), "RrsUniqueId must mirror ncclUniqueId");
    if (!nccl().lib) return api_fail(RRS_ERR_COMM, nccl().error);
    ncclUniqueId id;
    RRS_NCCL_CHECK(nccl().GetUniqueId(&id));
    std::memcpy(out->bytes, &id, sizeof(id));
    return RRS_OK;
}

int rrs_comm_init_rank(const RrsUniqueId* id, int world, int rank, int device, RrsComm** out) {
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return api_fail(RRS_ERR_INVALID, "bad argument");
    *out = nullptr;
    if (!nccl().lib) return api_fail(RRS_ERR_COMM, nccl().error);
    if (cudaSetDevice(device) != cudaSuccess) return api_fail(RRS_ERR_CUDA, "cudaSetDevice failed");
    ncclUniqueId nid;
    std::memcpy(&nid, id->bytes, sizeof(nid));
    ncclComm_t comm;
    RRS_NCCL_CHECK(nccl().CommInitRank(&comm, world, nid, rank));
    auto* c = new RrsComm();
    c->world = world;
    c->devices = {device};
    c->ranks = {rank};
    c->comms = {comm};
    *out = c;
    return RRS_OK;
}

void rrs_comm_destroy(RrsComm* comm) {
    if (!comm) return;
    for (size_t i = 0; i < comm->comms.size(); ++i) {
        cudaSetDevice(comm->devices[i]);
        nccl().CommDestroy(comm->comms[i]);
    }
    delete comm;
}

int rrs_render_multi(RrsScene* const* scenes, int n_local, RrsComm* comm, const RrsCamera* camera,
                     const RrsRenderParams* params, float* out_rgb, int out_is_device, void* const* cuda_streams) {
    if (!scenes || !comm || !camera || !params || n_local < 1) return api_fail(RRS_ERR_INVALID, "null argument");
    if (n_local != (int)comm->devices.size()) return api_fail(RRS_ERR_INVALID, "n_local does not match the communicator");
    int root_local = -1;
    for (int i = 0; i < n_local; ++i) {
        if (!scenes[i]) return api_fail(RRS_ERR_INVALID, "null scene");
        if (scenes[i]->impl.device != comm->devices[i]) return api_fail(RRS_ERR_INVALID, "scene i must live on device i of the communicator");
        if (comm->ranks[i] == 0) root_local = i;
    }
    if (root_local >= 0 && !out_rgb) return api_fail(RRS_ERR_INVALID, "rank 0 needs an output buffer");
    const size_t npix = (size_t)params->width * params->height;
    if (npix == 0) return api_fail(RRS_ERR_INVALID, "empty image");
    const uint32_t spp_total = params->spp_total ? params->spp_total : params->spp;
    std::string err;
    std::vector<cudaStream_t> streams(n_local);
    // ---- every local GPU renders its sample range, asynchronously ----
    for (int i = 0; i < n_local; ++i) {
        SceneImpl& s = scenes[i]->impl;
        if (cudaSetDevice(s.device) != cudaSuccess) return api_fail(RRS_ERR_CUDA, "cudaSetDevice failed");
        if (cuda_streams) {
            streams[i] = static_cast<cudaStream_t>(cuda_streams[i]);
        } else {
            if (!s.own_stream && cudaStreamCreateWithFlags(&s.own_stream, cudaStreamNonBlocking) != cudaSuccess)
                return api_fail(RRS_ERR_CUDA, "cudaStreamCreate failed");
            streams[i] = s.own_stream;
        }
        if (s.accum_pixels != npix) {
            cudaFree(s.accum);
            s.accum = nullptr;
            s.accum_pixels = 0;
            if (cudaMalloc(&s.accum, sizeof(float4) * npix) != cudaSuccess) return api_fail(RRS_ERR_NOMEM, "cudaMalloc(accumulator) failed");
            s.accum_pixels = npix;
        }
        if (cudaMemsetAsync(s.accum, 0, sizeof(float4) * npix, streams[i]) != cudaSuccess) return api_fail(RRS_ERR_CUDA, "cudaMemsetAsync failed");
        RrsRenderParams p = *params;
        uint32_t first = 0, count = 0;
        rrs_sample_range((uint32_t)comm->ranks[i], (uint32_t)comm->world, params->spp, &first, &count);
        p.spp = count;
        p.sample_offset = params->sample_offset + first;
        p.spp_total = spp_total;
        if (count > 0) {
            int rc = wf_render_accumulate(&s, camera, &p, s.accum, streams[i], err);
            if (rc != RRS_OK) return api_fail(rc, err);
        } else {
            wf_finish_stats(&s);
            s.stats = RrsStats{};  // more GPUs than samples: this one only joins the reduce
        }
    }
    // ---- the path's single exchange step ----
    if (comm->world > 1) {
        RRS_NCCL_CHECK(nccl().GroupStart());
        for (int i = 0; i < n_local; ++i) {
            SceneImpl& s = scenes[i]->impl;
            ncclResult_t r = nccl().Reduce(s.accum, s.accum, npix * 4, ncclFloat, ncclSum, 0, comm->comms[i], streams[i]);
            if (r != ncclSuccess) {
                nccl().GroupEnd();
                return api_fail(RRS_ERR_COMM, std::string("ncclReduce: ") + nccl().GetErrorString(r));
            }
        }
        RRS_NCCL_CHECK(nccl().GroupEnd());
        for (int i = 0; i < n_local; ++i) scenes[i]->impl.stats.kernel_launches += 1;  // NCCL's reduce kernel
    }
    // ---- rank 0: sum -> mean, NaN / negative census (main.rs:81-89); waits for the image ----
    if (root_local >= 0) {
        SceneImpl& s = scenes[root_local]->impl;
        int rc = wf_resolve(&s, s.accum, params->width, params->height, spp_total ? spp_total : 1u, out_rgb, out_is_device != 0,
                            streams[root_local], err);
        if (rc != RRS_OK) return api_fail(rc, err);
    }
    return RRS_OK;
}

}  // extern "C"
