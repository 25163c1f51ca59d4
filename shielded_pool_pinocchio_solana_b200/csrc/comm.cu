// comm.cu -- the one collective on the path: a tiny NCCL all-gather of partial MSM sums.
//
// SURVEY.md 8(e) row 2 / BASELINE.json configs[3]: a single large proof splits each of its MSM point
// sets into `world` contiguous index ranges (one per GPU, resident from load time); every rank runs the
// full Pippenger pipeline on its slice and the partial G1 / G2 sums are all-gathered over NVLink and
// added on every rank.  Batches of independent proofs (configs[2]) never come here: they shard by proof
// and need no collective.
//
// NCCL is bound at run time (dlopen) so that single-GPU users of libg16b200.so do not need it installed;
// in a torch process the already loaded libnccl.so.2 is reused (the Python mirror preloads the NCCL wheel that
// PyTorch ships before the first g16_comm_* call, so that a later `import torch` finds its own build).
#include <dlfcn.h>
#include <nccl.h>

#include "capi.cuh"

namespace g16 {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        // 1. a libnccl the process already holds (e.g. the one bundled with PyTorch): two different NCCL builds in
        //    one process resolve each other's symbols by SONAME and break whichever is loaded second
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        // 2. an explicit path, 3. the system library
        if (!a.handle && getenv("G16_NCCL_LIB")) a.handle = dlopen(getenv("G16_NCCL_LIB"), RTLD_NOW | RTLD_LOCAL);
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            if (a.handle) break;
            a.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        }
        if (!a.handle) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
        a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.AllGather && a.CommDestroy && a.GetErrorString;
        return a;
    }();
    return api;
}

int nccl_fail(const char* what, ncclResult_t r) {
    set_error(std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
    return G16_E_CUDA;
}

// out[i] = sum over ranks of gathered[r * count + i]   (affine, Montgomery; infinity = (0, 0))
template <class F>
__global__ void __launch_bounds__(64) k_sum_ranks(const Affine<F>* __restrict__ gathered, int world, uint32_t count,
                                                  Affine<F>* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int r = 0; r < world; r++) acc.madd(gathered[(size_t)r * count + i]);
    out[i] = acc.to_affine();
}

}  // namespace

void shard_range(size_t n, int rank, int world, size_t* lo, size_t* hi) {
    *lo = n * (size_t)rank / (size_t)world;
    *hi = n * (size_t)(rank + 1) / (size_t)world;
}

template <class F>
int comm_sum_points(g16_ctx* ctx, Affine<F>* d_pts, size_t count, cudaStream_t st) {
    if (ctx->world <= 1 || count == 0) return G16_OK;
    const size_t bytes = sizeof(Affine<F>) * count;
    G16_TRY(ctx->comm_recv.ensure(bytes * ctx->world));
    ncclResult_t r = nccl().AllGather(d_pts, ctx->comm_recv.ptr, bytes, ncclUint8, (ncclComm_t)ctx->comm, st);
    if (r != ncclSuccess) return nccl_fail("ncclAllGather", r);
    k_sum_ranks<F><<<cdiv(count, 64), 64, 0, st>>>((const Affine<F>*)ctx->comm_recv.ptr, ctx->world, (uint32_t)count, d_pts);
    G16_CUDA(cudaGetLastError());
    return G16_OK;
}
template int comm_sum_points<Fp>(g16_ctx*, Affine<Fp>*, size_t, cudaStream_t);
template int comm_sum_points<Fp2>(g16_ctx*, Affine<Fp2>*, size_t, cudaStream_t);

void comm_release(g16_ctx* ctx) {
    if (ctx->comm && nccl().ok) nccl().CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
}

}  // namespace g16

using namespace g16;

extern "C" {

int g16_comm_unique_id(uint8_t id[128]) {
    if (!id) return G16_E_ARG;
    if (!nccl().ok) {
        set_error("g16_comm_unique_id: libnccl.so.2 not found");
        return G16_E_CUDA;
    }
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
    ncclUniqueId u;
    ncclResult_t r = nccl().GetUniqueId(&u);
    if (r != ncclSuccess) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id, &u, 128);
    return G16_OK;
}

int g16_comm_init(g16_ctx* ctx, const uint8_t id[128], int rank, int world) {
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) {
        set_error("g16_comm_init: bad arguments");
        return G16_E_ARG;
    }
    if (!nccl().ok) {
        set_error("g16_comm_init: libnccl.so.2 not found");
        return G16_E_CUDA;
    }
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    comm_release(ctx);
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t comm = nullptr;
    ncclResult_t r = nccl().CommInitRank(&comm, world, u, rank);
    if (r != ncclSuccess) return nccl_fail("ncclCommInitRank", r);
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return G16_OK;
}

}  // extern "C"
