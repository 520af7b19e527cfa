// comm.cu -- the optional inter-GPU communicator of libr3dgpu.
//
// Image pairs shard over GPUs with no exchange at all; the bundle adjustment is the
// one path with a real exchange step: the partial reduced camera systems of the point partitions
// are summed once per LM iteration.  That sum is an ncclAllReduce over NVLink on the worker's own
// stream, so it is ordered with the Schur kernel before it and the Cholesky after it without a host
// round trip.  NCCL is resolved with dlopen at run time: the library has no link-time dependency on
// it and loads (and exports every symbol) on a machine without NCCL or without a GPU.
#include "r3d_internal.cuh"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {

struct NcclId { char internal[R3D_COMM_ID_BYTES]; };  // ncclUniqueId: 128 opaque bytes, passed by value

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
};

constexpr int kNcclFloat64 = 8;  // ncclDataType_t::ncclFloat64 (nccl.h; stable since NCCL 2.0)

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[3] = {nullptr, nullptr, "libnccl.so.2"};
    // 1. the copy the process already holds (torch bundles one): two NCCL instances in one process
    //    would each grab their own NVLink buffers
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) {
      names[1] = std::getenv("R3D_NCCL_LIB");
      for (int k = 1; k < 3 && !h; ++k)
        if (names[k]) h = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) {
      const char* e = dlerror();
      api.error = std::string("libnccl.so.2 not found (set R3D_NCCL_LIB): ") + (e ? e : "");
      return;
    }
    api.lib = h;
    api.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    api.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
      api.error = "libnccl.so.2 lacks a required symbol";
      api.lib = nullptr;
    }
  });
  return api;
}

std::string nccl_err(int rc) {
  NcclApi& a = nccl();
  return a.GetErrorString ? std::string(a.GetErrorString(rc)) : ("ncclResult " + std::to_string(rc));
}

}  // namespace

namespace r3d {

int comm_allreduce(r3d_ctx* ctx, cudaStream_t stream, double* buf, size_t n, CommOp op) {
  if (!ctx->nccl_comm || n == 0) return R3D_OK;
  const int rc = nccl().AllReduce(buf, buf, n, kNcclFloat64, (int)op, ctx->nccl_comm, stream);
  if (rc != 0) return fail(ctx, R3D_ERR_CUDA, "ncclAllReduce: " + nccl_err(rc));
  return R3D_OK;
}

}  // namespace r3d

using namespace r3d;

extern "C" {

int r3d_comm_unique_id(r3d_ctx* ctx, uint8_t id[R3D_COMM_ID_BYTES]) {
  if (!ctx || !id) return fail(ctx, R3D_ERR_INVALID, "r3d_comm_unique_id: bad arguments");
  NcclApi& a = nccl();
  if (!a.lib) return fail(ctx, R3D_ERR_UNSUPPORTED, "r3d_comm_unique_id: " + a.error);
  NcclId nid;
  const int rc = a.GetUniqueId(&nid);
  if (rc != 0) return fail(ctx, R3D_ERR_CUDA, "ncclGetUniqueId: " + nccl_err(rc));
  std::memcpy(id, nid.internal, R3D_COMM_ID_BYTES);
  return R3D_OK;
}

int r3d_comm_init(r3d_ctx* ctx, int world, int rank, const uint8_t id[R3D_COMM_ID_BYTES]) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return fail(ctx, R3D_ERR_INVALID, "r3d_comm_init: bad arguments");
  if (ctx->nccl_comm) return fail(ctx, R3D_ERR_INVALID, "r3d_comm_init: a communicator is already attached");
  NcclApi& a = nccl();
  if (!a.lib) return fail(ctx, R3D_ERR_UNSUPPORTED, "r3d_comm_init: " + a.error);
  R3D_CUDA_TRY(ctx, cudaSetDevice(ctx->workers[0].device));  // one process per GPU: rank <-> workers[0]
  NcclId nid;
  std::memcpy(nid.internal, id, R3D_COMM_ID_BYTES);
  void* comm = nullptr;
  const int rc = a.CommInitRank(&comm, world, nid, rank);
  if (rc != 0) return fail(ctx, R3D_ERR_CUDA, "ncclCommInitRank: " + nccl_err(rc));
  ctx->nccl_comm = comm;
  ctx->comm_world = world;
  ctx->comm_rank = rank;
  return R3D_OK;
}

int r3d_comm_destroy(r3d_ctx* ctx) {
  if (!ctx) return R3D_ERR_INVALID;
  if (ctx->nccl_comm) {
    cudaSetDevice(ctx->workers[0].device);
    cudaStreamSynchronize(ctx->workers[0].stream);
    nccl().CommDestroy(ctx->nccl_comm);
  }
  ctx->nccl_comm = nullptr;
  ctx->comm_world = 1;
  ctx->comm_rank = 0;
  return R3D_OK;
}

int r3d_comm_world(const r3d_ctx* ctx) { return ctx ? ctx->comm_world : 1; }

}  // extern "C"
