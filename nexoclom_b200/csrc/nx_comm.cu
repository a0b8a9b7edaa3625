// nexoclom_b200 -- the one collective of the path: packets are sharded over
// the GPUs of a box by global id, every GPU integrates its own slice, and the per-GPU products
// (image f64 + counts i64; LOS radiance f64 + hit counts i64) are combined with ONE NCCL
// all-reduce (sum) per product.  Exposed through the C ABI so that a non-Python caller of
// libnexo_b200.so can combine its ranks' products too.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2"): inside a torch process that is the
// copy torch already loaded (same SONAME), elsewhere the system library; the single-GPU paths
// of libnexo_b200.so therefore do not depend on NCCL being installed.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>

#include "../../include/nexoclom_b200.h"

struct nx_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  void* staging = nullptr;
  size_t staging_bytes = 0;
};

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi& api() {
  static NcclApi a;            // resolved once; immutable afterwards
  if (a.handle || !a.err.empty()) return a;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  if (!a.handle) { a.err = std::string("cannot load NCCL: ") + dlerror(); return a; }
  auto sym = [&](const char* n) { return dlsym(a.handle, n); };
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy || !a.GetErrorString) {
    a.err = "NCCL library lacks a required symbol";
    a.handle = nullptr;
  }
  return a;
}

thread_local std::string g_comm_err;
int fail(const std::string& what, ncclResult_t r) {
  g_comm_err = what + ": " + (api().GetErrorString ? api().GetErrorString(r) : "nccl error");
  return -1000 - (int)r;
}
}  // namespace

extern "C" {

const char* nx_comm_last_error(void) { return g_comm_err.c_str(); }

int nx_comm_unique_id(void* out128) {
  static_assert(sizeof(ncclUniqueId) == NX_COMM_ID_BYTES, "ncclUniqueId size");
  NcclApi& a = api();
  if (!a.handle) { g_comm_err = a.err; return -1; }
  ncclUniqueId id;
  ncclResult_t r = a.GetUniqueId(&id);
  if (r != ncclSuccess) return fail("ncclGetUniqueId", r);
  std::memcpy(out128, &id, sizeof(id));
  return 0;
}

int nx_comm_create(int device, const void* unique_id128, int rank, int world, nx_comm** out) {
  if (!out) return -1;
  *out = nullptr;
  NcclApi& a = api();
  if (!a.handle) { g_comm_err = a.err; return -1; }
  if (cudaSetDevice(device) != cudaSuccess) { g_comm_err = "cudaSetDevice failed"; return -1; }
  ncclUniqueId id;
  std::memcpy(&id, unique_id128, sizeof(id));
  nx_comm* c = new nx_comm();
  c->rank = rank; c->world = world;
  ncclResult_t r = a.CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) { delete c; return fail("ncclCommInitRank", r); }
  *out = c;
  return 0;
}

// In-place SUM over the ranks of `comm`, enqueued on `cuda_stream` (nullptr: default stream).
// dtype: NX_DTYPE_F64 or NX_DTYPE_I64.
int nx_allreduce(nx_comm* comm, void* dev_buffer, long long count, int dtype, void* cuda_stream) {
  if (!comm || !comm->comm) { g_comm_err = "null communicator"; return -1; }
  if (count <= 0) return 0;
  const ncclDataType_t t = dtype == NX_DTYPE_F64 ? ncclFloat64 : ncclInt64;
  ncclResult_t r = api().AllReduce(dev_buffer, dev_buffer, (size_t)count, t, ncclSum, comm->comm,
                                   reinterpret_cast<cudaStream_t>(cuda_stream));
  if (r != ncclSuccess) return fail("ncclAllReduce", r);
  return 0;
}

// Host-buffer variant: H2D into a staging buffer owned by the communicator, all-reduce, D2H.
int nx_allreduce_host(nx_comm* comm, void* host_buffer, long long count, int dtype) {
  if (!comm || !comm->comm) { g_comm_err = "null communicator"; return -1; }
  if (count <= 0) return 0;
  const size_t bytes = (size_t)count * 8;
  if (bytes > comm->staging_bytes) {
    cudaFree(comm->staging);
    comm->staging = nullptr; comm->staging_bytes = 0;
    if (cudaMalloc(&comm->staging, bytes) != cudaSuccess) { g_comm_err = "staging allocation failed"; return -1; }
    comm->staging_bytes = bytes;
  }
  cudaError_t e = cudaMemcpy(comm->staging, host_buffer, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { g_comm_err = cudaGetErrorString(e); return -(int)e; }
  int r = nx_allreduce(comm, comm->staging, count, dtype, nullptr);
  if (r) return r;
  e = cudaMemcpy(host_buffer, comm->staging, bytes, cudaMemcpyDeviceToHost);   // syncs stream 0
  if (e != cudaSuccess) { g_comm_err = cudaGetErrorString(e); return -(int)e; }
  return 0;
}

int nx_comm_rank(nx_comm* comm, int* rank, int* world) {
  if (!comm) return -1;
  if (rank) *rank = comm->rank;
  if (world) *world = comm->world;
  return 0;
}

int nx_comm_destroy(nx_comm* comm) {
  if (!comm) return 0;
  if (comm->comm) api().CommDestroy(comm->comm);
  cudaFree(comm->staging);
  delete comm;
  return 0;
}

}  // extern "C"
