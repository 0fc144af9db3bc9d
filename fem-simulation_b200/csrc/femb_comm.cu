// Multi-GPU plumbing of the mesh-partitioned steps behind the C ABI: communicator, halo exchange, scalar all-reduce.
//
// The reference is single-GPU (zero NCCL / MPI call sites); these entry points are what a partitioned
// caller of the operator ABI needs (SURVEY.md 8b "femb_comm_init / halo_exchange / allreduce", 8e).
//
// A halo exchange is three stream-ordered operations on the caller's stream: a pack kernel gathers the rows every
// neighbour needs into one staging buffer, one NCCL group of ncclSend / ncclRecv pairs moves the per-neighbour segments
// over NVLink, an unpack kernel scatters the received segments into the ghost rows.  Nothing synchronises with the host,
// so a whole step -- ~120 exchanges between its kernels -- is captured into ONE CUDA graph per rank.
//
// NCCL is loaded with dlopen at femb_comm_init (the library itself has no link-time dependency on it: single-GPU users
// never need it); in a PyTorch process this resolves to the libnccl.so.2 torch already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include "femb_common.cuh"

namespace femb {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;            // function table of the process-wide NCCL library: written once, read-only afterwards

static int load_nccl() {
  if (g_nccl.handle != nullptr) return FEMB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h != nullptr) break;
  }
  if (h == nullptr) {
    set_error("femb_comm: cannot load libnccl.so.2: %s", dlerror());
    return FEMB_E_CUDA;
  }
  NcclApi a;
  a.handle = h;
#define FEMB_SYM(field, name)                                                   \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));                 \
  if (a.field == nullptr) {                                                      \
    set_error("femb_comm: libnccl has no symbol %s", name);                      \
    return FEMB_E_CUDA;                                                          \
  }
  FEMB_SYM(GetUniqueId, "ncclGetUniqueId")
  FEMB_SYM(CommInitRank, "ncclCommInitRank")
  FEMB_SYM(CommDestroy, "ncclCommDestroy")
  FEMB_SYM(CommAbort, "ncclCommAbort")
  FEMB_SYM(GroupStart, "ncclGroupStart")
  FEMB_SYM(GroupEnd, "ncclGroupEnd")
  FEMB_SYM(Send, "ncclSend")
  FEMB_SYM(Recv, "ncclRecv")
  FEMB_SYM(AllReduce, "ncclAllReduce")
  FEMB_SYM(GetErrorString, "ncclGetErrorString")
#undef FEMB_SYM
  g_nccl = a;
  return FEMB_OK;
}

#define FEMB_NCCL(call, what)                                                                   \
  do {                                                                                          \
    const ncclResult_t r_ = (call);                                                             \
    if (r_ != ncclSuccess) {                                                                    \
      set_error("%s: NCCL error: %s", what, g_nccl.GetErrorString(r_));                         \
      return FEMB_E_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

// rows of a vector (width floats per row) <-> contiguous staging buffer
__global__ void __launch_bounds__(256) halo_pack_kernel(int n, int width, const int* __restrict__ rows, const float* __restrict__ vec,
                                                       float* __restrict__ buf) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  const int i = t / width, c = t - i * width;
  buf[t] = vec[(size_t)__ldg(rows + i) * width + c];
}
__global__ void __launch_bounds__(256) halo_unpack_kernel(int n, int width, const int* __restrict__ rows, const float* __restrict__ buf,
                                                         float* __restrict__ vec) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * width) return;
  const int i = t / width, c = t - i * width;
  vec[(size_t)__ldg(rows + i) * width + c] = buf[t];
}

}  // namespace femb

using namespace femb;

struct FembComm {
  ncclComm_t comm;
  int rank, nranks, device;
};

extern "C" {

int femb_comm_unique_id(unsigned char* id_host) {
  FEMB_REQUIRE(id_host, "null pointer");
  int rc = load_nccl();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == FEMB_COMM_ID_BYTES, "FEMB_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
  ncclUniqueId id;
  FEMB_NCCL(g_nccl.GetUniqueId(&id), "femb_comm_unique_id");
  memcpy(id_host, &id, sizeof(id));
  return FEMB_OK;
}

int femb_comm_init(FembComm** out, const unsigned char* id_host, int rank, int nranks) {
  FEMB_REQUIRE(out && id_host && nranks >= 1 && rank >= 0 && rank < nranks, "bad arguments");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  FembComm* c = new FembComm();
  c->rank = rank; c->nranks = nranks;
  cudaGetDevice(&c->device);
  const ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    set_error("femb_comm_init: NCCL error: %s", g_nccl.GetErrorString(r));
    delete c;
    return FEMB_E_CUDA;
  }
  *out = c;
  return FEMB_OK;
}

int femb_comm_destroy(FembComm* c) {
  if (c == nullptr) return FEMB_OK;
  if (g_nccl.handle != nullptr) g_nccl.CommDestroy(c->comm);
  delete c;
  return FEMB_OK;
}

int femb_halo_exchange(FembComm* c, const FembHalo* h, float* vec, int width, void* stream) {
  FEMB_REQUIRE(c && h && vec && width >= 1, "bad arguments");
  FEMB_REQUIRE(h->n_nb >= 0 && h->n_nb <= FEMB_MAX_NEIGHBOURS, "neighbour count");
  cudaStream_t st = as_stream(stream);
  int n_send = 0, n_recv = 0;
  for (int k = 0; k < h->n_nb; ++k) { n_send += h->n_send[k]; n_recv += h->n_recv[k]; }
  if (n_send > 0) {
    FEMB_REQUIRE(h->send_rows && h->send_buf, "send lists");
    halo_pack_kernel<<<cdiv((long long)n_send * width, 256), 256, 0, st>>>(n_send, width, h->send_rows, vec, h->send_buf);
    count_launch();
  }
  if (n_send + n_recv > 0) {
    FEMB_REQUIRE(n_recv == 0 || (h->recv_rows && h->recv_buf), "recv lists");
    FEMB_NCCL(g_nccl.GroupStart(), "femb_halo_exchange");
    size_t so = 0, ro = 0;
    for (int k = 0; k < h->n_nb; ++k) {
      if (h->n_send[k] > 0) FEMB_NCCL(g_nccl.Send(h->send_buf + so, (size_t)h->n_send[k] * width, ncclFloat32, h->nb[k], c->comm, st), "ncclSend");
      if (h->n_recv[k] > 0) FEMB_NCCL(g_nccl.Recv(h->recv_buf + ro, (size_t)h->n_recv[k] * width, ncclFloat32, h->nb[k], c->comm, st), "ncclRecv");
      so += (size_t)h->n_send[k] * width; ro += (size_t)h->n_recv[k] * width;
    }
    FEMB_NCCL(g_nccl.GroupEnd(), "femb_halo_exchange");
  }
  if (n_recv > 0) {
    halo_unpack_kernel<<<cdiv((long long)n_recv * width, 256), 256, 0, st>>>(n_recv, width, h->recv_rows, h->recv_buf, vec);
    count_launch();
  }
  return check_launch("femb_halo_exchange");
}

int femb_comm_allreduce(FembComm* c, float* buf, int n, int op, void* stream) {
  FEMB_REQUIRE(c && buf && n >= 0 && (op == FEMB_REDUCE_SUM || op == FEMB_REDUCE_MAX), "bad arguments");
  if (n == 0) return FEMB_OK;
  FEMB_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, op == FEMB_REDUCE_SUM ? ncclSum : ncclMax, c->comm, as_stream(stream)),
            "femb_comm_allreduce");
  return FEMB_OK;
}

}  // extern "C"
