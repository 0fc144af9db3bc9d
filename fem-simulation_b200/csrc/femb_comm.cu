// Multi-GPU plumbing of the mesh-partitioned steps behind the C ABI: communicator, halo exchange, scalar all-reduce.
//
// The reference is single-GPU (zero NCCL / MPI call sites); these entry points are what a partitioned
// caller of the operator ABI needs (SURVEY.md 8b "femb_comm_init / halo_exchange / allreduce", 8e).
//
// A halo exchange is stream-ordered work on the caller's stream; nothing synchronises with the host, so a whole step -- ~120
// exchanges between its kernels -- is captured into ONE CUDA graph per rank.  Two transports:
//   * peer memory (femb_comm_p2p_enable; the default of the Python drivers on one NVLink box): ONE kernel.  Every CTA
//     gathers its share of the rows a neighbour needs and stores them straight into that neighbour's receive arena through a
//     CUDA-IPC mapping (NVLink stores, no staging copy, no NCCL proxy), the last CTA to arrive publishes an epoch flag with a
//     system-scope release store; then the CTAs spin (acquire, bounded) on the flags of their own arena and scatter the
//     received segments into the ghost rows.  Arena slots are double-buffered by epoch parity: a neighbour can be one exchange ahead, never two
//     (its push e+2 follows its pull e+1, which waited for my push e+1, which follows my pull e).
//   * NCCL: a pack kernel, one group of ncclSend / ncclRecv pairs, an unpack kernel (any topology; ~3x the latency).
//
// NCCL is loaded with dlopen at femb_comm_init (the library itself has no link-time dependency on it: single-GPU users
// never need it); in a PyTorch process this resolves to the libnccl.so.2 torch already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <vector>

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
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
  FEMB_SYM(AllGather, "ncclAllGather")
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

// ---- peer-memory transport --------------------------------------------------------------------------------------------------
constexpr int NB = FEMB_MAX_NEIGHBOURS;
// device state words: [0, NB) epochs of my pushes per neighbour slot, [NB, 2 NB) epochs of my pulls, then two "CTAs done"
// counters and the error word (1 = a pull gave up waiting)
constexpr int ST_SEND = 0, ST_RECV = NB, ST_DONE_PUSH = 2 * NB, ST_DONE_PULL = 2 * NB + 1, ST_ERR = 2 * NB + 2, ST_WORDS = 2 * NB + 4;
constexpr long long P2P_SPIN_CYCLES = 60000000000ll;            // ~30 s: a lost peer becomes an error, not a hung GPU

struct P2PArgs {
  int n_nb, width;
  int off[NB + 1];                 // prefix sums of the rows per neighbour (send rows for push, recv rows for pull)
  int slot[NB];                    // push: my slot in the neighbour's arena; pull: the neighbour's slot in mine
  int ep[NB];                      // index of the neighbour in the communicator's table: where its epochs live in `state`
  unsigned char* data[NB];         // push: the neighbour's arena; pull: my arena
  unsigned* flag[NB];              // push: my flag word in the neighbour's arena; pull: the neighbour's flag word in mine
  size_t slot_bytes;
  const int* rows;
  float* vec;
  unsigned* state;
};
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int nb_of_row(const P2PArgs& a, int i) {
  int k = 0;
#pragma unroll
  for (int j = 1; j < NB; ++j) k += (j < a.n_nb && i >= a.off[j]) ? 1 : 0;
  return k;
}

// ONE kernel per exchange.  Every CTA first stores its share of the boundary rows into the neighbours' arenas and arrives; the
// last arriver publishes the epoch flags.  Then every CTA waits for the neighbours' flags (they depend on the neighbours' pushes
// only, never on this kernel's waiting CTAs: no circular wait as long as the grid is co-resident -- the launcher caps it at
// P2P_MAX_CTAS and the loops stride) and scatters its share of the received rows.
constexpr int P2P_MAX_CTAS = 296;                               // 2 per SM of the smallest part this is built for: always co-resident
__global__ void __launch_bounds__(256) halo_exchange_kernel(P2PArgs a, P2PArgs b) {
  const int stride = gridDim.x * blockDim.x;
  // ---- push (a: send side) ----------------------------------------------------------------------------------------------
  const int total_s = a.off[a.n_nb] * a.width;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total_s; t += stride) {
    const int i = t / a.width, c = t - i * a.width;
    const int k = nb_of_row(a, i);
    const unsigned e = a.state[ST_SEND + a.ep[k]] + 1u;      // epoch of this push to neighbour k
    float* dst = reinterpret_cast<float*>(a.data[k] + ((size_t)(e & 1u) * NB + a.slot[k]) * a.slot_bytes);
    dst[(size_t)(i - a.off[k]) * a.width + c] = a.vec[(size_t)__ldg(a.rows + i) * a.width + c];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(a.state + ST_DONE_PUSH, 1u);
    if (prev == gridDim.x - 1) {                                  // every CTA has stored and fenced: publish
      __threadfence_system();
      for (int k = 0; k < a.n_nb; ++k)
        if (a.off[k + 1] > a.off[k]) {
          const unsigned e = a.state[ST_SEND + a.ep[k]] + 1u;
          st_release_sys(a.flag[k], e);
          a.state[ST_SEND + a.ep[k]] = e;
        }
      a.state[ST_DONE_PUSH] = 0u;
    }
    // ---- wait for the neighbours (b: receive side) ----------------------------------------------------------------------
    const long long t0 = clock64();
    for (int k = 0; k < b.n_nb; ++k)
      if (b.off[k + 1] > b.off[k]) {
        const unsigned e = b.state[ST_RECV + b.ep[k]] + 1u;
        while ((int)(ld_acquire_sys(b.flag[k]) - e) < 0) {
          // a lost peer is reported once; after that nobody waits any more (the results are garbage, the error word says so)
          if (*reinterpret_cast<volatile unsigned*>(b.state + ST_ERR) != 0u) break;
          if (clock64() - t0 > P2P_SPIN_CYCLES) { b.state[ST_ERR] = 1u; break; }
          __nanosleep(32);
        }
      }
  }
  __syncthreads();
  // ---- pull ------------------------------------------------------------------------------------------------------------
  const int total_r = b.off[b.n_nb] * b.width;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total_r; t += stride) {
    const int i = t / b.width, c = t - i * b.width;
    const int k = nb_of_row(b, i);
    const unsigned e = b.state[ST_RECV + b.ep[k]] + 1u;
    const float* src = reinterpret_cast<const float*>(b.data[k] + ((size_t)(e & 1u) * NB + b.slot[k]) * b.slot_bytes);
    // the arena is written by the peer over NVLink: read through L2 (an L1 line could be two exchanges old)
    b.vec[(size_t)__ldg(b.rows + i) * b.width + c] = __ldcg(src + (size_t)(i - b.off[k]) * b.width + c);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(b.state + ST_DONE_PULL, 1u);
    if (prev == gridDim.x - 1) {
      for (int k = 0; k < b.n_nb; ++k)
        if (b.off[k + 1] > b.off[k]) b.state[ST_RECV + b.ep[k]] += 1u;
      b.state[ST_DONE_PULL] = 0u;
    }
  }
}

}  // namespace femb

using namespace femb;

struct FembComm {
  ncclComm_t comm;
  int rank, nranks, device;
  // peer-memory transport (femb_comm_p2p_enable)
  bool p2p = false;
  size_t slot_bytes = 0, data_bytes = 0;
  unsigned char* arena = nullptr;                 // [2][NB][slot_bytes] receive slots, then NB flag words
  unsigned* state = nullptr;                      // ST_WORDS device words
  int n_nb = 0;
  int nb[FEMB_MAX_NEIGHBOURS];                    // my neighbour table: slot k of my arena belongs to rank nb[k]
  int my_slot_at[FEMB_MAX_NEIGHBOURS];            // my slot in the arena of nb[k]
  unsigned char* peer[FEMB_MAX_NEIGHBOURS];       // IPC mapping of the arena of nb[k]
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
  if (c->p2p) {
    // the caller destroys the communicators of all ranks together, after a barrier of its own (comm.py): no rank may still be
    // storing into this arena
    cudaDeviceSynchronize();
    for (int k = 0; k < c->n_nb; ++k)
      if (c->peer[k] != nullptr) cudaIpcCloseMemHandle(c->peer[k]);
  }
  if (c->arena != nullptr) cudaFree(c->arena);
  if (c->state != nullptr) cudaFree(c->state);
  if (g_nccl.handle != nullptr) g_nccl.CommDestroy(c->comm);
  delete c;
  return FEMB_OK;
}

int femb_comm_p2p_enable(FembComm* c, const int* nb_host, int n_nb, size_t slot_bytes, void* stream) {
  FEMB_REQUIRE(c && n_nb >= 0 && n_nb <= FEMB_MAX_NEIGHBOURS && (n_nb == 0 || nb_host) && slot_bytes > 0, "bad arguments");
  FEMB_REQUIRE(!c->p2p, "already enabled");
  cudaStream_t st = as_stream(stream);
  slot_bytes = (slot_bytes + 255) / 256 * 256;
  // one record per rank: IPC handle of its arena, its neighbour table, its slot size
  struct Rec { cudaIpcMemHandle_t h; int n_nb; int nb[FEMB_MAX_NEIGHBOURS]; unsigned long long slot_bytes; int ok; int pad[3]; };
  static_assert(sizeof(Rec) % 16 == 0, "record size");
  Rec mine;
  memset(&mine, 0, sizeof(mine));
  mine.n_nb = n_nb; mine.slot_bytes = slot_bytes; mine.ok = 1;
  for (int k = 0; k < n_nb; ++k) mine.nb[k] = nb_host[k];
  const size_t data_bytes = (size_t)2 * NB * slot_bytes;
  unsigned char* arena = nullptr;
  unsigned* state = nullptr;
  if (cudaMalloc(&arena, data_bytes + 256) != cudaSuccess || cudaMalloc(&state, ST_WORDS * sizeof(unsigned)) != cudaSuccess ||
      cudaMemsetAsync(arena, 0, data_bytes + 256, st) != cudaSuccess || cudaMemsetAsync(state, 0, ST_WORDS * sizeof(unsigned), st) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine.h, arena) != cudaSuccess) {
    cudaGetLastError();
    mine.ok = 0;                                   // still take part in the collective below: every rank must learn it
  }
  Rec* all_dev = nullptr;
  if (cudaMalloc(&all_dev, sizeof(Rec) * c->nranks) != cudaSuccess) {
    set_error("femb_comm_p2p_enable: out of memory");
    return FEMB_E_CUDA;
  }
  cudaMemcpyAsync(all_dev + c->rank, &mine, sizeof(Rec), cudaMemcpyHostToDevice, st);
  FEMB_NCCL(g_nccl.AllGather(all_dev + c->rank, all_dev, sizeof(Rec), ncclInt8, c->comm, st), "femb_comm_p2p_enable");
  std::vector<Rec> all(c->nranks);
  cudaMemcpyAsync(all.data(), all_dev, sizeof(Rec) * c->nranks, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  cudaFree(all_dev);
  bool ok = true;
  for (int r = 0; r < c->nranks; ++r) ok = ok && all[r].ok && all[r].slot_bytes == slot_bytes;
  // my slot in every neighbour's table; the relation must be symmetric
  int my_slot_at[FEMB_MAX_NEIGHBOURS];
  for (int k = 0; k < n_nb && ok; ++k) {
    const int q = nb_host[k];
    my_slot_at[k] = -1;
    if (q < 0 || q >= c->nranks || q == c->rank) { ok = false; break; }
    for (int j = 0; j < all[q].n_nb; ++j)
      if (all[q].nb[j] == c->rank) my_slot_at[k] = j;
    if (my_slot_at[k] < 0) ok = false;
  }
  unsigned char* peer[FEMB_MAX_NEIGHBOURS] = {};
  int opened = 0;
  if (ok)
    for (; opened < n_nb; ++opened) {
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[nb_host[opened]].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); break; }
      peer[opened] = static_cast<unsigned char*>(ptr);
    }
  // second collective: peer memory is used only if EVERY rank could map all its neighbours
  int* flag_dev = nullptr;
  cudaMalloc(&flag_dev, sizeof(int));
  const int good = (ok && opened == n_nb) ? 1 : 0;
  cudaMemcpyAsync(flag_dev, &good, sizeof(int), cudaMemcpyHostToDevice, st);
  FEMB_NCCL(g_nccl.AllReduce(flag_dev, flag_dev, 1, ncclInt32, ncclMin, c->comm, st), "femb_comm_p2p_enable");
  int all_good = 0;
  cudaMemcpyAsync(&all_good, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  cudaFree(flag_dev);
  if (!all_good) {
    for (int k = 0; k < opened; ++k) cudaIpcCloseMemHandle(peer[k]);
    // nobody frees its arena before every rank has closed its mappings of it
    int* bar_dev = nullptr;
    if (cudaMalloc(&bar_dev, sizeof(int)) == cudaSuccess) {
      cudaMemsetAsync(bar_dev, 0, sizeof(int), st);
      g_nccl.AllReduce(bar_dev, bar_dev, 1, ncclInt32, ncclSum, c->comm, st);
      cudaStreamSynchronize(st);
      cudaFree(bar_dev);
    }
    if (arena) cudaFree(arena);
    if (state) cudaFree(state);
    set_error("femb_comm_p2p_enable: peer memory not available between all neighbours (CUDA IPC / NVLink peer access)");
    return FEMB_E_CUDA;
  }
  c->p2p = true; c->slot_bytes = slot_bytes; c->data_bytes = data_bytes; c->arena = arena; c->state = state; c->n_nb = n_nb;
  for (int k = 0; k < n_nb; ++k) { c->nb[k] = nb_host[k]; c->my_slot_at[k] = my_slot_at[k]; c->peer[k] = peer[k]; }
  return FEMB_OK;
}

int femb_comm_p2p_error(FembComm* c, int* err_host) {
  FEMB_REQUIRE(c && err_host, "bad arguments");
  *err_host = 0;
  if (!c->p2p) return FEMB_OK;
  unsigned e = 0;
  if (cudaMemcpy(&e, c->state + ST_ERR, sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("femb_comm_p2p_error: %s", cudaGetErrorString(cudaGetLastError()));
    return FEMB_E_CUDA;
  }
  *err_host = (int)e;
  return FEMB_OK;
}

// vec[recv rows] <- neighbours' vec[send rows] through peer memory; returns false if this plan cannot use it
static bool halo_exchange_p2p(FembComm* c, const FembHalo* h, float* vec, int width, cudaStream_t st) {
  P2PArgs push, pull;
  push.n_nb = pull.n_nb = h->n_nb; push.width = pull.width = width;
  push.slot_bytes = pull.slot_bytes = c->slot_bytes; push.state = pull.state = c->state;
  push.rows = h->send_rows; pull.rows = h->recv_rows; push.vec = pull.vec = vec;
  push.off[0] = pull.off[0] = 0;
  unsigned* my_flags = reinterpret_cast<unsigned*>(c->arena + c->data_bytes);
  for (int k = 0; k < h->n_nb; ++k) {
    int j = -1;
    for (int q = 0; q < c->n_nb; ++q)
      if (c->nb[q] == h->nb[k]) j = q;
    if (j < 0) return false;
    if ((size_t)h->n_send[k] * width * sizeof(float) > c->slot_bytes || (size_t)h->n_recv[k] * width * sizeof(float) > c->slot_bytes) return false;
    push.off[k + 1] = push.off[k] + h->n_send[k];
    pull.off[k + 1] = pull.off[k] + h->n_recv[k];
    push.slot[k] = c->my_slot_at[j];
    push.data[k] = c->peer[j];
    push.flag[k] = reinterpret_cast<unsigned*>(c->peer[j] + c->data_bytes) + c->my_slot_at[j];
    pull.slot[k] = j;
    pull.data[k] = c->arena;
    pull.flag[k] = my_flags + j;
    push.ep[k] = pull.ep[k] = j;
  }
  const int n_send = push.off[h->n_nb], n_recv = pull.off[h->n_nb];
  if (n_send + n_recv > 0) {
    const long long work = (long long)(n_send > n_recv ? n_send : n_recv) * width;
    unsigned grid = cdiv(work, 256);
    if (grid > (unsigned)P2P_MAX_CTAS) grid = P2P_MAX_CTAS;
    halo_exchange_kernel<<<grid, 256, 0, st>>>(push, pull);
    count_launch();
  }
  return true;
}

int femb_halo_exchange(FembComm* c, const FembHalo* h, float* vec, int width, void* stream) {
  FEMB_REQUIRE(c && h && vec && width >= 1, "bad arguments");
  FEMB_REQUIRE(h->n_nb >= 0 && h->n_nb <= FEMB_MAX_NEIGHBOURS, "neighbour count");
  cudaStream_t st = as_stream(stream);
  int n_send = 0, n_recv = 0;
  for (int k = 0; k < h->n_nb; ++k) { n_send += h->n_send[k]; n_recv += h->n_recv[k]; }
  FEMB_REQUIRE((n_send == 0 || h->send_rows) && (n_recv == 0 || h->recv_rows), "row lists");
  if (c->p2p) {
    // (both sides of a pair take the same decision: the table and the slot size are communicator-wide, the sizes symmetric)
    if (halo_exchange_p2p(c, h, vec, width, st)) return check_launch("femb_halo_exchange (peer memory)");
    set_error("femb_halo_exchange: plan does not fit the peer-memory arena (neighbour not registered or segment > slot)");
    return FEMB_E_INVALID;
  }
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
