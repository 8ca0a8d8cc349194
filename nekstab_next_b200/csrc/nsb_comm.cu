// Inter-GPU communication: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// Replaces the reference's MPI layer reached through Nek5000: gop -> MPI_Allreduce inside glsc3
// (core/nek_vectors.f90:7-12, 94-102) and gslib's pairwise exchange inside dssum
// (core/utils.f90:287, 338).  NCCL is dlopen'ed on first use so a single-GPU process never needs
// it; the communicator is bootstrapped from an ncclUniqueId the host broadcasts with its own
// transport (MPI_Bcast in Nek, torch.distributed in the Python tests).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>

#include "nsb_internal.h"

namespace nsb {

struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};

static Nccl g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return NSB_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    set_error("cannot dlopen libnccl.so.2: %s", dlerror());
    return NSB_ENCCL;
  }
#define SYM(field, name)                                         \
  do {                                                           \
    *(void **)(&g_nccl.field) = dlsym(lib, name);                \
    if (!g_nccl.field) {                                         \
      set_error("NCCL symbol %s missing", name);                 \
      return NSB_ENCCL;                                          \
    }                                                            \
  } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(GetErrorString, "ncclGetErrorString");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
#undef SYM
  g_nccl.lib = lib;
  return NSB_OK;
}

#define NSB_NCCL(call)                                                                        \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess) {                                                                  \
      set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_));       \
      return NSB_ENCCL;                                                                       \
    }                                                                                         \
  } while (0)

// ------------------------------------------------------------------------------------------------
// NVLink peer-memory path.  Every rank exports one "mailbox" allocation through CUDA IPC and maps
// all the others (NVSwitch gives every GPU a direct path to every peer).  Collectives on the hot
// path then are plain kernels:
//   all-reduce of <= 1032 doubles (the H column): each rank STORES its vector into slot
//     [seq&1][rank] of every mailbox, publishes seq with a release store, waits until all peers'
//     flags show seq, and sums the P vectors in rank order (bitwise identical on every rank);
//   halo exchange: the pack kernel stores the interface sums straight into the peer's mailbox.
// Two slots suffice: a rank can only be one collective ahead of the slowest one, because finishing
// collective s+1 needs everybody's contribution to it, which they issue after finishing s.
// NCCL stays for bootstrap, set-up and as the fallback when no mailbox is connected.
// ------------------------------------------------------------------------------------------------
constexpr int ARN = kMaxK + 8;                    // doubles per all-reduce vector slot
struct PeerPtrs { double *p[nsb_context_s::kMaxPeers]; };

// mailbox layout (in doubles): [ar data 2*P*ARN][ar flags 2*P][halo flags 2*P][halo data ...]
__host__ __device__ inline size_t mb_ar_data(int P, int slot, int r) { return ((size_t)slot * P + r) * ARN; }
__host__ __device__ inline size_t mb_ar_flag(int P, int slot, int r) { return (size_t)2 * P * ARN + (size_t)slot * P + r; }
__host__ __device__ inline size_t mb_hx_flag(int P, int slot, int r) { return (size_t)2 * P * ARN + 2 * P + (size_t)slot * P + r; }
__host__ __device__ inline size_t mb_halo_base(int P) { return (((size_t)2 * P * ARN + 4 * P) + 31) & ~(size_t)31; }

__device__ __forceinline__ void st_release_sys(uint64_t *p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
p2p_allreduce_kernel(double *__restrict__ buf, int n, uint64_t seq, int rank, int P, PeerPtrs mail) {
  const int slot = (int)(seq & 1);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double v = buf[j];
    for (int r = 0; r < P; ++r) mail.p[r][mb_ar_data(P, slot, rank) + j] = v;   // peer stores over NVLink
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < P)
    st_release_sys(reinterpret_cast<uint64_t *>(mail.p[threadIdx.x] + mb_ar_flag(P, slot, rank)), seq);
  if (threadIdx.x < P) {
    const uint64_t *f = reinterpret_cast<const uint64_t *>(mail.p[rank] + mb_ar_flag(P, slot, threadIdx.x));
    while (ld_acquire_sys(f) != seq) { }
  }
  __syncthreads();
  const double *mine = mail.p[rank];
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < P; ++r) s += mine[mb_ar_data(P, slot, r) + j];
    buf[j] = s;
  }
}

// pack the interface sums of nf fields straight into the peer's mailbox
__global__ void pack_p2p_kernel(const double *__restrict__ node_sum, const int32_t *__restrict__ nodes, int64_t n,
                                double *__restrict__ dst, int64_t ns_stride) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (t < n) dst[(int64_t)f * n + t] = node_sum[(int64_t)f * ns_stride + nodes[t]];
}

struct HaloFlags { uint64_t *p[nsb_context_s::kMaxPeers]; int n; };
__global__ void p2p_flags_kernel(HaloFlags fl, uint64_t seq) {
  if ((int)threadIdx.x < fl.n) st_release_sys(fl.p[threadIdx.x], seq);
}

__global__ void unpack_wait_add_kernel(double *__restrict__ node_sum, const int32_t *__restrict__ nodes, int64_t n,
                                       const double *__restrict__ src, int64_t ns_stride,
                                       const uint64_t *__restrict__ flag, uint64_t seq) {
  if (threadIdx.x == 0)
    while (ld_acquire_sys(flag) != seq) { }
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (t < n) node_sum[(int64_t)f * ns_stride + nodes[t]] += src[(int64_t)f * n + t];
}

int halo_exchange_p2p(nsb_sem_t S, int nf, cudaStream_t st) {
  nsb_context_t ctx = S->ctx;
  const int P = ctx->nranks;
  const uint64_t seq = ++ctx->hx_seq;
  const int slot = (int)(seq & 1);
  const int64_t nifc = S->nshared - S->n_local;
  HaloFlags fl;
  fl.n = 0;
  for (auto &Pr : S->peers) {
    const unsigned nb = (unsigned)((Pr.n + 255) / 256);
    double *dst = ctx->peer_mail[Pr.rank] + mb_halo_base(P) + Pr.peer_off + (int64_t)slot * S->ns_fields * Pr.n;
    pack_p2p_kernel<<<dim3(nb, nf), 256, 0, st>>>(S->node_sum_d, Pr.idx_d, Pr.n, dst, nifc);
    fl.p[fl.n++] = reinterpret_cast<uint64_t *>(ctx->peer_mail[Pr.rank] + mb_hx_flag(P, slot, ctx->rank));
    ctx->launches++;
  }
  // stores of the pack kernels are complete at the kernel boundary; then publish the sequence number
  p2p_flags_kernel<<<1, 32, 0, st>>>(fl, seq);
  ctx->launches++;
  for (auto &Pr : S->peers) {
    const unsigned nb = (unsigned)((Pr.n + 255) / 256);
    const double *src = ctx->mail_d + mb_halo_base(P) + Pr.my_off + (int64_t)slot * S->ns_fields * Pr.n;
    const uint64_t *flag = reinterpret_cast<const uint64_t *>(ctx->mail_d + mb_hx_flag(P, slot, Pr.rank));
    unpack_wait_add_kernel<<<dim3(nb, nf), 256, 0, st>>>(S->node_sum_d, Pr.idx_d, Pr.n, src, nifc, flag, seq);
    ctx->launches++;
  }
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

int comm_init(nsb_context_t ctx, const void *unique_id) {
  NSB_CHECK(load_nccl());
  static_assert(sizeof(ncclUniqueId) <= NSB_UNIQUE_ID_BYTES, "unique id size");
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm;
  cudaSetDevice(ctx->device);
  NSB_NCCL(g_nccl.CommInitRank(&comm, ctx->nranks, id, ctx->rank));
  ctx->nccl_comm = comm;
  return NSB_OK;
}

int comm_destroy(nsb_context_t ctx) {
  if (ctx->nccl_comm && g_nccl.lib) {
    g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  return NSB_OK;
}

int allreduce_sum_d(nsb_context_t ctx, double *buf_d, int n) {
  if (ctx->nranks == 1 || n == 0) return NSB_OK;
  if (ctx->p2p && n <= kMaxK + 8) {
    PeerPtrs mail;
    for (int r = 0; r < ctx->nranks; ++r) mail.p[r] = ctx->peer_mail[r];
    const uint64_t seq = ++ctx->ar_seq;
    ProfScope ps(ctx, PC_SMALL, 8.0 * n * ctx->nranks);
    p2p_allreduce_kernel<<<1, 256, 0, ctx->stream>>>(buf_d, n, seq, ctx->rank, ctx->nranks, mail);
    ctx->launches++;
    NSB_CUDA(cudaGetLastError());
    return NSB_OK;
  }
  NSB_REQUIRE(ctx->nccl_comm, "allreduce: no communicator");
  NSB_NCCL(g_nccl.AllReduce(buf_d, buf_d, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm,
                            ctx->stream));
  return NSB_OK;
}

int sendrecv_d(nsb_context_t ctx, const std::vector<nsb_sem_s::Peer> &peers, int nf, cudaStream_t st) {
  if (peers.empty()) return NSB_OK;
  NSB_REQUIRE(ctx->nccl_comm, "sendrecv: no communicator");
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  NSB_NCCL(g_nccl.GroupStart());
  for (const auto &P : peers) {
    NSB_NCCL(g_nccl.Send(P.send_d, (size_t)P.n * nf, ncclDouble, P.rank, comm, st));
    NSB_NCCL(g_nccl.Recv(P.recv_d, (size_t)P.n * nf, ncclDouble, P.rank, comm, st));
  }
  NSB_NCCL(g_nccl.GroupEnd());
  return NSB_OK;
}

// Pure host part of the exchange set-up: intersect this rank's node ids with every other rank's
// sorted id list, move interface nodes behind the private ones, and list each peer's shared nodes
// in ascending global id (both sides of a pair then agree on the order of the packed buffer).
int exchange_plan(int rank, int nranks, const std::vector<int64_t> &gid, const std::vector<int64_t> &cnt,
                  const int64_t *all_sorted, int64_t mx, ExchangePlan &plan) {
  const int64_t nn = (int64_t)gid.size();
  std::vector<std::pair<int64_t, int32_t>> my(nn);
  for (int64_t n = 0; n < nn; ++n) my[n] = {gid[n], (int32_t)n};
  std::sort(my.begin(), my.end());
  std::vector<std::vector<int32_t>> shared_with(nranks);
  std::vector<char> is_ifc(nn, 0);
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) continue;
    const int64_t *other = all_sorted + (size_t)r * mx;
    int64_t a = 0, b = 0;
    while (a < nn && b < cnt[r]) {
      if (my[a].first < other[b]) ++a;
      else if (my[a].first > other[b]) ++b;
      else { shared_with[r].push_back(my[a].second); is_ifc[my[a].second] = 1; ++a; ++b; }
    }
  }
  plan.newpos.assign(nn, 0);
  int64_t nloc = 0;
  for (int64_t n = 0; n < nn; ++n) if (!is_ifc[n]) plan.newpos[n] = nloc++;
  int64_t w = nloc;
  for (int64_t n = 0; n < nn; ++n) if (is_ifc[n]) plan.newpos[n] = w++;
  plan.n_local = nloc;
  plan.peer_nodes.assign(nranks, {});
  for (int r = 0; r < nranks; ++r) {
    plan.peer_nodes[r].resize(shared_with[r].size());
    for (size_t t = 0; t < shared_with[r].size(); ++t)
      plan.peer_nodes[r][t] = (int32_t)(plan.newpos[shared_with[r][t]] - nloc);
  }
  return NSB_OK;
}

// Discover which gather-scatter nodes are shared with which rank: all-gather the (sorted) global
// ids of every rank's nodes and intersect.  Both sides order a pair's nodes by global id, so the
// packed buffers line up without further negotiation.
int exchange_setup(nsb_sem_t S) {
  nsb_context_t ctx = S->ctx;
  NSB_REQUIRE(ctx->nccl_comm, "exchange_setup: no communicator");
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  cudaSetDevice(ctx->device);
  const int P = ctx->nranks;
  // 1. counts
  int64_t *cnt_d = nullptr;
  NSB_CUDA(cudaMalloc(&cnt_d, sizeof(int64_t) * (P + 1)));
  int64_t mine = S->nshared;
  NSB_CUDA(cudaMemcpyAsync(cnt_d + P, &mine, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  NSB_NCCL(g_nccl.AllGather(cnt_d + P, cnt_d, 1, ncclInt64, comm, ctx->stream));
  std::vector<int64_t> cnt(P);
  NSB_CUDA(cudaMemcpyAsync(cnt.data(), cnt_d, sizeof(int64_t) * P, cudaMemcpyDeviceToHost, ctx->stream));
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(cnt_d);
  const int64_t mx = std::max<int64_t>(1, *std::max_element(cnt.begin(), cnt.end()));
  // 2. all-gather every rank's sorted ids
  std::vector<int64_t> send(S->node_gid.begin(), S->node_gid.end());
  std::sort(send.begin(), send.end());
  send.resize(mx, -1);
  int64_t *send_d = nullptr, *all_d = nullptr;
  NSB_CUDA(cudaMalloc(&send_d, sizeof(int64_t) * mx));
  NSB_CUDA(cudaMalloc(&all_d, sizeof(int64_t) * mx * P));
  NSB_CUDA(cudaMemcpyAsync(send_d, send.data(), sizeof(int64_t) * mx, cudaMemcpyHostToDevice, ctx->stream));
  NSB_NCCL(g_nccl.AllGather(send_d, all_d, (size_t)mx, ncclInt64, comm, ctx->stream));
  std::vector<int64_t> all((size_t)mx * P);
  NSB_CUDA(cudaMemcpyAsync(all.data(), all_d, sizeof(int64_t) * mx * P, cudaMemcpyDeviceToHost, ctx->stream));
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(send_d);
  cudaFree(all_d);
  // 3. host plan: intersect, private nodes first / interface nodes last
  for (auto &Pr : S->peers) {
    cudaFree(Pr.idx_d);
    cudaFree(Pr.send_d);
    cudaFree(Pr.recv_d);
  }
  S->peers.clear();
  ExchangePlan plan;
  NSB_CHECK(exchange_plan(ctx->rank, P, S->node_gid, cnt, all.data(), mx, plan));
  const std::vector<int64_t> &newpos = plan.newpos;
  const int64_t nloc = plan.n_local;
  // 4. reorder the gather-scatter lists accordingly
  std::vector<int64_t> off2(S->nshared + 1), gid2(S->nshared), inv(S->nshared);
  for (int64_t n = 0; n < S->nshared; ++n) inv[newpos[n]] = n;
  std::vector<int32_t> idx2(S->gs_idx_h.size());
  int64_t q = 0;
  for (int64_t m = 0; m < S->nshared; ++m) {
    const int64_t n = inv[m];
    off2[m] = q;
    for (int64_t t = S->gs_off_h[n]; t < S->gs_off_h[n + 1]; ++t) idx2[q++] = S->gs_idx_h[t];
    gid2[m] = S->node_gid[n];
  }
  off2[S->nshared] = q;
  S->gs_off_h.swap(off2);
  S->gs_idx_h.swap(idx2);
  S->node_gid.swap(gid2);
  S->n_local = nloc;
  NSB_CUDA(cudaMemcpy(S->gs_off_d, S->gs_off_h.data(), sizeof(int64_t) * (S->nshared + 1), cudaMemcpyHostToDevice));
  NSB_CUDA(cudaMemcpy(S->gs_idx_d, S->gs_idx_h.data(), sizeof(int32_t) * S->gs_idx_h.size(), cudaMemcpyHostToDevice));
  const int64_t nifc = S->nshared - nloc;
  if (S->node_sum_d) cudaFree(S->node_sum_d);
  S->node_sum_d = nullptr;
  NSB_CUDA(cudaMalloc(&S->node_sum_d, sizeof(double) * S->ns_fields * (nifc > 0 ? nifc : 1)));
  for (int r = 0; r < P; ++r) {
    const std::vector<int32_t> &nodes = plan.peer_nodes[r];
    if (nodes.empty()) continue;
    nsb_sem_s::Peer Pr;
    Pr.rank = r;
    Pr.n = (int64_t)nodes.size();
    NSB_CUDA(cudaMalloc(&Pr.idx_d, sizeof(int32_t) * Pr.n));
    NSB_CUDA(cudaMalloc(&Pr.send_d, sizeof(double) * Pr.n * S->ns_fields));
    NSB_CUDA(cudaMalloc(&Pr.recv_d, sizeof(double) * Pr.n * S->ns_fields));
    NSB_CUDA(cudaMemcpy(Pr.idx_d, nodes.data(), sizeof(int32_t) * Pr.n, cudaMemcpyHostToDevice));
    S->peers.push_back(Pr);
  }
  // 5. peer-memory halo: lay out one region per peer in MY mailbox and tell every peer where its
  //    data goes (all-gather of the offset tables)
  S->p2p_halo = false;
  if (ctx->p2p) {
    std::vector<int64_t> mine_off(P, -1);
    int64_t cur = 0;
    for (auto &Pr : S->peers) {
      Pr.my_off = cur;
      mine_off[Pr.rank] = cur;
      cur += 2 * (int64_t)S->ns_fields * Pr.n;           // two slots
      cur = (cur + 31) & ~(int64_t)31;
    }
    int64_t fits = (mb_halo_base(P) + (size_t)cur) * sizeof(double) <= ctx->mail_bytes ? 1 : 0;
    int64_t *tab_d = nullptr;
    NSB_CUDA(cudaMalloc(&tab_d, sizeof(int64_t) * (size_t)(P + 1) * (P + 1)));
    std::vector<int64_t> sendt(mine_off);
    sendt.push_back(fits);
    NSB_CUDA(cudaMemcpyAsync(tab_d + (size_t)P * (P + 1), sendt.data(), sizeof(int64_t) * (P + 1), cudaMemcpyHostToDevice,
                             ctx->stream));
    NSB_NCCL(g_nccl.AllGather(tab_d + (size_t)P * (P + 1), tab_d, (size_t)(P + 1), ncclInt64, comm, ctx->stream));
    std::vector<int64_t> tab((size_t)P * (P + 1));
    NSB_CUDA(cudaMemcpyAsync(tab.data(), tab_d, sizeof(int64_t) * tab.size(), cudaMemcpyDeviceToHost, ctx->stream));
    NSB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(tab_d);
    bool all_fit = true;
    for (int r = 0; r < P; ++r) all_fit = all_fit && tab[(size_t)r * (P + 1) + P] == 1;
    if (all_fit) {
      for (auto &Pr : S->peers) Pr.peer_off = tab[(size_t)Pr.rank * (P + 1) + ctx->rank];
      S->p2p_halo = true;
    }
  }
  return NSB_OK;
}

}  // namespace nsb

// Host-only (no CUDA / NCCL): the exchange plan for CPU tests of the multi-rank logic.
//   gid[nnodes]            this rank's node ids in gather-scatter order
//   counts[nranks]         nodes per rank;  all_sorted[nranks][mx] every rank's ids, ascending
//   newpos[nnodes]         out: position of each node after the private/interface reorder
//   peer_count[nranks]     out: nodes shared with each rank
//   peer_nodes[nranks][mx] out: interface-relative node indices per rank, ascending global id
extern "C" int nsb_host_exchange_plan(int rank, int nranks, int64_t nnodes, const int64_t *gid,
                                      const int64_t *counts, const int64_t *all_sorted, int64_t mx,
                                      int64_t *newpos, int64_t *n_local, int64_t *peer_count,
                                      int32_t *peer_nodes) {
  NSB_REQUIRE(gid && counts && all_sorted && newpos && n_local && peer_count && peer_nodes && nranks >= 1 &&
                  rank >= 0 && rank < nranks && nnodes >= 0 && mx >= 1,
              "nsb_host_exchange_plan: bad argument");
  std::vector<int64_t> g(gid, gid + nnodes), c(counts, counts + nranks);
  nsb::ExchangePlan plan;
  NSB_CHECK(nsb::exchange_plan(rank, nranks, g, c, all_sorted, mx, plan));
  memcpy(newpos, plan.newpos.data(), sizeof(int64_t) * nnodes);
  *n_local = plan.n_local;
  for (int r = 0; r < nranks; ++r) {
    peer_count[r] = (int64_t)plan.peer_nodes[r].size();
    memcpy(peer_nodes + (size_t)r * mx, plan.peer_nodes[r].data(), sizeof(int32_t) * plan.peer_nodes[r].size());
  }
  return NSB_OK;
}

// Peer-memory mailbox: create + export (64-byte CUDA IPC handle), then map every rank's mailbox.
extern "C" int nsb_p2p_mailbox_create(nsb_context_t ctx, int64_t halo_bytes, void *handle_out) {
  NSB_REQUIRE(ctx && handle_out && halo_bytes >= 0, "nsb_p2p_mailbox_create: bad argument");
  NSB_REQUIRE(ctx->nranks <= nsb_context_s::kMaxPeers, "nsb_p2p_mailbox_create: more than %d ranks",
              nsb_context_s::kMaxPeers);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaSetDevice(ctx->device);
  if (ctx->mail_d) cudaFree(ctx->mail_d);
  ctx->halo_bytes = (size_t)halo_bytes;
  ctx->mail_bytes = nsb::mb_halo_base(ctx->nranks) * sizeof(double) + (size_t)halo_bytes;
  NSB_CUDA(cudaMalloc(&ctx->mail_d, ctx->mail_bytes));
  NSB_CUDA(cudaMemset(ctx->mail_d, 0, ctx->mail_bytes));
  cudaIpcMemHandle_t h;
  NSB_CUDA(cudaIpcGetMemHandle(&h, ctx->mail_d));
  memcpy(handle_out, &h, sizeof(h));
  return NSB_OK;
}

extern "C" int nsb_p2p_mailbox_connect(nsb_context_t ctx, const void *all_handles) {
  NSB_REQUIRE(ctx && all_handles && ctx->mail_d, "nsb_p2p_mailbox_connect: create the mailbox first");
  cudaSetDevice(ctx->device);
  for (int r = 0; r < ctx->nranks; ++r) {
    if (r == ctx->rank) {
      ctx->peer_mail[r] = ctx->mail_d;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)all_handles + (size_t)r * 64, 64);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      nsb::set_error("nsb_p2p_mailbox_connect: cannot map the mailbox of rank %d: %s", r, cudaGetErrorString(e));
      return NSB_ECUDA;
    }
    ctx->peer_mail[r] = (double *)p;
  }
  ctx->p2p = true;
  return NSB_OK;
}

extern "C" int nsb_p2p_enabled(nsb_context_t ctx, int *allreduce_p2p) {
  NSB_REQUIRE(ctx && allreduce_p2p, "nsb_p2p_enabled: NULL argument");
  *allreduce_p2p = ctx->p2p ? 1 : 0;
  return NSB_OK;
}

extern "C" int nsb_get_unique_id(void *id_out) {
  NSB_REQUIRE(id_out, "nsb_get_unique_id: NULL");
  NSB_CHECK(nsb::load_nccl());
  ncclUniqueId id;
  ncclResult_t r = nsb::g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) {
    nsb::set_error("ncclGetUniqueId: %s", nsb::g_nccl.GetErrorString(r));
    return NSB_ENCCL;
  }
  memset(id_out, 0, NSB_UNIQUE_ID_BYTES);
  memcpy(id_out, &id, sizeof(id));
  return NSB_OK;
}
