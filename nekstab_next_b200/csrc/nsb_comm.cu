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
#include <cstdio>
#include <cstring>

#include "nsb_internal.h"
#include "nsb_tail.cuh"

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
// (mailbox layout, flag loads / stores and the block-level all-reduce live in nsb_tail.cuh: the sweep kernels
// of the orthogonalisation run the same exchange in their own tails)

// stand-alone all-reduce of n <= ARN doubles for the callers without a tail (single dots, CG scalars)
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(double *buf, int n, PeerComm comm) {
  p2p_allreduce_block(buf, n, comm);
}

// ---- halo exchange of the gather-scatter through the mailboxes ---------------------------------
// Every mesh (nsb_sem_t) owns a reservation in the halo area of the mailbox: its flag words [2][P] and one
// data region per neighbour (two slots each), plus its own device-resident sequence number -- two meshes on
// one context (velocity and pressure mesh) neither overlap nor share a sequence.
//   pack   : interface sums -> the neighbour's region (peer stores), slot = (seq + 1) & 1
//   flags  : seq += 1; release-store seq into every neighbour's flag word for this rank
//   unpack : acquire-spin on the neighbour's flag word in MY mailbox, then add its data
__global__ void pack_p2p_kernel(const double *__restrict__ node_sum, const int32_t *__restrict__ nodes, int64_t n,
                                double *__restrict__ dst0, int64_t slot_stride, int64_t ns_stride,
                                const unsigned long long *__restrict__ seq_d) {
  const int slot = (int)((*seq_d + 1ull) & 1ull);
  double *dst = dst0 + (int64_t)slot * slot_stride;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (t < n) dst[(int64_t)f * n + t] = node_sum[(int64_t)f * ns_stride + nodes[t]];
}

struct HaloFlags { uint64_t *p[nsb_context_s::kMaxPeers]; int n; int P; };   // p[i]: slot-0 flag word at peer i
__global__ void p2p_flags_kernel(HaloFlags fl, unsigned long long *seq_d) {
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) s_seq = ++(*seq_d);
  __syncthreads();
  const uint64_t seq = s_seq;
  if ((int)threadIdx.x < fl.n) st_release_sys(fl.p[threadIdx.x] + (seq & 1) * fl.P, seq);
}

__global__ void unpack_wait_add_kernel(double *__restrict__ node_sum, const int32_t *__restrict__ nodes, int64_t n,
                                       const double *src0, int64_t slot_stride, int64_t ns_stride,
                                       const uint64_t *flag0, int P, const unsigned long long *__restrict__ seq_d,
                                       int *err) {
  const uint64_t seq = *seq_d;
  const int slot = (int)(seq & 1);
  if (threadIdx.x == 0) spin_until(flag0 + (size_t)slot * P, seq, err, DEVERR_HALO_TIMEOUT);
  __syncthreads();
  const double *src = src0 + (int64_t)slot * slot_stride;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (t < n) node_sum[(int64_t)f * ns_stride + nodes[t]] += ld_relaxed_sys(src + (int64_t)f * n + t);
}

int halo_exchange_p2p(nsb_sem_t S, int nf, cudaStream_t st) {
  nsb_context_t ctx = S->ctx;
  const int P = ctx->nranks;
  const int64_t nifc = S->nshared - S->n_local;
  HaloFlags fl;
  fl.n = 0;
  fl.P = P;
  for (auto &Pr : S->peers) {
    const unsigned nb = (unsigned)((Pr.n + 255) / 256);
    double *dst = ctx->peer_mail[Pr.rank] + mb_halo_base(P) + Pr.peer_off;
    pack_p2p_kernel<<<dim3(nb, nf), 256, 0, st>>>(S->node_sum_d, Pr.idx_d, Pr.n, dst, (int64_t)S->ns_fields * Pr.n, nifc,
                                                  S->hx_seq_d);
    fl.p[fl.n++] = reinterpret_cast<uint64_t *>(ctx->peer_mail[Pr.rank] + mb_halo_base(P) + S->peer_flag_off[Pr.rank]) +
                   ctx->rank;
    ctx->launches++;
  }
  // stores of the pack kernels are complete at the kernel boundary; then publish the sequence number
  p2p_flags_kernel<<<1, 32, 0, st>>>(fl, S->hx_seq_d);
  ctx->launches++;
  for (auto &Pr : S->peers) {
    const unsigned nb = (unsigned)((Pr.n + 255) / 256);
    const double *src = ctx->mail_d + mb_halo_base(P) + Pr.my_off;
    const uint64_t *flag = reinterpret_cast<const uint64_t *>(ctx->mail_d + mb_halo_base(P) + S->halo_flag_off) + Pr.rank;
    unpack_wait_add_kernel<<<dim3(nb, nf), 256, 0, st>>>(S->node_sum_d, Pr.idx_d, Pr.n, src, (int64_t)S->ns_fields * Pr.n,
                                                         nifc, flag, P, S->hx_seq_d, ctx->dev_err_d);
    ctx->launches++;
  }
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}


// ---- fused interface path of the gather-scatter (peer-memory transport) --------------------------
// send : one thread per interface node -- sum its local copies (all nf fields), keep the sum in node_sum and
//        store it straight into the mailbox region of every peer sharing the node; the CTA drawing the last
//        ticket bumps the mesh's sequence number and release-stores it into the peers' flag words.
// recv : thread 0 of every CTA waits for all neighbours' flags (bounded), then one thread per interface node adds
//        the neighbours' contributions in ascending rank order (bitwise the same on both sides of an interface)
//        and scatters the result (plain dssum or the fused operator tail alpha u + beta bmask sum).
// Two launches instead of 3 + 2 P (sum, pack x P, flags, wait+unpack x P, scatter).
struct HaloPeers {
  double *dst[nsb_context_s::kMaxPeers];            // my region in the peer's mailbox (slot 0)
  const double *src[nsb_context_s::kMaxPeers];      // the peer's region in my mailbox (slot 0)
  uint64_t *flag_out[nsb_context_s::kMaxPeers];     // my flag word in the peer's mailbox (slot 0)
  const uint64_t *flag_in[nsb_context_s::kMaxPeers];
  int64_t n[nsb_context_s::kMaxPeers];              // nodes shared with the peer
  int npeers, P, ns_fields;
};

__global__ void __launch_bounds__(256)
halo_send_kernel(const double *__restrict__ v, const int64_t *__restrict__ off, const int32_t *__restrict__ idx,
                 int64_t n0, int64_t nifc, int nf, int64_t fstride, double *__restrict__ node_sum,
                 const int32_t *__restrict__ poff, const int64_t *__restrict__ pent, HaloPeers hp,
                 unsigned long long *seq_d, unsigned int *ticket) {
  const int slot = (int)((*seq_d + 1ull) & 1ull);
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < nifc) {
    const int64_t a = off[n0 + m], b = off[n0 + m + 1];
    const int pa = poff[m], pb = poff[m + 1];
    for (int f = 0; f < nf; ++f) {
      const double *vf = v + (int64_t)f * fstride;
      double sum = 0.0;
      for (int64_t q = a; q < b; ++q) sum += vf[idx[q]];
      node_sum[(int64_t)f * nifc + m] = sum;
      for (int e = pa; e < pb; ++e) {
        const int64_t ent = pent[e];
        const int pi = (int)(ent >> 32);
        const int64_t pos = ent & 0xffffffffll;
        hp.dst[pi][(int64_t)slot * hp.ns_fields * hp.n[pi] + (int64_t)f * hp.n[pi] + pos] = sum;   // peer store
      }
    }
  }
  __shared__ unsigned int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) {
    s_seq = ++(*seq_d);
    *ticket = 0;
  }
  __syncthreads();
  if ((int)threadIdx.x < hp.npeers) st_release_sys(hp.flag_out[threadIdx.x] + (s_seq & 1) * hp.P, s_seq);
}

// 3: v <- sum ; 4: v <- alpha uin + beta bnode sum ; 5 (C0 layout): wu[node] <- alpha uu[node] + beta bnode sum
template <int EPI>
__global__ void __launch_bounds__(256)
halo_recv_kernel(double *__restrict__ v, const int64_t *__restrict__ off, const int32_t *__restrict__ idx, int64_t n0,
                 int64_t nifc, int nf, int64_t fstride, const double *__restrict__ node_sum,
                 const int32_t *__restrict__ poff, const int64_t *__restrict__ pent, HaloPeers hp,
                 const unsigned long long *__restrict__ seq_d, int *err, const double *__restrict__ uin, double alpha,
                 double beta, const double *__restrict__ bnode, const double *__restrict__ bmask) {
  const uint64_t seq = *seq_d;
  const int slot = (int)(seq & 1);
  if ((int)threadIdx.x < hp.npeers) spin_until(hp.flag_in[threadIdx.x] + (size_t)slot * hp.P, seq, err, DEVERR_HALO_TIMEOUT);
  __syncthreads();
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nifc) return;
  const int64_t a = off[n0 + m], b = off[n0 + m + 1];
  const int pa = poff[m], pb = poff[m + 1];
  for (int f = 0; f < nf; ++f) {
    double sum = node_sum[(int64_t)f * nifc + m];
    for (int e = pa; e < pb; ++e) {                 // entries are stored in ascending peer rank
      const int64_t ent = pent[e];
      const int pi = (int)(ent >> 32);
      const int64_t pos = ent & 0xffffffffll;
      sum += ld_relaxed_sys(hp.src[pi] + (int64_t)slot * hp.ns_fields * hp.n[pi] + (int64_t)f * hp.n[pi] + pos);
    }
    if (EPI == 5) {       // v / uin point at the first interface node of the unique layout, fstride = its field stride
      v[(int64_t)f * fstride + m] = alpha * uin[(int64_t)f * fstride + m] + beta * bnode[n0 + m] * sum;
      continue;
    }
    double *vf = v + (int64_t)f * fstride;
    for (int64_t q = a; q < b; ++q) {
      const int32_t p = idx[q];
      if (EPI == 3) vf[p] = sum;
      else vf[p] = alpha * uin[(int64_t)f * fstride + p] + beta * (bnode ? bnode[n0 + m] : bmask[p]) * sum;
    }
  }
}

static void fill_halo_peers(nsb_sem_t S, HaloPeers &hp);

// C0 layout: the copies are read from the element-local scratch, the assembled interface nodes are written to the
// unique layout (wu / uu point at boundary node 0 of field 0)
int halo_exchange_fused_c0(nsb_sem_t S, const double *wloc, int nf, int64_t fs_loc, double *wu, const double *uu,
                           int64_t fs_u, double alpha, double beta, cudaStream_t st) {
  nsb_context_t ctx = S->ctx;
  const int64_t nifc = S->nshared - S->n_local;
  HaloPeers hp;
  fill_halo_peers(S, hp);
  const unsigned nb = (unsigned)((nifc + 255) / 256);
  halo_send_kernel<<<nb, 256, 0, st>>>(wloc, S->gs_off_d, S->gs_idx_d, S->n_local, nifc, nf, fs_loc, S->node_sum_d,
                                      S->ifc_poff_d, S->ifc_pent_d, hp, S->hx_seq_d, S->hx_ticket_d);
  halo_recv_kernel<5><<<nb, 256, 0, st>>>(wu + S->n_local, S->gs_off_d, S->gs_idx_d, S->n_local, nifc, nf, fs_u, S->node_sum_d,
                                         S->ifc_poff_d, S->ifc_pent_d, hp, S->hx_seq_d, ctx->dev_err_d, uu + S->n_local, alpha,
                                         beta, S->bnode_d, nullptr);
  ctx->launches += 2;
  NSB_CUDA(cudaGetLastError());
  return NSB_OK;
}

static void fill_halo_peers(nsb_sem_t S, HaloPeers &hp) {
  nsb_context_t ctx = S->ctx;
  const int P = ctx->nranks;
  hp.npeers = (int)S->peers.size();
  hp.P = P;
  hp.ns_fields = S->ns_fields;
  for (int i = 0; i < hp.npeers; ++i) {
    const auto &Pr = S->peers[i];
    hp.dst[i] = ctx->peer_mail[Pr.rank] + mb_halo_base(P) + Pr.peer_off;
    hp.src[i] = ctx->mail_d + mb_halo_base(P) + Pr.my_off;
    hp.flag_out[i] = reinterpret_cast<uint64_t *>(ctx->peer_mail[Pr.rank] + mb_halo_base(P) + S->peer_flag_off[Pr.rank]) + ctx->rank;
    hp.flag_in[i] = reinterpret_cast<const uint64_t *>(ctx->mail_d + mb_halo_base(P) + S->halo_flag_off) + Pr.rank;
    hp.n[i] = Pr.n;
  }
}

int halo_exchange_fused(nsb_sem_t S, double *v, int nf, int64_t fstride, int epi, const double *uin, double alpha,
                        double beta, const double *bmask, cudaStream_t st) {
  nsb_context_t ctx = S->ctx;
  const int P = ctx->nranks;
  const int64_t nifc = S->nshared - S->n_local;
  HaloPeers hp;
  hp.npeers = (int)S->peers.size();
  hp.P = P;
  hp.ns_fields = S->ns_fields;
  for (int i = 0; i < hp.npeers; ++i) {
    const auto &Pr = S->peers[i];
    hp.dst[i] = ctx->peer_mail[Pr.rank] + mb_halo_base(P) + Pr.peer_off;
    hp.src[i] = ctx->mail_d + mb_halo_base(P) + Pr.my_off;
    hp.flag_out[i] = reinterpret_cast<uint64_t *>(ctx->peer_mail[Pr.rank] + mb_halo_base(P) + S->peer_flag_off[Pr.rank]) + ctx->rank;
    hp.flag_in[i] = reinterpret_cast<const uint64_t *>(ctx->mail_d + mb_halo_base(P) + S->halo_flag_off) + Pr.rank;
    hp.n[i] = Pr.n;
  }
  const unsigned nb = (unsigned)((nifc + 255) / 256);
  halo_send_kernel<<<nb, 256, 0, st>>>(v, S->gs_off_d, S->gs_idx_d, S->n_local, nifc, nf, fstride, S->node_sum_d,
                                      S->ifc_poff_d, S->ifc_pent_d, hp, S->hx_seq_d, S->hx_ticket_d);
  const double *bnode = (bmask && bmask == S->bmask_d) ? S->bnode_d : nullptr;
  if (epi == 0)
    halo_recv_kernel<3><<<nb, 256, 0, st>>>(v, S->gs_off_d, S->gs_idx_d, S->n_local, nifc, nf, fstride, S->node_sum_d,
                                           S->ifc_poff_d, S->ifc_pent_d, hp, S->hx_seq_d, ctx->dev_err_d, nullptr, 0.0,
                                           0.0, nullptr, nullptr);
  else
    halo_recv_kernel<4><<<nb, 256, 0, st>>>(v, S->gs_off_d, S->gs_idx_d, S->n_local, nifc, nf, fstride, S->node_sum_d,
                                           S->ifc_poff_d, S->ifc_pent_d, hp, S->hx_seq_d, ctx->dev_err_d, uin, alpha, beta,
                                           bnode, bmask);
  ctx->launches += 2;
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
    PeerComm comm;
    comm.P = ctx->nranks;
    comm.rank = ctx->rank;
    comm.seq = ctx->seq_d;
    comm.err = ctx->dev_err_d;
    for (int r = 0; r < nsb_context_s::kMaxPeers; ++r) comm.mail.p[r] = r < ctx->nranks ? ctx->peer_mail[r] : nullptr;
    ProfScope ps(ctx, PC_SMALL, 8.0 * n * ctx->nranks);
    p2p_allreduce_kernel<<<1, 256, 0, ctx->stream>>>(buf_d, n, comm);
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
  std::vector<int32_t> slab2(S->node_slab.size());
  for (int64_t n = 0; n < S->nshared; ++n) inv[newpos[n]] = n;
  std::vector<int32_t> idx2(S->gs_idx_h.size());
  int64_t q = 0;
  for (int64_t m = 0; m < S->nshared; ++m) {
    const int64_t n = inv[m];
    off2[m] = q;
    for (int64_t t = S->gs_off_h[n]; t < S->gs_off_h[n + 1]; ++t) idx2[q++] = S->gs_idx_h[t];
    gid2[m] = S->node_gid[n];
    if (!slab2.empty()) slab2[m] = S->node_slab[n];
  }
  off2[S->nshared] = q;
  S->gs_off_h.swap(off2);
  S->gs_idx_h.swap(idx2);
  S->node_gid.swap(gid2);
  S->node_slab.swap(slab2);
  S->n_local = nloc;
  compute_slab_ends(S);
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
  // per interface node: the peers sharing it and its position in each peer's packed list (ascending peer rank),
  // for the fused send / recv kernels
  {
    std::vector<std::vector<int64_t>> ent((size_t)(nifc > 0 ? nifc : 0));
    for (size_t pi = 0; pi < S->peers.size(); ++pi) {
      const std::vector<int32_t> &nodes = plan.peer_nodes[S->peers[pi].rank];
      for (size_t t = 0; t < nodes.size(); ++t) ent[nodes[t]].push_back(((int64_t)pi << 32) | (int64_t)t);
    }
    std::vector<int32_t> poff((size_t)nifc + 1, 0);
    std::vector<int64_t> pent;
    for (int64_t m = 0; m < nifc; ++m) {
      poff[m] = (int32_t)pent.size();
      pent.insert(pent.end(), ent[m].begin(), ent[m].end());
    }
    poff[nifc] = (int32_t)pent.size();
    if (S->ifc_poff_d) cudaFree(S->ifc_poff_d);
    if (S->ifc_pent_d) cudaFree(S->ifc_pent_d);
    S->ifc_poff_d = nullptr;
    S->ifc_pent_d = nullptr;
    NSB_CUDA(cudaMalloc(&S->ifc_poff_d, sizeof(int32_t) * poff.size()));
    NSB_CUDA(cudaMalloc(&S->ifc_pent_d, sizeof(int64_t) * (pent.empty() ? 1 : pent.size())));
    NSB_CUDA(cudaMemcpy(S->ifc_poff_d, poff.data(), sizeof(int32_t) * poff.size(), cudaMemcpyHostToDevice));
    if (!pent.empty())
      NSB_CUDA(cudaMemcpy(S->ifc_pent_d, pent.data(), sizeof(int64_t) * pent.size(), cudaMemcpyHostToDevice));
    if (!S->hx_ticket_d) {
      NSB_CUDA(cudaMalloc(&S->hx_ticket_d, sizeof(unsigned int)));
      NSB_CUDA(cudaMemset(S->hx_ticket_d, 0, sizeof(unsigned int)));
    }
  }
  // 5. peer-memory halo: reserve this mesh's flag words and one data region per peer in MY mailbox (bump
  //    allocator over the halo area, so several meshes on one context never overlap) and tell every peer
  //    where its data and its flag go (all-gather of the offset tables)
  S->p2p_halo = false;
  if (ctx->p2p) {
    const int64_t flag_doubles = ((int64_t)2 * P + 31) & ~(int64_t)31;
    int64_t need = flag_doubles;
    for (auto &Pr : S->peers) need += ((2 * (int64_t)S->ns_fields * Pr.n) + 31) & ~(int64_t)31;
    const bool reuse = S->halo_flag_off >= 0 && need <= S->halo_region_doubles;   // repeated set-up of this mesh
    const int64_t base = reuse ? S->halo_flag_off : (int64_t)ctx->halo_used;
    std::vector<int64_t> mine_off(P, -1);
    int64_t cur = base + flag_doubles;
    for (auto &Pr : S->peers) {
      Pr.my_off = cur;
      mine_off[Pr.rank] = cur;
      cur += ((2 * (int64_t)S->ns_fields * Pr.n) + 31) & ~(int64_t)31;   // two slots
    }
    const int64_t fits = (mb_halo_base(P) + (size_t)cur) * sizeof(double) <= ctx->mail_bytes ? 1 : 0;
    if (!S->hx_seq_d) NSB_CUDA(cudaMalloc(&S->hx_seq_d, sizeof(unsigned long long)));
    NSB_CUDA(cudaMemsetAsync(S->hx_seq_d, 0, sizeof(unsigned long long), ctx->stream));
    // flags (and data) of the reservation start from zero BEFORE any peer can learn the offsets
    if (fits) NSB_CUDA(cudaMemsetAsync(ctx->mail_d + mb_halo_base(P) + base, 0, sizeof(double) * (size_t)(cur - base), ctx->stream));
    const int W = P + 2;   // per rank: P data offsets, the flag offset, fits
    int64_t *tab_d = nullptr;
    NSB_CUDA(cudaMalloc(&tab_d, sizeof(int64_t) * (size_t)(P + 1) * W));
    std::vector<int64_t> sendt(mine_off);
    sendt.push_back(base);
    sendt.push_back(fits);
    NSB_CUDA(cudaMemcpyAsync(tab_d + (size_t)P * W, sendt.data(), sizeof(int64_t) * W, cudaMemcpyHostToDevice, ctx->stream));
    NSB_NCCL(g_nccl.AllGather(tab_d + (size_t)P * W, tab_d, (size_t)W, ncclInt64, comm, ctx->stream));
    std::vector<int64_t> tab((size_t)P * W);
    NSB_CUDA(cudaMemcpyAsync(tab.data(), tab_d, sizeof(int64_t) * tab.size(), cudaMemcpyDeviceToHost, ctx->stream));
    NSB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(tab_d);
    bool all_fit = true;
    for (int r = 0; r < P; ++r) all_fit = all_fit && tab[(size_t)r * W + P + 1] == 1;
    if (all_fit) {
      for (auto &Pr : S->peers) Pr.peer_off = tab[(size_t)Pr.rank * W + ctx->rank];
      S->peer_flag_off.assign(P, -1);
      for (int r = 0; r < P; ++r) S->peer_flag_off[r] = tab[(size_t)r * W + P];
      if (!reuse) {
        if (S->halo_flag_off < 0) ctx->halo_users++;
        S->halo_flag_off = base;
        S->halo_region_doubles = cur - base;
        ctx->halo_used = (size_t)cur;
      }
      S->p2p_halo = true;
    }
  }
  S->c0_l2u_nshared = -1;   // the node order changed: the C0 index map is rebuilt on next use
  clear_step_graphs(ctx);
  return NSB_OK;
}

// MIN over all ranks of a host flag (set-up only, NCCL)
int allreduce_min_int(nsb_context_t ctx, int mine, int *out) {
  *out = mine;
  if (ctx->nranks == 1) return NSB_OK;
  NSB_REQUIRE(ctx->nccl_comm, "allreduce_min_int: no communicator");
  int *v_d = nullptr;
  NSB_CUDA(cudaMalloc(&v_d, sizeof(int)));
  NSB_CUDA(cudaMemcpyAsync(v_d, &mine, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NSB_NCCL(g_nccl.AllReduce(v_d, v_d, 1, ncclInt32, ncclMin, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  NSB_CUDA(cudaMemcpyAsync(out, v_d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  NSB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(v_d);
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
  int ok = 1;
  char why[256] = "";
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
      snprintf(why, sizeof(why), "cannot map the mailbox of rank %d: %s", r, cudaGetErrorString(e));
      cudaGetLastError();
      ok = 0;
      break;
    }
    ctx->peer_mail[r] = (double *)p;
  }
  // The peer-memory path is used only if EVERY rank mapped EVERY mailbox: a rank left on NCCL while the others
  // spin on its flag would hang them.  The verdict is a MIN all-reduce over the NCCL communicator, so all
  // ranks also agree on the number of collectives issued during set-up.
  int all_ok = ok;
  NSB_CHECK(nsb::allreduce_min_int(ctx, ok, &all_ok));
  if (!all_ok) {
    for (int r = 0; r < ctx->nranks; ++r) {
      if (r != ctx->rank && ctx->peer_mail[r]) cudaIpcCloseMemHandle(ctx->peer_mail[r]);
      ctx->peer_mail[r] = nullptr;
    }
    ctx->p2p = false;
    if (!ok) nsb::set_error("nsb_p2p_mailbox_connect: %s; all ranks stay on NCCL", why);
    return NSB_OK;   // not an error: NCCL remains the transport on every rank (query nsb_p2p_enabled)
  }
  ctx->p2p = true;
  ctx->halo_used = 0;
  ctx->halo_users = 0;
  nsb::clear_step_graphs(ctx);
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
