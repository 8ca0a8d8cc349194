// Internal structures shared by the translation units of libnekstab_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "nekstab_b200.h"

namespace nsb {

void set_error(const char *fmt, ...);

#define NSB_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      nsb::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return NSB_ECUDA;                                                                  \
    }                                                                                    \
  } while (0)

#define NSB_CHECK(call)      \
  do {                       \
    int r_ = (call);         \
    if (r_ != NSB_OK) return r_; \
  } while (0)

#define NSB_REQUIRE(cond, ...)     \
  do {                             \
    if (!(cond)) {                 \
      nsb::set_error(__VA_ARGS__); \
      return NSB_EINVAL;           \
    }                              \
  } while (0)

constexpr int kRowPad = 1024;   // every region of a column is padded to this many rows
constexpr int kMaxK = 1024;     // largest number of columns one orthogonalisation handles
constexpr int kNumSM = 148;     // B200

struct Nccl;  // dlopen'ed NCCL entry points (nsb_comm.cu)

}  // namespace nsb

struct nsb_context_s {
  int device = 0, rank = 0, nranks = 1;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t gs_stream = nullptr;    // gather-scatter of finished slabs, concurrent with the next slab's axhelm
  double ax_slab_mb = 0.0;             // NSB_AX_SLAB_MB: u + w of one slab (3 fields) in MB; 0 (default) = no slab pipeline
                                       // (measured slower: profiles/ncu_r02_matvec_slabs.md)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t launches = 0;
  int num_sms = nsb::kNumSM;
  void *nccl_comm = nullptr;  // ncclComm_t
  // scratch for reductions: partial sums [max_ctas][kMaxK+1], results, flags
  double *partial_d = nullptr;
  int64_t partial_rows = 0;
  double *hvec_d = nullptr;    // 4 x (kMaxK+8): h1, h2, hsum, scalars
  double *hpin = nullptr;      // pinned mirror
  double *flush_d = nullptr;
  size_t flush_bytes = 0;
  // NVLink peer-memory mailbox (nsb_comm.cu): one-shot all-reduce and halo exchange written
  // straight into the peers' memory instead of NCCL launches
  static constexpr int kMaxPeers = 16;
  bool p2p = false;
  double *mail_d = nullptr;            // own mailbox (cudaMalloc, IPC-exported)
  size_t mail_bytes = 0, halo_bytes = 0;
  double *peer_mail[kMaxPeers] = {};   // every rank's mailbox mapped here ([rank] = own)
  size_t halo_used = 0;                // bump allocator over the halo area (doubles); one region set per mesh
  int halo_users = 0;                  // meshes holding a region; the allocator rewinds when the last one goes
  // device-side state of the kernel tails (nsb_tail.cuh): last-CTA tickets, the DGKS second-pass flag, the
  // sequence number of the peer-memory all-reduces (device-resident so that a step replays as a CUDA graph)
  unsigned int *ticket_d = nullptr;    // [8], zero between launches
  int *flag_d = nullptr;               // [4]
  unsigned long long *seq_d = nullptr; // [2]
  int *dev_err = nullptr;              // mapped pinned word the kernels set on a spin timeout (host address)
  int *dev_err_d = nullptr;            // its device address
  // whole Arnoldi steps captured as CUDA graphs (nsb_krylov.cu), keyed by (basis, operator, step, mode)
  bool use_graph = true;               // NSB_GRAPH=0: plain launches
  struct StepGraph { cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
  std::map<std::tuple<const void *, const void *, int, int>, StepGraph> step_graphs;
  double *hstage = nullptr;            // pinned staging of the H columns of one factorisation
  size_t hstage_elems = 0;
  // pipelined host upload (nsb_orth.cu): h1 computed chunk by chunk while the vector arrives
  std::vector<cudaEvent_t> chunk_ev;
  int h1_ready_k = -1;
  const double *h1_ready_col = nullptr;
  bool pipeline_upload = true;   // NSB_PIPELINE_UPLOAD=0: plain upload, then the usual first sweep
  // per-kernel-class device timing (bench / roofline): events around every launch when enabled
  bool no_fused = false;       // NSB_NO_FUSED=1: CGS2 with separate update / multidot kernels
  bool ns_generic = false;     // NSB_NS_GENERIC=1: generic opdiv / opgradt kernels, multi-launch coarse solve (A/B numbers)
  bool ns_no_coarse = false;   // NSB_NS_NO_COARSE=1: pressure preconditioner without the coarse level (A/B numbers)
  bool fold_norm = true;       // NSB_FOLD_NORM=0: explicit norm reduction in the third sweep + normalize_kernel
  int fused_loader = 3;        // NSB_FUSED_LOADER: 1 cp.async, 3 TMA 2-D tensor loads (default)
  int fused_rc = 0;            // NSB_FUSED_RC: force rows per block (tuning)
  int fused_reg_min_k = 54;    // NSB_FUSED_REG_MIN_K: smallest k for the register-retention variant
  bool ax_generic = false;     // NSB_AX_GENERIC=1: use the generic-order axhelm kernel for N = 7 too
  bool rotate_simple = false;  // NSB_ROTATE_SIMPLE=1: first (untiled) rotation kernel
  bool rotate_dmma = true;     // NSB_ROTATE_DMMA=0: register-tiled FMA rotation instead of the fp64 tensor-core kernel
  double *gram_d = nullptr;    // per-CTA tile sums of nsb_basis_gram
  size_t gram_elems = 0;
  double *rot_d = nullptr;     // Z and the saved %time row of nsb_basis_rotate
  size_t rot_elems = 0;
  int ax_stages = 0;           // NSB_AX_STAGES: ring depth of the N = 7 axhelm kernels (0: default; DMMA: = warp groups pins one buffer per group)
  bool fused_priv = true;      // NSB_FUSED_PRIV=0: per-block warp reduction in the fused kernel's second projection
  bool fused_allwarps = true;  // NSB_FUSED_ALLWARPS=0: warp 0 alone combines the row sums (two barriers per block)
  double dgks_eta2 = 0.5;      // DGKS: second projection when |w'|^2 < eta^2 |w|^2 (nsb_set_dgks_eta; default eta = 1/sqrt 2)
  bool halo_fused = true;      // NSB_HALO_FUSED=0: sum / pack / flags / unpack / scatter as separate launches
  bool tail = true;            // NSB_TAIL=0: separate reduce / all-reduce / add launches (round-1 structure)
  bool ax_dmma = true;         // NSB_AX_DMMA=0: vector-FMA contraction in the ring kernel instead of DMMA
  bool ax_ring = true;         // NSB_AX_RING=0: warp-per-element kernel instead of the TMA ring (N = 7)
  bool prof = false;
  struct ProfRec { int cls; cudaEvent_t e0, e1; double bytes; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
};

struct nsb_layout_s {
  nsb_context_t ctx = nullptr;
  int nfields = 0;
  std::vector<int64_t> len;      // stored length of each field (rows in a column)
  std::vector<int64_t> hlen;     // length of the host-side array of each field (differs for C0 fields)
  nsb_sem_t c0_sem = nullptr;    // C0 layout (nsb_c0.cu): fields [0, c0_nfields) live on the distinct nodes of this mesh
  int c0_nfields = 0;
  std::vector<int> in_dot;
  std::vector<int64_t> off;      // row offset of each field inside a column
  int time_in_dot = 0;
  int64_t time_row = 0;          // row holding %time
  int64_t ndot = 0;              // rows [0, ndot) are covered by the inner product (padded)
  int64_t ld = 0;                // rows of a column (padded)
  int64_t ndof_dot = 0;          // unpadded dofs in the inner product
  int64_t nact = 0;              // unpadded rows of a column (all fields + %time)
  double *w_d = nullptr;         // weight, ndot rows (zeros on pads)
};

struct nsb_basis_s {
  nsb_layout_t lay = nullptr;
  int ncols = 0;
  double *v_d = nullptr;         // [ld * ncols]
  inline double *col(int c) const { return v_d + (size_t)c * (size_t)lay->ld; }
};

struct nsb_sem_s {
  nsb_context_t ctx = nullptr;
  int dim = 3, N = 7, lx = 8;
  int64_t nel = 0, npts = 0;
  int ng = 6;                    // geometric factors per point (6 in 3-D, 3 in 2-D)
  double *g_d = nullptr;         // [ng][npts]  G1..G6  (2-D: G1, G2, G4)
  double *bm1_d = nullptr, *jac_d = nullptr, *binv_d = nullptr, *vmult_d = nullptr,
         *mask_d = nullptr, *bmask_d = nullptr;  // bmask = binvm1 * mask
  double *rst_d = nullptr;       // [dim*dim][npts] rx..tz (times jac), for the convective term
  // element range of the axhelm launch in flight (launch_axhelm sets them; slab pipeline of the fused operator)
  const double *ax_g = nullptr, *ax_bm1 = nullptr, *ax_bmask = nullptr;
  int64_t ax_nel = 0;
  double *ax_dotp = nullptr;     // request: the axhelm launch also leaves partial sums of (w_raw, u) per field here
  int ax_dot_rows = 0;           // answer: rows of [rows][nf] partials written (0: this kernel variant cannot)
  // slab pipeline: elements [slab_e0[s], slab_e0[s+1]) ; private gather-scatter nodes whose LAST copy lies in
  // slab s are [slab_node_end[s-1], slab_node_end[s]) -- their gather-scatter runs right after that slab's
  // axhelm, while w and u of the slab are still in L2
  int nslab = 1;
  std::vector<int64_t> slab_e0, slab_node_end;
  std::vector<int32_t> node_slab;        // slab of the last copy of every gs node
  std::vector<cudaEvent_t> slab_ev;
  cudaEvent_t ev_c = nullptr;
  double *D_d = nullptr;         // (N+1)^2, D[i + lx*j] = dxm1(i,j)
  std::vector<double> D_h, z_h, w_h;
  // gather-scatter: unique nodes owning >= 1 element-boundary point, CSR
  int64_t nshared = 0;           // number of such nodes (local)
  int64_t *gs_off_d = nullptr;   // [nshared+1]
  int32_t *gs_idx_d = nullptr;   // local point indices
  int64_t gs_nnz = 0;
  int64_t n_local = 0;           // nodes [0, n_local) are private to this rank, the rest are interface nodes
  double *bnode_d = nullptr;     // [nshared] bmask of every node (all copies of a node carry the same value)
  int ns_fields = 8;             // fields one batched gather-scatter can carry
  double *node_sum_d = nullptr;  // [ns_fields][nshared - n_local] interface node sums (multi-rank)
  std::vector<int64_t> node_gid; // global id of each gs node
  std::vector<int64_t> gs_off_h; // host copies of the CSR lists (reordered by exchange_setup)
  std::vector<int32_t> gs_idx_h;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  // inter-rank exchange (nsb_sem.cu / nsb_comm.cu)
  struct Peer {
    int rank;
    int64_t n;                   // shared nodes with that rank
    int32_t *idx_d;              // interface-node index (node - n_local) of every node shared with that rank, sorted by gid
    double *send_d, *recv_d;
    int64_t my_off = -1;         // P2P: offset (doubles) of this peer's region in MY mailbox halo area
    int64_t peer_off = -1;       // P2P: offset of MY region in the PEER's mailbox halo area
  };
  std::vector<Peer> peers;
  std::vector<int64_t> glo_h;    // kept for exchange setup
  bool exchange_ready = false;
  double *c0_scratch_d = nullptr;   // element-local scratch of the C0 layout (2 x nf x npts)
  size_t c0_scratch_elems = 0;
  int32_t *c0_l2u_d = nullptr;      // local point -> row of the unique layout
  int64_t c0_l2u_nshared = -1, c0_l2u_nlocal = -1;
  double *pcg_d = nullptr;       // work vectors of nsb_sem_hmholtz (r, p, w, z per system, d)
  double *diagA_d = nullptr;     // diagonal of A per local point (setprec), computed at the first solve
  bool p2p_halo = false;         // interface data is written straight into the peers' mailboxes
  unsigned long long *hx_seq_d = nullptr;  // device sequence number of this mesh's halo exchanges
  unsigned int *hx_ticket_d = nullptr;     // last-CTA ticket of the fused send kernel
  int32_t *ifc_poff_d = nullptr;           // per interface node: CSR offsets into ifc_pent_d
  int64_t *ifc_pent_d = nullptr;           // (peer index << 32 | position in that peer's packed list), ascending rank
  int64_t halo_flag_off = -1;    // offset (doubles, from the halo base) of this mesh's flag words [2][P] in MY mailbox
  std::vector<int64_t> peer_flag_off;   // the same offset in every peer's mailbox
  int64_t halo_region_doubles = 0;      // size of this mesh's reservation in the halo area
  // dealiased convection (nsb_conv.cu): lxd Gauss-Legendre points per direction
  int lxd = 0;
  double *J_d = nullptr, *Dg_d = nullptr;     // [lxd][lx] GLL -> GL interpolation, [lxd][lxd] derivative on GL
  double *rxf_d = nullptr;                    // [9][nel lxd^3] Gauss weights x metrics on the fine mesh
  double *cfine_d[2] = {nullptr, nullptr};    // [3][nel lxd^3] contravariant convecting fields (two slots)
  // pressure mesh of the P_N - P_N-2 splitting (nsb_ns.cu): lx2 = lx - 2 Gauss-Legendre points per direction
  int lx2 = 0;
  int64_t n2 = 0;                             // nel * lx2^dim pressure points on this rank
  double *i12_d = nullptr, *d12_d = nullptr;  // [lx2][lx] GLL -> GL interpolation, derivative at the Gauss points
  double *rx2_d = nullptr;                    // [dim*dim][n2] Gauss weights x metrics on the pressure mesh
  double *bm2inv_d = nullptr;                 // [n2] 1 / (w3m2 jacm2)
  double *ns_work_d = nullptr;                // pressure CG: r, p, w [n2 each] + dim velocity-shaped fields
  double *ns_state_d = nullptr;               // device scalars of the pressure CG
  // two-level preconditioner of the pressure operator (set up by the first solve that asks for it)
  double *fdm_S_d = nullptr, *fdm_lam_d = nullptr;   // [lx2][lx2] 1-D generalised eigenvectors, [lx2] eigenvalues
  double *fdm_c_d = nullptr;                  // [nel][3] direction weights of the element-wise solves
  int *cc_rowptr_d = nullptr, *cc_col_d = nullptr;   // coarse operator (one constant per element), CSR
  double *cc_val_d = nullptr, *cc_dinv_d = nullptr, *cc_vec_d = nullptr, *cc_partial_d = nullptr, *cc_state_d = nullptr;
  int64_t cc_nnz = 0;
  bool cc_coop = false;                       // the coarse solve runs as one cooperative kernel
  int cc_width = 0;                           // ELL width of the coarse operator
  int cc_maxit = 400;                         // most coarse CG iterations enqueued per application
  int cc_launch = 400;                        // currently enqueued (stops early on the device; adapted at every poll)
};

struct nsb_op_s {
  int kind = 0;                  // 0 sem, 1 host callback, 2 composition outer(inner(.)), 3 / 4 device time-steppers, 5 alpha outer + beta inner,
                                 // 6 finite-difference Frechet derivative of outer about (base_b, base_c)
  nsb_op_t outer = nullptr, inner = nullptr;
  nsb_basis_t tmp = nullptr;     // work vector of the composition
  nsb_sem_t sem = nullptr;
  int nfields_apply = 0;
  double alpha = 0, beta = 1, h1 = 1, h2 = 0;
  double *c_d = nullptr;         // [dim][npts] convecting velocity or null
  nsb_layout_t lay = nullptr;
  nsb_host_matvec_fn fn = nullptr;
  void *user = nullptr;
  std::vector<double *> hin, hout;  // pinned staging buffers of the host operator
  bool linear = false;              // nsb_op_set_linear: M(a x) = a M(x) may be used (un-normalised hand-over)
  int64_t napply = 0;
  // kind 3 (nsb_conv.cu): nsteps BDF3/EXT3 advection-diffusion steps; tmp holds the 7 work columns
  int slot = -1, nsteps = 0, maxit = 0;
  double kappa = 0, rho = 1, dt = 0, tol = 0;
  int64_t helm_iters = 0;        // Helmholtz iterations spent so far
  bool adjoint = false;          // kind 3: apply the discrete BM1-adjoint of the stepper (rmatvec)
  // kind 4 (nsb_ns.cu): pressure-coupled perturbation step (linearised Navier-Stokes, P_N - P_N-2)
  double nu = 0, tol_p = 0;
  int mean_free = 0, precond = 0;
  bool has_base = false;         // base flow in column 8 of tmp, its contravariant field in convection slot 0
  int64_t pres_iters = 0;        // pressure iterations spent so far
  nsb_basis_t orbit_b = nullptr; // kind 4: time-periodic base flow, step n linearises about column orbit_c0 + (n-1) orbit_stride
  int orbit_c0 = 0, orbit_stride = 1;
  bool c_dirty = false;          // kind 4, adjoint: c_d holds the gradient of an orbit column, not of the steady base flow
  // kind 6: forward_finite_difference_map (core/matvec.f90:246-379)
  nsb_basis_t base_b = nullptr;
  int base_c = 0, fd_order = 2;
};

namespace nsb {
int stepper_apply(nsb_op_t op, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout);   // nsb_conv.cu
int ns_stepper_apply(nsb_op_t op, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout);   // nsb_ns.cu
void ns_free(nsb_sem_t S);                                                                // nsb_ns.cu
enum ProfClass { PC_MULTIDOT = 0, PC_UPDATE, PC_NORMALIZE, PC_AXHELM, PC_GS, PC_BLAS1, PC_SMALL,
                 PC_ROTATE, PC_GEMV, PC_DOT, PC_FUSED, PC_COUNT };
// RAII: records a start event now and a stop event at scope exit on the context stream
struct ProfScope {
  nsb_context_t c;
  int slot = -1;
  ProfScope(nsb_context_t ctx, int cls, double bytes);
  ~ProfScope();
};
// implemented in nsb_core.cu
int ensure_partial(nsb_context_t ctx, int64_t rows);
// implemented in nsb_comm.cu
int comm_init(nsb_context_t ctx, const void *unique_id);
int comm_destroy(nsb_context_t ctx);
int allreduce_sum_d(nsb_context_t ctx, double *buf_d, int n);  // on ctx->stream, in place
int check_dev_err(nsb_context_t ctx);                          // NSB_ECUDA if a kernel reported a spin timeout
void clear_step_graphs(nsb_context_t ctx);                     // any handle a captured step refers to went away
int sendrecv_d(nsb_context_t ctx, const std::vector<nsb_sem_s::Peer> &peers, int nf, cudaStream_t st);
int exchange_setup(nsb_sem_t sem);
int halo_exchange_p2p(nsb_sem_t S, int nf, cudaStream_t st);  // pack -> peer stores -> flags -> wait -> add
// sum + peer stores + flags in one kernel, wait + add + scatter in a second (peer-memory transport)
int halo_exchange_fused(nsb_sem_t S, double *v, int nf, int64_t fstride, int epi, const double *uin, double alpha,
                        double beta, const double *bmask, cudaStream_t st);
// host-only plans (also reachable through nsb_host_gs_plan / nsb_host_exchange_plan for CPU tests)
int gs_plan(int dim, int lx, int64_t nel, const int64_t *glo_num, const double *mask, std::vector<int64_t> &off,
            std::vector<int32_t> &idx, std::vector<int64_t> &gid, std::vector<double> *vmult,
            int64_t slab_elems = 0, std::vector<int32_t> *node_slab = nullptr);
void compute_slab_ends(nsb_sem_t S);
struct ExchangePlan {
  std::vector<int64_t> newpos;                   // node -> position after the private/interface reorder
  int64_t n_local = 0;                           // private nodes
  std::vector<std::vector<int32_t>> peer_nodes;  // per rank: interface-relative node indices, ascending gid
};
int exchange_plan(int rank, int nranks, const std::vector<int64_t> &gid, const std::vector<int64_t> &cnt,
                  const int64_t *all_sorted, int64_t mx, ExchangePlan &plan);
// implemented in nsb_c0.cu (unique-node storage) / nsb_sem.cu / nsb_comm.cu
int c0_upload(nsb_basis_t B, int col, const double *const *fields);
int c0_download(nsb_basis_t B, int col, double *const *fields);
int c0_set_weight(nsb_layout_t L, int f, const double *w_host);
int c0_apply_sem(nsb_op_t op, nsb_basis_t bin, int cin, nsb_basis_t bout, int cout);
int64_t c0_nint(nsb_sem_t S);
int launch_axhelm_ext(nsb_sem_t S, const double *u, double *w, int nf, int64_t fstride, double h1, double h2,
                      const double *cv, int epi, double alpha, double beta, const double *bmask);
// gather-scatter of nf equally spaced fields incl. the inter-rank exchange (epi 0: dssum; 1: alpha uin + beta bmask sum)
int launch_gs_ext(nsb_sem_t S, double *v, int nf, int64_t fstride, int epi, const double *uin, double alpha, double beta,
                  const double *bmask);
int halo_exchange_fused_c0(nsb_sem_t S, const double *wloc, int nf, int64_t fs_loc, double *wu, const double *uu,
                           int64_t fs_u, double alpha, double beta, cudaStream_t st);
// implemented in nsb_orth.cu
int weighted_multidot(nsb_basis_t b, int k, const double *w_col_d, double *h_d);
int upload_multidot_pipelined(nsb_basis_t B, int col_w, const double *const *fields, double time, int k,
                              const double *scale_by_inv_d = nullptr);
struct StreamOut {            // host (pinned) destinations of a vector streamed out chunk by chunk
  double *const *fields;
  double *time;
};
int orth_enqueue(nsb_basis_t B, int k, int col_w, int mode, const StreamOut *so);
int orthonormalize_stream_out(nsb_basis_t B, int k, int col_w, int mode, double *h, const StreamOut *so);
}  // namespace nsb
