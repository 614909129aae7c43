// evp_halo.h -- halo update of (uvel,vvel) on the device sub-domains: replaces the reference's
// dyn_haloUpdate -> ice_HaloUpdate for the dyn fields (ice_dyn_shared.F90:2518-2574,
// ice_boundary.F90:1066-1760) with NCCL point-to-point over NVLink plus two small kernels.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include "evp_internal.h"

#include <string>
#include <vector>

namespace evp {

struct CommState {
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
};

int comm_get_unique_id(void *id128, char *err, size_t nerr);
int comm_init(CommState &cs, int rank, int nranks, const void *id128, char *err, size_t nerr);
void comm_destroy(CommState &cs);

int halo_plan_host(int nranks, const int *rects, int rank, int nxg, int nyg, int ew, int ns, int *n, int *out, int cap);
int p2p_plan_host(int nranks, const int *rects, int rank, int nxg, int nyg, int ew, int ns, int *n_push, int *push_out, int *n_fold,
                  int *fold_out, int cap);
int dom_pitch(int nx);
size_t dom_cells(int nx, int ny);  // pitch * (ny + 2 ghost rows + 2 staging rows)

// What one exchange does, for every destination cell e of this rank (ghost ring, plus the top interior
// row on a tripole grid):  dst = op(code, s1, s2), where s1/s2 were staged from this rank's interior
// (pack kernel) or received from a neighbour rank.
struct HaloPlan {
  int wrap_ew = 0, wrap_ns = 0;  // ghost columns/rows the compute kernels fill themselves (on-rank cyclic wrap)

  int n_dst = 0;     // destination cells handled by the apply kernel
  int n_pack = 0;    // slots gathered by the pack kernel: [n_loc local | sends to peer 0 | peer 1 ...]
  int n_loc = 0;
  int n_recv = 0;
  int *d_pack_idx = nullptr;                 // [n_pack] dom index to read
  int *d_dst = nullptr, *d_s1 = nullptr, *d_s2 = nullptr;  // [n_dst] dom index to write, slot refs
  signed char *d_code = nullptr;             // [n_dst]
  double *d_packbuf = nullptr, *d_recvbuf = nullptr;       // 2 doubles (u,v) per slot
  // one rank, every source local (tripole fold on one GPU): the whole update as one kernel (p2p_fold_kernel, evp_kernels.cu)
  int fold_n = 0;
  bool no_fold_kernel = false;                             // EVP_B200_P2P=0: keep the staged pack + apply form
  bool fold_pdl = false;                                   // the kernel in front of it is a fused subcycle kernel (PDL chain)
  int *d_fold_dst = nullptr, *d_fold_c1 = nullptr, *d_fold_c2 = nullptr;
  signed char *d_fold_code = nullptr;
  int build_local_fold(int nxg, int nyg, int ew, int ns, int max_entries, char *err, size_t nerr);

  struct Peer { int rank, send_off, nsend, recv_off, nrecv; };
  std::vector<Peer> peers;
  bool allow_graph = true;

  std::vector<int> rects;  // 4 ints per rank {gi0, gj0, nx, ny}, filled by build
  bool has_fold = false;   // some entry is not a plain copy (tripole)

  int build(CommState &cs, int gi0, int gj0, int nx, int ny, int ld, int nxg, int nyg, int ew, int ns, bool allow_partial, char *err,
            size_t nerr);
  int exchange(CommState &cs, double *U, double *V, cudaStream_t s, int *launches, char *err, size_t nerr);
  bool graph_safe() const { return allow_graph; }
  std::string describe() const;
  void release();
};

// Tripole grids, stresses resident: the top physical row of the 12 stress arrays of every rank of the top row, gathered into
// rowtop[12][nxg] on each of those ranks (what ice_HaloUpdate_stress moves through its tripole buffer, ice_boundary.F90:7596-7688).
// rects as HaloPlan::rects; sig: the 12 dom arrays of the current copy.  Ranks below the top row return at once; rowtop is zeroed
// first (columns nobody holds -- eliminated land blocks -- stay at the reference's fill value).
// host side of it: which other ranks of the top row this rank swaps its row segment with -- 3 ints {rank, gi0, nx} each, in rank order --
// and, per ghost cell of its north ghost row (dst column 0 .. nx+1), the rank and the column of that rank's top row that is mirrored
// into it -- 3 ints {dst column, source rank or -1 (nobody holds the column: land), source column 1 .. nx_source}.  Empty on ranks
// below the top row and on grids without a fold.  Returns 1 when `cap` entries do not hold a list (counts are still returned).
int stress_fold_plan_host(int nranks, const int *rects, int rank, int nxg, int nyg, int ns, int *n_seg, int *seg_out, int *n_cell,
                          int *cell_out, int cap);
int stress_rows_exchange(CommState &cs, const std::vector<int> &rects, int nxg, int nyg, double *const *sig, int ld, int nx, int ny, int gi0,
                         int gj0, double *rowtop, cudaStream_t s, char *err, size_t nerr);

// In-kernel NVLink halo: peers' velocity arrays and flags mapped through CUDA IPC (one process per GPU).
struct P2PState {
  bool enabled = false;
  std::string why = "not set up";
  P2PParams prm{};
  int swapped = 0;
  int npeers = 0;
  double *peer_base[P2P_MAXPEER] = {};
  size_t peer_ndom[P2P_MAXPEER] = {};
  int *d_tile_order = nullptr, *d_push_start = nullptr, *d_push_peer = nullptr, *d_push_dst = nullptr;
  // what this rank combines itself once its peers' stores have arrived (tripole fold): p2p_fold_kernel after every subcycle kernel
  int fold_n = 0;
  int *d_fold_dst = nullptr, *d_fold_c1 = nullptr, *d_fold_c2 = nullptr;
  signed char *d_fold_code = nullptr;
  unsigned long long *d_done = nullptr, *d_epoch = nullptr;
  int *d_err = nullptr;
  unsigned long long *d_dbg = nullptr;
  int *d_push_ll = nullptr;
  unsigned char *d_ll_fed = nullptr;

  // dshare = [u0 | u1 | v0 | v1] (ndom doubles each), 64 u64 flags, then the low-latency slots (p2p_share_bytes), one cudaMalloc
  int setup(CommState &cs, const HaloPlan &plan, double *dshare, size_t ndom, int nx, int ny, int ld, int nxg, int nyg,
            int ew, int ns, int max_fold, char *err, size_t nerr);
  void set_parity(int swapped_);
  static size_t share_bytes(size_t ndom, int nx, int ny) { return 4 * ndom * sizeof(double) + 64 * 8 + (size_t)P2P_LL_SLOTS * ring_cells(nx, ny) * 4 * 8; }
  void release();
};

}  // namespace evp
