// evp_halo.h -- halo update of (uvel,vvel) on the device sub-domains: replaces the reference's
// dyn_haloUpdate -> ice_HaloUpdate for the dyn fields (ice_dyn_shared.F90:2518-2574,
// ice_boundary.F90:1066-1760) with NCCL point-to-point over NVLink plus two small kernels.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

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
int dom_pitch(int nx);

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

  struct Peer { int rank, send_off, nsend, recv_off, nrecv; };
  std::vector<Peer> peers;
  bool allow_graph = true;

  int build(CommState &cs, int gi0, int gj0, int nx, int ny, int ld, int nxg, int nyg, int ew, int ns, char *err, size_t nerr);
  int exchange(CommState &cs, double *U, double *V, cudaStream_t s, int *launches, char *err, size_t nerr);
  bool graph_safe() const { return allow_graph; }
  std::string describe() const;
  void release();
};

}  // namespace evp
