// evp_halo.cu -- see evp_halo.h.
//
// Semantics (restated from ice_HaloUpdate2DR8, ice_boundary.F90:1066-1760, as the per-cell values
// the reference's halochk unit test expects, halochk.F90:530-830) for a NE-corner vector field:
//   * a ghost cell whose global index lies in the domain gets the owner's interior value;
//     cyclic directions wrap the index;
//   * ghost cells outside an open/closed edge are not touched (no fillValue is passed for the
//     dyn fields, ice_boundary.F90:1173-1181);
//   * tripole (u-fold): ghost row ny_global+1 <- -f(nx_global-i, ny_global-1); the top row
//     ny_global is symmetrised, f <- 0.5*(f(i) - f(nx_global-i)), except the two pole points
//     i = nx_global/2 and nx_global, which become -f(i)  (ice_boundary.F90:1630-1649, 1689-1722).
//     (for i > nx_global/2 the reference forms -(0.5*(f(nx_global-i) - f(i))): same value, but a zero keeps that sign)
// Every output is a function of the pre-update values only, so one staged exchange suffices.
#include "evp_halo.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "evp_b200.h"

namespace evp {

enum { OP_COPY = 0, OP_NEG = 1, OP_AVGM = 2, OP_AVGH = 3 };

#define HFAIL(...) do { snprintf(err, nerr, __VA_ARGS__); return 1; } while (0)
#define HCK(call)                                                                         \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) HFAIL("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#define NCK(call)                                                                         \
  do {                                                                                    \
    ncclResult_t r_ = (call);                                                             \
    if (r_ != ncclSuccess) HFAIL("%s:%d %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_)); \
  } while (0)

int comm_get_unique_id(void *id128, char *err, size_t nerr) {
  static_assert(sizeof(ncclUniqueId) <= EVP_B200_UNIQUE_ID_BYTES, "unique id size");
  if (!id128) HFAIL("evp_b200_get_unique_id: null buffer");
  ncclUniqueId id;
  NCK(ncclGetUniqueId(&id));
  memset(id128, 0, EVP_B200_UNIQUE_ID_BYTES);
  memcpy(id128, &id, sizeof id);
  return 0;
}

int comm_init(CommState &cs, int rank, int nranks, const void *id128, char *err, size_t nerr) {
  if (nranks < 1 || rank < 0 || rank >= nranks) HFAIL("evp_b200_comm_init: rank %d of %d", rank, nranks);
  comm_destroy(cs);
  cs.rank = rank;
  cs.nranks = nranks;
  if (nranks == 1) return 0;
  if (!id128) HFAIL("evp_b200_comm_init: null id");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NCK(ncclCommInitRank(&cs.comm, nranks, id, rank));
  return 0;
}

void comm_destroy(CommState &cs) {
  if (cs.comm) ncclCommDestroy(cs.comm);
  cs = CommState{};
}

int stress_fold_plan_host(int nranks, const int *rects, int rank, int nxg, int nyg, int ns, int *n_seg, int *seg_out, int *n_cell,
                          int *cell_out, int cap) {
  *n_seg = 0; *n_cell = 0;
  const int gi0 = rects[4 * rank], gj0 = rects[4 * rank + 1], nx = rects[4 * rank + 2], ny = rects[4 * rank + 3];
  if (ns != EVP_B200_BNDY_TRIPOLE || nx < 1 || ny < 1 || gj0 + ny - 1 != nyg) return 0;
  int rc = 0;
  auto top = [&](int t) { return rects[4 * t + 2] >= 1 && rects[4 * t + 3] >= 1 && rects[4 * t + 1] + rects[4 * t + 3] - 1 == nyg; };
  for (int t = 0; t < nranks; ++t) {
    if (t == rank || !top(t)) continue;
    if (*n_seg < cap && seg_out) { seg_out[3 * *n_seg] = t; seg_out[3 * *n_seg + 1] = rects[4 * t]; seg_out[3 * *n_seg + 2] = rects[4 * t + 2]; }
    else rc = 1;
    ++*n_seg;
  }
  for (int i = 0; i <= nx + 1; ++i) {
    const int im = fold_mirror_col(nxg, gi0, i);
    int src = -1, col = 0;
    for (int t = 0; t < nranks; ++t)
      if (top(t) && im >= rects[4 * t] && im < rects[4 * t] + rects[4 * t + 2]) { src = t; col = im - rects[4 * t] + 1; break; }
    if (*n_cell < cap && cell_out) { cell_out[3 * *n_cell] = i; cell_out[3 * *n_cell + 1] = src; cell_out[3 * *n_cell + 2] = col; }
    else rc = 1;
    ++*n_cell;
  }
  return rc;
}

int stress_rows_exchange(CommState &cs, const std::vector<int> &rects, int nxg, int nyg, double *const *sig, int ld, int nx, int ny, int gi0,
                         int gj0, double *rowtop, cudaStream_t s, char *err, size_t nerr) {
  if (gj0 + ny - 1 != nyg) return 0;
  HCK(cudaMemsetAsync(rowtop, 0, sizeof(double) * 12 * (size_t)nxg, s));
  for (int q = 0; q < 12; ++q)
    HCK(cudaMemcpyAsync(rowtop + (size_t)q * nxg + (gi0 - 1), sig[q] + (size_t)ny * ld + 1, sizeof(double) * nx, cudaMemcpyDeviceToDevice, s));
  if (cs.nranks < 2) return 0;
  if (!cs.comm || (int)rects.size() < 4 * cs.nranks) HFAIL("stress symmetrisation: no communicator / rank table");
  // every pair of top-row ranks swaps its twelve row segments; same order on both sides (the segment list: stress_fold_plan_host)
  std::vector<int> seg(3 * (size_t)cs.nranks), cell(3 * (size_t)(nx + 2));
  int nseg = 0, ncell = 0;
  if (stress_fold_plan_host(cs.nranks, rects.data(), cs.rank, nxg, nyg, EVP_B200_BNDY_TRIPOLE, &nseg, seg.data(), &ncell, cell.data(),
                            std::max(cs.nranks, nx + 2)))
    HFAIL("stress symmetrisation: internal: plan overflow");
  NCK(ncclGroupStart());
  for (int k = 0; k < nseg; ++k) {
    const int t = seg[3 * k], ti0 = seg[3 * k + 1], tnx = seg[3 * k + 2];
    for (int q = 0; q < 12; ++q) {
      NCK(ncclSend(sig[q] + (size_t)ny * ld + 1, (size_t)nx, ncclDouble, t, cs.comm, s));
      NCK(ncclRecv(rowtop + (size_t)q * nxg + (ti0 - 1), (size_t)tnx, ncclDouble, t, cs.comm, s));
    }
  }
  NCK(ncclGroupEnd());
  return 0;
}

__global__ void halo_pack(const double *__restrict__ U, const double *__restrict__ V, const int *__restrict__ idx, int n,
                          double *__restrict__ buf) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int c = idx[k];
    buf[2 * k] = U[c];
    buf[2 * k + 1] = V[c];
  }
}

__device__ __forceinline__ double2 slot(const double *__restrict__ packbuf, const double *__restrict__ recvbuf, int nloc, int r) {
  const double *b = (r < nloc) ? packbuf + 2 * (size_t)r : recvbuf + 2 * (size_t)(r - nloc);
  return make_double2(b[0], b[1]);
}

__global__ void halo_apply(double *__restrict__ U, double *__restrict__ V, const int *__restrict__ dst, const int *__restrict__ s1,
                           const int *__restrict__ s2, const signed char *__restrict__ code, int n, int nloc,
                           const double *__restrict__ packbuf, const double *__restrict__ recvbuf) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const double2 a = slot(packbuf, recvbuf, nloc, s1[k]);
    double u = a.x, v = a.y;
    const int op = code[k];
    if (op == OP_NEG) {
      u = -u; v = -v;
    } else if (op == OP_AVGM || op == OP_AVGH) {
      // xavg = 0.5*(x1 + isign*x2) with isign = -1, x1 the partner with the lower i (ice_boundary.F90:1641-1646).
      // The lower partner receives isign*(isign*xavg) = xavg, the upper one isign*xavg: written exactly so, because
      // -(0.5*(x1-x2)) and 0.5*(x2-x1) differ in the sign of a zero result.
      const double2 b = slot(packbuf, recvbuf, nloc, s2[k]);
      u = 0.5 * (a.x - b.x);
      v = 0.5 * (a.y - b.y);
      if (op == OP_AVGH) { u = -u; v = -v; }
    }
    U[dst[k]] = u;
    V[dst[k]] = v;
  }
}

struct Rect { int gi0, gj0, nx, ny; };
static inline int ld_of(int nx) { return ((nx + 2 + 15) / 16) * 16; }
// rows of a dom array: ny + ghost ring + two staging rows for raw values that cross a tripole fold (P2PState::setup)
size_t dom_cells(int nx, int ny) { return (size_t)ld_of(nx) * (size_t)(ny + 4); }

struct Entry { int dst; int r1, c1, r2, c2; signed char code; };  // sources as (rank, dom index)

// destination entries of rank r, in a fixed order every rank can reproduce
static void enumerate(const std::vector<Rect> &R, int r, int nxg, int nyg, int ew, int ns, std::vector<Entry> &out) {
  const Rect &A = R[r];
  const int lda = ld_of(A.nx);
  const bool tripole = (ns == EVP_B200_BNDY_TRIPOLE);
  const bool wew = (ew == EVP_B200_BNDY_CYCLIC && A.nx == nxg), wns = (ns == EVP_B200_BNDY_CYCLIC && A.ny == nyg);
  auto owner = [&](int gi, int gj, int &rk, int &c) {
    for (int q = 0; q < (int)R.size(); ++q)
      if (gi >= R[q].gi0 && gi < R[q].gi0 + R[q].nx && gj >= R[q].gj0 && gj < R[q].gj0 + R[q].ny) {
        rk = q;
        c = (gj - R[q].gj0 + 1) * ld_of(R[q].nx) + (gi - R[q].gi0 + 1);
        return true;
      }
    return false;
  };
  auto wrapi = [&](int gi) { return ((gi - 1) % nxg + nxg) % nxg + 1; };
  for (int dj = 0; dj <= A.ny + 1; ++dj)
    for (int di = 0; di <= A.nx + 1; ++di) {
      const bool ring = (di == 0 || di == A.nx + 1 || dj == 0 || dj == A.ny + 1);
      int gi = A.gi0 + di - 1, gj = A.gj0 + dj - 1;
      const bool toprow = tripole && gj == nyg;
      if (!ring && !toprow) continue;
      if (gi < 1 || gi > nxg) {
        if (ew != EVP_B200_BNDY_CYCLIC) continue;
        gi = wrapi(gi);
      }
      Entry e{dj * lda + di, -1, -1, -1, -1, OP_COPY};
      if (gj < 1) {
        if (ns != EVP_B200_BNDY_CYCLIC) continue;
        gj += nyg;
      } else if (gj > nyg && ns == EVP_B200_BNDY_CYCLIC) {
        gj -= nyg;
      }
      if (gj > nyg) {
        if (!tripole) continue;
        const int it = wrapi(nxg - gi);
        if (!owner(it, nyg - 1, e.r1, e.c1)) continue;
        e.code = OP_NEG;
      } else if (toprow) {
        if (!owner(gi, nyg, e.r1, e.c1)) continue;
        if (gi == nxg / 2 || gi == nxg) {
          e.code = OP_NEG;
        } else if (gi < nxg / 2) {
          if (!owner(wrapi(nxg - gi), nyg, e.r2, e.c2)) continue;
          e.code = OP_AVGM;  // s1 = own (lower partner), s2 = mirror
        } else {
          // upper partner: s1 = mirror (the lower partner), s2 = own, result negated
          e.r2 = e.r1; e.c2 = e.c1;
          if (!owner(wrapi(nxg - gi), nyg, e.r1, e.c1)) continue;
          e.code = OP_AVGH;
        }
      } else {
        if (!owner(gi, gj, e.r1, e.c1)) continue;
        // ghost cells the compute kernels fill themselves by wrap stores
        const bool inj = (dj >= 1 && dj <= A.ny), ini = (di >= 1 && di <= A.nx);
        const bool ghost_i = (di == 0 || di == A.nx + 1), ghost_j = (dj == 0 || dj == A.ny + 1);
        if (e.r1 == r && ((wew && ghost_i && (inj || (wns && ghost_j))) || (wns && ghost_j && ini))) continue;
      }
      out.push_back(e);
    }
}

int halo_plan_host(int nranks, const int *rects, int rank, int nxg, int nyg, int ew, int ns, int *n, int *out, int cap) {
  std::vector<Rect> R(nranks);
  for (int q = 0; q < nranks; ++q) R[q] = Rect{rects[4 * q], rects[4 * q + 1], rects[4 * q + 2], rects[4 * q + 3]};
  std::vector<Entry> es;
  enumerate(R, rank, nxg, nyg, ew, ns, es);
  *n = (int)es.size();
  for (int k = 0; k < (int)es.size() && k < cap; ++k) {
    int *o = out + 6 * k;
    o[0] = es[k].dst; o[1] = es[k].r1; o[2] = es[k].c1; o[3] = es[k].r2; o[4] = es[k].c2; o[5] = es[k].code;
  }
  return 0;
}
int dom_pitch(int nx) { return ld_of(nx); }

template <class T>
static int up(T *&dptr, const std::vector<T> &h, char *err, size_t nerr) {
  HCK(cudaMalloc(&dptr, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) HCK(cudaMemcpy(dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

int HaloPlan::build(CommState &cs, int gi0, int gj0, int nx, int ny, int ld, int nxg, int nyg, int ew, int ns, bool allow_partial,
                    char *err, size_t nerr) {
  release();
  wrap_ew = (ew == EVP_B200_BNDY_CYCLIC && nx == nxg);
  wrap_ns = (ns == EVP_B200_BNDY_CYCLIC && ny == nyg);
  if (ld != ld_of(nx)) HFAIL("halo: pitch mismatch");

  // every rank learns every rectangle
  std::vector<Rect> R(cs.nranks);
  R[cs.rank] = Rect{gi0, gj0, nx, ny};
  if (cs.nranks > 1) {
    int *dbuf = nullptr;
    HCK(cudaMalloc(&dbuf, sizeof(Rect) * cs.nranks));
    HCK(cudaMemcpy(dbuf + 4 * cs.rank, &R[cs.rank], sizeof(Rect), cudaMemcpyHostToDevice));
    NCK(ncclAllGather(dbuf + 4 * cs.rank, dbuf, 4, ncclInt32, cs.comm, 0));
    HCK(cudaStreamSynchronize(0));
    HCK(cudaMemcpy(R.data(), dbuf, sizeof(Rect) * cs.nranks, cudaMemcpyDeviceToHost));
    HCK(cudaFree(dbuf));
    // the bounding rectangles must be disjoint (cartesian distribution); they need not cover the domain: ghost cells that
    // face eliminated land blocks have no source and keep the zeros the host's halo update put there
    for (int a = 0; a < cs.nranks; ++a)
      for (int b = a + 1; b < cs.nranks; ++b) {
        const bool sep = R[a].gi0 + R[a].nx <= R[b].gi0 || R[b].gi0 + R[b].nx <= R[a].gi0 || R[a].gj0 + R[a].ny <= R[b].gj0 ||
                         R[b].gj0 + R[b].ny <= R[a].gj0;
        if (!sep) HFAIL("halo: the blocks of ranks %d and %d interleave (bounding rectangles overlap); use distribution_type = 'cartesian'", a, b);
      }
  } else if ((nx != nxg || ny != nyg) && !allow_partial) {
    HFAIL("halo: one rank but its blocks span %dx%d of the %dx%d domain: call evp_b200_comm_init first, or -- if the rest of the "
          "domain is eliminated land blocks -- evp_b200_allow_partial_domain(1)", nx, ny, nxg, nyg);
  }

  rects.assign(4 * (size_t)cs.nranks, 0);
  for (int q = 0; q < cs.nranks; ++q) { rects[4 * q] = R[q].gi0; rects[4 * q + 1] = R[q].gj0; rects[4 * q + 2] = R[q].nx; rects[4 * q + 3] = R[q].ny; }
  const int me = cs.rank;
  std::vector<Entry> mine;
  enumerate(R, me, nxg, nyg, ew, ns, mine);
  has_fold = false;
  for (const Entry &e : mine) has_fold = has_fold || e.code != OP_COPY || e.r1 == me;

  // pack list: local sources first, then what each peer needs from me (in the peer's entry order)
  std::vector<int> pack_idx, h_dst, h_s1, h_s2;
  std::vector<signed char> h_code;
  std::vector<std::vector<int>> recv_from(cs.nranks);  // per source rank: count of slots, assigned in my entry order
  std::vector<int> recv_count(cs.nranks, 0);
  struct Ref { int rank, pos; };
  std::vector<Ref> ref1(mine.size()), ref2(mine.size());
  for (size_t k = 0; k < mine.size(); ++k) {
    const Entry &e = mine[k];
    auto add = [&](int rk, int c) -> Ref {
      if (rk < 0) return Ref{-1, -1};
      if (rk == me) { pack_idx.push_back(c); return Ref{me, (int)pack_idx.size() - 1}; }
      return Ref{rk, recv_count[rk]++};
    };
    ref1[k] = add(e.r1, e.c1);
    ref2[k] = add(e.r2, e.c2);
  }
  n_loc = (int)pack_idx.size();
  // receive offsets per peer, ascending rank
  std::vector<int> recv_off(cs.nranks, 0), send_off(cs.nranks, 0), send_cnt(cs.nranks, 0);
  int roff = 0;
  for (int q = 0; q < cs.nranks; ++q) { recv_off[q] = roff; roff += recv_count[q]; }
  n_recv = roff;
  // sends: walk every other rank's entries in its order
  for (int q = 0; q < cs.nranks; ++q) {
    if (q == me) continue;
    std::vector<Entry> theirs;
    enumerate(R, q, nxg, nyg, ew, ns, theirs);
    send_off[q] = (int)pack_idx.size();
    for (const Entry &e : theirs) {
      if (e.r1 == me) pack_idx.push_back(e.c1);
      if (e.r2 == me) pack_idx.push_back(e.c2);
    }
    send_cnt[q] = (int)pack_idx.size() - send_off[q];
  }
  n_pack = (int)pack_idx.size();
  for (size_t k = 0; k < mine.size(); ++k) {
    auto slot_of = [&](const Ref &r) { return r.rank < 0 ? -1 : (r.rank == me ? r.pos : n_loc + recv_off[r.rank] + r.pos); };
    h_dst.push_back(mine[k].dst);
    h_s1.push_back(slot_of(ref1[k]));
    h_s2.push_back(slot_of(ref2[k]));
    h_code.push_back(mine[k].code);
  }
  n_dst = (int)mine.size();
  for (int q = 0; q < cs.nranks; ++q)
    if (q != me && (send_cnt[q] || recv_count[q])) peers.push_back(Peer{q, send_off[q], send_cnt[q], recv_off[q], recv_count[q]});

  if (up(d_pack_idx, pack_idx, err, nerr) || up(d_dst, h_dst, err, nerr) || up(d_s1, h_s1, err, nerr) || up(d_s2, h_s2, err, nerr) ||
      up(d_code, h_code, err, nerr))
    return 1;
  HCK(cudaMalloc(&d_packbuf, sizeof(double) * 2 * std::max(n_pack, 1)));
  HCK(cudaMalloc(&d_recvbuf, sizeof(double) * 2 * std::max(n_recv, 1)));
  const char *eg = getenv("EVP_B200_GRAPH");
  allow_graph = !(eg && eg[0] == '0');
  const char *ep = getenv("EVP_B200_P2P");
  no_fold_kernel = (ep && ep[0] == '0');
  return 0;
}

int HaloPlan::exchange(CommState &cs, double *U, double *V, cudaStream_t s, int *launches, char *err, size_t nerr) {
  *launches = 0;
  if (n_dst == 0 && n_pack == 0) return 0;
  if (fold_n > 0 && !no_fold_kernel) {
    // one rank: every source is local and the whole update is the fold list -- one kernel instead of pack + apply
    static const P2PParams nopeers{};
    HCK(exact::launch_fold(nopeers, U, V, d_fold_dst, d_fold_c1, d_fold_c2, d_fold_code, fold_n, 0, fold_pdl ? 1 : 0, s));
    ++*launches;
    return 0;
  }
  if (n_pack) {
    halo_pack<<<std::min((n_pack + 127) / 128, 296), 128, 0, s>>>(U, V, d_pack_idx, n_pack, d_packbuf);
    ++*launches;
  }
  if (!peers.empty()) {
    NCK(ncclGroupStart());
    for (const Peer &p : peers) {
      if (p.nsend) NCK(ncclSend(d_packbuf + 2 * (size_t)p.send_off, 2 * (size_t)p.nsend, ncclDouble, p.rank, cs.comm, s));
      if (p.nrecv) NCK(ncclRecv(d_recvbuf + 2 * (size_t)p.recv_off, 2 * (size_t)p.nrecv, ncclDouble, p.rank, cs.comm, s));
    }
    NCK(ncclGroupEnd());
    ++*launches;
  }
  if (n_dst) {
    halo_apply<<<std::min((n_dst + 127) / 128, 296), 128, 0, s>>>(U, V, d_dst, d_s1, d_s2, d_code, n_dst, n_loc, d_packbuf, d_recvbuf);
    ++*launches;
  }
  HCK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// P2P: map the neighbours' velocity arrays, build the push table, the local fold list and the edge-first tile order
// ------------------------------------------------------------------------------------------------
// How one rank's plan entries (enumerate) are served without a staged exchange:
//   copy / negate of ANOTHER rank's cell   -> that rank's subcycle kernel stores the value (negated across the fold) straight into
//                                             the destination cell over NVLink (push table);
//   0.5*(a-b) forms (tripole top row)      -> every source that lives on another rank is pushed RAW into a staging slot of the
//                                             destination rank (rows ny+2, ny+3 of its own u/v arrays, slots numbered in the
//                                             entry order every rank can reproduce); the destination rank combines them itself
//                                             (fold list, p2p_fold_kernel) once the peers' epoch flags are up;
//   anything whose sources are all local   -> fold list as well (ghost row below the fold fed from the rank's own row ny-1,
//                                             pole points, symmetrised pairs held by one rank).
struct FoldEntry { int dst, c1, c2; signed char code; };
struct PushEntry { int src, rank, dst, neg; };

// fold list of rank q and (optionally) what rank `me` has to push to q; returns false when q's staging rows overflow
static bool split_entries(const std::vector<Rect> &R, int q, int me, int nxg, int nyg, int ew, int ns, std::vector<FoldEntry> *fold,
                          std::vector<PushEntry> *push_from_me, std::vector<int> *remote_ranks) {
  std::vector<Entry> es;
  enumerate(R, q, nxg, nyg, ew, ns, es);
  const int ldq = ld_of(R[q].nx), stg_base = (R[q].ny + 2) * ldq, stg_cap = 2 * ldq;
  int slot = 0;
  for (const Entry &e : es) {
    const bool avg = (e.code == OP_AVGM || e.code == OP_AVGH);
    if (!avg) {
      if (e.r1 == q) {
        if (fold) fold->push_back(FoldEntry{e.dst, e.c1, e.c1, e.code});
      } else {
        if (remote_ranks) remote_ranks->push_back(e.r1);
        if (push_from_me && e.r1 == me) push_from_me->push_back(PushEntry{e.c1, q, e.dst, e.code == OP_NEG ? 1 : 0});
      }
      continue;
    }
    int s[2] = {e.c1, e.c2};
    const int rr[2] = {e.r1, e.r2};
    for (int w = 0; w < 2; ++w) {
      if (rr[w] == q) continue;
      if (slot >= stg_cap) return false;
      const int idx = stg_base + slot++;
      if (remote_ranks) remote_ranks->push_back(rr[w]);
      if (push_from_me && rr[w] == me) push_from_me->push_back(PushEntry{s[w], q, idx, 0});
      s[w] = idx;
    }
    if (fold) fold->push_back(FoldEntry{e.dst, s[0], s[1], e.code});
  }
  return true;
}

// host-only: the push list and the fold list of `rank` (what P2PState::setup builds), for the CPU tests of the multi-rank logic
int p2p_plan_host(int nranks, const int *rects, int rank, int nxg, int nyg, int ew, int ns, int *n_push, int *push_out, int *n_fold,
                  int *fold_out, int cap) {
  std::vector<Rect> R(nranks);
  for (int q = 0; q < nranks; ++q) R[q] = Rect{rects[4 * q], rects[4 * q + 1], rects[4 * q + 2], rects[4 * q + 3]};
  std::vector<PushEntry> pushes;
  std::vector<FoldEntry> fold;
  for (int q = 0; q < nranks; ++q) {
    const bool ok = (q == rank) ? split_entries(R, q, rank, nxg, nyg, ew, ns, &fold, nullptr, nullptr)
                                : split_entries(R, q, rank, nxg, nyg, ew, ns, nullptr, &pushes, nullptr);
    if (!ok) return 2;
  }
  *n_push = (int)pushes.size();
  *n_fold = (int)fold.size();
  for (int k = 0; k < (int)pushes.size() && k < cap; ++k) {
    int *o = push_out + 4 * k;
    o[0] = pushes[k].src; o[1] = pushes[k].rank; o[2] = pushes[k].dst; o[3] = pushes[k].neg;
  }
  for (int k = 0; k < (int)fold.size() && k < cap; ++k) {
    int *o = fold_out + 4 * k;
    o[0] = fold[k].dst; o[1] = fold[k].c1; o[2] = fold[k].c2; o[3] = fold[k].code;
  }
  return 0;
}

static int upload_fold(const std::vector<FoldEntry> &f, int *&d_dst, int *&d_c1, int *&d_c2, signed char *&d_code, char *err, size_t nerr) {
  std::vector<int> a, b, c;
  std::vector<signed char> k;
  for (const FoldEntry &e : f) { a.push_back(e.dst); b.push_back(e.c1); c.push_back(e.c2); k.push_back(e.code); }
  if (up(d_dst, a, err, nerr) || up(d_c1, b, err, nerr) || up(d_c2, c, err, nerr) || up(d_code, k, err, nerr)) return 1;
  return 0;
}

// one rank: every source of the fold is local, the whole exchange is the fold list (p2p_fold_kernel without peers)
int HaloPlan::build_local_fold(int nxg, int nyg, int ew, int ns, int max_entries, char *err, size_t nerr) {
  fold_n = 0;
  if (rects.size() != 4 || n_dst == 0) return 0;
  std::vector<Rect> R(1, Rect{rects[0], rects[1], rects[2], rects[3]});
  std::vector<FoldEntry> fold;
  std::vector<int> remote;
  if (!split_entries(R, 0, 0, nxg, nyg, ew, ns, &fold, nullptr, &remote) || !remote.empty()) return 0;
  if ((int)fold.size() > max_entries || fold.size() != (size_t)n_dst) return 0;
  if (upload_fold(fold, d_fold_dst, d_fold_c1, d_fold_c2, d_fold_code, err, nerr)) return 1;
  fold_n = (int)fold.size();
  return 0;
}

// setup only: tell a neighbour GPU which of its ghost cells this rank feeds (tag 1 in every word of the slot, both parities)
__global__ void ll_mark_kernel(unsigned long long *ll, const int *slot, int n, int ring) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || slot[k] < 0) return;
  for (int par = 0; par < P2P_LL_SLOTS; ++par)
    for (int w = 0; w < 4; ++w) ll[((size_t)par * ring + slot[k]) * 4 + w] = 1ULL << 32;
}

int P2PState::setup(CommState &cs, const HaloPlan &plan, double *dshare, size_t ndom, int nx, int ny, int ld, int nxg, int nyg,
                    int ew, int ns, int max_fold, char *err, size_t nerr) {
  release();
  enabled = false;
  const bool selftest = cs.nranks < 2 && getenv("EVP_B200_P2P_SELFTEST");
  if (cs.nranks < 2 && !selftest) { why = "single rank"; return 0; }
  const int ntx = (nx + 30) / 31, nty = (ny + 6) / 7;
  if (selftest) {
    // one rank, no peers: runs the edge-first tile order, the edge-CTA counter and the hand-shake kernels
    // without any NVLink traffic (timing / parity of the kernel structure itself)
    std::vector<int> order, start(3 * nx + 2 * (ny - 2) + 1, 0), none;
    int n_edge = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int t = 0; t < ntx * nty; ++t) {
        const int bx = t % ntx, by = t / ntx;
        const bool edge = (bx == 0 || bx == ntx - 1 || by == 0 || by == nty - 1);
        if (edge == (pass == 0)) { order.push_back(t); n_edge += edge; }
      }
    if (ntx > 0xffff || nty > 0x7fff) HFAIL("p2p: tile grid too large");
    for (int &t : order) t = ((t / ntx) << 16) | (t % ntx);  // packed (tby, tbx): see fused_kernel
    if (up(d_tile_order, order, err, nerr) || up(d_push_start, start, err, nerr) || up(d_push_peer, none, err, nerr) || up(d_push_dst, none, err, nerr)) return 1;
    HCK(cudaMalloc(&d_done, sizeof(unsigned long long))); HCK(cudaMalloc(&d_epoch, sizeof(unsigned long long))); HCK(cudaMalloc(&d_err, sizeof(int)));
    HCK(cudaMemset(d_done, 0, 8)); HCK(cudaMemset(d_err, 0, 4)); HCK(cudaMemset(d_epoch, 0, 8));
    prm = P2PParams{};
    prm.enabled = 1; prm.npeers = 0; prm.n_edge_tiles = n_edge; prm.ntx = ntx; prm.nty = nty;
    prm.tile_order = d_tile_order; prm.n_push = 0; prm.push_start = d_push_start; prm.push_peer = d_push_peer; prm.push_dst = d_push_dst;
    prm.my_flags = (const unsigned long long *)(dshare + 4 * ndom);
    prm.done_ctr = d_done; prm.epoch_base = d_epoch; prm.err = d_err;
    enabled = true; why = "selftest (no peers)";
    return 0;
  }
  const char *env = getenv("EVP_B200_P2P");
  int want = !(env && env[0] == '0');
  const int me = cs.rank;
  std::vector<Rect> R(cs.nranks);
  for (int q = 0; q < cs.nranks; ++q) R[q] = Rect{plan.rects[4 * q], plan.rects[4 * q + 1], plan.rects[4 * q + 2], plan.rects[4 * q + 3]};
  const bool tripole = (ns == EVP_B200_BNDY_TRIPOLE);
  const int fold_row = (tripole && R[me].gj0 + ny - 1 == nyg) ? ny - 1 : 0;

  // what I push to whom, what I combine myself, who pushes to me
  std::vector<PushEntry> pushes;
  std::vector<FoldEntry> fold;
  std::vector<int> peer_ranks;
  auto slot_of = [&](int rk) {
    for (size_t q = 0; q < peer_ranks.size(); ++q) if (peer_ranks[q] == rk) return (int)q;
    peer_ranks.push_back(rk);
    return (int)peer_ranks.size() - 1;
  };
  for (int q = 0; q < cs.nranks; ++q) {
    std::vector<int> remote;
    const bool ok = (q == me) ? split_entries(R, q, me, nxg, nyg, ew, ns, &fold, nullptr, &remote)
                              : split_entries(R, q, me, nxg, nyg, ew, ns, nullptr, &pushes, nullptr);
    if (!ok) { want = 0; why = "fold staging rows too small"; }
    for (int rk : remote) slot_of(rk);   // ranks that push to me
  }
  for (const PushEntry &p : pushes) slot_of(p.rank);
  if ((int)peer_ranks.size() > P2P_MAXPEER) { want = 0; why = "too many peers"; }
  if ((int)fold.size() > max_fold) { want = 0; why = "too many fold entries for one CTA"; }
  if (nx < 2 || ny < 3) { want = 0; why = "sub-domain too small"; }
  for (const PushEntry &p : pushes) {
    const int i = p.src % ld, j = p.src / ld;
    if (!(i == 1 || i == nx || j == 1 || j == ny || (fold_row && j == fold_row))) { want = 0; why = "a push source is not on the sub-domain edge"; }
  }

  // exchange IPC handles of the shared segment and try to map every peer
  cudaIpcMemHandle_t myh;
  memset(&myh, 0, sizeof myh);
  if (want) {
    cudaError_t e = cudaIpcGetMemHandle(&myh, dshare);
    if (e != cudaSuccess) { want = 0; why = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); }
  }
  unsigned char *dbuf = nullptr;
  const size_t hs = sizeof(cudaIpcMemHandle_t);
  HCK(cudaMalloc(&dbuf, hs * cs.nranks));
  HCK(cudaMemcpy(dbuf + hs * me, &myh, hs, cudaMemcpyHostToDevice));
  NCK(ncclAllGather(dbuf + hs * me, dbuf, hs, ncclChar, cs.comm, 0));
  HCK(cudaStreamSynchronize(0));
  std::vector<cudaIpcMemHandle_t> hall(cs.nranks);
  HCK(cudaMemcpy(hall.data(), dbuf, hs * cs.nranks, cudaMemcpyDeviceToHost));
  HCK(cudaFree(dbuf));
  npeers = (int)peer_ranks.size();
  if (want) {
    for (int q = 0; q < npeers; ++q) {
      void *ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, hall[peer_ranks[q]], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        want = 0; why = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); cudaGetLastError();
        break;
      }
      peer_base[q] = (double *)ptr;
      const Rect &A = R[peer_ranks[q]];
      peer_ndom[q] = dom_cells(A.nx, A.ny);
    }
  }
  // every rank must take the same path
  int *dflag = nullptr;
  HCK(cudaMalloc(&dflag, sizeof(int)));
  HCK(cudaMemcpy(dflag, &want, sizeof(int), cudaMemcpyHostToDevice));
  NCK(ncclAllReduce(dflag, dflag, 1, ncclInt32, ncclMin, cs.comm, 0));
  HCK(cudaStreamSynchronize(0));
  int all = 0;
  HCK(cudaMemcpy(&all, dflag, sizeof(int), cudaMemcpyDeviceToHost));
  HCK(cudaFree(dflag));
  if (!all) {
    if (want) why = "a peer rank could not use the in-kernel halo";
    for (int q = 0; q < npeers; ++q) if (peer_base[q]) { cudaIpcCloseMemHandle(peer_base[q]); peer_base[q] = nullptr; }
    npeers = 0;
    return 0;
  }

  // CSR over the edge index of the source cell (device twin: edge_index in evp_kernels.cu)
  const int nedge = 3 * nx + 2 * (ny - 2);
  std::vector<int> start(nedge + 1, 0), ppeer(pushes.size()), pdst(pushes.size());
  auto eidx = [&](int c) {
    const int i = c % ld, j = c / ld;
    if (j == 1) return i - 1;
    if (j == ny) return nx + i - 1;
    if (fold_row && j == fold_row) return 2 * nx + 2 * (ny - 2) + (i - 1);
    if (i == 1) return 2 * nx + (j - 2);
    return 2 * nx + (ny - 2) + (j - 2);
  };
  for (const PushEntry &p : pushes) start[eidx(p.src) + 1]++;
  for (int e = 0; e < nedge; ++e) start[e + 1] += start[e];
  std::vector<int> pll(pushes.size(), -1);   // ring index of the destination ghost cell in the peer's sub-domain (-1: not a ring cell)
  {
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (const PushEntry &p : pushes) {
      const int k = fill[eidx(p.src)]++;
      ppeer[k] = slot_of(p.rank) | (p.neg ? 0x100 : 0);
      pdst[k] = p.dst;
      const Rect &A = R[p.rank];
      const int ldr = ((A.nx + 2 + 15) / 16) * 16, di = p.dst % ldr, dj = p.dst / ldr;
      if (dj <= A.ny + 1 && (di == 0 || di == A.nx + 1 || dj == 0 || dj == A.ny + 1)) pll[k] = ring_index(A.nx, A.ny, di, dj);
    }
  }
  // tile order of the 32x8 fused kernel, edge tiles first; below a tripole fold the tile row that holds row ny-1 counts as edge
  const int fold_tby = fold_row ? (fold_row - 1) / 7 : -1;
  std::vector<int> order;
  int n_edge = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int t = 0; t < ntx * nty; ++t) {
      const int bx = t % ntx, by = t / ntx;
      const bool edge = (bx == 0 || bx == ntx - 1 || by == 0 || by == nty - 1 || by == fold_tby);
      if (edge == (pass == 0)) { order.push_back(t); n_edge += edge; }
    }
  if (ntx > 0xffff || nty > 0x7fff) HFAIL("p2p: tile grid too large");
  for (int &t : order) t = ((t / ntx) << 16) | (t % ntx);  // packed (tby, tbx): see fused_kernel
  if (up(d_tile_order, order, err, nerr) || up(d_push_start, start, err, nerr) || up(d_push_peer, ppeer, err, nerr) ||
      up(d_push_dst, pdst, err, nerr) || up(d_push_ll, pll, err, nerr))
    return 1;
  if (upload_fold(fold, d_fold_dst, d_fold_c1, d_fold_c2, d_fold_code, err, nerr)) return 1;
  fold_n = (int)fold.size();
  HCK(cudaMalloc(&d_done, sizeof(unsigned long long)));
  HCK(cudaMalloc(&d_epoch, sizeof(unsigned long long)));
  HCK(cudaMalloc(&d_err, sizeof(int)));
  HCK(cudaMalloc(&d_dbg, (8 + 8 * 1024) * sizeof(unsigned long long)));
  HCK(cudaMemset(d_dbg, 0, (8 + 8 * 1024) * sizeof(unsigned long long)));
  HCK(cudaMemset(d_done, 0, sizeof(unsigned long long)));
  HCK(cudaMemset(d_err, 0, sizeof(int)));
  const unsigned long long one = 1;  // flags start at 0: epoch 0 is "nothing yet"
  HCK(cudaMemcpy(d_epoch, &one, sizeof one, cudaMemcpyHostToDevice));
  HCK(cudaMemset(dshare + 4 * ndom, 0, 64 * sizeof(unsigned long long)));

  prm = P2PParams{};
  prm.enabled = 1; prm.npeers = npeers; prm.n_edge_tiles = n_edge; prm.ntx = ntx; prm.nty = nty;
  prm.tile_order = d_tile_order;
  prm.n_push = (int)pushes.size(); prm.fold_row = fold_row;
  prm.push_start = d_push_start; prm.push_peer = d_push_peer; prm.push_dst = d_push_dst;
  prm.my_flags = (const unsigned long long *)(dshare + 4 * ndom);
  for (int q = 0; q < npeers; ++q) {
    prm.peer_rank[q] = peer_ranks[q];
    prm.peer_flag[q] = (unsigned long long *)(peer_base[q] + 4 * peer_ndom[q]) + me;
  }
  prm.done_ctr = d_done; prm.epoch_base = d_epoch; prm.err = d_err;
  prm.dbg = getenv("EVP_B200_P2P_DEBUG") ? d_dbg : nullptr;
  set_parity(0);
  // nobody may be written to before every rank has zeroed its flags
  NCK(ncclAllReduce(d_err, d_err, 1, ncclInt32, ncclMax, cs.comm, 0));
  HCK(cudaStreamSynchronize(0));
  // low-latency slots: learn which of my ghost cells the neighbour GPUs feed (they mark them), then clear the slots again
  prm.my_ring = ring_cells(nx, ny);
  prm.my_ll = (unsigned long long *)(dshare + 4 * ndom) + 64;
  prm.push_ll = d_push_ll;
  for (int q = 0; q < npeers; ++q) {
    const Rect &A = R[peer_ranks[q]];
    prm.peer_ring[q] = ring_cells(A.nx, A.ny);
    prm.peer_ll[q] = (unsigned long long *)(peer_base[q] + 4 * peer_ndom[q]) + 64;
    std::vector<int> mine;
    for (size_t k = 0; k < pll.size(); ++k) if ((ppeer[k] & 0xff) == q) mine.push_back(pll[k]);
    int *dm = nullptr;
    if (up(dm, mine, err, nerr)) return 1;
    if (!mine.empty()) ll_mark_kernel<<<((int)mine.size() + 127) / 128, 128>>>(prm.peer_ll[q], dm, (int)mine.size(), prm.peer_ring[q]);
    HCK(cudaDeviceSynchronize());
    HCK(cudaFree(dm));
  }
  NCK(ncclAllReduce(d_err, d_err, 1, ncclInt32, ncclMax, cs.comm, 0));   // every mark has landed
  HCK(cudaStreamSynchronize(0));
  {
    std::vector<unsigned long long> w((size_t)P2P_LL_SLOTS * prm.my_ring * 4);
    HCK(cudaMemcpy(w.data(), prm.my_ll, w.size() * 8, cudaMemcpyDeviceToHost));
    std::vector<unsigned char> fed(prm.my_ring, 0);
    for (int r = 0; r < prm.my_ring; ++r) fed[r] = (w[(size_t)r * 4] >> 32) == 1ULL;
    if (up(d_ll_fed, fed, err, nerr)) return 1;
    prm.ll_fed = d_ll_fed;
    HCK(cudaMemset(prm.my_ll, 0, w.size() * 8));
  }
  NCK(ncclAllReduce(d_err, d_err, 1, ncclInt32, ncclMax, cs.comm, 0));   // ... and every rank has cleared its slots before anyone runs
  HCK(cudaStreamSynchronize(0));
  enabled = true;
  char b[200];
  snprintf(b, sizeof b, "in-kernel NVLink stores to %d peer(s), %zu pushed cells, %d edge tiles of %d, %d fold entries", npeers, pushes.size(),
           n_edge, ntx * nty, fold_n);
  why = b;
  return 0;
}

void P2PState::set_parity(int s) {
  swapped = s;
  for (int q = 0; q < npeers; ++q) {
    for (int b = 0; b < 2; ++b) {
      prm.peer_u[b][q] = peer_base[q] + (size_t)(b ^ s) * peer_ndom[q];
      prm.peer_v[b][q] = peer_base[q] + (size_t)(2 + (b ^ s)) * peer_ndom[q];
    }
  }
}

void P2PState::release() {
  for (int q = 0; q < P2P_MAXPEER; ++q) {
    if (peer_base[q]) cudaIpcCloseMemHandle(peer_base[q]);
    peer_base[q] = nullptr;
  }
  auto F = [](auto *&p) { if (p) cudaFree(p); p = nullptr; };
  F(d_tile_order); F(d_push_start); F(d_push_peer); F(d_push_dst); F(d_done); F(d_epoch); F(d_err); F(d_dbg); F(d_push_ll); F(d_ll_fed);
  F(d_fold_dst); F(d_fold_c1); F(d_fold_c2); F(d_fold_code);
  fold_n = 0;
  enabled = false;
  npeers = 0;
  prm = P2PParams{};
}

std::string HaloPlan::describe() const {
  char b[160];
  snprintf(b, sizeof b, "wrap_ew=%d wrap_ns=%d dst=%d local=%d recv=%d peers=%zu", wrap_ew, wrap_ns, n_dst, n_loc, n_recv, peers.size());
  return b;
}

void HaloPlan::release() {
  auto F = [](auto *&p) { if (p) cudaFree(p); p = nullptr; };
  F(d_pack_idx); F(d_dst); F(d_s1); F(d_s2); F(d_code); F(d_packbuf); F(d_recvbuf);
  F(d_fold_dst); F(d_fold_c1); F(d_fold_c2); F(d_fold_code);
  fold_n = 0;
  peers.clear();
  n_dst = n_pack = n_loc = n_recv = 0;
  wrap_ew = wrap_ns = 0;
}

}  // namespace evp
