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
// Every output is a function of the pre-update values only, so one staged exchange suffices.
#include "evp_halo.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "evp_b200.h"

namespace evp {

enum { OP_COPY = 0, OP_NEG = 1, OP_AVGM = 2 };

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
    } else if (op == OP_AVGM) {
      const double2 b = slot(packbuf, recvbuf, nloc, s2[k]);
      u = 0.5 * (a.x - b.x);  // 0.5*(x1 + isign*x2), isign = -1   (ice_boundary.F90:1641-1646)
      v = 0.5 * (a.y - b.y);
    }
    U[dst[k]] = u;
    V[dst[k]] = v;
  }
}

struct Rect { int gi0, gj0, nx, ny; };
static inline int ld_of(int nx) { return ((nx + 2 + 15) / 16) * 16; }

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
        } else {
          if (!owner(wrapi(nxg - gi), nyg, e.r2, e.c2)) continue;
          e.code = OP_AVGM;
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

int HaloPlan::build(CommState &cs, int gi0, int gj0, int nx, int ny, int ld, int nxg, int nyg, int ew, int ns, char *err,
                    size_t nerr) {
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
    size_t cells = 0;
    for (auto &q : R) cells += (size_t)q.nx * q.ny;
    if (cells != (size_t)nxg * nyg) HFAIL("halo: the ranks' rectangles cover %zu of %zu cells (land-block elimination is not supported)", cells, (size_t)nxg * nyg);
  } else if (nx != nxg || ny != nyg) {
    HFAIL("halo: one rank but its blocks cover %dx%d of the %dx%d domain; call evp_b200_comm_init first", nx, ny, nxg, nyg);
  }

  const int me = cs.rank;
  std::vector<Entry> mine;
  enumerate(R, me, nxg, nyg, ew, ns, mine);

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
  return 0;
}

int HaloPlan::exchange(CommState &cs, double *U, double *V, cudaStream_t s, int *launches, char *err, size_t nerr) {
  *launches = 0;
  if (n_dst == 0 && n_pack == 0) return 0;
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

std::string HaloPlan::describe() const {
  char b[160];
  snprintf(b, sizeof b, "wrap_ew=%d wrap_ns=%d dst=%d local=%d recv=%d peers=%zu", wrap_ew, wrap_ns, n_dst, n_loc, n_recv, peers.size());
  return b;
}

void HaloPlan::release() {
  auto F = [](auto *&p) { if (p) cudaFree(p); p = nullptr; };
  F(d_pack_idx); F(d_dst); F(d_s1); F(d_s2); F(d_code); F(d_packbuf); F(d_recvbuf);
  peers.clear();
  n_dst = n_pack = n_loc = n_recv = 0;
  wrap_ew = wrap_ns = 0;
}

}  // namespace evp
