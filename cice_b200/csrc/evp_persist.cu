// evp_persist.cu -- KERNEL_PERSISTENT: all ndte subcycles of the EVP loop (ice_dyn_evp.F90:859-913) in ONE cooperative launch.
//
// One CTA per SM owns a bx x by tile of U points (and the (bx+1) x (by+1) T cells that close them) for the whole loop:
//   * the carried state of the tile -- 12 stresses per T cell, (u,v) with a one-cell ring -- the 8 `str` terms and the static
//     operands that are needed first (KT of the 10 T arrays, KU of the 11 U arrays) live in SHARED MEMORY for the whole loop;
//     the remaining static operands are read through L2 where their latency hides: dxhy, dyhx, strength, DminTarea are requested
//     when a T cell is started and used hundreds of cycles later, the momentum operands are requested BEFORE the CTA barrier
//     that separates the two phases;
//   * threads are not tied to a corner or a lane role: every thread advances PERSIST_SLOTS T cells and PERSIST_SLOTS U points per
//     subcycle, which ones is a table built on the host (evp_persist_plan.h).  The table orders the work so that the exchange
//     between tiles is off the critical path:
//        phase A (stress)   : T cells that read only the tile's own velocities first; the tile-edge T cells -- the only readers of
//                             the ring -- last, by the last warps, after those warps have refreshed the ring;
//        phase C (momentum) : the tile-edge U points first; their warps store them to the ping-pong arrays in global memory (L2)
//                             and publish (fence + one atomic per warp) BEFORE the interior U points are advanced.
//     A neighbour therefore has the whole interior momentum phase and the whole interior stress phase of slack before it reads.
//   * no grid-wide barrier: a tile waits for its <= 8 neighbours' counters only (acquire), tiles drift by at most one subcycle,
//     which the two copies of (u,v) in global memory cover (the same argument as for the launch-per-subcycle kernels' ping-pong);
//   * T cells on a tile's E/N overlap edge are relaxed by both tiles (same inputs, same arithmetic -> identical bits), exactly like
//     the reference evaluates N/E ghost T cells per block (ice_dyn_shared.F90:740-749).
// The arithmetic is evp_math.cuh's (stress_point / stepu_cv, interleaved division / square-root form), so the bits are those of
// the other kernels.  Compiled twice like evp_kernels.cu (namespace exact: -fmad=false; namespace fast).
#include "evp_math.cuh"
#include "evp_dom.cuh"
#include "evp_ptx.cuh"

#include <type_traits>

#ifndef EVP_NS
#error "compile with -DEVP_NS=exact or -DEVP_NS=fast"
#endif

namespace evp {
namespace EVP_NS {

// static per-cell arrays in the order they are needed (= shared-memory priority); on chip they are stored as pairs (double2)
enum { T_CXP = 0, T_CYP, T_CXM, T_CYM, T_DXT, T_DYT, T_DMIN, T_STRENGTH, T_DXHY, T_DYHX, T_COUNT };
enum { U_CV = 0, U_UOCN, U_VOCN, U_UMASSDTI, U_FM, U_WATERX, U_WATERY, U_FORCEX, U_FORCEY, U_UAREAR, U_TBU, U_COUNT };

__device__ __forceinline__ double2 mk2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }

// P2P: the sub-domain has neighbour GPUs.  Inside the loop their edge velocities arrive over NVLink as self-validating words in this
// rank's low-latency slots (P2PParams: value halves tagged with epoch + subcycle; no fence, no flag -- one store latency), and the
// tiles on this rank's edge push theirs the same way.  Only the last subcycle's values go into the neighbours' arrays themselves,
// handed over with a counter, one system-scope fence and the per-peer epoch flags (as fused_kernel<..,P2P> does every subcycle).
template <int NT, int KT, int KU, bool DBG = false, bool P2P = false>
__global__ void __launch_bounds__(NT, 1)
persist_kernel(const __grid_constant__ Dom d, const __grid_constant__ KParams k, const __grid_constant__ PersistPlan pp,
               const __grid_constant__ P2PParams px) {
  static_assert(KT % 2 == 0 && KU % 2 == 1 && KU >= 1, "static operands are kept on chip as pairs; cv (U operand 0) always");
  constexpr int NW = NT / 32, SL = PERSIST_SLOTS;
  double *sm = reinterpret_cast<double *>(dyn_smem());
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int ttx = tile % pp.ntx, tty = tile / pp.ntx;
  const int bx = pp.bx, by = pp.by;
  const int i0 = 1 + ttx * bx, j0 = 1 + tty * by;  // first U point / T cell of the tile (dom coordinates)
  const int tw = bx + 1, nT = pp.nT, nU = pp.nU;    // T pitch
  const int uw = bx + 2;                            // u/v pitch (ring included)
  // a tile in the last column / row may be narrower than bx x by: its ring sits right after its last valid column / row
  const int ebx = (d.nx - i0 + 1 < bx) ? d.nx - i0 + 1 : bx, eby = (d.ny - j0 + 1 < by) ? d.ny - j0 + 1 : by;
  const int shape = (ttx == pp.ntx - 1 ? 1 : 0) | (tty == pp.nty - 1 ? 2 : 0);
  const int ewT = pp.ewT[shape], ewU = pp.ewU[shape];
  // who refreshes the ring: the light warps while the others relax their first cell, or (none) the edge warps before their last
  const int nlw = pp.nlw[shape];
  const int rw0 = nlw ? pp.lw0[shape] : NW - ewT, nrw = nlw ? nlw : ewT;
  const bool edge_warp = warp >= NW - ewT, refresh_warp = warp >= rw0 && warp < rw0 + nrw;
  // tiles on the sub-domain edge talk to the neighbour GPUs
  const bool rank_edge = P2P && (ttx == 0 || ttx == pp.ntx - 1 || tty == 0 || tty == pp.nty - 1);
  const unsigned long long ebase = P2P ? *px.epoch_base : 0ULL;

  double2 *suv = reinterpret_cast<double2 *>(sm + pp.off_uv);     // (u, v) of the tile with its ring
  double2 *sstr = reinterpret_cast<double2 *>(sm + pp.off_str);   // 4 x nT: (str1,str5) (str2,str7) (str3,str6) (str4,str8) by T position
  double2 *ssig = reinterpret_cast<double2 *>(sm + pp.off_sig);   // 6 x nT by pidx: stressp_1..4, stressm_1..4, stress12_1..4 in pairs
  double2 *sT = reinterpret_cast<double2 *>(sm + pp.off_T);       // KT/2 x nT by pidx
  double2 *sU = reinterpret_cast<double2 *>(sm + pp.off_U);       // (KU-1)/2 x nU by pidx, then cv
  double *scv = sm + pp.off_U + (KU - 1) * nU;

  const double *gT[T_COUNT] = {d.cxp, d.cyp, d.cxm, d.cym, d.dxT, d.dyT, d.DminTarea, d.strength, d.dxhy, d.dyhx};
  const double *gU[U_COUNT] = {nullptr, d.uocn, d.vocn, d.umassdti, d.fm, d.waterx, d.watery, d.forcex, d.forcey, d.uarear, d.TbU};

  // ---------------- prologue: make the tile resident ------------------------------------------------------------------
  unsigned wT[SL], wU[SL];   // packed table words: t | li << 10 | lj << 16 | pidx << 22
  unsigned flags = 0;        // bit r: T cell of slot r carries ice; bit 2+r: ... and this tile stores it; bit 4+r: U point of slot r is advanced
#pragma unroll
  for (int r = 0; r < SL; ++r) {
    wT[r] = pp.tslot[(shape * SL + r) * NT + tid];
    wU[r] = pp.uslot[(shape * SL + r) * NT + tid];
    if (wT[r] != PERSIST_NONE_W) {
      const int li = (wT[r] >> 10) & 63, lj = (wT[r] >> 16) & 63, p = wT[r] >> 22;
      const int g = at(d, i0 + li, j0 + lj);
      if (d.maskT[g]) {
        flags |= 1u << r;
        if ((li < bx || i0 + li == d.nx + 1) && (lj < by || j0 + lj == d.ny + 1)) flags |= 4u << r;
      }
#pragma unroll
      for (int q = 0; q < KT / 2; ++q) sT[q * nT + p] = mk2(gT[2 * q][g], gT[2 * q + 1][g]);
#pragma unroll
      for (int q = 0; q < 6; ++q) ssig[q * nT + p] = mk2(d.sig[0][2 * q][g], d.sig[0][2 * q + 1][g]);
    }
    if (wU[r] != PERSIST_NONE_W) {
      const int li = (wU[r] >> 10) & 63, lj = (wU[r] >> 16) & 63, p = wU[r] >> 22;
      const int g = at(d, i0 + li, j0 + lj);
      if (d.maskU[g]) flags |= 16u << r;
      scv[p] = d.aiu[g] * k.rhow * d.cdn[g];  // aiX*rhow*Cw, left to right (stepu_cv)
#pragma unroll
      for (int q = 0; q < (KU - 1) / 2; ++q) sU[q * nU + p] = mk2(gU[1 + 2 * q][g], gU[2 + 2 * q][g]);
    }
  }
  for (int t = tid; t < 4 * nT; t += NT) sstr[t] = mk2(0.0, 0.0);  // cells off the ice are never visited: `str(:,:,:) = c0` (ice_dyn_evp.F90:1537)
  for (int c = tid; c < pp.nring; c += NT) {  // u, v of the tile and its ring: copy 0 holds the state at entry
    const int li = c % uw, lj = c / uw;
    const int i = i0 - 1 + li, j = j0 - 1 + lj;
    const bool in = (i <= d.nx + 1) && (j <= d.ny + 1);
    const int g = in ? at(d, i, j) : 0;
    suv[c] = in ? mk2(d.u[0][g], d.v[0][g]) : mk2(0.0, 0.0);
  }
  // neighbour tiles whose edge values feed my ring (E-W / N-S wrap where the ghost ring aliases the interior); lane q < 8 of the
  // ring-refreshing warps watches neighbour q, which publishes ewU[its shape] times per subcycle
  int nb = -1;
  unsigned nb_per = 0;
  if (lane < 8) {
    const int dxs[8] = {-1, 1, 0, 0, -1, 1, -1, 1}, dys[8] = {0, 0, -1, 1, -1, -1, 1, 1};
    int nx_ = ttx + dxs[lane], ny_ = tty + dys[lane];
    if (d.wrap_ew) nx_ = (nx_ + pp.ntx) % pp.ntx;
    if (d.wrap_ns) ny_ = (ny_ + pp.nty) % pp.nty;
    if (nx_ >= 0 && nx_ < pp.ntx && ny_ >= 0 && ny_ < pp.nty) nb = ny_ * pp.ntx + nx_;
    if (nb == tile) nb = -1;
    if (nb >= 0) nb_per = (unsigned)pp.ewU[(nx_ == pp.ntx - 1 ? 1 : 0) | (ny_ == pp.nty - 1 ? 2 : 0)];
  }
  __syncthreads();

  // per-warp cycle accounting (EVP_B200_PERSIST_DEBUG=1): [0] phase A, [1] of which ring wait + refresh, [2] barrier after A,
  // [3] phase C, [4] barrier after C
  long long acc[5] = {0, 0, 0, 0, 0};
  constexpr bool dbg = DBG;
  // timeline (DBG): %globaltimer of six events of subcycles PERSIST_TL0 .. PERSIST_TL0+3, per tile, behind the cycle counters
  unsigned long long *tl = dbg ? reinterpret_cast<unsigned long long *>(pp.dbg + (size_t)pp.ntx * pp.nty * NW * 5) + (size_t)tile * 4 * 8 : nullptr;
  auto mark = [&](int ksub, int ev) {
    if (dbg && lane == 0 && ksub >= PERSIST_TL0 && ksub < PERSIST_TL0 + 4) tl[(ksub - PERSIST_TL0) * 8 + ev] = gtime();
  };

  // one T cell: relax the stresses in place (shared memory), str -> shared memory
  auto relax = [&](unsigned w) {
    const int t = w & 1023, li = (w >> 10) & 63, lj = (w >> 16) & 63, p = w >> 22;
    const int c = (lj + 1) * uw + li + 1;
    double tv[T_COUNT];
    if (KT < T_COUNT) {
      const int g = at(d, i0 + li, j0 + lj);
#pragma unroll
      for (int q = KT; q < T_COUNT; ++q) tv[q] = ld_nc_f64(gT[q] + g);  // used late: the L2 round trip hides behind the strain rates
    }
#pragma unroll
    for (int q = 0; q < KT / 2; ++q) { const double2 x = sT[q * nT + p]; tv[2 * q] = x.x; tv[2 * q + 1] = x.y; }
    const double2 cc = suv[c], ee = suv[c - 1], se = suv[c - uw], ne = suv[c - uw - 1];
    Sigma sg;
    {
      const double2 a = ssig[0 * nT + p], b = ssig[1 * nT + p], m0 = ssig[2 * nT + p], m1 = ssig[3 * nT + p], s0 = ssig[4 * nT + p], s1 = ssig[5 * nT + p];
      sg.p[0] = a.x; sg.p[1] = a.y; sg.p[2] = b.x; sg.p[3] = b.y;
      sg.m[0] = m0.x; sg.m[1] = m0.y; sg.m[2] = m1.x; sg.m[3] = m1.y;
      sg.s12[0] = s0.x; sg.s12[1] = s0.y; sg.s12[2] = s1.x; sg.s12[3] = s1.y;
    }
    double str[8];
    stress_point<true>(cc.x, cc.y, ee.x, ee.y, se.x, se.y, ne.x, ne.y, tv[T_DXT], tv[T_DYT], tv[T_DXHY], tv[T_DYHX], tv[T_CXP], tv[T_CYP], tv[T_CXM],
                       tv[T_CYM], tv[T_DMIN], tv[T_STRENGTH], k, sg, str);
    ssig[0 * nT + p] = mk2(sg.p[0], sg.p[1]); ssig[1 * nT + p] = mk2(sg.p[2], sg.p[3]);
    ssig[2 * nT + p] = mk2(sg.m[0], sg.m[1]); ssig[3 * nT + p] = mk2(sg.m[2], sg.m[3]);
    ssig[4 * nT + p] = mk2(sg.s12[0], sg.s12[1]); ssig[5 * nT + p] = mk2(sg.s12[2], sg.s12[3]);
    // paired by the U point that reads them: (str1,str5) at (i,j), (str2,str7) at (i+1,j), (str3,str6) at (i,j+1), (str4,str8) at (i+1,j+1)
    sstr[0 * nT + t] = mk2(str[0], str[4]); sstr[1 * nT + t] = mk2(str[1], str[6]);
    sstr[2 * nT + t] = mk2(str[2], str[5]); sstr[3 * nT + t] = mk2(str[3], str[7]);
  };

  // one edge value into the low-latency slots of the neighbour GPUs that hold (i,j) as a ghost cell; tagged for subcycle ksub
  auto push_ll = [&](int i, int j, double u, double v, int nxt, int ksub) {
    const unsigned long long tag = ((ebase + (unsigned long long)ksub + 1ULL) & 0xffffffffULL) << 32;
    const int e = edge_index(d, 0, i, j);
    for (int q = px.push_start[e]; q < px.push_start[e + 1]; ++q) {
      const int pr = px.push_peer[q] & 0xff;
      unsigned long long *sl = px.peer_ll[pr] + ((size_t)((ksub + 1) % P2P_LL_SLOTS) * px.peer_ring[pr] + px.push_ll[q]) * 4;
      const unsigned long long ub = (unsigned long long)__double_as_longlong(u), vb = (unsigned long long)__double_as_longlong(v);
      st_relaxed_sys(sl + 0, (ub & 0xffffffffULL) | tag);
      st_relaxed_sys(sl + 1, (ub >> 32) | tag);
      st_relaxed_sys(sl + 2, (vb & 0xffffffffULL) | tag);
      st_relaxed_sys(sl + 3, (vb >> 32) | tag);
    }
  };

  // N U points of this thread at once (N = 2: both slots in one instruction stream; called with at least one of them active)
  auto advance = [&](auto nconst, const unsigned (&ws)[decltype(nconst)::value], const bool (&act)[decltype(nconst)::value],
                     const double (&pre)[decltype(nconst)::value][U_COUNT - KU + 1], bool publish_edge, int nxt, bool last, int ksub) {
    constexpr int N = decltype(nconst)::value;
    double uold[N], vold[N], uo[U_COUNT][N], ui[N], vi[N], s[N][8];
    int cs[N], gs[N], is[N], js[N];
#pragma unroll
    for (int n = 0; n < N; ++n) {
      // an inactive slot repeats the thread's active one (same operands, nothing stored): operands of an arbitrary cell -- zeros on
      // land -- would send the square root and both divisions through their out-of-range calls every subcycle
      const unsigned w = act[n] ? ws[n] : ws[N - 1 - n];
      const int t0 = w & 1023, li = (w >> 10) & 63, lj = (w >> 16) & 63, p = w >> 22;
      const int t = lj * tw + li;
      (void)t0;
      cs[n] = (lj + 1) * uw + li + 1;
      is[n] = i0 + li; js[n] = j0 + lj;
      gs[n] = at(d, is[n], js[n]);
#pragma unroll
      for (int q = KU; q < U_COUNT; ++q) uo[q][n] = act[n] ? pre[n][q - KU] : pre[N - 1 - n][q - KU];
      uo[U_CV][n] = scv[p];
#pragma unroll
      for (int q = 0; q < (KU - 1) / 2; ++q) { const double2 x = sU[q * nU + p]; uo[1 + 2 * q][n] = x.x; uo[2 + 2 * q][n] = x.y; }
      const double2 uv = suv[cs[n]];
      uold[n] = uv.x; vold[n] = uv.y;
      const double2 a = sstr[0 * nT + t], b = sstr[1 * nT + t + 1], c2 = sstr[2 * nT + t + tw], e = sstr[3 * nT + t + tw + 1];
      s[n][0] = a.x; s[n][1] = b.x; s[n][2] = c2.x; s[n][3] = e.x;   // str1(i,j) str2(i+1,j) str3(i,j+1) str4(i+1,j+1)
      s[n][4] = a.y; s[n][5] = c2.y; s[n][6] = b.y; s[n][7] = e.y;   // str5(i,j) str6(i,j+1) str7(i+1,j) str8(i+1,j+1)
      ui[n] = 0.0; vi[n] = 0.0;
      if (k.revp != 0.0 || uold[n] == 0.0 || vold[n] == 0.0) { ui[n] = d.uinit[gs[n]]; vi[n] = d.vinit[gs[n]]; }  // see load_uin (evp_dom.cuh)
    }
    UOut o[N];
    stepu_cv_n<N>(uold, vold, uo[U_CV], uo[U_UOCN], uo[U_VOCN], uo[U_WATERX], uo[U_WATERY], uo[U_FORCEX], uo[U_FORCEY], uo[U_UMASSDTI], uo[U_FM],
                  uo[U_UAREAR], uo[U_TBU], ui, vi, s, k, o);
#pragma unroll
    for (int n = 0; n < N; ++n) {
      if (!act[n]) continue;
      suv[cs[n]] = mk2(o[n].u, o[n].v);
      if (publish_edge || last) {
        double *Un = d.u[nxt];
        double *Vn = d.v[nxt];
        const int i = is[n], j = js[n], g = gs[n];
        __stcg(Un + g, o[n].u);
        __stcg(Vn + g, o[n].v);
        int ig = -1, jg = -1;
        if (d.wrap_ew) ig = (i == 1) ? d.nx + 1 : (i == d.nx ? 0 : -1);
        if (d.wrap_ns) jg = (j == 1) ? d.ny + 1 : (j == d.ny ? 0 : -1);
        if (ig >= 0) { __stcg(Un + at(d, ig, j), o[n].u); __stcg(Vn + at(d, ig, j), o[n].v); }
        if (jg >= 0) { __stcg(Un + at(d, i, jg), o[n].u); __stcg(Vn + at(d, i, jg), o[n].v); }
        if (ig >= 0 && jg >= 0) { __stcg(Un + at(d, ig, jg), o[n].u); __stcg(Vn + at(d, ig, jg), o[n].v); }
        // a 1-wide interior aliases both ghosts (store_uv)
        if (d.wrap_ew && d.nx == 1) { __stcg(Un + at(d, 0, j), o[n].u); __stcg(Vn + at(d, 0, j), o[n].v); }
        if (d.wrap_ns && d.ny == 1) { __stcg(Un + at(d, i, 0), o[n].u); __stcg(Vn + at(d, i, 0), o[n].v); }
        if (P2P && publish_edge && is_push_point(d, 0, i, j)) {
          // a ghost cell of up to three other sub-domains: stored there over NVLink right away (push table: evp_halo.cu)
          const int e = edge_index(d, 0, i, j);
          if (last) {
            // the loop's last values go into the neighbours' arrays themselves (what their download and their next loop read);
            // handed over with the counter + fence + flag below
            for (int q = px.push_start[e]; q < px.push_start[e + 1]; ++q) {
              const int pr = px.push_peer[q] & 0xff, dst = px.push_dst[q];
              px.peer_u[nxt][pr][dst] = o[n].u;
              px.peer_v[nxt][pr][dst] = o[n].v;
            }
          } else {
            push_ll(i, j, o[n].u, o[n].v, nxt, ksub);  // inside the loop: self-validating words, nothing else
          }
        }
      }
      if (last) {
        // only the last subcycle's values survive (ice_dyn_shared.F90:948-965; calc_diag_1d, ice_dyn_core1d.F90:607)
        const int g = gs[n];
        d.strintx[g] = o[n].strintx;
        d.strinty[g] = o[n].strinty;
        d.taubx[g] = o[n].taubx;
        d.tauby[g] = o[n].tauby;
      }
    }
  };

  // ---------------- the subcycle loop ------------------------------------------------------------------------------------
  for (int ksub = 0; ksub < pp.ndte; ++ksub) {
    const int cur = ksub & 1, nxt = cur ^ 1;
    const bool last = (ksub == pp.ndte - 1);
    long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
    if (dbg) tk0 = clk();
    if (warp == 0) mark(ksub, 0);

    // ---- phase A: relax my T cells, str -> shared ----------------------------------------------------------------------
    // The tile-edge T cells (slot 1 of the edge warps) are the only readers of the ring.  It is refreshed from copy `cur` -- once
    // every neighbour has published subcycle ksub-1 -- by the light warps at the top of the phase (they have no second cell), or,
    // when the tile has none, by the edge warps themselves before their second cell.
    auto refresh_ring = [&]() {
      long long tw0 = 0;
      if (dbg) tw0 = clk();
      if (warp == rw0) {  // one warp polls (lane q watches neighbour q), the other refreshing warps wait for it at a named barrier
        if (nb >= 0) wait_progress(pp.progress + PERSIST_CTR_STRIDE * nb, nb_per * (unsigned)ksub, pp.err);
        syncwarp();
        mark(ksub, 1);
      }
      if (nrw > 1) bar_sync_n(2, 32 * nrw);
      const double *U = d.u[cur];
      const double *V = d.v[cur];
      const int nedge = 2 * (ebx + 2) + 2 * eby;
      for (int e = (warp - rw0) * 32 + lane; e < nedge; e += 32 * nrw) {
        int li, lj;
        if (e < ebx + 2) { li = e; lj = 0; }
        else if (e < 2 * (ebx + 2)) { li = e - (ebx + 2); lj = eby + 1; }
        else if (e < 2 * (ebx + 2) + eby) { li = 0; lj = 1 + e - 2 * (ebx + 2); }
        else { li = ebx + 1; lj = 1 + e - 2 * (ebx + 2) - eby; }
        const int gi = i0 - 1 + li, gj = j0 - 1 + lj;
        if (P2P && rank_edge && (gi == 0 || gi == d.nx + 1 || gj == 0 || gj == d.ny + 1)) {
          const int r = ring_index(d.nx, d.ny, gi, gj);
          if (px.ll_fed[r]) {
            // fed by a neighbour GPU: spin on the slot of parity `cur` until all four words carry this subcycle's tag
            const unsigned long long *sl = px.my_ll + ((size_t)(ksub % P2P_LL_SLOTS) * px.my_ring + r) * 4;
            const unsigned long long want = (ebase + (unsigned long long)ksub) & 0xffffffffULL;
            unsigned long long w0, w1, w2, w3;
            const unsigned long long t0 = gtime();
            for (;;) {
              w0 = ld_relaxed_sys(sl + 0); w1 = ld_relaxed_sys(sl + 1); w2 = ld_relaxed_sys(sl + 2); w3 = ld_relaxed_sys(sl + 3);
              if ((w0 >> 32) == want && (w1 >> 32) == want && (w2 >> 32) == want && (w3 >> 32) == want) break;
              if (gtime() - t0 > g_wait_timeout_ns) { atomicExch(px.err, 1); break; }
            }
            suv[lj * uw + li] = mk2(__longlong_as_double((long long)((w0 & 0xffffffffULL) | (w1 << 32))),
                                    __longlong_as_double((long long)((w2 & 0xffffffffULL) | (w3 << 32))));
            continue;
          }
        }
        const int g = at(d, gi, gj);
        suv[lj * uw + li] = mk2(ld_cg_f64(U + g), ld_cg_f64(V + g));  // written by another SM: served by L2, never a stale L1 line
      }
      if (dbg) acc[1] += clk() - tw0;
      if (warp == rw0) mark(ksub, 2);
    };
    if (ksub > 0 && nlw && refresh_warp) {
      refresh_ring();
      bar_arrive_n(1, 32 * (nrw + ewT));
    }
    if (flags & 1u) relax(wT[0]);
    if (ksub > 0 && edge_warp) {
      if (!nlw) refresh_ring();
      bar_sync_n(1, nlw ? 32 * (nrw + ewT) : 32 * ewT);
    }
    if (flags & 2u) relax(wT[1]);
    // momentum operands that are not on chip: requested HERE (the asm's memory clobber keeps the loads below the stores above), so the
    // L2 round trip overlaps with the wait at the barrier instead of heading the momentum phase
    double upre[SL][U_COUNT - KU + 1];
#pragma unroll
    for (int r = 0; r < SL; ++r) {
#pragma unroll
      for (int q = 0; q < U_COUNT - KU + 1; ++q) upre[r][q] = 0.0;
      if (flags & (16u << r)) {
        const int g = at(d, i0 + (int)((wU[r] >> 10) & 63), j0 + (int)((wU[r] >> 16) & 63));
#pragma unroll
        for (int q = KU; q < U_COUNT; ++q) upre[r][q - KU] = ld_nc_f64_pinned(gU[q] + g);
      }
    }
    if (dbg) tk1 = clk();
    __syncthreads();
    if (dbg) tk2 = clk();
    if (warp == 0) mark(ksub, 3);
    if (warp == NW - 1) mark(ksub, 6);

    // ---- phase C: advance my U points -----------------------------------------------------------------------------------
    if (warp < ewU) {
      // slot 0 of the first ewU warps: the points other tiles (or wrapped ghost copies) read.  Publish: the stores of this warp's
      // lanes are ordered before lane 0's fence by the warp barrier, the fence makes them visible at gpu scope before the counter moves
      const unsigned w1[1] = {wU[0]};
      const bool a1[1] = {(flags & 16u) != 0};
      double p1[1][U_COUNT - KU + 1], p2[1][U_COUNT - KU + 1];
#pragma unroll
      for (int q = 0; q < U_COUNT - KU + 1; ++q) { p1[0][q] = upre[0][q]; p2[0][q] = upre[1][q]; }
      if (P2P && rank_edge && ksub == 0) {
        // nothing may be stored into a neighbour GPU before it has entered this loop (p2p_start_kernel over there)
        if (lane < px.npeers) wait_flag(px.my_flags + px.peer_rank[lane], ebase, px.err);
        syncwarp();
      }
      if (a1[0]) advance(std::integral_constant<int, 1>{}, w1, a1, p1, true, nxt, last, ksub);
      if (P2P && rank_edge && !last && !a1[0] && wU[0] != PERSIST_NONE_W) {
        // a point off the ice never changes, but the neighbour GPU's ring refresh waits for EVERY ghost cell it is fed: send what is there
        const int li = (wU[0] >> 10) & 63, lj = (wU[0] >> 16) & 63;
        if (is_push_point(d, 0, i0 + li, j0 + lj)) {
          const double2 uv = suv[(lj + 1) * uw + li + 1];
          push_ll(i0 + li, j0 + lj, uv.x, uv.y, nxt, ksub);
        }
      }
      syncwarp();
      if (lane == 0) publish_progress(pp.progress + PERSIST_CTR_STRIDE * tile);
      if (P2P && rank_edge && last && lane == 0) {
        // hand-over of the LAST subcycle's plain stores: every publishing warp of every edge tile counts itself done with gpu-scope
        // ordering; the one that arrives last issues the ONE system-scope fence -- cumulative over the NVLink stores of all the warps it
        // has synchronised with through the counter -- and raises the peers' flags to epoch + ndte, which p2p_finish_kernel waits for
        __threadfence();
        const unsigned long long old = atomicAdd(px.done_ctr, 1ULL);
        if (old + 1 == (unsigned long long)pp.n_sig) {
          __threadfence_system();
          for (int q = 0; q < px.npeers; ++q) st_relaxed_sys(px.peer_flag[q], ebase + (unsigned long long)ksub + 1ULL);
        }
      }
      if (warp == 0) mark(ksub, 4);
      const unsigned w2[1] = {wU[1]};
      const bool a2[1] = {(flags & 32u) != 0};
      if (a2[0]) advance(std::integral_constant<int, 1>{}, w2, a2, p2, false, nxt, last, ksub);
    } else if (flags & 48u) {
      const unsigned w2[2] = {wU[0], wU[1]};
      const bool a2[2] = {(flags & 16u) != 0, (flags & 32u) != 0};
      advance(std::integral_constant<int, 2>{}, w2, a2, upre, false, nxt, last, ksub);
    }
    if (dbg) tk3 = clk();
    __syncthreads();
    if (warp == 0) mark(ksub, 5);
    if (dbg) { acc[0] += tk1 - tk0; acc[2] += tk2 - tk1; acc[3] += tk3 - tk2; acc[4] += clk() - tk3; }
  }
  if (dbg && lane == 0) {
#pragma unroll
    for (int q = 0; q < 5; ++q) pp.dbg[((size_t)tile * NW + warp) * 5 + q] = acc[q];
  }

  // ---------------- epilogue: the carried stress state goes back to copy (ndte & 1) ------------------------------------------
  const int fin = pp.ndte & 1;
#pragma unroll
  for (int r = 0; r < SL; ++r)
    if (flags & (4u << r)) {
      const int li = (wT[r] >> 10) & 63, lj = (wT[r] >> 16) & 63, p = wT[r] >> 22;
      const int g = at(d, i0 + li, j0 + lj);
#pragma unroll
      for (int q = 0; q < 6; ++q) { const double2 x = ssig[q * nT + p]; d.sig[fin][2 * q][g] = x.x; d.sig[fin][2 * q + 1][g] = x.y; }
    }
}

#ifndef EVP_HOST_EMU  // launcher: not part of the host emulation (tests/emu_persist.cpp)
template <int KT, int KU, bool DBG, bool P2P>
static cudaError_t launch_persist_t(const Dom &d, const KParams &p, const PersistPlan &pp, const P2PParams &px, cudaStream_t s) {
  static bool attr_set = false;
  auto kern = persist_kernel<PERSIST_THREADS, KT, KU, DBG, P2P>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  void *args[] = {(void *)&d, (void *)&p, (void *)&pp, (void *)&px};
  return cudaLaunchCooperativeKernel((const void *)kern, dim3(pp.ntx * pp.nty), dim3(PERSIST_THREADS), args, pp.smem_bytes, s);
}
template <int KT, int KU>
static cudaError_t launch_persist_k(const Dom &d, const KParams &p, const PersistPlan &pp, const P2PParams *px, cudaStream_t s) {
  static const P2PParams nop2p{};
  if (px) return launch_persist_t<KT, KU, false, true>(d, p, pp, *px, s);   // (no cycle accounting in the multi-GPU form)
  return pp.dbg ? launch_persist_t<KT, KU, true, false>(d, p, pp, nop2p, s) : launch_persist_t<KT, KU, false, false>(d, p, pp, nop2p, s);
}

// how many CTAs of the instantiation the plan selects fit one SM (0: it cannot be launched at all): a cooperative launch needs every
// tile co-resident, and the planner's own shared-memory arithmetic is checked against the driver's here
cudaError_t persist_ctas_per_sm(const PersistPlan &pp, bool p2p, int *n) {
  *n = 0;
  const void *kern = nullptr;
  if (pp.kT == 10 && pp.kU == 11) kern = p2p ? (const void *)persist_kernel<PERSIST_THREADS, 10, 11, false, true> : (const void *)persist_kernel<PERSIST_THREADS, 10, 11, false, false>;
  else if (pp.kT == 6 && pp.kU == 3) kern = p2p ? (const void *)persist_kernel<PERSIST_THREADS, 6, 3, false, true> : (const void *)persist_kernel<PERSIST_THREADS, 6, 3, false, false>;
  else return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, kern, PERSIST_THREADS, pp.smem_bytes);
}

cudaError_t set_wait_timeout_persist(unsigned long long ns) { return cudaMemcpyToSymbol(g_wait_timeout_ns, &ns, sizeof ns); }

// px: null on a single rank, else the in-kernel NVLink halo parameters (evp_halo.cu)
cudaError_t launch_persist(const Dom &d, const KParams &p, const PersistPlan &pp, const P2PParams *px, cudaStream_t s) {
  if (pp.nthreads != PERSIST_THREADS) return cudaErrorInvalidValue;
  if (pp.kT == 10 && pp.kU == 11) return launch_persist_k<10, 11>(d, p, pp, px, s);
  if (pp.kT == 6 && pp.kU == 3) return launch_persist_k<6, 3>(d, p, pp, px, s);
  return cudaErrorInvalidValue;
}
#endif  // EVP_HOST_EMU

}  // namespace EVP_NS
}  // namespace evp
