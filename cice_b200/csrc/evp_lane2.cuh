// evp_lane2.cuh -- KERNEL_FUSED with TWO LANES PER T CELL (lane2_relax / lane2_str of evp_math.cuh).
//
// Why: at gx1 the fused kernel is bound by the length of the dependent fp64 chain of a T cell (~5.7 k cycles at 14 resident
// warps per SM), not by a pipe or by memory.  Here the chain of a cell is cut in two: the north lane
// relaxes the corners NE and NW, the south lane SE and SW, the lanes swap their six stresses, and each forms the four
// `str` terms of the U points on its own row.  Patch geometry, ownership rule, ping-pong copies, on-rank wrap stores and
// the programmatic-dependent-launch protocol are those of fused_kernel; the momentum step runs on the first PX*PY
// threads, one per U point.  Bit-identical to the one-thread-per-cell form (tests/test_host_math.py, tests/test_emu_bgrid.py).
//
// Included by evp_kernels.cu inside namespace evp::EVP_NS.  Everything CUDA-specific it touches is a qualifier,
// __syncthreads, one warp shuffle, one named barrier and the two PDL calls, so the same text also compiles for the
// host emulation in tests/emu_bgrid.cpp.
//
//   MAP 0: lanes 2q and 2q+1 of a warp share T cell q; stresses swapped with __shfl_xor_sync
//   MAP 2: MAP 1 with the two roles as two specialised code paths (`north` a compile-time constant in each)
//   MAP 1: threads [0, PX*PY) are the north lanes, [PX*PY, 2*PX*PY) the south lanes of the same cells (`north` is
//          warp-uniform, loads fully coalesced); stresses swapped through shared memory behind a 64-thread named
//          barrier per row (needs PX == 32: one warp per row and role)
#pragma once

//   SPT:   every operand of the lane requested before the ice mask is known (the speculative form of fused_kernel: one L2
//          round trip instead of mask -> operands), static ones even before the grid dependency resolves
//   P2P:   the in-kernel NVLink halo of fused_kernel<..., P2P = true>: (tbx, tby) come from the edge-first tile
//          table, boundary U points are also stored into the neighbour GPUs' ghost cells
// the stress half of the kernel for one lane.  `north` may be a compile-time constant at the call site (MAP 2): the function is
// always inlined, so the selections on it fold away
template <int PX, int PY, bool IL, int MAP, bool SPT, int NSX>
__device__ __forceinline__ void lane2_stress(const Dom &d, const KParams &k, int cur, const bool north, int cx, int cy, int i, int j,
                                             bool inT, int c, double (&sstr)[8][PY][PX], double (&sx)[NSX][PY][PX]) {
  const int nxt = cur ^ 1;
  if (SPT) {
    // corner numbering 0 NE, 1 NW, 2 SW, 3 SE: the lane's E corner is NE or SE, its W corner NW or SW
    const int qE = north ? NE : SE, qW = north ? NW : SW;
    const unsigned mT = ld_nc_u8(d.maskT + c);
    const double dxT_ = ld_nc_f64(d.dxT + c), dyT_ = ld_nc_f64(d.dyT + c), dxhy = ld_nc_f64(d.dxhy + c), dyhx = ld_nc_f64(d.dyhx + c);
    const double cxp = ld_nc_f64(d.cxp + c), cyp = ld_nc_f64(d.cyp + c), cxm = ld_nc_f64(d.cxm + c), cym = ld_nc_f64(d.cym + c);
    const double dmin = ld_nc_f64(d.DminTarea + c), strength = ld_nc_f64(d.strength + c);
#if EVP_USE_PDL
    cudaGridDependencySynchronize();
#endif
    const int ca = north ? c : c - d.ld, cb = north ? c - d.ld : c;
    const double *U = d.u[cur], *V = d.v[cur];
    const double ua_c = ld_f64(U + ca), va_c = ld_f64(V + ca), ua_e = ld_f64(U + ca - 1), va_e = ld_f64(V + ca - 1);
    const double ub_c = ld_f64(U + cb), vb_c = ld_f64(V + cb), ub_e = ld_f64(U + cb - 1), vb_e = ld_f64(V + cb - 1);
    Half own = {ld_f64(d.sig[cur][qE] + c), ld_f64(d.sig[cur][qW] + c), ld_f64(d.sig[cur][4 + qE] + c), ld_f64(d.sig[cur][4 + qW] + c),
                ld_f64(d.sig[cur][8 + qE] + c), ld_f64(d.sig[cur][8 + qW] + c)};
    const bool active = inT && mT;
    if (active) {
      lane2_relax<IL>(north, ua_c, va_c, ua_e, va_e, ub_c, vb_c, ub_e, vb_e, dxT_, dyT_, cxp, cyp, cxm, cym, dmin, strength, k, own);
      const bool ownT = (cx < PX - 1 || i == d.nx + 1) && (cy < PY - 1 || j == d.ny + 1);
      if (ownT) {
        d.sig[nxt][qE][c] = own.pE; d.sig[nxt][qW][c] = own.pW;
        d.sig[nxt][4 + qE][c] = own.mE; d.sig[nxt][4 + qW][c] = own.mW;
        d.sig[nxt][8 + qE][c] = own.sE; d.sig[nxt][8 + qW][c] = own.sW;
      }
    } else {
      own = Half{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    }
    Half oth;
    if (MAP == 0) {
      oth.pE = __shfl_xor_sync(0xffffffffu, own.pE, 1); oth.pW = __shfl_xor_sync(0xffffffffu, own.pW, 1);
      oth.mE = __shfl_xor_sync(0xffffffffu, own.mE, 1); oth.mW = __shfl_xor_sync(0xffffffffu, own.mW, 1);
      oth.sE = __shfl_xor_sync(0xffffffffu, own.sE, 1); oth.sW = __shfl_xor_sync(0xffffffffu, own.sW, 1);
    } else {
      const int mine = north ? 0 : 6, theirs = north ? 6 : 0;
      sx[mine + 0][cy][cx] = own.pE; sx[mine + 1][cy][cx] = own.pW; sx[mine + 2][cy][cx] = own.mE;
      sx[mine + 3][cy][cx] = own.mW; sx[mine + 4][cy][cx] = own.sE; sx[mine + 5][cy][cx] = own.sW;
      bar_sync64(cy + 1);
      oth.pE = sx[theirs + 0][cy][cx]; oth.pW = sx[theirs + 1][cy][cx]; oth.mE = sx[theirs + 2][cy][cx];
      oth.mW = sx[theirs + 3][cy][cx]; oth.sE = sx[theirs + 4][cy][cx]; oth.sW = sx[theirs + 5][cy][cx];
    }
    double out[4] = {0.0, 0.0, 0.0, 0.0};
    if (active) lane2_str(north, own, oth, dxT_, dyT_, dxhy, dyhx, out);
    sstr[north ? 0 : 2][cy][cx] = out[0];
    sstr[north ? 1 : 3][cy][cx] = out[1];
    sstr[north ? 4 : 5][cy][cx] = out[2];
    sstr[north ? 6 : 7][cy][cx] = out[3];
  } else {
#if EVP_USE_PDL
  cudaGridDependencySynchronize();
#endif
  const bool active = inT && d.maskT[c];
  Half own = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double dxT = 0.0, dyT = 0.0;
  if (active) {
    // the lane's own velocity row a and the other row b (evp_math.cuh: lane2_relax)
    const int ca = north ? c : c - d.ld, cb = north ? c - d.ld : c;
    const double *__restrict__ U = d.u[cur];
    const double *__restrict__ V = d.v[cur];
    // corner numbering 0 NE, 1 NW, 2 SW, 3 SE: the lane's E corner is NE or SE, its W corner NW or SW
    const int qE = north ? NE : SE, qW = north ? NW : SW;
    own.pE = d.sig[cur][qE][c]; own.pW = d.sig[cur][qW][c];
    own.mE = d.sig[cur][4 + qE][c]; own.mW = d.sig[cur][4 + qW][c];
    own.sE = d.sig[cur][8 + qE][c]; own.sW = d.sig[cur][8 + qW][c];
    dxT = d.dxT[c]; dyT = d.dyT[c];
    lane2_relax<IL>(north, U[ca], V[ca], U[ca - 1], V[ca - 1], U[cb], V[cb], U[cb - 1], V[cb - 1], dxT, dyT, d.cxp[c], d.cyp[c],
                    d.cxm[c], d.cym[c], d.DminTarea[c], d.strength[c], k, own);
    // each T cell is stored by exactly one CTA: the one that holds it off its E/N overlap edge
    const bool ownT = (cx < PX - 1 || i == d.nx + 1) && (cy < PY - 1 || j == d.ny + 1);
    if (ownT) {
      d.sig[nxt][qE][c] = own.pE; d.sig[nxt][qW][c] = own.pW;
      d.sig[nxt][4 + qE][c] = own.mE; d.sig[nxt][4 + qW][c] = own.mW;
      d.sig[nxt][8 + qE][c] = own.sE; d.sig[nxt][8 + qW][c] = own.sW;
    }
  }
  // the swap is unconditional (every lane of the warp / both warps of the row take part); off the ice the lanes swap zeros
  Half oth;
  if (MAP == 0) {
    oth.pE = __shfl_xor_sync(0xffffffffu, own.pE, 1); oth.pW = __shfl_xor_sync(0xffffffffu, own.pW, 1);
    oth.mE = __shfl_xor_sync(0xffffffffu, own.mE, 1); oth.mW = __shfl_xor_sync(0xffffffffu, own.mW, 1);
    oth.sE = __shfl_xor_sync(0xffffffffu, own.sE, 1); oth.sW = __shfl_xor_sync(0xffffffffu, own.sW, 1);
  } else {
    const int mine = north ? 0 : 6, theirs = north ? 6 : 0;
    sx[mine + 0][cy][cx] = own.pE; sx[mine + 1][cy][cx] = own.pW; sx[mine + 2][cy][cx] = own.mE;
    sx[mine + 3][cy][cx] = own.mW; sx[mine + 4][cy][cx] = own.sE; sx[mine + 5][cy][cx] = own.sW;
    bar_sync64(cy + 1);  // named barrier over the two warps of this patch row (evp_ptx.cuh)
    oth.pE = sx[theirs + 0][cy][cx]; oth.pW = sx[theirs + 1][cy][cx]; oth.mE = sx[theirs + 2][cy][cx];
    oth.mW = sx[theirs + 3][cy][cx]; oth.sE = sx[theirs + 4][cy][cx]; oth.sW = sx[theirs + 5][cy][cx];
  }
  double out[4] = {0.0, 0.0, 0.0, 0.0};
  if (active) lane2_str(north, own, oth, dxT, dyT, d.dxhy[c], d.dyhx[c], out);
  // north: str1 str2 str5 str7, south: str3 str4 str6 str8 (0-based rows of sstr as fused_kernel uses them)
  sstr[north ? 0 : 2][cy][cx] = out[0];
  sstr[north ? 1 : 3][cy][cx] = out[1];
  sstr[north ? 4 : 5][cy][cx] = out[2];
  sstr[north ? 6 : 7][cy][cx] = out[3];
  }  // !SPT
}

template <int PX, int PY, bool IL, int MAP, bool SPT, bool P2P>
__device__ __forceinline__ void lane2_body(const Dom &d, const KParams &k, int cur, int last, int tbx, int tby, const P2PParams *pp) {
  static_assert(MAP == 0 || (PX == 32 && PY <= 15), "MAP 1, 2 pair one north warp with one south warp per row");
  static_assert(MAP != 0 || PX % 16 == 0, "MAP 0 keeps the 16 cells of a warp on one row");
  __shared__ double sstr[8][PY][PX];
  constexpr int NSX = MAP == 0 ? 1 : 12;
  __shared__ double sx[NSX][PY][PX];  // MAP 1, 2: [0..5] the north lanes' stresses, [6..11] the south lanes'
  const int t = threadIdx.x;
  const bool north = MAP == 0 ? !(t & 1) : t < PX * PY;
  const int cell = MAP == 0 ? (t >> 1) : (north ? t : t - PX * PY);
  const int cx = cell % PX, cy = cell / PX;
  const int i = 1 + tbx * (PX - 1) + cx;  // T cell of this lane pair
  const int j = 1 + tby * (PY - 1) + cy;
  const int nxt = cur ^ 1;
  const bool inT = (i <= d.nx + 1) && (j <= d.ny + 1);
  const int c = at(d, inT ? i : 1, inT ? j : 1);
  if (MAP == 2) {
    // warp-uniform roles as MAP 1, and each role runs its own specialised copy of the stress code (no selections on `north`)
    if (north) lane2_stress<PX, PY, IL, 1, SPT, NSX>(d, k, cur, true, cx, cy, i, j, inT, c, sstr, sx);
    else lane2_stress<PX, PY, IL, 1, SPT, NSX>(d, k, cur, false, cx, cy, i, j, inT, c, sstr, sx);
  } else {
    lane2_stress<PX, PY, IL, MAP, SPT, NSX>(d, k, cur, north, cx, cy, i, j, inT, c, sstr, sx);
  }
  __syncthreads();

  if (t < PX * PY) {
    const int tx = t % PX, ty = t / PX;
    const int iu = 1 + tbx * (PX - 1) + tx, ju = 1 + tby * (PY - 1) + ty;
    if (tx < PX - 1 && ty < PY - 1 && iu <= d.nx && ju <= d.ny) {
      const int cu = at(d, iu, ju);
      if (d.maskU[cu]) {
        double uin[16];
        load_uin(d, k, cur, cu, uin);
        const UOut o = stepu_point<IL>(uin[0], uin[1], uin[2], uin[3], uin[4], uin[5], uin[6], uin[7], uin[8], uin[9], uin[10], uin[11],
                                       uin[12], uin[13], uin[14], uin[15], sstr[0][ty][tx], sstr[1][ty][tx + 1], sstr[2][ty + 1][tx],
                                       sstr[3][ty + 1][tx + 1], sstr[4][ty][tx], sstr[5][ty + 1][tx], sstr[6][ty][tx + 1],
                                       sstr[7][ty + 1][tx + 1], k);
        store_uv(d, d.u[nxt], d.v[nxt], iu, ju, o.u, o.v);
        if (last) {  // see fused_kernel
          d.strintx[cu] = o.strintx;
          d.strinty[cu] = o.strinty;
          d.taubx[cu] = o.taubx;
          d.tauby[cu] = o.tauby;
        }
        if (P2P && (iu == 1 || iu == d.nx || ju == 1 || ju == d.ny)) {  // a ghost cell of up to three neighbour GPUs
          const int e = edge_index(d, iu, ju);
          for (int q = pp->push_start[e]; q < pp->push_start[e + 1]; ++q) {
            const int pr = pp->push_peer[q];
            const int dst = pp->push_dst[q];
            pp->peer_u[nxt][pr][dst] = o.u;
            pp->peer_v[nxt][pr][dst] = o.v;
          }
        }
      }
    }
  }
}

template <int PX, int PY, int MINB, bool IL, int MAP, bool SPT = false>
__global__ void __launch_bounds__(2 * PX *PY, MINB) fused2_kernel(const __grid_constant__ Dom d, const __grid_constant__ KParams k,
                                                                  int cur, int flags) {
#if EVP_USE_PDL
  if (flags & 2) cudaTriggerProgrammaticLaunchCompletion();  // see fused_kernel
#endif
  lane2_body<PX, PY, IL, MAP, SPT, false>(d, k, cur, flags & 1, blockIdx.x, blockIdx.y, nullptr);
}

// The in-kernel-halo form: the protocol of fused_kernel<..., P2P = true> around lane2_body (edge tiles first, wait for the
// neighbours' epoch flags before touching ghost cells, count the edge CTAs, the last one fences at system scope and raises
// the peers' flags).  The tile table is read from constant memory when `ctiles` is set (evp_b200 P2P_CONST_TILES).
template <int PX, int PY, int MINB, bool IL, int MAP>
__global__ void __launch_bounds__(2 * PX *PY, MINB) fused2_p2p_kernel(const __grid_constant__ Dom d, const __grid_constant__ KParams k,
                                                                      int cur, const __grid_constant__ P2PParams pp, int ksub, int flags,
                                                                      int ctiles) {
#if EVP_USE_PDL
  if (flags & 2) cudaTriggerProgrammaticLaunchCompletion();
#endif
  const int t = threadIdx.x, b = blockIdx.x;
  const int tile = ctiles ? c_tile_order[b] : pp.tile_order[b];
  const bool edge_tile = b < pp.n_edge_tiles;
  unsigned long long base = 0;
  if (edge_tile) {
    // the ghost ring of copy `cur` was written by the neighbour GPUs during their previous subcycle
    base = *pp.epoch_base;
    if (t < pp.npeers) wait_flag(pp.my_flags + pp.peer_rank[t], base + (unsigned long long)ksub, pp.err);
    __syncthreads();
  }
  lane2_body<PX, PY, IL, MAP, false, true>(d, k, cur, flags & 1, tile & 0xffff, tile >> 16, &pp);
  if (edge_tile) {
    __syncthreads();
    if (t == 0) {
      __threadfence();
      const unsigned long long old = atomicAdd(pp.done_ctr, 1ULL);
      if (old + 1 == (unsigned long long)pp.n_edge_tiles * (unsigned long long)(ksub + 1)) {
        __threadfence_system();
        for (int q = 0; q < pp.npeers; ++q) st_relaxed_sys(pp.peer_flag[q], base + (unsigned long long)ksub + 1ULL);
      }
    }
  }
}
