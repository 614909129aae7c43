// evp_tstream.cu -- KERNEL_TSTREAM: the EVP subcycle for sub-domains that stream from HBM (arrays far larger than L2), with every
// operand staged through shared memory by the Tensor Memory Accelerator.
//
// Compiled twice like the other kernel units (-DEVP_NS=exact -fmad=false / -DEVP_NS=fast).
//
// Why another kernel form.  The launch-per-subcycle kernel (evp_kernels.cu: fused_kernel) advances a cell in the same time whether
// its operands come from L2 or from HBM (9.2 us x 70 at 3600x2400 = 647 us against 641-671 measured): it is bound by the latency its
// warps see -- request, wait, compute, CTA barrier, request, wait -- and by the 18 % of T cells its overlapping 32 x 8 patches relax
// twice, not by bandwidth.  Here the requests are taken off the warps altogether:
//   * the sub-domain is cut into column STRIPS of 31 T cells (30 U points; neighbouring strips share one T column) and every strip
//     into SEGMENTS of rows*nb T rows; (strip, segment) items are dealt round robin to one persistent CTA per SM (or two).  A box
//     load must start on a 16-byte boundary of its row (measured: anything else is an illegal instruction, scripts/micro/tma_probe.cu),
//     hence the even strip stride: every fp64 box is the 32 columns xs .. xs+31 with xs = 30*strip, T cells xs+1 .. xs+31 and their
//     west neighbours included; the byte masks start at the 16-column boundary below xs and are 48 wide;
//   * a CTA walks its segment upwards in BLOCKS of `rows` T rows, one thread per T cell.  Everything a block reads -- the 12
//     stresses, (u,v) with the west/south neighbours, dxT, dyT, the two metric arrays the other seven geometry values derive from
//     (evp_math.cuh: derive_geometry, after evp_b200_set_metric), strength, both ice masks and the 12 momentum operands of the U
//     points the block closes: 33 arrays -- is fetched by 33 `cp.async.bulk.tensor.2d` box loads issued by ONE thread into one of
//     two shared-memory stages, a whole block (3 us of arithmetic) ahead of its use; the loads complete on the stage's mbarrier.
//     No thread ever waits for a global load it issued itself, no register holds a value in flight;
//   * the stress divergence terms go from the stress phase to the momentum phase through shared memory (as in fused_kernel), and the
//     top row of every block is CARRIED to the next block of the segment, so a T row is relaxed twice only where two segments meet
//     (1 row in rows*nb) and a T column only where two strips meet (1 in 31): 4 % redundant work instead of 18 %;
//   * reads only copy `cur` of the carried state, writes only copy `cur^1` (plain coalesced stores): no race between CTAs, the
//     launch boundary is the only synchronisation, exactly as for fused_kernel -- and the same bits.
// Shared memory: 2 stages x 94.9 KB + 28 KB at rows = 12 (one CTA of 384 threads per SM, the default), 2 x 48 KB + 16 KB at rows = 6 (two
// CTAs of 192; EVP_B200_TSTREAM_ROWS).  Measured on B200 at 3600x2400 (DESIGN.md sections 3, 4, 8; profiles/r2_tstream_*): 563 us per
// subcycle against 641 for the launch-per-subcycle kernel (605 / 685 once the board sits at its power cap), DRAM traffic 0.99 x the
// algorithmic 3.11 GB per launch.
#include "evp_math.cuh"
#include "evp_dom.cuh"
#include "evp_ptx.cuh"
#include "evp_tma.cuh"
#include <stdio.h>
#include <stdlib.h>

#ifndef EVP_USE_PDL
#define EVP_USE_PDL 1
#endif

#ifndef EVP_NS
#error "compile with -DEVP_NS=exact or -DEVP_NS=fast"
#endif

namespace evp {
namespace EVP_NS {

// shared-memory layout of a stage (byte offsets, every box 128-byte aligned) and of the CTA
template <int R>
struct TsL {
  static constexpr int pad(int x) { return (x + 127) / 128 * 128; }
  static constexpr int ROW = TS_W * 8;                         // one box row of 32 doubles: columns xs .. xs+31
  static constexpr int MROW = TS_MW;                           // one box row of 48 mask bytes: columns 16*(xs/16) ..
  static constexpr int SIG = 0;                                // [12][R][32]     stresses, rows jb ..
  static constexpr int UOP = SIG + 12 * R * ROW;               // [12][R][32]     momentum operands of U rows jb-1 .. jb+R-2
  static constexpr int GEO = UOP + 12 * R * ROW;               // [3][R][32]      dxT, dyT, strength
  static constexpr int HTN = GEO + 3 * R * ROW;                // [R+1][32]       HTN rows jb-1 .. jb+R-1
  static constexpr int HTE = HTN + (R + 1) * ROW;              // [R][32]         HTE (west neighbour = previous column of the box)
  static constexpr int UU = HTE + R * ROW;                     // [R+1][32]       u rows jb-1 ..
  static constexpr int VV = UU + (R + 1) * ROW;
  static constexpr int MT = VV + (R + 1) * ROW;                // [R][48] bytes   T ice mask
  static constexpr int MU = MT + pad(R * MROW);                // [R][48] bytes   U ice mask of rows jb-1 ..
  static constexpr int STAGE = MU + pad(R * MROW);
  static constexpr unsigned TX_BYTES = (31u * R + 3u) * ROW + 2u * R * MROW;
  static constexpr int STR = 2 * STAGE;                        // [8][R][32]      str terms of the block
  static constexpr int CARRY = STR + 8 * R * ROW;              // [2][8][32]      ... of the top row of the previous block
  static constexpr int BARS = CARRY + 2 * 8 * ROW;             // one mbarrier per stage
  static constexpr int TOTAL = BARS + 128;
};

// position of a CTA in its sequence of blocks
struct TsIter { int item, b, nb, i0, j0; };
template <int R>
__device__ __forceinline__ void ts_setup(const Dom &d, const TsPlan &ts, TsIter &it) {
  it.b = 0; it.nb = 0; it.i0 = 0; it.j0 = 0;
  if (it.item >= ts.nitems) return;
  const int seg = it.item / ts.nstrips, strip = it.item - seg * ts.nstrips;
  it.i0 = 1 + TS_STRIDE * strip;   // first T column of the strip; its boxes start one column to the west (even)
  it.j0 = 1 + seg * (R * ts.nb - 1);
  const int need = (d.ny + 2 - it.j0 + R - 1) / R;   // blocks that reach T row ny+1
  it.nb = need < ts.nb ? need : ts.nb;
}
template <int R>
__device__ __forceinline__ void ts_next(const Dom &d, const TsPlan &ts, TsIter &it, int stride) {
  if (++it.b >= it.nb) { it.item += stride; ts_setup<R>(d, ts, it); }
}

// the 33 box loads of the block whose first T cell is (i0, jb), all completing on `bar`
template <int R>
__device__ __forceinline__ void ts_issue(const TsMaps &tm, int cur, unsigned char *st, MBar *bar, int i0, int jb) {
  using L = TsL<R>;
  const TmaMap *m = tm.m;
  const int xs = i0 - 1, xm = xs & ~15;   // box starts: fp64 columns (16-byte boundary: xs is even), mask bytes
  tma_load_2d(st + L::MT, m + TS_MAP_MASKT, xm, jb, bar);
  tma_load_2d(st + L::UU, m + TS_MAP_U + cur, xs, jb - 1, bar);
  tma_load_2d(st + L::VV, m + TS_MAP_V + cur, xs, jb - 1, bar);
  for (int q = 0; q < 12; ++q) tma_load_2d(st + L::SIG + q * R * L::ROW, m + TS_MAP_SIG + 12 * cur + q, xs, jb, bar);
  tma_load_2d(st + L::GEO, m + TS_MAP_DXT, xs, jb, bar);
  tma_load_2d(st + L::GEO + R * L::ROW, m + TS_MAP_DYT, xs, jb, bar);
  tma_load_2d(st + L::GEO + 2 * R * L::ROW, m + TS_MAP_STRENGTH, xs, jb, bar);
  tma_load_2d(st + L::HTN, m + TS_MAP_HTN, xs, jb - 1, bar);
  tma_load_2d(st + L::HTE, m + TS_MAP_HTE, xs, jb, bar);
  tma_load_2d(st + L::MU, m + TS_MAP_MASKU, xm, jb - 1, bar);
  for (int q = 0; q < 12; ++q) tma_load_2d(st + L::UOP + q * R * L::ROW, m + TS_MAP_UOP + q, xs, jb - 1, bar);
  mbar_arrive_expect_tx(bar, L::TX_BYTES);
}

template <int R, int MINB>
__global__ void __launch_bounds__(32 * R, MINB) tstream_kernel(const __grid_constant__ Dom d, const __grid_constant__ KParams k,
                                                               const __grid_constant__ TsPlan ts, const __grid_constant__ TsMaps tm,
                                                               int cur, int last) {
  using L = TsL<R>;
  const int t = threadIdx.x, tx = t & 31, ty = t >> 5;
  unsigned char *smem = smem_align128(dyn_smem());
  MBar *bars = (MBar *)(smem + L::BARS);
  double *sstr = (double *)(smem + L::STR);
  double *carry = (double *)(smem + L::CARRY);
  if (t == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
#if EVP_USE_PDL
  cudaGridDependencySynchronize();   // the previous subcycle's stores are what the first box loads fetch
#endif
  const int nxt = cur ^ 1, G = gridDim.x, H = R * ts.nb;
  TsIter ci, pi;
  ci.item = blockIdx.x;
  ts_setup<R>(d, ts, ci);
  pi = ci;
  // Thread 0 issues the box loads.  (An elected lane of warp 0 inside a warp-uniform branch keeps the operands in uniform registers --
  // 5 instead of ~15 instructions per load, no ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop -- and was 4 % SLOWER at 12 rows and
  // 15 % at 6 in the one-box A/B, profiles/r2_tstream_ab.txt: 160 registers instead of 151 and uniform-register spills.)
  const bool issuer = (t == 0);
  // (every thread steps the prefetch iterator; only the box loads themselves are thread 0's)
  for (int s = 0; s < 2; ++s)
    if (pi.item < ts.nitems) {
      if (issuer) ts_issue<R>(tm, cur, smem + s * L::STAGE, &bars[s], pi.i0, pi.j0 + R * pi.b);
      ts_next<R>(d, ts, pi, G);
    }
  for (unsigned q = 0; ci.item < ts.nitems; ++q) {
    const int s = q & 1;
    unsigned char *st = smem + s * L::STAGE;
    const int jb = ci.j0 + R * ci.b;          // first T row of the block
    const int i = ci.i0 + tx;                 // column of this thread's T cell and U point
    mbar_wait(&bars[s], q >> 1, ts.err);

    // ---- stress phase: T cell (i, jb + ty) -------------------------------------------------------------------------------
    const int j = jb + ty;
    const bool inT = (tx < TS_W - 1) && (i <= d.nx + 1) && (j <= d.ny + 1);   // lane 31 has no T cell: the box holds 31 and a west neighbour
    const double *sU = (const double *)(st + L::UU), *sV = (const double *)(st + L::VV);
    const int o = (ty + 1) * TS_W + tx + 1;   // this T cell in a box that starts one row / one column earlier
    const int oc = (tx < TS_W - 1) ? o : o - 1;   // (lane 31 reads inside the box and drops the values)
    const double ucc = sU[oc], vcc = sV[oc], uee = sU[oc - 1], vee = sV[oc - 1];
    const double use_ = sU[oc - TS_W], vse = sV[oc - TS_W], une = sU[oc - TS_W - 1], vne = sV[oc - TS_W - 1];
    const int mo = ty * L::MROW + ((ci.i0 - 1) & 15) + 1 + tx;   // this cell / point in the mask boxes
    double str[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (inT && (st + L::MT)[mo]) {
      const int tc = ty * TS_W + tx + 1;       // ... in a box that starts one column earlier
      const double *sS = (const double *)(st + L::SIG) + tc;
      Sigma sg;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        sg.p[c4] = sS[c4 * R * TS_W];
        sg.m[c4] = sS[(4 + c4) * R * TS_W];
        sg.s12[c4] = sS[(8 + c4) * R * TS_W];
      }
      const double *sG = (const double *)(st + L::GEO) + tc;
      const double dxT = sG[0], dyT = sG[R * TS_W], strength = sG[2 * R * TS_W];
      const double *sN = (const double *)(st + L::HTN), *sE = (const double *)(st + L::HTE);
      double dxhy, dyhx, cxp, cyp, cxm, cym, dmin;
      derive_geometry(sN[tc + TS_W], sN[tc], sE[tc], sE[tc - 1], dxT, dyT, ts.deltamin, dxhy, dyhx, cxp, cyp, cxm, cym, dmin);
      stress_point<true>(ucc, vcc, uee, vee, use_, vse, une, vne, dxT, dyT, dxhy, dyhx, cxp, cyp, cxm, cym, dmin, strength, k, sg, str);
      // a T cell is stored by exactly one CTA: not by the one that holds it as the east column of its strip or as the top row of
      // its segment (the neighbour strip / the next segment recomputes and stores it), unless the sub-domain ends there
      const bool own = (tx < TS_STRIDE || i == d.nx + 1) && (j < ci.j0 + H - 1 || j == d.ny + 1);
      if (own) {
        const int c = at(d, i, j);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          d.sig[nxt][c4][c] = sg.p[c4];
          d.sig[nxt][4 + c4][c] = sg.m[c4];
          d.sig[nxt][8 + c4][c] = sg.s12[c4];
        }
      }
    }
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) sstr[c8 * R * TS_W + t] = str[c8];
    if (ty == R - 1) {
      double *cr = carry + (ci.b & 1) * 8 * TS_W + tx;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) cr[c8 * TS_W] = str[c8];
    }
    __syncthreads();

    // ---- momentum phase: U point (i, jb - 1 + ty); its T cells are rows jb-1+ty (lower) and jb+ty (upper) ------------------
    const int ju = jb - 1 + ty;
    const bool uspot = tx < TS_STRIDE && i <= d.nx && ju <= d.ny && (ty > 0 || ci.b > 0);
    const unsigned mU = (st + L::MU)[mo];
    // (u,v) at the U point are the stress phase's south neighbours.  An off-ice point of the row below a tripole fold travels to
    // the other ping-pong copy like a computed one (see fused_body, evp_kernels.cu)
    const bool doU = uspot && mU, carryf = uspot && !mU && d.fold_top && ju == d.ny;
    double un = use_, vn = vse;
    if (doU) {
      const int c = at(d, i, ju);
      const double *sO = (const double *)(st + L::UOP) + t + 1;
      double uo[12];
#pragma unroll
      for (int c12 = 0; c12 < 12; ++c12) uo[c12] = sO[c12 * R * TS_W];
      double ui = 0.0, vi = 0.0;
      if (k.revp != 0.0 || use_ == 0.0 || vse == 0.0) { ui = d.uinit[c]; vi = d.vinit[c]; }  // see load_uin (evp_dom.cuh)
      const double *lo = (ty > 0) ? sstr + (ty - 1) * TS_W + tx : carry + ((ci.b & 1) ^ 1) * 8 * TS_W + tx;
      const int ls = (ty > 0) ? R * TS_W : TS_W;
      const double *up = sstr + ty * TS_W + tx;
      const int us = R * TS_W;
      const UOut out = stepu_point<true>(use_, vse, uo[0], uo[1], uo[2], uo[3], uo[4], uo[5], uo[6], uo[7], uo[8], uo[9], uo[10], uo[11], ui, vi,
                                         lo[0], lo[ls + 1], up[2 * us], up[3 * us + 1], lo[4 * ls], up[5 * us], lo[6 * ls + 1], up[7 * us + 1], k);
      un = out.u; vn = out.v;
      if (last) {   // diagnostics: only the last subcycle's values survive (calc_diag_1d, ice_dyn_core1d.F90:607)
        d.strintx[c] = out.strintx;
        d.strinty[c] = out.strinty;
        d.taubx[c] = out.taubx;
        d.tauby[c] = out.tauby;
      }
    }
    if (doU || carryf) store_uv(d, d.u[nxt], d.v[nxt], i, ju, un, vn);
    __syncthreads();   // the stage and the str terms have been read by everyone

    // ---- the stage is free: fetch the block after the next one into it -------------------------------------------------------
    if (pi.item < ts.nitems) {
      if (issuer) {
        fence_proxy_async();
        ts_issue<R>(tm, cur, st, &bars[s], pi.i0, pi.j0 + R * pi.b);
      }
      ts_next<R>(d, ts, pi, G);
    }
    ts_next<R>(d, ts, ci, G);
  }
}

// The cut into strips and segments (plain host code, shared with the host emulation).  nb is chosen so that the CTAs' block counts come
// out as even as the round-robin deal allows.
static void tstream_cut(int nx, int ny, int num_sms, int rows, TsPlan *ts) {
  const int minb = (rows == 12) ? 1 : 2;
  const int nstrips = (nx + TS_STRIDE - 1) / TS_STRIDE;
  long best_cost = -1;
  int best_nb = 0;
  for (int nb = 2; nb <= 64; ++nb) {
    const int H = rows * nb, nseg = (ny + H - 2) / (H - 1), nitems = nstrips * nseg;
    const int ctas = nitems < num_sms * minb ? nitems : num_sms * minb;
    // CTA c takes items c, c + ctas, ...; item = seg * nstrips + strip; blocks of an item: those that reach T row ny+1, at most nb
    long worst = 0;
    for (int c = 0; c < ctas; ++c) {
      long blocks = 0;
      for (int it = c; it < nitems; it += ctas) {
        const int j0 = 1 + (it / nstrips) * (H - 1), need = (ny + 2 - j0 + rows - 1) / rows;
        blocks += (need < nb ? need : nb);
      }
      if (blocks > worst) worst = blocks;
    }
    if (best_cost < 0 || worst < best_cost) { best_cost = worst; best_nb = nb; }
    if (H - 1 >= ny) break;   // one segment per strip already
  }
  const int H = rows * best_nb;
  ts->rows = rows; ts->nstrips = nstrips; ts->nb = best_nb; ts->nseg = (ny + H - 2) / (H - 1);
  ts->nitems = ts->nstrips * ts->nseg;
  ts->ctas = ts->nitems < num_sms * minb ? ts->nitems : num_sms * minb;
}

#ifndef EVP_HOST_EMU

template <int R, int MINB>
static cudaError_t launch_tstream_t(const Dom &d, const KParams &p, const TsPlan &ts, int cur, int last, bool pdl, cudaStream_t s) {
  const int smem = TsL<R>::TOTAL + 128;
  cudaError_t e = cudaFuncSetAttribute(tstream_kernel<R, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ts.ctas); cfg.blockDim = dim3(32 * R); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, tstream_kernel<R, MINB>, d, p, ts, *(const TsMaps *)ts.maps, cur, last);
}

cudaError_t launch_tstream(const Dom &d, const KParams &p, const TsPlan &ts, int cur, int last, bool pdl, cudaStream_t s) {
  switch (ts.rows) {
    case 12: return launch_tstream_t<12, 1>(d, p, ts, cur, last, pdl, s);
    case 6: return launch_tstream_t<6, 2>(d, p, ts, cur, last, pdl, s);
    default: return cudaErrorInvalidValue;
  }
}

size_t tstream_map_bytes() { return sizeof(TsMaps); }

// Host side of the plan: the tensor maps (one per array; the box extents follow `rows`) into host_maps (tstream_map_bytes()), and the
// cut into strips and segments.
// Returns 0 on success, 1 with a reason in `why` when the form cannot be used.
int tstream_plan(const Dom &d, size_t dom_rows, const double *HTN, const double *HTE, int num_sms, int rows, void *host_maps, TsPlan *ts,
                 char *why, size_t nwhy) {
  if (rows != 12 && rows != 6) { snprintf(why, nwhy, "rows must be 12 or 6"); return 1; }
  if (!HTN || !HTE) { snprintf(why, nwhy, "needs the metric arrays (evp_b200_set_metric)"); return 1; }
  typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn || qr != cudaDriverEntryPointSuccess) {
    snprintf(why, nwhy, "the driver does not export cuTensorMapEncodeTiled");
    return 1;
  }
  const Encode encode = (Encode)fn;
  TmaMap *m = (TmaMap *)host_maps;
  int bad = 0;
  // Driver workaround, as CUTLASS applies it in make_tma_copy_desc: drivers up to CUDA 13.1 may set bit 21 of the second descriptor word
  // for tensors of less than 128 KiB (small sub-domains, byte masks), which the TMA unit does not accept.  (Driver 580 on the B200 pool
  // did not set it; what its TMA unit rejected were box starts off a 16-byte boundary -- see the header comment.)
  // L2 promotion of the box loads (the granule the TMA unit asks L2 to fetch): none, 128 B and 256 B measured the same within 1 %
  // (profiles/r2_tstream_ab.txt)
  const CUtensorMapL2promotion l2p = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  int drv = 0;
  cudaDriverGetVersion(&drv);
  auto small_fix = [&](TmaMap &t, size_t bytes) {
    if (drv <= 13010 && bytes < 131072) reinterpret_cast<uint64_t *>(&t.m)[1] &= ~(1ull << 21);
  };
  auto f64 = [&](int idx, const double *base, int bx, int by) {
    const cuuint64_t dims[2] = {(cuuint64_t)d.ld, (cuuint64_t)dom_rows}, strides[1] = {(cuuint64_t)d.ld * 8};
    const cuuint32_t box[2] = {(cuuint32_t)bx, (cuuint32_t)by}, es[2] = {1, 1};
    if (encode(&m[idx].m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      ++bad;
    small_fix(m[idx], (size_t)d.ld * dom_rows * 8);
  };
  auto u8 = [&](int idx, const unsigned char *base) {
    const cuuint64_t dims[2] = {(cuuint64_t)d.ld, (cuuint64_t)dom_rows}, strides[1] = {(cuuint64_t)d.ld};
    const cuuint32_t box[2] = {TS_MW, (cuuint32_t)rows}, es[2] = {1, 1};
    if (encode(&m[idx].m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      ++bad;
    small_fix(m[idx], (size_t)d.ld * dom_rows);
  };
  for (int b = 0; b < 2; ++b) {
    f64(TS_MAP_U + b, d.u[b], TS_W, rows + 1);
    f64(TS_MAP_V + b, d.v[b], TS_W, rows + 1);
    for (int q = 0; q < 12; ++q) f64(TS_MAP_SIG + 12 * b + q, d.sig[b][q], TS_W, rows);
  }
  f64(TS_MAP_STRENGTH, d.strength, TS_W, rows);
  f64(TS_MAP_DXT, d.dxT, TS_W, rows);
  f64(TS_MAP_DYT, d.dyT, TS_W, rows);
  f64(TS_MAP_HTN, HTN, TS_W, rows + 1);
  f64(TS_MAP_HTE, HTE, TS_W, rows);
  const double *uop[12] = {d.cdn, d.aiu, d.uocn, d.vocn, d.waterx, d.watery, d.forcex, d.forcey, d.umassdti, d.fm, d.uarear, d.TbU};
  for (int q = 0; q < 12; ++q) f64(TS_MAP_UOP + q, uop[q], TS_W, rows);
  u8(TS_MAP_MASKT, d.maskT);
  u8(TS_MAP_MASKU, d.maskU);
  if (bad) { snprintf(why, nwhy, "cuTensorMapEncodeTiled refused %d of %d arrays (ld %d)", bad, TS_NMAPS, d.ld); return 1; }

  tstream_cut(d.nx, d.ny, num_sms, rows, ts);
  return 0;
}

#endif  // EVP_HOST_EMU

}  // namespace EVP_NS
}  // namespace evp
