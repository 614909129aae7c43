// evp_abi.cu -- the C ABI of include/evp_b200.h: device memory, host<->device staging in the
// reference's block layout, sub-domain stitching, the subcycle driver (CUDA-graphed) and the halo.
//
// Host side of the seam this replaces: dyn_evp1d_init / dyn_evp1d_run / dyn_evp1d_finalize
// (cicecore/cicedyn/dynamics/ice_dyn_evp1d.F90:73-330) as dispatched from evp()
// (ice_dyn_evp.F90:153-155, 846-914).  Where the reference's 1-D solver gathers every field to the
// master task and converts 2-D -> 1-D (ice_dyn_evp1d.F90:178-208), this library keeps the data
// distributed: each rank stitches its own blocks into one rectangle ("dom") on its GPU.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "evp_b200.h"
#include "evp_internal.h"
#include "evp_persist_plan.h"
#include "evp_halo.h"

namespace evp {

static char g_err[1024] = "";
static int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return 1;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// pack / unpack kernels between the reference's block layout and the dom layout
// ------------------------------------------------------------------------------------------------
__global__ void pack_f64(double *__restrict__ dst, const double *__restrict__ src, const int *__restrict__ gsrc, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int s = gsrc[k];
    dst[k] = (s >= 0) ? src[s] : 0.0;
  }
}
// same, into both ping-pong copies of a carried field (they must start identical: cells off the ice are never written)
__global__ void pack2_f64(double *__restrict__ dst0, double *__restrict__ dst1, const double *__restrict__ src,
                          const int *__restrict__ gsrc, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int s = gsrc[k];
    const double v = (s >= 0) ? src[s] : 0.0;
    dst0[k] = v;
    dst1[k] = v;
  }
}
// device-resident stresses: what dyn_prep2 does to the carried stresses on the host before every
// loop -- zero them where there is no ice (ice_dyn_shared.F90:717-730) -- applied to both ping-pong copies
struct SigPtrs { double *p[24]; };
__global__ void zero_stress_off_ice(const __grid_constant__ SigPtrs sp, const unsigned char *__restrict__ maskT, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    if (!maskT[k]) {
#pragma unroll
      for (int q = 0; q < 24; ++q) sp.p[q][k] = 0.0;
    }
  }
}
// block-layout bookkeeping of that zeroing for the cells the device does not own (W/S ghost cells, padding): the host
// zeroes the WHOLE block where iceTmask is false, so when resident stresses are fetched the same cells must read zero
__global__ void note_off_ice(unsigned char *__restrict__ ever_off, const int *__restrict__ maskT_blk, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    if (maskT_blk[k] == 0) ever_off[k] = 1;
}
__global__ void zero_where(double *__restrict__ a, const unsigned char *__restrict__ ever_off, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    if (ever_off[k]) a[k] = 0.0;
}
__global__ void pack_mask(unsigned char *__restrict__ dst, const int *__restrict__ src, const int *__restrict__ gsrc,
                          int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int s = gsrc[k];
    dst[k] = (s >= 0 && src[s] != 0) ? 1 : 0;
  }
}
__global__ void unpack_f64(double *__restrict__ dst_blk, const double *__restrict__ src_dom, const int *__restrict__ lin,
                           const int *__restrict__ dom, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) dst_blk[lin[k]] = src_dom[dom[k]];
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
enum FieldId {
  // inout state (order of evp_b200_fields_t)
  F_SIG0 = 0,                                    // 12 stresses
  F_STRINTX = 12, F_STRINTY, F_TAUBX, F_TAUBY, F_U, F_V,  // 18 inout in total
  // per-step inputs
  F_STRENGTH = 18, F_CDN, F_AIU, F_UOCN, F_VOCN, F_WATERX, F_WATERY, F_FORCEX, F_FORCEY, F_UMASSDTI, F_FM, F_TBU,
  NF_STEP = 30,
  // static geometry
  G_DXT = 30, G_DYT, G_DXHY, G_DYHX, G_CXP, G_CYP, G_CXM, G_CYM, G_DMIN, G_UAREAR,
  NF_TOTAL = 40
};

struct Ctx {
  bool inited = false;
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  // host-side description
  int nx_block = 0, ny_block = 0, nblocks = 0, max_blocks = 0, nxg = 0, nyg = 0, ew = 0, ns = 0;
  size_t nblk_elems = 0;  // nx_block*ny_block*max_blocks
  int gi0 = 0, gj0 = 0;   // global index of dom interior (1,1)
  size_t ndom = 0;        // ld*nyd

  Dom dom{};
  // device memory
  double *dfield[NF_TOTAL] = {};  // dom-layout arrays (sig/u/v: copy 0)
  double *dsig1[12] = {}, *du1 = nullptr, *dv1 = nullptr;  // ping-pong copy 1
  double *dshare = nullptr;  // [u0|u1|v0|v1] + 64 u64 flags in one allocation (CUDA-IPC shareable)
  P2PState p2p;
  double *duinit = nullptr, *dvinit = nullptr;
  double *dstr[8] = {};
  unsigned char *dmaskT = nullptr, *dmaskU = nullptr;
  double *stage[NF_STEP] = {};    // block-layout staging, one per time-varying field
  int *stage_mask = nullptr, *stage_mask2 = nullptr;
  cudaStream_t xfer = nullptr;          // copy-engine stream: H2D/D2H run beside the pack/unpack kernels
  cudaEvent_t ev_field[34] = {};         // one per staged field
  bool stress_resident = false;
  bool stress_uploaded_this_call = false;
  unsigned char *d_ever_off = nullptr;   // block layout: iceTmask was false in some step since the host last saw the stresses          // the device holds the current stresses (evp_b200_run_bgrid_resident)
  int *d_gsrc = nullptr;
  int *d_uv_lin = nullptr, *d_uv_dom = nullptr, n_uv = 0;
  int *d_sig_lin = nullptr, *d_sig_dom = nullptr, n_sig = 0;
  int *d_int_lin = nullptr, *d_int_dom = nullptr, n_int = 0;
  int *d_symw_lin = nullptr, *d_symw_dom = nullptr, n_symw = 0;  // tripole: west-ghost corner of the north ghost row of every top block
  bool sym_applied = false;                                       // the device stresses carry the symmetrisation across the fold
  double *d_symrow = nullptr;                                     // [12][nx_global]: the top physical row of every stress array (ranks of the top row)

  // loop state
  int cur = 0;
  bool uploaded = false;
  float last_ms = 0.f;
  int64_t last_launches = 0;
  std::string desc;

  // cached graph of the whole ndte loop
  cudaGraphExec_t gexec = nullptr;
  evp_b200_params_t gparams{};
  int64_t glaunches = 0;
  int gcur_end = 0;

  HaloPlan halo;

  // C grid
  bool cinit = false;
  CDom cdom{};
  std::vector<double *> dbuf;      // deformations: dxU, dyU, tarear + 5 outputs
  std::vector<double *> dstage;    // their own block-layout staging (g.stage still holds the fields of an upload that a later
                                   // evp_b200_download builds on)
  std::vector<double *> cbuf;      // every C-grid device array (freed together)
  std::vector<double *> cstage;    // staging, one per C field
  std::vector<double *> cdstage;   // staging, one per CD field
  bool cdinit = false;             // CD-only device arrays allocated
  cudaGraphExec_t cdexec = nullptr;
  evp_b200_params_t cdparams{};
  unsigned char *cmask[4] = {};
  cudaGraphExec_t cexec = nullptr;  // cached graph of the C-grid loop (dies with the context: it bakes pointers and wrap flags)
  evp_b200_params_t cparams{};
  int cgraph_launches = 0;
  bool c_fused = true;               // three kernels per subcycle (kA, kB, k5); params->kernel == SPLIT selects the five-kernel form

  // KERNEL_PERSISTENT
  PersistPlan pplan{};
  // step preparation on the device (evp_b200_prep_init / evp_b200_step_resident)
  double *prep_static[4] = {};   // hm, tarea, uarea, fcor
  unsigned char *prep_umask = nullptr;
  double *prepT[10] = {};        // tmass, aice_init, cdn_ocn, uocn, vocn, ss_tltx, ss_tlty, strairxT, strairyT, TbU
  bool prep_ok = false, state_resident = false;
  bool persist_ok = false;
  bool persist_p2p = false;  // the persistent kernel's in-kernel NVLink halo form (neighbour ranks)
  std::string persist_why;
  unsigned *d_progress = nullptr;
  unsigned *d_ptab[2] = {nullptr, nullptr};  // (slot, thread) -> cell tables
  int *d_perr = nullptr;
  long long *d_pdbg = nullptr;
  int num_sms = 0;
  bool streaming = false;    // the sub-domain's arrays exceed L2: HBM-streaming form of the fused kernel
  bool derived_ok = false;   // evp_b200_set_metric: HTN/HTE reproduce the seven derived geometry arrays bit for bit
  int metric_mismatches = -1;
  // KERNEL_TSTREAM (evp_tstream.cu): the TMA tile-streaming form for sub-domains that stream from HBM
  double *d_HTN = nullptr, *d_HTE = nullptr;   // metric arrays on the dom layout (owned by cbuf)
  double deltamin = 0.0;
  std::vector<unsigned char> ts_hmaps;         // TS_NMAPS tensor maps (host copy: they travel as a kernel parameter)
  int *d_tserr = nullptr;
  TsPlan tsplan{};
  int ts_rows = 12;                            // T rows per block: 12 (one CTA per SM) or 6 (two); EVP_B200_TSTREAM_ROWS
  const void *ts_key[3] = {nullptr, nullptr, nullptr};   // what the maps were encoded for: u[0], sig[0][0], rows
  std::string ts_why;
};
static Ctx g;
static CommState g_comm;
static bool g_allow_partial = false;  // evp_b200_allow_partial_domain

static int grid_blocks(size_t n) { return (int)std::min<size_t>((n + 255) / 256, 148 * 16); }

static void destroy_graph() {
  if (g.gexec) {
    cudaGraphExecDestroy(g.gexec);
    g.gexec = nullptr;
  }
  if (g.cexec) {
    cudaGraphExecDestroy(g.cexec);
    g.cexec = nullptr;
  }
  if (g.cdexec) {
    cudaGraphExecDestroy(g.cdexec);
    g.cdexec = nullptr;
  }
}

static int free_all() {
  destroy_graph();
  g.cinit = false;
  g.cdinit = false;
  auto F = [](auto *&p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  g.p2p.release();
  for (auto &p : g.cbuf) F(p);
  for (auto &p : g.dbuf) F(p);
  for (auto &p : g.dstage) F(p);
  for (auto &p : g.cstage) F(p);
  for (auto &p : g.cdstage) F(p);
  for (auto &p : g.cmask) F(p);
  g.dfield[F_U] = g.dfield[F_V] = g.du1 = g.dv1 = nullptr;  // live inside dshare
  F(g.dshare);
  for (auto &p : g.dfield) F(p);
  for (auto &p : g.dsig1) F(p);
  F(g.du1); F(g.dv1); F(g.duinit); F(g.dvinit);
  for (auto &p : g.dstr) F(p);
  F(g.dmaskT); F(g.dmaskU);
  for (auto &p : g.stage) F(p);
  F(g.stage_mask); F(g.stage_mask2); F(g.d_ever_off); F(g.d_gsrc); F(g.d_progress); F(g.d_ptab[0]); F(g.d_ptab[1]); F(g.d_perr); F(g.d_pdbg);
  F(g.d_tserr);
  for (auto &p : g.prep_static) F(p);
  for (auto &p : g.prepT) F(p);
  F(g.prep_umask);
  F(g.d_symw_lin); F(g.d_symw_dom); F(g.d_symrow); F(g.d_uv_lin); F(g.d_uv_dom); F(g.d_sig_lin); F(g.d_sig_dom); F(g.d_int_lin); F(g.d_int_dom);
  g.halo.release();
  for (auto &e : g.ev_field) if (e) cudaEventDestroy(e);
  if (g.xfer) cudaStreamDestroy(g.xfer);
  if (g.ev0) cudaEventDestroy(g.ev0);
  if (g.ev1) cudaEventDestroy(g.ev1);
  if (g.stream) cudaStreamDestroy(g.stream);
  const int dev = g.device;
  g = Ctx{};
  g.device = dev;  // the device chosen with evp_b200_set_device survives a failed or repeated init
  return 0;
}

template <class T>
static int upload_vec(T *&dptr, const std::vector<T> &h) {
  CK(cudaMalloc(&dptr, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) CK(cudaMemcpy(dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// KERNEL_PERSISTENT: tiling, shared-memory layout and (slot, thread) tables (evp_persist_plan.h)
static int plan_persist() {
  g.persist_ok = false;
  // between subcycles the tiles exchange through the rank's own arrays; neighbour ranks are reached by in-kernel NVLink stores
  // (P2PState).  A staged exchange (NCCL fallback) or a tripole fold kernel between subcycles cannot run inside one launch.
  const bool halo_in_kernel = g.p2p.enabled && g.p2p.npeers > 0 && g.p2p.fold_n == 0 && g.ns != EVP_B200_BNDY_TRIPOLE;
  if ((g.halo.n_dst != 0 || !g.halo.peers.empty()) && !halo_in_kernel) {
    g.persist_why = "the halo needs an exchange between subcycles (staged exchange with neighbour ranks, or tripole fold)";
    return 0;
  }
  PersistPlan pp{};
  PersistTables tb;
  if (!persist_plan(g.dom.nx, g.dom.ny, g.num_sms, PERSIST_THREADS, 232448, pp, tb, g.persist_why)) return 0;
  {
    int per_sm = 0;
    CK(exact::persist_ctas_per_sm(pp, halo_in_kernel, &per_sm));
    if (per_sm < 1 || pp.ntx * pp.nty > per_sm * g.num_sms) {
      g.persist_why = "the driver cannot keep every tile's CTA resident at once (cooperative launch)";
      return 0;
    }
  }
  unsigned *dt = nullptr, *du = nullptr;
  if (upload_vec(dt, tb.tslot) || upload_vec(du, tb.uslot)) return 1;
  g.d_ptab[0] = dt; g.d_ptab[1] = du;
  CK(cudaMalloc(&g.d_progress, sizeof(unsigned) * PERSIST_CTR_STRIDE * pp.ntx * pp.nty));
  CK(cudaMalloc(&g.d_perr, sizeof(int)));
  CK(cudaMemset(g.d_perr, 0, sizeof(int)));
  pp.tslot = dt; pp.uslot = du; pp.progress = g.d_progress; pp.err = g.d_perr;
  pp.n_sig = 0;
  for (int ty = 0; ty < pp.nty; ++ty)
    for (int tx = 0; tx < pp.ntx; ++tx)
      if (tx == 0 || tx == pp.ntx - 1 || ty == 0 || ty == pp.nty - 1) pp.n_sig += pp.ewU[(tx == pp.ntx - 1 ? 1 : 0) | (ty == pp.nty - 1 ? 2 : 0)];
  g.persist_p2p = halo_in_kernel;
  if (getenv("EVP_B200_PERSIST_DEBUG")) {  // per-warp cycle accounting, printed after every loop
    CK(cudaMalloc(&g.d_pdbg, sizeof(long long) * (5 * (PERSIST_THREADS / 32) + 32) * pp.ntx * pp.nty));
    CK(cudaMemset(g.d_pdbg, 0, sizeof(long long) * (5 * (PERSIST_THREADS / 32) + 32) * pp.ntx * pp.nty));
    pp.dbg = g.d_pdbg;
  }
  g.pplan = pp;
  g.persist_ok = true;
  return 0;
}

static int do_init(const evp_b200_grid_t *gr) {
  if (!gr) return fail("evp_b200_init: null grid");
  if (gr->abi_version != EVP_B200_ABI_VERSION) return fail("evp_b200_init: abi_version %d != %d", gr->abi_version, EVP_B200_ABI_VERSION);
  if (gr->nghost != 1) return fail("evp_b200_init: nghost must be 1 (ice_blocks.F90:48), got %d", gr->nghost);
  if (gr->nblocks < 1 || gr->nblocks > gr->max_blocks) return fail("evp_b200_init: nblocks=%d max_blocks=%d", gr->nblocks, gr->max_blocks);
  if (gr->nx_block < 3 || gr->ny_block < 3) return fail("evp_b200_init: block too small");
  if (gr->ns_boundary_type == EVP_B200_BNDY_TRIPOLE && gr->ew_boundary_type != EVP_B200_BNDY_CYCLIC)
    return fail("evp_b200_init: tripole requires ew_boundary_type cyclic");
  if (g.inited) free_all();

  if (g.device < 0) CK(cudaGetDevice(&g.device));
  CK(cudaSetDevice(g.device));
  CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&g.ev0));
  CK(cudaEventCreate(&g.ev1));
  CK(cudaStreamCreateWithFlags(&g.xfer, cudaStreamNonBlocking));
  for (auto &e : g.ev_field) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

  g.nx_block = gr->nx_block; g.ny_block = gr->ny_block; g.nblocks = gr->nblocks; g.max_blocks = gr->max_blocks;
  g.nxg = gr->nx_global; g.nyg = gr->ny_global; g.ew = gr->ew_boundary_type; g.ns = gr->ns_boundary_type;
  const int nxb = g.nx_block, nyb = g.ny_block;
  g.nblk_elems = (size_t)nxb * nyb * g.max_blocks;
  if (g.nblk_elems > 0x7fffffffULL) return fail("evp_b200_init: block arrays exceed 2^31 elements");

  // ---- the rank's rectangle --------------------------------------------------------------------
  int gi0 = 1 << 30, gi1 = -1, gj0 = 1 << 30, gj1 = -1;
  size_t ncell = 0;
  for (int b = 0; b < g.nblocks; ++b) {
    const int32_t *ig = gr->i_glob + (size_t)b * nxb, *jg = gr->j_glob + (size_t)b * nyb;
    const int ilo = gr->ilo[b], ihi = gr->ihi[b], jlo = gr->jlo[b], jhi = gr->jhi[b];
    if (ilo != 2 || jlo != 2 || ihi < ilo || jhi < jlo || ihi > nxb - 1 || jhi > nyb - 1)
      return fail("evp_b200_init: block %d has bad interior bounds", b + 1);
    if (ig[ilo - 1] < 1 || ig[ihi - 1] > g.nxg || jg[jlo - 1] < 1 || jg[jhi - 1] > g.nyg || ig[ihi - 1] - ig[ilo - 1] != ihi - ilo ||
        jg[jhi - 1] - jg[jlo - 1] != jhi - jlo)
      return fail("evp_b200_init: block %d has inconsistent i_glob/j_glob", b + 1);
    gi0 = std::min(gi0, (int)ig[ilo - 1]); gi1 = std::max(gi1, (int)ig[ihi - 1]);
    gj0 = std::min(gj0, (int)jg[jlo - 1]); gj1 = std::max(gj1, (int)jg[jhi - 1]);
    ncell += (size_t)(ihi - ilo + 1) * (jhi - jlo + 1);
  }
  const int nx = gi1 - gi0 + 1, ny = gj1 - gj0 + 1;
  // Land-block elimination (ice_domain.F90: blocks without ocean are not distributed) leaves holes in the rectangle:
  // allowed.  A hole cell is land for every purpose -- masks 0, fields 0, exactly what the reference's halo update puts
  // into the ghost cells that face an eliminated block (ice_boundary.F90:1398-1408) -- unless a neighbouring block holds
  // it as a ghost cell, whose (halo-filled) value is then used.
  if (ncell > (size_t)nx * ny) return fail("evp_b200_init: the rank's blocks overlap (%zu cells in a %dx%d rectangle)", ncell, nx, ny);
  const size_t nholes = (size_t)nx * ny - ncell;
  g.gi0 = gi0; g.gj0 = gj0;
  Dom &d = g.dom;
  d.nx = nx; d.ny = ny; d.nyd = ny + 2;
  d.ld = ((nx + 2 + 15) / 16) * 16;
  g.ndom = dom_cells(nx, ny);  // ghost ring + two staging rows (tripole fold, evp_halo.cu)
  if (g.ndom > 0x7fffffffULL) return fail("evp_b200_init: sub-domain exceeds 2^31 cells");

  // ---- index maps -------------------------------------------------------------------------------
  std::vector<int> gsrc(g.ndom, -1), uv_lin, uv_dom, sig_lin, sig_dom, int_lin, int_dom, symw_lin, symw_dom;
  std::vector<unsigned char> owned(g.ndom, 0);
  for (int b = 0; b < g.nblocks; ++b) {
    const int32_t *ig = gr->i_glob + (size_t)b * nxb, *jg = gr->j_glob + (size_t)b * nyb;
    const int ilo = gr->ilo[b], ihi = gr->ihi[b], jlo = gr->jlo[b], jhi = gr->jhi[b];
    const int oi = ig[ilo - 1] - gi0 + 1 - ilo, oj = jg[jlo - 1] - gj0 + 1 - jlo;  // dom = local + o
    for (int lj = jlo - 1; lj <= jhi + 1; ++lj)
      for (int li = ilo - 1; li <= ihi + 1; ++li) {
        const int di = li + oi, dj = lj + oj;
        if (di < 0 || di > nx + 1 || dj < 0 || dj > ny + 1) return fail("evp_b200_init: internal: dom index out of range");
        const int lin = (li - 1) + nxb * ((lj - 1) + nyb * b);
        const int dm = dj * d.ld + di;
        const bool interior = (li >= ilo && li <= ihi && lj >= jlo && lj <= jhi);
        const bool ring = (di == 0 || di == nx + 1 || dj == 0 || dj == ny + 1);
        if (interior) {
          if (owned[dm]) return fail("evp_b200_init: blocks overlap");
          owned[dm] = 1;
          gsrc[dm] = lin;
          int_lin.push_back(lin); int_dom.push_back(dm);
        } else if (!owned[dm] && gsrc[dm] < 0) {
          gsrc[dm] = lin;  // ghost ring of the rectangle, or a hole cell next to a block: some block's ghost copy of it
        }
        uv_lin.push_back(lin); uv_dom.push_back(dm);
        if (li >= ilo && lj >= jlo) { sig_lin.push_back(lin); sig_dom.push_back(dm); }
        // ice_HaloUpdate_stress also writes the west ghost column of the north ghost row (ice_boundary.F90:8141: i = 1 ..)
        if (gr->ns_boundary_type == EVP_B200_BNDY_TRIPOLE && jg[jhi - 1] == g.nyg && li == ilo - 1 && lj == jhi + 1) {
          symw_lin.push_back(lin); symw_dom.push_back(dm);
        }
      }
  }
  g.n_symw = (int)symw_lin.size();
  if (g.n_symw && (upload_vec(g.d_symw_lin, symw_lin) || upload_vec(g.d_symw_dom, symw_dom))) return 1;
  g.n_uv = (int)uv_lin.size(); g.n_sig = (int)sig_lin.size(); g.n_int = (int)int_lin.size();
  if (upload_vec(g.d_gsrc, gsrc) || upload_vec(g.d_uv_lin, uv_lin) || upload_vec(g.d_uv_dom, uv_dom) ||
      upload_vec(g.d_sig_lin, sig_lin) || upload_vec(g.d_sig_dom, sig_dom) || upload_vec(g.d_int_lin, int_lin) ||
      upload_vec(g.d_int_dom, int_dom))
    return 1;

  // ---- device memory ----------------------------------------------------------------------------
  const size_t bdom = g.ndom * sizeof(double), bblk = g.nblk_elems * sizeof(double);
  CK(cudaMalloc(&g.dshare, P2PState::share_bytes(g.ndom, nx, ny)));
  CK(cudaMemsetAsync(g.dshare, 0, P2PState::share_bytes(g.ndom, nx, ny), g.stream));
  for (int f = 0; f < NF_TOTAL; ++f) {
    if (f == F_U || f == F_V) continue;
    CK(cudaMalloc(&g.dfield[f], bdom)); CK(cudaMemsetAsync(g.dfield[f], 0, bdom, g.stream));
  }
  g.dfield[F_U] = g.dshare; g.du1 = g.dshare + g.ndom; g.dfield[F_V] = g.dshare + 2 * g.ndom; g.dv1 = g.dshare + 3 * g.ndom;
  for (int q = 0; q < 12; ++q) { CK(cudaMalloc(&g.dsig1[q], bdom)); CK(cudaMemsetAsync(g.dsig1[q], 0, bdom, g.stream)); }
  CK(cudaMalloc(&g.duinit, bdom)); CK(cudaMalloc(&g.dvinit, bdom));
  CK(cudaMemsetAsync(g.duinit, 0, bdom, g.stream)); CK(cudaMemsetAsync(g.dvinit, 0, bdom, g.stream));
  for (int q = 0; q < 8; ++q) { CK(cudaMalloc(&g.dstr[q], bdom)); CK(cudaMemsetAsync(g.dstr[q], 0, bdom, g.stream)); }
  CK(cudaMalloc(&g.dmaskT, g.ndom)); CK(cudaMalloc(&g.dmaskU, g.ndom));
  CK(cudaMemsetAsync(g.dmaskT, 0, g.ndom, g.stream)); CK(cudaMemsetAsync(g.dmaskU, 0, g.ndom, g.stream));
  for (int f = 0; f < NF_STEP; ++f) CK(cudaMalloc(&g.stage[f], bblk));
  CK(cudaMalloc(&g.stage_mask, g.nblk_elems * sizeof(int)));
  CK(cudaMalloc(&g.stage_mask2, g.nblk_elems * sizeof(int)));
  CK(cudaMalloc(&g.d_ever_off, g.nblk_elems));
  CK(cudaMemsetAsync(g.d_ever_off, 0, g.nblk_elems, g.stream));

  // ---- static geometry --------------------------------------------------------------------------
  const double *geo[10] = {gr->dxT, gr->dyT, gr->dxhy, gr->dyhx, gr->cxp, gr->cyp, gr->cxm, gr->cym, gr->DminTarea, gr->uarear};
  for (int q = 0; q < 10; ++q) {
    if (!geo[q]) return fail("evp_b200_init: null geometry array %d", q);
    CK(cudaMemcpyAsync(g.stage[q], geo[q], bblk, cudaMemcpyHostToDevice, g.stream));
    pack_f64<<<grid_blocks(g.ndom), 256, 0, g.stream>>>(g.dfield[G_DXT + q], g.stage[q], g.d_gsrc, (int)g.ndom);
  }
  CK(cudaGetLastError());

  // ---- Dom ----------------------------------------------------------------------------------------
  for (int q = 0; q < 12; ++q) { d.sig[0][q] = g.dfield[F_SIG0 + q]; d.sig[1][q] = g.dsig1[q]; }
  d.u[0] = g.dfield[F_U]; d.u[1] = g.du1; d.v[0] = g.dfield[F_V]; d.v[1] = g.dv1;
  d.strength = g.dfield[F_STRENGTH];
  d.dxT = g.dfield[G_DXT]; d.dyT = g.dfield[G_DYT]; d.dxhy = g.dfield[G_DXHY]; d.dyhx = g.dfield[G_DYHX];
  d.cxp = g.dfield[G_CXP]; d.cyp = g.dfield[G_CYP]; d.cxm = g.dfield[G_CXM]; d.cym = g.dfield[G_CYM];
  d.DminTarea = g.dfield[G_DMIN]; d.uarear = g.dfield[G_UAREAR];
  d.cdn = g.dfield[F_CDN]; d.aiu = g.dfield[F_AIU]; d.uocn = g.dfield[F_UOCN]; d.vocn = g.dfield[F_VOCN];
  d.waterx = g.dfield[F_WATERX]; d.watery = g.dfield[F_WATERY]; d.forcex = g.dfield[F_FORCEX]; d.forcey = g.dfield[F_FORCEY];
  d.umassdti = g.dfield[F_UMASSDTI]; d.fm = g.dfield[F_FM]; d.TbU = g.dfield[F_TBU];
  d.uinit = g.duinit; d.vinit = g.dvinit;
  d.strintx = g.dfield[F_STRINTX]; d.strinty = g.dfield[F_STRINTY]; d.taubx = g.dfield[F_TAUBX]; d.tauby = g.dfield[F_TAUBY];
  for (int q = 0; q < 8; ++q) d.str[q] = g.dstr[q];
  d.maskT = g.dmaskT; d.maskU = g.dmaskU;

  // ---- halo plan ----------------------------------------------------------------------------------
  if (g.halo.build(g_comm, gi0, gj0, nx, ny, d.ld, g.nxg, g.nyg, g.ew, g.ns, g_allow_partial, g_err, sizeof g_err)) return 1;
  d.wrap_ew = g.halo.wrap_ew; d.wrap_ns = g.halo.wrap_ns;
  d.fold_top = (g.ns == EVP_B200_BNDY_TRIPOLE && gj0 + ny - 1 == g.nyg) ? 1 : 0;
  CK(cudaStreamSynchronize(g.stream));
  if (g_comm.nranks == 1 && g.halo.build_local_fold(g.nxg, g.nyg, g.ew, g.ns, exact::fold_max_entries(), g_err, sizeof g_err)) return 1;
  if (g.p2p.setup(g_comm, g.halo, g.dshare, g.ndom, nx, ny, d.ld, g.nxg, g.nyg, g.ew, g.ns, exact::fold_max_entries(), g_err, sizeof g_err)) return 1;

  // fused kernel form.  Sub-domains whose arrays fit the 126 MB L2 run latency-bound: both masks requested at once and the
  // IEEE division / square-root expansions of the four corners interleaved.  Larger ones stream from HBM and want every operand
  // in flight early: speculative T-cell loads + momentum operands through cp.async, and -- once evp_b200_set_metric has verified
  // them -- seven geometry arrays derived from two.  Measured on B200 (profiles/): gx1 2.22 vs 2.39 ms per step; 3600x2400
  // 164 (derived) / 170 (streaming) / 182 (resident form).
  g.streaming = (g.ndom * sizeof(double) * 50 > (size_t)96 << 20);
  if (const char *e = getenv("EVP_B200_P2P_TIMEOUT_S")) {  // bound of the in-kernel halo waits (default 60 s)
    const unsigned long long ns = (unsigned long long)(std::max(atof(e), 0.001) * 1e9);
    CK(exact::set_wait_timeout(ns));
    CK(fast::set_wait_timeout(ns));
    CK(exact::set_wait_timeout_persist(ns));
    CK(fast::set_wait_timeout_persist(ns));
  }
  CK(cudaDeviceGetAttribute(&g.num_sms, cudaDevAttrMultiProcessorCount, g.device));
  // ---- persistent tiling ---------------------------------------------------------------------------
  if (plan_persist()) return 1;

  CK(cudaStreamSynchronize(g.stream));
  g.inited = true;
  char buf[512], pbuf[200];
  if (g.persist_ok)
    snprintf(pbuf, sizeof pbuf, "persistent: %dx%d tiles of %dx%d on %d SMs, smem %u B, kT=%d kU=%d, edge warps T %d U %d", g.pplan.ntx, g.pplan.nty,
             g.pplan.bx, g.pplan.by, g.num_sms, g.pplan.smem_bytes, g.pplan.kT, g.pplan.kU, g.pplan.ewT[0], g.pplan.ewU[0]);
  else
    snprintf(pbuf, sizeof pbuf, "persistent: unavailable (%s)", g.persist_why.c_str());
  snprintf(buf, sizeof buf, "dom %dx%d (ld %d) at global (%d,%d) of %dx%d; %d block(s) %dx%d, %zu hole cell(s); rank %d/%d; halo: %s; %s",
           nx, ny, d.ld, gi0, gj0, g.nxg, g.nyg, g.nblocks, nxb, nyb, nholes, g_comm.rank, g_comm.nranks, g.halo.describe().c_str(), pbuf);
  g.desc = buf;
  g.desc += std::string("; p2p: ") + (g.p2p.enabled ? "" : "off (") + g.p2p.why + (g.p2p.enabled ? "" : ")");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// upload / download
// ------------------------------------------------------------------------------------------------
static int do_upload(const evp_b200_fields_t *f, bool keep_stress = false) {
  if (!g.inited) return fail("evp_b200_upload: evp_b200_init has not been called");
  if (!f) return fail("evp_b200_upload: null fields");
  const double *src[NF_STEP] = {
      f->stressp_1, f->stressp_2, f->stressp_3, f->stressp_4, f->stressm_1, f->stressm_2, f->stressm_3, f->stressm_4,
      f->stress12_1, f->stress12_2, f->stress12_3, f->stress12_4, f->strintxU, f->strintyU, f->taubxU, f->taubyU,
      f->uvel, f->vvel, f->strength, f->cdn_ocnU, f->aiU, f->uocnU, f->vocnU, f->waterxU, f->wateryU, f->forcexU,
      f->forceyU, f->umassdti, f->fmU, f->TbU};
  const size_t bblk = g.nblk_elems * sizeof(double);
  CK(cudaSetDevice(g.device));
  keep_stress = keep_stress && g.stress_resident;
  for (int q = keep_stress ? 12 : 0; q < NF_STEP; ++q)
    if (!src[q]) return fail("evp_b200_upload: null field %d", q);
  if (!f->iceTmask || !f->iceUmask) return fail("evp_b200_upload: null mask");
  // everything carried is packed into BOTH ping-pong copies below, so copy 0 can simply be declared current; resident
  // stresses of a loop that ended on copy 1 are first brought over (no pointer swap: the cached graph bakes the pointers)
  if (keep_stress && g.cur == 1)
    for (int q = 0; q < 12; ++q)
      CK(cudaMemcpyAsync(g.dom.sig[0][q], g.dom.sig[1][q], g.ndom * sizeof(double), cudaMemcpyDeviceToDevice, g.stream));
  g.cur = 0;
  // the copy engine streams the block arrays in (g.xfer) while the pack kernels of earlier fields run (g.stream)
  CK(cudaEventRecord(g.ev_field[33], g.stream));
  CK(cudaStreamWaitEvent(g.xfer, g.ev_field[33], 0));  // staging buffers are free once earlier work on g.stream is done
  const int nb = grid_blocks(g.ndom), nd = (int)g.ndom;
  // masks first: the resident-stress path needs iceTmask early
  CK(cudaMemcpyAsync(g.stage_mask, f->iceTmask, g.nblk_elems * sizeof(int), cudaMemcpyHostToDevice, g.xfer));
  CK(cudaEventRecord(g.ev_field[30], g.xfer));
  CK(cudaMemcpyAsync(g.stage_mask2, f->iceUmask, g.nblk_elems * sizeof(int), cudaMemcpyHostToDevice, g.xfer));
  CK(cudaEventRecord(g.ev_field[31], g.xfer));
  for (int q = keep_stress ? 12 : 0; q < NF_STEP; ++q) {
    CK(cudaMemcpyAsync(g.stage[q], src[q], bblk, cudaMemcpyHostToDevice, g.xfer));
    CK(cudaEventRecord(g.ev_field[q], g.xfer));
  }
  CK(cudaStreamWaitEvent(g.stream, g.ev_field[30], 0));
  pack_mask<<<nb, 256, 0, g.stream>>>(g.dmaskT, g.stage_mask, g.d_gsrc, nd);
  CK(cudaStreamWaitEvent(g.stream, g.ev_field[31], 0));
  pack_mask<<<nb, 256, 0, g.stream>>>(g.dmaskU, g.stage_mask2, g.d_gsrc, nd);
  g.stress_uploaded_this_call = !keep_stress;
  if (!keep_stress) CK(cudaMemsetAsync(g.d_ever_off, 0, g.nblk_elems, g.stream));  // the host did its own zeroing
  if (keep_stress) {
    note_off_ice<<<grid_blocks(g.nblk_elems), 256, 0, g.stream>>>(g.d_ever_off, g.stage_mask, (int)g.nblk_elems);
    SigPtrs sp;
    for (int q = 0; q < 12; ++q) { sp.p[q] = g.dom.sig[0][q]; sp.p[12 + q] = g.dom.sig[1][q]; }
    zero_stress_off_ice<<<nb, 256, 0, g.stream>>>(sp, g.dmaskT, nd);
  }
  for (int q = keep_stress ? 12 : 0; q < NF_STEP; ++q) {
    CK(cudaStreamWaitEvent(g.stream, g.ev_field[q], 0));
    // carried state goes into both ping-pong copies: cells off the ice are never written again
    if (q < 12) pack2_f64<<<nb, 256, 0, g.stream>>>(g.dom.sig[0][q], g.dom.sig[1][q], g.stage[q], g.d_gsrc, nd);
    else if (q == F_U) pack2_f64<<<nb, 256, 0, g.stream>>>(g.dom.u[0], g.dom.u[1], g.stage[q], g.d_gsrc, nd);
    else if (q == F_V) pack2_f64<<<nb, 256, 0, g.stream>>>(g.dom.v[0], g.dom.v[1], g.stage[q], g.d_gsrc, nd);
    else pack_f64<<<nb, 256, 0, g.stream>>>(g.dfield[q], g.stage[q], g.d_gsrc, nd);
  }
  CK(cudaGetLastError());
  g.uploaded = true;
  g.stress_resident = true;
  return 0;
}

// what: bit 0 stresses, bit 1 diagnostics (strintxU, strintyU, taubxU, taubyU), bit 2 velocities
static int do_download(evp_b200_fields_t *f, int what = 7) {
  if (!g.inited || !g.uploaded) return fail("evp_b200_download: nothing uploaded");
  if (!f) return fail("evp_b200_download: null fields");
  double *dst[18] = {f->stressp_1, f->stressp_2, f->stressp_3, f->stressp_4, f->stressm_1, f->stressm_2,
                     f->stressm_3, f->stressm_4, f->stress12_1, f->stress12_2, f->stress12_3, f->stress12_4,
                     f->strintxU, f->strintyU, f->taubxU, f->taubyU, f->uvel, f->vvel};
  const size_t bblk = g.nblk_elems * sizeof(double);
  CK(cudaSetDevice(g.device));
  const Dom &d = g.dom;
  for (int q = 0; q < 18; ++q) {
    if (!((q < 12) ? (what & 1) : (q < 16) ? (what & 2) : (what & 4))) continue;
    if (!dst[q]) return fail("evp_b200_download: null field %d", q);
    if (q < 12 && !g.stress_uploaded_this_call) {
      // resident stresses are being fetched: the staging copy is stale, start from the caller's array and apply the zeroing
      // the host would have applied to the cells the device does not own
      CK(cudaMemcpyAsync(g.stage[q], dst[q], bblk, cudaMemcpyHostToDevice, g.stream));
      zero_where<<<grid_blocks(g.nblk_elems), 256, 0, g.stream>>>(g.stage[q], g.d_ever_off, (int)g.nblk_elems);
    }
    // the staging copy still holds the host array as uploaded: cells the loop does not own keep their values
    if (q < 12) {
      unpack_f64<<<grid_blocks(g.n_sig), 256, 0, g.stream>>>(g.stage[q], d.sig[g.cur][q], g.d_sig_lin, g.d_sig_dom, g.n_sig);
      if (g.sym_applied && g.n_symw)
        unpack_f64<<<grid_blocks(g.n_symw), 256, 0, g.stream>>>(g.stage[q], d.sig[g.cur][q], g.d_symw_lin, g.d_symw_dom, g.n_symw);
    } else if (q < 16)
      unpack_f64<<<grid_blocks(g.n_int), 256, 0, g.stream>>>(g.stage[q], g.dfield[q], g.d_int_lin, g.d_int_dom, g.n_int);
    else
      unpack_f64<<<grid_blocks(g.n_uv), 256, 0, g.stream>>>(g.stage[q], q == 16 ? d.u[g.cur] : d.v[g.cur], g.d_uv_lin, g.d_uv_dom, g.n_uv);
    CK(cudaEventRecord(g.ev_field[q], g.stream));
    CK(cudaStreamWaitEvent(g.xfer, g.ev_field[q], 0));
    CK(cudaMemcpyAsync(dst[q], g.stage[q], bblk, cudaMemcpyDeviceToHost, g.xfer));
  }
  if ((what & 1) && !g.stress_uploaded_this_call) CK(cudaMemsetAsync(g.d_ever_off, 0, g.nblk_elems, g.stream));  // host is current again
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaStreamSynchronize(g.xfer));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// the subcycle loop
// ------------------------------------------------------------------------------------------------
static KParams kparams(const evp_b200_params_t *p) {
  KParams k;
  k.arlx1i = p->arlx1i; k.denom1 = p->denom1; k.revp = p->revp; k.brlx = p->brlx;
  k.e_factor = p->e_factor; k.epp2i = p->epp2i; k.capping = p->capping; k.Ktens = p->Ktens;
  k.u0 = p->u0; k.cosw = p->cosw; k.sinw = p->sinw; k.rhow = p->rhow;
  k.deltaminEVP = p->deltaminEVP; k.visc_method = p->visc_method;
  return k;
}

// form of the fused kernel: 0 L2-resident, 1 HBM-streaming, 2 HBM-streaming with derived geometry (evp_kernels.cu)
static int fused_form(const evp_b200_params_t *p) {
  const bool stream = (p->kernel == EVP_B200_KERNEL_FUSED_STREAM) || (p->kernel != EVP_B200_KERNEL_FUSED_RESIDENT && g.streaming);
  return stream ? (g.derived_ok ? 2 : 1) : 0;
}

// AUTO: the persistent kernel wherever a sub-domain's carried state fits on chip (one tile per SM; measured on B200 at gx1
// 1.80 vs 2.20 ms per step, gx3 0.52 vs 0.63), else the fused kernel in the form fused_form() picks
// KERNEL_TSTREAM needs the metric arrays (derived geometry) and is the single-rank form (tested as such): between GPUs the fused
// kernel stays in charge, with the in-kernel NVLink halo or the staged exchange
static bool tstream_available() { return g.derived_ok && g.d_HTN && g.d_HTE && !g.p2p.enabled && g_comm.nranks == 1; }
static int choose_kernel(const evp_b200_params_t *p) {
  int kern = p->kernel;
  if (kern == EVP_B200_KERNEL_AUTO) {
    kern = g.persist_ok ? EVP_B200_KERNEL_PERSISTENT : EVP_B200_KERNEL_FUSED;
    if (kern == EVP_B200_KERNEL_FUSED && g.streaming && tstream_available()) kern = EVP_B200_KERNEL_TSTREAM;
  }
  if (kern == EVP_B200_KERNEL_FUSED_STREAM || kern == EVP_B200_KERNEL_FUSED_RESIDENT) kern = EVP_B200_KERNEL_FUSED;
  return kern;
}

// tensor maps and cut of KERNEL_TSTREAM for the arrays as they are bound right now (the ping-pong copies swap when a loop ends on
// copy 1); called outside graph capture
static int tstream_prepare() {
  if (!tstream_available())
    return fail("evp_b200_subcycle: the TMA tile-streaming kernel needs evp_b200_set_metric to have accepted the metric arrays and a single rank (%s)",
                (g.p2p.enabled || g_comm.nranks > 1) ? "several ranks" : (g.derived_ok ? "metric arrays missing" : "derived geometry unavailable"));
  if (const char *e = getenv("EVP_B200_TSTREAM_ROWS")) g.ts_rows = atoi(e);
  const void *key[3] = {g.dom.u[0], g.dom.sig[0][0], (const void *)(intptr_t)g.ts_rows};
  if (!g.ts_hmaps.empty() && memcmp(key, g.ts_key, sizeof key) == 0) return 0;
  g.ts_hmaps.resize(exact::tstream_map_bytes() + 64);
  void *hmaps = (void *)(((uintptr_t)g.ts_hmaps.data() + 63) & ~(uintptr_t)63);
  TsPlan ts{};
  char why[200] = "";
  if (exact::tstream_plan(g.dom, g.ndom / (size_t)g.dom.ld, g.d_HTN, g.d_HTE, g.num_sms, g.ts_rows, hmaps, &ts, why, sizeof why)) {
    g.ts_hmaps.clear();
    return fail("evp_b200_subcycle: TMA tile-streaming kernel unavailable: %s", why);
  }
  if (!g.d_tserr) { CK(cudaMalloc(&g.d_tserr, sizeof(int))); CK(cudaMemset(g.d_tserr, 0, sizeof(int))); }
  ts.maps = hmaps; ts.err = g.d_tserr; ts.deltamin = g.deltamin;
  g.tsplan = ts;
  memcpy(g.ts_key, key, sizeof key);
  if (g.desc.find("; tstream:") == std::string::npos) {
    char b[200];
    snprintf(b, sizeof b, "; tstream: %d CTAs x %d threads, %d strips x %d segments of %d blocks of %d rows", ts.ctas, 32 * ts.rows, ts.nstrips, ts.nseg,
             ts.nb, ts.rows);
    g.desc += b;
  }
  destroy_graph();
  return 0;
}

// enqueue the whole `do ksub = 1,ndte` loop (ice_dyn_evp.F90:859-913) on g.stream, starting from cur=0
static int enqueue_loop(const evp_b200_params_t *p, int *cur_end, int64_t *launches) {
  const KParams k = kparams(p);
  const bool exact = (p->mode == EVP_B200_MODE_EXACT);
  const int kern = choose_kernel(p);
  int cur = 0;
  int64_t nl = 0;
  if (kern == EVP_B200_KERNEL_PERSISTENT) {
    if (!g.persist_ok) return fail("evp_b200_subcycle: persistent kernel unavailable: %s", g.persist_why.c_str());
    if (p->ndte > 0) {
      PersistPlan pp = g.pplan;
      pp.ndte = p->ndte;
      pp.use_init = (p->revp != 0.0);
      CK(cudaMemsetAsync(g.d_progress, 0, sizeof(unsigned) * PERSIST_CTR_STRIDE * pp.ntx * pp.nty, g.stream));
      const P2PParams *px = g.persist_p2p ? &g.p2p.prm : nullptr;
      if (px) {  // loop hand-shake with the neighbour GPUs around the one launch, as for the fused kernel below
        CK(cudaMemsetAsync(g.p2p.d_done, 0, sizeof(unsigned long long), g.stream));
        CK(exact::launch_fused_p2p(g.dom, k, g.p2p.prm, 0, -1, 0, 0, g.stream));
        ++nl;
      }
      CK(exact ? exact::launch_persist(g.dom, k, pp, px, g.stream) : fast::launch_persist(g.dom, k, pp, px, g.stream));
      ++nl;
      if (px) {
        CK(exact::launch_fused_p2p(g.dom, k, g.p2p.prm, 0, -2 - p->ndte, 0, 0, g.stream));
        ++nl;
      }
    }
    *cur_end = p->ndte & 1;
    *launches = nl;
    return 0;
  }
  const bool p2p = g.p2p.enabled && kern == EVP_B200_KERNEL_FUSED;
  if (p2p) {
    CK(cudaMemsetAsync(g.p2p.d_done, 0, sizeof(unsigned long long), g.stream));
    CK(exact ? exact::launch_fused_p2p(g.dom, k, g.p2p.prm, 0, -1, 0, 0, g.stream) : fast::launch_fused_p2p(g.dom, k, g.p2p.prm, 0, -1, 0, 0, g.stream));
    ++nl;
  }
  // early programmatic-launch trigger: pays when the grid is more than one wave of co-resident CTAs (gx1 2.29 -> 2.21 ms), costs
  // on grids far smaller than the machine (gx3 0.56 -> 0.64 ms)
  const int pdl_trig = ((long)((g.dom.nx + 30) / 31) * ((g.dom.ny + 6) / 7) > 2L * g.num_sms) ? 2 : 0;
  // in-kernel NVLink halo: PDL attribute + early trigger (no-peer self test on one GPU: 2.74 -> 2.60 ms per step)
  const int p2p_pdl = 6;
  const int form = fused_form(p);
  for (int ksub = 0; ksub < p->ndte; ++ksub) {
    const int last = (ksub == p->ndte - 1);
    if (p2p) {
      CK(exact ? exact::launch_fused_p2p(g.dom, k, g.p2p.prm, cur, ksub, last | p2p_pdl, form, g.stream)
               : fast::launch_fused_p2p(g.dom, k, g.p2p.prm, cur, ksub, last | p2p_pdl, form, g.stream));
      cur ^= 1;
      ++nl;
      if (g.p2p.fold_n > 0) {  // tripole fold: what this rank combines itself once the peers' stores of this subcycle are in
        CK(exact::launch_fold(g.p2p.prm, g.dom.u[cur], g.dom.v[cur], g.p2p.d_fold_dst, g.p2p.d_fold_c1, g.p2p.d_fold_c2, g.p2p.d_fold_code,
                              g.p2p.fold_n, ksub, 1, g.stream));
        ++nl;
      }
      continue;
    }
    if (kern == EVP_B200_KERNEL_SPLIT) {
      CK(exact ? exact::launch_stress(g.dom, k, cur, g.stream) : fast::launch_stress(g.dom, k, cur, g.stream));
      CK(exact ? exact::launch_stepu(g.dom, k, cur, g.stream) : fast::launch_stepu(g.dom, k, cur, g.stream));
      nl += 2;
    } else if (kern == EVP_B200_KERNEL_FUSED) {
      CK(exact ? exact::launch_fused(g.dom, k, cur, g.stream, form, true, last | pdl_trig)
               : fast::launch_fused(g.dom, k, cur, g.stream, form, true, last | pdl_trig));
      cur ^= 1;
      nl += 1;
    } else if (kern == EVP_B200_KERNEL_TSTREAM) {
      CK(exact ? exact::launch_tstream(g.dom, k, g.tsplan, cur, last, true, g.stream)
               : fast::launch_tstream(g.dom, k, g.tsplan, cur, last, true, g.stream));
      cur ^= 1;
      nl += 1;
    } else {
      return fail("evp_b200_subcycle: kernel strategy %d not available", kern);
    }
    // the part of dyn_haloUpdate(uvel,vvel) that is not an on-rank wrap: neighbour ranks, tripole fold
    int hl = 0;
    g.halo.fold_pdl = (kern == EVP_B200_KERNEL_FUSED || kern == EVP_B200_KERNEL_TSTREAM);
    if (g.halo.exchange(g_comm, g.dom.u[cur], g.dom.v[cur], g.stream, &hl, g_err, sizeof g_err)) return 1;
    nl += hl;
  }
  if (p2p) {
    CK(exact ? exact::launch_fused_p2p(g.dom, k, g.p2p.prm, 0, -2 - p->ndte, 0, 0, g.stream)
             : fast::launch_fused_p2p(g.dom, k, g.p2p.prm, 0, -2 - p->ndte, 0, 0, g.stream));
    ++nl;
  }
  *cur_end = cur;
  *launches = nl;
  return 0;
}

static int do_subcycle(const evp_b200_params_t *p) {
  if (!g.inited || !g.uploaded) return fail("evp_b200_subcycle: no fields uploaded");
  if (!p) return fail("evp_b200_subcycle: null params");
  if (p->ndte < 0) return fail("evp_b200_subcycle: ndte < 0");
  if (p->mode != EVP_B200_MODE_EXACT && p->mode != EVP_B200_MODE_FAST) return fail("evp_b200_subcycle: unknown mode %d", p->mode);
  CK(cudaSetDevice(g.device));
  const size_t bdom = g.ndom * sizeof(double);

  // normalise to cur = 0 (a previous loop may have ended on copy 1)
  if (g.cur == 1) {
    std::swap(g.dom.u[0], g.dom.u[1]);
    std::swap(g.dom.v[0], g.dom.v[1]);
    for (int q = 0; q < 12; ++q) std::swap(g.dom.sig[0][q], g.dom.sig[1][q]);
    g.cur = 0;
    if (g.p2p.enabled) g.p2p.set_parity(g.p2p.swapped ^ 1);
    destroy_graph();  // pointers baked into the graph changed
  }
  // uvel_init = uvel at entry (ice_dyn_shared.F90:787-788; ice_dyn_evp1d.F90:939-940)
  CK(cudaMemcpyAsync(g.duinit, g.dom.u[0], bdom, cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaMemcpyAsync(g.dvinit, g.dom.v[0], bdom, cudaMemcpyDeviceToDevice, g.stream));

  if (choose_kernel(p) == EVP_B200_KERNEL_TSTREAM && tstream_prepare()) return 1;
  const bool use_graph = g.halo.graph_safe();
  int cur_end = 0;
  int64_t nl = 0;
  if (use_graph) {
    if (!g.gexec || memcmp(&g.gparams, p, sizeof *p) != 0) {
      destroy_graph();
      cudaGraph_t graph = nullptr;
      CK(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeThreadLocal));
      int rc = enqueue_loop(p, &cur_end, &nl);
      cudaError_t ce = cudaStreamEndCapture(g.stream, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return 1; }
      CK(ce);
      CK(cudaGraphInstantiate(&g.gexec, graph, 0));
      CK(cudaGraphDestroy(graph));
      g.gparams = *p; g.glaunches = nl; g.gcur_end = cur_end;
    }
    CK(cudaEventRecord(g.ev0, g.stream));
    CK(cudaGraphLaunch(g.gexec, g.stream));
    CK(cudaEventRecord(g.ev1, g.stream));
    cur_end = g.gcur_end; nl = g.glaunches;
  } else {
    CK(cudaEventRecord(g.ev0, g.stream));
    if (enqueue_loop(p, &cur_end, &nl)) return 1;
    CK(cudaEventRecord(g.ev1, g.stream));
  }
  g.cur = cur_end;
  g.sym_applied = false;
  g.last_launches = nl;
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1));
  if (choose_kernel(p) == EVP_B200_KERNEL_PERSISTENT && g.d_pdbg && p->ndte > 0) {
    const int nw = PERSIST_THREADS / 32, nt = g.pplan.ntx * g.pplan.nty;
    std::vector<long long> w((size_t)5 * nw * nt);
    CK(cudaMemcpy(w.data(), g.d_pdbg, w.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    // tile in the middle of the tile grid, every warp; then the mean over all tiles and warps
    const int mid = (g.pplan.nty / 2) * g.pplan.ntx + g.pplan.ntx / 2;
    fprintf(stderr, "[evp_b200 persistent] loop %.3f ms, %d subcycles; cycles per subcycle and warp of tile %d: phase A (ring wait+refresh) | barrier | phase C | barrier\n",
            g.last_ms, p->ndte, mid);
    for (int wv = 0; wv < nw; ++wv) {
      const long long *a = &w[((size_t)mid * nw + wv) * 5];
      fprintf(stderr, "  warp %2d: %6.0f (%5.0f) | %6.0f | %6.0f | %6.0f\n", wv, (double)a[0] / p->ndte, (double)a[1] / p->ndte, (double)a[2] / p->ndte,
              (double)a[3] / p->ndte, (double)a[4] / p->ndte);
    }
    if (p->ndte >= PERSIST_TL0 + 4) {  // timeline of four subcycles: the middle tile and its W, E, S, N neighbours, ns relative to the tile's own
      std::vector<unsigned long long> tl((size_t)nt * 32);      // phase A start of the first of them
      CK(cudaMemcpy(tl.data(), g.d_pdbg + (size_t)5 * nw * nt, tl.size() * 8, cudaMemcpyDeviceToHost));
      int slow = 0;   // the tile that waits least for its neighbours sets the pace
      long long wmin = 1LL << 60;
      for (int t = 0; t < nt; ++t) {
        long long wsum = 0;
        for (int ks = 0; ks < 4; ++ks) wsum += (long long)(tl[(size_t)t * 32 + ks * 8 + 1] - tl[(size_t)t * 32 + ks * 8]);
        if (wsum < wmin) { wmin = wsum; slow = t; }
      }
      const int nbs[5] = {mid, mid - 1, mid + 1, mid - g.pplan.ntx, slow};
      const unsigned long long t0 = tl[(size_t)mid * 32];
      fprintf(stderr, "  timeline (ns after tile %d began subcycle %d): A start | flags seen | ring refreshed | past barrier B (warp 0) | published | past barrier D | B (last warp)\n", mid, PERSIST_TL0);
      for (int q = 0; q < 5; ++q)
        for (int ks = 0; ks < 4; ++ks) {
          const unsigned long long *e = &tl[(size_t)nbs[q] * 32 + ks * 8];
          fprintf(stderr, "   tile %3d k+%d: %7lld %7lld %7lld %7lld %7lld %7lld %7lld\n", nbs[q], ks, (long long)(e[0] - t0), (long long)(e[1] - t0), (long long)(e[2] - t0),
                  (long long)(e[3] - t0), (long long)(e[4] - t0), (long long)(e[5] - t0), (long long)(e[6] - t0));
        }
    }
    if (p->ndte >= PERSIST_TL0 + 4 && getenv("EVP_B200_PERSIST_DEBUG") && atoi(getenv("EVP_B200_PERSIST_DEBUG")) > 1) {
      std::vector<unsigned long long> tl((size_t)nt * 32);
      CK(cudaMemcpy(tl.data(), g.d_pdbg + (size_t)5 * nw * nt, tl.size() * 8, cudaMemcpyDeviceToHost));
      const unsigned long long t0 = tl[(size_t)mid * 32 + 8];
      const char *what[5] = {"A start of subcycle k+1 (ns, relative to the middle tile)", "first warp at barrier B - A start", "flags seen - A start",
                             "published - last warp at barrier B", "next A start - A start (period)"};
      for (int w = 0; w < 5; ++w) {
        fprintf(stderr, "  %s, tile rows from the north:\n", what[w]);
        for (int ty = g.pplan.nty - 1; ty >= 0; --ty) {
          fprintf(stderr, "   ");
          for (int tx = 0; tx < g.pplan.ntx; ++tx) {
            const unsigned long long *e = &tl[(size_t)(ty * g.pplan.ntx + tx) * 32 + 8];
            long long v = 0;
            if (w == 0) v = (long long)(e[0] - t0);
            if (w == 1) v = (long long)(e[3] - e[0]);
            if (w == 2) v = (long long)(e[1] - e[0]);
            if (w == 3) v = (long long)(e[4] - e[6]);
            if (w == 4) v = (long long)(e[8] - e[0]);
            fprintf(stderr, " %6lld", v);
          }
          fprintf(stderr, "\n");
        }
      }
    }
    double m[5] = {0, 0, 0, 0, 0};
    for (int t = 0; t < nt * nw; ++t) for (int q = 0; q < 5; ++q) m[q] += (double)w[(size_t)t * 5 + q];
    fprintf(stderr, "  mean   : %6.0f (%5.0f) | %6.0f | %6.0f | %6.0f\n", m[0] / nt / nw / p->ndte, m[1] / nt / nw / p->ndte, m[2] / nt / nw / p->ndte,
            m[3] / nt / nw / p->ndte, m[4] / nt / nw / p->ndte);
  }
  if (choose_kernel(p) == EVP_B200_KERNEL_PERSISTENT && g.d_perr) {
    int e = 0;
    CK(cudaMemcpy(&e, g.d_perr, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) {
      CK(cudaMemset(g.d_perr, 0, sizeof(int)));
      return fail("evp_b200_subcycle: persistent kernel: a tile waited for a neighbour tile longer than 2 s (CTAs not co-resident?)");
    }
  }
  if (choose_kernel(p) == EVP_B200_KERNEL_TSTREAM && g.d_tserr) {
    int e = 0;
    CK(cudaMemcpy(&e, g.d_tserr, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) {
      CK(cudaMemset(g.d_tserr, 0, sizeof(int)));
      return fail("evp_b200_subcycle: TMA tile-streaming kernel: a box load did not complete within 2 s");
    }
  }
  if (g.p2p.enabled) {
    int e = 0;
    CK(cudaMemcpy(&e, g.p2p.d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (g.p2p.prm.dbg) {
      static std::vector<unsigned long long> w(8 + 8 * 1024);
      CK(cudaMemcpy(w.data(), g.p2p.d_dbg, w.size() * 8, cudaMemcpyDeviceToHost));
      std::vector<unsigned long long> init(w.size(), 0);
      for (int k = 0; k < 1024; ++k) init[8 + 8 * k] = ~0ULL;  // slot 0 is an atomicMin
      CK(cudaMemcpy(g.p2p.d_dbg, init.data(), w.size() * 8, cudaMemcpyHostToDevice));
      const int n = std::min(p->ndte, 1024);
      double dur = 0, gap = 0, t_edge = 0, t_push = 0, t_flag = 0, t_rel = 0;
      int cnt = 0;
      for (int k = 1; k + 1 < n; ++k) {
        const unsigned long long *a = &w[8 + 8 * k], *b = &w[8 + 8 * (k + 1)];
        if (a[0] == ~0ULL || a[0] == 0 || b[0] == ~0ULL) continue;
        dur += (double)(a[1] - a[0]); gap += (double)(b[0] - a[1]);
        t_edge += (double)(a[2] - a[0]); t_push += (double)(a[3] - a[0]); t_flag += (double)(a[4] - a[0]); t_rel += (double)(a[5] - a[0]);
        ++cnt;
      }
      if (cnt)
        fprintf(stderr, "[evp_b200 p2p rank %d] loop %.3f ms | per kernel (us): duration %.2f, gap to next %.2f, last edge CTA released from wait at %.2f, "
                        "last edge CTA done at %.2f, ring pushed+fenced at %.2f, flags written at %.2f | waits: mean %.0f cyc, max %llu\n",
                g_comm.rank, g.last_ms, dur / cnt / 1e3, gap / cnt / 1e3, t_rel / cnt / 1e3, t_edge / cnt / 1e3, t_push / cnt / 1e3, t_flag / cnt / 1e3,
                w[1] ? (double)w[0] / (double)w[1] : 0.0, w[2]);
    }
    if (e) {
      CK(cudaMemset(g.p2p.d_err, 0, sizeof(int)));  // reported once; a later loop starts clean
      return fail("evp_b200_subcycle: a neighbour GPU did not deliver its halo in time (in-kernel NVLink hand-over timed out; "
                  "EVP_B200_P2P_TIMEOUT_S sets the bound)");
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// C grid
// ------------------------------------------------------------------------------------------------
static int calloc_dom(double *&p) {
  CK(cudaMalloc(&p, g.ndom * sizeof(double)));
  CK(cudaMemsetAsync(p, 0, g.ndom * sizeof(double), g.stream));
  g.cbuf.push_back(p);
  return 0;
}

// deformations on the device-resident final velocities
// dom arrays + block-layout staging of their own for the post-loop entry points (deformations, dyn_finish)
static int ensure_aux_buffers() {
  if (!g.dbuf.empty()) return 0;
  for (int q = 0; q < 8; ++q) {
    double *p = nullptr;
    CK(cudaMalloc(&p, g.ndom * sizeof(double)));
    g.dbuf.push_back(p);
    double *st = nullptr;
    CK(cudaMalloc(&st, g.nblk_elems * sizeof(double)));
    g.dstage.push_back(st);
  }
  return 0;
}

static int do_dyn_finish(evp_b200_finish_t *ff) {
  if (!g.inited || !g.uploaded) return fail("evp_b200_dyn_finish: no velocities on the device (run the loop first)");
  if (!ff || !ff->strocnxU || !ff->strocnyU) return fail("evp_b200_dyn_finish: null argument");
  CK(cudaSetDevice(g.device));
  if (ensure_aux_buffers()) return 1;
  const size_t bblk = g.nblk_elems * sizeof(double);
  double *host[2] = {ff->strocnxU, ff->strocnyU};
  for (int q = 0; q < 2; ++q) {  // inout: points off the U list keep the caller's values
    CK(cudaMemcpyAsync(g.dstage[q], host[q], bblk, cudaMemcpyHostToDevice, g.stream));
    pack_f64<<<grid_blocks(g.ndom), 256, 0, g.stream>>>(g.dbuf[q], g.dstage[q], g.d_gsrc, (int)g.ndom);
  }
  CK(exact::launch_finish(g.dom, g.cur, g.dbuf[0], g.dbuf[1], ff->rhow, ff->cosw, ff->sinw, g.stream));
  for (int q = 0; q < 2; ++q) {
    unpack_f64<<<grid_blocks(g.n_int), 256, 0, g.stream>>>(g.dstage[q], g.dbuf[q], g.d_int_lin, g.d_int_dom, g.n_int);
    CK(cudaMemcpyAsync(host[q], g.dstage[q], bblk, cudaMemcpyDeviceToHost, g.stream));
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

// metric arrays for the derived-geometry kernels (variants 59 / 63): upload, check on the device that they reproduce the seven
// derived arrays of evp_b200_init bit for bit on every T cell the loop can touch, publish them to both kernel units
static int do_set_metric(const double *HTN, const double *HTE, double deltaminEVP, int32_t *mismatches) {
  if (!g.inited) return fail("evp_b200_set_metric: call evp_b200_init first");
  if (!HTN || !HTE) return fail("evp_b200_set_metric: null argument");
  CK(cudaSetDevice(g.device));
  const size_t bblk = g.nblk_elems * sizeof(double);
  double *dm[2] = {nullptr, nullptr}, *scratch = nullptr;
  int *dcount = nullptr;
  CK(cudaMalloc(&scratch, bblk));
  CK(cudaMalloc(&dcount, sizeof(int)));
  CK(cudaMemsetAsync(dcount, 0, sizeof(int), g.stream));
  const double *src[2] = {HTN, HTE};
  for (int q = 0; q < 2; ++q) {
    if (calloc_dom(dm[q])) return 1;
    CK(cudaMemcpyAsync(scratch, src[q], bblk, cudaMemcpyHostToDevice, g.stream));
    pack_f64<<<grid_blocks(g.ndom), 256, 0, g.stream>>>(dm[q], scratch, g.d_gsrc, (int)g.ndom);
  }
  // ghost T cells outside a closed or open domain edge are never ice (tmask is false there) and hold fill values: not checked.
  // A tripole top row is checked like any other; its sign-flipped dxhy/dyhx make the check fail and the arrays stay in use.
  const bool e_edge = (g.gi0 + g.dom.nx - 1 == g.nxg), n_edge = (g.gj0 + g.dom.ny - 1 == g.nyg);
  const int skip_e = (e_edge && g.ew != EVP_B200_BNDY_CYCLIC) ? 1 : 0;
  const int skip_n = (n_edge && g.ns != EVP_B200_BNDY_CYCLIC && g.ns != EVP_B200_BNDY_TRIPOLE) ? 1 : 0;
  CK(exact::launch_metric_verify(g.dom, dm[0], dm[1], deltaminEVP, skip_e, skip_n, dcount, g.stream));
  int bad = -1;
  CK(cudaMemcpyAsync(&bad, dcount, sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaFree(scratch));
  CK(cudaFree(dcount));
  g.metric_mismatches = bad;
  g.derived_ok = (bad == 0);   // (cells next to an eliminated land block fail the check by themselves: the hole holds zeros)
  if (g.derived_ok) {
    CK(exact::set_metric(dm[0], dm[1], deltaminEVP));
    CK(fast::set_metric(dm[0], dm[1], deltaminEVP));
    g.d_HTN = dm[0]; g.d_HTE = dm[1]; g.deltamin = deltaminEVP;
    g.ts_key[0] = nullptr;   // tensor maps of an earlier metric are stale
    destroy_graph();  // a cached graph may hold the array-reading form
  } else {
    g.d_HTN = g.d_HTE = nullptr;
  }
  if (mismatches) *mismatches = bad;
  char b[160];
  snprintf(b, sizeof b, "; metric: %s (%d cells differ)", g.derived_ok ? "derived geometry available" : "arrays kept", bad);
  g.desc += b;
  return 0;
}

static int do_deformations(evp_b200_deform_t *dd) {
  if (!g.inited || !g.uploaded) return fail("evp_b200_deformations: no velocities on the device (run the loop first)");
  if (!dd || !dd->dxU || !dd->dyU || !dd->tarear || !dd->divu || !dd->shear || !dd->vort || !dd->rdg_conv || !dd->rdg_shear)
    return fail("evp_b200_deformations: null argument");
  CK(cudaSetDevice(g.device));
  const size_t bblk = g.nblk_elems * sizeof(double);
  if (ensure_aux_buffers()) return 1;
  const double *src[8] = {dd->dxU, dd->dyU, dd->tarear, dd->divu, dd->shear, dd->vort, dd->rdg_conv, dd->rdg_shear};
  double *dst[5] = {dd->divu, dd->shear, dd->vort, dd->rdg_conv, dd->rdg_shear};
  for (int q = 0; q < 8; ++q) {
    CK(cudaMemcpyAsync(g.dstage[q], src[q], bblk, cudaMemcpyHostToDevice, g.stream));
    pack_f64<<<grid_blocks(g.ndom), 256, 0, g.stream>>>(g.dbuf[q], g.dstage[q], g.d_gsrc, (int)g.ndom);
  }
  // the exact build: the reference's `deformations` has no contraction-sensitive state, but keep one answer
  CK(exact::launch_deform(g.dom, g.cur, g.dbuf[0], g.dbuf[1], g.dbuf[2], g.dbuf[3], g.dbuf[4], g.dbuf[5], g.dbuf[6], g.dbuf[7],
                          dd->e_factor, g.stream));
  for (int q = 0; q < 5; ++q) {
    unpack_f64<<<grid_blocks(g.n_sig), 256, 0, g.stream>>>(g.dstage[3 + q], g.dbuf[3 + q], g.d_sig_lin, g.d_sig_dom, g.n_sig);
    CK(cudaMemcpyAsync(dst[q], g.dstage[3 + q], bblk, cudaMemcpyDeviceToHost, g.stream));
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

static int do_init_cgrid(const evp_b200_cgrid_t *cg) {
  if (!g.inited) return fail("evp_b200_init_cgrid: call evp_b200_init first");
  if (!cg) return fail("evp_b200_init_cgrid: null grid");
  if (g_comm.nranks > 1) return fail("evp_b200_init_cgrid: the C-grid path runs on one GPU in this version");
  if (g.ns == EVP_B200_BNDY_TRIPOLE) return fail("evp_b200_init_cgrid: tripole grids are not supported on the C-grid path");
  if (g.halo.n_dst != 0) return fail("evp_b200_init_cgrid: the halo of this decomposition is not an on-rank wrap");
  if (g.cinit) return fail("evp_b200_init_cgrid: already initialised");
  CK(cudaSetDevice(g.device));
  CDom &c = g.cdom;
  c.nx = g.dom.nx; c.ny = g.dom.ny; c.ld = g.dom.ld; c.nyd = g.dom.nyd; c.wrap_ew = g.dom.wrap_ew; c.wrap_ns = g.dom.wrap_ns;
  const size_t bblk = g.nblk_elems * sizeof(double);
  // static geometry: 20 arrays from the caller, 3 shared with the B-grid set
  const double *src[20] = {cg->dxN, cg->dyE, cg->dxE, cg->dyN, cg->dxU, cg->dyU, cg->tarea, cg->uarea, cg->earea, cg->narea,
                           cg->earear, cg->narear, cg->ratiodxN, cg->ratiodxNr, cg->ratiodyE, cg->ratiodyEr, cg->hm, cg->uvm, cg->epm, cg->npm};
  const double **dst[20] = {&c.dxN, &c.dyE, &c.dxE, &c.dyN, &c.dxU, &c.dyU, &c.tarea, &c.uarea, &c.earea, &c.narea,
                            &c.earear, &c.narear, &c.ratiodxN, &c.ratiodxNr, &c.ratiodyE, &c.ratiodyEr, &c.hm, &c.uvm, &c.epm, &c.npm};
  double *scratch = nullptr;
  CK(cudaMalloc(&scratch, bblk));
  for (int q = 0; q < 20; ++q) {
    if (!src[q]) { cudaFree(scratch); return fail("evp_b200_init_cgrid: null geometry array %d", q); }
    double *p = nullptr;
    if (calloc_dom(p)) return 1;
    // scratch staging of its own (freed below): g.stage may hold an upload that a later evp_b200_download builds on
    CK(cudaMemcpyAsync(scratch, src[q], bblk, cudaMemcpyHostToDevice, g.stream));
    pack_f64<<<grid_blocks(g.ndom), 256, 0, g.stream>>>(p, scratch, g.d_gsrc, (int)g.ndom);
    *dst[q] = p;
  }
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaFree(scratch));
  c.dxT = g.dom.dxT; c.dyT = g.dom.dyT; c.DminTarea = g.dom.DminTarea;
  {
    double *q[5] = {};
    for (auto &p : q) if (calloc_dom(p)) return 1;
    CK(exact::launch_cgrid_static(c, q[0], q[1], q[2], q[3], q[4], g.stream));
    c.rhalf_dyE = q[0]; c.r_dxE = q[1]; c.rhalf_dxN = q[2]; c.r_dyN = q[3]; c.uareaavgr = q[4];
  }
  // 43 time-varying fields, each with a staging buffer in block layout
  double **flds[43] = {&c.uvelE, &c.vvelE, &c.uvelN, &c.vvelN, &c.uvel, &c.vvel, &c.stresspT, &c.stressmT, &c.stress12T, &c.stress12U,
                       &c.zetax2T, &c.etax2T, &c.etax2U, &c.strengthU, &c.divergU, &c.tensionU, &c.shearU, &c.deltaU,
                       &c.strintxE, &c.strintyN, &c.taubxE, &c.taubyN,
                       (double **)&c.strength, (double **)&c.cdnE, (double **)&c.cdnN, (double **)&c.aiE, (double **)&c.aiN,
                       (double **)&c.uocnE, (double **)&c.vocnE, (double **)&c.uocnN, (double **)&c.vocnN, (double **)&c.waterxE,
                       (double **)&c.wateryN, (double **)&c.forcexE, (double **)&c.forceyN, (double **)&c.emassdti, (double **)&c.nmassdti,
                       (double **)&c.fmE, (double **)&c.fmN, (double **)&c.TbE, (double **)&c.TbN, (double **)&c.rheofactE, (double **)&c.rheofactN};
  for (int q = 0; q < 43; ++q) {
    if (calloc_dom(*flds[q])) return 1;
    double *st = nullptr;
    CK(cudaMalloc(&st, bblk));
    g.cstage.push_back(st);
  }
  if (calloc_dom(c.uvelE_init) || calloc_dom(c.vvelN_init) || calloc_dom(c.stress12Ub)) return 1;
  for (int q = 0; q < 4; ++q) { CK(cudaMalloc(&g.cmask[q], g.ndom)); CK(cudaMemsetAsync(g.cmask[q], 0, g.ndom, g.stream)); }
  c.maskT = g.cmask[0]; c.maskU = g.cmask[1]; c.maskE = g.cmask[2]; c.maskN = g.cmask[3];
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g.stream));
  g.cinit = true;
  return 0;
}

static int do_run_cgrid(const evp_b200_params_t *p, evp_b200_cfields_t *f) {
  if (!g.cinit) return fail("evp_b200_run_cgrid: evp_b200_init_cgrid has not been called");
  if (!p || !f) return fail("evp_b200_run_cgrid: null argument");
  if (p->ndte < 0) return fail("evp_b200_run_cgrid: ndte < 0");
  if (p->visc_method != EVP_B200_VISC_AVG_ZETA && p->visc_method != EVP_B200_VISC_AVG_STRENGTH) return fail("evp_b200_run_cgrid: unknown visc_method %d", p->visc_method);
  if (p->mode != EVP_B200_MODE_EXACT && p->mode != EVP_B200_MODE_FAST) return fail("evp_b200_run_cgrid: unknown mode %d", p->mode);
  CK(cudaSetDevice(g.device));
  CDom &c = g.cdom;
  g.c_fused = (p->kernel != EVP_B200_KERNEL_SPLIT);  // SPLIT: the five-kernel first form, cut where the reference has a halo point
  const size_t bblk = g.nblk_elems * sizeof(double), bdom = g.ndom * sizeof(double);
  // host pointer, device array, how it moves: 'i' in, 'r' inout ring, 's' inout T cells the loop owns (N/E ghost incl.),
  // 'n' inout interior, 'z' out: the reference zero-fills the whole block, writes interiors, halo-updates;
  // 'y' out: zero-filled whole block, interiors written, NOT halo-updated (block ghost cells stay 0);
  // 'R' like 'r' but the reference zero-fills the whole block (padding included) every subcycle before it
  //     re-interpolates and halo-updates (grid_average_X2YA, ice_grid.F90:4410)
  struct Fld { const double *h; double *dv; char kind; };
  const bool avgstr = (p->visc_method == EVP_B200_VISC_AVG_STRENGTH);
  Fld tab[43] = {
      {f->uvelE, c.uvelE, 'r'}, {f->vvelE, c.vvelE, 'R'}, {f->uvelN, c.uvelN, 'R'}, {f->vvelN, c.vvelN, 'r'}, {f->uvel, c.uvel, 'R'}, {f->vvel, c.vvel, 'R'},
      {f->stresspT, c.stresspT, 'r'}, {f->stressmT, c.stressmT, 'r'}, {f->stress12T, c.stress12T, 's'}, {f->stress12U, c.stress12U, 'r'},
      {f->zetax2T, c.zetax2T, 'r'}, {f->etax2T, c.etax2T, 'r'}, {f->etax2U, c.etax2U, avgstr ? '-' : 'y'}, {f->strengthU, c.strengthU, avgstr ? 'y' : '-'},
      {f->divergU, c.divergU, 'y'}, {f->tensionU, c.tensionU, 'y'}, {f->shearU, c.shearU, 'z'}, {f->deltaU, c.deltaU, 'y'},
      {f->strintxE, c.strintxE, 'n'}, {f->strintyN, c.strintyN, 'n'}, {f->taubxE, c.taubxE, 'n'}, {f->taubyN, c.taubyN, 'n'},
      {f->strength, (double *)c.strength, 'i'}, {f->cdn_ocnE, (double *)c.cdnE, 'i'}, {f->cdn_ocnN, (double *)c.cdnN, 'i'}, {f->aiE, (double *)c.aiE, 'i'},
      {f->aiN, (double *)c.aiN, 'i'}, {f->uocnE, (double *)c.uocnE, 'i'}, {f->vocnE, (double *)c.vocnE, 'i'}, {f->uocnN, (double *)c.uocnN, 'i'},
      {f->vocnN, (double *)c.vocnN, 'i'}, {f->waterxE, (double *)c.waterxE, 'i'}, {f->wateryN, (double *)c.wateryN, 'i'}, {f->forcexE, (double *)c.forcexE, 'i'},
      {f->forceyN, (double *)c.forceyN, 'i'}, {f->emassdti, (double *)c.emassdti, 'i'}, {f->nmassdti, (double *)c.nmassdti, 'i'}, {f->fmE, (double *)c.fmE, 'i'},
      {f->fmN, (double *)c.fmN, 'i'}, {f->TbE, (double *)c.TbE, 'i'}, {f->TbN, (double *)c.TbN, 'i'}, {f->rheofactE, (double *)c.rheofactE, 'i'},
      {f->rheofactN, (double *)c.rheofactN, 'i'}};
  for (int q = 0; q < 43; ++q) {
    const Fld &t = tab[q];
    if (t.kind == '-') continue;
    if (!t.h) return fail("evp_b200_run_cgrid: null field %d", q);
    if (t.kind == 'z' || t.kind == 'y') {
      CK(cudaMemsetAsync(t.dv, 0, bdom, g.stream));
      CK(cudaMemsetAsync(g.cstage[q], 0, bblk, g.stream));
      continue;
    }
    CK(cudaMemcpyAsync(g.cstage[q], t.h, bblk, cudaMemcpyHostToDevice, g.stream));
    if (t.dv == c.stress12U)  // ping-ponged by the fused form: both copies start identical
      pack2_f64<<<grid_blocks(g.ndom), 256, 0, g.stream>>>(c.stress12U, c.stress12Ub, g.cstage[q], g.d_gsrc, (int)g.ndom);
    else
      pack_f64<<<grid_blocks(g.ndom), 256, 0, g.stream>>>(t.dv, g.cstage[q], g.d_gsrc, (int)g.ndom);
  }
  const int32_t *hm[4] = {f->iceTmask, f->iceUmask, f->iceEmask, f->iceNmask};
  for (int q = 0; q < 4; ++q) {
    if (!hm[q]) return fail("evp_b200_run_cgrid: null mask %d", q);
    CK(cudaMemcpyAsync(g.stage_mask, hm[q], g.nblk_elems * sizeof(int), cudaMemcpyHostToDevice, g.stream));
    pack_mask<<<grid_blocks(g.ndom), 256, 0, g.stream>>>(g.cmask[q], g.stage_mask, g.d_gsrc, (int)g.ndom);
  }
  // uvelE_init / vvelN_init = values at entry (dyn_prep2, shared.F90:787-788)
  CK(cudaMemcpyAsync(c.uvelE_init, c.uvelE, bdom, cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaMemcpyAsync(c.vvelN_init, c.vvelN, bdom, cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaGetLastError());

  // the loop: 5 kernels per subcycle, captured once per parameter set
  const KParams k = kparams(p);
  const bool exact = (p->mode == EVP_B200_MODE_EXACT);
  cudaGraphExec_t &cexec = g.cexec;
  evp_b200_params_t &cparams = g.cparams;
  int nl = 0;
  if (!cexec || memcmp(&cparams, p, sizeof *p) != 0) {
    if (cexec) { cudaGraphExecDestroy(cexec); cexec = nullptr; }
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t le = cudaSuccess;
    if (g.c_fused) {
      for (int ksub = 0; ksub < p->ndte && le == cudaSuccess; ++ksub)
        le = exact ? exact::launch_cgrid_subcycle_fused(c, k, ksub & 1, g.stream, &nl) : fast::launch_cgrid_subcycle_fused(c, k, ksub & 1, g.stream, &nl);
    } else {
      for (int ksub = 0; ksub < p->ndte && le == cudaSuccess; ++ksub)
        le = exact ? exact::launch_cgrid_subcycle(c, k, g.stream, &nl) : fast::launch_cgrid_subcycle(c, k, g.stream, &nl);
    }
    g.cgraph_launches = nl;
    cudaError_t ce = cudaStreamEndCapture(g.stream, &graph);
    if (le != cudaSuccess || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return fail("evp_b200_run_cgrid: capture failed: %s", cudaGetErrorString(le != cudaSuccess ? le : ce)); }
    CK(cudaGraphInstantiate(&cexec, graph, 0));
    CK(cudaGraphDestroy(graph));
    cparams = *p;
  } else {
    nl = g.cgraph_launches;
  }
  CK(cudaEventRecord(g.ev0, g.stream));
  if (p->ndte > 0) CK(cudaGraphLaunch(cexec, g.stream));
  CK(cudaEventRecord(g.ev1, g.stream));
  g.last_launches = nl;

  // back to the caller's block arrays
  for (int q = 0; q < 22; ++q) {
    const Fld &t = tab[q];
    if (t.kind == '-') continue;
    if (t.kind == 'R' && p->ndte > 0) CK(cudaMemsetAsync(g.cstage[q], 0, bblk, g.stream));
    if (t.kind == 's')
      unpack_f64<<<grid_blocks(g.n_sig), 256, 0, g.stream>>>(g.cstage[q], t.dv, g.d_sig_lin, g.d_sig_dom, g.n_sig);
    else if (t.kind == 'n' || t.kind == 'y')
      unpack_f64<<<grid_blocks(g.n_int), 256, 0, g.stream>>>(g.cstage[q], t.dv, g.d_int_lin, g.d_int_dom, g.n_int);
    else {
      const bool in_b = (t.dv == c.stress12U) && g.c_fused && (p->ndte & 1);
      unpack_f64<<<grid_blocks(g.n_uv), 256, 0, g.stream>>>(g.cstage[q], in_b ? c.stress12Ub : t.dv, g.d_uv_lin, g.d_uv_dom, g.n_uv);
    }
    CK(cudaMemcpyAsync((void *)t.h, g.cstage[q], bblk, cudaMemcpyDeviceToHost, g.stream));
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// CD grid: same life cycle as the C grid, four kernels per subcycle
// ------------------------------------------------------------------------------------------------
static int do_run_cdgrid(const evp_b200_params_t *p, evp_b200_cdfields_t *f) {
  if (!g.cinit) return fail("evp_b200_run_cdgrid: evp_b200_init_cgrid has not been called (the CD grid uses the C grid's geometry)");
  if (!p || !f) return fail("evp_b200_run_cdgrid: null argument");
  if (p->ndte < 0) return fail("evp_b200_run_cdgrid: ndte < 0");
  if (p->visc_method != EVP_B200_VISC_AVG_ZETA && p->visc_method != EVP_B200_VISC_AVG_STRENGTH) return fail("evp_b200_run_cdgrid: unknown visc_method %d", p->visc_method);
  if (p->mode != EVP_B200_MODE_EXACT && p->mode != EVP_B200_MODE_FAST) return fail("evp_b200_run_cdgrid: unknown mode %d", p->mode);
  CK(cudaSetDevice(g.device));
  CDom &c = g.cdom;
  const size_t bblk = g.nblk_elems * sizeof(double), bdom = g.ndom * sizeof(double);
  constexpr int NCD = 54;
  if (!g.cdinit) {
    double **extra[13] = {&c.stresspU, &c.stressmU, &c.zetax2U, &c.strintyE, &c.strintxN, &c.taubyE, &c.taubxN, &c.vvelE_init, &c.uvelN_init,
                          (double **)&c.wateryE, (double **)&c.waterxN, (double **)&c.forceyE, (double **)&c.forcexN};
    for (auto pp : extra) if (calloc_dom(*pp)) return 1;
    for (int q = 0; q < NCD; ++q) {
      double *st = nullptr;
      CK(cudaMalloc(&st, bblk));
      g.cdstage.push_back(st);
    }
    g.cdinit = true;
  }
  // kinds as in do_run_cgrid: 'i' in; 'r' inout, ring refreshed by the loop; 'R' like 'r' but zero-filled whole block every subcycle
  // (grid_average output); 'n' inout interior; 'y' out: zero-filled, interiors written, not halo-updated; '-' untouched
  struct Fld { const double *h; double *dv; char kind; };
  const bool avgstr = (p->visc_method == EVP_B200_VISC_AVG_STRENGTH);
  Fld tab[NCD] = {
      {f->uvelE, c.uvelE, 'r'}, {f->vvelE, c.vvelE, 'r'}, {f->uvelN, c.uvelN, 'r'}, {f->vvelN, c.vvelN, 'r'}, {f->uvel, c.uvel, 'R'}, {f->vvel, c.vvel, 'R'},
      {f->stresspT, c.stresspT, 'r'}, {f->stressmT, c.stressmT, 'r'}, {f->stress12T, c.stress12T, 'r'},
      {f->stresspU, c.stresspU, 'r'}, {f->stressmU, c.stressmU, 'r'}, {f->stress12U, c.stress12U, 'r'},
      {f->zetax2T, c.zetax2T, 'r'}, {f->etax2T, c.etax2T, 'r'}, {f->zetax2U, c.zetax2U, avgstr ? '-' : 'y'}, {f->etax2U, c.etax2U, avgstr ? '-' : 'y'},
      {f->strengthU, c.strengthU, avgstr ? 'y' : '-'},
      {f->divergU, c.divergU, 'y'}, {f->tensionU, c.tensionU, 'y'}, {f->shearU, c.shearU, 'y'}, {f->deltaU, c.deltaU, 'y'},
      {f->strintxE, c.strintxE, 'n'}, {f->strintyE, c.strintyE, 'n'}, {f->strintxN, c.strintxN, 'n'}, {f->strintyN, c.strintyN, 'n'},
      {f->taubxE, c.taubxE, 'n'}, {f->taubyE, c.taubyE, 'n'}, {f->taubxN, c.taubxN, 'n'}, {f->taubyN, c.taubyN, 'n'},
      {f->strength, (double *)c.strength, 'i'}, {f->cdn_ocnE, (double *)c.cdnE, 'i'}, {f->cdn_ocnN, (double *)c.cdnN, 'i'}, {f->aiE, (double *)c.aiE, 'i'},
      {f->aiN, (double *)c.aiN, 'i'}, {f->uocnE, (double *)c.uocnE, 'i'}, {f->vocnE, (double *)c.vocnE, 'i'}, {f->uocnN, (double *)c.uocnN, 'i'},
      {f->vocnN, (double *)c.vocnN, 'i'}, {f->waterxE, (double *)c.waterxE, 'i'}, {f->wateryE, (double *)c.wateryE, 'i'},
      {f->waterxN, (double *)c.waterxN, 'i'}, {f->wateryN, (double *)c.wateryN, 'i'}, {f->forcexE, (double *)c.forcexE, 'i'},
      {f->forceyE, (double *)c.forceyE, 'i'}, {f->forcexN, (double *)c.forcexN, 'i'}, {f->forceyN, (double *)c.forceyN, 'i'},
      {f->emassdti, (double *)c.emassdti, 'i'}, {f->nmassdti, (double *)c.nmassdti, 'i'}, {f->fmE, (double *)c.fmE, 'i'}, {f->fmN, (double *)c.fmN, 'i'},
      {f->TbE, (double *)c.TbE, 'i'}, {f->TbN, (double *)c.TbN, 'i'}, {f->rheofactE, (double *)c.rheofactE, 'i'}, {f->rheofactN, (double *)c.rheofactN, 'i'}};
  for (int q = 0; q < NCD; ++q) {
    const Fld &t = tab[q];
    if (t.kind == '-') continue;
    if (!t.h) return fail("evp_b200_run_cdgrid: null field %d", q);
    if (t.kind == 'y') {
      CK(cudaMemsetAsync(t.dv, 0, bdom, g.stream));
      CK(cudaMemsetAsync(g.cdstage[q], 0, bblk, g.stream));
      continue;
    }
    CK(cudaMemcpyAsync(g.cdstage[q], t.h, bblk, cudaMemcpyHostToDevice, g.stream));
    pack_f64<<<grid_blocks(g.ndom), 256, 0, g.stream>>>(t.dv, g.cdstage[q], g.d_gsrc, (int)g.ndom);
  }
  const int32_t *hm[4] = {f->iceTmask, f->iceUmask, f->iceEmask, f->iceNmask};
  for (int q = 0; q < 4; ++q) {
    if (!hm[q]) return fail("evp_b200_run_cdgrid: null mask %d", q);
    CK(cudaMemcpyAsync(g.stage_mask, hm[q], g.nblk_elems * sizeof(int), cudaMemcpyHostToDevice, g.stream));
    pack_mask<<<grid_blocks(g.ndom), 256, 0, g.stream>>>(g.cmask[q], g.stage_mask, g.d_gsrc, (int)g.ndom);
  }
  // velocities at entry (dyn_prep2 at E and N, shared.F90:787-788)
  CK(cudaMemcpyAsync(c.uvelE_init, c.uvelE, bdom, cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaMemcpyAsync(c.vvelE_init, c.vvelE, bdom, cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaMemcpyAsync(c.uvelN_init, c.uvelN, bdom, cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaMemcpyAsync(c.vvelN_init, c.vvelN, bdom, cudaMemcpyDeviceToDevice, g.stream));
  CK(cudaGetLastError());

  const KParams k = kparams(p);
  const bool exact = (p->mode == EVP_B200_MODE_EXACT);
  int nl = 0;
  if (!g.cdexec || memcmp(&g.cdparams, p, sizeof *p) != 0) {
    if (g.cdexec) { cudaGraphExecDestroy(g.cdexec); g.cdexec = nullptr; }
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t le = cudaSuccess;
    for (int ksub = 0; ksub < p->ndte && le == cudaSuccess; ++ksub)
      le = exact ? exact::launch_cdgrid_subcycle(c, k, g.stream, &nl) : fast::launch_cdgrid_subcycle(c, k, g.stream, &nl);
    cudaError_t ce = cudaStreamEndCapture(g.stream, &graph);
    if (le != cudaSuccess || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return fail("evp_b200_run_cdgrid: capture failed: %s", cudaGetErrorString(le != cudaSuccess ? le : ce)); }
    CK(cudaGraphInstantiate(&g.cdexec, graph, 0));
    CK(cudaGraphDestroy(graph));
    g.cdparams = *p;
  } else {
    nl = 4 * p->ndte;
  }
  CK(cudaEventRecord(g.ev0, g.stream));
  if (p->ndte > 0) CK(cudaGraphLaunch(g.cdexec, g.stream));
  CK(cudaEventRecord(g.ev1, g.stream));
  g.last_launches = nl;

  for (int q = 0; q < 29; ++q) {   // everything that is not 'i'
    const Fld &t = tab[q];
    if (t.kind == '-') continue;
    if (t.kind == 'R' && p->ndte > 0) CK(cudaMemsetAsync(g.cdstage[q], 0, bblk, g.stream));
    if (t.kind == 'n' || t.kind == 'y')
      unpack_f64<<<grid_blocks(g.n_int), 256, 0, g.stream>>>(g.cdstage[q], t.dv, g.d_int_lin, g.d_int_dom, g.n_int);
    else
      unpack_f64<<<grid_blocks(g.n_uv), 256, 0, g.stream>>>(g.cdstage[q], t.dv, g.d_uv_lin, g.d_uv_dom, g.n_uv);
    CK(cudaMemcpyAsync((void *)t.h, g.cdstage[q], bblk, cudaMemcpyDeviceToHost, g.stream));
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g.stream));
  CK(cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1));
  return 0;
}

}  // namespace evp

// ------------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------------
using namespace evp;

extern "C" {

const char *evp_b200_last_error(void) { return g_err; }
const char *evp_b200_describe(void) { return g.desc.c_str(); }

int evp_b200_set_device(int32_t dev) {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (dev < 0 || dev >= n) return fail("evp_b200_set_device: device %d of %d", dev, n);
  if (g.inited && dev != g.device) return fail("evp_b200_set_device: already initialised on device %d", g.device);
  g.device = dev;
  CK(cudaSetDevice(dev));
  return 0;
}

int evp_b200_allow_partial_domain(int32_t yes) { g_allow_partial = (yes != 0); return 0; }

int evp_b200_get_unique_id(void *id128) { return comm_get_unique_id(id128, g_err, sizeof g_err); }

int evp_b200_comm_init(int32_t rank, int32_t nranks, const void *id128) {
  if (g.inited) return fail("evp_b200_comm_init: call before evp_b200_init");
  if (g.device < 0) CK(cudaGetDevice(&g.device));
  CK(cudaSetDevice(g.device));
  return comm_init(g_comm, rank, nranks, id128, g_err, sizeof g_err);
}

int evp_b200_init(const evp_b200_grid_t *grid) {
  int rc = do_init(grid);
  if (rc) { std::string keep = g_err; free_all(); snprintf(g_err, sizeof g_err, "%s", keep.c_str()); }
  return rc;
}

int evp_b200_finalize(void) {
  free_all();
  comm_destroy(g_comm);
  return 0;
}

int evp_b200_deformations(evp_b200_deform_t *d) { return do_deformations(d); }
int evp_b200_pin_host(void *ptr, size_t bytes) {
  if (!ptr || !bytes) return fail("evp_b200_pin_host: null argument");
  if (g.device >= 0) CK(cudaSetDevice(g.device));
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
  if (e != cudaSuccess) return fail("evp_b200_pin_host: %s", cudaGetErrorString(e));
  return 0;
}
int evp_b200_unpin_host(void *ptr) {
  if (!ptr) return fail("evp_b200_unpin_host: null argument");
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return 0; }
  if (e != cudaSuccess) return fail("evp_b200_unpin_host: %s", cudaGetErrorString(e));
  return 0;
}
int evp_b200_dyn_finish(evp_b200_finish_t *f) { return do_dyn_finish(f); }
int evp_b200_set_metric(const double *HTN, const double *HTE, double deltaminEVP, int32_t *mismatches) {
  return do_set_metric(HTN, HTE, deltaminEVP, mismatches);
}
int evp_b200_init_cgrid(const evp_b200_cgrid_t *cg) { return do_init_cgrid(cg); }
int evp_b200_run_cgrid(const evp_b200_params_t *p, evp_b200_cfields_t *f) { return do_run_cgrid(p, f); }
int evp_b200_run_cdgrid(const evp_b200_params_t *p, evp_b200_cdfields_t *f) { return do_run_cdgrid(p, f); }

int evp_b200_upload(const evp_b200_fields_t *f) {
  if (do_upload(f, false)) return 1;
  // "keeps no host pointer past the return of a call": the copies out of the caller's arrays are complete when this returns
  // (run_bgrid keeps them asynchronous internally; its download synchronises)
  CK(cudaStreamSynchronize(g.xfer));
  return 0;
}
int evp_b200_subcycle(const evp_b200_params_t *p) { return do_subcycle(p); }
int evp_b200_download(evp_b200_fields_t *f) { return do_download(f, 7); }

// Tripole grids: "force symmetry across the tripole seam" (ice_dyn_evp.F90:1321-1388) on the stresses the device holds.  Each of the
// twelve ice_HaloUpdate_stress(array1, array2) calls writes the north ghost row of array1 on the blocks of the top row -- every local
// column, ghost columns included -- from the top PHYSICAL row of array2 at the mirrored global column nx_global - i_glob + 1
// (ice_boundary.F90:7760-7794, addresses :8117-8157; centre location, scalar: no offsets, no sign).  Pairs 1<->3, 2<->4 of stressp,
// stressm, stress12.  Reads row ny, writes row ny+1 of the same copy: one thread per (array, ghost cell), no ordering needed.
__global__ void stress_fold_kernel(const __grid_constant__ Dom d, int cur, int gi0, int nxg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y;
  if (i > d.nx + 1) return;
  const int ic = fold_mirror_col(nxg, gi0, i) - gi0 + 1;   // dom column of the mirrored cell (the rank holds the whole top row)
  const int src = (q / 4) * 4 + ((q % 4) + 2) % 4;
  d.sig[cur][q][(size_t)(d.ny + 1) * d.ld + i] = d.sig[cur][src][(size_t)d.ny * d.ld + ic];
}
// the same with the mirrored cells taken from the gathered top row (rowtop[12][nxg]): the top row is spread over several ranks
__global__ void stress_fold_rows_kernel(const __grid_constant__ Dom d, int cur, int gi0, int nxg, const double *__restrict__ rowtop) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y;
  if (i > d.nx + 1) return;
  const int src = (q / 4) * 4 + ((q % 4) + 2) % 4;
  d.sig[cur][q][(size_t)(d.ny + 1) * d.ld + i] = rowtop[(size_t)src * nxg + fold_mirror_col(nxg, gi0, i) - 1];
}
static bool symmetrise_here() { return g.ns == EVP_B200_BNDY_TRIPOLE && g.gj0 + g.dom.ny - 1 == g.nyg; }
static int do_stress_symmetrise() {
  if (!g.inited || !g.uploaded) return fail("evp_b200_stress_symmetrise: no stresses on the device");
  if (g.ns != EVP_B200_BNDY_TRIPOLE) return 0;
  if (symmetrise_here()) {
    CK(cudaSetDevice(g.device));
    dim3 b(128), gr((g.dom.nx + 2 + b.x - 1) / b.x, 12);
    if (g.gi0 == 1 && g.dom.nx == g.nxg) {   // the whole top row is this rank's
      stress_fold_kernel<<<gr, b, 0, g.stream>>>(g.dom, g.cur, g.gi0, g.nxg);
    } else {                                 // gather it from the other ranks of the top row (NCCL, once per step)
      if (g_comm.nranks < 2) return fail("evp_b200_stress_symmetrise: the rank holds a part of the tripole top row and there is no communicator");
      if (!g.d_symrow) CK(cudaMalloc(&g.d_symrow, sizeof(double) * 12 * (size_t)g.nxg));
      if (stress_rows_exchange(g_comm, g.halo.rects, g.nxg, g.nyg, g.dom.sig[g.cur], g.dom.ld, g.dom.nx, g.dom.ny, g.gi0, g.gj0, g.d_symrow, g.stream,
                               g_err, sizeof g_err))
        return 1;
      stress_fold_rows_kernel<<<gr, b, 0, g.stream>>>(g.dom, g.cur, g.gi0, g.nxg, g.d_symrow);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(g.stream));
  }
  g.sym_applied = true;
  return 0;
}
int evp_b200_stress_symmetrise(void) { return do_stress_symmetrise(); }

int evp_b200_run_bgrid_resident(const evp_b200_params_t *p, evp_b200_fields_t *f, int32_t flags) {
  const bool sym = (flags & EVP_B200_KEEP_STRESS) && g.inited && g.ns == EVP_B200_BNDY_TRIPOLE;
  if (do_upload(f, (flags & EVP_B200_KEEP_STRESS) != 0)) return 1;
  if (do_subcycle(p)) return 1;
  if (sym && do_stress_symmetrise()) return 1;
  const bool stress_back = !(flags & EVP_B200_KEEP_STRESS) || (flags & EVP_B200_FETCH_STRESS);
  return do_download(f, stress_back ? 7 : 6);
}
int evp_b200_download_stress(evp_b200_fields_t *f) { return do_download(f, 1); }

// ---- the step preparation on the device --------------------------------------------------------------
__global__ void unpack_mask(int *__restrict__ dst_blk, const unsigned char *__restrict__ src_dom, const int *__restrict__ lin,
                            const int *__restrict__ dom, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) dst_blk[lin[k]] = src_dom[dom[k]] ? -1 : 0;  // .true.
}

int evp_b200_prep_init(const evp_b200_prep_static_t *st) {
  if (!g.inited) return fail("evp_b200_prep_init: call evp_b200_init first");
  if (!st || !st->hm || !st->tarea || !st->uarea || !st->fcor || !st->umask) return fail("evp_b200_prep_init: null argument");
  if (g.ns == EVP_B200_BNDY_TRIPOLE || g.halo.has_fold)
    return fail("evp_b200_prep_init: not for tripole grids in this version (the T->U averages and the velocity halo update after dyn_prep2 "
                "at the fold are not built; evp_b200_run_bgrid_resident keeps the stresses on the device there)");
  CK(cudaSetDevice(g.device));
  const size_t bdom = g.ndom * sizeof(double), bblk = g.nblk_elems * sizeof(double);
  const int nd = (int)g.ndom, nb = grid_blocks(g.ndom);
  const double *src[4] = {st->hm, st->tarea, st->uarea, st->fcor};
  for (int q = 0; q < 4; ++q) {
    if (!g.prep_static[q]) CK(cudaMalloc(&g.prep_static[q], bdom));
    CK(cudaMemcpyAsync(g.stage[q], src[q], bblk, cudaMemcpyHostToDevice, g.stream));
    pack_f64<<<nb, 256, 0, g.stream>>>(g.prep_static[q], g.stage[q], g.d_gsrc, nd);
  }
  for (auto &p : g.prepT)
    if (!p) { CK(cudaMalloc(&p, bdom)); CK(cudaMemsetAsync(p, 0, bdom, g.stream)); }
  if (!g.prep_umask) CK(cudaMalloc(&g.prep_umask, g.ndom));
  CK(cudaMemcpyAsync(g.stage_mask, st->umask, g.nblk_elems * sizeof(int), cudaMemcpyHostToDevice, g.stream));
  pack_mask<<<nb, 256, 0, g.stream>>>(g.prep_umask, g.stage_mask, g.d_gsrc, nd);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g.stream));
  g.prep_ok = true;
  g.state_resident = false;
  return 0;
}

int evp_b200_step_resident(const evp_b200_params_t *p, const evp_b200_prep_t *pr, evp_b200_fields_t *f, int32_t flags) {
  if (!g.inited || !g.prep_ok) return fail("evp_b200_step_resident: call evp_b200_init and evp_b200_prep_init first");
  if (!p || !pr || !f) return fail("evp_b200_step_resident: null argument");
  if (!pr->tmass || !pr->aice_init || !pr->cdn_ocn || !pr->uocn || !pr->vocn || !pr->strairxT || !pr->strairyT || !pr->strength || !pr->iceTmask)
    return fail("evp_b200_step_resident: null input array");
  if (pr->ssh_stress != 0 && pr->ssh_stress != 1) return fail("evp_b200_step_resident: ssh_stress %d (0 geostrophic, 1 coupled)", pr->ssh_stress);
  if (pr->ssh_stress == 1 && (!pr->ss_tltx || !pr->ss_tlty)) return fail("evp_b200_step_resident: coupled ssh_stress needs ss_tltx, ss_tlty");
  if (!(pr->dt > 0.0)) return fail("evp_b200_step_resident: dt must be positive");
  if (!f->uvel || !f->vvel) return fail("evp_b200_step_resident: fields->uvel / vvel are written every step");
  CK(cudaSetDevice(g.device));
  const size_t bblk = g.nblk_elems * sizeof(double);
  const int nd = (int)g.ndom, nb = grid_blocks(g.ndom);
  const bool init = (flags & EVP_B200_STEP_INIT_STATE) || !g.state_resident;
  // everything below is packed into copy 0 (and 1) of the carried state: normalise as do_subcycle would
  if (g.cur == 1) {
    std::swap(g.dom.u[0], g.dom.u[1]);
    std::swap(g.dom.v[0], g.dom.v[1]);
    for (int q = 0; q < 12; ++q) std::swap(g.dom.sig[0][q], g.dom.sig[1][q]);
    g.cur = 0;
    if (g.p2p.enabled) g.p2p.set_parity(g.p2p.swapped ^ 1);   // the neighbours' copies are addressed by the same parity
    destroy_graph();
  }
  if (init) {
    // the carried state comes from the host once: 12 stresses, velocities, the old iceUmask; the staging copies of the four
    // diagnostics too (cells the loop does not own are returned as they came)
    double *src[18] = {f->stressp_1, f->stressp_2, f->stressp_3, f->stressp_4, f->stressm_1, f->stressm_2, f->stressm_3, f->stressm_4,
                       f->stress12_1, f->stress12_2, f->stress12_3, f->stress12_4, f->strintxU, f->strintyU, f->taubxU, f->taubyU, f->uvel, f->vvel};
    for (int q = 0; q < 18; ++q) if (!src[q]) return fail("evp_b200_step_resident: EVP_B200_STEP_INIT_STATE needs field %d", q);
    if (!f->iceUmask) return fail("evp_b200_step_resident: EVP_B200_STEP_INIT_STATE needs iceUmask (the mask of the previous step)");
    for (int q = 0; q < 18; ++q) {
      CK(cudaMemcpyAsync(g.stage[q], src[q], bblk, cudaMemcpyHostToDevice, g.stream));
      if (q < 12) pack2_f64<<<nb, 256, 0, g.stream>>>(g.dom.sig[0][q], g.dom.sig[1][q], g.stage[q], g.d_gsrc, nd);
      else if (q == F_U) pack2_f64<<<nb, 256, 0, g.stream>>>(g.dom.u[0], g.dom.u[1], g.stage[q], g.d_gsrc, nd);
      else if (q == F_V) pack2_f64<<<nb, 256, 0, g.stream>>>(g.dom.v[0], g.dom.v[1], g.stage[q], g.d_gsrc, nd);
      else pack_f64<<<nb, 256, 0, g.stream>>>(g.dfield[q], g.stage[q], g.d_gsrc, nd);
    }
    CK(cudaMemcpyAsync(g.stage_mask2, f->iceUmask, g.nblk_elems * sizeof(int), cudaMemcpyHostToDevice, g.stream));
    pack_mask<<<nb, 256, 0, g.stream>>>(g.dmaskU, g.stage_mask2, g.d_gsrc, nd);
    CK(cudaMemsetAsync(g.d_ever_off, 0, g.nblk_elems, g.stream));
  }
  // per step: iceTmask, strength, the T-point inputs (copies on the transfer stream, packs on the compute stream as they land)
  CK(cudaEventRecord(g.ev_field[29], g.stream));
  CK(cudaStreamWaitEvent(g.xfer, g.ev_field[29], 0));   // the staging buffers are free again
  CK(cudaMemcpyAsync(g.stage_mask, pr->iceTmask, g.nblk_elems * sizeof(int), cudaMemcpyHostToDevice, g.xfer));
  CK(cudaEventRecord(g.ev_field[30], g.xfer));
  const double *tin[11] = {pr->strength, pr->tmass, pr->aice_init, pr->cdn_ocn, pr->uocn, pr->vocn, pr->ss_tltx, pr->ss_tlty, pr->strairxT, pr->strairyT, pr->TbU};
  for (int q = 0; q < 11; ++q) {
    if (!tin[q]) continue;
    CK(cudaMemcpyAsync(g.stage[18 + q], tin[q], bblk, cudaMemcpyHostToDevice, g.xfer));
    CK(cudaEventRecord(g.ev_field[18 + q], g.xfer));
  }
  CK(cudaStreamWaitEvent(g.stream, g.ev_field[30], 0));
  pack_mask<<<nb, 256, 0, g.stream>>>(g.dmaskT, g.stage_mask, g.d_gsrc, nd);
  {  // dyn_prep2's zeroing of the stresses off the ice (ice_dyn_shared.F90:717-730), on the device
    note_off_ice<<<grid_blocks(g.nblk_elems), 256, 0, g.stream>>>(g.d_ever_off, g.stage_mask, (int)g.nblk_elems);
    SigPtrs sp;
    for (int q = 0; q < 12; ++q) { sp.p[q] = g.dom.sig[0][q]; sp.p[12 + q] = g.dom.sig[1][q]; }
    zero_stress_off_ice<<<nb, 256, 0, g.stream>>>(sp, g.dmaskT, nd);
  }
  for (int q = 0; q < 11; ++q) {
    if (!tin[q]) continue;
    CK(cudaStreamWaitEvent(g.stream, g.ev_field[18 + q], 0));
    pack_f64<<<nb, 256, 0, g.stream>>>(q == 0 ? g.dfield[F_STRENGTH] : g.prepT[q - 1], g.stage[18 + q], g.d_gsrc, nd);
  }
  PrepArgs a{};
  a.hm = g.prep_static[0]; a.tarea = g.prep_static[1]; a.uarea = g.prep_static[2]; a.fcor = g.prep_static[3]; a.umask = g.prep_umask;
  a.tmass = g.prepT[0]; a.aice = g.prepT[1]; a.cdn = g.prepT[2]; a.uocn = g.prepT[3]; a.vocn = g.prepT[4];
  a.tltx = g.prepT[5]; a.tlty = g.prepT[6]; a.sax = g.prepT[7]; a.say = g.prepT[8];
  a.TbU_in = pr->TbU ? g.prepT[9] : nullptr;
  a.cdnU = g.dfield[F_CDN]; a.aiU = g.dfield[F_AIU]; a.uocnU = g.dfield[F_UOCN]; a.vocnU = g.dfield[F_VOCN];
  a.waterx = g.dfield[F_WATERX]; a.watery = g.dfield[F_WATERY]; a.forcex = g.dfield[F_FORCEX]; a.forcey = g.dfield[F_FORCEY];
  a.umassdti = g.dfield[F_UMASSDTI]; a.fm = g.dfield[F_FM]; a.TbU = g.dfield[F_TBU];
  a.strintx = g.dfield[F_STRINTX]; a.strinty = g.dfield[F_STRINTY]; a.taubx = g.dfield[F_TAUBX]; a.tauby = g.dfield[F_TAUBY];
  a.maskU = g.dmaskU;
  a.dt = pr->dt; a.cosw = p->cosw; a.sinw = p->sinw; a.area_min = pr->dyn_area_min; a.mass_min = pr->dyn_mass_min; a.gravit = pr->gravit;
  a.coupled_tilt = pr->ssh_stress;
  CK(exact::launch_prep(g.dom, a, g.stream));
  CK(cudaGetLastError());
  // "velocities may have changed in dyn_prep2" (ice_dyn_evp.F90:735-739): the ghost cells that neighbour ranks own are refreshed once,
  // in both ping-pong copies, by the staged exchange (the on-rank wrap was written by the kernel itself)
  if (g.halo.n_dst != 0 || !g.halo.peers.empty()) {
    for (int b = 0; b < 2; ++b) {
      int hl = 0;
      if (g.halo.exchange(g_comm, g.dom.u[b], g.dom.v[b], g.stream, &hl, g_err, sizeof g_err)) return 1;
    }
  }
  g.uploaded = true;
  g.stress_resident = true;
  g.state_resident = true;
  g.stress_uploaded_this_call = init;
  if (do_subcycle(p)) return 1;
  // out: the velocities always; diagnostics, stresses and the new iceUmask on request
  int what = 4;
  if (flags & EVP_B200_STEP_FETCH_DIAG) what |= 2;
  if (flags & EVP_B200_STEP_FETCH_STATE) what |= 1;
  if ((what & 2) && (!f->strintxU || !f->strintyU || !f->taubxU || !f->taubyU)) return fail("evp_b200_step_resident: EVP_B200_STEP_FETCH_DIAG needs the four arrays");
  if (flags & EVP_B200_STEP_FETCH_STATE) {
    if (!f->iceUmask) return fail("evp_b200_step_resident: EVP_B200_STEP_FETCH_STATE needs iceUmask");
    // the mask array is intent(inout) on the interior only: start from the caller's copy
    CK(cudaMemcpyAsync(g.stage_mask2, f->iceUmask, g.nblk_elems * sizeof(int), cudaMemcpyHostToDevice, g.stream));
    unpack_mask<<<grid_blocks(g.n_int), 256, 0, g.stream>>>(g.stage_mask2, g.dmaskU, g.d_int_lin, g.d_int_dom, g.n_int);
    CK(cudaMemcpyAsync(const_cast<int32_t *>(f->iceUmask), g.stage_mask2, g.nblk_elems * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  }
  return do_download(f, what);
}

int evp_b200_run_bgrid(const evp_b200_params_t *p, evp_b200_fields_t *f) {
  if (do_upload(f, false)) return 1;
  if (do_subcycle(p)) return 1;
  return do_download(f, 7);
}

int evp_b200_halo_plan(int32_t nranks, const int32_t *rects, int32_t rank, int32_t nxg, int32_t nyg, int32_t ew, int32_t ns,
                       int32_t *n, int32_t *out, int32_t cap) {
  if (!rects || !n || (!out && cap > 0) || nranks < 1 || rank < 0 || rank >= nranks) return fail("evp_b200_halo_plan: bad arguments");
  return halo_plan_host(nranks, rects, rank, nxg, nyg, ew, ns, n, out, cap);
}
int evp_b200_p2p_plan(int32_t nranks, const int32_t *rects, int32_t rank, int32_t nxg, int32_t nyg, int32_t ew, int32_t ns,
                      int32_t *n_push, int32_t *push_out, int32_t *n_fold, int32_t *fold_out, int32_t cap) {
  if (!rects || !n_push || !n_fold || (cap > 0 && (!push_out || !fold_out)) || nranks < 1 || rank < 0 || rank >= nranks)
    return fail("evp_b200_p2p_plan: bad arguments");
  const int rc = p2p_plan_host(nranks, rects, rank, nxg, nyg, ew, ns, n_push, push_out, n_fold, fold_out, cap);
  if (rc) return fail("evp_b200_p2p_plan: the staging rows of a rank overflow");
  return 0;
}
int evp_b200_stress_fold_plan(int32_t nranks, const int32_t *rects, int32_t rank, int32_t nxg, int32_t nyg, int32_t ns, int32_t *n_seg,
                              int32_t *seg_out, int32_t *n_cell, int32_t *cell_out, int32_t cap) {
  if (!rects || !n_seg || !n_cell || (cap > 0 && (!seg_out || !cell_out)) || nranks < 1 || rank < 0 || rank >= nranks)
    return fail("evp_b200_stress_fold_plan: bad arguments");
  const int rc = stress_fold_plan_host(nranks, rects, rank, nxg, nyg, ns, n_seg, seg_out, n_cell, cell_out, cap);
  if (rc && cap > 0) return fail("evp_b200_stress_fold_plan: %d entries do not hold the lists (%d segments, %d cells)", cap, *n_seg, *n_cell);
  return 0;
}
int32_t evp_b200_dom_pitch(int32_t nx) { return dom_pitch(nx); }
int64_t evp_b200_dom_cells(int32_t nx, int32_t ny) { return (int64_t)dom_cells(nx, ny); }

int evp_b200_last_loop_ms(double *ms) { if (!ms) return fail("null"); *ms = g.last_ms; return 0; }
int evp_b200_last_launches(int64_t *n) { if (!n) return fail("null"); *n = g.last_launches; return 0; }
int evp_b200_stream(void **s) { if (!s) return fail("null"); *s = (void *)g.stream; return 0; }

}  // extern "C"
