// evp_internal.h -- types shared by the C-ABI translation unit and the two kernel translation
// units (exact: -fmad=false, fast: default FMA contraction).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace evp {

// Device sub-domain ("dom"): this rank's rectangle of the global grid with its 1-cell ghost ring,
// every field one 2-D array, i fastest, row pitch ld doubles.  dom (i,j): 0 = W/S ghost,
// 1..nx / 1..ny interior, nx+1 / ny+1 = E/N ghost.  This is the reference's block with nghost=1
// (ice_blocks.F90:48-49) for the rank's stitched rectangle of blocks.
struct Dom {
  int nx, ny;  // interior extent
  int ld;      // row pitch (doubles)
  int nyd;     // ny + 2
  int wrap_ew; // E-W ghost columns are the rank's own opposite interior columns (cyclic, whole width local)
  int wrap_ns; // same for N-S (cyclic)
  int fold_top; // row ny is the row below a tripole fold: its off-ice U points are carried across the ping-pong copies (fused_body)

  // carried state
  double *u[2], *v[2];  // velocity ping-pong (fused kernels read [cur], write [cur^1])
  double *sig[2][12];   // stressp_1..4, stressm_1..4, stress12_1..4, same ping-pong
  // per-step inputs at T points
  const double *strength;
  // static geometry at T points
  const double *dxT, *dyT, *dxhy, *dyhx, *cxp, *cyp, *cxm, *cym, *DminTarea;
  // per-step inputs at U points
  const double *cdn, *aiu, *uocn, *vocn, *waterx, *watery, *forcex, *forcey, *umassdti, *fm, *TbU;
  const double *uarear;
  double *uinit, *vinit;
  // diagnostics written by stepu
  double *strintx, *strinty, *taubx, *tauby;
  // split-kernel temporaries: the reference's strtmp(nx_block,ny_block,8) (ice_dyn_evp.F90:328)
  double *str[8];
  // ice masks, 1 byte per cell
  const unsigned char *maskT, *maskU;
};

// C grid (grid_ice = 'C'): every array of the C-grid subcycle on the dom layout (see evp_cgrid.cu)
struct CDom {
  int nx, ny, ld, nyd, wrap_ew, wrap_ns;
  // static geometry
  const double *dxN, *dyE, *dxE, *dyN, *dxU, *dyU, *dxT, *dyT, *uarea, *DminTarea, *tarea, *hm, *earea, *narea, *earear, *narear;
  const double *ratiodxN, *ratiodxNr, *ratiodyE, *ratiodyEr, *epm, *npm, *uvm;
  const double *rhalf_dyE, *r_dxE, *rhalf_dxN, *r_dyN, *uareaavgr;  // static quotients, divided once at init (evp_cgrid.cu)
  // carried state
  double *uvelE, *vvelE, *uvelN, *vvelN, *uvel, *vvel, *stresspT, *stressmT, *stress12T, *stress12U;
  double *stress12Ub;  // second copy of stress12U for the fused three-kernel form (ping-pong, see evp_cgrid.cu)
  // work arrays the reference leaves behind
  double *zetax2T, *etax2T, *etax2U, *strengthU, *divergU, *tensionU, *shearU, *deltaU, *strintxE, *strintyN, *taubxE, *taubyN;
  // per-step inputs
  const double *strength, *cdnE, *cdnN, *aiE, *aiN, *uocnE, *vocnE, *uocnN, *vocnN, *waterxE, *wateryN, *forcexE, *forceyN;
  const double *emassdti, *nmassdti, *fmE, *fmN, *TbE, *TbN, *rheofactE, *rheofactN;
  double *uvelE_init, *vvelN_init;
  const unsigned char *maskT, *maskU, *maskE, *maskN;
  // CD grid only (grid_ice = 'CD', evp_b200_run_cdgrid): the rest of the stress tensor, the other momentum component at E and N
  double *stresspU, *stressmU, *zetax2U, *strintyE, *strintxN, *taubyE, *taubxN, *vvelE_init, *uvelN_init;
  const double *wateryE, *waterxN, *forceyE, *forcexN;
};

// scalars of set_evp_parameters (ice_dyn_shared.F90:453-486) and friends
struct KParams {
  double arlx1i, denom1, revp, brlx;
  double e_factor, epp2i, capping, Ktens;
  double u0, cosw, sinw, rhow;
  double deltaminEVP;
  int visc_method;
};

// In-kernel NVLink halo (see evp_halo.cu, fused_kernel): edge CTAs store new edge velocities straight into the
// neighbour GPUs' ghost cells (CUDA-IPC mapped peer memory) and hand over with per-peer epoch flags.
#define P2P_MAXPEER 16
#define P2P_LL_SLOTS 3   // copies of the low-latency slots, used round robin by subcycle (see P2PParams)
struct P2PParams {
  int enabled;
  int npeers;
  int n_edge_tiles;              // tiles touching the sub-domain boundary; they are scheduled first
  int ntx, nty;                  // tile grid of the fused kernel
  const int *tile_order;         // [ntx*nty] tile ids, edge tiles first
  int n_push;                    // cells of other ranks (ghost cells, fold staging slots) that are fed from this rank
  int fold_row;                  // 0, or the interior row (ny-1) below a tripole fold whose U points are push sources too
  const int *push_start;         // CSR over the edge index of a push source (see edge_index)
  const int *push_peer;          // peer slot in bits 0-7, bit 8: the value arrives negated (it crosses the tripole fold)
  const int *push_dst;           // cell index inside that peer's sub-domain array (staging rows ny+2, ny+3 included)
  double *peer_u[2][P2P_MAXPEER], *peer_v[2][P2P_MAXPEER];
  unsigned long long *peer_flag[P2P_MAXPEER];  // the peer's flags[my rank]
  const unsigned long long *my_flags;          // my flags[], indexed by peer rank
  int peer_rank[P2P_MAXPEER];
  unsigned long long *done_ctr;                // edge CTAs finished so far in this loop
  unsigned long long *epoch_base;              // device-resident: epoch of the current loop
  int *err;                                    // set when a wait times out
  unsigned long long *dbg;                     // [0] sum of wait cycles of lane 0 of edge CTAs, [1] number of waits, [2] max wait
  // Low-latency slots (persistent kernel): every ghost cell that a neighbour GPU feeds has, in each of P2P_LL_SLOTS copies used round
  // robin by subcycle (one more than the ping-pong needs: a sender is never less than two subcycles from a copy still being read,
  // whatever the tilings of the two ranks), four 8-byte words
  // (lo32(u) | tag << 32, hi32(u) | tag << 32, the same for v) in the receiver's memory.  A value validates itself -- tag = low 32
  // bits of epoch + subcycle -- so the exchange inside the loop needs no fence and no flag: one NVLink store latency.
  unsigned long long *peer_ll[P2P_MAXPEER];    // that peer's slots
  int peer_ring[P2P_MAXPEER];                  // ghost-ring cells of that peer's sub-domain (slots per parity)
  const int *push_ll;                          // per push entry: ring index of the destination ghost cell over there (ring_index)
  unsigned long long *my_ll;                   // this rank's slots
  const unsigned char *ll_fed;                 // [my_ring] 1 where a neighbour GPU feeds the ghost cell
  int my_ring;
};

// position of ghost cell (i,j) of an nx x ny sub-domain in its ring: S row, N row, W column, E column
__host__ __device__ inline int ring_cells(int nx, int ny) { return 2 * (nx + 2) + 2 * ny; }
__host__ __device__ inline int ring_index(int nx, int ny, int i, int j) {
  if (j == 0) return i;
  if (j == ny + 1) return (nx + 2) + i;
  if (i == 0) return 2 * (nx + 2) + (j - 1);
  return 2 * (nx + 2) + ny + (j - 1);
}

// Tripole fold, centre-located fields (the stress symmetrisation, ice_boundary.F90:8136-8157 with ioffset = 0): the global column whose
// top physical row is mirrored into ghost cell `i` (0 .. nx+1) of the north ghost row of a sub-domain that starts at global column gi0.
// One definition for the device kernels and for the host-side plan (evp_b200_stress_fold_plan).
__host__ __device__ inline int fold_mirror_col(int nxg, int gi0, int i) {
  int ig = gi0 + i - 1;
  if (ig < 1) ig += nxg;
  if (ig > nxg) ig -= nxg;
  return nxg - ig + 1;
}

// KERNEL_PERSISTENT (evp_persist.cu): one CTA per SM owns a tile of the sub-domain for the whole loop.  The plan is built on the
// host (evp_persist_plan.h); the tables say which T cell / U point each (slot, thread) of a CTA advances.
#define PERSIST_THREADS 512
#define PERSIST_SLOTS 2      // cells per thread and subcycle
#define PERSIST_NONE_W 0xffffffffu  // table word of an empty (slot, thread)
#define PERSIST_TL0 100              // first subcycle of the debug timeline
#define PERSIST_CTR_STRIDE 32       // progress counters sit 128 B apart: one L2 line (and slice) each
struct PersistPlan {
  int ntx, nty;   // tile grid (one CTA per tile, all co-resident)
  int bx, by;     // U points per full tile; tiles of the last column / row may be narrower
  int nT, nU;     // (bx+1)*(by+1) T cells, bx*by U points of a full tile
  int nring;      // (bx+2)*(by+2): u, v of the tile with its one-cell ring
  int ndte;
  int kT, kU;     // how many static T / U arrays the instantiation keeps in shared memory
  int use_init;   // revised EVP reads uvel_init/vvel_init
  int nthreads;   // threads per CTA the tables were built for
  int n_sig;      // multi-GPU: publishing warps of all tiles on the sub-domain edge (what the hand-over counter reaches per subcycle)
  unsigned *progress;   // [ntx*nty * PERSIST_CTR_STRIDE] edge-U warps that have published, summed over the subcycles so far
  // [4 tile shapes][PERSIST_SLOTS * nthreads] packed words (evp_persist_plan.h: persist_word); shape = (last column) + 2*(last row)
  const unsigned *tslot, *uslot;
  int ewT[4], ewU[4];   // warps (from the top) whose slot 1 holds the tile-edge T cells; warps (from 0) whose slot 0 holds the tile-edge U points
  int lw0[4], nlw[4];   // "light" warps (no slot-1 T cell): they refresh the ring for the edge warps; nlw = 0: the edge warps do it themselves
  int *err;             // set when a wait on a neighbour tile times out (never, unless CTAs are not co-resident)
  long long *dbg;       // null, or [tiles][warps][5] cycle counters (EVP_B200_PERSIST_DEBUG)
  int off_uv, off_str, off_sig, off_T, off_U;  // shared-memory offsets in doubles
  unsigned smem_bytes;
};

// KERNEL_TSTREAM (evp_tstream.cu): sub-domains that stream from HBM, advanced by persistent CTAs that walk column strips of the
// sub-domain; every operand of a block of cells arrives in shared memory as a TMA box load one block ahead of the arithmetic.
#define TS_W 32              // columns of an fp64 box: 31 T cells of a strip row and the west neighbour of the first
#define TS_STRIDE 30         // U points a strip advances per row = distance between strips (even: box loads start on 16-byte boundaries)
#define TS_MW 48             // columns of a mask box (bytes): it starts at the 16-column boundary at or below the fp64 box
#define TS_NMAPS 47          // tensor maps: u[2] v[2] sig[2][12] strength dxT dyT HTN HTE uop[12] maskT maskU
#define TS_MAP_U 0
#define TS_MAP_V 2
#define TS_MAP_SIG 4
#define TS_MAP_STRENGTH 28
#define TS_MAP_DXT 29
#define TS_MAP_DYT 30
#define TS_MAP_HTN 31
#define TS_MAP_HTE 32
#define TS_MAP_UOP 33
#define TS_MAP_MASKT 45
#define TS_MAP_MASKU 46
struct TsPlan {
  const void *maps;    // HOST copy of the TS_NMAPS tensor maps (evp_tma.cuh: TsMaps), boxes sized for `rows`; the launcher passes them
                       // to the kernel by value as a __grid_constant__ parameter (the TMA unit reads them from the constant bank)
  int rows;            // T rows per block (threads per CTA = 32 * rows)
  int nstrips;         // strips of TS_W - 1 T columns, stride TS_STRIDE
  int nseg;            // segments of a strip: rows * nb T rows each, stride rows * nb - 1
  int nb;              // blocks per (full) segment
  int nitems;          // nstrips * nseg work items, dealt round robin to the CTAs
  int ctas;            // grid size
  int *err;            // set when a wait for a box load times out (never, unless the byte count is wrong)
  double deltamin;     // deltaminEVP (derived geometry)
};

// arguments of prep_kernel (evp_kernels.cu): the step preparation on the device
struct PrepArgs {
  const double *hm, *tarea, *uarea, *fcor;      // static: T land mask as 0/1 real, T-cell area, U-cell area, Coriolis parameter at U
  const unsigned char *umask;                   // static: U-point land mask
  const double *tmass, *aice, *cdn, *uocn, *vocn, *tltx, *tlty, *sax, *say;   // T points, halos filled by the caller
  const double *TbU_in;                         // null: no seabed stress (TbU = 0)
  double *cdnU, *aiU, *uocnU, *vocnU, *waterx, *watery, *forcex, *forcey, *umassdti, *fm, *TbU, *strintx, *strinty, *taubx, *tauby;
  unsigned char *maskU;
  double dt, cosw, sinw, area_min, mass_min, gravit;
  int coupled_tilt;                             // ssh_stress: 0 geostrophic, 1 coupled
};

// launchers implemented once per arithmetic mode (namespace exact / fast)
#define EVP_DECLARE_LAUNCHERS(NS)                                                                   \
  namespace NS {                                                                                    \
  cudaError_t launch_stress(const Dom &d, const KParams &p, int cur, cudaStream_t s);               \
  cudaError_t launch_stepu(const Dom &d, const KParams &p, int cur, cudaStream_t s);                \
  cudaError_t launch_deform(const Dom &d, int cur, const double *dxU, const double *dyU, const double *tarear, double *divu, double *shear, double *vort, double *rdg_conv, double *rdg_shear, double e_factor, cudaStream_t s); \
  cudaError_t launch_prep(const Dom &d, const PrepArgs &args, cudaStream_t s); \
  cudaError_t launch_finish(const Dom &d, int cur, double *strocnx, double *strocny, double rhow, double cosw, double sinw, cudaStream_t s); \
  cudaError_t launch_fused(const Dom &d, const KParams &p, int cur, cudaStream_t s, int form, bool pdl, int last); \
  cudaError_t launch_fused_p2p(const Dom &d, const KParams &p, const P2PParams &pp, int cur, int ksub, int last, int form, cudaStream_t s); \
  cudaError_t launch_fold(const P2PParams &pp, double *U, double *V, const int *dst, const int *c1, const int *c2, const signed char *code, int n, int ksub, int pdl, cudaStream_t s); \
  int fold_max_entries(); \
  cudaError_t set_wait_timeout(unsigned long long ns); \
  cudaError_t set_wait_timeout_persist(unsigned long long ns); \
  cudaError_t persist_ctas_per_sm(const PersistPlan &pp, bool p2p, int *n); \
  cudaError_t set_metric(const double *HTN, const double *HTE, double deltamin); \
  cudaError_t launch_metric_verify(const Dom &d, const double *HTN, const double *HTE, double deltamin, int skip_e, int skip_n, int *mismatches, cudaStream_t s); \
  cudaError_t launch_persist(const Dom &d, const KParams &p, const PersistPlan &pp, const P2PParams *px, cudaStream_t s); \
  cudaError_t launch_tstream(const Dom &d, const KParams &p, const TsPlan &ts, int cur, int last, bool pdl, cudaStream_t s); \
  int tstream_plan(const Dom &d, size_t dom_rows, const double *HTN, const double *HTE, int num_sms, int rows, void *host_maps, TsPlan *ts, char *why, size_t nwhy); \
  size_t tstream_map_bytes(); \
  cudaError_t launch_cgrid_subcycle(const CDom &d, const KParams &p, cudaStream_t s, int *launches);  \
  cudaError_t launch_cgrid_static(const CDom &d, double *rhalf_dyE, double *r_dxE, double *rhalf_dxN, double *r_dyN, double *uareaavgr, cudaStream_t s); \
  cudaError_t launch_cgrid_subcycle_fused(const CDom &d, const KParams &p, int cur, cudaStream_t s, int *launches);  \
  cudaError_t launch_cdgrid_subcycle(const CDom &d, const KParams &p, cudaStream_t s, int *launches);  \
  }
EVP_DECLARE_LAUNCHERS(exact)
EVP_DECLARE_LAUNCHERS(fast)

}  // namespace evp
