/*
 * evp_b200.h -- C ABI of the B200-native EVP subcycling path.
 *
 * This is the drop-in boundary for ONE path of CICE: the elastic-viscous-plastic
 * subcycling loop of `evp(dt)` on the B grid,
 *
 *     do ksub = 1,ndte:  stress ; stepu ; halo(uvel,vvel)
 *     (cicecore/cicedyn/dynamics/ice_dyn_evp.F90:859-913)
 *
 * Every entry point below is what a Fortran ISO_C_BINDING interface (see
 * fortran/ice_dyn_evp_b200.F90 and INTEGRATION.md) binds.  The seam is the one the
 * reference already has for its own 1-D solver:
 *
 *     evp_b200_init        <->  dyn_evp1d_init      ice_dyn_evp1d.F90:73-115, called from ice_dyn_evp.F90:153-155
 *     evp_b200_run_bgrid   <->  dyn_evp1d_run       ice_dyn_evp1d.F90:119-310, called from ice_dyn_evp.F90:848-856
 *     evp_b200_finalize    <->  dyn_evp1d_finalize  ice_dyn_evp1d.F90:314-330
 *
 * Conventions
 *  - plain C, no C++ or torch types; pointers address the FIRST element of a Fortran
 *    array `real(dbl_kind) :: a(nx_block, ny_block, max_blocks)` (column major, i fastest)
 *    (cicecore/cicedyn/infrastructure/ice_blocks.F90:48-49,169-170).
 *  - Fortran `logical(log_kind)` is not C-interoperable; masks cross as int32 0/1.
 *  - all indices in evp_b200_grid_t are Fortran 1-based, exactly as in `type(block)`
 *    (ice_blocks.F90:22-41).
 *  - every function returns 0 on success and non-zero on failure and never calls exit();
 *    the message is available from evp_b200_last_error().  The Fortran shim turns
 *    non-zero into `abort_ice(...)` (comm/mpi/ice_exit.F90).
 *  - the library is entered single-threaded, once per MPI rank per dynamics step
 *    (the reference opens its OpenMP regions inside evp, ice_dyn_evp.F90:861);
 *    one rank <-> one GPU.  Not re-entrant.
 *  - the host owns all host arrays; the library owns device memory and keeps no host
 *    pointer past the return of a call.
 */
#ifndef EVP_B200_H
#define EVP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVP_B200_ABI_VERSION 3

/* boundary types: domain_nml ew_boundary_type / ns_boundary_type
 * (cicecore/cicedyn/infrastructure/ice_domain.F90:186-240) */
enum {
  EVP_B200_BNDY_OPEN    = 0,
  EVP_B200_BNDY_CLOSED  = 1,
  EVP_B200_BNDY_CYCLIC  = 2,
  EVP_B200_BNDY_TRIPOLE = 3   /* u-fold tripole, ns only */
};

/* arithmetic mode of the device kernels */
enum {
  EVP_B200_MODE_EXACT = 0,  /* no FMA contraction, source evaluation order: bit-identical to the
                               -ffp-contract=off CPU oracle */
  EVP_B200_MODE_FAST  = 1   /* FMA contraction allowed (nvcc default); within 1e-10 relative */
};

/* kernel strategy (0 lets the library choose from the sub-domain size) */
enum {
  EVP_B200_KERNEL_AUTO           = 0,
  EVP_B200_KERNEL_SPLIT          = 1,  /* stress kernel + stepu kernel per subcycle, the reference's two sweeps (first correct
                                          path; on the C grid: the five-kernel form cut at the reference's halo points) */
  EVP_B200_KERNEL_FUSED          = 2,  /* one fused stress+stepu kernel per subcycle, CUDA-graphed; the form follows the
                                          sub-domain size (L2-resident or HBM-streaming) */
  EVP_B200_KERNEL_PERSISTENT     = 3,  /* all ndte subcycles in one cooperative launch, state on chip */
  EVP_B200_KERNEL_FUSED_STREAM   = 4,  /* FUSED, HBM-streaming form forced (operands requested early, cp.async staging) */
  EVP_B200_KERNEL_FUSED_RESIDENT = 5,  /* FUSED, L2-resident form forced (interleaved division / square-root chains) */
  EVP_B200_KERNEL_TSTREAM        = 6   /* one launch per subcycle by persistent CTAs that walk column strips of the sub-domain; every
                                          operand staged global -> shared by TMA box loads (cp.async.bulk.tensor) one block ahead of
                                          the arithmetic.  AUTO's choice for sub-domains that stream from HBM once
                                          evp_b200_set_metric has accepted the metric arrays (single rank) */
};

/*
 * Static description of this rank's part of the domain; passed once.
 * Mirrors what dyn_evp1d_init gathers (ice_dyn_evp1d.F90:73-115) plus the block table
 * of ice_blocks.F90:22-41 and the static arrays the 2-D `stress`/`stepu` read
 * (ice_dyn_evp.F90:1457-1500, ice_dyn_shared.F90:847-890).
 */
typedef struct {
  int32_t abi_version;        /* EVP_B200_ABI_VERSION */
  int32_t nx_block, ny_block; /* block_size + 2*nghost                         ice_blocks.F90:169-170 */
  int32_t nblocks;            /* blocks owned by this rank                     ice_domain.F90 nblocks */
  int32_t max_blocks;         /* third extent of every field array             ice_domain_size.F90    */
  int32_t nghost;             /* must be 1                                     ice_blocks.F90:48      */
  int32_t nx_global, ny_global;
  int32_t ew_boundary_type;   /* EVP_B200_BNDY_*                                                      */
  int32_t ns_boundary_type;

  /* per local block n = 0..nblocks-1 (Fortran iblk = n+1) */
  const int32_t *ilo, *ihi, *jlo, *jhi; /* [nblocks]  first/last interior index, 1-based               */
  const int32_t *i_glob;      /* [nx_block*nblocks] global i of local column i; 0 = outside a closed edge */
  const int32_t *j_glob;      /* [ny_block*nblocks] global j of local row j                             */

  /* static geometry, each (nx_block,ny_block,max_blocks) f64 */
  const double *dxT, *dyT;    /* ice_grid.F90:3086-3131, 3197-3241 */
  const double *dxhy, *dyhx;  /* ice_dyn_shared.F90:401-424        */
  const double *cxp, *cyp, *cxm, *cym; /* ice_dyn_shared.F90:426-441 */
  const double *DminTarea;    /* ice_dyn_shared.F90:384-388        */
  const double *uarear;       /* ice_grid.F90:681-715              */
} evp_b200_grid_t;

/*
 * Scalars of one dynamics step, passed by value in the struct.
 * ice_dyn_shared.F90:66-89 (declarations), :453-486 (set_evp_parameters).
 */
typedef struct {
  int32_t ndte;               /* number of subcycles                         */
  int32_t mode;               /* EVP_B200_MODE_*                             */
  int32_t kernel;             /* EVP_B200_KERNEL_*                           */
  int32_t visc_method;        /* C grid only: EVP_B200_VISC_* (ice_init.F90:453) */
  double arlx1i, denom1, revp, brlx;
  double e_factor, epp2i, capping, Ktens;
  double u0, cosw, sinw;      /* ice_dyn_shared.F90:66-70                    */
  double rhow;                /* icepack_query_parameters(rhow_out=...)      */
  double deltaminEVP;         /* C grid, avg_strength only (ice_dyn_evp.F90:1960) */
} evp_b200_params_t;

/* C grid: how viscosities reach the U point (dynamics_nml visc_method) */
enum {
  EVP_B200_VISC_AVG_ZETA     = 0,  /* Bouillon et al. 2013 / C1 of Kimmritz et al. 2016 (default) */
  EVP_B200_VISC_AVG_STRENGTH = 1   /* C2 of Kimmritz et al. 2016 */
};

/*
 * Extra static geometry of grid_ice = 'C' (ice_grid.F90 dx*, dy*, areas, masks;
 * ice_dyn_evp.F90:218-240 ratio arrays).  Passed once with evp_b200_init_cgrid, after evp_b200_init.
 * Each (nx_block,ny_block,max_blocks) f64.
 */
typedef struct {
  const double *dxN, *dyE, *dxE, *dyN, *dxU, *dyU;
  const double *tarea, *uarea, *earea, *narea, *earear, *narear;
  const double *ratiodxN, *ratiodxNr, *ratiodyE, *ratiodyEr;
  const double *hm, *uvm, *epm, *npm;   /* 0/1 land masks as the reference holds them (real) */
} evp_b200_cgrid_t;

/*
 * Time-varying fields of one C-grid call: what the `grid_ice == "C"` branch of the subcycle loop
 * reads and writes (ice_dyn_evp.F90:936-1101).  All (nx_block,ny_block,max_blocks).
 */
typedef struct {
  /* inout: prognostic velocities at E and N points, their interpolants, the carried stresses */
  double *uvelE, *vvelE, *uvelN, *vvelN, *uvel, *vvel;
  double *stresspT, *stressmT, *stress12T, *stress12U;
  /* out: work arrays the reference leaves behind after the last subcycle */
  double *zetax2T, *etax2T, *etax2U, *strengthU;
  double *divergU, *tensionU, *shearU, *deltaU;
  double *strintxE, *strintyN, *taubxE, *taubyN;
  /* in */
  const double *strength;
  const double *cdn_ocnE, *cdn_ocnN, *aiE, *aiN, *uocnE, *vocnE, *uocnN, *vocnN;
  const double *waterxE, *wateryN, *forcexE, *forceyN, *emassdti, *nmassdti, *fmE, *fmN;
  const double *TbE, *TbN, *rheofactE, *rheofactN;
  /* in: 0/1 */
  const int32_t *iceTmask, *iceUmask, *iceEmask, *iceNmask;
} evp_b200_cfields_t;

/*
 * Time-varying fields of one call, the argument list of dyn_evp1d_run
 * (ice_dyn_evp1d.F90:119-153) in the same order.  All (nx_block,ny_block,max_blocks).
 */
typedef struct {
  /* inout: carried dynamics state (ice_restart_driver.F90:150-231) */
  double *stressp_1, *stressp_2, *stressp_3, *stressp_4;
  double *stressm_1, *stressm_2, *stressm_3, *stressm_4;
  double *stress12_1, *stress12_2, *stress12_3, *stress12_4;
  /* in */
  const double *strength;
  const double *cdn_ocnU, *aiU, *uocnU, *vocnU;
  const double *waterxU, *wateryU, *forcexU, *forceyU;
  const double *umassdti, *fmU;
  /* inout: written only where iceUmask is true (ice_dyn_shared.F90:925-966) */
  double *strintxU, *strintyU;
  /* in */
  const double *TbU;
  /* inout */
  double *taubxU, *taubyU;
  double *uvel, *vvel;
  /* in: 0/1 */
  const int32_t *iceTmask, *iceUmask;
} evp_b200_fields_t;

/* ---- multi-GPU bootstrap (optional; omit for one GPU) ---------------------------------
 * Replaces ice_boundary's MPI halo for the dyn fields (ice_boundary.F90:1066-1760 reached
 * through dyn_haloUpdate, ice_dyn_shared.F90:2518-2574).  Rank 0 obtains a 128-byte id,
 * the host broadcasts it with its own transport (MPI_Bcast in the Fortran shim,
 * torch.distributed in the Python tests), then every rank calls evp_b200_comm_init
 * BEFORE evp_b200_init. */
#define EVP_B200_UNIQUE_ID_BYTES 128
int evp_b200_get_unique_id(void *id128);
int evp_b200_comm_init(int32_t rank, int32_t nranks, const void *id128);

/* ---- life cycle ------------------------------------------------------------------------ */
int evp_b200_set_device(int32_t device_ordinal);           /* default: current device */
int evp_b200_init(const evp_b200_grid_t *grid);
int evp_b200_finalize(void);
const char *evp_b200_last_error(void);
/* Land-block elimination (ice_domain.F90: blocks without an ocean cell are not distributed): the blocks of a rank may
 * leave holes in the rectangle they span -- accepted as is; hole cells are land -- and several ranks' rectangles need
 * not cover the domain.  The one case that cannot be told from a forgotten evp_b200_comm_init is a SINGLE rank whose
 * blocks span less than the global domain; the caller, who knows that blocks were eliminated
 * (nblocks_tot vs. the distributed count), states it with this call before evp_b200_init. */
int evp_b200_allow_partial_domain(int32_t yes);

/* ---- the hot path ------------------------------------------------------------------------
 * One call = the whole `do ksub=1,ndte` loop of ice_dyn_evp.F90:859-913 on this rank's
 * blocks, including the per-subcycle (uvel,vvel) halo update.  Host buffers in, host
 * buffers out: H2D of the fields, the device loop, D2H of the inout fields. */
int evp_b200_run_bgrid(const evp_b200_params_t *params, evp_b200_fields_t *fields);

/* C grid (configs[2]): one call = the whole `do ksub=1,ndte` loop of ice_dyn_evp.F90:938-1097
 * (strain_rates_U, stressC_T, stressC_U, div_stress_Ex/Ny, stepu_C/stepv_C, the four velocity
 * re-interpolations and the seven halo points), one GPU.  evp_b200_init_cgrid must have been called. */
int evp_b200_init_cgrid(const evp_b200_cgrid_t *cgrid);
int evp_b200_run_cgrid(const evp_b200_params_t *params, evp_b200_cfields_t *fields);

/* ---- CD grid (grid_ice = 'CD', ice_dyn_evp.F90:1123-1293) --------------------------------------------
 * Both velocity components are prognostic at E and at N (stepuv_CD, ice_dyn_shared.F90:973-1085), the full stress tensor is
 * carried at T (stressCD_T, ice_dyn_evp.F90:1978-2080) and at U (stressCD_U, :2088-2178), the stress divergence has four parts
 * (div_stress_Ex/Ey/Nx/Ny, :2195-2416).  Same static geometry as the C grid (evp_b200_init_cgrid), one GPU, non-tripole.
 * All arrays (nx_block,ny_block,max_blocks). */
typedef struct {
  /* inout */
  double *uvelE, *vvelE, *uvelN, *vvelN, *uvel, *vvel;
  double *stresspT, *stressmT, *stress12T, *stresspU, *stressmU, *stress12U;
  /* out: work arrays the reference leaves behind after the last subcycle */
  double *zetax2T, *etax2T, *zetax2U, *etax2U, *strengthU;
  double *divergU, *tensionU, *shearU, *deltaU;
  double *strintxE, *strintyE, *strintxN, *strintyN, *taubxE, *taubyE, *taubxN, *taubyN;
  /* in */
  const double *strength;
  const double *cdn_ocnE, *cdn_ocnN, *aiE, *aiN, *uocnE, *vocnE, *uocnN, *vocnN;
  const double *waterxE, *wateryE, *waterxN, *wateryN, *forcexE, *forceyE, *forcexN, *forceyN;
  const double *emassdti, *nmassdti, *fmE, *fmN, *TbE, *TbN, *rheofactE, *rheofactN;
  /* in: 0/1 */
  const int32_t *iceTmask, *iceUmask, *iceEmask, *iceNmask;
} evp_b200_cdfields_t;
int evp_b200_run_cdgrid(const evp_b200_params_t *params, evp_b200_cdfields_t *fields);

/* ---- next row: deformations ------------------------------------------------
 * `deformations` (ice_dyn_shared.F90:1756-1860), the step right after the subcycle loop in evp()
 * (ice_dyn_evp.F90:920-934): divu, shear, vort, rdg_conv, rdg_shear at the ice T cells from the final
 * (uvel,vvel), which are still resident on the device after evp_b200_run_bgrid / evp_b200_subcycle.
 * The five arrays are inout: cells off the T list keep the caller's values (evp() zeroes them at :383-395). */
typedef struct {
  const double *dxU, *dyU, *tarear;                       /* static geometry, (nx_block,ny_block,max_blocks) */
  double *divu, *shear, *vort, *rdg_conv, *rdg_shear;     /* inout */
  double e_factor;                                        /* ice_dyn_shared.F90:465 */
} evp_b200_deform_t;
int evp_b200_deformations(evp_b200_deform_t *d);

/* ---- next row (second half): dyn_finish ------------------------------------------
 * `dyn_finish` (ice_dyn_shared.F90:1291-1365), called by evp() after the loop (ice_dyn_evp.F90:1392-1405): the ice-ocean stress
 * strocnxU, strocnyU at the ice U points from the final (uvel,vvel) and from cdn_ocnU, uocnU, vocnU, aiU, fmU, iceUmask of the
 * last upload -- all still resident on the device, so only the two results cross the boundary.
 * The two arrays are inout: points off the U list keep the caller's values, as in the reference. */
typedef struct {
  double *strocnxU, *strocnyU;   /* inout, (nx_block,ny_block,max_blocks) */
  double rhow, cosw, sinw;       /* icepack rhow; ice_dyn_shared.F90:66-70 */
} evp_b200_finish_t;
int evp_b200_dyn_finish(evp_b200_finish_t *f);

/* ---- optional: page-locking the caller's arrays -------------------------------------------------------
 * Fortran allocatables are pageable memory: copies from and to them are staged by the driver, run at a fraction of the PCIe
 * rate and block the calling thread.  A host that passes the same arrays every step (CICE does: module variables of
 * ice_dyn_evp / ice_flux / ice_state) can page-lock them once; evp_b200_run_bgrid then copies at the rate bench.py's `e2e`
 * line reports (which is measured from pinned memory).  Thin wrappers over cudaHostRegister / cudaHostUnregister; an array
 * that is already page-locked is accepted.  Unpin before the array is deallocated. */
int evp_b200_pin_host(void *ptr, size_t bytes);
int evp_b200_unpin_host(void *ptr);

/* ---- optional: metric arrays for the derived-geometry kernels ---------------------------------------------
 * dxhy, dyhx, cxp, cyp, cxm, cym and DminTarea of evp_b200_grid_t are functions of HTN, HTE, dxT, dyT and deltaminEVP
 * (ice_dyn_shared.F90:384-388, 401-441; the reference's own 1-D solver recomputes them from HTE, HTN every subcycle,
 * ice_dyn_core1d.F90:191-199).  Given HTN and HTE (ice_grid.F90, (nx_block,ny_block,max_blocks)), the library checks ON THE
 * DEVICE that the reference's expressions reproduce the seven arrays bit for bit on every T cell the loop can touch; only then
 * do the kernels of sub-domains that stream from HBM read two arrays instead of seven (360 instead of 400 B per cell and
 * subcycle; measured 164 vs 170 ms per step on 3600x2400).  *mismatches (may be NULL) receives the number of cells that differ;
 * a non-zero count is not an error, the arrays simply stay in use.  Call after evp_b200_init. */
int evp_b200_set_metric(const double *HTN, const double *HTE, double deltaminEVP, int32_t *mismatches);

/* The same call split in three so that a caller which keeps dynamics state on the device
 *  -- and bench.py's device-resident timing -- can run the loop alone.
 * run_bgrid == upload + subcycle + download. */
int evp_b200_upload(const evp_b200_fields_t *fields);
int evp_b200_subcycle(const evp_b200_params_t *params);
int evp_b200_download(evp_b200_fields_t *fields);

/* ---- next row: stresses resident on the device across dynamics steps ----------------
 * The 12 stress arrays and the velocities are the whole carried dynamics state (ice_restart_driver.F90:150-231).
 * Between two calls of evp() the reference touches the stresses in exactly one place: dyn_prep2 zeroes them where
 * iceTmask is false (ice_dyn_shared.F90:717-730).  With EVP_B200_KEEP_STRESS the library keeps them on the device,
 * applies that zeroing itself from the iceTmask of the call, and neither reads nor writes the host's stress arrays
 * (24 of the 48 field transfers of a step).  The first call after evp_b200_init uploads them regardless.
 * EVP_B200_FETCH_STRESS additionally copies them back in this call (restart / history steps);
 * evp_b200_download_stress does the same outside a step.
 * Tripole grids: after the loop the reference forces the symmetry of the stress tensor across the fold -- twelve
 * ice_HaloUpdate_stress calls (ice_dyn_evp.F90:1321-1388; ice_boundary.F90:7440-7825) that copy the mirrored top physical row
 * of stressX_3 into the north ghost row of stressX_1 and so on.  With the stresses on the device the library does that itself
 * (evp_b200_stress_symmetrise, called by evp_b200_run_bgrid_resident after the loop): a rank that holds the whole top row of the
 * grid mirrors it locally, ranks that share the top row first swap their twelve row segments (NCCL send/recv, once per step). */
enum {
  EVP_B200_KEEP_STRESS  = 1,
  EVP_B200_FETCH_STRESS = 2
};
int evp_b200_run_bgrid_resident(const evp_b200_params_t *params, evp_b200_fields_t *fields, int32_t flags);
int evp_b200_download_stress(evp_b200_fields_t *fields);
/* the symmetrisation alone, on the stresses the device holds (for callers of the upload / subcycle / download split); no-op on
 * grids without a tripole fold and on ranks below the top row.  Collective over the ranks of the top row when they are several. */
int evp_b200_stress_symmetrise(void);

/* ---- next rows: the step preparation on the device, the whole dynamics state resident ----------
 * Replaces, between the halo updates of the T-point inputs and the subcycle loop of evp() (ice_dyn_evp.F90:428-560):
 *   grid_average_X2Y('S', tmass|aice_init|cdn_ocn|uocn|vocn|ss_tltx|ss_tlty, 'T', .., 'U')   ice_grid.F90:4159-4211
 *   grid_average_X2Y('F', strairxT|strairyT, 'T', .., 'U')                                       ice_grid.F90:4620-4660
 *   dyn_prep2                                                                                    ice_dyn_shared.F90:593-839
 * so that a step uploads NINE T-point arrays + strength + iceTmask instead of the eleven U-point inputs, the velocities, both
 * masks and (first form) the stresses, and downloads the velocities only (diagnostics on request).  Carried on the device
 * between steps: velocities, stresses, and iceUmask -- dyn_prep2 needs the OLD mask to find new ice points (:765-783).
 * The velocity halo update that follows dyn_prep2 (ice_dyn_evp.F90:735-739) is the on-rank wrap plus, between ranks, one staged
 * exchange per step.  Not for tripole grids in this version (the averages and that halo update at the fold are not built; the stresses
 * alone can stay resident there, EVP_B200_KEEP_STRESS).  Ice strength stays with the caller (Icepack). */
typedef struct {
  /* static, arrays (nx_block, ny_block, max_blocks) like everything else */
  const double *hm;      /* T land mask as 0/1 real (ice_grid.F90: hm)   */
  const double *tarea;   /* T-cell area                                  */
  const double *uarea;   /* U-cell area                                  */
  const double *fcor;    /* Coriolis parameter at U points (fcor_blk)    */
  const int32_t *umask;  /* U-point land mask, Fortran logical           */
} evp_b200_prep_static_t;
int evp_b200_prep_init(const evp_b200_prep_static_t *st);

typedef struct {
  /* per step, T points, ghost cells filled (the caller's ice_HaloUpdate calls of ice_dyn_evp.F90:419-426, 471-474 stay) */
  const double *tmass, *aice_init, *cdn_ocn, *uocn, *vocn;
  const double *ss_tltx, *ss_tlty;     /* read only when ssh_stress is coupled; may be NULL otherwise */
  const double *strairxT, *strairyT;
  const double *strength;              /* after its halo update (ice_dyn_evp.F90:731-733) */
  const int32_t *iceTmask;             /* after its halo update (:415-418) */
  const double *TbU;                   /* seabed stress factor at U points, or NULL (seabed_stress = .false.) */
  double dt, dyn_area_min, dyn_mass_min, gravit;
  int32_t ssh_stress;                  /* 0 geostrophic, 1 coupled */
} evp_b200_prep_t;
enum {
  EVP_B200_STEP_INIT_STATE = 1,   /* take uvel, vvel, iceUmask and the 12 stresses from `fields` first (first step, after a restart read) */
  EVP_B200_STEP_FETCH_DIAG = 2,   /* copy strintxU, strintyU, taubxU, taubyU back as well (history steps) */
  EVP_B200_STEP_FETCH_STATE = 4   /* copy the 12 stresses and iceUmask back as well (restart steps) */
};
/* one dynamics step: preparation + the whole subcycle loop; fields->uvel, vvel are written, the rest of `fields` as the flags say */
int evp_b200_step_resident(const evp_b200_params_t *params, const evp_b200_prep_t *prep, evp_b200_fields_t *fields, int32_t flags);

/* ---- measurement hooks (not part of the reference seam) ---------------------------------- */
/* device time of the most recent evp_b200_subcycle in ms (CUDA events on the library's stream) */
int evp_b200_last_loop_ms(double *ms);
/* number of kernel launches issued by the most recent evp_b200_subcycle */
int evp_b200_last_launches(int64_t *n);
/* raw CUDA stream handle (cudaStream_t) the library launches on, for external event timing */
int evp_b200_stream(void **stream);
/* report the chosen tiling/kernel as a short static string */
const char *evp_b200_describe(void);


/* ---- host-only planning hook (no GPU needed; used by the CPU tests of the multi-rank logic) ----
 * The (uvel,vvel) halo plan of `rank` when rank r owns the rectangle rects[4r..4r+3] =
 * {gi0, gj0, nx, ny} (global index of its first interior cell, interior extent).  Writes up to `cap`
 * entries of 6 ints {dst, src1_rank, src1, src2_rank, src2, op} into `out` (dst/src are indices into
 * the owning rank's (nx+2)-wide, pitch-padded sub-domain array: see evp_b200_dom_pitch; op 0 copy,
 * 1 negate, 2 0.5*(src1 - src2), 3 -(0.5*(src1 - src2))) and the total count into *n.  Ghost cells the compute kernels fill
 * themselves (on-rank cyclic wrap) are not listed. */
int evp_b200_halo_plan(int32_t nranks, const int32_t *rects, int32_t rank, int32_t nx_global, int32_t ny_global,
                       int32_t ew_boundary_type, int32_t ns_boundary_type, int32_t *n, int32_t *out, int32_t cap);
int32_t evp_b200_dom_pitch(int32_t nx);
/* cells of one sub-domain array: pitch * (ny + 2 ghost rows + 2 staging rows for raw values that cross a tripole fold) */
int64_t evp_b200_dom_cells(int32_t nx, int32_t ny);
/* The same halo as the in-kernel NVLink form serves it (no staged exchange): what `rank` stores into OTHER ranks' arrays while it
 * advances a subcycle -- up to `cap` entries of 4 ints {src cell of rank, destination rank, destination cell, negate 0/1} in
 * push_out; destination cells include the staging rows ny+2, ny+3 -- and what it combines itself once its peers' stores have
 * arrived -- entries {dst, src1, src2, op} (all cells of its own array, op as in evp_b200_halo_plan) in fold_out.
 * Replaces ice_HaloUpdate's message build for the dyn fields (ice_boundary.F90:7910-9050) incl. the tripole buffers (:8079-8157). */
int evp_b200_p2p_plan(int32_t nranks, const int32_t *rects, int32_t rank, int32_t nx_global, int32_t ny_global,
                      int32_t ew_boundary_type, int32_t ns_boundary_type, int32_t *n_push, int32_t *push_out, int32_t *n_fold,
                      int32_t *fold_out, int32_t cap);

/* The stress symmetrisation across a tripole fold between ranks (evp_b200_stress_symmetrise), as a plan: the other ranks of the top row
 * `rank` swaps its top-row segment with -- entries {rank, gi0, nx} in seg_out -- and, for every cell of its north ghost row (column
 * 0 .. nx+1), the rank and the column (1 .. nx of that rank) of the top physical row that is mirrored into it -- entries {dst column,
 * source rank or -1 where no rank holds the column, source column} in cell_out.  What ice_HaloUpdate_stress derives from its tripole
 * buffer addresses (ice_boundary.F90:8117-8157, field_loc_center).  Empty below the top row and without a fold.  cap = 0: counts only. */
int evp_b200_stress_fold_plan(int32_t nranks, const int32_t *rects, int32_t rank, int32_t nx_global, int32_t ny_global,
                              int32_t ns_boundary_type, int32_t *n_seg, int32_t *seg_out, int32_t *n_cell, int32_t *cell_out, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* EVP_B200_H */
