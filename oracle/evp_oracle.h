/*
 * evp_oracle.h -- CPU oracle for the EVP subcycling path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker or the timed CPU baseline.  The product
 * (cice_b200/) never links, imports or calls it.
 *
 * PINNING: the reference (Fortran) cannot be compiled in this image (no f951 / MPI / csh) and ships no
 * golden vectors for this path.  The restatement is pinned
 *   (1) by vectors generated FROM THE REFERENCE'S SOURCE TEXT: tests/golden/ref_translit.py reads the Fortran
 *       subroutines under /root/reference, transliterates them statement by statement into Python (same operators,
 *       operand order and parentheses; IEEE doubles, no contraction) and executes them; this oracle reproduces those
 *       vectors bit for bit (tests/golden/ref_source_vectors.{npz,json}; B grid: stress, stepu, strain_rates,
 *       visc_replpress; C grid: strain_rates_U/Tdt, stressC_T/U, div_stress_Ex/Ny, stepu_C, stepv_C, grid_average_X2Y*;
 *       deformations; ice_constants).  Not covered that way: the MPI halo (ice_boundary.F90) and the 1-D solver;
 *   (2) by the properties the reference itself asserts (bit-for-bit under any block decomposition and with eliminated
 *       land blocks, 2-D == 1-D formulation, halochk closed-form halo values).
 * It is still not pinned by outputs of a COMPILED reference binary: compiler-specific reassociation or FMA contraction
 * in a real Fortran build is outside what a source-level transliteration can show.
 *
 * The entry points take the SAME structs as the product's C ABI (include/evp_b200.h) so a
 * test feeds both sides from one set of buffers.
 */
#ifndef EVP_ORACLE_H
#define EVP_ORACLE_H

#include "evp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* 2-D blocked path: ice_dyn_evp.F90:859-913 (loop), :1457-1743 (stress),
 * ice_dyn_shared.F90:847-968 (stepu), halo via the halochk semantics.
 * `grid` must describe ALL blocks of the global domain (serial-comm view).
 * nthreads <= 0: use omp_get_max_threads(). Returns 0 on success. */
int orc_evp_run_bgrid(const evp_b200_grid_t *grid, const evp_b200_params_t *params,
                      evp_b200_fields_t *fields, int nthreads);

/* 1-D gather-indexed path: ice_dyn_evp1d.F90:119-310 + ice_dyn_core1d.F90.
 * Needs HTE/HTN instead of the precomputed cxp.. arrays, exactly like the reference 1-D
 * solver (ice_dyn_core1d.F90:191-199).  Single block covering the whole domain only
 * (the reference gathers to one global array first, ice_dyn_evp1d.F90:178-195). */
int orc_evp_run_bgrid_1d(const evp_b200_grid_t *grid, const double *HTE, const double *HTN,
                         double deltaminEVP, const evp_b200_params_t *params,
                         evp_b200_fields_t *fields, int nthreads);

/* C grid (grid_ice='C'): ice_dyn_evp.F90:936-1101 and callees, see evp_oracle_cgrid.c.  Non-tripole only. */
int orc_evp_run_cgrid(const evp_b200_grid_t *grid, const evp_b200_cgrid_t *cgrid, const evp_b200_params_t *params,
                      evp_b200_cfields_t *fields, int nthreads);
/* grid_ice = 'CD': evp.F90:1123-1275 */
int orc_evp_run_cdgrid(const evp_b200_grid_t *grid, const evp_b200_cgrid_t *cgrid, const evp_b200_params_t *params,
                        evp_b200_cdfields_t *fields, int nthreads);

/* deformations: ice_dyn_shared.F90:1756-1860 over the T list (ilo:ihi+1, jlo:jhi+1 where iceTmask) */
int orc_deformations(const evp_b200_grid_t *grid, const int32_t *iceTmask, const double *uvel, const double *vvel,
                     evp_b200_deform_t *d);
/* dyn_finish: ice_dyn_shared.F90:1291-1365 over the U list (ilo:ihi, jlo:jhi where iceUmask) */
int orc_dyn_finish(const evp_b200_grid_t *grid, const int32_t *iceUmask, const double *uvel, const double *vvel, const double *Cw,
                   const double *uocn, const double *vocn, const double *aiX, const double *fm, evp_b200_finish_t *f);

/* NE-corner / vector halo update of nfld fields, the dyn_haloUpdate call of
 * ice_dyn_evp.F90:908-910 (ice_boundary.F90:1066-1760; expectations halochk.F90:530-830).
 * field_loc: 0 center, 1 NE corner; field_type: 0 scalar, 1 vector. */
int orc_halo_update(const evp_b200_grid_t *grid, double **flds, int nfld,
                    int field_loc, int field_type);

/* Tripole grids: the twelve ice_HaloUpdate_stress calls that force symmetry of the stress tensor across the fold after
 * the loop (ice_dyn_evp.F90:1321-1388; ice_boundary.F90:7440-7825).  sig: the 12 stress arrays in the order of
 * evp_b200_fields_t.  No-op on other grids. */
int orc_stress_symmetrise(const evp_b200_grid_t *grid, double **sig);

/* single calls, for kernel-level tests (one block, arrays (nx_block,ny_block)) */
void orc_stress_block(int nx_block, int ny_block, int icellT, const int *indxTi, const int *indxTj,
                      const double *uvel, const double *vvel,
                      const double *dxT, const double *dyT, const double *dxhy, const double *dyhx,
                      const double *cxp, const double *cyp, const double *cxm, const double *cym,
                      const double *DminTarea, const double *strength,
                      double *stressp_1, double *stressp_2, double *stressp_3, double *stressp_4,
                      double *stressm_1, double *stressm_2, double *stressm_3, double *stressm_4,
                      double *stress12_1, double *stress12_2, double *stress12_3, double *stress12_4,
                      double *str, const evp_b200_params_t *p);

void orc_stepu_block(int nx_block, int ny_block, int icellU, const int *indxUi, const int *indxUj,
                     const double *Cw, const double *aiX, const double *str,
                     const double *uocn, const double *vocn, const double *waterx, const double *watery,
                     const double *forcex, const double *forcey, const double *Umassdti,
                     const double *fm, const double *uarear,
                     double *strintx, double *strinty, double *taubx, double *tauby,
                     const double *uvel_init, const double *vvel_init,
                     double *uvel, double *vvel, const double *TbU, const evp_b200_params_t *p);

const char *orc_last_error(void);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
