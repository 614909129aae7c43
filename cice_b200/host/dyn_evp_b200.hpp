// dyn_evp_b200.hpp -- C++ host side above the C ABI, mirroring the reference's seam for this path.
//
// The reference is Fortran and no Fortran compiler exists in this image, so the tested host layer is
// this C++ mirror of the three routines the reference dispatches for its own 1-D solver
// (cicecore/cicedyn/dynamics/ice_dyn_evp1d.F90:25):
//
//     dyn_evp1d_init      -> cice_b200::dyn_evp_b200_init      (ice_dyn_evp.F90:153-155)
//     dyn_evp1d_run       -> cice_b200::dyn_evp_b200_run       (ice_dyn_evp.F90:848-856; same 31 arrays, same order)
//     dyn_evp1d_finalize  -> cice_b200::dyn_evp_b200_finalize
//
// Arrays are passed exactly as a Fortran caller holds them: pointer to the first element of
// a(nx_block,ny_block,max_blocks), column major.  Masks are Fortran default logicals (4 bytes, any
// non-zero bit pattern is .true.: gfortran stores 1, ifort -1).  Errors follow the reference's
// convention -- abort_ice(message) (comm/mpi/ice_exit.F90) -- as a thrown AbortIce carrying the same
// "(subname) ERROR: ..." text; nothing calls exit().  The Fortran twin is fortran/ice_dyn_evp_b200.F90.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "evp_b200.h"

namespace cice_b200 {

struct AbortIce : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// what ice_blocks / ice_domain hold for this rank (ice_blocks.F90:22-41, ice_domain.F90 domain_nml)
struct BlockTable {
  int nx_block = 0, ny_block = 0, nblocks = 0, max_blocks = 0, nghost = 1;
  int nx_global = 0, ny_global = 0;
  std::string ew_boundary_type = "cyclic", ns_boundary_type = "open";
  const int *ilo = nullptr, *ihi = nullptr, *jlo = nullptr, *jhi = nullptr;  // [nblocks]
  const int *i_glob = nullptr, *j_glob = nullptr;                          // [nx_block*nblocks], [ny_block*nblocks]
};

// module variables of ice_grid / ice_dyn_shared the 2-D kernels read
struct Geometry {
  const double *dxT, *dyT, *dxhy, *dyhx, *cxp, *cyp, *cxm, *cym, *DminTarea, *uarear;
};

// module scalars of ice_dyn_shared (ice_dyn_shared.F90:66-89) + rhow from icepack
struct EvpScalars {
  int ndte = 120;
  double arlx1i = 0, denom1 = 0, revp = 0, brlx = 0, e_factor = 0, epp2i = 0, capping = 1, Ktens = 0;
  double u0 = 5e-5, cosw = 1, sinw = 0, rhow = 1026, deltaminEVP = 1e-11;
  int mode = EVP_B200_MODE_EXACT, kernel = EVP_B200_KERNEL_AUTO;
  // EVP_B200_KEEP_STRESS / EVP_B200_FETCH_STRESS: leave the stresses on the device between steps;
  // the driver sets FETCH on steps that write a restart or history file (ice_restart_driver.F90:150-231)
  int resident_flags = 0;
};

// optional: before init, for more than one rank (the host broadcasts the id itself, e.g. MPI_Bcast)
std::vector<char> dyn_evp_b200_unique_id();
void dyn_evp_b200_comm_init(int my_task, int nprocs, const std::vector<char> &id, int device);

void dyn_evp_b200_init(const BlockTable &blocks, const Geometry &geom);

void dyn_evp_b200_run(double *stressp_1, double *stressp_2, double *stressp_3, double *stressp_4,
                      double *stressm_1, double *stressm_2, double *stressm_3, double *stressm_4,
                      double *stress12_1, double *stress12_2, double *stress12_3, double *stress12_4,
                      const double *strength,
                      const double *cdn_ocnU, const double *aiU, const double *uocnU, const double *vocnU,
                      const double *waterxU, const double *wateryU, const double *forcexU, const double *forceyU,
                      const double *umassdti, const double *fmU, double *strintxU, double *strintyU,
                      const double *TbU, double *taubxU, double *taubyU, double *uvel,
                      double *vvel, const int *iceTmask, const int *iceUmask, const EvpScalars &s);

void dyn_evp_b200_finalize();

}  // namespace cice_b200
