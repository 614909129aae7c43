// dyn_evp_b200.cpp -- see dyn_evp_b200.hpp.  Plain host C++; links against libevp_b200.so.
#include "dyn_evp_b200.hpp"

#include <cstring>

namespace cice_b200 {

static size_t g_nelem = 0;
static std::vector<int32_t> g_maskT, g_maskU;

static void check(int rc, const char *subname) {
  if (rc != 0) throw AbortIce(std::string("(") + subname + ") ERROR: " + evp_b200_last_error());
}

static int bndy_code(const std::string &s) {
  if (s == "cyclic") return EVP_B200_BNDY_CYCLIC;
  if (s == "closed") return EVP_B200_BNDY_CLOSED;
  if (s == "tripole") return EVP_B200_BNDY_TRIPOLE;
  if (s == "open") return EVP_B200_BNDY_OPEN;
  throw AbortIce("(dyn_evp_b200_init) ERROR: boundary type '" + s + "' not supported");
}

std::vector<char> dyn_evp_b200_unique_id() {
  std::vector<char> id(EVP_B200_UNIQUE_ID_BYTES);
  check(evp_b200_get_unique_id(id.data()), "dyn_evp_b200_unique_id");
  return id;
}

void dyn_evp_b200_comm_init(int my_task, int nprocs, const std::vector<char> &id, int device) {
  check(evp_b200_set_device(device), "dyn_evp_b200_comm_init");
  if (id.size() < EVP_B200_UNIQUE_ID_BYTES) throw AbortIce("(dyn_evp_b200_comm_init) ERROR: short unique id");
  check(evp_b200_comm_init(my_task, nprocs, id.data()), "dyn_evp_b200_comm_init");
}

void dyn_evp_b200_init(const BlockTable &b, const Geometry &geo) {
  evp_b200_grid_t g;
  std::memset(&g, 0, sizeof g);
  g.abi_version = EVP_B200_ABI_VERSION;
  g.nx_block = b.nx_block; g.ny_block = b.ny_block; g.nblocks = b.nblocks; g.max_blocks = b.max_blocks;
  g.nghost = b.nghost; g.nx_global = b.nx_global; g.ny_global = b.ny_global;
  g.ew_boundary_type = bndy_code(b.ew_boundary_type);
  g.ns_boundary_type = bndy_code(b.ns_boundary_type);
  g.ilo = b.ilo; g.ihi = b.ihi; g.jlo = b.jlo; g.jhi = b.jhi; g.i_glob = b.i_glob; g.j_glob = b.j_glob;
  g.dxT = geo.dxT; g.dyT = geo.dyT; g.dxhy = geo.dxhy; g.dyhx = geo.dyhx;
  g.cxp = geo.cxp; g.cyp = geo.cyp; g.cxm = geo.cxm; g.cym = geo.cym;
  g.DminTarea = geo.DminTarea; g.uarear = geo.uarear;
  check(evp_b200_init(&g), "dyn_evp_b200_init");
  g_nelem = (size_t)b.nx_block * b.ny_block * b.max_blocks;
  g_maskT.assign(g_nelem, 0);
  g_maskU.assign(g_nelem, 0);
}

void dyn_evp_b200_run(double *stressp_1, double *stressp_2, double *stressp_3, double *stressp_4,
                      double *stressm_1, double *stressm_2, double *stressm_3, double *stressm_4,
                      double *stress12_1, double *stress12_2, double *stress12_3, double *stress12_4,
                      const double *strength,
                      const double *cdn_ocnU, const double *aiU, const double *uocnU, const double *vocnU,
                      const double *waterxU, const double *wateryU, const double *forcexU, const double *forceyU,
                      const double *umassdti, const double *fmU, double *strintxU, double *strintyU,
                      const double *TbU, double *taubxU, double *taubyU, double *uvel,
                      double *vvel, const int *iceTmask, const int *iceUmask, const EvpScalars &s) {
  if (g_nelem == 0) throw AbortIce("(dyn_evp_b200_run) ERROR: dyn_evp_b200_init has not been called");
  // Fortran logical -> 0/1 (the only conversion at the boundary)
  for (size_t k = 0; k < g_nelem; ++k) {
    g_maskT[k] = iceTmask[k] != 0;
    g_maskU[k] = iceUmask[k] != 0;
  }
  evp_b200_params_t p;
  std::memset(&p, 0, sizeof p);
  p.ndte = s.ndte; p.mode = s.mode; p.kernel = s.kernel;
  p.arlx1i = s.arlx1i; p.denom1 = s.denom1; p.revp = s.revp; p.brlx = s.brlx;
  p.e_factor = s.e_factor; p.epp2i = s.epp2i; p.capping = s.capping; p.Ktens = s.Ktens;
  p.u0 = s.u0; p.cosw = s.cosw; p.sinw = s.sinw; p.rhow = s.rhow; p.deltaminEVP = s.deltaminEVP;
  evp_b200_fields_t f;
  f.stressp_1 = stressp_1; f.stressp_2 = stressp_2; f.stressp_3 = stressp_3; f.stressp_4 = stressp_4;
  f.stressm_1 = stressm_1; f.stressm_2 = stressm_2; f.stressm_3 = stressm_3; f.stressm_4 = stressm_4;
  f.stress12_1 = stress12_1; f.stress12_2 = stress12_2; f.stress12_3 = stress12_3; f.stress12_4 = stress12_4;
  f.strength = strength; f.cdn_ocnU = cdn_ocnU; f.aiU = aiU; f.uocnU = uocnU; f.vocnU = vocnU;
  f.waterxU = waterxU; f.wateryU = wateryU; f.forcexU = forcexU; f.forceyU = forceyU;
  f.umassdti = umassdti; f.fmU = fmU; f.strintxU = strintxU; f.strintyU = strintyU; f.TbU = TbU;
  f.taubxU = taubxU; f.taubyU = taubyU; f.uvel = uvel; f.vvel = vvel;
  f.iceTmask = g_maskT.data(); f.iceUmask = g_maskU.data();
  if (s.resident_flags)
    check(evp_b200_run_bgrid_resident(&p, &f, s.resident_flags), "dyn_evp_b200_run");
  else
    check(evp_b200_run_bgrid(&p, &f), "dyn_evp_b200_run");
}

void dyn_evp_b200_finalize() {
  check(evp_b200_finalize(), "dyn_evp_b200_finalize");
  g_nelem = 0;
  g_maskT.clear();
  g_maskU.clear();
}

}  // namespace cice_b200
