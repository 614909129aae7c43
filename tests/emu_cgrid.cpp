// emu_cgrid.cpp -- the C-grid and CD-grid CUDA kernels of cice_b200/csrc/evp_cgrid.cu run THREAD BY THREAD ON THE HOST.
//
// Test infrastructure only (tests/test_emu_cgrid.py); nothing in the product links this.  The kernel translation unit is included
// unchanged with EVP_HOST_EMU defined (launchers and the cooperative single-launch kernel compiled out) on top of
// tests/cuda_emu.h and compiled with g++ -ffp-contract=off like the `exact` namespace.  One block in the reference's layout
// (nghost = 1) IS a device sub-domain (ld = nx_block), so the kernels work directly on the caller's arrays; what the library
// does around them (zero-filling the work arrays the reference zero-fills, the second stress12U copy, uvelE_init / vvelN_init,
// the static quotients) is repeated here in a few lines.
#include "cuda_emu.h"

#define EVP_HOST_EMU 1
#define EVP_NS exact
#include "evp_b200.h"
#include "evp_cgrid.cu"

using namespace evp;
using namespace evp::exact;

namespace {
KParams kparams(const evp_b200_params_t *p) {
  KParams k;
  k.arlx1i = p->arlx1i; k.denom1 = p->denom1; k.revp = p->revp; k.brlx = p->brlx;
  k.e_factor = p->e_factor; k.epp2i = p->epp2i; k.capping = p->capping; k.Ktens = p->Ktens;
  k.u0 = p->u0; k.cosw = p->cosw; k.sinw = p->sinw; k.rhow = p->rhow;
  k.deltaminEVP = p->deltaminEVP; k.visc_method = p->visc_method;
  return k;
}
struct Host {
  CDom c{};
  std::vector<unsigned char> m[4];
  std::vector<double> q[5], s12b, uE0, vN0, vE0, uN0, spare[4];
};
void common(Host &h, const evp_b200_grid_t *g, const evp_b200_cgrid_t *cg, const int32_t *const masks[4]) {
  CDom &c = h.c;
  const size_t n = (size_t)g->nx_block * g->ny_block;
  c.nx = g->nx_block - 2; c.ny = g->ny_block - 2; c.ld = g->nx_block; c.nyd = g->ny_block;
  c.wrap_ew = g->ew_boundary_type == EVP_B200_BNDY_CYCLIC; c.wrap_ns = g->ns_boundary_type == EVP_B200_BNDY_CYCLIC;
  c.dxN = cg->dxN; c.dyE = cg->dyE; c.dxE = cg->dxE; c.dyN = cg->dyN; c.dxU = cg->dxU; c.dyU = cg->dyU; c.tarea = cg->tarea;
  c.uarea = cg->uarea; c.earea = cg->earea; c.narea = cg->narea; c.earear = cg->earear; c.narear = cg->narear;
  c.ratiodxN = cg->ratiodxN; c.ratiodxNr = cg->ratiodxNr; c.ratiodyE = cg->ratiodyE; c.ratiodyEr = cg->ratiodyEr;
  c.hm = cg->hm; c.uvm = cg->uvm; c.epm = cg->epm; c.npm = cg->npm;
  c.dxT = g->dxT; c.dyT = g->dyT; c.DminTarea = g->DminTarea;
  for (int k = 0; k < 4; ++k) {
    h.m[k].resize(n);
    for (size_t i = 0; i < n; ++i) h.m[k][i] = masks[k][i] != 0;
  }
  c.maskT = h.m[0].data(); c.maskU = h.m[1].data(); c.maskE = h.m[2].data(); c.maskN = h.m[3].data();
  for (auto &v : h.q) v.assign(n, 0.0);
  emu::launch({(c.nx + 2 + 31) / 32, (c.ny + 2 + 7) / 8, 1}, {32, 8, 1},
              [&] { cgrid_static_quotients(c, h.q[0].data(), h.q[1].data(), h.q[2].data(), h.q[3].data(), h.q[4].data()); });
  c.rhalf_dyE = h.q[0].data(); c.r_dxE = h.q[1].data(); c.rhalf_dxN = h.q[2].data(); c.r_dyN = h.q[3].data(); c.uareaavgr = h.q[4].data();
}
void zero(double *a, size_t n) { memset(a, 0, n * sizeof(double)); }
}  // namespace

// form 0: k1..k5; 1: kA<8,4> kB<8,4> k5 with momentum_at; 2: <16,2>; 3: <12,3>; 4: <4,8>; 5: <8,4> with momentum_il_at (the default)
extern "C" int emu_cgrid_run(int form, const evp_b200_grid_t *g, const evp_b200_cgrid_t *cg, const evp_b200_params_t *p, evp_b200_cfields_t *f) {
  if (g->nblocks != 1) return 1;
  Host h;
  const int32_t *masks[4] = {f->iceTmask, f->iceUmask, f->iceEmask, f->iceNmask};
  common(h, g, cg, masks);
  CDom &c = h.c;
  const size_t n = (size_t)g->nx_block * g->ny_block;
  const bool avgstr = p->visc_method == EVP_B200_VISC_AVG_STRENGTH;
  c.uvelE = f->uvelE; c.vvelE = f->vvelE; c.uvelN = f->uvelN; c.vvelN = f->vvelN; c.uvel = f->uvel; c.vvel = f->vvel;
  c.stresspT = f->stresspT; c.stressmT = f->stressmT; c.stress12T = f->stress12T; c.stress12U = f->stress12U;
  c.zetax2T = f->zetax2T; c.etax2T = f->etax2T; c.etax2U = f->etax2U; c.strengthU = f->strengthU;
  c.divergU = f->divergU; c.tensionU = f->tensionU; c.shearU = f->shearU; c.deltaU = f->deltaU;
  c.strintxE = f->strintxE; c.strintyN = f->strintyN; c.taubxE = f->taubxE; c.taubyN = f->taubyN;
  c.strength = f->strength; c.cdnE = f->cdn_ocnE; c.cdnN = f->cdn_ocnN; c.aiE = f->aiE; c.aiN = f->aiN;
  c.uocnE = f->uocnE; c.vocnE = f->vocnE; c.uocnN = f->uocnN; c.vocnN = f->vocnN; c.waterxE = f->waterxE; c.wateryN = f->wateryN;
  c.forcexE = f->forcexE; c.forceyN = f->forceyN; c.emassdti = f->emassdti; c.nmassdti = f->nmassdti; c.fmE = f->fmE; c.fmN = f->fmN;
  c.TbE = f->TbE; c.TbN = f->TbN; c.rheofactE = f->rheofactE; c.rheofactN = f->rheofactN;
  // arrays the reference zero-fills (kinds 'z' and 'y' of evp_abi.cu)
  zero(avgstr ? c.strengthU : c.etax2U, n); zero(c.divergU, n); zero(c.tensionU, n); zero(c.shearU, n); zero(c.deltaU, n);
  h.s12b.assign(c.stress12U, c.stress12U + n); c.stress12Ub = h.s12b.data();
  h.uE0.assign(c.uvelE, c.uvelE + n); h.vN0.assign(c.vvelN, c.vvelN + n); c.uvelE_init = h.uE0.data(); c.vvelN_init = h.vN0.data();
  const KParams k = kparams(p);
  const emu::Idx b{32, 8, 1}, gU{(c.nx + 31) / 32, (c.ny + 7) / 8, 1}, gT{(c.nx + 1 + 31) / 32, (c.ny + 1 + 7) / 8, 1};
  for (int ks = 0; ks < p->ndte; ++ks) {
    const int cur = ks & 1;
    auto AB = [&](auto gby, auto minb) {
      constexpr int GBY = decltype(gby)::value, MINB = decltype(minb)::value;
      const emu::Idx bb{GBX, GBY, 1}, gg{(c.nx + 1 + GBX - 2) / (GBX - 1), (c.ny + 1 + GBY - 2) / (GBY - 1), 1};
      emu::launch(gg, bb, [&] { kA_strainU_stressT<GBY, MINB>(c, k); });
      emu::launch(gg, bb, [&] { kB_stressU_momentum<GBY, MINB, false>(c, k, cur); });
    };
    using std::integral_constant;
    switch (form) {
      case 0:
        emu::launch(gU, b, [&] { k1_strain_U(c, k); });
        emu::launch(gT, b, [&] { k2_stress_T(c, k); });
        emu::launch(gU, b, [&] { k3_stress_U(c, k); });
        emu::launch(gU, b, [&] { k4_momentum(c, k); });
        break;
      case 1: AB(integral_constant<int, 8>{}, integral_constant<int, 4>{}); break;
      case 2: AB(integral_constant<int, 16>{}, integral_constant<int, 2>{}); break;
      case 3: AB(integral_constant<int, 12>{}, integral_constant<int, 3>{}); break;
      case 4: AB(integral_constant<int, 4>{}, integral_constant<int, 8>{}); break;
      case 5: {  // interleaved sqrt / division in the momentum step
        const emu::Idx bb{GBX, 8, 1}, gg{(c.nx + 1 + GBX - 2) / (GBX - 1), (c.ny + 1 + 8 - 2) / (8 - 1), 1};
        emu::launch(gg, bb, [&] { kA_strainU_stressT<8, 4>(c, k); });
        emu::launch(gg, bb, [&] { kB_stressU_momentum<8, 4, true>(c, k, cur); });
        break;
      }
      default: return 1;
    }
    emu::launch(gU, b, [&] { k5_interp(c); });
  }
  if (form != 0 && (p->ndte & 1)) memcpy(f->stress12U, h.s12b.data(), n * sizeof(double));  // the fused form ends on the second copy
  return 0;
}

extern "C" int emu_cdgrid_run(const evp_b200_grid_t *g, const evp_b200_cgrid_t *cg, const evp_b200_params_t *p, evp_b200_cdfields_t *f) {
  if (g->nblocks != 1) return 1;
  Host h;
  const int32_t *masks[4] = {f->iceTmask, f->iceUmask, f->iceEmask, f->iceNmask};
  common(h, g, cg, masks);
  CDom &c = h.c;
  const size_t n = (size_t)g->nx_block * g->ny_block;
  const bool avgstr = p->visc_method == EVP_B200_VISC_AVG_STRENGTH;
  c.uvelE = f->uvelE; c.vvelE = f->vvelE; c.uvelN = f->uvelN; c.vvelN = f->vvelN; c.uvel = f->uvel; c.vvel = f->vvel;
  c.stresspT = f->stresspT; c.stressmT = f->stressmT; c.stress12T = f->stress12T;
  c.stresspU = f->stresspU; c.stressmU = f->stressmU; c.stress12U = f->stress12U;
  c.zetax2T = f->zetax2T; c.etax2T = f->etax2T; c.zetax2U = f->zetax2U; c.etax2U = f->etax2U; c.strengthU = f->strengthU;
  c.divergU = f->divergU; c.tensionU = f->tensionU; c.shearU = f->shearU; c.deltaU = f->deltaU;
  c.strintxE = f->strintxE; c.strintyE = f->strintyE; c.strintxN = f->strintxN; c.strintyN = f->strintyN;
  c.taubxE = f->taubxE; c.taubyE = f->taubyE; c.taubxN = f->taubxN; c.taubyN = f->taubyN;
  c.strength = f->strength; c.cdnE = f->cdn_ocnE; c.cdnN = f->cdn_ocnN; c.aiE = f->aiE; c.aiN = f->aiN;
  c.uocnE = f->uocnE; c.vocnE = f->vocnE; c.uocnN = f->uocnN; c.vocnN = f->vocnN;
  c.waterxE = f->waterxE; c.wateryE = f->wateryE; c.waterxN = f->waterxN; c.wateryN = f->wateryN;
  c.forcexE = f->forcexE; c.forceyE = f->forceyE; c.forcexN = f->forcexN; c.forceyN = f->forceyN;
  c.emassdti = f->emassdti; c.nmassdti = f->nmassdti; c.fmE = f->fmE; c.fmN = f->fmN;
  c.TbE = f->TbE; c.TbN = f->TbN; c.rheofactE = f->rheofactE; c.rheofactN = f->rheofactN;
  if (avgstr) zero(c.strengthU, n); else { zero(c.zetax2U, n); zero(c.etax2U, n); }
  zero(c.divergU, n); zero(c.tensionU, n); zero(c.shearU, n); zero(c.deltaU, n);
  h.uE0.assign(c.uvelE, c.uvelE + n); h.vE0.assign(c.vvelE, c.vvelE + n); h.uN0.assign(c.uvelN, c.uvelN + n); h.vN0.assign(c.vvelN, c.vvelN + n);
  c.uvelE_init = h.uE0.data(); c.vvelE_init = h.vE0.data(); c.uvelN_init = h.uN0.data(); c.vvelN_init = h.vN0.data();
  const KParams k = kparams(p);
  const emu::Idx b{32, 8, 1}, gU{(c.nx + 31) / 32, (c.ny + 7) / 8, 1}, gT{(c.nx + 1 + 31) / 32, (c.ny + 1 + 7) / 8, 1};
  for (int ks = 0; ks < p->ndte; ++ks) {
    emu::launch(gT, b, [&] { kcd1_stress_T(c, k); });
    emu::launch(gU, b, [&] { kcd2_stress_U(c, k); });
    emu::launch(gU, b, [&] { kcd3_momentum(c, k); });
    emu::launch(gU, b, [&] { kcd4_interp(c); });
  }
  return 0;
}
