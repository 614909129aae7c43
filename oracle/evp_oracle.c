/*
 * evp_oracle.c -- CPU oracle for CICE's EVP subcycling loop.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, double-precision restatement of the reference algorithm, written to be read
 * next to the Fortran it follows.  Citations are relative to /root/reference/.
 *
 *   evp.F90    = cicecore/cicedyn/dynamics/ice_dyn_evp.F90
 *   shared.F90 = cicecore/cicedyn/dynamics/ice_dyn_shared.F90
 *   core1d.F90 = cicecore/cicedyn/dynamics/ice_dyn_core1d.F90
 *   evp1d.F90  = cicecore/cicedyn/dynamics/ice_dyn_evp1d.F90
 *   boundary.F90 = cicecore/cicedyn/infrastructure/comm/mpi/ice_boundary.F90
 *   halochk.F90  = cicecore/drivers/unittest/halochk/halochk.F90
 *
 * PINNING (see evp_oracle.h): no Fortran compiler exists in this image; the restatement reproduces, bit for bit, vectors
 * generated from the reference's own source text (tests/golden/ref_translit.py) and satisfies the properties the
 * reference asserts (decomposition invariance, 2-D == 1-D, halochk values).  Not pinned by a compiled reference binary.
 *
 * Arithmetic contract: expressions keep the Fortran source's operator order (left to right
 * at equal precedence, `x**2` as x*x).  Built with -O2 -ffp-contract=off this file defines the
 * "exact" answer; built with -O3 -ffp-contract=fast it is the timed CPU baseline.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this.
 */
#include "evp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static char g_err[512] = "";
const char *orc_last_error(void) { return g_err; }
#define ORC_FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return 1; } while (0)

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ice_constants.F90:79-85 -- computed, not decimal literals */
static const double c0 = 0.0, c1 = 1.0, c2 = 2.0, c3 = 3.0, c4 = 4.0, c6 = 6.0, c9 = 9.0;
static const double p5 = 0.5, p25 = 0.25;
#define P111 (c1 / c9)
#define P055 (P111 * p5)
#define P027 (P055 * p5)
#define P166 (c1 / c6)
#define P222 (c2 / c9)
#define P333 (c1 / c3)

/* Fortran (i,j) 1-based, column major */
#define IX(i, j) ((size_t)((i)-1) + (size_t)nx_block * (size_t)((j)-1))

/* ---------------------------------------------------------------------------------------
 * strain_rates: shared.F90:2083-2163
 * ------------------------------------------------------------------------------------- */
typedef struct {
  double divune, divunw, divuse, divusw;
  double tensionne, tensionnw, tensionse, tensionsw;
  double shearne, shearnw, shearse, shearsw;
  double Deltane, Deltanw, Deltase, Deltasw;
} strain_t;

static inline void strain_rates_pt(double u_cc, double v_cc,   /* (i  ,j  ) */
                                   double u_ee, double v_ee,   /* (i-1,j  ) */
                                   double u_se, double v_se,   /* (i  ,j-1) */
                                   double u_ne, double v_ne,   /* (i-1,j-1) */
                                   double dxT, double dyT, double cxp, double cyp,
                                   double cxm, double cym, double e_factor, strain_t *s) {
  /* divergence = e_11 + e_22                                   shared.F90:2125-2133 */
  s->divune = cyp * u_cc - dyT * u_ee + cxp * v_cc - dxT * v_se;
  s->divunw = cym * u_ee + dyT * u_cc + cxp * v_ee - dxT * v_ne;
  s->divusw = cym * u_ne + dyT * u_se + cxm * v_ne + dxT * v_ee;
  s->divuse = cyp * u_se - dyT * u_ne + cxm * v_se + dxT * v_cc;
  /* tension strain rate = e_11 - e_22                          shared.F90:2136-2143 */
  s->tensionne = -cym * u_cc - dyT * u_ee + cxm * v_cc + dxT * v_se;
  s->tensionnw = -cyp * u_ee + dyT * u_cc + cxm * v_ee + dxT * v_ne;
  s->tensionsw = -cyp * u_ne + dyT * u_se + cxp * v_ne - dxT * v_ee;
  s->tensionse = -cym * u_se - dyT * u_ne + cxp * v_se - dxT * v_cc;
  /* shearing strain rate = 2*e_12                              shared.F90:2146-2153 */
  s->shearne = -cym * v_cc - dyT * v_ee - cxm * u_cc - dxT * u_se;
  s->shearnw = -cyp * v_ee + dyT * v_cc - cxm * u_ee - dxT * u_ne;
  s->shearsw = -cyp * v_ne + dyT * v_se - cxp * u_ne + dxT * u_ee;
  s->shearse = -cym * v_se - dyT * v_ne - cxp * u_se + dxT * u_cc;
  /* Delta                                                      shared.F90:2156-2159 */
  s->Deltane = sqrt(s->divune * s->divune + e_factor * (s->tensionne * s->tensionne + s->shearne * s->shearne));
  s->Deltanw = sqrt(s->divunw * s->divunw + e_factor * (s->tensionnw * s->tensionnw + s->shearnw * s->shearnw));
  s->Deltasw = sqrt(s->divusw * s->divusw + e_factor * (s->tensionsw * s->tensionsw + s->shearsw * s->shearsw));
  s->Deltase = sqrt(s->divuse * s->divuse + e_factor * (s->tensionse * s->tensionse + s->shearse * s->shearse));
}

/* visc_replpress: shared.F90:2446-2475 */
static inline void visc_replpress(double strength, double DminArea, double Delta,
                                  double capping, double Ktens, double epp2i,
                                  double *zetax2, double *etax2, double *rep_prs) {
  double tmpcalc = capping * (strength / fmax(Delta, DminArea)) +
                   (c1 - capping) * (strength / (Delta + DminArea));
  *zetax2 = (c1 + Ktens) * tmpcalc;
  *rep_prs = (c1 - Ktens) * tmpcalc * Delta;
  *etax2 = epp2i * (*zetax2);
}

/* The per-cell body shared by `stress` (evp.F90:1539-1741) and `stress_1d`
 * (core1d.F90:178-365): everything after the operands have been fetched. */
static inline void stress_cell(const strain_t *s, double strength, double DminTarea,
                               double dxT, double dyT, double dxhy, double dyhx,
                               const evp_b200_params_t *p,
                               double *sp1, double *sp2, double *sp3, double *sp4,
                               double *sm1, double *sm2, double *sm3, double *sm4,
                               double *s121, double *s122, double *s123, double *s124,
                               double str[8]) {
  const double arlx1i = p->arlx1i, denom1 = p->denom1, revp = p->revp;
  const double p111 = P111, p055 = P055, p027 = P027, p166 = P166, p222 = P222, p333 = P333;
  double zetax2ne, zetax2nw, zetax2se, zetax2sw, etax2ne, etax2nw, etax2se, etax2sw;
  double rep_prsne, rep_prsnw, rep_prsse, rep_prssw;

  /* evp.F90:1567-1577 (order ne, nw, sw, se) */
  visc_replpress(strength, DminTarea, s->Deltane, p->capping, p->Ktens, p->epp2i, &zetax2ne, &etax2ne, &rep_prsne);
  visc_replpress(strength, DminTarea, s->Deltanw, p->capping, p->Ktens, p->epp2i, &zetax2nw, &etax2nw, &rep_prsnw);
  visc_replpress(strength, DminTarea, s->Deltasw, p->capping, p->Ktens, p->epp2i, &zetax2sw, &etax2sw, &rep_prssw);
  visc_replpress(strength, DminTarea, s->Deltase, p->capping, p->Ktens, p->epp2i, &zetax2se, &etax2se, &rep_prsse);

  /* evp.F90:1585-1610 */
  *sp1 = (*sp1 * (c1 - arlx1i * revp) + arlx1i * (zetax2ne * s->divune - rep_prsne)) * denom1;
  *sp2 = (*sp2 * (c1 - arlx1i * revp) + arlx1i * (zetax2nw * s->divunw - rep_prsnw)) * denom1;
  *sp3 = (*sp3 * (c1 - arlx1i * revp) + arlx1i * (zetax2sw * s->divusw - rep_prssw)) * denom1;
  *sp4 = (*sp4 * (c1 - arlx1i * revp) + arlx1i * (zetax2se * s->divuse - rep_prsse)) * denom1;

  *sm1 = (*sm1 * (c1 - arlx1i * revp) + arlx1i * etax2ne * s->tensionne) * denom1;
  *sm2 = (*sm2 * (c1 - arlx1i * revp) + arlx1i * etax2nw * s->tensionnw) * denom1;
  *sm3 = (*sm3 * (c1 - arlx1i * revp) + arlx1i * etax2sw * s->tensionsw) * denom1;
  *sm4 = (*sm4 * (c1 - arlx1i * revp) + arlx1i * etax2se * s->tensionse) * denom1;

  *s121 = (*s121 * (c1 - arlx1i * revp) + arlx1i * p5 * etax2ne * s->shearne) * denom1;
  *s122 = (*s122 * (c1 - arlx1i * revp) + arlx1i * p5 * etax2nw * s->shearnw) * denom1;
  *s123 = (*s123 * (c1 - arlx1i * revp) + arlx1i * p5 * etax2sw * s->shearsw) * denom1;
  *s124 = (*s124 * (c1 - arlx1i * revp) + arlx1i * p5 * etax2se * s->shearse) * denom1;

  /* evp.F90:1646-1690 */
  const double stressp_1 = *sp1, stressp_2 = *sp2, stressp_3 = *sp3, stressp_4 = *sp4;
  const double stressm_1 = *sm1, stressm_2 = *sm2, stressm_3 = *sm3, stressm_4 = *sm4;
  const double stress12_1 = *s121, stress12_2 = *s122, stress12_3 = *s123, stress12_4 = *s124;

  double ssigpn = stressp_1 + stressp_2;
  double ssigps = stressp_3 + stressp_4;
  double ssigpe = stressp_1 + stressp_4;
  double ssigpw = stressp_2 + stressp_3;
  double ssigp1 = (stressp_1 + stressp_3) * p055;
  double ssigp2 = (stressp_2 + stressp_4) * p055;

  double ssigmn = stressm_1 + stressm_2;
  double ssigms = stressm_3 + stressm_4;
  double ssigme = stressm_1 + stressm_4;
  double ssigmw = stressm_2 + stressm_3;
  double ssigm1 = (stressm_1 + stressm_3) * p055;
  double ssigm2 = (stressm_2 + stressm_4) * p055;

  double ssig12n = stress12_1 + stress12_2;
  double ssig12s = stress12_3 + stress12_4;
  double ssig12e = stress12_1 + stress12_4;
  double ssig12w = stress12_2 + stress12_3;
  double ssig121 = (stress12_1 + stress12_3) * p111;
  double ssig122 = (stress12_2 + stress12_4) * p111;

  double csigpne = p111 * stressp_1 + ssigp2 + p027 * stressp_3;
  double csigpnw = p111 * stressp_2 + ssigp1 + p027 * stressp_4;
  double csigpsw = p111 * stressp_3 + ssigp2 + p027 * stressp_1;
  double csigpse = p111 * stressp_4 + ssigp1 + p027 * stressp_2;

  double csigmne = p111 * stressm_1 + ssigm2 + p027 * stressm_3;
  double csigmnw = p111 * stressm_2 + ssigm1 + p027 * stressm_4;
  double csigmsw = p111 * stressm_3 + ssigm2 + p027 * stressm_1;
  double csigmse = p111 * stressm_4 + ssigm1 + p027 * stressm_2;

  double csig12ne = p222 * stress12_1 + ssig122 + p055 * stress12_3;
  double csig12nw = p222 * stress12_2 + ssig121 + p055 * stress12_4;
  double csig12sw = p222 * stress12_3 + ssig122 + p055 * stress12_1;
  double csig12se = p222 * stress12_4 + ssig121 + p055 * stress12_2;

  double str12ew = p5 * dxT * (p333 * ssig12e + p166 * ssig12w);
  double str12we = p5 * dxT * (p333 * ssig12w + p166 * ssig12e);
  double str12ns = p5 * dyT * (p333 * ssig12n + p166 * ssig12s);
  double str12sn = p5 * dyT * (p333 * ssig12s + p166 * ssig12n);

  /* for dF/dx (u momentum)                                     evp.F90:1695-1714 */
  double strp_tmp = p25 * dyT * (p333 * ssigpn + p166 * ssigps);
  double strm_tmp = p25 * dyT * (p333 * ssigmn + p166 * ssigms);
  /* northeast (i,j) */
  str[0] = -strp_tmp - strm_tmp - str12ew + dxhy * (-csigpne + csigmne) + dyhx * csig12ne;
  /* northwest (i+1,j) */
  str[1] = strp_tmp + strm_tmp - str12we + dxhy * (-csigpnw + csigmnw) + dyhx * csig12nw;

  strp_tmp = p25 * dyT * (p333 * ssigps + p166 * ssigpn);
  strm_tmp = p25 * dyT * (p333 * ssigms + p166 * ssigmn);
  /* southeast (i,j+1) */
  str[2] = -strp_tmp - strm_tmp + str12ew + dxhy * (-csigpse + csigmse) + dyhx * csig12se;
  /* southwest (i+1,j+1) */
  str[3] = strp_tmp + strm_tmp + str12we + dxhy * (-csigpsw + csigmsw) + dyhx * csig12sw;

  /* for dF/dy (v momentum)                                     evp.F90:1719-1739 */
  strp_tmp = p25 * dxT * (p333 * ssigpe + p166 * ssigpw);
  strm_tmp = p25 * dxT * (p333 * ssigme + p166 * ssigmw);
  /* northeast (i,j) */
  str[4] = -strp_tmp + strm_tmp - str12ns - dyhx * (csigpne + csigmne) + dxhy * csig12ne;
  /* southeast (i,j+1) */
  str[5] = strp_tmp - strm_tmp - str12sn - dyhx * (csigpse + csigmse) + dxhy * csig12se;

  strp_tmp = p25 * dxT * (p333 * ssigpw + p166 * ssigpe);
  strm_tmp = p25 * dxT * (p333 * ssigmw + p166 * ssigme);
  /* northwest (i+1,j) */
  str[6] = -strp_tmp + strm_tmp + str12ns - dyhx * (csigpnw + csigmnw) + dxhy * csig12nw;
  /* southwest (i+1,j+1) */
  str[7] = strp_tmp - strm_tmp + str12sn - dyhx * (csigpsw + csigmsw) + dxhy * csig12sw;
}

/* ---------------------------------------------------------------------------------------
 * stress: evp.F90:1457-1743.  One block; str is (nx_block,ny_block,8).
 * ------------------------------------------------------------------------------------- */
void orc_stress_block(int nx_block, int ny_block, int icellT, const int *indxTi, const int *indxTj,
                      const double *uvel, const double *vvel,
                      const double *dxT, const double *dyT, const double *dxhy, const double *dyhx,
                      const double *cxp, const double *cyp, const double *cxm, const double *cym,
                      const double *DminTarea, const double *strength,
                      double *stressp_1, double *stressp_2, double *stressp_3, double *stressp_4,
                      double *stressm_1, double *stressm_2, double *stressm_3, double *stressm_4,
                      double *stress12_1, double *stress12_2, double *stress12_3, double *stress12_4,
                      double *str, const evp_b200_params_t *p) {
  const size_t npl = (size_t)nx_block * (size_t)ny_block;
  memset(str, 0, 8 * npl * sizeof(double)); /* str(:,:,:) = c0      evp.F90:1537 */

  for (int ij = 0; ij < icellT; ++ij) {
    const int i = indxTi[ij], j = indxTj[ij];
    const size_t c = IX(i, j);
    strain_t s;
    strain_rates_pt(uvel[c], vvel[c], uvel[IX(i - 1, j)], vvel[IX(i - 1, j)],
                    uvel[IX(i, j - 1)], vvel[IX(i, j - 1)], uvel[IX(i - 1, j - 1)], vvel[IX(i - 1, j - 1)],
                    dxT[c], dyT[c], cxp[c], cyp[c], cxm[c], cym[c], p->e_factor, &s);
    double st[8];
    stress_cell(&s, strength[c], DminTarea[c], dxT[c], dyT[c], dxhy[c], dyhx[c], p,
                &stressp_1[c], &stressp_2[c], &stressp_3[c], &stressp_4[c],
                &stressm_1[c], &stressm_2[c], &stressm_3[c], &stressm_4[c],
                &stress12_1[c], &stress12_2[c], &stress12_3[c], &stress12_4[c], st);
    for (int k = 0; k < 8; ++k) str[c + (size_t)k * npl] = st[k];
  }
}

/* ---------------------------------------------------------------------------------------
 * stepu: shared.F90:847-968.
 * ------------------------------------------------------------------------------------- */
static inline void stepu_cell(double uold, double vold, double Cw, double aiX, double uocn, double vocn,
                              double waterx, double watery, double forcex, double forcey,
                              double Umassdti, double fm, double uarear, double TbU,
                              double uvel_init, double vvel_init,
                              double s1, double s2, double s3, double s4,   /* str1(i,j) str2(i+1,j) str3(i,j+1) str4(i+1,j+1) */
                              double s5, double s6, double s7, double s8,   /* str5(i,j) str6(i,j+1) str7(i+1,j) str8(i+1,j+1) */
                              const evp_b200_params_t *p,
                              double *unew, double *vnew, double *strintx, double *strinty,
                              double *taubx, double *tauby) {
  const double rhow = p->rhow, brlx = p->brlx, revp = p->revp, u0 = p->u0, cosw = p->cosw, sinw = p->sinw;
  /* (magnitude of relative ocean current)*rhow*drag*aice      shared.F90:933-934 */
  double vrel = aiX * rhow * Cw * sqrt((uocn - uold) * (uocn - uold) + (vocn - vold) * (vocn - vold));
  /* ice/ocean stress                                           shared.F90:936-937 */
  double taux = vrel * waterx;
  double tauy = vrel * watery;
  double Cb = TbU / (sqrt(uold * uold + vold * vold) + u0);     /* shared.F90:939 */
  double cca = (brlx + revp) * Umassdti + vrel * cosw + Cb;     /* shared.F90:941 */
  double ccb = fm + copysign(c1, fm) * vrel * sinw;             /* shared.F90:943 */
  double ab2 = cca * cca + ccb * ccb;
  /* divergence of the internal stress tensor                   shared.F90:948-951 */
  *strintx = uarear * (s1 + s2 + s3 + s4);
  *strinty = uarear * (s5 + s6 + s7 + s8);
  /* finally, the velocity components                           shared.F90:954-960 */
  double cc1 = *strintx + forcex + taux + Umassdti * (brlx * uold + revp * uvel_init);
  double cc2 = *strinty + forcey + tauy + Umassdti * (brlx * vold + revp * vvel_init);
  *unew = (cca * cc1 + ccb * cc2) / ab2;
  *vnew = (cca * cc2 - ccb * cc1) / ab2;
  /* seabed stress component                                    shared.F90:964-965 */
  *taubx = -(*unew) * Cb;
  *tauby = -(*vnew) * Cb;
}

void orc_stepu_block(int nx_block, int ny_block, int icellU, const int *indxUi, const int *indxUj,
                     const double *Cw, const double *aiX, const double *str,
                     const double *uocn, const double *vocn, const double *waterx, const double *watery,
                     const double *forcex, const double *forcey, const double *Umassdti,
                     const double *fm, const double *uarear,
                     double *strintx, double *strinty, double *taubx, double *tauby,
                     const double *uvel_init, const double *vvel_init,
                     double *uvel, double *vvel, const double *TbU, const evp_b200_params_t *p) {
  const size_t npl = (size_t)nx_block * (size_t)ny_block;
  const double *str1 = str, *str2 = str + npl, *str3 = str + 2 * npl, *str4 = str + 3 * npl;
  const double *str5 = str + 4 * npl, *str6 = str + 5 * npl, *str7 = str + 6 * npl, *str8 = str + 7 * npl;
  for (int ij = 0; ij < icellU; ++ij) {
    const int i = indxUi[ij], j = indxUj[ij];
    const size_t c = IX(i, j), e = IX(i + 1, j), n = IX(i, j + 1), ne = IX(i + 1, j + 1);
    double un, vn;
    stepu_cell(uvel[c], vvel[c], Cw[c], aiX[c], uocn[c], vocn[c], waterx[c], watery[c],
               forcex[c], forcey[c], Umassdti[c], fm[c], uarear[c], TbU[c], uvel_init[c], vvel_init[c],
               str1[c], str2[e], str3[n], str4[ne], str5[c], str6[n], str7[e], str8[ne], p,
               &un, &vn, &strintx[c], &strinty[c], &taubx[c], &tauby[c]);
    uvel[c] = un;
    vvel[c] = vn;
  }
}

/* ---------------------------------------------------------------------------------------
 * Halo update for the dyn fields.
 *
 * Semantics restated from ice_HaloUpdate2DR8 (boundary.F90:1066-1760) as the closed-form
 * expectations halochk checks cell by cell (halochk.F90:530-830):
 *  - a ghost cell whose (i_glob,j_glob) lies inside the global domain receives the owning
 *    block's interior value (on-rank copies boundary.F90:1372-1409 / messages :1419-1449);
 *  - 'cyclic' wraps the index (ice_blocks.F90:222-276 already stores wrapped i_glob/j_glob);
 *  - 'open'/'closed' outer ghost cells are NOT touched when no fillValue is passed
 *    (ewfillouter/nsfillouter false, boundary.F90:1173-1181; nobody sends into them,
 *    :296-420), halochk.F90:541-566;
 *  - 'tripole' (u-fold), NE-corner field: on blocks whose top interior row is ny_global,
 *    ghost row je+1 <- isign * f(nx_global - ig, ny_global-1); row je is replaced by the
 *    symmetrised value 0.5*(own + isign*f(nx_global-ig, ny_global)) except the two pole
 *    points ig = nx_global/2 and nx_global which become isign*own
 *    (boundary.F90:1630-1649 averaging, :1689-1722 copy-out; halochk.F90:688-830);
 *    center field: ghost row je+1 <- isign * f(nx_global - ig + 1, ny_global), row je kept.
 *  - padded cells (global index 0, ice_blocks.F90:150-166) are never touched.
 * ------------------------------------------------------------------------------------- */
typedef struct {
  int nx_global, ny_global;
  int *blk;   /* [ny_global*nx_global] owning block (0-based) or -1 */
  int *loc;   /* local linear index inside the block plane */
} owner_map_t;

static int build_owner_map(const evp_b200_grid_t *g, owner_map_t *m) {
  const int nxg = g->nx_global, nyg = g->ny_global, nx_block = g->nx_block;
  m->nx_global = nxg;
  m->ny_global = nyg;
  m->blk = (int *)malloc(sizeof(int) * (size_t)nxg * nyg);
  m->loc = (int *)malloc(sizeof(int) * (size_t)nxg * nyg);
  if (!m->blk || !m->loc) ORC_FAIL("owner map: out of memory");
  for (size_t k = 0; k < (size_t)nxg * nyg; ++k) m->blk[k] = -1;
  for (int b = 0; b < g->nblocks; ++b) {
    const int *ig = g->i_glob + (size_t)b * g->nx_block;
    const int *jg = g->j_glob + (size_t)b * g->ny_block;
    for (int j = g->jlo[b]; j <= g->jhi[b]; ++j)
      for (int i = g->ilo[b]; i <= g->ihi[b]; ++i) {
        int gi = ig[i - 1], gj = jg[j - 1];
        if (gi < 1 || gi > nxg || gj < 1 || gj > nyg) ORC_FAIL("interior cell with global index outside domain");
        size_t k = (size_t)(gj - 1) * nxg + (gi - 1);
        m->blk[k] = b;
        m->loc[k] = (int)IX(i, j);
      }
  }
  return 0;
}

static void free_owner_map(owner_map_t *m) {
  free(m->blk);
  free(m->loc);
}

int orc_halo_update(const evp_b200_grid_t *g, double **flds, int nfld, int field_loc, int field_type) {
  owner_map_t m;
  if (build_owner_map(g, &m)) return 1;
  const int nxg = g->nx_global, nyg = g->ny_global, nx_block = g->nx_block;
  const size_t npl = (size_t)g->nx_block * g->ny_block;
  const int tripole = (g->ns_boundary_type == EVP_B200_BNDY_TRIPOLE);
  const double sgn = (field_type == 1) ? -1.0 : 1.0;

  for (int f = 0; f < nfld; ++f) {
    double *a = flds[f];
    /* pass 0: snapshot of the tripole source rows (the reference copies them into bufTripole
       before any ghost cell is written, boundary.F90:1389-1395, 1437-1443) */
    double *rowtop = NULL, *rowbelow = NULL;
    if (tripole) {
      rowtop = (double *)malloc(sizeof(double) * nxg);
      rowbelow = (double *)malloc(sizeof(double) * nxg);
      for (int gi = 1; gi <= nxg; ++gi) {
        size_t k1 = (size_t)(nyg - 1) * nxg + (gi - 1), k0 = (size_t)(nyg - 2) * nxg + (gi - 1);
        rowtop[gi - 1] = (m.blk[k1] >= 0) ? a[(size_t)m.blk[k1] * npl + m.loc[k1]] : 0.0;
        rowbelow[gi - 1] = (m.blk[k0] >= 0) ? a[(size_t)m.blk[k0] * npl + m.loc[k0]] : 0.0;
      }
    }
    /* pass 1: regular ghost cells */
    for (int b = 0; b < g->nblocks; ++b) {
      const int *ig = g->i_glob + (size_t)b * g->nx_block;
      const int *jg = g->j_glob + (size_t)b * g->ny_block;
      const int ilo = g->ilo[b], ihi = g->ihi[b], jlo = g->jlo[b], jhi = g->jhi[b];
      double *ab = a + (size_t)b * npl;
      for (int j = jlo - 1; j <= jhi + 1; ++j)
        for (int i = ilo - 1; i <= ihi + 1; ++i) {
          if (i >= ilo && i <= ihi && j >= jlo && j <= jhi) continue;
          int gi = ig[i - 1], gj = jg[j - 1];
          if (gi < 1 || gi > nxg || gj < 1 || gj > nyg) continue; /* outside / padding / tripole row */
          size_t k = (size_t)(gj - 1) * nxg + (gi - 1);
          if (m.blk[k] < 0) { ab[IX(i, j)] = 0.0; continue; } /* eliminated land block: fill, boundary.F90:1398-1408 */
          ab[IX(i, j)] = a[(size_t)m.blk[k] * npl + m.loc[k]];
        }
    }
    /* pass 2: tripole fold on blocks whose top interior row is ny_global */
    if (tripole) {
      for (int b = 0; b < g->nblocks; ++b) {
        const int *ig = g->i_glob + (size_t)b * g->nx_block;
        const int *jg = g->j_glob + (size_t)b * g->ny_block;
        const int ilo = g->ilo[b], ihi = g->ihi[b], jhi = g->jhi[b];
        if (jg[jhi - 1] != nyg) continue;
        double *ab = a + (size_t)b * npl;
        for (int i = ilo - 1; i <= ihi + 1; ++i) {
          int gi = ig[i - 1];
          if (gi < 1 || gi > nxg) continue;
          if (field_loc == 1) { /* NE corner */
            int it = nxg - gi;               /* itrip with ioffset, halochk.F90:735-777 */
            it = ((it + nxg - 1) % nxg) + 1;
            ab[IX(i, jhi + 1)] = sgn * rowbelow[it - 1];
            if (gi == nxg / 2 || gi == nxg) {
              ab[IX(i, jhi)] = sgn * rowtop[gi - 1];           /* pole points: halochk.F90:796-800 */
            } else if (gi < nxg / 2) {
              double x1 = rowtop[gi - 1], x2 = rowtop[it - 1];
              double xavg = 0.5 * (x1 + sgn * x2);             /* boundary.F90:1641-1646 */
              ab[IX(i, jhi)] = sgn * (sgn * xavg);
            } else {
              double x1 = rowtop[it - 1], x2 = rowtop[gi - 1];
              double xavg = 0.5 * (x1 + sgn * x2);
              ab[IX(i, jhi)] = sgn * xavg;
            }
          } else { /* center */
            int it = nxg - gi + 1;
            it = ((it + nxg - 1) % nxg) + 1;
            ab[IX(i, jhi + 1)] = sgn * rowtop[it - 1];
          }
        }
      }
      free(rowtop);
      free(rowbelow);
    }
  }
  free_owner_map(&m);
  return 0;
}

/* ---------------------------------------------------------------------------------------
 * Tripole grids: "Force symmetry across the tripole seam" after the loop, evp.F90:1321-1388 --
 * twelve calls ice_HaloUpdate_stress(array1, array2, halo_info, field_loc_center, field_type_scalar)
 * with (array1, array2) = (stressp_1, stressp_3), (stressp_3, stressp_1), (stressp_2, stressp_4),
 * (stressp_4, stressp_2) and the same for stressm and stress12.
 *
 * ice_HaloUpdate_stress (boundary.F90:7440-7825) moves nothing but the tripole rows: the top physical
 * row(s) of ARRAY2 go into the global tripole buffer (local copies :7641-7660, messages :7674-7688; rows
 * ny_global-1 and ny_global are buffer rows 1 and 2, address build :8117-8133), and the copy-out
 * (:7760-7794) writes ARRAY1 on every block whose top interior row is ny_global, for every local column
 * i = 1 .. ihi+nghost (ghost columns included, :8136-8157):
 *     array1(i, jhi+1) = isign * buf(nx_global - i_glob(i) + 1, 2)        isign = +1 (scalar)
 * -- centre location on a u-fold: ioffset = joffset = 0 (:7729-7732), so buffer row 3 - 1 = "replace the top
 * physical row" falls out of range and is skipped (:7779-7785).  A source cell of an eliminated land block
 * leaves the buffer at its fill value 0 (:7538).  All twelve calls read physical rows and write ghost
 * rows only: their order does not matter.
 * sig: the 12 arrays in the order of evp_b200_fields_t (stressp_1..4, stressm_1..4, stress12_1..4).
 * ------------------------------------------------------------------------------------- */
int orc_stress_symmetrise(const evp_b200_grid_t *g, double **sig) {
  if (g->ns_boundary_type != EVP_B200_BNDY_TRIPOLE) return 0;
  owner_map_t m;
  if (build_owner_map(g, &m)) return 1;
  const int nxg = g->nx_global, nyg = g->ny_global, nx_block = g->nx_block;
  const size_t npl = (size_t)g->nx_block * g->ny_block;
  static const int partner[4] = {2, 3, 0, 1}; /* 1<->3, 2<->4 */
  double *rowtop = (double *)malloc(sizeof(double) * 12 * (size_t)nxg);
  if (!rowtop) { free_owner_map(&m); ORC_FAIL("stress symmetrise: out of memory"); }
  for (int q = 0; q < 12; ++q)
    for (int gi = 1; gi <= nxg; ++gi) {
      const size_t k = (size_t)(nyg - 1) * nxg + (gi - 1);
      rowtop[(size_t)q * nxg + gi - 1] = (m.blk[k] >= 0) ? sig[q][(size_t)m.blk[k] * npl + m.loc[k]] : 0.0;
    }
  for (int b = 0; b < g->nblocks; ++b) {
    const int *ig = g->i_glob + (size_t)b * g->nx_block;
    const int *jg = g->j_glob + (size_t)b * g->ny_block;
    const int ihi = g->ihi[b], jhi = g->jhi[b];
    if (jg[jhi - 1] != nyg) continue;
    for (int q = 0; q < 12; ++q) {
      const int src = (q / 4) * 4 + partner[q % 4];
      double *ab = sig[q] + (size_t)b * npl;
      for (int i = 1; i <= ihi + 1; ++i) {
        const int gi = ig[i - 1];
        if (gi < 1 || gi > nxg) continue; /* padding */
        ab[IX(i, jhi + 1)] = rowtop[(size_t)src * nxg + (nxg - gi + 1) - 1];
      }
    }
  }
  free(rowtop);
  free_owner_map(&m);
  return 0;
}

/* ---------------------------------------------------------------------------------------
 * The subcycling loop, 2-D blocked: evp.F90:859-913.
 * Index lists are rebuilt from the masks exactly as dyn_prep2 does (shared.F90:740-789):
 * T cells over (ilo:ihi+1, jlo:jhi+1), U cells over (ilo:ihi, jlo:jhi), j outer / i inner.
 * ------------------------------------------------------------------------------------- */
static int check_grid(const evp_b200_grid_t *g) {
  if (!g) ORC_FAIL("null grid");
  if (g->abi_version != EVP_B200_ABI_VERSION) ORC_FAIL("abi_version mismatch");
  if (g->nghost != 1) ORC_FAIL("nghost must be 1");
  if (g->nblocks < 1 || g->nblocks > g->max_blocks) ORC_FAIL("bad nblocks");
  return 0;
}

int orc_evp_run_bgrid(const evp_b200_grid_t *g, const evp_b200_params_t *p, evp_b200_fields_t *f, int nthreads) {
  if (check_grid(g)) return 1;
  const int nx_block = g->nx_block, ny_block = g->ny_block, nb = g->nblocks;
  const size_t npl = (size_t)nx_block * ny_block;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif

  int *icellT = (int *)calloc(nb, sizeof(int)), *icellU = (int *)calloc(nb, sizeof(int));
  int *indxTi = (int *)malloc(sizeof(int) * npl * nb), *indxTj = (int *)malloc(sizeof(int) * npl * nb);
  int *indxUi = (int *)malloc(sizeof(int) * npl * nb), *indxUj = (int *)malloc(sizeof(int) * npl * nb);
  double *uvel_init = (double *)malloc(sizeof(double) * npl * nb);
  double *vvel_init = (double *)malloc(sizeof(double) * npl * nb);
  double *strtmp = (double *)malloc(sizeof(double) * 8 * npl * (size_t)nthreads);
  if (!icellT || !icellU || !indxTi || !indxTj || !indxUi || !indxUj || !uvel_init || !vvel_init || !strtmp)
    ORC_FAIL("out of memory");

  for (int b = 0; b < nb; ++b) {
    const int32_t *tm = f->iceTmask + (size_t)b * npl, *um = f->iceUmask + (size_t)b * npl;
    int nT = 0, nU = 0;
    for (int j = g->jlo[b]; j <= g->jhi[b] + 1; ++j)
      for (int i = g->ilo[b]; i <= g->ihi[b] + 1; ++i)
        if (tm[IX(i, j)]) { indxTi[b * npl + nT] = i; indxTj[b * npl + nT] = j; ++nT; }
    for (int j = g->jlo[b]; j <= g->jhi[b]; ++j)
      for (int i = g->ilo[b]; i <= g->ihi[b]; ++i)
        if (um[IX(i, j)]) { indxUi[b * npl + nU] = i; indxUj[b * npl + nU] = j; ++nU; }
    icellT[b] = nT;
    icellU[b] = nU;
  }
  /* uvel_init = uvel at entry (shared.F90:787-788; evp1d.F90:939-940) */
  memcpy(uvel_init, f->uvel, sizeof(double) * npl * nb);
  memcpy(vvel_init, f->vvel, sizeof(double) * npl * nb);

  int rc = 0;
  for (int ksub = 1; ksub <= p->ndte; ++ksub) {                     /* evp.F90:859 */
#pragma omp parallel for schedule(static) num_threads(nthreads)    /* evp.F90:861 */
    for (int b = 0; b < nb; ++b) {
#ifdef _OPENMP
      double *str = strtmp + (size_t)omp_get_thread_num() * 8 * npl;
#else
      double *str = strtmp;
#endif
      const size_t o = (size_t)b * npl;
      orc_stress_block(nx_block, ny_block, icellT[b], indxTi + o, indxTj + o, f->uvel + o, f->vvel + o,
                       g->dxT + o, g->dyT + o, g->dxhy + o, g->dyhx + o, g->cxp + o, g->cyp + o,
                       g->cxm + o, g->cym + o, g->DminTarea + o, f->strength + o,
                       f->stressp_1 + o, f->stressp_2 + o, f->stressp_3 + o, f->stressp_4 + o,
                       f->stressm_1 + o, f->stressm_2 + o, f->stressm_3 + o, f->stressm_4 + o,
                       f->stress12_1 + o, f->stress12_2 + o, f->stress12_3 + o, f->stress12_4 + o, str, p);
      orc_stepu_block(nx_block, ny_block, icellU[b], indxUi + o, indxUj + o, f->cdn_ocnU + o, f->aiU + o, str,
                      f->uocnU + o, f->vocnU + o, f->waterxU + o, f->wateryU + o, f->forcexU + o,
                      f->forceyU + o, f->umassdti + o, f->fmU + o, g->uarear + o,
                      f->strintxU + o, f->strintyU + o, f->taubxU + o, f->taubyU + o,
                      uvel_init + o, vvel_init + o, f->uvel + o, f->vvel + o, f->TbU + o, p);
    }
    /* U fields at NE corner                                         evp.F90:908-910 */
    double *uv[2] = {f->uvel, f->vvel};
    if (orc_halo_update(g, uv, 2, 1, 1)) { rc = 1; break; }
  }

  free(icellT); free(icellU); free(indxTi); free(indxTj); free(indxUi); free(indxUj);
  free(uvel_init); free(vvel_init); free(strtmp);
  return rc;
}

/* ---------------------------------------------------------------------------------------
 * The 1-D gather-indexed path: evp1d.F90:119-310 + core1d.F90.
 *
 * One global array (nx = nx_global+2, ny = ny_global+2).  Active T list = every cell of
 * (2:nx, 2:ny) (calc_2d_indices_init with an all-ocean tmask, evp1d.F90:532-561), neighbour
 * tables ee/ne/se/nw/sw/sse (evp1d.F90:820-826), skipTcell/skipUcell from the ice masks
 * (set_skipMe, evp1d.F90:502-528).  Geometry is recomputed per cell from HTE/HTN exactly as
 * core1d.F90:191-199 does, including the split tmparea product.
 * str1..8 are zeroed once before the loop (evp1d.F90:236-243), strintx/y are diagnosed once
 * after it (calc_diag_1d) and taub from the last Cb (evp1d.F90:1030-1031).
 * ------------------------------------------------------------------------------------- */
int orc_evp_run_bgrid_1d(const evp_b200_grid_t *g, const double *HTE, const double *HTN, double deltaminEVP,
                         const evp_b200_params_t *p, evp_b200_fields_t *f, int nthreads) {
  if (check_grid(g)) return 1;
  if (g->nblocks != 1) ORC_FAIL("1-d oracle: single global block only (the reference gathers to one array first)");
  if (g->ns_boundary_type == EVP_B200_BNDY_TRIPOLE) ORC_FAIL("1-d evp not supported with tripole (shared.F90:300-304)");
  const int nx_block = g->nx_block, ny = g->ny_block, nx = nx_block;
  const size_t n = (size_t)nx * ny;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  (void)nthreads;
  const double c1p5 = 1.5;

  double *str[8];
  for (int k = 0; k < 8; ++k) {
    str[k] = (double *)calloc(n, sizeof(double));
    if (!str[k]) ORC_FAIL("out of memory");
  }
  double *Cb = (double *)calloc(n, sizeof(double));
  double *uinit = (double *)malloc(n * sizeof(double)), *vinit = (double *)malloc(n * sizeof(double));
  unsigned char *skipT = (unsigned char *)malloc(n), *skipU = (unsigned char *)malloc(n);
  if (!Cb || !uinit || !vinit || !skipT || !skipU) ORC_FAIL("out of memory");
  memcpy(uinit, f->uvel, n * sizeof(double));
  memcpy(vinit, f->vvel, n * sizeof(double));
  for (int j = 1; j <= ny; ++j)
    for (int i = 1; i <= nx; ++i) {
      size_t c = IX(i, j);
      skipT[c] = !f->iceTmask[c];
      skipU[c] = !f->iceUmask[c] || i == nx || j == ny; /* evp1d.F90:522-523 */
    }

  for (int ksub = 1; ksub <= p->ndte; ++ksub) {
    /* stress_1d: core1d.F90:178-365 */
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int j = 2; j <= ny; ++j)
      for (int i = 2; i <= nx; ++i) {
        const size_t iw = IX(i, j);
        if (skipT[iw]) continue;
        const size_t ee = IX(i - 1, j), ne = IX(i - 1, j - 1), se = IX(i, j - 1);
        const double tmp_dxT = g->dxT[iw], tmp_dyT = g->dyT[iw];
        const double hte = HTE[iw], htn = HTN[iw], htem1 = HTE[ee], htnm1 = HTN[se];
        const double tmp_cxp = c1p5 * htn - p5 * htnm1;
        const double tmp_cyp = c1p5 * hte - p5 * htem1;
        const double tmp_cxm = -(c1p5 * htnm1 - p5 * htn);
        const double tmp_cym = -(c1p5 * htem1 - p5 * hte);
        const double tmparea = tmp_dxT * tmp_dyT;           /* split on purpose, core1d.F90:196 */
        const double tmp_DminTarea = deltaminEVP * tmparea;
        const double tmp_dxhy = p5 * (hte - htem1);
        const double tmp_dyhx = p5 * (htn - htnm1);
        strain_t s;
        strain_rates_pt(f->uvel[iw], f->vvel[iw], f->uvel[ee], f->vvel[ee], f->uvel[se], f->vvel[se],
                        f->uvel[ne], f->vvel[ne], tmp_dxT, tmp_dyT, tmp_cxp, tmp_cyp, tmp_cxm, tmp_cym,
                        p->e_factor, &s);
        double st[8];
        stress_cell(&s, f->strength[iw], tmp_DminTarea, tmp_dxT, tmp_dyT, tmp_dxhy, tmp_dyhx, p,
                    &f->stressp_1[iw], &f->stressp_2[iw], &f->stressp_3[iw], &f->stressp_4[iw],
                    &f->stressm_1[iw], &f->stressm_2[iw], &f->stressm_3[iw], &f->stressm_4[iw],
                    &f->stress12_1[iw], &f->stress12_2[iw], &f->stress12_3[iw], &f->stress12_4[iw], st);
        for (int k = 0; k < 8; ++k) str[k][iw] = st[k];
      }
    /* stepu_1d: core1d.F90:541-600 (nw=(i+1,j) sw=(i+1,j+1) sse=(i,j+1)) */
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int j = 2; j <= ny; ++j)
      for (int i = 2; i <= nx; ++i) {
        const size_t iw = IX(i, j);
        if (skipU[iw]) continue;
        const size_t nw = IX(i + 1, j), sw = IX(i + 1, j + 1), sse = IX(i, j + 1);
        double un, vn, sx, sy, tx, ty;
        const double uold = f->uvel[iw], vold = f->vvel[iw];
        stepu_cell(uold, vold, f->cdn_ocnU[iw], f->aiU[iw], f->uocnU[iw], f->vocnU[iw],
                   f->waterxU[iw], f->wateryU[iw], f->forcexU[iw], f->forceyU[iw], f->umassdti[iw],
                   f->fmU[iw], g->uarear[iw], f->TbU[iw], uinit[iw], vinit[iw],
                   str[0][iw], str[1][nw], str[2][sse], str[3][sw], str[4][iw], str[5][sse], str[6][nw], str[7][sw],
                   p, &un, &vn, &sx, &sy, &tx, &ty);
        Cb[iw] = f->TbU[iw] / (sqrt(uold * uold + vold * vold) + p->u0);
        f->uvel[iw] = un;
        f->vvel[iw] = vn;
      }
    /* evp1d_halo_update: in-array cyclic halo, evp1d.F90:1274-1301 */
    double *uv[2] = {f->uvel, f->vvel};
    if (orc_halo_update(g, uv, 2, 1, 1)) return 1;
  }
  /* calc_diag_1d (core1d.F90:607-669) and taub from Cb (evp1d.F90:1030-1031) */
  for (int j = 2; j <= ny; ++j)
    for (int i = 2; i <= nx; ++i) {
      const size_t iw = IX(i, j);
      if (skipU[iw]) continue;
      const size_t nw = IX(i + 1, j), sw = IX(i + 1, j + 1), sse = IX(i, j + 1);
      f->strintxU[iw] = g->uarear[iw] * (str[0][iw] + str[1][nw] + str[2][sse] + str[3][sw]);
      f->strintyU[iw] = g->uarear[iw] * (str[4][iw] + str[5][sse] + str[6][nw] + str[7][sw]);
      f->taubxU[iw] = -f->uvel[iw] * Cb[iw];
      f->taubyU[iw] = -f->vvel[iw] * Cb[iw];
    }
  for (int k = 0; k < 8; ++k) free(str[k]);
  free(Cb); free(uinit); free(vinit); free(skipT); free(skipU);
  return 0;
}

/* ---------------------------------------------------------------------------------------
 * deformations: shared.F90:1756-1860 (called from evp.F90:920-934 after the loop)
 * ------------------------------------------------------------------------------------- */
int orc_deformations(const evp_b200_grid_t *g, const int32_t *iceTmask, const double *uvel, const double *vvel,
                     evp_b200_deform_t *d) {
  if (check_grid(g)) return 1;
  const int nx_block = g->nx_block;
  const size_t npl = (size_t)g->nx_block * g->ny_block;
  for (int b = 0; b < g->nblocks; ++b) {
    const size_t o = (size_t)b * npl;
    const double *u = uvel + o, *v = vvel + o, *dxU = d->dxU + o, *dyU = d->dyU + o, *tarear = d->tarear + o;
    for (int j = g->jlo[b]; j <= g->jhi[b] + 1; ++j)
      for (int i = g->ilo[b]; i <= g->ihi[b] + 1; ++i) {
        const size_t c = IX(i, j), w = IX(i - 1, j), s = IX(i, j - 1), sw = IX(i - 1, j - 1);
        if (!iceTmask[o + c]) continue;
        strain_t st;
        strain_rates_pt(u[c], v[c], u[w], v[w], u[s], v[s], u[sw], v[sw], g->dxT[o + c], g->dyT[o + c], g->cxp[o + c],
                        g->cyp[o + c], g->cxm[o + c], g->cym[o + c], d->e_factor, &st);
        /* shared.F90:1825-1848 */
        const double divu = p25 * (st.divune + st.divunw + st.divuse + st.divusw) * tarear[c];
        const double tmp = p25 * (st.Deltane + st.Deltanw + st.Deltase + st.Deltasw) * tarear[c];
        d->divu[o + c] = divu;
        d->rdg_conv[o + c] = -fmin(divu, c0);
        d->rdg_shear[o + c] = p5 * (tmp - fabs(divu));
        const double tsum = st.tensionne + st.tensionnw + st.tensionse + st.tensionsw;
        const double ssum = st.shearne + st.shearnw + st.shearse + st.shearsw;
        d->shear[o + c] = p25 * tarear[c] * sqrt(tsum * tsum + ssum * ssum);
        const double dvdxn = dyU[c] * v[c] - dyU[w] * v[w];
        const double dvdxs = dyU[s] * v[s] - dyU[sw] * v[sw];
        const double dudye = dxU[c] * u[c] - dxU[s] * u[s];
        const double dudyw = dxU[w] * u[w] - dxU[sw] * u[sw];
        d->vort[o + c] = p5 * tarear[c] * (dvdxn + dvdxs - dudye - dudyw);
      }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------
 * dyn_finish: shared.F90:1291-1365 (called from evp.F90:1392-1405 after the loop), over the U list
 * (ilo:ihi, jlo:jhi where iceUmask)
 * ------------------------------------------------------------------------------------- */
int orc_dyn_finish(const evp_b200_grid_t *g, const int32_t *iceUmask, const double *uvel, const double *vvel, const double *Cw,
                   const double *uocn, const double *vocn, const double *aiX, const double *fm, evp_b200_finish_t *f) {
  if (check_grid(g)) return 1;
  const int nx_block = g->nx_block;
  const size_t npl = (size_t)g->nx_block * g->ny_block;
  for (int b = 0; b < g->nblocks; ++b) {
    const size_t o = (size_t)b * npl;
    for (int j = g->jlo[b]; j <= g->jhi[b]; ++j)
      for (int i = g->ilo[b]; i <= g->ihi[b]; ++i) {
        const size_t c = o + IX(i, j);
        if (!iceUmask[c]) continue;
        /* shared.F90:1343-1353 */
        const double du = uocn[c] - uvel[c], dv = vocn[c] - vvel[c];
        double vrel = f->rhow * Cw[c] * sqrt(du * du + dv * dv);
        vrel = vrel * aiX[c];
        f->strocnxU[c] = vrel * (du * f->cosw - dv * f->sinw * copysign(c1, fm[c]));
        f->strocnyU[c] = vrel * (dv * f->cosw + du * f->sinw * copysign(c1, fm[c]));
      }
  }
  return 0;
}
