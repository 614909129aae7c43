/*
 * evp_oracle_cgrid.c -- CPU oracle for the C-grid branch of CICE's EVP subcycling loop.
 * TEST INFRASTRUCTURE ONLY (see evp_oracle.h).  Pinned bit for bit to vectors generated from the reference's source text
 * (tests/golden/ref_translit.py, ccase*); not pinned by a compiled reference binary (no Fortran compiler here).
 *
 * Restates, with the Fortran operator order, the `grid_ice == "C"` loop of
 *   evp.F90:936-1101 (cicecore/cicedyn/dynamics/ice_dyn_evp.F90) and the routines it calls:
 *   strain_rates_U    shared.F90:2319-2430        strain_rates_Tdt  shared.F90:2250-2311
 *   stressC_T         evp.F90:1758-1883           stressC_U         evp.F90:1898-1970
 *   div_stress_Ex     evp.F90:2195-2248           div_stress_Ny     evp.F90:2364-2416
 *   stepu_C, stepv_C  shared.F90:1090-1283
 *   grid_average_X2YS 'NE'                         grid.F90:4159-4211
 *   grid_average_X2YA 'NW','SE','N','E'            grid.F90:4388-4606
 * Halo points use orc_halo_update (non-tripole boundaries; the field location only matters on a fold).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "evp_oracle.h"

static const double c0 = 0.0, c1 = 1.0, p5 = 0.5;
#define IX(i, j) ((size_t)((i)-1) + (size_t)nx_block * (size_t)((j)-1))

/* visc_replpress: shared.F90:2446-2475 */
static inline void visc_replpress_c(double strength, double DminArea, double Delta, const evp_b200_params_t *p,
                                    double *zetax2, double *etax2, double *rep_prs) {
  double tmpcalc = p->capping * (strength / fmax(Delta, DminArea)) + (c1 - p->capping) * (strength / (Delta + DminArea));
  *zetax2 = (c1 + p->Ktens) * tmpcalc;
  *rep_prs = (c1 - p->Ktens) * tmpcalc * Delta;
  *etax2 = p->epp2i * (*zetax2);
}

/* strain_rates_U: shared.F90:2319-2430.  Whole output arrays are zeroed first (:2376-2379). */
static void strain_rates_U(int nx_block, int ny_block, int icellU, const int *indxUi, const int *indxUj,
                           const double *uvelE, const double *vvelE, const double *uvelN, const double *vvelN,
                           const double *uvelU, const double *vvelU, const double *dxE, const double *dyN,
                           const double *dxU, const double *dyU, const double *ratiodxN, const double *ratiodxNr,
                           const double *ratiodyE, const double *ratiodyEr, const double *epm, const double *npm,
                           double *divergU, double *tensionU, double *shearU, double *DeltaU, double e_factor) {
  const size_t n = (size_t)nx_block * ny_block;
  memset(divergU, 0, n * sizeof(double));
  memset(tensionU, 0, n * sizeof(double));
  memset(shearU, 0, n * sizeof(double));
  memset(DeltaU, 0, n * sizeof(double));
  for (int ij = 0; ij < icellU; ++ij) {
    const int i = indxUi[ij], j = indxUj[ij];
    const size_t c = IX(i, j), e = IX(i + 1, j), nn = IX(i, j + 1);
    double uNip1j = uvelN[e] * npm[e] + (npm[c] - npm[e]) * npm[c] * ratiodxN[c] * uvelN[c];
    double uNij = uvelN[c] * npm[c] + (npm[e] - npm[c]) * npm[e] * ratiodxNr[c] * uvelN[e];
    double vEijp1 = vvelE[nn] * epm[nn] + (epm[c] - epm[nn]) * epm[c] * ratiodyE[c] * vvelE[c];
    double vEij = vvelE[c] * epm[c] + (epm[nn] - epm[c]) * epm[nn] * ratiodyEr[c] * vvelE[nn];
    divergU[c] = dyU[c] * (uNip1j - uNij) + uvelU[c] * (dyN[e] - dyN[c]) + dxU[c] * (vEijp1 - vEij) +
                 vvelU[c] * (dxE[nn] - dxE[c]);
    tensionU[c] = dyU[c] * (uNip1j - uNij) - uvelU[c] * (dyN[e] - dyN[c]) - dxU[c] * (vEijp1 - vEij) +
                  vvelU[c] * (dxE[nn] - dxE[c]);
    double uEijp1 = uvelE[nn] * epm[nn] + (epm[c] - epm[nn]) * epm[c] * ratiodyE[c] * uvelE[c];
    double uEij = uvelE[c] * epm[c] + (epm[nn] - epm[c]) * epm[nn] * ratiodyEr[c] * uvelE[nn];
    double vNip1j = vvelN[e] * npm[e] + (npm[c] - npm[e]) * npm[c] * ratiodxN[c] * vvelN[c];
    double vNij = vvelN[c] * npm[c] + (npm[e] - npm[c]) * npm[e] * ratiodxNr[c] * vvelN[e];
    shearU[c] = dxU[c] * (uEijp1 - uEij) - uvelU[c] * (dxE[nn] - dxE[c]) + dyU[c] * (vNip1j - vNij) -
                vvelU[c] * (dyN[e] - dyN[c]);
    DeltaU[c] = sqrt(divergU[c] * divergU[c] + e_factor * (tensionU[c] * tensionU[c] + shearU[c] * shearU[c]));
  }
}

/* stressC_T with strain_rates_Tdt inlined: evp.F90:1758-1883, shared.F90:2250-2311 */
static void stressC_T(int nx_block, int ny_block, int icellT, const int *indxTi, const int *indxTj,
                      const double *uvelE, const double *vvelN, const double *dxN, const double *dyE,
                      const double *dxT, const double *dyT, const double *uarea, const double *DminTarea,
                      const double *strength, const double *shearU, double *zetax2T, double *etax2T,
                      double *stresspT, double *stressmT, double *stress12T, const evp_b200_params_t *p) {
  const double arlx1i = p->arlx1i, revp = p->revp, denom1 = p->denom1;
  for (int ij = 0; ij < icellT; ++ij) {
    const int i = indxTi[ij], j = indxTj[ij];
    const size_t c = IX(i, j), w = IX(i - 1, j), s = IX(i, j - 1), sw = IX(i - 1, j - 1);
    /* strain_rates_Tdt: shared.F90:2299-2307 */
    double divT = dyE[c] * uvelE[c] - dyE[w] * uvelE[w] + dxN[c] * vvelN[c] - dxN[s] * vvelN[s];
    double tensionT = (dyT[c] * dyT[c]) * (uvelE[c] / dyE[c] - uvelE[w] / dyE[w]) -
                      (dxT[c] * dxT[c]) * (vvelN[c] / dxN[c] - vvelN[s] / dxN[s]);
    /* evp.F90:1841-1855 */
    double uareaavgr = c1 / (uarea[c] + uarea[s] + uarea[sw] + uarea[w]);
    double shearTsqr = (shearU[c] * shearU[c] * uarea[c] + shearU[s] * shearU[s] * uarea[s] +
                        shearU[sw] * shearU[sw] * uarea[sw] + shearU[w] * shearU[w] * uarea[w]) * uareaavgr;
    double shearT = (shearU[c] * uarea[c] + shearU[s] * uarea[s] + shearU[sw] * uarea[sw] + shearU[w] * uarea[w]) * uareaavgr;
    double DeltaT = sqrt(divT * divT + p->e_factor * (tensionT * tensionT + shearTsqr));
    double rep_prsT;
    visc_replpress_c(strength[c], DminTarea[c], DeltaT, p, &zetax2T[c], &etax2T[c], &rep_prsT);
    stresspT[c] = (stresspT[c] * (c1 - arlx1i * revp) + arlx1i * (zetax2T[c] * divT - rep_prsT)) * denom1;
    stressmT[c] = (stressmT[c] * (c1 - arlx1i * revp) + arlx1i * etax2T[c] * tensionT) * denom1;
    stress12T[c] = (stress12T[c] * (c1 - arlx1i * revp) + arlx1i * p5 * etax2T[c] * shearT) * denom1;
  }
}

/* stressC_U: evp.F90:1898-1970 */
static void stressC_U(int nx_block, int icellU, const int *indxUi, const int *indxUj, const double *uarea,
                      const double *etax2U, const double *deltaU, const double *strengthU, const double *shearU,
                      double *stress12U, const evp_b200_params_t *p) {
  const double arlx1i = p->arlx1i, revp = p->revp, denom1 = p->denom1;
  if (p->visc_method == EVP_B200_VISC_AVG_ZETA) {
    for (int ij = 0; ij < icellU; ++ij) {
      const size_t c = IX(indxUi[ij], indxUj[ij]);
      stress12U[c] = (stress12U[c] * (c1 - arlx1i * revp) + arlx1i * p5 * etax2U[c] * shearU[c]) * denom1;
    }
  } else {
    for (int ij = 0; ij < icellU; ++ij) {
      const size_t c = IX(indxUi[ij], indxUj[ij]);
      double DminUarea = p->deltaminEVP * uarea[c];
      double lzetax2U, letax2U, lrep_prsU;
      visc_replpress_c(strengthU[c], DminUarea, deltaU[c], p, &lzetax2U, &letax2U, &lrep_prsU);
      stress12U[c] = (stress12U[c] * (c1 - arlx1i * revp) + arlx1i * p5 * letax2U * shearU[c]) * denom1;
    }
  }
}

/* div_stress_Ex: evp.F90:2232-2245; div_stress_Ny: evp.F90:2401-2414 */
static void div_stress_Ex(int nx_block, int icell, const int *indxi, const int *indxj, const double *dxE,
                          const double *dyE, const double *dxU, const double *dyT, const double *arear,
                          const double *rheofactE, const double *stressp, const double *stressm,
                          const double *stress12, double *strintx) {
  for (int ij = 0; ij < icell; ++ij) {
    const int i = indxi[ij], j = indxj[ij];
    const size_t c = IX(i, j), e = IX(i + 1, j), s = IX(i, j - 1);
    strintx[c] = rheofactE[c] * arear[c] *
                 (p5 * dyE[c] * (stressp[e] - stressp[c]) +
                  (p5 / dyE[c]) * ((dyT[e] * dyT[e]) * stressm[e] - (dyT[c] * dyT[c]) * stressm[c]) +
                  (c1 / dxE[c]) * ((dxU[c] * dxU[c]) * stress12[c] - (dxU[s] * dxU[s]) * stress12[s]));
  }
}
static void div_stress_Ny(int nx_block, int icell, const int *indxi, const int *indxj, const double *dxN,
                          const double *dyN, const double *dxT, const double *dyU, const double *arear,
                          const double *rheofactN, const double *stressp, const double *stressm,
                          const double *stress12, double *strinty) {
  for (int ij = 0; ij < icell; ++ij) {
    const int i = indxi[ij], j = indxj[ij];
    const size_t c = IX(i, j), n = IX(i, j + 1), w = IX(i - 1, j);
    strinty[c] = rheofactN[c] * arear[c] *
                 (p5 * dxN[c] * (stressp[n] - stressp[c]) -
                  (p5 / dxN[c]) * ((dxT[n] * dxT[n]) * stressm[n] - (dxT[c] * dxT[c]) * stressm[c]) +
                  (c1 / dyN[c]) * ((dyU[c] * dyU[c]) * stress12[c] - (dyU[w] * dyU[w]) * stress12[w]));
  }
}

/* stepu_C: shared.F90:1090-1184 */
static void stepu_C(int nx_block, int icell, const int *indxi, const int *indxj, const double *Cw, const double *aiE,
                    const double *uocn, const double *vocn, const double *waterx, const double *forcex,
                    const double *massdti, const double *fm, const double *strintx, double *taubx,
                    const double *uvel_init, double *uvel, const double *vvel, const double *Tb,
                    const evp_b200_params_t *p) {
  const double rhow = p->rhow, brlx = p->brlx, revp = p->revp, u0 = p->u0, cosw = p->cosw, sinw = p->sinw;
  for (int ij = 0; ij < icell; ++ij) {
    const size_t c = IX(indxi[ij], indxj[ij]);
    double uold = uvel[c], vold = vvel[c];
    double vrel = aiE[c] * rhow * Cw[c] * sqrt((uocn[c] - uold) * (uocn[c] - uold) + (vocn[c] - vold) * (vocn[c] - vold));
    double taux = vrel * waterx[c];
    double ccc = sqrt(uold * uold + vold * vold) + u0;
    double Cb = Tb[c] / ccc;
    double cca = (brlx + revp) * massdti[c] + vrel * cosw + Cb;
    double ccb = fm[c] + copysign(c1, fm[c]) * vrel * sinw;
    double cc1 = strintx[c] + forcex[c] + taux + massdti[c] * (brlx * uold + revp * uvel_init[c]);
    uvel[c] = (ccb * vold + cc1) / cca;
    taubx[c] = -uvel[c] * Cb;
  }
}
/* stepv_C: shared.F90:1189-1283 */
static void stepv_C(int nx_block, int icell, const int *indxi, const int *indxj, const double *Cw, const double *aiN,
                    const double *uocn, const double *vocn, const double *watery, const double *forcey,
                    const double *massdti, const double *fm, const double *strinty, double *tauby,
                    const double *vvel_init, const double *uvel, double *vvel, const double *Tb,
                    const evp_b200_params_t *p) {
  const double rhow = p->rhow, brlx = p->brlx, revp = p->revp, u0 = p->u0, cosw = p->cosw, sinw = p->sinw;
  for (int ij = 0; ij < icell; ++ij) {
    const size_t c = IX(indxi[ij], indxj[ij]);
    double uold = uvel[c], vold = vvel[c];
    double vrel = aiN[c] * rhow * Cw[c] * sqrt((uocn[c] - uold) * (uocn[c] - uold) + (vocn[c] - vold) * (vocn[c] - vold));
    double tauy = vrel * watery[c];
    double ccc = sqrt(uold * uold + vold * vold) + u0;
    double Cb = Tb[c] / ccc;
    double cca = (brlx + revp) * massdti[c] + vrel * cosw + Cb;
    double ccb = fm[c] + copysign(c1, fm[c]) * vrel * sinw;
    double cc2 = strinty[c] + forcey[c] + tauy + massdti[c] * (brlx * vold + revp * vvel_init[c]);
    vvel[c] = (-ccb * uold + cc2) / cca;
    tauby[c] = -vvel[c] * Cb;
  }
}

/* grid_average_X2YS 'NE' (state, masked): grid.F90:4180-4205.  work2 zeroed over the whole block first. */
static void avg_S_NE(int nx_block, int ny_block, int ilo, int ihi, int jlo, int jhi, const double *work1,
                     const double *wght1, const double *mask1, double *work2) {
  memset(work2, 0, sizeof(double) * (size_t)nx_block * ny_block);
  for (int j = jlo; j <= jhi; ++j)
    for (int i = ilo; i <= ihi; ++i) {
      const size_t c = IX(i, j), e = IX(i + 1, j), n = IX(i, j + 1), ne = IX(i + 1, j + 1);
      double wtmp = (mask1[c] * wght1[c] + mask1[e] * wght1[e] + mask1[n] * wght1[n] + mask1[ne] * wght1[ne]);
      if (wtmp != c0)
        work2[c] = (mask1[c] * work1[c] * wght1[c] + mask1[e] * work1[e] * wght1[e] + mask1[n] * work1[n] * wght1[n] +
                    mask1[ne] * work1[ne] * wght1[ne]) / wtmp;
    }
}
/* grid_average_X2YA (state, unmasked): grid.F90:4388-4606; dir 0 'NW' 1 'SE' 2 'N' 3 'E' */
static void avg_A(int dir, int nx_block, int ny_block, int ilo, int ihi, int jlo, int jhi, const double *work1,
                  const double *wght1, double *work2) {
  memset(work2, 0, sizeof(double) * (size_t)nx_block * ny_block);
  for (int j = jlo; j <= jhi; ++j)
    for (int i = ilo; i <= ihi; ++i) {
      const size_t c = IX(i, j);
      if (dir == 0) { /* NW: (i-1,j) (i,j) (i-1,j+1) (i,j+1) */
        const size_t a = IX(i - 1, j), b = c, cc = IX(i - 1, j + 1), d = IX(i, j + 1);
        double wtmp = (wght1[a] + wght1[b] + wght1[cc] + wght1[d]);
        if (wtmp != c0) work2[c] = (work1[a] * wght1[a] + work1[b] * wght1[b] + work1[cc] * wght1[cc] + work1[d] * wght1[d]) / wtmp;
      } else if (dir == 1) { /* SE: (i,j-1) (i+1,j-1) (i,j) (i+1,j) */
        const size_t a = IX(i, j - 1), b = IX(i + 1, j - 1), cc = c, d = IX(i + 1, j);
        double wtmp = (wght1[a] + wght1[b] + wght1[cc] + wght1[d]);
        if (wtmp != c0) work2[c] = (work1[a] * wght1[a] + work1[b] * wght1[b] + work1[cc] * wght1[cc] + work1[d] * wght1[d]) / wtmp;
      } else if (dir == 2) { /* N: (i,j) (i,j+1) */
        const size_t b = IX(i, j + 1);
        double wtmp = (wght1[c] + wght1[b]);
        if (wtmp != c0) work2[c] = (work1[c] * wght1[c] + work1[b] * wght1[b]) / wtmp;
      } else { /* E: (i,j) (i+1,j) */
        const size_t b = IX(i + 1, j);
        double wtmp = (wght1[c] + wght1[b]);
        if (wtmp != c0) work2[c] = (work1[c] * wght1[c] + work1[b] * wght1[b]) / wtmp;
      }
    }
}

static int build_list(const evp_b200_grid_t *g, int b, const int32_t *mask, int ext, int *li, int *lj) {
  const int nx_block = g->nx_block;
  int n = 0;
  for (int j = g->jlo[b]; j <= g->jhi[b] + ext; ++j)
    for (int i = g->ilo[b]; i <= g->ihi[b] + ext; ++i)
      if (mask[IX(i, j)]) { li[n] = i; lj[n] = j; ++n; }
  return n;
}

int orc_evp_run_cgrid(const evp_b200_grid_t *g, const evp_b200_cgrid_t *cg, const evp_b200_params_t *p,
                      evp_b200_cfields_t *f, int nthreads) {
  if (!g || !cg || !p || !f) return 1;
  if (g->ns_boundary_type == EVP_B200_BNDY_TRIPOLE) return 1; /* not restated for the C grid */
  const int nx_block = g->nx_block, ny_block = g->ny_block, nb = g->nblocks;
  const size_t npl = (size_t)nx_block * ny_block;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  int *nT = calloc(nb, sizeof(int)), *nU = calloc(nb, sizeof(int)), *nE = calloc(nb, sizeof(int)), *nN = calloc(nb, sizeof(int));
  int *Ti = malloc(sizeof(int) * npl * nb), *Tj = malloc(sizeof(int) * npl * nb), *Ui = malloc(sizeof(int) * npl * nb),
      *Uj = malloc(sizeof(int) * npl * nb), *Ei = malloc(sizeof(int) * npl * nb), *Ej = malloc(sizeof(int) * npl * nb),
      *Ni = malloc(sizeof(int) * npl * nb), *Nj = malloc(sizeof(int) * npl * nb);
  double *uvelE_init = malloc(sizeof(double) * npl * nb), *vvelN_init = malloc(sizeof(double) * npl * nb);
  if (!nT || !nU || !nE || !nN || !Ti || !Tj || !Ui || !Uj || !Ei || !Ej || !Ni || !Nj || !uvelE_init || !vvelN_init) return 1;
  for (int b = 0; b < nb; ++b) {
    const size_t o = (size_t)b * npl;
    nT[b] = build_list(g, b, f->iceTmask + o, 1, Ti + o, Tj + o); /* T incl. N/E ghost: shared.F90:740-749 */
    nU[b] = build_list(g, b, f->iceUmask + o, 0, Ui + o, Uj + o);
    nE[b] = build_list(g, b, f->iceEmask + o, 0, Ei + o, Ej + o);
    nN[b] = build_list(g, b, f->iceNmask + o, 0, Ni + o, Nj + o);
  }
  memcpy(uvelE_init, f->uvelE, sizeof(double) * npl * nb);
  memcpy(vvelN_init, f->vvelN, sizeof(double) * npl * nb);

  int rc = 0;
#define PARFOR _Pragma("omp parallel for schedule(static) num_threads(nthreads)")
#define HALO(nf, loc, type, ...)                                   \
  do {                                                             \
    double *fl_[] = {__VA_ARGS__};                                 \
    if (orc_halo_update(g, fl_, nf, loc, type)) { rc = 1; goto done; } \
  } while (0)
  for (int ksub = 1; ksub <= p->ndte; ++ksub) { /* evp.F90:938 */
    PARFOR for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * npl;
      strain_rates_U(nx_block, ny_block, nU[b], Ui + o, Uj + o, f->uvelE + o, f->vvelE + o, f->uvelN + o, f->vvelN + o,
                     f->uvel + o, f->vvel + o, cg->dxE + o, cg->dyN + o, cg->dxU + o, cg->dyU + o, cg->ratiodxN + o,
                     cg->ratiodxNr + o, cg->ratiodyE + o, cg->ratiodyEr + o, cg->epm + o, cg->npm + o, f->divergU + o,
                     f->tensionU + o, f->shearU + o, f->deltaU + o, p->e_factor);
    }
    HALO(1, 1, 0, f->shearU); /* evp.F90:964-966 */
    PARFOR for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * npl;
      stressC_T(nx_block, ny_block, nT[b], Ti + o, Tj + o, f->uvelE + o, f->vvelN + o, cg->dxN + o, cg->dyE + o, g->dxT + o,
                g->dyT + o, cg->uarea + o, g->DminTarea + o, f->strength + o, f->shearU + o, f->zetax2T + o, f->etax2T + o,
                f->stresspT + o, f->stressmT + o, f->stress12T + o, p);
    }
    HALO(4, 0, 0, f->zetax2T, f->etax2T, f->stresspT, f->stressmT); /* evp.F90:988-990 */
    PARFOR for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * npl;
      if (p->visc_method == EVP_B200_VISC_AVG_STRENGTH)
        avg_S_NE(nx_block, ny_block, g->ilo[b], g->ihi[b], g->jlo[b], g->jhi[b], f->strength + o, cg->tarea + o, cg->hm + o, f->strengthU + o);
      else
        avg_S_NE(nx_block, ny_block, g->ilo[b], g->ihi[b], g->jlo[b], g->jhi[b], f->etax2T + o, cg->tarea + o, cg->hm + o, f->etax2U + o);
      stressC_U(nx_block, nU[b], Ui + o, Uj + o, cg->uarea + o, f->etax2U + o, f->deltaU + o, f->strengthU + o, f->shearU + o,
                f->stress12U + o, p);
    }
    HALO(1, 1, 0, f->stress12U); /* evp.F90:1011-1013 */
    PARFOR for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * npl;
      div_stress_Ex(nx_block, nE[b], Ei + o, Ej + o, cg->dxE + o, cg->dyE + o, cg->dxU + o, g->dyT + o, cg->earear + o,
                    f->rheofactE + o, f->stresspT + o, f->stressmT + o, f->stress12U + o, f->strintxE + o);
      div_stress_Ny(nx_block, nN[b], Ni + o, Nj + o, cg->dxN + o, cg->dyN + o, g->dxT + o, cg->dyU + o, cg->narear + o,
                    f->rheofactN + o, f->stresspT + o, f->stressmT + o, f->stress12U + o, f->strintyN + o);
    }
    PARFOR for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * npl;
      stepu_C(nx_block, nE[b], Ei + o, Ej + o, f->cdn_ocnE + o, f->aiE + o, f->uocnE + o, f->vocnE + o, f->waterxE + o,
              f->forcexE + o, f->emassdti + o, f->fmE + o, f->strintxE + o, f->taubxE + o, uvelE_init + o, f->uvelE + o,
              f->vvelE + o, f->TbE + o, p);
      stepv_C(nx_block, nN[b], Ni + o, Nj + o, f->cdn_ocnN + o, f->aiN + o, f->uocnN + o, f->vocnN + o, f->wateryN + o,
              f->forceyN + o, f->nmassdti + o, f->fmN + o, f->strintyN + o, f->taubyN + o, vvelN_init + o, f->uvelN + o,
              f->vvelN + o, f->TbN + o, p);
    }
    HALO(1, 1, 1, f->uvelE); /* E face / N face vectors: location-independent off a fold */
    HALO(1, 1, 1, f->vvelN);
    PARFOR for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * npl;
      avg_A(0, nx_block, ny_block, g->ilo[b], g->ihi[b], g->jlo[b], g->jhi[b], f->uvelE + o, cg->earea + o, f->uvelN + o); /* E2NA 'NW' */
      avg_A(1, nx_block, ny_block, g->ilo[b], g->ihi[b], g->jlo[b], g->jhi[b], f->vvelN + o, cg->narea + o, f->vvelE + o); /* N2EA 'SE' */
      for (size_t k = 0; k < npl; ++k) {
        f->uvelN[o + k] = f->uvelN[o + k] * cg->npm[o + k];
        f->vvelE[o + k] = f->vvelE[o + k] * cg->epm[o + k];
      }
    }
    HALO(1, 1, 1, f->uvelN);
    HALO(1, 1, 1, f->vvelE);
    PARFOR for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * npl;
      avg_A(2, nx_block, ny_block, g->ilo[b], g->ihi[b], g->jlo[b], g->jhi[b], f->uvelE + o, cg->earea + o, f->uvel + o); /* E2UA 'N' */
      avg_A(3, nx_block, ny_block, g->ilo[b], g->ihi[b], g->jlo[b], g->jhi[b], f->vvelN + o, cg->narea + o, f->vvel + o); /* N2UA 'E' */
      for (size_t k = 0; k < npl; ++k) {
        f->uvel[o + k] = f->uvel[o + k] * cg->uvm[o + k];
        f->vvel[o + k] = f->vvel[o + k] * cg->uvm[o + k];
      }
    }
    HALO(2, 1, 1, f->uvel, f->vvel); /* evp.F90:1091-1093 */
  }
done:
  free(nT); free(nU); free(nE); free(nN); free(Ti); free(Tj); free(Ui); free(Uj); free(Ei); free(Ej); free(Ni); free(Nj);
  free(uvelE_init); free(vvelN_init);
  return rc;
}

/* =============================================================================================
 * grid_ice = 'CD': evp.F90:1123-1275 and the routines it calls
 *   strain_rates_Tdtsd  shared.F90:2171-2243     stressCD_T   evp.F90:1978-2080     stressCD_U  evp.F90:2088-2178
 *   div_stress_Ey       evp.F90:2252-2304        div_stress_Nx evp.F90:2308-2360    stepuv_CD   shared.F90:973-1085
 * (strain_rates_U, div_stress_Ex/Ny, the T->U and E/N->U averages are shared with the C grid above).
 * Pinned bit for bit to the transliterated reference source (tests/golden/ref_translit.py, cdcase*).
 * ============================================================================================= */
static void stressCD_T(int nx_block, int ny_block, int icellT, const int *indxTi, const int *indxTj, const double *uvelE,
                       const double *vvelE, const double *uvelN, const double *vvelN, const double *dxN, const double *dyE,
                       const double *dxT, const double *dyT, const double *DminTarea, const double *strength, double *zetax2T,
                       double *etax2T, double *stresspT, double *stressmT, double *stress12T, const evp_b200_params_t *p) {
  (void)ny_block;
  const double arlx1i = p->arlx1i, revp = p->revp, denom1 = p->denom1;
  for (int ij = 0; ij < icellT; ++ij) {
    const int i = indxTi[ij], j = indxTj[ij];
    const size_t c = IX(i, j), w = IX(i - 1, j), s = IX(i, j - 1);
    /* strain_rates_Tdt: shared.F90:2299-2307 */
    double divT = dyE[c] * uvelE[c] - dyE[w] * uvelE[w] + dxN[c] * vvelN[c] - dxN[s] * vvelN[s];
    double tensionT = (dyT[c] * dyT[c]) * (uvelE[c] / dyE[c] - uvelE[w] / dyE[w]) -
                      (dxT[c] * dxT[c]) * (vvelN[c] / dxN[c] - vvelN[s] / dxN[s]);
    /* strain_rates_Tdtsd: shared.F90:2232-2239 */
    double shearT = (dxT[c] * dxT[c]) * (uvelN[c] / dxN[c] - uvelN[s] / dxN[s]) +
                    (dyT[c] * dyT[c]) * (vvelE[c] / dyE[c] - vvelE[w] / dyE[w]);
    double DeltaT = sqrt(divT * divT + p->e_factor * (tensionT * tensionT + shearT * shearT));
    double rep_prsT;
    visc_replpress_c(strength[c], DminTarea[c], DeltaT, p, &zetax2T[c], &etax2T[c], &rep_prsT);
    stresspT[c] = (stresspT[c] * (c1 - arlx1i * revp) + arlx1i * (zetax2T[c] * divT - rep_prsT)) * denom1;
    stressmT[c] = (stressmT[c] * (c1 - arlx1i * revp) + arlx1i * etax2T[c] * tensionT) * denom1;
    stress12T[c] = (stress12T[c] * (c1 - arlx1i * revp) + arlx1i * p5 * etax2T[c] * shearT) * denom1;
  }
}

static void stressCD_U(int nx_block, int icellU, const int *indxUi, const int *indxUj, const double *uarea, const double *zetax2U,
                       const double *etax2U, const double *strengthU, const double *divergU, const double *tensionU,
                       const double *shearU, const double *deltaU, double *stresspU, double *stressmU, double *stress12U,
                       const evp_b200_params_t *p) {
  const double arlx1i = p->arlx1i, revp = p->revp, denom1 = p->denom1;
  for (int ij = 0; ij < icellU; ++ij) {
    const size_t c = IX(indxUi[ij], indxUj[ij]);
    double lzetax2U, letax2U, lrep_prsU;
    if (p->visc_method == EVP_B200_VISC_AVG_ZETA) {
      lzetax2U = zetax2U[c];
      letax2U = etax2U[c];
      lrep_prsU = (c1 - p->Ktens) / (c1 + p->Ktens) * lzetax2U * deltaU[c];
    } else {
      double DminUarea = p->deltaminEVP * uarea[c];
      visc_replpress_c(strengthU[c], DminUarea, deltaU[c], p, &lzetax2U, &letax2U, &lrep_prsU);
    }
    stresspU[c] = (stresspU[c] * (c1 - arlx1i * revp) + arlx1i * (lzetax2U * divergU[c] - lrep_prsU)) * denom1;
    stressmU[c] = (stressmU[c] * (c1 - arlx1i * revp) + arlx1i * letax2U * tensionU[c]) * denom1;
    stress12U[c] = (stress12U[c] * (c1 - arlx1i * revp) + arlx1i * p5 * letax2U * shearU[c]) * denom1;
  }
}

static void div_stress_Ey(int nx_block, int icell, const int *indxi, const int *indxj, const double *dxE, const double *dyE,
                          const double *dxU, const double *dyT, const double *arear, const double *rheofactE,
                          const double *stressp, const double *stressm, const double *stress12, double *strinty) {
  for (int ij = 0; ij < icell; ++ij) {
    const int i = indxi[ij], j = indxj[ij];
    const size_t c = IX(i, j), e = IX(i + 1, j), s = IX(i, j - 1);
    strinty[c] = rheofactE[c] * arear[c] *
                 (p5 * dxE[c] * (stressp[c] - stressp[s]) -
                  (p5 / dxE[c]) * ((dxU[c] * dxU[c]) * stressm[c] - (dxU[s] * dxU[s]) * stressm[s]) +
                  (c1 / dyE[c]) * ((dyT[e] * dyT[e]) * stress12[e] - (dyT[c] * dyT[c]) * stress12[c]));
  }
}
static void div_stress_Nx(int nx_block, int icell, const int *indxi, const int *indxj, const double *dxN, const double *dyN,
                          const double *dxT, const double *dyU, const double *arear, const double *rheofactN,
                          const double *stressp, const double *stressm, const double *stress12, double *strintx) {
  for (int ij = 0; ij < icell; ++ij) {
    const int i = indxi[ij], j = indxj[ij];
    const size_t c = IX(i, j), n = IX(i, j + 1), w = IX(i - 1, j);
    strintx[c] = rheofactN[c] * arear[c] *
                 (p5 * dyN[c] * (stressp[c] - stressp[w]) +
                  (p5 / dyN[c]) * ((dyU[c] * dyU[c]) * stressm[c] - (dyU[w] * dyU[w]) * stressm[w]) +
                  (c1 / dxN[c]) * ((dxT[n] * dxT[n]) * stress12[n] - (dxT[c] * dxT[c]) * stress12[c]));
  }
}

static void stepuv_CD(int nx_block, int icell, const int *indxi, const int *indxj, const double *Cw, const double *aiX,
                      const double *uocn, const double *vocn, const double *waterx, const double *watery, const double *forcex,
                      const double *forcey, const double *massdti, const double *fm, const double *strintx, const double *strinty,
                      double *taubx, double *tauby, const double *uvel_init, const double *vvel_init, double *uvel, double *vvel,
                      const double *Tb, const evp_b200_params_t *p) {
  const double rhow = p->rhow, brlx = p->brlx, revp = p->revp, u0 = p->u0, cosw = p->cosw, sinw = p->sinw;
  for (int ij = 0; ij < icell; ++ij) {
    const size_t c = IX(indxi[ij], indxj[ij]);
    double uold = uvel[c], vold = vvel[c];
    double vrel = aiX[c] * rhow * Cw[c] * sqrt((uocn[c] - uold) * (uocn[c] - uold) + (vocn[c] - vold) * (vocn[c] - vold));
    double taux = vrel * waterx[c];
    double tauy = vrel * watery[c];
    double ccc = sqrt(uold * uold + vold * vold) + u0;
    double Cb = Tb[c] / ccc;
    double cca = (brlx + revp) * massdti[c] + vrel * cosw + Cb;
    double ccb = fm[c] + copysign(c1, fm[c]) * vrel * sinw;
    double ab2 = cca * cca + ccb * ccb;
    double cc1 = strintx[c] + forcex[c] + taux + massdti[c] * (brlx * uold + revp * uvel_init[c]);
    double cc2 = strinty[c] + forcey[c] + tauy + massdti[c] * (brlx * vold + revp * vvel_init[c]);
    uvel[c] = (cca * cc1 + ccb * cc2) / ab2;
    vvel[c] = (cca * cc2 - ccb * cc1) / ab2;
    taubx[c] = -uvel[c] * Cb;
    tauby[c] = -vvel[c] * Cb;
  }
}

int orc_evp_run_cdgrid(const evp_b200_grid_t *g, const evp_b200_cgrid_t *cg, const evp_b200_params_t *p,
                       evp_b200_cdfields_t *f, int nthreads) {
  if (!g || !cg || !p || !f) return 1;
  if (g->ns_boundary_type == EVP_B200_BNDY_TRIPOLE) return 1;
  const int nx_block = g->nx_block, ny_block = g->ny_block, nb = g->nblocks;
  const size_t npl = (size_t)nx_block * ny_block;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  int *nT = calloc(nb, sizeof(int)), *nU = calloc(nb, sizeof(int)), *nE = calloc(nb, sizeof(int)), *nN = calloc(nb, sizeof(int));
  int *Ti = malloc(sizeof(int) * npl * nb), *Tj = malloc(sizeof(int) * npl * nb), *Ui = malloc(sizeof(int) * npl * nb),
      *Uj = malloc(sizeof(int) * npl * nb), *Ei = malloc(sizeof(int) * npl * nb), *Ej = malloc(sizeof(int) * npl * nb),
      *Ni = malloc(sizeof(int) * npl * nb), *Nj = malloc(sizeof(int) * npl * nb);
  double *init[4];
  double *cur[4] = {f->uvelE, f->vvelE, f->uvelN, f->vvelN};
  for (int q = 0; q < 4; ++q) init[q] = malloc(sizeof(double) * npl * nb);
  if (!nT || !nU || !nE || !nN || !Ti || !Tj || !Ui || !Uj || !Ei || !Ej || !Ni || !Nj || !init[0] || !init[1] || !init[2] || !init[3]) return 1;
  for (int b = 0; b < nb; ++b) {
    const size_t o = (size_t)b * npl;
    nT[b] = build_list(g, b, f->iceTmask + o, 1, Ti + o, Tj + o);
    nU[b] = build_list(g, b, f->iceUmask + o, 0, Ui + o, Uj + o);
    nE[b] = build_list(g, b, f->iceEmask + o, 0, Ei + o, Ej + o);
    nN[b] = build_list(g, b, f->iceNmask + o, 0, Ni + o, Nj + o);
  }
  for (int q = 0; q < 4; ++q) memcpy(init[q], cur[q], sizeof(double) * npl * nb);
  int rc = 0;
  for (int ksub = 1; ksub <= p->ndte; ++ksub) { /* evp.F90:1125 */
    PARFOR for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * npl;
      stressCD_T(nx_block, ny_block, nT[b], Ti + o, Tj + o, f->uvelE + o, f->vvelE + o, f->uvelN + o, f->vvelN + o, cg->dxN + o,
                 cg->dyE + o, g->dxT + o, g->dyT + o, g->DminTarea + o, f->strength + o, f->zetax2T + o, f->etax2T + o,
                 f->stresspT + o, f->stressmT + o, f->stress12T + o, p);
    }
    HALO(2, 0, 0, f->zetax2T, f->etax2T); /* evp.F90:1143-1145 */
    PARFOR for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * npl;
      if (p->visc_method == EVP_B200_VISC_AVG_STRENGTH) {
        avg_S_NE(nx_block, ny_block, g->ilo[b], g->ihi[b], g->jlo[b], g->jhi[b], f->strength + o, cg->tarea + o, cg->hm + o, f->strengthU + o);
      } else {
        avg_S_NE(nx_block, ny_block, g->ilo[b], g->ihi[b], g->jlo[b], g->jhi[b], f->zetax2T + o, cg->tarea + o, cg->hm + o, f->zetax2U + o);
        avg_S_NE(nx_block, ny_block, g->ilo[b], g->ihi[b], g->jlo[b], g->jhi[b], f->etax2T + o, cg->tarea + o, cg->hm + o, f->etax2U + o);
      }
      strain_rates_U(nx_block, ny_block, nU[b], Ui + o, Uj + o, f->uvelE + o, f->vvelE + o, f->uvelN + o, f->vvelN + o,
                     f->uvel + o, f->vvel + o, cg->dxE + o, cg->dyN + o, cg->dxU + o, cg->dyU + o, cg->ratiodxN + o,
                     cg->ratiodxNr + o, cg->ratiodyE + o, cg->ratiodyEr + o, cg->epm + o, cg->npm + o, f->divergU + o,
                     f->tensionU + o, f->shearU + o, f->deltaU + o, p->e_factor);
      stressCD_U(nx_block, nU[b], Ui + o, Uj + o, cg->uarea + o, f->zetax2U + o, f->etax2U + o, f->strengthU + o, f->divergU + o,
                 f->tensionU + o, f->shearU + o, f->deltaU + o, f->stresspU + o, f->stressmU + o, f->stress12U + o, p);
    }
    HALO(3, 0, 0, f->stresspT, f->stressmT, f->stress12T); /* evp.F90:1181-1186 */
    HALO(3, 1, 0, f->stresspU, f->stressmU, f->stress12U);
    PARFOR for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * npl;
      div_stress_Ex(nx_block, nE[b], Ei + o, Ej + o, cg->dxE + o, cg->dyE + o, cg->dxU + o, g->dyT + o, cg->earear + o,
                    f->rheofactE + o, f->stresspT + o, f->stressmT + o, f->stress12U + o, f->strintxE + o);
      div_stress_Ey(nx_block, nE[b], Ei + o, Ej + o, cg->dxE + o, cg->dyE + o, cg->dxU + o, g->dyT + o, cg->earear + o,
                    f->rheofactE + o, f->stresspU + o, f->stressmU + o, f->stress12T + o, f->strintyE + o);
      div_stress_Nx(nx_block, nN[b], Ni + o, Nj + o, cg->dxN + o, cg->dyN + o, g->dxT + o, cg->dyU + o, cg->narear + o,
                    f->rheofactN + o, f->stresspU + o, f->stressmU + o, f->stress12T + o, f->strintxN + o);
      div_stress_Ny(nx_block, nN[b], Ni + o, Nj + o, cg->dxN + o, cg->dyN + o, g->dxT + o, cg->dyU + o, cg->narear + o,
                    f->rheofactN + o, f->stresspT + o, f->stressmT + o, f->stress12U + o, f->strintyN + o);
      stepuv_CD(nx_block, nE[b], Ei + o, Ej + o, f->cdn_ocnE + o, f->aiE + o, f->uocnE + o, f->vocnE + o, f->waterxE + o,
                f->wateryE + o, f->forcexE + o, f->forceyE + o, f->emassdti + o, f->fmE + o, f->strintxE + o, f->strintyE + o,
                f->taubxE + o, f->taubyE + o, init[0] + o, init[1] + o, f->uvelE + o, f->vvelE + o, f->TbE + o, p);
      stepuv_CD(nx_block, nN[b], Ni + o, Nj + o, f->cdn_ocnN + o, f->aiN + o, f->uocnN + o, f->vocnN + o, f->waterxN + o,
                f->wateryN + o, f->forcexN + o, f->forceyN + o, f->nmassdti + o, f->fmN + o, f->strintxN + o, f->strintyN + o,
                f->taubxN + o, f->taubyN + o, init[2] + o, init[3] + o, f->uvelN + o, f->vvelN + o, f->TbN + o, p);
    }
    HALO(2, 1, 1, f->uvelE, f->vvelE); /* evp.F90:1247-1252 */
    HALO(2, 1, 1, f->uvelN, f->vvelN);
    PARFOR for (int b = 0; b < nb; ++b) {
      const size_t o = (size_t)b * npl;
      avg_A(2, nx_block, ny_block, g->ilo[b], g->ihi[b], g->jlo[b], g->jhi[b], f->uvelE + o, cg->earea + o, f->uvel + o); /* E2UA 'N' */
      avg_A(3, nx_block, ny_block, g->ilo[b], g->ihi[b], g->jlo[b], g->jhi[b], f->vvelN + o, cg->narea + o, f->vvel + o); /* N2UA 'E' */
      for (size_t k = 0; k < npl; ++k) {
        f->uvel[o + k] = f->uvel[o + k] * cg->uvm[o + k];
        f->vvel[o + k] = f->vvel[o + k] * cg->uvm[o + k];
      }
    }
    HALO(2, 1, 1, f->uvel, f->vvel); /* evp.F90:1262-1264 */
  }
done:
  free(nT); free(nU); free(nE); free(nN); free(Ti); free(Tj); free(Ui); free(Uj); free(Ei); free(Ej); free(Ni); free(Nj);
  for (int q = 0; q < 4; ++q) free(init[q]);
  return rc;
}
