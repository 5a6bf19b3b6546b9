// =====================================================================================
// c2ray_oracle.cpp  --  TEST INFRASTRUCTURE ONLY (parity oracle + reported CPU baseline).
//
// A CPU restatement, in plain C++17 / FP64, of the C2-Ray H+He hot path of the reference
// (garrelt/C2-Ray3Dm1D_Helium, Fortran 90).  Every function cites the reference file:line it
// follows, keeps the reference's order of operations, and reproduces the reference's
// default-real (binary32) literal semantics: a literal written without `_dp`/`d0` in the
// Fortran source is rounded to float first (helper F()).
//
// PARITY UNPINNED: the reference ships no tests / golden vectors / expected outputs and cannot be
// compiled in this image (no Fortran compiler), so this restatement is pinned only by
// self-consistency checks (serial-order vs shell-order sweep, analytic limits, an independent
// numpy restatement of the scalar kernels in tests/).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// load this library.  The product (c2-ray3dm1d_helium_b200/) never includes or links it.
//
// Build:  g++ -O2 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile)
// =====================================================================================
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "band_data.h"

namespace {

// (double)(float)literal : Fortran default-real literal promoted to real(dp)
static inline double F(float x) { return (double)x; }

// ---------------------------------------------------------------------------------------------
// Constants.  code/mathconstants.f90:21, abundances.f90:23-29, atomic.f90, cgsconstants.f90:26-103,
// cgsphotoconstants.f90:25-50, c2ray_parameters.f90:26-89, cgsastroconstants.f90
// ---------------------------------------------------------------------------------------------
const double pi = F(3.141592654f);
const double abu_he = F(0.074f);
const double abu_c = F(7.1e-7f);
const double gamma_ad = 5.0 / 3.0;
const double gamma1 = gamma_ad - 1.0;
const double c_light = 2.997925e+10;
const double hplanck = 6.6260755e-27;
const double sigma_SB = 5.670e-5;
const double k_B = 1.381e-16;
const double ev2k = F(1.0f / 8.617e-05f);
const double ev2fr = F(0.241838e15f);
const double two_pi_over_c_square = F(2.0f) * pi / (c_light * c_light);
const double eth0 = F(13.598f);
const double temph0 = eth0 * ev2k;
const double xih0 = F(1.0f);
const double fh0 = F(0.83f);
const double colh0 = F(1.3e-8f) * fh0 * xih0 / (eth0 * eth0);
const double ethe[2] = {F(24.587f), F(54.416f)};
const double temphe[2] = {ethe[0] * ev2k, ethe[1] * ev2k};
const double xihe[2] = {F(2.0f), F(1.0f)};
const double fhe[2] = {F(0.63f), F(1.30f)};
const double colhe[2] = {F(1.3e-8f) * fhe[0] * xihe[0] / (ethe[0] * ethe[0]),
                         F(1.3e-8f) * fhe[1] * xihe[1] / (ethe[1] * ethe[1])};
const double sigma_HI_at_ion_freq = F(6.346e-18f);
const double sigma_HeI_at_ion_freq = F(7.430e-18f);
const double sigma_HeII_at_ion_freq = F(1.589e-18f);
const double ion_freq_HI = ev2fr * eth0;
const double ion_freq_HeI = ev2fr * ethe[0];
const double ion_freq_HeII = ev2fr * ethe[1];
const double sigma_H_heth = 1.238e-18;
const double sigma_H_heLya = 9.907e-22;
const double sigma_He_heLya = 1.301e-20;
const double sigma_He_he2 = 1.690780687052975e-18;
const double sigma_H_he2 = 1.230695924714239e-19;
const double R_SOLAR = F(6.9599e10f);

const double epsilon = 1.0e-20;
const double convergence_fraction = F(2.5e-4f);
const double minimum_fractional_change = F(1.0e-2f);
const double minimum_fraction_of_atoms = F(1.0e-8f);
const double minitemp = F(1.0f);
const double relative_denergy = F(0.1f);

// radiation_sizes.f90:17-23
constexpr int NumFreq = 512, NumTau = 2000, NumBndin1 = 1, NumBndin2 = 26, NumBndin3 = 20;
constexpr int NumFreqBnd = NumBndin1 + NumBndin2 + NumBndin3;           // 47
constexpr int NumheatBin = NumBndin1 + NumBndin2 * 2 + NumBndin3 * 3;   // 113
// radiation_tables.f90:59-61
const double minlogtau = F(-20.0f), maxlogtau = F(4.0f);
const double dlogtau = (maxlogtau - minlogtau) / (double)(float)NumTau;

// ---------------------------------------------------------------------------------------------
// Module state (the Fortran `use`-associated globals)
// ---------------------------------------------------------------------------------------------
struct RecCol {  // cgsconstants.f90:105-133 module variables
  double arech0, brech0, areche0, breche0, oreche0, areche1, breche1, treche1;
  double colli_HI, colli_HeI, colli_HeII, v;
};

struct SedTables {  // one SED: tables (0:NumTau, 1:nb) column-major => [band][tau]
  std::vector<double> photo_thick, photo_thin, heat_thick, heat_thin;
  int lo = 1, hi = NumFreqBnd;
};

struct Ctx {
  // cooling_h.f90 tables (linear values), 1-based in Fortran -> [0..800]
  double h0_cool[801], h1_cool[801], he0_cool[801], he1_cool[801], he2_cool[801];
  double mintemp = 1.0, dtemp = 0.01;
  // radiation_sizes
  double freq_min[NumFreqBnd + 1], freq_max[NumFreqBnd + 1], delta_freq[NumFreqBnd + 1];
  double sigma_HI[NumFreqBnd + 1], sigma_HeI[NumFreqBnd + 1], sigma_HeII[NumFreqBnd + 1];
  double pl_HI[NumFreqBnd + 1], pl_HeI[NumFreqBnd + 1], pl_HeII[NumFreqBnd + 1];
  double f1ion_HI[NumFreqBnd + 1], f1ion_HeI[NumFreqBnd + 1], f1ion_HeII[NumFreqBnd + 1];
  double f2ion_HI[NumFreqBnd + 1], f2ion_HeI[NumFreqBnd + 1], f2ion_HeII[NumFreqBnd + 1];
  double f1heat_HI[NumFreqBnd + 1], f1heat_HeI[NumFreqBnd + 1], f1heat_HeII[NumFreqBnd + 1];
  double f2heat_HI[NumFreqBnd + 1], f2heat_HeI[NumFreqBnd + 1], f2heat_HeII[NumFreqBnd + 1];
  // romberg weights romw(0:512, px=9)
  double romw[NumFreq + 1];
  // SED parameters
  double T_eff, R_star, R_star2, S_star, h_over_kT;
  double pl_index, pl_minfreq, pl_maxfreq, pl_scaling, pl_S_star;
  double qpl_index, qpl_minfreq, qpl_maxfreq, qpl_scaling, qpl_S_star;
  bool use_pl = false, use_qpl = false;
  SedTables bb, pl, qpl;
  // material / grid / parameters
  int mesh[3] = {0, 0, 0};
  bool isothermal = false;
  double temper_val = 1e4;
  float clumping = 1.0f;
  // mat_ini_test.F90:37,43-45: position dependent clumping (type_of_clumping == 5) and Lyman-limit systems
  // (use_LLS; type_of_LLS 1: one value, 2: LLS_grid).  Both arrays are default real.
  std::vector<float> clumping_grid, LLS_grid;
  bool use_clumping_grid = false;
  int type_of_LLS = 0;
  double coldensh_LLS = 0.0;
  double dr[3], vol;
  double zred = 9.0, H0 = 0, Omega0 = 0.27;
  bool cosmological = true;
  int subboxsize = 10, max_subbox = 1150;
  RecCol rc;  // module globals of cgsconstants
  // sources
  int NumSrc = 0;
  std::vector<int> srcpos;  // (3,NumSrc) 1-based
  std::vector<double> NormFlux, NormFluxPL, NormFluxQPL;  // 0-based [ns-1]
  // state grids (Fortran layout, i fastest, component slowest)
  std::vector<double> ndens, xh, xhe, xh_av, xhe_av, xh_intermed, xhe_intermed;
  std::vector<float> temperature_grid;  // (N3, 0:2)
  std::vector<double> phih_grid, phihe_grid, phiheat;
  double photon_loss_all[NumFreqBnd];
  long sum_nbox_all = 0;
  long rt_updates = 0;
  // stats of the last global pass
  long nit_total = 0;
  int nit_max = 0;
};

Ctx G;

inline size_t ncell() { return (size_t)G.mesh[0] * G.mesh[1] * G.mesh[2]; }

// ---------------------------------------------------------------------------------------------
// cgsconstants.f90:140-266  ini_rec_colion_factors
// ---------------------------------------------------------------------------------------------
void ini_rec_colion_factors(double temperature, RecCol& r) {
  double lambda;
  // ini_hydrogen_recombination :171-173
  lambda = 2.0 * (temph0 / temperature);
  r.arech0 = F(1.269e-13f) * pow(lambda, 1.503) / pow(1.0 + pow(lambda / F(0.522f), F(0.470f)), F(1.923f));
  r.brech0 = F(2.753e-14f) * pow(lambda, 1.500) / pow(1.0 + pow(lambda / F(2.740f), F(0.407f)), F(2.242f));
  // ini_helium0_recombination :190-200
  if (temperature < 9.e3) {
    lambda = 2.0 * (temph0 / temperature);
    r.areche0 = 1.269e-13 * pow(lambda, 1.503) / pow(1.0 + pow(lambda / F(0.522f), F(0.470f)), F(1.923f));
    r.breche0 = 2.753e-14 * pow(lambda, 1.500) / pow(1.0 + pow(lambda / F(2.740f), F(0.407f)), F(2.242f));
  } else {
    lambda = 2.0 * (temphe[0] / temperature);
    double dielectronic = 1.9e-3 * pow(temperature, -1.5) * exp(-4.7e5 / temperature) *
                          (1.0 + 0.3 * exp(-9.4e4 / temperature));
    r.areche0 = 3.000e-14 * pow(lambda, 0.654) + dielectronic;
    r.breche0 = 1.260e-14 * pow(lambda, 0.750) + dielectronic;
  }
  r.oreche0 = r.areche0 - r.breche0;
  // ini_helium1_recombination :229-238
  lambda = 2.0 * (temphe[1] / temperature);
  r.breche1 = 5.5060e-14 * pow(lambda, 1.5) / pow(1.0 + pow(lambda / 2.740, 0.407), 2.242);
  r.areche1 = F(2.538e-13f) * pow(lambda, 1.503) / pow(1.0 + pow(lambda / 0.522, 0.470), 1.923);
  r.treche1 = 3.4e-13 * pow(temperature / 1.0e4, -0.6);
  r.v = 0.285 * pow(temperature / 1.0e4, 0.119);
  // ini_hydrogen_helium_collisional_ionization :252-255
  double sqrtt0 = sqrt(temperature);
  r.colli_HI = colh0 * sqrtt0 * exp(-temph0 / temperature);
  r.colli_HeI = colhe[0] * sqrtt0 * exp(-temphe[0] / temperature);
  r.colli_HeII = colhe[1] * sqrtt0 * exp(-temphe[1] / temperature);
}

// ---------------------------------------------------------------------------------------------
// types  (mat_ini_test.F90:70-77 ionstates; radiation_photoionrates.f90:59-81 photrates)
// ---------------------------------------------------------------------------------------------
struct IonStates {
  double h[2], he[3], h_av[2], he_av[3], h_old[2], he_old[3];
};
struct PhotRates {
  double photo_cell_HI = 0, photo_cell_HeI = 0, photo_cell_HeII = 0;
  double heat = 0, photo_in = 0, photo_out = 0;
  // the heat_cell_*, *_in_*, *_out_* members of the reference type are never assigned a non-zero
  // value on this path (set_photrates_to_zero + photrates_add only), so they are not carried.
};
inline void photrates_add(PhotRates& a, const PhotRates& b) {  // radiation_photoionrates.f90:827
  a.photo_cell_HI = a.photo_cell_HI + b.photo_cell_HI;
  a.photo_cell_HeI = a.photo_cell_HeI + b.photo_cell_HeI;
  a.photo_cell_HeII = a.photo_cell_HeII + b.photo_cell_HeII;
  a.heat = a.heat + b.heat;
  a.photo_in = a.photo_in + b.photo_in;
  a.photo_out = a.photo_out + b.photo_out;
}

// tped.f90:75-84
inline double electrondens(double ndens, const double* xh, const double* xhe) {
  return ndens * (xh[1] * (1.0 - abu_he) + abu_c + abu_he * (xhe[1] + 2.0 * xhe[2]));
}
inline double temper2pressr(double temper, double ndens, double eldens) {  // tped.f90:41-53
  return (ndens + eldens) * k_B * temper;
}
inline double pressr2temper(double pressr, double ndens, double eldens) {  // tped.f90:58-70
  return pressr / (k_B * (ndens + eldens));
}

// cooling_h.f90:40-71
double coolin(double nucldens, double eldens, const double* xh, const double* xhe, double temp0) {
  double tpos = (log10(temp0) - G.mintemp) / G.dtemp + 1.0;
  int itpos = std::min(801 - 1, std::max(1, (int)tpos));
  double dtpos = tpos - (double)(float)itpos;
  int itpos1 = std::min(801, itpos + 1);
  const int a = itpos - 1, b = itpos1 - 1;  // 0-based
  return nucldens * eldens *
         ((xh[0] * (G.h0_cool[a] + (G.h0_cool[b] - G.h0_cool[a]) * dtpos) +
           xh[1] * (G.h1_cool[a] + (G.h1_cool[b] - G.h1_cool[a]) * dtpos)) *
              (1.0 - abu_he) +
          (xhe[0] * (G.he0_cool[a] + (G.he0_cool[b] - G.he0_cool[a]) * dtpos) +
           xhe[1] * (G.he1_cool[a] + (G.he1_cool[b] - G.he1_cool[a]) * dtpos) +
           xhe[2] * (G.he2_cool[a] + (G.he2_cool[b] - G.he2_cool[a]) * dtpos)) *
              abu_he);
}

// cosmology.f90:207-234
double cosmo_cool(double e_int) {
  const double one = F(1.0f);
  double zp1 = one + G.zred;
  double dzdt = G.H0 * zp1 * sqrt(G.Omega0 * (zp1 * zp1 * zp1) + one - G.Omega0);
  return e_int * F(2.0f) / (one + G.zred) * dzdt;
}

// ---------------------------------------------------------------------------------------------
// doric.f90:317-351 prepare_doric_factors, :358-372 coldens
// ---------------------------------------------------------------------------------------------
inline double coldens(double path, double neufrac, double ndens, double abundance) {
  return neufrac * ndens * path * abundance;
}
void prepare_doric_factors(double NH, const double* NHe, double& yfrac, double& zfrac, double& y2afrac,
                           double& y2bfrac) {
  double tau_H_heth = NH * sigma_H_heth;
  double tau_He_heth = NHe[0] * sigma_HeI_at_ion_freq;
  double tau_H_heLya = NH * sigma_H_heLya;
  double tau_He_heLya = NHe[0] * sigma_He_heLya;
  double tau_H_he2th = NH * sigma_H_he2;
  double tau_He_he2th = NHe[0] * sigma_He_he2;
  double tau_He2_he2th = NHe[1] * sigma_HeII_at_ion_freq;
  yfrac = tau_H_heth / (tau_H_heth + tau_He_heth);
  zfrac = tau_H_heLya / (tau_H_heLya + tau_He_heLya);
  y2afrac = tau_He2_he2th / (tau_He2_he2th + tau_He_he2th + tau_H_he2th);
  y2bfrac = tau_He_he2th / (tau_He2_he2th + tau_He_he2th + tau_H_he2th);
}

// ---------------------------------------------------------------------------------------------
// doric.f90:35-313
// ---------------------------------------------------------------------------------------------
void doric(double dt, double rhe, double /*rhh*/, IonStates& ion, const PhotRates& phi, double yfrac, double zfrac,
           double y2afrac, double y2bfrac, const RecCol& rc, double clumping) {
  const double pfrac = 0.96;
  const double heliumfraction = abu_he / (1.0 - abu_he);
  const double ffrac = std::max(std::min(10.0 * ion.h[0], 1.0), 0.01);
  const double wfrac = (1.425 - 0.737) + 0.737 * yfrac;
  const double v = rc.v;

  const double alpha_h_B = clumping * rc.brech0;
  const double alpha_he_1 = clumping * rc.oreche0;
  const double alpha_he_B = clumping * rc.breche0;
  const double alpha_he_A = clumping * rc.areche0;
  const double alpha_he2_B = clumping * rc.breche1;
  const double alpha_he2_A = clumping * rc.areche1;
  const double alpha_he2_2 = clumping * rc.treche1;
  const double alpha_he2_1 = alpha_he2_A - alpha_he2_B;

  const double aih0 = std::max(phi.photo_cell_HI + rhe * rc.colli_HI, 1.0e-200);
  const double aihe0 = std::max(phi.photo_cell_HeI + rhe * rc.colli_HeI, 1.0e-200);
  const double aihe1 = std::max(phi.photo_cell_HeII + rhe * rc.colli_HeII, 1.0e-200);

  const double Lmat = -(aih0 + rhe * alpha_h_B);
  const double Mmat = (yfrac * rhe * alpha_he_1 + pfrac * rhe * alpha_he_B) * heliumfraction;
  const double Nmat = ((ffrac * zfrac * (1.0 - v) + v * wfrac) * alpha_he2_B + alpha_he2_2 +
                       (1.0 - y2afrac - y2bfrac) * alpha_he2_1) *
                      heliumfraction * rhe;
  const double Pmat = -aihe0 - aihe1 - rhe * (alpha_he_A - (1.0 - yfrac) * alpha_he_1);
  const double Emat = -rhe * (alpha_he2_A - y2afrac * alpha_he2_1);
  const double Qmat = -aihe0 + rhe * alpha_he2_B * (ffrac * (1.0 - zfrac) * (1.0 - v) + v * (1.425 - wfrac)) - Emat +
                      alpha_he2_1 * y2bfrac * rhe;

  const double Bcoef = Emat - Pmat;
  const double Scoef = sqrt(Bcoef * Bcoef + 4.0 * aihe1 * Qmat);
  const double QHEPcoef = 1.0 / (Qmat * aihe1 - Emat * Pmat);
  const double BminusS = Bcoef - Scoef;
  const double BplusS = Bcoef + Scoef;

  const double lambda1 = Lmat;
  const double lambda2 = 0.5 * (Emat + Pmat - Scoef);
  const double lambda3 = 0.5 * (Emat + Pmat + Scoef);

  const double rx = -1.0 / Lmat * (aih0 + (Mmat * Emat - Nmat * aihe1) * (aihe0 * QHEPcoef));
  const double ry = aihe0 * (Emat * QHEPcoef);
  const double rz = -aihe0 * (aihe1 * QHEPcoef);

  const double twoaihe1 = 2.0 * aihe1;
  const double eigv2x = -Nmat / (Lmat - lambda2) + (Mmat / twoaihe1) * BplusS / (Lmat - lambda2);
  const double eigv3x = (-twoaihe1 * Nmat + Mmat * (BminusS)) / (twoaihe1 * (Lmat - lambda3));
  const double eigv2y = (-BplusS) / (twoaihe1);
  const double eigv3y = (-BminusS) / (twoaihe1);

  const double Rcoef = twoaihe1 * (ry - ion.he_old[1]);
  const double Tcoef = rz - ion.he_old[2];

  const double coef2 = (Rcoef + (BminusS)*Tcoef) / (2.0 * Scoef);
  const double coef3 = -(Rcoef + (BplusS)*Tcoef) / (2.0 * Scoef);
  const double coef1 = -rx + (eigv3x - eigv2x) * (Rcoef / (2.0 * Scoef)) +
                       Tcoef * ((BplusS * eigv3x / (2.0 * Scoef) - BminusS * eigv2x / (2.0 * Scoef))) + ion.h_old[1];

  const double lam1dt = dt * lambda1, lam2dt = dt * lambda2, lam3dt = dt * lambda3;
  const double elam1dt = exp(lam1dt), elam2dt = exp(lam2dt), elam3dt = exp(lam3dt);

  ion.h[1] = coef1 * elam1dt + coef2 * elam2dt * eigv2x + coef3 * elam3dt * eigv3x + rx;
  ion.he[1] = coef2 * elam2dt * eigv2y + coef3 * elam3dt * eigv3y + ry;
  ion.he[2] = coef2 * elam2dt + coef3 * elam3dt + rz;
  ion.h[0] = 1.0 - ion.h[1];
  ion.he[0] = 1.0 - ion.he[1] - ion.he[2];

  if (ion.h[0] < epsilon) { ion.h[0] = epsilon; ion.h[1] = 1.0 - epsilon; }
  if (ion.h[1] < epsilon) { ion.h[1] = epsilon; ion.h[0] = 1.0 - epsilon; }
  if ((ion.he[0] <= epsilon) || (ion.he[1] <= epsilon) || (ion.he[2] <= epsilon)) {
    if (ion.he[0] < epsilon) ion.he[0] = epsilon;
    if (ion.he[1] < epsilon) ion.he[1] = epsilon;
    if (ion.he[2] < epsilon) ion.he[2] = epsilon;
    double normfac = ion.he[0] + ion.he[1] + ion.he[2];
    ion.he[0] = ion.he[0] / normfac;
    ion.he[1] = ion.he[1] / normfac;
    ion.he[2] = ion.he[2] / normfac;
  }

  const double lim = F(1.0e-8f);
  double avg_factor_1, avg_factor_2, avg_factor_3;
  if (fabs(lam1dt) < lim) avg_factor_1 = coef1; else avg_factor_1 = coef1 * (elam1dt - 1.0) / lam1dt;
  if (fabs(lam2dt) < lim) avg_factor_2 = coef2; else avg_factor_2 = coef2 * (elam2dt - 1.0) / lam2dt;
  if (fabs(lam3dt) < lim) avg_factor_3 = coef3; else avg_factor_3 = coef3 * (elam3dt - 1.0) / lam3dt;

  ion.h_av[1] = rx + avg_factor_1 + eigv2x * avg_factor_2 + eigv3x * avg_factor_3;
  ion.he_av[1] = ry + eigv2y * avg_factor_2 + eigv3y * avg_factor_3;
  ion.he_av[2] = rz + avg_factor_2 + avg_factor_3;
  ion.h_av[0] = 1.0 - ion.h_av[1];
  ion.he_av[0] = 1.0 - ion.he_av[1] - ion.he_av[2];

  if (ion.h_av[1] < epsilon) { ion.h_av[1] = epsilon; ion.h_av[0] = 1.0 - epsilon; }
  if (ion.h_av[0] < epsilon) { ion.h_av[0] = epsilon; ion.h_av[1] = 1.0 - epsilon; }
  if ((ion.he_av[0] <= epsilon) || (ion.he_av[1] <= epsilon) || (ion.he_av[2] <= epsilon)) {
    if (ion.he_av[1] < epsilon) ion.he_av[1] = epsilon;
    if (ion.he_av[2] < epsilon) ion.he_av[2] = epsilon;
    if (ion.he_av[0] < epsilon) ion.he_av[0] = epsilon;
    double normfac = ion.he_av[0] + ion.he_av[1] + ion.he_av[2];
    ion.he_av[0] = ion.he_av[0] / normfac;
    ion.he_av[1] = ion.he_av[1] / normfac;
    ion.he_av[2] = ion.he_av[2] / normfac;
  }
}

// ---------------------------------------------------------------------------------------------
// thermal.f90:22-174
// ---------------------------------------------------------------------------------------------
void thermal(double dt, double& end_temper, double& avg_temper, double ndens_electron, double ndens_atom,
             const IonStates& ion, const PhotRates& phi, int* n_sub = nullptr) {
  double heating = phi.heat;
  double internal_energy =
      temper2pressr(end_temper, ndens_atom, electrondens(ndens_atom, ion.h_old, ion.he_old)) / (gamma1);
  double cosmo_cool_rate;
  if (G.cosmological) cosmo_cool_rate = cosmo_cool(internal_energy); else cosmo_cool_rate = 0.0;
  int i_heating = 0;
  if (end_temper > minitemp) {
    double cumulative_time = 0.0;
    avg_temper = 0.0;
    double initial_temp = end_temper;
    for (;;) {
      i_heating = i_heating + 1;
      double cooling = coolin(ndens_atom, ndens_electron, ion.h_av, ion.he_av, end_temper) + cosmo_cool_rate;
      double thermal_rate = std::max(1e-50, fabs(cooling - heating));
      double thermal_timescale = internal_energy / fabs(thermal_rate);
      double dt_thermal = relative_denergy * thermal_timescale;
      double dt_ODE = std::min(dt_thermal, dt - cumulative_time);
      internal_energy = internal_energy + dt_ODE * (heating - cooling);
      avg_temper = avg_temper + F(0.5f) * end_temper * dt_ODE;
      end_temper = pressr2temper(internal_energy * gamma1, ndens_atom, electrondens(ndens_atom, ion.h_av, ion.he_av));
      avg_temper = avg_temper + F(0.5f) * end_temper * dt_ODE;
      if (end_temper < minitemp) {
        internal_energy = temper2pressr(minitemp, ndens_atom, electrondens(ndens_atom, ion.h_av, ion.he_av));
        end_temper = minitemp;
      }
      cumulative_time = cumulative_time + dt_ODE;
      if (cumulative_time >= dt || fabs(cumulative_time - dt) < F(1e-6f) * dt) break;
      if (i_heating > 10000) break;
    }
    if (dt > 0.0) avg_temper = avg_temper / dt; else avg_temper = initial_temp;
    end_temper = pressr2temper(internal_energy * gamma1, ndens_atom, electrondens(ndens_atom, ion.h, ion.he));
  }
  if (n_sub) *n_sub = i_heating;
}

// ---------------------------------------------------------------------------------------------
// romberg.f90:22-96  romberg_initialisation (weights for px = log2(nmax) only are kept)
// ---------------------------------------------------------------------------------------------
void romberg_initialisation(int nmax, double* romw_out) {
  const int maxpow = 14;
  int pmax = (int)lround(log((double)nmax) / F(logf(2.0f)));
  std::vector<double> a(maxpow + 1, 0.0), b(maxpow + 1, 0.0);
  std::vector<std::vector<double>> s(maxpow + 1, std::vector<double>(maxpow + 1, 0.0));
  std::vector<std::vector<double>> romw(pmax + 1, std::vector<double>((1 << pmax) + 1, 0.0));  // romw[i][j]
  for (int k = 1; k <= pmax; k++) {
    float four_k = powf(4.0f, (float)k);  // 4.0**k, exact in binary32 for k<=11
    b[k] = (double)(-1.0f / (four_k - 1.0f));
    a[k] = -b[k] * (double)four_k;
  }
  for (int i = 1; i <= pmax; i++) s[i][0] = 0.0;
  for (int k = 0; k <= pmax; k++) {
    s[k][0] = 1.0;
    for (int j = 1; j <= pmax; j++)
      for (int i = pmax; i >= j; i--) s[i][j] = a[j] * s[i][j - 1] + b[j] * s[i - 1][j - 1];
    for (int i = k; i <= pmax; i++)
      for (int j = 0; j <= (1 << k); j++) {
        int idx = (1 << (i - k)) * j;
        romw[i][idx] = s[i][i] * (double)(1 << (i - k)) + romw[i][idx];
      }
    s[k][0] = 0.0;
  }
  for (int i = 0; i <= pmax; i++) {
    romw[i][0] = F(0.5f) * romw[i][0];
    romw[i][1 << i] = F(0.5f) * romw[i][1 << i];
  }
  for (int j = 0; j <= nmax; j++) romw_out[j] = romw[pmax][j];
}

// romberg.f90:100-142 scalar_romberg with ny=0 (romw(0,-1)=1)
double scalar_romberg(const double* f, const double* w, int nx) {
  double integral = 0.0;
  for (int x = 0; x <= nx; x++) integral = integral + f[x] * w[x] * G.romw[x] * 1.0;
  return integral;
}

// ---------------------------------------------------------------------------------------------
// radiation_sizes.f90:62-688 setup_scalingfactors (NumBndin1/2/3 = 1/26/20 branches only)
// ---------------------------------------------------------------------------------------------
void setup_scalingfactors() {
  G.freq_max[1] = ion_freq_HeI;
  for (int i = 0; i < 25; i++) G.freq_max[2 + i] = ion_freq_HeI * BD_FREQMAX_MULT_HEI[i];
  G.freq_max[NumBndin1 + NumBndin2] = ion_freq_HeII;
  for (int i = 0; i < 20; i++) G.freq_max[NumBndin1 + NumBndin2 + 1 + i] = ion_freq_HeII * BD_FREQMAX_MULT_HEII[i];
  G.freq_min[1] = ion_freq_HI;
  for (int i = 2; i <= NumFreqBnd; i++) G.freq_min[i] = G.freq_max[i - 1];
  for (int i = 1; i <= NumFreqBnd; i++) G.delta_freq[i] = (G.freq_max[i] - G.freq_min[i]) / (double)(float)NumFreq;
  G.sigma_HI[1] = sigma_HI_at_ion_freq; G.sigma_HeI[1] = 0.0; G.sigma_HeII[1] = 0.0;
  G.pl_HI[1] = BD_PLIDX_HI_B1; G.pl_HeI[1] = 0; G.pl_HeII[1] = 0;
  for (int i = 0; i < 26; i++) {
    int b = NumBndin1 + 1 + i;
    G.sigma_HI[b] = BD_SIGMA_HI_B2[i]; G.sigma_HeI[b] = BD_SIGMA_HEI_B2[i]; G.sigma_HeII[b] = 0.0;
    G.pl_HI[b] = BD_PLIDX_HI_B2[i]; G.pl_HeI[b] = BD_PLIDX_HEI_B2[i]; G.pl_HeII[b] = 0;
    G.f1ion_HI[b] = BD_F1ION_HI_B2[i]; G.f1ion_HeI[b] = BD_F1ION_HEI_B2[i]; G.f1ion_HeII[b] = BD_F1ION_HEII_B2[i];
    G.f2ion_HI[b] = BD_F2ION_HI_B2[i]; G.f2ion_HeI[b] = BD_F2ION_HEI_B2[i]; G.f2ion_HeII[b] = BD_F2ION_HEII_B2[i];
    G.f1heat_HI[b] = BD_F1HEAT_HI_B2[i]; G.f1heat_HeI[b] = BD_F1HEAT_HEI_B2[i]; G.f1heat_HeII[b] = BD_F1HEAT_HEII_B2[i];
    G.f2heat_HI[b] = BD_F2HEAT_HI_B2[i]; G.f2heat_HeI[b] = BD_F2HEAT_HEI_B2[i]; G.f2heat_HeII[b] = BD_F2HEAT_HEII_B2[i];
  }
  for (int i = 0; i < 20; i++) {
    int b = NumBndin1 + NumBndin2 + 1 + i;
    G.sigma_HI[b] = BD_SIGMA_HI_B3[i]; G.sigma_HeI[b] = BD_SIGMA_HEI_B3[i]; G.sigma_HeII[b] = BD_SIGMA_HEII_B3[i];
    G.pl_HI[b] = BD_PLIDX_HI_B3[i]; G.pl_HeI[b] = BD_PLIDX_HEI_B3[i]; G.pl_HeII[b] = BD_PLIDX_HEII_B3[i];
    G.f1ion_HI[b] = BD_F1ION_HI_B3[i]; G.f1ion_HeI[b] = BD_F1ION_HEI_B3[i]; G.f1ion_HeII[b] = BD_F1ION_HEII_B3[i];
    G.f2ion_HI[b] = BD_F2ION_HI_B3[i]; G.f2ion_HeI[b] = BD_F2ION_HEI_B3[i]; G.f2ion_HeII[b] = BD_F2ION_HEII_B3[i];
    G.f1heat_HI[b] = BD_F1HEAT_HI_B3[i]; G.f1heat_HeI[b] = BD_F1HEAT_HEI_B3[i]; G.f1heat_HeII[b] = BD_F1HEAT_HEII_B3[i];
    G.f2heat_HI[b] = BD_F2HEAT_HI_B3[i]; G.f2heat_HeI[b] = BD_F2HEAT_HEI_B3[i]; G.f2heat_HeII[b] = BD_F2HEAT_HEII_B3[i];
  }
}

// radiation_sed_parameters.f90:803-842 blackbody_sed / powerlaw_sed ("S" = photon sense)
double blackbody_sed_S(double frequency) {
  if (frequency * G.h_over_kT <= 709.0)
    return two_pi_over_c_square * frequency * frequency / (exp(frequency * G.h_over_kT) - 1.0);
  return two_pi_over_c_square * frequency * frequency / (exp((frequency * G.h_over_kT) / 2.0)) /
         (exp((frequency * G.h_over_kT) / 2.0));
}
// radiation_sed_parameters.f90:746-800 integrate_sed(...,"S")
double integrate_sed_S(double fmin, double fmax, char sourcetype) {
  double frequency[NumFreq + 1], weight[NumFreq + 1], integrand[NumFreq + 1];
  double freq_step = (fmax - fmin) / (double)(float)NumFreq;
  for (int i = 0; i <= NumFreq; i++) { frequency[i] = fmin + freq_step * (double)(float)i; weight[i] = freq_step; }
  if (sourcetype == 'B') {
    for (int i = 0; i <= NumFreq; i++) integrand[i] = blackbody_sed_S(frequency[i]);
    return F(4.0f) * pi * G.R_star * G.R_star * scalar_romberg(integrand, weight, NumFreq);
  } else if (sourcetype == 'P') {
    for (int i = 0; i <= NumFreq; i++) integrand[i] = pow(frequency[i], -G.pl_index);
    return G.pl_scaling * scalar_romberg(integrand, weight, NumFreq);
  } else {
    for (int i = 0; i <= NumFreq; i++) integrand[i] = pow(frequency[i], -G.qpl_index);
    return G.qpl_scaling * scalar_romberg(integrand, weight, NumFreq);
  }
}

// radiation_tables.f90:172-422 spec_integration and helpers :527-899
void spec_integration() {
  std::vector<double> tau(NumTau + 1);
  for (int i = 1; i <= NumTau; i++) tau[i] = pow(F(10.0f), minlogtau + dlogtau * (double)(float)(i - 1));
  tau[0] = 0.0;
  // :194-199
  G.bb.lo = 1; G.bb.hi = NumFreqBnd;
  for (int b = 1; b <= NumFreqBnd; b++)
    if (G.freq_min[b] * G.h_over_kT > F(25.f)) { G.bb.hi = b - 1; break; }
  // :208-247
  G.pl.hi = NumFreqBnd;
  for (int b = 1; b <= NumFreqBnd; b++) if (G.freq_min[b] > G.pl_maxfreq) { G.pl.hi = b - 1; break; }
  G.pl.lo = 1;
  for (int b = NumFreqBnd; b >= 1; b--) if (G.freq_min[b] < G.pl_minfreq) { G.pl.lo = b; break; }
  G.qpl.hi = NumFreqBnd;
  for (int b = 1; b <= NumFreqBnd; b++) if (G.freq_min[b] > G.qpl_maxfreq) { G.qpl.hi = b - 1; break; }
  G.qpl.lo = 1;
  for (int b = NumFreqBnd; b >= 1; b--) if (G.freq_min[b] < G.qpl_minfreq) { G.qpl.lo = b; break; }

  SedTables* seds[3] = {&G.bb, &G.pl, &G.qpl};
  for (auto* s : seds) {
    s->photo_thick.assign((size_t)NumFreqBnd * (NumTau + 1), 0.0);
    s->photo_thin.assign((size_t)NumFreqBnd * (NumTau + 1), 0.0);
    s->heat_thick.assign((size_t)NumheatBin * (NumTau + 1), 0.0);
    s->heat_thin.assign((size_t)NumheatBin * (NumTau + 1), 0.0);
  }
  const bool active[3] = {true, G.use_pl, G.use_qpl};

#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 1; b <= NumFreqBnd; b++) {
    double frequency[NumFreq + 1], csfd[NumFreq + 1];
    // set_frequency_array :527-537
    for (int i = 0; i <= NumFreq; i++) frequency[i] = G.freq_min[b] + G.delta_freq[b] * (double)(float)i;
    // set_cross_section_freq_dependence :566-589 (band 1: HI index, band 2: HeI index, band 3: HeII index)
    double idx = (b <= NumBndin1) ? G.pl_HI[b] : (b <= NumBndin1 + NumBndin2 ? G.pl_HeI[b] : G.pl_HeII[b]);
    for (int i = 0; i <= NumFreq; i++) csfd[i] = pow(frequency[i] / G.freq_min[b], -idx);
    const int nsp = (b <= NumBndin1) ? 1 : (b <= NumBndin1 + NumBndin2 ? 2 : 3);
    const int hcol0 = (b <= NumBndin1) ? 1 : (b <= NumBndin1 + NumBndin2 ? 2 * b - NumBndin1 - 1
                                                                         : 3 * b - NumBndin2 - NumBndin1 * 2 - 2);
    const double ionf[3] = {ion_freq_HI, ion_freq_HeI, ion_freq_HeII};
    const double w = G.delta_freq[b];  // set_integration_weights :789-795
    for (int sidx = 0; sidx < 3; sidx++) {
      if (!active[sidx]) continue;
      SedTables& S = *seds[sidx];
      for (int it = 0; it <= NumTau; it++) {
        double thick[NumFreq + 1], thin[NumFreq + 1];
        // fill_photo_integrands :593-660
        for (int i = 0; i <= NumFreq; i++) {
          if (tau[it] * csfd[i] < F(700.0f)) {
            if (sidx == 0) {
              if (frequency[i] * G.h_over_kT < F(700.0f)) {
                thick[i] = 4.0 * pi * G.R_star2 * two_pi_over_c_square * frequency[i] * frequency[i] *
                           exp(-tau[it] * csfd[i]) / (exp(frequency[i] * G.h_over_kT) - F(1.0f));
                thin[i] = 4.0 * pi * G.R_star2 * two_pi_over_c_square * frequency[i] * frequency[i] * csfd[i] *
                          exp(-tau[it] * csfd[i]) / (exp(frequency[i] * G.h_over_kT) - F(1.0f));
              } else { thick[i] = 0.0; thin[i] = 0.0; }
            } else {
              double scal = (sidx == 1) ? G.pl_scaling : G.qpl_scaling;
              double pidx = (sidx == 1) ? G.pl_index : G.qpl_index;
              thick[i] = scal * pow(frequency[i], -pidx) * exp(-tau[it] * csfd[i]);
              thin[i] = scal * pow(frequency[i], -pidx) * csfd[i] * exp(-tau[it] * csfd[i]);
            }
          } else { thick[i] = 0.0; thin[i] = 0.0; }
        }
        // make_photo_tables :798-822 -> vector_romberg (romberg.f90:158-188)
        double a1 = 0.0, a2 = 0.0;
        for (int x = 0; x <= NumFreq; x++) { a1 = a1 + thick[x] * w * G.romw[x]; a2 = a2 + thin[x] * w * G.romw[x]; }
        S.photo_thick[(size_t)(b - 1) * (NumTau + 1) + it] = a1;
        S.photo_thin[(size_t)(b - 1) * (NumTau + 1) + it] = a2;
        if (!G.isothermal) {
          // fill_heating_integrands_* :664-783, make_heat_tables_* :825-899
          for (int sp = 0; sp < nsp; sp++) {
            double h1 = 0.0, h2 = 0.0;
            for (int x = 0; x <= NumFreq; x++) {
              double ft = hplanck * (frequency[x] - ionf[sp]) * thick[x];
              double fn = hplanck * (frequency[x] - ionf[sp]) * thin[x];
              h1 = h1 + ft * w * G.romw[x];
              h2 = h2 + fn * w * G.romw[x];
            }
            S.heat_thick[(size_t)(hcol0 + sp - 1) * (NumTau + 1) + it] = h1;
            S.heat_thin[(size_t)(hcol0 + sp - 1) * (NumTau + 1) + it] = h2;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// radiation_photoionrates.f90
// ---------------------------------------------------------------------------------------------
struct TablePos {  // :91-97
  double tau[NumFreqBnd + 1], odpos[NumFreqBnd + 1], residual[NumFreqBnd + 1];
  int ipos[NumFreqBnd + 1], ipos_p1[NumFreqBnd + 1];
};
void set_tau_table_positions(const double* tau, TablePos& p) {  // :282-306
  for (int b = 1; b <= NumFreqBnd; b++) {
    p.tau[b] = log10(std::max(1.0e-20, tau[b]));
    p.odpos[b] = std::min((double)NumTau, std::max(0.0, F(1.0f) + (p.tau[b] - minlogtau) / dlogtau));
    p.ipos[b] = (int)p.odpos[b];
    p.residual[b] = p.odpos[b] - (double)p.ipos[b];
    p.ipos_p1[b] = std::min(NumTau, p.ipos[b] + 1);
  }
}
inline double read_table(const std::vector<double>& table, const TablePos& p, int b, int col) {  // :310-326
  const double* t = &table[(size_t)(col - 1) * (NumTau + 1)];
  return t[p.ipos[b]] + (t[p.ipos_p1[b]] - t[p.ipos[b]]) * p.residual[b];
}
const double CR1[4] = {0, 0.3908, 0.0554, 1.0}, bR1[4] = {0, 0.4092, 0.4614, 0.2663}, dR1[4] = {0, 1.7592, 1.6660, 1.3163};
const double CR2[4] = {0, 0.6941, 0.0984, 3.9811}, aR2[4] = {0, 0.2, 0.2, 0.4}, bR2[4] = {0, 0.38, 0.38, 0.34};

PhotRates photo_lookuptable(const TablePos& pin, const TablePos& pout, const double* tau_in, const double* tau_out,
                            double NFlux, const SedTables& T, double vol, const double* sHI, const double* sHeI,
                            const double* sHeII) {  // :331-462
  const double tau_photo_limit = F(1.0e-7f);
  PhotRates r;
  for (int b = T.lo; b <= T.hi; b++) {
    double phi_photo_in_all = NFlux * read_table(T.photo_thick, pin, b, b);
    r.photo_in = r.photo_in + phi_photo_in_all;
    double phi_photo_out_all, phi_photo_all;
    if (fabs(tau_out[b] - tau_in[b]) > tau_photo_limit) {
      phi_photo_out_all = NFlux * read_table(T.photo_thick, pout, b, b);
      phi_photo_all = phi_photo_in_all - phi_photo_out_all;
    } else {
      phi_photo_all = NFlux * (tau_out[b] - tau_in[b]) * read_table(T.photo_thin, pin, b, b);
      phi_photo_out_all = phi_photo_in_all - phi_photo_all;
    }
    r.photo_out = r.photo_out + phi_photo_out_all;
    if (b <= NumBndin1) {
      r.photo_cell_HI = r.photo_cell_HI + phi_photo_all / vol;
    } else if (b <= NumBndin1 + NumBndin2) {
      r.photo_cell_HI = r.photo_cell_HI + sHI[b] * phi_photo_all / vol;
      r.photo_cell_HeI = r.photo_cell_HeI + sHeI[b] * phi_photo_all / vol;
    } else {
      r.photo_cell_HI = r.photo_cell_HI + sHI[b] * phi_photo_all / vol;
      r.photo_cell_HeI = r.photo_cell_HeI + sHeI[b] * phi_photo_all / vol;
      r.photo_cell_HeII = r.photo_cell_HeII + sHeII[b] * phi_photo_all / vol;
    }
  }
  return r;
}

PhotRates heat_lookuptable(const TablePos& pin, const TablePos& pout, const double* tau_in, const double* tau_out,
                           const double* tcHI, const double* tcHeI, const double* tcHeII, double NFlux,
                           const SedTables& T, double vol, double i_state, const double* sHI, const double* sHeI,
                           const double* sHeII) {  // :470-779
  const double tau_heat_limit = F(1.0e-4f);
  PhotRates r;
  double f_heat = 0.0, f_ion_HI = 0.0, f_ion_HeI = 0.0;
  double fra_sum1 = 0, fra_sum2 = 0, fra_sum3 = 0, fra_sum4 = 0, df_ion_HI = 0, df_ion_HeI = 0, df_heat = 0;
  double y1R[4], y2R[4];
  for (int i = 1; i <= 3; i++) {  // :557-565
    y1R[i] = CR1[i] * pow(1.0 - pow(i_state, bR1[i]), dR1[i]);
    double xeb = 1.0 - pow(i_state, bR2[i]);
    y2R[i] = CR2[i] * pow(i_state, aR2[i]) * xeb * xeb;
  }
  for (int b = T.lo; b <= T.hi; b++) {
    double phi_heat_HI = 0.0, phi_heat_HeI = 0.0, phi_heat_HeII = 0.0;
    const bool thick = fabs(tau_out[b] - tau_in[b]) > tau_heat_limit;
    if (b <= NumBndin1) {  // :586-613
      double in_HI = NFlux * read_table(T.heat_thick, pin, b, b);
      if (thick) {
        double out_HI = NFlux * read_table(T.heat_thick, pout, b, b);
        phi_heat_HI = (in_HI - out_HI) / vol;
      } else {
        phi_heat_HI = NFlux * tcHI[b] * read_table(T.heat_thin, pin, b, b);
        phi_heat_HI = phi_heat_HI / vol;
      }
      df_heat = phi_heat_HI;
    } else if (b <= NumBndin1 + NumBndin2) {  // :616-676
      const int c1 = 2 * b - NumBndin1 - 1, c2 = 2 * b - NumBndin1;
      double in_HI = NFlux * read_table(T.heat_thick, pin, b, c1);
      double in_HeI = NFlux * read_table(T.heat_thick, pin, b, c2);
      if (thick) {
        double out_HI = NFlux * read_table(T.heat_thick, pout, b, c1);
        phi_heat_HI = sHI[b] * (in_HI - out_HI) / vol;
        double out_HeI = NFlux * read_table(T.heat_thick, pout, b, c2);
        phi_heat_HeI = sHeI[b] * (in_HeI - out_HeI) / vol;
      } else {
        phi_heat_HI = NFlux * tcHI[b] * read_table(T.heat_thin, pin, b, c1);
        phi_heat_HI = phi_heat_HI / vol;
        phi_heat_HeI = NFlux * tcHeI[b] * read_table(T.heat_thin, pin, b, c2);
        phi_heat_HeI = phi_heat_HeI / vol;
      }
      df_heat = phi_heat_HI + phi_heat_HeI;
      fra_sum1 = G.f1ion_HI[b] * phi_heat_HI + G.f1ion_HeI[b] * phi_heat_HeI;
      fra_sum2 = G.f2ion_HI[b] * phi_heat_HI + G.f2ion_HeI[b] * phi_heat_HeI;
      fra_sum3 = G.f1heat_HI[b] * phi_heat_HI + G.f1heat_HeI[b] * phi_heat_HeI;
      fra_sum4 = G.f2heat_HI[b] * phi_heat_HI + G.f2heat_HeI[b] * phi_heat_HeI;
      df_ion_HeI = y1R[2] * fra_sum1 - y2R[2] * fra_sum2;
      df_ion_HI = y1R[1] * fra_sum1 - y2R[1] * fra_sum2;
      df_heat = df_heat - y1R[3] * fra_sum3 + y2R[3] * fra_sum4;
    } else {  // :679-760
      const int c1 = 3 * b - NumBndin2 - NumBndin1 * 2 - 2, c2 = c1 + 1, c3 = c1 + 2;
      double in_HI = NFlux * read_table(T.heat_thick, pin, b, c1);
      double in_HeI = NFlux * read_table(T.heat_thick, pin, b, c2);
      double in_HeII = NFlux * read_table(T.heat_thick, pin, b, c3);
      if (thick) {
        double out_HI = NFlux * read_table(T.heat_thick, pout, b, c1);
        phi_heat_HI = sHI[b] * (in_HI - out_HI) / vol;
        double out_HeI = NFlux * read_table(T.heat_thick, pout, b, c2);
        phi_heat_HeI = sHeI[b] * (in_HeI - out_HeI) / vol;
        double out_HeII = NFlux * read_table(T.heat_thick, pout, b, c3);
        phi_heat_HeII = sHeII[b] * (in_HeII - out_HeII) / vol;
      } else {
        phi_heat_HI = NFlux * tcHI[b] * read_table(T.heat_thin, pin, b, c1);
        phi_heat_HI = phi_heat_HI / vol;
        phi_heat_HeI = NFlux * tcHeI[b] * read_table(T.heat_thin, pin, b, c2);
        phi_heat_HeI = phi_heat_HeI / vol;
        phi_heat_HeII = NFlux * tcHeII[b] * read_table(T.heat_thin, pin, b, c3);
        phi_heat_HeII = phi_heat_HeII / vol;
      }
      df_heat = phi_heat_HI + phi_heat_HeI + phi_heat_HeII;
      fra_sum1 = G.f1ion_HI[b] * phi_heat_HI + G.f1ion_HeI[b] * phi_heat_HeI + G.f1ion_HeII[b] * phi_heat_HeII;
      fra_sum2 = G.f2ion_HI[b] * phi_heat_HI + G.f2ion_HeI[b] * phi_heat_HeI + G.f2ion_HeII[b] * phi_heat_HeII;
      fra_sum3 = G.f1heat_HI[b] * phi_heat_HI + G.f1heat_HeI[b] * phi_heat_HeI + G.f1heat_HeII[b] * phi_heat_HeII;
      fra_sum4 = G.f2heat_HI[b] * phi_heat_HI + G.f2heat_HeI[b] * phi_heat_HeI + G.f2heat_HeII[b] * phi_heat_HeII;
      df_ion_HeI = y1R[2] * fra_sum1 - y2R[2] * fra_sum2;
      df_ion_HI = y1R[1] * fra_sum1 - y2R[1] * fra_sum2;
      df_heat = df_heat - y1R[3] * fra_sum3 + y2R[3] * fra_sum4;
    }
    f_heat = f_heat + df_heat;
    f_ion_HI = f_ion_HI + df_ion_HI;
    f_ion_HeI = f_ion_HeI + df_ion_HeI;
  }
  r.heat = f_heat;
  r.photo_cell_HI = f_ion_HI / (ion_freq_HI * hplanck);
  r.photo_cell_HeI = f_ion_HeI / (ion_freq_HeI * hplanck);
  return r;
}

// :108-277.  nflux = {NormFlux(ns), NormFluxPL(ns), NormFluxQPL(ns)}
PhotRates photoion_rates(double colum_in_HI, double colum_out_HI, double colum_in_HeI, double colum_out_HeI,
                         double colum_in_HeII, double colum_out_HeII, double vol, const double* nflux,
                         double i_state) {
  PhotRates phi;
  double colum_cell_HI = colum_out_HI - colum_in_HI;
  double colum_cell_HeI = colum_out_HeI - colum_in_HeI;
  double colum_cell_HeII = colum_out_HeII - colum_in_HeII;
  double tau_in_all[NumFreqBnd + 1], tau_out_all[NumFreqBnd + 1];
  double tcHI[NumFreqBnd + 1], tcHeI[NumFreqBnd + 1], tcHeII[NumFreqBnd + 1];
  double sHI[NumFreqBnd + 1], sHeI[NumFreqBnd + 1], sHeII[NumFreqBnd + 1];
  for (int b = 1; b <= NumFreqBnd; b++)
    tau_in_all[b] = colum_in_HI * G.sigma_HI[b] + colum_in_HeI * G.sigma_HeI[b] + colum_in_HeII * G.sigma_HeII[b];
  for (int b = 1; b <= NumFreqBnd; b++)
    tau_out_all[b] = colum_out_HI * G.sigma_HI[b] + colum_out_HeI * G.sigma_HeI[b] + colum_out_HeII * G.sigma_HeII[b];
  TablePos pin, pout;
  set_tau_table_positions(tau_in_all, pin);
  set_tau_table_positions(tau_out_all, pout);
  for (int b = NumBndin1 + 1; b <= NumBndin1 + NumBndin2; b++) {  // scale_int2 :787-802
    double forscaleing = 1.0 / (G.sigma_HI[b] * colum_cell_HI + G.sigma_HeI[b] * colum_cell_HeI);
    sHI[b] = G.sigma_HI[b] * colum_cell_HI * forscaleing;
    sHeI[b] = G.sigma_HeI[b] * colum_cell_HeI * forscaleing;
  }
  for (int b = NumBndin1 + NumBndin2 + 1; b <= NumFreqBnd; b++) {  // scale_int3 :808-825
    double forscaleing =
        1.0 / (G.sigma_HI[b] * colum_cell_HI + G.sigma_HeI[b] * colum_cell_HeI + G.sigma_HeII[b] * colum_cell_HeII);
    sHI[b] = colum_cell_HI * G.sigma_HI[b] * forscaleing;
    sHeI[b] = colum_cell_HeI * G.sigma_HeI[b] * forscaleing;
    sHeII[b] = colum_cell_HeII * G.sigma_HeII[b] * forscaleing;
  }
  if (nflux[0] > 0.0)
    photrates_add(phi, photo_lookuptable(pin, pout, tau_in_all, tau_out_all, nflux[0], G.bb, vol, sHI, sHeI, sHeII));
  if (G.use_pl && nflux[1] > 0.0)
    photrates_add(phi, photo_lookuptable(pin, pout, tau_in_all, tau_out_all, nflux[1], G.pl, vol, sHI, sHeI, sHeII));
  if (G.use_qpl && nflux[2] > 0.0)
    photrates_add(phi, photo_lookuptable(pin, pout, tau_in_all, tau_out_all, nflux[2], G.qpl, vol, sHI, sHeI, sHeII));
  if (!G.isothermal) {
    for (int b = 1; b <= NumFreqBnd; b++) {
      tcHI[b] = colum_cell_HI * G.sigma_HI[b];
      tcHeI[b] = colum_cell_HeI * G.sigma_HeI[b];
      tcHeII[b] = colum_cell_HeII * G.sigma_HeII[b];
    }
    if (nflux[0] > 0.0)
      photrates_add(phi, heat_lookuptable(pin, pout, tau_in_all, tau_out_all, tcHI, tcHeI, tcHeII, nflux[0], G.bb, vol,
                                          i_state, sHI, sHeI, sHeII));
    if (G.use_pl && nflux[1] > 0.0)
      photrates_add(phi, heat_lookuptable(pin, pout, tau_in_all, tau_out_all, tcHI, tcHeI, tcHeII, nflux[1], G.pl, vol,
                                          i_state, sHI, sHeI, sHeII));
    if (G.use_qpl && nflux[2] > 0.0)
      photrates_add(phi, heat_lookuptable(pin, pout, tau_in_all, tau_out_all, tcHI, tcHeI, tcHeII, nflux[2], G.qpl,
                                          vol, i_state, sHI, sHeI, sHeII));
  }
  return phi;
}

// ---------------------------------------------------------------------------------------------
// column_density.f90:351-376 weightf, :28-345 cinterp
// ---------------------------------------------------------------------------------------------
inline double weightf(double cd, int id) {
  double sig = (id == 0) ? sigma_HI_at_ion_freq : (id == 1 ? sigma_HeI_at_ion_freq : sigma_HeII_at_ion_freq);
  return F(1.0f) / std::max(0.6, cd * sig);
}
inline int fmodulo(int a, int n) { int m = a % n; return m < 0 ? m + n : m; }   // Fortran modulo()
inline int isign1(int b) { return b >= 0 ? 1 : -1; }                             // sign(1,b)

// Per-worker (thread / MPI rank analogue) scratch: evolve_data.F90 coldensh_out, coldenshe_out + private rate grids
struct Worker {
  std::vector<double> cdh, cdhe0, cdhe1;
  double *phih, *phihe0, *phihe1, *phiheat;
  std::vector<double> own_rates;
  double photon_loss = 0.0;
  double photon_loss_src_thread = 0.0;
  long sum_nbox = 0;
  long updates = 0;
  int last_l[3], last_r[3];
};

inline size_t cidx(int i, int j, int k) {  // 1-based wrapped indices -> flat
  return (size_t)(i - 1) + (size_t)G.mesh[0] * ((size_t)(j - 1) + (size_t)G.mesh[1] * (size_t)(k - 1));
}

void cinterp(const Worker& W, const int* pos, const int* srcpos, double& cdensi, double& cdensihe0, double& cdensihe1,
             double& path) {
  const double sqrt3 = F(sqrtf(3.0f)), sqrt2 = F(sqrtf(2.0f));
  const int i = pos[0], j = pos[1], k = pos[2], i0 = srcpos[0], j0 = srcpos[1], k0 = srcpos[2];
  const int idel = i - i0, jdel = j - j0, kdel = k - k0;
  const int idela = abs(idel), jdela = abs(jdel), kdela = abs(kdel);
  const int sgni = isign1(idel), sgnj = isign1(jdel), sgnk = isign1(kdel);
  const int im = i - sgni, jm = j - sgnj, km = k - sgnk;
  const double di = (double)(float)idel, dj = (double)(float)jdel, dk = (double)(float)kdel;
  const int* mesh = G.mesh;
  double c1, c2, c3, c4, c1he0, c2he0, c3he0, c4he0, c1he1, c2he1, c3he1, c4he1, s1, s2, s3, s4;
  bool diag3, diag2;
  const double one = F(1.f), two = F(2.0f);
  if (kdela >= jdela && kdela >= idela) {
    double alam = (double)((float)(km - k0) + (float)sgnk * 0.5f) / dk;
    double xc = alam * di + (double)(float)i0;
    double yc = alam * dj + (double)(float)j0;
    double dx = two * fabs(xc - (double)((float)im + 0.5f * (float)sgni));
    double dy = two * fabs(yc - (double)((float)jm + 0.5f * (float)sgnj));
    s1 = (one - dx) * (one - dy); s2 = (one - dy) * dx; s3 = (one - dx) * dy; s4 = dx * dy;
    int ip = fmodulo(i - 1, mesh[0]) + 1, imp = fmodulo(im - 1, mesh[0]) + 1;
    int jp = fmodulo(j - 1, mesh[1]) + 1, jmp = fmodulo(jm - 1, mesh[1]) + 1;
    int kmp = fmodulo(km - 1, mesh[2]) + 1;
    size_t a1 = cidx(imp, jmp, kmp), a2 = cidx(ip, jmp, kmp), a3 = cidx(imp, jp, kmp), a4 = cidx(ip, jp, kmp);
    c1 = W.cdh[a1]; c2 = W.cdh[a2]; c3 = W.cdh[a3]; c4 = W.cdh[a4];
    c1he0 = W.cdhe0[a1]; c2he0 = W.cdhe0[a2]; c3he0 = W.cdhe0[a3]; c4he0 = W.cdhe0[a4];
    c1he1 = W.cdhe1[a1]; c2he1 = W.cdhe1[a2]; c3he1 = W.cdhe1[a3]; c4he1 = W.cdhe1[a4];
    diag2 = (kdela == 1 && (idela == 1 || jdela == 1));
    diag3 = (idela == 1 && jdela == 1);
    path = sqrt((di * di + dj * dj) / (dk * dk) + one);
  } else if (jdela >= idela && jdela >= kdela) {
    double alam = (double)((float)(jm - j0) + (float)sgnj * 0.5f) / dj;
    double zc = alam * dk + (double)(float)k0;
    double xc = alam * di + (double)(float)i0;
    double dz = two * fabs(zc - (double)((float)km + 0.5f * (float)sgnk));
    double dx = two * fabs(xc - (double)((float)im + 0.5f * (float)sgni));
    s1 = (one - dx) * (one - dz); s2 = (one - dz) * dx; s3 = (one - dx) * dz; s4 = dx * dz;
    int ip = fmodulo(i - 1, mesh[0]) + 1, imp = fmodulo(im - 1, mesh[0]) + 1;
    int jmp = fmodulo(jm - 1, mesh[1]) + 1;
    int kp = fmodulo(k - 1, mesh[2]) + 1, kmp = fmodulo(km - 1, mesh[2]) + 1;
    size_t a1 = cidx(imp, jmp, kmp), a2 = cidx(ip, jmp, kmp), a3 = cidx(imp, jmp, kp), a4 = cidx(ip, jmp, kp);
    c1 = W.cdh[a1]; c2 = W.cdh[a2]; c3 = W.cdh[a3]; c4 = W.cdh[a4];
    c1he0 = W.cdhe0[a1]; c2he0 = W.cdhe0[a2]; c3he0 = W.cdhe0[a3]; c4he0 = W.cdhe0[a4];
    c1he1 = W.cdhe1[a1]; c2he1 = W.cdhe1[a2]; c3he1 = W.cdhe1[a3]; c4he1 = W.cdhe1[a4];
    diag2 = (jdela == 1 && (idela == 1 || kdela == 1));
    diag3 = (idela == 1 && kdela == 1);
    path = sqrt((di * di + dk * dk) / (dj * dj) + one);
  } else {
    double alam = (double)((float)(im - i0) + (float)sgni * 0.5f) / di;
    double zc = alam * dk + (double)(float)k0;
    double yc = alam * dj + (double)(float)j0;
    double dz = two * fabs(zc - (double)((float)km + 0.5f * (float)sgnk));
    double dy = two * fabs(yc - (double)((float)jm + 0.5f * (float)sgnj));
    s1 = (one - dz) * (one - dy); s2 = (one - dz) * dy; s3 = (one - dy) * dz; s4 = dy * dz;
    int imp = fmodulo(im - 1, mesh[0]) + 1;
    int jp = fmodulo(j - 1, mesh[1]) + 1, jmp = fmodulo(jm - 1, mesh[1]) + 1;
    int kp = fmodulo(k - 1, mesh[2]) + 1, kmp = fmodulo(km - 1, mesh[2]) + 1;
    size_t a1 = cidx(imp, jmp, kmp), a2 = cidx(imp, jp, kmp), a3 = cidx(imp, jmp, kp), a4 = cidx(imp, jp, kp);
    c1 = W.cdh[a1]; c2 = W.cdh[a2]; c3 = W.cdh[a3]; c4 = W.cdh[a4];
    c1he0 = W.cdhe0[a1]; c2he0 = W.cdhe0[a2]; c3he0 = W.cdhe0[a3]; c4he0 = W.cdhe0[a4];
    c1he1 = W.cdhe1[a1]; c2he1 = W.cdhe1[a2]; c3he1 = W.cdhe1[a3]; c4he1 = W.cdhe1[a4];
    diag2 = (idela == 1 && (jdela == 1 || kdela == 1));
    diag3 = (jdela == 1 && kdela == 1);
    path = sqrt(one + (dj * dj + dk * dk) / (di * di));
  }
  double w1 = s1 * weightf(c1, 0), w2 = s2 * weightf(c2, 0), w3 = s3 * weightf(c3, 0), w4 = s4 * weightf(c4, 0);
  double w1he0 = s1 * weightf(c1he0, 1), w2he0 = s2 * weightf(c2he0, 1), w3he0 = s3 * weightf(c3he0, 1),
         w4he0 = s4 * weightf(c4he0, 1);
  double w1he1 = s1 * weightf(c1he1, 2), w2he1 = s2 * weightf(c2he1, 2), w3he1 = s3 * weightf(c3he1, 2),
         w4he1 = s4 * weightf(c4he1, 2);
  cdensi = (c1 * w1 + c2 * w2 + c3 * w3 + c4 * w4) / (w1 + w2 + w3 + w4);
  cdensihe0 = (c1he0 * w1he0 + c2he0 * w2he0 + c3he0 * w3he0 + c4he0 * w4he0) / (w1he0 + w2he0 + w3he0 + w4he0);
  cdensihe1 = (c1he1 * w1he1 + c2he1 * w2he1 + c3he1 * w3he1 + c4he1 * w4he1) / (w1he1 + w2he1 + w3he1 + w4he1);
  if (diag2) {
    if (diag3) { cdensi = sqrt3 * cdensi; cdensihe0 = sqrt3 * cdensihe0; cdensihe1 = sqrt3 * cdensihe1; }
    else { cdensi = sqrt2 * cdensi; cdensihe0 = sqrt2 * cdensihe0; cdensihe1 = sqrt2 * cdensihe1; }
  }
}

// ---------------------------------------------------------------------------------------------
// evolve_point.F90:79-319 evolve0D  (niter /= -1 path; use_LLS=.false.)
// ---------------------------------------------------------------------------------------------
// loss_out / upd_out (parallel shell traversal only): the cell's photon-loss contribution and update count are handed
// back instead of being accumulated in W, so that the caller can sum them in a fixed order.
void evolve0D(Worker& W, const int* rtpos, int ns, double* loss_out = nullptr, long* upd_out = nullptr) {
  const double max_coldensh = F(2e29f);
  const int* mesh = G.mesh;
  const size_t N3 = ncell();
  int pos[3];
  for (int d = 0; d < 3; d++) pos[d] = fmodulo(rtpos[d] - 1, mesh[d]) + 1;
  const size_t p = cidx(pos[0], pos[1], pos[2]);
  if (W.cdh[p] == 0.0) {
    const double h_av0 = std::max(G.xh_av[p], epsilon), h_av1 = std::max(G.xh_av[p + N3], epsilon);
    const double he_av0 = std::max(G.xhe_av[p], epsilon), he_av1 = std::max(G.xhe_av[p + N3], epsilon);
    const double ndens_p = G.ndens[p];
    const int* sp = &G.srcpos[3 * (size_t)(ns - 1)];
    double coldensh_in, coldenshe_in[2], path, vol_ph;
    if (rtpos[0] == sp[0] && rtpos[1] == sp[1] && rtpos[2] == sp[2]) {
      coldensh_in = 0.0; coldenshe_in[0] = 0.0; coldenshe_in[1] = 0.0;
      path = F(0.5f) * G.dr[0];
      vol_ph = G.dr[0] * G.dr[1] * G.dr[2];
    } else {
      cinterp(W, rtpos, sp, coldensh_in, coldenshe_in[0], coldenshe_in[1], path);
      path = path * G.dr[0];
      double xs = G.dr[0] * (double)(float)(rtpos[0] - sp[0]);
      double ys = G.dr[1] * (double)(float)(rtpos[1] - sp[1]);
      double zs = G.dr[2] * (double)(float)(rtpos[2] - sp[2]);
      double dist2 = xs * xs + ys * ys + zs * zs;
      vol_ph = F(4.0f) * pi * dist2 * path;
      if (G.type_of_LLS != 0) {  // evolve_point.F90:177-180 (LLS_point: coldensh_LLS = LLS_grid(i,j,k))
        const double coldensh_LLS = G.type_of_LLS == 2 ? (double)G.LLS_grid[p] : G.coldensh_LLS;
        coldensh_in = coldensh_in + coldensh_LLS * path / G.dr[0];
      }
    }
    W.cdh[p] = coldensh_in + coldens(path, h_av0, ndens_p, (1.0 - abu_he));
    W.cdhe0[p] = coldenshe_in[0] + coldens(path, he_av0, ndens_p, abu_he);
    W.cdhe1[p] = coldenshe_in[1] + coldens(path, he_av1, ndens_p, abu_he);
    PhotRates phi;
    if (coldensh_in < max_coldensh) {
      double nflux[3] = {G.NormFlux[ns - 1], G.use_pl ? G.NormFluxPL[ns - 1] : 0.0, G.use_qpl ? G.NormFluxQPL[ns - 1] : 0.0};
      phi = photoion_rates(coldensh_in, W.cdh[p], coldenshe_in[0], W.cdhe0[p], coldenshe_in[1], W.cdhe1[p], vol_ph,
                           nflux, h_av1);
      phi.photo_cell_HI = phi.photo_cell_HI / (h_av0 * ndens_p * (1.0 - abu_he));
      phi.photo_cell_HeI = phi.photo_cell_HeI / (he_av0 * ndens_p * abu_he);
      phi.photo_cell_HeII = phi.photo_cell_HeII / (he_av1 * ndens_p * abu_he);
    }
    W.phih[p] = W.phih[p] + phi.photo_cell_HI;
    W.phihe0[p] = W.phihe0[p] + phi.photo_cell_HeI;
    W.phihe1[p] = W.phihe1[p] + phi.photo_cell_HeII;
    if (!G.isothermal) W.phiheat[p] = W.phiheat[p] + phi.heat;
    bool on_l = rtpos[0] == W.last_l[0] || rtpos[1] == W.last_l[1] || rtpos[2] == W.last_l[2];
    bool on_r = rtpos[0] == W.last_r[0] || rtpos[1] == W.last_r[1] || rtpos[2] == W.last_r[2];
    if (loss_out) {
      if (on_l || on_r) *loss_out = phi.photo_out * G.vol / vol_ph;
      *upd_out = 1;
    } else {
      if (on_l || on_r) W.photon_loss_src_thread = W.photon_loss_src_thread + phi.photo_out * G.vol / vol_ph;
      W.updates++;
    }
  }
}

// evolve_source.F90:244-284 evolve2D
void evolve2D(Worker& W, int* rtpos, int ns) {
  const int* sp = &G.srcpos[3 * (size_t)(ns - 1)];
  for (int j = sp[1]; j <= W.last_r[1]; j++) {
    rtpos[1] = j;
    for (int i = sp[0]; i <= W.last_r[0]; i++) { rtpos[0] = i; evolve0D(W, rtpos, ns); }
    for (int i = sp[0] - 1; i >= W.last_l[0]; i--) { rtpos[0] = i; evolve0D(W, rtpos, ns); }
  }
  for (int j = sp[1] - 1; j >= W.last_l[1]; j--) {
    rtpos[1] = j;
    for (int i = sp[0]; i <= W.last_r[0]; i++) { rtpos[0] = i; evolve0D(W, rtpos, ns); }
    for (int i = sp[0] - 1; i >= W.last_l[0]; i--) { rtpos[0] = i; evolve0D(W, rtpos, ns); }
  }
}

// Offsets of max-norm shell r in ascending (dk, dj, di) order, O(r^2) per shell.
template <class Fn>
inline void for_each_shell_offset(int r, Fn fn) {
  for (int dk = -r; dk <= r; dk++) {
    if (abs(dk) == r) {
      for (int dj = -r; dj <= r; dj++)
        for (int di = -r; di <= r; di++) fn(di, dj, dk);
    } else {
      for (int dj = -r; dj <= r; dj++) {
        if (abs(dj) == r) {
          for (int di = -r; di <= r; di++) fn(di, dj, dk);
        } else {
          fn(-r, dj, dk);
          fn(r, dj, dk);  // r > 0 here (|dk| < r)
        }
      }
    }
  }
}

// Shell-order traversal of the current sub-box (the wavefront order the GPU uses); SURVEY H2.
void sweep_box_shell_order(Worker& W, int ns) {
  const int* sp = &G.srcpos[3 * (size_t)(ns - 1)];
  int rmax = 0;
  for (int d = 0; d < 3; d++) rmax = std::max(rmax, std::max(sp[d] - W.last_l[d], W.last_r[d] - sp[d]));
  for (int r = 0; r <= rmax; r++)
    for_each_shell_offset(r, [&](int di, int dj, int dk) {
      int rtpos[3] = {sp[0] + di, sp[1] + dj, sp[2] + dk};
      bool in = true;
      for (int d = 0; d < 3; d++) in = in && rtpos[d] >= W.last_l[d] && rtpos[d] <= W.last_r[d];
      if (in) evolve0D(W, rtpos, ns);
    });
}

// The same shell-order traversal with the cells of one shell spread over `nthreads` OpenMP threads (cells of a shell
// are independent, SURVEY H2: same-shell corners enter cinterp with weight exactly 0).  Per-cell results are
// bitwise those of sweep_box_shell_order; the photon loss is summed in the same cell order afterwards.  Only there to
// make single-source full-size parity runs (BASELINE configs[0]) finish in seconds.
void sweep_box_shell_parallel(Worker& W, int ns, int nthreads) {
  const int* sp = &G.srcpos[3 * (size_t)(ns - 1)];
  int rmax = 0;
  for (int d = 0; d < 3; d++) rmax = std::max(rmax, std::max(sp[d] - W.last_l[d], W.last_r[d] - sp[d]));
  std::vector<int> cells;
  std::vector<double> loss;
  std::vector<long> upd;
  for (int r = 0; r <= rmax; r++) {
    cells.clear();
    for_each_shell_offset(r, [&](int di, int dj, int dk) {
      const int rt[3] = {sp[0] + di, sp[1] + dj, sp[2] + dk};
      bool in = true;
      for (int d = 0; d < 3; d++) in = in && rt[d] >= W.last_l[d] && rt[d] <= W.last_r[d];
      if (in) { cells.push_back(rt[0]); cells.push_back(rt[1]); cells.push_back(rt[2]); }
    });
    const long n = (long)cells.size() / 3;
    loss.assign(n, 0.0); upd.assign(n, 0);
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (long q = 0; q < n; q++) evolve0D(W, &cells[3 * q], ns, &loss[q], &upd[q]);
    for (long q = 0; q < n; q++) {
      if (loss[q] != 0.0) W.photon_loss_src_thread = W.photon_loss_src_thread + loss[q];
      W.updates += upd[q];
    }
  }
}

// evolve_source.F90:66-238 do_source (serial branch, periodic_bc=.true.)
// shell_order: 0 reference serial order, 1 shell order, 2 shell order with `inner_threads` threads per shell
int do_source(Worker& W, int ns, int shell_order, int inner_threads = 1) {
  const int* mesh = G.mesh;
  const int* sp = &G.srcpos[3 * (size_t)(ns - 1)];
  std::fill(W.cdh.begin(), W.cdh.end(), 0.0);
  std::fill(W.cdhe0.begin(), W.cdhe0.end(), 0.0);
  std::fill(W.cdhe1.begin(), W.cdhe1.end(), 0.0);
  int lastpos_r[3], lastpos_l[3];
  for (int d = 0; d < 3; d++) {
    lastpos_r[d] = sp[d] + std::min(G.max_subbox, mesh[d] / 2 - 1 + mesh[d] % 2);
    lastpos_l[d] = sp[d] - std::min(G.max_subbox, mesh[d] / 2);
  }
  int nbox = 0;
  double total_source_flux = G.NormFlux[ns - 1] * G.S_star;
  if (G.use_pl) total_source_flux = total_source_flux + G.NormFluxPL[ns - 1] * G.pl_S_star;
  if (G.use_qpl) total_source_flux = total_source_flux + G.NormFluxQPL[ns - 1] * G.qpl_S_star;
  double photon_loss_src = total_source_flux;
  for (int d = 0; d < 3; d++) { W.last_r[d] = sp[d]; W.last_l[d] = sp[d]; }
  while (photon_loss_src > F(1e-10f) * total_source_flux && W.last_r[2] < lastpos_r[2] && W.last_l[2] > lastpos_l[2]) {
    nbox = nbox + 1;
    photon_loss_src = 0.0;
    W.photon_loss_src_thread = 0.0;
    for (int d = 0; d < 3; d++) {
      W.last_r[d] = std::min(sp[d] + G.subboxsize * nbox, lastpos_r[d]);
      W.last_l[d] = std::max(sp[d] - G.subboxsize * nbox, lastpos_l[d]);
    }
    if (shell_order == 2) {
      sweep_box_shell_parallel(W, ns, inner_threads);
    } else if (shell_order == 1) {
      sweep_box_shell_order(W, ns);
    } else {
      int rtpos[3];
      for (int k = sp[2]; k <= W.last_r[2]; k++) { rtpos[2] = k; evolve2D(W, rtpos, ns); }
      for (int k = sp[2] - 1; k >= W.last_l[2]; k--) { rtpos[2] = k; evolve2D(W, rtpos, ns); }
    }
    photon_loss_src = W.photon_loss_src_thread;
  }
  W.photon_loss = W.photon_loss + photon_loss_src;
  W.sum_nbox = W.sum_nbox + nbox;
  return nbox;
}

// mat_ini_test.F90:469-502 get/set_temperature_point (temperature_grid is real(kind=si))
inline void get_temperature_point(size_t p, double& temper_inter, double& av_temper, double& temper) {
  if (G.isothermal) { temper = G.temper_val; av_temper = G.temper_val; temper_inter = G.temper_val; }
  else {
    const size_t N3 = ncell();
    temper_inter = (double)G.temperature_grid[p];
    av_temper = (double)G.temperature_grid[p + N3];
    temper = (double)G.temperature_grid[p + 2 * N3];
  }
}

// evolve_point.F90:444-646 do_chemistry (local=.false. branch)
int do_chemistry(double dt, double ndens_p, IonStates& ion, const PhotRates& phi, double temper1_in,
                 double& avg_temper, double& temper1_out, RecCol& rc, long* therm_sub = nullptr,
                 double clumping = -1.0) {
  // clumping: the material module's scalar, which :484 (clumping_point) overwrites with clumping_grid(i,j,k) when
  // type_of_clumping == 5; passed in rather than stored so that cells can run on several threads
  if (clumping < 0.0) clumping = (double)G.clumping;
  const double path = 1.0;
  double temper1 = temper1_in, temper0 = temper1, temper2;
  int nit = 0;
  for (;;) {
    nit = nit + 1;
    temper2 = temper1;
    const double yh0_av_old = ion.h_av[0], yhe0_av_old = ion.he_av[0], yhe2_av_old = ion.he_av[2];
    double de = electrondens(ndens_p, ion.h_av, ion.he_av);
    if (!G.isothermal) ini_rec_colion_factors(avg_temper, rc);
    double coldensh_cell = coldens(path, ion.h[0], ndens_p, (1.0 - abu_he));
    double coldenshe_cell[2] = {coldens(path, ion.he[0], ndens_p, abu_he), coldens(path, ion.he[1], ndens_p, abu_he)};
    double yfrac, zfrac, y2afrac, y2bfrac;
    prepare_doric_factors(coldensh_cell, coldenshe_cell, yfrac, zfrac, y2afrac, y2bfrac);
    doric(dt, de, ndens_p, ion, phi, yfrac, zfrac, y2afrac, y2bfrac, rc, clumping);
    de = electrondens(ndens_p, ion.h_av, ion.he_av);
    coldensh_cell = coldens(path, ion.h[0], ndens_p, (1.0 - abu_he));
    coldenshe_cell[0] = coldens(path, ion.he[0], ndens_p, abu_he);
    coldenshe_cell[1] = coldens(path, ion.he[1], ndens_p, abu_he);
    prepare_doric_factors(coldensh_cell, coldenshe_cell, yfrac, zfrac, y2afrac, y2bfrac);
    const double ionh0old = ion.h[0], ionh1old = ion.h[1], ionhe0old = ion.he[0], ionhe1old = ion.he[1],
                 ionhe2old = ion.he[2], oldhav = ion.h_av[0], oldhe0av = ion.he_av[0], oldhe1av = ion.he_av[1];
    doric(dt, de, ndens_p, ion, phi, yfrac, zfrac, y2afrac, y2bfrac, rc, clumping);
    ion.h[0] = (ion.h[0] + ionh0old) / 2.0;
    ion.h[1] = (ion.h[1] + ionh1old) / 2.0;
    ion.he[0] = (ion.he[0] + ionhe0old) / 2.0;
    ion.he[1] = (ion.he[1] + ionhe1old) / 2.0;
    ion.he[2] = (ion.he[2] + ionhe2old) / 2.0;
    ion.h_av[0] = (ion.h_av[0] + oldhav) / 2.0;
    ion.he_av[0] = (ion.he_av[0] + oldhe0av) / 2.0;
    ion.he_av[1] = (ion.he_av[1] + oldhe1av) / 2.0;
    de = electrondens(ndens_p, ion.h_av, ion.he_av);
    temper1 = temper0;
    if (!G.isothermal) {
      int nsub = 0;
      thermal(dt, temper1, avg_temper, de, ndens_p, ion, phi, &nsub);
      if (therm_sub) *therm_sub += nsub;
    }
    if ((fabs((ion.h_av[0] - yh0_av_old) / ion.h_av[0]) < minimum_fractional_change ||
         (ion.h_av[0] < minimum_fraction_of_atoms)) &&
        (fabs((ion.he_av[0] - yhe0_av_old) / ion.he_av[0]) < minimum_fractional_change ||
         (ion.he_av[0] < minimum_fraction_of_atoms)) &&
        (fabs((ion.he_av[2] - yhe2_av_old) / ion.he_av[2]) < minimum_fractional_change ||
         (ion.he_av[2] < minimum_fraction_of_atoms)) &&
        (fabs((temper1 - temper2) / temper1) < minimum_fractional_change))
      break;
    if (nit > 400) break;
  }
  temper1_out = temper1;
  return nit;
}

// evolve_point.F90:325-440 evolve0D_global
int evolve0D_global(double dt, size_t p, int& conv_flag, long* therm_sub) {
  const size_t N3 = ncell();
  IonStates ion;
  for (int nx = 0; nx < 2; nx++) {
    ion.h[nx] = std::max(epsilon, G.xh_intermed[p + nx * N3]);
    ion.h_old[nx] = std::max(epsilon, G.xh[p + nx * N3]);
    ion.h_av[nx] = std::max(epsilon, G.xh_av[p + nx * N3]);
  }
  for (int nx = 0; nx < 3; nx++) {
    ion.he[nx] = std::max(epsilon, G.xhe_intermed[p + nx * N3]);
    ion.he_old[nx] = std::max(epsilon, G.xhe[p + nx * N3]);
    ion.he_av[nx] = std::max(epsilon, G.xhe_av[p + nx * N3]);
  }
  const double ndens_p = G.ndens[p];
  double temper_inter, temp_av_old, temper_old;
  get_temperature_point(p, temper_inter, temp_av_old, temper_old);
  PhotRates phi;
  phi.photo_cell_HI = G.phih_grid[p];
  phi.photo_cell_HeI = G.phihe_grid[p];
  phi.photo_cell_HeII = G.phihe_grid[p + N3];
  if (!G.isothermal) phi.heat = G.phiheat[p];
  // do_chemistry :479-481: (temper_inter, avg_temper, temper1) = get_temperature_point
  double avg_temper = temp_av_old, temper1;
  const double clumping = G.use_clumping_grid ? (double)G.clumping_grid[p] : (double)G.clumping;
  int nit = do_chemistry(dt, ndens_p, ion, phi, temper_old, avg_temper, temper1, G.rc, therm_sub, clumping);
  if (!G.isothermal) {  // set_temperature_point :644
    G.temperature_grid[p] = (float)temper1;
    G.temperature_grid[p + N3] = (float)avg_temper;
  }
  const double yh0_av_old = G.xh_av[p], yhe0_av_old = G.xhe_av[p], yhe2_av_old = G.xhe_av[p + 2 * N3];
  double temp_av_new, tdum1, tdum2;
  get_temperature_point(p, tdum1, temp_av_new, tdum2);
  if ((fabs((ion.h_av[0] - yh0_av_old)) > minimum_fractional_change &&
       fabs((ion.h_av[0] - yh0_av_old) / ion.h_av[0]) > minimum_fractional_change &&
       (ion.h_av[0] > minimum_fraction_of_atoms)) ||
      (fabs((ion.he_av[0] - yhe0_av_old)) > minimum_fractional_change &&
       fabs((ion.he_av[0] - yhe0_av_old) / ion.he_av[0]) > minimum_fractional_change &&
       (ion.he_av[0] > minimum_fraction_of_atoms)) ||
      (fabs((ion.he_av[2] - yhe2_av_old)) > minimum_fractional_change &&
       fabs((ion.he_av[2] - yhe2_av_old) / ion.he_av[2]) > minimum_fractional_change &&
       (ion.he_av[2] > minimum_fraction_of_atoms)) ||
      ((fabs((temp_av_old - temp_av_new) / temp_av_new) > 1.0e-1) && (fabs(temp_av_new - temp_av_old) > 100.0))) {
    conv_flag = conv_flag + 1;
  }
  for (int nx = 0; nx < 2; nx++) { G.xh_intermed[p + nx * N3] = ion.h[nx]; G.xh_av[p + nx * N3] = ion.h_av[nx]; }
  for (int nx = 0; nx < 3; nx++) { G.xhe_intermed[p + nx * N3] = ion.he[nx]; G.xhe_av[p + nx * N3] = ion.he_av[nx]; }
  return nit;
}

std::vector<Worker> workers;

void setup_workers(int nw) {
  const size_t N3 = ncell();
  workers.resize(nw);
  for (int w = 0; w < nw; w++) {
    Worker& W = workers[w];
    W.cdh.assign(N3, 0.0); W.cdhe0.assign(N3, 0.0); W.cdhe1.assign(N3, 0.0);
    if (w == 0) {
      W.own_rates.clear();
      W.phih = G.phih_grid.data(); W.phihe0 = G.phihe_grid.data(); W.phihe1 = G.phihe_grid.data() + N3;
      W.phiheat = G.phiheat.data();
    } else {
      W.own_rates.assign(4 * N3, 0.0);
      W.phih = W.own_rates.data(); W.phihe0 = W.phih + N3; W.phihe1 = W.phih + 2 * N3; W.phiheat = W.phih + 3 * N3;
    }
    W.photon_loss = 0; W.sum_nbox = 0; W.updates = 0;
  }
}

}  // namespace

// =====================================================================================
// C entry points (ctypes)
// =====================================================================================
extern "C" {

void orc_set_cooling(const double* logT, const double* h0, const double* h1, const double* he0, const double* he1,
                     const double* he2) {  // cooling_h.f90:76-171
  G.mintemp = logT[0];
  G.dtemp = logT[1] - logT[0];
  for (int i = 0; i < 801; i++) {
    G.h0_cool[i] = pow(10.0, h0[i]); G.h1_cool[i] = pow(10.0, h1[i]); G.he0_cool[i] = pow(10.0, he0[i]);
    G.he1_cool[i] = pow(10.0, he1[i]); G.he2_cool[i] = pow(10.0, he2[i]);
  }
}

// rad_ini (radiation_tables.f90:141-168) with nominal-value SEDs (radiation_sed_parameters.f90:208-244).
// pl/qpl: pass S_star<=0 to disable the SED (the -DPL / -DQUASARS build switches).
void orc_rad_ini(double T_eff, double S_star, double pl_index, double pl_minfreq, double pl_maxfreq, double pl_S_star,
                 double qpl_index, double qpl_minfreq, double qpl_maxfreq, double qpl_S_star, int isothermal) {
  G.isothermal = isothermal != 0;
  G.T_eff = T_eff; G.S_star = S_star; G.R_star = R_SOLAR;
  G.h_over_kT = hplanck / (k_B * T_eff);
  G.use_pl = pl_S_star > 0; G.use_qpl = qpl_S_star > 0;
  G.pl_index = pl_index; G.pl_minfreq = pl_minfreq; G.pl_maxfreq = pl_maxfreq; G.pl_S_star = pl_S_star; G.pl_scaling = 1.0;
  G.qpl_index = qpl_index; G.qpl_minfreq = qpl_minfreq; G.qpl_maxfreq = qpl_maxfreq; G.qpl_S_star = qpl_S_star; G.qpl_scaling = 1.0;
  setup_scalingfactors();
  romberg_initialisation(NumFreq, G.romw);
  // normalize_blackbody :637-675 (S_star given)
  double S_star_unscaled = integrate_sed_S(G.freq_min[1], G.freq_max[NumFreqBnd], 'B');
  double S_scaling = G.S_star / S_star_unscaled;
  G.R_star = sqrt(S_scaling) * G.R_star;
  G.R_star2 = G.R_star * G.R_star;
  if (G.use_pl) G.pl_scaling = G.pl_S_star / integrate_sed_S(G.pl_minfreq, G.pl_maxfreq, 'P');   // :695-699
  if (G.use_qpl) G.qpl_scaling = G.qpl_S_star / integrate_sed_S(G.qpl_minfreq, G.qpl_maxfreq, 'Q');  // :727-731
  spec_integration();
}

void orc_get_sed_info(double* out) {  // R_star2, h_over_kT, pl_scaling, qpl_scaling, band limits
  out[0] = G.R_star2; out[1] = G.h_over_kT; out[2] = G.pl_scaling; out[3] = G.qpl_scaling;
  out[4] = G.bb.lo; out[5] = G.bb.hi; out[6] = G.pl.lo; out[7] = G.pl.hi; out[8] = G.qpl.lo; out[9] = G.qpl.hi;
}
void orc_get_band_data(double* freq_min, double* freq_max, double* sHI, double* sHeI, double* sHeII) {
  for (int b = 1; b <= NumFreqBnd; b++) {
    freq_min[b - 1] = G.freq_min[b]; freq_max[b - 1] = G.freq_max[b];
    sHI[b - 1] = G.sigma_HI[b]; sHeI[b - 1] = G.sigma_HeI[b]; sHeII[b - 1] = G.sigma_HeII[b];
  }
}
void orc_get_romw(double* out) { memcpy(out, G.romw, sizeof(G.romw)); }
// sed: 0=BB 1=PL 2=QPL ; kind: 0 photo_thick 1 photo_thin 2 heat_thick 3 heat_thin ; out is [band][0:NumTau]
void orc_get_table(int sed, int kind, double* out) {
  SedTables& S = sed == 0 ? G.bb : (sed == 1 ? G.pl : G.qpl);
  std::vector<double>& t = kind == 0 ? S.photo_thick : kind == 1 ? S.photo_thin : kind == 2 ? S.heat_thick : S.heat_thin;
  memcpy(out, t.data(), t.size() * sizeof(double));
}

void orc_rec_colion(double T, double* out) {
  RecCol r; ini_rec_colion_factors(T, r);
  double v[12] = {r.arech0, r.brech0, r.areche0, r.breche0, r.oreche0, r.areche1, r.breche1, r.treche1,
                  r.colli_HI, r.colli_HeI, r.colli_HeII, r.v};
  memcpy(out, v, sizeof(v));
}

void orc_set_params(int isothermal, double temper_val, float clumping, double zred, double H0, double Omega0,
                    int cosmological, int subboxsize, int max_subbox) {
  G.isothermal = isothermal != 0; G.temper_val = temper_val; G.clumping = clumping; G.zred = zred; G.H0 = H0;
  G.Omega0 = Omega0; G.cosmological = cosmological != 0; G.subboxsize = subboxsize; G.max_subbox = max_subbox;
  if (G.isothermal) ini_rec_colion_factors(temper_val, G.rc);  // mat_ini_test.F90:168
}

double orc_coolin(double n, double ne, const double* xh, const double* xhe, double T) { return coolin(n, ne, xh, xhe, T); }

// ion15 = h(0:1) he(0:2) h_av(0:1) he_av(0:2) h_old(0:1) he_old(0:2)
static void unpack_ion(const double* v, IonStates& ion) {
  ion.h[0] = v[0]; ion.h[1] = v[1]; ion.he[0] = v[2]; ion.he[1] = v[3]; ion.he[2] = v[4];
  ion.h_av[0] = v[5]; ion.h_av[1] = v[6]; ion.he_av[0] = v[7]; ion.he_av[1] = v[8]; ion.he_av[2] = v[9];
  ion.h_old[0] = v[10]; ion.h_old[1] = v[11]; ion.he_old[0] = v[12]; ion.he_old[1] = v[13]; ion.he_old[2] = v[14];
}
static void pack_ion(const IonStates& ion, double* v) {
  v[0] = ion.h[0]; v[1] = ion.h[1]; v[2] = ion.he[0]; v[3] = ion.he[1]; v[4] = ion.he[2];
  v[5] = ion.h_av[0]; v[6] = ion.h_av[1]; v[7] = ion.he_av[0]; v[8] = ion.he_av[1]; v[9] = ion.he_av[2];
  v[10] = ion.h_old[0]; v[11] = ion.h_old[1]; v[12] = ion.he_old[0]; v[13] = ion.he_old[1]; v[14] = ion.he_old[2];
}

// one doric call at temperature T (coefficients from ini_rec_colion_factors(T))
void orc_doric(double dt, double rhe, double rhh, double* ion15, const double* phi3, const double* fr4, double T) {
  IonStates ion; unpack_ion(ion15, ion);
  PhotRates phi; phi.photo_cell_HI = phi3[0]; phi.photo_cell_HeI = phi3[1]; phi.photo_cell_HeII = phi3[2];
  RecCol rc; ini_rec_colion_factors(T, rc);
  doric(dt, rhe, rhh, ion, phi, fr4[0], fr4[1], fr4[2], fr4[3], rc, (double)G.clumping);
  pack_ion(ion, ion15);
}

void orc_thermal(double dt, double* end_temper, double* avg_temper, double ne, double n, const double* ion15, double heat,
                 int* nsub) {
  IonStates ion; unpack_ion(ion15, ion);
  PhotRates phi; phi.heat = heat;
  thermal(dt, *end_temper, *avg_temper, ne, n, ion, phi, nsub);
}

static int* g_nsub_out = nullptr;
// batch of independent cells through do_chemistry (evolve0D_global without the grid): in/out arrays of n cells.
// ion15 [n][15], phi4 [n][4] (HI,HeI,HeII,heat), T3 [n][3] (inter, avg, old) doubles holding float values
void orc_chemistry_batch(int n, double dt, const double* ndens, double* ion15, const double* phi4, double* T3, int* nit_out) {
  for (int c = 0; c < n; c++) {
    IonStates ion; unpack_ion(ion15 + 15 * (size_t)c, ion);
    PhotRates phi; phi.photo_cell_HI = phi4[4 * c]; phi.photo_cell_HeI = phi4[4 * c + 1]; phi.photo_cell_HeII = phi4[4 * c + 2];
    phi.heat = phi4[4 * c + 3];
    double avg = T3[3 * c + 1], t1;
    RecCol rc = G.rc;
    long nsub = 0;
    nit_out[c] = do_chemistry(dt, ndens[c], ion, phi, T3[3 * c + 2], avg, t1, rc, &nsub);
    if (g_nsub_out) g_nsub_out[c] = (int)nsub;
    T3[3 * c] = t1; T3[3 * c + 1] = avg;
    pack_ion(ion, ion15 + 15 * (size_t)c);
  }
}
// diagnostics: total thermal sub-steps per cell of the next orc_chemistry_batch call (NULL to switch off)
void orc_set_nsub_out(int* p) { g_nsub_out = p; }

// photoion_rates for n independent cells.  col6 [n][6] = in_HI,out_HI,in_HeI,out_HeI,in_HeII,out_HeII ;
// out6 [n][6] = photo_cell_HI, HeI, HeII, heat, photo_in, photo_out
void orc_photoion_rates_batch(int n, const double* col6, const double* vol, const double* nflux3, const double* i_state,
                              double* out6) {
  for (int c = 0; c < n; c++) {
    const double* q = col6 + 6 * (size_t)c;
    PhotRates p = photoion_rates(q[0], q[1], q[2], q[3], q[4], q[5], vol[c], nflux3, i_state[c]);
    double* o = out6 + 6 * (size_t)c;
    o[0] = p.photo_cell_HI; o[1] = p.photo_cell_HeI; o[2] = p.photo_cell_HeII; o[3] = p.heat; o[4] = p.photo_in; o[5] = p.photo_out;
  }
}

// ---- grid state -----------------------------------------------------------------------
void orc_grid_init(const int* mesh, const double* dr, double vol) {
  for (int d = 0; d < 3; d++) { G.mesh[d] = mesh[d]; G.dr[d] = dr[d]; }
  G.vol = vol;
  G.use_clumping_grid = false; G.clumping_grid.clear();
  G.type_of_LLS = 0; G.coldensh_LLS = 0.0; G.LLS_grid.clear();
  const size_t N3 = ncell();
  G.ndens.assign(N3, 0); G.xh.assign(2 * N3, 0); G.xhe.assign(3 * N3, 0);
  G.xh_av.assign(2 * N3, 0); G.xhe_av.assign(3 * N3, 0); G.xh_intermed.assign(2 * N3, 0); G.xhe_intermed.assign(3 * N3, 0);
  G.temperature_grid.assign(3 * N3, 0.f);
  G.phih_grid.assign(N3, 0); G.phihe_grid.assign(2 * N3, 0); G.phiheat.assign(N3, 0);
  workers.clear();
}
void orc_set_geometry(const double* dr, double vol) { for (int d = 0; d < 3; d++) G.dr[d] = dr[d]; G.vol = vol; }
// type_of_clumping == 5: clumping_grid (NULL switches back to the scalar)
void orc_set_clumping_grid(const float* grid) {
  G.use_clumping_grid = grid != nullptr;
  if (grid) G.clumping_grid.assign(grid, grid + ncell()); else G.clumping_grid.clear();
}
// use_LLS: type 0 off, 1 coldensh_LLS for every cell, 2 LLS_grid
void orc_set_LLS(int type_of_LLS, double coldensh_LLS, const float* grid) {
  G.type_of_LLS = type_of_LLS;
  G.coldensh_LLS = type_of_LLS == 1 ? coldensh_LLS : 0.0;
  if (type_of_LLS == 2 && grid) G.LLS_grid.assign(grid, grid + ncell()); else G.LLS_grid.clear();
}
void orc_set_state(const double* ndens, const double* xh, const double* xhe, const float* temperature_grid) {
  const size_t N3 = ncell();
  memcpy(G.ndens.data(), ndens, N3 * 8); memcpy(G.xh.data(), xh, 2 * N3 * 8); memcpy(G.xhe.data(), xhe, 3 * N3 * 8);
  if (temperature_grid) memcpy(G.temperature_grid.data(), temperature_grid, 3 * N3 * 4);
}
void orc_set_work_state(const double* xh_av, const double* xhe_av, const double* xh_intermed, const double* xhe_intermed) {
  const size_t N3 = ncell();
  memcpy(G.xh_av.data(), xh_av, 2 * N3 * 8); memcpy(G.xhe_av.data(), xhe_av, 3 * N3 * 8);
  memcpy(G.xh_intermed.data(), xh_intermed, 2 * N3 * 8); memcpy(G.xhe_intermed.data(), xhe_intermed, 3 * N3 * 8);
}
void orc_set_rates(const double* phih, const double* phihe, const double* phiheat) {
  const size_t N3 = ncell();
  memcpy(G.phih_grid.data(), phih, N3 * 8); memcpy(G.phihe_grid.data(), phihe, 2 * N3 * 8); memcpy(G.phiheat.data(), phiheat, N3 * 8);
}
void orc_get_state(double* xh, double* xhe, float* temperature_grid) {
  const size_t N3 = ncell();
  memcpy(xh, G.xh.data(), 2 * N3 * 8); memcpy(xhe, G.xhe.data(), 3 * N3 * 8);
  if (temperature_grid) memcpy(temperature_grid, G.temperature_grid.data(), 3 * N3 * 4);
}
void orc_get_work_state(double* xh_av, double* xhe_av, double* xh_intermed, double* xhe_intermed) {
  const size_t N3 = ncell();
  memcpy(xh_av, G.xh_av.data(), 2 * N3 * 8); memcpy(xhe_av, G.xhe_av.data(), 3 * N3 * 8);
  memcpy(xh_intermed, G.xh_intermed.data(), 2 * N3 * 8); memcpy(xhe_intermed, G.xhe_intermed.data(), 3 * N3 * 8);
}
void orc_get_rates(double* phih, double* phihe, double* phiheat) {
  const size_t N3 = ncell();
  memcpy(phih, G.phih_grid.data(), N3 * 8); memcpy(phihe, G.phihe_grid.data(), 2 * N3 * 8); memcpy(phiheat, G.phiheat.data(), N3 * 8);
}
void orc_set_sources(int NumSrc, const int* srcpos, const double* NormFlux, const double* NormFluxPL, const double* NormFluxQPL) {
  G.NumSrc = NumSrc;
  G.srcpos.assign(srcpos, srcpos + 3 * (size_t)NumSrc);
  G.NormFlux.assign(NormFlux, NormFlux + NumSrc);
  if (NormFluxPL) G.NormFluxPL.assign(NormFluxPL, NormFluxPL + NumSrc); else G.NormFluxPL.assign(NumSrc, 0.0);
  if (NormFluxQPL) G.NormFluxQPL.assign(NormFluxQPL, NormFluxQPL + NumSrc); else G.NormFluxQPL.assign(NumSrc, 0.0);
}

// evolve.F90:371-381 set_rates_to_zero
void orc_set_rates_to_zero() {
  std::fill(G.phih_grid.begin(), G.phih_grid.end(), 0.0);
  std::fill(G.phihe_grid.begin(), G.phihe_grid.end(), 0.0);
  std::fill(G.phiheat.begin(), G.phiheat.end(), 0.0);
  for (int b = 0; b < NumFreqBnd; b++) G.photon_loss_all[b] = 0.0;
}

// evolve.F90:385-431 pass_all_sources over the sources rank, rank+npr, ... (master_slave.F90:85 do_grid_static)
// with nthreads OpenMP workers standing in for MPI ranks (private rate grids, summed in rank order afterwards:
// evolve.F90:505-548).  order: 0 = reference serial sweep order, 1 = shell order, 2 = shell order with the threads
// that the source loop leaves idle (fewer sources than threads) working inside each source's shells.
// Returns the number of source x cell updates done.
long orc_pass_all_sources(int nthreads, int order, int rank, int npr, int* nbox_per_source) {
  const size_t N3 = ncell();
  if (nthreads < 1) nthreads = 1;
  std::vector<int> mine;
  for (int ns1 = 1 + rank; ns1 <= G.NumSrc; ns1 += npr) mine.push_back(ns1);
  int inner = 1;
  if (order == 2) {  // workers = min(sources, threads); the rest of the threads go inside the shells
    const int outer = std::max(1, std::min(nthreads, (int)mine.size()));
    inner = std::max(1, nthreads / outer);
    nthreads = outer;
#ifdef _OPENMP
    omp_set_max_active_levels(2);
    omp_set_nested(1);  // older libgomp builds (the copy bundled with torch wins when torch was imported first) need this one
#endif
  }
  if ((int)workers.size() != nthreads) setup_workers(nthreads);
  for (auto& W : workers) {
    W.photon_loss = 0; W.sum_nbox = 0; W.updates = 0;
    if (!W.own_rates.empty()) std::fill(W.own_rates.begin(), W.own_rates.end(), 0.0);
  }
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
  for (int q = 0; q < (int)mine.size(); q++) {
#ifdef _OPENMP
    int w = omp_get_thread_num();
#else
    int w = 0;
#endif
    int nb = do_source(workers[w], mine[q], order, inner);
    if (nbox_per_source) nbox_per_source[mine[q] - 1] = nb;
  }
  long upd = 0;
  double loss = 0.0; long nbox = 0;
  for (int w = 0; w < nthreads; w++) {
    Worker& W = workers[w];
    if (w > 0) {
#pragma omp parallel for num_threads(nthreads)
      for (long p = 0; p < (long)N3; p++) {
        G.phih_grid[p] += W.phih[p]; G.phihe_grid[p] += W.phihe0[p]; G.phihe_grid[p + N3] += W.phihe1[p];
        G.phiheat[p] += W.phiheat[p];
      }
    }
    loss += W.photon_loss; nbox += W.sum_nbox; upd += W.updates;
  }
  G.photon_loss_all[0] += loss;  // photon_loss(1) only (evolve_source.F90:233)
  G.sum_nbox_all = nbox;
  G.rt_updates = upd;
  return upd;
}
double orc_photon_loss() { return G.photon_loss_all[0]; }
long orc_sum_nbox() { return G.sum_nbox_all; }

// evolve.F90:435-501 global_pass (the loop :477-484); nit_out may be NULL; nthreads>1 only changes wall time
// evolve0D_global over the cells [p0, p1) only (serial): what one rank does when the global pass is split over the ranks
// (the reference runs the whole mesh on every rank, evolve.F90:477-484; cells are independent there).  Returns the number
// of cells of the range that vote "not converged".
int orc_global_pass_range(double dt, long p0, long p1) {
  int conv_flag = 0;
  long therm_total = 0;
  for (long p = p0; p < p1; p++) evolve0D_global(dt, (size_t)p, conv_flag, &therm_total);
  return conv_flag;
}

int orc_global_pass(double dt, int nthreads, int* nit_out) {
  const long N3 = (long)ncell();
  int conv_flag = 0;
  long nit_total = 0, therm_total = 0; int nit_max = 0;
  if (nthreads <= 1) {
    for (long p = 0; p < N3; p++) {
      int nit = evolve0D_global(dt, (size_t)p, conv_flag, &therm_total);
      if (nit_out) nit_out[p] = nit;
      nit_total += nit; nit_max = std::max(nit_max, nit);
    }
  } else {
    // cells are independent; G.rc (module globals) is only read in isothermal mode and private otherwise
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 4096) reduction(+ : conv_flag, nit_total, therm_total) reduction(max : nit_max)
    for (long p = 0; p < N3; p++) {
      int cf = 0; long ts = 0;
      // NOTE: in non-isothermal mode do_chemistry writes G.rc; give each call a private copy via thread-local swap
      int nit;
      {
        const size_t N3s = (size_t)N3;
        (void)N3s;
        static thread_local RecCol rc_local;
        // re-implementation of evolve0D_global with a private RecCol
        IonStates ion;
        for (int nx = 0; nx < 2; nx++) {
          ion.h[nx] = std::max(epsilon, G.xh_intermed[p + nx * N3]);
          ion.h_old[nx] = std::max(epsilon, G.xh[p + nx * N3]);
          ion.h_av[nx] = std::max(epsilon, G.xh_av[p + nx * N3]);
        }
        for (int nx = 0; nx < 3; nx++) {
          ion.he[nx] = std::max(epsilon, G.xhe_intermed[p + nx * N3]);
          ion.he_old[nx] = std::max(epsilon, G.xhe[p + nx * N3]);
          ion.he_av[nx] = std::max(epsilon, G.xhe_av[p + nx * N3]);
        }
        double temper_inter, temp_av_old, temper_old;
        get_temperature_point((size_t)p, temper_inter, temp_av_old, temper_old);
        PhotRates phi;
        phi.photo_cell_HI = G.phih_grid[p]; phi.photo_cell_HeI = G.phihe_grid[p]; phi.photo_cell_HeII = G.phihe_grid[p + N3];
        if (!G.isothermal) phi.heat = G.phiheat[p];
        double avg_temper = temp_av_old, temper1;
        rc_local = G.rc;
        nit = do_chemistry(dt, G.ndens[p], ion, phi, temper_old, avg_temper, temper1, rc_local, &ts,
                           G.use_clumping_grid ? (double)G.clumping_grid[p] : (double)G.clumping);
        if (!G.isothermal) { G.temperature_grid[p] = (float)temper1; G.temperature_grid[p + N3] = (float)avg_temper; }
        const double yh0 = G.xh_av[p], yhe0 = G.xhe_av[p], yhe2 = G.xhe_av[p + 2 * N3];
        double temp_av_new, d1, d2;
        get_temperature_point((size_t)p, d1, temp_av_new, d2);
        if ((fabs((ion.h_av[0] - yh0)) > minimum_fractional_change &&
             fabs((ion.h_av[0] - yh0) / ion.h_av[0]) > minimum_fractional_change && (ion.h_av[0] > minimum_fraction_of_atoms)) ||
            (fabs((ion.he_av[0] - yhe0)) > minimum_fractional_change &&
             fabs((ion.he_av[0] - yhe0) / ion.he_av[0]) > minimum_fractional_change && (ion.he_av[0] > minimum_fraction_of_atoms)) ||
            (fabs((ion.he_av[2] - yhe2)) > minimum_fractional_change &&
             fabs((ion.he_av[2] - yhe2) / ion.he_av[2]) > minimum_fractional_change && (ion.he_av[2] > minimum_fraction_of_atoms)) ||
            ((fabs((temp_av_old - temp_av_new) / temp_av_new) > 1.0e-1) && (fabs(temp_av_new - temp_av_old) > 100.0)))
          cf = 1;
        for (int nx = 0; nx < 2; nx++) { G.xh_intermed[p + nx * N3] = ion.h[nx]; G.xh_av[p + nx * N3] = ion.h_av[nx]; }
        for (int nx = 0; nx < 3; nx++) { G.xhe_intermed[p + nx * N3] = ion.he[nx]; G.xhe_av[p + nx * N3] = ion.he_av[nx]; }
      }
      if (nit_out) nit_out[p] = nit;
      conv_flag += cf; nit_total += nit; therm_total += ts; nit_max = std::max(nit_max, nit);
    }
  }
  G.nit_total = nit_total; G.nit_max = nit_max;
  return conv_flag;
}
long orc_last_nit_total() { return G.nit_total; }
int orc_last_nit_max() { return G.nit_max; }

// evolve.F90:78-229 evolve3D (restart == 0).  stats: [0]=niter [1]=last conv_flag [2]=conv_criterion
// [3]=sum_nbox_all (last iteration) [4]=total RT updates ; conv_hist[niter] conv_flag after each iteration (<=512)
void orc_evolve3d(double dt, int nthreads, int order, long* stats, int* conv_hist) {
  const size_t N3 = ncell();
  G.xh_av = G.xh; G.xh_intermed = G.xh; G.xhe_av = G.xhe; G.xhe_intermed = G.xhe;
  int niter = 0;
  int conv_flag = G.mesh[0] * G.mesh[1] * G.mesh[2];
  int conv_criterion = std::min((int)(convergence_fraction * G.mesh[0] * G.mesh[1] * G.mesh[2]), G.NumSrc);
  long updates = 0;
  for (;;) {
    if (conv_flag < conv_criterion && niter > 1) {
      G.xh = G.xh_intermed; G.xhe = G.xhe_intermed;
      if (!G.isothermal) memcpy(&G.temperature_grid[2 * N3], &G.temperature_grid[0], N3 * sizeof(float));
      break;
    } else if (niter > 500) break;
    niter = niter + 1;
    orc_set_rates_to_zero();
    if (G.NumSrc > 0) updates += orc_pass_all_sources(nthreads, order, 0, 1, nullptr);
    conv_flag = orc_global_pass(dt, nthreads, nullptr);
    if (conv_hist && niter < 512) conv_hist[niter] = conv_flag;
  }
  stats[0] = niter; stats[1] = conv_flag; stats[2] = conv_criterion; stats[3] = G.sum_nbox_all; stats[4] = updates;
}

// photonstatistics.f90:117-147 state_before / :208-247 state_after : 5 sums in i,j,k order
void orc_state_sums(const double* xh, const double* xhe, double* out5) {
  const size_t N3 = ncell();
  double s[5] = {0, 0, 0, 0, 0};
  for (size_t p = 0; p < N3; p++) {
    s[0] = s[0] + G.ndens[p] * xh[p]; s[1] = s[1] + G.ndens[p] * xh[p + N3];
    s[2] = s[2] + G.ndens[p] * xhe[p]; s[3] = s[3] + G.ndens[p] * xhe[p + N3]; s[4] = s[4] + G.ndens[p] * xhe[p + 2 * N3];
  }
  out5[0] = s[0] * G.vol * (1.0 - abu_he); out5[1] = s[1] * G.vol * (1.0 - abu_he);
  out5[2] = s[2] * G.vol * abu_he; out5[3] = s[3] * G.vol * abu_he; out5[4] = s[4] * G.vol * abu_he;
}

// photonstatistics.f90:150-204 total_rates with the module-global coefficients as the last do_chemistry left them
// (serial global pass only), then *vol*dt (:201-203).  out3 = totrec, totcollisions, recomions
void orc_total_rates(double dt, const double* xh_l, const double* xhe_l, double* out3) {
  const size_t N3 = ncell();
  double totrec = 0.0, totcollisions = 0.0, recomions = 0.0;
  double clumping = (double)G.clumping;
  for (size_t p = 0; p < N3; p++) {
    if (G.use_clumping_grid) clumping = (double)G.clumping_grid[p];  // photonstatistics.f90:176
    const double yh[2] = {xh_l[p], xh_l[p + N3]};
    const double yhe[3] = {xhe_l[p], xhe_l[p + N3], xhe_l[p + 2 * N3]};
    const double ndens_p = G.ndens[p];
    totrec = totrec + ndens_p * (yh[1] * G.rc.brech0 * (1.0 - abu_he) + yhe[1] * G.rc.breche0 * abu_he * 0.04) *
                          electrondens(ndens_p, yh, yhe) * clumping;
    totcollisions = totcollisions + ndens_p * electrondens(ndens_p, yh, yhe) *
                                        (yh[0] * G.rc.colli_HI + yhe[0] * G.rc.colli_HeI + yhe[1] * G.rc.colli_HeII);
    recomions = recomions + ndens_p * abu_he * clumping * (yhe[2] * 1.121 * G.rc.breche1 + yhe[1] * G.rc.breche0 * 0.96) *
                                abu_he * electrondens(ndens_p, yh, yhe);
  }
  out3[0] = totrec * G.vol * dt; out3[1] = totcollisions * G.vol * dt; out3[2] = recomions * G.vol * dt;
}

// mrgrnk.f90:21-215 R_mrgrnk: rank of a real array, stable under ties (merge sort); ranks are 1-based.
void orc_mrgrnk(int n, const float* x, int* irngt) {
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return x[a] < x[b]; });
  for (int i = 0; i < n; i++) irngt[i] = idx[i] + 1;
}

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// A launcher may have exported OMP_NUM_THREADS=1 for its children (torch.distributed.run does): the CPU arms of
// bench.py set the thread count they state explicitly.
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

// standalone cinterp on a caller-provided scratch (for the cinterp_batch parity hook)
void orc_cinterp(const int* mesh, const double* cdh, const double* cdhe0, const double* cdhe1, const int* pos,
                 const int* srcpos, double* out4) {
  int save[3] = {G.mesh[0], G.mesh[1], G.mesh[2]};
  for (int d = 0; d < 3; d++) G.mesh[d] = mesh[d];
  Worker W; size_t N3 = ncell();
  W.cdh.assign(cdh, cdh + N3); W.cdhe0.assign(cdhe0, cdhe0 + N3); W.cdhe1.assign(cdhe1, cdhe1 + N3);
  cinterp(W, pos, srcpos, out4[0], out4[1], out4[2], out4[3]);
  for (int d = 0; d < 3; d++) G.mesh[d] = save[d];
}

}  // extern "C"
