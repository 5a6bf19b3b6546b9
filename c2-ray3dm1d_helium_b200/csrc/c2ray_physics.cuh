// c2ray_physics.cuh -- per-cell device physics of the C2-Ray H+He hot path (sm_100a, FP64).
//
// Written for registers, not as a translation: the reference's module globals (recombination /
// collisional coefficients, cgsconstants.f90:105-133) become a per-thread struct, its 47-element per-cell
// work arrays (radiation_photoionrates.f90:147-159) are never materialised -- photoion_rates streams over the
// frequency bands keeping only accumulators -- and the secondary-ionisation factors are hoisted out of the
// per-SED loop.  Each function cites the reference lines whose arithmetic it must reproduce.
#pragma once
#include "c2ray_consts.cuh"
#include "c2ray_fastmath.cuh"

#ifndef C2RAY_FAST_COEF
#define C2RAY_FAST_COEF 1
#endif
// doric's exponentials feed (e^x - 1)/x and differences of O(1) terms: libdevice exp (< 1 ulp near 1) keeps the
// fractions at the 1.5e-11 noise level, the cheaper fast_exp (~2 ulp) measured 2.5e-10 for no gain in time.
#ifndef C2RAY_DORIC_EXP
#define C2RAY_DORIC_EXP exp
#endif
#ifndef C2RAY_DORIC_IEEE_DIV
#define C2RAY_DORIC_IEEE_DIV 0
#endif
#if C2RAY_DORIC_IEEE_DIV
#define DDIV(a, b) ((a) / (b))
#define DDIVR(a, b, r) ((a) / (b))
#else
#define DDIV(a, b) fdiv((a), (b))
#define DDIVR(a, b, r) fdiv_r((a), (b), (r))
#endif

namespace c2 {

// ------------------------------------------------------------------------------------------------
// Device-visible tables and per-run constants
// ------------------------------------------------------------------------------------------------
struct SedDev {
  const double* photo_thick;  // [band][0:NumTau]
  const double* photo_thin;
  const double* heat_thick;   // [heatbin][0:NumTau]
  const double* heat_thin;
  const double* packed;       // the same four tables as band-major 64-byte rows (c2ray_photo.cuh)
  int lo, hi;                 // 1-based FreqBnd limits (hi < lo : SED absent)
  double S_star;
};

// radiation_sizes.f90 band arrays, one 128-byte record per band (0-based band index): a band step reads up to 15 of
// these through dynamically indexed constant loads, and one or two constant-cache lines per band serve them where 15
// separate arrays cost 15 lines (the LDC waits showed up as "short scoreboard" stalls: 13 % of the samples at 16
// sources, 24 % at 1000)
struct BandRec {
  double sigma_HI, sigma_HeI, sigma_HeII;
  double f1ion_HI, f1ion_HeI, f1ion_HeII;
  double f2ion_HI, f2ion_HeI, f2ion_HeII;
  double f1heat_HI, f1heat_HeI, f1heat_HeII;
  double f2heat_HI, f2heat_HeI, f2heat_HeII;
  double dead_bb;  // optical depth from which the black-body tables of this band are all zero (d_dead, c2ray_photo.cuh)
};

struct RunConst {
  SedDev sed[3];
  int isothermal, cosmological;
  double temper_val;
  double clumping;        // (double) of the reference's real
  double cosmo_coef;      // 2.0/(1.0+zred)*dzdt  applied as e_int*2.0/(1.0+zred)*dzdt
  double zp1, dzdt;
  double dr[3], vol;
  double cool_mintemp, cool_dtemp, cool_rdtemp;  // rdtemp = 1/dtemp
  const double* cool;     // 5 x 801 linear cooling tables: h0,h1,he0,he1,he2
  int mesh[3];
};

__constant__ BandRec d_band[NumFreqBnd];
__constant__ RunConst d_run;

struct RecCol {
  double arech0, brech0, areche0, breche0, oreche0, areche1, breche1, treche1;
  double colli_HI, colli_HeI, colli_HeII, v;
};

struct Ion {
  double h0, h1, he0, he1, he2;
  double h_av0, h_av1, he_av0, he_av1, he_av2;
  double h_old0, h_old1, he_old0, he_old1, he_old2;
};

// ------------------------------------------------------------------------------------------------
// cgsconstants.f90:140-266 ini_rec_colion_factors (literal kinds reproduced, see FL())
// ------------------------------------------------------------------------------------------------
// The 13 pow() of the fits share four bases (lambda and T up to constant factors), so one log and one exp per term
// replace them: x^p = exp(p ln x), ln(lambda_He) = ln(lambda_H) + ln(T_He/T_H), ln(T/1e4) = ln(2 T_H/1e4) - ln(lambda_H).
// ln of the (binary32 / _dp) fit constants, to 20 digits:
constexpr double LN_F0522 = -0.65008766278158320814;   // ln((double)0.522f)
constexpr double LN_F2740 = 1.0079579238805420492;     // ln((double)2.740f)
constexpr double LN_D0522 = -0.65008769109949833266;   // ln(0.522_dp)
constexpr double LN_D2740 = 1.0079579203999788567;     // ln(2.740_dp)
constexpr double LN_HE0_OVER_H = 0.59229515194333738436;  // ln(temphe(0)/temph0)
constexpr double LN_HE1_OVER_H = 1.3867355433101860653;   // ln(temphe(1)/temph0)
constexpr double LN_2TH0_OVER_1E4 = 3.4519179578675728116;  // ln(2*temph0/1e4)

// lam^pA / (1 + (lam/div)^pB)^pC  with lnl = ln(lam), lnd = ln(div)
__device__ __forceinline__ double hg_fit(double lnl, double lnd, double pA, double pB, double pC) {
  const double u = fast_exp(pB * (lnl - lnd));
  return fast_exp(fma(pA, lnl, -pC * fast_log(1.0 + u)));
}

#if C2RAY_FAST_COEF
__device__ __forceinline__ void ini_rec_colion_factors(double T, RecCol& r) {
  const double rT = fast_rcp(T);
  const double lambda = 2.0 * (temph0 * rT);
  const double lnl = fast_log(lambda);
  // the H fit is shared by arech0/brech0 and (T<9e3) areche0/breche0 up to the leading literal
  const double hA = hg_fit(lnl, LN_F0522, 1.503, FL(0.470f), FL(1.923f));
  const double hB = hg_fit(lnl, LN_F2740, 1.500, FL(0.407f), FL(2.242f));
  r.arech0 = FL(1.269e-13f) * hA;
  r.brech0 = FL(2.753e-14f) * hB;
  const double sqrtt0 = sqrt(T);
  if (T < 9.e3) {
    r.areche0 = 1.269e-13 * hA;
    r.breche0 = 2.753e-14 * hB;
  } else {
    const double lnl0 = lnl + LN_HE0_OVER_H;  // lambda = 2*temphe(0)/T
    const double dielectronic = 1.9e-3 * (rT * fast_rcp(sqrtt0)) * fast_exp(-4.7e5 * rT) * (1.0 + 0.3 * fast_exp(-9.4e4 * rT));
    r.areche0 = 3.000e-14 * fast_exp(0.654 * lnl0) + dielectronic;
    r.breche0 = 1.260e-14 * fast_exp(0.750 * lnl0) + dielectronic;
  }
  r.oreche0 = r.areche0 - r.breche0;
  const double lnl1 = lnl + LN_HE1_OVER_H;    // lambda = 2*temphe(1)/T
  r.breche1 = 5.5060e-14 * hg_fit(lnl1, LN_D2740, 1.5, 0.407, 2.242);
  r.areche1 = FL(2.538e-13f) * hg_fit(lnl1, LN_D0522, 1.503, 0.470, 1.923);
  const double lnt4 = LN_2TH0_OVER_1E4 - lnl;  // ln(T/1e4)
  r.treche1 = 3.4e-13 * fast_exp(-0.6 * lnt4);
  r.v = 0.285 * fast_exp(0.119 * lnt4);
  r.colli_HI = colh0 * sqrtt0 * fast_exp(-temph0 * rT);
  r.colli_HeI = colhe0 * sqrtt0 * fast_exp(-temphe0 * rT);
  r.colli_HeII = colhe1 * sqrtt0 * fast_exp(-temphe1 * rT);
}
#else
// libdevice pow/exp version (13 pow): 2x slower global pass; kept for A/B checks (-DC2RAY_FAST_COEF=0).
__device__ __forceinline__ void ini_rec_colion_factors(double T, RecCol& r) {
  double lambda = 2.0 * (temph0 / T);
  // the H fit is shared by arech0/brech0 and (T<9e3) areche0/breche0 up to the leading literal
  const double hA = pow(lambda, 1.503) / pow(1.0 + pow(lambda / FL(0.522f), FL(0.470f)), FL(1.923f));
  const double hB = pow(lambda, 1.500) / pow(1.0 + pow(lambda / FL(2.740f), FL(0.407f)), FL(2.242f));
  r.arech0 = FL(1.269e-13f) * hA;
  r.brech0 = FL(2.753e-14f) * hB;
  if (T < 9.e3) {
    r.areche0 = 1.269e-13 * hA;
    r.breche0 = 2.753e-14 * hB;
  } else {
    const double lam0 = 2.0 * (temphe0 / T);
    const double dielectronic = 1.9e-3 * pow(T, -1.5) * exp(-4.7e5 / T) * (1.0 + 0.3 * exp(-9.4e4 / T));
    r.areche0 = 3.000e-14 * pow(lam0, 0.654) + dielectronic;
    r.breche0 = 1.260e-14 * pow(lam0, 0.750) + dielectronic;
  }
  r.oreche0 = r.areche0 - r.breche0;
  lambda = 2.0 * (temphe1 / T);
  r.breche1 = 5.5060e-14 * pow(lambda, 1.5) / pow(1.0 + pow(lambda / 2.740, 0.407), 2.242);
  r.areche1 = FL(2.538e-13f) * pow(lambda, 1.503) / pow(1.0 + pow(lambda / 0.522, 0.470), 1.923);
  r.treche1 = 3.4e-13 * pow(T / 1.0e4, -0.6);
  r.v = 0.285 * pow(T / 1.0e4, 0.119);
  const double sqrtt0 = sqrt(T);
  r.colli_HI = colh0 * sqrtt0 * exp(-temph0 / T);
  r.colli_HeI = colhe0 * sqrtt0 * exp(-temphe0 / T);
  r.colli_HeII = colhe1 * sqrtt0 * exp(-temphe1 / T);
}
#endif

// tped.f90:75-84
__device__ __forceinline__ double electrondens(double n, double xh1, double xhe1, double xhe2) {
  return n * (xh1 * (1.0 - abu_he) + abu_c + abu_he * (xhe1 + 2.0 * xhe2));
}

// cooling_h.f90:40-71
__device__ __forceinline__ double coolin(double n, double ne, double h_av0, double h_av1, double he_av0, double he_av1,
                                         double he_av2, double T, const double* __restrict__ ct) {
  const double tpos = fma(fast_log10(T) - d_run.cool_mintemp, d_run.cool_rdtemp, 1.0);
  const int itpos = min(TEMPPOINTS - 1, max(1, (int)tpos));
  const double dtpos = tpos - (double)itpos;
  const int a = itpos - 1, b = min(TEMPPOINTS, itpos + 1) - 1;
  const double c0 = ct[a] + (ct[b] - ct[a]) * dtpos;
  const double c1 = ct[TEMPPOINTS + a] + (ct[TEMPPOINTS + b] - ct[TEMPPOINTS + a]) * dtpos;
  const double c2v = ct[2 * TEMPPOINTS + a] + (ct[2 * TEMPPOINTS + b] - ct[2 * TEMPPOINTS + a]) * dtpos;
  const double c3 = ct[3 * TEMPPOINTS + a] + (ct[3 * TEMPPOINTS + b] - ct[3 * TEMPPOINTS + a]) * dtpos;
  const double c4 = ct[4 * TEMPPOINTS + a] + (ct[4 * TEMPPOINTS + b] - ct[4 * TEMPPOINTS + a]) * dtpos;
  return n * ne * ((h_av0 * c0 + h_av1 * c1) * (1.0 - abu_he) + (he_av0 * c2v + he_av1 * c3 + he_av2 * c4) * abu_he);
}

// doric.f90:317-351 prepare_doric_factors with coldens (:358-372) at path = 1
struct DoricFrac { double y, z, y2a, y2b; };
__device__ __forceinline__ DoricFrac prepare_doric_factors(double n, double h0, double he0, double he1) {
  const double NH = h0 * n * 1.0 * (1.0 - abu_he);
  const double NHe0 = he0 * n * 1.0 * abu_he;
  const double NHe1 = he1 * n * 1.0 * abu_he;
  const double tau_H_heth = NH * sigma_H_heth, tau_He_heth = NHe0 * sigma_HeI_at_ion_freq;
  const double tau_H_heLya = NH * sigma_H_heLya, tau_He_heLya = NHe0 * sigma_He_heLya;
  const double tau_H_he2th = NH * sigma_H_he2, tau_He_he2th = NHe0 * sigma_He_he2;
  const double tau_He2_he2th = NHe1 * sigma_HeII_at_ion_freq;
  DoricFrac f;
  f.y = DDIV(tau_H_heth, tau_H_heth + tau_He_heth);
  f.z = DDIV(tau_H_heLya, tau_H_heLya + tau_He_heLya);
  const double dden = tau_He2_he2th + tau_He_he2th + tau_H_he2th;
  const double rden = fast_rcp(dden);
  f.y2a = DDIVR(tau_He2_he2th, dden, rden);
  f.y2b = DDIVR(tau_He_he2th, dden, rden);
  return f;
}

// doric.f90:35-313
__device__ __forceinline__ void doric(double dt, double rhe, Ion& ion, double phiHI, double phiHeI, double phiHeII,
                                      const DoricFrac& fr, const RecCol& rc, double clumping) {
  const double pfrac = 0.96;
  const double heliumfraction = abu_he / (1.0 - abu_he);
  const double ffrac = fmax(fmin(10.0 * ion.h0, 1.0), 0.01);
  const double wfrac = (1.425 - 0.737) + 0.737 * fr.y;
  const double v = rc.v;
  const double alpha_h_B = clumping * rc.brech0;
  const double alpha_he_1 = clumping * rc.oreche0;
  const double alpha_he_B = clumping * rc.breche0;
  const double alpha_he_A = clumping * rc.areche0;
  const double alpha_he2_B = clumping * rc.breche1;
  const double alpha_he2_A = clumping * rc.areche1;
  const double alpha_he2_2 = clumping * rc.treche1;
  const double alpha_he2_1 = alpha_he2_A - alpha_he2_B;
  const double aih0 = fmax(phiHI + rhe * rc.colli_HI, 1.0e-200);
  const double aihe0 = fmax(phiHeI + rhe * rc.colli_HeI, 1.0e-200);
  const double aihe1 = fmax(phiHeII + rhe * rc.colli_HeII, 1.0e-200);

  const double Lmat = -(aih0 + rhe * alpha_h_B);
  const double Mmat = (fr.y * rhe * alpha_he_1 + pfrac * rhe * alpha_he_B) * heliumfraction;
  const double Nmat = ((ffrac * fr.z * (1.0 - v) + v * wfrac) * alpha_he2_B + alpha_he2_2 +
                       (1.0 - fr.y2a - fr.y2b) * alpha_he2_1) * heliumfraction * rhe;
  const double Pmat = -aihe0 - aihe1 - rhe * (alpha_he_A - (1.0 - fr.y) * alpha_he_1);
  const double Emat = -rhe * (alpha_he2_A - fr.y2a * alpha_he2_1);
  const double Qmat = -aihe0 + rhe * alpha_he2_B * (ffrac * (1.0 - fr.z) * (1.0 - v) + v * (1.425 - wfrac)) - Emat +
                      alpha_he2_1 * fr.y2b * rhe;
  const double Bcoef = Emat - Pmat;
  const double Scoef = sqrt(Bcoef * Bcoef + 4.0 * aihe1 * Qmat);
  const double QHEPcoef = DDIV(1.0, Qmat * aihe1 - Emat * Pmat);
  const double BminusS = Bcoef - Scoef, BplusS = Bcoef + Scoef;
  const double lambda1 = Lmat;
  const double lambda2 = 0.5 * (Emat + Pmat - Scoef);
  const double lambda3 = 0.5 * (Emat + Pmat + Scoef);
  const double rx = DDIV(-1.0, Lmat) * (aih0 + (Mmat * Emat - Nmat * aihe1) * (aihe0 * QHEPcoef));
  const double ry = aihe0 * (Emat * QHEPcoef);
  const double rz = -aihe0 * (aihe1 * QHEPcoef);
  const double twoaihe1 = 2.0 * aihe1;
  // doric's coefficients are differences of large terms (noise amplification ~1e5): every
  // quotient below is corrected to ~0.5 ulp so the GPU stays at the noise level of the reference's own arithmetic
  const double dL2 = Lmat - lambda2, dL3t = twoaihe1 * (Lmat - lambda3), twoS = 2.0 * Scoef;
  const double r2a = fast_rcp(twoaihe1), rL2 = fast_rcp(dL2), rL3t = fast_rcp(dL3t), r2S = fast_rcp(twoS);
  const double eigv2x = DDIVR(-Nmat, dL2, rL2) + DDIVR(DDIVR(Mmat, twoaihe1, r2a) * BplusS, dL2, rL2);
  const double eigv3x = DDIVR(-twoaihe1 * Nmat + Mmat * BminusS, dL3t, rL3t);
  const double eigv2y = DDIVR(-BplusS, twoaihe1, r2a);
  const double eigv3y = DDIVR(-BminusS, twoaihe1, r2a);
  const double Rcoef = twoaihe1 * (ry - ion.he_old1);
  const double Tcoef = rz - ion.he_old2;
  const double coef2 = DDIVR(Rcoef + BminusS * Tcoef, twoS, r2S);
  const double coef3 = -DDIVR(Rcoef + BplusS * Tcoef, twoS, r2S);
  const double coef1 = -rx + (eigv3x - eigv2x) * DDIVR(Rcoef, twoS, r2S) +
                       Tcoef * (DDIVR(BplusS * eigv3x, twoS, r2S) - DDIVR(BminusS * eigv2x, twoS, r2S)) + ion.h_old1;
  const double lam1dt = dt * lambda1, lam2dt = dt * lambda2, lam3dt = dt * lambda3;
  const double elam1dt = C2RAY_DORIC_EXP(lam1dt), elam2dt = C2RAY_DORIC_EXP(lam2dt), elam3dt = C2RAY_DORIC_EXP(lam3dt);

  ion.h1 = coef1 * elam1dt + coef2 * elam2dt * eigv2x + coef3 * elam3dt * eigv3x + rx;
  ion.he1 = coef2 * elam2dt * eigv2y + coef3 * elam3dt * eigv3y + ry;
  ion.he2 = coef2 * elam2dt + coef3 * elam3dt + rz;
  ion.h0 = 1.0 - ion.h1;
  ion.he0 = 1.0 - ion.he1 - ion.he2;
  if (ion.h0 < epsilon) { ion.h0 = epsilon; ion.h1 = 1.0 - epsilon; }
  if (ion.h1 < epsilon) { ion.h1 = epsilon; ion.h0 = 1.0 - epsilon; }
  if ((ion.he0 <= epsilon) || (ion.he1 <= epsilon) || (ion.he2 <= epsilon)) {
    if (ion.he0 < epsilon) ion.he0 = epsilon;
    if (ion.he1 < epsilon) ion.he1 = epsilon;
    if (ion.he2 < epsilon) ion.he2 = epsilon;
    const double normfac = ion.he0 + ion.he1 + ion.he2;
    ion.he0 = ion.he0 / normfac; ion.he1 = ion.he1 / normfac; ion.he2 = ion.he2 / normfac;
  }
  const double lim = FL(1.0e-8f);
  const double af1 = (fabs(lam1dt) < lim) ? coef1 : DDIV(coef1 * (elam1dt - 1.0), lam1dt);
  const double af2 = (fabs(lam2dt) < lim) ? coef2 : DDIV(coef2 * (elam2dt - 1.0), lam2dt);
  const double af3 = (fabs(lam3dt) < lim) ? coef3 : DDIV(coef3 * (elam3dt - 1.0), lam3dt);
  ion.h_av1 = rx + af1 + eigv2x * af2 + eigv3x * af3;
  ion.he_av1 = ry + eigv2y * af2 + eigv3y * af3;
  ion.he_av2 = rz + af2 + af3;
  ion.h_av0 = 1.0 - ion.h_av1;
  ion.he_av0 = 1.0 - ion.he_av1 - ion.he_av2;
  if (ion.h_av1 < epsilon) { ion.h_av1 = epsilon; ion.h_av0 = 1.0 - epsilon; }
  if (ion.h_av0 < epsilon) { ion.h_av0 = epsilon; ion.h_av1 = 1.0 - epsilon; }
  if ((ion.he_av0 <= epsilon) || (ion.he_av1 <= epsilon) || (ion.he_av2 <= epsilon)) {
    if (ion.he_av1 < epsilon) ion.he_av1 = epsilon;
    if (ion.he_av2 < epsilon) ion.he_av2 = epsilon;
    if (ion.he_av0 < epsilon) ion.he_av0 = epsilon;
    const double normfac = ion.he_av0 + ion.he_av1 + ion.he_av2;
    ion.he_av0 = ion.he_av0 / normfac; ion.he_av1 = ion.he_av1 / normfac; ion.he_av2 = ion.he_av2 / normfac;
  }
}

// thermal.f90:22-174, split into begin / sub-step / end so that the queue-driven global pass (k_global_pass_q) can
// suspend a cell between sub-steps; thermal() below composes the three and is what the batch hook runs.
struct ThermState {
  double internal_energy, end_temper, avg_acc, cumulative_time, cosmo_cool_rate, ne_av, rkn_av, initial_temp;
  int i_heating;
  bool active;   // end_temper > minitemp at entry (thermal.f90:83)
};

__device__ __forceinline__ void thermal_begin(ThermState& S, double end_temper, double n, const Ion& ion) {
  S.internal_energy = (n + electrondens(n, ion.h_old1, ion.he_old1, ion.he_old2)) * k_B * end_temper / gamma1;  // :68
  S.cosmo_cool_rate = d_run.cosmological ? S.internal_energy * FL(2.0f) / d_run.zp1 * d_run.dzdt : 0.0;  // cosmology.f90:232
  S.i_heating = 0;
  S.end_temper = end_temper;
  S.initial_temp = end_temper;
  S.active = end_temper > minitemp;
  S.ne_av = electrondens(n, ion.h_av1, ion.he_av1, ion.he_av2);
  S.rkn_av = fast_rcp(k_B * (n + S.ne_av));
  S.cumulative_time = 0.0;
  S.avg_acc = 0.0;
}

// one explicit sub-step (:98-157); returns true when the loop of the reference would exit
__device__ __forceinline__ bool thermal_substep(ThermState& S, double dt, double ne, double n, const Ion& ion,
                                                double heating) {
  S.i_heating++;
  const double cooling =
      coolin(n, ne, ion.h_av0, ion.h_av1, ion.he_av0, ion.he_av1, ion.he_av2, S.end_temper, d_run.cool) + S.cosmo_cool_rate;
  const double thermal_rate = fmax(1e-50, fabs(cooling - heating));
  const double thermal_timescale = fdiv(S.internal_energy, thermal_rate);
  const double dt_thermal = relative_denergy * thermal_timescale;
  const double dt_ODE = fmin(dt_thermal, dt - S.cumulative_time);
  S.internal_energy = S.internal_energy + dt_ODE * (heating - cooling);
  S.avg_acc = S.avg_acc + FL(0.5f) * S.end_temper * dt_ODE;
  S.end_temper = (S.internal_energy * gamma1) * S.rkn_av;
  S.avg_acc = S.avg_acc + FL(0.5f) * S.end_temper * dt_ODE;
  if (S.end_temper < minitemp) {
    S.internal_energy = (n + S.ne_av) * k_B * minitemp;  // thermal.f90:141 (no /gamma1, as in the reference)
    S.end_temper = minitemp;
  }
  S.cumulative_time = S.cumulative_time + dt_ODE;
  if (S.cumulative_time >= dt || fabs(S.cumulative_time - dt) < FL(1e-6f) * dt) return true;
  return S.i_heating > 10000;
}

// :159-172 ; avg_temper is left untouched when the cell was at or below minitemp (quirk q3)
__device__ __forceinline__ void thermal_end(const ThermState& S, double dt, double n, const Ion& ion, double& end_temper,
                                            double& avg_temper) {
  if (!S.active) return;
  avg_temper = (dt > 0.0) ? S.avg_acc / dt : S.initial_temp;
  end_temper = (S.internal_energy * gamma1) / (k_B * (n + electrondens(n, ion.h1, ion.he1, ion.he2)));
}

__device__ __forceinline__ int thermal(double dt, double& end_temper, double& avg_temper, double ne, double n,
                                       const Ion& ion, double heating) {
  ThermState S;
  thermal_begin(S, end_temper, n, ion);
  if (S.active)
    while (!thermal_substep(S, dt, ne, n, ion, heating)) {}
  thermal_end(S, dt, n, ion, end_temper, avg_temper);
  return S.i_heating;
}

// evolve_point.F90:444-646 do_chemistry (local=.false.), split at the thermal call.
struct ChemIter {  // values the convergence test of one iteration needs (:488-496)
  double yh0_av_old, yhe0_av_old, yhe2_av_old, temper2;
};

// :488-597: coefficients at avg_temper, doric twice with the partial averaging; returns the electron density for thermal
__device__ __forceinline__ double chem_ionization(double dt, double n, Ion& ion, double phiHI, double phiHeI,
                                                  double phiHeII, double avg_temper, double temper1, RecCol& rc,
                                                  ChemIter& it, double clumping) {
  // clumping: material's module scalar, or clumping_grid(i,j,k) of the cell when type_of_clumping == 5
  // (evolve_point.F90:484 clumping_point)
  const bool iso = d_run.isothermal != 0;
  it.temper2 = temper1;
  it.yh0_av_old = ion.h_av0; it.yhe0_av_old = ion.he_av0; it.yhe2_av_old = ion.he_av2;
  double de = electrondens(n, ion.h_av1, ion.he_av1, ion.he_av2);
  if (!iso) ini_rec_colion_factors(avg_temper, rc);
  DoricFrac fr = prepare_doric_factors(n, ion.h0, ion.he0, ion.he1);
  doric(dt, de, ion, phiHI, phiHeI, phiHeII, fr, rc, clumping);
  de = electrondens(n, ion.h_av1, ion.he_av1, ion.he_av2);
  fr = prepare_doric_factors(n, ion.h0, ion.he0, ion.he1);
  const double ionh0old = ion.h0, ionh1old = ion.h1, ionhe0old = ion.he0, ionhe1old = ion.he1, ionhe2old = ion.he2;
  const double oldhav = ion.h_av0, oldhe0av = ion.he_av0, oldhe1av = ion.he_av1;
  doric(dt, de, ion, phiHI, phiHeI, phiHeII, fr, rc, clumping);
  // evolve_point.F90:588-595: h_av(1) and he_av(2) keep their pass-2 values
  ion.h0 = (ion.h0 + ionh0old) * 0.5;
  ion.h1 = (ion.h1 + ionh1old) * 0.5;
  ion.he0 = (ion.he0 + ionhe0old) * 0.5;
  ion.he1 = (ion.he1 + ionhe1old) * 0.5;
  ion.he2 = (ion.he2 + ionhe2old) * 0.5;
  ion.h_av0 = (ion.h_av0 + oldhav) * 0.5;
  ion.he_av0 = (ion.he_av0 + oldhe0av) * 0.5;
  ion.he_av1 = (ion.he_av1 + oldhe1av) * 0.5;
  return electrondens(n, ion.h_av1, ion.he_av1, ion.he_av2);
}

// :607-628
__device__ __forceinline__ bool chem_converged(const Ion& ion, const ChemIter& it, double temper1) {
  return (fabs((ion.h_av0 - it.yh0_av_old) / ion.h_av0) < minimum_fractional_change || ion.h_av0 < minimum_fraction_of_atoms) &&
         (fabs((ion.he_av0 - it.yhe0_av_old) / ion.he_av0) < minimum_fractional_change || ion.he_av0 < minimum_fraction_of_atoms) &&
         (fabs((ion.he_av2 - it.yhe2_av_old) / ion.he_av2) < minimum_fractional_change || ion.he_av2 < minimum_fraction_of_atoms) &&
         (fabs((temper1 - it.temper2) / temper1) < minimum_fractional_change);
}

// temper_old: T at the start of the step (grid(..,2)); avg_temper in: grid(..,1)
__device__ __forceinline__ int do_chemistry(double dt, double n, Ion& ion, double phiHI, double phiHeI, double phiHeII,
                                            double heat, double temper_old, double& avg_temper, double& temper1_out,
                                            RecCol& rc, double clumping, int* nsub_total = nullptr,
                                            double* last_coef_T = nullptr) {
  const bool iso = d_run.isothermal != 0;
  double temper1 = temper_old;
  const double temper0 = temper1;
  int nit = 0;
  for (;;) {
    nit++;
    ChemIter it;
    if (last_coef_T && !iso) *last_coef_T = avg_temper;  // what the reference's module globals hold afterwards
    const double de = chem_ionization(dt, n, ion, phiHI, phiHeI, phiHeII, avg_temper, temper1, rc, it, clumping);
    temper1 = temper0;
    if (!iso) {
      const int ns = thermal(dt, temper1, avg_temper, de, n, ion, heat);
      if (nsub_total) *nsub_total += ns;
    }
    if (chem_converged(ion, it, temper1)) break;
    if (nit > 400) break;
  }
  temper1_out = temper1;
  return nit;
}

// column_density.f90:351-376
__device__ __forceinline__ double weightf(double cd, double sig) { return 1.0 / fmax(0.6, cd * sig); }

}  // namespace c2
