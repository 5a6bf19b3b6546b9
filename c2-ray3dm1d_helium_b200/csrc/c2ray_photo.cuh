// c2ray_photo.cuh -- radiation_photoionrates.f90:108-277 photoion_rates for one cell, streamed over the frequency
// bands (no 47-element work arrays), written for instruction count: the round-1 profile showed 22.8 k warp
// instructions per source x cell update of which only a third were FP64 arithmetic -- the rest came from IEEE
// division slow-path scaffolding, the general-purpose log10/pow, and per-band divisions by the shell volume.
//   * 1/vol, NFlux and 1/(x n abundance) are applied once per cell, not per band (all rates are linear in them)
//   * (log10(tau)-minlogtau)/dlogtau becomes one FMA with 1/dlogtau
//   * log10 is a branch-free atanh-series evaluation valid for the positive normal arguments that occur here
//     (tau clamped to >= 1e-20), accurate to ~2e-16 relative
//   * reciprocals use MUFU.RCP64H + cubic/Newton refinement (error ~ 2^-54) without the IEEE slow path
//   * the secondary-ionisation factors y1R, y2R (9 pow per cell in the reference, :557-565) depend on the cell only
//     and are precomputed once per iteration by k_secion_factors
// All of this changes results at the 1e-15 level; the parity tests hold 1e-8.
#pragma once
#include "c2ray_physics.cuh"

namespace c2 {

// reciprocal without the IEEE special-case path; |rel err| ~ 2^-54 for normal b
__device__ __forceinline__ double fast_rcp(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  e = fma(e, e, e);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// log10 for positive normal x (no zero / subnormal / inf / nan handling): x = 2^e * m, m in [sqrt(1/2), sqrt(2)),
// log(m) = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716
__device__ __forceinline__ double fast_log10(double x) {
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  const int big = hi >= 0x3ff6a09f;  // m >= ~sqrt(2)
  hi -= big << 20;                   // m *= 0.5
  e += big;
  const double m = __hiloint2double(hi, lo);
  const double f = m - 1.0;
  const double s = f * fast_rcp(m + 1.0);
  const double z = s * s;
  double p = 1.0 / 19.0;
  p = fma(p, z, 1.0 / 17.0);
  p = fma(p, z, 1.0 / 15.0);
  p = fma(p, z, 1.0 / 13.0);
  p = fma(p, z, 1.0 / 11.0);
  p = fma(p, z, 1.0 / 9.0);
  p = fma(p, z, 1.0 / 7.0);
  p = fma(p, z, 1.0 / 5.0);
  p = fma(p, z, 1.0 / 3.0);
  p = p * z;                                   // atanh(s)/s - 1
  const double l = fma(s, p, s);               // atanh(s)
  // log10(x) = e*log10(2) + 2*atanh(s)*log10(e)
  return fma((double)e, 0.30102999566398119521, l * 0.86858896380650365530);
}

// column_density.f90:351-376 with the fast reciprocal
__device__ __forceinline__ double weightf_fast(double cd, double sig) { return fast_rcp(fmax(0.6, cd * sig)); }

struct TauPos { int ipos; double residual; };

__device__ __forceinline__ TauPos tau_table_position(double tau) {  // :282-306
  const double lt = fast_log10(fmax(1.0e-20, tau));
  // odpos = min(NumTau, max(0, 1 + (lt - minlogtau)/dlogtau))
  const double odpos = fmin((double)NumTau, fmax(0.0, fma(lt - minlogtau, 1.0 / dlogtau, 1.0)));
  TauPos p;
  p.ipos = (int)odpos;
  p.residual = odpos - (double)p.ipos;
  return p;
}
// :310-326 ; ipos_p1 = min(NumTau, ipos+1): at ipos == NumTau the residual is 0, so re-reading row ipos is exact
__device__ __forceinline__ double read_table(const double* __restrict__ col, const TauPos& p) {
  const double a = __ldg(col + p.ipos), b = __ldg(col + min(NumTau, p.ipos + 1));
  return fma(b - a, p.residual, a);
}

struct PhotOut { double photo_HI, photo_HeI, photo_HeII, heat, photo_in, photo_out; };

// secondary ionisation (Ricotti et al. 2002), radiation_photoionrates.f90:49-55, :557-565
struct SecIon { double y1R0, y1R1, y1R2, y2R0, y2R1, y2R2; };
__device__ __forceinline__ SecIon secion_factors(double i_state) {
  SecIon y;
  y.y1R0 = 0.3908 * pow(1.0 - pow(i_state, 0.4092), 1.7592);
  y.y1R1 = 0.0554 * pow(1.0 - pow(i_state, 0.4614), 1.6660);
  y.y1R2 = 1.0 * pow(1.0 - pow(i_state, 0.2663), 1.3163);
  const double xeb01 = 1.0 - pow(i_state, 0.38);  // bR2(1) == bR2(2)
  const double xeb2 = 1.0 - pow(i_state, 0.34);
  const double p02 = pow(i_state, 0.2);            // aR2(1) == aR2(2)
  y.y2R0 = 0.6941 * p02 * xeb01 * xeb01;
  y.y2R1 = 0.0984 * p02 * xeb01 * xeb01;
  y.y2R2 = 3.9811 * pow(i_state, 0.4) * xeb2 * xeb2;
  return y;
}

// vol: the shell-cell volume the rates are diluted over; nflux: NormFlux, NormFluxPL, NormFluxQPL of the source.
template <bool ISO>
__device__ __forceinline__ PhotOut photoion_rates(double in_HI, double out_HI, double in_HeI, double out_HeI,
                                                  double in_HeII, double out_HeII, double vol, const double nflux[3],
                                                  const SecIon& y) {
  const double cell_HI = out_HI - in_HI, cell_HeI = out_HeI - in_HeI, cell_HeII = out_HeII - in_HeII;
  bool act[3];
  int blo = NumFreqBnd + 1, bhi = 0;
#pragma unroll
  for (int s = 0; s < 3; s++) {
    act[s] = (d_run.sed[s].hi >= d_run.sed[s].lo) && (nflux[s] > 0.0);
    if (act[s]) { blo = min(blo, d_run.sed[s].lo); bhi = max(bhi, d_run.sed[s].hi); }
  }
  // accumulators, all still to be multiplied by 1/vol
  double a_in = 0.0, a_out = 0.0, a_HI = 0.0, a_HeI = 0.0, a_HeII = 0.0;
  double f_heat = 0.0, f_ion_HI = 0.0, f_ion_HeI = 0.0;

  for (int b = blo; b <= bhi; b++) {  // 1-based band
    const int q = b - 1;
    const double sHI = d_band.sigma_HI[q], sHeI = d_band.sigma_HeI[q], sHeII = d_band.sigma_HeII[q];
    const double tau_in = in_HI * sHI + in_HeI * sHeI + in_HeII * sHeII;     // :172-176
    const double tau_out = out_HI * sHI + out_HeI * sHeI + out_HeII * sHeII;  // :179-183
    const double dtau = tau_out - tau_in;
    const bool thick_p = fabs(dtau) > tau_photo_limit;
    const bool thick_h = fabs(dtau) > tau_heat_limit;
    const TauPos pin = tau_table_position(tau_in);
    TauPos pout = pin;
    if (thick_p) pout = tau_table_position(tau_out);
    // species shares of the band's absorption (:787-825) and per-species cell optical depths (:236-240)
    const double tcHI = cell_HI * sHI, tcHeI = cell_HeI * sHeI, tcHeII = cell_HeII * sHeII;
    int nsp = 1, hcol = 0;
    double scHI = 1.0, scHeI = 0.0, scHeII = 0.0;
    if (b > NumBndin1 + NumBndin2) {
      const double f = fast_rcp(tcHI + tcHeI + tcHeII);
      scHI = tcHI * f; scHeI = tcHeI * f; scHeII = tcHeII * f;
      nsp = 3; hcol = 3 * b - NumBndin2 - NumBndin1 * 2 - 3;
    } else if (b > NumBndin1) {
      const double f = fast_rcp(tcHI + tcHeI);
      scHI = tcHI * f; scHeI = tcHeI * f;
      nsp = 2; hcol = 2 * b - NumBndin1 - 2;
    }
    double phot = 0.0;                                  // this band's absorbed photons, all SEDs
    double ph_HI = 0.0, ph_HeI = 0.0, ph_HeII = 0.0;    // this band's heating per species, all SEDs
#pragma unroll
    for (int s = 0; s < 3; s++) {
      if (!act[s] || b < d_run.sed[s].lo || b > d_run.sed[s].hi) continue;
      const SedDev& T = d_run.sed[s];
      const double NFlux = nflux[s];
      const size_t off = (size_t)q * (NumTau + 1);
      // photo_lookuptable :390-460
      const double phi_in = NFlux * read_table(T.photo_thick + off, pin);
      double phi_all, phi_out;
      if (thick_p) {
        phi_out = NFlux * read_table(T.photo_thick + off, pout);
        phi_all = phi_in - phi_out;
      } else {
        phi_all = NFlux * dtau * read_table(T.photo_thin + off, pin);
        phi_out = phi_in - phi_all;
      }
      a_in += phi_in;
      a_out += phi_out;
      phot += phi_all;
      // heat_lookuptable :586-760
      if (!ISO) {
        const double* ht = T.heat_thick + (size_t)hcol * (NumTau + 1);
        const double* hn = T.heat_thin + (size_t)hcol * (NumTau + 1);
        if (thick_h) {
          ph_HI += scHI * (NFlux * (read_table(ht, pin) - read_table(ht, pout)));
          if (nsp >= 2) ph_HeI += scHeI * (NFlux * (read_table(ht + (NumTau + 1), pin) - read_table(ht + (NumTau + 1), pout)));
          if (nsp == 3) ph_HeII += scHeII * (NFlux * (read_table(ht + 2 * (NumTau + 1), pin) - read_table(ht + 2 * (NumTau + 1), pout)));
        } else {
          ph_HI += NFlux * tcHI * read_table(hn, pin);
          if (nsp >= 2) ph_HeI += NFlux * tcHeI * read_table(hn + (NumTau + 1), pin);
          if (nsp == 3) ph_HeII += NFlux * tcHeII * read_table(hn + 2 * (NumTau + 1), pin);
        }
      }
    }
    a_HI = fma(scHI, phot, a_HI);
    a_HeI = fma(scHeI, phot, a_HeI);
    a_HeII = fma(scHeII, phot, a_HeII);
    if (!ISO) {
      // the secondary-ionisation bookkeeping is linear in the per-species heating, so the SED sum is taken first
      // (:654-669, :739-759)
      double df_heat = ph_HI + ph_HeI + ph_HeII;
      if (b > NumBndin1) {
        const double fs1 = d_band.f1ion_HI[q] * ph_HI + d_band.f1ion_HeI[q] * ph_HeI + d_band.f1ion_HeII[q] * ph_HeII;
        const double fs2 = d_band.f2ion_HI[q] * ph_HI + d_band.f2ion_HeI[q] * ph_HeI + d_band.f2ion_HeII[q] * ph_HeII;
        const double fs3 = d_band.f1heat_HI[q] * ph_HI + d_band.f1heat_HeI[q] * ph_HeI + d_band.f1heat_HeII[q] * ph_HeII;
        const double fs4 = d_band.f2heat_HI[q] * ph_HI + d_band.f2heat_HeI[q] * ph_HeI + d_band.f2heat_HeII[q] * ph_HeII;
        f_ion_HeI += y.y1R1 * fs1 - y.y2R1 * fs2;
        f_ion_HI += y.y1R0 * fs1 - y.y2R0 * fs2;
        df_heat = df_heat - y.y1R2 * fs3 + y.y2R2 * fs4;
      }
      f_heat += df_heat;
    }
  }
  const double rvol = fast_rcp(vol);
  PhotOut r;
  r.photo_in = a_in;
  r.photo_out = a_out;
  r.photo_HI = a_HI * rvol;
  r.photo_HeI = a_HeI * rvol;
  r.photo_HeII = a_HeII * rvol;
  r.heat = 0.0;
  if (!ISO) {
    r.heat = f_heat * rvol;
    r.photo_HI += f_ion_HI * rvol * (1.0 / (ion_freq_HI * hplanck));
    r.photo_HeI += f_ion_HeI * rvol * (1.0 / (ion_freq_HeI * hplanck));
  }
  return r;
}

}  // namespace c2
