// c2ray_photo.cuh -- radiation_photoionrates.f90:108-277 photoion_rates for one cell, streamed over the frequency
// bands (no 47-element work arrays), written first for instruction count -- the first profile showed 22.8 k
// instructions per source x cell update of which only a third were FP64 arithmetic, the rest IEEE division slow-path
// scaffolding, the general-purpose log10/pow, per-band divisions by the shell volume and scalar gathers from 16
// table columns per band -- and then for the footprint of its look-ups in the L1 and the constant cache.
//   * 1/vol is applied once per cell, not per band (all rates are linear in it)
//   * (log10(tau)-minlogtau)/dlogtau becomes one FMA with 1/dlogtau; the clamps of odpos move to the integer index
//   * log10 is a branch-free atanh-series evaluation valid for the positive normal arguments that occur here
//     (tau clamped to >= 1e-20), accurate to ~2e-16 relative, coefficients as constant-bank operands
//   * reciprocals use MUFU.RCP64H + one cubic step (|1 - b r| <= 2^-53 measured) without the IEEE slow path
//   * the secondary-ionisation factors y1R, y2R (9 pow per cell in the reference, :557-565) depend on the cell only:
//     they are computed once per iteration into the cell records (k_cell_records) and, being band independent factors
//     of sums that are linear in the per-band heating, applied once per cell after the band loop (photoion_finish)
//   * the band constants (3 cross sections, 12 f-factors) sit in one 128-byte record per band in constant memory
//   * a source with several SEDs is traced SED by SED (rates are linear in the SEDs), see photoion_bands
//   * the four tables of an SED are re-packed band-major into 32-byte rows, thick and thin values apart:
//       thick(band, itau) = [photo_thick, heat_thick(HI,HeI,HeII)]   thin(band, itau) = [photo_thin, heat_thin(HI,HeI,HeII)]
//     so one table position is two adjacent rows (64 contiguous bytes, 16-byte vector loads) instead of up to 8
//     scalar gathers from different columns, and the common optically thick cell never pulls thin values through
//     the L1; row NumTau+1 duplicates row NumTau (ipos_p1 = min(NumTau, ipos+1))
//   * the band loop is split by band group (1 / 26 / 20 sub-bands) with the species count as a template parameter
// All of this changes results at the 1e-15 level; the parity tests hold 1e-8.
#pragma once
#include "c2ray_physics.cuh"

#ifndef C2RAY_LD256
#define C2RAY_LD256 1
#endif
#ifndef C2RAY_TAUOUT_SUM
#define C2RAY_TAUOUT_SUM 1
#endif
#ifndef C2RAY_DEAD_BANDS
#define C2RAY_DEAD_BANDS 1
#endif

// tuning builds only: -DC2RAY_BAND_UNROLL=n unrolls the band loops
#ifdef C2RAY_BAND_UNROLL
#define C2_STR2(x) #x
#define C2_STR(x) C2_STR2(x)
#define C2_BAND_UNROLL _Pragma(C2_STR(unroll C2RAY_BAND_UNROLL))
#else
#define C2_BAND_UNROLL
#endif

namespace c2 {

// (the band records in shared memory instead of constant memory: 5 % slower at 16 sources, 5 % faster at 1000 --
// profiles/README.md)
#define BANDREC(q) d_band[q]

constexpr int PK_ROW = 8;                 // doubles per table position: 4 thick + 4 thin values
constexpr int PK_HALF = 4;                // doubles per packed row (thick rows and thin rows are separate arrays)
constexpr int PK_ROWS = NumTau + 2;       // rows per band (one duplicate at the end)
constexpr size_t PK_THIN_OFF = (size_t)NumFreqBnd * PK_ROWS * PK_HALF;  // the thin rows follow all thick rows
// isothermal runs read no heating values: a third copy holds photo_thick alone (8 bytes per table position, four
// times as many positions per cache line), then photo_thin alone
constexpr size_t PK_ISO_OFF = 2 * PK_THIN_OFF;
constexpr size_t PK_ISO_THIN_OFF = PK_ISO_OFF + (size_t)NumFreqBnd * PK_ROWS;
constexpr size_t PK_TOTAL = PK_ISO_THIN_OFF + (size_t)NumFreqBnd * PK_ROWS;  // doubles per SED

// column_density.f90:351-376 with the fast reciprocal
__device__ __forceinline__ double weightf_fast(double cd, double sig) { return fast_rcp(fmax(0.6, cd * sig)); }

// Dead bands.  The table build zeroes every integrand with tau * (nu/nu_min)^-index >= 700 (radiation_tables.f90:607), so
// beyond some row every entry of a band -- photo and heating, thick and thin -- is exactly 0.0 and the band contributes
// exactly nothing to any rate of a cell whose incoming optical depth lies there (tau_out >= tau_in reads zero rows as
// well).  d_dead[sed][band] is the optical depth of the row AFTER the first all-zero row of that SED's band (one row of
// margin against the rounding of the table position), +inf for a band without such rows.  A band step returns before
// its logarithms when tau_in >= d_dead: bit-identical results, and behind a thick neutral shell (an optical depth of
// 40-70 per cell at the hydrogen edge on the 256^3 cosmological box) most low-energy bands are dead.
__constant__ double d_dead[3][NumFreqBnd];

struct TauPos { int ipos; double residual; };

// radiation_photoionrates.f90:282-306: odpos = min(NumTau, max(0, 1+(log10(max(1e-20,tau))-minlogtau)/dlogtau)).
// With tau >= 1e-20 the lower clamp never binds; the upper one is applied to the integer index: for odpos > NumTau
// both interpolation rows are row NumTau, so the residual is irrelevant.
#ifndef C2RAY_TABLOG
#define C2RAY_TABLOG 1
#endif
// Table position through a 256-entry table of the leading mantissa byte, staged in shared memory by the kernel
// (postab_stage): x = 2^e m, m in [1,2), i = top 8 mantissa bits, c_i = 1 + (i + 1/2)/256, entry i holds
// r_i = fl(1/c_i) and P_i = 1 + (log10(1/r_i) + 20)/dlogtau; t = m r_i - 1 is one exact-product FMA with |t| <= 2^-9, and
// odpos = e log10(2)/dlogtau + P_i + log10(1+t)/dlogtau with a degree-6 series for log10(1+t) (next term 3e-18 rows).
// 12 FP64 operations per position instead of 29 with the table-free logarithm (no reciprocal, a shorter series, the
// shift and scale of :288 folded into the table entry).
__device__ double2 g_postab[256];
__constant__ double d_posc[7];  // [0] log10(2)/dlogtau  [1] log10(e)/dlogtau  [2..6] log10(e)/dlogtau * (-1/2, 1/3, -1/4, 1/5, -1/6)
__device__ __forceinline__ double2* postab_smem() {
  __shared__ double2 tab[256];
  return tab;
}
__device__ __forceinline__ void postab_stage() {
  double2* tab = postab_smem();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = g_postab[i];
  __syncthreads();
}
__device__ __forceinline__ TauPos tau_table_position(double tau) {
  static_assert(minlogtau == -20.0 && dlogtau == 24.0 / 2000.0, "d_lit[6] holds 1/dlogtau");
#if C2RAY_TABLOG
#ifndef C2RAY_POS_INTCLAMP
#define C2RAY_POS_INTCLAMP 1
#endif
#if C2RAY_POS_INTCLAMP
  // The clamp max(1e-20, tau) of :288 is applied to the result instead (an fmax on doubles sits 26 cycles in front of
  // everything else, tools/latency_probe.cu): tau <= 1e-20 (tau = 0 in the source cell) gives odpos < 1 -- hugely
  // negative but finite for 0 and denormals, whose exponent field reads as 2^-1023 -- and is put on row 1, residual 0.
  const double x = tau;
#else
  const double x = fmax(d_lit[3], tau);
#endif
  const int hi = __double2hiint(x);
  const double2 ent = postab_smem()[(hi >> 12) & 0xff];
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double ef = (double)((hi >> 20) - 1023);
  const double t = fma(m, ent.x, -1.0);
  const double t2 = t * t;
  const double q0 = fma(d_posc[3], t, d_posc[2]), q1 = fma(d_posc[5], t, d_posc[4]);
  const double q = fma(fma(d_posc[6], t2, q1), t2, q0);
  const double big = fma(ef, d_posc[0], ent.y);
  const double small = fma(t2, q, t * d_posc[1]);
  const double odpos = big + small;
#else
  const double lt = fast_log10(fmax(d_lit[3], tau));
  const double odpos = fma(lt - minlogtau, d_lit[6], 1.0);
#endif
  TauPos p;
#if C2RAY_TABLOG && C2RAY_POS_INTCLAMP
  const int raw = (int)odpos;
  p.ipos = min(max(raw, 1), NumTau);
  p.residual = raw < 1 ? 0.0 : odpos - (double)p.ipos;
#else
  p.ipos = min((int)odpos, NumTau);
  p.residual = odpos - (double)p.ipos;
#endif
  return p;
}

// (ld.global.nc.L1::evict_last on these loads was tried: no change at 16 sources, 8 % slower at 1000 -- profiles/README.md)
__device__ __forceinline__ double2 ld2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ double lerp(double a, double b, double t) { return fma(b - a, t, a); }  // :321-324
// one packed table row (32 bytes) in one 256-bit load (sm_100: LDG.E.256)
struct Row4 { double x, y, z, w; };
__device__ __forceinline__ Row4 ld4(const double* p) {
  Row4 r;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
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

// The same factors for the once-per-iteration full-grid pass (k_secion_factors): the six powers of i_state share one
// logarithm and the general-purpose pow (~200 instructions, 9 of them per cell) becomes exp(p ln x) without slow paths.
// i_state is in [epsilon, 1], so 1 - i_state^p is in [0, 1]; an exact 0 (fully ionised cell) must stay an exact 0.
__device__ __forceinline__ double pow01(double base, double p) { return base > 0.0 ? fast_exp(p * fast_log(base)) : 0.0; }
__device__ __forceinline__ SecIon secion_factors_fast(double i_state) {
  SecIon y;
  const double L = fast_log(i_state);
  y.y1R0 = 0.3908 * pow01(1.0 - fast_exp(0.4092 * L), 1.7592);
  y.y1R1 = 0.0554 * pow01(1.0 - fast_exp(0.4614 * L), 1.6660);
  y.y1R2 = 1.0 * pow01(1.0 - fast_exp(0.2663 * L), 1.3163);
  const double xeb01 = 1.0 - fast_exp(0.38 * L);
  const double xeb2 = 1.0 - fast_exp(0.34 * L);
  const double p02 = fast_exp(0.2 * L);
  y.y2R0 = 0.6941 * p02 * xeb01 * xeb01;
  y.y2R1 = 0.0984 * p02 * xeb01 * xeb01;
  y.y2R2 = 3.9811 * fast_exp(0.4 * L) * xeb2 * xeb2;
  return y;
}

struct PhotAcc {  // per-cell accumulators, still to be multiplied by 1/vol (except a_in, a_out)
  // s1..s4: band sums of f1ion.ph, f2ion.ph, f1heat.ph, f2heat.ph (ph = heating per species of a band).  The
  // secondary-ionisation bookkeeping (:654-669, :739-759) is linear in them with band-independent factors y1R, y2R,
  // so the factors are applied once per cell after the band loop instead of once per band.
  double a_in, a_out, a_HI, a_HeI, a_HeII, f_heat, s1, s2, s3, s4;
};

struct CellCols {
  double in_HI, in_HeI, in_HeII, out_HI, out_HeI, out_HeII, cell_HI, cell_HeI, cell_HeII;
};

// One frequency band (1-based b, NSP species absorb in it) for every active SED.
// MULTI = false: only the black-body SED exists in this run; NFlux is then factored out of the band loop.
template <bool ISO, int NSP, bool MULTI, bool NEED_IN = true, bool SCALED = false>
__device__ __forceinline__ void band_step(int b, const CellCols& c, const double nflux[3], unsigned actmask,
                                          PhotAcc& A, const double* __restrict__ pk_single = nullptr, int sed_single = 0,
                                          double nf_single = 1.0) {
  const int q = b - 1;
  const double sHI = BANDREC(q).sigma_HI;
  const double sHeI = NSP >= 2 ? BANDREC(q).sigma_HeI : 0.0;
  const double sHeII = NSP == 3 ? BANDREC(q).sigma_HeII : 0.0;
#if C2RAY_TAUOUT_SUM
  // tau_out as tau_in + the cell's own optical depth (the sum the species shares below need anyway): 3 FP64
  // operations fewer per band; differs from :172-183's direct sum by an ulp of tau_out
  const double tcHI = c.cell_HI * sHI, tcHeI = NSP >= 2 ? c.cell_HeI * sHeI : 0.0, tcHeII = NSP == 3 ? c.cell_HeII * sHeII : 0.0;
  double tau_in = c.in_HI * sHI;
  if (NSP >= 2) tau_in = fma(c.in_HeI, sHeI, tau_in);
  if (NSP == 3) tau_in = fma(c.in_HeII, sHeII, tau_in);
  const double dtau = NSP == 3 ? (tcHI + tcHeI + tcHeII) : (NSP == 2 ? tcHI + tcHeI : tcHI);
#if !defined(C2RAY_MULTI_SED_INNER) && C2RAY_DEAD_BANDS
  // (the single-SED kernels read the black body's threshold from the band record they have in cache anyway)
  // (the black body's threshold rides in the band record, the constant-cache line this step reads anyway)
  if (tau_in >= ((pk_single && sed_single != 0) ? d_dead[sed_single][q] : BANDREC(q).dead_bb)) return;  // every row this band would read is exactly zero
#endif
  const double tau_out = tau_in + dtau;
#else
  double tau_in = c.in_HI * sHI, tau_out = c.out_HI * sHI;             // :172-183
  if (NSP >= 2) { tau_in = fma(c.in_HeI, sHeI, tau_in); tau_out = fma(c.out_HeI, sHeI, tau_out); }
  if (NSP == 3) { tau_in = fma(c.in_HeII, sHeII, tau_in); tau_out = fma(c.out_HeII, sHeII, tau_out); }
  const double dtau = tau_out - tau_in;
  const double tcHI = c.cell_HI * sHI, tcHeI = c.cell_HeI * sHeI, tcHeII = c.cell_HeII * sHeII;
#endif
  const bool thick_p = fabs(dtau) > d_lit[4];  // tau_photo_limit, :342
  const bool thick_h = fabs(dtau) > d_lit[5];  // tau_heat_limit, :482
  // both positions unconditionally: the two log10 evaluations are independent and interleave (the thin branch, which
  // does not need pout, is the rare one)
  const TauPos pin = tau_table_position(tau_in);
  const TauPos pout = tau_table_position(tau_out);
  // species shares of the band's absorption (:787-825) and per-species cell optical depths (:236-240)
  double scHI = 1.0, scHeI = 0.0, scHeII = 0.0;
  if (NSP == 3) {
#if C2RAY_TAUOUT_SUM
    const double f = fast_rcp(dtau);
#else
    const double f = fast_rcp(tcHI + tcHeI + tcHeII);
#endif
    scHI = tcHI * f; scHeI = tcHeI * f; scHeII = tcHeII * f;
  } else if (NSP == 2) {
    const double f = fast_rcp(tcHI + tcHeI);
    scHI = tcHI * f; scHeI = tcHeI * f;
  }
  double phot = 0.0;                                  // absorbed photons of this band, all SEDs
  double ph_HI = 0.0, ph_HeI = 0.0, ph_HeII = 0.0;    // heating per species of this band, all SEDs
  const size_t row_in = ((size_t)q * PK_ROWS + pin.ipos) * PK_HALF;
  const size_t row_out = ((size_t)q * PK_ROWS + pout.ipos) * PK_HALF;
  // (a rolled loop over the SEDs was tried for the MULTI kernels: fewer registers, but 8 % slower at 4 CTAs/SM and
  // 27 % slower at 5 -- profiles/README.md)
#pragma unroll
  for (int s = 0; s < (MULTI ? 3 : 1); s++) {
    if (MULTI) {
      if (!((actmask >> s) & 1u) || b < d_run.sed[s].lo || b > d_run.sed[s].hi) continue;
    }
    const double* __restrict__ pk = MULTI ? d_run.sed[s].packed : (pk_single ? pk_single : d_run.sed[0].packed);
    const double NFlux = MULTI ? nflux[s] : (SCALED ? nf_single : 1.0);  // SCALED: one SED of a multi-SED source, its flux applied per band
    if (ISO) {  // photo_lookuptable :390-425 on the photo-only copy of the tables
      const double* pi = pk + PK_ISO_OFF + (size_t)q * PK_ROWS + pin.ipos;
      const double phi_in = NFlux * lerp(__ldg(pi), __ldg(pi + 1), pin.residual);
      double phi_all, phi_out;
      if (thick_p) {
        const double* po = pk + PK_ISO_OFF + (size_t)q * PK_ROWS + pout.ipos;
        phi_out = NFlux * lerp(__ldg(po), __ldg(po + 1), pout.residual);
        phi_all = phi_in - phi_out;
      } else {
        const double* pt = pi + (PK_ISO_THIN_OFF - PK_ISO_OFF);
        phi_all = NFlux * dtau * lerp(__ldg(pt), __ldg(pt + 1), pin.residual);
        phi_out = phi_in - phi_all;
      }
      if (NEED_IN) A.a_in += phi_in;
      A.a_out += phi_out;
      phot += phi_all;
      continue;
    }
    const double* ri = pk + row_in;
    const double* ro = pk + row_out;
    const double* ti = ri + PK_THIN_OFF;  // thin rows at tau_in
#if C2RAY_LD256
    if (NSP >= 2) {
      // rows i and i+1 of the thick table at tau_in: [photo_thick, heat_thick HI, HeI, HeII] each, one 256-bit load per row
      const Row4 i0 = ld4(ri), i1 = ld4(ri + PK_HALF);
      const double phi_in = NFlux * lerp(i0.x, i1.x, pin.residual);  // photo_lookuptable :390-396
      double phi_all, phi_out;
      Row4 o0 = i0, o1 = i1;
      if (thick_p) {
        o0 = ld4(ro); o1 = ld4(ro + PK_HALF);
        phi_out = NFlux * lerp(o0.x, o1.x, pout.residual);
        phi_all = phi_in - phi_out;
      } else {
        const double thin = lerp(__ldg(ti), __ldg(ti + PK_HALF), pin.residual);
        phi_all = NFlux * dtau * thin;
        phi_out = phi_in - phi_all;
      }
      if (NEED_IN) A.a_in += phi_in;
      A.a_out += phi_out;
      phot += phi_all;
      if (thick_h) {
        ph_HI = fma(scHI, NFlux * (lerp(i0.y, i1.y, pin.residual) - lerp(o0.y, o1.y, pout.residual)), ph_HI);
        ph_HeI = fma(scHeI, NFlux * (lerp(i0.z, i1.z, pin.residual) - lerp(o0.z, o1.z, pout.residual)), ph_HeI);
        if (NSP == 3) ph_HeII = fma(scHeII, NFlux * (lerp(i0.w, i1.w, pin.residual) - lerp(o0.w, o1.w, pout.residual)), ph_HeII);
      } else {
        const Row4 t0 = ld4(ti), t1 = ld4(ti + PK_HALF);
        ph_HI = fma(NFlux * tcHI, lerp(t0.y, t1.y, pin.residual), ph_HI);
        ph_HeI = fma(NFlux * tcHeI, lerp(t0.z, t1.z, pin.residual), ph_HeI);
        if (NSP == 3) ph_HeII = fma(NFlux * tcHeII, lerp(t0.w, t1.w, pin.residual), ph_HeII);
      }
      continue;
    }
#endif
    // thick values at tau_in: [photo_thick, heat_thick HI | heat_thick HeI, heat_thick HeII]
    const double2 i0a = ld2(ri), i1a = ld2(ri + PK_HALF);
    const double phi_in = NFlux * lerp(i0a.x, i1a.x, pin.residual);  // photo_lookuptable :390-396
    double phi_all, phi_out;
    double2 o0a = i0a, o1a = i1a;
    if (thick_p) {
      o0a = ld2(ro); o1a = ld2(ro + PK_HALF);  // (skipping these loads when pout.ipos == pin.ipos: 7 % slower)
      phi_out = NFlux * lerp(o0a.x, o1a.x, pout.residual);
      phi_all = phi_in - phi_out;
    } else {
      const double thin = lerp(__ldg(ti), __ldg(ti + PK_HALF), pin.residual);
      phi_all = NFlux * dtau * thin;
      phi_out = phi_in - phi_all;
    }
    if (NEED_IN) A.a_in += phi_in;
    A.a_out += phi_out;
    phot += phi_all;
    if (!ISO) {  // heat_lookuptable :586-760
      if (thick_h) {
        ph_HI = fma(scHI, NFlux * (lerp(i0a.y, i1a.y, pin.residual) - lerp(o0a.y, o1a.y, pout.residual)), ph_HI);
        if (NSP >= 2) {
          const double2 i0b = ld2(ri + 2), i1b = ld2(ri + PK_HALF + 2), o0b = ld2(ro + 2), o1b = ld2(ro + PK_HALF + 2);
          ph_HeI = fma(scHeI, NFlux * (lerp(i0b.x, i1b.x, pin.residual) - lerp(o0b.x, o1b.x, pout.residual)), ph_HeI);
          if (NSP == 3)
            ph_HeII = fma(scHeII, NFlux * (lerp(i0b.y, i1b.y, pin.residual) - lerp(o0b.y, o1b.y, pout.residual)), ph_HeII);
        }
      } else {
        // thin rows at tau_in: [photo_thin, heat_thin HI | heat_thin HeI, heat_thin HeII]
        const double2 t0a = ld2(ti), t1a = ld2(ti + PK_HALF);
        ph_HI = fma(NFlux * tcHI, lerp(t0a.y, t1a.y, pin.residual), ph_HI);
        if (NSP >= 2) {
          const double2 t0b = ld2(ti + 2), t1b = ld2(ti + PK_HALF + 2);
          ph_HeI = fma(NFlux * tcHeI, lerp(t0b.x, t1b.x, pin.residual), ph_HeI);
          if (NSP == 3) ph_HeII = fma(NFlux * tcHeII, lerp(t0b.y, t1b.y, pin.residual), ph_HeII);
        }
      }
    }
  }
  A.a_HI = fma(scHI, phot, A.a_HI);                  // :428-456
  if (NSP >= 2) A.a_HeI = fma(scHeI, phot, A.a_HeI);
  if (NSP == 3) A.a_HeII = fma(scHeII, phot, A.a_HeII);
  if (!ISO) {
    // the secondary-ionisation bookkeeping is linear in the per-species heating, so the SED sum is taken first
    // (:654-669, :739-759)
    double df_heat = ph_HI + ph_HeI + ph_HeII;
    if (NSP >= 2) {
      double fs1 = BANDREC(q).f1ion_HI * ph_HI + BANDREC(q).f1ion_HeI * ph_HeI;
      double fs2 = BANDREC(q).f2ion_HI * ph_HI + BANDREC(q).f2ion_HeI * ph_HeI;
      double fs3 = BANDREC(q).f1heat_HI * ph_HI + BANDREC(q).f1heat_HeI * ph_HeI;
      double fs4 = BANDREC(q).f2heat_HI * ph_HI + BANDREC(q).f2heat_HeI * ph_HeI;
      if (NSP == 3) {
        fs1 = fma(BANDREC(q).f1ion_HeII, ph_HeII, fs1);
        fs2 = fma(BANDREC(q).f2ion_HeII, ph_HeII, fs2);
        fs3 = fma(BANDREC(q).f1heat_HeII, ph_HeII, fs3);
        fs4 = fma(BANDREC(q).f2heat_HeII, ph_HeII, fs4);
      }
      A.s1 += fs1; A.s2 += fs2; A.s3 += fs3; A.s4 += fs4;
    }
    A.f_heat += df_heat;
  }
}

// vol: the shell-cell volume the rates are diluted over; nflux: NormFlux, NormFluxPL, NormFluxQPL of the source.
// The band loop: everything of photoion_rates that does not need the cell's secondary-ionisation factors.  Returns the
// accumulators and the flux scale still to be applied (photoion_finish).
// LANES > 1: the bands of a cell are dealt to LANES adjacent lanes of a warp (lane_j = 0..LANES-1 takes every
// LANES-th band of each band group); the caller sums the accumulators over those lanes (reduce_bands).
template <bool ISO, bool MULTI, int LANES = 1, bool NEED_IN = true>
__device__ __forceinline__ PhotAcc photoion_bands(double in_HI, double out_HI, double in_HeI, double out_HeI,
                                                  double in_HeII, double out_HeII, const double nflux[3], double& scale_out,
                                                  int lane_j = 0) {
  CellCols c;
  c.in_HI = in_HI; c.in_HeI = in_HeI; c.in_HeII = in_HeII;
  c.out_HI = out_HI; c.out_HeI = out_HeI; c.out_HeII = out_HeII;
  c.cell_HI = out_HI - in_HI; c.cell_HeI = out_HeI - in_HeI; c.cell_HeII = out_HeII - in_HeII;  // :167-169
#ifndef C2RAY_MULTI_SED_INNER
  if (MULTI) {
    // Every rate is linear in the SEDs' contributions and the optical depths do not depend on the SED, so a source
    // with several SEDs is traced as the sum of single-SED sources: for each SED that the source emits in, the lean
    // single-SED band loop over that SED's own band range (NormFlux factored out), then one multiply-add per
    // accumulator.  The reference's order -- sum over the SEDs inside every band (:390-456) -- costs predicated
    // copies of the table look-ups in every band for a sum that usually has one term (the black body ends at band
    // 33-36, PL / QPL start at 38).
    PhotAcc T = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#ifndef C2RAY_MULTI_SCALED
#define C2RAY_MULTI_SCALED 1
#endif
#pragma unroll 1
    for (int s = 0; s < 3; s++) {
      const int lo = d_run.sed[s].lo, hi = d_run.sed[s].hi;
#if C2RAY_MULTI_SCALED
      // (warp-uniform decision, so that the band loops stay in convergent code -- k_sweep_shell; a lane whose source does
      // not emit in this SED runs along with flux 0 and adds exact zeros)
      const bool on = hi >= lo && nflux[s] > 0.0;
      if (!__any_sync(0xffffffffu, on)) continue;
      const double nf = on ? nflux[s] : 0.0;
#else
      const double nf = nflux[s];
      if (hi < lo || !(nf > 0.0)) continue;
#endif
      const double* __restrict__ pk = d_run.sed[s].packed;
#if C2RAY_MULTI_SCALED
      // the SED's flux is applied inside the band step (five multiplications per band) instead of to a second set of ten
      // accumulators afterwards: 20 registers less in a kernel that spills
      if (lo <= NumBndin1 && lane_j == 0) band_step<ISO, 1, false, NEED_IN, true>(1, c, nflux, 1u, T, pk, s, nf);
      for (int b = max(lo, NumBndin1 + 1) + lane_j; b <= min(hi, NumBndin1 + NumBndin2); b += LANES)
        band_step<ISO, 2, false, NEED_IN, true>(b, c, nflux, 1u, T, pk, s, nf);
      for (int b = max(lo, NumBndin1 + NumBndin2 + 1) + lane_j; b <= hi; b += LANES)
        band_step<ISO, 3, false, NEED_IN, true>(b, c, nflux, 1u, T, pk, s, nf);
#else
      PhotAcc B = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      if (lo <= NumBndin1 && lane_j == 0) band_step<ISO, 1, false, NEED_IN>(1, c, nflux, 1u, B, pk, s);
      for (int b = max(lo, NumBndin1 + 1) + lane_j; b <= min(hi, NumBndin1 + NumBndin2); b += LANES)
        band_step<ISO, 2, false, NEED_IN>(b, c, nflux, 1u, B, pk, s);
      for (int b = max(lo, NumBndin1 + NumBndin2 + 1) + lane_j; b <= hi; b += LANES)
        band_step<ISO, 3, false, NEED_IN>(b, c, nflux, 1u, B, pk, s);
      T.a_in = fma(nf, B.a_in, T.a_in); T.a_out = fma(nf, B.a_out, T.a_out);
      T.a_HI = fma(nf, B.a_HI, T.a_HI); T.a_HeI = fma(nf, B.a_HeI, T.a_HeI); T.a_HeII = fma(nf, B.a_HeII, T.a_HeII);
      if (!ISO) {
        T.f_heat = fma(nf, B.f_heat, T.f_heat);
        T.s1 = fma(nf, B.s1, T.s1); T.s2 = fma(nf, B.s2, T.s2); T.s3 = fma(nf, B.s3, T.s3); T.s4 = fma(nf, B.s4, T.s4);
      }
#endif
    }
    scale_out = 1.0;
    return T;
  }
#endif
  unsigned act = 1u;  // bit s: SED s exists and this source emits in it
  int blo, bhi;
  double scale = 1.0;
  unsigned long long bands = ~0ull;  // bit b-1: some active SED covers band b (BB and QPL ranges can leave a gap)
  if (MULTI) {
    blo = NumFreqBnd + 1; bhi = 0; bands = 0ull;
#pragma unroll
    for (int s = 0; s < 3; s++) {
      const bool on = (d_run.sed[s].hi >= d_run.sed[s].lo) && (nflux[s] > 0.0);
      if (s == 0) act = 0u;
      if (on) {
        act |= 1u << s;
        blo = min(blo, d_run.sed[s].lo); bhi = max(bhi, d_run.sed[s].hi);
        bands |= (~0ull >> (64 - d_run.sed[s].hi)) & (~0ull << (d_run.sed[s].lo - 1));
      }
    }
  } else {
    blo = d_run.sed[0].lo; bhi = d_run.sed[0].hi;
    scale = nflux[0];
    // (scale == 0: every rate comes out as 0 x finite = 0, as :207 `if (NormFlux(nsrc) > 0.0)` would leave it)
  }
  PhotAcc A = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (blo <= NumBndin1 && bhi >= 1 && lane_j == 0) band_step<ISO, 1, MULTI, NEED_IN>(1, c, nflux, act, A);
  C2_BAND_UNROLL
  for (int b = max(blo, NumBndin1 + 1) + lane_j; b <= min(bhi, NumBndin1 + NumBndin2); b += LANES)
    if (!MULTI || ((bands >> (b - 1)) & 1ull)) band_step<ISO, 2, MULTI, NEED_IN>(b, c, nflux, act, A);
  C2_BAND_UNROLL
  for (int b = max(blo, NumBndin1 + NumBndin2 + 1) + lane_j; b <= bhi; b += LANES)
    if (!MULTI || ((bands >> (b - 1)) & 1ull)) band_step<ISO, 3, MULTI, NEED_IN>(b, c, nflux, act, A);
  scale_out = scale;
  return A;
}

// Sum of the band accumulators over the LANES lanes that shared a cell (mask: exactly those lanes).
template <bool ISO, int LANES>
__device__ __forceinline__ void reduce_bands(PhotAcc& A, unsigned mask) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
    A.a_in += __shfl_xor_sync(mask, A.a_in, o);
    A.a_out += __shfl_xor_sync(mask, A.a_out, o);
    A.a_HI += __shfl_xor_sync(mask, A.a_HI, o);
    A.a_HeI += __shfl_xor_sync(mask, A.a_HeI, o);
    A.a_HeII += __shfl_xor_sync(mask, A.a_HeII, o);
    if (!ISO) {
      A.f_heat += __shfl_xor_sync(mask, A.f_heat, o);
      A.s1 += __shfl_xor_sync(mask, A.s1, o);
      A.s2 += __shfl_xor_sync(mask, A.s2, o);
      A.s3 += __shfl_xor_sync(mask, A.s3, o);
      A.s4 += __shfl_xor_sync(mask, A.s4, o);
    }
  }
}

template <bool ISO>
__device__ __forceinline__ PhotOut photoion_finish(const PhotAcc& A, double scale, double vol, const SecIon& y) {
  const double rvol = fast_rcp(vol) * scale;
  PhotOut r;
  r.photo_in = A.a_in * scale;
  r.photo_out = A.a_out * scale;
  r.photo_HI = A.a_HI * rvol;
  r.photo_HeI = A.a_HeI * rvol;
  r.photo_HeII = A.a_HeII * rvol;
  r.heat = 0.0;
  if (!ISO) {
    r.heat = (A.f_heat - y.y1R2 * A.s3 + y.y2R2 * A.s4) * rvol;                                    // :669, :759
    r.photo_HI += (y.y1R0 * A.s1 - y.y2R0 * A.s2) * rvol * (1.0 / (ion_freq_HI * hplanck));        // :773-777
    r.photo_HeI += (y.y1R1 * A.s1 - y.y2R1 * A.s2) * rvol * (1.0 / (ion_freq_HeI * hplanck));
  }
  return r;
}

template <bool ISO, bool MULTI>
__device__ __forceinline__ PhotOut photoion_rates(double in_HI, double out_HI, double in_HeI, double out_HeI,
                                                  double in_HeII, double out_HeII, double vol, const double nflux[3],
                                                  const SecIon& y) {
  double scale;
  const PhotAcc A = photoion_bands<ISO, MULTI>(in_HI, out_HI, in_HeI, out_HeI, in_HeII, out_HeII, nflux, scale);
  return photoion_finish<ISO>(A, scale, vol, y);
}

// Re-pack the four (0:NumTau, 1:nb) tables of one SED into band-major 32-byte thick and thin rows.  One thread per (band, row).
__global__ void k_pack_tables(const double* __restrict__ photo_thick, const double* __restrict__ photo_thin,
                              const double* __restrict__ heat_thick, const double* __restrict__ heat_thin,
                              double* __restrict__ packed) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NumFreqBnd * PK_ROWS) return;
  const int q = t / PK_ROWS, row = t % PK_ROWS, it = min(row, NumTau), b = q + 1;
  const int nsp = (b <= NumBndin1) ? 1 : (b <= NumBndin1 + NumBndin2 ? 2 : 3);
  const int hcol = (b <= NumBndin1) ? 0 : (b <= NumBndin1 + NumBndin2 ? 2 * b - NumBndin1 - 2 : 3 * b - NumBndin2 - NumBndin1 * 2 - 3);
  double v[PK_ROW] = {0, 0, 0, 0, 0, 0, 0, 0};
  v[0] = photo_thick[(size_t)q * (NumTau + 1) + it];
  v[4] = photo_thin[(size_t)q * (NumTau + 1) + it];
  if (heat_thick && heat_thin)
    for (int sp = 0; sp < nsp; sp++) {
      v[1 + sp] = heat_thick[(size_t)(hcol + sp) * (NumTau + 1) + it];
      v[5 + sp] = heat_thin[(size_t)(hcol + sp) * (NumTau + 1) + it];
    }
  for (int k = 0; k < PK_HALF; k++) {
    packed[(size_t)t * PK_HALF + k] = v[k];
    packed[PK_THIN_OFF + (size_t)t * PK_HALF + k] = v[PK_HALF + k];
  }
  packed[PK_ISO_OFF + t] = v[0];
  packed[PK_ISO_THIN_OFF + t] = v[PK_HALF];
}

// First all-zero row of every band of one SED's tables -> optical depth from which the band is dead (see d_dead).
// One block per band; tables as the host holds them: (0:NumTau, 1:nb), band-major.
__global__ void k_band_dead(const double* __restrict__ photo_thick, const double* __restrict__ photo_thin,
                            const double* __restrict__ heat_thick, const double* __restrict__ heat_thin,
                            double* __restrict__ dead_tau) {
  const int q = blockIdx.x, b = q + 1;
  const int nsp = (b <= NumBndin1) ? 1 : (b <= NumBndin1 + NumBndin2 ? 2 : 3);
  const int hcol = (b <= NumBndin1) ? 0 : (b <= NumBndin1 + NumBndin2 ? 2 * b - NumBndin1 - 2 : 3 * b - NumBndin2 - NumBndin1 * 2 - 3);
  __shared__ int last_nonzero;
  if (threadIdx.x == 0) last_nonzero = -1;
  __syncthreads();
  int mine = -1;
  for (int it = threadIdx.x; it <= NumTau; it += blockDim.x) {
    bool nz = photo_thick[(size_t)q * (NumTau + 1) + it] != 0.0 || photo_thin[(size_t)q * (NumTau + 1) + it] != 0.0;
    if (heat_thick && heat_thin)
      for (int sp = 0; sp < nsp; sp++)
        nz = nz || heat_thick[(size_t)(hcol + sp) * (NumTau + 1) + it] != 0.0 || heat_thin[(size_t)(hcol + sp) * (NumTau + 1) + it] != 0.0;
    if (nz) mine = max(mine, it);
  }
  atomicMax(&last_nonzero, mine);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int first_zero = last_nonzero + 1, row = first_zero + 1;   // one row of margin
    // row i >= 1 of the table is tau = 10^(minlogtau + dlogtau (i-1)) (radiation_tables.f90:181-186)
    dead_tau[q] = row <= NumTau ? pow(10.0, minlogtau + dlogtau * (double)(row - 1)) : __longlong_as_double(0x7ff0000000000000LL);
  }
}

}  // namespace c2
