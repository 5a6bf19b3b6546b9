// c2ray_consts.cuh -- physical constants and compile-time sizes of the C2-Ray H+He hot path (sm_100a).
//
// The reference is compiled WITHOUT -r8 (code/files_for_3D/Makefile:43), so every literal written without
// `_dp`/`d0` is a binary32 value promoted to real(dp).  FL(x) reproduces that: (double)(float)x.
// Sources: code/mathconstants.f90:21, abundances.f90:23-29, atomic.f90, cgsconstants.f90:26-103,
// cgsphotoconstants.f90:25-50, c2ray_parameters.f90:26-89, radiation_sizes.f90:17-23,
// radiation_tables.f90:59-61, files_for_3D/column_density.f90:53-54.
#pragma once

#define FL(x) ((double)(float)(x))

namespace c2 {

constexpr double pi = FL(3.141592654f);
constexpr double abu_he = FL(0.074f);
constexpr double abu_c = FL(7.1e-7f);
constexpr double abu_h = 1.0 - abu_he;  // (1.0_dp-abu_he) as used on the path
constexpr double gamma_ad = 5.0 / 3.0;
constexpr double gamma1 = gamma_ad - 1.0;
constexpr double c_light = 2.997925e+10;
constexpr double hplanck = 6.6260755e-27;
constexpr double k_B = 1.381e-16;
constexpr double ev2k = FL(1.0f / 8.617e-05f);
constexpr double ev2fr = FL(0.241838e15f);
constexpr double two_pi_over_c_square = FL(2.0f) * pi / (c_light * c_light);
constexpr double eth0 = FL(13.598f);
constexpr double temph0 = eth0 * ev2k;
constexpr double colh0 = FL(1.3e-8f) * FL(0.83f) * FL(1.0f) / (eth0 * eth0);
constexpr double ethe0 = FL(24.587f), ethe1 = FL(54.416f);
constexpr double temphe0 = ethe0 * ev2k, temphe1 = ethe1 * ev2k;
constexpr double colhe0 = FL(1.3e-8f) * FL(0.63f) * FL(2.0f) / (ethe0 * ethe0);
constexpr double colhe1 = FL(1.3e-8f) * FL(1.30f) * FL(1.0f) / (ethe1 * ethe1);
constexpr double sigma_HI_at_ion_freq = FL(6.346e-18f);
constexpr double sigma_HeI_at_ion_freq = FL(7.430e-18f);
constexpr double sigma_HeII_at_ion_freq = FL(1.589e-18f);
constexpr double ion_freq_HI = ev2fr * eth0;
constexpr double ion_freq_HeI = ev2fr * ethe0;
constexpr double ion_freq_HeII = ev2fr * ethe1;
constexpr double sigma_H_heth = 1.238e-18;
constexpr double sigma_H_heLya = 9.907e-22;
constexpr double sigma_He_heLya = 1.301e-20;
constexpr double sigma_He_he2 = 1.690780687052975e-18;
constexpr double sigma_H_he2 = 1.230695924714239e-19;
constexpr double R_SOLAR = FL(6.9599e10f);

constexpr double epsilon = 1.0e-20;
constexpr double convergence_fraction = FL(2.5e-4f);
constexpr double minimum_fractional_change = FL(1.0e-2f);
constexpr double minimum_fraction_of_atoms = FL(1.0e-8f);
constexpr double minitemp = FL(1.0f);
constexpr double relative_denergy = FL(0.1f);
constexpr double max_coldensh = FL(2e29f);          // evolve_point.F90:91
constexpr double loss_fraction = FL(1e-10f);        // evolve_source.F90:136
constexpr double tau_photo_limit = FL(1.0e-7f);     // radiation_photoionrates.f90:342
constexpr double tau_heat_limit = FL(1.0e-4f);      // radiation_photoionrates.f90:482
constexpr double sqrt3 = 1.7320507764816284;        // real(sqrt(3.0)) -> dp, column_density.f90:53
constexpr double sqrt2 = 1.4142135381698608;        // real(sqrt(2.0)) -> dp, column_density.f90:54

constexpr int NumFreq = 512, NumTau = 2000, NumBndin1 = 1, NumBndin2 = 26, NumBndin3 = 20;
constexpr int NumFreqBnd = NumBndin1 + NumBndin2 + NumBndin3;          // 47
constexpr int NumheatBin = NumBndin1 + NumBndin2 * 2 + NumBndin3 * 3;  // 113
constexpr double minlogtau = -20.0, maxlogtau = 4.0;
constexpr double dlogtau = (maxlogtau - minlogtau) / 2000.0;
constexpr int TEMPPOINTS = 801;  // cooling_h.f90:25

}  // namespace c2
