/* A compiled-language host of libc2ray_b200.so: what the Fortran side does through iso_c_binding, written in C because
 * the image has no Fortran compiler.  It owns plain host arrays in the reference's memory layout
 * (column-major A(i,j,k,c), srcpos(3,NumSrc) 1-based, temperature_grid real(si)), and drives one time step exactly as
 * fortran/c2ray_b200_iso_c.F90 does: init -> cooling curves -> rad_ini -> sources -> geometry -> evolve3d_host -> rates.
 *
 *   evolve3d_driver <problem.bin> <result.bin>
 *
 * problem.bin (written by tests/test_gpu_c_driver.py, little endian):
 *   int32  mesh[3], NumSrc, isothermal, cosmological, subboxsize, max_subbox
 *   double temper_val, H0, Omega0, clumping, T_eff, S_star, dr[3], vol, zred, dt
 *   double logT[801], logLambda[5][801]
 *   int32  srcpos[NumSrc][3] ; double NormFlux[NumSrc]
 *   double ndens[N3], xh[2][N3], xhe[3][N3] ; float temperature_grid[3][N3]
 * result.bin: int32 niter, conv_flag ; int64 rt_updates ; double photon_loss_all ; xh, xhe (double), temperature_grid
 *   (float), phih_grid (double)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "c2ray_b200.h"

static void die(const char* what, int rc) {
  fprintf(stderr, "evolve3d_driver: %s failed (%d): %s\n", what, rc, c2ray_b200_last_error());
  exit(1);
}
#define CHECK(call) do { int rc_ = (call); if (rc_ != C2RAY_OK) die(#call, rc_); } while (0)
static void rd(void* p, size_t n, FILE* f) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s problem.bin result.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  int32_t hdr[8];
  double sc[14];
  rd(hdr, sizeof hdr, f);
  rd(sc, sizeof sc, f);
  const int32_t mesh[3] = {hdr[0], hdr[1], hdr[2]}, NumSrc = hdr[3];
  const size_t N3 = (size_t)mesh[0] * mesh[1] * mesh[2];
  double* logT = malloc(801 * sizeof(double));
  double* logL = malloc(5 * 801 * sizeof(double));
  rd(logT, 801 * sizeof(double), f);
  rd(logL, 5 * 801 * sizeof(double), f);
  int32_t* srcpos = malloc(3 * (size_t)NumSrc * sizeof(int32_t));
  double* nflux = malloc((size_t)NumSrc * sizeof(double));
  rd(srcpos, 3 * (size_t)NumSrc * sizeof(int32_t), f);
  rd(nflux, (size_t)NumSrc * sizeof(double), f);
  double* ndens = malloc(N3 * sizeof(double));
  double* xh = malloc(2 * N3 * sizeof(double));
  double* xhe = malloc(3 * N3 * sizeof(double));
  float* temp = malloc(3 * N3 * sizeof(float));
  rd(ndens, N3 * sizeof(double), f);
  rd(xh, 2 * N3 * sizeof(double), f);
  rd(xhe, 3 * N3 * sizeof(double), f);
  rd(temp, 3 * N3 * sizeof(float), f);
  fclose(f);

  c2ray_params par;
  memset(&par, 0, sizeof par);
  par.isothermal = hdr[4]; par.cosmological = hdr[5]; par.subboxsize = hdr[6]; par.max_subbox = hdr[7];
  par.temper_val = sc[0]; par.H0 = sc[1]; par.Omega0 = sc[2]; par.clumping = (float)sc[3];
  par.max_slots = 0; par.deterministic = 1;
  c2ray_ctx* ctx = NULL;
  CHECK(c2ray_b200_init(&par, mesh, -1, &ctx));
  CHECK(c2ray_b200_set_cooling_tables(ctx, logT, logL));
  c2ray_sed_params sed;
  memset(&sed, 0, sizeof sed);
  sed.T_eff = sc[4]; sed.S_star = sc[5];
  sed.pl_index = 1.0; sed.pl_minfreq = 1.0; sed.pl_maxfreq = 2.0; sed.pl_S_star = 0.0;      /* S_star <= 0: SED absent */
  sed.qpl_index = 1.0; sed.qpl_minfreq = 1.0; sed.qpl_maxfreq = 2.0; sed.qpl_S_star = 0.0;
  CHECK(c2ray_b200_rad_ini(ctx, &sed));
  CHECK(c2ray_b200_set_sources(ctx, NumSrc, srcpos, nflux, NULL, NULL));
  const double dr[3] = {sc[6], sc[7], sc[8]};
  CHECK(c2ray_b200_set_geometry(ctx, dr, sc[9], sc[10]));
  c2ray_stats st;
  CHECK(c2ray_b200_evolve3d_host(ctx, 0.0, sc[11], 0, ndens, xh, xhe, par.isothermal ? NULL : temp, &st));
  double* phih = malloc(N3 * sizeof(double));
  double* phihe = malloc(2 * N3 * sizeof(double));
  double* phiheat = malloc(N3 * sizeof(double));
  CHECK(c2ray_b200_get_rates(ctx, phih, phihe, phiheat));
  /* an out-of-range restart flag must be refused, not ignored */
  if (c2ray_b200_evolve3d(ctx, 0.0, sc[11], 7, NULL) != C2RAY_ERR_ARG) { fprintf(stderr, "restart=7 accepted\n"); return 3; }
  CHECK(c2ray_b200_destroy(ctx));

  f = fopen(argv[2], "wb");
  if (!f) { perror(argv[2]); return 2; }
  fwrite(&st.niter, 4, 1, f); fwrite(&st.conv_flag, 4, 1, f); fwrite(&st.rt_updates, 8, 1, f); fwrite(&st.photon_loss_all, 8, 1, f);
  fwrite(xh, sizeof(double), 2 * N3, f); fwrite(xhe, sizeof(double), 3 * N3, f); fwrite(temp, sizeof(float), 3 * N3, f);
  fwrite(phih, sizeof(double), N3, f);
  fclose(f);
  printf("evolve3d_driver: mesh %d %d %d, %d sources, niter %d, conv_flag %d, %lld updates\n", mesh[0], mesh[1], mesh[2], NumSrc,
         st.niter, st.conv_flag, (long long)st.rt_updates);
  return 0;
}
