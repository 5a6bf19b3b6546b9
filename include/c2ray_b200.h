/* c2ray_b200.h -- C ABI of the B200-native C2-Ray H+He hot path (libc2ray_b200.so).
 *
 * The reference (garrelt/C2-Ray3Dm1D_Helium) has no FFI: its seam is Fortran module procedures plus
 * `use`-associated module arrays.  Each entry point below names the reference routine it replaces
 * (file:line under code/), so a Fortran host keeps `evolve3D(time,dt,restart)` / `do_source(dt,ns1,niter)`
 * and forwards to these through iso_c_binding (see INTEGRATION.md and fortran/c2ray_b200_iso_c.F90).
 *
 * Conventions: plain pointers and sizes only.  Grid arrays are Fortran column-major A(i,j,k[,c]) --
 * i fastest, component slowest -- exactly as the reference's allocatables lie in memory; srcpos is
 * (3,NumSrc), 1-based.  All host pointers are owned by the caller and only copied from/to.
 * Every function returns 0 on success, a negative code otherwise; c2ray_b200_last_error() returns a
 * message.  There is no CPU fallback: without a CUDA device every compute entry point fails with
 * C2RAY_ERR_CUDA.
 *
 * Threading (the reference's routines are not re-entrant either: module globals, cgsconstants.f90:105-133): one
 * context per process and device, driven from one host thread -- one MPI rank / one torchrun rank per GPU.  The
 * per-run constants of the active context live in device __constant__ memory; two contexts alive on the same
 * device must not have work in flight at the same time.  c2ray_b200_last_error() is per host thread.
 * Arguments that become device indices are range-checked at the boundary: srcpos must lie in 1..mesh(d), the
 * FreqBnd limits of uploaded tables in 1..47 (C2RAY_ERR_ARG otherwise).
 */
#ifndef C2RAY_B200_H
#define C2RAY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C2RAY_NUMFREQBND 47   /* radiation_sizes.f90:22 */
#define C2RAY_NUMHEATBIN 113  /* radiation_sizes.f90:23 */
#define C2RAY_NUMTAU 2000     /* radiation_sizes.f90:18 ; tables are (0:NumTau, 1:nbands) */
#define C2RAY_MAX_ITER_HIST 512

#define C2RAY_OK 0
#define C2RAY_ERR_ARG (-1)
#define C2RAY_ERR_CUDA (-2)
#define C2RAY_ERR_STATE (-3)
#define C2RAY_ERR_NCCL (-4)

typedef struct c2ray_ctx c2ray_ctx;

/* Run-time values of the compile-time parameters of c2ray_parameters.f90:26-89 and of the material /
 * cosmology module scalars the hot path reads. */
typedef struct c2ray_params {
  int32_t isothermal;    /* material: isothermal (mat_ini_test.F90) */
  int32_t cosmological;  /* c2ray_parameters.f90:84 -> thermal.f90:74 */
  int32_t subboxsize;    /* c2ray_parameters.f90:51 (TEST4: mesh(1)) */
  int32_t max_subbox;    /* c2ray_parameters.f90:56 */
  double temper_val;     /* isothermal temperature (mat_ini_test.F90:168) */
  double H0;             /* cosmoparms.f90 H0 (cgs) -> cosmology.f90:229 */
  double Omega0;         /* cosmoparms.f90 */
  float clumping;        /* material: clumping (real), type_of_clumping=1 */
  int32_t max_slots;     /* sources traced concurrently on the device (0 = default 1024) */
  int32_t deterministic; /* 1: one source at a time in source order (rate-grid sums in the reference's order) */
} c2ray_params;

/* Nominal-SED parameters of radiation_sed_parameters.f90:208-244 / sed_parameters.f90.  S_star <= 0 disables
 * the PL / QPL SED (the -DPL / -DQUASARS build switches). */
typedef struct c2ray_sed_params {
  double T_eff, S_star;
  double pl_index, pl_minfreq, pl_maxfreq, pl_S_star;
  double qpl_index, qpl_minfreq, qpl_maxfreq, qpl_S_star;
} c2ray_sed_params;

/* The radiation_tables.f90 module arrays for one SED: *_photo_thick_table(0:NumTau,1:NumFreqBnd) etc.,
 * column-major, plus its FreqBnd_LowerLimit / UpperLimit (1-based) and S_star.  heat_* may be NULL when
 * isothermal.  photo_thick == NULL marks the SED as absent. */
typedef struct c2ray_sed_tables {
  const double* photo_thick;
  const double* photo_thin;
  const double* heat_thick;
  const double* heat_thin;
  int32_t freqbnd_lower, freqbnd_upper;
  double S_star;
} c2ray_sed_tables;

typedef struct c2ray_stats {
  int32_t niter;           /* evolve.F90:185 */
  int32_t conv_flag;       /* evolve.F90:163 last global_pass result */
  int32_t conv_criterion;  /* evolve.F90:147 */
  int32_t nit_max;         /* largest do_chemistry iteration count of the last global pass */
  int64_t sum_nbox_all;    /* evolve.F90:544 (last iteration) */
  int64_t rt_updates;      /* source x cell updates done in this call */
  int64_t chem_cells;      /* cells passed through evolve0D_global in this call */
  int64_t nit_total;       /* sum of do_chemistry iterations of the last global pass */
  double photon_loss_all;  /* photon_loss_all(1), evolve.F90:423/511 */
  double ms_sweep, ms_chem, ms_allreduce, ms_total; /* CUDA-event device times of this call */
  double sums_before[5];   /* photonstatistics.f90:117 state_before: H0,H+,He0,He+,He++ */
  double sums_after[5];    /* photonstatistics.f90:208 state_after */
  /* photonstatistics.f90:150-270 for the finished step: total_rates with (xh_av, xhe_av), total_ionizations,
   * the source photon budget and the photon-conservation ratio written to PhotonCounts.out (:286-298) */
  double totrec, totcollisions, recomions, total_ion, totalsrc, photcons;
  int32_t conv_hist[C2RAY_MAX_ITER_HIST]; /* conv_flag after iteration i+1 */
} c2ray_stats;

const char* c2ray_b200_last_error(void);

/* ---- life cycle ------------------------------------------------------------------------------------ */
/* evolve_data.F90:74 evolve_ini + mpi.F90:83 mpi_setup (device selection).  device < 0: use LOCAL_RANK or 0. */
int c2ray_b200_init(const c2ray_params* params, const int32_t mesh[3], int32_t device, c2ray_ctx** out);
int c2ray_b200_destroy(c2ray_ctx* ctx);
int c2ray_b200_set_params(c2ray_ctx* ctx, const c2ray_params* params);

/* cooling_h.f90:76 setup_cool: logT[801] and 5 x log10(Lambda)[801] (H0, H1 caseB, He0, He1, He2) as read
 * from the tables/ directory (H0-cool.tab ...); converted to linear (10**x) on upload as the reference does (:163-169). */
int c2ray_b200_set_cooling_tables(c2ray_ctx* ctx, const double* logT, const double* logLambda5x801);

/* radiation_tables.f90:141 rad_ini executed on the device (table build kernel).  Alternatively the host's own
 * rad_ini output can be passed with c2ray_b200_upload_tables (sed index 0=BB "B", 1=PL "P", 2=QPL "Q"). */
int c2ray_b200_rad_ini(c2ray_ctx* ctx, const c2ray_sed_params* sed);
int c2ray_b200_upload_tables(c2ray_ctx* ctx, int32_t sed, const c2ray_sed_tables* tables);
int c2ray_b200_download_table(c2ray_ctx* ctx, int32_t sed, int32_t kind /*0 photo_thick 1 photo_thin 2 heat_thick 3 heat_thin*/,
                              double* out, int32_t* lower, int32_t* upper, double* S_star);

/* sourceprops: NumSrc, srcpos(3,NumSrc), NormFlux(1:NumSrc) [, NormFluxPL, NormFluxQPL] (may be NULL). */
int c2ray_b200_set_sources(c2ray_ctx* ctx, int32_t NumSrc, const int32_t* srcpos, const double* NormFlux,
                           const double* NormFluxPL, const double* NormFluxQPL);
/* grid: dr(3), vol (proper, change every step: cosmology.f90:159) ; cosmology: zred */
int c2ray_b200_set_geometry(c2ray_ctx* ctx, const double dr[3], double vol, double zred);

/* material, type_of_clumping == 5 (c2ray_parameters.f90:67): clumping_grid(N3) real, read per cell by do_chemistry
 * (evolve_point.F90:484 clumping_point) and by total_rates (photonstatistics.f90:176).  NULL: back to the scalar
 * c2ray_params.clumping. */
int c2ray_b200_set_clumping_grid(c2ray_ctx* ctx, const float* clumping_grid);
/* material, use_LLS (c2ray_parameters.f90:72-77): type_of_LLS 0 none | 1 coldensh_LLS for every cell | 2 LLS_grid(N3)
 * real (already scaled to a column density per cell, mat_ini_test.F90:742-743).  evolve0D adds
 * coldensh_LLS*path/dr(1) to the incoming HI column density of every cell but the source's (evolve_point.F90:177-180).
 * The LLS_loss counter (:277) is not kept: the reference feeds it photo_in_HI, a member photoion_rates never sets. */
int c2ray_b200_set_LLS(c2ray_ctx* ctx, int32_t type_of_LLS, double coldensh_LLS, const float* LLS_grid);
/* material: ndens(N3), xh(N3,0:1), xhe(N3,0:2), temperature_grid(N3,0:2) real(si) (NULL when isothermal) */
int c2ray_b200_set_state(c2ray_ctx* ctx, const double* ndens, const double* xh, const double* xhe,
                         const float* temperature_grid);
int c2ray_b200_get_state(c2ray_ctx* ctx, double* xh, double* xhe, float* temperature_grid);
/* evolve_data: phih_grid(N3), phihe_grid(N3,0:1), phiheat(N3) ; xh_av, xhe_av, xh_intermed, xhe_intermed */
int c2ray_b200_get_rates(c2ray_ctx* ctx, double* phih_grid, double* phihe_grid, double* phiheat);
int c2ray_b200_set_rates(c2ray_ctx* ctx, const double* phih_grid, const double* phihe_grid, const double* phiheat);
int c2ray_b200_get_work_state(c2ray_ctx* ctx, double* xh_av, double* xhe_av, double* xh_intermed, double* xhe_intermed);
int c2ray_b200_set_work_state(c2ray_ctx* ctx, const double* xh_av, const double* xhe_av, const double* xh_intermed,
                              const double* xhe_intermed);
/* device-resident snapshot / restore of (xh, xhe, temperature_grid): lets a benchmark repeat one time step
 * without host traffic. */
int c2ray_b200_snapshot_state(c2ray_ctx* ctx);
int c2ray_b200_restore_state(c2ray_ctx* ctx);

/* ---- the hot path ---------------------------------------------------------------------------------- */
/* evolve.F90:78 evolve3D(time,dt,restart) on the device-resident state.  restart = 0: a fresh time step;
 * 1, 2, 3: resume the iteration from iterdump1.bin, iterdump2.bin, iterdump.bin in the dump directory (:137-141).
 * With a communicator attached the global pass is split over the ranks (reduce-scatter of the rate grids, 1/npr of
 * the cells per rank, all-gather of the fractions the next sweep reads) unless C2RAY_SPLIT_CHEM=0; results are those
 * of the reference's allreduce + replicated pass. */
int c2ray_b200_evolve3d(c2ray_ctx* ctx, double time, double dt, int32_t restart, c2ray_stats* stats);
/* The drop-in for the Fortran evolve3D body: H2D of the module arrays, evolve3D, D2H of the results. */
int c2ray_b200_evolve3d_host(c2ray_ctx* ctx, double time, double dt, int32_t restart, const double* ndens,
                             double* xh, double* xhe, float* temperature_grid, c2ray_stats* stats);

/* evolve.F90:130-136 (restart==0 initialisation of xh_av, xh_intermed, ...) */
int c2ray_b200_begin_step(c2ray_ctx* ctx);
/* evolve.F90:371 set_rates_to_zero */
int c2ray_b200_set_rates_to_zero(c2ray_ctx* ctx);
/* evolve.F90:385 pass_all_sources -> master_slave.F90:74 do_grid_static -> evolve_source.F90:66 do_source, for the
 * sources of this rank (ns1 = 1+rank, NumSrc, npr), followed by evolve.F90:505 mpi_accumulate_grid_quantities
 * (NCCL allreduce when a communicator is attached). */
int c2ray_b200_pass_all_sources(c2ray_ctx* ctx, double dt, int32_t niter, int64_t* rt_updates);
/* evolve_source.F90:66 do_source(dt,ns1,niter) for one source (1-based); nbox returns the sub-box count. */
int c2ray_b200_do_source(c2ray_ctx* ctx, double dt, int32_t ns1, int32_t niter, int32_t* nbox, double* photon_loss_src);
/* evolve.F90:435 global_pass(conv_flag,dt): evolve_point.F90:325 evolve0D_global over the whole mesh.
 * nit_out (N3 int32, may be NULL) receives do_chemistry's iteration count per cell. */
int c2ray_b200_global_pass(c2ray_ctx* ctx, double dt, int32_t* conv_flag, int32_t* nit_out);
/* evolve.F90:164-166 copy-back at convergence */
int c2ray_b200_end_step(c2ray_ctx* ctx);
/* photonstatistics.f90:117/:208 : sums of ndens*x over the mesh times vol*abundance for (xh, xhe) [which=0]
 * or (xh_intermed, xhe_intermed) [which=1] */
int c2ray_b200_state_sums(c2ray_ctx* ctx, int32_t which, double out5[5]);

/* ---- fine-grained parity hooks (batches of independent cells, host pointers) ------------------------ */
/* radiation_photoionrates.f90:108 photoion_rates: col6[n][6] = in_HI,out_HI,in_HeI,out_HeI,in_HeII,out_HeII;
 * out6[n][6] = photo_cell_HI, photo_cell_HeI, photo_cell_HeII, heat, photo_in, photo_out */
int c2ray_b200_photoion_rates_batch(c2ray_ctx* ctx, int32_t n, const double* col6, const double* vol,
                                    const double nflux3[3], const double* i_state, double* out6);
/* evolve_point.F90:444 do_chemistry (local=.false.): doric x2 + thermal to convergence for n independent cells.
 * ion15[n][15] = h(0:1) he(0:2) h_av(0:1) he_av(0:2) h_old(0:1) he_old(0:2); phi4[n][4] = HI,HeI,HeII,heat;
 * T3[n][3] = (T_inter, T_avg, T_old) ; nit_out[n] */
int c2ray_b200_chemistry_batch(c2ray_ctx* ctx, int32_t n, double dt, const double* ndens, double* ion15,
                               const double* phi4, double* T3, int32_t* nit_out);
/* doric.f90:35 doric(dt,rhe,rhh,ion,phi,yfrac,zfrac,y2afrac,y2bfrac) for n independent states, with the module
 * coefficients set by ini_rec_colion_factors(T[i]) and material's scalar clumping: rhe[n], ion15[n][15] in/out
 * (h, he, h_av, he_av overwritten; h_old, he_old read), phi3[n][3] = photo_cell_HI, HeI, HeII, fr4[n][4] = yfrac,
 * zfrac, y2afrac, y2bfrac (doric.f90:317 prepare_doric_factors). */
int c2ray_b200_doric_batch(c2ray_ctx* ctx, int32_t n, double dt, const double* rhe, double* ion15, const double* phi3,
                           const double* fr4, const double* T);
/* thermal.f90:22 thermal(dt,end_temper,avg_temper,ndens_electron,ndens_atom,ion,phi) for n independent states:
 * end_temper[n] in/out, avg_temper[n] in/out (left untouched when end_temper <= minitemp, thermal.f90:83), heat[n] =
 * phi%heat, nsub[n] = explicit sub-steps taken (thermal.f90:98-157). */
int c2ray_b200_thermal_batch(c2ray_ctx* ctx, int32_t n, double dt, double* end_temper, double* avg_temper,
                             const double* ndens_electron, const double* ndens_atom, const double* ion15, const double* heat,
                             int32_t* nsub);
/* cgsconstants.f90:140 ini_rec_colion_factors: out12 = arech0,brech0,areche0,breche0,oreche0,areche1,breche1,
 * treche1,colli_HI,colli_HeI,colli_HeII,v */
int c2ray_b200_rec_colion_batch(c2ray_ctx* ctx, int32_t n, const double* T, double* out12);
/* column_density.f90:28 cinterp for n target cells against a caller-provided full-grid scratch (coldensh_out,
 * coldenshe_out(:,:,:,0:1), Fortran layout): pos[n][3] unwrapped, srcpos[3]; out4[n][4] = cdensi, he0, he1, path */
int c2ray_b200_cinterp_batch(c2ray_ctx* ctx, int32_t n, const int32_t* pos, const int32_t srcpos[3],
                             const double* coldensh_out, const double* coldenshe_out, double* out4);

/* ---- iteration dumps, restart, output streams (Fortran form="unformatted" sequential files) --------- */
/* evolve.F90:199-213: during evolve3d rank 0 writes <dump_dir>/iterdump1.bin and iterdump2.bin alternately whenever
 * more than interval_s seconds have passed since the call started or the last dump (the reference: 15*60).
 * interval_s < 0 disables the dumps.  dump_dir is also where evolve3d(restart=1|2|3) looks for iterdump1.bin |
 * iterdump2.bin | iterdump.bin (evolve.F90:279 start_from_dump; every rank reads the file itself). */
int c2ray_b200_set_dump(c2ray_ctx* ctx, const char* dump_dir, double interval_s);
/* evolve.F90:233 write_iteration_dump(niter) to an explicit path: records niter, photon_loss_all(47), phih_grid,
 * xh_av, xh_intermed, phihe_grid, xhe_av, xhe_intermed [, phiheat, temperature_grid(real(si))] */
int c2ray_b200_write_iteration_dump(c2ray_ctx* ctx, const char* path, int32_t niter);
/* evolve.F90:279 start_from_dump: loads the records above into the device-resident arrays, returns niter */
int c2ray_b200_read_iteration_dump(c2ray_ctx* ctx, const char* path, int32_t* niter);
/* output.F90:249 write_stream2: <results_dir>/xfrac3d_<z>.bin, xfrac3dHe1_<z>.bin, xfrac3dHe2_<z>.bin with
 * <z> = trim(adjustl(f6.3 of zred_now)); each file: record mesh(1:3) int32, record N3 real(dp).  Rank 0 only. */
int c2ray_b200_write_stream2(c2ray_ctx* ctx, const char* results_dir, double zred_now);
/* output.F90:312 write_stream3: Temper3D_<z>.bin (temperature_grid(:,:,:,0), not when isothermal),
 * IonRates3D_<z>.bin (real(phih_grid)), HeatRates3D_<z>.bin (real(phiheat)); N3 real(si) each.  Rank 0 only. */
int c2ray_b200_write_stream3(c2ray_ctx* ctx, const char* results_dir, double zred_now);
/* The record layer itself on host memory (no device, no context): n records data[i] of bytes[i] bytes.  Records
 * above max_subrecord bytes (<= 0: the compilers' 2^31-9) are split into subrecords with signed markers as gfortran
 * and ifort do.  read fails unless every record has exactly the expected length. */
int c2ray_b200_fortran_records_write(const char* path, int32_t n, const void* const* data, const int64_t* bytes,
                                     int64_t max_subrecord);
int c2ray_b200_fortran_records_read(const char* path, int32_t n, void* const* data, const int64_t* bytes);

/* mrgrnk.f90 R_mrgrnk(XVALT, IRNGT) as ctrper.f90:108-113 uses it: irngt[i] = 1-based position of the (i+1)-th smallest
 * key, equal keys in input order.  Bit-exact against the merge sort; off the hot path (the reference computes the
 * source permutation but never uses it, evolve_source.F90:88-90). */
int c2ray_b200_mrgrnk(c2ray_ctx* ctx, int32_t n, const float* xvalt, int32_t* irngt);

/* ---- multi-GPU (mpi.F90 my_mpi: rank, npr ; evolve.F90:505-548 allreduce) ---------------------------- */
/* 128-byte NCCL unique id created on rank 0 and broadcast by the host (MPI_BCAST / torch.distributed). */
int c2ray_b200_comm_unique_id(uint8_t id[128]);
int c2ray_b200_comm_init(c2ray_ctx* ctx, const uint8_t id[128], int32_t rank, int32_t npr);
/* rank / npr without a communicator: source partition only (used by tests and by hosts that reduce themselves) */
int c2ray_b200_set_rank(c2ray_ctx* ctx, int32_t rank, int32_t npr);
/* Which rank traces which source.  0: the reference's static model, ns1 = 1+rank, NumSrc, npr (master_slave.F90:74-96
 * do_grid_static), the default.  1: balanced -- the stand-in for the master/slave model (master_slave.F90:124-326): after
 * every pass the ranks exchange each source's sub-box count and the sources are re-dealt, most expensive first, to the
 * least loaded rank (needs a communicator; the first pass is the round-robin deal).  Results differ from schedule 0 only
 * in the order the rate grids are summed. */
int c2ray_b200_set_source_schedule(c2ray_ctx* ctx, int32_t mode);
/* the deal itself, on host data: owner[i] = rank of source i given cost[i] (deterministic; every rank computes it) */
int c2ray_b200_balanced_partition(int32_t NumSrc, const int64_t* cost, int32_t npr, int32_t* owner);
/* 1-based numbers of the sources this rank traces in the next pass (ids may be NULL to query the count) */
int c2ray_b200_my_sources(c2ray_ctx* ctx, int32_t* ids, int32_t cap, int32_t* n);
/* device pointer + element count of the contiguous [phih | phihe(0) | phihe(1) | phiheat | photon_loss(47) |
 * sum_nbox] FP64 buffer, for hosts that run their own reduction on it */
int c2ray_b200_rates_device_buffer(c2ray_ctx* ctx, void** dptr, int64_t* count);

/* ---- measurement helpers ----------------------------------------------------------------------------- */
/* global-pass-only microbenchmark on the resident state (BASELINE config 5): runs `reps` global passes, each
 * from the same start state, returns average device ms per pass */
int c2ray_b200_bench_global_pass(c2ray_ctx* ctx, double dt, int32_t reps, double* ms_per_pass, int32_t* conv_flag);
/* kernels launched by this context since init (bench.py's gpu_launches) */
int64_t c2ray_b200_launch_count(c2ray_ctx* ctx);
/* launches of the ray-tracing kernel alone (the dominant kernel: bench.py's mean launch duration) */
int64_t c2ray_b200_sweep_launch_count(c2ray_ctx* ctx);
/* FP64 FMA throughput microbenchmark (TFLOP/s), for the FP64 roofline denominator */
int c2ray_b200_measure_fp64(c2ray_ctx* ctx, double* tflops);
/* CUDA-event stopwatch on the context's stream (device time of everything enqueued between the two calls) */
int c2ray_b200_timer_start(c2ray_ctx* ctx);
int c2ray_b200_timer_stop(c2ray_ctx* ctx, double* ms);
/* CUDA stream handle (cudaStream_t) the context launches on */
int c2ray_b200_stream(c2ray_ctx* ctx, void** stream);

#ifdef __cplusplus
}
#endif
#endif
