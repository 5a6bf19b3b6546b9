// c2ray_api.cu -- host side of libc2ray_b200.so: context, device residency, the evolve3D iteration
// (code/files_for_3D/evolve.F90:78-229), the source loop (master_slave.F90:74-96, evolve_source.F90:66-238) as
// batched shell-wavefront launches, the device rad_ini (radiation_tables.f90:141-168), the NCCL rate-grid
// reduction (evolve.F90:505-548) and the C ABI declared in include/c2ray_b200.h.
//
// No CPU fallback: every compute entry point needs a CUDA device and fails with C2RAY_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/c2ray_b200.h"
#include "band_data.h"
#include "c2ray_io.h"
#include "c2ray_kernels.cuh"

using namespace c2;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      cudaGetLastError(); /* reported here: do not leave it for a later, unrelated cudaGetLastError() */ \
      return fail(C2RAY_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                      std::to_string(__LINE__) + ")");                                   \
    }                                                                                                    \
  } while (0)

#define NC_(call)                                                                              \
  do {                                                                                         \
    int r_ = (call);                                                                           \
    if (r_ != 0) return fail(C2RAY_ERR_NCCL, std::string(#call) + " failed: " + std::to_string(r_)); \
  } while (0)

// ---- NCCL through dlopen (no link-time dependency; torch's bundled libnccl is reused when already loaded) ----
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(nccl_comm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
} g_nccl;

int nccl_load() {
  if (g_nccl.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return fail(C2RAY_ERR_NCCL, std::string("dlopen libnccl.so.2 failed: ") + dlerror());
  g_nccl.GetUniqueId = (int (*)(nccl_uid*))dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(nccl_comm*, int, nccl_uid, int))dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t))dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.ReduceScatter = (int (*)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t))dlsym(g_nccl.h, "ncclReduceScatter");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm, cudaStream_t))dlsym(g_nccl.h, "ncclAllGather");
  g_nccl.GroupStart = (int (*)())dlsym(g_nccl.h, "ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())dlsym(g_nccl.h, "ncclGroupEnd");
  g_nccl.CommDestroy = (int (*)(nccl_comm))dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.ReduceScatter ||
      !g_nccl.AllGather || !g_nccl.GroupStart || !g_nccl.GroupEnd)
    return fail(C2RAY_ERR_NCCL, "libnccl: missing symbols");
  return 0;
}
constexpr int NCCL_INT32 = 2, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2;

}  // namespace

constexpr int MAX_SWEEP_GROUPS = 8;
constexpr int SPLIT_LANES = 8;  // lanes sharing a cell in latency-bound sweep launches

struct c2ray_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  c2ray_params par{};
  int mesh[3] = {0, 0, 0};
  size_t N3 = 0;
  // device-resident state (Fortran layout)
  double *ndens = nullptr, *xh = nullptr, *xhe = nullptr, *xh_av = nullptr, *xhe_av = nullptr, *xh_int = nullptr,
         *xhe_int = nullptr, *rates = nullptr;
  float* temp = nullptr;
  double *snap_xh = nullptr, *snap_xhe = nullptr;
  float* snap_temp = nullptr;
  size_t rates_count = 0;  // 4*N3 + 47 + 1
  // sources
  int NumSrc = 0;
  int* d_srcpos = nullptr;
  double *d_nf = nullptr, *d_nfpl = nullptr, *d_nfqpl = nullptr;
  int n_mine = 0;
  bool have_pl_flux = false, have_qpl_flux = false;
  double sum_nf[3] = {0, 0, 0};  // sum(NormFlux), sum(NormFluxPL), sum(NormFluxQPL) in source order
  // radiation tables
  double* tab[3][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
  int lo[3] = {1, 1, 1}, hi[3] = {0, 0, 0};
  double S_star[3] = {0, 0, 0};
  double* packed[3] = {nullptr, nullptr, nullptr};
  TableBuild* d_tb = nullptr;
  // cooling
  double* d_cool = nullptr;
  double cool_mintemp = 1.0, cool_dtemp = 0.01;
  bool have_cool = false;
  // geometry / cosmology
  double dr[3] = {1, 1, 1}, vol = 1, zred = 0;
  // sweep work space
  Slot* d_slots = nullptr;
  int* d_active = nullptr;
  SweepTotals* d_tot = nullptr;
  SweepTotals* d_gtot = nullptr;            // per stream group
  int* h_nact = nullptr;                    // pinned: active sources per group after a sub-box decision
  cudaStream_t gstream[MAX_SWEEP_GROUPS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_SWEEP_GROUPS] = {};
  int sweep_groups = 2;                     // env C2RAY_SWEEP_GROUPS
  int sweep_split = 1;                      // env C2RAY_SWEEP_SPLIT: band-split kernel for launches that cannot fill the GPU
  int sweep_lanes_mode = 1;                 // env C2RAY_SWEEP_LANES_MODE: 1 graded 2..16 lanes per cell, 0 the 8-lanes-or-1 rule
  double sweep_lanes_fill = 1.0;            // env C2RAY_SWEEP_LANES_FILL: resident waves the split threads may fill
  int sparse_records = 1;                   // env C2RAY_SPARSE_RECORDS: per-level cell records when the sources cover little of the mesh
  long long last_pass_updates = -1;         // updates of this rank's previous pass (-1: none yet)
  int dead_bands = 1;                       // env C2RAY_DEAD_BANDS: skip bands whose table rows are all zero from tau_in on (c2ray_photo.cuh)
  int sweep_pdl = 1;                        // env C2RAY_SWEEP_PDL: programmatic dependent launch between the shells of a level
  double* d_scratch = nullptr;
  double* d_lossbuf = nullptr;              // deterministic mode only
  int lossbuf_cap = 0;
  int slots_cap = 0;
  bool slots_budget_limited = false;
  SweepGeom geom{};
  ChemTotals* d_chem = nullptr;
  double* d_sums = nullptr;
  double* d_cellrec = nullptr;  // per-cell sweep inputs, rebuilt every iteration (k_cell_records)
  unsigned long long* d_next_cell = nullptr;
  BandRec h_band[NumFreqBnd];  // host copy of the band records in constant memory
  int n_sm = 148;          // multiprocessors of this context's device
  int chemq_per_sm = 0;    // resident CTAs per SM of k_global_pass_q (queried once per context)
  int chem_thermal_min = 0;     // env C2RAY_CHEM_TMIN: lanes that must wait in THERMAL before the queue kernel runs a burst (0: any)
  int chem_burst = CHEM_BURST;  // env C2RAY_CHEM_BURST: thermal sub-steps per turn of the queue kernel's state machine
  double last_nit_per_cell = 0.0;  // do_chemistry iterations per cell of the previous global pass
  int chem_mode = -1;      // -1 auto, 0 one cell per thread, 1 queue-driven (env C2RAY_CHEM_QUEUE overrides)
  double last_nsub_per_cell = 0.0;  // thermal sub-steps per cell of the previous global pass
  int* d_nit = nullptr;
  // multi-GPU
  int rank = 0, npr = 1;
  nccl_comm comm = nullptr;
  int schedule = 0;            // 0: do_grid_static round robin (master_slave.F90:85); 1: balanced by last pass's cost
  int* d_nbox_all = nullptr;   // sub-box count per source of the last pass (this rank's sources; summed over ranks)
  std::vector<int> my_ids;     // 0-based ids of this rank's sources
  std::vector<char> src_pl, src_qpl;  // per source: NormFluxPL > 0, NormFluxQPL > 0
  std::vector<double> src_flux;       // per source: total photon rate (evolve_source.F90:122-128), the balanced schedule's first cost model
  int* d_srcids_run = nullptr; // this rank's sources ordered for the sweep: black-body-only sources first
  int n_single = 0;            // how many of them take the single-SED kernel
  int run_key = -1;            // what d_srcids_run was built for (table presence bits); -1: rebuild
  int split_chem = 1;          // env C2RAY_SPLIT_CHEM: evolve3d splits the global pass over the ranks (see split_active)
  double* d_chemred = nullptr; // 4 sums (FP64) + 1 maximum (int32) of the global-pass counters, for the cross-rank combination
  // material: position-dependent clumping (type_of_clumping == 5) and Lyman-limit systems (use_LLS)
  float* d_clump = nullptr;
  float* d_lls = nullptr;
  int lls_type = 0;
  double coldensh_LLS = 0.0;
  // iteration dumps (evolve.F90:233-367) and output streams (output.F90:249-379)
  std::string dump_dir;
  double dump_interval_s = -1.0;  // < 0: no dumps; the reference writes one when 15 minutes have passed (:207)
  int ndump = 0;                  // evolve.F90:239, saved between calls
  void* h_stage = nullptr;        // pinned staging buffer for device <-> file traffic
  // bookkeeping
  int64_t launches = 0;
  int64_t sweep_launches = 0;  // k_sweep_* launches only (bench.py: mean duration of the dominant kernel)
  bool run_dirty = true;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_timer[2] = {nullptr, nullptr};
};

namespace {

c2ray_ctx* g_bound = nullptr;  // context whose RunConst is currently in __constant__ memory

#define LAUNCH(ctx, kernel, grid, block, ...)                   \
  do {                                                          \
    kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__); \
    (ctx)->launches++;                                          \
  } while (0)

int bind(c2ray_ctx* c) {
  CK(cudaSetDevice(c->device));
  if (g_bound == c && !c->run_dirty) return 0;
  RunConst rc;
  memset(&rc, 0, sizeof(rc));
  for (int s = 0; s < 3; s++) {
    rc.sed[s].photo_thick = c->tab[s][0]; rc.sed[s].photo_thin = c->tab[s][1];
    rc.sed[s].heat_thick = c->tab[s][2]; rc.sed[s].heat_thin = c->tab[s][3];
    rc.sed[s].packed = c->packed[s];
    rc.sed[s].lo = c->lo[s]; rc.sed[s].hi = c->tab[s][0] ? c->hi[s] : c->lo[s] - 1;
    rc.sed[s].S_star = c->S_star[s];
  }
  rc.isothermal = c->par.isothermal; rc.cosmological = c->par.cosmological;
  rc.temper_val = c->par.temper_val;
  rc.clumping = (double)c->par.clumping;
  // cosmology.f90:229  dzdt=H0*(1.+zred)*sqrt(Omega0*(1.+zred)**3+1.-Omega0)
  const double one = FL(1.0f);
  const double zp1 = one + c->zred;
  rc.zp1 = zp1;
  rc.dzdt = c->par.H0 * zp1 * sqrt(c->par.Omega0 * (zp1 * zp1 * zp1) + one - c->par.Omega0);
  rc.cosmo_coef = 0;
  for (int d = 0; d < 3; d++) { rc.dr[d] = c->dr[d]; rc.mesh[d] = c->mesh[d]; }
  rc.vol = c->vol;
  rc.cool_mintemp = c->cool_mintemp; rc.cool_dtemp = c->cool_dtemp; rc.cool_rdtemp = 1.0 / c->cool_dtemp; rc.cool = c->d_cool;
  CK(cudaMemcpyToSymbolAsync(d_run, &rc, sizeof(rc), 0, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));  // rc is a stack object
  g_bound = c;
  c->run_dirty = false;
  return 0;
}

int upload_band_const(c2ray_ctx* c) {
  BandRec* bc = c->h_band;
  memset(bc, 0, sizeof(c->h_band));
  bc[0].sigma_HI = sigma_HI_at_ion_freq;  // radiation_sizes.f90:381-383
  for (int i = 0; i < 26; i++) {
    const int q = NumBndin1 + i;
    bc[q].sigma_HI = BD_SIGMA_HI_B2[i]; bc[q].sigma_HeI = BD_SIGMA_HEI_B2[i]; bc[q].sigma_HeII = 0.0;
    bc[q].f1ion_HI = BD_F1ION_HI_B2[i]; bc[q].f1ion_HeI = BD_F1ION_HEI_B2[i]; bc[q].f1ion_HeII = BD_F1ION_HEII_B2[i];
    bc[q].f2ion_HI = BD_F2ION_HI_B2[i]; bc[q].f2ion_HeI = BD_F2ION_HEI_B2[i]; bc[q].f2ion_HeII = BD_F2ION_HEII_B2[i];
    bc[q].f1heat_HI = BD_F1HEAT_HI_B2[i]; bc[q].f1heat_HeI = BD_F1HEAT_HEI_B2[i]; bc[q].f1heat_HeII = BD_F1HEAT_HEII_B2[i];
    bc[q].f2heat_HI = BD_F2HEAT_HI_B2[i]; bc[q].f2heat_HeI = BD_F2HEAT_HEI_B2[i]; bc[q].f2heat_HeII = BD_F2HEAT_HEII_B2[i];
  }
  for (int i = 0; i < 20; i++) {
    const int q = NumBndin1 + NumBndin2 + i;
    bc[q].sigma_HI = BD_SIGMA_HI_B3[i]; bc[q].sigma_HeI = BD_SIGMA_HEI_B3[i]; bc[q].sigma_HeII = BD_SIGMA_HEII_B3[i];
    bc[q].f1ion_HI = BD_F1ION_HI_B3[i]; bc[q].f1ion_HeI = BD_F1ION_HEI_B3[i]; bc[q].f1ion_HeII = BD_F1ION_HEII_B3[i];
    bc[q].f2ion_HI = BD_F2ION_HI_B3[i]; bc[q].f2ion_HeI = BD_F2ION_HEI_B3[i]; bc[q].f2ion_HeII = BD_F2ION_HEII_B3[i];
    bc[q].f1heat_HI = BD_F1HEAT_HI_B3[i]; bc[q].f1heat_HeI = BD_F1HEAT_HEI_B3[i]; bc[q].f1heat_HeII = BD_F1HEAT_HEII_B3[i];
    bc[q].f2heat_HI = BD_F2HEAT_HI_B3[i]; bc[q].f2heat_HeI = BD_F2HEAT_HEI_B3[i]; bc[q].f2heat_HeII = BD_F2HEAT_HEII_B3[i];
  }
  for (int q = 0; q < NumFreqBnd; q++) bc[q].dead_bb = INFINITY;  // until tables are packed (pack_tables)
  CK(cudaMemcpyToSymbol(d_band, bc, sizeof(c->h_band)));
#if C2RAY_TABLOG
  {
    // tau_table_position's mantissa table and series coefficients (c2ray_photo.cuh), formed in long double
    const long double K = 1.0L / (long double)dlogtau, l10e = 1.0L / logl(10.0L);
    static double2 tab[256];
    for (int i = 0; i < 256; i++) {
      const double ci = 1.0 + (i + 0.5) / 256.0;
      const double ri = (double)(1.0L / (long double)ci);
      tab[i].x = ri;
      tab[i].y = (double)(1.0L + (log10l(1.0L / (long double)ri) - (long double)minlogtau) * K);
    }
    const double pc[7] = {(double)(log10l(2.0L) * K), (double)(l10e * K), (double)(-l10e * K / 2.0L), (double)(l10e * K / 3.0L),
                          (double)(-l10e * K / 4.0L), (double)(l10e * K / 5.0L), (double)(-l10e * K / 6.0L)};
    CK(cudaMemcpyToSymbol(g_postab, tab, sizeof(tab)));
    CK(cudaMemcpyToSymbol(d_posc, pc, sizeof(pc)));
  }
#endif
  return 0;
}

// band edges, radiation_sizes.f90:96-192
void band_edges(double* fmin, double* fmax, double* dfreq) {
  fmax[0] = ion_freq_HeI;
  for (int i = 0; i < 25; i++) fmax[1 + i] = ion_freq_HeI * BD_FREQMAX_MULT_HEI[i];
  fmax[NumBndin1 + NumBndin2 - 1] = ion_freq_HeII;
  for (int i = 0; i < 20; i++) fmax[NumBndin1 + NumBndin2 + i] = ion_freq_HeII * BD_FREQMAX_MULT_HEII[i];
  fmin[0] = ion_freq_HI;
  for (int q = 1; q < NumFreqBnd; q++) fmin[q] = fmax[q - 1];
  for (int q = 0; q < NumFreqBnd; q++) dfreq[q] = (fmax[q] - fmin[q]) / (double)NumFreq;
}

// Weights of the fixed-weight "Romberg" quadrature over 2^p+1 points (romberg.f90:22-96).  The reference forms the
// Richardson factors in default real: b_k = -1.0/(4.0**k-1.0) is a binary32 value; a_k = -b_k*4^k in double.
void romberg_weights(int npow, double* w /* 2^npow + 1 */) {
  const int n = 1 << npow;
  std::vector<double> a(npow + 1, 0.0), b(npow + 1, 0.0);
  for (int k = 1; k <= npow; k++) {
    const float f4k = (float)(1u << (2 * k));
    b[k] = (double)(-1.0f / (f4k - 1.0f));
    a[k] = -b[k] * (double)f4k;
  }
  for (int j = 0; j <= n; j++) w[j] = 0.0;
  // Each trapezoid level k (stride n/2^k) enters the final extrapolate with coefficient T(npow,npow | unit at level k).
  std::vector<double> col(npow + 1);
  for (int k = 0; k <= npow; k++) {
    std::vector<std::vector<double>> s(npow + 1, std::vector<double>(npow + 1, 0.0));
    s[k][0] = 1.0;
    for (int j = 1; j <= npow; j++)
      for (int i = npow; i >= j; i--) s[i][j] = a[j] * s[i][j - 1] + b[j] * s[i - 1][j - 1];
    const int stride = 1 << (npow - k);
    for (int j = 0; j <= (1 << k); j++) w[stride * j] = s[npow][npow] * (double)stride + w[stride * j];
  }
  w[0] = FL(0.5f) * w[0];
  w[n] = FL(0.5f) * w[n];
}

int alloc_sweep(c2ray_ctx* c, int want_slots) {
  // reach of the trace: evolve_source.F90:103-105
  int rmax = 0;
  for (int d = 0; d < 3; d++) {
    c->geom.R[d] = std::min(c->par.max_subbox, c->mesh[d] / 2 - 1 + c->mesh[d] % 2);
    c->geom.L[d] = std::min(c->par.max_subbox, c->mesh[d] / 2);
    rmax = std::max(rmax, std::max(c->geom.R[d], c->geom.L[d]));
  }
  c->geom.subboxsize = c->par.subboxsize;
  const int cap = rmax == 0 ? 1 : 24 * rmax * rmax + 2;
  // slots_cap below the request means the scratch budget, not the request, set it: asking again changes nothing
  if (c->d_scratch && cap == c->geom.cap && (want_slots <= c->slots_cap || c->slots_budget_limited)) return 0;
  c->geom.cap = cap;
  if (c->d_scratch) { cudaFree(c->d_scratch); cudaFree(c->d_slots); cudaFree(c->d_active); c->d_scratch = nullptr; }
  size_t per_slot = (size_t)6 * cap * sizeof(double);
  size_t freeb = 0, totb = 0;
  CK(cudaMemGetInfo(&freeb, &totb));
  const size_t budget = std::min<size_t>(freeb / 3, (size_t)32 << 30);
  int slots = (int)std::max<size_t>(1, std::min<size_t>((size_t)want_slots, budget / per_slot));
  CK(cudaMalloc(&c->d_scratch, per_slot * slots));
  CK(cudaMalloc(&c->d_slots, sizeof(Slot) * slots));
  CK(cudaMalloc(&c->d_active, sizeof(int) * slots));
  c->slots_cap = slots;
  c->slots_budget_limited = slots < want_slots;
  return 0;
}

int upload_my_sources(c2ray_ctx* c) {
  c->n_mine = (int)c->my_ids.size();
  c->run_key = -1;  // sweep_all rebuilds and uploads the run-ordered list
  return 0;
}

void balanced_partition(int NumSrc, const long long* cost, int npr, int* owner);

int rebuild_my_sources(c2ray_ctx* c) {
  c->my_ids.clear();
  if (c->schedule == 1 && c->npr > 1 && (int)c->src_flux.size() == c->NumSrc && c->NumSrc > 0) {
    // Balanced schedule before any pass has left cost records: deal by photon rate, brightest first to the least loaded
    // rank (how far a source traces grows with its flux; the records of the first pass replace this guess).
    double fmax = 0.0;
    for (double f : c->src_flux) fmax = std::max(fmax, f);
    std::vector<long long> cost(c->NumSrc);
    for (int i = 0; i < c->NumSrc; i++) cost[i] = fmax > 0.0 ? (long long)(1.0e9 * c->src_flux[i] / fmax) + 1 : 1;
    std::vector<int> owner(c->NumSrc);
    balanced_partition(c->NumSrc, cost.data(), c->npr, owner.data());
    for (int i = 0; i < c->NumSrc; i++) if (owner[i] == c->rank) c->my_ids.push_back(i);
  } else {
    for (int ns1 = 1 + c->rank; ns1 <= c->NumSrc; ns1 += c->npr) c->my_ids.push_back(ns1 - 1);  // master_slave.F90:85
  }
  if (c->d_nbox_all) { cudaFree(c->d_nbox_all); c->d_nbox_all = nullptr; }
  if (c->NumSrc > 0) {
    CK(cudaMalloc(&c->d_nbox_all, sizeof(int) * c->NumSrc));
    CK(cudaMemset(c->d_nbox_all, 0, sizeof(int) * c->NumSrc));
  }
  return upload_my_sources(c);
}

// Longest-processing-time-first assignment of sources to ranks: sources in order of decreasing cost (ties: lower
// source number first) each go to the rank with the smallest load so far (ties: lowest rank).  Every rank computes the
// same table from the same costs.  This is the static stand-in for the reference's master/slave hand-out
// (master_slave.F90:124-326), where a free slave asks the master for the next source.
void balanced_partition(int NumSrc, const long long* cost, int npr, int* owner) {
  std::vector<int> order(NumSrc);
  for (int i = 0; i < NumSrc; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  std::vector<long long> load(npr, 0);
  for (int i : order) {
    int best = 0;
    for (int r = 1; r < npr; r++) if (load[r] < load[best]) best = r;
    owner[i] = best;
    load[best] += cost[i] > 0 ? cost[i] : 1;
  }
}

// cells a source with `nbox` sub-boxes traces: prod_d (last_r - last_l + 1), evolve_source.F90:143-144
long long source_cost(const c2ray_ctx* c, int nbox) {
  long long cells = 1;
  for (int d = 0; d < 3; d++) {
    const long long reach = (long long)c->par.subboxsize * std::max(nbox, 1);
    cells *= std::min<long long>(reach, c->geom.L[d]) + std::min<long long>(reach, c->geom.R[d]) + 1;
  }
  return cells;
}

// After a pass: share the sub-box counts and re-deal the sources for the next pass (balanced schedule only).
int rebalance_sources(c2ray_ctx* c) {
  if (c->schedule != 1 || !c->comm || c->npr <= 1 || c->NumSrc <= 0) return 0;
  NC_(g_nccl.AllReduce(c->d_nbox_all, c->d_nbox_all, (size_t)c->NumSrc, NCCL_INT32, NCCL_SUM, c->comm, c->stream));
  std::vector<int> nbox(c->NumSrc);
  CK(cudaMemcpyAsync(nbox.data(), c->d_nbox_all, sizeof(int) * c->NumSrc, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemsetAsync(c->d_nbox_all, 0, sizeof(int) * c->NumSrc, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  std::vector<long long> cost(c->NumSrc);
  for (int i = 0; i < c->NumSrc; i++) cost[i] = source_cost(c, nbox[i]);
  std::vector<int> owner(c->NumSrc);
  balanced_partition(c->NumSrc, cost.data(), c->npr, owner.data());
  c->my_ids.clear();
  for (int i = 0; i < c->NumSrc; i++) if (owner[i] == c->rank) c->my_ids.push_back(i);
  return upload_my_sources(c);
}

// Sums the per-group totals into the context's aggregate and packs [photon_loss(1:47) | sum_nbox] behind the rate
// grids; only photon_loss(1) is ever filled (evolve_source.F90:233).
__global__ void k_pack_tail(const SweepTotals* gtot, int ngroups, SweepTotals* tot, double* tail) {
  const int t = threadIdx.x;
  if (t == 0) {
    SweepTotals a;
    a.photon_loss = 0.0; a.sum_nbox = 0; a.updates = 0; a.nactive = 0;
    for (int g = 0; g < ngroups; g++) { a.photon_loss += gtot[g].photon_loss; a.sum_nbox += gtot[g].sum_nbox; a.updates += gtot[g].updates; }
    *tot = a;
    tail[0] = a.photon_loss;
    tail[NumFreqBnd] = (double)a.sum_nbox;
  } else if (t < NumFreqBnd) tail[t] = 0.0;
}

// A launch that may begin while its predecessor in the stream drains (the kernel waits itself, griddepcontrol.wait).
template <typename... KArgs, typename... Args>
cudaError_t launch_overlapped(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, bool overlap, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = overlap ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

#define LAUNCH_S(ctx, strm, kernel, grid, block, ...)     \
  do {                                                    \
    kernel<<<(grid), (block), 0, (strm)>>>(__VA_ARGS__);  \
    (ctx)->launches++;                                    \
  } while (0)

// evolve.F90:385 pass_all_sources for this rank's sources (device work only, no host sync).
// The sources of a batch are split into groups that run on separate streams: a shell launch ends with a partially
// filled last wave (and the innermost shells are a single short wave), so the next group's launch fills the SMs the
// previous one is draining.  Groups share nothing but the atomically accumulated rate grids.
int sweep_all(c2ray_ctx* c) {
  int rc = bind(c);
  if (rc) return rc;
  int ngroups = 1;
  if (c->n_mine > 0) {
    int want = 1;
    if (!c->par.deterministic) {  // a multiple of the group count, so that every group gets the same number of slots
      const int gw = std::max(1, std::min(c->sweep_groups, c->n_mine));
      want = std::min(c->n_mine, c->par.max_slots > 0 ? c->par.max_slots : 1024);
      want = (want + gw - 1) / gw * gw;
    }
    rc = alloc_sweep(c, want);
    if (rc) return rc;
    const SweepGeom g = c->geom;
    int rmax = 0;
    for (int d = 0; d < 3; d++) rmax = std::max(rmax, std::max(g.R[d], g.L[d]));
    if (!c->d_cellrec) CK(cudaMalloc(&c->d_cellrec, CELLREC * c->N3 * sizeof(double)));
    // Records for the whole mesh, or -- when the previous pass of this rank touched less than half of it -- only for
    // the cells each sub-box level is about to trace (k_cell_records_level, launched per level below): on a 512^3 mesh
    // with 1250 sources that stop after their first sub-box the full pass costs 3.3 ms of an 8.6 ms RT pass.
    const bool sparse_records = c->sparse_records && !c->par.deterministic && c->last_pass_updates >= 0 &&
                                (double)c->last_pass_updates * 2.0 < (double)c->N3;
    if (!sparse_records)
      LAUNCH(c, k_cell_records, (unsigned)((c->N3 + 255) / 256), 256, c->ndens, c->xh_av, c->xhe_av, c->N3,
             c->par.isothermal ? 1 : 0, c->d_cellrec);
    GridPtrs G{c->d_cellrec, c->rates, c->N3, c->lls_type, c->coldensh_LLS, c->d_lls};
    double* lossbuf = nullptr;   // deterministic mode: per-cell photon-loss contributions of the current shell (k_loss_sum)
    if (c->par.deterministic) {
      if (c->lossbuf_cap < g.cap) {
        if (c->d_lossbuf) cudaFree(c->d_lossbuf);
        c->d_lossbuf = nullptr;
        CK(cudaMalloc(&c->d_lossbuf, sizeof(double) * (size_t)g.cap));
        CK(cudaMemset(c->d_lossbuf, 0, sizeof(double) * (size_t)g.cap));
        c->lossbuf_cap = g.cap;
      }
      lossbuf = c->d_lossbuf;
    }
    ngroups = c->par.deterministic ? 1 : std::max(1, std::min(c->sweep_groups, std::min(c->slots_cap, c->n_mine)));
    const int region = c->par.deterministic ? 1 : c->slots_cap / ngroups;  // slots per stream group
    const int batch = region * ngroups;                                     // sources in flight at a time
    // A source whose PL and QPL fluxes are zero (or whose tables are absent) contributes through the black-body
    // tables only and takes the single-SED kernel, whatever other sources need: in a -DQUASARS run with QPL flux
    // on a few bright sources the rest do not pay for the three-SED loop.  The list is ordered single-SED sources
    // first; deterministic mode keeps source order and picks the kernel per source.
    const bool has_bb = c->tab[0][0] != nullptr, has_pl = c->tab[1][0] != nullptr, has_qpl = c->tab[2][0] != nullptr;
    auto is_multi = [&](int id) { return !has_bb || (has_pl && c->src_pl[id]) || (has_qpl && c->src_qpl[id]); };
    const int key = (has_bb ? 1 : 0) | (has_pl ? 2 : 0) | (has_qpl ? 4 : 0) | (c->par.deterministic ? 8 : 0);
    if (c->run_key != key) {
      std::vector<int> order;
      order.reserve(c->n_mine);
      if (c->par.deterministic) {
        order = c->my_ids;
        c->n_single = 0;
      } else {
        for (int id : c->my_ids) if (!is_multi(id)) order.push_back(id);
        c->n_single = (int)order.size();
        for (int id : c->my_ids) if (is_multi(id)) order.push_back(id);
      }
      if (c->d_srcids_run) { cudaFree(c->d_srcids_run); c->d_srcids_run = nullptr; }
      CK(cudaMalloc(&c->d_srcids_run, sizeof(int) * order.size()));
      CK(cudaMemcpy(c->d_srcids_run, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice));
      c->run_key = key;
    }
    const size_t slot_stride = (size_t)6 * g.cap;
    CK(cudaMemsetAsync(c->d_gtot, 0, sizeof(SweepTotals) * MAX_SWEEP_GROUPS, c->stream));
    CK(cudaEventRecord(c->ev_fork, c->stream));
    for (int q = 0; q < ngroups; q++) CK(cudaStreamWaitEvent(c->gstream[q], c->ev_fork, 0));
    for (int first = 0; first < c->n_mine;) {
      // a batch never mixes the two kinds
      bool multi_sed;
      int ns;
      if (c->par.deterministic) { multi_sed = is_multi(c->my_ids[first]); ns = 1; }
      else if (first < c->n_single) { multi_sed = false; ns = std::min(batch, c->n_single - first); }
      else { multi_sed = true; ns = std::min(batch, c->n_mine - first); }
      struct Advance { int& f; int n; ~Advance() { f += n; } } advance{first, ns};
      // Group q owns the fixed slot range [q*region, (q+1)*region) in every batch: batches follow each other on a
      // group's own stream without any cross-stream wait, so a slot (its state, active-list entry and shell scratch)
      // must never move to another group.  soff: where a group's sources start within the batch.
      const int per = (ns + ngroups - 1) / ngroups;
      int goff[MAX_SWEEP_GROUPS], soff[MAX_SWEEP_GROUPS], gns[MAX_SWEEP_GROUPS];
      for (int q = 0; q < ngroups; q++) { goff[q] = q * region; soff[q] = std::min(q * per, ns); gns[q] = std::min(per, ns - soff[q]); }
      for (int q = 0; q < ngroups; q++)
        if (gns[q] > 0)
          LAUNCH_S(c, c->gstream[q], k_slots_init, (gns[q] + 127) / 128, 128, c->d_slots + goff[q], gns[q],
                   c->d_srcids_run + first + soff[q], c->d_srcpos, c->d_nf, c->have_pl_flux ? c->d_nfpl : nullptr,
                   c->have_qpl_flux ? c->d_nfqpl : nullptr, c->d_gtot + q, c->d_active + goff[q]);
      const int reach3 = std::min(g.R[2], g.L[2]);
      int nact[MAX_SWEEP_GROUPS];
      for (int q = 0; q < ngroups; q++) nact[q] = gns[q];
      for (int b = 1;; b++) {
        for (int q = 0; q < ngroups; q++)
          if (nact[q] > 0) LAUNCH_S(c, c->gstream[q], k_decide, 1, 256, c->d_slots + goff[q], gns[q], g, c->d_gtot + q, c->d_active + goff[q], c->d_nbox_all);
        // (enqueueing the levels the previous pass reached without this wait was tried for the few-source case: no change,
        // 53.2 ms of sweeps per step for one source at 128^3 either way -- the wait overlaps the draining level;
        // profiles/r2_ab8_speculate.log)
        if (b > 1) {
          // From the second sub-box on most sources have dropped out (photon loss below 1e-10 of the flux): fetch the
          // number still active so that the shells of a level nobody traces are not launched at all and the grids
          // of the others are sized to the work that exists.  One short host wait per sub-box level.
          for (int q = 0; q < ngroups; q++)
            if (nact[q] > 0) CK(cudaMemcpyAsync(c->h_nact + q, &c->d_gtot[q].nactive, sizeof(int), cudaMemcpyDeviceToHost, c->gstream[q]));
          bool any = false;
          for (int q = 0; q < ngroups; q++)
            if (nact[q] > 0) { CK(cudaStreamSynchronize(c->gstream[q])); nact[q] = c->h_nact[q]; any = any || nact[q] > 0; }
          if (!any) break;
        }
        const int r_lo = b == 1 ? 0 : g.subboxsize * (b - 1) + 1;
        const int r_hi = (int)std::min<long long>((long long)g.subboxsize * b, rmax);
        if (sparse_records)
          for (int q = 0; q < ngroups; q++) {
            if (nact[q] <= 0) continue;
            const long long per = (long long)(2 * r_hi + 1) * (2 * r_hi + 1) * (2 * r_hi + 1) -
                                  (r_lo > 0 ? (long long)(2 * r_lo - 1) * (2 * r_lo - 1) * (2 * r_lo - 1) : 0);
            const int blocks = (int)std::min<long long>((per * nact[q] + 255) / 256, 148 * 16);
            LAUNCH_S(c, c->gstream[q], k_cell_records_level, blocks, 256, c->d_slots + goff[q], c->d_active + goff[q], c->d_gtot + q,
                     r_lo, r_hi, c->ndens, c->xh_av, c->xhe_av, c->N3, c->par.isothermal ? 1 : 0, c->d_cellrec);
          }
        for (int r = r_lo; r <= r_hi; r++) {
          for (int q = 0; q < ngroups; q++) {
            if (nact[q] <= 0) continue;
            const long long cells = (long long)nact[q] * (r == 0 ? 1 : 24LL * r * r + 2);
            // Fewer cells than resident threads: the launch is bound by the latency of one update's dependency chain,
            // not by throughput, so 2..16 lanes share a cell (see k_sweep_shell) -- as many as still fit the resident
            // threads (x sweep_lanes_fill).  sweep_lanes_mode 0: the round-1 rule (8 lanes below a quarter wave).
            const long long resident = 148LL * 128 * C2RAY_SWEEP_MINBLOCKS_SPLIT;  // of the LANES > 1 instances
            int lanes = 1;
            if (c->sweep_split) {
              if (c->sweep_lanes_mode == 0) {
                if (cells * ngroups * 4 <= resident) lanes = SPLIT_LANES;
              } else {
                for (int L = 16; L >= 2; L >>= 1)
                  if ((double)(cells * ngroups * L) <= c->sweep_lanes_fill * (double)resident) { lanes = L; break; }
              }
            }
            const long long items = cells * lanes;
            const unsigned cta = lanes > 1 ? 128u : (unsigned)C2RAY_SWEEP_THREADS;
            const int blocks = (int)((items + cta - 1) / cta);   // one work item per thread
            // the predecessor in this group's stream is the previous shell of the same level (not k_decide): overlap
            // Measured: +6 % on a single source (one group: nothing else fills the draining tail), -0.7 % with two
            // groups on two streams (they already overlap each other's tails) -> only used with a single group.
            const bool pdl = c->sweep_pdl && ngroups == 1 && r > r_lo;
#define SWEEP(ISO, MULTI, LANES)                                                                                              \
  do {                                                                                                                        \
    CK(launch_overlapped(k_sweep_shell<ISO, MULTI, LANES>, (unsigned)blocks, cta, c->gstream[q], pdl, c->d_slots + goff[q],  \
                         (const int*)(c->d_active + goff[q]), c->d_gtot + q, g, G,                                            \
                         c->d_scratch + (size_t)goff[q] * slot_stride, r, lossbuf));                                          \
    c->launches++; c->sweep_launches++;                                                                                        \
  } while (0)
#define SWEEP2(ISO, MULTI)                         \
  do {                                             \
    switch (lanes) {                               \
      case 16: SWEEP(ISO, MULTI, 16); break;       \
      case 8: SWEEP(ISO, MULTI, 8); break;         \
      case 4: SWEEP(ISO, MULTI, 4); break;         \
      case 2: SWEEP(ISO, MULTI, 2); break;         \
      default: SWEEP(ISO, MULTI, 1); break;        \
    }                                              \
  } while (0)
            if (multi_sed) { if (c->par.isothermal) SWEEP2(true, true); else SWEEP2(false, true); }
            else { if (c->par.isothermal) SWEEP2(true, false); else SWEEP2(false, false); }
#undef SWEEP2
#undef SWEEP
            if (lossbuf)  // deterministic mode: one source, its slot is the first of the group
              LAUNCH_S(c, c->gstream[q], k_loss_sum, 1, 1024, c->d_slots + goff[q], lossbuf, (int)(r == 0 ? 1 : 24LL * r * r + 2));
          }
        }
        if ((long long)g.subboxsize * b >= reach3) break;  // the do-while's extent test fails for every source
      }
      for (int q = 0; q < ngroups; q++)  // close the sources still active
        if (nact[q] > 0) LAUNCH_S(c, c->gstream[q], k_decide, 1, 256, c->d_slots + goff[q], gns[q], g, c->d_gtot + q, c->d_active + goff[q], c->d_nbox_all);
    }
    for (int q = 0; q < ngroups; q++) {
      CK(cudaEventRecord(c->ev_join[q], c->gstream[q]));
      CK(cudaStreamWaitEvent(c->stream, c->ev_join[q], 0));
    }
  } else {
    CK(cudaMemsetAsync(c->d_gtot, 0, sizeof(SweepTotals) * MAX_SWEEP_GROUPS, c->stream));
  }
  LAUNCH(c, k_pack_tail, 1, 64, c->d_gtot, ngroups, c->d_tot, c->rates + 4 * c->N3);
  CK(cudaGetLastError());
  return 0;
}

int allreduce_rates(c2ray_ctx* c) {  // evolve.F90:505-548
  if (!c->comm || c->npr <= 1) return 0;
  int r = g_nccl.AllReduce(c->rates, c->rates, c->rates_count, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream);
  if (r != 0) return fail(C2RAY_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  return 0;
}

// ---- files ------------------------------------------------------------------------------------------------------
constexpr size_t STAGE_BYTES = (size_t)64 << 20;

int ensure_stage(c2ray_ctx* c) {
  if (!c->h_stage) CK(cudaMallocHost(&c->h_stage, STAGE_BYTES));
  return 0;
}

// appends `bytes` of device memory to the record being written; to_f32: the source is FP64 and the file gets real(x)
int put_device(c2ray_ctx* c, c2io::RecordWriter& w, const void* dptr, size_t bytes, bool to_f32 = false) {
  int rc = ensure_stage(c);
  if (rc) return rc;
  const char* d = static_cast<const char*>(dptr);
  for (size_t off = 0; off < bytes; off += STAGE_BYTES) {
    const size_t k = std::min(STAGE_BYTES, bytes - off);
    CK(cudaMemcpyAsync(c->h_stage, d + off, k, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (to_f32) {
      const double* src = static_cast<const double*>(c->h_stage);
      float* dst = static_cast<float*>(c->h_stage);  // in place: element i is read before slot i/2 is overwritten
      const size_t n = k / 8;
      for (size_t i = 0; i < n; i++) { const double v = src[i]; dst[i] = (float)v; }
      if (!w.put(dst, n * 4)) return fail(C2RAY_ERR_STATE, "short write");
    } else if (!w.put(c->h_stage, k)) {
      return fail(C2RAY_ERR_STATE, "short write");
    }
  }
  return 0;
}

int get_device(c2ray_ctx* c, c2io::RecordReader& r, void* dptr, size_t bytes) {
  int rc = ensure_stage(c);
  if (rc) return rc;
  char* d = static_cast<char*>(dptr);
  for (size_t off = 0; off < bytes; off += STAGE_BYTES) {
    const size_t k = std::min(STAGE_BYTES, bytes - off);
    if (!r.get(c->h_stage, k)) return fail(C2RAY_ERR_STATE, "iteration dump: record does not match this mesh / isothermal setting");
    CK(cudaMemcpyAsync(d + off, c->h_stage, k, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

std::string join_dir(const std::string& dir, const std::string& name) {
  // the reference concatenates trim(adjustl(dump_dir))//iterfile, i.e. the directory string carries its own separator
  if (dir.empty() || dir.back() == '/') return dir + name;
  return dir + "/" + name;
}

// evolve.F90:233-275 write_iteration_dump: niter, photon_loss_all, phih_grid, xh_av, xh_intermed, phihe_grid, xhe_av,
// xhe_intermed [, phiheat, temperature_grid]
int write_iteration_dump_file(c2ray_ctx* c, const std::string& path, int niter) {
  const size_t N3 = c->N3;
  c2io::RecordWriter w;
  if (!w.open(path)) return fail(C2RAY_ERR_STATE, "cannot open " + path + " for writing");
  int rc = 0;
  const int32_t ni = niter;
  if (!w.record(&ni, 4)) return fail(C2RAY_ERR_STATE, "short write");
  if (!w.begin(NumFreqBnd * 8) || (rc = put_device(c, w, c->rates + 4 * N3, NumFreqBnd * 8))) return rc ? rc : fail(C2RAY_ERR_STATE, "short write");
  struct Rec { const void* p; size_t bytes; };
  const Rec recs[] = {{c->rates, N3 * 8},          {c->xh_av, 2 * N3 * 8},  {c->xh_int, 2 * N3 * 8}, {c->rates + N3, 2 * N3 * 8},
                      {c->xhe_av, 3 * N3 * 8},     {c->xhe_int, 3 * N3 * 8}, {c->rates + 3 * N3, N3 * 8}, {c->temp, 3 * N3 * 4}};
  const int nrec = c->par.isothermal ? 6 : 8;
  for (int i = 0; i < nrec; i++) {
    if (!w.begin(recs[i].bytes)) return fail(C2RAY_ERR_STATE, "short write");
    if ((rc = put_device(c, w, recs[i].p, recs[i].bytes))) return rc;
  }
  if (!w.close()) return fail(C2RAY_ERR_STATE, "error closing " + path);
  return 0;
}

// evolve.F90:279-367 start_from_dump (every rank reads the file itself instead of rank 0 reading and broadcasting)
int read_iteration_dump_file(c2ray_ctx* c, const std::string& path, int* niter) {
  const size_t N3 = c->N3;
  c2io::RecordReader r;
  if (!r.open(path)) return fail(C2RAY_ERR_STATE, "cannot open iteration dump " + path);
  int32_t ni = 0;
  if (!r.record(&ni, 4)) return fail(C2RAY_ERR_STATE, "iteration dump: bad first record in " + path);
  int rc;
  if (!r.begin(NumFreqBnd * 8)) return fail(C2RAY_ERR_STATE, "iteration dump: bad photon_loss record");
  if ((rc = get_device(c, r, c->rates + 4 * N3, NumFreqBnd * 8))) return rc;
  struct Rec { void* p; size_t bytes; };
  const Rec recs[] = {{c->rates, N3 * 8},          {c->xh_av, 2 * N3 * 8},  {c->xh_int, 2 * N3 * 8}, {c->rates + N3, 2 * N3 * 8},
                      {c->xhe_av, 3 * N3 * 8},     {c->xhe_int, 3 * N3 * 8}, {c->rates + 3 * N3, N3 * 8}, {c->temp, 3 * N3 * 4}};
  const int nrec = c->par.isothermal ? 6 : 8;
  for (int i = 0; i < nrec; i++) {
    if (!r.begin(recs[i].bytes)) return fail(C2RAY_ERR_STATE, "iteration dump: record does not match this mesh / isothermal setting");
    if ((rc = get_device(c, r, recs[i].p, recs[i].bytes))) return rc;
  }
  if (niter) *niter = ni;
  return 0;
}

// output.F90: write(unit) mesh(1),mesh(2),mesh(3) ; write(unit) (((a(i,j,k),i=..),j=..),k=..)
int write_plane_file(c2ray_ctx* c, const std::string& path, const void* dplane, size_t elem_bytes, bool to_f32) {
  c2io::RecordWriter w;
  if (!w.open(path)) return fail(C2RAY_ERR_STATE, "cannot open " + path + " for writing");
  const int32_t m[3] = {c->mesh[0], c->mesh[1], c->mesh[2]};
  if (!w.record(m, 12)) return fail(C2RAY_ERR_STATE, "short write");
  if (!w.begin(c->N3 * (to_f32 ? 4 : elem_bytes))) return fail(C2RAY_ERR_STATE, "short write");
  int rc = put_device(c, w, dplane, c->N3 * elem_bytes, to_f32);
  if (rc) return rc;
  if (!w.close()) return fail(C2RAY_ERR_STATE, "error closing " + path);
  return 0;
}

#define NC(call)                                                                                              \
  do {                                                                                                        \
    int r_ = (call);                                                                                          \
    if (r_ != 0)                                                                                              \
      return fail(C2RAY_ERR_NCCL, std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?")); \
  } while (0)

// ---- the global pass split over the ranks ------------------------------------------------------
// The reference runs the global pass replicated on every rank after an allreduce of the four rate grids
// (evolve.F90:477-548).  Cells are independent there, so with a communicator attached evolve3d gives rank r the cells
// [r*N3/npr, (r+1)*N3/npr) of every plane: the allreduce becomes its first half (reduce-scatter: a rank only needs the
// summed rates of its own cells), the pass runs on 1/npr of the mesh, and the second half (all-gather) moves the four
// planes the next sweep reads -- xh_av(0:1), xhe_av(0:1) -- instead of the rates.  Same bytes on the wire as the
// allreduce, 1/npr of the chemistry.  The planes the sweep does not read are gathered once, when the iteration ends.
bool split_active(const c2ray_ctx* c) { return c->comm && c->npr > 1 && c->split_chem && c->N3 % (size_t)c->npr == 0; }

int reduce_scatter_rates(c2ray_ctx* c) {
  const size_t chunk = c->N3 / c->npr;
  const int planes = c->par.isothermal ? 3 : 4;
  NC(g_nccl.GroupStart());
  for (int q = 0; q < planes; q++) {
    double* plane = c->rates + (size_t)q * c->N3;
    NC(g_nccl.ReduceScatter(plane, plane + (size_t)c->rank * chunk, chunk, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
  }
  double* tail = c->rates + 4 * c->N3;  // photon_loss(1:47), sum_nbox: every rank needs them
  NC(g_nccl.AllReduce(tail, tail, NumFreqBnd + 1, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
  NC(g_nccl.GroupEnd());
  return 0;
}

int allgather_f64(c2ray_ctx* c, double* const* planes, int n) {
  const size_t chunk = c->N3 / c->npr;
  NC(g_nccl.GroupStart());
  for (int q = 0; q < n; q++)
    NC(g_nccl.AllGather(planes[q] + (size_t)c->rank * chunk, planes[q], chunk, NCCL_FLOAT64, c->comm, c->stream));
  NC(g_nccl.GroupEnd());
  return 0;
}

// conv_flag, sum(nit), sum(thermal sub-steps), last_coef_T: sums over ranks; nit_max: maximum
int combine_chem_totals(c2ray_ctx* c) {
  int* d_max = reinterpret_cast<int*>(c->d_chemred + 4);
  LAUNCH(c, k_chem_pack, 1, 1, c->d_chem, c->d_chemred, d_max);
  NC(g_nccl.GroupStart());
  NC(g_nccl.AllReduce(c->d_chemred, c->d_chemred, 4, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
  NC(g_nccl.AllReduce(d_max, d_max, 1, NCCL_INT32, NCCL_MAX, c->comm, c->stream));
  NC(g_nccl.GroupEnd());
  LAUNCH(c, k_chem_unpack, 1, 1, c->d_chem, c->d_chemred, d_max);
  return 0;
}

// what the sweep of the next iteration reads
int allgather_sweep_inputs(c2ray_ctx* c) {
  double* planes[4] = {c->xh_av, c->xh_av + c->N3, c->xhe_av, c->xhe_av + c->N3};
  return allgather_f64(c, planes, 4);
}

// everything else a rank owns only its share of, once the iteration has ended: xh_intermed, xhe_intermed, xhe_av(2),
// temperature_grid(0:1) and the summed rate grids (output.F90:354,364 writes them)
int allgather_final(c2ray_ctx* c) {
  const size_t N3 = c->N3, chunk = N3 / c->npr;
  double* planes[10] = {c->xh_int, c->xh_int + N3, c->xhe_int, c->xhe_int + N3, c->xhe_int + 2 * N3, c->xhe_av + 2 * N3,
                        c->rates, c->rates + N3, c->rates + 2 * N3, c->rates + 3 * N3};
  int rc = allgather_f64(c, planes, c->par.isothermal ? 9 : 10);
  if (rc) return rc;
  if (!c->par.isothermal) {
    NC(g_nccl.GroupStart());
    for (int q = 0; q < 2; q++)
      NC(g_nccl.AllGather(c->temp + (size_t)q * N3 + (size_t)c->rank * chunk, c->temp + (size_t)q * N3, chunk, NCCL_FLOAT32,
                          c->comm, c->stream));
    NC(g_nccl.GroupEnd());
  }
  return 0;
}

int global_pass_launch(c2ray_ctx* c, double dt, int* d_nit, size_t p_begin = 0, size_t p_end = 0) {
  if (p_end == 0) p_end = c->N3;
  const size_t ncell = p_end - p_begin;
  int rc = bind(c);
  if (rc) return rc;
  if (!c->par.isothermal && !c->have_cool) return fail(C2RAY_ERR_STATE, "cooling tables not set (c2ray_b200_set_cooling_tables)");
  CK(cudaMemsetAsync(c->d_chem, 0, sizeof(ChemTotals), c->stream));
  ChemPtrs P{c->ndens, c->xh, c->xhe, c->xh_av, c->xhe_av, c->xh_int, c->xhe_int, c->temp, c->rates, c->N3, c->d_clump};
  // auto: the queue-driven kernel pays off once cells need many thermal sub-steps (divergence); measured break-even
  // between 10 and 40 sub-steps per cell (config 2: ~10, simple kernel faster; config 5: 42, queue 2x faster)
  // (isothermal passes have no sub-steps: there the queue pays once cells iterate, 4.78 -> 5.60 G cells/s on the config-5
  // inputs -- profiles/r2_ab4_chem.log)
  const bool use_queue = c->chem_mode == 1 || (c->chem_mode < 0 && (c->last_nsub_per_cell > 20.0 ||
                                                                      (c->par.isothermal && c->last_nit_per_cell > 1.2)));
  if (use_queue) {
    // queue-driven: a persistent grid of lanes drawing cells from a counter (k_global_pass_q)
    CK(cudaMemsetAsync(c->d_next_cell, 0, sizeof(unsigned long long), c->stream));
    if (!c->chemq_per_sm) {
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->chemq_per_sm, k_global_pass_q, 128, 0));
      c->chemq_per_sm = std::max(c->chemq_per_sm, 1);
    }
    const unsigned blocks = (unsigned)std::min<size_t>((ncell + 127) / 128, (size_t)c->n_sm * c->chemq_per_sm);
    LAUNCH(c, k_global_pass_q, blocks, 128, P, dt, c->d_chem, d_nit, c->d_next_cell, p_begin, p_end, c->chem_burst, c->chem_thermal_min);
  } else {
    const unsigned blocks = (unsigned)((ncell + 127) / 128);
    LAUNCH(c, k_global_pass, blocks, 128, P, dt, c->d_chem, d_nit, p_begin, p_end);
  }
  CK(cudaGetLastError());
  return 0;
}

int begin_step(c2ray_ctx* c) {  // evolve.F90:131-134
  const size_t N3 = c->N3;
  CK(cudaMemcpyAsync(c->xh_av, c->xh, 2 * N3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->xh_int, c->xh, 2 * N3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->xhe_av, c->xhe, 3 * N3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->xhe_int, c->xhe, 3 * N3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
int end_step(c2ray_ctx* c) {  // evolve.F90:164-166
  const size_t N3 = c->N3;
  CK(cudaMemcpyAsync(c->xh, c->xh_int, 2 * N3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->xhe, c->xhe_int, 3 * N3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  if (!c->par.isothermal)
    CK(cudaMemcpyAsync(c->temp + 2 * N3, c->temp, N3 * 4, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

int state_sums(c2ray_ctx* c, const double* xh, const double* xhe, double* out5) {
  CK(cudaMemsetAsync(c->d_sums, 0, 5 * sizeof(double), c->stream));
  LAUNCH(c, k_state_sums, 148 * 4, 256, c->ndens, xh, xhe, c->N3, c->d_sums);
  double s[5];
  CK(cudaMemcpyAsync(s, c->d_sums, sizeof(s), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  // photonstatistics.f90:141-146
  out5[0] = s[0] * c->vol * (1.0 - abu_he); out5[1] = s[1] * c->vol * (1.0 - abu_he);
  out5[2] = s[2] * c->vol * abu_he; out5[3] = s[3] * c->vol * abu_he; out5[4] = s[4] * c->vol * abu_he;
  return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* c2ray_b200_last_error(void) { return g_err.c_str(); }

static int init_device_state(c2ray_ctx* c);

int c2ray_b200_init(const c2ray_params* params, const int32_t mesh[3], int32_t device, c2ray_ctx** out) {
  if (!params || !mesh || !out) return fail(C2RAY_ERR_ARG, "null argument");
  if (mesh[0] < 2 || mesh[1] < 2 || mesh[2] < 2) return fail(C2RAY_ERR_ARG, "mesh must be >= 2 in every dimension");
  if ((double)sqrtf(3.0f) != sqrt3 || (double)sqrtf(2.0f) != sqrt2) return fail(C2RAY_ERR_STATE, "binary32 sqrt constants mismatch");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(C2RAY_ERR_CUDA, std::string("no CUDA device (the hot path has no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0) {
    const char* lr = getenv("LOCAL_RANK");
    device = lr ? atoi(lr) % ndev : 0;
  }
  if (device >= ndev) return fail(C2RAY_ERR_ARG, "device index out of range");
  c2ray_ctx* c = new c2ray_ctx();
  c->device = device;
  c->par = *params;
  for (int d = 0; d < 3; d++) c->mesh[d] = mesh[d];
  c->N3 = (size_t)mesh[0] * mesh[1] * mesh[2];
  const int rc_alloc = init_device_state(c);
  if (rc_alloc) {  // e.g. out of device memory half way: release what was allocated, keep the message
    const std::string msg = g_err;
    c2ray_b200_destroy(c);
    return fail(rc_alloc, msg);
  }
  *out = c;
  return C2RAY_OK;
}

static int init_device_state(c2ray_ctx* c) {
  const int device = c->device;
  const size_t N3 = c->N3;
  CK(cudaSetDevice(device));
  CK(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device));
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& ev : c->ev) CK(cudaEventCreate(&ev));
  for (auto& ev : c->ev_timer) CK(cudaEventCreate(&ev));
  c->rates_count = 4 * N3 + NumFreqBnd + 1;
  CK(cudaMalloc(&c->ndens, N3 * 8));
  CK(cudaMalloc(&c->xh, 2 * N3 * 8)); CK(cudaMalloc(&c->xhe, 3 * N3 * 8));
  CK(cudaMalloc(&c->xh_av, 2 * N3 * 8)); CK(cudaMalloc(&c->xhe_av, 3 * N3 * 8));
  CK(cudaMalloc(&c->xh_int, 2 * N3 * 8)); CK(cudaMalloc(&c->xhe_int, 3 * N3 * 8));
  CK(cudaMalloc(&c->temp, 3 * N3 * 4));
  CK(cudaMalloc(&c->rates, c->rates_count * 8));
  CK(cudaMemset(c->rates, 0, c->rates_count * 8));
  CK(cudaMemset(c->temp, 0, 3 * N3 * 4));
  CK(cudaMalloc(&c->d_tot, sizeof(SweepTotals)));
  CK(cudaMalloc(&c->d_gtot, sizeof(SweepTotals) * MAX_SWEEP_GROUPS));
  CK(cudaMallocHost(&c->h_nact, sizeof(int) * MAX_SWEEP_GROUPS));
  for (auto& st : c->gstream) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  for (auto& ev : c->ev_join) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  if (const char* e = getenv("C2RAY_SWEEP_SPLIT")) c->sweep_split = atoi(e);
  if (const char* e = getenv("C2RAY_SWEEP_PDL")) c->sweep_pdl = atoi(e);
  if (const char* e = getenv("C2RAY_DEAD_BANDS")) c->dead_bands = atoi(e);
  {
    // Shared-memory carve-out of the sweep instances, percent of the SM's 228 KB; what is left is their L1.  They need
    // 5 x (4 KB + 1 KB reserved); the driver's default leaves more to shared memory than that.  Measured
    // (profiles/r2_ab17_carveout.log): 10 % 12.69 ms per configs[1] pass / 249.2 ms per configs[2] pass, default 12.85 / 253.6,
    // 50 % 13.05 / 264.8, 100 % 13.73 / 281.3.
    int pct = 10;
    if (const char* e = getenv("C2RAY_SWEEP_CARVEOUT")) pct = atoi(e);
    if (pct >= 0) {
#define CARVE(ISO, MULTI, LANES) CK(cudaFuncSetAttribute(k_sweep_shell<ISO, MULTI, LANES>, cudaFuncAttributePreferredSharedMemoryCarveout, pct))
#define CARVE5(ISO, MULTI) CARVE(ISO, MULTI, 1); CARVE(ISO, MULTI, 2); CARVE(ISO, MULTI, 4); CARVE(ISO, MULTI, 8); CARVE(ISO, MULTI, 16)
      CARVE5(false, false); CARVE5(false, true); CARVE5(true, false); CARVE5(true, true);
#undef CARVE5
#undef CARVE
    }
  }
  if (const char* e = getenv("C2RAY_SWEEP_LANES_MODE")) c->sweep_lanes_mode = atoi(e);
  if (const char* e = getenv("C2RAY_SWEEP_LANES_FILL")) c->sweep_lanes_fill = std::max(0.1, atof(e));
  if (const char* e = getenv("C2RAY_SPARSE_RECORDS")) c->sparse_records = atoi(e);
  if (const char* e = getenv("C2RAY_SWEEP_GROUPS")) c->sweep_groups = std::max(1, std::min(MAX_SWEEP_GROUPS, atoi(e)));
  CK(cudaMalloc(&c->d_chem, sizeof(ChemTotals)));
  CK(cudaMalloc(&c->d_sums, 5 * sizeof(double)));
  CK(cudaMalloc(&c->d_next_cell, sizeof(unsigned long long)));
  CK(cudaMalloc(&c->d_chemred, 6 * sizeof(double)));
  if (const char* e = getenv("C2RAY_SPLIT_CHEM")) c->split_chem = atoi(e);
  if (const char* e = getenv("C2RAY_CHEM_QUEUE")) c->chem_mode = atoi(e);
  if (const char* e = getenv("C2RAY_CHEM_BURST")) c->chem_burst = std::max(1, atoi(e));
  if (const char* e = getenv("C2RAY_CHEM_TMIN")) c->chem_thermal_min = std::min(32, std::max(0, atoi(e)));
  CK(cudaMalloc(&c->d_cool, 5 * TEMPPOINTS * sizeof(double)));
  CK(cudaMalloc(&c->d_tb, sizeof(TableBuild)));
  int rc = upload_band_const(c);
  if (rc) return rc;
  c->run_dirty = true;
  return C2RAY_OK;
}

int c2ray_b200_destroy(c2ray_ctx* c) {
  if (!c) return C2RAY_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  void* ptrs[] = {c->ndens, c->xh, c->xhe, c->xh_av, c->xhe_av, c->xh_int, c->xhe_int, c->rates, c->temp, c->snap_xh,
                  c->snap_xhe, c->snap_temp, c->d_srcpos, c->d_nf, c->d_nfpl, c->d_nfqpl, c->d_tb, c->d_cool,
                  c->d_slots, c->d_active, c->d_tot, c->d_gtot, c->d_scratch, c->d_chem, c->d_sums, c->d_nit, c->d_cellrec, c->d_next_cell, c->d_chemred, c->d_clump, c->d_lls, c->d_nbox_all, c->d_srcids_run, c->d_lossbuf};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int s = 0; s < 3; s++) for (int k = 0; k < 4; k++) if (c->tab[s][k]) cudaFree(c->tab[s][k]);
  for (int s = 0; s < 3; s++) if (c->packed[s]) cudaFree(c->packed[s]);
  if (c->h_nact) cudaFreeHost(c->h_nact);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->ev_timer) if (ev) cudaEventDestroy(ev);
  for (auto& st : c->gstream) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  for (auto& ev : c->ev_join) if (ev) cudaEventDestroy(ev);
  cudaStreamDestroy(c->stream);
  if (g_bound == c) g_bound = nullptr;
  delete c;
  return C2RAY_OK;
}

int c2ray_b200_set_params(c2ray_ctx* c, const c2ray_params* p) {
  if (!c || !p) return fail(C2RAY_ERR_ARG, "null argument");
  c->par = *p;
  c->run_dirty = true;
  return C2RAY_OK;
}

int c2ray_b200_set_cooling_tables(c2ray_ctx* c, const double* logT, const double* logL) {
  if (!c || !logT || !logL) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  std::vector<double> lin(5 * TEMPPOINTS);
  for (int i = 0; i < 5 * TEMPPOINTS; i++) lin[i] = pow(10.0, logL[i]);  // cooling_h.f90:163-169
  c->cool_mintemp = logT[0];
  c->cool_dtemp = logT[1] - logT[0];  // :95-96
  CK(cudaMemcpy(c->d_cool, lin.data(), lin.size() * 8, cudaMemcpyHostToDevice));
  c->have_cool = true;
  c->run_dirty = true;
  return C2RAY_OK;
}

static int pack_tables(c2ray_ctx* c, int s) {
  if (!c->tab[s][0]) {
    if (c->packed[s]) { cudaFree(c->packed[s]); c->packed[s] = nullptr; }
    return 0;
  }
  const size_t n = PK_TOTAL;
  if (!c->packed[s]) CK(cudaMalloc(&c->packed[s], n * sizeof(double)));
  const int items = NumFreqBnd * PK_ROWS;
  LAUNCH(c, k_pack_tables, (items + 255) / 256, 256, c->tab[s][0], c->tab[s][1], c->tab[s][2], c->tab[s][3], c->packed[s]);
  CK(cudaGetLastError());
  // optical depth beyond which a band of this SED reads only all-zero table rows (d_dead, c2ray_photo.cuh)
  double* d_tmp = nullptr;
  double dead[NumFreqBnd];
  CK(cudaMalloc(&d_tmp, sizeof(dead)));
  LAUNCH(c, k_band_dead, NumFreqBnd, 256, c->tab[s][0], c->tab[s][1], c->tab[s][2], c->tab[s][3], d_tmp);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dead, d_tmp, sizeof(dead), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaFree(d_tmp));
  if (!c->dead_bands)  // env C2RAY_DEAD_BANDS=0: no band is ever skipped (the bit-identity test's reference run)
    for (double& v : dead) v = INFINITY;
  CK(cudaMemcpyToSymbol(d_dead, dead, sizeof(dead), (size_t)s * sizeof(dead)));
  if (s == 0) {
    for (int q = 0; q < NumFreqBnd; q++) c->h_band[q].dead_bb = dead[q];
    CK(cudaMemcpyToSymbol(d_band, c->h_band, sizeof(c->h_band)));
  }
  return 0;
}

static int ensure_tables(c2ray_ctx* c, int s, bool heat) {
  const size_t np = (size_t)NumFreqBnd * (NumTau + 1), nh = (size_t)NumheatBin * (NumTau + 1);
  for (int k = 0; k < 4; k++) {
    if (k >= 2 && !heat) continue;
    if (!c->tab[s][k]) {
      CK(cudaMalloc(&c->tab[s][k], (k < 2 ? np : nh) * 8));
      CK(cudaMemset(c->tab[s][k], 0, (k < 2 ? np : nh) * 8));
    }
  }
  return 0;
}

int c2ray_b200_upload_tables(c2ray_ctx* c, int32_t s, const c2ray_sed_tables* t) {
  if (!c || !t || s < 0 || s > 2) return fail(C2RAY_ERR_ARG, "bad argument");
  CK(cudaSetDevice(c->device));
  c->run_dirty = true;
  if (!t->photo_thick) {  // SED absent
    for (int k = 0; k < 4; k++) if (c->tab[s][k]) { cudaFree(c->tab[s][k]); c->tab[s][k] = nullptr; }
    c->lo[s] = 1; c->hi[s] = 0;
    return pack_tables(c, s);
  }
  if (!t->photo_thin) return fail(C2RAY_ERR_ARG, "photo_thin missing");
  // the band limits index the packed tables and the band records on the device (radiation_tables.f90:194-247)
  if (t->freqbnd_lower < 1 || t->freqbnd_upper > NumFreqBnd || t->freqbnd_lower > t->freqbnd_upper + 1)
    return fail(C2RAY_ERR_ARG, "FreqBnd limits must satisfy 1 <= lower, upper <= NumFreqBnd (47), lower <= upper + 1");
  if (!std::isfinite(t->S_star) || t->S_star < 0.0) return fail(C2RAY_ERR_ARG, "S_star must be finite and non-negative");
  const bool heat = t->heat_thick && t->heat_thin;
  if (!heat && !c->par.isothermal) return fail(C2RAY_ERR_ARG, "heating tables required unless isothermal");
  int rc = ensure_tables(c, s, heat);
  if (rc) return rc;
  const size_t np = (size_t)NumFreqBnd * (NumTau + 1) * 8, nh = (size_t)NumheatBin * (NumTau + 1) * 8;
  CK(cudaMemcpy(c->tab[s][0], t->photo_thick, np, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(c->tab[s][1], t->photo_thin, np, cudaMemcpyHostToDevice));
  if (heat) {
    CK(cudaMemcpy(c->tab[s][2], t->heat_thick, nh, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->tab[s][3], t->heat_thin, nh, cudaMemcpyHostToDevice));
  }
  c->lo[s] = t->freqbnd_lower; c->hi[s] = t->freqbnd_upper; c->S_star[s] = t->S_star;
  return pack_tables(c, s);
}

int c2ray_b200_download_table(c2ray_ctx* c, int32_t s, int32_t kind, double* out, int32_t* lower, int32_t* upper,
                              double* S_star) {
  if (!c || s < 0 || s > 2 || kind < 0 || kind > 3) return fail(C2RAY_ERR_ARG, "bad argument");
  CK(cudaSetDevice(c->device));
  if (lower) *lower = c->lo[s];
  if (upper) *upper = c->tab[s][0] ? c->hi[s] : c->lo[s] - 1;
  if (S_star) *S_star = c->S_star[s];
  if (out) {
    if (!c->tab[s][kind]) return fail(C2RAY_ERR_STATE, "table not present");
    const size_t n = (size_t)(kind < 2 ? NumFreqBnd : NumheatBin) * (NumTau + 1) * 8;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out, c->tab[s][kind], n, cudaMemcpyDeviceToHost));
  }
  return C2RAY_OK;
}

// radiation_tables.f90:141 rad_ini with nominal-value SEDs (radiation_sed_parameters.f90:208-244, :637-742)
int c2ray_b200_rad_ini(c2ray_ctx* c, const c2ray_sed_params* sp) {
  if (!c || !sp) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  std::vector<TableBuild> tb_store(1);  // per call (the struct holds this context's device pointers)
  TableBuild& tb = tb_store[0];
  memset(&tb, 0, sizeof(tb));
  double fmin[NumFreqBnd], fmax[NumFreqBnd], dfreq[NumFreqBnd];
  band_edges(fmin, fmax, dfreq);
  romberg_weights(9, tb.romw);
  // power-law index of the cross-section used for each band (radiation_tables.f90:280-340): HI, HeI, HeII
  tb.plidx[0] = BD_PLIDX_HI_B1;
  for (int i = 0; i < 26; i++) tb.plidx[NumBndin1 + i] = BD_PLIDX_HEI_B2[i];
  for (int i = 0; i < 20; i++) tb.plidx[NumBndin1 + NumBndin2 + i] = BD_PLIDX_HEII_B3[i];
  for (int q = 0; q < NumFreqBnd; q++) { tb.freq_min[q] = fmin[q]; tb.delta_freq[q] = dfreq[q]; }
  const double h_over_kT = hplanck / (k_B * sp->T_eff);
  // integrate_sed over one 512-interval grid (radiation_sed_parameters.f90:746-800)
  auto quad = [&](double a, double b, auto fn) {
    const double step = (b - a) / (double)NumFreq;
    double acc = 0.0;
    for (int i = 0; i <= NumFreq; i++) acc = acc + fn(a + step * (double)i) * step * tb.romw[i];
    return acc;
  };
  auto bb = [&](double f) {
    const double x = f * h_over_kT;
    if (x <= 709.0) return two_pi_over_c_square * f * f / (exp(x) - 1.0);
    return two_pi_over_c_square * f * f / exp(x / 2.0) / exp(x / 2.0);
  };
  // normalize_blackbody :637-675 with S_star specified: R_star scaled so that the BB emits S_star photons/s
  double R_star = R_SOLAR;
  const double S_unscaled = FL(4.0f) * pi * R_star * R_star * quad(fmin[0], fmax[NumFreqBnd - 1], bb);
  R_star = sqrt(sp->S_star / S_unscaled) * R_star;
  tb.R_star2 = R_star * R_star;
  tb.h_over_kT = h_over_kT;
  tb.isothermal = c->par.isothermal;
  tb.active[0] = 1; tb.active[1] = sp->pl_S_star > 0; tb.active[2] = sp->qpl_S_star > 0;
  const double idx[3] = {0, sp->pl_index, sp->qpl_index};
  const double lof[3] = {0, sp->pl_minfreq, sp->qpl_minfreq}, hif[3] = {0, sp->pl_maxfreq, sp->qpl_maxfreq};
  const double Ss[3] = {sp->S_star, sp->pl_S_star, sp->qpl_S_star};
  c->lo[0] = 1; c->hi[0] = NumFreqBnd;
  for (int b = 1; b <= NumFreqBnd; b++)  // radiation_tables.f90:194-199
    if (fmin[b - 1] * h_over_kT > FL(25.f)) { c->hi[0] = b - 1; break; }
  for (int s = 1; s < 3; s++) {
    if (!tb.active[s]) { c->lo[s] = 1; c->hi[s] = 0; continue; }
    const double ix = idx[s];
    tb.index[s] = ix;
    tb.scaling[s] = Ss[s] / quad(lof[s], hif[s], [&](double f) { return pow(f, -ix); });  // :695-699 / :727-731
    c->hi[s] = NumFreqBnd;  // radiation_tables.f90:208-247
    for (int b = 1; b <= NumFreqBnd; b++) if (fmin[b - 1] > hif[s]) { c->hi[s] = b - 1; break; }
    c->lo[s] = 1;
    for (int b = NumFreqBnd; b >= 1; b--) if (fmin[b - 1] < lof[s]) { c->lo[s] = b; break; }
  }
  for (int s = 0; s < 3; s++) {
    c->S_star[s] = Ss[s];
    if (!tb.active[s]) {
      for (int k = 0; k < 4; k++) if (c->tab[s][k]) { cudaFree(c->tab[s][k]); c->tab[s][k] = nullptr; }
      continue;
    }
    int rc = ensure_tables(c, s, !c->par.isothermal);
    if (rc) return rc;
    tb.photo_thick[s] = c->tab[s][0]; tb.photo_thin[s] = c->tab[s][1];
    tb.heat_thick[s] = c->tab[s][2]; tb.heat_thin[s] = c->tab[s][3];
  }
  CK(cudaMemcpyAsync(c->d_tb, &tb, sizeof(tb), cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, k_build_tables, dim3(NumTau + 1, NumFreqBnd), 128, c->d_tb);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  c->run_dirty = true;
  for (int s = 0; s < 3; s++) {
    int rc = pack_tables(c, s);
    if (rc) return rc;
  }
  return C2RAY_OK;
}

int c2ray_b200_set_sources(c2ray_ctx* c, int32_t NumSrc, const int32_t* srcpos, const double* nf, const double* nfpl,
                           const double* nfqpl) {
  if (!c || NumSrc < 0 || (NumSrc > 0 && (!srcpos || !nf))) return fail(C2RAY_ERR_ARG, "bad argument");
  // srcpos indexes the mesh on the device (1-based, sourceprops_test.F90:93-106): anything else would make the sweep
  // read and atomically add outside the grids.  Fluxes must be finite and non-negative (they scale every rate).
  for (int i = 0; i < NumSrc; i++) {
    for (int d = 0; d < 3; d++)
      if (srcpos[3 * i + d] < 1 || srcpos[3 * i + d] > c->mesh[d])
        return fail(C2RAY_ERR_ARG, "srcpos(" + std::to_string(d + 1) + "," + std::to_string(i + 1) + ") = " +
                                       std::to_string(srcpos[3 * i + d]) + " outside 1..mesh (1-based mesh positions expected)");
    const double f3[3] = {nf[i], nfpl ? nfpl[i] : 0.0, nfqpl ? nfqpl[i] : 0.0};
    for (double f : f3)
      if (!(f >= 0.0) || !std::isfinite(f))
        return fail(C2RAY_ERR_ARG, "NormFlux of source " + std::to_string(i + 1) + " is negative or not finite");
  }
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  for (double** p : {&c->d_nf, &c->d_nfpl, &c->d_nfqpl}) if (*p) { cudaFree(*p); *p = nullptr; }
  if (c->d_srcpos) { cudaFree(c->d_srcpos); c->d_srcpos = nullptr; }
  c->NumSrc = NumSrc;
  c->have_pl_flux = nfpl != nullptr; c->have_qpl_flux = nfqpl != nullptr;
  c->sum_nf[0] = c->sum_nf[1] = c->sum_nf[2] = 0.0;
  c->last_pass_updates = -1;
  c->src_pl.assign(NumSrc, 0); c->src_qpl.assign(NumSrc, 0);
  c->src_flux.assign(NumSrc, 0.0);
  for (int i = 0; i < NumSrc; i++) {
    c->src_flux[i] = nf[i] * c->S_star[0] + (nfpl ? nfpl[i] * c->S_star[1] : 0.0) + (nfqpl ? nfqpl[i] * c->S_star[2] : 0.0);
    c->sum_nf[0] += nf[i];
    if (nfpl) { c->sum_nf[1] += nfpl[i]; c->src_pl[i] = nfpl[i] > 0.0; }
    if (nfqpl) { c->sum_nf[2] += nfqpl[i]; c->src_qpl[i] = nfqpl[i] > 0.0; }
  }
  if (NumSrc > 0) {
    CK(cudaMalloc(&c->d_srcpos, sizeof(int) * 3 * NumSrc));
    CK(cudaMemcpy(c->d_srcpos, srcpos, sizeof(int) * 3 * NumSrc, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c->d_nf, 8 * (size_t)NumSrc));
    CK(cudaMemcpy(c->d_nf, nf, 8 * (size_t)NumSrc, cudaMemcpyHostToDevice));
    if (nfpl) { CK(cudaMalloc(&c->d_nfpl, 8 * (size_t)NumSrc)); CK(cudaMemcpy(c->d_nfpl, nfpl, 8 * (size_t)NumSrc, cudaMemcpyHostToDevice)); }
    if (nfqpl) { CK(cudaMalloc(&c->d_nfqpl, 8 * (size_t)NumSrc)); CK(cudaMemcpy(c->d_nfqpl, nfqpl, 8 * (size_t)NumSrc, cudaMemcpyHostToDevice)); }
  }
  return rebuild_my_sources(c);
}

int c2ray_b200_set_geometry(c2ray_ctx* c, const double dr[3], double vol, double zred) {
  if (!c || !dr) return fail(C2RAY_ERR_ARG, "null argument");
  for (int d = 0; d < 3; d++) c->dr[d] = dr[d];
  c->vol = vol; c->zred = zred;
  c->run_dirty = true;
  return C2RAY_OK;
}

int c2ray_b200_set_clumping_grid(c2ray_ctx* c, const float* clumping_grid) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  if (!clumping_grid) {  // back to the scalar of c2ray_params (type_of_clumping 1-4)
    if (c->d_clump) { CK(cudaFree(c->d_clump)); c->d_clump = nullptr; }
    return C2RAY_OK;
  }
  if (!c->d_clump) CK(cudaMalloc(&c->d_clump, c->N3 * sizeof(float)));
  CK(cudaMemcpy(c->d_clump, clumping_grid, c->N3 * sizeof(float), cudaMemcpyHostToDevice));
  return C2RAY_OK;
}

int c2ray_b200_set_LLS(c2ray_ctx* c, int32_t type_of_LLS, double coldensh_LLS, const float* LLS_grid) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  if (type_of_LLS < 0 || type_of_LLS > 2) return fail(C2RAY_ERR_ARG, "type_of_LLS must be 0 (none), 1 (one value) or 2 (LLS_grid)");
  if (type_of_LLS == 2 && !LLS_grid) return fail(C2RAY_ERR_ARG, "type_of_LLS = 2 needs LLS_grid");
  CK(cudaSetDevice(c->device));
  if (type_of_LLS == 2) {
    if (!c->d_lls) CK(cudaMalloc(&c->d_lls, c->N3 * sizeof(float)));
    CK(cudaMemcpy(c->d_lls, LLS_grid, c->N3 * sizeof(float), cudaMemcpyHostToDevice));
  } else if (c->d_lls) {
    CK(cudaFree(c->d_lls)); c->d_lls = nullptr;
  }
  c->lls_type = type_of_LLS;
  c->coldensh_LLS = type_of_LLS == 1 ? coldensh_LLS : 0.0;
  return C2RAY_OK;
}

int c2ray_b200_set_state(c2ray_ctx* c, const double* ndens, const double* xh, const double* xhe, const float* temp) {
  if (!c || !ndens || !xh || !xhe) return fail(C2RAY_ERR_ARG, "null argument");
  if (!temp && !c->par.isothermal) return fail(C2RAY_ERR_ARG, "temperature_grid required unless isothermal");
  CK(cudaSetDevice(c->device));
  const size_t N3 = c->N3;
  CK(cudaMemcpyAsync(c->ndens, ndens, N3 * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->xh, xh, 2 * N3 * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->xhe, xhe, 3 * N3 * 8, cudaMemcpyHostToDevice, c->stream));
  if (temp) CK(cudaMemcpyAsync(c->temp, temp, 3 * N3 * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return C2RAY_OK;
}

int c2ray_b200_get_state(c2ray_ctx* c, double* xh, double* xhe, float* temp) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  const size_t N3 = c->N3;
  if (xh) CK(cudaMemcpyAsync(xh, c->xh, 2 * N3 * 8, cudaMemcpyDeviceToHost, c->stream));
  if (xhe) CK(cudaMemcpyAsync(xhe, c->xhe, 3 * N3 * 8, cudaMemcpyDeviceToHost, c->stream));
  if (temp) CK(cudaMemcpyAsync(temp, c->temp, 3 * N3 * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return C2RAY_OK;
}

int c2ray_b200_get_rates(c2ray_ctx* c, double* phih, double* phihe, double* phiheat) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  const size_t N3 = c->N3;
  if (phih) CK(cudaMemcpyAsync(phih, c->rates, N3 * 8, cudaMemcpyDeviceToHost, c->stream));
  if (phihe) CK(cudaMemcpyAsync(phihe, c->rates + N3, 2 * N3 * 8, cudaMemcpyDeviceToHost, c->stream));
  if (phiheat) CK(cudaMemcpyAsync(phiheat, c->rates + 3 * N3, N3 * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return C2RAY_OK;
}

int c2ray_b200_set_rates(c2ray_ctx* c, const double* phih, const double* phihe, const double* phiheat) {
  if (!c || !phih || !phihe) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  const size_t N3 = c->N3;
  CK(cudaMemcpyAsync(c->rates, phih, N3 * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->rates + N3, phihe, 2 * N3 * 8, cudaMemcpyHostToDevice, c->stream));
  if (phiheat) CK(cudaMemcpyAsync(c->rates + 3 * N3, phiheat, N3 * 8, cudaMemcpyHostToDevice, c->stream));
  else CK(cudaMemsetAsync(c->rates + 3 * N3, 0, N3 * 8, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return C2RAY_OK;
}

int c2ray_b200_get_work_state(c2ray_ctx* c, double* xh_av, double* xhe_av, double* xh_int, double* xhe_int) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  const size_t N3 = c->N3;
  if (xh_av) CK(cudaMemcpyAsync(xh_av, c->xh_av, 2 * N3 * 8, cudaMemcpyDeviceToHost, c->stream));
  if (xhe_av) CK(cudaMemcpyAsync(xhe_av, c->xhe_av, 3 * N3 * 8, cudaMemcpyDeviceToHost, c->stream));
  if (xh_int) CK(cudaMemcpyAsync(xh_int, c->xh_int, 2 * N3 * 8, cudaMemcpyDeviceToHost, c->stream));
  if (xhe_int) CK(cudaMemcpyAsync(xhe_int, c->xhe_int, 3 * N3 * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return C2RAY_OK;
}

int c2ray_b200_set_work_state(c2ray_ctx* c, const double* xh_av, const double* xhe_av, const double* xh_int,
                              const double* xhe_int) {
  if (!c || !xh_av || !xhe_av || !xh_int || !xhe_int) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  const size_t N3 = c->N3;
  CK(cudaMemcpyAsync(c->xh_av, xh_av, 2 * N3 * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->xhe_av, xhe_av, 3 * N3 * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->xh_int, xh_int, 2 * N3 * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->xhe_int, xhe_int, 3 * N3 * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return C2RAY_OK;
}

int c2ray_b200_snapshot_state(c2ray_ctx* c) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  const size_t N3 = c->N3;
  if (!c->snap_xh) { CK(cudaMalloc(&c->snap_xh, 2 * N3 * 8)); CK(cudaMalloc(&c->snap_xhe, 3 * N3 * 8)); CK(cudaMalloc(&c->snap_temp, 3 * N3 * 4)); }
  CK(cudaMemcpyAsync(c->snap_xh, c->xh, 2 * N3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->snap_xhe, c->xhe, 3 * N3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->snap_temp, c->temp, 3 * N3 * 4, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return C2RAY_OK;
}
int c2ray_b200_restore_state(c2ray_ctx* c) {
  if (!c || !c->snap_xh) return fail(C2RAY_ERR_STATE, "no snapshot");
  CK(cudaSetDevice(c->device));
  const size_t N3 = c->N3;
  CK(cudaMemcpyAsync(c->xh, c->snap_xh, 2 * N3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->xhe, c->snap_xhe, 3 * N3 * 8, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->temp, c->snap_temp, 3 * N3 * 4, cudaMemcpyDeviceToDevice, c->stream));
  return C2RAY_OK;
}

int c2ray_b200_begin_step(c2ray_ctx* c) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  return begin_step(c);
}
int c2ray_b200_end_step(c2ray_ctx* c) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  return end_step(c);
}
int c2ray_b200_set_rates_to_zero(c2ray_ctx* c) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  CK(cudaMemsetAsync(c->rates, 0, c->rates_count * 8, c->stream));  // evolve.F90:371-381
  return C2RAY_OK;
}

int c2ray_b200_pass_all_sources(c2ray_ctx* c, double /*dt*/, int32_t /*niter*/, int64_t* rt_updates) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  if (c->NumSrc > 0 && !c->tab[0][0] && !c->tab[1][0] && !c->tab[2][0]) return fail(C2RAY_ERR_STATE, "no radiation tables");
  int rc = sweep_all(c);
  if (rc) return rc;
  SweepTotals t;
  CK(cudaMemcpyAsync(&t, c->d_tot, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
  rc = allreduce_rates(c);
  if (rc) return rc;
  CK(cudaStreamSynchronize(c->stream));
  if ((rc = rebalance_sources(c))) return rc;
  c->last_pass_updates = (long long)t.updates;
  if (rt_updates) *rt_updates = (int64_t)t.updates;
  return C2RAY_OK;
}

int c2ray_b200_do_source(c2ray_ctx* c, double /*dt*/, int32_t ns1, int32_t /*niter*/, int32_t* nbox, double* loss) {
  if (!c || ns1 < 1 || ns1 > c->NumSrc) return fail(C2RAY_ERR_ARG, "bad source number");
  // temporarily trace just this source
  CK(cudaSetDevice(c->device));
  std::vector<int> saved;
  saved.swap(c->my_ids);
  const int saved_n = c->n_mine;
  c->my_ids.assign(1, ns1 - 1); c->n_mine = 1; c->run_key = -1;
  int rc = sweep_all(c);
  SweepTotals t;
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(&t, c->d_tot, sizeof(t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = fail(C2RAY_ERR_CUDA, std::string("do_source: ") + cudaGetErrorString(e));
  }
  c->my_ids.swap(saved); c->n_mine = saved_n; c->run_key = -1;
  if (rc) return rc;
  if (nbox) *nbox = (int)t.sum_nbox;
  if (loss) *loss = t.photon_loss;
  return C2RAY_OK;
}

int c2ray_b200_global_pass(c2ray_ctx* c, double dt, int32_t* conv_flag, int32_t* nit_out) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  if (nit_out && !c->d_nit) CK(cudaMalloc(&c->d_nit, c->N3 * sizeof(int)));
  int rc = global_pass_launch(c, dt, nit_out ? c->d_nit : nullptr);
  if (rc) return rc;
  ChemTotals t;
  CK(cudaMemcpyAsync(&t, c->d_chem, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
  if (nit_out) CK(cudaMemcpyAsync(nit_out, c->d_nit, c->N3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->last_nsub_per_cell = (double)t.nsub_total / (double)c->N3;
  c->last_nit_per_cell = (double)t.nit_total / (double)c->N3;
  if (conv_flag) *conv_flag = t.conv_flag;
  return C2RAY_OK;
}

int c2ray_b200_state_sums(c2ray_ctx* c, int32_t which, double out5[5]) {
  if (!c || !out5) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  return which == 0 ? state_sums(c, c->xh, c->xhe, out5) : state_sums(c, c->xh_int, c->xhe_int, out5);
}

int c2ray_b200_evolve3d(c2ray_ctx* c, double /*time*/, double dt, int32_t restart, c2ray_stats* st) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  if (restart < 0 || restart > 3) return fail(C2RAY_ERR_ARG, "restart must be 0 (fresh step) or 1, 2, 3 (iterdump1.bin, iterdump2.bin, iterdump.bin; evolve.F90:300-307)");
  if (restart != 0 && c->dump_dir.empty()) return fail(C2RAY_ERR_STATE, "restart requested but no dump directory set (c2ray_b200_set_dump)");
  auto wallclock1 = std::chrono::steady_clock::now();  // evolve.F90:122
  if (c->NumSrc > 0 && !c->tab[0][0] && !c->tab[1][0] && !c->tab[2][0]) return fail(C2RAY_ERR_STATE, "no radiation tables");
  CK(cudaSetDevice(c->device));
  c2ray_stats S;
  memset(&S, 0, sizeof(S));
  int rc;
  if ((rc = state_sums(c, c->xh, c->xhe, S.sums_before))) return rc;  // evolve.F90:127
  int niter = 0;
  // conv_flag=mesh(1)*mesh(2)*mesh(3) ; conv_criterion=min(int(convergence_fraction*mesh1*mesh2*mesh3),NumSrc)  :136,:147
  int conv_flag = c->mesh[0] * c->mesh[1] * c->mesh[2];
  if (restart == 0) {
    if ((rc = begin_step(c))) return rc;
  }
  const int conv_criterion = std::min((int)(convergence_fraction * c->mesh[0] * c->mesh[1] * c->mesh[2]), c->NumSrc);
  float ms;
  SweepTotals swt;
  memset(&swt, 0, sizeof(swt));
  ChemTotals cht;
  memset(&cht, 0, sizeof(cht));
  const bool split = split_active(c);
  const size_t chunk = split ? c->N3 / c->npr : c->N3;
  // the global pass of one iteration: on this rank's cells + counters + the next sweep's inputs when split over the
  // ranks, on the whole mesh otherwise.  Records ev[4] after the chemistry kernel.
  auto chem_phase = [&]() -> int {
    int r;
    if (split) {
      if ((r = global_pass_launch(c, dt, nullptr, (size_t)c->rank * chunk, (size_t)(c->rank + 1) * chunk))) return r;  // :217
      CK(cudaEventRecord(c->ev[4], c->stream));
      if ((r = combine_chem_totals(c))) return r;
      if ((r = allgather_sweep_inputs(c))) return r;
    } else {
      if ((r = global_pass_launch(c, dt, nullptr))) return r;  // :217
      CK(cudaEventRecord(c->ev[4], c->stream));
    }
    CK(cudaMemcpyAsync(&cht, c->d_chem, sizeof(cht), cudaMemcpyDeviceToHost, c->stream));
    return 0;
  };
  if (restart != 0) {
    // evolve.F90:137-141: reload xh_av, xh_intermed, rates, photon_loss, niter ; call global_pass (conv_flag,dt)
    static const char* names[4] = {"", "iterdump1.bin", "iterdump2.bin", "iterdump.bin"};
    if ((rc = read_iteration_dump_file(c, join_dir(c->dump_dir, names[restart]), &niter))) return rc;
    if ((rc = chem_phase())) return rc;
    CK(cudaStreamSynchronize(c->stream));
    conv_flag = cht.conv_flag;
    c->last_nsub_per_cell = (double)cht.nsub_total / (double)c->N3;
    c->last_nit_per_cell = (double)cht.nit_total / (double)c->N3;
    S.chem_cells += (int64_t)chunk;
  }
  for (;;) {
    const bool converged = conv_flag < conv_criterion && niter > 1;  // :163
    if (converged || niter > 500) {                                  // :177
      if (split && (rc = allgather_final(c))) return rc;
      if (converged && (rc = end_step(c))) return rc;
      break;
    }
    niter++;
    CK(cudaEventRecord(c->ev[1], c->stream));
    CK(cudaMemsetAsync(c->rates, 0, c->rates_count * 8, c->stream));  // :188
    if (c->NumSrc > 0) {
      if ((rc = sweep_all(c))) return rc;  // :192
      CK(cudaMemcpyAsync(&swt, c->d_tot, sizeof(swt), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaEventRecord(c->ev[2], c->stream));
    if (c->NumSrc > 0) {
      if (split) rc = reduce_scatter_rates(c); else rc = allreduce_rates(c);
      if (rc) return rc;
      if ((rc = rebalance_sources(c))) return rc;  // balanced schedule: re-deal the sources for the next pass
      // evolve.F90:199-213: rank 0 writes an iteration dump when more than the interval (15 minutes) has passed
      if (c->dump_interval_s >= 0.0 && !c->dump_dir.empty()) {
        int due = 0;
        const auto wallclock2 = std::chrono::steady_clock::now();
        if (c->rank == 0 && std::chrono::duration<double>(wallclock2 - wallclock1).count() > c->dump_interval_s) due = 1;
        if (c->comm && c->npr > 1) {  // rank 0's clock decides for everybody
          int* d_flag = reinterpret_cast<int*>(c->d_chemred + 5);
          CK(cudaMemcpyAsync(d_flag, &due, sizeof(int), cudaMemcpyHostToDevice, c->stream));
          NC(g_nccl.AllReduce(d_flag, d_flag, 1, NCCL_INT32, NCCL_SUM, c->comm, c->stream));
          CK(cudaMemcpyAsync(&due, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
          CK(cudaStreamSynchronize(c->stream));
        }
        if (due) {
          if (split && (rc = allgather_final(c))) return rc;  // rank 0 needs every cell of what it writes
          if (c->rank == 0) {
            c->ndump++;  // :243-248
            if ((rc = write_iteration_dump_file(c, join_dir(c->dump_dir, c->ndump % 2 == 0 ? "iterdump2.bin" : "iterdump1.bin"), niter))) return rc;
          }
          wallclock1 = wallclock2;
        }
      }
    }
    CK(cudaEventRecord(c->ev[3], c->stream));
    if ((rc = chem_phase())) return rc;
    CK(cudaEventRecord(c->ev[0], c->stream));  // iteration end
    CK(cudaStreamSynchronize(c->stream));
    conv_flag = cht.conv_flag;
    c->last_nsub_per_cell = (double)cht.nsub_total / (double)c->N3;
    c->last_nit_per_cell = (double)cht.nit_total / (double)c->N3;
    if (niter <= C2RAY_MAX_ITER_HIST) S.conv_hist[niter - 1] = conv_flag;
    S.rt_updates += (int64_t)swt.updates;
    if (c->NumSrc > 0) c->last_pass_updates = (long long)swt.updates;
    S.chem_cells += (int64_t)chunk;
    CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); S.ms_sweep += ms;
    CK(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3])); S.ms_allreduce += ms;
    CK(cudaEventElapsedTime(&ms, c->ev[3], c->ev[4])); S.ms_chem += ms;
    CK(cudaEventElapsedTime(&ms, c->ev[4], c->ev[0])); S.ms_allreduce += ms;  // counters + all-gather (split pass only)
  }
  if ((rc = state_sums(c, c->xh, c->xhe, S.sums_after))) return rc;  // :225 (state_after on the final xh)
  {
    // total_rates(dt, xh_av, xhe_av) with the coefficients the reference's module globals hold at this point
    const double coef_T = c->par.isothermal ? c->par.temper_val : cht.last_coef_T;
    CK(cudaMemsetAsync(c->d_sums, 0, 3 * sizeof(double), c->stream));
    if (coef_T > 0.0) LAUNCH(c, k_total_rates, 148 * 4, 256, c->ndens, c->xh_av, c->xhe_av, c->N3, coef_T, c->d_sums, c->d_clump);
    double t3[3];
    CK(cudaMemcpyAsync(t3, c->d_sums, sizeof(t3), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    S.totrec = t3[0] * c->vol * dt; S.totcollisions = t3[1] * c->vol * dt; S.recomions = t3[2] * c->vol * dt;   // :201-203
    // total_ionizations :251-260
    S.total_ion = (S.sums_before[0] - S.sums_after[0]) + (S.sums_before[2] - S.sums_after[2]) + (S.sums_after[4] - S.sums_before[4]);
    // report_photonstatistics :280-291
    S.totalsrc = c->sum_nf[0] * c->S_star[0] * dt;
    if (c->tab[1][0]) S.totalsrc = S.totalsrc + c->sum_nf[1] * c->S_star[1] * dt;
    if (c->tab[2][0]) S.totalsrc = S.totalsrc + c->sum_nf[2] * c->S_star[2] * dt;
    S.photcons = S.totalsrc > 0.0 ? (S.total_ion - S.totcollisions - S.recomions) / S.totalsrc : 0.0;
  }
  CK(cudaStreamSynchronize(c->stream));
  S.ms_total = S.ms_sweep + S.ms_allreduce + S.ms_chem;
  S.niter = niter; S.conv_flag = conv_flag; S.conv_criterion = conv_criterion;
  S.nit_max = cht.nit_max; S.nit_total = (int64_t)cht.nit_total;
  // after the allreduce the tail of the rate buffer holds the global sums
  double tail[NumFreqBnd + 1];
  CK(cudaMemcpy(tail, c->rates + 4 * c->N3, sizeof(tail), cudaMemcpyDeviceToHost));
  S.photon_loss_all = tail[0];
  S.sum_nbox_all = (int64_t)llround(tail[NumFreqBnd]);
  if (st) *st = S;
  return C2RAY_OK;
}

int c2ray_b200_set_dump(c2ray_ctx* c, const char* dump_dir, double interval_s) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  c->dump_dir = dump_dir ? dump_dir : "";
  c->dump_interval_s = interval_s;
  return C2RAY_OK;
}

int c2ray_b200_write_iteration_dump(c2ray_ctx* c, const char* path, int32_t niter) {
  if (!c || !path) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  return write_iteration_dump_file(c, path, niter);
}

int c2ray_b200_read_iteration_dump(c2ray_ctx* c, const char* path, int32_t* niter) {
  if (!c || !path) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  int ni = 0;
  int rc = read_iteration_dump_file(c, path, &ni);
  if (rc) return rc;
  if (niter) *niter = ni;
  return C2RAY_OK;
}

int c2ray_b200_write_stream2(c2ray_ctx* c, const char* results_dir, double zred_now) {
  if (!c || !results_dir) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  if (c->rank != 0) return C2RAY_OK;  // output.F90:260 if (rank == 0)
  const std::string z = c2io::f6_3(zred_now), dir = results_dir;
  int rc;
  if ((rc = write_plane_file(c, join_dir(dir, "xfrac3d_" + z + ".bin"), c->xh + c->N3, 8, false))) return rc;      // xh(:,:,:,1)
  if ((rc = write_plane_file(c, join_dir(dir, "xfrac3dHe1_" + z + ".bin"), c->xhe + c->N3, 8, false))) return rc;  // xhe(:,:,:,1)
  return write_plane_file(c, join_dir(dir, "xfrac3dHe2_" + z + ".bin"), c->xhe + 2 * c->N3, 8, false);             // xhe(:,:,:,2)
}

int c2ray_b200_write_stream3(c2ray_ctx* c, const char* results_dir, double zred_now) {
  if (!c || !results_dir) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  if (c->rank != 0) return C2RAY_OK;  // output.F90:323
  const std::string z = c2io::f6_3(zred_now), dir = results_dir;
  int rc;
  if (!c->par.isothermal &&
      (rc = write_plane_file(c, join_dir(dir, "Temper3D_" + z + ".bin"), c->temp, 4, false))) return rc;  // real(temperature_grid(i,j,k,0))
  if ((rc = write_plane_file(c, join_dir(dir, "IonRates3D_" + z + ".bin"), c->rates, 8, true))) return rc;  // real(phih_grid)
  return write_plane_file(c, join_dir(dir, "HeatRates3D_" + z + ".bin"), c->rates + 3 * c->N3, 8, true);    // real(phiheat)
}

int c2ray_b200_fortran_records_write(const char* path, int32_t n, const void* const* data, const int64_t* bytes,
                                     int64_t max_subrecord) {
  if (!path || n < 0 || (n > 0 && (!data || !bytes))) return fail(C2RAY_ERR_ARG, "bad argument");
  c2io::RecordWriter w(max_subrecord > 0 ? (uint64_t)max_subrecord : c2io::MAX_SUBRECORD);
  if (!w.open(path)) return fail(C2RAY_ERR_STATE, std::string("cannot open ") + path + " for writing");
  for (int i = 0; i < n; i++)
    if (bytes[i] < 0 || !w.record(data[i], (uint64_t)bytes[i])) return fail(C2RAY_ERR_STATE, "short write");
  if (!w.close()) return fail(C2RAY_ERR_STATE, "error closing file");
  return C2RAY_OK;
}

int c2ray_b200_fortran_records_read(const char* path, int32_t n, void* const* data, const int64_t* bytes) {
  if (!path || n < 0 || (n > 0 && (!data || !bytes))) return fail(C2RAY_ERR_ARG, "bad argument");
  c2io::RecordReader r;
  if (!r.open(path)) return fail(C2RAY_ERR_STATE, std::string("cannot open ") + path);
  for (int i = 0; i < n; i++)
    if (bytes[i] < 0 || !r.record(data[i], (uint64_t)bytes[i]))
      return fail(C2RAY_ERR_STATE, "record " + std::to_string(i) + " does not have the expected length or the file is corrupt");
  return C2RAY_OK;
}

int c2ray_b200_evolve3d_host(c2ray_ctx* c, double time, double dt, int32_t restart, const double* ndens, double* xh,
                             double* xhe, float* temp, c2ray_stats* st) {
  int rc = c2ray_b200_set_state(c, ndens, xh, xhe, temp);
  if (rc) return rc;
  rc = c2ray_b200_evolve3d(c, time, dt, restart, st);
  if (rc) return rc;
  return c2ray_b200_get_state(c, xh, xhe, (c->par.isothermal || !temp) ? nullptr : temp);
}

// ---- parity hooks -------------------------------------------------------------------------------
int c2ray_b200_photoion_rates_batch(c2ray_ctx* c, int32_t n, const double* col6, const double* vol, const double nflux3[3],
                                    const double* i_state, double* out6) {
  if (!c || n <= 0 || !col6 || !vol || !nflux3 || !i_state || !out6) return fail(C2RAY_ERR_ARG, "bad argument");
  int rc = bind(c);
  if (rc) return rc;
  double *d_col, *d_vol, *d_is, *d_out;
  CK(cudaMalloc(&d_col, 48 * (size_t)n)); CK(cudaMalloc(&d_vol, 8 * (size_t)n)); CK(cudaMalloc(&d_is, 8 * (size_t)n));
  CK(cudaMalloc(&d_out, 48 * (size_t)n));
  CK(cudaMemcpy(d_col, col6, 48 * (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_vol, vol, 8 * (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_is, i_state, 8 * (size_t)n, cudaMemcpyHostToDevice));
  LAUNCH(c, k_photoion_batch, (n + 127) / 128, 128, n, d_col, d_vol, nflux3[0], nflux3[1], nflux3[2], d_is, d_out);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(out6, d_out, 48 * (size_t)n, cudaMemcpyDeviceToHost));
  cudaFree(d_col); cudaFree(d_vol); cudaFree(d_is); cudaFree(d_out);
  return C2RAY_OK;
}

int c2ray_b200_chemistry_batch(c2ray_ctx* c, int32_t n, double dt, const double* ndens, double* ion15, const double* phi4,
                               double* T3, int32_t* nit_out) {
  if (!c || n <= 0 || !ndens || !ion15 || !phi4 || !T3 || !nit_out) return fail(C2RAY_ERR_ARG, "bad argument");
  int rc = bind(c);
  if (rc) return rc;
  if (!c->par.isothermal && !c->have_cool) return fail(C2RAY_ERR_STATE, "cooling tables not set");
  double *d_n, *d_ion, *d_phi, *d_T; int* d_nit;
  CK(cudaMalloc(&d_n, 8 * (size_t)n)); CK(cudaMalloc(&d_ion, 120 * (size_t)n)); CK(cudaMalloc(&d_phi, 32 * (size_t)n));
  CK(cudaMalloc(&d_T, 24 * (size_t)n)); CK(cudaMalloc(&d_nit, 4 * (size_t)n));
  CK(cudaMemcpy(d_n, ndens, 8 * (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ion, ion15, 120 * (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_phi, phi4, 32 * (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_T, T3, 24 * (size_t)n, cudaMemcpyHostToDevice));
  LAUNCH(c, k_chemistry_batch, (n + 127) / 128, 128, n, dt, d_n, d_ion, d_phi, d_T, d_nit);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(ion15, d_ion, 120 * (size_t)n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(T3, d_T, 24 * (size_t)n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(nit_out, d_nit, 4 * (size_t)n, cudaMemcpyDeviceToHost));
  cudaFree(d_n); cudaFree(d_ion); cudaFree(d_phi); cudaFree(d_T); cudaFree(d_nit);
  return C2RAY_OK;
}

int c2ray_b200_doric_batch(c2ray_ctx* c, int32_t n, double dt, const double* rhe, double* ion15, const double* phi3,
                           const double* fr4, const double* T) {
  if (!c || n <= 0 || !rhe || !ion15 || !phi3 || !fr4 || !T) return fail(C2RAY_ERR_ARG, "bad argument");
  int rc = bind(c);
  if (rc) return rc;
  const size_t N = (size_t)n;
  double* d = nullptr;   // [rhe | ion15 | phi3 | fr4 | T]
  CK(cudaMalloc(&d, 8 * N * (1 + 15 + 3 + 4 + 1)));
  double *d_rhe = d, *d_ion = d + N, *d_phi = d_ion + 15 * N, *d_fr = d_phi + 3 * N, *d_T = d_fr + 4 * N;
  CK(cudaMemcpy(d_rhe, rhe, 8 * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ion, ion15, 120 * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_phi, phi3, 24 * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_fr, fr4, 32 * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_T, T, 8 * N, cudaMemcpyHostToDevice));
  LAUNCH(c, k_doric_batch, (n + 127) / 128, 128, n, dt, d_rhe, d_ion, d_phi, d_fr, d_T);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(ion15, d_ion, 120 * N, cudaMemcpyDeviceToHost));
  cudaFree(d);
  return C2RAY_OK;
}

int c2ray_b200_thermal_batch(c2ray_ctx* c, int32_t n, double dt, double* end_temper, double* avg_temper, const double* ne,
                             const double* ndens, const double* ion15, const double* heat, int32_t* nsub) {
  if (!c || n <= 0 || !end_temper || !avg_temper || !ne || !ndens || !ion15 || !heat || !nsub) return fail(C2RAY_ERR_ARG, "bad argument");
  int rc = bind(c);
  if (rc) return rc;
  if (!c->have_cool) return fail(C2RAY_ERR_STATE, "cooling tables not set");
  const size_t N = (size_t)n;
  double* d = nullptr;   // [end | avg | ne | ndens | heat | ion15]
  int* d_ns = nullptr;
  CK(cudaMalloc(&d, 8 * N * (5 + 15)));
  CK(cudaMalloc(&d_ns, 4 * N));
  double *d_e = d, *d_a = d + N, *d_ne = d + 2 * N, *d_n = d + 3 * N, *d_h = d + 4 * N, *d_ion = d + 5 * N;
  CK(cudaMemcpy(d_e, end_temper, 8 * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_a, avg_temper, 8 * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ne, ne, 8 * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_n, ndens, 8 * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_h, heat, 8 * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ion, ion15, 120 * N, cudaMemcpyHostToDevice));
  LAUNCH(c, k_thermal_batch, (n + 127) / 128, 128, n, dt, d_e, d_a, d_ne, d_n, d_ion, d_h, d_ns);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(end_temper, d_e, 8 * N, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(avg_temper, d_a, 8 * N, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(nsub, d_ns, 4 * N, cudaMemcpyDeviceToHost));
  cudaFree(d); cudaFree(d_ns);
  return C2RAY_OK;
}

int c2ray_b200_rec_colion_batch(c2ray_ctx* c, int32_t n, const double* T, double* out12) {
  if (!c || n <= 0 || !T || !out12) return fail(C2RAY_ERR_ARG, "bad argument");
  int rc = bind(c);
  if (rc) return rc;
  double *d_T, *d_o;
  CK(cudaMalloc(&d_T, 8 * (size_t)n)); CK(cudaMalloc(&d_o, 96 * (size_t)n));
  CK(cudaMemcpy(d_T, T, 8 * (size_t)n, cudaMemcpyHostToDevice));
  LAUNCH(c, k_rec_colion_batch, (n + 127) / 128, 128, n, d_T, d_o);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(out12, d_o, 96 * (size_t)n, cudaMemcpyDeviceToHost));
  cudaFree(d_T); cudaFree(d_o);
  return C2RAY_OK;
}

int c2ray_b200_cinterp_batch(c2ray_ctx* c, int32_t n, const int32_t* pos, const int32_t srcpos[3], const double* cdh,
                             const double* cdhe, double* out4) {
  if (!c || n <= 0 || !pos || !srcpos || !cdh || !cdhe || !out4) return fail(C2RAY_ERR_ARG, "bad argument");
  int rc = bind(c);
  if (rc) return rc;
  const size_t N3 = c->N3;
  int* d_pos; double *d_h, *d_he, *d_o;
  CK(cudaMalloc(&d_pos, 12 * (size_t)n)); CK(cudaMalloc(&d_h, 8 * N3)); CK(cudaMalloc(&d_he, 16 * N3)); CK(cudaMalloc(&d_o, 32 * (size_t)n));
  CK(cudaMemcpy(d_pos, pos, 12 * (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_h, cdh, 8 * N3, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_he, cdhe, 16 * N3, cudaMemcpyHostToDevice));
  LAUNCH(c, k_cinterp_batch, (n + 127) / 128, 128, n, d_pos, srcpos[0], srcpos[1], srcpos[2], d_h, d_he, N3, d_o);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(out4, d_o, 32 * (size_t)n, cudaMemcpyDeviceToHost));
  cudaFree(d_pos); cudaFree(d_h); cudaFree(d_he); cudaFree(d_o);
  return C2RAY_OK;
}

// ---- multi-GPU ----------------------------------------------------------------------------------
int c2ray_b200_comm_unique_id(uint8_t id[128]) {
  int rc = nccl_load();
  if (rc) return rc;
  nccl_uid u;
  if (g_nccl.GetUniqueId(&u) != 0) return fail(C2RAY_ERR_NCCL, "ncclGetUniqueId failed");
  memcpy(id, u.internal, 128);
  return C2RAY_OK;
}

int c2ray_b200_comm_init(c2ray_ctx* c, const uint8_t id[128], int32_t rank, int32_t npr) {
  if (!c || !id || rank < 0 || npr < 1 || rank >= npr) return fail(C2RAY_ERR_ARG, "bad argument");
  int rc = nccl_load();
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  nccl_uid u;
  memcpy(u.internal, id, 128);
  int r = g_nccl.CommInitRank(&c->comm, npr, u, rank);
  if (r != 0) return fail(C2RAY_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  c->rank = rank; c->npr = npr;
  return rebuild_my_sources(c);
}

int c2ray_b200_set_rank(c2ray_ctx* c, int32_t rank, int32_t npr) {
  if (!c || rank < 0 || npr < 1 || rank >= npr) return fail(C2RAY_ERR_ARG, "bad argument");
  CK(cudaSetDevice(c->device));
  c->rank = rank; c->npr = npr;
  return rebuild_my_sources(c);
}

int c2ray_b200_set_source_schedule(c2ray_ctx* c, int32_t mode) {
  if (!c || mode < 0 || mode > 1) return fail(C2RAY_ERR_ARG, "schedule must be 0 (static round robin) or 1 (balanced)");
  c->schedule = mode;
  CK(cudaSetDevice(c->device));
  return rebuild_my_sources(c);  // static: round robin; balanced: dealt by photon rate until a pass has left cost records
}

int c2ray_b200_balanced_partition(int32_t NumSrc, const int64_t* cost, int32_t npr, int32_t* owner) {
  if (NumSrc < 0 || npr < 1 || (NumSrc > 0 && (!cost || !owner))) return fail(C2RAY_ERR_ARG, "bad argument");
  std::vector<long long> cst(cost, cost + NumSrc);
  std::vector<int> own(NumSrc);
  balanced_partition(NumSrc, cst.data(), npr, own.data());
  for (int i = 0; i < NumSrc; i++) owner[i] = own[i];
  return C2RAY_OK;
}

int c2ray_b200_my_sources(c2ray_ctx* c, int32_t* ids, int32_t cap, int32_t* n) {
  if (!c || !n) return fail(C2RAY_ERR_ARG, "null argument");
  *n = (int32_t)c->my_ids.size();
  if (ids) for (int i = 0; i < *n && i < cap; i++) ids[i] = c->my_ids[i] + 1;
  return C2RAY_OK;
}

// mrgrnk.f90:21-215 R_mrgrnk as the reference calls it on the source fluxes (ctrper.f90:108-113): IRNGT(i) = 1-based index
// of the i-th smallest element, equal keys in their original order (merge sort with `<=` comparisons is stable).
// A stable LSD radix sort of (key, index) pairs gives the same permutation bit for bit; -0.0 is folded onto +0.0
// first because Fortran compares them equal while their bit patterns differ.  Off the hot path.
__global__ void k_mrgrnk_prepare(int n, const float* __restrict__ x, float* __restrict__ key, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  key[i] = v == 0.0f ? 0.0f : v;
  idx[i] = i + 1;
}

int c2ray_b200_mrgrnk(c2ray_ctx* c, int32_t n, const float* xvalt, int32_t* irngt) {
  if (!c || n < 0 || (n > 0 && (!xvalt || !irngt))) return fail(C2RAY_ERR_ARG, "bad argument");
  if (n == 0) return C2RAY_OK;
  CK(cudaSetDevice(c->device));
  float *d_x = nullptr, *d_k0 = nullptr, *d_k1 = nullptr;
  int *d_i0 = nullptr, *d_i1 = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  auto cleanup = [&]() { for (void* p : {(void*)d_x, (void*)d_k0, (void*)d_k1, (void*)d_i0, (void*)d_i1, d_tmp}) if (p) cudaFree(p); };
  cudaError_t e = cudaSuccess;
  auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
  ok(cudaMalloc(&d_x, 4 * (size_t)n)) && ok(cudaMalloc(&d_k0, 4 * (size_t)n)) && ok(cudaMalloc(&d_k1, 4 * (size_t)n)) &&
      ok(cudaMalloc(&d_i0, 4 * (size_t)n)) && ok(cudaMalloc(&d_i1, 4 * (size_t)n)) &&
      ok(cudaMemcpyAsync(d_x, xvalt, 4 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  if (e == cudaSuccess) {
    LAUNCH(c, k_mrgrnk_prepare, (n + 255) / 256, 256, n, d_x, d_k0, d_i0);
    ok(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_k0, d_k1, d_i0, d_i1, n, 0, 32, c->stream)) &&
        ok(cudaMalloc(&d_tmp, tmp_bytes)) &&
        ok(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_k0, d_k1, d_i0, d_i1, n, 0, 32, c->stream)) &&
        ok(cudaMemcpyAsync(irngt, d_i1, 4 * (size_t)n, cudaMemcpyDeviceToHost, c->stream)) && ok(cudaStreamSynchronize(c->stream));
    c->launches += 4;
  }
  cleanup();
  if (e != cudaSuccess) return fail(C2RAY_ERR_CUDA, std::string("mrgrnk: ") + cudaGetErrorString(e));
  return C2RAY_OK;
}

int c2ray_b200_rates_device_buffer(c2ray_ctx* c, void** dptr, int64_t* count) {
  if (!c || !dptr || !count) return fail(C2RAY_ERR_ARG, "null argument");
  *dptr = c->rates; *count = (int64_t)c->rates_count;
  return C2RAY_OK;
}

// ---- measurement ----------------------------------------------------------------------------------
int c2ray_b200_bench_global_pass(c2ray_ctx* c, double dt, int32_t reps, double* ms_per_pass, int32_t* conv_flag) {
  if (!c || reps < 1) return fail(C2RAY_ERR_ARG, "bad argument");
  CK(cudaSetDevice(c->device));
  int rc;
  float total = 0.f, ms;
  ChemTotals t;
  memset(&t, 0, sizeof(t));
  for (int i = 0; i < reps; i++) {
    if ((rc = begin_step(c))) return rc;  // same start state every repetition
    if (!c->par.isothermal && c->snap_temp) CK(cudaMemcpyAsync(c->temp, c->snap_temp, 3 * c->N3 * 4, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaEventRecord(c->ev[1], c->stream));
    if ((rc = global_pass_launch(c, dt, nullptr))) return rc;
    CK(cudaEventRecord(c->ev[2], c->stream));
    CK(cudaMemcpyAsync(&t, c->d_chem, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]));
    total += ms;
    c->last_nsub_per_cell = (double)t.nsub_total / (double)c->N3;
    c->last_nit_per_cell = (double)t.nit_total / (double)c->N3;
  }
  if (ms_per_pass) *ms_per_pass = total / reps;
  if (conv_flag) *conv_flag = t.conv_flag;
  return C2RAY_OK;
}

int64_t c2ray_b200_launch_count(c2ray_ctx* c) { return c ? c->launches : 0; }
int64_t c2ray_b200_sweep_launch_count(c2ray_ctx* c) { return c ? c->sweep_launches : 0; }

int c2ray_b200_measure_fp64(c2ray_ctx* c, double* tflops) {
  if (!c || !tflops) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  const int blocks = 148 * 8, threads = 256, iters = 20000;
  double* d;
  CK(cudaMalloc(&d, sizeof(double) * blocks * threads));
  LAUNCH(c, k_fp64_probe, blocks, threads, d, 100);
  float best = 1e30f, ms;
  for (int i = 0; i < 3; i++) {
    CK(cudaEventRecord(c->ev[1], c->stream));
    LAUNCH(c, k_fp64_probe, blocks, threads, d, iters);
    CK(cudaEventRecord(c->ev[2], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]));
    best = std::min(best, ms);
  }
  cudaFree(d);
  *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
  return C2RAY_OK;
}

int c2ray_b200_timer_start(c2ray_ctx* c) {
  if (!c) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaEventRecord(c->ev_timer[0], c->stream));
  return C2RAY_OK;
}
int c2ray_b200_timer_stop(c2ray_ctx* c, double* ms) {
  if (!c || !ms) return fail(C2RAY_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  CK(cudaEventRecord(c->ev_timer[1], c->stream));
  CK(cudaEventSynchronize(c->ev_timer[1]));
  float f;
  CK(cudaEventElapsedTime(&f, c->ev_timer[0], c->ev_timer[1]));
  *ms = f;
  return C2RAY_OK;
}

int c2ray_b200_stream(c2ray_ctx* c, void** stream) {
  if (!c || !stream) return fail(C2RAY_ERR_ARG, "null argument");
  *stream = (void*)c->stream;
  return C2RAY_OK;
}

}  // extern "C"
