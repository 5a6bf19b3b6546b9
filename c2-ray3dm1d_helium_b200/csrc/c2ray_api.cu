// c2ray_api.cu -- host side of libc2ray_b200.so: context, device residency, the evolve3D iteration
// (code/files_for_3D/evolve.F90:78-229), the source loop (master_slave.F90:74-96, evolve_source.F90:66-238) as
// batched shell-wavefront launches, the device rad_ini (radiation_tables.f90:141-168), the NCCL rate-grid
// reduction (evolve.F90:505-548) and the C ABI declared in include/c2ray_b200.h.
//
// No CPU fallback: every compute entry point needs a CUDA device and fails with C2RAY_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/c2ray_b200.h"
#include "band_data.h"
#include "c2ray_kernels.cuh"

using namespace c2;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(C2RAY_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                      std::to_string(__LINE__) + ")");                                   \
  } while (0)

// ---- NCCL through dlopen (no link-time dependency; torch's bundled libnccl is reused when already loaded) ----
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
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
  g_nccl.CommDestroy = (int (*)(nccl_comm))dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    return fail(C2RAY_ERR_NCCL, "libnccl: missing symbols");
  return 0;
}
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

}  // namespace

constexpr int MAX_SWEEP_GROUPS = 8;

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
  int* d_srcids = nullptr;  // 0-based ids of this rank's sources, in source order
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
  cudaStream_t gstream[MAX_SWEEP_GROUPS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_SWEEP_GROUPS] = {};
  int sweep_groups = 2;                     // env C2RAY_SWEEP_GROUPS
  double* d_scratch = nullptr;
  int slots_cap = 0;
  SweepGeom geom{};
  ChemTotals* d_chem = nullptr;
  double* d_sums = nullptr;
  double* d_secion = nullptr;
  unsigned long long* d_next_cell = nullptr;
  int chem_mode = -1;      // -1 auto, 0 one cell per thread, 1 queue-driven (env C2RAY_CHEM_QUEUE overrides)
  double last_nsub_per_cell = 0.0;  // thermal sub-steps per cell of the previous global pass
  int* d_nit = nullptr;
  // multi-GPU
  int rank = 0, npr = 1;
  nccl_comm comm = nullptr;
  // bookkeeping
  int64_t launches = 0;
  bool run_dirty = true;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
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
  static BandConst bc;
  memset(&bc, 0, sizeof(bc));
  bc.sigma_HI[0] = sigma_HI_at_ion_freq;  // radiation_sizes.f90:381-383
  for (int i = 0; i < 26; i++) {
    const int q = NumBndin1 + i;
    bc.sigma_HI[q] = BD_SIGMA_HI_B2[i]; bc.sigma_HeI[q] = BD_SIGMA_HEI_B2[i]; bc.sigma_HeII[q] = 0.0;
    bc.f1ion_HI[q] = BD_F1ION_HI_B2[i]; bc.f1ion_HeI[q] = BD_F1ION_HEI_B2[i]; bc.f1ion_HeII[q] = BD_F1ION_HEII_B2[i];
    bc.f2ion_HI[q] = BD_F2ION_HI_B2[i]; bc.f2ion_HeI[q] = BD_F2ION_HEI_B2[i]; bc.f2ion_HeII[q] = BD_F2ION_HEII_B2[i];
    bc.f1heat_HI[q] = BD_F1HEAT_HI_B2[i]; bc.f1heat_HeI[q] = BD_F1HEAT_HEI_B2[i]; bc.f1heat_HeII[q] = BD_F1HEAT_HEII_B2[i];
    bc.f2heat_HI[q] = BD_F2HEAT_HI_B2[i]; bc.f2heat_HeI[q] = BD_F2HEAT_HEI_B2[i]; bc.f2heat_HeII[q] = BD_F2HEAT_HEII_B2[i];
  }
  for (int i = 0; i < 20; i++) {
    const int q = NumBndin1 + NumBndin2 + i;
    bc.sigma_HI[q] = BD_SIGMA_HI_B3[i]; bc.sigma_HeI[q] = BD_SIGMA_HEI_B3[i]; bc.sigma_HeII[q] = BD_SIGMA_HEII_B3[i];
    bc.f1ion_HI[q] = BD_F1ION_HI_B3[i]; bc.f1ion_HeI[q] = BD_F1ION_HEI_B3[i]; bc.f1ion_HeII[q] = BD_F1ION_HEII_B3[i];
    bc.f2ion_HI[q] = BD_F2ION_HI_B3[i]; bc.f2ion_HeI[q] = BD_F2ION_HEI_B3[i]; bc.f2ion_HeII[q] = BD_F2ION_HEII_B3[i];
    bc.f1heat_HI[q] = BD_F1HEAT_HI_B3[i]; bc.f1heat_HeI[q] = BD_F1HEAT_HEI_B3[i]; bc.f1heat_HeII[q] = BD_F1HEAT_HEII_B3[i];
    bc.f2heat_HI[q] = BD_F2HEAT_HI_B3[i]; bc.f2heat_HeI[q] = BD_F2HEAT_HEI_B3[i]; bc.f2heat_HeII[q] = BD_F2HEAT_HEII_B3[i];
  }
  CK(cudaMemcpyToSymbol(d_band, &bc, sizeof(bc)));
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
  if (c->d_scratch && cap == c->geom.cap && want_slots <= c->slots_cap) return 0;
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
  return 0;
}

int rebuild_my_sources(c2ray_ctx* c) {
  if (c->d_srcids) { cudaFree(c->d_srcids); c->d_srcids = nullptr; }
  std::vector<int> ids;
  for (int ns1 = 1 + c->rank; ns1 <= c->NumSrc; ns1 += c->npr) ids.push_back(ns1 - 1);  // master_slave.F90:85
  c->n_mine = (int)ids.size();
  if (c->n_mine) {
    CK(cudaMalloc(&c->d_srcids, sizeof(int) * ids.size()));
    CK(cudaMemcpy(c->d_srcids, ids.data(), sizeof(int) * ids.size(), cudaMemcpyHostToDevice));
  }
  return 0;
}

// Sums the per-group totals into the context's aggregate and packs [photon_loss(1:47) | sum_nbox] behind the rate
// grids; only photon_loss(1) is ever filled (evolve_source.F90:233).
__global__ void k_pack_tail(const SweepTotals* gtot, int ngroups, SweepTotals* tot, double* tail) {
  const int t = threadIdx.x;
  if (t == 0) {
    SweepTotals a;
    a.photon_loss = 0.0; a.sum_nbox = 0; a.updates = 0; a.nactive = 0; a.pad = 0;
    for (int g = 0; g < ngroups; g++) { a.photon_loss += gtot[g].photon_loss; a.sum_nbox += gtot[g].sum_nbox; a.updates += gtot[g].updates; }
    *tot = a;
    tail[0] = a.photon_loss;
    tail[NumFreqBnd] = (double)a.sum_nbox;
  } else if (t < NumFreqBnd) tail[t] = 0.0;
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
    const int want = c->par.deterministic ? 1 : std::min(c->n_mine, c->par.max_slots > 0 ? c->par.max_slots : 1024);
    rc = alloc_sweep(c, want);
    if (rc) return rc;
    const SweepGeom g = c->geom;
    int rmax = 0;
    for (int d = 0; d < 3; d++) rmax = std::max(rmax, std::max(g.R[d], g.L[d]));
    if (!c->par.isothermal) {
      if (!c->d_secion) CK(cudaMalloc(&c->d_secion, 6 * c->N3 * sizeof(double)));
      LAUNCH(c, k_secion_factors, (unsigned)((c->N3 + 255) / 256), 256, c->xh_av, c->N3, c->d_secion);
    }
    GridPtrs G{c->ndens, c->xh_av, c->xhe_av, c->rates, c->d_secion, c->N3};
    const int batch = c->par.deterministic ? 1 : c->slots_cap;
    ngroups = c->par.deterministic ? 1 : std::max(1, std::min(c->sweep_groups, std::min(batch, c->n_mine)));
    const int max_blocks = 148 * 16;
    const bool multi_sed = c->tab[1][0] != nullptr || c->tab[2][0] != nullptr;  // PL / QPL tables present
    const size_t slot_stride = (size_t)6 * g.cap;
    CK(cudaMemsetAsync(c->d_gtot, 0, sizeof(SweepTotals) * MAX_SWEEP_GROUPS, c->stream));
    CK(cudaEventRecord(c->ev_fork, c->stream));
    for (int q = 0; q < ngroups; q++) CK(cudaStreamWaitEvent(c->gstream[q], c->ev_fork, 0));
    for (int first = 0; first < c->n_mine; first += batch) {
      const int ns = std::min(batch, c->n_mine - first);
      const int per = (ns + ngroups - 1) / ngroups;
      int goff[MAX_SWEEP_GROUPS], gns[MAX_SWEEP_GROUPS];
      for (int q = 0; q < ngroups; q++) { goff[q] = std::min(q * per, ns); gns[q] = std::min(per, ns - goff[q]); }
      for (int q = 0; q < ngroups; q++)
        if (gns[q] > 0)
          LAUNCH_S(c, c->gstream[q], k_slots_init, (gns[q] + 127) / 128, 128, c->d_slots + goff[q], gns[q],
                   c->d_srcids + first + goff[q], c->d_srcpos, c->d_nf, c->have_pl_flux ? c->d_nfpl : nullptr,
                   c->have_qpl_flux ? c->d_nfqpl : nullptr, c->d_gtot + q, c->d_active + goff[q]);
      const int reach3 = std::min(g.R[2], g.L[2]);
      for (int b = 1;; b++) {
        for (int q = 0; q < ngroups; q++)
          if (gns[q] > 0) LAUNCH_S(c, c->gstream[q], k_decide, 1, 256, c->d_slots + goff[q], gns[q], g, c->d_gtot + q, c->d_active + goff[q]);
        const int r_lo = b == 1 ? 0 : g.subboxsize * (b - 1) + 1;
        const int r_hi = (int)std::min<long long>((long long)g.subboxsize * b, rmax);
        for (int r = r_lo; r <= r_hi; r++) {
          for (int q = 0; q < ngroups; q++) {
            if (gns[q] <= 0) continue;
            const long long items = (long long)gns[q] * (r == 0 ? 1 : 24LL * r * r + 2);
            const int blocks = (int)std::min<long long>((items + 127) / 128, max_blocks);
#define SWEEP(ISO, MULTI)                                                                                              \
  LAUNCH_S(c, c->gstream[q], (k_sweep_shell<ISO, MULTI>), blocks, 128, c->d_slots + goff[q], c->d_active + goff[q], \
           c->d_gtot + q, g, G, c->d_scratch + (size_t)goff[q] * slot_stride, r)
            if (multi_sed) { if (c->par.isothermal) SWEEP(true, true); else SWEEP(false, true); }
            else { if (c->par.isothermal) SWEEP(true, false); else SWEEP(false, false); }
#undef SWEEP
          }
        }
        if ((long long)g.subboxsize * b >= reach3) break;  // the do-while's extent test fails for every source
      }
      for (int q = 0; q < ngroups; q++)  // close the sources still active
        if (gns[q] > 0) LAUNCH_S(c, c->gstream[q], k_decide, 1, 256, c->d_slots + goff[q], gns[q], g, c->d_gtot + q, c->d_active + goff[q]);
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

int global_pass_launch(c2ray_ctx* c, double dt, int* d_nit) {
  int rc = bind(c);
  if (rc) return rc;
  if (!c->par.isothermal && !c->have_cool) return fail(C2RAY_ERR_STATE, "cooling tables not set (c2ray_b200_set_cooling_tables)");
  CK(cudaMemsetAsync(c->d_chem, 0, sizeof(ChemTotals), c->stream));
  ChemPtrs P{c->ndens, c->xh, c->xhe, c->xh_av, c->xhe_av, c->xh_int, c->xhe_int, c->temp, c->rates, c->N3};
  // auto: the queue-driven kernel pays off once cells need many thermal sub-steps (divergence); measured break-even
  // between 10 and 40 sub-steps per cell (config 2: ~10, simple kernel faster; config 5: 42, queue 2x faster)
  const bool use_queue = c->chem_mode == 1 || (c->chem_mode < 0 && c->last_nsub_per_cell > 20.0);
  if (use_queue) {
    // queue-driven: a persistent grid of lanes drawing cells from a counter (k_global_pass_q)
    CK(cudaMemsetAsync(c->d_next_cell, 0, sizeof(unsigned long long), c->stream));
    static int per_sm = 0, n_sm = 0;
    if (!per_sm) {
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_global_pass_q, 128, 0));
      CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device));
      per_sm = std::max(per_sm, 1);
    }
    const unsigned blocks = (unsigned)std::min<size_t>((c->N3 + 127) / 128, (size_t)n_sm * per_sm);
    LAUNCH(c, k_global_pass_q, blocks, 128, P, dt, c->d_chem, d_nit, c->d_next_cell);
  } else {
    const unsigned blocks = (unsigned)((c->N3 + 127) / 128);
    LAUNCH(c, k_global_pass, blocks, 128, P, dt, c->d_chem, d_nit);
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
  const size_t N3 = c->N3;
  CK(cudaSetDevice(device));
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
  for (auto& st : c->gstream) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  for (auto& ev : c->ev_join) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  if (const char* e = getenv("C2RAY_SWEEP_GROUPS")) c->sweep_groups = std::max(1, std::min(MAX_SWEEP_GROUPS, atoi(e)));
  CK(cudaMalloc(&c->d_chem, sizeof(ChemTotals)));
  CK(cudaMalloc(&c->d_sums, 5 * sizeof(double)));
  CK(cudaMalloc(&c->d_next_cell, sizeof(unsigned long long)));
  if (const char* e = getenv("C2RAY_CHEM_QUEUE")) c->chem_mode = atoi(e);
  CK(cudaMalloc(&c->d_cool, 5 * TEMPPOINTS * sizeof(double)));
  CK(cudaMalloc(&c->d_tb, sizeof(TableBuild)));
  int rc = upload_band_const(c);
  if (rc) return rc;
  c->run_dirty = true;
  *out = c;
  return C2RAY_OK;
}

int c2ray_b200_destroy(c2ray_ctx* c) {
  if (!c) return C2RAY_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  void* ptrs[] = {c->ndens, c->xh, c->xhe, c->xh_av, c->xhe_av, c->xh_int, c->xhe_int, c->rates, c->temp, c->snap_xh,
                  c->snap_xhe, c->snap_temp, c->d_srcpos, c->d_nf, c->d_nfpl, c->d_nfqpl, c->d_srcids, c->d_tb, c->d_cool,
                  c->d_slots, c->d_active, c->d_tot, c->d_gtot, c->d_scratch, c->d_chem, c->d_sums, c->d_nit, c->d_secion, c->d_next_cell};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (int s = 0; s < 3; s++) for (int k = 0; k < 4; k++) if (c->tab[s][k]) cudaFree(c->tab[s][k]);
  for (int s = 0; s < 3; s++) if (c->packed[s]) cudaFree(c->packed[s]);
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
  const size_t n = (size_t)NumFreqBnd * PK_ROWS * PK_ROW;
  if (!c->packed[s]) CK(cudaMalloc(&c->packed[s], n * sizeof(double)));
  const int items = NumFreqBnd * PK_ROWS;
  LAUNCH(c, k_pack_tables, (items + 255) / 256, 256, c->tab[s][0], c->tab[s][1], c->tab[s][2], c->tab[s][3], c->packed[s]);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
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
  static TableBuild tb;
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
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  for (double** p : {&c->d_nf, &c->d_nfpl, &c->d_nfqpl}) if (*p) { cudaFree(*p); *p = nullptr; }
  if (c->d_srcpos) { cudaFree(c->d_srcpos); c->d_srcpos = nullptr; }
  c->NumSrc = NumSrc;
  c->have_pl_flux = nfpl != nullptr; c->have_qpl_flux = nfqpl != nullptr;
  c->sum_nf[0] = c->sum_nf[1] = c->sum_nf[2] = 0.0;
  for (int i = 0; i < NumSrc; i++) {
    c->sum_nf[0] += nf[i];
    if (nfpl) c->sum_nf[1] += nfpl[i];
    if (nfqpl) c->sum_nf[2] += nfqpl[i];
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
  if (rt_updates) *rt_updates = (int64_t)t.updates;
  return C2RAY_OK;
}

int c2ray_b200_do_source(c2ray_ctx* c, double /*dt*/, int32_t ns1, int32_t /*niter*/, int32_t* nbox, double* loss) {
  if (!c || ns1 < 1 || ns1 > c->NumSrc) return fail(C2RAY_ERR_ARG, "bad source number");
  // temporarily trace just this source
  int* saved = c->d_srcids; const int saved_n = c->n_mine;
  int id = ns1 - 1; int* d_id = nullptr;
  CK(cudaSetDevice(c->device));
  CK(cudaMalloc(&d_id, sizeof(int)));
  CK(cudaMemcpy(d_id, &id, sizeof(int), cudaMemcpyHostToDevice));
  c->d_srcids = d_id; c->n_mine = 1;
  int rc = sweep_all(c);
  c->d_srcids = saved; c->n_mine = saved_n;
  if (rc) { cudaFree(d_id); return rc; }
  SweepTotals t;
  CK(cudaMemcpyAsync(&t, c->d_tot, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_id);
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
  if (restart != 0) return fail(C2RAY_ERR_ARG, "restart from iteration dumps is not supported (evolve.F90:279)");
  if (c->NumSrc > 0 && !c->tab[0][0] && !c->tab[1][0] && !c->tab[2][0]) return fail(C2RAY_ERR_STATE, "no radiation tables");
  CK(cudaSetDevice(c->device));
  c2ray_stats S;
  memset(&S, 0, sizeof(S));
  int rc;
  if ((rc = state_sums(c, c->xh, c->xhe, S.sums_before))) return rc;  // evolve.F90:127
  if ((rc = begin_step(c))) return rc;
  int niter = 0;
  // conv_flag=mesh(1)*mesh(2)*mesh(3) ; conv_criterion=min(int(convergence_fraction*mesh1*mesh2*mesh3),NumSrc)  :136,:147
  int conv_flag = c->mesh[0] * c->mesh[1] * c->mesh[2];
  const int conv_criterion = std::min((int)(convergence_fraction * c->mesh[0] * c->mesh[1] * c->mesh[2]), c->NumSrc);
  float ms;
  SweepTotals swt;
  memset(&swt, 0, sizeof(swt));
  ChemTotals cht;
  memset(&cht, 0, sizeof(cht));
  for (;;) {
    if (conv_flag < conv_criterion && niter > 1) {  // :163
      if ((rc = end_step(c))) return rc;
      break;
    } else if (niter > 500) break;  // :177
    niter++;
    CK(cudaEventRecord(c->ev[1], c->stream));
    CK(cudaMemsetAsync(c->rates, 0, c->rates_count * 8, c->stream));  // :188
    if (c->NumSrc > 0) {
      if ((rc = sweep_all(c))) return rc;  // :192
      CK(cudaMemcpyAsync(&swt, c->d_tot, sizeof(swt), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaEventRecord(c->ev[2], c->stream));
    if (c->NumSrc > 0 && (rc = allreduce_rates(c))) return rc;
    CK(cudaEventRecord(c->ev[3], c->stream));
    if ((rc = global_pass_launch(c, dt, nullptr))) return rc;  // :217
    CK(cudaMemcpyAsync(&cht, c->d_chem, sizeof(cht), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev[0], c->stream));  // iteration end
    CK(cudaStreamSynchronize(c->stream));
    conv_flag = cht.conv_flag;
    c->last_nsub_per_cell = (double)cht.nsub_total / (double)c->N3;
    if (niter <= C2RAY_MAX_ITER_HIST) S.conv_hist[niter - 1] = conv_flag;
    S.rt_updates += (int64_t)swt.updates;
    S.chem_cells += (int64_t)c->N3;
    CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); S.ms_sweep += ms;
    CK(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3])); S.ms_allreduce += ms;
    CK(cudaEventElapsedTime(&ms, c->ev[3], c->ev[0])); S.ms_chem += ms;
  }
  if ((rc = state_sums(c, c->xh, c->xhe, S.sums_after))) return rc;  // :225 (state_after on the final xh)
  {
    // total_rates(dt, xh_av, xhe_av) with the coefficients the reference's module globals hold at this point
    const double coef_T = c->par.isothermal ? c->par.temper_val : cht.last_coef_T;
    CK(cudaMemsetAsync(c->d_sums, 0, 3 * sizeof(double), c->stream));
    if (coef_T > 0.0) LAUNCH(c, k_total_rates, 148 * 4, 256, c->ndens, c->xh_av, c->xhe_av, c->N3, coef_T, c->d_sums);
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
  }
  if (ms_per_pass) *ms_per_pass = total / reps;
  if (conv_flag) *conv_flag = t.conv_flag;
  return C2RAY_OK;
}

int64_t c2ray_b200_launch_count(c2ray_ctx* c) { return c ? c->launches : 0; }

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
