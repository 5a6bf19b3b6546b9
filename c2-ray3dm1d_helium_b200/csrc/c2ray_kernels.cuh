// c2ray_kernels.cuh -- CUDA kernels of the hot path (sm_100a).
//   (a)+(b) k_sweep_shell : evolve_source.F90 do_source / evolve_point.F90 evolve0D / column_density.f90 cinterp /
//                           radiation_photoionrates.f90 photoion_rates as a max-norm shell wavefront over many
//                           sources at once: one launch per shell radius, every (active source, cell)
//                           pair of that shell is one work item.
//   (c)     k_global_pass : evolve_point.F90 evolve0D_global / do_chemistry + doric + thermal, one cell per thread,
//                           HBM-streaming SoA loads/stores, warp-shuffle reductions for the convergence vote.
//   tables  k_build_tables: radiation_tables.f90 spec_integration on the device.
#pragma once
#include <cuda_runtime.h>
#include "c2ray_photo.cuh"

namespace c2 {

// ------------------------------------------------------------------------------------------------
// Sweep bookkeeping
// ------------------------------------------------------------------------------------------------
// Two 128-byte lines per slot.  Every thread of a shell launch reads the first (and again after its band loop); the
// second takes the photon-loss atomics of the boundary cells.  In one line, each atomic threw the line out of the issuing
// SM's L1 and the next reader queued behind the atomics at the L2 -- the same for SweepTotals::nactive next to the update
// counter, where the first instruction after the load held 7.6 % of the kernel's warp time
// (profiles/r2c_ncu_sweep_cfg1_stalls.txt).
struct alignas(128) Slot {  // one source being traced (evolve_source.F90:66-238 local state)
  double nflux[3];     // NormFlux, NormFluxPL, NormFluxQPL of the source
  double total_flux;   // :122-128
  int src;             // 0-based source number
  int s[3];            // srcpos (1-based mesh position)
  int lo[3], hi[3];    // current sub-box reach: last_l = srcpos - lo, last_r = srcpos + hi (:143-144)
  int nbox;
  int active;
  alignas(128) double loss;  // photon_loss_src
};
static_assert(sizeof(Slot) == 256, "Slot: one read-mostly line, one line for the atomics");

struct SweepTotals {
  double photon_loss;               // photon_loss(1), evolve_source.F90:233
  unsigned long long sum_nbox;      // :236
  unsigned long long updates;       // evolve0D calls that did work
  alignas(128) int nactive;         // written by k_slots_init / k_decide only
};
static_assert(sizeof(SweepTotals) == 256, "SweepTotals: counters and nactive in separate lines");

struct SweepGeom {
  int L[3], R[3];   // lastpos_l / lastpos_r reach (:103-105)
  int subboxsize;
  int cap;          // entries per species per shell buffer
};

__device__ __forceinline__ int shell_cells(int r) { return r == 0 ? 1 : 24 * r * r + 2; }

// index of offset (di,dj,dk), max-norm r, in the face-ordered shell layout (also the thread order)
__device__ __forceinline__ int shell_index(int di, int dj, int dk, int r) {
  if (r == 0) return 0;
  const int n1 = 2 * r + 1, n0 = 2 * r - 1;
  if (dk == -r) return (dj + r) * n1 + (di + r);
  if (dk == r) return n1 * n1 + (dj + r) * n1 + (di + r);
  int base = 2 * n1 * n1;
  if (dj == -r) return base + (dk + r - 1) * n1 + (di + r);
  base += n0 * n1;
  if (dj == r) return base + (dk + r - 1) * n1 + (di + r);
  base += n0 * n1;
  if (di == -r) return base + (dk + r - 1) * n0 + (dj + r - 1);
  return base + n0 * n0 + (dk + r - 1) * n0 + (dj + r - 1);
}
__device__ __forceinline__ void shell_decode(int c, int r, int& di, int& dj, int& dk) {
  if (r == 0) { di = dj = dk = 0; return; }
  const int n1 = 2 * r + 1, n0 = 2 * r - 1;
  const int fz = n1 * n1, fy = n0 * n1, fx = n0 * n0;
  if (c < 2 * fz) {
    const int f = c >= fz; c -= f * fz;
    dk = f ? r : -r; dj = c / n1 - r; di = c % n1 - r;
  } else if ((c -= 2 * fz) < 2 * fy) {
    const int f = c >= fy; c -= f * fy;
    dj = f ? r : -r; dk = c / n1 - r + 1; di = c % n1 - r;
  } else {
    c -= 2 * fy;
    const int f = c >= fx; c -= f * fx;
    di = f ? r : -r; dk = c / n0 - r + 1; dj = c % n0 - r + 1;
  }
}

__device__ __forceinline__ int wrap0(int p1, int n) {  // 1-based unwrapped -> 0-based periodic (modulo(p-1,n))
  // traced offsets reach at most n/2 cells from a mesh position, so one conditional wrap is the whole modulo
  int m = p1 - 1;
  m += (m < 0) ? n : 0;
  m -= (m >= n) ? n : 0;
  return m;
}

__global__ void k_slots_init(Slot* slots, int nslots, const int* __restrict__ src_ids, const int* __restrict__ srcpos,
                             const double* __restrict__ nf, const double* __restrict__ nfpl,
                             const double* __restrict__ nfqpl, SweepTotals* tot, int* active_list) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) tot->nactive = nslots;
  if (t >= nslots) return;
  Slot s;
  const int ns = src_ids[t];
  s.src = ns;
  for (int d = 0; d < 3; d++) { s.s[d] = srcpos[3 * ns + d]; s.lo[d] = 0; s.hi[d] = 0; }
  s.nflux[0] = nf[ns];
  s.nflux[1] = (d_run.sed[1].hi >= d_run.sed[1].lo && nfpl) ? nfpl[ns] : 0.0;
  s.nflux[2] = (d_run.sed[2].hi >= d_run.sed[2].lo && nfqpl) ? nfqpl[ns] : 0.0;
  double tf = s.nflux[0] * d_run.sed[0].S_star;                       // evolve_source.F90:122
  if (d_run.sed[1].hi >= d_run.sed[1].lo) tf = tf + s.nflux[1] * d_run.sed[1].S_star;  // :124
  if (d_run.sed[2].hi >= d_run.sed[2].lo) tf = tf + s.nflux[2] * d_run.sed[2].S_star;  // :127
  s.total_flux = tf;
  s.loss = tf;  // :129
  s.nbox = 0;
  s.active = 1;
  slots[t] = s;
  active_list[t] = t;
}

// The `do while` test of evolve_source.F90:136-144 for every slot, then compaction of the active slots.
// Single block.
__global__ void k_decide(Slot* slots, int nslots, SweepGeom g, SweepTotals* tot, int* active_list, int* nbox_all) {
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < nslots; t += blockDim.x) {
    Slot& s = slots[t];
    if (!s.active) continue;
    const bool go = s.loss > loss_fraction * s.total_flux && s.hi[2] < g.R[2] && s.lo[2] < g.L[2];
    if (go) {
      s.nbox = s.nbox + 1;
      s.loss = 0.0;
      for (int d = 0; d < 3; d++) {
        s.hi[d] = min(g.subboxsize * s.nbox, g.R[d]);
        s.lo[d] = min(g.subboxsize * s.nbox, g.L[d]);
      }
      active_list[atomicAdd(&cnt, 1)] = t;
    } else {
      s.active = 0;
      if (nbox_all) nbox_all[s.src] = s.nbox;                           // per-source cost record for the balanced schedule
      atomicAdd(&tot->photon_loss, s.loss);                             // :233
      atomicAdd(&tot->sum_nbox, (unsigned long long)s.nbox);            // :236
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) tot->nactive = cnt;
}

// What the sweep reads of a cell, gathered once per iteration into one record per cell (k_cell_records):
//   [ xh_av(0) ndens , xhe_av(0) ndens , xhe_av(1) ndens , - , y1R(1:3) , y2R(1:3) ]      (fractions floored at epsilon)
// 80 bytes (isothermal: the first 32).  The x-faces of a shell are strided in memory (i is the fast axis), where ten
// separate planes cost ten 64-byte DRAM granules per cell; one record costs two.
constexpr int CELLREC = 10, CELLREC_ISO = 4;

struct GridPtrs {
  const double* cellrec;  // (N3, CELLREC or CELLREC_ISO)
  double* rates;          // phih | phihe0 | phihe1 | phiheat, N3 each
  size_t N3;
  // Lyman-limit systems (evolve_point.F90:170-180): type_of_LLS 0 none, 1 one column density per cell for the whole
  // mesh, 2 LLS_grid(i,j,k) (material's real array)
  int lls_type;
  double coldensh_LLS;
  const float* lls_grid;
};

__device__ __forceinline__ void write_cell_record(const double* __restrict__ ndens, const double* __restrict__ xh_av,
                                                  const double* __restrict__ xhe_av, size_t N3, bool iso, size_t p,
                                                  double* __restrict__ out) {
  const double n = ndens[p];
  double2* o = reinterpret_cast<double2*>(out + p * (iso ? CELLREC_ISO : CELLREC));
  o[0] = make_double2(fmax(xh_av[p], epsilon) * n, fmax(xhe_av[p], epsilon) * n);
  o[1] = make_double2(fmax(xhe_av[p + N3], epsilon) * n, 0.0);
  if (!iso) {
    const SecIon y = secion_factors_fast(fmax(xh_av[p + N3], epsilon));
    o[2] = make_double2(y.y1R0, y.y1R1); o[3] = make_double2(y.y1R2, y.y2R0); o[4] = make_double2(y.y2R1, y.y2R2);
  }
}

// The cell records.  The secondary-ionisation factors (radiation_photoionrates.f90:557-565) depend on xh_av(1) only, so
// they are evaluated once per iteration here, not once per source x cell.
__global__ void k_cell_records(const double* __restrict__ ndens, const double* __restrict__ xh_av,
                               const double* __restrict__ xhe_av, size_t N3, int iso, double* __restrict__ out) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N3) return;
  write_cell_record(ndens, xh_av, xhe_av, N3, iso != 0, p, out);  // evolve_point.F90:117-124
}

// The records of the cells one sub-box level is about to trace, and no others: shells r_lo..r_hi around every active
// source (the same enumeration as the sweep's).  Used instead of k_cell_records when the sources cover a small part of
// the mesh (1250 sources x 21^3 cells on a 512^3 mesh: 12 M records instead of 134 M).  Cells reached by several sources
// are written several times with identical bits.
__global__ void k_cell_records_level(const Slot* __restrict__ slots, const int* __restrict__ active_list,
                                     const SweepTotals* __restrict__ tot, int r_lo, int r_hi, const double* __restrict__ ndens,
                                     const double* __restrict__ xh_av, const double* __restrict__ xhe_av, size_t N3, int iso,
                                     double* __restrict__ out) {
  const int nact = tot->nactive;
  const int m0 = d_run.mesh[0], m1 = d_run.mesh[1], m2 = d_run.mesh[2];
  // cells of the shells r_lo..r_hi: (2 r_hi + 1)^3 - (2 r_lo - 1)^3 (r_lo = 0: the whole cube)
  const long long per = (long long)(2 * r_hi + 1) * (2 * r_hi + 1) * (2 * r_hi + 1) -
                        (r_lo > 0 ? (long long)(2 * r_lo - 1) * (2 * r_lo - 1) * (2 * r_lo - 1) : 0);
  const long long total = per * nact;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(t / per);
    long long c = t - (long long)a * per;
    int r = r_lo;
    for (;; r++) {  // which shell (at most subboxsize steps)
      const long long n = shell_cells(r);
      if (c < n) break;
      c -= n;
    }
    const Slot& S = slots[active_list[a]];
    int di, dj, dk;
    shell_decode((int)c, r, di, dj, dk);
    if (di < -S.lo[0] || di > S.hi[0] || dj < -S.lo[1] || dj > S.hi[1] || dk < -S.lo[2] || dk > S.hi[2]) continue;
    const size_t p = (size_t)wrap0(S.s[0] + di, m0) + (size_t)m0 * ((size_t)wrap0(S.s[1] + dj, m1) + (size_t)m1 * wrap0(S.s[2] + dk, m2));
    write_cell_record(ndens, xh_av, xhe_av, N3, iso != 0, p, out);
  }
}

// One shell radius r of every active source.  Work item = (active slot, cell of the shell).
// Resident CTAs per SM: with the secondary-ionisation factors out of the band loop the kernels fit 96 registers with
// a spill of a few words and gain 3-4 % from the fifth CTA (A/B on one box: 17.86 -> 17.31 ms per pass at 16 sources,
// 615 -> 592 ms at 1000 sources with the multi-SED kernel); a sixth (80 registers) loses 1 %.
#ifndef C2RAY_SWEEP_MINBLOCKS
#define C2RAY_SWEEP_MINBLOCKS 5
#endif
#ifndef C2RAY_SWEEP_MINBLOCKS_MULTI
#define C2RAY_SWEEP_MINBLOCKS_MULTI 5
#endif
// The LANES > 1 instances only ever run in launches below one resident wave, where occupancy buys nothing and the
// latency of the one chain is everything: they get 128 registers (4 CTAs/SM, no spills).  One source at 128^3:
// 51.2 -> 45.2 ms of sweeps per time step with every instance at 4 CTAs/SM, while the full-wave launches of configs[1]
// and [2] lose 6 % there (profiles/r2_ab10_unroll.log) -- hence per instance.
#ifndef C2RAY_SWEEP_MINBLOCKS_SPLIT
#define C2RAY_SWEEP_MINBLOCKS_SPLIT 4
#endif
// threads per CTA of the LANES == 1 instances (the LANES > 1 instances keep 128)
#ifndef C2RAY_SWEEP_THREADS
#define C2RAY_SWEEP_THREADS 128
#endif
// LANES (1 or a power of two <= 32): lanes of a warp that share one cell, each taking every LANES-th frequency band.
// One update is a dependent chain of ~9 k instructions, ~25 us for a warp on its own; a launch that cannot fill the
// machine (the inner shells, few sources) is bound by that latency, not by throughput, and finishes LANES times sooner
// when the chain is cut into LANES pieces.  The geometry part is computed redundantly by the sharing lanes; lane 0 of a
// cell writes.  Launches with more cells than resident threads use LANES = 1.
template <bool ISO, bool MULTI, int LANES>
__global__ void __launch_bounds__(LANES > 1 ? 128 : C2RAY_SWEEP_THREADS,
                                  LANES > 1 ? C2RAY_SWEEP_MINBLOCKS_SPLIT : (MULTI ? C2RAY_SWEEP_MINBLOCKS_MULTI : C2RAY_SWEEP_MINBLOCKS))
k_sweep_shell(Slot* slots, const int* __restrict__ active_list, SweepTotals* tot, SweepGeom g,
              GridPtrs G, double* __restrict__ scratch, int r, double* __restrict__ lossbuf) {
  // lossbuf != nullptr (deterministic mode, one source at a time): every cell of the shell writes its photon-loss
  // contribution (0 for cells off the sub-box boundary or outside the box) and k_loss_sum adds them in a fixed order,
  // instead of atomic adds whose order varies from run to run
  // Programmatic dependent launch: within a sub-box level the next shell's launch is allowed to start filling SMs
  // while this one drains (its threads decode their cell, test the box and fetch the cell record, then wait below
  // before they touch the shell scratch).  Without the launch attribute both instructions are no-ops.
  asm volatile("griddepcontrol.launch_dependents;");
  const int nact = tot->nactive;  // (requested before the barrier below: its latency passes under the staging)
#if C2RAY_TABLOG
  postab_stage();  // 4 KB, L2-resident; before griddepcontrol.wait, so it overlaps the previous shell's tail
#endif
  const int ncell = shell_cells(r);
  const long long total = (long long)nact * ncell * LANES;
  const int lane_j = LANES > 1 ? (int)(threadIdx.x & (LANES - 1)) : 0;
  const unsigned lane_mask = LANES > 1 ? ((LANES == 32 ? 0xffffffffu : ((1u << LANES) - 1u)) << (threadIdx.x & 31 & ~(LANES - 1))) : 0u;
  const size_t slot_stride = (size_t)6 * g.cap;                 // [parity][species][cap]
  const int par = r & 1;
  constexpr bool iso = ISO;
  const int m0 = d_run.mesh[0], m1 = d_run.mesh[1], m2 = d_run.mesh[2];
  unsigned int done = 0;
  {
    // One work item per thread.  No thread leaves early: threads beyond the work, and threads whose cell lies outside
    // its source's sub-box, run along on harmless values (item 0 / unit columns) with `inbox` false, so that the band
    // loop below sits in convergent code.
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = t0 < total;
    const long long t = in_range ? t0 : 0;
    const long long item = LANES > 1 ? t / LANES : t;   // the LANES lanes of an aligned group share the item
    const int a = (int)(item / ncell);
    const int c = (int)(item - (long long)a * ncell);
    const int sid = active_list[a];
    Slot& S = slots[sid];
    int di, dj, dk;
    shell_decode(c, r, di, dj, dk);
    const bool inbox = in_range && !(di < -S.lo[0] || di > S.hi[0] || dj < -S.lo[1] || dj > S.hi[1] || dk < -S.lo[2] || dk > S.hi[2]);
    if (in_range && !inbox && lossbuf && lane_j == 0) lossbuf[c] = 0.0;
    double* cur = scratch + sid * slot_stride + (size_t)par * 3 * g.cap;
    const double* prev = scratch + sid * slot_stride + (size_t)(par ^ 1) * 3 * g.cap;
    const int i0 = S.s[0], j0 = S.s[1], k0 = S.s[2];
    const size_t p = (size_t)wrap0(i0 + di, m0) + (size_t)m0 * ((size_t)wrap0(j0 + dj, m1) + (size_t)m1 * wrap0(k0 + dk, m2));
    double2 rec0 = make_double2(1.0, 1.0), rec1 = make_double2(1.0, 0.0);
    if (inbox) {
      rec0 = ld2(G.cellrec + p * (ISO ? CELLREC_ISO : CELLREC));      // xh_av(0) n, xhe_av(0) n
      rec1 = ld2(G.cellrec + p * (ISO ? CELLREC_ISO : CELLREC) + 2);  // xhe_av(1) n, -
    }
    // Everything above reads only what is constant within a sub-box level (slots, active list, cell records).  From
    // here on the previous shell's column densities are read and the buffer it is still reading is overwritten.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    double cin_H = 1.0, cin_He0 = 1.0, cin_He1 = 1.0, path = 1.0, vol_ph = 1.0;
    if (!inbox) {
      // (nothing to interpolate)
    } else if (r == 0) {  // evolve_point.F90:140-150
      cin_H = 0.0; cin_He0 = 0.0; cin_He1 = 0.0;
      path = FL(0.5f) * d_run.dr[0];
      vol_ph = d_run.dr[0] * d_run.dr[1] * d_run.dr[2];
    } else {
      // ---- cinterp, column_density.f90:28-345.  The geometric part must not be contracted into FMAs: the
      // weights of not-yet-computed same-shell corners are exactly 0 only in plain IEEE mul/add.
      const int ia = abs(di), ja = abs(dj), ka = abs(dk);
      const int sgi = di >= 0 ? 1 : -1, sgj = dj >= 0 ? 1 : -1, sgk = dk >= 0 ? 1 : -1;
      const double ddi = (double)di, ddj = (double)dj, ddk = (double)dk;
      // dominant axis: z if ka>=ja&&ka>=ia ; else y if ja>=ia&&ja>=ka ; else x
      int ax;  // 2=z 1=y 0=x
      if (ka >= ja && ka >= ia) ax = 2; else if (ja >= ia && ja >= ka) ax = 1; else ax = 0;
      // generic (u,v) transverse coordinates; w the dominant one
      int du, dv, dw, sgu, sgv, sgw, u0, v0;
      double fu, fv, fw;
      if (ax == 2) { du = di; dv = dj; dw = dk; sgu = sgi; sgv = sgj; sgw = sgk; u0 = i0; v0 = j0; fu = ddi; fv = ddj; fw = ddk; }
      else if (ax == 1) { du = di; dv = dk; dw = dj; sgu = sgi; sgv = sgk; sgw = sgj; u0 = i0; v0 = k0; fu = ddi; fv = ddk; fw = ddj; }
      else { du = dj; dv = dk; dw = di; sgu = sgj; sgv = sgk; sgw = sgi; u0 = j0; v0 = k0; fu = ddj; fv = ddk; fw = ddi; }
      // alam=(real(wm-w0)+sgw*0.5)/dw with wm-w0 = dw-sgw  (values exact in binary32)
      const double alam = __ddiv_rn((double)((float)(dw - sgw) + (float)sgw * 0.5f), fw);
      const double uc = __dadd_rn(__dmul_rn(alam, fu), (double)u0);
      const double vc = __dadd_rn(__dmul_rn(alam, fv), (double)v0);
      const double um = (double)((float)(u0 + du - sgu) + 0.5f * (float)sgu);
      const double vm = (double)((float)(v0 + dv - sgv) + 0.5f * (float)sgv);
      const double du_ = __dmul_rn(2.0, fabs(__dsub_rn(uc, um)));
      const double dv_ = __dmul_rn(2.0, fabs(__dsub_rn(vc, vm)));
      // weights: z-plane: s1=(1-dx)(1-dy) s2=(1-dy)dx s3=(1-dx)dy s4=dx dy  with (u,v)=(x,y)
      //          y-plane: s1=(1-dx)(1-dz) s2=(1-dz)dx s3=(1-dx)dz s4=dx dz  with (u,v)=(x,z)
      //          x-plane: s1=(1-dz)(1-dy) s2=(1-dz)dy s3=(1-dy)dz s4=dy dz  with (u,v)=(y,z): c2<->u-shifted
      const double omu = __dsub_rn(1.0, du_), omv = __dsub_rn(1.0, dv_);
      double s1, s2, s3, s4;
      if (ax == 0) { s1 = __dmul_rn(omv, omu); s2 = __dmul_rn(omv, du_); s3 = __dmul_rn(omu, dv_); s4 = __dmul_rn(du_, dv_); }
      else { s1 = __dmul_rn(omu, omv); s2 = __dmul_rn(omv, du_); s3 = __dmul_rn(omu, dv_); s4 = __dmul_rn(du_, dv_); }
      // corner offsets: c1=(um,vm) c2=(u,vm) c3=(um,v) c4=(u,v), all at w-sgw
      const int rm = r - 1;
      int o1[3], o2[3], o3[3], o4[3];
      {
        const int uu = du, um_i = du - sgu, vv = dv, vm_i = dv - sgv, wm_i = dw - sgw;
        int A[4][3];
        const int cu[4] = {um_i, uu, um_i, uu}, cv[4] = {vm_i, vm_i, vv, vv};
        for (int q = 0; q < 4; q++) {
          if (ax == 2) { A[q][0] = cu[q]; A[q][1] = cv[q]; A[q][2] = wm_i; }
          else if (ax == 1) { A[q][0] = cu[q]; A[q][2] = cv[q]; A[q][1] = wm_i; }
          else { A[q][1] = cu[q]; A[q][2] = cv[q]; A[q][0] = wm_i; }
        }
        for (int d = 0; d < 3; d++) { o1[d] = A[0][d]; o2[d] = A[1][d]; o3[d] = A[2][d]; o4[d] = A[3][d]; }
      }
      double cH[4], cHe0[4], cHe1[4];
      const int* oo[4] = {o1, o2, o3, o4};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int* o = oo[q];
        const int nr = max(abs(o[0]), max(abs(o[1]), abs(o[2])));
        if (nr == rm) {
          const int id = shell_index(o[0], o[1], o[2], rm);
          cH[q] = prev[id]; cHe0[q] = prev[g.cap + id]; cHe1[q] = prev[2 * g.cap + id];
        } else {  // same-shell corner: its weight is exactly zero, the reference reads 0 or a finite value
          cH[q] = 0.0; cHe0[q] = 0.0; cHe1[q] = 0.0;
        }
      }
      const double sw[4] = {s1, s2, s3, s4};
      double num, den;
      num = 0.0; den = 0.0;
#pragma unroll
      for (int q = 0; q < 4; q++) { const double w = sw[q] * weightf_fast(cH[q], sigma_HI_at_ion_freq); num += cH[q] * w; den += w; }
      cin_H = fdiv(num, den);
      num = 0.0; den = 0.0;
#pragma unroll
      for (int q = 0; q < 4; q++) { const double w = sw[q] * weightf_fast(cHe0[q], sigma_HeI_at_ion_freq); num += cHe0[q] * w; den += w; }
      cin_He0 = fdiv(num, den);
      num = 0.0; den = 0.0;
#pragma unroll
      for (int q = 0; q < 4; q++) { const double w = sw[q] * weightf_fast(cHe1[q], sigma_HeII_at_ion_freq); num += cHe1[q] * w; den += w; }
      cin_He1 = fdiv(num, den);
      const int wa = abs(dw), ua = abs(du), va = abs(dv);
      if (wa == 1 && (ua == 1 || va == 1)) {  // :174-184
        const double f = (ua == 1 && va == 1) ? sqrt3 : sqrt2;
        cin_H = f * cin_H; cin_He0 = f * cin_He0; cin_He1 = f * cin_He1;
      }
      path = sqrt(fdiv(fu * fu + fv * fv, fw * fw) + 1.0);  // :194
      path = path * d_run.dr[0];                            // evolve_point.F90:158
      const double xs = d_run.dr[0] * ddi, ys = d_run.dr[1] * ddj, zs = d_run.dr[2] * ddk;
      const double dist2 = xs * xs + ys * ys + zs * zs;
      vol_ph = FL(4.0f) * pi * dist2 * path;                // :168
      if (G.lls_type) {                                      // :177-180 coldensh_in + coldensh_LLS * path/dr(1)
        const double cl = G.lls_type == 2 ? (double)G.lls_grid[p] : G.coldensh_LLS;
        cin_H = cin_H + __ddiv_rn(cl * path, d_run.dr[0]);
      }
    }
    // evolve_point.F90:237-244
    const double cout_H = cin_H + rec0.x * path * (1.0 - abu_he);
    const double cout_He0 = cin_He0 + rec0.y * path * abu_he;
    const double cout_He1 = cin_He1 + rec1.x * path * abu_he;
    if (inbox && lane_j == 0) { cur[c] = cout_H; cur[g.cap + c] = cout_He0; cur[2 * g.cap + c] = cout_He1; }

    PhotOut phi = {0, 0, 0, 0, 0, 0};
    const bool do_bands = inbox && cin_H < max_coldensh;  // :250-270
    // The band loop runs in convergent code -- every lane of the warp enters it when any lane has a cell to do, lanes
    // without one on harmless columns -- so that the compiler may keep the band index, and with it the band constants,
    // in the uniform datapath.
    if (__any_sync(0xffffffffu, do_bands)) {
      double scale;
      PhotAcc A = photoion_bands<ISO, MULTI, LANES, false>(cin_H, cout_H, cin_He0, cout_He0, cin_He1, cout_He1, S.nflux, scale, lane_j);
      if (LANES > 1) reduce_bands<ISO, LANES>(A, lane_mask);  // (do_bands is uniform over the lanes of a cell)
      if (do_bands) {
        // the cell's secondary-ionisation factors are only needed now: loading them after the band loop keeps twelve
        // registers free while it runs (the record's address is formed again from the cell index behind an
        // optimisation barrier, so that only the index stays live across the loop)
        size_t p2 = p;
        asm volatile("" : "+l"(p2));
        const double* rec = G.cellrec + p2 * (ISO ? CELLREC_ISO : CELLREC);
        SecIon yR = {0, 0, 0, 0, 0, 0};
        if (!iso) {
          const double2 y0 = ld2(rec + 4), y1 = ld2(rec + 6), y2 = ld2(rec + 8);
          yR.y1R0 = y0.x; yR.y1R1 = y0.y; yR.y1R2 = y1.x; yR.y2R0 = y1.y; yR.y2R1 = y2.x; yR.y2R2 = y2.y;
        }
        phi = photoion_finish<ISO>(A, scale, vol_ph, yR);
        // the cell's densities again (cache hits) rather than three values held in registers across the band loop
        const double2 d0 = ld2(rec), d1 = ld2(rec + 2);
        phi.photo_HI = fdiv(phi.photo_HI, d0.x * (1.0 - abu_he));
        phi.photo_HeI = fdiv(phi.photo_HeI, d0.y * abu_he);
        phi.photo_HeII = fdiv(phi.photo_HeII, d1.x * abu_he);
      }
    }
    if (inbox && !(LANES > 1 && lane_j != 0)) {            // one lane per cell publishes
      atomicAdd(G.rates + p, phi.photo_HI);                // :299-306
      atomicAdd(G.rates + G.N3 + p, phi.photo_HeI);
      atomicAdd(G.rates + 2 * G.N3 + p, phi.photo_HeII);
      if (!iso) atomicAdd(G.rates + 3 * G.N3 + p, phi.heat);
      // :310-314 photon loss over the current sub-box boundary.  On the outermost shells every cell is a loss cell of
      // one of a handful of sources: a full warp working on one source sums its contributions by shuffles first, so the
      // per-source counter sees one atomic per warp instead of 32.
      const bool is_loss = di == -S.lo[0] || dj == -S.lo[1] || dk == -S.lo[2] || di == S.hi[0] || dj == S.hi[1] || dk == S.hi[2];
      if (lossbuf) {
        lossbuf[c] = is_loss ? fdiv(phi.photo_out * d_run.vol, vol_ph) : 0.0;
      } else {
        const unsigned am = __activemask();
        if (__any_sync(am, is_loss)) {
          double lv = is_loss ? fdiv(phi.photo_out * d_run.vol, vol_ph) : 0.0;
          int same = 0;
          if (am == 0xffffffffu) __match_all_sync(am, sid, &same);
          if (same) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lv += __shfl_xor_sync(0xffffffffu, lv, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(&S.loss, lv);
          } else if (is_loss) {
            atomicAdd(&S.loss, lv);
          }
        }
      }
      done++;
    }
  }
  __syncwarp();
  done = __reduce_add_sync(0xffffffffu, done);
  if ((threadIdx.x & 31) == 0 && done) atomicAdd(&tot->updates, (unsigned long long)done);
}

// Deterministic mode: photon_loss_src += sum of the shell's per-cell contributions, in an order that depends on
// nothing but the shell size (thread t adds cells t, t+1024, ... in sequence; then a fixed tree over the 1024 partial sums).
__global__ void __launch_bounds__(1024) k_loss_sum(Slot* slot, const double* __restrict__ lossbuf, int ncell) {
  __shared__ double sh[1024];
  double v = 0.0;
  for (int c = threadIdx.x; c < ncell; c += 1024) v += lossbuf[c];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) slot->loss += sh[0];
}

// ------------------------------------------------------------------------------------------------
// (c) global pass: evolve_point.F90:325-440 evolve0D_global
// ------------------------------------------------------------------------------------------------
struct ChemPtrs {
  const double* ndens;
  const double* xh;      // (N3,0:1)   start-of-step state (frozen)
  const double* xhe;     // (N3,0:2)
  double* xh_av;         // (N3,0:1)
  double* xhe_av;        // (N3,0:2)
  double* xh_int;        // (N3,0:1)
  double* xhe_int;       // (N3,0:2)
  float* temp;           // (N3,0:2) real(si)
  const double* rates;   // phih | phihe0 | phihe1 | phiheat
  size_t N3;
  const float* clumping_grid;  // type_of_clumping == 5: material's clumping_grid(i,j,k); else nullptr
};
struct ChemTotals {
  int conv_flag;
  int nit_max;
  unsigned long long nit_total;
  unsigned long long nsub_total;  // explicit thermal sub-steps taken (drives the choice of global-pass kernel)
  double last_coef_T;             // avg_temper of the last ini_rec_colion_factors call of the last mesh cell: the state
                                  // the reference's module globals are left in (photonstatistics.f90:180-194 reads them)
};

// 4 CTAs/SM (128 registers, ~110 bytes of spills) measured 3-5 % faster than the natural 168 registers / 3 CTAs on the
// 128^3 pass and on both config-5 variants
__global__ void __launch_bounds__(128, 4)
k_global_pass(ChemPtrs P, double dt, ChemTotals* tot, int* __restrict__ nit_out, size_t p_begin, size_t p_end) {
  // cells [p_begin, p_end) of the mesh: the whole mesh on one rank, this rank's share when the pass is split over
  // ranks (cells are independent, evolve.F90:477-484)
  const size_t N3 = P.N3;
  const size_t p = p_begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool iso = d_run.isothermal != 0;
  int vote = 0, nit = 0, nsub = 0;
  if (p < p_end) {
    Ion ion;
    // evolve_point.F90:368-378.  Only the live members are loaded: ion%h(1), ion%he(2) are
    // overwritten by doric before any use and ion%h_old(0), ion%he_old(0) are never read.
    ion.h0 = fmax(epsilon, P.xh_int[p]);
    ion.he0 = fmax(epsilon, P.xhe_int[p]);
    ion.he1 = fmax(epsilon, P.xhe_int[p + N3]);
    ion.h1 = 0.0; ion.he2 = 0.0;
    ion.h_old0 = 0.0; ion.he_old0 = 0.0;
    ion.h_old1 = fmax(epsilon, P.xh[p + N3]);
    ion.he_old1 = fmax(epsilon, P.xhe[p + N3]);
    ion.he_old2 = fmax(epsilon, P.xhe[p + 2 * N3]);
    const double yh0_av_old = P.xh_av[p], yh1 = P.xh_av[p + N3];
    const double yhe0_av_old = P.xhe_av[p], yhe1 = P.xhe_av[p + N3], yhe2_av_old = P.xhe_av[p + 2 * N3];
    ion.h_av0 = fmax(epsilon, yh0_av_old); ion.h_av1 = fmax(epsilon, yh1);
    ion.he_av0 = fmax(epsilon, yhe0_av_old); ion.he_av1 = fmax(epsilon, yhe1); ion.he_av2 = fmax(epsilon, yhe2_av_old);
    const double n = P.ndens[p];
    double temp_av_old, temper_old;
    if (iso) { temp_av_old = d_run.temper_val; temper_old = d_run.temper_val; }
    else { temp_av_old = (double)P.temp[p + N3]; temper_old = (double)P.temp[p + 2 * N3]; }
    const double phiHI = P.rates[p], phiHeI = P.rates[N3 + p], phiHeII = P.rates[2 * N3 + p];
    const double heat = iso ? 0.0 : P.rates[3 * N3 + p];
    RecCol rc;
    if (iso) ini_rec_colion_factors(d_run.temper_val, rc);  // mat_ini_test.F90:168
    double avg_temper = temp_av_old, temper1;
    const double clumping = P.clumping_grid ? (double)P.clumping_grid[p] : d_run.clumping;  // evolve_point.F90:484
    nit = do_chemistry(dt, n, ion, phiHI, phiHeI, phiHeII, heat, temper_old, avg_temper, temper1, rc, clumping, &nsub,
                       p == N3 - 1 ? &tot->last_coef_T : nullptr);
    double temp_av_new = temp_av_old;
    if (!iso) {  // set_temperature_point: stored as real(si), read back as such (:404)
      const float t0 = (float)temper1, t1 = (float)avg_temper;
      P.temp[p] = t0; P.temp[p + N3] = t1;
      temp_av_new = (double)t1;
    }
    // :406-424
    const double mfc = minimum_fractional_change, mfa = minimum_fraction_of_atoms;
    if ((fabs(ion.h_av0 - yh0_av_old) > mfc && fabs((ion.h_av0 - yh0_av_old) / ion.h_av0) > mfc && ion.h_av0 > mfa) ||
        (fabs(ion.he_av0 - yhe0_av_old) > mfc && fabs((ion.he_av0 - yhe0_av_old) / ion.he_av0) > mfc && ion.he_av0 > mfa) ||
        (fabs(ion.he_av2 - yhe2_av_old) > mfc && fabs((ion.he_av2 - yhe2_av_old) / ion.he_av2) > mfc && ion.he_av2 > mfa) ||
        ((fabs((temp_av_old - temp_av_new) / temp_av_new) > 1.0e-1) && (fabs(temp_av_new - temp_av_old) > 100.0)))
      vote = 1;
    P.xh_int[p] = ion.h0; P.xh_int[p + N3] = ion.h1;
    P.xh_av[p] = ion.h_av0; P.xh_av[p + N3] = ion.h_av1;
    P.xhe_int[p] = ion.he0; P.xhe_int[p + N3] = ion.he1; P.xhe_int[p + 2 * N3] = ion.he2;
    P.xhe_av[p] = ion.he_av0; P.xhe_av[p + N3] = ion.he_av1; P.xhe_av[p + 2 * N3] = ion.he_av2;
    if (nit_out) nit_out[p] = nit;
  }
  const unsigned int v = __reduce_add_sync(0xffffffffu, (unsigned)vote);
  const unsigned int ns = __reduce_add_sync(0xffffffffu, (unsigned)nit);
  const unsigned int nm = __reduce_max_sync(0xffffffffu, (unsigned)nit);
  const unsigned int nsb = __reduce_add_sync(0xffffffffu, (unsigned)nsub);
  if ((threadIdx.x & 31) == 0) {
    if (v) atomicAdd(&tot->conv_flag, (int)v);
    atomicAdd(&tot->nit_total, (unsigned long long)ns);
    if (nsb) atomicAdd(&tot->nsub_total, (unsigned long long)nsb);
    atomicMax(&tot->nit_max, (int)nm);
  }
}

// ------------------------------------------------------------------------------------------------
// (c') queue-driven global pass.  do_chemistry's cost per cell varies by three orders of magnitude (1..401 outer
// iterations x 1..10001 explicit thermal sub-steps), so with one cell per thread a warp idles behind its slowest lane
// (measured on the BASELINE config-5 inputs: mean 42 sub-steps per cell, mean of the per-warp maximum 272).  Here every
// lane is a small state machine -- IONIZE (coefficients + doric x2) -> THERMAL (a bounded burst of sub-steps) ->
// TEST (convergence; store or loop) -- and a lane that finishes its cell immediately draws the next cell index from a
// global counter (one atomic per warp refill).  The arithmetic per cell is exactly that of do_chemistry.
// ------------------------------------------------------------------------------------------------
constexpr int CHEM_BURST = 32;  // default thermal sub-steps per state-machine turn (8: 36 ms, 32: 30 ms on config 5 at 256^3)
// (holding lanes back until a dozen of them wait for IONIZE, so that block runs better filled: no change, 28.2-28.5 ms at
// any threshold -- profiles/r2_ab4_chem.log)

__global__ void __launch_bounds__(128, 4)
k_global_pass_q(ChemPtrs P, double dt, ChemTotals* tot, int* __restrict__ nit_out, unsigned long long* next_cell,
                size_t p_begin, size_t p_end, int burst, int thermal_min) {
  const size_t N3 = P.N3;
  const bool iso = d_run.isothermal != 0;
  const unsigned lane = threadIdx.x & 31;
  enum { IDLE = 0, IONIZE = 1, THERMAL = 2, TEST = 3 };
  int phase = IDLE;
  bool exhausted = false;
  long long p = -1;
  Ion ion;
  RecCol rc;
  ThermState TS;
  ChemIter it;
  double n = 0, phiHI = 0, phiHeI = 0, phiHeII = 0, heat = 0, de = 0, clumping = d_run.clumping;
  double temper0 = 0, temper1 = 0, avg_temper = 0, temp_av_old = 0, yh0_old = 0, yhe0_old = 0, yhe2_old = 0;
  int nit = 0, votes = 0, nit_sum = 0, nit_max = 0, nsub = 0;
  if (iso) ini_rec_colion_factors(d_run.temper_val, rc);  // mat_ini_test.F90:168

  for (;;) {
    // ---- refill idle lanes ----------------------------------------------------------------------------------
    const unsigned want = __ballot_sync(0xffffffffu, phase == IDLE && !exhausted);
    if (want) {
      unsigned long long base = 0;
      if (lane == (unsigned)(__ffs(want) - 1)) base = atomicAdd(next_cell, (unsigned long long)__popc(want));
      base = __shfl_sync(0xffffffffu, base, __ffs(want) - 1);
      if (phase == IDLE && !exhausted) {
        const unsigned long long idx = p_begin + base + __popc(want & ((1u << lane) - 1));
        if (idx >= p_end) {
          exhausted = true;
        } else {
          p = (long long)idx;
          // evolve_point.F90:368-390, live members only (see k_global_pass)
          ion.h0 = fmax(epsilon, P.xh_int[p]);
          ion.he0 = fmax(epsilon, P.xhe_int[p]);
          ion.he1 = fmax(epsilon, P.xhe_int[p + N3]);
          ion.h1 = 0.0; ion.he2 = 0.0; ion.h_old0 = 0.0; ion.he_old0 = 0.0;
          ion.h_old1 = fmax(epsilon, P.xh[p + N3]);
          ion.he_old1 = fmax(epsilon, P.xhe[p + N3]);
          ion.he_old2 = fmax(epsilon, P.xhe[p + 2 * N3]);
          yh0_old = P.xh_av[p]; yhe0_old = P.xhe_av[p]; yhe2_old = P.xhe_av[p + 2 * N3];
          ion.h_av0 = fmax(epsilon, yh0_old); ion.h_av1 = fmax(epsilon, P.xh_av[p + N3]);
          ion.he_av0 = fmax(epsilon, yhe0_old); ion.he_av1 = fmax(epsilon, P.xhe_av[p + N3]);
          ion.he_av2 = fmax(epsilon, yhe2_old);
          n = P.ndens[p];
          if (P.clumping_grid) clumping = (double)P.clumping_grid[p];  // evolve_point.F90:484
          if (iso) { temp_av_old = d_run.temper_val; temper0 = d_run.temper_val; }
          else { temp_av_old = (double)P.temp[p + N3]; temper0 = (double)P.temp[p + 2 * N3]; }
          phiHI = P.rates[p]; phiHeI = P.rates[N3 + p]; phiHeII = P.rates[2 * N3 + p];
          heat = iso ? 0.0 : P.rates[3 * N3 + p];
          avg_temper = temp_av_old; temper1 = temper0; nit = 0;
          phase = IONIZE;
        }
      }
    }
    if (__all_sync(0xffffffffu, phase == IDLE)) break;  // every lane idle and the queue empty

    // thermal_min > 0 holds the THERMAL burst back until that many lanes wait in it (or nothing else is left to do this
    // turn), so that cheap cells stream through the other lanes while expensive ones gather.  Measured: no effect at any
    // threshold (profiles/r2_ab5_chem_thermal_batching.log), default 0.  What does matter is the warp-wide vote below:
    // it makes the lanes that left the divergent IONIZE block reconverge BEFORE the burst, so the sub-step loop runs once
    // for the warp instead of once per diverged group (config-5 thermal pass at 256^3: 28.4 -> 20.7 ms).
    const bool others_busy = __any_sync(0xffffffffu, phase == IONIZE);
    // ---- IONIZE: one do_chemistry iteration up to the thermal call (evolve_point.F90:488-600) ---------------------
    if (phase == IONIZE) {
      nit++;
      if (!iso && (size_t)p == N3 - 1) tot->last_coef_T = avg_temper;
      de = chem_ionization(dt, n, ion, phiHI, phiHeI, phiHeII, avg_temper, temper1, rc, it, clumping);
      temper1 = temper0;
      if (iso) {
        phase = TEST;
      } else {
        thermal_begin(TS, temper1, n, ion);
        phase = TS.active ? THERMAL : TEST;
      }
    }
    // ---- THERMAL: a burst of explicit sub-steps (thermal.f90:98-157) -----------------------------------------------
    const int n_thermal = __popc(__ballot_sync(0xffffffffu, phase == THERMAL));
    if (phase == THERMAL && (n_thermal >= thermal_min || !others_busy)) {
      bool done = false;
#pragma unroll 1
      for (int k = 0; k < burst && !done; k++) { done = thermal_substep(TS, dt, de, n, ion, heat); nsub++; }
      if (done) phase = TEST;
    }
    // ---- TEST: finish thermal, convergence test, store (evolve_point.F90:607-644, :397-435) --------------------------
    if (phase == TEST) {
      if (!iso) thermal_end(TS, dt, n, ion, temper1, avg_temper);
      if (chem_converged(ion, it, temper1) || nit > 400) {
        double temp_av_new = temp_av_old;
        if (!iso) {
          const float t0 = (float)temper1, t1 = (float)avg_temper;
          P.temp[p] = t0; P.temp[p + N3] = t1;
          temp_av_new = (double)t1;
        }
        const double mfc = minimum_fractional_change, mfa = minimum_fraction_of_atoms;
        if ((fabs(ion.h_av0 - yh0_old) > mfc && fabs((ion.h_av0 - yh0_old) / ion.h_av0) > mfc && ion.h_av0 > mfa) ||
            (fabs(ion.he_av0 - yhe0_old) > mfc && fabs((ion.he_av0 - yhe0_old) / ion.he_av0) > mfc && ion.he_av0 > mfa) ||
            (fabs(ion.he_av2 - yhe2_old) > mfc && fabs((ion.he_av2 - yhe2_old) / ion.he_av2) > mfc && ion.he_av2 > mfa) ||
            ((fabs((temp_av_old - temp_av_new) / temp_av_new) > 1.0e-1) && (fabs(temp_av_new - temp_av_old) > 100.0)))
          votes++;
        P.xh_int[p] = ion.h0; P.xh_int[p + N3] = ion.h1;
        P.xh_av[p] = ion.h_av0; P.xh_av[p + N3] = ion.h_av1;
        P.xhe_int[p] = ion.he0; P.xhe_int[p + N3] = ion.he1; P.xhe_int[p + 2 * N3] = ion.he2;
        P.xhe_av[p] = ion.he_av0; P.xhe_av[p + N3] = ion.he_av1; P.xhe_av[p + 2 * N3] = ion.he_av2;
        if (nit_out) nit_out[p] = nit;
        nit_sum += nit; nit_max = max(nit_max, nit);
        phase = IDLE;
      } else {
        phase = IONIZE;
      }
    }
  }
  const unsigned int v = __reduce_add_sync(0xffffffffu, (unsigned)votes);
  const unsigned int ns = __reduce_add_sync(0xffffffffu, (unsigned)nit_sum);
  const unsigned int nm = __reduce_max_sync(0xffffffffu, (unsigned)nit_max);
  const unsigned int nsb = __reduce_add_sync(0xffffffffu, (unsigned)nsub);
  if (lane == 0) {
    if (v) atomicAdd(&tot->conv_flag, (int)v);
    atomicAdd(&tot->nit_total, (unsigned long long)ns);
    if (nsb) atomicAdd(&tot->nsub_total, (unsigned long long)nsb);
    atomicMax(&tot->nit_max, (int)nm);
  }
}

// Cross-rank combination of the global-pass counters when the pass is split over ranks: sums travel as FP64 (exact
// below 2^53), the maximum as int32.
__global__ void k_chem_pack(const ChemTotals* tot, double* sum4, int* max1) {
  sum4[0] = (double)tot->conv_flag; sum4[1] = (double)tot->nit_total; sum4[2] = (double)tot->nsub_total;
  sum4[3] = tot->last_coef_T;  // written by the rank that owns the last mesh cell only, 0 elsewhere
  max1[0] = tot->nit_max;
}
__global__ void k_chem_unpack(ChemTotals* tot, const double* sum4, const int* max1) {
  tot->conv_flag = (int)sum4[0]; tot->nit_total = (unsigned long long)sum4[1];
  tot->nsub_total = (unsigned long long)sum4[2]; tot->last_coef_T = sum4[3];
  tot->nit_max = max1[0];
}

// photonstatistics.f90:117-147 / :208-247 : sum_p ndens*x for 5 species
__global__ void k_state_sums(const double* __restrict__ ndens, const double* __restrict__ xh, const double* __restrict__ xhe,
                             size_t N3, double* out5) {
  double s[5] = {0, 0, 0, 0, 0};
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < N3; p += (size_t)gridDim.x * blockDim.x) {
    const double n = ndens[p];
    s[0] += n * xh[p]; s[1] += n * xh[p + N3];
    s[2] += n * xhe[p]; s[3] += n * xhe[p + N3]; s[4] += n * xhe[p + 2 * N3];
  }
  __shared__ double sh[5][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int q = 0; q < 5; q++) {
    double v = s[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[q][w] = v;
  }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    for (int q = 0; q < 5; q++) {
      double v = lane < nw ? sh[q][lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(out5 + q, v);
    }
  }
}

// photonstatistics.f90:150-204 total_rates: recombinations that do not ionize, collisional ionizations and ionizing
// He recombinations, summed over the mesh with the coefficients the module globals hold (coef_T, see ChemTotals).
// out3 += (totrec, totcollisions, recomions) before the *vol*dt factor.
__global__ void k_total_rates(const double* __restrict__ ndens, const double* __restrict__ xh_av,
                              const double* __restrict__ xhe_av, size_t N3, double coef_T, double* out3,
                              const float* __restrict__ clumping_grid) {
  __shared__ RecCol rcs;
  __shared__ double sh[3][32];
  if (threadIdx.x == 0) ini_rec_colion_factors(coef_T, rcs);
  __syncthreads();
  const RecCol rc = rcs;
  double clumping = d_run.clumping;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < N3; p += (size_t)gridDim.x * blockDim.x) {
    const double n = ndens[p];
    if (clumping_grid) clumping = (double)clumping_grid[p];  // photonstatistics.f90:176
    const double h0 = xh_av[p], h1 = xh_av[p + N3], he0 = xhe_av[p], he1 = xhe_av[p + N3], he2 = xhe_av[p + 2 * N3];
    const double ne = electrondens(n, h1, he1, he2);
    s0 += n * (h1 * rc.brech0 * (1.0 - abu_he) + he1 * rc.breche0 * abu_he * 0.04) * ne * clumping;
    s1 += n * ne * (h0 * rc.colli_HI + he0 * rc.colli_HeI + he1 * rc.colli_HeII);
    s2 += n * abu_he * clumping * (he2 * 1.121 * rc.breche1 + he1 * rc.breche0 * 0.96) * abu_he * ne;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double v[3] = {s0, s1, s2};
  for (int q = 0; q < 3; q++) {
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
    if (lane == 0) sh[q][w] = v[q];
  }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    for (int q = 0; q < 3; q++) {
      double x = lane < nw ? sh[q][lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) atomicAdd(out3 + q, x);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fine-grained parity hooks
// ------------------------------------------------------------------------------------------------
__global__ void k_photoion_batch(int n, const double* __restrict__ col6, const double* __restrict__ vol, double nf0,
                                 double nf1, double nf2, const double* __restrict__ i_state, double* __restrict__ out6) {
#if C2RAY_TABLOG
  postab_stage();
#endif
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double* q = col6 + 6 * (size_t)t;
  const double nflux[3] = {nf0, nf1, nf2};
  const SecIon y = secion_factors(i_state[t]);
  const PhotOut r = d_run.isothermal ? photoion_rates<true, true>(q[0], q[1], q[2], q[3], q[4], q[5], vol[t], nflux, y)
                                     : photoion_rates<false, true>(q[0], q[1], q[2], q[3], q[4], q[5], vol[t], nflux, y);
  double* o = out6 + 6 * (size_t)t;
  o[0] = r.photo_HI; o[1] = r.photo_HeI; o[2] = r.photo_HeII; o[3] = r.heat; o[4] = r.photo_in; o[5] = r.photo_out;
}

__global__ void k_chemistry_batch(int n, double dt, const double* __restrict__ ndens, double* __restrict__ ion15,
                                  const double* __restrict__ phi4, double* __restrict__ T3, int* __restrict__ nit_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double* v = ion15 + 15 * (size_t)t;
  Ion ion;
  ion.h0 = v[0]; ion.h1 = v[1]; ion.he0 = v[2]; ion.he1 = v[3]; ion.he2 = v[4];
  ion.h_av0 = v[5]; ion.h_av1 = v[6]; ion.he_av0 = v[7]; ion.he_av1 = v[8]; ion.he_av2 = v[9];
  ion.h_old0 = v[10]; ion.h_old1 = v[11]; ion.he_old0 = v[12]; ion.he_old1 = v[13]; ion.he_old2 = v[14];
  RecCol rc;
  if (d_run.isothermal) ini_rec_colion_factors(d_run.temper_val, rc);
  double avg = T3[3 * t + 1], t1;
  const int nit = do_chemistry(dt, ndens[t], ion, phi4[4 * t], phi4[4 * t + 1], phi4[4 * t + 2], phi4[4 * t + 3],
                               T3[3 * t + 2], avg, t1, rc, d_run.clumping);
  T3[3 * t] = t1; T3[3 * t + 1] = avg;
  v[0] = ion.h0; v[1] = ion.h1; v[2] = ion.he0; v[3] = ion.he1; v[4] = ion.he2;
  v[5] = ion.h_av0; v[6] = ion.h_av1; v[7] = ion.he_av0; v[8] = ion.he_av1; v[9] = ion.he_av2;
  nit_out[t] = nit;
}

// doric.f90:35-313 alone, coefficients from ini_rec_colion_factors(T[t]) (cgsconstants.f90:140): one call per state
__global__ void k_doric_batch(int n, double dt, const double* __restrict__ rhe, double* __restrict__ ion15,
                              const double* __restrict__ phi3, const double* __restrict__ fr4, const double* __restrict__ T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double* v = ion15 + 15 * (size_t)t;
  Ion ion;
  ion.h0 = v[0]; ion.h1 = v[1]; ion.he0 = v[2]; ion.he1 = v[3]; ion.he2 = v[4];
  ion.h_av0 = v[5]; ion.h_av1 = v[6]; ion.he_av0 = v[7]; ion.he_av1 = v[8]; ion.he_av2 = v[9];
  ion.h_old0 = v[10]; ion.h_old1 = v[11]; ion.he_old0 = v[12]; ion.he_old1 = v[13]; ion.he_old2 = v[14];
  RecCol rc;
  ini_rec_colion_factors(T[t], rc);
  DoricFrac fr;
  fr.y = fr4[4 * t]; fr.z = fr4[4 * t + 1]; fr.y2a = fr4[4 * t + 2]; fr.y2b = fr4[4 * t + 3];
  doric(dt, rhe[t], ion, phi3[3 * t], phi3[3 * t + 1], phi3[3 * t + 2], fr, rc, d_run.clumping);
  v[0] = ion.h0; v[1] = ion.h1; v[2] = ion.he0; v[3] = ion.he1; v[4] = ion.he2;
  v[5] = ion.h_av0; v[6] = ion.h_av1; v[7] = ion.he_av0; v[8] = ion.he_av1; v[9] = ion.he_av2;
}

// thermal.f90:22-174 alone: end_temper in/out, avg_temper in/out (untouched at or below minitemp), sub-step count out
__global__ void k_thermal_batch(int n, double dt, double* __restrict__ end_temper, double* __restrict__ avg_temper,
                                const double* __restrict__ ne, const double* __restrict__ ndens,
                                const double* __restrict__ ion15, const double* __restrict__ heat, int* __restrict__ nsub) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double* v = ion15 + 15 * (size_t)t;
  Ion ion;
  ion.h0 = v[0]; ion.h1 = v[1]; ion.he0 = v[2]; ion.he1 = v[3]; ion.he2 = v[4];
  ion.h_av0 = v[5]; ion.h_av1 = v[6]; ion.he_av0 = v[7]; ion.he_av1 = v[8]; ion.he_av2 = v[9];
  ion.h_old0 = v[10]; ion.h_old1 = v[11]; ion.he_old0 = v[12]; ion.he_old1 = v[13]; ion.he_old2 = v[14];
  double e = end_temper[t], a = avg_temper[t];
  nsub[t] = thermal(dt, e, a, ne[t], ndens[t], ion, heat[t]);
  end_temper[t] = e; avg_temper[t] = a;
}

__global__ void k_rec_colion_batch(int n, const double* __restrict__ T, double* __restrict__ out12) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  RecCol r;
  ini_rec_colion_factors(T[t], r);
  double* o = out12 + 12 * (size_t)t;
  o[0] = r.arech0; o[1] = r.brech0; o[2] = r.areche0; o[3] = r.breche0; o[4] = r.oreche0; o[5] = r.areche1;
  o[6] = r.breche1; o[7] = r.treche1; o[8] = r.colli_HI; o[9] = r.colli_HeI; o[10] = r.colli_HeII; o[11] = r.v;
}

// ------------------------------------------------------------------------------------------------
// radiation_tables.f90:172-422 spec_integration on the device.  One block per (tau index, band).
// ------------------------------------------------------------------------------------------------
struct TableBuild {
  double freq_min[NumFreqBnd], delta_freq[NumFreqBnd], plidx[NumFreqBnd];  // per band; plidx = index used for the band
  double romw[NumFreq + 1];
  double R_star2, h_over_kT;
  double scaling[3], index[3];  // [1],[2]: pl, qpl
  int active[3];
  int isothermal;
  double* photo_thick[3];
  double* photo_thin[3];
  double* heat_thick[3];
  double* heat_thin[3];
};

__global__ void __launch_bounds__(128) k_build_tables(const TableBuild* __restrict__ tb) {
  const int it = blockIdx.x;      // 0..NumTau
  const int q = blockIdx.y;       // band 0..46
  const int b = q + 1;
  const double tau = it == 0 ? 0.0 : pow(FL(10.0f), minlogtau + dlogtau * (double)(it - 1));
  const int nsp = (b <= NumBndin1) ? 1 : (b <= NumBndin1 + NumBndin2 ? 2 : 3);
  const int hcol0 = (b <= NumBndin1) ? 0 : (b <= NumBndin1 + NumBndin2 ? 2 * b - NumBndin1 - 2 : 3 * b - NumBndin2 - NumBndin1 * 2 - 3);
  const double ionf[3] = {ion_freq_HI, ion_freq_HeI, ion_freq_HeII};
  const double w = tb->delta_freq[q];
  __shared__ double red[4][8];
  for (int s = 0; s < 3; s++) {
    if (!tb->active[s]) continue;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // thick, thin, heat thick x3, heat thin x3
    for (int x = threadIdx.x; x <= NumFreq; x += blockDim.x) {
      const double f = tb->freq_min[q] + tb->delta_freq[q] * (double)x;
      const double cs = pow(f / tb->freq_min[q], -tb->plidx[q]);
      double thick = 0.0, thin = 0.0;
      if (tau * cs < FL(700.0f)) {
        if (s == 0) {
          if (f * tb->h_over_kT < FL(700.0f)) {
            const double base = 4.0 * pi * tb->R_star2 * two_pi_over_c_square * f * f;
            const double ex = exp(-tau * cs), den = exp(f * tb->h_over_kT) - 1.0;
            thick = base * ex / den;
            thin = base * cs * ex / den;
          }
        } else {
          const double pw = tb->scaling[s] * pow(f, -tb->index[s]);
          const double ex = exp(-tau * cs);
          thick = pw * ex;
          thin = pw * cs * ex;
        }
      }
      const double wr = w * tb->romw[x];
      acc[0] += thick * wr; acc[1] += thin * wr;
      if (!tb->isothermal)
        for (int sp = 0; sp < nsp; sp++) {
          const double e = hplanck * (f - ionf[sp]);
          acc[2 + sp] += e * thick * wr; acc[5 + sp] += e * thin * wr;
        }
    }
    // block reduction
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int k = 0; k < 8; k++) {
      double v = acc[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) red[wp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      const int k = threadIdx.x;
      const double v = red[0][k] + red[1][k] + red[2][k] + red[3][k];
      const size_t o = (size_t)q * (NumTau + 1) + it;
      if (k == 0) tb->photo_thick[s][o] = v;
      else if (k == 1) tb->photo_thin[s][o] = v;
      else if (!tb->isothermal) {
        const int sp = (k - 2) % 3;
        if (sp < nsp) {
          const size_t ho = (size_t)(hcol0 + sp) * (NumTau + 1) + it;
          if (k < 5) tb->heat_thick[s][ho] = v; else tb->heat_thin[s][ho] = v;
        }
      }
    }
    __syncthreads();
  }
}

// cinterp against a full-grid scratch (parity hook for column_density.f90:28; the sweep kernel carries the same
// arithmetic on its shell buffers)
__global__ void k_cinterp_batch(int n, const int* __restrict__ pos, int i0, int j0, int k0,
                                const double* __restrict__ cdh, const double* __restrict__ cdhe, size_t N3,
                                double* __restrict__ out4) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int m0 = d_run.mesh[0], m1 = d_run.mesh[1], m2 = d_run.mesh[2];
  const int i = pos[3 * t], j = pos[3 * t + 1], k = pos[3 * t + 2];
  const int di = i - i0, dj = j - j0, dk = k - k0;
  const int ia = abs(di), ja = abs(dj), ka = abs(dk);
  const int sgi = di >= 0 ? 1 : -1, sgj = dj >= 0 ? 1 : -1, sgk = dk >= 0 ? 1 : -1;
  int ax;
  if (ka >= ja && ka >= ia) ax = 2; else if (ja >= ia && ja >= ka) ax = 1; else ax = 0;
  int du, dv, dw, sgu, sgv, sgw, u0, v0;
  if (ax == 2) { du = di; dv = dj; dw = dk; sgu = sgi; sgv = sgj; sgw = sgk; u0 = i0; v0 = j0; }
  else if (ax == 1) { du = di; dv = dk; dw = dj; sgu = sgi; sgv = sgk; sgw = sgj; u0 = i0; v0 = k0; }
  else { du = dj; dv = dk; dw = di; sgu = sgj; sgv = sgk; sgw = sgi; u0 = j0; v0 = k0; }
  const double fu = (double)du, fv = (double)dv, fw = (double)dw;
  const double alam = __ddiv_rn((double)((float)(dw - sgw) + (float)sgw * 0.5f), fw);
  const double uc = __dadd_rn(__dmul_rn(alam, fu), (double)u0);
  const double vc = __dadd_rn(__dmul_rn(alam, fv), (double)v0);
  const double um = (double)((float)(u0 + du - sgu) + 0.5f * (float)sgu);
  const double vm = (double)((float)(v0 + dv - sgv) + 0.5f * (float)sgv);
  const double du_ = __dmul_rn(2.0, fabs(__dsub_rn(uc, um)));
  const double dv_ = __dmul_rn(2.0, fabs(__dsub_rn(vc, vm)));
  const double omu = __dsub_rn(1.0, du_), omv = __dsub_rn(1.0, dv_);
  const double sw[4] = {__dmul_rn(omu, omv), __dmul_rn(omv, du_), __dmul_rn(omu, dv_), __dmul_rn(du_, dv_)};
  const int cu[4] = {du - sgu, du, du - sgu, du}, cv[4] = {dv - sgv, dv - sgv, dv, dv};
  double cH[4], c0[4], c1[4];
  for (int q = 0; q < 4; q++) {
    int o[3];
    if (ax == 2) { o[0] = cu[q]; o[1] = cv[q]; o[2] = dw - sgw; }
    else if (ax == 1) { o[0] = cu[q]; o[2] = cv[q]; o[1] = dw - sgw; }
    else { o[1] = cu[q]; o[2] = cv[q]; o[0] = dw - sgw; }
    const size_t p = (size_t)wrap0(i0 + o[0], m0) + (size_t)m0 * ((size_t)wrap0(j0 + o[1], m1) + (size_t)m1 * wrap0(k0 + o[2], m2));
    cH[q] = cdh[p]; c0[q] = cdhe[p]; c1[q] = cdhe[p + N3];
  }
  double r3[3];
  const double* cc[3] = {cH, c0, c1};
  const double sg[3] = {sigma_HI_at_ion_freq, sigma_HeI_at_ion_freq, sigma_HeII_at_ion_freq};
  for (int s = 0; s < 3; s++) {
    double num = 0.0, den = 0.0;
    for (int q = 0; q < 4; q++) { const double w = sw[q] * weightf_fast(cc[s][q], sg[s]); num += cc[s][q] * w; den += w; }
    r3[s] = fdiv(num, den);
  }
  const int wa = abs(dw), ua = abs(du), va = abs(dv);
  if (wa == 1 && (ua == 1 || va == 1)) {
    const double f = (ua == 1 && va == 1) ? sqrt3 : sqrt2;
    r3[0] *= f; r3[1] *= f; r3[2] *= f;
  }
  out4[4 * t] = r3[0]; out4[4 * t + 1] = r3[1]; out4[4 * t + 2] = r3[2];
  out4[4 * t + 3] = sqrt(fdiv(fu * fu + fv * fv, fw * fw) + 1.0);
}

// FP64 FMA throughput probe
__global__ void k_fp64_probe(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace c2
