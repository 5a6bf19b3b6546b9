// Dependent-issue latencies of the instruction kinds the sweep's critical path is made of (one warp, one SM).
// Build and run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/latency_probe tools/latency_probe.cu && /tmp/latency_probe
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double c_tab[1024];
#define N 4096
template <int KIND>
__global__ void probe(double* out, long long* cyc, double x, int idx, const double* g) {
  __shared__ double sh[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = (double)((i * 7 + 1) & 1023);
  __syncthreads();
  double a = x, b = x * 0.5 + 1e-3;
  int k = idx;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) {
    if (KIND == 0) a = fma(a, b, b);                                   // DFMA chain
    if (KIND == 1) a = a + b;                                          // DADD chain
    if (KIND == 2) a = a * b;                                          // DMUL chain
    if (KIND == 3) { k = (int)a; a = (double)k + b; }                  // F2I + I2F + DADD
    if (KIND == 4) { k = (int)c_tab[k & 1023]; }                       // LDC indexed + F2I
    if (KIND == 5) { k = (int)sh[k & 1023]; }                          // LDS + F2I
    if (KIND == 6) { k = (int)__ldg(g + (k & 1023)); }                 // LDG (L1 hit) + F2I
    if (KIND == 7) { a = fmax(a, b) + b; }                             // DMNMX? + DADD
    if (KIND == 8) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r + b; }  // MUFU.RCP64H + DADD
    if (KIND == 9) { a = __hiloint2double(__double2hiint(a) + 1, __double2loint(a)) + b; }  // int ALU on the high word + DADD
    if (KIND == 10) { k = (int)(float)k + 1; }                         // I2F.F32 + F2I.F32 + IADD
    if (KIND == 11) { a = (double)(int)rint(a) + b; }                   // rint path
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[KIND] = t1 - t0; }
  out[threadIdx.x + 32 * KIND] = a + k;
}
int main() {
  double *out, *g; long long* cyc;
  cudaMalloc(&out, 32 * 16 * 8); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&g, 1024 * 8);
  double h[1024]; for (int i = 0; i < 1024; i++) h[i] = (double)((i * 7 + 1) & 1023);
  cudaMemcpy(g, h, sizeof(h), cudaMemcpyHostToDevice); cudaMemcpyToSymbol(c_tab, h, sizeof(h));
  const char* names[] = {"DFMA", "DADD", "DMUL", "F2I.F64+I2F.F64+DADD", "LDC indexed + F2I.F64", "LDS + F2I.F64", "LDG(L1 hit) + F2I.F64", "fmax + DADD", "MUFU.RCP64H + DADD", "hi-word int add + DADD", "I2F.F32+F2I.F32+IADD", "rint+F2I+I2F+DADD"};
  for (int rep = 0; rep < 2; rep++) {
    probe<0><<<1, 32>>>(out, cyc, 1.0000001, 3, g); probe<1><<<1, 32>>>(out, cyc, 1.0000001, 3, g); probe<2><<<1, 32>>>(out, cyc, 1.0000001, 3, g);
    probe<3><<<1, 32>>>(out, cyc, 1.0000001, 3, g); probe<4><<<1, 32>>>(out, cyc, 1.0000001, 3, g); probe<5><<<1, 32>>>(out, cyc, 1.0000001, 3, g);
    probe<6><<<1, 32>>>(out, cyc, 1.0000001, 3, g); probe<7><<<1, 32>>>(out, cyc, 1.0000001, 3, g); probe<8><<<1, 32>>>(out, cyc, 1.0000001, 3, g);
    probe<9><<<1, 32>>>(out, cyc, 1.0000001, 3, g); probe<10><<<1, 32>>>(out, cyc, 1.0000001, 3, g); probe<11><<<1, 32>>>(out, cyc, 1.0000001, 3, g);
    cudaDeviceSynchronize();
  }
  long long hc[16]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  for (int i = 0; i < 12; i++) printf("%-28s %7.2f cycles per link\n", names[i], (double)hc[i] / N);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
