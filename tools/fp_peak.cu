// FP issue-rate micro-benchmark ("FP32/FP64 issue peaks are not in MEASURED_PEAKS.json ->
// measure with an FMA micro-benchmark").  Independent dependency chains per thread, enough warps to fill the SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp_peak.cu -o /tmp/fp_peak && /tmp/fp_peak
// Prints one JSON object: FP64 FMA, FP64 separate add+mul (what --fmad=false code issues), FP32 FMA, in Tflop/s
// and in G warp-instructions/s.
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, bool FUSED>
__global__ void Spin(T *out, int iters, T a, T b) {
  T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    if (FUSED) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    } else {
      x0 = x0 * a; x1 = x1 * a; x2 = x2 * a; x3 = x3 * a; x4 = x4 * a; x5 = x5 * a; x6 = x6 * a; x7 = x7 * a;
      x0 = x0 + b; x1 = x1 + b; x2 = x2 + b; x3 = x3 + b; x4 = x4 + b; x5 = x5 + b; x6 = x6 + b; x7 = x7 + b;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <typename T, bool FUSED>
double Measure(int iters) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, threads = 256;
  T *out;
  cudaMalloc(&out, sizeof(T) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  Spin<T, FUSED><<<blocks, threads>>>(out, iters / 10, (T)1.000001, (T)1e-7);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    Spin<T, FUSED><<<blocks, threads>>>(out, iters, (T)1.000001, (T)1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaFree(out);
  const double instr = (double)blocks * threads * iters * (FUSED ? 8.0 : 16.0);  // thread-level instructions
  return instr / (best * 1e-3);
}

int main() {
  const double d_fma = Measure<double, true>(20000);
  const double d_sep = Measure<double, false>(10000);
  const double f_fma = Measure<float, true>(40000);
  const double f_sep = Measure<float, false>(20000);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f, \"fp64_addmul_tflops\": %.2f, \"fp64_ginstr_per_s\": %.1f, "
         "\"fp32_fma_tflops\": %.2f, \"fp32_addmul_tflops\": %.2f, \"fp32_ginstr_per_s\": %.1f}\n",
         p.name, p.multiProcessorCount, d_fma * 2 / 1e12, d_sep / 1e12, d_sep / 1e9, f_fma * 2 / 1e12, f_sep / 1e12, f_sep / 1e9);
  return 0;
}
