// Pipe-peak microbenchmarks for the roofline denominators that MEASURED_PEAKS.json does not carry:
//   dfma   FP64 FMA pipe            (the banded LDL^T env-step kernel)
//   dmma   mma.sync.m8n8k4.f64      (the dense DMMA Cholesky variant)
//   hmma   mma.sync.m16n8k16 f16    (the warp-level products of the actor kernel's operand generators)
// Every thread runs ILP independent dependent-chains; the grid fills every SM with `warps` warps per scheduler.
// Prints one JSON line per measurement.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks peaks.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int ITERS = 4096;

__global__ void dfma_kernel(double* out, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void dmma_kernel(double* out, double a, double b) {
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

__global__ void hmma_kernel(float* out, uint32_t a, uint32_t b) {
  float c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a), "r"(a), "r"(a), "r"(a), "r"(b), "r"(b));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678f) out[0] = s;
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double* dout;
  cudaMalloc(&dout, 64);
  const int warps_list[] = {1, 2, 4, 8};
  for (int w : warps_list) {                       // warps per scheduler
    const int threads = 4 * w * 32 > 1024 ? 1024 : 4 * w * 32, ctas = sms * ((4 * w * 32 + threads - 1) / threads);
    const double nthreads = (double)ctas * threads;
    float ms = time_ms([&] { dfma_kernel<<<ctas, threads>>>(dout, 1.0000001, 1e-9); });
    printf("{\"pipe\": \"dfma\", \"warps_per_scheduler\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"flop_per_clk_per_sm\": %.1f}\n", w, ms,
           nthreads * 8.0 * ITERS * 2 / (ms * 1e-3) / 1e12, nthreads * 8.0 * ITERS * 2 / (ms * 1e-3) / sms / (khz * 1e3));
    ms = time_ms([&] { dmma_kernel<<<ctas, threads>>>(dout, 1.0000001, 1e-9); });
    printf("{\"pipe\": \"dmma_m8n8k4\", \"warps_per_scheduler\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"flop_per_clk_per_sm\": %.1f}\n", w, ms,
           nthreads / 32 * 4.0 * ITERS * (8 * 8 * 4 * 2) / (ms * 1e-3) / 1e12,
           nthreads / 32 * 4.0 * ITERS * (8 * 8 * 4 * 2) / (ms * 1e-3) / sms / (khz * 1e3));
    ms = time_ms([&] { hmma_kernel<<<ctas, threads>>>((float*)dout, 0x3c003c00u, 0x3c003c00u); });
    printf("{\"pipe\": \"hmma_m16n8k16_f16_f32\", \"warps_per_scheduler\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"flop_per_clk_per_sm\": %.1f, "
           "\"cycles_per_hmma_per_scheduler\": %.2f}\n", w, ms,
           nthreads / 32 * 4.0 * ITERS * (16 * 8 * 16 * 2) / (ms * 1e-3) / 1e12,
           nthreads / 32 * 4.0 * ITERS * (16 * 8 * 16 * 2) / (ms * 1e-3) / sms / (khz * 1e3),
           (ms * 1e-3) * (khz * 1e3) / (w * 4.0 * ITERS));
  }
  printf("{\"sms\": %d, \"clock_khz_attr\": %d}\n", sms, khz);
  cudaFree(dout);
  return 0;
}
