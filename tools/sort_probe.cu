// sort_probe.cu -- where the cycles of seed_sort_kernel go (clock64 probes of thread 0), on synthetic keys.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDPX_SORT_PROBE -Ideplex_b200/csrc -Iinclude tools/sort_probe.cu -o gpurun_out/sort_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "seed_sort.cuh"
using namespace dpx;
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 19200;
  std::vector<int16_t> bin(n);
  std::vector<float> mse(n);
  srand(1);
  for (int i = 0; i < n; ++i) { bin[i] = (rand() % 100 < 95) ? (rand() % 7 == 0 ? rand() % 400 : 37 + rand() % 3) : -1; mse[i] = (rand() % 100000) * 1e-4f; }
  int16_t* d_bin; float* d_mse; unsigned long long *d_a, *d_b;
  cudaMalloc(&d_bin, n * 2); cudaMalloc(&d_mse, n * 4); cudaMalloc(&d_a, n * 8); cudaMalloc(&d_b, n * 8);
  cudaMemcpy(d_bin, bin.data(), n * 2, cudaMemcpyHostToDevice); cudaMemcpy(d_mse, mse.data(), n * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    long long zero[16] = {};
    cudaMemcpyToSymbol(g_sort_probe, zero, sizeof(zero));
    cudaEventRecord(e0);
    launch_seed_sort(d_bin, d_mse, d_a, d_b, 1, n, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long p[16]; cudaMemcpyFromSymbol(p, g_sort_probe, sizeof(p));
    printf("n=%d  %.1f us  chunk-sort cycles of CTA 0: build %lld setup %lld match %lld prefix %lld scatter %lld writeback %lld  err=%s\n", n, ms * 1e3,
           p[0], p[1], p[2], p[3], p[4], p[5], cudaGetErrorString(cudaGetLastError()));
  }
  std::vector<unsigned long long> out(n);
  cudaMemcpy(out.data(), d_a, n * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 1; i < n; ++i) bad += out[i - 1] > out[i];
  printf("order violations: %d\n", bad);
  return 0;
}
