// tools/hash_bench.cu — microbenchmark of k_hash variants (development aid, not part of the product or of the tests).
// Build: nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -o tools/hash_bench tools/hash_bench.cu oracle/kr_oracle.c -lpthread
// Run on the GPU box: tools/hash_bench [n_messages] ; prints the mean time of every variant with the L2 flushed between launches
// and checks each variant's output against the CPU SHA-1 of the oracle.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../kuberay_b200/csrc/kr_kernels.cuh"
extern "C" void kr_oracle_hash32(const uint8_t *msg, uint64_t len, char out32[32]);

using namespace kr;

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <typename F>
float time_it(F launch, void *flush, size_t flush_bytes, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float tot = 0;
  for (int i = 0; i < iters + 3; i++) {
    CKC(cudaMemsetAsync(flush, i, flush_bytes));
    CKC(cudaEventRecord(a));
    launch();
    CKC(cudaEventRecord(b));
    CKC(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (i >= 3) tot += ms;
  }
  return tot / iters * 1000.f;  // us
}

int main(int argc, char **argv) {
  uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 10000;
  const uint32_t sizes[4] = {1536 - 40, 2560 - 30, 4096 - 20, 6144 - 10};
  std::vector<uint64_t> off(n + 1);
  std::vector<uint32_t> len(n);
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; i++) { len[i] = sizes[i % 4] + (i * 7) % 13; off[i] = total; total += (len[i] + 15) & ~15u; }
  off[n] = total;
  std::vector<uint8_t> bytes(total + 64);
  uint64_t x = 88172645463325252ull;
  for (auto &b : bytes) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; b = (uint8_t)x; }
  std::vector<char> want(32 * (size_t)n), got(32 * (size_t)n);
  for (uint32_t i = 0; i < n; i++) kr_oracle_hash32(bytes.data() + off[i], len[i], &want[32 * (size_t)i]);

  // message ids by descending block count (what the engine's commit builds from c_json_len)
  std::vector<uint32_t> order(n);
  for (uint32_t i = 0; i < n; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return (len[a] + 8) / 64 > (len[b] + 8) / 64; });
  uint32_t *d_order;
  CKC(cudaMalloc(&d_order, 4 * (size_t)n));
  CKC(cudaMemcpy(d_order, order.data(), 4 * (size_t)n, cudaMemcpyHostToDevice));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t ngroups = (n + 31) / 32;
  uint8_t *d_bytes; uint64_t *d_off; uint32_t *d_len; char *d_out; void *flush;
  size_t flush_bytes = 512ull << 20;
  CKC(cudaMalloc(&d_bytes, bytes.size())); CKC(cudaMalloc(&d_off, 8 * (n + 1))); CKC(cudaMalloc(&d_len, 4 * n)); CKC(cudaMalloc(&d_out, 32 * (size_t)n));
  CKC(cudaMalloc(&flush, flush_bytes));
  CKC(cudaMemcpy(d_bytes, bytes.data(), bytes.size(), cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(d_off, off.data(), 8 * (n + 1), cudaMemcpyHostToDevice));
  CKC(cudaMemcpy(d_len, len.data(), 4 * n, cudaMemcpyHostToDevice));

  auto check = [&](const char *name, float us) {
    CKC(cudaMemcpy(got.data(), d_out, 32 * (size_t)n, cudaMemcpyDeviceToHost));
    bool ok = memcmp(got.data(), want.data(), 32 * (size_t)n) == 0;
    printf("%-28s %9.2f us  %7.1f GB/s  %s\n", name, us, (double)total / us / 1e3, ok ? "OK" : "MISMATCH");
    CKC(cudaMemset(d_out, 0, 32 * (size_t)n));
  };
  printf("n=%u messages, %.1f MB\n", n, total / 1e6);
  CKC(cudaFuncSetAttribute(k_hash3<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)sizeof(H3Smem)));
#define RUN(NAME, ...) { float us = time_it([&]() { __VA_ARGS__; }, flush, flush_bytes, 20); CKC(cudaGetLastError()); check(NAME, us); }
  for (int per_sm : {1, 2, 4, 6}) {
    const uint32_t G = std::min<uint32_t>(ngroups, (uint32_t)(sms * per_sm));
    char nm[64];
    snprintf(nm, sizeof nm, "k_hash3<1> sorted %d CTA/SM", per_sm);
    RUN(nm, (k_hash3<1, 0><<<G, 64, sizeof(H3Smem)>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
    snprintf(nm, sizeof nm, "k_hash3<1> unsorted %d CTA/SM", per_sm);
    RUN(nm, (k_hash3<1, 0><<<G, 64, sizeof(H3Smem)>>>(d_bytes, d_off, d_len, nullptr, n, d_out)));
  }
  for (int per_sm : {1, 2, 3}) {
    const uint32_t G = std::min<uint32_t>((ngroups + 1) / 2, (uint32_t)(sms * per_sm));
    char nm[64];
    snprintf(nm, sizeof nm, "k_hash3<2> sorted %d CTA/SM", per_sm);
    RUN(nm, (k_hash3<2, 0><<<G, 128, 2 * sizeof(H3Smem)>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
  }
  {  // each side alone (results are wrong by construction): which warp bounds a block?
    const uint32_t G = std::min<uint32_t>(ngroups, (uint32_t)sms);
    RUN("k_hash3<1> 1/SM producer only", (k_hash3<1, 1><<<G, 64, sizeof(H3Smem)>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
    RUN("k_hash3<1> 1/SM consumer only", (k_hash3<1, 2><<<G, 64, sizeof(H3Smem)>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
    RUN("k_hash3<1> 1/SM both", (k_hash3<1, 0><<<G, 64, sizeof(H3Smem)>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
    RUN("k_hash3<1> 1/SM hand-offs only", (k_hash3<1, 3><<<G, 64, sizeof(H3Smem)>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
    RUN("k_hash3<1> 1/SM no hand-offs", (k_hash3<1, 4><<<G, 64, sizeof(H3Smem)>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
    const uint32_t G2 = std::min<uint32_t>(ngroups, (uint32_t)sms * 2);
    RUN("k_hash3<1> 2/SM hand-offs only", (k_hash3<1, 3><<<G2, 64, sizeof(H3Smem)>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
    RUN("k_hash3<1> 2/SM no hand-offs", (k_hash3<1, 4><<<G2, 64, sizeof(H3Smem)>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
    RUN("k_hash3<1> 2/SM producer only", (k_hash3<1, 1><<<G2, 64, sizeof(H3Smem)>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
    RUN("k_hash3<1> 2/SM consumer only", (k_hash3<1, 2><<<G2, 64, sizeof(H3Smem)>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
  }
  RUN("k_hash2<1,0> sorted", (k_hash2<1, 0><<<(n + 31) / 32, 32>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
  RUN("k_hash2<4,1> sorted", (k_hash2<4, 1><<<(n + 127) / 128, 128>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
  RUN("k_hash2<4,1> sorted 2CTA/SM", (k_hash2<4, 1><<<std::min<uint32_t>((n + 127) / 128, sms * 2), 128>>>(d_bytes, d_off, d_len, d_order, n, d_out)));
  RUN("k_hash2<1,0> cpasync", (k_hash2<1, 0><<<(n + 31) / 32, 32>>>(d_bytes, d_off, d_len, nullptr, n, d_out)));
  RUN("k_hash2<1,1> cpasync+mad", (k_hash2<1, 1><<<(n + 31) / 32, 32>>>(d_bytes, d_off, d_len, nullptr, n, d_out)));
  RUN("k_hash2<1,5> fma: w+K", (k_hash2<1, 5><<<(n + 31) / 32, 32>>>(d_bytes, d_off, d_len, nullptr, n, d_out)));
  RUN("k_hash2<1,2> fma: w+K rol30", (k_hash2<1, 2><<<(n + 31) / 32, 32>>>(d_bytes, d_off, d_len, nullptr, n, d_out)));
  RUN("k_hash2<1,3> fma: w+K rol1", (k_hash2<1, 3><<<(n + 31) / 32, 32>>>(d_bytes, d_off, d_len, nullptr, n, d_out)));
  RUN("k_hash2<1,4> fma: all three", (k_hash2<1, 4><<<(n + 31) / 32, 32>>>(d_bytes, d_off, d_len, nullptr, n, d_out)));
  RUN("k_hash2<4,2> fma: w+K rol30", (k_hash2<4, 2><<<(n + 127) / 128, 128>>>(d_bytes, d_off, d_len, nullptr, n, d_out)));
  RUN("k_hash2<4,4> fma: all three", (k_hash2<4, 4><<<(n + 127) / 128, 128>>>(d_bytes, d_off, d_len, nullptr, n, d_out)));
  RUN("k_hash2<4,0> cpasync", (k_hash2<4, 0><<<(n + 127) / 128, 128>>>(d_bytes, d_off, d_len, nullptr, n, d_out)));
  RUN("k_hash2<4,1> cpasync+mad", (k_hash2<4, 1><<<(n + 127) / 128, 128>>>(d_bytes, d_off, d_len, nullptr, n, d_out)));
  return 0;
}
