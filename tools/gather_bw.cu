// Micro-benchmark (evidence, not product): HBM bandwidth of RANDOM fixed-size chunk reads as a function of chunk size
// and bytes in flight per SM — the access pattern of the candidate gather (stage 2 of the batched three-stage search:
// 256k pooled pages of ~6 KB each, picked at random from a 6 GB store). One thread per CTA keeps `stages` bulk copies
// (cp.async.bulk global -> shared, mbarrier completion) in flight and re-issues as they land; nothing is computed.
// Gives the roofline of the pattern: what a perfect gather kernel could reach.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_bw.bin tools/gather_bw.cu && tools/gather_bw.bin
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Chunk starts come from an in-register LCG (random) or arithmetic (sequential): no dependent global load on the issue path.
__global__ void gather_bw_kernel(const uint8_t* __restrict__ buf, unsigned long long slots, int sequential, long long n_chunks,
                                 int chunk_bytes, int stages) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[64];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  long long it = 0;
  const int lg = 31 - __clz(stages);
  unsigned long long x = 0x9E3779B97F4A7C15ull * (blockIdx.x + 1);
  for (long long i = blockIdx.x; i < n_chunks; i += gridDim.x, ++it) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    // slots and stages are powers of two: no integer division on the issue path
    const unsigned long long start = (sequential ? static_cast<unsigned long long>(i) * (chunk_bytes >> 8) : (x >> 20)) & (slots - 1);
    const int s = static_cast<int>(it & (stages - 1));
    if (it >= stages) {
      const uint32_t parity = static_cast<uint32_t>(((it >> lg) - 1) & 1);
      while (!mbar_try(&full[s], parity)) {}
    }
    bulk_load(smem + static_cast<size_t>(s) * chunk_bytes, buf + start * 256, chunk_bytes, &full[s]);
    mbar_expect(&full[s], chunk_bytes);
  }
  for (long long j = (it > stages ? it - stages : 0); j < it; ++j) {   // drain
    const int s = static_cast<int>(j & (stages - 1));
    while (!mbar_try(&full[s], static_cast<uint32_t>((j >> lg) & 1))) {}
  }
}

int main(int argc, char** argv) {
  const size_t buf_bytes = (argc > 1 ? atoll(argv[1]) : 8ll) << 30;
  uint8_t* buf;
  if (cudaMalloc(&buf, buf_bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(buf, 1, buf_bytes);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  cudaFuncSetAttribute(gather_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const long long total_bytes = 3ll << 30;   // bytes moved per run
  printf("{\"gather_bw\": [\n");
  bool first = true;
  for (int seq = 0; seq < 2; ++seq) {          // 0: random chunk starts, 1: consecutive chunks (the streaming reference)
    for (int chunk : {2048, 4096, 6144, 8192, 16384, 32768}) {
      const long long n = total_bytes / chunk;
      unsigned long long slots = 1;
      while (slots * 2 * 256 + chunk <= buf_bytes) slots *= 2;
      for (int inflight_kb : {32, 64, 128, 192}) {
        int stages = 1;
        while (stages * 2 * chunk <= inflight_kb * 1024) stages *= 2;
        if (stages * chunk > 200 * 1024 || stages > 64) continue;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          gather_bw_kernel<<<sms, 32, static_cast<size_t>(stages) * chunk>>>(buf, slots, seq, n, chunk, stages);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        printf("%s {\"pattern\": \"%s\", \"chunk_bytes\": %d, \"inflight_kb_per_sm\": %d, \"stages\": %d, \"ms\": %.4f, \"gbs\": %.1f, \"err\": \"%s\"}",
               first ? "" : ",\n", seq ? "sequential" : "random", chunk, inflight_kb, stages, best, n * (double)chunk / best / 1e6,
               err == cudaSuccess ? "" : cudaGetErrorString(err));
        first = false;
      }
    }
  }
  printf("\n]}\n");
  return 0;
}
