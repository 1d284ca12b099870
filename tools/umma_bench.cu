// Microbenchmark: cost of back-to-back tcgen05.mma (kind::f16, bf16, cta_group::1) as a function of the N extent
// and of the swizzle mode / K-chunk layout.  One CTA per SM (148), one issuing thread each; operands are whatever is
// in shared memory (values do not matter), the accumulator lives in TMEM.  Prints cycles per MMA.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../video-anomaly-detection_b200/csrc \
//        -I../include -o umma_bench umma_bench.cu && ./umma_bench
#include <cstdio>
#include <cstdlib>

#include "vad_ptx.cuh"

using namespace vad;
__device__ __forceinline__ int distinct_a_c(int d) { return d >= 9 ? 9 : (d >= 3 ? 3 : 1); }

template <int M, int N, int ROWB, int DA, int DD, int SBOR = 8>  // ROWB = bytes per operand row (128: SWIZZLE_128B, 64: SWIZZLE_64B)
__global__ void __launch_bounds__(128, 1) umma_issue_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc<512>(&tmem_slot); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16_f32(M, N);
    constexpr uint32_t layout = ROWB == 128 ? 2u : 4u;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);  // A: <=18 KB, B: <=32 KB
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      // fully unrolled groups of 18 MMAs with compile-time descriptor offsets (9 shifted A rows x 2 K steps, like the
      // halo kernel): no integer work between issues, so the loop measures the tensor pipe and not the issuing thread
      const uint64_t da0 = umma_smem_desc(a0, SBOR * ROWB, layout);
      const uint64_t db0 = umma_smem_desc(b0, 8 * ROWB, layout);
      for (int i = 0; i < iters; i += 18) {
#pragma unroll
        for (int j = 0; j < 18; ++j) {
                    umma_bf16(tmem + (DD > 1 ? static_cast<uint32_t>((j % DD) * 128) : 0u), da0 + static_cast<uint64_t>((DA == 1 ? 0 : (DA == 3 ? (j >> 1) / 3 : (j >> 1))) * (ROWB >> 4) + (j & 1) * 2),
                    db0 + static_cast<uint64_t>((j & 1) * 2), idesc, (i > 0 || j >= DD) ? 1u : 0u);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (elect_one()) {
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int M, int N, int ROWB, int DA = 9, int DD = 1, int SBOR = 8>
void run(int iters) {
  const int distinct_a = DA, distinct_d = DD;
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 96 * 1024 + 1024;
  cudaFuncSetAttribute(umma_issue_kernel<M, N, ROWB, DA, DD, SBOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_issue_kernel<M, N, ROWB, DA, DD, SBOR><<<148, 128, smem>>>(iters, d);  // warm-up
  umma_issue_kernel<M, N, ROWB, DA, DD, SBOR><<<148, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  double avg = 0;
  for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; avg += h[i]; }
  avg /= 148;
  printf("SBO=%2d rows M=%3d N=%3d rowbytes=%3d distinctA=%2d distinctD=%d: %7.1f cycles/MMA (max over SMs %7.1f)  floor %5.1f  %s\n", SBOR, M, N, ROWB,
         distinct_a, distinct_d, avg / iters, double(mx) / iters, M * N / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int iters = 4608;
  run<128, 16, 64>(iters);
  run<128, 32, 64>(iters);
  run<128, 64, 64>(iters);
  run<128, 64, 128>(iters);
  run<128, 16, 64, 9, 1, 10>(iters);
  run<128, 32, 64, 9, 1, 10>(iters);
  run<128, 64, 64, 9, 1, 10>(iters);
  run<128, 64, 128, 9, 1, 10>(iters);
  run<128, 32, 64, 9, 1, 18>(iters);
  run<128, 64, 128, 9, 1, 18>(iters);
  run<128, 32, 64, 9, 1, 16>(iters);
  run<128, 32, 64, 9, 1, 12>(iters);
  return 0;
}
