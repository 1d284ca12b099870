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

template <int N, int ROWB>  // ROWB = bytes per operand row (128: SWIZZLE_128B, 64: SWIZZLE_64B)
__global__ void __launch_bounds__(128, 1) umma_issue_kernel(int iters, int distinct_a, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc<256>(&tmem_slot); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, N);
    constexpr uint32_t layout = ROWB == 128 ? 2u : 4u;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);  // A: <=18 KB, B: <=32 KB
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        // distinct_a different A tiles (like the taps of a conv) so the operand really streams from smem
        const uint32_t a_addr = a0 + (i % distinct_a) * ROWB;  // shifted-row descriptors, as the halo kernel does
        const uint64_t da = umma_smem_desc(a_addr, 8 * ROWB, layout);
        const uint64_t db = umma_smem_desc(b0, 8 * ROWB, layout);
        umma_bf16(tmem, da, db, idesc, i > 0);
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
  if (warp == 1) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

template <int N, int ROWB>
void run(int iters, int distinct_a) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 96 * 1024 + 1024;
  cudaFuncSetAttribute(umma_issue_kernel<N, ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_issue_kernel<N, ROWB><<<148, 128, smem>>>(iters, distinct_a, d);  // warm-up
  umma_issue_kernel<N, ROWB><<<148, 128, smem>>>(iters, distinct_a, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  double avg = 0;
  for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; avg += h[i]; }
  avg /= 148;
  printf("M=128 N=%3d rowbytes=%3d distinctA=%2d: %7.1f cycles/MMA (max over SMs %7.1f)  floor %5.1f  %s\n", N, ROWB,
         distinct_a, avg / iters, double(mx) / iters, 128.0 * N / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int iters = 4096;
  for (int da : {1, 9}) {
    run<16, 64>(iters, da);
    run<32, 64>(iters, da);
    run<64, 64>(iters, da);
    run<32, 128>(iters, da);
    run<64, 128>(iters, da);
    run<128, 128>(iters, da);
    run<256, 128>(iters, da);
  }
  return 0;
}
