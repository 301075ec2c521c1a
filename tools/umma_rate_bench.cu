// Micro-benchmark: issue rate of tcgen05.mma (M=128, bf16, fp32 accumulate) from shared memory in the no-swizzle
// canonical layouts the GEMM kernel uses (K-major / MN-major operands), without any TMA traffic or epilogue.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I drvae_b200/csrc tools/umma_rate_bench.cu -o /tmp/umma
// Prints cycles per MMA instruction; the hardware floor is 128 * N / 256 cycles (N=256 -> 128).
#include <stdio.h>

#include "common.cuh"

using namespace drvae;

// mode 0: A, B K-major (forward);  1: A K-major, B MN-major (dX);  2: A, B MN-major (dW)
// walk: 0 = every MMA reads the same operand slabs, 1 = walk over 4 k-steps x `stages` stages like the mainloop
__global__ void __launch_bounds__(128, 1) k(int mode, int BN, int iters, int stages, int walk, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const int stage_bytes = 16384 + BN * 128;
  for (int i = threadIdx.x; i < stages * stage_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0) {
    const bool a_mn = mode == 2, b_mn = mode != 0;
    const uint32_t idesc = umma_idesc_bf16(BN, a_mn, b_mn);
    uint32_t a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step;
    if (!a_mn) { a_lbo = 128 * 16; a_sbo = 128; a_step = 2 * 128 * 16; } else { a_lbo = 128; a_sbo = 64 * 16; a_step = 256; }
    if (!b_mn) { b_lbo = BN * 16; b_sbo = 128; b_step = 2 * BN * 16; } else { b_lbo = 128; b_sbo = 64 * 16; b_step = 256; }
    const uint64_t a0 = umma_smem_desc(0, a_lbo, a_sbo), b0 = umma_smem_desc(0, b_lbo, b_sbo);
    const uint32_t base = smem_u32(smem);
    if (walk == 2) {
      // tightest possible issue loop: one elected lane, 8 MMAs unrolled with constant descriptor increments
      const uint32_t As = base, Bs = base + 16384;
      const uint64_t ad = a0 | (uint64_t)((As >> 4) & 0x3FFFu), bd = b0 | (uint64_t)((Bs >> 4) & 0x3FFFu);
      const uint64_t da = a_step >> 4, db = b_step >> 4;
      const long long t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base + (u & 1) * 256),
                "l"(ad + (u & 3) * da), "l"(bd + (u & 3) * db), "r"(idesc), "r"(1)
                : "memory");
          }
        }
      }
      __syncwarp();
      umma_commit(&bar);
      mbar_wait(&bar, 0, nullptr, 0);
      const long long t1 = clock64();
      if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    } else {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int s = walk ? (i >> 2) % stages : 0, q = walk ? (i & 3) : 0;
      const uint32_t As = base + s * stage_bytes, Bs = As + 16384;
      const uint64_t ad = a0 | (uint64_t)(((As + q * a_step) >> 4) & 0x3FFFu);
      const uint64_t bd = b0 | (uint64_t)(((Bs + q * b_step) >> 4) & 0x3FFFu);
      umma_bf16(tmem_base + (walk ? ((i >> 6) & 1) * 256 : 0), ad, bd, idesc, i > 0);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0, nullptr, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[3] = {"NT (A,B K-major)", "DX (A K-major, B MN-major)", "DW (A,B MN-major)"};
  for (int grid : {1})
    for (int mode = 0; mode < 3; ++mode)
      for (int BN : {64, 128, 208, 256})
        for (int walk = 0; walk < 3; ++walk) {
          const int iters = 4096, stages = 4;
          k<<<grid, 128, stages * (16384 + BN * 128)>>>(mode, BN, iters, stages, walk, out);
          cudaError_t err = cudaDeviceSynchronize();
          long long h[148];
          cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("grid %3d  %-28s N=%3d walk=%d: %7.1f cycles / MMA (floor %5.1f)  %s\n", grid, names[mode], BN, walk, (double)mx / iters,
                 128.0 * BN / 256.0, err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
  return 0;
}
