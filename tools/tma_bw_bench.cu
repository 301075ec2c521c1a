// Micro-benchmark: per-SM global->shared throughput of 1-D bulk copies (cp.async.bulk, what the GEMM producer
// uses) vs 2-D/3-D TMA tensor loads over the chunk8 layout, one CTA per SM, ring of S stages.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void wait(uint64_t* b, uint32_t par) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(dst)),
               "l"((unsigned long long)m), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
}

// operand: chunk8 buffer [nchunks][rcap][8] bf16; a "stage" = A tile (128 rows x 8 chunks = 16 KB) + B tile (256 rows x 8 chunks = 32 KB)
template <int MODE>  // 0: 1-D bulk copies from one thread; 1: 1-D bulk copies from 3 warps; 2: TMA tensor (2 KB inner boxes)
                     // 3: dX pattern (B operand MN-major: 32 copies of 1 KB) from 3 warps; 4: dX pattern, TMA tensor (1 KB inner box x 32 chunks)
__global__ void __launch_bounds__(128) k(const uint8_t* base, const CUtensorMap* tm, const CUtensorMap* tm2, int rcap, int nchunks, int iters, int S) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8];
  const int stage_bytes = 48 * 1024;
  if (threadIdx.x == 0) for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 384) % (rcap - 384);
  // consumer = thread 96 waits; producers issue. No compute: we measure the copy path only (stage reuse right after landing).
  for (int it = 0; it < iters + S; ++it) {
    const int s = it % S;
    if (it >= S) { if (threadIdx.x == 96) wait(&full[s], ((it - S) / S) & 1); }
    __syncthreads();
    if (it < iters) {
      const int kb = it % (nchunks / 8);
      uint8_t* dst = smem + (size_t)s * stage_bytes;
      if (MODE == 0) {
        if (threadIdx.x == 0) {
          expect(&full[s], stage_bytes);
          for (int c = 0; c < 8; ++c) bulk(dst + c * 2048, base + ((size_t)(kb * 8 + c) * rcap + row0) * 16, 2048, &full[s]);
          for (int c = 0; c < 8; ++c) bulk(dst + 16384 + c * 4096, base + ((size_t)(kb * 8 + c) * rcap + row0 + 128) * 16, 4096, &full[s]);
        }
      } else if (MODE == 1) {
        if (threadIdx.x == 0) expect(&full[s], stage_bytes);
        __syncwarp();
        if (warp < 3) for (int c = warp + 3 * lane; c < 16; c += 96) {
          if (c < 8) bulk(dst + c * 2048, base + ((size_t)(kb * 8 + c) * rcap + row0) * 16, 2048, &full[s]);
          else bulk(dst + 16384 + (c - 8) * 4096, base + ((size_t)(kb * 8 + c - 8) * rcap + row0 + 128) * 16, 4096, &full[s]);
        }
      } else if (MODE == 3) {
        if (threadIdx.x == 0) expect(&full[s], stage_bytes);
        __syncwarp();
        const int r64 = row0 + (it % 4) * 64;
        if (warp < 3) for (int c = warp + 3 * lane; c < 40; c += 96) {
          if (c < 8) bulk(dst + c * 2048, base + ((size_t)(kb * 8 + c) * rcap + row0) * 16, 2048, &full[s]);
          else bulk(dst + 16384 + (c - 8) * 1024, base + ((size_t)(c - 8) * rcap + r64) * 16, 1024, &full[s]);
        }
      } else if (MODE == 4) {
        if (threadIdx.x == 0) {
          const int r64 = row0 + (it % 4) * 64;
          expect(&full[s], stage_bytes);
          tma3(dst, tm, row0 * 2, kb * 8, 0, &full[s]);   // A: 128 rows x 8 chunks
          tma3(dst + 16384, tm2, r64 * 2, 0, 0, &full[s]);  // B: 64 contraction rows x 32 feature chunks
        }
      } else {
        if (threadIdx.x == 0) {
          expect(&full[s], stage_bytes);
          tma3(dst, tm, row0 * 2, kb * 8, 0, &full[s]);               // 128 rows x 8 chunks
          tma3(dst + 16384, tm, (row0 + 128) * 2, kb * 8, 0, &full[s]);  // 2 x (128 rows x 8 chunks)
          tma3(dst + 32768, tm, (row0 + 256) * 2, kb * 8, 0, &full[s]);
        }
      }
    }
  }
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rcap = 148 * 384 + 512, nchunks = 128;  // 8 GB? no: 57k rows x 128 chunks x 16 B = 117 MB (fits L2 mostly)
  size_t bytes = (size_t)rcap * nchunks * 16;
  uint8_t* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)rcap * 2, (cuuint64_t)nchunks, 1};
  cuuint64_t strides[2] = {(cuuint64_t)rcap * 16, (cuuint64_t)rcap * 16 * nchunks};
  cuuint32_t box[3] = {256, 8, 1}, es[3] = {1, 1, 1};
  CUresult r = ((PFN)fp)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  CUtensorMap* dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
  CUtensorMap m2;
  cuuint32_t box2[3] = {128, 32, 1};
  r = ((PFN)fp)(&m2, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, buf, dims, strides, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode2: %d\n", (int)r);
  CUtensorMap* dm2; cudaMalloc(&dm2, sizeof(m2)); cudaMemcpy(dm2, &m2, sizeof(m2), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  auto run = [&](const char* name, auto kern, int S, int grid) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    kern<<<grid, 128, S * 48 * 1024>>>(buf, dm, dm2, rcap, nchunks, 200, S);
    cudaEventRecord(e0);
    kern<<<grid, 128, S * 48 * 1024>>>(buf, dm, dm2, rcap, nchunks, iters, S);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    double per_sm = 48.0 * 1024 * iters / (ms * 1e-3) / 1e9;
    printf("%-40s S=%d grid=%3d  %8.1f us  %6.1f GB/s per SM  %7.0f GB/s total  %.2f us per 48 KB stage  %s\n", name, S, grid, ms * 1e3, per_sm, per_sm * grid,
           ms * 1e3 / iters, err == cudaSuccess ? "" : cudaGetErrorString(err));
  };
  for (int grid : {148, 32}) for (int S : {2, 4}) {
    run("1-D bulk copies, one thread", k<0>, S, grid);
    run("1-D bulk copies, 3 warps", k<1>, S, grid);
    run("TMA tensor 3-D (2 KB inner box)", k<2>, S, grid);
    run("dX pattern: 1-D bulk 8x2KB+32x1KB, 3 warps", k<3>, S, grid);
    run("dX pattern: TMA tensor 2 instr", k<4>, S, grid);
  }
  return 0;
}
