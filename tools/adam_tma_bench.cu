// Micro-benchmark (round 2): the Adam read-modify-write of p, m, v (fp32, reference [rows][ld] layout) + bf16 chunk8
// shadow, with the optimizer state moved by the TMA unit instead of per-thread loads/stores.
//
// Round 1 left the fused dW+Adam epilogue at the ceiling of its per-thread access pattern (4.7-4.9 TB/s physical in the
// persistent one-CTA-per-SM geometry vs 6.3 TB/s for a linear stream, tools/adam_pattern_bench.cu): memory-level
// parallelism was bounded by resident warps x registers.  Here a tile's state is staged through shared memory by 3-D
// tensor loads ({128 k, R weight rows, 1 model} boxes of 512-byte row pieces: 3 instructions per stage) and written back
// by tensor stores, so the bytes in flight per SM no longer depend on the epilogue warps; the warps only run the
// update on shared memory (lane = input feature k, conflict-free).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/adam_tma_bench.cu -o gpurun_out/adam_tma_bench
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct H { float lr, b1, b2, eps, wd, isb; };
__device__ __forceinline__ void upd(float g, float& p, float& m, float& v, const H& h) {
  float gr = g + h.wd * p;
  m = h.b1 * m + (1.f - h.b1) * gr;
  v = h.b2 * v + (1.f - h.b2) * gr * gr;
  float s;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(s) : "f"(v));
  p = p - h.lr * __fdividef(m, s * h.isb + h.eps);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    __nanosleep(32);
    if (clock64() - t0 > 2000000000LL) {
      printf("mbar timeout: block %d thread %d tag %d parity %u\n", blockIdx.x, threadIdx.x, tag, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst)), "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((unsigned long long)map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// reference point: linear float4 stream
__global__ void k_linear(float4* P, float4* M, float4* V, __nv_bfloat162* S, long long n4, H h) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 p = P[i], m = M[i], v = V[i];
    upd(1e-3f, p.x, m.x, v.x, h); upd(1e-3f, p.y, m.y, v.y, h); upd(1e-3f, p.z, m.z, v.z, h); upd(1e-3f, p.w, m.w, v.w, h);
    P[i] = p; M[i] = m; V[i] = v;
    S[2 * i] = __floats2bfloat162_rn(p.x, p.y); S[2 * i + 1] = __floats2bfloat162_rn(p.z, p.w);
  }
}

// round-1 epilogue pattern (16-byte accesses after a quad transpose, persistent geometry): the number to beat
template <int EW>
__global__ void __launch_bounds__(EW * 32 + 128, 1) k_tile_v4(float* P, float* M, float* V, __nv_bfloat16* S, int rows, int ld, int BN, int tiles_m,
                                                              int tiles_n, long long ms, long long sms, H h, int total_tiles) {
  extern __shared__ float sm[];
  if (sm[0] == 123.f) return;
  const int warp = (threadIdx.x >> 5) - 4, lane = threadIdx.x & 31;
  if (warp < 0) return;
  const int cg = warp >> 2, ngroups = EW / 4, q = warp & 3;
  const int ci = lane & 3, kg = lane >> 2;
  const int nhc = BN >> 3;
  struct Buf { float4 p[2], m[2], v[2]; };
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int per = tiles_m * tiles_n;
    const int model = tile / per, mn = tile % per;
    const int tile_m = mn % tiles_m, tile_n = mn / tiles_m;
    const int k4 = tile_m * 128 + q * 32 + kg * 4;
    const bool kok = k4 + 3 < ld;
    float* p = P + model * ms + k4; float* m = M + model * ms + k4; float* v = V + model * ms + k4;
    __nv_bfloat16* s = S + model * sms;
    const int rcap = tiles_n * BN;
    for (int hc = cg; hc < nhc; hc += ngroups) {
      Buf b;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = tile_n * BN + hc * 8 + j * 4 + ci;
        b.p[j] = b.m[j] = b.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < rows && kok) {
          const long long idx = (long long)n * ld;
          b.p[j] = *reinterpret_cast<const float4*>(p + idx); b.m[j] = *reinterpret_cast<const float4*>(m + idx); b.v[j] = *reinterpret_cast<const float4*>(v + idx);
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        upd(1e-3f, b.p[j].x, b.m[j].x, b.v[j].x, h); upd(1e-3f, b.p[j].y, b.m[j].y, b.v[j].y, h);
        upd(1e-3f, b.p[j].z, b.m[j].z, b.v[j].z, h); upd(1e-3f, b.p[j].w, b.m[j].w, b.v[j].w, h);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = tile_n * BN + hc * 8 + j * 4 + ci;
        if (n < rows && kok) {
          const long long idx = (long long)n * ld;
          *reinterpret_cast<float4*>(p + idx) = b.p[j]; *reinterpret_cast<float4*>(m + idx) = b.m[j]; *reinterpret_cast<float4*>(v + idx) = b.v[j];
          __nv_bfloat162 lo = __floats2bfloat162_rn(b.p[j].x, b.p[j].y), hi = __floats2bfloat162_rn(b.p[j].z, b.p[j].w);
          uint2 pk = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
          *reinterpret_cast<uint2*>(s + ((long long)(k4 >> 3) * rcap + n) * 8 + (k4 & 7)) = pk;
        }
      }
    }
  }
}

// TMA-staged: warp 0 = loader, warp 1 = storer, warps 4.. = EW update warps (k quarter w&3, row group w>>2).
// Stage = R weight rows x 128 k of p, m, v (3 x R x 512 B) [+ R x 256 B of bf16 shadow when SHMODE == 2].
// SHMODE: 0 no shadow, 1 scattered 2-byte global stores (what the round-1 scalar epilogue does), 2 staged + tensor store.
// Barriers per stage: full (tx), done (EW arrivals: updated and fenced), empty (storer: shared memory read by the TMA unit).
template <int EW, int R, int NST, int SHMODE, int LAG>
__global__ void __launch_bounds__(EW * 32 + 128, 1) k_tile_tma(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmM,
                                                               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmS,
                                                               __nv_bfloat16* S, int rows, int ld, int BN, int tiles_m, int tiles_n, long long sms,
                                                               H h, int total_tiles, int reserve_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[NST], done_bar[NST], empty_bar[NST];
  constexpr int ARR = R * 128 * 4;                        // one array of one stage
  constexpr int SHB = (SHMODE == 2) ? R * 256 : 0;        // bf16 shadow of the stage: [16 chunks][R][8]
  constexpr int STAGE = 3 * ARR + SHB;
  uint8_t* ring = smem + reserve_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&done_bar[s], EW);
      mbar_init(&empty_bar[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int subs = BN / R;  // stages per tile
  const int rcap = tiles_n * BN;
  if (warp == 0) {
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int per = tiles_m * tiles_n;
      const int model = tile / per, mn = tile % per;
      const int tile_m = mn % tiles_m, tile_n = mn / tiles_m;
      for (int sub = 0; sub < subs; ++sub, ++it) {
        const int s = it % NST;
        const uint32_t ph = (it / NST) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 1);
        const int n0 = tile_n * BN + sub * R;
        if (n0 >= rows) {  // nothing to load: complete the phase with a plain arrival
          if (elect_one()) mbar_arrive(&full_bar[s]);
        } else if (elect_one()) {
          uint8_t* st = ring + (size_t)s * STAGE;
          mbar_expect_tx(&full_bar[s], 3 * ARR);
          tma_load_3d(st, &tmP, tile_m * 128, n0, model, &full_bar[s]);
          tma_load_3d(st + ARR, &tmM, tile_m * 128, n0, model, &full_bar[s]);
          tma_load_3d(st + 2 * ARR, &tmV, tile_m * 128, n0, model, &full_bar[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int per = tiles_m * tiles_n;
      const int model = tile / per, mn = tile % per;
      const int tile_m = mn % tiles_m, tile_n = mn / tiles_m;
      for (int sub = 0; sub < subs; ++sub, ++it) {
        const int s = it % NST;
        const uint32_t ph = (it / NST) & 1;
        mbar_wait(&done_bar[s], ph, 2);
        const int n0 = tile_n * BN + sub * R;
        if (elect_one()) {
          if (n0 < rows) {
            uint8_t* st = ring + (size_t)s * STAGE;
            tma_store_3d(&tmP, st, tile_m * 128, n0, model);
            tma_store_3d(&tmM, st + ARR, tile_m * 128, n0, model);
            tma_store_3d(&tmV, st + 2 * ARR, tile_m * 128, n0, model);
            if (SHMODE == 2) tma_store_3d(&tmS, st + 3 * ARR, n0 * 2, tile_m * 16, model);
          }
          tma_commit();
          if (LAG == 0) {
            tma_wait_read<0>();  // the unit has read this stage's shared memory
            mbar_arrive(&empty_bar[s]);
          } else if (it > 0) {
            tma_wait_read<1>();  // ... the previous stage's (one store group stays in flight)
            mbar_arrive(&empty_bar[(it - 1) % NST]);
          }
        }
        __syncwarp();
      }
    }
    if (elect_one()) {
      if (LAG == 1 && it > 0) {
        tma_wait_read<0>();
        mbar_arrive(&empty_bar[(it - 1) % NST]);
      }
      tma_wait_all();
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int w = warp - 4, q = w & 3, g = w >> 2;
    constexpr int G = EW / 4, RPG = R / G;  // rows per group and stage
    const int kl = q * 32 + lane;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int per = tiles_m * tiles_n;
      const int model = tile / per, mn = tile % per;
      const int tile_m = mn % tiles_m, tile_n = mn / tiles_m;
      const int k = tile_m * 128 + kl;
      for (int sub = 0; sub < subs; ++sub, ++it) {
        const int s = it % NST;
        const uint32_t ph = (it / NST) & 1;
        mbar_wait(&full_bar[s], ph, 3);
        uint8_t* st = ring + (size_t)s * STAGE;
        float* sp = reinterpret_cast<float*>(st);
        float* smm = reinterpret_cast<float*>(st + ARR);
        float* sv = reinterpret_cast<float*>(st + 2 * ARR);
        const int n0 = tile_n * BN + sub * R;
        float pv[RPG], mv[RPG], vv[RPG];
#pragma unroll
        for (int i = 0; i < RPG; ++i) {
          const int r = g * RPG + i;
          pv[i] = sp[r * 128 + kl], mv[i] = smm[r * 128 + kl], vv[i] = sv[r * 128 + kl];
        }
#pragma unroll
        for (int i = 0; i < RPG; ++i) upd(1e-3f, pv[i], mv[i], vv[i], h);
#pragma unroll
        for (int i = 0; i < RPG; ++i) {
          const int r = g * RPG + i;
          sp[r * 128 + kl] = pv[i], smm[r * 128 + kl] = mv[i], sv[r * 128 + kl] = vv[i];
          if (SHMODE == 1) {
            const int n = n0 + r;
            if (n < rows && k < ld) S[model * sms + ((long long)(k >> 3) * rcap + n) * 8 + (k & 7)] = __float2bfloat16_rn(pv[i]);
          } else if (SHMODE == 2) {
            reinterpret_cast<__nv_bfloat16*>(st + 3 * ARR)[((kl >> 3) * R + r) * 8 + (kl & 7)] = __float2bfloat16_rn(pv[i]);
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&done_bar[s]);
      }
    }
  }
}

typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn g_encode;

static CUtensorMap map_f32(float* base, int ld, int rows, int E, int R) {
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)rows, (cuuint64_t)E};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)rows * ld * 4};
  const cuuint32_t box[3] = {128, (cuuint32_t)R, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode f32 map failed: %d\n", (int)r); exit(1); }
  return m;
}
// chunk8 bf16 shadow [E][kc/8][rcap][8] as {rcap x 16 B (2 x 8-byte elements), chunks, models}; box {R rows, 16 chunks, 1}
static CUtensorMap map_c8(__nv_bfloat16* base, int rcap, int nchunks, long long sms, int E, int R) {
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)rcap * 2, (cuuint64_t)nchunks, (cuuint64_t)E};
  const cuuint64_t strides[2] = {(cuuint64_t)rcap * 16, (cuuint64_t)sms * 2};
  const cuuint32_t box[3] = {(cuuint32_t)R * 2, 16, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode c8 map failed: %d\n", (int)r); exit(1); }
  return m;
}

int main(int argc, char** argv) {
  setvbuf(stdout, NULL, _IONBF, 0);
  {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encoder\n"); return 1; }
    g_encode = (TmapEncodeFn)fp;
  }
  const int E = 32;
  struct Shape { const char* name; int rows, ld; };
  // decoder heads ([2 x 978 stacked as 1956] x 600), first encoder layer with rows padded to a 16-byte stride (800 x 980)
  const Shape shapes[2] = {{"dec.head 1956x600", 1956, 600}, {"enc.h0 800x980", 800, 980}};
  H h{5e-4f, 0.9f, 0.999f, 1e-8f, 0.05f, 1.f};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (const Shape& sh : shapes) {
    const int rows = sh.rows, ld = sh.ld, BN = 256;
    const int tiles_m = (ld + 127) / 128, tiles_n = (rows + BN - 1) / BN, rcap = tiles_n * BN;
    const int kc = (ld + 16) / 16 * 16;
    const long long ms = (long long)rows * ld, total = ms * E, sms = (long long)kc * rcap;
    float *P, *M, *V;
    __nv_bfloat16* S;
    cudaMalloc(&P, total * 4); cudaMalloc(&M, total * 4); cudaMalloc(&V, total * 4); cudaMalloc(&S, sms * E * 2);
    cudaMemset(P, 0, total * 4); cudaMemset(M, 0, total * 4); cudaMemset(V, 0, total * 4); cudaMemset(S, 0, sms * E * 2);
    const int total_tiles = tiles_m * tiles_n * E;
    printf("== %s, %d models: %d tiles of 128 x %d ==\n", sh.name, E, total_tiles, BN);
    auto timeit = [&](const char* name, auto launch) {
      for (int i = 0; i < 3; ++i) launch();
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("%-64s FAILED: %s\n", name, cudaGetErrorString(err)); exit(1); }
      cudaEventRecord(e0);
      const int Rn = 10;
      for (int i = 0; i < Rn; ++i) launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms_;
      cudaEventElapsedTime(&ms_, e0, e1);
      ms_ /= Rn;
      err = cudaGetLastError();
      printf("%-64s %8.1f us  %7.0f GB/s (26 B/param)  %s\n", name, ms_ * 1e3, 26.0 * total / (ms_ * 1e-3) / 1e9,
             err == cudaSuccess ? "" : cudaGetErrorString(err));
    };
    timeit("linear float4, 148*32 CTAs x 256", [&] { k_linear<<<148 * 32, 256>>>((float4*)P, (float4*)M, (float4*)V, (__nv_bfloat162*)S, total / 4, h); });
    cudaFuncSetAttribute(k_tile_v4<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    timeit("round-1 pattern: vec4, 24 warps, 100 KB smem", [&] {
      k_tile_v4<24><<<148, 24 * 32 + 128, 100 * 1024>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, sms, h, total_tiles);
    });
    // correctness spot check of the TMA path: after one launch from zeros every element must equal the linear kernel's
#define RUN_TMA(EW, R, NST, SHM, RES_KB, LAG)                                                                                          \
  {                                                                                                                               \
    CUtensorMap tp = map_f32(P, ld, rows, E, R), tm = map_f32(M, ld, rows, E, R), tv = map_f32(V, ld, rows, E, R);                \
    CUtensorMap ts = map_c8(S, rcap, kc / 8, sms, E, R);                                                                          \
    const int stage = 3 * R * 512 + ((SHM) == 2 ? R * 256 : 0);                                                                   \
    const int smem = RES_KB * 1024 + NST * stage;                                                                                 \
    cudaFuncSetAttribute(k_tile_tma<EW, R, NST, SHM, LAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                         \
    char nm[128];                                                                                                                 \
    snprintf(nm, sizeof(nm), "TMA-staged: %d warps, R=%d, %d stages, shadow=%d, lag=%d, +%d KB (%d KB)", EW, R, NST, SHM, LAG, RES_KB, smem / 1024); \
    timeit(nm, [&] {                                                                                                              \
      k_tile_tma<EW, R, NST, SHM, LAG><<<148, EW * 32 + 128, smem>>>(tp, tm, tv, ts, S, rows, ld, BN, tiles_m, tiles_n, sms, h,        \
                                                                total_tiles, RES_KB * 1024);                                      \
    });                                                                                                                           \
  }
    RUN_TMA(16, 16, 3, 0, 96, 0)
    RUN_TMA(16, 16, 4, 0, 96, 0)
    RUN_TMA(16, 16, 4, 0, 96, 1)
    RUN_TMA(16, 32, 2, 0, 96, 0)
    RUN_TMA(16, 32, 3, 0, 0, 1)
    RUN_TMA(16, 32, 4, 0, 0, 1)
    RUN_TMA(8, 16, 4, 0, 96, 1)
    RUN_TMA(8, 32, 2, 0, 96, 0)
    RUN_TMA(16, 16, 4, 1, 96, 1)
    RUN_TMA(16, 16, 4, 2, 96, 0)
    RUN_TMA(16, 16, 4, 2, 96, 1)
    RUN_TMA(16, 32, 2, 2, 96, 0)
    RUN_TMA(8, 16, 4, 2, 96, 1)
    RUN_TMA(16, 16, 6, 2, 48, 1)
    RUN_TMA(16, 8, 8, 2, 96, 1)
    // verification: reset, one linear pass vs one TMA pass (values identical: same update on zeros)
    {
      cudaMemset(P, 0, total * 4); cudaMemset(M, 0, total * 4); cudaMemset(V, 0, total * 4);
      CUtensorMap tp = map_f32(P, ld, rows, E, 16), tm = map_f32(M, ld, rows, E, 16), tv = map_f32(V, ld, rows, E, 16);
      CUtensorMap ts = map_c8(S, rcap, kc / 8, sms, E, 16);
      const int smem = 4 * (3 * 16 * 512 + 16 * 256);
      cudaFuncSetAttribute(k_tile_tma<16, 16, 4, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      k_tile_tma<16, 16, 4, 2, 1><<<148, 16 * 32 + 128, smem>>>(tp, tm, tv, ts, S, rows, ld, BN, tiles_m, tiles_n, sms, h, total_tiles, 0);
      k_tile_tma<16, 16, 4, 2, 1><<<148, 16 * 32 + 128, smem>>>(tp, tm, tv, ts, S, rows, ld, BN, tiles_m, tiles_n, sms, h, total_tiles, 0);
      cudaDeviceSynchronize();
      const long long probe[6] = {0, 1, ms - 1, ms, total / 2 + 7, total - 1};
      float got[6], want_p;
      for (int i = 0; i < 6; ++i) cudaMemcpy(&got[i], P + probe[i], 4, cudaMemcpyDeviceToHost);
      // host replica of two updates from zero
      {
        float p = 0.f, m = 0.f, v = 0.f;
        for (int t = 0; t < 2; ++t) {
          float gr = 1e-3f + h.wd * p;
          m = h.b1 * m + (1.f - h.b1) * gr;
          v = h.b2 * v + (1.f - h.b2) * gr * gr;
          p = p - h.lr * (m / (sqrtf(v) * h.isb + h.eps));
        }
        want_p = p;
      }
      int bad = 0;
      for (int i = 0; i < 6; ++i) bad += fabsf(got[i] - want_p) > 1e-6f * fabsf(want_p) + 1e-9f;
      // full check of p through a device-side reduction would need another kernel; sample the min/max instead
      float* hp = (float*)malloc(ms * 4);
      cudaMemcpy(hp, P + (E - 1) * ms, ms * 4, cudaMemcpyDeviceToHost);
      float lo = 1e30f, hi = -1e30f;
      for (long long i = 0; i < ms; ++i) { lo = fminf(lo, hp[i]); hi = fmaxf(hi, hp[i]); }
      __nv_bfloat16 s0;
      cudaMemcpy(&s0, S + (long long)(E - 1) * sms + 8 * 3 + 5, 2, cudaMemcpyDeviceToHost);  // chunk 0, row 3, k 5
      printf("verify: want p %.9g, probes bad %d, last model min %.9g max %.9g, shadow sample %.6g\n", want_p, bad, lo, hi,
             __bfloat162float(s0));
      free(hp);
    }
    cudaFree(P); cudaFree(M); cudaFree(V); cudaFree(S);
  }
  return 0;
}
