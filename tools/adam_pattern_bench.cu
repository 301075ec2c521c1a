// Micro-benchmark: how fast can the Adam read-modify-write of p, m, v (fp32) + bf16 shadow stream
// through HBM under different access patterns?  Informs the layout of the fused dW+Adam epilogue.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/adam_pattern_bench.cu -o gpurun_out/adam_pattern_bench
#include <cuda_bf16.h>
#include <cuda_runtime.h>
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

// P3: linear float4
__global__ void k_linear(float4* P, float4* M, float4* V, __nv_bfloat162* S, long long n4, H h) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 p = P[i], m = M[i], v = V[i];
    upd(1e-3f, p.x, m.x, v.x, h); upd(1e-3f, p.y, m.y, v.y, h); upd(1e-3f, p.z, m.z, v.z, h); upd(1e-3f, p.w, m.w, v.w, h);
    P[i] = p; M[i] = m; V[i] = v;
    S[2 * i] = __floats2bfloat162_rn(p.x, p.y); S[2 * i + 1] = __floats2bfloat162_rn(p.z, p.w);
  }
}

// P1: tiles [128 k x BN n] of a [rows n][ld] matrix per model; CTA 256 threads: warp w: k quarter w&3, column group w>>2;
// each thread batches NB columns (loads first, then math, then stores), like the GEMM epilogue.
template <int NB>
__global__ void k_tile(float* P, float* M, float* V, __nv_bfloat16* S, int rows, int ld, int BN, int tiles_m, int tiles_n, long long ms, H h,
                       int m_fast) {
  int tile_m, tile_n;
  if (m_fast) { tile_m = blockIdx.x % tiles_m; tile_n = blockIdx.x / tiles_m; } else { tile_n = blockIdx.x % tiles_n; tile_m = blockIdx.x / tiles_n; }
  const int model = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = tile_m * 128 + (warp & 3) * 32 + lane;
  const int cg = warp >> 2;
  if (k >= ld) return;
  float* p = P + model * ms + k; float* m = M + model * ms + k; float* v = V + model * ms + k;
  __nv_bfloat16* s = S + model * ms + k;
  for (int c = cg * NB; c < BN; c += 2 * NB) {
    float pv[NB], mv[NB], vv[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      int n = tile_n * BN + c + i;
      pv[i] = mv[i] = vv[i] = 0.f;
      if (n < rows) { long long idx = (long long)n * ld; pv[i] = p[idx]; mv[i] = m[idx]; vv[i] = v[idx]; }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) upd(1e-3f, pv[i], mv[i], vv[i], h);
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      int n = tile_n * BN + c + i;
      if (n < rows) { long long idx = (long long)n * ld; p[idx] = pv[i]; m[idx] = mv[i]; v[idx] = vv[i]; s[idx] = __float2bfloat16_rn(pv[i]); }
    }
  }
}

// P1b: same tile pattern under the fused GEMM kernel's occupancy (2 CTAs/SM: 96 KB dynamic smem each), double-buffered
// register prefetch of 8 columns (what the epilogue does today).
__global__ void __launch_bounds__(256, 2) k_tile_occ2(float* P, float* M, float* V, __nv_bfloat16* S, int rows, int ld, int BN, int tiles_m,
                                                      int tiles_n, long long ms, H h) {
  extern __shared__ float sm[];
  const int tile_m = blockIdx.x % tiles_m, tile_n = blockIdx.x / tiles_m;
  const int model = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = tile_m * 128 + (warp & 3) * 32 + lane;
  const int cg = warp >> 2;
  if (k >= ld) return;
  if (sm[0] == 123.f) return;
  float* p = P + model * ms + k; float* m = M + model * ms + k; float* v = V + model * ms + k;
  __nv_bfloat16* s = S + model * ms + k;
  for (int c = cg * 8; c < BN; c += 16) {
    float pv[8], mv[8], vv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int n = tile_n * BN + c + i;
      pv[i] = mv[i] = vv[i] = 0.f;
      if (n < rows) { long long idx = (long long)n * ld; pv[i] = p[idx]; mv[i] = m[idx]; vv[i] = v[idx]; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) upd(1e-3f, pv[i], mv[i], vv[i], h);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int n = tile_n * BN + c + i;
      if (n < rows) { long long idx = (long long)n * ld; p[idx] = pv[i]; m[idx] = mv[i]; v[idx] = vv[i]; s[idx] = __float2bfloat16_rn(pv[i]); }
    }
  }
}

// P1d: persistent-kernel geometry: 1 CTA/SM (200 KB smem), 16 epilogue warps = 4 k-quarters x 4 column groups, NB=8 with
// double buffering; SHADOW: 0 none, 1 linear bf16, 2 chunk8-scattered bf16 ([k/8][n][k%8], what the GEMMs read)
// P1e: persistent, EW epilogue warps per CTA (EW/4 column groups), register budget set by the thread count
template <int EW, int NB>
__global__ void __launch_bounds__(EW * 32 + 128, 1) k_tile_pw(float* P, float* M, float* V, __nv_bfloat16* S, int rows, int ld, int BN,
                                                              int tiles_m, int tiles_n, long long ms, H h, int total_tiles) {
  extern __shared__ float sm[];
  if (sm[0] == 123.f) return;
  const int warp = (threadIdx.x >> 5) - 4, lane = threadIdx.x & 31;
  if (warp < 0) return;  // the producer / MMA warps of the real kernel
  const int cg = warp >> 2, ngroups = EW / 4;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int per = tiles_m * tiles_n;
    const int model = tile / per, mn = tile % per;
    const int tile_m = mn % tiles_m, tile_n = mn / tiles_m;
    const int k = tile_m * 128 + (warp & 3) * 32 + lane;
    if (k >= ld) continue;
    float* p = P + model * ms + k; float* m = M + model * ms + k; float* v = V + model * ms + k;
    __nv_bfloat16* s = S + model * ms;
    const int rcap = tiles_n * BN;
    for (int c = cg * NB; c < BN; c += ngroups * NB) {
      float pv[NB], mv[NB], vv[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        int n = tile_n * BN + c + i;
        pv[i] = mv[i] = vv[i] = 0.f;
        if (n < rows && c + i < BN) { long long idx = (long long)n * ld; pv[i] = p[idx]; mv[i] = m[idx]; vv[i] = v[idx]; }
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) upd(1e-3f, pv[i], mv[i], vv[i], h);
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        int n = tile_n * BN + c + i;
        if (n < rows && c + i < BN) {
          long long idx = (long long)n * ld; p[idx] = pv[i]; m[idx] = mv[i]; v[idx] = vv[i];
          s[((long long)(k >> 3) * rcap + n) * 8 + (k & 7)] = __float2bfloat16_rn(pv[i]);
        }
      }
    }
  }
}

template <int SHADOW>
__global__ void __launch_bounds__(512) k_tile_p16(float* P, float* M, float* V, __nv_bfloat16* S, int rows, int ld, int BN, int tiles_m,
                                                     int tiles_n, long long ms, H h, int total_tiles, int variant) {
  extern __shared__ float sm[];
  if (sm[0] == 123.f) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = warp >> 2;
  for (int tile0 = blockIdx.x; tile0 < total_tiles; tile0 += gridDim.x) {
    int tile = tile0;
    if (variant == 2) {  // contiguous tile ranges per CTA instead of strided
      const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
      tile = blockIdx.x * per_cta + (tile0 / gridDim.x);
      if (tile >= total_tiles || (tile0 / (int)gridDim.x) >= per_cta) continue;
    }
    const int per = tiles_m * tiles_n;
    const int model = tile / per, mn = tile % per;
    int tile_m = mn % tiles_m, tile_n = mn / tiles_m;
    if (variant == 1) { tile_n = mn % tiles_n; tile_m = mn / tiles_n; }
    const int k = tile_m * 128 + (warp & 3) * 32 + lane;
    if (k >= ld) continue;
    float* p = P + model * ms + k; float* m = M + model * ms + k; float* v = V + model * ms + k;
    __nv_bfloat16* s = S + model * ms;
    const int rcap = tiles_n * BN;
    const int nb = BN / 32;
    const int rot = (variant == 3) ? (int)(blockIdx.x * 5u) % nb : 0;  // per-CTA rotation of the batch order
    for (int jb = 0; jb < nb; ++jb) {
      const int c = cg * 8 + 32 * ((jb + rot) % nb);
      float pv[8], mv[8], vv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int n = tile_n * BN + c + i;
        pv[i] = mv[i] = vv[i] = 0.f;
        if (n < rows) { long long idx = (long long)n * ld; pv[i] = p[idx]; mv[i] = m[idx]; vv[i] = v[idx]; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) upd(1e-3f, pv[i], mv[i], vv[i], h);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int n = tile_n * BN + c + i;
        if (n < rows) {
          long long idx = (long long)n * ld; p[idx] = pv[i]; m[idx] = mv[i]; v[idx] = vv[i];
          if (SHADOW == 1) s[idx + k] = __float2bfloat16_rn(pv[i]);
          if (SHADOW == 2) s[((long long)(k >> 3) * rcap + n) * 8 + (k & 7)] = __float2bfloat16_rn(pv[i]);
        }
      }
    }
  }
}

// P1c: per-thread cp.async (4 B) ring in shared memory, depth R stages of 8 columns x 3 arrays; no registers hold data in
// flight.  Same 2 CTAs/SM occupancy.
template <int R>
__global__ void __launch_bounds__(256, 2) k_tile_cpasync(float* P, float* M, float* V, __nv_bfloat16* S, int rows, int ld, int BN, int tiles_m,
                                                         int tiles_n, long long ms, H h) {
  extern __shared__ float sm[];  // [R][24][256]
  const int tile_m = blockIdx.x % tiles_m, tile_n = blockIdx.x / tiles_m;
  const int model = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = tile_m * 128 + (warp & 3) * 32 + lane;
  const int cg = warp >> 2;
  if (k >= ld) return;
  float* p = P + model * ms + k; float* m = M + model * ms + k; float* v = V + model * ms + k;
  __nv_bfloat16* s = S + model * ms + k;
  const int nh = (BN / 8 - cg + 1) / 2;  // half-chunks of this group: columns c = (cg + 2 j) * 8
  auto issue = [&](int j) {
    if (j < nh) {
      const int c = (cg + 2 * j) * 8;
      float* dst = sm + (size_t)(j % R) * 24 * 256 + threadIdx.x;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int n = tile_n * BN + c + i;
        if (n < rows) {
          long long idx = (long long)n * ld;
          unsigned d0 = (unsigned)__cvta_generic_to_shared(dst + (3 * i + 0) * 256);
          unsigned d1 = (unsigned)__cvta_generic_to_shared(dst + (3 * i + 1) * 256);
          unsigned d2 = (unsigned)__cvta_generic_to_shared(dst + (3 * i + 2) * 256);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0), "l"(p + idx) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d1), "l"(m + idx) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d2), "l"(v + idx) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int j = 0; j < R - 1; ++j) issue(j);
  for (int j = 0; j < nh; ++j) {
    issue(j + R - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(R - 1) : "memory");
    const int c = (cg + 2 * j) * 8;
    const float* src = sm + (size_t)(j % R) * 24 * 256 + threadIdx.x;
    float pv[8], mv[8], vv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { pv[i] = src[(3 * i) * 256]; mv[i] = src[(3 * i + 1) * 256]; vv[i] = src[(3 * i + 2) * 256]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) upd(1e-3f, pv[i], mv[i], vv[i], h);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int n = tile_n * BN + c + i;
      if (n < rows) { long long idx = (long long)n * ld; p[idx] = pv[i]; m[idx] = mv[i]; v[idx] = vv[i]; s[idx] = __float2bfloat16_rn(pv[i]); }
    }
  }
}


// P4: persistent geometry, optimizer state staged through shared memory by the bulk-copy (TMA) engine: producer warps
// issue one 512-byte copy per (array, weight row) of a half-chunk (8 weight rows x 128 k) into a ring of NST stages,
// 4 quarter-warps of a column group consume a stage.  In-flight bytes are set by the ring, not by registers x warps.
// BULK_ST: results go back into the stage and a store warp writes them out with bulk shared->global copies.
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mb_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(unsigned long long* b, unsigned par, int tag = 0) {
  for (unsigned spin = 0;; ++spin) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    if (ok) return;
    if (spin > (1u << 22)) {  // watchdog: a protocol bug must not hang the box
      printf("mb_wait timeout: block %d warp %d tag %d parity %u\n", blockIdx.x, threadIdx.x >> 5, tag, par);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(src)), "r"(bytes) : "memory");
}

template <int EW, int NST, int PW, int BULK_ST>
__global__ void __launch_bounds__((EW + 4) * 32, 1) k_tile_bulk(float* P, float* M, float* V, __nv_bfloat16* S, int rows, int ld, int BN, int tiles_m,
                                                                 int tiles_n, long long ms, H h, int total_tiles, int reserve_bytes) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ __align__(8) unsigned long long full[NST], empty[NST], done[NST];
  constexpr int PITCH = 132;  // floats per staged weight-row segment (128 + one 16-byte slot for misaligned rows)
  constexpr int STAGE = 3 * 8 * PITCH;
  constexpr int GROUPS = EW / 4;
  float* ring = reinterpret_cast<float*>(smraw + reserve_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], BULK_ST ? 1 : 4); mb_init(&done[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nhc = BN >> 3, per = tiles_m * tiles_n;
  if (warp < PW) {
    unsigned g = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int model = tile / per, mn = tile % per, tile_m = mn % tiles_m, tile_n = mn / tiles_m;
      const int k0 = tile_m * 128;
      const unsigned kbytes = (unsigned)min(128, ld - k0) * 4u;
      for (int hc = 0; hc < nhc; ++hc, ++g) {
        const int slot = g % NST;
        const unsigned ph = (g / NST) & 1;
        if (lane == 0) mb_wait(&empty[slot], ph ^ 1, 100 + slot);
        __syncwarp();
        const int n0 = tile_n * BN + hc * 8;
        const int nr = max(0, min(8, rows - n0));
        if (warp == 0 && lane == 0) {
          if (nr > 0) mb_expect(&full[slot], 3u * nr * kbytes); else mb_arrive(&full[slot]);
        }
        __syncwarp();
        const int c = warp + PW * lane;
        if (c < 24) {
          const int a = c >> 3, i = c & 7;
          if (i < nr) {
            const float* src = (a == 0 ? P : (a == 1 ? M : V)) + model * ms + (long long)(n0 + i) * ld + k0;
            bulk_g2s(ring + (size_t)slot * STAGE + (a * 8 + i) * PITCH, src, kbytes, &full[slot]);
          }
        }
      }
    }
  } else if (BULK_ST && warp == 3) {
    unsigned g = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int model = tile / per, mn = tile % per, tile_m = mn % tiles_m, tile_n = mn / tiles_m;
      const int k0 = tile_m * 128;
      const unsigned kbytes = (unsigned)min(128, ld - k0) * 4u;
      for (int hc = 0; hc < nhc; ++hc, ++g) {
        const int slot = g % NST;
        const unsigned ph = (g / NST) & 1;
        if (lane == 0) mb_wait(&done[slot], ph, 200 + slot);
        __syncwarp();
        const int n0 = tile_n * BN + hc * 8;
        const int nr = max(0, min(8, rows - n0));
        if (lane < 24) {
          const int a = lane >> 3, i = lane & 7;
          if (i < nr) {
            float* dst = (a == 0 ? P : (a == 1 ? M : V)) + model * ms + (long long)(n0 + i) * ld + k0;
            bulk_s2g(dst, ring + (size_t)slot * STAGE + (a * 8 + i) * PITCH, kbytes);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) mb_arrive(&empty[slot]);
      }
    }
    if (lane < 24) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp >= 4) {
    const int q = warp & 3, cg = (warp - 4) >> 2;
    unsigned j = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
      const int model = tile / per, mn = tile % per, tile_m = mn % tiles_m, tile_n = mn / tiles_m;
      const int kk = q * 32 + lane, k = tile_m * 128 + kk;
      const bool kok = k < ld;
      const int rcap = tiles_n * BN;
      float* p = P + model * ms + k; float* m = M + model * ms + k; float* v = V + model * ms + k;
      __nv_bfloat16* s = S + model * ms;
      for (int hc = cg; hc < nhc; hc += GROUPS) {
        const unsigned g = j * nhc + hc;
        const int slot = g % NST;
        const unsigned ph = (g / NST) & 1;
        if (lane == 0) mb_wait(&full[slot], ph, 300 + slot);
        __syncwarp();
        float* st = ring + (size_t)slot * STAGE + kk;
        const int n0 = tile_n * BN + hc * 8;
        float pv[8], mv[8], vv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { pv[i] = st[i * PITCH]; mv[i] = st[(8 + i) * PITCH]; vv[i] = st[(16 + i) * PITCH]; }
#pragma unroll
        for (int i = 0; i < 8; ++i) upd(1e-3f, pv[i], mv[i], vv[i], h);
        if (kok) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n = n0 + i;
            if (n < rows) {
              long long idx = (long long)n * ld;
              if (BULK_ST) { st[i * PITCH] = pv[i]; st[(8 + i) * PITCH] = mv[i]; st[(16 + i) * PITCH] = vv[i]; }
              else { p[idx] = pv[i]; m[idx] = mv[i]; v[idx] = vv[i]; }
              s[((long long)(k >> 3) * rcap + n) * 8 + (k & 7)] = __float2bfloat16_rn(pv[i]);
            }
          }
        }
        if (BULK_ST) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mb_arrive(BULK_ST ? &done[slot] : &empty[slot]);
      }
    }
  }
}


// P5: persistent geometry, 16-byte accesses: after a 4x4 lane-quad transpose of the accumulators (not modelled here) a
// thread owns 4 consecutive k of 2 of the 8 weight rows of a half-chunk: 6 x LDG.128 / STG.128 instead of 24 scalar
// accesses for the same bytes.  VEC = 4: float4 (rows 16-byte aligned), VEC = 2: two float2 per row (8-byte aligned
// rows, e.g. ld = 978).  DB = 1: the loads of the next half-chunk are issued before the math of the current one.
template <int EW, int VEC, int DB>
__global__ void __launch_bounds__(EW * 32 + 128, 1) k_tile_v4(float* P, float* M, float* V, __nv_bfloat16* S, int rows, int ld, int BN, int tiles_m,
                                                              int tiles_n, long long ms, H h, int total_tiles) {
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
    __nv_bfloat16* s = S + model * ms;
    const int rcap = tiles_n * BN;
    auto load = [&](int hc, Buf& b) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = tile_n * BN + hc * 8 + j * 4 + ci;
        b.p[j] = b.m[j] = b.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < rows && kok) {
          const long long idx = (long long)n * ld;
          if (VEC == 4) {
            b.p[j] = *reinterpret_cast<const float4*>(p + idx); b.m[j] = *reinterpret_cast<const float4*>(m + idx); b.v[j] = *reinterpret_cast<const float4*>(v + idx);
          } else {
            float2 a0 = *reinterpret_cast<const float2*>(p + idx), a1 = *reinterpret_cast<const float2*>(p + idx + 2);
            float2 b0 = *reinterpret_cast<const float2*>(m + idx), b1 = *reinterpret_cast<const float2*>(m + idx + 2);
            float2 c0 = *reinterpret_cast<const float2*>(v + idx), c1 = *reinterpret_cast<const float2*>(v + idx + 2);
            b.p[j] = make_float4(a0.x, a0.y, a1.x, a1.y); b.m[j] = make_float4(b0.x, b0.y, b1.x, b1.y); b.v[j] = make_float4(c0.x, c0.y, c1.x, c1.y);
          }
        }
      }
    };
    auto apply = [&](int hc, Buf& b) {
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
          if (VEC == 4) {
            *reinterpret_cast<float4*>(p + idx) = b.p[j]; *reinterpret_cast<float4*>(m + idx) = b.m[j]; *reinterpret_cast<float4*>(v + idx) = b.v[j];
          } else {
            *reinterpret_cast<float2*>(p + idx) = make_float2(b.p[j].x, b.p[j].y); *reinterpret_cast<float2*>(p + idx + 2) = make_float2(b.p[j].z, b.p[j].w);
            *reinterpret_cast<float2*>(m + idx) = make_float2(b.m[j].x, b.m[j].y); *reinterpret_cast<float2*>(m + idx + 2) = make_float2(b.m[j].z, b.m[j].w);
            *reinterpret_cast<float2*>(v + idx) = make_float2(b.v[j].x, b.v[j].y); *reinterpret_cast<float2*>(v + idx + 2) = make_float2(b.v[j].z, b.v[j].w);
          }
          __nv_bfloat162 lo = __floats2bfloat162_rn(b.p[j].x, b.p[j].y), hi = __floats2bfloat162_rn(b.p[j].z, b.p[j].w);
          uint2 pk = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
          *reinterpret_cast<uint2*>(s + ((long long)(k4 >> 3) * rcap + n) * 8 + (k4 & 7)) = pk;
        }
      }
    };
    if (DB) {
      Buf A, B;
      int hc = cg;
      if (hc < nhc) load(hc, A);
      for (; hc < nhc; hc += 2 * ngroups) {
        if (hc + ngroups < nhc) load(hc + ngroups, B);
        apply(hc, A);
        if (hc + 2 * ngroups < nhc) load(hc + 2 * ngroups, A);
        if (hc + ngroups < nhc) apply(hc + ngroups, B);
      }
    } else {
      Buf A;
      for (int hc = cg; hc < nhc; hc += ngroups) { load(hc, A); apply(hc, A); }
    }
  }
}


// P6: as P5 (16-byte accesses, persistent geometry, 16 epilogue warps) but with the tile's k-span KS as a parameter:
// KS = 128 reads 512-byte pieces of each weight row (the GEMM epilogue's pattern), 256 / 512 read 1 / 2 KB pieces
// (tile = KS x 32768/KS).  Tests whether DRAM page locality limits the tile pattern.
template <int KS>
__global__ void __launch_bounds__(16 * 32 + 128, 1) k_tile_ks(float* P, float* M, float* V, __nv_bfloat16* S, int rows, int ld, long long ms, H h,
                                                             int n_models) {
  extern __shared__ float sm[];
  if (sm[0] == 123.f) return;
  const int warp = (threadIdx.x >> 5) - 4, lane = threadIdx.x & 31;
  if (warp < 0) return;
  constexpr int NS = 32768 / KS, KSLOTS = KS / 32, GROUPS = 16 / KSLOTS, NHC = NS / 8;
  const int ks = warp % KSLOTS, cg = warp / KSLOTS;
  const int ci = lane & 3, kg = lane >> 2;
  const int tiles_m = (ld + KS - 1) / KS, tiles_n = (rows + NS - 1) / NS;
  const int total_tiles = tiles_m * tiles_n * n_models;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int per = tiles_m * tiles_n;
    const int model = tile / per, mn = tile % per;
    const int tile_m = mn % tiles_m, tile_n = mn / tiles_m;
    const int k4 = tile_m * KS + ks * 32 + kg * 4;
    const bool kok = k4 + 3 < ld;
    float* p = P + model * ms + k4; float* m = M + model * ms + k4; float* v = V + model * ms + k4;
    for (int hc = cg; hc < NHC; hc += GROUPS) {
      float4 bp[2], bm[2], bv[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = tile_n * NS + hc * 8 + j * 4 + ci;
        bp[j] = bm[j] = bv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < rows && kok) {
          const long long idx = (long long)n * ld;
          bp[j] = *reinterpret_cast<const float4*>(p + idx); bm[j] = *reinterpret_cast<const float4*>(m + idx); bv[j] = *reinterpret_cast<const float4*>(v + idx);
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        upd(1e-3f, bp[j].x, bm[j].x, bv[j].x, h); upd(1e-3f, bp[j].y, bm[j].y, bv[j].y, h);
        upd(1e-3f, bp[j].z, bm[j].z, bv[j].z, h); upd(1e-3f, bp[j].w, bm[j].w, bv[j].w, h);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = tile_n * NS + hc * 8 + j * 4 + ci;
        if (n < rows && kok) {
          const long long idx = (long long)n * ld;
          *reinterpret_cast<float4*>(p + idx) = bp[j]; *reinterpret_cast<float4*>(m + idx) = bm[j]; *reinterpret_cast<float4*>(v + idx) = bv[j];
          __nv_bfloat162 lo = __floats2bfloat162_rn(bp[j].x, bp[j].y), hi = __floats2bfloat162_rn(bp[j].z, bp[j].w);
          uint2 pk = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
          *reinterpret_cast<uint2*>(S + model * ms + ((long long)(k4 >> 3) * 2048 + n) * 8 + (k4 & 7)) = pk;
        }
      }
    }
  }
}


// P7: as P5 (vec4, persistent geometry) but the optimizer state is fetched with per-thread 16-byte cp.async.cg
// (LDGSTS.128, L1 bypass) into a private shared-memory ring of D half-chunks per warp: data in flight is bounded by
// shared memory, not by registers.  The thread that issued a copy is the one that reads it back (no cross-thread sync).
template <int EW, int D>
__global__ void __launch_bounds__(EW * 32 + 128, 1) k_tile_cpa16(float* P, float* M, float* V, __nv_bfloat16* S, int rows, int ld, int BN, int tiles_m,
                                                                 int tiles_n, long long ms, H h, int total_tiles, int reserve) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int warp = (threadIdx.x >> 5) - 4, lane = threadIdx.x & 31;
  if (warp < 0) return;
  float4* ring = reinterpret_cast<float4*>(smraw + reserve) + (size_t)warp * D * 6 * 32;  // [D][6][32 lanes] float4
  const int cg = warp >> 2, ngroups = EW / 4, q = warp & 3;
  const int ci = lane & 3, kg = lane >> 2;
  const int nhc = BN >> 3;
  const int nh = (nhc - cg + ngroups - 1) / ngroups;           // half-chunks of this warp per tile
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_tiles * nh;                              // work items of this warp
  auto item = [&](int w, int& model, int& k4, int& n0, bool& kok) {
    const int tile = blockIdx.x + (w / nh) * gridDim.x, hc = cg + (w % nh) * ngroups;
    const int per = tiles_m * tiles_n;
    model = tile / per;
    const int mn = tile % per, tile_m = mn % tiles_m, tile_n = mn / tiles_m;
    k4 = tile_m * 128 + q * 32 + kg * 4;
    kok = k4 + 3 < ld;
    n0 = tile_n * BN + hc * 8;
  };
  auto issue = [&](int w) {
    if (w < total) {
      int model, k4, n0; bool kok;
      item(w, model, k4, n0, kok);
      float4* dst = ring + (size_t)(w % D) * 6 * 32 + lane;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = n0 + j * 4 + ci;
        if (n < rows && kok) {
          const long long idx = model * ms + (long long)n * ld + k4;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + (3 * j + 0) * 32)), "l"(P + idx) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + (3 * j + 1) * 32)), "l"(M + idx) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + (3 * j + 2) * 32)), "l"(V + idx) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int w = 0; w < D - 1; ++w) issue(w);
  for (int w = 0; w < total; ++w) {
    issue(w + D - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
    int model, k4, n0; bool kok;
    item(w, model, k4, n0, kok);
    const float4* src = ring + (size_t)(w % D) * 6 * 32 + lane;
    const int rcap = tiles_n * BN;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + j * 4 + ci;
      if (n < rows && kok) {
        float4 bp = src[(3 * j + 0) * 32], bm = src[(3 * j + 1) * 32], bv = src[(3 * j + 2) * 32];
        upd(1e-3f, bp.x, bm.x, bv.x, h); upd(1e-3f, bp.y, bm.y, bv.y, h); upd(1e-3f, bp.z, bm.z, bv.z, h); upd(1e-3f, bp.w, bm.w, bv.w, h);
        const long long idx = model * ms + (long long)n * ld + k4;
        *reinterpret_cast<float4*>(P + idx) = bp; *reinterpret_cast<float4*>(M + idx) = bm; *reinterpret_cast<float4*>(V + idx) = bv;
        __nv_bfloat162 lo = __floats2bfloat162_rn(bp.x, bp.y), hi = __floats2bfloat162_rn(bp.z, bp.w);
        uint2 pk = make_uint2(*reinterpret_cast<unsigned*>(&lo), *reinterpret_cast<unsigned*>(&hi));
        *reinterpret_cast<uint2*>(S + model * ms + ((long long)(k4 >> 3) * rcap + n) * 8 + (k4 & 7)) = pk;
      }
    }
  }
}

// P2: row-linear: CTA handles RN consecutive rows n, all k; thread t handles k = t, t+256, ... for each row; NB rows batched.
template <int NB>
__global__ void k_rows(float* P, float* M, float* V, __nv_bfloat16* S, int rows, int ld, int RN, long long ms, H h) {
  const int model = blockIdx.y;
  const int n0 = blockIdx.x * RN;
  float* p = P + model * ms; float* m = M + model * ms; float* v = V + model * ms;
  __nv_bfloat16* s = S + model * ms;
  for (int k = threadIdx.x; k < ld; k += blockDim.x) {
    for (int c = 0; c < RN; c += NB) {
      float pv[NB], mv[NB], vv[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        int n = n0 + c + i;
        pv[i] = mv[i] = vv[i] = 0.f;
        if (n < rows) { long long idx = (long long)n * ld + k; pv[i] = p[idx]; mv[i] = m[idx]; vv[i] = v[idx]; }
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) upd(1e-3f, pv[i], mv[i], vv[i], h);
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        int n = n0 + c + i;
        if (n < rows) { long long idx = (long long)n * ld + k; p[idx] = pv[i]; m[idx] = mv[i]; v[idx] = vv[i]; s[idx] = __float2bfloat16_rn(pv[i]); }
      }
    }
  }
}

int main(int argc, char** argv) {
  setvbuf(stdout, NULL, _IONBF, 0);
  const bool only_bulk = argc > 1;
  const int E = 32, rows = 1956, ld = 600;  // decoder heads
  const long long ms = (long long)rows * ld, total = ms * E;
  float *P, *M, *V; __nv_bfloat16* S;
  cudaMalloc(&P, total * 4); cudaMalloc(&M, total * 4); cudaMalloc(&V, total * 4); cudaMalloc(&S, total * 2);
  cudaMemset(P, 0, total * 4); cudaMemset(M, 0, total * 4); cudaMemset(V, 0, total * 4);
  H h{5e-4f, 0.9f, 0.999f, 1e-8f, 0.05f, 1.f};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto launch) {
    if (only_bulk && strncmp(name, "bulk", 4) != 0 && strncmp(name, "linear", 6) != 0 && strncmp(name, "persistent 148 CTA", 18) != 0) return;
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    const int R = 10;
    for (int i = 0; i < R; ++i) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms_; cudaEventElapsedTime(&ms_, e0, e1); ms_ /= R;
    cudaError_t err = cudaGetLastError();
    printf("%-44s %8.1f us  %7.0f GB/s (26 B/param)  %s\n", name, ms_ * 1e3, 26.0 * total / (ms_ * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
  };
  timeit("linear float4, 148*8 CTAs x 256", [&] { k_linear<<<148 * 8, 256>>>((float4*)P, (float4*)M, (float4*)V, (__nv_bfloat162*)S, total / 4, h); });
  timeit("linear float4, 148*32 CTAs x 256", [&] { k_linear<<<148 * 32, 256>>>((float4*)P, (float4*)M, (float4*)V, (__nv_bfloat162*)S, total / 4, h); });
  for (int BN : {256, 64}) {
    int tiles_m = (ld + 127) / 128, tiles_n = (rows + BN - 1) / BN;
    char nm[128];
    for (int mf = 0; mf < 2; ++mf) {
      snprintf(nm, sizeof(nm), "tile 128 x %d, NB=8, m_fast=%d", BN, mf);
      timeit(nm, [&] { k_tile<8><<<dim3(tiles_m * tiles_n, E), 256>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, mf); });
      snprintf(nm, sizeof(nm), "tile 128 x %d, NB=16, m_fast=%d", BN, mf);
      timeit(nm, [&] { k_tile<16><<<dim3(tiles_m * tiles_n, E), 256>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, mf); });
    }
  }
  {
    const int BN = 256, tiles_m = (ld + 127) / 128, tiles_n = (rows + BN - 1) / BN;
    cudaFuncSetAttribute(k_tile_occ2, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    cudaFuncSetAttribute(k_tile_cpasync<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    cudaFuncSetAttribute(k_tile_cpasync<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    cudaFuncSetAttribute(k_tile_cpasync<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    timeit("tile 128x256 regs NB=8, 2 CTA/SM (96KB smem)", [&] { k_tile_occ2<<<dim3(tiles_m * tiles_n, E), 256, 98304>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h); });
    timeit("tile 128x256 cp.async ring R=2, 2 CTA/SM", [&] { k_tile_cpasync<2><<<dim3(tiles_m * tiles_n, E), 256, 98304>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h); });
    timeit("tile 128x256 cp.async ring R=3, 2 CTA/SM", [&] { k_tile_cpasync<3><<<dim3(tiles_m * tiles_n, E), 256, 98304>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h); });
    timeit("tile 128x256 cp.async ring R=4, 2 CTA/SM", [&] { k_tile_cpasync<4><<<dim3(tiles_m * tiles_n, E), 256, 98304>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h); });
  }
  {
    const int BN = 256, tiles_m = (ld + 127) / 128, tiles_n = (rows + BN - 1) / BN;
    const int total_tiles = tiles_m * tiles_n * E;
    cudaFuncSetAttribute(k_tile_p16<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_tile_p16<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_tile_p16<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int var = 0; var < 4; ++var) {
      char nm[128];
      snprintf(nm, sizeof(nm), "persistent 148x16 warps, no shadow, variant %d", var);
      timeit(nm, [&] { k_tile_p16<0><<<148, 512, 200 * 1024>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, total_tiles, var); });
    }
    timeit("persistent 148x16 warps, chunk8 shadow, v0", [&] { k_tile_p16<2><<<148, 512, 200 * 1024>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, total_tiles, 0); });
#define PW(EW, NB, SM)                                                                                                  \
    {                                                                                                                    \
      cudaFuncSetAttribute(k_tile_pw<EW, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM * 1024);                   \
      char nm[128];                                                                                                      \
      snprintf(nm, sizeof(nm), "persistent 148 CTA, %d epi warps, NB=%d, %d KB smem", EW, NB, SM);                       \
      timeit(nm, [&] { k_tile_pw<EW, NB><<<148, EW * 32 + 128, SM * 1024>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, total_tiles); }); \
    }
    PW(16, 8, 200) PW(16, 8, 100) PW(20, 8, 100) PW(24, 8, 100) PW(28, 8, 100) PW(28, 4, 100) PW(28, 8, 200) PW(24, 4, 100) PW(16, 16, 100)

#define BK_(EW, NST, PW, BST, RES)                                                                                        \
    {                                                                                                                    \
      const int smem = RES * 1024 + NST * 3 * 8 * 132 * 4;                                                                \
      cudaFuncSetAttribute(k_tile_bulk<EW, NST, PW, BST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);             \
      char nm[128];                                                                                                      \
      snprintf(nm, sizeof(nm), "bulk-staged: %d epi warps, %d stages, %d prod, st=%d, +%dKB", EW, NST, PW, BST, RES);     \
      timeit(nm, [&] { k_tile_bulk<EW, NST, PW, BST><<<148, (EW + 4) * 32, smem>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, total_tiles, RES * 1024); }); \
    }

#define V4_(EW, VEC, DB, SM)                                                                                              \
    {                                                                                                                    \
      cudaFuncSetAttribute(k_tile_v4<EW, VEC, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM * 1024);               \
      char nm[128];                                                                                                      \
      snprintf(nm, sizeof(nm), "bulk-alt vec%d: persistent, %d epi warps, db=%d, %d KB smem", VEC, EW, DB, SM);           \
      timeit(nm, [&] { k_tile_v4<EW, VEC, DB><<<148, EW * 32 + 128, SM * 1024>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, total_tiles); }); \
    }

#define KS_(KS)                                                                                                           \
    {                                                                                                                    \
      cudaFuncSetAttribute(k_tile_ks<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);                       \
      char nm[128];                                                                                                      \
      snprintf(nm, sizeof(nm), "bulk-alt k-span %d: persistent vec4, 16 epi warps, 100 KB smem", KS);                     \
      timeit(nm, [&] { k_tile_ks<KS><<<148, 16 * 32 + 128, 100 * 1024>>>(P, M, V, S, rows, ld, ms, h, E); });             \
    }
    KS_(128)
    V4_(24, 4, 0, 1) V4_(24, 4, 0, 32) V4_(24, 4, 0, 64) V4_(24, 4, 0, 100) V4_(24, 4, 0, 132) V4_(24, 4, 0, 164) V4_(16, 4, 1, 1) V4_(16, 4, 1, 32) V4_(16, 4, 1, 64) V4_(28, 4, 0, 32)

#define CPA_(EW, D, RES)                                                                                                  \
    {                                                                                                                    \
      const int smem = RES * 1024 + EW * D * 6 * 32 * 16;                                                                 \
      cudaFuncSetAttribute(k_tile_cpa16<EW, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                       \
      char nm[128];                                                                                                      \
      snprintf(nm, sizeof(nm), "bulk-alt cp.async.cg 16B: %d epi warps, depth %d, +%d KB (%d KB smem)", EW, D, RES, smem / 1024); \
      timeit(nm, [&] { k_tile_cpa16<EW, D><<<148, EW * 32 + 128, smem>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, total_tiles, RES * 1024); }); \
    }
    CPA_(16, 2, 0) CPA_(8, 2, 0) CPA_(16, 4, 0)

    V4_(16, 4, 0, 100) V4_(16, 4, 1, 100) V4_(24, 4, 0, 100) V4_(24, 4, 1, 100) V4_(16, 4, 1, 200) V4_(16, 2, 1, 100) V4_(24, 2, 0, 100) V4_(24, 2, 1, 100) V4_(12, 4, 1, 100) V4_(8, 4, 1, 100)
    PW(16, 8, 100) PW(24, 8, 100)
    BK_(16, 4, 3, 0, 96) BK_(16, 8, 3, 0, 96) BK_(16, 8, 1, 0, 96) BK_(16, 8, 3, 0, 0) BK_(16, 12, 3, 0, 0) BK_(16, 8, 3, 1, 96)
    timeit("p16: grid 1280 (one tile each), 512 thr, 200KB", [&] { k_tile_p16<0><<<1280, 512, 200 * 1024>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, total_tiles, 0); });
    timeit("p16: grid 1280 (one tile each), 512 thr, 96KB", [&] { k_tile_p16<0><<<1280, 512, 98304>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, total_tiles, 0); });
    timeit("p16: grid 296 persistent, 512 thr, 96KB (2/SM)", [&] { k_tile_p16<0><<<296, 512, 98304>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, total_tiles, 0); });
    timeit("p16: grid 148 persistent, 512 thr, 96KB", [&] { k_tile_p16<0><<<148, 512, 98304>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, total_tiles, 0); });
    timeit("p16: grid 640 persistent, 512 thr, 0KB", [&] { k_tile_p16<0><<<640, 512, 0>>>(P, M, V, S, rows, ld, BN, tiles_m, tiles_n, ms, h, total_tiles, 0); });
  }
  for (int RN : {16}) {
    char nm[128];
    snprintf(nm, sizeof(nm), "rows: %d full rows per CTA, NB=8", RN);
    timeit(nm, [&] { k_rows<8><<<dim3((rows + RN - 1) / RN, E), 256>>>(P, M, V, S, rows, ld, RN, ms, h); });
    snprintf(nm, sizeof(nm), "rows: %d full rows per CTA, NB=16", RN);
    timeit(nm, [&] { k_rows<16><<<dim3((rows + RN - 1) / RN, E), 256>>>(P, M, V, S, rows, ld, RN, ms, h); });
  }
  return 0;
}
