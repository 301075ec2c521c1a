// Data-parallel exchange over NVLink peer memory (BASELINE configs[3]; SURVEY.md §8(e): one model, rows of a large
// minibatch sharded over the GPUs of one box, gradients summed, replicated Adam).
//
// The reference has no distributed code.  Design: every rank's flat gradient lives in a symmetric allocation all ranks
// have mapped (the caller passes the peer pointers: torch.distributed._symmetric_memory, cudaIpc, ...), and the
// all-reduce is FUSED INTO THE OPTIMIZER: adam_peer_kernel reads element i of every rank's gradient over NVLink
// (or one switch-reduced value through the NVLS multicast mapping: multimem.ld_reduce), sums in rank order (bit-identical
// on all ranks) and applies Adam + the shadow refresh in the same pass.  No NCCL call, no separate reduction buffer,
// no extra pass over the 9.3 MB gradient; the whole step (gradient kernels + barrier + optimizer) is captured as ONE
// CUDA graph on every rank, which is what a shard of a few hundred rows needs (the host cannot enqueue ~50 launches
// and 7 collectives in the ~0.3 ms such a shard takes: round-1 DP got slower from 4 to 8 GPUs).
//
// Control words (int64, in each rank's symmetric control block):
//   [4 r .. 4 r + 3]   counts slot of rank r: {N, Np, Nlab, tag}     (dp_counts_kernel: batch-global normalisers)
//   [64 + r]           "gradient of rank r complete" flag = step tag (dp_barrier_kernel)
// Tags increase monotonically with the step, nothing is ever reset.  The counts exchange at the head of step s+1 doubles
// as the "every rank has finished reading my gradient of step s" barrier (it follows the optimizer in stream order).
#pragma once

#include "optim.cuh"

namespace drvae {

constexpr int DP_MAX_RANKS = 16;
constexpr int DP_CTL_WORDS = 128;  // int64 words per rank

struct DpPeers {
  int rank, world;
  const float* grads[DP_MAX_RANKS];  // every rank's gradient vector (this rank's own included), peer-mapped
  long long* ctl[DP_MAX_RANKS];      // every rank's control block
  const float* grads_mc;             // NVLS multicast mapping of the gradient vectors (null: plain peer loads)
};

__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// bounded spin on a control word of THIS rank's block: a peer that never arrives becomes a trapped kernel, not a hang
__device__ __forceinline__ void wait_tag(const long long* p, long long tag, DebugWord* dbg, unsigned code) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) < tag) {
    __nanosleep(100);
    if (clock64() - t0 > 20000000000LL) {  // ~10 s
      if (dbg) {
        dbg->code = code;
        dbg->info[0] = threadIdx.x;
        __threadfence_system();
      }
      __trap();
    }
  }
}

// Batch-global normalisers: every rank posts {N, Np, Nlab} of its shard to all peers and sums what it received.
// out3: int64 [3] in this rank's memory (drvae_hparams_t.global_counts_dev).  grid 1, block 32.
__global__ void dp_counts_kernel(DpPeers pe, long long N, long long Np, long long Nlab, long long tag, long long* out3, DebugWord* dbg) {
  const int r = threadIdx.x;
  if (r < pe.world) {
    long long* slot = pe.ctl[r] + 4 * pe.rank;
    slot[0] = N, slot[1] = Np, slot[2] = Nlab;
    __threadfence_system();
    st_release_sys(slot + 3, tag);
    wait_tag(pe.ctl[pe.rank] + 4 * r + 3, tag, dbg, 0xD1000000u);
  }
  __syncwarp();
  if (r < 3) {
    long long s = 0;
    for (int k = 0; k < pe.world; ++k) s += pe.ctl[pe.rank][4 * k + r];
    out3[r] = s;
  }
}

// "my gradient is complete" -> every peer; wait for every peer's.  The tag comes from the per-step scalars in device
// memory, so the kernel can sit inside a replayed CUDA graph.  grid 1, block 32.
__global__ void dp_barrier_kernel(DpPeers pe, const StepDyn* dyn, DebugWord* dbg) {
  const int r = threadIdx.x;
  const long long tag = (long long)dyn->noise_step + 1;
  if (r < pe.world) {
    __threadfence_system();  // this rank's gradient (earlier kernels of the stream) before the flag
    st_release_sys(pe.ctl[r] + 64 + pe.rank, tag);
    wait_tag(pe.ctl[pe.rank] + 64 + r, tag, dbg, 0xD2000000u);
  }
}

__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// one value reduced inside the NVSwitch over all ranks' copies (NVLS): same bits on every rank
__device__ __forceinline__ float4 ld_reduce_mc_f4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// All-reduce fused into Adam: thread = 4 consecutive parameters.  `nred` floats are reduced: the P gradient elements
// followed by the 8 additive loss shares of the step (written to losses_out, not optimised).
// grid ceil(nred / 1024), block 256
__global__ void __launch_bounds__(256) adam_peer_kernel(AdamArgs a, DpPeers pe, int nred, float* losses_out) {
  const int i4 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (i4 >= nred) return;
  float4 g;
  if (pe.grads_mc) {
    g = ld_reduce_mc_f4(pe.grads_mc + i4);
  } else {
    g = ld_peer_f4(pe.grads[0] + i4);
    for (int r = 1; r < pe.world; ++r) {
      const float4 x = ld_peer_f4(pe.grads[r] + i4);
      g.x += x.x, g.y += x.y, g.z += x.z, g.w += x.w;
    }
  }
  if (i4 >= a.P) {  // loss shares
    if (i4 < a.P + 8) *reinterpret_cast<float4*>(losses_out + (i4 - a.P)) = g;
    return;
  }
  float* p = a.params.p;
  float4 pv = *reinterpret_cast<float4*>(p + i4), m1 = *reinterpret_cast<float4*>(a.m.p + i4), v1 = *reinterpret_cast<float4*>(a.v.p + i4);
  const AdamHyper h = *a.h;
  bool skip_range = false;
  if (a.skip_hi > a.skip_lo) skip_range = (a.dyn->s.gN > 0 ? a.dyn->s.gNp : a.counts[CNT_NP]) == 0;
  if (!(skip_range && i4 >= a.skip_lo && i4 < a.skip_hi)) {
    adam_update(g.x, pv.x, m1.x, v1.x, h);
    adam_update(g.y, pv.y, m1.y, v1.y, h);
    adam_update(g.z, pv.z, m1.z, v1.z, h);
    adam_update(g.w, pv.w, m1.w, v1.w, h);
    *reinterpret_cast<float4*>(p + i4) = pv;
    *reinterpret_cast<float4*>(a.m.p + i4) = m1;
    *reinterpret_cast<float4*>(a.v.p + i4) = v1;
  }
  // kernel-facing copies of the four elements
  const float vals[4] = {pv.x, pv.y, pv.z, pv.w};
  int si = seg_find(a.segs, a.nseg, i4);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int idx = i4 + u;
    if (si + 1 < a.nseg && idx >= a.segs[si + 1].off) ++si;
    if (idx >= a.segs[si].off && idx - a.segs[si].off < a.segs[si].rows * a.segs[si].ld) write_derived(a.segs[si], idx, vals[u], a.shadow.p, a.derived.p);
  }
}

}  // namespace drvae
