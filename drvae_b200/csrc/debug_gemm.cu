// Bring-up / validation entry point: run ONE grouped GEMM (tensor-core or SIMT mainloop) on
// caller-provided chunk8 operands and return the fp32 accumulator tile matrix.  Used by
// tests/test_gemm_gpu.py to pin the UMMA descriptor encodings against a torch matmul.
#include <vector>

#include "drvae_b200.h"
#include "errors.h"
#include "gemm.cuh"

using namespace drvae;

extern "C" int drvae_debug_gemm(int impl, int mode, const void* A, int a_rcap, int a_nchunks, long long a_ms,
                                const void* B, int b_rcap, int b_nchunks, long long b_ms, float* D, int ldd,
                                long long d_ms, int M, int N, int K, int BN, const int* dyn_dev, int ksplit,
                                int desc_variant, int n_models, void* stream) {
  if (BN % 16 != 0 || BN < 16 || BN > 256) return set_error("drvae_debug_gemm: BN must be a multiple of 16 in [16,256]");
  if (mode < 0 || mode > 2) return set_error("drvae_debug_gemm: bad mode");
  cudaStream_t st = (cudaStream_t)stream;
  static DebugWord* dbg = nullptr;
  if (!dbg) {
    if (cudaMalloc(&dbg, sizeof(DebugWord)) != cudaSuccess) return set_error("cudaMalloc(debug word) failed");
    cudaMemset(dbg, 0, sizeof(DebugWord));
  }
  GemmProblem p{};
  p.A = GemmOperand{(const bf16*)A, a_ms, a_rcap, a_nchunks, 0};
  p.B = GemmOperand{(const bf16*)B, b_ms, b_rcap, b_nchunks, 0};
  p.mode = mode;
  p.M = M;
  p.N = N;
  p.K = K;
  p.dyn = dyn_dev;
  p.dyn_stride = 1;
  p.BN = BN;
  p.tiles_n = (N + BN - 1) / BN;
  p.tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  p.ksplit = ksplit < 1 ? 1 : ksplit;
  p.desc_variant = desc_variant;
  p.dbg = dbg;
  EpiParams e{};
  int epi = EPI_STORE_F32;
  if (p.ksplit > 1) {
    // exercise the atomic path through the gradient epilogue with an identity row map: the
    // epilogue stores D transposed (reference weight layout [out = D col][in = D row]), so the
    // caller's buffer holds D^T as [N][M]
    epi = EPI_GRAD;
    e.grad = D;
    e.grad_ms = d_ms;
    const int ncap = p.tiles_n * BN;
    std::vector<int> tab(3 * (size_t)ncap, -1);
    for (int s = 0; s < N; ++s) tab[s] = s * M;
    int* d_tab = nullptr;  // test-only entry point: the table is leaked on purpose (a few KB)
    if (cudaMalloc(&d_tab, tab.size() * sizeof(int)) != cudaSuccess) return set_error("cudaMalloc(table) failed");
    cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice);
    e.g_tab = d_tab;
    e.g_tab_n = ncap;
    e.g_kin = M;
    e.g_kaug = M;
  } else {
    e.out_f32 = D;
    e.out_f32_ms = d_ms;
    e.out_ld = ldd;
    e.out_row0 = 0;
    e.n_valid = N;
  }
  cudaError_t err = gemm_launch(epi, p, e, n_models, impl, st);
  if (err != cudaSuccess) return set_cuda_error("drvae_debug_gemm launch", err);
  err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) {
    DebugWord h{};
    cudaMemcpy(&h, dbg, sizeof(h), cudaMemcpyDeviceToHost);  // likely fails after a trap; best effort
    char msg[256];
    snprintf(msg, sizeof(msg), "drvae_debug_gemm: %s (debug code 0x%08x block %u,%u,%u thread %u)",
             cudaGetErrorString(err), h.code, h.info[0], h.info[1], h.info[2], h.info[3]);
    return set_error(msg);
  }
  return 0;
}
