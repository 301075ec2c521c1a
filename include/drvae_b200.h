/* drvae_b200 — C ABI of the B200-native DrVAE / PertVAE / VFAE training step.
 *
 * (bring-up revision: GEMM validation entry only; the step API follows)
 */
#ifndef DRVAE_B200_H
#define DRVAE_B200_H
#ifdef __cplusplus
extern "C" {
#endif
const char* drvae_last_error(void);
int drvae_debug_gemm(int impl, int mode, const void* A, int a_rcap, int a_nchunks, long long a_ms,
                     const void* B, int b_rcap, int b_nchunks, long long b_ms, float* D, int ldd,
                     long long d_ms, int M, int N, int K, int BN, const int* dyn_dev, int ksplit,
                     int desc_variant, int n_models, void* stream);
#ifdef __cplusplus
}
#endif
#endif
