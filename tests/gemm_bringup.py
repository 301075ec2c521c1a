"""First-contact script for the tcgen05 GEMM (run under gpurun, each variant in its own process
so a trapped kernel cannot poison the next one).  Prints max relative error per mode/variant."""
import subprocess
import sys

CASE = r'''
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_gemm_gpu import *
impl, mode, variant = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
shapes = {0: [(128, 16, 16, 16), (128, 64, 64, 64), (256, 208, 112, 208), (225, 800, 978, 208)],
          1: [(128, 16, 16, 16), (128, 64, 64, 64), (256, 112, 208, 112), (600, 600, 2048, 208)],
          2: [(128, 16, 16, 16), (128, 64, 64, 64), (208, 112, 300, 112), (2048, 600, 600, 208)]}[mode]
for (M, N, K, BN) in shapes:
    A, B = make(mode, M, N, K, 1, seed=1)
    D = run_gemm(impl, mode, A, B, M, N, K, BN, variant=variant)
    ref = ref_gemm(mode, A[0], B[0])
    err = ((D[0] - ref).abs().max() / ref.abs().max()).item()
    print("impl=%d mode=%d variant=%d %dx%dx%d BN=%d: rel err %.3e %s" % (impl, mode, variant, M, N, K, BN, err, "OK" if err < 2e-3 else "BAD"), flush=True)
'''

for impl in (1, 0):
    for mode in (0, 1, 2):
        for variant in ((0,) if impl == 1 else (0, 1)):
            r = subprocess.run([sys.executable, "-c", CASE, str(impl), str(mode), str(variant)], capture_output=True,
                               text=True, timeout=300)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                print("impl=%d mode=%d variant=%d FAILED rc=%d: %s" % (impl, mode, variant, r.returncode,
                                                                      r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ""))
            sys.stdout.flush()
