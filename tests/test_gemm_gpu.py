"""GPU bring-up tests for the grouped tcgen05 GEMM (drvae_b200/csrc/gemm.cuh).

Each contraction mode is checked, through the C ABI, against torch.matmul on the same
bf16-rounded operands (fp32 accumulate): tensor-core mainloop, SIMT validation mainloop, ragged
row counts, split-K, and several ensemble members per launch.  Tolerance: 2e-3 relative to the
row/column scale (fp32 accumulation-order noise only; the inputs are identical bf16 values).
"""
import ctypes

import pytest
import torch

from drvae_b200 import _lib
from drvae_b200.layout import pack_c8, round_up

pytestmark = pytest.mark.gpu

NT, DX, DW = 0, 1, 2
TC, SIMT = 0, 1


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def run_gemm(impl, mode, A, B, M, N, K, BN, dyn=None, ksplit=1, variant=0):
    """A, B: lists (one per model) of logical fp32 matrices.
    NT: A [M,K], B [N,K] -> D = A B^T.   DX: A [M,K], B [K,N] (W[out=K,in=N]) -> D = A B.
    DW: A [K,M] (dY rows x out), B [K,N] (X rows x in) -> D = A^T B."""
    lib = _lib.load()
    nm = len(A)
    dev = A[0].device
    if mode == NT:
        Ac = [pack_c8(a, rcap=round_up(M, 128), fcap=round_up(K, 16)) for a in A]
        Bc = [pack_c8(b, rcap=round_up(N, BN), fcap=round_up(K, 16)) for b in B]
    elif mode == DX:
        Ac = [pack_c8(a, rcap=round_up(M, 128), fcap=round_up(K, 16)) for a in A]
        Bc = [pack_c8(b, rcap=round_up(K, 16), fcap=round_up(N, 8)) for b in B]
    else:
        Ac = [pack_c8(a, rcap=round_up(K, 128), fcap=round_up(M, 8)) for a in A]
        Bc = [pack_c8(b, rcap=round_up(K, 128), fcap=round_up(N, 8)) for b in B]
    Ast = torch.stack(Ac).contiguous()
    Bst = torch.stack(Bc).contiguous()
    D = torch.zeros(nm, M, N, dtype=torch.float32, device=dev)
    dyn_t = None
    if dyn is not None:
        dyn_t = torch.tensor(dyn, dtype=torch.int32, device=dev)
    Kc = round_up(K, 16) if mode != DW else K
    st = lib.drvae_debug_gemm(
        impl, mode, _ptr(Ast), Ast.shape[2], Ast.shape[1], Ast[0].numel(), _ptr(Bst), Bst.shape[2], Bst.shape[1],
        Bst[0].numel(), _ptr(D), N, M * N, M, N, Kc, BN, _ptr(dyn_t) if dyn_t is not None else None, ksplit,
        variant, nm, None)
    _lib.check(st, "debug_gemm")
    torch.cuda.synchronize()
    if ksplit > 1:  # the gradient epilogue stores D^T ([out][in], the reference weight layout)
        D = D.view(nm, N, M).transpose(1, 2).contiguous()
    return D


def ref_gemm(mode, a, b):
    a = a.to(torch.bfloat16).float()
    b = b.to(torch.bfloat16).float()
    if mode == NT:
        return a @ b.t()
    if mode == DX:
        return a @ b
    return a.t() @ b


def make(mode, M, N, K, nm, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A, B = [], []
    for _ in range(nm):
        if mode == NT:
            a, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
        elif mode == DX:
            a, b = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g)
        else:
            a, b = torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)
        A.append(a.cuda())
        B.append(b.cuda())
    return A, B


def check(D, mode, A, B, dyn=None, tag=""):
    for m in range(len(A)):
        a, b = A[m], B[m]
        ref = ref_gemm(mode, a, b)
        got = D[m]
        if dyn is not None:
            if mode == DW:
                ref = ref_gemm(mode, a[:dyn[m]], b[:dyn[m]])
            else:
                ref = ref[:dyn[m]]
                assert torch.all(got[dyn[m]:] == 0), tag + " rows beyond the dynamic count were written"
                got = got[:dyn[m]]
        if ref.numel() == 0:
            continue
        scale = ref.abs().max().item() + 1e-6
        err = (got - ref).abs().max().item() / scale
        assert err < 2e-3, "%s model %d: max rel err %.3e" % (tag, m, err)


SHAPES = [
    # mode, M, N, K, BN
    (NT, 128, 16, 16, 16),
    (NT, 225, 800, 978, 208),
    (NT, 600, 2048, 608, 256),
    (NT, 300, 200, 100, 208),
    (DX, 128, 16, 16, 16),
    (DX, 600, 600, 2048, 208),
    (DX, 400, 100, 208, 112),
    (DW, 128, 16, 16, 16),
    (DW, 2048, 600, 600, 208),
    (DW, 800, 978, 225, 256),
    (DW, 200, 100, 300, 112),
]


@pytest.mark.parametrize("impl", [SIMT, TC], ids=["simt", "tc"])
@pytest.mark.parametrize("mode,M,N,K,BN", SHAPES)
def test_gemm_matches_torch(impl, mode, M, N, K, BN):
    A, B = make(mode, M, N, K, 2, seed=M * 7 + N * 3 + K + mode)
    if mode == DW:
        # rows beyond K must be zero padding: pack_c8 pads with zeros
        pass
    D = run_gemm(impl, mode, A, B, M, N, K, BN)
    check(D, mode, A, B, tag="impl=%d mode=%d %dx%dx%d" % (impl, mode, M, N, K))


@pytest.mark.parametrize("impl", [SIMT, TC], ids=["simt", "tc"])
def test_gemm_dynamic_rows(impl):
    M, N, K, BN = 600, 208, 112, 208
    A, B = make(NT, M, N, K, 3, seed=5)
    dyn = [600, 130, 0]
    D = run_gemm(impl, NT, A, B, M, N, K, BN, dyn=dyn)
    check(D, NT, A, B, dyn=dyn, tag="dyn-NT")
    # dW with a dynamic contraction length; rows beyond the count are zero in the operands
    M, N, K = 208, 112, 600
    A, B = make(DW, M, N, K, 3, seed=6)
    dyn = [600, 77, 0]
    for m, d in enumerate(dyn):
        A[m][d:] = 0
        B[m][d:] = 0
    D = run_gemm(impl, DW, A, B, M, N, K, 112, dyn=dyn)
    check(D, DW, A, B, dyn=dyn, tag="dyn-DW")


def test_gemm_splitk_atomic():
    M, N, K = 800, 978, 4096
    A, B = make(DW, M, N, K, 1, seed=9)
    D = run_gemm(TC, DW, A, B, M, N, K, 256, ksplit=8)
    check(D, DW, A, B, tag="splitk")
