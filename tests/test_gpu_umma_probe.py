"""Pins the tcgen05 shared-memory operand conventions (awb_tc.cuh) on the device: every (A, B) layout
combination the fused kernels use is run through awb_debug_umma_probe and compared with a matmul of the
same fp16 operands accumulated in fp32."""
import ctypes as C

import pytest
import torch

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def built():
    entry.build()


def pack(mat: torch.Tensor) -> torch.Tensor:
    """[R, Cc] fp16 -> tile[chunk][row][8] bytes (Cc multiple of 8)."""
    R, Cc = mat.shape
    return mat.reshape(R, Cc // 8, 8).permute(1, 0, 2).contiguous()


def probe(a_tile, b_tile, N, K, a_mn, b_mn, a_desc, b_desc):
    from awesome_b200 import _lib
    lib = _lib.load()
    a = a_tile.to(DEV).contiguous()
    b = b_tile.to(DEV).contiguous()
    D = torch.zeros((128, N), dtype=torch.float32, device=DEV)
    ad = (C.c_uint32 * 4)(*a_desc)
    bd = (C.c_uint32 * 4)(*b_desc)
    _lib.check(lib.awb_debug_umma_probe(a.data_ptr(), a.numel() * 2, b.data_ptr(), b.numel() * 2, D.data_ptr(),
                                        N, K, a_mn, b_mn, ad, bd, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return D.cpu()


def close(D, ref):
    torch.testing.assert_close(D, ref, rtol=1e-4, atol=1e-3)


def test_forward_layout_kmajor_kmajor():
    """D[px,out] = sum_in A[px,in] W[out,in]: A K-major (R=128), B K-major (R=144)."""
    torch.manual_seed(0)
    A = torch.randn(128, 144).half()
    W = torch.randn(144, 144).half()
    D = probe(pack(A), pack(W), N=144, K=144, a_mn=0, b_mn=0,
              a_desc=(0, 128 * 16, 128, 2 * 128 * 16), b_desc=(0, 144 * 16, 128, 2 * 144 * 16))
    close(D, A.float() @ W.float().T)


def test_dgrad_layout_kmajor_mnmajor():
    """D[px,in] = sum_out d[px,out] W[out,in]: same W bytes read as an MN-major B operand."""
    torch.manual_seed(1)
    d = torch.randn(128, 144).half()
    W = torch.randn(144, 144).half()
    D = probe(pack(d), pack(W), N=144, K=144, a_mn=0, b_mn=1,
              a_desc=(0, 128 * 16, 128, 2 * 128 * 16), b_desc=(0, 128, 144 * 16, 256))
    close(D, d.float() @ W.float())


@pytest.mark.parametrize("window", [0, 16])
def test_wgrad_layout_mnmajor_mnmajor_windows(window):
    """D[out_m,in] = sum_px d[px,out_m] Z[px,in] for the 128-row window of d's columns starting at `window`."""
    torch.manual_seed(2)
    d = torch.randn(128, 144).half()
    Z = torch.randn(128, 144).half()
    D = probe(pack(d), pack(Z), N=144, K=128, a_mn=1, b_mn=1,
              a_desc=((window // 8) * 128 * 16, 128, 128 * 16, 256), b_desc=(0, 128, 128 * 16, 256))
    close(D, d.float()[:, window:window + 128].T @ Z.float())


def test_narrow_n_block_of_columns():
    """D[in_m, j] = sum_px Z[px,in_m] d[px,128+j], N=16: B operand = last two column chunks of the d tile."""
    torch.manual_seed(3)
    d = torch.randn(128, 144).half()
    Z = torch.randn(128, 144).half()
    D = probe(pack(Z), pack(d), N=16, K=128, a_mn=1, b_mn=1,
              a_desc=(0, 128, 128 * 16, 256), b_desc=(16 * 128 * 16, 128, 128 * 16, 256))
    close(D, Z.float()[:, :128].T @ d.float()[:, 128:144])


def test_input_layer_k16():
    """D[px,out] = X[px,0:16] Win[out,0:16]: single K=16 step."""
    torch.manual_seed(4)
    X = torch.randn(128, 16).half()
    W = torch.randn(144, 16).half()
    D = probe(pack(X), pack(W), N=144, K=16, a_mn=0, b_mn=0,
              a_desc=(0, 128 * 16, 128, 0), b_desc=(0, 144 * 16, 128, 0))
    close(D, X.float() @ W.float().T)
