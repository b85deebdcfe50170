"""GPU bring-up: the tcgen05 TF32 tile (arah_umma.cuh) on its own, before anything depends on it."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('a_in_tmem', [0, 1])
@pytest.mark.parametrize('K,N', [(32, 256), (256, 256), (128, 128), (96, 256)])
def test_umma_tf32_tile_matches_matmul(K, N, a_in_tmem):
    """tcgen05 TF32 tile (descriptors, 128B swizzle, TMEM read-back) against torch fp32 matmul; tolerance = TF32 operand rounding."""
    import ctypes as C
    from arah_release_b200 import _lib
    g = torch.Generator(device='cpu').manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g).to(DEV)
    Wt = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    D = torch.zeros(128, N, device=DEV)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().arah_debug_umma_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(Wt.data_ptr()), K, N, C.c_void_p(D.data_ptr()), a_in_tmem, st))
    ref = A.double() @ Wt.double().t()
    err = (D.double() - ref).abs().max().item()
    assert err < 6e-3, err          # |a||w| K 2^-11 scale
    assert err > 0 or K == 0


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('K,N', [(64, 128), (128, 128), (128, 32), (128, 256)])
def test_umma_f16_split_tile_matches_matmul(K, N, mode):
    """tcgen05 kind::f16 tile with A in tensor memory as packed half pairs and pre-scaled hi / lo weight images
    (arah_f16x3.cuh).  mode 0: hi.hi only -> fp16 operand rounding; mode 1: the three-pass split product the persistent
    root-finding kernels use -> fp32-grade (the lo.lo term, 2^-22 relative, is dropped)."""
    import ctypes as C
    from arah_release_b200 import _lib
    g = torch.Generator(device='cpu').manual_seed(K * 1000 + N + mode)
    A = torch.rand(128, K, generator=g).to(DEV) * 2 - 0.5              # activations of O(1), both signs
    A[:, 0] = 1e-6                                                      # below fp16's normal range: absolute error must stay tiny
    Wt = (torch.randn(N, K, generator=g) / K ** 0.5 * 0.05).to(DEV)     # small weights: exercises the power-of-two pre-scale
    D = torch.zeros(128, N, device=DEV)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().arah_debug_umma_f16(C.c_void_p(A.data_ptr()), C.c_void_p(Wt.data_ptr()), K, N, C.c_void_p(D.data_ptr()), mode, st))
    ref = A.double() @ Wt.double().t()
    err = (D.double() - ref).abs().max().item()
    scale = (A.abs().double() @ Wt.abs().double().t()).max().item()     # sum |a||w|: what operand rounding multiplies
    print(f'f16 tile K={K} N={N} mode={mode}: max abs err {err:.3e}, sum|a||w| {scale:.3e}')
    if mode == 0:
        assert err < scale * 2 ** -10 and err > 0
    else:
        assert err < scale * 2e-6, (err, scale)
