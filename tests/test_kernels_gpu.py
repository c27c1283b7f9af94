"""GPU unit tests of the two tensor-core kernels through the C ABI debug entries.

The checker here is a plain PyTorch fp32 restatement of the same op on the same bf16-rounded inputs (floating-point
kernels keep a torch fp32 reference; the end-to-end parity tests use oracle/)."""
import math

import pytest
import torch

from mode_diffusion_policy_b200 import _lib

pytestmark = pytest.mark.gpu


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _stream():
    return torch.cuda.current_stream().cuda_stream


def run_gemm(M, N, K, epi, seed=0, pair=False, stream_k=False, bn=256, skip_b=False):
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    Mp = (M + 255) // 256 * 256
    A = torch.zeros(Mp, K, dtype=torch.bfloat16)
    A[:M] = (torch.randn(M, K, generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, generator=g) * 0.1
    A, W, bias = A.cuda(), W.cuda(), bias.cuda()
    ref = A[:M].float() @ W.float().t()
    resid = None
    if epi == 0:
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        want = (ref + bias).bfloat16().float()
    elif epi == 1:
        resid = torch.randn(M, N, generator=g).cuda()
        out = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
        want = resid + ref
    elif epi == 2:
        out = torch.full((M, N // 2), float("nan"), dtype=torch.bfloat16, device="cuda")
        z = (ref + bias).view(M, N // 256, 2, 128)
        want = (z[:, :, 0] * torch.nn.functional.silu(z[:, :, 1])).reshape(M, N // 2).bfloat16().float()
    elif epi == 3:
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        want = ref.bfloat16().float()
    else:
        out = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
        want = ref
    _lib.check(lib.mode_debug_gemm(_ptr(A), _ptr(W), _ptr(bias), _ptr(resid), _ptr(out), M, N, K, epi | (0x100 if pair else 0) | (0x200 if stream_k else 0) | ((bn // 16) << 16 if bn != 256 else 0) | (0x400 if skip_b else 0), _stream()))
    torch.cuda.synchronize()
    return out.float(), want


@pytest.mark.parametrize("epi", [4, 0, 1, 2, 3])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (256, 512, 1024), (448, 1024, 1024),
                                   (100, 256, 512), (3584, 1024, 4096), (1000, 3072, 1024),
                                   (3584, 3072, 1024), (5000, 2048, 512), (7168, 1024, 4096), (3584, 8192, 1024)])
@pytest.mark.parametrize("mode", ["cta1", "cta2", "cta2_streamk"])
def test_gemm_matches_fp32_reference(M, N, K, epi, mode):
    got, want = run_gemm(M, N, K, epi, pair=mode != "cta1", stream_k=mode == "cta2_streamk")
    assert torch.isfinite(got).all()
    err = (got - want).abs()
    scale = want.abs().max().item() + 1e-6
    noise = 2e-5 * scale * math.sqrt(K / 64)  # fp32 accumulation-order noise
    if epi in (1, 4):
        assert err.max().item() <= noise, (err.max().item(), scale)
    else:  # bf16 outputs: at most one bf16 ulp (plus the fp32 noise floor) from the rounded reference
        assert (err <= want.abs() * 2 ** -7 + noise).all(), err.max().item()
        assert (err > 0).float().mean().item() < 0.02  # almost all elements round identically


@pytest.mark.parametrize("epi", [0, 1, 3, 4])
@pytest.mark.parametrize("bn", [208, 160, 128, 240, 176, 96, 64])
@pytest.mark.parametrize("M,N,K", [(3584, 1024, 1024), (3584, 3072, 1024), (7168, 1024, 4096), (300, 1024, 512), (1000, 3072, 256)])
def test_gemm_tile_widths_are_bit_identical(M, N, K, epi, bn):
    """CTA-pair kernel with `bn`-wide tiles (wave filling, engine.cu choose_bn): the column tiling changes which tile owns
    an output column — overhanging last tile, 16-column register-stored tails — but not the K order of any element, so
    the result must equal the 256-wide tiling bit for bit."""
    got, want = run_gemm(M, N, K, epi, pair=True, bn=bn)
    base, _ = run_gemm(M, N, K, epi, pair=True)
    assert torch.isfinite(got).all()
    assert torch.equal(got, base), (got - base).abs().max().item()
    del want


def attention_reference(qkv, gq, gk, B, T, H, Dh, eps):
    d = H * Dh
    x = qkv.float().view(B, T, 3, H, Dh)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)

    def rms(t, g):
        n = t.norm(dim=-1, keepdim=True) * Dh ** -0.5
        return (t / n.clamp(min=eps) * g).bfloat16().float()

    q, k = rms(q, gq), rms(k, gk)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(Dh)
    mask = torch.ones(T, T, dtype=torch.bool, device=qkv.device).tril()
    s = s.masked_fill(~mask, float("-inf"))
    m = s.max(dim=-1, keepdim=True).values
    p = torch.exp(s - m)
    o = (p.bfloat16().float() @ v) / p.sum(dim=-1, keepdim=True)
    return o.transpose(1, 2).reshape(B * T, d)


@pytest.mark.parametrize("B,T,H,Dh", [(3, 14, 8, 128), (2, 32, 8, 64), (5, 14, 4, 64), (2, 16, 8, 32), (1, 50, 2, 128),
                                      (256, 14, 8, 128)])
def test_attention_matches_reference(B, T, H, Dh):
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + T)
    d = H * Dh
    qkv = torch.randn(B * T, 3 * d, generator=g).bfloat16().cuda()
    gq = (1 + 0.1 * torch.randn(Dh, generator=g)).cuda()
    gk = (1 + 0.1 * torch.randn(Dh, generator=g)).cuda()
    out = torch.full((B * T, d), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.mode_debug_attention(_ptr(qkv), _ptr(gq), _ptr(gk), _ptr(out), B, T, H, Dh, 1e-6, _stream()))
    torch.cuda.synchronize()
    want = attention_reference(qkv, gq, gk, B, T, H, Dh, 1e-6)
    got = out.float()
    assert torch.isfinite(got).all()
    rel = (got - want).norm() / want.norm()
    assert rel.item() < 4e-3, rel.item()  # bf16 output rounding dominates (2^-9 rms)
    assert (got - want).abs().max().item() < 0.05


@pytest.mark.parametrize("rows,n_out,k_out", [(64, 128, 256), (256, 128, 256), (1792, 1024, 1024), (3584, 3072, 1024),
                                              (3584, 1024, 4096), (448, 256, 2048)])
def test_wgrad_gemm_matches_fp32_reference(rows, n_out, k_out):
    """dW = dY^T X with both operands consumed MN-major from the row-major activations (csrc/gemm_wgrad.cuh)."""
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(rows + n_out)
    dy = (torch.randn(rows, n_out, generator=g) * 0.5).bfloat16().cuda()
    x = (torch.randn(rows, k_out, generator=g) / math.sqrt(rows)).bfloat16().cuda()
    out = torch.full((n_out, k_out), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(lib.mode_debug_wgrad(_ptr(dy), _ptr(x), _ptr(out), rows, n_out, k_out, 0, _stream()))
    torch.cuda.synchronize()
    want = dy.float().t() @ x.float()
    assert torch.isfinite(out).all()
    err = (out - want).abs().max().item()
    assert err <= 2e-5 * (want.abs().max().item() + 1e-6) * math.sqrt(rows / 64), err


def test_wgrad_gemm_unpacks_swiglu_rows():
    lib = _lib.load()
    rows, n_out, k_out, half = 256, 1024, 256, 512
    g = torch.Generator(device="cpu").manual_seed(5)
    dy = torch.randn(rows, n_out, generator=g).bfloat16().cuda()  # columns in the packed [128 proj | 128 gate] order
    x = torch.randn(rows, k_out, generator=g).bfloat16().cuda()
    out = torch.zeros(n_out, k_out, dtype=torch.float32, device="cuda")
    _lib.check(lib.mode_debug_wgrad(_ptr(dy), _ptr(x), _ptr(out), rows, n_out, k_out, half, _stream()))
    torch.cuda.synchronize()
    packed = (dy.float().t() @ x.float()).view(n_out // 256, 2, 128, k_out)  # [block, proj/gate, 128, K]
    want = torch.cat([packed[:, 0].reshape(half, k_out), packed[:, 1].reshape(half, k_out)])
    assert (out - want).abs().max().item() <= 1e-3 * want.abs().max().item()
