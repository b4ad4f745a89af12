"""GPU parity tests of the individual sm_100a kernels, called through the C ABI (ctypes).

Each kernel is compared with a plain PyTorch fp32 evaluation of the same arithmetic on the same
(bf16-rounded) inputs. Tolerances are written next to each check: outputs stored as bf16 carry one
bf16 rounding (2^-9 relative), fp32 outputs only accumulation-order noise.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from climb_b200 import _lib
    return _lib


def _rel_err(got, ref):
    got = got.float()
    ref = ref.float()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def _max_err(got, ref):
    return (got.float() - ref.float()).abs().max().item()


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
GEMM_SHAPES = [
    (128, 256, 64), (128, 128, 128), (256, 256, 256), (300, 768, 768), (948, 2304, 768),
    (948, 768, 3072), (64, 1536, 768), (64, 3129, 1536), (1000, 72, 200), (130, 3072, 768),
]


@pytest.mark.parametrize("block_n", [0, 64, 128, 256])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_kmajor(M, N, K, block_n):
    L = _lib()
    torch.manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    ldc = (N + 7) // 8 * 8
    out = torch.full((M, ldc), float("nan"), device="cuda", dtype=torch.float32)
    L.gemm(a, b, out, N=N, block_n=block_n)
    ref = a.float() @ b.float().t()
    torch.cuda.synchronize()
    assert torch.isfinite(out[:, :N]).all()
    assert _rel_err(out[:, :N], ref) < 1e-5, (_rel_err(out[:, :N], ref), _max_err(out[:, :N], ref))
    if ldc > N:   # padding columns untouched
        assert torch.isnan(out[:, N:]).all()


@pytest.mark.parametrize("a_mn,b_mn", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("block_n", [64, 128, 256])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 128, 192), (768, 768, 952), (3072, 768, 504), (200, 136, 72)])
def test_gemm_mn_major(M, N, K, block_n, a_mn, b_mn):
    L = _lib()
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    a_in = a.t().contiguous() if a_mn else a      # [K, M] storage when MN-major
    b_in = b.t().contiguous() if b_mn else b
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    L.gemm(a_in, b_in, out, a_mn_major=a_mn, b_mn_major=b_mn, block_n=block_n)
    ref = a.float() @ b.float().t()
    assert _rel_err(out, ref) < 1e-5, (_rel_err(out, ref), _max_err(out, ref))


@pytest.mark.parametrize("split_k", [0, 2, 5])
def test_gemm_wgrad_splitk_accumulate(split_k):
    """dW[N_out, K_in] += dY^T X, both operands read in place (MN-major), split over tokens."""
    L = _lib()
    torch.manual_seed(3)
    tokens, n_out, k_in = 237 * 8, 768, 768
    dy = torch.randn(tokens, n_out, device="cuda").bfloat16()
    x = torch.randn(tokens, k_in, device="cuda").bfloat16()
    dw = torch.ones(n_out, k_in, device="cuda", dtype=torch.float32)
    L.gemm(dy, x, dw, a_mn_major=True, b_mn_major=True, accumulate=True, split_k=split_k)
    ref = 1.0 + dy.float().t() @ x.float()
    assert _rel_err(dw, ref) < 1e-5, _rel_err(dw, ref)


@pytest.fixture(params=[0, 1], ids=["one_cta", "cta_pair"])
def pair_mode(request):
    """Run a GEMM test on the one-CTA kernels (mode 0) and on every CTA-pair (tcgen05 cta_group::2) kernel (mode 1);
    the library default (mode 2) picks per epilogue kind."""
    from climb_b200 import _lib as lib
    old = lib.climb_gemm_pair_mode(request.param)
    yield request.param
    lib.climb_gemm_pair_mode(old)


@pytest.mark.parametrize("tokens,n_out,k_in", [(15168, 3072, 768), (15168, 768, 3072), (15168, 2304, 768), (3001, 768, 768),
                                               (12544, 768, 3072), (1100, 256, 512)])
def test_gemm_wgrad_layer_shapes(tokens, n_out, k_in, pair_mode):
    """The weight gradients of the ViLT-base Linears at the bench's token count (and a token count that is not a multiple
    of the 64-row k-block): generic split-K kernel by default, the CTA-pair (cta_group::2) wgrad kernel when enabled."""
    L = _lib()
    torch.manual_seed(tokens + n_out)
    dy = (torch.randn(tokens, n_out, device="cuda") * 0.5).bfloat16()
    x = torch.randn(tokens, k_in, device="cuda").bfloat16()
    dw = torch.full((n_out, k_in), 0.5, device="cuda", dtype=torch.float32)
    L.gemm(dy, x, dw, a_mn_major=True, b_mn_major=True, accumulate=True)
    ref = 0.5 + dy.float().t() @ x.float()
    assert _rel_err(dw, ref) < 1e-5, _rel_err(dw, ref)
    L.gemm(dy, x, dw, a_mn_major=True, b_mn_major=True, accumulate=True)          # accumulates on top
    assert _rel_err(dw, 2 * ref - 0.5) < 1e-5


@pytest.mark.parametrize("tokens,n_out,k_in", [(15168, 3072, 768), (15168, 768, 3072), (15168, 768, 768), (15168, 2304, 768),
                                               (3001, 768, 768), (30336, 768, 3072), (1100, 200, 512)])
def test_gemm_wgrad_with_bias_gradient(tokens, n_out, k_in, pair_mode):
    """colsum_a: the bias gradient (column sums of dY) produced next to dW. Inside the CTA-pair weight-gradient kernel it is
    one more MMA per k-step against a tile of ones when no pair owns more than one tile (FC1 / FC2 / O shapes at B = 64);
    multi-round shapes, ragged shapes and the one-CTA kernels fall back to the streaming pass. Both accumulate."""
    L = _lib()
    torch.manual_seed(tokens + k_in)
    dy = (torch.randn(tokens, n_out, device="cuda") * 0.5 + 0.1).bfloat16()
    x = torch.randn(tokens, k_in, device="cuda").bfloat16()
    dw = torch.full((n_out, k_in), 0.5, device="cuda", dtype=torch.float32)
    db = torch.full((n_out,), 0.25, device="cuda", dtype=torch.float32)
    L.gemm(dy, x, dw, a_mn_major=True, b_mn_major=True, accumulate=True, colsum_a=db)
    ref = 0.5 + dy.float().t() @ x.float()
    ref_b = 0.25 + dy.double().sum(0).float()
    assert _rel_err(dw, ref) < 3e-5, _rel_err(dw, ref)          # fp32 split-K atomics over up to 30 336 tokens
    assert _rel_err(db, ref_b) < 2e-6, _rel_err(db, ref_b)
    L.gemm(dy, x, dw, a_mn_major=True, b_mn_major=True, accumulate=True, colsum_a=db)          # accumulates on top
    assert _rel_err(dw, 2 * ref - 0.5) < 3e-5
    assert _rel_err(db, 2 * ref_b - 0.25) < 2e-6


@pytest.mark.parametrize("M,K,N", [(2500, 768, 2304), (15168, 768, 2304), (2400, 200, 2304), (2433, 3072, 2304),
                                   # N = 768 at the bench's M: 357 tiles = 2 rounds + 61 -> the last 61 run as 122 half-width tiles
                                   (15168, 768, 768), (15168, 3072, 768), (15104, 768, 1024), (9600, 256, 512)])
def test_gemm_fast_kinds(M, K, N, pair_mode):
    """The specialised 16-epilogue-warp kernels (N % 256 == 0, >= 148 tiles): bf16 + bias, GELU with saved
    derivative, multiply-by-aux, fp32 + bias + residual; ragged last row tile, ragged K; tail tiles at half width."""
    L = _lib()
    torch.manual_seed(M + K)
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    acc = a.float() @ b.float().t()
    # FK_BF16 (with and without bias); padding rows / columns of a wider buffer stay untouched
    buf = torch.full((M + 3, N + 8), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.gemm(a, b, buf[:M], N=N, bias=bias)
    assert _rel_err(buf[:M, :N], acc + bias) < 4e-3
    assert torch.isnan(buf[M:]).all() and torch.isnan(buf[:M, N:]).all()
    outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a, b, outb)
    assert _rel_err(outb, acc) < 4e-3
    # B operand MN-major (dgrad form)
    L.gemm(a, b.t().contiguous(), outb, b_mn_major=True)
    assert _rel_err(outb, acc) < 4e-3
    # FK_GELU_SAVE
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a, b, outb, bias=bias, epilogue=L.EPI_GELU_SAVE_GRAD, aux=aux)
    pre = (acc + bias).requires_grad_(True)
    y = torch.nn.functional.gelu(pre)
    y.sum().backward()
    assert _rel_err(outb, y.detach()) < 4e-3
    assert _rel_err(aux, pre.grad) < 4e-3
    assert _max_err(outb, y.detach()) < 2 ** -7 * max(1.0, y.abs().max().item())
    # FK_MUL_AUX
    u = torch.randn(M, N, device="cuda").bfloat16()
    L.gemm(a, b, outb, epilogue=L.EPI_MUL_AUX, aux=u)
    assert _rel_err(outb, acc * u.float()) < 4e-3
    # FK_RES_F32, also in place (C == residual)
    res = torch.randn(M, N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    L.gemm(a, b, out, bias=bias, residual=res)
    assert _rel_err(out, acc + bias + res) < 1e-5
    inplace = res.clone()
    L.gemm(a, b, inplace, residual=inplace)
    assert _rel_err(inplace, acc + res) < 1e-5
    # ... with bf16 copies of the pre-residual (aux) and final (c2) values: the adapter sites
    aux2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    c2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out2 = torch.empty(M, N, device="cuda")
    L.gemm(a, b, out2, bias=bias, residual=res, aux=aux2, c2=c2)
    assert torch.equal(out2, out)
    assert _rel_err(aux2, acc + bias) < 4e-3 and _rel_err(c2, acc + bias + res) < 4e-3
    inplace2, c2b = res.clone(), torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a, b, inplace2, residual=inplace2, c2=c2b)           # dgrad-down of an adapter: dx += dz Wd, bf16 copy refreshed
    assert torch.equal(inplace2, inplace) and _rel_err(c2b, acc + res) < 4e-3


def test_gemm_epilogues():
    L = _lib()
    torch.manual_seed(11)
    M, N, K = 500, 384, 256
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    acc = a.float() @ b.float().t() + bias
    # bias + residual, fp32 out
    out = torch.empty(M, N, device="cuda")
    L.gemm(a, b, out, bias=bias, residual=res)
    assert _rel_err(out, acc + res) < 1e-5
    # bias, bf16 out  (one bf16 rounding: 2^-9)
    outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a, b, outb, bias=bias)
    assert _max_err(outb, acc.bfloat16()) <= 2 ** -6 * acc.abs().max().item() * 2 ** -2
    assert _rel_err(outb, acc) < 4e-3
    # GELU with pre-activation saved
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a, b, outb, bias=bias, epilogue=L.EPI_GELU, aux=aux)
    assert _rel_err(aux, acc) < 4e-3
    assert _rel_err(outb, torch.nn.functional.gelu(acc)) < 4e-3
    # dGELU: out = acc * gelu'(aux)
    u = torch.randn(M, N, device="cuda").bfloat16()
    L.gemm(a, b, outb, epilogue=L.EPI_DGELU, aux=u)
    uf = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uf).sum().backward()
    assert _rel_err(outb, (acc - bias) * uf.grad) < 4e-3
    # swish / dswish / relu / drelu / tanh
    L.gemm(a, b, out, bias=bias, epilogue=L.EPI_SWISH)
    assert _rel_err(out, torch.nn.functional.silu(acc)) < 1e-5
    L.gemm(a, b, out, epilogue=L.EPI_DSWISH, aux=u)
    uf = u.float().requires_grad_(True)
    torch.nn.functional.silu(uf).sum().backward()
    assert _rel_err(out, (acc - bias) * uf.grad) < 1e-5
    L.gemm(a, b, out, bias=bias, epilogue=L.EPI_RELU)
    assert _rel_err(out, torch.relu(acc)) < 1e-5
    L.gemm(a, b, out, epilogue=L.EPI_DRELU, aux=u)
    assert _rel_err(out, (acc - bias) * (u.float() > 0)) < 1e-5
    L.gemm(a, b, out, bias=bias, epilogue=L.EPI_TANH)
    assert _max_err(out, torch.tanh(acc)) < 1e-5
    # bf16 copy of the final value (after residual) next to the fp32 output, pre-residual aux
    c2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a, b, out, bias=bias, residual=res, c2=c2, aux=aux)
    assert _rel_err(out, acc + res) < 1e-5
    assert _rel_err(c2, acc + res) < 4e-3
    assert _rel_err(aux, acc) < 4e-3
    # GELU with the derivative saved, then the multiply-by-aux backward epilogue with fused column sums
    gaux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a, b, outb, bias=bias, epilogue=L.EPI_GELU_SAVE_GRAD, aux=gaux)
    accg = acc.clone().requires_grad_(True)
    torch.nn.functional.gelu(accg).sum().backward()
    assert _rel_err(outb, torch.nn.functional.gelu(acc)) < 4e-3
    assert _rel_err(gaux, accg.grad) < 4e-3
    cs = torch.ones(N, device="cuda")
    L.gemm(a, b, outb, epilogue=L.EPI_MUL_AUX, aux=gaux, colsum=cs)
    ref_mul = (acc - bias) * gaux.float()
    assert _rel_err(outb, ref_mul) < 4e-3
    assert _rel_err(cs, 1.0 + outb.float().sum(0)) < 1e-4
    # alpha
    L.gemm(a, b, out, alpha=0.25)
    assert _rel_err(out, 0.25 * (acc - bias)) < 1e-5


def test_gemm_large_persistent():
    """Config-2 sized forward GEMM (B=64 sequences): many tiles per CTA, both accumulator stages."""
    L = _lib()
    torch.manual_seed(5)
    M, N, K = 64 * 237, 3072, 768
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    L.gemm(a, b, out)
    ref = a.float() @ b.float().t()
    assert _rel_err(out, ref) < 4e-3


def test_gemm_rejects_bad_args():
    L = _lib()
    a = torch.zeros(128, 60, device="cuda", dtype=torch.bfloat16)   # K*2 not a multiple of 16
    b = torch.zeros(128, 60, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(128, 128, device="cuda")
    with pytest.raises(L.ClimbError):
        L.gemm(a, b, out)


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def _attn_ref(qkv, key_bias, B, L_, H, scale):
    q, k, v = qkv.float().view(B, L_, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    if key_bias is not None:
        s = s + key_bias[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    ctx = (p @ v).permute(0, 2, 1, 3).reshape(B, L_, H * 64)
    return ctx, torch.logsumexp(s, dim=-1)


@pytest.mark.parametrize("B,L_,H,masked", [(2, 237, 12, False), (3, 237, 12, True), (2, 13, 2, True),
                                           (1, 64, 1, False), (2, 200, 3, True), (1, 281, 2, True),
                                           (2, 256, 2, True), (2, 128, 1, False), (3, 129, 2, True), (5, 1, 1, False),
                                           # more (b, h) items than SMs: the persistent kernels walk 2-5 items per CTA
                                           (40, 237, 12, True), (64, 237, 12, False), (30, 100, 12, True), (26, 129, 7, True)])
def test_attention_fwd_bwd(B, L_, H, masked):
    L = _lib()
    torch.manual_seed(B * 100 + L_)
    qkv = torch.randn(B, L_, 3 * H * 64, device="cuda").bfloat16()
    key_bias = None
    if masked:
        lens = torch.randint(max(1, L_ // 2), L_ + 1, (B,), device="cuda")
        mask = (torch.arange(L_, device="cuda")[None, :] < lens[:, None]).float()
        key_bias = (1.0 - mask) * -10000.0
    scale = 1.0 / math.sqrt(64)
    ctx, lse = L.attention_fwd(qkv, key_bias, B, L_, H, scale)
    qkv_ref = qkv.float().requires_grad_(True)
    ctx_ref, lse_ref = _attn_ref(qkv_ref, key_bias, B, L_, H, scale)
    # ctx is stored in bf16 and P is rounded to bf16 before P.V: 2^-8 relative
    assert _rel_err(ctx, ctx_ref) < 6e-3, _rel_err(ctx, ctx_ref)
    assert _max_err(lse, lse_ref) < 2e-3, _max_err(lse, lse_ref)
    dctx = torch.randn(B, L_, H * 64, device="cuda").bfloat16()
    cs = torch.zeros(3 * H * 64, device="cuda")
    dqkv = L.attention_bwd(qkv, key_bias, ctx, dctx, lse, B, L_, H, scale, colsum=cs)
    # q / k / v bias gradients = column sums of dq / dk / dv over all tokens, by the kernel's own definitions
    # (attention_tc.cu): q = sums of the bf16 dq it stores; k = 0 analytically (sum_keys dS = 0: nothing is added);
    # v = sum_q dO (softmax rows sum to one). Each is exact up to fp32 summation order ...
    d = H * 64
    cs_ref = dqkv.float().sum((0, 1))
    scale_q = max(1.0, cs_ref[:d].abs().max().item())
    if L_ <= 256:
        assert (cs[:d] - cs_ref[:d]).abs().max().item() <= 2e-4 * scale_q
        assert cs[d:2 * d].abs().max().item() == 0.0
        dsum = dctx.float().sum((0, 1))
        assert (cs[2 * d:] - dsum).abs().max().item() <= 2e-4 * max(1.0, dsum.abs().max().item())
    # ... and all agree with the sums of the stored bf16 gradients up to their rounding noise (a random walk over the
    # B * L rows: 2^-9 relative per element) plus the bf16 rounding of P (rows of P sum to 1 +- 2^-9 / sqrt(L))
    noise = 2 ** -8 * math.sqrt(B * L_) * max(1.0, dqkv.float().abs().max().item()) / 4
    assert (cs - cs_ref).abs().max().item() <= 2e-3 * max(1.0, cs_ref.abs().max().item()) + noise
    ctx_ref.backward(dctx.float())
    ref = qkv_ref.grad.view(B, L_, 3, H * 64)
    got = dqkv.float().view(B, L_, 3, H * 64)
    gscale = ref.norm().item()
    for i, name in enumerate("qkv"):
        if ref[:, :, i].norm().item() < 1e-6 * gscale:      # analytically zero (e.g. dq, dk when L == 1)
            assert got[:, :, i].norm().item() < 1e-3 * gscale, name
            continue
        e = _rel_err(got[:, :, i], ref[:, :, i])
        assert e < 1.5e-2, (name, e)


# ------------------------------------------------------------------------------------------------
# layernorm
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,d,eps", [(1000, 768, 1e-12), (37, 128, 1e-12), (64, 1536, 1e-5), (9, 256, 1e-5)])
@pytest.mark.parametrize("act", [0, 1])
def test_layernorm_fwd_bwd(rows, d, eps, act):
    L = _lib()
    torch.manual_seed(rows + d)
    x = torch.randn(rows, d, device="cuda") * 2 + 0.5
    g = torch.randn(d, device="cuda")
    b = torch.randn(d, device="cuda")
    yb, yf, mean, rstd = L.layernorm_fwd(x, g, b, eps, out_bf16=True, out_f32=True, act=act)
    xr = x.clone().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), gr, br, eps)
    if act:
        ref = torch.nn.functional.gelu(ref)
    assert _max_err(yf, ref) < 2e-5 * max(1.0, ref.abs().max().item())
    assert _rel_err(yb, ref) < 4e-3
    dy = torch.randn(rows, d, device="cuda")
    dres = torch.randn(rows, d, device="cuda")
    ref.backward(dy)
    dx = torch.empty(rows, d, device="cuda")
    dxb = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    dg = torch.zeros(d, device="cuda")
    db = torch.zeros(d, device="cuda")
    L.layernorm_bwd(dy, x, g, b, mean, rstd, dres=dres, dx_f32=dx, dx_bf16=dxb, dgamma=dg, dbeta=db, act=act)
    assert _rel_err(dx, xr.grad + dres) < 2e-5, _rel_err(dx, xr.grad + dres)
    assert _rel_err(dxb, xr.grad + dres) < 4e-3
    assert _rel_err(dg, gr.grad) < 2e-5, _rel_err(dg, gr.grad)
    assert _rel_err(db, br.grad) < 2e-5
    if d <= 1024:
        # fused column sums of the output dx (= bias gradient of the Linear feeding this LayerNorm), accumulating
        dx3 = torch.empty(rows, d, device="cuda")
        cs = torch.ones(d, device="cuda")
        dg3, db3 = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
        L.layernorm_bwd(dy, x, g, b, mean, rstd, dres=dres, dx_f32=dx3, dgamma=dg3, dbeta=db3, act=act, dx_colsum=cs)
        assert torch.equal(dx3, dx) and _rel_err(dg3, gr.grad) < 2e-5
        assert _rel_err(cs, 1.0 + (xr.grad + dres).sum(0)) < 2e-5, _rel_err(cs, 1.0 + (xr.grad + dres).sum(0))
    # bf16 dy path
    dx2 = torch.empty(rows, d, device="cuda")
    L.layernorm_bwd(dy.bfloat16(), x, g, b, mean, rstd, dx_f32=dx2, act=act)
    xr.grad = None
    ref2 = torch.nn.functional.layer_norm(xr, (d,), gr, br, eps)
    if act:
        ref2 = torch.nn.functional.gelu(ref2)
    ref2.backward(dy.bfloat16().float())
    assert _rel_err(dx2, xr.grad) < 2e-5


@pytest.mark.parametrize("rows", [1024, 3001, 15168])
def test_layernorm_bwd_streaming_kernel(rows):
    """The bulk-copy (cp.async.bulk ring) backward used by the encoder's hot calls: d = 768, bf16 dy, residual-gradient
    add, fp32 + bf16 outputs, dgamma / dbeta accumulation."""
    L = _lib()
    torch.manual_seed(rows)
    d = 768
    x = torch.randn(rows, d, device="cuda") * 3 + 1
    g = torch.randn(d, device="cuda")
    b = torch.randn(d, device="cuda")
    # (the forward with bf16 output only and >= 1024 dense rows is the streaming forward kernel)
    yb, _, mean, rstd = L.layernorm_fwd(x, g, b, 1e-12, out_bf16=True, out_f32=False)
    y_ref = torch.nn.functional.layer_norm(x, (d,), g, b, 1e-12)
    assert _rel_err(yb, y_ref) < 4e-3, _rel_err(yb, y_ref)
    assert (yb.float() - y_ref.bfloat16().float()).abs().max().item() <= 2 ** -6 * y_ref.abs().max().item()
    assert _max_err(mean, x.mean(1)) < 1e-5 and _rel_err(rstd, x.var(1, unbiased=False).add(1e-12).rsqrt()) < 1e-5
    dy = torch.randn(rows, d, device="cuda").bfloat16()
    dres = torch.randn(rows, d, device="cuda")
    xr, gr, br = x.clone().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-12).backward(dy.float())
    dx = torch.full((rows, d), float("nan"), device="cuda")
    dxb = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    dg, db = torch.ones(d, device="cuda"), torch.ones(d, device="cuda")
    L.layernorm_bwd(dy, x, g, b, mean, rstd, dres=dres, dx_f32=dx, dx_bf16=dxb, dgamma=dg, dbeta=db)
    assert _rel_err(dx, xr.grad + dres) < 2e-5, _rel_err(dx, xr.grad + dres)
    assert _rel_err(dxb, xr.grad + dres) < 4e-3
    assert _rel_err(dg, 1.0 + gr.grad) < 5e-5 and _rel_err(db, 1.0 + br.grad) < 5e-5
    # in-place over the residual gradient is NOT how the engine calls it, but a second call must give the same result
    dx2 = torch.empty_like(dx)
    L.layernorm_bwd(dy, x, g, b, mean, rstd, dres=dres, dx_f32=dx2, dx_bf16=dxb)
    assert torch.equal(dx, dx2)


def test_layernorm_strided_cls_rows():
    L = _lib()
    torch.manual_seed(0)
    B, L_, d = 5, 237, 768
    x = torch.randn(B, L_, d, device="cuda")
    g = torch.randn(d, device="cuda")
    b = torch.randn(d, device="cuda")
    yb, yf, mean, rstd = L.layernorm_fwd(x, g, b, 1e-12, rows=B, ldx=L_ * d, out_f32=True)
    ref = torch.nn.functional.layer_norm(x[:, 0], (d,), g, b, 1e-12)
    assert _max_err(yf, ref) < 2e-5 * ref.abs().max().item()


@pytest.mark.parametrize("M,d,r,act", [(15168, 768, 48, "swish"), (300, 768, 48, "relu"), (77, 128, 64, "swish"), (129, 256, 16, "relu")])
def test_adapter_fused_bottleneck_fwd_bwd(M, d, r, act):
    """climb_adapter_fused (down -> act -> up -> residual in one launch) against fp32 torch on the bf16-rounded operands:
    ViLT-base with CLiMB's reduction factor 16 (d = 768, r = 48) at the benchmark's row count, ragged row counts, r = 64 / 16."""
    import torch.nn.functional as F
    from climb_b200 import _lib
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(M + r)
    rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(dev)
    A = rnd(M, d).bfloat16()
    wd, wu = rnd(r, d, scale=0.05).bfloat16(), rnd(d, r, scale=0.05).bfloat16()
    bd, bu = rnd(r, scale=0.1), rnd(d, scale=0.1)
    c_in = rnd(M, d)
    act_id = _lib.EPI_SWISH if act == "swish" else _lib.EPI_RELU
    f = F.silu if act == "swish" else F.relu
    # ---- forward ----
    pre = torch.empty(M, r, dtype=torch.bfloat16, device=dev)
    z = torch.empty_like(pre)
    out = c_in.clone()
    c2 = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
    _lib.check(_lib.climb_adapter_fused(0, M, d, r, act_id, _lib.ptr(A), _lib.ptr(wd), _lib.ptr(wu), _lib.ptr(bd), _lib.ptr(bu),
                                        _lib.ptr(pre), _lib.ptr(z), _lib.ptr(out), _lib.ptr(out), _lib.ptr(c2), None, _lib.stream()))
    ref_pre = A.float() @ wd.float().t() + bd
    ref_z = f(ref_pre)
    ref_out = c_in + ref_z.bfloat16().float() @ wu.float().t() + bu
    assert _rel_err(pre.float(), ref_pre) < 4e-3 and _rel_err(z.float(), ref_z) < 6e-3
    assert _rel_err(out, ref_out) < 2e-3 and _rel_err(c2.float(), ref_out) < 5e-3
    # ---- backward: A = dout (bf16), c_in = dout (fp32) ----
    dout = rnd(M, d)
    dout_h = dout.bfloat16()
    dpre = torch.empty_like(pre)
    dx = dout.clone()
    dx_h = dout_h.clone()
    cs = torch.zeros(r, dtype=torch.float32, device=dev)
    _lib.check(_lib.climb_adapter_fused(1, M, d, r, act_id, _lib.ptr(dx_h), _lib.ptr(wd), _lib.ptr(wu), None, None, _lib.ptr(pre),
                                        _lib.ptr(dpre), _lib.ptr(dx), _lib.ptr(dx), _lib.ptr(dx_h), _lib.ptr(cs), _lib.stream()))
    p32 = pre.float().requires_grad_(True)
    (f(p32)).backward(dout_h.float() @ wu.float())
    ref_dpre = p32.grad
    ref_dx = dout + ref_dpre.bfloat16().float() @ wd.float()
    assert _rel_err(dpre.float(), ref_dpre) < 6e-3
    assert _rel_err(dx, ref_dx) < 2e-3 and _rel_err(dx_h.float(), ref_dx) < 5e-3
    assert _rel_err(cs, dpre.float().sum(0)) < 2e-3
    # bf16-only output (the mh site's backward): c_out = NULL
    dmh = torch.empty_like(dx_h)
    _lib.check(_lib.climb_adapter_fused(1, M, d, r, act_id, _lib.ptr(dout_h), _lib.ptr(wd), _lib.ptr(wu), None, None, _lib.ptr(pre),
                                        _lib.ptr(dpre), _lib.ptr(dout), None, _lib.ptr(dmh), None, _lib.stream()))
    assert _rel_err(dmh.float(), ref_dx) < 5e-3
