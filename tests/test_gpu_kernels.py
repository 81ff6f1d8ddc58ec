"""GPU: each hand-written kernel through the C ABI unit entry points vs a plain PyTorch fp32 statement
of the reference op (tolerances written per test)."""
import math

import pytest
import torch

from boficap_b200.layout import BofiConfig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    from boficap_b200.engine import BofiEngine
    e32 = BofiEngine(BofiConfig(), 0, "fp32")
    e16 = BofiEngine(BofiConfig(), 0, "bf16")
    yield e32, e16
    e32.close()
    e16.close()


def test_layernorm_matches_reference_formula(engines):
    e32, _ = engines
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(777, 512, generator=g) * 3 + 0.5).cuda()
    a = (1 + 0.1 * torch.randn(512, generator=g)).cuda()
    b = (0.1 * torch.randn(512, generator=g)).cuda()
    out = e32.layernorm(x, a, b)
    xd = x.double()
    ref = a.double() * (xd - xd.mean(-1, keepdim=True)) / (xd.std(-1, keepdim=True) + 1e-6) + b.double()
    assert (out.double() - ref).abs().max().item() < 2e-6     # fp32 rounding only
    # NOT nn.LayerNorm: biased variance / eps inside the sqrt would differ at the 1e-3 level
    wrong = torch.nn.functional.layer_norm(x, (512,), a, b, 1e-6)
    assert (out - wrong).abs().max().item() > 1e-4


@pytest.mark.parametrize("M,N,K,relu,resid", [(300, 512, 512, False, True), (128, 128, 64, True, False),
                                              (1000, 1536, 2048, False, False), (77, 200, 512, True, False),
                                              (260, 203, 512, False, False)])
def test_linear_fp32_simt(engines, M, N, K, relu, resid):
    e32, _ = engines
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda() if resid else None
    out = e32.linear(a, w, bias, res, relu)
    ref = a.double() @ w.double().T + bias.double()
    if relu:
        ref = ref.relu()
    if resid:
        ref = ref + res.double()
    assert (out.double() - ref).abs().max().item() < 5e-5     # fp32 accumulation order only


@pytest.mark.parametrize("M,N,K,relu,resid", [(128, 128, 64, False, False), (300, 512, 512, False, True),
                                              (1000, 1536, 2048, True, False), (257, 9496, 512, False, False),
                                              (36864, 512, 2048, True, False),
                                              # wide K=512 shapes on 2-CTA tile pairs (M tail, N tail, K tail)
                                              (5000, 2048, 512, True, False), (2304, 9496, 512, False, False),
                                              (40000, 1024, 512, False, False), (2048, 1536, 448, False, False)])
def test_linear_bf16_tcgen05(engines, M, N, K, relu, resid):
    _, e16 = engines
    g = torch.Generator().manual_seed(M + N + 1)
    a = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda() if resid else None
    out = e16.linear(a, w, bias, res, relu)
    # operands are rounded to bf16 once, products accumulate in fp32: compare against exactly that
    ref = a.bfloat16().double() @ w.bfloat16().double().T + bias.double()
    if relu:
        ref = ref.relu()
    if resid:
        ref = ref + res.double()
    err = (out.double() - ref).abs().max().item()
    assert err < 2e-4, err


@pytest.mark.parametrize("M,N,K", [(36864, 512, 2048), (20480, 512, 512), (300, 512, 512), (1000, 1024, 512), (2049, 520, 128)])
def test_linear_bf16_inplace_residual(engines, M, N, K):
    """x += a . W^T + b in place: the epilogue hands (acc + b) to the L2 as a TMA reduce-add (2-CTA pairs for M >= 2048,
    1-CTA narrow tiles below); must equal the out-of-place residual form exactly (same fp32 sum) and the fp64 reference."""
    _, e16 = engines
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    out_of_place = e16.linear(a, w, bias, res, False)
    x = res.clone()
    got = e16.linear(a, w, bias, x, False, inplace=True)
    assert got.data_ptr() == x.data_ptr()
    ref = a.bfloat16().double() @ w.bfloat16().double().T + bias.double() + res.double()
    assert (x.double() - ref).abs().max().item() < 2e-4
    assert torch.equal(x, out_of_place)


@pytest.mark.parametrize("M,K,rows", [(36864, 512, None), (20480, 2048, None), (300, 512, None), (2049, 128, None), (5000, 512, 3000)])
def test_linear_bf16_residual_layernorm_epilogue(engines, M, K, rows):
    """gemm_tc2_ln_kernel: x += a . W^T + b in place (must equal the plain in-place residual GEMM bit for bit: same fp32 sum) and
    y = LayerNorm(x) with the reference's formula (unbiased std, eps outside) out of the GEMM's epilogue, rounded to bf16."""
    _, e16 = engines
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(512, K, generator=g) / math.sqrt(K)).cuda()
    bias = torch.randn(512, generator=g).cuda()
    res = (torch.randn(M, 512, generator=g) * 2 + 0.3).cuda()
    ln_a = (1 + 0.1 * torch.randn(512, generator=g)).cuda()
    ln_b = (0.1 * torch.randn(512, generator=g)).cuda()
    x_ref = e16.linear(a, w, bias, res.clone(), False, inplace=True)
    x = res.clone()
    rows_dev = torch.tensor([rows], dtype=torch.int32, device="cuda") if rows is not None else None
    y = e16.linear_resid_ln(a, w, bias, x, ln_a, ln_b, rows_dev)
    torch.cuda.synchronize()
    n = rows if rows is not None else M
    assert torch.equal(x[:n], x_ref[:n])
    if rows is not None:                       # whole 256-row panels above the device-side row count are never touched
        first_untouched = (rows + 255) // 256 * 256
        assert torch.equal(x[first_untouched:], res[first_untouched:])
    xd = x[:n].double()
    ref = ln_a.double() * (xd - xd.mean(-1, keepdim=True)) / (xd.std(-1, keepdim=True) + 1e-6) + ln_b.double()
    err = (y[:n].double() - ref).abs()
    # bf16 output: half an ulp of the value (2^-9 relative) plus fp32 rounding of the statistics
    tol = ref.abs() * 2.0 ** -8 + 1e-5
    assert bool((err <= tol).all()), float((err - tol).max())
    # and against the stand-alone kernel's bf16 rounding: identical except where a value sits on a rounding boundary
    y2 = e16.layernorm(x[:n].contiguous(), ln_a, ln_b).bfloat16().float()
    assert float((y[:n] != y2).float().mean()) < 1e-3


def test_attention_prefix_masks_and_nan_rows(engines):
    e32, _ = engines
    g = torch.Generator().manual_seed(5)
    B, Tq, Tk, h = 9, 20, 100, 8
    q = torch.randn(B, Tq, 512, generator=g).cuda()
    k = torch.randn(B, Tk, 512, generator=g).cuda()
    v = torch.randn(B, Tk, 512, generator=g).cuda()
    vis = torch.randint(0, Tk + 1, (B, Tq), generator=g).int()
    vis[0, 0] = 0                     # all-masked row -> NaN like softmax over -inf
    vis[1, :] = Tk
    out = e32.attention(q, k, v, vis.cuda()).cpu()
    qh = q.cpu().view(B, Tq, h, 64).transpose(1, 2)
    kh = k.cpu().view(B, Tk, h, 64).transpose(1, 2)
    vh = v.cpu().view(B, Tk, h, 64).transpose(1, 2)
    sc = qh @ kh.transpose(-2, -1) / 8.0
    mask = torch.arange(Tk)[None, None, :] < vis[:, :, None]
    sc = sc.masked_fill(~mask[:, None], float("-inf"))
    ref = (torch.softmax(sc, -1) @ vh).transpose(1, 2).reshape(B, Tq, 512)
    assert torch.isnan(out[0, 0]).all() and torch.isnan(ref[0, 0]).all()
    ok = ~torch.isnan(ref)
    assert (torch.isnan(out) == torch.isnan(ref)).all()
    assert (out[ok] - ref[ok]).abs().max().item() < 2e-5


def _attention_reference(q, k, v, vis, h=8):
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    qh = q.view(B, Tq, h, 64).transpose(1, 2)
    kh = k.view(B, Tk, h, 64).transpose(1, 2)
    vh = v.view(B, Tk, h, 64).transpose(1, 2)
    sc = qh @ kh.transpose(-2, -1) / 8.0
    mask = torch.arange(Tk)[None, None, :] < vis[:, :, None]
    sc = sc.masked_fill(~mask[:, None], float("-inf"))
    return (torch.softmax(sc, -1) @ vh).transpose(1, 2).reshape(B, Tq, 512)


@pytest.mark.parametrize("Tq,Tk", [(36, 36), (20, 20), (22, 22), (20, 36), (100, 100), (20, 100), (1, 36), (1, 100), (50, 77)])
def test_attention_bf16_engine_kernels(engines, Tq, Tk):
    """bf16 engines: mma.sync tensor-core attention (Tq > 1) and the single-query row kernel (Tq == 1) on
    bf16-rounded q/k/v; probabilities are rounded to bf16 before P.V, outputs are bf16: 2e-2 absolute."""
    _, e16 = engines
    g = torch.Generator().manual_seed(Tq * 131 + Tk)
    B = 5
    q = torch.randn(B, Tq, 512, generator=g).bfloat16().float()
    k = torch.randn(B, Tk, 512, generator=g).bfloat16().float()
    v = torch.randn(B, Tk, 512, generator=g).bfloat16().float()
    vis = torch.randint(1, Tk + 1, (B, Tq), generator=g).int()
    vis[0, 0] = 0
    vis[1, :] = Tk
    out = e16.attention(q.cuda(), k.cuda(), v.cuda(), vis.cuda()).cpu()
    ref = _attention_reference(q, k, v, vis)
    assert torch.isnan(out[0, 0]).all()
    assert (torch.isnan(out) == torch.isnan(ref)).all()
    ok = ~torch.isnan(ref)
    assert (out[ok] - ref[ok]).abs().max().item() < 2e-2


@pytest.mark.parametrize("Tk", [36, 100])
def test_attention_single_query_fp32(engines, Tk):
    e32, _ = engines
    g = torch.Generator().manual_seed(Tk)
    B = 7
    q = torch.randn(B, 1, 512, generator=g)
    k = torch.randn(B, Tk, 512, generator=g)
    v = torch.randn(B, Tk, 512, generator=g)
    vis = torch.randint(1, Tk + 1, (B, 1), generator=g).int()
    out = e32.attention(q.cuda(), k.cuda(), v.cuda(), vis.cuda()).cpu()
    assert (out - _attention_reference(q, k, v, vis)).abs().max().item() < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("env", ["BOFI_GEMM2", "BOFI_KPS", "BOFI_ARES", "BOFI_LNFUSE"])
def test_optional_gemm_variants_reproduce_the_default_path(env, monkeypatch):
    """The other tcgen05 variants (1-CTA tiles via BOFI_GEMM2=0, one k-block per stage via BOFI_KPS=1, the opt-in A-resident 2-CTA tiles and LayerNorm-fused GEMM) must give
    the default bf16 path's results: same boxes, logits within bf16 rounding of the LayerNorm output."""
    import torch
    from boficap_b200 import synth
    from boficap_b200.engine import BofiEngine
    from boficap_b200.layout import BofiConfig
    cfg = BofiConfig()
    sd = synth.synth_state_dict(cfg, 0, "s_real")
    fc, att_all, _ = synth.synth_inputs(2304, 36, seed=8)      # M >= 73k encoder rows: the wide-tile paths are taken
    att_all = att_all.cuda()

    def run(att):
        eng = BofiEngine(cfg, 0, "bf16").load_state_dict(sd)
        eng.encode(att, None)
        out = [t.clone() for t in eng.decode("NAIC", 1, 0, True)]
        eng.close()
        return out

    monkeypatch.delenv(env, raising=False)
    for B in range(2048, 2304):                                # end the batch on an image that predicts a phrase (no NaN batch)
        base = run(att_all[:B])
        if not bool(base[1].isnan().any()):
            break
    else:
        raise AssertionError("no usable batch")
    monkeypatch.setenv(env, "0" if env == "BOFI_GEMM2" else "1")
    monkeypatch.setenv("BOFI_LNFUSE_MIN", "1")
    var = run(att_all[:B])
    same = (base[3] == var[3]).all(1)
    assert float(same.float().mean()) > 0.95
    a, b = base[1][same], var[1][same]
    assert not bool(b.isnan().any())
    err = float((a - b).abs().max())
    assert err < (2e-2 if env == "BOFI_LNFUSE" else 1e-5), err
