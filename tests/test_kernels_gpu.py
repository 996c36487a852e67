"""GPU parity of every C-ABI kernel against plain torch fp32 math on the same (bf16-rounded) inputs.
Run on the B200 box:  python -m pytest tests -m gpu -q"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from theia_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"


def S():
    return torch.cuda.current_stream().cuda_stream


def rnd(*shape, scale=1.0, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (scale * torch.randn(*shape, generator=g)).to(dtype).to(DEV)


def relerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def gemm(lib, **kw):
    d = L.GemmDesc()
    d.splits, d.batch_z = 1, 1
    for k, v in kw.items():
        if k == "conv":
            d.conv = v
        elif torch.is_tensor(v):
            setattr(d, k, v.data_ptr())
        else:
            setattr(d, k, v)
    L.check(lib.theia_gemm(C.byref(d), S()), "theia_gemm")


# Every GEMM / implicit-GEMM test runs three ways: the launcher's own choice, single-CTA kernels only
# (cta_group::1), and CTA pairs wherever the tile allows (cta_group::2, including the conv gathers).
@pytest.fixture(params=["auto", "single", "pair"])
def cta_mode(request, lib):
    lib.theia_debug_set(8, {"auto": 0, "single": 1, "pair": 2}[request.param])
    yield request.param
    lib.theia_debug_set(8, 0)


# ----------------------------------------------------------------------------- GEMM, K-major
@pytest.mark.parametrize("M,N,K,bn", [(256, 256, 128, 0), (591, 192, 192, 0), (394, 1280, 768, 0),
                                      (128, 128, 64, 128), (1000, 576, 192, 192), (300, 768, 3072, 256), (591, 512, 192, 256),
                                      (197, 32, 192, 0)])
def test_gemm_k2d_bias(lib, M, N, K, bn, cta_mode):
    a, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05)
    bias = rnd(N, seed=3, dtype=torch.float32)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    gemm(lib, M=M, N=N, K=K, A=a, lda=K, B=b, ldb=K, out=out, ldo=N, bias=bias, bn=bn)
    ref = a.float() @ b.float().t() + bias
    torch.cuda.synchronize()
    assert relerr(out.float(), ref) < 6e-3


def test_gemm_epilogues(lib, cta_mode):
    M, N, K = 394, 768, 192
    a, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.08)
    bias = rnd(N, seed=3, dtype=torch.float32)
    aux = rnd(M, N, seed=4)
    acc = a.float() @ b.float().t() + bias
    # GELU: out2 = pre-activation, out = gelu
    out, out2 = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV), torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    gemm(lib, M=M, N=N, K=K, A=a, lda=K, B=b, ldb=K, out=out, ldo=N, bias=bias, out2=out2, epi=L.EPI_GELU)
    xg = acc.clone().requires_grad_(True)
    F.gelu(xg).sum().backward()
    assert relerr(out2.float(), xg.grad) < 6e-3  # out2 = gelu'(pre-activation), saved for the backward pass
    assert relerr(out.float(), F.gelu(acc)) < 8e-3
    # the same with the derivative saved as an 8-bit code (EPI_AUX_U8): q = rint((g' + 0.129) * 255 / 1.258)
    out8 = torch.zeros(M, N, dtype=torch.uint8, device=DEV)
    outg = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    gemm(lib, M=M, N=N, K=K, A=a, lda=K, B=b, ldb=K, out=outg, ldo=N, bias=bias, out2=out8, epi=L.EPI_GELU | L.EPI_AUX_U8)
    assert torch.equal(outg, out)
    dec = out8.float() * (1.258 / 255.0) - 0.129
    assert (dec - xg.grad).abs().max().item() < 3.5e-3  # half a code step (2.5e-3) + the erf approximation
    assert out8.min().item() >= 0 and xg.grad.min().item() > -0.129 and xg.grad.max().item() < 1.129
    # ... and consumed by the dgrad epilogue: v *= decode(aux)
    cs8 = torch.zeros(N, device=DEV)
    o8 = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    gemm(lib, M=M, N=N, K=K, A=a, lda=K, B=b, ldb=K, out=o8, ldo=N, aux=out8, epi=L.EPI_MUL_AUX | L.EPI_COLSUM | L.EPI_AUX_U8,
         colsum=cs8)
    assert relerr(o8.float(), (acc - bias) * dec) < 6e-3
    assert relerr(cs8, o8.float().sum(0)) < 1e-4
    # residual
    gemm(lib, M=M, N=N, K=K, A=a, lda=K, B=b, ldb=K, out=out, ldo=N, bias=bias, aux=aux, epi=L.EPI_RESID)
    assert relerr(out.float(), acc + aux.float()) < 6e-3
    # accumulate in place (aux aliases out)
    out.copy_(aux)
    gemm(lib, M=M, N=N, K=K, A=a, lda=K, B=b, ldb=K, out=out, ldo=N, aux=out, epi=L.EPI_RESID)
    assert relerr(out.float(), acc - bias + aux.float()) < 6e-3
    # fp32 output
    o32 = torch.zeros(M, N, dtype=torch.float32, device=DEV)
    gemm(lib, M=M, N=N, K=K, A=a, lda=K, B=b, ldb=K, out=o32, ldo=N, bias=bias, epi=L.EPI_OUT_F32)
    assert relerr(o32, acc) < 1e-5
    # gelu' and relu-mask multipliers
    cs = torch.zeros(N, device=DEV)
    gemm(lib, M=M, N=N, K=K, A=a, lda=K, B=b, ldb=K, out=out, ldo=N, aux=aux, epi=L.EPI_MUL_AUX | L.EPI_COLSUM,
         colsum=cs)
    assert relerr(out.float(), (acc - bias) * aux.float()) < 6e-3
    assert relerr(cs, out.float().sum(0)) < 1e-4
    gemm(lib, M=M, N=N, K=K, A=a, lda=K, B=b, ldb=K, out=out, ldo=N, aux=aux, epi=L.EPI_MUL_RELUMASK)
    assert relerr(out.float(), (acc - bias) * (aux.float() > 0)) < 6e-3
    # relu
    gemm(lib, M=M, N=N, K=K, A=a, lda=K, B=b, ldb=K, out=out, ldo=N, bias=bias, epi=L.EPI_RELU)
    assert relerr(out.float(), acc.clamp_min(0)) < 6e-3


def test_gemm_poscls(lib, cta_mode):
    Bn, D, K = 3, 192, 768
    M = Bn * 197
    a = rnd(M, K, seed=1)
    a.view(Bn, 197, K)[:, 0] = 0  # CLS slot rows are zero patches
    w = rnd(D, K, seed=2, scale=0.03)
    bias, pos, cls = (rnd(D, seed=3, dtype=torch.float32), rnd(197, D, seed=4, dtype=torch.float32),
                      rnd(D, seed=5, dtype=torch.float32))
    out = torch.zeros(M, D, dtype=torch.bfloat16, device=DEV)
    table = pos.clone()
    table[0] += cls  # per-token table: CLS row holds cls_token + pos[0]
    gemm(lib, M=M, N=D, K=K, A=a, lda=K, B=w, ldb=K, out=out, ldo=D, bias=bias, pos=table, tokens=197, tok_p0=1,
         tok_p1=197, epi=L.EPI_POSCLS)
    ref = (a.float() @ w.float().t() + bias).view(Bn, 197, D) + pos
    ref[:, 0] = cls + pos[0]
    assert relerr(out.float().view(Bn, 197, D), ref) < 6e-3


# ----------------------------------------------------------------------------- GEMM, MN-major (wgrad)
@pytest.mark.parametrize("Mtok,Nout,Kin,splits,bn", [(512, 256, 256, 1, 256), (1000, 320, 192, 3, 192),
                                                     (788, 768, 768, 4, 256), (640, 32, 192, 2, 0),
                                                     (985, 576, 192, 5, 128)])
def test_gemm_wgrad_mn_major(lib, Mtok, Nout, Kin, splits, bn, cta_mode):
    dy, x = rnd(Mtok, Nout, seed=1), rnd(Mtok, Kin, seed=2)
    dw = torch.zeros(Nout, Kin, dtype=torch.float32, device=DEV)
    gemm(lib, M=Nout, N=Kin, K=Mtok, a_mode=L.OP_MN2D, b_mode=L.OP_MN2D, A=dy, lda=Nout, B=x, ldb=Kin, out=dw,
         ldo=Kin, epi=L.EPI_ATOMIC, splits=splits, bn=bn)
    ref = dy.float().t() @ x.float()
    assert relerr(dw, ref) < 1e-4


def test_gemm_mixed_major_b(lib, cta_mode):
    """dgrad without a transposed weight copy: B operand MN-major (W stored [K_gemm][N_gemm])."""
    M, Nout, Kin = 300, 320, 256  # y = dy[M,Nout] @ W[Nout,Kin]
    dy, w = rnd(M, Nout, seed=1), rnd(Nout, Kin, seed=2, scale=0.05)
    out = torch.zeros(M, Kin, dtype=torch.bfloat16, device=DEV)
    gemm(lib, M=M, N=Kin, K=Nout, a_mode=L.OP_K2D, b_mode=L.OP_MN2D, A=dy, lda=Nout, B=w, ldb=Kin, out=out, ldo=Kin)
    assert relerr(out.float(), dy.float() @ w.float()) < 6e-3


# ----------------------------------------------------------------------------- implicit-GEMM conv
def geom16(Cc, Hin, Bn, sw, sh, sb, shift0):
    g = L.ConvGeom()
    g.C, g.H, g.W, g.B = Cc, Hin, Hin, Bn
    g.stride_w, g.stride_h, g.stride_b = sw, sh, sb
    g.ntaps = 9
    for t in range(9):
        g.dh[t], g.dw[t] = t // 3 + shift0, t % 3 + shift0
    g.tile_w, g.tile_h = 16, 8
    g.out_h, g.out_w, g.out_img_rows, g.out_row_off, g.out_wpitch = 16, 16, 256, 0, 16
    g.sy = g.sx = 1
    return g


@pytest.mark.parametrize("Cc,Bn", [(128, 3), (192, 2), (256, 3)])
def test_conv3x3_fwd_relu_stats(lib, Cc, Bn, cta_mode):
    x = rnd(Bn, 16, 16, Cc, seed=1)  # NHWC
    w = rnd(Cc, Cc, 3, 3, seed=2, scale=0.05, dtype=torch.float32)  # Conv2d [co,ci,kh,kw]
    bias = rnd(Cc, seed=3, dtype=torch.float32)
    wp = w.permute(0, 2, 3, 1).reshape(Cc, 9 * Cc).to(torch.bfloat16).contiguous()  # [co][tap][ci]
    out = torch.zeros(Bn * 256, Cc, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(Bn, 2, dtype=torch.float32, device=DEV)
    g = geom16(Cc, 16, Bn, Cc, 16 * Cc, 256 * Cc, -1)
    gemm(lib, M=Bn * 256, N=Cc, K=9 * Cc, a_mode=L.OP_CONV_K, A=x, B=wp, ldb=9 * Cc, conv=g, out=out, ldo=Cc,
         bias=bias, stats=stats, epi=L.EPI_RELU | L.EPI_STATS)
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wp.float().view(Cc, 3, 3, Cc).permute(0, 3, 1, 2), bias,
                          padding=1)).permute(0, 2, 3, 1).reshape(Bn * 256, Cc)
    assert relerr(out.float(), ref) < 6e-3
    o = out.float().view(Bn, -1)
    torch.testing.assert_close(stats[:, 0], o.sum(1), rtol=2e-4, atol=1e-2)
    torch.testing.assert_close(stats[:, 1], (o * o).sum(1), rtol=2e-4, atol=1e-2)


def test_pad_convtranspose_fwd_and_dgrad(lib, cta_mode):
    Cc, Bn, D = 128, 2, 128
    tok = rnd(Bn, 197, D, seed=1)
    wt = rnd(Cc, Cc, 3, 3, seed=2, scale=0.05, dtype=torch.float32).to(torch.bfloat16).float()  # [ci,co,kh,kw]
    bias = rnd(Cc, seed=3, dtype=torch.float32)
    # fwd pack F[co][tap2][ci] = Wt[ci][co][2-kh2][2-kw2]
    wf = wt.flip(2, 3).permute(1, 2, 3, 0).reshape(Cc, 9 * Cc).to(torch.bfloat16).contiguous()
    out = torch.zeros(Bn * 256, Cc, dtype=torch.bfloat16, device=DEV)
    g = geom16(Cc, 14, Bn, D, 14 * D, 197 * D, -2)
    a_ptr = tok.data_ptr() + 2 * D
    gemm(lib, M=Bn * 256, N=Cc, K=9 * Cc, a_mode=L.OP_CONV_K, A=a_ptr, B=wf, ldb=9 * Cc, conv=g, out=out, ldo=Cc,
         bias=bias)
    xin = tok[:, 1:].float().reshape(Bn, 14, 14, D).permute(0, 3, 1, 2)
    ref = F.conv_transpose2d(xin, wt, bias).permute(0, 2, 3, 1).reshape(Bn * 256, Cc)
    assert relerr(out.float(), ref) < 6e-3
    # dgrad onto the token grid (rows 1..196 of each image), accumulated into an existing tensor
    dy = rnd(Bn * 256, Cc, seed=5)
    wd = wt.permute(0, 2, 3, 1).reshape(Cc, 9 * Cc).to(torch.bfloat16).contiguous()  # D[ci][tap][co]
    dtok = rnd(Bn, 197, D, seed=6)
    dtok0 = dtok.clone()
    g2 = geom16(Cc, 16, Bn, Cc, 16 * Cc, 256 * Cc, 0)
    g2.out_h, g2.out_w, g2.out_img_rows, g2.out_row_off, g2.out_wpitch = 14, 14, 197, 1, 14
    gemm(lib, M=Bn * 256, N=D, K=9 * Cc, a_mode=L.OP_CONV_K, A=dy, B=wd, ldb=9 * Cc, conv=g2, out=dtok, ldo=D,
         aux=dtok, epi=L.EPI_RESID)
    xr = xin.clone().requires_grad_(True)
    (F.conv_transpose2d(xr, wt, None) * dy.float().view(Bn, 16, 16, Cc).permute(0, 3, 1, 2)).sum().backward()
    want = dtok0.float().clone()
    want[:, 1:] += xr.grad.permute(0, 2, 3, 1).reshape(Bn, 196, D)
    assert relerr(dtok.float(), want) < 6e-3
    assert torch.equal(dtok[:, 0], dtok0[:, 0])  # CLS rows untouched


@pytest.mark.parametrize("Cc", [128, 256])
def test_conv_wgrad(lib, Cc, cta_mode):
    Bn = 3
    x = rnd(Bn, 16, 16, Cc, seed=1)
    dy = rnd(Bn * 256, Cc, seed=2)
    ws = torch.zeros(9, Cc, Cc, dtype=torch.float32, device=DEV)
    g = geom16(Cc, 16, Bn, Cc, 16 * Cc, 256 * Cc, -1)
    gemm(lib, M=Cc, N=Cc, K=Bn * 256, a_mode=L.OP_MN2D, b_mode=L.OP_CONV_MN, A=dy, lda=Cc, B=x, conv=g, out=ws,
         ldo=Cc, epi=L.EPI_ATOMIC, batch_z=9, out_z_stride=Cc * Cc, splits=2, bn=Cc)
    w = torch.zeros(Cc, Cc, 3, 3, device=DEV, requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, None, padding=1)
    (y * dy.float().view(Bn, 16, 16, Cc).permute(0, 3, 1, 2)).sum().backward()
    ref = w.grad.permute(2, 3, 0, 1).reshape(9, Cc, Cc)  # [tap][co][ci]
    assert relerr(ws, ref) < 1e-4


def test_pad_conv_wgrad(lib, cta_mode):
    Cc, Bn, D = 128, 2, 128
    tok = rnd(Bn, 197, D, seed=1)
    dy = rnd(Bn * 256, Cc, seed=2)
    ws = torch.zeros(9, Cc, Cc, dtype=torch.float32, device=DEV)
    g = geom16(Cc, 14, Bn, D, 14 * D, 197 * D, -2)
    gemm(lib, M=Cc, N=Cc, K=Bn * 256, a_mode=L.OP_MN2D, b_mode=L.OP_CONV_MN, A=dy, lda=Cc, B=tok.data_ptr() + 2 * D,
         conv=g, out=ws, ldo=Cc, epi=L.EPI_ATOMIC, batch_z=9, out_z_stride=Cc * Cc, splits=1, bn=128)
    wt = torch.zeros(Cc, Cc, 3, 3, device=DEV, requires_grad=True)
    xin = tok[:, 1:].float().reshape(Bn, 14, 14, D).permute(0, 3, 1, 2)
    y = F.conv_transpose2d(xin, wt, None)
    (y * dy.float().view(Bn, 16, 16, Cc).permute(0, 3, 1, 2)).sum().backward()
    # ws[tap2][co][ci] = dWt[ci][co][2-kh2][2-kw2]
    ref = wt.grad.flip(2, 3).permute(2, 3, 1, 0).reshape(9, Cc, Cc)
    assert relerr(ws, ref) < 1e-4


# ----------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("M,D", [(591, 192), (300, 384), (394, 768), (4133, 192), (4500, 768), (4099, 1024),
                                 (30011, 192), (40000, 768)])  # >= 4096 rows: ring-staged forward; the last two wrap the rings
def test_layernorm_fwd_bwd(lib, M, D):
    x, dy, dadd = rnd(M, D, seed=1, scale=2.0), rnd(M, D, seed=2), rnd(M, D, seed=3)
    gamma = 1 + 0.1 * rnd(D, seed=4, dtype=torch.float32)
    beta = 0.1 * rnd(D, seed=5, dtype=torch.float32)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    L.check(lib.theia_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                    rstd.data_ptr(), M, D, 1e-12, S()))
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (D,), gr, br, 1e-12)
    assert relerr(y.float(), ref) < 5e-3
    ref.backward(dy.float())
    dx = torch.empty_like(x)
    dg, db, dxs = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    L.check(lib.theia_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                    dadd.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), dxs.data_ptr(), M, D,
                                    S()))
    assert relerr(dx.float(), xr.grad + dadd.float()) < 6e-3
    assert relerr(dg, gr.grad) < 1e-3
    assert relerr(db, br.grad) < 1e-3
    # fused column sums of dx (bias gradient of the producer Linear): accumulated in fp32 before the bf16
    # rounding of dx, so they follow the fp32 reference more closely than a sum over the stored tensor
    assert relerr(dxs, (xr.grad + dadd.float()).sum(0)) < 2e-4
    assert relerr(dxs, dx.float().sum(0)) < 5e-3


def test_ln3d_apply_and_bwd(lib):
    Bn, Cc = 19, 64
    n = 256 * Cc
    x = F.relu(rnd(Bn, n, seed=1)).contiguous()
    dy = rnd(Bn, n, seed=2)
    gamma = 1 + 0.1 * rnd(n, seed=3, dtype=torch.float32)
    beta = 0.1 * rnd(n, seed=4, dtype=torch.float32)
    xf = x.float()
    stats = torch.stack([xf.sum(1), (xf * xf).sum(1)], 1).contiguous()
    y = torch.empty_like(x)
    L.check(lib.theia_ln3d_apply(x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), Bn,
                                 n, 1e-5, Cc, 0, 0, S()))
    xr = xf.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (n,), gr, br, 1e-5)
    assert relerr(y.float(), ref) < 5e-3
    ref.backward(dy.float())
    red = torch.empty(Bn, 2, device=DEV)
    dx = torch.empty_like(x)
    dg, db = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    L.check(lib.theia_ln3d_bwd(dy.data_ptr(), x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), red.data_ptr(),
                               dx.data_ptr(), dg.data_ptr(), db.data_ptr(), Bn, n, 1e-5, 1, Cc, 0, 0, S()))
    want = xr.grad * (xf > 0)
    assert relerr(dx.float(), want) < 6e-3
    assert relerr(dg, gr.grad) < 1e-3
    assert relerr(db, br.grad) < 1e-3


# ----------------------------------------------------------------------------- loss
@pytest.mark.parametrize("tgt_dtype", [torch.float32, torch.bfloat16])
def test_loss_fwd_bwd(lib, tgt_dtype):
    from oracle import theia_oracle as O
    Bn, n = 5, 256 * 96
    pred = 1.5 * rnd(Bn, 256, 96, seed=1, dtype=torch.float32)
    tgt = rnd(Bn, 256, 96, seed=2, dtype=tgt_dtype)
    acc = torch.empty(Bn, 5, device=DEV)
    out = torch.empty(3, device=DEV)
    L.check(lib.theia_loss_fwd(pred.data_ptr(), tgt.data_ptr(), int(tgt_dtype == torch.bfloat16), acc.data_ptr(),
                               out.data_ptr(), Bn, n, S()))
    pr = pred.clone().requires_grad_(True)
    ref = O.get_loss({"t": pr}, {"t": tgt.float()})
    want = torch.stack([ref["mse_loss"], ref["cos_loss"], ref["l1_loss"]]).detach()
    torch.testing.assert_close(out, want, rtol=2e-5, atol=1e-6)
    coef = torch.tensor([0.3, 0.9, 0.1], device=DEV)
    (0.3 * ref["mse_loss"] + 0.9 * ref["cos_loss"] + 0.1 * ref["l1_loss"]).backward()
    d32 = torch.empty_like(pred)
    dcp = torch.empty(pred.shape, dtype=torch.bfloat16, device=DEV)  # bf16 copy written in the same pass
    L.check(lib.theia_loss_bwd(pred.data_ptr(), tgt.data_ptr(), int(tgt_dtype == torch.bfloat16), acc.data_ptr(),
                               coef.data_ptr(), d32.data_ptr(), 1, dcp.data_ptr(), Bn, n, S()))
    assert relerr(d32, pr.grad) < 1e-4
    assert torch.equal(dcp, d32.to(torch.bfloat16))
    d16 = torch.empty(pred.shape, dtype=torch.bfloat16, device=DEV)
    L.check(lib.theia_loss_bwd(pred.data_ptr(), tgt.data_ptr(), int(tgt_dtype == torch.bfloat16), acc.data_ptr(),
                               coef.data_ptr(), d16.data_ptr(), 0, 0, Bn, n, S()))
    assert relerr(d16.float(), pr.grad) < 5e-3


# ----------------------------------------------------------------------------- preprocess
@pytest.mark.parametrize("chw", [0, 1])
def test_preprocess(lib, chw):
    from oracle import theia_oracle as O
    Bn = 3
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (Bn, 224, 224, 3), dtype=torch.uint8, generator=g)
    if chw:
        img = img.permute(0, 3, 1, 2).contiguous()
    out = torch.empty(Bn * 197, 768, dtype=torch.bfloat16, device=DEV)
    mean, std = (C.c_float * 3)(*O.IMAGE_MEAN), (C.c_float * 3)(*O.IMAGE_STD)
    d = img.to(DEV)
    L.check(lib.theia_preprocess(d.data_ptr(), out.data_ptr(), Bn, chw, 0, 1, 1, mean, std, 197, 1, S()))
    pix = O.preprocess(img, do_resize=False)  # [B,3,224,224]
    ref = F.unfold(pix, 16, stride=16).transpose(1, 2)  # [B,196, c*256+i*16+j]
    o = out.float().view(Bn, 197, 768).cpu()
    assert torch.all(o[:, 0] == 0)
    assert relerr(o[:, 1:], ref) < 4e-3


@pytest.mark.parametrize("chw", [0, 1])
def test_preprocess_with_bicubic_resize(lib, chw):
    """do_resize=True (the reference's default): resize 256 bicubic-antialias on uint8 + centre crop 224.
    Checked against the oracle running the torchvision CUDA path the reference itself takes for device tensors."""
    from oracle import theia_oracle as O
    Bn = 2
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (Bn, 224, 224, 3), dtype=torch.uint8, generator=g)
    # smooth content too (random noise alone under-tests the interpolation)
    yy, xx = torch.meshgrid(torch.arange(224), torch.arange(224), indexing="ij")
    img[1] = ((yy[..., None] * 0.7 + xx[..., None] * 0.4 + torch.arange(3) * 40) % 256).to(torch.uint8)
    if chw:
        img = img.permute(0, 3, 1, 2).contiguous()
    d = img.to(DEV)
    out = torch.empty(Bn * 197, 768, dtype=torch.bfloat16, device=DEV)
    mean, std = (C.c_float * 3)(*O.IMAGE_MEAN), (C.c_float * 3)(*O.IMAGE_STD)
    L.check(lib.theia_preprocess(d.data_ptr(), out.data_ptr(), Bn, chw, 1, 1, 1, mean, std, 197, 1, S()))
    pix = O.preprocess(d, do_resize=True)  # device tensor -> torchvision float path + round
    ref = F.unfold(pix, 16, stride=16).transpose(1, 2)
    o = out.float().view(Bn, 197, 768)
    assert torch.all(o[:, 0] == 0)
    assert relerr(o[:, 1:], ref) < 4e-3
    # uint8 levels agree exactly almost everywhere (float summation order may move a .5 boundary)
    lev = lambda x, c: x * (255 * O.IMAGE_STD[c]) + 255 * O.IMAGE_MEAN[c]
    diff = (lev(o[:, 1:, :256], 0) - lev(ref[:, :, :256], 0)).abs()
    assert (diff > 1.6).float().mean().item() < 1e-4


@pytest.mark.parametrize("chw", [0, 1])
def test_resize_byte_stage_exact(lib, chw):
    """The BYTE stage of A1 (hf image_processing_backends.py:361-414): uint8 in, bicubic-antialias resize 224 -> 256,
    round, uint8 out, centre crop 224.  The kernel's resized uint8 image (test hook) must EQUAL
    torchvision `resize(uint8 CUDA tensor, antialias=True)` + centre crop -- the path the reference takes for the
    device tensors train_rvfm.py:101 feeds it.  The CPU uint8 path of torchvision (fixed-point weights, the path the
    CPU goldens used) may differ from the float path by one level on a few pixels: measured and bounded."""
    import torchvision.transforms.v2.functional as tvF
    from oracle import theia_oracle as O
    Bn = 3
    g = torch.Generator().manual_seed(5)
    img = torch.randint(0, 256, (Bn, 224, 224, 3), dtype=torch.uint8, generator=g)
    yy, xx = torch.meshgrid(torch.arange(224), torch.arange(224), indexing="ij")
    img[1] = ((yy[..., None] * 0.7 + xx[..., None] * 0.4 + torch.arange(3) * 40) % 256).to(torch.uint8)
    img[2] = ((yy[..., None] // 16 + xx[..., None] // 16) % 2 * 255).to(torch.uint8)  # hard edges: clamp + rounding
    nchw = img.permute(0, 3, 1, 2).contiguous()
    d = (nchw if chw else img).to(DEV)
    dbg = torch.zeros(Bn, 224, 224, 3, dtype=torch.uint8, device=DEV)
    out = torch.empty(Bn * 197, 768, dtype=torch.bfloat16, device=DEV)
    mean, std = (C.c_float * 3)(*O.IMAGE_MEAN), (C.c_float * 3)(*O.IMAGE_STD)
    L.check(lib.theia_preprocess_debug_u8(dbg.data_ptr()))
    try:
        L.check(lib.theia_preprocess(d.data_ptr(), out.data_ptr(), Bn, chw, 1, 1, 1, mean, std, 197, 1, S()))
        torch.cuda.synchronize()
    finally:
        L.check(lib.theia_preprocess_debug_u8(0))
    ours = dbg.permute(0, 3, 1, 2).clone()
    ref_cuda = tvF.center_crop(tvF.resize(nchw.to(DEV), [256, 256], interpolation=tvF.InterpolationMode.BICUBIC,
                                          antialias=True), [224, 224])
    assert ref_cuda.dtype == torch.uint8
    assert torch.equal(ours, ref_cuda)  # do_resize = 1: byte-exact with torchvision's CUDA (float) path
    # do_resize = 2: torchvision's CPU uint8 path (int16 fixed-point weights, horizontal pass rounded + clamped to
    # uint8, then the vertical pass) -- what the reference's processor does when the images arrive on the CPU
    # (eval loop train_rvfm.py:165, PIL / numpy inputs).  Also byte-exact.
    L.check(lib.theia_preprocess_debug_u8(dbg.data_ptr()))
    try:
        L.check(lib.theia_preprocess(d.data_ptr(), out.data_ptr(), Bn, chw, 2, 1, 1, mean, std, 197, 1, S()))
        torch.cuda.synchronize()
    finally:
        L.check(lib.theia_preprocess_debug_u8(0))
    ref_cpu = tvF.center_crop(tvF.resize(nchw, [256, 256], interpolation=tvF.InterpolationMode.BICUBIC,
                                         antialias=True), [224, 224]).to(DEV)
    assert torch.equal(dbg.permute(0, 3, 1, 2), ref_cpu)
    # the two torchvision paths are different algorithms: they differ from each other on noise / hard edges
    diff = (ours.int() - ref_cpu.int()).abs().float()
    print(f"torchvision CUDA float path vs CPU fixed-point path: {100 * (diff > 0).float().mean().item():.2f} % of pixels "
          f"differ, max {int(diff.max())} levels (both reproduced bit for bit)")


@pytest.mark.parametrize("H,W,chw", [(300, 240, 0), (160, 200, 1), (480, 640, 0), (256, 100, 0), (1000, 700, 1), (17, 33, 0),
                                     (224, 320, 0), (231, 224, 1)])
def test_resize_any_extent_byte_stage_exact(lib, H, W, chw):
    """A1 for images of any extent (theia_preprocess_hw): the byte stage -- resize H x W -> 256 x 256 (up- and
    down-sampling: the tap support widens with the scale), centre crop 224 -- equals torchvision bit for bit, in float
    arithmetic (CUDA tensors) and in fixed point (CPU uint8 tensors); without resize: centre crop / zero padding as
    hf center_crop does it."""
    import torchvision.transforms.v2.functional as tvF
    from oracle import theia_oracle as O
    Bn = 2
    g = torch.Generator().manual_seed(H * 7 + W)
    img = torch.randint(0, 256, (Bn, H, W, 3), dtype=torch.uint8, generator=g)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    img[1] = ((yy[..., None] // 9 + xx[..., None] // 5) % 2 * 255).to(torch.uint8)  # hard edges: clamp + rounding
    nchw = img.permute(0, 3, 1, 2).contiguous()
    d = (nchw if chw else img).to(DEV)
    dbg = torch.zeros(Bn, 224, 224, 3, dtype=torch.uint8, device=DEV)
    out = torch.empty(Bn * 197, 768, dtype=torch.bfloat16, device=DEV)
    mean, std = (C.c_float * 3)(*O.IMAGE_MEAN), (C.c_float * 3)(*O.IMAGE_STD)

    def run(mode):
        L.check(lib.theia_preprocess_debug_u8(dbg.data_ptr()))
        try:
            L.check(lib.theia_preprocess_hw(d.data_ptr(), H, W, out.data_ptr(), Bn, chw, mode, 1, 1, mean, std, 197, 1, S()))
            torch.cuda.synchronize()
        finally:
            L.check(lib.theia_preprocess_debug_u8(0))
        return dbg.permute(0, 3, 1, 2).clone()

    rs = lambda t: tvF.resize(t, [256, 256], interpolation=tvF.InterpolationMode.BICUBIC, antialias=True)
    crop = lambda t: t[..., 16:240, 16:240]
    assert torch.equal(run(1), crop(rs(nchw.to(DEV))))      # float path = torchvision on CUDA tensors
    assert torch.equal(run(2), crop(rs(nchw)).to(DEV))      # fixed-point path = torchvision on CPU uint8 tensors
    # the patch rows carry the normalised values of exactly those bytes
    u8 = run(2)
    ref = ((u8.float() / 255.0 - torch.tensor(O.IMAGE_MEAN, device=DEV)[:, None, None]) /
           torch.tensor(O.IMAGE_STD, device=DEV)[:, None, None])
    pat = ref.unfold(2, 16, 16).unfold(3, 16, 16).permute(0, 2, 3, 1, 4, 5).reshape(Bn, 196, 768)
    got = out.view(Bn, 197, 768)
    assert torch.count_nonzero(got[:, 0]) == 0
    assert relerr(got[:, 1:].float(), pat) < 4e-3
    # no resize: centre crop, zero padding for the smaller axis (processor semantics, restated by the oracle)
    want = O.preprocess(img, do_resize=False, do_rescale=False, do_normalize=False)  # uint8 [B,3,224,224]
    assert want.dtype == torch.uint8 and torch.equal(run(0).cpu(), want)


# ----------------------------------------------------------------------------- attention
@pytest.mark.parametrize("Bn,H", [(2, 3), (3, 12), (40, 6), (100, 12)])  # the last: 8 (image, head) items per CTA, every ring wraps
def test_attention_fwd_bwd(lib, Bn, H):
    N, D = 197, H * 64
    qkv = rnd(Bn * N, 3 * D, seed=1, scale=1.5)
    do = rnd(Bn * N, D, seed=2)
    out = torch.full((Bn * N, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(Bn, H, N, device=DEV)
    fwd = lib.theia_attention_tc_fwd
    L.check(fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), Bn, N, H, S()))
    x = qkv.float().view(Bn, N, 3, H, 64).requires_grad_(True)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    s = (q @ k.transpose(2, 3)) * 0.125
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(Bn * N, D)
    assert relerr(out.float(), ref) < 8e-3
    torch.testing.assert_close(lse, torch.logsumexp(s, -1).detach(), rtol=1e-4, atol=1e-4)
    ref.backward(do.float())
    dqkv = torch.full_like(qkv, float("nan"))
    bwd = lib.theia_attention_tc_bwd
    L.check(bwd(qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), Bn, N, H, S()))
    g = x.grad.view(Bn * N, 3, D)
    got = dqkv.float().view(Bn * N, 3, D)
    for i, nm in enumerate("qkv"):
        assert relerr(got[:, i], g[:, i]) < 1.5e-2, nm


@pytest.mark.parametrize("Bn,H,N", [(2, 16, 257), (40, 16, 257), (3, 12, 272), (2, 4, 209), (5, 16, 256),
                                    (11, 16, 258), (3, 6, 259)])  # 257 / 258: tail rows on the CUDA cores; 259: a third tile
def test_attention_fwd_long_sequences(lib, Bn, H, N):
    """209..272 tokens: the 257-token (ViT-L/14 at 224) teacher forward (feature_extraction.py); forward only"""
    D = H * 64
    qkv = rnd(Bn * N, 3 * D, seed=3, scale=1.5)
    out = torch.full((Bn * N, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(Bn, H, N, device=DEV)
    L.check(lib.theia_attention_tc_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), Bn, N, H, S()))
    x = qkv.float().view(Bn, N, 3, H, 64)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    s = (q @ k.transpose(2, 3)) * 0.125
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(Bn * N, D)
    assert relerr(out.float(), ref) < 8e-3
    torch.testing.assert_close(lse, torch.logsumexp(s, -1), rtol=1e-4, atol=1e-4)
    L.check(lib.theia_attention_tc_fwd(qkv.data_ptr(), out.data_ptr(), 0, Bn, N, H, S()))  # lse optional
    assert relerr(out.float(), ref) < 8e-3
    assert lib.theia_attention_tc_fwd(qkv.data_ptr(), out.data_ptr(), 0, Bn, 273, H, S()) != 0


@pytest.mark.parametrize("Bn,H,N", [(2, 16, 257), (20, 16, 257), (3, 5, 272), (4, 16, 197), (7, 3, 258), (2, 2, 130)])
def test_attention_fwd_head_dim_80(lib, Bn, H, N):
    """google/vit-huge-patch14-224-in21k (vit.py:36): 16 heads of 80 dims, 257 tokens; dims 64..79 ride in a second,
    32-byte-swizzled operand tile"""
    hd, D = 80, H * 80
    qkv = rnd(Bn * N, 3 * D, seed=5, scale=1.5)
    out = torch.full((Bn * N, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(Bn, H, N, device=DEV)
    L.check(lib.theia_attention_fwd_hd80(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), Bn, N, H, S()))
    x = qkv.float().view(Bn, N, 3, H, hd)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    s = (q @ k.transpose(2, 3)) * hd ** -0.5
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(Bn * N, D)
    assert relerr(out.float(), ref) < 8e-3
    # per 16-column group: a wrong extra tile would show up in columns 64..79 of every head only
    gn = lambda t: t.view(Bn * N, H, 5, 16).double().pow(2).sum(dim=(0, 1, 3)).sqrt()
    g = gn(out.float() - ref) / gn(ref)
    assert g.max().item() < 1e-2, g.tolist()
    torch.testing.assert_close(lse, torch.logsumexp(s, -1), rtol=1e-4, atol=1e-4)
    assert lib.theia_attention_fwd_hd80(qkv.data_ptr(), out.data_ptr(), 0, Bn, 273, H, S()) != 0


# ----------------------------------------------------------------------------- packing / reductions
def test_pack_and_reductions(lib):
    w = rnd(96, 160, seed=1, dtype=torch.float32)
    o = torch.empty(160, 96, dtype=torch.bfloat16, device=DEV)
    L.check(lib.theia_transpose_cast_bf16(w.data_ptr(), o.data_ptr(), 96, 160, S()))
    assert torch.equal(o, w.t().to(torch.bfloat16))
    o2 = torch.empty(96 * 160 + 3, dtype=torch.bfloat16, device=DEV)
    w2 = rnd(96 * 160 + 3, seed=2, dtype=torch.float32)
    L.check(lib.theia_cast_bf16(w2.data_ptr(), o2.data_ptr(), w2.numel(), S()))
    assert torch.equal(o2, w2.to(torch.bfloat16))
    # conv weight [co][ci][3][3] -> dgrad pack D[ci][tap2][co] = W[co][ci][8-tap2]
    Cc = 16
    cw = rnd(Cc, Cc, 3, 3, seed=3, dtype=torch.float32)
    pk = torch.empty(Cc, 9, Cc, dtype=torch.bfloat16, device=DEV)
    L.check(lib.theia_gather4(cw.data_ptr(), pk.data_ptr(), 1, 0, Cc, 9, Cc, 1, 9, -1, 9 * Cc, 0, 8, S()))
    assert torch.equal(pk, cw.flip(2, 3).reshape(Cc, Cc, 9).permute(1, 2, 0).to(torch.bfloat16))
    x = rnd(985, 200, seed=4)
    cs = torch.zeros(200, device=DEV)
    L.check(lib.theia_colsum(x.data_ptr(), cs.data_ptr(), 985, 200, 200, 197, S()))
    mask = (torch.arange(985, device=DEV) % 197 != 0).float()[:, None]
    assert relerr(cs, (x.float() * mask).sum(0)) < 1e-5
    bs = torch.empty(197 * 64, device=DEV)
    xb = rnd(7, 197 * 64, seed=5)
    L.check(lib.theia_batchsum(xb.data_ptr(), bs.data_ptr(), 7, 197 * 64, S()))
    assert relerr(bs, xb.float().sum(0)) < 1e-5


# ----------------------------------------------------------------------------- stride-2 transposed convs (64x64 heads)
def _classes(p):
    """output parity -> list of (kh, dh) with input index = sub-grid index + dh (kernel 3, stride 2, padding p)"""
    return {par: [(kh, (par + p - kh) // 2) for kh in range(3) if (par + p - kh) % 2 == 0] for par in (0, 1)}


def _convt_geom(Cc, Bn, Hin, pitch_in, tile_w, tile_h, sub_h, sub_w, Hout_pitch, py, px, taps):
    g = L.ConvGeom()
    g.C, g.H, g.W, g.B = Cc, Hin, Hin, Bn
    g.stride_w, g.stride_h, g.stride_b = Cc, pitch_in * Cc, pitch_in * pitch_in * Cc
    g.ntaps = len(taps)
    for t, (wt_idx, dh, dw) in enumerate(taps):
        g.dh[t], g.dw[t], g.wtap[t] = dh, dw, wt_idx
    g.tile_w, g.tile_h = tile_w, tile_h
    g.out_h, g.out_w = sub_h, sub_w
    g.out_img_rows, g.out_row_off, g.out_wpitch = Hout_pitch * Hout_pitch, 0, Hout_pitch
    g.sy = g.sx = 2
    g.py, g.px = py, px
    g.in_stride = 1
    g.b_tap_rows = Cc
    return g


@pytest.mark.parametrize("which", ["t1_16to31", "t2_31to64"])
def test_convtranspose_stride2_fwd_dgrad_wgrad(lib, which, cta_mode):
    Cc, Bn = 128, 2
    if which == "t1_16to31":
        Hin, pin, Hout, pout, pad, opad, tw, th = 16, 16, 31, 32, 1, 0, 16, 8
    else:
        Hin, pin, Hout, pout, pad, opad, tw, th = 31, 32, 64, 64, 0, 1, 32, 4
    xin = torch.zeros(Bn, pin, pin, Cc, dtype=torch.bfloat16, device=DEV)
    xin[:, :Hin, :Hin] = rnd(Bn, Hin, Hin, Cc, seed=1)
    wt = rnd(Cc, Cc, 3, 3, seed=2, scale=0.05, dtype=torch.float32).to(torch.bfloat16).float()  # [ci,co,kh,kw]
    bias = rnd(Cc, seed=3, dtype=torch.float32)
    wF = wt.permute(2, 3, 1, 0).reshape(9, Cc, Cc).to(torch.bfloat16).contiguous()  # [tap][co][ci]
    wD = wt.permute(2, 3, 0, 1).reshape(9, Cc, Cc).to(torch.bfloat16).contiguous()  # [tap][ci][co]
    # ---------------- forward: 4 output-parity classes, each a dense stride-1 gather ----------------
    y = torch.zeros(Bn, pout, pout, Cc, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(Bn, 2, device=DEV)
    cls = _classes(pad)
    for py in (0, 1):
        for px in (0, 1):
            taps = [(kh * 3 + kw, dh, dw) for kh, dh in cls[py] for kw, dw in cls[px]]
            sub_h, sub_w = (Hout - 1 - py) // 2 + 1, (Hout - 1 - px) // 2 + 1
            g = _convt_geom(Cc, Bn, Hin, pin, tw, th, sub_h, sub_w, pout, py, px, taps)
            tiles = (sub_h + th - 1) // th
            gemm(lib, M=Bn * tiles * 128, N=Cc, K=len(taps) * Cc, a_mode=L.OP_CONV_K, A=xin, B=wF, conv=g, out=y, ldo=Cc,
                 bias=bias, stats=stats, epi=L.EPI_RELU | L.EPI_STATS)
    xr = xin[:, :Hin, :Hin].float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    pre = F.conv_transpose2d(xr, wr, bias, stride=2, padding=pad, output_padding=opad)
    ref = F.relu(pre).permute(0, 2, 3, 1)
    assert relerr(y[:, :Hout, :Hout].float(), ref) < 6e-3
    if pout > Hout:
        assert torch.all(y[:, Hout:] == 0) and torch.all(y[:, :, Hout:] == 0)  # padding row/col never written
    o = y.float().view(Bn, -1)
    torch.testing.assert_close(stats[:, 0], o.sum(1), rtol=2e-4, atol=2e-2)
    # ---------------- backward operands ----------------
    dy = torch.zeros(Bn, pout, pout, Cc, dtype=torch.bfloat16, device=DEV)
    dy[:, :Hout, :Hout] = rnd(Bn, Hout, Hout, Cc, seed=5)
    (F.conv_transpose2d(xr, wr, None, stride=2, padding=pad, output_padding=opad) *
     dy[:, :Hout, :Hout].float().permute(0, 3, 1, 2)).sum().backward()
    # dgrad: dX[ih] = sum_k dY[2 ih - p + k] W : one GEMM, 9 taps, TMA element stride 2
    taps9 = [(t, t // 3 - pad, t % 3 - pad) for t in range(9)]
    g = _convt_geom(Cc, Bn, pout if pout > Hout else Hout, pout, tw, th, Hin, Hin, pin, 0, 0, taps9)
    g.sy = g.sx = 1
    g.in_stride = 2
    tiles = (Hin + th - 1) // th
    dx = torch.zeros(Bn, pin, pin, Cc, dtype=torch.bfloat16, device=DEV)
    gemm(lib, M=Bn * tiles * 128, N=Cc, K=9 * Cc, a_mode=L.OP_CONV_K, A=dy, B=wD, conv=g, out=dx, ldo=Cc)
    assert relerr(dx[:, :Hin, :Hin].float(), xr.grad.permute(0, 2, 3, 1)) < 6e-3
    # wgrad: ws[tap][ci][co] = sum_pix X[pix, ci] * dY[2 pix - p + k, co]
    ws = torch.zeros(9, Cc, Cc, device=DEV)
    g2 = _convt_geom(Cc, Bn, pout if pout > Hout else Hout, pout, tw, 64 // tw, pin, pin, pin, 0, 0, taps9)
    g2.in_stride = 2
    gemm(lib, M=Cc, N=Cc, K=Bn * pin * pin, a_mode=L.OP_MN2D, b_mode=L.OP_CONV_MN, A=xin, lda=Cc, B=dy, conv=g2, out=ws,
         ldo=Cc, epi=L.EPI_ATOMIC, batch_z=9, out_z_stride=Cc * Cc, splits=2, bn=128)
    assert relerr(ws, wr.grad.permute(2, 3, 0, 1).reshape(9, Cc, Cc)) < 1e-4


def test_ln3d_padded_31_in_32(lib):
    """LayerNorm([C,31,31]) of the 64x64 heads, stored zero-padded at pitch 32 (NHWC)."""
    Bn, Cc, Wv, Wp = 5, 64, 31, 32
    xv = F.relu(rnd(Bn, Wv, Wv, Cc, seed=1))
    dyv = rnd(Bn, Wv, Wv, Cc, seed=2)
    x = torch.zeros(Bn, Wp, Wp, Cc, dtype=torch.bfloat16, device=DEV)
    dy = torch.zeros_like(x)
    x[:, :Wv, :Wv], dy[:, :Wv, :Wv] = xv, dyv
    g_chw = 1 + 0.1 * rnd(Cc, Wv, Wv, seed=3, dtype=torch.float32)
    b_chw = 0.1 * rnd(Cc, Wv, Wv, seed=4, dtype=torch.float32)
    g_hwc, b_hwc = torch.empty(Wp * Wp * Cc, device=DEV), torch.empty(Wp * Wp * Cc, device=DEV)
    L.check(lib.theia_chw_to_hwc(g_chw.data_ptr(), g_hwc.data_ptr(), Cc, Wv, Wv, Wp, Wp, S()))
    L.check(lib.theia_chw_to_hwc(b_chw.data_ptr(), b_hwc.data_ptr(), Cc, Wv, Wv, Wp, Wp, S()))
    assert torch.equal(g_hwc.view(Wp, Wp, Cc)[:Wv, :Wv], g_chw.permute(1, 2, 0)) and g_hwc.view(Wp, Wp, Cc)[Wv:].abs().sum() == 0
    back = torch.empty_like(g_chw)
    L.check(lib.theia_hwc_to_chw(g_hwc.data_ptr(), back.data_ptr(), Cc, Wv, Wv, Wp, Wp, S()))
    assert torch.equal(back, g_chw)
    xf = xv.float().reshape(Bn, -1)
    stats = torch.stack([xf.sum(1), (xf * xf).sum(1)], 1).contiguous()
    n = Wp * Wp * Cc
    y = torch.full_like(x, float("nan"))
    L.check(lib.theia_ln3d_apply(x.data_ptr(), stats.data_ptr(), g_hwc.data_ptr(), b_hwc.data_ptr(), y.data_ptr(), Bn, n,
                                 1e-5, Cc, Wp, Wv, S()))
    xr = xv.float().permute(0, 3, 1, 2).clone().requires_grad_(True)  # [B,C,31,31]
    gr, br = g_chw.clone().requires_grad_(True), b_chw.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (Cc, Wv, Wv), gr, br, 1e-5)
    assert relerr(y[:, :Wv, :Wv].float(), ref.permute(0, 2, 3, 1)) < 5e-3
    assert torch.all(y[:, Wv:] == 0) and torch.all(y[:, :, Wv:] == 0)
    ref.backward(dyv.float().permute(0, 3, 1, 2))
    red = torch.empty(Bn, 2, device=DEV)
    dx = torch.full_like(x, float("nan"))
    dg, db = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    L.check(lib.theia_ln3d_bwd(dy.data_ptr(), x.data_ptr(), stats.data_ptr(), g_hwc.data_ptr(), red.data_ptr(),
                               dx.data_ptr(), dg.data_ptr(), db.data_ptr(), Bn, n, 1e-5, 1, Cc, Wp, Wv, S()))
    want = (xr.grad * (xr > 0)).permute(0, 2, 3, 1)
    assert relerr(dx[:, :Wv, :Wv].float(), want) < 6e-3
    assert torch.all(dx[:, Wv:] == 0) and torch.all(dx[:, :, Wv:] == 0)
    assert relerr(dg.view(Wp, Wp, Cc)[:Wv, :Wv], gr.grad.permute(1, 2, 0)) < 1e-3
    assert relerr(db.view(Wp, Wp, Cc)[:Wv, :Wv], br.grad.permute(1, 2, 0)) < 1e-3


@pytest.mark.parametrize("C_,H_", [(1024, 16), (256, 64), (32, 64), (1280, 16)])
def test_target_ingest_bit_exact_with_reference_bf16_math(lib, C_, H_):
    """dataset/data_utils.py:152-153 (rearrange 'c h w -> (h w) c') + :342-355 ((x - mean) / std in bf16)."""
    from theia_b200.data import ingest_targets
    Bn = 3
    emb = rnd(Bn, C_, H_, H_, seed=1, scale=3.0)
    mean, std = rnd(C_, seed=2), (rnd(C_, seed=3).float().abs() + 0.5).to(torch.bfloat16)
    got = ingest_targets(emb, mean, std)
    want = (emb.flatten(2).transpose(1, 2) - mean) / std  # torch bf16 arithmetic, exactly the reference's
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == (Bn, H_ * H_, C_)
    assert torch.equal(got, want)
    assert torch.equal(ingest_targets(emb), emb.flatten(2).transpose(1, 2))
