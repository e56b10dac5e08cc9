"""GPU parity tests, one per kernel family, called through the C ABI (ctypes -> libb200dm.so) and
checked against the torch-fp32 expression of the reference lines each kernel replaces
(models/generative/diffusion/ddpm.py; the same expressions the oracle uses).

Tolerances: fp32 kernels 1e-5..1e-4 relative (accumulation order only); bf16 kernels are compared with
the fp32 expression evaluated on the bf16-rounded inputs, so the only differences are the fp32
accumulation order and the final bf16 rounding of the output (rel-L2 <= 4e-3).
"""
import ctypes
import math

import pytest
import torch
import torch.nn.functional as F

from _refcuda import report
from b200dm import _lib as L
from b200dm.tensor import View

pytestmark = pytest.mark.gpu
DEV = "cuda"
# the torch expressions below are the fp32 yardstick: keep cuDNN/cuBLAS from silently using TF32
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DT = {L.F32: torch.float32, L.BF16: torch.bfloat16}


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def tol(dtype):
    return 2e-5 if dtype == L.F32 else 4e-3


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def q(x, dtype):  # round to the storage dtype
    return x.to(DT[dtype]).float()


def nhwc(x_nchw, dtype, ld=None, off=0):
    B, C, H, W = x_nchw.shape
    v = View.zeros(B, H, W, C, DT[dtype], DEV, ld=ld, off=off)
    v.buf.fill_(7.0)  # poison the padding channels
    return v.from_nchw(x_nchw)


# ------------------------------------------------------------------------------------------------
# scheduler / loss kernels
# ------------------------------------------------------------------------------------------------
def noise_desc(buf, img, t, noise, normalize, seed=0, stream_id=0, offset=None, strength=0.0):
    d = L.NoiseDesc(img=img.data_ptr(), t=t.data_ptr(), noise=None if noise is None else noise.data_ptr(),
                    offset=None if offset is None else offset.data_ptr(),
                    sqrt_ac=buf["sqrt_alphas_cumprod"].data_ptr(),
                    sqrt_1mac=buf["sqrt_one_minus_alphas_cumprod"].data_ptr(), offset_strength=strength,
                    normalize=normalize, B=img.shape[0], chw=img[0].numel(), hw=img.shape[-1] * img.shape[-2],
                    seed=seed, stream_id=stream_id, elem_offset=0)
    d._keep = (img, t, noise, offset)
    return d


def _buffers(schedule="sigmoid", objective="pred_v"):
    from b200dm.schedule import make_buffers
    return {k: v.to(DEV) for k, v in make_buffers(1000, schedule, objective).items()}


def test_q_sample_injected_noise_exact():
    buf = _buffers()
    B, C, S = 5, 3, 32
    img, noise = torch.rand(B, C, S, S, device=DEV), rnd(B, C, S, S, seed=1)
    t = torch.tensor([0, 999, 17, 500, 998], device=DEV)
    xt, eo, x0 = (torch.empty_like(img) for _ in range(3))
    d = noise_desc(buf, img, t, noise, normalize=1)
    L.call("b200dm_q_sample", ctypes.byref(d), xt.data_ptr(), eo.data_ptr(), x0.data_ptr())
    xs = img * 2 - 1
    ref = (buf["sqrt_alphas_cumprod"][t].view(B, 1, 1, 1) * xs
           + buf["sqrt_one_minus_alphas_cumprod"][t].view(B, 1, 1, 1) * noise)
    assert torch.equal(x0, xs) and torch.equal(eo, noise)
    assert (xt - ref).abs().max().item() <= 2.4e-7 * 4  # <= 1 ulp at |x| < 4 (FMA contraction in torch)


def test_philox_noise_statistics_and_shard_invariance():
    n = 1 << 22
    a = torch.empty(n, device=DEV)
    L.call("b200dm_randn", a.data_ptr(), n, 1234, 7, 0)
    assert abs(a.mean().item()) < 3e-3 and abs(a.std().item() - 1) < 3e-3
    assert abs((a ** 3).mean().item()) < 1e-2 and abs((a ** 4).mean().item() - 3) < 3e-2
    # the same stream generated in two shards (elem_offset) is bit-identical
    b = torch.empty(n, device=DEV)
    h = n // 2
    L.call("b200dm_randn", b.data_ptr(), h, 1234, 7, 0)
    L.call("b200dm_randn", b.data_ptr() + 4 * h, h, 1234, 7, h)
    assert torch.equal(a, b)
    c = torch.empty(n, device=DEV)
    L.call("b200dm_randn", c.data_ptr(), n, 1234, 8, 0)       # another stream id decorrelates
    assert abs((a * c).mean().item()) < 3e-3
    # q_sample without an injected noise pointer uses the same generator
    buf = _buffers()
    img = torch.rand(4, 3, 32, 32, device=DEV)
    t = torch.tensor([10, 20, 30, 40], device=DEV)
    xt, eo = torch.empty_like(img), torch.empty_like(img)
    d = noise_desc(buf, img, t, None, normalize=1, seed=1234, stream_id=7)
    L.call("b200dm_q_sample", ctypes.byref(d), xt.data_ptr(), eo.data_ptr(), None)
    assert torch.equal(eo.flatten(), a[:eo.numel()])
    # ... and the loss kernel regenerates exactly that eps from the descriptor (nothing stored in between):
    # with out = target of pred_noise = eps the loss is exactly 0
    acc = torch.zeros(1, device=DEV)
    L.call("b200dm_loss_fwd_bwd", ctypes.byref(d), eo.data_ptr(), buf["loss_weight"].data_ptr(), acc.data_ptr(), None,
           L.OBJECTIVES["pred_noise"])
    assert acc.item() == 0.0


@pytest.mark.parametrize("objective", ["pred_noise", "pred_x0", "pred_v"])
def test_loss_fwd_bwd(objective):
    buf = _buffers(objective=objective)
    B, C, S = 6, 3, 32
    out = rnd(B, C, S, S, seed=2).requires_grad_(True)
    x0, noise = rnd(B, C, S, S, seed=3), rnd(B, C, S, S, seed=4)
    t = torch.tensor([0, 999, 17, 500, 998, 250], device=DEV)
    sa = buf["sqrt_alphas_cumprod"][t].view(B, 1, 1, 1)
    sb = buf["sqrt_one_minus_alphas_cumprod"][t].view(B, 1, 1, 1)
    target = {"pred_noise": noise, "pred_x0": x0, "pred_v": sa * noise - sb * x0}[objective]
    loss = (F.mse_loss(out, target, reduction="none").flatten(1).mean(1) * buf["loss_weight"][t]).mean()
    loss.backward()
    acc = torch.zeros(1, device=DEV)
    dout = torch.empty_like(out)
    d = noise_desc(buf, x0, t, noise, normalize=0)
    L.call("b200dm_loss_fwd_bwd", ctypes.byref(d), out.data_ptr(), buf["loss_weight"].data_ptr(), acc.data_ptr(),
           dout.data_ptr(), L.OBJECTIVES[objective])
    assert abs(acc.item() - loss.item()) <= 2e-6 * abs(loss.item()) + 1e-9
    assert rel(dout, out.grad) < 1e-6


def test_offset_noise_q_sample_and_loss():
    """ddpm.py:889-891: noise += strength * randn(b, c)[:, :, None, None], used by q_sample AND the target."""
    buf = _buffers(objective="pred_v")
    B, Cc, S = 4, 3, 32
    img, noise, out = torch.rand(B, Cc, S, S, device=DEV), rnd(B, Cc, S, S, seed=11), rnd(B, Cc, S, S, seed=12)
    off = rnd(B, Cc, seed=13)
    t = torch.tensor([3, 999, 400, 77], device=DEV)
    d = noise_desc(buf, img, t, noise, normalize=1, offset=off, strength=0.1)
    xt, eo = torch.empty_like(img), torch.empty_like(img)
    L.call("b200dm_q_sample", ctypes.byref(d), xt.data_ptr(), eo.data_ptr(), None)
    n2 = noise + 0.1 * off.view(B, Cc, 1, 1)
    sa = buf["sqrt_alphas_cumprod"][t].view(B, 1, 1, 1)
    sb = buf["sqrt_one_minus_alphas_cumprod"][t].view(B, 1, 1, 1)
    xs = img * 2 - 1
    assert (eo - n2).abs().max().item() <= 1e-6
    assert (xt - (sa * xs + sb * n2)).abs().max().item() <= 2e-6
    acc = torch.zeros(1, device=DEV)
    L.call("b200dm_loss_fwd_bwd", ctypes.byref(d), out.data_ptr(), buf["loss_weight"].data_ptr(), acc.data_ptr(), None,
           L.OBJECTIVES["pred_v"])
    target = sa * n2 - sb * xs
    ref = (F.mse_loss(out, target, reduction="none").flatten(1).mean(1) * buf["loss_weight"][t]).mean()
    assert abs(acc.item() - ref.item()) <= 1e-5 * abs(ref.item())


@pytest.mark.parametrize("objective", ["pred_noise", "pred_x0", "pred_v"])
@pytest.mark.parametrize("last,sigma", [(0, 0.0), (0, 0.3), (1, 0.0)])
def test_ddim_step(objective, last, sigma):
    buf = _buffers("linear", objective)
    n = 2 * 3 * 32 * 32
    x, out, z = rnd(n, seed=5), rnd(n, seed=6), rnd(n, seed=7)
    t = 640
    sac, s1m = buf["sqrt_alphas_cumprod"][t], buf["sqrt_one_minus_alphas_cumprod"][t]
    sr, srm1 = buf["sqrt_recip_alphas_cumprod"][t], buf["sqrt_recipm1_alphas_cumprod"][t]
    if objective == "pred_noise":
        x0 = (sr * x - srm1 * out).clamp(-1, 1)
    elif objective == "pred_x0":
        x0 = out.clamp(-1, 1)
    else:
        x0 = (sac * x - s1m * out).clamp(-1, 1)
    eps = (sr * x - x0) / srm1
    san, c = 0.8, 0.55
    ref = x0 if last else x0 * san + c * eps + sigma * z
    xn, x0o = torch.empty_like(x), torch.empty_like(x)
    L.call("b200dm_ddim_step", x.data_ptr(), out.data_ptr(), z.data_ptr(), xn.data_ptr(), x0o.data_ptr(),
           sac.item(), s1m.item(), sr.item(), srm1.item(), san, c, sigma, last, L.OBJECTIVES[objective],
           n, 0, 0, 0)
    assert (x0o - x0).abs().max().item() < 1e-5
    assert (xn - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("objective", ["pred_noise", "pred_x0", "pred_v"])
@pytest.mark.parametrize("add_noise", [0, 1])
def test_ddpm_step(objective, add_noise):
    buf = _buffers("sigmoid", objective)
    n = 2 * 3 * 32 * 32
    x, out, z = rnd(n, seed=8), rnd(n, seed=9), rnd(n, seed=10)
    t = 333
    sac, s1m = buf["sqrt_alphas_cumprod"][t], buf["sqrt_one_minus_alphas_cumprod"][t]
    sr, srm1 = buf["sqrt_recip_alphas_cumprod"][t], buf["sqrt_recipm1_alphas_cumprod"][t]
    if objective == "pred_noise":
        x0 = sr * x - srm1 * out
    elif objective == "pred_x0":
        x0 = out
    else:
        x0 = sac * x - s1m * out
    x0 = x0.clamp(-1, 1)
    c1, c2 = buf["posterior_mean_coef1"][t], buf["posterior_mean_coef2"][t]
    std = (0.5 * buf["posterior_log_variance_clipped"][t]).exp()
    ref = c1 * x0 + c2 * x + (std * z if add_noise else 0.0)
    xp, x0o = torch.empty_like(x), torch.empty_like(x)
    L.call("b200dm_ddpm_step", x.data_ptr(), out.data_ptr(), z.data_ptr(), xp.data_ptr(), x0o.data_ptr(),
           sac.item(), s1m.item(), sr.item(), srm1.item(), c1.item(), c2.item(), std.item(), add_noise,
           L.OBJECTIVES[objective], n, 0, 0, 0)
    assert (x0o - x0).abs().max().item() < 1e-5 and (xp - ref).abs().max().item() < 1e-5


def test_elementwise_argument_errors():
    x = torch.zeros(6, device=DEV)
    with pytest.raises(L.B200dmError):
        L.call("b200dm_randn", x.data_ptr(), 6, 0, 0, 0)          # not a multiple of 4
    with pytest.raises(L.B200dmError):
        L.call("b200dm_unnormalize", x.data_ptr() + 4, x.data_ptr(), 4)  # misaligned


# ------------------------------------------------------------------------------------------------
# convolutions
# ------------------------------------------------------------------------------------------------
def pack_w(w_oihw, dtype, transpose=False, flip=False):
    """OIHW fp32 -> [taps][Cout][Cin] (or the dgrad operand [taps'][Cin][Cout]) in the storage dtype."""
    co, ci, kh, kw = w_oihw.shape
    p = w_oihw.permute(2, 3, 0, 1).reshape(kh * kw, co, ci)
    if transpose:
        p = p.transpose(1, 2)
        if flip:
            p = p.flip(0)
    return p.contiguous().to(DT[dtype])


def run_conv(dtype, impl, mode, ksize, xv, w_packed, bias, yv, Cin, Cout, H, W, res=None, accumulate=0):
    d = L.ConvDesc(dtype=dtype, mode=mode, ksize=ksize, impl=impl, B=xv.B, H=H, W=W, Cin=Cin, Cout=Cout,
                   x=xv.ptr, x_ld=xv.ld, w=w_packed.data_ptr(), bias=L.ptr(bias), y=yv.ptr, y_ld=yv.ld,
                   res=None if res is None else res.ptr, res_ld=0 if res is None else res.ld,
                   accumulate=accumulate)
    L.call("b200dm_conv_fwd", d)


CONV_CASES = [  # B, H, Cin, Cout, ksize
    (2, 32, 64, 64, 3), (3, 16, 128, 128, 3), (2, 8, 192, 128, 3), (3, 4, 256, 512, 3),
    (9, 4, 512, 512, 3), (2, 16, 64, 384, 1), (1, 64, 128, 64, 3), (5, 8, 384, 256, 1),
]


def _impls():
    return [0, 1]


def _skip_tc(impl, dtype):
    if impl == 1:
        if dtype != L.BF16:
            pytest.skip("tcgen05 path is bf16 only")
        assert L.load().b200dm_tc_available() == 1, "tcgen05 path unavailable on this GPU box"


@pytest.mark.parametrize("impl", [0, 1], ids=["simt", "tc"])
@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_same(impl, dtype, case):
    _skip_tc(impl, dtype)
    B, S, Cin, Cout, k = case
    x = q(rnd(B, Cin, S, S, seed=11), dtype)
    w = q(rnd(Cout, Cin, k, k, seed=12, scale=1 / math.sqrt(Cin * k * k)), dtype)
    bias = rnd(Cout, seed=13)
    ref = F.conv2d(x, w, bias, padding=k // 2)
    xv = nhwc(x, dtype, ld=Cin + 64, off=32)            # channel slice of a wider (concat) buffer
    yv = View.zeros(B, S, S, Cout, DT[dtype], DEV, ld=Cout + 8, off=8)
    run_conv(dtype, impl, 0, k, xv, pack_w(w, dtype), bias, yv, Cin, Cout, S, S)
    assert rel(yv.to_nchw(), ref) < tol(dtype)
    # residual + accumulate epilogue
    r = q(rnd(B, Cout, S, S, seed=14), dtype)
    y0 = q(rnd(B, Cout, S, S, seed=15), dtype)
    rv = nhwc(r, dtype)
    yv.from_nchw(y0)
    run_conv(dtype, impl, 0, k, xv, pack_w(w, dtype), None, yv, Cin, Cout, S, S, res=rv, accumulate=1)
    ref2 = F.conv2d(x, w, None, padding=k // 2) + r + y0
    assert rel(yv.to_nchw(), ref2) < tol(dtype)


@pytest.mark.parametrize("impl", [0, 1], ids=["simt", "tc"])
@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
@pytest.mark.parametrize("case", [(2, 32, 64, 64), (3, 16, 64, 128), (5, 8, 128, 256)])
def test_conv_unshuffle_and_its_transpose(impl, dtype, case):
    """Downsample (ddpm.py:100-104): 'b c (h p1) (w p2) -> b (c p1 p2) h w' + 1x1 conv == mode 1;
    its data gradient == mode 2."""
    _skip_tc(impl, dtype)
    B, S, C, Cout = case          # input [B, C, S, S] -> output [B, Cout, S/2, S/2]
    x = q(rnd(B, C, S, S, seed=21), dtype).requires_grad_(True)
    w = q(rnd(Cout, 4 * C, 1, 1, seed=22, scale=1 / math.sqrt(4 * C)), dtype)
    bias = rnd(Cout, seed=23)
    xu = x.reshape(B, C, S // 2, 2, S // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, 4 * C, S // 2, S // 2)
    ref = F.conv2d(xu, w, bias)
    # packed [tap = p1*2+p2][Cout][C] from W[Cout][c*4 + p1*2 + p2]
    wp = w.reshape(Cout, C, 4).permute(2, 0, 1).contiguous()
    xv = nhwc(x.detach(), dtype, ld=C + 64, off=0)
    yv = View.zeros(B, S // 2, S // 2, Cout, DT[dtype], DEV)
    run_conv(dtype, impl, 1, 1, xv, wp.to(DT[dtype]), bias, yv, C, Cout, S // 2, S // 2)
    assert rel(yv.to_nchw(), ref) < tol(dtype)
    # transpose: dX from dY
    dy = q(rnd(B, Cout, S // 2, S // 2, seed=24), dtype)
    ref.backward(dy)
    wt = wp.transpose(1, 2).contiguous()         # [tap][C][Cout]: GEMM columns (tap, c), K = Cout
    dyv = nhwc(dy, dtype)
    dxv = View.zeros(B, S, S, C, DT[dtype], DEV, ld=C + 8, off=8)
    run_conv(dtype, impl, 2, 1, dyv, wt.to(DT[dtype]), None, dxv, Cout, C, S // 2, S // 2)
    assert rel(dxv.to_nchw(), x.grad) < tol(dtype)


@pytest.mark.parametrize("case", [(2, 32, 128, 64), (3, 16, 256, 128), (5, 8, 512, 256), (2, 4, 64, 64)])
def test_conv_fused_nearest_upsample_3x3(case):
    """Upsample (ddpm.py:93-97): nn.Upsample(scale_factor=2, mode='nearest') + Conv2d(3x3, padding=1) as ONE launch
    over the low-resolution tensor (conv mode 3: one 2x2 conv per output phase, weights summed by
    b200dm_pack_upconv_weight)."""
    _skip_tc(1, L.BF16)
    B, S, Cin, Cout = case                   # input [B, Cin, S, S] -> output [B, Cout, 2S, 2S]
    x = q(rnd(B, Cin, S, S, seed=31), L.BF16)
    w = rnd(Cout, Cin, 3, 3, seed=32, scale=1 / math.sqrt(9 * Cin))
    bias = rnd(Cout, seed=33)
    master = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()          # arena layout [ky*3+kx][Cout][Cin]
    wp = torch.full((16 * Cout * Cin,), 7.0, dtype=torch.bfloat16, device=DEV)
    L.call("b200dm_pack_upconv_weight", master.data_ptr(), wp.data_ptr(), Cout, Cin)
    sel = ([[0], [1, 2]], [[0, 1], [2]])      # phase -> 2x2 tap -> 3x3 taps that land on the same source pixel
    exp = torch.empty(4, 4, Cout, Cin, device=DEV)
    for a in range(2):
        for b in range(2):
            for r in range(2):
                for c in range(2):
                    acc = torch.zeros(Cout, Cin, device=DEV)          # same fp32 summation order as the kernel
                    for ky in sel[a][r]:
                        for kx in sel[b][c]:
                            acc = acc + w[:, :, ky, kx]
                    exp[r * 2 + c, a * 2 + b] = acc
    assert torch.equal(wp.view(4, 4, Cout, Cin).float(), exp.to(torch.bfloat16).float())
    xv = nhwc(x, L.BF16, ld=Cin + 64, off=0)
    yv = View.zeros(B, 2 * S, 2 * S, Cout, torch.bfloat16, DEV, ld=Cout + 64, off=64)
    run_conv(L.BF16, 1, 3, 3, xv, wp, bias, yv, Cin, Cout, S, S)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, bias, padding=1)
    assert rel(yv.to_nchw(), ref) < tol(L.BF16)
    with pytest.raises(L.B200dmError):        # the SIMT path does not build it
        run_conv(L.BF16, 0, 3, 3, xv, wp, bias, yv, Cin, Cout, S, S)


@pytest.mark.parametrize("impl", [0, 1], ids=["simt", "tc"])
@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
@pytest.mark.parametrize("case", [(2, 32, 64, 64, 3), (3, 8, 128, 64, 3), (4, 4, 256, 128, 3),
                                  (2, 16, 128, 192, 1)])
def test_conv_dgrad_via_flipped_weights(impl, dtype, case):
    _skip_tc(impl, dtype)
    B, S, Cin, Cout, k = case
    x = q(rnd(B, Cin, S, S, seed=31), dtype).requires_grad_(True)
    w = q(rnd(Cout, Cin, k, k, seed=32, scale=1 / math.sqrt(Cin * k * k)), dtype)
    dy = q(rnd(B, Cout, S, S, seed=33), dtype)
    F.conv2d(x, w, None, padding=k // 2).backward(dy)
    dyv = nhwc(dy, dtype)
    dxv = View.zeros(B, S, S, Cin, DT[dtype], DEV)
    run_conv(dtype, impl, 0, k, dyv, pack_w(w, dtype, transpose=True, flip=True), None, dxv, Cout, Cin, S, S)
    assert rel(dxv.to_nchw(), x.grad) < tol(dtype)


@pytest.mark.parametrize("impl", [0, 1], ids=["simt", "tc"])
@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
@pytest.mark.parametrize("case", [(2, 32, 64, 64, 3, 0), (3, 16, 128, 192, 3, 0), (9, 4, 256, 128, 3, 0),
                                  (2, 8, 192, 64, 1, 0), (3, 16, 64, 128, 1, 1), (2, 64, 64, 64, 3, 0)])
def test_conv_wgrad(impl, dtype, case):
    _skip_tc(impl, dtype)
    B, S, Cin, Cout, k, mode = case
    if mode == 0:
        x = q(rnd(B, Cin, S, S, seed=41), dtype)
        w = rnd(Cout, Cin, k, k, seed=42).requires_grad_(True)
        dy = q(rnd(B, Cout, S, S, seed=43), dtype)
        F.conv2d(x, w, None, padding=k // 2).backward(dy)
        ref = w.grad.permute(2, 3, 0, 1).reshape(k * k, Cout, Cin)
        xv, dyv, H = nhwc(x, dtype, ld=Cin + 8, off=8), nhwc(dy, dtype), S
    else:  # unshuffle: x is [B, Cin, S, S], output spatial S/2
        x = q(rnd(B, Cin, S, S, seed=41), dtype)
        w = rnd(Cout, 4 * Cin, 1, 1, seed=42).requires_grad_(True)
        dy = q(rnd(B, Cout, S // 2, S // 2, seed=43), dtype)
        xu = x.reshape(B, Cin, S // 2, 2, S // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, 4 * Cin, S // 2, S // 2)
        F.conv2d(xu, w).backward(dy)
        ref = w.grad.reshape(Cout, Cin, 4).permute(2, 0, 1)
        xv, dyv, H = nhwc(x, dtype, ld=Cin + 8, off=0), nhwc(dy, dtype), S // 2
    taps = ref.shape[0]
    dw = torch.full((taps, Cout, Cin), 3.0, device=DEV)      # accumulate = 0 must overwrite
    d = L.WgradDesc(dtype=dtype, mode=mode, ksize=k, impl=impl, B=B, H=H, W=H, Cin=Cin, Cout=Cout,
                    x=xv.ptr, x_ld=xv.ld, dy=dyv.ptr, dy_ld=dyv.ld, dw=dw.data_ptr(), accumulate=0)
    L.call("b200dm_conv_wgrad", d)
    assert rel(dw, ref) < 2e-5 if dtype == L.F32 else rel(dw, ref) < 1e-4
    d.accumulate = 1
    L.call("b200dm_conv_wgrad", d)
    assert rel(dw, 2 * ref) < 1e-4


def test_conv_rejects_bad_shapes():
    x = View.zeros(1, 8, 8, 24, torch.bfloat16, DEV)
    w = torch.zeros(1, 64, 24, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(L.B200dmError):
        run_conv(L.BF16, 1, 0, 1, x, w, None, View.zeros(1, 8, 8, 64, torch.bfloat16, DEV), 24, 64, 8, 8)
    with pytest.raises(L.B200dmError):
        run_conv(L.F32, 1, 0, 1, x, w, None, View.zeros(1, 8, 8, 64, torch.float32, DEV), 64, 64, 8, 8)
    with pytest.raises(L.B200dmError):
        run_conv(L.BF16, 0, 0, 5, x, w, None, View.zeros(1, 8, 8, 64, torch.bfloat16, DEV), 16, 64, 8, 8)


@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
@pytest.mark.parametrize("C,S,B", [(3, 32, 3), (1, 32, 2), (3, 64, 2)])
def test_init_conv_fwd_and_wgrad(dtype, C, S, B):
    x = rnd(B, C, S, S, seed=51)
    w = rnd(64, C, 7, 7, seed=52, scale=0.1).requires_grad_(True)
    bias = rnd(64, seed=53)
    ref = F.conv2d(x, w, bias, padding=3)
    yv = View.zeros(B, S, S, 64, DT[dtype], DEV, ld=128, off=64)
    L.call("b200dm_init_conv_fwd", dtype, x.data_ptr(), w.data_ptr(), bias.data_ptr(), yv.ptr, yv.ld, B, C, S, S, 64)
    assert rel(yv.to_nchw(), ref) < tol(dtype)
    dy = q(rnd(B, 64, S, S, seed=54), dtype)
    ref.backward(dy)
    dw = torch.zeros_like(w)
    dyv = nhwc(dy, dtype)
    L.call("b200dm_init_conv_wgrad", dtype, x.data_ptr(), dyv.ptr, dyv.ld, dw.data_ptr(), B, C, S, S, 64)
    assert rel(dw, w.grad) < 2e-5


@pytest.mark.parametrize("C,S,B", [(3, 32, 4), (1, 32, 3), (3, 16, 5), (3, 64, 2), (2, 8, 3)])
def test_stem_on_tensor_cores_im2col_gemm_and_wgrad(C, S, B):
    """7x7 stem as im2col + tcgen05 1x1 GEMM, and its weight gradient through conv_wgrad(cin_valid)."""
    if not L.load().b200dm_tc_available():
        pytest.skip("needs the tcgen05 path")
    K, KP = C * 49, (C * 49 + 63) // 64 * 64
    x = rnd(B, C, S, S, seed=71)
    w = rnd(64, C, 7, 7, seed=72, scale=0.1)
    bias = rnd(64, seed=73, scale=0.1)
    P = torch.full((B * S * S, KP), 7.0, dtype=torch.bfloat16, device=DEV)
    L.call("b200dm_im2col7", x.data_ptr(), P.data_ptr(), B, C, S, S, KP)
    ref_cols = F.unfold(q(x, L.BF16), 7, padding=3).transpose(1, 2).reshape(B * S * S, K)
    assert torch.equal(P[:, :K].float(), ref_cols) and P[:, K:].abs().max().item() == 0
    wp = torch.empty(64, KP, dtype=torch.bfloat16, device=DEV)
    L.call("b200dm_pack_stem_weight", w.data_ptr(), wp.data_ptr(), 64, K, KP)
    assert torch.equal(wp[:, :K].float(), q(w.reshape(64, K), L.BF16)) and wp[:, K:].abs().max().item() == 0
    yv = View.zeros(B, S, S, 64, torch.bfloat16, DEV, ld=128, off=64)
    d = L.ConvDesc(dtype=L.BF16, mode=0, ksize=1, impl=1, B=B, H=S, W=S, Cin=KP, Cout=64, x=P.data_ptr(), x_ld=KP,
                   w=wp.data_ptr(), bias=bias.data_ptr(), y=yv.ptr, y_ld=yv.ld, res=None, res_ld=0, accumulate=0)
    L.call("b200dm_conv_fwd", d)
    ref = F.conv2d(q(x, L.BF16), q(w, L.BF16), bias, padding=3)
    assert rel(yv.to_nchw(), ref) < tol(L.BF16)
    dy = q(rnd(B, 64, S, S, seed=74), L.BF16)
    dyv = nhwc(dy, L.BF16)
    dw = torch.zeros(64, C, 7, 7, device=DEV)
    guard = dw.clone()
    wd = L.WgradDesc(dtype=L.BF16, mode=0, ksize=1, impl=1, B=B, H=S, W=S, Cin=KP, Cout=64, x=P.data_ptr(), x_ld=KP,
                     dy=dyv.ptr, dy_ld=dyv.ld, dw=dw.data_ptr(), accumulate=1, cin_valid=K, s_tap=64 * K, s_co=K,
                     s_ci=1)
    L.call("b200dm_conv_wgrad", wd)
    xr = q(x, L.BF16).requires_grad_(False)
    wr = torch.zeros(64, C, 7, 7, device=DEV, requires_grad=True)
    F.conv2d(xr, wr, None, padding=3).backward(dy)
    assert rel(dw, wr.grad) < 2e-3
    del guard


@pytest.mark.parametrize("C,S,B", [(3, 32, 4), (1, 32, 3), (3, 64, 2), (3, 16, 5), (2, 128, 1), (3, 64, 37), (4, 16, 2), (4, 32, 1)])
def test_stem_one_launch_patches_in_shared_memory(C, S, B):
    """b200dm_stem7_fwd (csrc/stem_tc.cu) against F.conv2d on the bf16-rounded operands, output as a channel slice."""
    if not L.load().b200dm_tc_available():
        pytest.skip("needs the tcgen05 path")
    KP = (C * 56 + 63) // 64 * 64
    x = rnd(B, C, S, S, seed=81)
    w = rnd(64, C, 7, 7, seed=82, scale=0.1)
    bias = rnd(64, seed=83, scale=0.1)
    wp = torch.full((64, KP), 7.0, dtype=torch.bfloat16, device=DEV)
    L.call("b200dm_pack_stem_rows", w.data_ptr(), wp.data_ptr(), 64, C, KP)
    rows = wp[:, :C * 56].float().view(64, C, 7, 8)
    assert torch.equal(rows[..., :7], q(w, L.BF16)) and rows[..., 7].abs().max().item() == 0
    assert wp[:, C * 56:].abs().max().item() == 0
    yv = View.zeros(B, S, S, 64, torch.bfloat16, DEV, ld=128, off=64)
    yv.buf.fill_(7.0)
    assert L.load().b200dm_stem7_supported(B, C, S, S, KP, yv.ld) == 1
    L.call("b200dm_stem7_fwd", x.data_ptr(), wp.data_ptr(), bias.data_ptr(), yv.ptr, yv.ld, B, C, S, S, KP)
    ref = F.conv2d(q(x, L.BF16), q(w, L.BF16), bias, padding=3)
    assert rel(yv.to_nchw(), ref) < tol(L.BF16)
    assert (yv.buf[..., :64] == 7.0).all()                 # the other half of the concat buffer is untouched
    assert L.load().b200dm_stem7_supported(B, C, 8, 8, KP, yv.ld) == 0       # fewer than 128 pixels per image
    assert L.load().b200dm_stem7_supported(B, C, 48, 48, KP, yv.ld) == 0     # W does not divide 128
    assert L.load().b200dm_stem7_supported(B, 6, 64, 64, 384, yv.ld) == 0    # six channels: K = 384 is more than five blocks


@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
@pytest.mark.parametrize("C", [1, 3])
def test_final_conv_fwd_bwd(dtype, C):
    B, S = 3, 16
    x = q(rnd(B, 64, S, S, seed=61), dtype).requires_grad_(True)
    w = rnd(C, 64, 1, 1, seed=62, scale=0.125).requires_grad_(True)
    bias = rnd(C, seed=63).requires_grad_(True)
    ref = F.conv2d(x, w, bias)
    xv = nhwc(x.detach(), dtype, ld=72, off=8)
    y = torch.empty(B, C, S, S, device=DEV)
    L.call("b200dm_final_conv_fwd", dtype, xv.ptr, xv.ld, w.data_ptr(), bias.data_ptr(), y.data_ptr(), B, S * S, 64, C)
    assert rel(y, ref) < 2e-5
    dy = rnd(B, C, S, S, seed=64)
    ref.backward(dy)
    dxv = View.zeros(B, S, S, 64, DT[dtype], DEV)
    dw, db = torch.zeros(C, 64, device=DEV), torch.zeros(C, device=DEV)
    L.call("b200dm_final_conv_bwd", dtype, xv.ptr, xv.ld, w.data_ptr(), dy.data_ptr(), dxv.ptr, dxv.ld,
           dw.data_ptr(), db.data_ptr(), B, S * S, 64, C)
    assert rel(dxv.to_nchw(), x.grad) < tol(dtype)
    assert rel(dw, w.grad.view(C, 64)) < 2e-5 and rel(db, bias.grad) < 2e-5


@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
def test_upsample_and_colsum(dtype):
    B, S, Cc = 2, 8, 128
    x = q(rnd(B, Cc, S, S, seed=71), dtype)
    xv = nhwc(x, dtype, ld=Cc + 8, off=8)
    yv = View.zeros(B, 2 * S, 2 * S, Cc, DT[dtype], DEV)
    L.call("b200dm_upsample2x_fwd", dtype, xv.ptr, xv.ld, yv.ptr, yv.ld, B, S, S, Cc)
    ref = x.repeat_interleave(2, 2).repeat_interleave(2, 3)
    assert torch.equal(yv.to_nchw(), ref)
    dy = q(rnd(B, Cc, 2 * S, 2 * S, seed=72), dtype)
    dyv = nhwc(dy, dtype)
    dxv = View.zeros(B, S, S, Cc, DT[dtype], DEV)
    L.call("b200dm_upsample2x_bwd", dtype, dyv.ptr, dyv.ld, dxv.ptr, dxv.ld, B, S, S, Cc)
    assert rel(dxv.to_nchw(), F.avg_pool2d(dy, 2) * 4) < tol(dtype)
    out = torch.full((Cc,), 5.0, device=DEV)
    L.call("b200dm_colsum", dtype, dyv.ptr, dyv.ld, B * 4 * S * S, Cc, out.data_ptr(), 0)
    assert rel(out, dy.sum((0, 2, 3))) < 2e-5


@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
def test_colsum_batched_ragged_items(dtype):
    """One launch, items of very different sizes (one with a single row, one strided slice of a wider buffer, one with
    more than 512 channels = several column blocks); every item accumulates onto what `out` already holds."""
    shapes = [(2, 32, 64, 0), (1, 1, 8, 0), (3, 8, 512, 0), (2, 16, 128, 8), (1, 4, 1024, 0), (5, 2, 64, 0)]
    arr = (L.ColsumItem * len(shapes))()
    keep, want, outs = [], [], []
    for i, (B, S, Cc, off) in enumerate(shapes):
        dy = q(rnd(B, Cc, S, S, seed=300 + i), dtype)
        v = nhwc(dy, dtype, ld=Cc + off, off=off) if off else nhwc(dy, dtype)
        out = torch.full((Cc,), float(i), device=DEV)
        arr[i].x, arr[i].out, arr[i].rows, arr[i].ld, arr[i].C = v.ptr, out.data_ptr(), B * S * S, v.ld, Cc
        keep.append(v)
        outs.append(out)
        want.append(dy.double().sum((0, 2, 3)) + i)
    L.call("b200dm_colsum_batched", dtype, arr, len(shapes))
    torch.cuda.synchronize()
    for out, w in zip(outs, want):
        assert rel(out, w) < 2e-5
    # refusals: too many items, a channel count that is not a multiple of 8
    big = (L.ColsumItem * 17)()
    with pytest.raises(L.B200dmError):
        L.call("b200dm_colsum_batched", dtype, big, 17)
    arr[0].C = 12
    with pytest.raises(L.B200dmError):
        L.call("b200dm_colsum_batched", dtype, arr, 1)


def test_final_conv_large_and_ragged_pixel_count():
    """B*HW that is not a multiple of the pixels one CTA covers, bf16, against the fp32 definition."""
    B, S, C = 5, 24, 3
    x = q(rnd(B, 64, S, S, seed=65), L.BF16).requires_grad_(True)
    w = rnd(C, 64, 1, 1, seed=66, scale=0.125).requires_grad_(True)
    bias = rnd(C, seed=67).requires_grad_(True)
    ref = F.conv2d(x, w, bias)
    xv = nhwc(x.detach(), L.BF16)
    y = torch.empty(B, C, S, S, device=DEV)
    L.call("b200dm_final_conv_fwd", L.BF16, xv.ptr, xv.ld, w.data_ptr(), bias.data_ptr(), y.data_ptr(), B, S * S, 64, C)
    assert rel(y, ref) < 2e-5
    dy = rnd(B, C, S, S, seed=68)
    ref.backward(dy)
    dxv = View.zeros(B, S, S, 64, DT[L.BF16], DEV)
    dw, db = torch.zeros(C, 64, device=DEV), torch.zeros(C, device=DEV)
    L.call("b200dm_final_conv_bwd", L.BF16, xv.ptr, xv.ld, w.data_ptr(), dy.data_ptr(), dxv.ptr, dxv.ld,
           dw.data_ptr(), db.data_ptr(), B, S * S, 64, C)
    assert rel(dxv.to_nchw(), x.grad) < tol(L.BF16)
    assert rel(dw, w.grad.view(C, 64)) < 2e-5 and rel(db, bias.grad) < 2e-5


# ------------------------------------------------------------------------------------------------
# fused inference LinearAttention block (csrc/linattn_tc.cu) against the oracle's linear_attention + skip
# ------------------------------------------------------------------------------------------------
def _linattn_block_case(B, S, Cc, seed, big_k=False, ld_extra=0):
    from oracle import ddpm_oracle as O
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cc, S, S, generator=g).to(DEV)
    x = x.to(torch.bfloat16).float()                                   # the activation tensor is bf16
    sd = {
        "a.norm.g": (1 + 0.2 * torch.randn(1, Cc, 1, 1, generator=g)).to(DEV),
        "a.to_qkv.weight": (torch.randn(384, Cc, 1, 1, generator=g) * (3.0 if big_k else 1.0) / Cc ** 0.5).to(DEV),
        "a.mem_kv": torch.randn(2, 4, 32, 4, generator=g).to(DEV),
        "a.to_out.0.weight": (torch.randn(Cc, 128, 1, 1, generator=g) / 128 ** 0.5).to(DEV),
        "a.to_out.0.bias": (0.1 * torch.randn(Cc, generator=g)).to(DEV),
        "a.to_out.1.g": (1 + 0.2 * torch.randn(1, Cc, 1, 1, generator=g)).to(DEV),
    }
    ref = O.linear_attention(sd, "a", x, O.Emu(None)) + x
    n = S * S
    xv = nhwc(x, L.BF16, ld=Cc + ld_extra, off=ld_extra) if ld_extra else nhwc(x, L.BF16)
    yv = View.zeros(B, S, S, Cc, DT[L.BF16], DEV)
    wq = torch.empty(384 * Cc, dtype=torch.bfloat16, device=DEV)
    L.call("b200dm_pack_linattn_qkv", sd["a.to_qkv.weight"].data_ptr(), sd["a.norm.g"].data_ptr(), wq.data_ptr(), Cc)
    wo = sd["a.to_out.0.weight"].view(Cc, 128).to(torch.bfloat16).contiguous()
    ws = torch.zeros(L.load().b200dm_linattn_block_ws_floats(B, n, Cc), device=DEV)
    d = L.LinAttnBlockDesc(B=B, n=n, C=Cc, x_ld=xv.ld, y_ld=yv.ld, x=xv.ptr, y=yv.ptr, wqkv=wq.data_ptr(),
                           wout=wo.data_ptr(), bout=sd["a.to_out.0.bias"].data_ptr(),
                           gout=sd["a.to_out.1.g"].data_ptr(), mem_kv=sd["a.mem_kv"].data_ptr(), ws=ws.data_ptr())
    assert L.load().b200dm_linattn_block_supported(ctypes.byref(d)) == 1
    L.call("b200dm_linattn_block_fwd", ctypes.byref(d))
    torch.cuda.synchronize()
    return yv.to_nchw().float(), ref, ws, d, (xv, wq, wo, sd)


@pytest.mark.parametrize("B,S,Cc", [(2, 16, 64), (3, 32, 64), (2, 16, 128), (5, 32, 128), (150, 16, 64)])
def test_linattn_block_fused_inference(B, S, Cc):
    y, ref, _, _, _ = _linattn_block_case(B, S, Cc, seed=500 + B + S + Cc)
    e = rel(y, ref)
    report(test="linattn_block", B=B, S=S, C=Cc, rel=e)
    assert e < 1.5e-2, e            # bf16 operands (x, q, k-softmax, v, context, out) against the fp32 definition


def test_linattn_block_large_logits_and_strided_input():
    """Keys whose exp would overflow without the exact row maximum (|k| up to ~40), input as a channel slice."""
    y, ref, _, _, _ = _linattn_block_case(2, 32, 64, seed=77, big_k=True, ld_extra=64)
    assert torch.isfinite(y).all()
    assert rel(y, ref) < 3e-2


def test_linattn_block_refuses_unsupported_shapes():
    d = L.LinAttnBlockDesc(B=1, n=64, C=64, x_ld=64, y_ld=64)
    assert L.load().b200dm_linattn_block_supported(ctypes.byref(d)) == 0          # n not a multiple of 128
    d = L.LinAttnBlockDesc(B=1, n=256, C=256, x_ld=256, y_ld=256)
    assert L.load().b200dm_linattn_block_supported(ctypes.byref(d)) == 0          # C = 256 does not fit
    with pytest.raises(L.B200dmError):
        L.call("b200dm_linattn_block_fwd", ctypes.byref(d))


# ------------------------------------------------------------------------------------------------
# norms
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
@pytest.mark.parametrize("B,S,Cc,film,res", [(3, 16, 64, True, False), (2, 8, 256, False, True),
                                             (2, 32, 128, True, True), (5, 4, 512, False, False)])
def test_groupnorm_film_silu_fwd_bwd(dtype, B, S, Cc, film, res):
    G = 8
    x = (q(rnd(B, Cc, S, S, seed=81) + 0.5, dtype)).requires_grad_(True)
    gamma = (1 + 0.1 * rnd(Cc, seed=82)).requires_grad_(True)
    beta = (0.1 * rnd(Cc, seed=83)).requires_grad_(True)
    fm = rnd(B, 2 * Cc + 6, seed=84, scale=0.3).requires_grad_(True) if film else None
    r = q(rnd(B, Cc, S, S, seed=85), dtype) if res else None
    h = F.group_norm(x, G, gamma, beta, eps=1e-5)
    if film:
        sc, sh = fm[:, 3:3 + Cc], fm[:, 3 + Cc:3 + 2 * Cc]
        h = h * (sc[:, :, None, None] + 1) + sh[:, :, None, None]
    ref = F.silu(h) + (r if res else 0)
    xv = nhwc(x.detach(), dtype, ld=Cc + 8, off=8)
    stats = torch.empty(B, G, 2, device=DEV)
    L.call("b200dm_gn_stats", dtype, xv.ptr, xv.ld, stats.data_ptr(), B, S * S, Cc, G, 1e-5)
    xs = x.detach().reshape(B, G, -1)
    assert rel(stats[..., 0], xs.mean(-1)) < 1e-5
    assert rel(stats[..., 1], (xs.var(-1, unbiased=False) + 1e-5).rsqrt()) < 1e-5
    yv = View.zeros(B, S, S, Cc, DT[dtype], DEV, ld=2 * Cc, off=Cc)
    rv = nhwc(r, dtype) if res else None
    fptr = fm.detach().data_ptr() + 3 * 4 if film else None
    L.call("b200dm_gn_apply_fwd", dtype, xv.ptr, xv.ld, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
           fptr, 2 * Cc + 6, None if rv is None else rv.ptr, 0 if rv is None else rv.ld, yv.ptr, yv.ld,
           B, S * S, Cc, G)
    assert rel(yv.to_nchw(), ref) < tol(dtype)
    # one-launch cluster kernel (statistics + apply): same outputs, same statistics
    stats2 = torch.zeros(B, G, 2, device=DEV)
    yv2 = View.zeros(B, S, S, Cc, DT[dtype], DEV, ld=2 * Cc, off=Cc)
    L.call("b200dm_gn_fwd", dtype, xv.ptr, xv.ld, stats2.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
           fptr, 2 * Cc + 6, None if rv is None else rv.ptr, 0 if rv is None else rv.ld, yv2.ptr, yv2.ld,
           B, S * S, Cc, G, 1e-5)
    assert rel(stats2, stats) < 2e-5
    assert rel(yv2.to_nchw(), ref) < tol(dtype)
    # backward
    dy = q(rnd(B, Cc, S, S, seed=86), dtype)
    ref.backward(dy)
    dyv = nhwc(dy, dtype)
    dxv = View.zeros(B, S, S, Cc, DT[dtype], DEV)
    dgamma, dbeta = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    dfilm = torch.zeros(B, 2 * Cc + 6, device=DEV)
    sums = torch.empty(L.load().b200dm_gn_bwd_ws_floats(B, S * S, Cc), device=DEV)
    gmeans = torch.empty(B, G, 2, device=DEV)
    dbias = torch.zeros(Cc, device=DEV)
    L.call("b200dm_gn_apply_bwd", dtype, dyv.ptr, dyv.ld, xv.ptr, xv.ld, stats.data_ptr(), gamma.data_ptr(),
           beta.data_ptr(), fptr, 2 * Cc + 6, dxv.ptr, dxv.ld, dgamma.data_ptr(), dbeta.data_ptr(),
           dfilm.data_ptr() + 3 * 4 if film else None, dbias.data_ptr(), sums.data_ptr(), gmeans.data_ptr(),
           B, S * S, Cc, G)
    # closed-form bias gradient of the producing conv == pixel/batch sum of dx
    ref_db = x.grad.sum((0, 2, 3))
    assert (dbias - ref_db).abs().max().item() <= 2e-4 * max(1.0, x.grad.abs().sum((0, 2, 3)).max().item())
    assert rel(dxv.to_nchw(), x.grad) < (1e-4 if dtype == L.F32 else 6e-3)
    assert rel(dgamma, gamma.grad) < 1e-4 and rel(dbeta, beta.grad) < 1e-4
    if film:
        assert rel(dfilm, fm.grad) < 1e-4


@pytest.mark.parametrize("B,S,Cin,Cout,film", [(3, 32, 64, 64, True), (2, 16, 128, 128, False), (5, 8, 128, 256, True),
                                               (9, 4, 256, 512, False), (1, 64, 64, 64, False)])
def test_conv_epilogue_groupnorm_statistics_and_one_pass_norm(B, S, Cin, Cout, film):
    """3x3 conv (tcgen05) emitting per-slot GroupNorm partial sums + the one-pass norm that consumes them,
    against conv2d -> group_norm -> FiLM -> SiLU of the reference Block (ddpm.py:164-173)."""
    if not L.load().b200dm_tc_available():
        pytest.skip("needs the tcgen05 path")
    G = 8
    x = q(rnd(B, Cin, S, S, seed=91), L.BF16)
    w = q(rnd(Cout, Cin, 3, 3, seed=92, scale=1 / math.sqrt(Cin * 9)), L.BF16)
    bias = rnd(Cout, seed=93, scale=0.3)
    gamma, beta = 1 + 0.1 * rnd(Cout, seed=94), 0.1 * rnd(Cout, seed=95)
    fm = rnd(B, 2 * Cout, seed=96, scale=0.3) if film else None
    conv = F.conv2d(x, w, bias, padding=1)
    h = F.group_norm(conv, G, gamma, beta, eps=1e-5)
    if film:
        h = h * (fm[:, :Cout, None, None] + 1) + fm[:, Cout:, None, None]
    ref = F.silu(h)
    xv = nhwc(x, L.BF16)
    cv = View.zeros(B, S, S, Cout, torch.bfloat16, DEV)
    yv = View.zeros(B, S, S, Cout, torch.bfloat16, DEV)
    slots = S * S // min(32, S * S)
    part = torch.full((B, slots, Cout // 8, 2), float("nan"), device=DEV)   # per pixel slot and 8-channel chunk
    wp = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous().to(torch.bfloat16)
    d = L.ConvDesc(dtype=L.BF16, mode=0, ksize=3, impl=1, B=B, H=S, W=S, Cin=Cin, Cout=Cout, x=xv.ptr, x_ld=xv.ld,
                   w=wp.data_ptr(), bias=bias.data_ptr(), y=cv.ptr, y_ld=cv.ld, res=None, res_ld=0, accumulate=0,
                   gn_part=part.data_ptr(), gn_groups=G)
    L.call("b200dm_conv_fwd", d)
    assert torch.isfinite(part).all()
    sums = part.sum(1).reshape(B, G, Cout // 8 // G, 2).sum(2)               # [B, G, 2]
    cg = conv.reshape(B, G, -1)
    assert rel(sums[..., 0], cg.sum(-1)) < 2e-3 and rel(sums[..., 1], (cg * cg).sum(-1)) < 2e-3
    stats = torch.empty(B, G, 2, device=DEV)
    L.call("b200dm_gn_fwd_pre", L.BF16, cv.ptr, cv.ld, part.data_ptr(), slots, stats.data_ptr(), gamma.data_ptr(),
           beta.data_ptr(), None if fm is None else fm.data_ptr(), 2 * Cout, None, 0, yv.ptr, yv.ld, B, S * S, Cout,
           G, 1e-5)
    assert rel(stats[..., 0], cg.mean(-1)) < 2e-3
    assert rel(stats[..., 1], (cg.var(-1, unbiased=False) + 1e-5).rsqrt()) < 2e-3
    assert rel(yv.to_nchw(), ref) < 8e-3                 # conv output rounded to bf16 before the norm
    part2 = torch.full_like(part, float("nan"))          # deterministic: bit-identical on a second run
    d.gn_part = part2.data_ptr()
    L.call("b200dm_conv_fwd", d)
    assert torch.equal(part, part2)


@pytest.mark.parametrize("B,S,Cin,Cout,film,res,raw", [
    (3, 32, 64, 64, True, False, True),        # resident weights, cluster of 4 (small batch)
    (80, 32, 64, 64, False, True, False),      # one CTA per sample, all 8 accumulators (512 TMEM columns)
    (2, 16, 128, 128, True, False, True),      # N = 128 tiles
    (2, 16, 192, 128, False, True, True),      # Cin = 192 (concat input), streamed weights
    (2, 64, 64, 64, False, False, True),       # 64x64: cluster of 8, four tiles per CTA
    (2, 32, 256, 128, True, True, False),      # cluster, N = 128, residual
    (2, 16, 256, 256, True, True, True),       # two N tiles per pixel tile
    (100, 16, 64, 64, True, True, True),       # two accumulators per CTA
    (40, 64, 128, 64, False, True, True),      # cluster of 4 x 8 tiles at 64x64, Cin = 128
    # small levels: tile-local statistics (conv_tc_gn_kernel)
    (5, 8, 128, 256, True, True, True),        # 8x8: a sample spans two epilogue warps; 2.5 tiles of samples
    (3, 8, 128, 128, False, False, True),      # groups of 16 channels: four per 64-column unit
    (9, 4, 256, 512, True, False, True),       # 4x4: two samples per warp, groups of 64; last tile partly empty
    (16, 4, 512, 512, False, True, False),     # two full tiles of 8 samples
    (130, 8, 256, 256, True, True, True),      # persistent CTAs walk several tiles (65 x 1 tiles at N = 256)
])
def test_conv_groupnorm_film_silu_one_launch(B, S, Cin, Cout, film, res, raw):
    """b200dm_conv_gn_fwd: Block.forward (ddpm.py:164-173) + the residual of ResnetBlock (:200) in ONE launch, with the
    conv accumulators resident in TMEM; against conv2d -> group_norm -> FiLM -> SiLU (+ res) in fp32."""
    if not L.load().b200dm_tc_available():
        pytest.skip("needs the tcgen05 path")
    G = 8
    x = q(rnd(B, Cin, S, S, seed=191), L.BF16)
    w = q(rnd(Cout, Cin, 3, 3, seed=192, scale=1 / math.sqrt(Cin * 9)), L.BF16)
    bias = rnd(Cout, seed=193, scale=0.3)
    gamma, beta = 1 + 0.1 * rnd(Cout, seed=194), 0.1 * rnd(Cout, seed=195)
    fm = rnd(B, 2 * Cout + 8, seed=196, scale=0.3) if film else None
    r = q(rnd(B, Cout, S, S, seed=197), L.BF16) if res else None
    conv = F.conv2d(x, w, bias, padding=1)
    h = F.group_norm(conv, G, gamma, beta, eps=1e-5)
    if film:
        h = h * (fm[:, :Cout, None, None] + 1) + fm[:, Cout:2 * Cout, None, None]
    ref = F.silu(h) + (r if res else 0)
    xv = nhwc(x, L.BF16, ld=Cin + 64, off=64)            # channel slice of a wider (concat) buffer
    yv = View.zeros(B, S, S, Cout, torch.bfloat16, DEV, ld=2 * Cout, off=Cout)
    cv = View.zeros(B, S, S, Cout, torch.bfloat16, DEV) if raw else None
    rv = nhwc(r, L.BF16) if res else None
    stats = torch.full((B, G, 2), float("nan"), device=DEV)
    wp = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous().to(torch.bfloat16)
    d = L.ConvDesc(dtype=L.BF16, mode=0, ksize=3, impl=1, B=B, H=S, W=S, Cin=Cin, Cout=Cout, x=xv.ptr, x_ld=xv.ld,
                   w=wp.data_ptr(), bias=bias.data_ptr(), y=yv.ptr, y_ld=yv.ld, res=None if rv is None else rv.ptr,
                   res_ld=0 if rv is None else rv.ld, accumulate=0, gn_part=None, gn_groups=0)
    gnd = L.GnDesc(gamma=gamma.data_ptr(), beta=beta.data_ptr(), film=None if fm is None else fm.data_ptr(),
                   film_ld=0 if fm is None else fm.shape[1], groups=G, eps=1e-5, raw_ld=0 if cv is None else cv.ld,
                   stats=stats.data_ptr(), raw=None if cv is None else cv.ptr)
    assert L.load().b200dm_conv_gn_supported(ctypes.byref(d), ctypes.byref(gnd)) == 1
    L.call("b200dm_conv_gn_fwd", ctypes.byref(d), ctypes.byref(gnd))
    cg = conv.reshape(B, G, -1)
    assert rel(stats[..., 0], cg.mean(-1)) < 1e-3
    assert rel(stats[..., 1], (cg.var(-1, unbiased=False) + 1e-5).rsqrt()) < 1e-3
    out = yv.to_nchw()
    assert rel(out, ref) < 4e-3, rel(out, ref)            # one bf16 rounding of the output + tanh.approx
    assert (out - ref).abs().max().item() < 0.06
    assert torch.equal(yv.buf[..., :Cout], torch.zeros_like(yv.buf[..., :Cout]))     # the other slice is untouched
    if raw:
        assert rel(cv.to_nchw(), conv) < 4e-3
    # deterministic (no atomics anywhere on the statistics path): bit-identical on a second run
    first, st1 = yv.buf.clone(), stats.clone()
    L.call("b200dm_conv_gn_fwd", ctypes.byref(d), ctypes.byref(gnd))
    assert torch.equal(first, yv.buf) and torch.equal(st1, stats)
    # unsupported layers are reported, not mis-executed
    d2 = L.ConvDesc(dtype=L.BF16, mode=0, ksize=3, impl=1, B=B, H=2, W=2, Cin=Cin, Cout=Cout, x=xv.ptr, x_ld=xv.ld,
                    w=wp.data_ptr(), bias=bias.data_ptr(), y=yv.ptr, y_ld=yv.ld, res=None, res_ld=0, accumulate=0,
                    gn_part=None, gn_groups=0)
    assert L.load().b200dm_conv_gn_supported(ctypes.byref(d2), ctypes.byref(gnd)) == 0


@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
@pytest.mark.parametrize("Cc,res", [(64, False), (128, True), (256, True), (512, False)])
def test_rmsnorm_fwd_bwd(dtype, Cc, res):
    B, S = 3, 8
    x = q(rnd(B, Cc, S, S, seed=91), dtype).requires_grad_(True)
    g = (1 + 0.1 * rnd(1, Cc, 1, 1, seed=92)).requires_grad_(True)
    r = q(rnd(B, Cc, S, S, seed=93), dtype) if res else None
    ref = F.normalize(x, dim=1) * g * (Cc ** 0.5) + (r if res else 0)
    xv = nhwc(x.detach(), dtype, ld=Cc + 8, off=0)
    yv = View.zeros(B, S, S, Cc, DT[dtype], DEV)
    rv = nhwc(r, dtype) if res else None
    L.call("b200dm_rmsnorm_fwd", dtype, xv.ptr, xv.ld, g.data_ptr(), None if rv is None else rv.ptr,
           0 if rv is None else rv.ld, yv.ptr, yv.ld, B * S * S, Cc)
    assert rel(yv.to_nchw(), ref) < tol(dtype)
    dy = q(rnd(B, Cc, S, S, seed=94), dtype)
    ref.backward(dy)
    dyv = nhwc(dy, dtype)
    dxv = View.zeros(B, S, S, Cc, DT[dtype], DEV)
    dg = torch.zeros(Cc, device=DEV)
    L.call("b200dm_rmsnorm_bwd", dtype, dyv.ptr, dyv.ld, xv.ptr, xv.ld, g.data_ptr(),
           None if rv is None else rv.ptr, 0 if rv is None else rv.ld, dxv.ptr, dxv.ld, dg.data_ptr(),
           B * S * S, Cc)
    expect_dx = x.grad + (r if res else 0)     # `res` of the backward kernel is an additive gradient
    assert rel(dxv.to_nchw(), expect_dx) < tol(dtype)
    assert rel(dg, g.grad.flatten()) < 1e-4


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def _lin_attn_ref(qkv, mem):  # qkv [B, 384, n] fp32; mem [2,4,32,4]   (ddpm.py:222-238)
    B, _, n = qkv.shape
    qh, kh, vh = (t.reshape(B, 4, 32, n) for t in qkv.chunk(3, dim=1))
    k = torch.cat((mem[0][None].expand(B, -1, -1, -1), kh), dim=-1)
    v = torch.cat((mem[1][None].expand(B, -1, -1, -1), vh), dim=-1)
    qs = qh.softmax(dim=-2) * (32 ** -0.5)
    ks = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", ks, v)
    return torch.einsum("bhde,bhdn->bhen", ctx, qs).reshape(B, 128, n)


@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
@pytest.mark.parametrize("B,S", [(2, 8), (3, 16), (1, 64)])
def test_linear_attention_fwd_bwd(dtype, B, S):
    n = S * S
    qkv = q(rnd(B, 384, S, S, seed=101), dtype).requires_grad_(True)
    mem = rnd(2, 4, 32, 4, seed=102).requires_grad_(True)
    ref = _lin_attn_ref(qkv.flatten(2), mem)
    qv = nhwc(qkv.detach(), dtype)
    ctx, kstat = torch.empty(B, 4, 32, 32, device=DEV), torch.empty(B, 4, 32, 2, device=DEV)
    ov = View.zeros(B, S, S, 128, DT[dtype], DEV)
    L.call("b200dm_linattn_fwd", dtype, qv.ptr, qv.ld, mem.data_ptr(), ctx.data_ptr(), kstat.data_ptr(),
           ov.ptr, ov.ld, B, n)
    assert rel(ov.to_nchw().flatten(2), ref) < tol(dtype)
    do = q(rnd(B, 128, S, S, seed=103), dtype)
    ref.backward(do.flatten(2))
    dov = nhwc(do, dtype)
    dqv = View.zeros(B, S, S, 384, DT[dtype], DEV)
    dctx, dmem = torch.empty(B, 4, 33, 32, device=DEV), torch.zeros(2, 4, 32, 4, device=DEV)
    L.call("b200dm_linattn_bwd", dtype, dov.ptr, dov.ld, qv.ptr, qv.ld, mem.data_ptr(), ctx.data_ptr(),
           kstat.data_ptr(), dctx.data_ptr(), dqv.ptr, dqv.ld, dmem.data_ptr(), B, n)
    assert rel(dqv.to_nchw(), qkv.grad) < (1e-4 if dtype == L.F32 else 6e-3)
    # bf16 mode runs the 32x32 contractions on tensor cores with bf16 operands (as the reference's autocast
    # einsum does), so the memory-kv gradient carries operand rounding; fp32 mode stays at 1e-4
    assert rel(dmem, mem.grad) < (1e-4 if dtype == L.F32 else 5e-3)


@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
@pytest.mark.parametrize("B,S", [(3, 4), (2, 8), (2, 5), (1, 7)])
def test_full_attention_fwd_bwd(dtype, B, S):
    n = S * S
    qkv = q(rnd(B, 384, S, S, seed=111), dtype).requires_grad_(True)
    mem = rnd(2, 4, 4, 32, seed=112).requires_grad_(True)
    qh, kh, vh = (t.reshape(B, 4, 32, n).transpose(-1, -2) for t in qkv.flatten(2).chunk(3, dim=1))
    k = torch.cat((mem[0][None].expand(B, -1, -1, -1), kh), dim=-2)
    v = torch.cat((mem[1][None].expand(B, -1, -1, -1), vh), dim=-2)
    attn = (torch.einsum("bhid,bhjd->bhij", qh, k) * 32 ** -0.5).softmax(-1)
    ref = torch.einsum("bhij,bhjd->bhid", attn, v).transpose(-1, -2).reshape(B, 128, S, S)
    qv = nhwc(qkv.detach(), dtype)
    ov = View.zeros(B, S, S, 128, DT[dtype], DEV)
    L.call("b200dm_attn_fwd", dtype, qv.ptr, qv.ld, mem.data_ptr(), ov.ptr, ov.ld, B, n)
    assert rel(ov.to_nchw(), ref) < tol(dtype)
    do = q(rnd(B, 128, S, S, seed=113), dtype)
    ref.backward(do)
    dov = nhwc(do, dtype)
    dqv = View.zeros(B, S, S, 384, DT[dtype], DEV)
    dmem = torch.zeros(2, 4, 4, 32, device=DEV)
    L.call("b200dm_attn_bwd", dtype, dov.ptr, dov.ld, qv.ptr, qv.ld, mem.data_ptr(), dqv.ptr, dqv.ld,
           dmem.data_ptr(), B, n)
    assert rel(dqv.to_nchw(), qkv.grad) < (1e-4 if dtype == L.F32 else 6e-3)
    assert rel(dmem, mem.grad) < (1e-4 if dtype == L.F32 else 5e-3)     # bf16: P and dS are bf16 MMA operands
    with pytest.raises(L.B200dmError):       # 16x16 softmax attention never occurs
        L.call("b200dm_attn_fwd", dtype, qv.ptr, qv.ld, mem.data_ptr(), ov.ptr, ov.ld, B, 256)


# ------------------------------------------------------------------------------------------------
# time embedding, linears, optimiser
# ------------------------------------------------------------------------------------------------
def test_sinusoidal_and_linears():
    B = 37
    t = torch.randint(0, 1000, (B,), device=DEV)
    emb = torch.empty(B, 64, device=DEV)
    L.call("b200dm_sinusoidal", t.data_ptr(), emb.data_ptr(), B, 64, 10000.0)
    half = 32
    fr = torch.exp(torch.arange(half, device=DEV) * -(math.log(10000) / (half - 1)))
    a = t[:, None] * fr[None, :]
    assert (emb - torch.cat((a.sin(), a.cos()), -1)).abs().max().item() < 2e-4  # argument up to ~1e3 rad
    for act, fn in ((0, lambda v: v), (1, F.gelu), (2, F.silu)):
        M, N, K = B, 200, 64
        X = rnd(M, K, seed=121).requires_grad_(True)
        W = rnd(N, K, seed=122, scale=0.1).requires_grad_(True)
        b = rnd(N, seed=123).requires_grad_(True)
        ref = fn(F.linear(X, W, b))
        Y, pre = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
        L.call("b200dm_linear_fwd", X.data_ptr(), W.data_ptr(), b.data_ptr(), Y.data_ptr(), pre.data_ptr(), M, N, K, act)
        assert rel(Y, ref) < 1e-5
        dY = rnd(M, N, seed=124)
        ref.backward(dY)
        dX = torch.empty(M, K, device=DEV)
        dW, db = torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
        g = dY.clone()
        L.call("b200dm_linear_bwd", X.data_ptr(), W.data_ptr(), pre.data_ptr(), g.data_ptr(), dX.data_ptr(),
               dW.data_ptr(), db.data_ptr(), M, N, K, act)
        assert rel(dX, X.grad) < 1e-5 and rel(dW, W.grad) < 1e-5 and rel(db, b.grad) < 1e-5
    # the wide FiLM projection shape: M = batch, N = 8064, K = 256 (split-K dX path)
    M, N, K = 16, 8064, 256
    X, W, dY = rnd(M, K, seed=125), rnd(N, K, seed=126, scale=0.05), rnd(M, N, seed=127)
    dX = torch.empty(M, K, device=DEV)
    L.call("b200dm_linear_bwd", X.data_ptr(), W.data_ptr(), None, dY.data_ptr(), dX.data_ptr(), None, None, M, N, K, 0)
    assert rel(dX, dY @ W) < 1e-5


@pytest.mark.parametrize("dtype", [L.F32, L.BF16], ids=["f32", "bf16"])
def test_pack_conv_weight(dtype):
    w = rnd(96, 160, 3, 3, seed=131)
    master = w.permute(2, 3, 0, 1).reshape(9, 96, 160).contiguous()
    wf = torch.empty(9, 96, 160, dtype=DT[dtype], device=DEV)
    wt = torch.empty(9, 160, 96, dtype=DT[dtype], device=DEV)
    L.call("b200dm_pack_conv_weight", dtype, master.data_ptr(), wf.data_ptr(), wt.data_ptr(), 9, 96, 160, 1, 0, 0, 0)
    assert torch.equal(wf, pack_w(w, dtype)) and torch.equal(wt, pack_w(w, dtype, transpose=True, flip=True))


def test_adam_matches_torch_and_ema():
    n = 100003
    p0, g = rnd(n, seed=141), rnd(n, seed=142)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=2e-5, betas=(0.9, 0.99))
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        ref.grad = g * step
        opt.step()
        gs = (g * step * 2).contiguous()
        L.call("b200dm_adam_step", p.data_ptr(), gs.data_ptr(), m.data_ptr(), v.data_ptr(), n, 2e-5, 0.9, 0.99,
               1e-8, 0.0, step, 0.5)
    assert (p - ref.detach()).abs().max().item() < 1e-7
    ema = p0.clone()
    L.call("b200dm_ema_update", ema.data_ptr(), p.data_ptr(), n, 0.995)
    assert (ema - torch.lerp(p0, p, 0.005)).abs().max().item() < 1e-7


def test_batched_weight_pack_matches_reference_layouts():
    """WeightPack (one launch for all 73 GEMM convs) against per-tensor torch permutations, including the
    pixel-unshuffle conv that keeps the reference [Cout, 4C] master layout."""
    from b200dm.engine import WeightPack
    from b200dm.params import ParamArena
    arena = ParamArena(64, 3, DEV)
    g = torch.Generator().manual_seed(5)
    arena.flat.copy_(torch.randn(arena.flat.numel(), generator=g))
    pk = WeightPack(arena, L.BF16, with_dgrad=True)
    pk.refresh(force=True)
    for nm, ci in arena.convs.items():
        w = arena.views[nm + ".weight"]                      # logical reference shape
        if ci.mode == 0:
            assert torch.equal(pk.fwd[nm].view(ci.taps, ci.cout, ci.cin), pack_w(w, L.BF16)), nm
            assert torch.equal(pk.tr[nm].view(ci.taps, ci.cin, ci.cout), pack_w(w, L.BF16, True, True)), nm
        else:
            wp = w.reshape(ci.cout, ci.cin, 4).permute(2, 0, 1).contiguous().to(torch.bfloat16)
            assert torch.equal(pk.fwd[nm].view(4, ci.cout, ci.cin), wp), nm
            assert torch.equal(pk.tr[nm].view(4, ci.cin, ci.cout), wp.transpose(1, 2).contiguous()), nm
