"""Kernels of the training step (csrc/train_ops.cu), the conv launch plans and b200ir_pack_weights on the B200, each against
the torch restatement of the same entry point in tests/cabi_sim.py: the same ops.* wrapper is called once on CUDA tensors
(libb200ir.so) and once on CPU copies with the simulator installed.  This also pins the simulator — the host-logic tests of
the `-m "not gpu"` suite rely on it — to the real kernels."""
import math

import pytest
import torch

from tests import cabi_sim

pytestmark = pytest.mark.gpu


def both(fn, tensors):
    """fn(dict of tensors) -> None (writes outputs in place).  Returns (gpu dict, cpu dict) after running it on both sides."""
    gpu = {k: (v.cuda() if v is not None else None) for k, v in tensors.items()}
    fn(gpu)
    torch.cuda.synchronize()
    cpu = {k: (v.clone() if v is not None else None) for k, v in tensors.items()}
    with cabi_sim.installed():
        fn(cpu)
    return gpu, cpu


def close(name, got, exp, rel=2e-3, abs_=None):
    got, exp = got.float().cpu(), exp.float()
    scale = exp.abs().max().item()
    tol = abs_ if abs_ is not None else rel * max(scale, 1e-6)
    err = (got - exp).abs().max().item()
    assert err <= tol, (name, err, tol, scale)


def rn(*s, g=None, std=1.0):
    return torch.randn(*s, generator=g) * std


@pytest.mark.parametrize('B,h,w,C,c_sft,mod', [(3, 16, 48, 64, 32, True), (2, 8, 24, 512, 256, True), (2, 5, 7, 32, 32, False),
                                              (2, 6, 10, 24, 0, True)])
def test_sft_mod_forward_backward(B, h, w, C, c_sft, mod):
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(B * C + h)
    t = dict(a=rn(B, h, w, C, g=g).half(), g=rn(B, h, w, C, g=g).half(),
             sc=(1 + 0.3 * rn(B, h, w, c_sft, g=g)).half() if c_sft else None, sh=rn(B, h, w, c_sft, g=g).half() if c_sft else None,
             s=(1 + 0.5 * rn(B, C, g=g)) if mod else None, out=torch.zeros(B, h, w, C).half(),
             da=rn(B, h, w, C, g=g).half(), dsc=torch.zeros(B, h, w, max(c_sft, 8)).half() if c_sft else None,
             dsh=torch.zeros(B, h, w, max(c_sft, 8)).half() if c_sft else None, ds=torch.zeros(B, C))

    def fwd(d):
        ops.sft_mod(d['a'], d['sc'], d['sh'], d['s'], d['out'])
    gpu, cpu = both(fwd, t)
    close('sft_mod', gpu['out'], cpu['out'])

    for acc in (False, True):
        def bwd(d):
            ops.sft_mod_bwd(d['g'], d['a'], d['sc'], d['sh'], d['s'], d['da'], acc, d['dsc'], d['dsh'], d['ds'])
        gpu, cpu = both(bwd, t)
        close('da', gpu['da'], cpu['da'])
        close('ds', gpu['ds'], cpu['ds'], rel=3e-3)
        if c_sft:
            close('dscale', gpu['dsc'], cpu['dsc'])
            close('dshift', gpu['dsh'], cpu['dsh'])


def test_sft_mod_bwd_broadcast_constant_input():
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, h, w, C = 4, 4, 12, 512
    t = dict(g=rn(B, h, w, C, g=g).half(), a=rn(h, w, C, g=g).half(), ds=torch.zeros(B, C))

    def bwd(d):
        ops.sft_mod_bwd(d['g'], d['a'], None, None, None, None, False, None, None, d['ds'], a_broadcast=True)
    gpu, cpu = both(bwd, t)
    close('ds', gpu['ds'], cpu['ds'], rel=3e-3)


@pytest.mark.parametrize('B,h,w,C,noise,osc', [(3, 16, 48, 64, True, True), (2, 8, 24, 512, True, True), (2, 5, 7, 32, False, False)])
def test_style_act_bwd(B, h, w, C, noise, osc):
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(C)
    a = rn(B, h, w, C, g=g)
    a = torch.where(a.abs() < 0.02, torch.full_like(a, 0.05), a).half()        # keep clear of the branch point
    t = dict(da=rn(B, h, w, C, g=g).half(), a=a, nz=rn(B, 1, h, w, g=g) if noise else None, gain=torch.tensor([0.3]),
             bias=0.1 * rn(C, g=g), osc=(1 + 0.3 * rn(B, C, g=g)) if osc else None, out=torch.zeros(B, h, w, C).half(),
             dd=torch.zeros(B, C))

    def run(d):
        ops.style_act_bwd(d['da'], d['a'], d['nz'], d['gain'], d['bias'], d['osc'], 4.0, d['out'], d['dd'])
    gpu, cpu = both(run, t)
    close('out', gpu['out'], cpu['out'])
    close('dd', gpu['dd'], cpu['dd'], rel=3e-3)


@pytest.mark.parametrize('B,h,w,C', [(3, 16, 48, 64), (2, 4, 12, 512), (2, 5, 7, 32)])
def test_to_rgb_bwd_and_skip_adjoint(B, h, w, C):
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(C + 1)
    t = dict(drgb=rn(B, 3, h, w, g=g), a=rn(B, h, w, C, g=g).half(), w=rn(3, C, g=g) / math.sqrt(C), s=1 + 0.5 * rn(B, C, g=g),
             da=rn(B, h, w, C, g=g).half(), ds=torch.zeros(B, C))
    for acc in (False, True):
        def run(d):
            ops.to_rgb_bwd(d['drgb'], d['a'], d['w'], d['s'], d['da'], acc, d['ds'])
        gpu, cpu = both(run, t)
        close('da', gpu['da'], cpu['da'])
        close('ds', gpu['ds'], cpu['ds'], rel=3e-3)
    t2 = dict(d=rn(B, 3, 2 * h, 2 * w, g=g), out=torch.zeros(B, 3, h, w))
    gpu, cpu = both(lambda d: ops.rgb_up_adjoint(d['d'], d['out']), t2)
    close('rgb_up_adjoint', gpu['out'], cpu['out'], rel=1e-5)


@pytest.mark.parametrize('B,h,w,C', [(3, 16, 48, 64), (2, 8, 24, 512), (2, 5, 7, 32)])
def test_decoder_parameter_reductions(B, h, w, C):
    """fix_decoder=False: the *_params variants of the pointwise adjoints (activate.bias, noise gain, ToRGB weight partials) and
    plane_sums, against the simulator; their primary outputs must equal the plain variants'."""
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(C + 2)
    a = rn(B, h, w, C, g=g)
    a = torch.where(a.abs() < 0.02, torch.full_like(a, 0.05), a).half()
    t = dict(da=rn(B, h, w, C, g=g).half(), a=a, nz=rn(B, 1, h, w, g=g), gain=torch.tensor([0.3]), bias=0.1 * rn(C, g=g),
             osc=1 + 0.3 * rn(B, C, g=g), out=torch.zeros(B, h, w, C).half(), dd=torch.zeros(B, C), db=torch.zeros(B, C),
             dn=torch.zeros(B, C), out0=torch.zeros(B, h, w, C).half(), dd0=torch.zeros(B, C))

    def run(d):
        ops.style_act_bwd_params(d['da'], d['a'], d['nz'], d['gain'], d['bias'], d['osc'], 4.0, d['out'], d['dd'], d['db'], d['dn'])
        ops.style_act_bwd(d['da'], d['a'], d['nz'], d['gain'], d['bias'], d['osc'], 4.0, d['out0'], d['dd0'])
    gpu, cpu = both(run, t)
    assert torch.equal(gpu['out'], gpu['out0'])
    close('dd', gpu['dd'], gpu['dd0'].cpu(), rel=1e-5)
    for k in ('out', 'dd', 'db', 'dn'):
        close(k, gpu[k], cpu[k], rel=3e-3)
    t = dict(drgb=rn(B, 3, h, w, g=g), a=rn(B, h, w, C, g=g).half(), w=rn(3, C, g=g) / math.sqrt(C), s=1 + 0.5 * rn(B, C, g=g),
             da=torch.zeros(B, h, w, C).half(), ds=torch.zeros(B, C), R=torch.zeros(B, 3, C), bias=torch.zeros(3))

    def run2(d):
        ops.to_rgb_bwd_params(d['drgb'], d['a'], d['w'], d['s'], d['da'], False, d['ds'], d['R'])
        ops.plane_sums(d['drgb'], d['bias'])
    gpu, cpu = both(run2, t)
    for k in ('da', 'ds', 'R', 'bias'):
        close(k, gpu[k], cpu[k], rel=3e-3)


def test_decoder_parameter_folds():
    """table_colsum (fp32 / fp16 input, with and without the per-image weights), mod_linear_wgrad, modconv_wgrad (both
    GEMM-result layouts) against the simulator."""
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, cin, cout, L, Fd = 5, 96, 40, 6, 256
    t = dict(tab=rn(B, 3, cin, g=g), s=1 + 0.5 * rn(B, cin, g=g), g0=rn(B, 4, 12, cin, g=g).half(), ds=rn(B, cin, g=g),
             lat=rn(B, L, Fd, g=g), G=rn(cout, 9, cin, g=g), Gt=rn(cin, 9, cout, g=g), W=rn(cout, cin, 3, 3, g=g),
             dd=rn(B, cout, g=g), d=torch.rand(B, cout, generator=g) + 0.5)
    res = {}

    def run(d):
        res['plain'] = ops.table_colsum(d['ds'])
        res['rgb'] = ops.table_colsum(d['tab'], mul=d['s'], scale=0.25)
        res['const'] = ops.table_colsum(d['g0'], mul=d['s'])
        res['total'] = ops.table_colsum(res['plain'].view(cin, 1))
        res['mod_w'] = ops.mod_linear_wgrad(d['ds'], d['lat'], 1.0 / math.sqrt(Fd), 4)
        res['dw'] = ops.modconv_wgrad(d['G'], False, d['W'], d['s'], d['dd'], d['d'], 1.0 / math.sqrt(cin * 9))
        res['dwt'] = ops.modconv_wgrad(d['Gt'], True, d['W'], d['s'], d['dd'], d['d'], 1.0 / math.sqrt(cin * 9))
    gpu = {k: v.cuda() for k, v in t.items()}
    run(gpu)
    torch.cuda.synchronize()
    got = dict(res)
    with cabi_sim.installed():
        run(t)
    assert got['const'].shape == (4 * 12 * cin,) and got['dw'].shape == (cout, cin, 3, 3)
    for k in got:
        close(k, got[k], res[k], rel=1e-4)


def test_style_tables_backward():
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, cin, cout, L, Fd = 5, 512, 128, 12, 256
    t = dict(ds=rn(B, cin, g=g), s=1 + 0.5 * rn(B, cin, g=g), dd=rn(B, cout, g=g), d=torch.rand(B, cout, generator=g) + 0.5,
             wsq=torch.rand(cout, cin, generator=g) * 9, w=rn(cin, Fd, g=g), dlat=rn(B, L, Fd, g=g))

    def run(d):
        ops.demod_bwd(d['ds'], d['s'], d['dd'], d['d'], d['wsq'], 1.0 / (cin * 9))
        ops.mod_linear_bwd(d['ds'], d['w'], 1.0 / math.sqrt(Fd), d['dlat'], 7)
    gpu, cpu = both(run, t)
    close('ds', gpu['ds'], cpu['ds'], rel=1e-4)
    close('dlat', gpu['dlat'], cpu['dlat'], rel=1e-4)


@pytest.mark.parametrize('B,H,W,cout', [(3, 16, 48, 32), (2, 9, 11, 128)])
def test_first_conv_dgrad_and_heads(B, H, W, cout):
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(cout)
    t = dict(dz=rn(B, H, W, cout, g=g).half(), w=rn(cout, 3, g=g), dx=rn(B, 3, H, W, g=g))
    for acc in (False, True):
        gpu, cpu = both(lambda d: ops.first_conv_dgrad(d['dz'], d['w'], d['dx'], acc), t)
        close('dx', gpu['dx'], cpu['dx'], rel=1e-4)
    t = dict(head=rn(B, H, W, 16, g=g).half(), rgb=torch.zeros(B, 3, H, W), drgb=rn(B, 3, H, W, g=g), dhead=torch.ones(B, H, W, 16).half())

    def run(d):
        ops.head_to_nchw(d['head'], d['rgb'])
        ops.nchw_to_head(d['drgb'], d['dhead'])
    gpu, cpu = both(run, t)
    assert torch.equal(gpu['rgb'].cpu(), cpu['rgb']) and torch.equal(gpu['dhead'].cpu(), cpu['dhead'])


def test_losses():
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(2)
    n = 3 * 3 * 128 * 384 + 5
    x = rn(n, g=g)
    tt = rn(n, g=g)
    tt[:7] = x[:7]                                                          # ties: gradient 0, as torch.sign
    t = dict(x=x, t=tt, loss=torch.zeros(1), grad=torch.zeros(n))
    gpu, cpu = both(lambda d: ops.l1_loss(d['x'], d['t'], 0.1, 8192.0, d['loss'], d['grad']), t)
    close('l1', gpu['loss'], cpu['loss'], rel=1e-5)
    close('l1 grad', gpu['grad'], cpu['grad'], rel=1e-6)
    assert torch.equal(torch.sign(gpu['grad'].cpu()), torch.sign(cpu['grad']))
    for sign in (-1.0, 1.0):
        buf = (rn(37, 16, g=g) * 20).half()                                  # scores are column 0 of a 16-wide GEMM output
        t = dict(pred=buf, loss=torch.zeros(1), dpred=torch.zeros(37, 16).half())

        def run(d):
            ops.softplus_loss(d['pred'][:, :1], sign, 0.1, 4096.0, d['loss'], d['dpred'][:, :1])
        gpu, cpu = both(run, t)
        close('softplus', gpu['loss'], cpu['loss'], rel=1e-5)
        close('dpred', gpu['dpred'], cpu['dpred'], rel=2e-3)
        assert (gpu['dpred'][:, 1:] == 0).all()


@pytest.mark.parametrize('cout,cin,k', [(64, 32, 3), (256, 256, 3), (32, 64, 1), (16, 3, 3)])
def test_pack_weights_matches_the_torch_packs(cout, cin, k):
    from image_restoration_b200 import ops
    from image_restoration_b200.backward import pack_equal_conv
    g = torch.Generator().manual_seed(cin)
    w = rn(cout, cin, k, k, g=g).cuda()
    scale = 1.0 / math.sqrt(cin * k * k)
    ref, _ = pack_equal_conv(w)
    got = ops.pack_weights(w, scale, 0)
    assert torch.equal(got, ref)
    if k == 3 and cin % 8 == 0:
        assert torch.equal(ops.pack_weights(w, scale, 1), ops.conv_dgrad_weight(ref, cin))
    pad = (cin + 15) // 16 * 16
    gp = ops.pack_weights(w, scale, 0, cin_pad=pad).view(cout, k * k, pad)
    assert torch.equal(gp[..., :cin].reshape(cout, -1), ref) and (gp[..., cin:] == 0).all()


def test_conv_plan_launch_equals_direct_launch():
    """b200ir_conv_plan_create / launch / destroy against b200ir_conv_igemm on the same descriptor (generic and row kernels),
    and the plan cache of ops.ConvOp hands the same handle back for an identical descriptor."""
    import ctypes as C
    from image_restoration_b200 import _lib, ops
    lib = _lib.lib()
    g = torch.Generator().manual_seed(4)
    for (B, H, W, cin, cout) in ((2, 16, 48, 256, 256), (8, 128, 384, 32, 32)):
        x = rn(B, H, W, cin, g=g).half().cuda()
        wgt = (rn(cout, 9 * cin, g=g) / math.sqrt(9 * cin)).half().cuda()
        bias = rn(cout, g=g).cuda()
        o1 = torch.empty(B, H, W, cout, device='cuda', dtype=torch.float16)
        o2 = torch.empty_like(o1)
        op1 = ops.conv_same(x, wgt, o1, 3, bias=bias, act=True)
        _lib.check(lib.b200ir_conv_igemm(C.byref(op1.desc), ops._stream()), 'direct')
        op2 = ops.conv_same(x, wgt, o2, 3, bias=bias, act=True)
        op2()
        op2()
        torch.cuda.synchronize()
        assert torch.equal(o1, o2)
        op3 = ops.conv_same(x, wgt, o2, 3, bias=bias, act=True)
        op3()
        assert op3._plan is op2._plan
    h = C.c_void_p()
    bad = ops.conv_same(x, wgt, o2, 3).desc
    bad.block_n = 48
    assert lib.b200ir_conv_plan_create(C.byref(bad), C.byref(h)) != 0 and not h.value
    lib.b200ir_conv_plan_destroy(None)


@pytest.mark.parametrize('B,h,w,C,group', [(8, 4, 12, 512, 4), (4, 4, 12, 64, 4), (6, 3, 5, 32, 2), (2, 4, 4, 24, 2)])
def test_minibatch_stddev_tangent_kernels_and_sum_squares(B, h, w, C, group):
    """First and second derivative of the minibatch standard deviation along a tangent (the two non-GEMM pieces of the R1
    penalty) against torch.autograd.functional.jvp / hvp of the oracle formula (tests/cabi_sim.py)."""
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(B * C)
    t = dict(x=rn(B, h, w, C, g=g).half(), t=rn(B, h, w, C, g=g).half(), a=rn(B // group, g=g),
             sq=rn(B, 3, h, w, g=g), out=torch.zeros(1))
    res = {}

    def run(d):
        res['tcat'] = ops.minibatch_stddev_jvp(d['x'], d['t'], group)
        res['q'] = ops.minibatch_stddev_hvp(d['x'], d['t'], d['a'], group)
        ops.sum_squares(d['sq'], 0.25, d['out'])
    gpu = {k: v.cuda() for k, v in t.items()}
    run(gpu)
    torch.cuda.synchronize()
    got = dict(res)
    cpu = {k: v.clone() for k, v in t.items()}
    with cabi_sim.installed():
        run(cpu)
    c_pad = got['tcat'].shape[3]
    assert c_pad % 16 == 0 and c_pad > C
    close('tcat', got['tcat'], res['tcat'])
    assert (got['tcat'][..., C + 1:] == 0).all()
    close('stat tangent', got['tcat'][..., C], res['tcat'][..., C], rel=2e-3)
    close('hvp', got['q'], res['q'], rel=3e-3)
    close('sum_squares', gpu['out'], cpu['out'], rel=1e-5)


@pytest.mark.parametrize('B,H,W,C', [(3, 16, 48, 64), (2, 6, 10, 512), (2, 128, 384, 64), (5, 2, 2, 24)])
def test_maxpool2_relu_and_l1_f16(B, H, W, C):
    """ReLU + MaxPool2d(2, 2) of the VGG feature extractor, its backward (first maximum of the window, only when positive,
    plus the addend) and the fp16 L1 loss, against torch (tests/cabi_sim.py)."""
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    z = rn(B, H, W, C, g=g).half()
    z[0, 0, 0] = z[0, 0, 1]                                                                          # an exact tie in one window
    t = dict(z=z, out=torch.zeros(B, H // 2, W // 2, C).half(), dpool=rn(B, H // 2, W // 2, C, g=g).half(),
             add=rn(B, H, W, C, g=g).half(), dz=torch.zeros(B, H, W, C).half(), dz2=torch.zeros(B, H, W, C).half(),
             t2=rn(B, H, W, C, g=g).half(), loss=torch.zeros(1), grad=torch.zeros(B, H, W, C).half())

    def run(d):
        ops.maxpool2_relu(d['z'], d['out'])
        ops.maxpool2_relu_bwd(d['z'], d['dpool'], d['add'], d['dz'])
        ops.maxpool2_relu_bwd(d['z'], d['dpool'], None, d['dz2'])
        ops.l1_loss_f16(d['z'], d['t2'], 0.1, 4096.0, d['loss'], d['grad'])
    gpu, cpu = both(run, t)
    assert torch.equal(gpu['out'].cpu(), cpu['out'])
    close('dz', gpu['dz'], cpu['dz'], rel=1e-3)
    assert torch.equal(gpu['dz2'].cpu(), cpu['dz2'])
    close('l1_f16', gpu['loss'], cpu['loss'], rel=1e-4)
    close('l1_f16 grad', gpu['grad'], cpu['grad'], rel=1e-3)


@pytest.mark.parametrize('B,h,w,C', [(3, 16, 48, 64), (5, 8, 24, 256), (2, 128, 384, 64), (4, 4, 12, 512), (3, 6, 10, 128)])
def test_gram_batched_and_per_image_weights(B, h, w, C):
    """The two batched pieces of the style loss: per-image Gram matrices on the weight-gradient GEMM (no sum over the batch)
    and the 1x1 conv whose weight matrix differs per image (w_per_image), both against torch (tests/cabi_sim.py)."""
    from image_restoration_b200 import ops
    g = torch.Generator().manual_seed(B * C + h)
    t = dict(f=rn(B, h, w, C, g=g).half(), gram=torch.zeros(B, C, C), sm=torch.randint(-2, 3, (B, C, C), generator=g).half(),
             scale=torch.full((B, C), 0.37), grad=rn(B, h, w, C, g=g).half())

    def run(d):
        ops.gram_batched(d['f'], out=d['gram'])
        ops.conv_same(d['f'], d['sm'], d['grad'], 1, demod=d['scale'], res=d['grad'], res_mode=1, res_strides=(C, w * C, h * w * C),
                      res_wh=(w, h), res_scale=1.0, res_mul=1.0, w_per_image=True)()
    gpu, cpu = both(run, t)
    close('gram', gpu['gram'], cpu['gram'], rel=1e-3)
    assert torch.allclose(gpu['gram'].cpu(), gpu['gram'].cpu().transpose(1, 2), rtol=1e-4, atol=1e-3 * cpu['gram'].abs().max().item())
    close('per-image conv', gpu['grad'], cpu['grad'], rel=2e-3)
