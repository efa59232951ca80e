"""CNN stem kernels (csrc/stem.cu) against torch.nn.functional.conv2d + autograd on the CPU (the calls the reference's
CNNStemNetwork issues, builders.py:8-33): forward on uint8 / float frames, weight / bias gradients through the slab
reduction, input gradient of the second convolution."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _kernel_layout(w):
    """torch [COUT][C][KS][KS] -> kernel layout [(c, ky, kx)][COUT]."""
    co = w.shape[0]
    return w.reshape(co, -1).t().contiguous()


@pytest.mark.parametrize('layer,C_in,H,W,frames,u8', [
    (1, 4, 84, 84, 37, True),       # cfg4 frame
    (1, 2, 20, 20, 6, True),        # golden-fixture frame
    (1, 3, 33, 45, 5, False),       # odd sizes, float input
    (2, 16, 20, 20, 37, False),     # second conv on the first one's 20x20 output
    (2, 16, 4, 4, 6, False),        # golden-fixture size: 1x1 output
    (2, 16, 7, 10, 9, False),
])
def test_conv_forward_and_backward_match_torch(layer, C_in, H, W, frames, u8):
    from surreal_b200 import _lib, ops
    L = _lib.lib()
    KS, ST, CO = (8, 4, 16) if layer == 1 else (4, 2, 32)
    g = torch.Generator().manual_seed(layer * 100 + H + frames)
    w = (torch.rand(CO, C_in, KS, KS, generator=g) * 2 - 1) / np.sqrt(C_in * KS * KS)
    b = (torch.rand(CO, generator=g) * 2 - 1) * 0.1
    if u8:
        x = torch.randint(0, 256, (frames, C_in, H, W), generator=g, dtype=torch.uint8)
        xf, scale = x.float(), 1.0 / 255.0
    else:
        x = torch.rand(frames, C_in, H, W, generator=g) * (2.0 if layer == 2 else 1.0)
        if layer == 2:
            x = torch.relu(x - 0.7)                                # a post-ReLU activation map with real zeros
        xf, scale = x, 1.0
    xin = (xf * scale).requires_grad_(True)
    wt, bt = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = torch.relu(F.conv2d(xin, wt, bt, stride=ST))
    HO, WO = y_ref.shape[2], y_ref.shape[3]
    dy = torch.randn(y_ref.shape, generator=g) * (y_ref > 0).float()        # gradient w.r.t. the PRE-activation
    F.conv2d(xin, wt, bt, stride=ST).backward(dy)
    xd = x.to(DEV).contiguous()
    wk, bk = _kernel_layout(w).to(DEV), b.to(DEV)
    y = torch.empty(frames, CO, HO, WO, device=DEV)
    st = ops._stream()
    _lib.check(L.sb200_conv_forward_f32(layer, _p(xd), int(u8), frames, C_in, H, W, _p(wk), _p(bk), scale, _p(y), st), 'conv fwd')
    torch.cuda.synchronize()
    assert float((y.cpu() - y_ref.detach()).abs().max()) <= 1e-5
    # weight / bias gradient through slabs
    taps = C_in * KS * KS
    stride = taps * CO + CO
    splits = min(frames, 7)
    slabs = torch.zeros(splits, stride, device=DEV)
    dyd = dy.to(DEV).contiguous()
    _lib.check(L.sb200_conv_backward_dw_f32(layer, _p(xd), int(u8), _p(dyd), frames, C_in, H, W, scale, _p(slabs),
                                            C.c_void_p(slabs.data_ptr() + 4 * taps * CO), stride, splits, st), 'conv dw')
    torch.cuda.synchronize()
    gsum = slabs.sum(0).cpu()
    dw = gsum[:taps * CO].view(taps, CO).t().reshape(CO, C_in, KS, KS)
    db = gsum[taps * CO:]
    tol = 1e-5 * max(1.0, float(wt.grad.abs().max()))
    assert float((dw - wt.grad).abs().max()) <= tol * 5, float((dw - wt.grad).abs().max())
    assert float((db - bt.grad).abs().max()) <= 1e-4 * max(1.0, float(bt.grad.abs().max()))
    if layer == 2:
        dx = torch.empty(frames, C_in, H, W, device=DEV)
        _lib.check(L.sb200_conv_backward_dx_f32(2, _p(dyd), _p(wk), _p(xd), frames, C_in, H, W, _p(dx), st), 'conv dx')
        torch.cuda.synchronize()
        ref = xin.grad * (x > 0).float()
        assert float((dx.cpu() - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
