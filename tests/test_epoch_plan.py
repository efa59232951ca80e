"""Host side of the one-launch learner kernel (csrc/epoch2.cu): the shared-memory / work plan that decides whether a pair of
networks is taken by the kernel or falls back to the launch chain, and the workspace it asks for.  No GPU needed: planning is
host code behind the C-ABI (`sb200_ppo_epochs2_supported`, `sb200_ppo_epochs2_workspace_bytes`)."""
import ctypes as C

import torch

from surreal_b200 import _lib, ops


def _block(dims, mode, M=1024):
    acts = [ops.ACT_RELU] * (len(dims) - 2) + [ops.ACT_TANH if mode != 2 else ops.ACT_NONE]
    net = ops.FlatNet(dims, acts, 'cpu', extra=(dims[-1] if mode != 2 else 0))
    tr = ops.MlpTrainer(net, M, 1e-4)
    x, stats = torch.zeros(M, dims[0]), torch.zeros(32)
    if mode != 2:
        A = dims[-1]
        return ops.EpochKernel(tr, mode, x, dims[0], M, None, 1e-2, stats, 10, actions=torch.zeros(M, A), lda=A,
                               adv=torch.zeros(M), behave_pd=torch.zeros(M, 2 * A), ldb=2 * A, ref_pd=torch.zeros(M, 2 * A),
                               ldr=2 * A, hyper=torch.zeros(2, dtype=torch.float64), grid=148)
    return ops.EpochKernel(tr, 2, x, dims[0], M, None, 1e-2, stats, 10, returns=torch.zeros(M), grid=148)


def _ws(p, v):
    return int(_lib.lib().sb200_ppo_epochs2_workspace_bytes(C.byref(p.args), C.byref(v.args) if v is not None else None))


def test_bench_shape_is_taken_and_workspace_scales_with_the_batch():
    p, v = _block([64, 256, 256, 8], 0), _block([64, 256, 256, 1], 2)
    assert ops.EpochPair.supported(p, v)
    assert ops.EpochPair.supported(p, None)                      # a single optimiser is a valid launch too
    small = _ws(p, v)
    p4, v4 = _block([64, 256, 256, 8], 1, M=4096), _block([64, 256, 256, 1], 2, M=4096)
    assert ops.EpochPair.supported(p4, v4)
    big = _ws(p4, v4)
    # per row block (16 rows) and network: KL partial + loss partials + the dW3 partial ((H2 + 1) * ru4(out) floats)
    assert big > small > 2 * 256 * 256 * 4                       # at least the two transposed W2 copies
    assert big - small >= (4096 - 1024) // 16 * ((256 + 1) * 8 * 4)


def test_ragged_shapes_are_taken():
    assert ops.EpochPair.supported(_block([17, 300, 200, 6], 0, M=200), _block([17, 100, 68, 1], 2, M=200))
    assert ops.EpochPair.supported(_block([20, 64, 64, 1], 1, M=77), _block([20, 64, 64, 1], 2, M=77))


def test_shapes_the_kernel_rejects_fall_back():
    v = _block([64, 256, 256, 1], 2)
    assert not ops.EpochPair.supported(_block([64, 600, 256, 8], 0), v)          # wider than the shared-memory plan allows
    assert not ops.EpochPair.supported(_block([64, 256, 256, 40], 0), v)         # head wider than 32
    four = ops.FlatNet([64, 128, 128, 128, 8], [ops.ACT_RELU] * 3 + [ops.ACT_TANH], 'cpu', extra=8)
    assert four.n_layers == 4                                                    # PPOLearner._epoch_kernels screens these out
