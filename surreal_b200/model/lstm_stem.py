"""LSTM stem of RNN-mode PPO -- host-side mirror of the ``nn.LSTM(rnn_insize, rnn_hidden, rnn_layer, batch_first=True)`` the
reference's PPOModel puts in front of actor and critic (surreal/model/ppo_net.py:143-152,277-279,342-351; the DEFAULT PPO
config, ppo_configs.py:57-62).  One layer (the reference default; more layers raise).

Parameters live in ONE flat buffer as two FlatNet segments -- W_ih^T [D][4H] + b_ih and W_hh^T [H][4H] + b_hh -- so the
input projection and both weight gradients are ordinary MLP-kernel calls, and each of the two optimisers that train the
shared stem (ppo_net.py:202-224) keeps its own flat Adam state over it.  The recurrence itself is csrc/lstm.cu."""
import ctypes as C

import torch

from .. import _lib, ops
from .._lib import check


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class LSTMStem:
    def __init__(self, in_dim, hidden, device, layers=1):
        if layers != 1:
            raise NotImplementedError('rnn_layer = %d: only the reference default of one LSTM layer is built' % layers)
        assert 4 * hidden <= 1024, 'rnn_hidden up to 256'
        self.D, self.H = int(in_dim), int(hidden)
        self.device = torch.device(device)
        self.ih = ops.FlatNet([self.D, 4 * self.H], [ops.ACT_NONE], self.device)
        self.hh = ops.FlatNet([self.H, 4 * self.H], [ops.ACT_NONE], self.device)
        self.off = {'ih': (0, self.ih.size), 'hh': (self.ih.size, self.hh.size)}
        self.size = self.ih.size + self.hh.size
        self.params = torch.zeros(self.size, dtype=torch.float32, device=self.device)
        self.ih.params, self.hh.params = self.seg('ih'), self.seg('hh')
        ref = torch.nn.LSTM(self.D, self.H, 1, batch_first=True)            # torch's default initialisation
        self.load_torch({k: v.detach() for k, v in ref.state_dict().items()})

    def seg(self, name, buf=None):
        o, n = self.off[name]
        return (self.params if buf is None else buf)[o:o + n]

    def load_torch(self, sd, prefix=''):
        f = lambda k: torch.as_tensor(sd[prefix + k], dtype=torch.float32)  # noqa: E731
        self.ih.set_layer(0, f('weight_ih_l0'), f('bias_ih_l0'))
        self.hh.set_layer(0, f('weight_hh_l0'), f('bias_hh_l0'))

    def state_items(self, prefix='rnn_stem.'):
        wi, bi = self.ih.get_layer(0)
        wh, bh = self.hh.get_layer(0)
        return [(prefix + 'weight_ih_l0', wi), (prefix + 'weight_hh_l0', wh), (prefix + 'bias_ih_l0', bi), (prefix + 'bias_hh_l0', bh)]

    def buffers(self, B, L, save=True):
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)  # noqa: E731
        H, R = self.H, B * L
        Hp = ops._ru(H, 4)                                        # row stride of the h buffers (16-byte rows for the GEMM kernels)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)  # noqa: E731
        b = dict(pre_x=f(R, 4 * H), h_out=z(R, Hp)[:, :H])
        if save:
            b.update(h_prev=z(R, Hp)[:, :H], gates=f(R, 4 * H), c_seq=f(R, H))
        return b

    def forward(self, xf, h0, c0, ld_cells, B, L, bufs, h_last=None, c_last=None, params=None):
        """xf: [B*L, D] (already z-filtered, sequence-major rows b*L + t); h0 / c0: tensors whose row b starts at element
        b*ld_cells (or None: zeros).  -> h_out [B*L, H]."""
        P = self.params if params is None else params
        ops.mlp_forward(self.ih, xf, params=self.seg('ih', P), out=bufs['pre_x'])
        lay = self.hh.layout[0]
        hh = self.seg('hh', P)
        check(_lib.lib().sb200_lstm_forward_f32(
            _p(bufs['pre_x']), C.c_void_p(hh.data_ptr() + 4 * lay['w']), C.c_void_p(hh.data_ptr() + 4 * lay['b']), _p(h0), _p(c0),
            int(ld_cells), B, L, self.H, bufs['h_out'].stride(0), _p(bufs['h_out']), _p(bufs.get('h_prev')), _p(bufs.get('gates')), _p(bufs.get('c_seq')),
            _p(h_last), _p(c_last), ops._stream()), 'sb200_lstm_forward_f32')
        return bufs['h_out']


class RnnTrainer:
    """Saved activations, BPTT and gradient slabs of the LSTM stem for ONE optimiser on B sequences of L steps."""

    def __init__(self, stem, B, L, splits=None):
        self.stem, self.B, self.L = stem, B, L
        dev = stem.device
        R = B * L
        self.splits = splits if splits is not None else max(1, min(16, R // 128))
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)  # noqa: E731
        self.bufs = stem.buffers(B, L, save=True)
        self.slabs = z(self.splits, stem.size)
        self.exp_avg, self.exp_avg_sq = z(stem.size), z(stem.size)
        self.dpre = z(R, 4 * stem.H)
        self.xf = self.h0 = self.c0 = None
        self.ld_cells = 0

    def forward(self, xf, h0, c0, ld_cells):
        self.xf, self.h0, self.c0, self.ld_cells = xf, h0, c0, ld_cells
        return self.stem.forward(xf, h0, c0, ld_cells, self.B, self.L, self.bufs)

    def backward(self, dh_out):
        """dh_out [B*L, >= H]: gradient w.r.t. the LSTM outputs (the head's input gradient, no activation in between)."""
        s, L_, st = self.stem, _lib.lib(), ops._stream()
        b, R = self.bufs, self.B * self.L
        lay_h, lay_i = s.hh.layout[0], s.ih.layout[0]
        check(L_.sb200_lstm_backward_f32(_p(dh_out), dh_out.stride(0), _p(b['gates']), _p(b['c_seq']), _p(self.c0), int(self.ld_cells),
                                         C.c_void_p(s.hh.params.data_ptr() + 4 * lay_h['w']), self.B, self.L, s.H, _p(self.dpre), st),
              'sb200_lstm_backward_f32')
        base = self.slabs.data_ptr()
        o_ih, o_hh = s.off['ih'][0], s.off['hh'][0]
        G = 4 * s.H
        check(L_.sb200_linear_bwd_dw_f32(_p(self.xf), self.xf.stride(0), _p(self.dpre), G, C.c_void_p(base + 4 * (o_ih + lay_i['w'])),
                                         C.c_void_p(base + 4 * (o_ih + lay_i['b'])), s.size, self.splits, lay_i['ldw'], R, s.D, G, st),
              'sb200_linear_bwd_dw_f32(lstm ih)')
        check(L_.sb200_linear_bwd_dw_f32(_p(b['h_prev']), b['h_prev'].stride(0), _p(self.dpre), G, C.c_void_p(base + 4 * (o_hh + lay_h['w'])),
                                         C.c_void_p(base + 4 * (o_hh + lay_h['b'])), s.size, self.splits, lay_h['ldw'], R, s.H, G, st),
              'sb200_linear_bwd_dw_f32(lstm hh)')

    def state_dict(self):
        return {'exp_avg': self.exp_avg.detach().clone(), 'exp_avg_sq': self.exp_avg_sq.detach().clone()}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(torch.as_tensor(sd['exp_avg']).to(self.stem.device))
        self.exp_avg_sq.copy_(torch.as_tensor(sd['exp_avg_sq']).to(self.stem.device))


def rows_zfilter(x, row_stride, batch_stride, B, L, D, zf_stats, zf_eps, out):
    """out[b*L + t, :D] = zfilter(x[b*batch_stride + t*row_stride : +D]) (plain gather when zf_stats is None)."""
    check(_lib.lib().sb200_rows_zfilter_f32(_p(x), int(row_stride), int(batch_stride), B, L, D, _p(zf_stats), float(zf_eps), _p(out),
                                            out.stride(0), ops._stream()), 'sb200_rows_zfilter_f32')
    return out
