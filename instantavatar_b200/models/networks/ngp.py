"""Host-side mirror of instant_avatar/models/networks/ngp.py::NeRFNGPNet.

State-dict layout is the reference's: `encoder.params` (flat fp32, [W1 64x32 | W2 16x64 | hash grid]) and
`color_net.params` (flat fp32, [W3 64x16 | W4 64x64 | W5 16x64]) -- the tiny-cuda-nn module parameter order --
plus buffers `center`, `scale`.  The arithmetic (16-level hash grid, 64-wide fused MLPs) runs in libia_b200.so on
fp16 working copies of the parameters that are refreshed whenever the fp32 masters change (`mark_dirty`).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ... import _lib, ops

EPS = 1e-3


class _ParamModule(nn.Module):
    """stand-in for tcnn.NetworkWithInputEncoding / tcnn.Network: one flat fp32 Parameter named `params`"""

    def __init__(self, n: int):
        super().__init__()
        self.params = nn.Parameter(torch.zeros(n, dtype=torch.float32))


def tcnn_like_init(seed: int = 1337):
    """tiny-cuda-nn style initialisation: hash tables U(-1e-4, 1e-4), Xavier-uniform weight matrices."""
    rng = np.random.default_rng(seed)
    total = _lib.hashgrid_layout()["total"] if _lib_available() else 6513496
    lim = lambda fi, fo: np.sqrt(6.0 / (fi + fo))
    mk = lambda o, i: rng.uniform(-lim(i, o), lim(i, o), (o, i)).astype(np.float32).ravel()
    enc = np.concatenate([mk(64, 32), mk(16, 64), rng.uniform(-1e-4, 1e-4, total * 2).astype(np.float32)])
    col = np.concatenate([mk(64, 16), mk(64, 64), mk(16, 64)])
    return enc, col


def _lib_available():
    import os
    return os.path.exists(_lib.LIB_PATH)


class NeRFNGPNet(nn.Module):
    def __init__(self, opt=None, seed: int = 1337):
        super().__init__()
        total = 6513496
        self.encoder = _ParamModule(_lib.IA_ENC_MLP_PARAMS + 2 * total)
        self.color_net = _ParamModule(_lib.IA_COL_MLP_PARAMS)
        enc, col = tcnn_like_init(seed)
        with torch.no_grad():
            self.encoder.params.copy_(torch.from_numpy(enc))
            self.color_net.params.copy_(torch.from_numpy(col))
        center = getattr(opt, "center", None) if opt is not None and not isinstance(opt, dict) else (opt or {}).get("center")
        scale = getattr(opt, "scale", None) if opt is not None and not isinstance(opt, dict) else (opt or {}).get("scale")
        self.register_buffer("center", torch.FloatTensor(list(center) if center is not None else [0.0, 0.0, 0.0]))
        self.register_buffer("scale", torch.FloatTensor(list(scale) if scale is not None else [1.0, 1.0, 1.0]))
        self.opt = opt
        self._table_h = None
        self._mlp_h = None
        self._dirty = True

    def initialize(self, bbox):
        """ngp.py:64-71"""
        if hasattr(self, "bbox"):
            return
        c = (bbox[0] + bbox[1]) / 2
        s = bbox[1] - bbox[0]
        self.center = c
        self.scale = s
        self.bbox = bbox

    def adopt_half_table(self, table_h):
        """use `table_h` ([total, 2] fp16, e.g. a view into the optimiser's flat fp16 image) as the working copy of the
        hash table from now on; refreshed from the fp32 masters on the next half_params() call"""
        self._table_h = table_h
        self._dirty = True

    def mark_dirty(self):
        self._dirty = True

    def mark_clean(self):
        """the optimiser refreshed the fp16 copies itself"""
        self._dirty = False

    def half_buffers(self):
        """persistent fp16 working copies (allocated on first use), without refreshing them"""
        if self._table_h is None or self._mlp_h is None or self._table_h.device != self.encoder.params.device:
            self.half_params()
        return self._table_h, self._mlp_h

    def grad_buffers(self):
        """persistent flat fp32 gradient buffers the fused backward accumulates into (aliased as `.grad`)"""
        p, c = self.encoder.params, self.color_net.params
        if p.grad is None or p.grad.device != p.device:
            p.grad = torch.zeros_like(p)
        if c.grad is None or c.grad.device != c.device:
            c.grad = torch.zeros_like(c)
        return p.grad, c.grad

    def load_flat_params(self, enc, col):
        with torch.no_grad():
            self.encoder.params.copy_(torch.as_tensor(enc, dtype=torch.float32))
            self.color_net.params.copy_(torch.as_tensor(col, dtype=torch.float32))
        self._dirty = True

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._dirty = True

    def half_params(self):
        """fp16 working copies (tcnn casts its fp32 params to fp16 each forward; here once per change)"""
        if self._dirty or self._table_h is None or self._table_h.device != self.encoder.params.device:
            self._table_h, self._mlp_h = ops.params_to_half(self.encoder.params.detach(), self.color_net.params.detach(),
                                                            self._table_h if self._table_h is not None and self._table_h.device == self.encoder.params.device else None,
                                                            self._mlp_h if self._mlp_h is not None and self._mlp_h.device == self.encoder.params.device else None)
            self._dirty = False
        return self._table_h, self._mlp_h

    def forward(self, x, d=None, cond=None):
        """ngp.py:73-83: canonical points -> (rgb [P,3] f32, sigma [P] f32).  The fused train kernels bypass this entry
        (instantavatar_b200/autograd.py); called directly it is differentiable w.r.t. the parameters and the points."""
        table_h, mlp_h = self.half_params()
        sc = ops.Scene(table_h=table_h, mlp_h=mlp_h, net_center=self.center.reshape(3).contiguous().float(),
                       net_scale=self.scale.reshape(3).contiguous().float())
        if torch.is_grad_enabled() and (x.requires_grad or self.encoder.params.requires_grad):
            from ...autograd import _NGPForward
            accum = self.grad_buffers() if self.encoder.params.requires_grad else None
            return _NGPForward.apply(x, self.encoder.params, self.color_net.params, sc, accum)
        rgb, sigma = ops.ngp_forward(sc, x.reshape(-1, 3).float().contiguous())
        return rgb, sigma
