"""NHWC activation views with an explicit pixel stride (ld).

A `View` is a channel slice [off, off+C) of a wider NHWC buffer [B, H, W, ld].  Producers write their
output straight into the slice of the consumer's concat buffer, which is how torch.cat
(reference ddpm.py:459,462,468) disappears from the B200 path.
"""
from __future__ import annotations

import torch


class View:
    __slots__ = ("buf", "off", "C")

    def __init__(self, buf: torch.Tensor, off: int = 0, C: int | None = None):
        assert buf.dim() == 4 and buf.is_contiguous()
        self.buf = buf
        self.off = off
        self.C = buf.shape[3] - off if C is None else C
        assert 0 <= off and off + self.C <= buf.shape[3]

    # geometry -------------------------------------------------------------------------------
    @property
    def B(self):
        return self.buf.shape[0]

    @property
    def H(self):
        return self.buf.shape[1]

    @property
    def W(self):
        return self.buf.shape[2]

    @property
    def ld(self):
        return self.buf.shape[3]

    @property
    def ptr(self):
        return self.buf.data_ptr() + self.off * self.buf.element_size()

    @property
    def dtype(self):
        return self.buf.dtype

    def slice(self, off: int, C: int) -> "View":
        return View(self.buf, self.off + off, C)

    # conversions (tests / API boundary only) ---------------------------------------------------
    def to_nchw(self) -> torch.Tensor:
        return self.buf[..., self.off:self.off + self.C].permute(0, 3, 1, 2).float().contiguous()

    def from_nchw(self, x: torch.Tensor):
        self.buf[..., self.off:self.off + self.C] = x.permute(0, 2, 3, 1).to(self.buf.dtype)
        return self

    @staticmethod
    def empty(B, H, W, C, dtype, device="cuda", ld=None, off=0):
        ld = C if ld is None else ld
        return View(torch.empty(B, H, W, ld, dtype=dtype, device=device), off, C)

    @staticmethod
    def zeros(B, H, W, C, dtype, device="cuda", ld=None, off=0):
        ld = C if ld is None else ld
        return View(torch.zeros(B, H, W, ld, dtype=dtype, device=device), off, C)
