"""Host mirror of tcnn::default_rng_t (pcg32, PCG-XSH-RR 64/32) — the state `Testbed::m_rng` carries between calls.

The device draws the numbers (csrc/nsb_device.cuh::Pcg32); the host only seeds and advances the state it hands to
nsb_update_density_grid, like Testbed does (`m_rng = default_rng_t{m_seed}`, testbed.cu; `m_rng.advance()` after each
sample-generation launch, testbed_nerf.cu:3576,3589).
"""
from __future__ import annotations

_M64 = (1 << 64) - 1
PCG32_DEFAULT_STATE = 0x853C49E6748FEA9B
PCG32_DEFAULT_STREAM = 0xDA3E39CB94B95BDB
PCG32_MULT = 0x5851F42D4C957F2D


class Pcg32:
    def __init__(self, initstate: int = PCG32_DEFAULT_STATE, initseq: int = PCG32_DEFAULT_STREAM):
        self.state = 0
        self.inc = ((initseq << 1) | 1) & _M64
        self.next_uint()
        self.state = (self.state + initstate) & _M64
        self.next_uint()

    def next_uint(self) -> int:
        old = self.state
        self.state = (old * PCG32_MULT + self.inc) & _M64
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def advance(self, delta: int = 1 << 32) -> None:
        """LCG skip-ahead; the default delta is tcnn's `advance()` default (2^32)."""
        delta &= _M64
        cur_mult, cur_plus, acc_mult, acc_plus = PCG32_MULT, self.inc, 1, 0
        while delta > 0:
            if delta & 1:
                acc_mult = (acc_mult * cur_mult) & _M64
                acc_plus = (acc_plus * cur_mult + cur_plus) & _M64
            cur_plus = ((cur_mult + 1) * cur_plus) & _M64
            cur_mult = (cur_mult * cur_mult) & _M64
            delta >>= 1
        self.state = (acc_mult * self.state + acc_plus) & _M64

    def copy(self) -> "Pcg32":
        r = Pcg32.__new__(Pcg32)
        r.state, r.inc = self.state, self.inc
        return r
