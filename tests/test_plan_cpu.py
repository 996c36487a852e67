"""Host logic of the GEMM launcher (no GPU): the split-K / CTA-pair plan of every weight-gradient shape of the
BASELINE configurations fills the persistent grid.  (A planner that left the 192x192 wgrads of the tiny heads
unsplit -- 2 tiles on 148 SMs, 281 us instead of 20 -- is what this guards against.)"""
import ctypes as C

import pytest

from theia_b200 import _lib as L

SMS = 148  # the planner's fallback when no device is visible, and the B200's SM count


def plan(nout, kin, mtok, z=1):
    bn, pair, splits = C.c_int(), C.c_int(), C.c_int()
    L.check(L.lib().theia_plan_wgrad(nout, kin, mtok, z, C.byref(bn), C.byref(pair), C.byref(splits)), "plan")
    return bn.value, pair.value, splits.value


def shapes():
    out = []
    for D, B in ((192, 256), (768, 256)):
        tok = B * 197
        out += [(D, D, tok, 1), (3 * D, D, tok, 1), (4 * D, D, tok, 1), (D, 4 * D, tok, 1), (D, 768, tok, 1)]  # backbone
        pix = B * 256
        out += [(D, D, pix, 9), (1024, D, pix, 1), (1280, D, pix, 1)]                                          # 16x16 heads
        out += [(256, D, B * 4096, 1), (32, D, B * 4096, 1), (D, D, B * 1024, 9)]                               # 64x64 heads
    return out


@pytest.mark.parametrize("nout,kin,mtok,z", shapes())
def test_wgrad_plan_fills_the_grid(nout, kin, mtok, z):
    bn, pair, splits = plan(nout, kin, mtok, z)
    assert bn in (128, 192, 256) and pair in (1, 2) and splits >= 1
    assert pair == 1 or bn == 256
    num_kb = (mtok + 63) // 64
    assert num_kb // splits >= 8 or splits == 1, "a split keeps at least 8 K blocks"
    m_tiles = (nout + 127) // 128
    units = ((m_tiles + pair - 1) // pair) * ((kin + bn - 1) // bn) * z
    resident = SMS // pair
    items = units * splits
    waves = -(-items // resident)
    eff = items / (waves * resident)
    assert eff >= 0.6, (bn, pair, splits, items, resident)


def test_bad_arguments_are_rejected():
    bn, pair, splits = C.c_int(), C.c_int(), C.c_int()
    assert L.lib().theia_plan_wgrad(0, 192, 1000, 1, C.byref(bn), C.byref(pair), C.byref(splits)) != 0
    assert L.lib().theia_plan_wgrad(192, 192, 1000, 1, None, C.byref(pair), C.byref(splits)) != 0
