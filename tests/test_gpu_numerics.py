"""Bit-level checks of the numerical building blocks the scene-flow kernel substitutes for IEEE library calls:
the shared-reciprocal double division (one refined reciprocal per divisor, several numerators) must equal
__ddiv_rn -- the reference's divsd -- for every operand pair."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from moving_object_detector_b200 import Context
    c = Context(64, 64, max_batch=1)
    yield c
    c.close()


def _check(ctx, a, b):
    import torch
    bad = ctx.selfcheck_ddiv(a.contiguous(), b.contiguous())
    torch.cuda.synchronize()
    assert bad == 0, f"{bad} of {a.numel()} operand pairs differ from IEEE double division"


def test_ddiv_random_bit_patterns(ctx):
    """Uniformly random 64-bit patterns: all exponents, subnormals, infinities and NaNs on both sides."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(7)
    for _ in range(4):
        a = torch.randint(-2**63, 2**63 - 1, (1 << 25,), generator=g, device="cuda", dtype=torch.int64)
        b = torch.randint(-2**63, 2**63 - 1, (1 << 25,), generator=g, device="cuda", dtype=torch.int64)
        _check(ctx, a.view(torch.float64), b.view(torch.float64))


def test_ddiv_scene_like_operands(ctx):
    """What the kernel divides: (fx*X + Tx) / Z with float-valued X, Z widened to double, and float differences / dt."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(8)
    n = 1 << 25
    for fx in (554.25, 700.0, 721.5377, 1400.0, 2800.0):
        Z = (torch.rand(n, generator=g, device="cuda") * 80.0 + 0.3).to(torch.float32).to(torch.float64)
        X = ((torch.rand(n, generator=g, device="cuda") - 0.5) * 60.0).to(torch.float32).to(torch.float64)
        _check(ctx, fx * X + 0.0, Z)
    for dt in (1.0 / 30.0, 1.0 / 15.0, 0.1, 0.25, 1e-9, 3600.0):
        a = ((torch.rand(n, generator=g, device="cuda") - 0.5) * 4.0).to(torch.float32).to(torch.float64)
        _check(ctx, a, torch.full_like(a, dt))


def test_ddiv_exponent_extremes_and_specials(ctx):
    """Quotients that underflow / overflow, subnormal operands, and every special value crossed with every other."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(9)
    n = 1 << 22
    m1 = torch.rand(n, generator=g, device="cuda", dtype=torch.float64) + 1.0
    m2 = torch.rand(n, generator=g, device="cuda", dtype=torch.float64) + 1.0
    for ea, eb in ((-1000, 20), (-1020, 3), (-1060, 0), (1000, -30), (1023, -2), (-500, 520), (0, 1023), (0, -1022),
                   (-1074 + 52, 0), (300, -800)):
        _check(ctx, torch.ldexp(m1, torch.tensor(ea, device="cuda")), torch.ldexp(m2, torch.tensor(eb, device="cuda")))
    sp = torch.tensor([0.0, -0.0, 1.0, -1.0, float("inf"), float("-inf"), float("nan"), 5e-324, -5e-324,
                       2.2250738585072014e-308, 1.7976931348623157e308, -1.7976931348623157e308, 3.0, 1.0 / 3.0,
                       1e-300, 1e300, 0.1, 1.0 / 15.0], device="cuda", dtype=torch.float64)
    aa, bb = torch.meshgrid(sp, sp, indexing="ij")
    _check(ctx, aa.reshape(-1), bb.reshape(-1))
