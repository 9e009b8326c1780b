"""Bit-level checks of the numerical building blocks the scene-flow kernel substitutes for IEEE library calls."""
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
    bad = ctx.selfcheck_fdiv(a.contiguous(), b.contiguous())
    torch.cuda.synchronize()
    assert bad == 0, f"{bad} of {a.numel()} operand pairs differ from IEEE float division"


def test_fdiv_random_bit_patterns(ctx):
    """Uniformly random 32-bit patterns: all exponents, subnormals, infinities and NaNs on both sides."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(7)
    for _ in range(4):
        a = torch.randint(-2**31, 2**31 - 1, (1 << 26,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
        b = torch.randint(-2**31, 2**31 - 1, (1 << 26,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
        _check(ctx, a.view(torch.float32), b.view(torch.float32))


def test_fdiv_disparity_like_operands(ctx):
    """The operands the kernel really sees: f*T constants over disparities in (0, 128], dense in mantissa space."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(8)
    for fT in (554.25 * 0.07, 700.0 * 0.12, 721.5377 * 0.5372, 1400.0 * 0.12, 2800.0 * 0.12):
        b = torch.rand((1 << 26,), generator=g, device="cuda", dtype=torch.float32) * 128.0
        a = torch.full_like(b, float(np.float32(fT)))
        _check(ctx, a, b)


def test_fdiv_adversarial(ctx):
    """Quotients engineered to sit next to float rounding boundaries, subnormal results, overflow, special values."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(9)
    n = 1 << 24
    # (1) b with random 24-bit mantissa, a = nearest float to b * (m + 1/2 ulp): quotients hugging midpoints
    mb = torch.randint(1 << 23, 1 << 24, (n,), generator=g, device="cuda", dtype=torch.int64)
    mq = torch.randint(1 << 23, 1 << 24, (n,), generator=g, device="cuda", dtype=torch.int64)
    prod = (2 * mq + 1) * mb  # (q + half ulp) * b as exact 49-bit integer
    a = (prod.to(torch.float64) * 2.0 ** -24).to(torch.float32)  # rounded to float: a/b lands close to a midpoint
    b = mb.to(torch.float32)
    _check(ctx, a, b)
    _check(ctx, a * 2.0 ** -60, b * 2.0 ** 40)   # tiny quotients
    _check(ctx, a * 2.0 ** -100, b * 2.0 ** 26)  # subnormal results
    _check(ctx, a * 2.0 ** 80, b * 2.0 ** -40)   # overflow to inf
    # (2) special values crossed with each other
    sp = torch.tensor([0.0, -0.0, 1.0, -1.0, float("inf"), float("-inf"), float("nan"), 1e-45, -1e-45, 1.17549435e-38,
                       3.4028235e38, -3.4028235e38, 128.0, 0.5, 3.0, 1e-20, 1e20], device="cuda")
    aa, bb = torch.meshgrid(sp, sp, indexing="ij")
    _check(ctx, aa.reshape(-1), bb.reshape(-1))
