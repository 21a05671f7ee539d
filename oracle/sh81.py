"""Real spherical harmonics up to degree 8 (81 values) of unit vectors, by recurrence.

TEST INFRASTRUCTURE (oracle for the UniDepthV1 ray embedding): the
reference embeds V1's rays with `rsh_cart_8` (unidepth/utils/sht.py:833-1393, called at
unidepthv1/decoder.py:218-220), an auto-generated list of 81 expanded polynomials.  This restates the same
functions from their definition instead: with (x + iy)^m = A_m + i B_m and Pi_l^m(z) = P_l^m(z) / sin^m(theta)
(Condon-Shortley phase included),

    Y_l^0  = K_l^0 Pi_l^0(z)
    Y_l^m  = sqrt(2) K_l^m A_m Pi_l^m(z)        (m > 0)
    Y_l^-m = sqrt(2) K_l^m B_m Pi_l^m(z)
    K_l^m  = sqrt((2l+1)/(4 pi) * (l-m)!/(l+m)!)

stored at index l*(l+1) + m like the reference.  Pinned against the reference's output in
tests/golden/sh81.npz (oracle/make_golden.py); tests/test_oracle_golden.py checks it."""
import math

import torch


def rsh_cart(xyz: torch.Tensor, degree: int = 8) -> torch.Tensor:
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    out = [None] * ((degree + 1) ** 2)
    a, b = torch.ones_like(x), torch.zeros_like(x)       # A_0, B_0
    pmm = torch.ones_like(x)                             # Pi_m^m = (-1)^m (2m-1)!!
    for m in range(degree + 1):
        if m > 0:
            a, b = x * a - y * b, x * b + y * a
            pmm = pmm * (-(2 * m - 1))
        p_prev, p_cur = None, pmm
        for l in range(m, degree + 1):
            if l == m + 1:
                p_prev, p_cur = p_cur, (2 * m + 1) * z * p_cur
            elif l > m + 1:
                p_prev, p_cur = p_cur, ((2 * l - 1) * z * p_cur - (l + m - 1) * p_prev) / (l - m)
            k = math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - m) / math.factorial(l + m))
            if m == 0:
                out[l * (l + 1)] = k * p_cur
            else:
                out[l * (l + 1) + m] = math.sqrt(2.0) * k * a * p_cur
                out[l * (l + 1) - m] = math.sqrt(2.0) * k * b * p_cur
    return torch.stack(out, dim=-1)
