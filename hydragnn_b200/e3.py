"""Host-side equivariance tables for the MACE path (set-up time only; nothing here runs per step).

The reference gets these from e3nn 0.5.1 (`o3.Irreps`, `o3.wigner_3j`, `o3.SphericalHarmonics`; call sites
hydragnn/models/MACEStack.py:144-159,195-311 and hydragnn/utils/model/mace_utils/tools/cg.py:58).  The engine needs
only numbers: which (l1, l2, l3) paths exist, the real coupling tensors, the generalised coupling tensors of the
symmetric contraction, and the closed-form harmonics.  Irreps are plain lists of (mul, l, parity) here; features on the
device are stored per degree l as [N, 2l+1, channels] ("channel-last"), not in e3nn's mul-major rows.
"""
import functools
import math
from fractions import Fraction

import torch


# ---- irreps as [(mul, l, p)] ------------------------------------------------------------------------------------
def parse_irreps(text):
    out = []
    for part in str(text).split("+"):
        part = part.strip()
        if not part:
            continue
        mul, ir = part.split("x") if "x" in part else ("1", part)
        out.append((int(mul), int(ir[:-1]), 1 if ir[-1] == "e" else -1))
    return out


def hidden_irreps(channels, lmax):
    """create_irreps_string (hydragnn/utils/model/irreps_tools.py:105-109): parity (-1)^l."""
    return [(channels, l, (-1) ** l) for l in range(lmax + 1)]


def irreps_dim(irreps):
    return sum(m * (2 * l + 1) for m, l, _ in irreps)


def irreps_str(irreps):
    return "+".join("%dx%d%s" % (m, l, "e" if p == 1 else "o") for m, l, p in irreps)


def tp_paths(lmax_in, lmax_sh, lmax_out):
    """Instruction list of tp_out_irreps_with_instructions (irreps_tools.py:15-44) for inputs F x (0e .. lmax_in),
    spherical harmonics up to lmax_sh and a target holding every l <= lmax_out with parity (-1)^l.  Returned in the
    order of the per-edge weight blocks: sorted by output degree, ties in generation order (l1 outer, l2 inner)."""
    gen = []
    for l1 in range(lmax_in + 1):
        for l2 in range(lmax_sh + 1):
            for l3 in range(abs(l1 - l2), l1 + l2 + 1):
                if l3 <= lmax_out and (l1 + l2 + l3) % 2 == 0:
                    gen.append((l1, l2, l3))
    return sorted(gen, key=lambda t: t[2])       # stable


# ---- real Wigner 3j --------------------------------------------------------------------------------------------
def _fact(n):
    return math.factorial(round(n))


def _cg_complex(j1, m1, j2, m2, j3, m3):
    if m3 != m1 + m2:
        return 0.0
    lo = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    hi = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))
    pre = ((2.0 * j3 + 1.0) * Fraction(_fact(j3 + j1 - j2) * _fact(j3 - j1 + j2) * _fact(j1 + j2 - j3) * _fact(j3 + m3) * _fact(j3 - m3),
                                       _fact(j1 + j2 + j3 + 1) * _fact(j1 - m1) * _fact(j1 + m1) * _fact(j2 - m2) * _fact(j2 + m2))) ** 0.5
    tot = Fraction(0)
    for v in range(lo, hi + 1):
        tot += (-1) ** int(v + j2 + m2) * Fraction(_fact(j2 + j3 + m1 - v) * _fact(j1 - m1 + v),
                                                   _fact(v) * _fact(j3 - j1 + j2 - v) * _fact(j3 + m3 - v) * _fact(v + j1 - j2 - m3))
    return pre * float(tot)


def _q_real_to_complex(l):
    q = torch.zeros(2 * l + 1, 2 * l + 1, dtype=torch.complex128)
    s = 1 / math.sqrt(2)
    for m in range(1, l + 1):
        q[l - m, l + m], q[l - m, l - m] = s, -1j * s
        q[l + m, l + m], q[l + m, l - m] = (-1) ** m * s, 1j * (-1) ** m * s
    q[l, l] = 1
    return (-1j) ** l * q


@functools.lru_cache(maxsize=None)
def w3j(l1, l2, l3):
    """Real coupling tensor [2l1+1, 2l2+1, 2l3+1] (float64, unit Frobenius norm) in the basis where l = 1 is (x, y, z)."""
    c = torch.zeros(2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1, dtype=torch.complex128)
    for a in range(-l1, l1 + 1):
        for b in range(-l2, l2 + 1):
            if abs(a + b) <= l3:
                c[l1 + a, l2 + b, l3 + a + b] = _cg_complex(l1, a, l2, b, l3, a + b)
    c = torch.einsum("ij,kl,mn,ikn->jlm", _q_real_to_complex(l1), _q_real_to_complex(l2), torch.conj(_q_real_to_complex(l3).T), c)
    c = c.real.clone()
    return c / c.norm()


# ---- spherical harmonics, closed forms up to l = 3 ('component' normalisation, unit input) ------------------------------
def spherical_harmonics_cl(lmax, u):
    """u [E, 3] unit vectors (or zero) -> [E, (lmax+1)^2]; ATen elementwise glue (9-16 floats per edge)."""
    if lmax > 3:
        raise NotImplementedError("b200 engine: spherical harmonics are implemented up to max_ell = 3")
    x, y, z = u[:, 0], u[:, 1], u[:, 2]
    out = [torch.ones_like(x)]
    if lmax >= 1:
        s = math.sqrt(3.0)
        out += [s * x, s * y, s * z]
    if lmax >= 2:
        s3, s5 = math.sqrt(3.0), math.sqrt(5.0)
        x2z2 = x * x + z * z
        sh20, sh24 = s3 * x * z, (s3 / 2) * (z * z - x * x)
        out += [s5 * sh20, s5 * s3 * x * y, s5 * (y * y - 0.5 * x2z2), s5 * s3 * y * z, s5 * sh24]
    if lmax >= 3:
        s7 = math.sqrt(7.0)
        a = 4 * y * y - x2z2
        out += [s7 * math.sqrt(5 / 6) * (sh20 * z + sh24 * x), s7 * math.sqrt(5.0) * sh20 * y, s7 * math.sqrt(3 / 8) * a * x,
                s7 * 0.5 * y * (2 * y * y - 3 * x2z2), s7 * math.sqrt(3 / 8) * z * a, s7 * math.sqrt(5.0) * sh24 * y,
                s7 * math.sqrt(5 / 6) * (sh24 * z - sh20 * x)]
    return torch.stack(out, dim=1)


# ---- generalised coupling tensors of the symmetric contraction -----------------------------------------------------------
def _couple(ls, parities, nu):
    """All couplings of `nu` copies of the input irreps, as (l_out, p_out, tensor[2l_out+1, D, ..., D]) sorted like
    hydragnn/utils/model/mace_utils/tools/cg.py:22-91 sorts them (by (l, p), ties in generation order)."""
    offs, d = [], 0
    for l in ls:
        offs.append(d)
        d += 2 * l + 1
    level = []
    for l, p, o in zip(ls, parities, offs):
        t = torch.zeros(2 * l + 1, d, dtype=torch.float64)
        t[:, o:o + 2 * l + 1] = torch.eye(2 * l + 1, dtype=torch.float64)
        level.append((l, p, t))
    for depth in range(1, nu):
        nxt = []
        for ll, lp, lt in level:
            for l, p, o in zip(ls, parities, offs):
                for lo in range(abs(ll - l), ll + l + 1):
                    c = w3j(lo, ll, l) * math.sqrt(2 * lo + 1)
                    t = torch.einsum("jk,ijl->ikl", lt.flatten(1), c).reshape(2 * lo + 1, *([d] * depth), 2 * l + 1)
                    full = torch.zeros(2 * lo + 1, *([d] * (depth + 1)), dtype=torch.float64)
                    full[..., o:o + 2 * l + 1] = t
                    nxt.append((lo, lp * p, full))
        level = sorted(nxt, key=lambda e: (e[0], e[1]))
    return level


@functools.lru_cache(maxsize=None)
def u_matrix(lmax_in, l_out, nu):
    """U tensor of U_matrix_real (cg.py:94-136) for inputs 1x0e+1x1o+..(parity (-1)^l) and output (l_out, (-1)^l_out):
    shape [2 l_out + 1] + [D] * nu + [num_params] with the singleton degree axis squeezed, float64."""
    if nu > 3:
        raise NotImplementedError("b200 engine: correlation <= 3")
    ls = list(range(lmax_in + 1))
    sel = [t for l, p, t in _couple(ls, [(-1) ** l for l in ls], nu) if l == l_out and p == (-1) ** l_out]
    return torch.stack([t.squeeze() for t in sel], dim=-1)


@functools.lru_cache(maxsize=None)
def silu_second_moment_constant():
    """e3nn normalize2mom for SiLU: E[silu(z)^2]^(-1/2) over 1e6 float64 normal samples of a generator seeded with 0."""
    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
    return float(torch.nn.functional.silu(z).pow(2).mean().pow(-0.5))
