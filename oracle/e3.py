"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the parts of **e3nn == 0.5.1** (requirements-torch.txt:4) that the reference's MACE path calls.
e3nn is NOT under /root/reference and cannot be installed here, so this file restates its *published algorithms*
(Geiger & Smidt, "e3nn: Euclidean Neural Networks", arXiv:2207.09453, and the e3nn 0.5 documentation) and anchors on the
reference's own call sites:

  o3.Irreps / o3.Irrep        hydragnn/models/MACEStack.py:144-145,195-212,284-311; utils/model/irreps_tools.py:15-44,66-109
  o3.wigner_3j                hydragnn/utils/model/mace_utils/tools/cg.py:58
  o3.SphericalHarmonics       hydragnn/models/MACEStack.py:155-159
  o3.Linear                   mace_utils/modules/blocks.py:61,307,329,355,367; MACEStack.py:349
  o3.TensorProduct ("uvu")    mace_utils/modules/blocks.py:320-327
  nn.FullyConnectedNet        mace_utils/modules/blocks.py:344-349

PARITY UNPINNED (this file only; the reference's in-repo MACE code on top of it IS pinned, see oracle/mace.py): the
reference tests hold no value pins for MACE and e3nn cannot be run here, so the
conventions below (real basis, signs, normalisations, parameter order) are from the published algorithm, checked only
through properties: orthogonality of the Clebsch-Gordan tensors, their invariance under rotations, equivariance of the
spherical harmonics, and against sympy (SU(2) Clebsch-Gordan coefficients; real spherical harmonics Z_lm with the polar
axis on y, up to the Condon-Shortley sign pattern) -- tests/test_oracle_mace.py.
"""
import math
from fractions import Fraction

import torch


# ---------------------------------------------------------------------------------------------------------------
# Irrep / Irreps bookkeeping
# ---------------------------------------------------------------------------------------------------------------
class Irrep(tuple):
    """(l, p) with p = +1 ('e') or -1 ('o').  Ordering is the tuple ordering, as in e3nn (Irrep subclasses tuple)."""

    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                s = l.strip()
                return super().__new__(cls, (int(s[:-1]), {"e": 1, "o": -1}[s[-1]]))
            if isinstance(l, tuple):
                l, p = l
        assert p in (1, -1) and l >= 0
        return super().__new__(cls, (int(l), int(p)))

    @property
    def l(self):
        return self[0]

    @property
    def p(self):
        return self[1]

    @property
    def dim(self):
        return 2 * self[0] + 1

    def __mul__(self, other):
        other = Irrep(other)
        return [Irrep(l, self.p * other.p) for l in range(abs(self.l - other.l), self.l + other.l + 1)]

    def __repr__(self):
        return "%d%s" % (self.l, "e" if self.p == 1 else "o")


class MulIr(tuple):
    """One (mul, Irrep) entry; e3nn exposes ``.mul`` / ``.ir`` and prints as ``64x1o``."""

    def __new__(cls, mul, ir):
        return super().__new__(cls, (int(mul), Irrep(ir)))

    @property
    def mul(self):
        return self[0]

    @property
    def ir(self):
        return self[1]

    @property
    def dim(self):
        return self[0] * self[1].dim

    def __repr__(self):
        return "%dx%r" % (self[0], self[1])


class Irreps(tuple):
    """Tuple of (mul, Irrep).  Mirrors the subset of e3nn.o3.Irreps the reference uses."""

    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return irreps
        out = []
        if irreps is None:
            pass
        elif isinstance(irreps, MulIr):
            out.append(tuple(irreps))
        elif isinstance(irreps, Irrep):
            out.append((1, irreps))
        elif isinstance(irreps, str):
            if irreps.strip():
                for part in irreps.split("+"):
                    part = part.strip()
                    if "x" in part:
                        mul, ir = part.split("x")
                        out.append((int(mul), Irrep(ir)))
                    else:
                        out.append((1, Irrep(part)))
        else:
            for item in irreps:
                if isinstance(item, Irrep):
                    out.append((1, item))
                elif isinstance(item, str):
                    out.append((1, Irrep(item)))
                else:
                    mul, ir = item
                    out.append((int(mul), Irrep(ir)))
        return super().__new__(cls, [MulIr(m, ir) for m, ir in out])

    @staticmethod
    def spherical_harmonics(lmax, p=-1):
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    @property
    def dim(self):
        return sum(mul * ir.dim for mul, ir in self)

    @property
    def num_irreps(self):
        return sum(mul for mul, _ in self)

    @property
    def lmax(self):
        return max(ir.l for _, ir in self)

    def count(self, ir):
        ir = Irrep(ir)
        return sum(mul for mul, i in self if i == ir)

    def slices(self):
        out, i = [], 0
        for mul, ir in self:
            out.append(slice(i, i + mul * ir.dim))
            i += mul * ir.dim
        return out

    def sort(self):
        """-> (irreps, p, inv) as e3nn: stable sort by (ir, original index); p[i] = new position of entry i."""
        order = sorted((ir, i, mul) for i, (mul, ir) in enumerate(self))
        inv = tuple(i for _, i, _ in order)
        p = [0] * len(inv)
        for new, old in enumerate(inv):
            p[old] = new
        return Irreps([(mul, ir) for ir, _, mul in order]), tuple(p), inv

    def simplify(self):
        out = []
        for mul, ir in self:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + mul, ir)
            elif mul > 0:
                out.append((mul, ir))
        return Irreps(out)

    def __contains__(self, ir):
        try:
            ir = Irrep(ir)
        except Exception:
            return False
        return any(i == ir for _, i in self)

    def __add__(self, other):
        return Irreps(tuple(self) + tuple(Irreps(other)))

    def __mul__(self, n):
        return Irreps(tuple(self) * int(n))

    def __getitem__(self, i):
        x = tuple.__getitem__(self, i)
        return Irreps(x) if isinstance(i, slice) else x

    def __repr__(self):
        return "+".join("%dx%r" % (mul, ir) for mul, ir in self)


def create_irreps_string(n, ell):
    """hydragnn/utils/model/irreps_tools.py:105-109."""
    return " + ".join("%dx%d%s" % (n, l, "e" if l % 2 == 0 else "o") for l in range(ell + 1))


# ---------------------------------------------------------------------------------------------------------------
# Clebsch-Gordan / Wigner 3j in e3nn's real basis
# ---------------------------------------------------------------------------------------------------------------
def _f(n):
    return math.factorial(round(n))


def _su2_cg_coeff(j1, m1, j2, m2, j3, m3):
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))
    c = ((2.0 * j3 + 1.0) * Fraction(_f(j3 + j1 - j2) * _f(j3 - j1 + j2) * _f(j1 + j2 - j3) * _f(j3 + m3) * _f(j3 - m3),
                                     _f(j1 + j2 + j3 + 1) * _f(j1 - m1) * _f(j1 + m1) * _f(j2 - m2) * _f(j2 + m2))) ** 0.5
    s = 0
    for v in range(vmin, vmax + 1):
        s += (-1) ** int(v + j2 + m2) * Fraction(_f(j2 + j3 + m1 - v) * _f(j1 - m1 + v),
                                                 _f(v) * _f(j3 - j1 + j2 - v) * _f(j3 + m3 - v) * _f(v + j1 - j2 - m3))
    return float(c * s)


def _su2_cg(j1, j2, j3):
    mat = torch.zeros(2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1, dtype=torch.float64)
    if abs(j1 - j2) <= j3 <= j1 + j2:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                if abs(m1 + m2) <= j3:
                    mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _real_to_complex(l):
    """Change of basis real -> complex spherical harmonics, with the (-i)^l phase that makes the real CG real."""
    q = torch.zeros(2 * l + 1, 2 * l + 1, dtype=torch.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / 2 ** 0.5
        q[l + m, l - abs(m)] = -1j / 2 ** 0.5
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / 2 ** 0.5
        q[l + m, l - abs(m)] = 1j * (-1) ** m / 2 ** 0.5
    return (-1j) ** l * q


_W3J = {}


def wigner_3j(l1, l2, l3, dtype=torch.float64):
    """Real-basis Wigner 3j tensor [2l1+1, 2l2+1, 2l3+1], Frobenius norm 1 (zero if the triangle rule fails)."""
    key = (l1, l2, l3)
    if key not in _W3J:
        q1, q2, q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
        c = _su2_cg(l1, l2, l3).to(torch.complex128)
        c = torch.einsum("ij,kl,mn,ikn->jlm", q1, q2, torch.conj(q3.T), c)
        assert float(c.imag.abs().max()) < 1e-9
        c = c.real
        n = c.norm()
        _W3J[key] = c / n if float(n) > 0 else c
    return _W3J[key].to(dtype)


# ---------------------------------------------------------------------------------------------------------------
# Real spherical harmonics (e3nn basis: Y^1 = (x, y, z); Y^{l+1} from Y^l (x) Y^1; Y^l_0(e_y) > 0)
# ---------------------------------------------------------------------------------------------------------------
_SH_SCALE = {}


def _sh_norm(l):
    """constants c_l such that the recursion below gives |Y^l| = 1 on unit vectors."""
    if l not in _SH_SCALE:
        y = torch.tensor([[0.0, 1.0, 0.0]], dtype=torch.float64)
        cur = y
        for k in range(1, l):
            nxt = torch.einsum("ijk,ni,nj->nk", wigner_3j(1, k, k + 1), y, cur)
            scale = 1.0 / float(nxt[0, k + 1])        # m = 0 component at the pole is +1
            _SH_SCALE[k + 1] = scale
            cur = nxt * scale
        _SH_SCALE.setdefault(1, 1.0)
        _SH_SCALE.setdefault(0, 1.0)
    return _SH_SCALE[l]


def spherical_harmonics(lmax, vec, normalize=True, normalization="component"):
    """[..., 3] -> [..., (lmax+1)^2], e3nn o3.SphericalHarmonics(Irreps.spherical_harmonics(lmax), normalize, normalization)."""
    assert normalization in ("component", "norm", "integral")
    if normalize:
        vec = torch.nn.functional.normalize(vec, dim=-1)
    shp = vec.shape[:-1]
    v = vec.reshape(-1, 3)
    out = [torch.ones(v.shape[0], 1, dtype=v.dtype, device=v.device)]
    cur = v
    for l in range(1, lmax + 1):
        if l > 1:
            _sh_norm(l)
            cur = torch.einsum("ijk,ni,nj->nk", wigner_3j(1, l - 1, l, dtype=v.dtype), v, cur) * _SH_SCALE[l]
        out.append(cur)
    if normalization == "component":
        out = [y * math.sqrt(2 * l + 1) for l, y in enumerate(out)]
    elif normalization == "integral":
        out = [y * math.sqrt((2 * l + 1) / (4 * math.pi)) for l, y in enumerate(out)]
    return torch.cat(out, dim=-1).reshape(*shp, (lmax + 1) ** 2)


class SphericalHarmonics(torch.nn.Module):
    """o3.SphericalHarmonics(Irreps.spherical_harmonics(lmax), normalize, normalization) as a module."""

    def __init__(self, irreps_out, normalize, normalization="integral"):
        super().__init__()
        self.lmax, self.normalize, self.normalization = Irreps(irreps_out).lmax, normalize, normalization

    def forward(self, vec):
        return spherical_harmonics(self.lmax, vec, self.normalize, self.normalization)


# ---------------------------------------------------------------------------------------------------------------
# o3.Linear
# ---------------------------------------------------------------------------------------------------------------
class Linear(torch.nn.Module):
    """o3.Linear(irreps_in, irreps_out): one [mul_in, mul_out] weight per (i_in, i_out) pair of equal irreps, all
    ~ N(0, 1) in ONE flat `weight`; output block = sum_paths x W / sqrt(sum over the block's paths of mul_in)
    (path_normalization="element").  No biases (e3nn default)."""

    def __init__(self, irreps_in, irreps_out, **_e3nn_defaults):      # internal_weights / shared_weights: e3nn defaults
        super().__init__()
        self.irreps_in, self.irreps_out = Irreps(irreps_in), Irreps(irreps_out)
        self.paths = [(i, o) for i, (_, ir_i) in enumerate(self.irreps_in) for o, (_, ir_o) in enumerate(self.irreps_out)
                      if ir_i == ir_o]
        fan = {}
        for i, o in self.paths:
            fan[o] = fan.get(o, 0) + self.irreps_in[i][0]
        self.alpha = [1.0 / math.sqrt(fan[o]) for _, o in self.paths]
        self.weight_numel = sum(self.irreps_in[i][0] * self.irreps_out[o][0] for i, o in self.paths)
        self.weight = torch.nn.Parameter(torch.randn(self.weight_numel))

    def forward(self, x):
        sl_in, sl_out = self.irreps_in.slices(), self.irreps_out.slices()
        outs = [None] * len(self.irreps_out)
        off = 0
        for (i, o), a in zip(self.paths, self.alpha):
            mi, ir = self.irreps_in[i]
            mo = self.irreps_out[o][0]
            w = self.weight[off:off + mi * mo].reshape(mi, mo)
            off += mi * mo
            xi = x[:, sl_in[i]].reshape(-1, mi, ir.dim)
            y = torch.einsum("nui,uw->nwi", xi, w) * a
            outs[o] = y if outs[o] is None else outs[o] + y
        res = []
        for o, (mo, ir) in enumerate(self.irreps_out):
            res.append(outs[o].reshape(x.shape[0], mo * ir.dim) if outs[o] is not None
                       else x.new_zeros(x.shape[0], mo * ir.dim))
        return torch.cat(res, dim=1)


# ---------------------------------------------------------------------------------------------------------------
# o3.TensorProduct, "uvu" instructions, external per-edge weights
# ---------------------------------------------------------------------------------------------------------------
class TensorProductUVU(torch.nn.Module):
    """o3.TensorProduct(irreps1, irreps2, irreps_out, instructions=[(i1, i2, io, "uvu", True)], shared_weights=False,
    internal_weights=False): out_io[u] = c * sum_v w[u, v] sum_{m1 m2} C[m1, m2, m3] x1[u, m1] x2[v, m2], with
    c = sqrt((2 l_out + 1) / sum over instructions into io of mul_2)  (irrep_normalization="component",
    path_normalization="element", all variances 1)."""

    def __init__(self, irreps1, irreps2, irreps_out, instructions, shared_weights=False, internal_weights=False):
        super().__init__()
        assert not shared_weights and not internal_weights, "only the external-weight form the reference uses is restated"
        self.irreps1, self.irreps2, self.irreps_out = Irreps(irreps1), Irreps(irreps2), Irreps(irreps_out)
        self.instructions = [tuple(ins[:3]) for ins in instructions]
        fan = {}
        for _, i2, io in self.instructions:
            fan[io] = fan.get(io, 0) + self.irreps2[i2][0]
        self.coeff = [math.sqrt(self.irreps_out[io][1].dim / fan[io]) for _, _, io in self.instructions]
        self.weight_numel = sum(self.irreps1[i1][0] * self.irreps2[i2][0] for i1, i2, _ in self.instructions)

    def forward(self, x1, x2, weight):
        s1, s2 = self.irreps1.slices(), self.irreps2.slices()
        outs = [None] * len(self.irreps_out)
        off = 0
        for (i1, i2, io), c in zip(self.instructions, self.coeff):
            m1, ir1 = self.irreps1[i1]
            m2, ir2 = self.irreps2[i2]
            _, ir3 = self.irreps_out[io]
            w = weight[:, off:off + m1 * m2].reshape(-1, m1, m2)
            off += m1 * m2
            a = x1[:, s1[i1]].reshape(-1, m1, ir1.dim)
            b = x2[:, s2[i2]].reshape(-1, m2, ir2.dim)
            cg = wigner_3j(ir1.l, ir2.l, ir3.l, dtype=x1.dtype).to(x1.device)
            y = torch.einsum("ijk,eui,evj,euv->euk", cg, a, b, w) * c
            outs[io] = y if outs[io] is None else outs[io] + y
        res = []
        for io, (mo, ir) in enumerate(self.irreps_out):
            res.append(outs[io].reshape(x1.shape[0], mo * ir.dim) if outs[io] is not None
                       else x1.new_zeros(x1.shape[0], mo * ir.dim))
        return torch.cat(res, dim=1)


# ---------------------------------------------------------------------------------------------------------------
# nn.FullyConnectedNet
# ---------------------------------------------------------------------------------------------------------------
_ACT_CST = {}


def normalize2mom_const(act):
    """e3nn.math.normalize2mom: 1/sqrt(E[act(z)^2]), z ~ N(0,1) estimated from 1e6 float64 samples of a generator seeded
    with 0; a constant within 1e-4 of 1 is replaced by exactly 1."""
    if act not in _ACT_CST:
        gen = torch.Generator(device="cpu").manual_seed(0)
        z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
        cst = float(act(z).pow(2).mean().pow(-0.5))
        _ACT_CST[act] = 1.0 if abs(cst - 1) < 1e-4 else cst
    return _ACT_CST[act]


class _FCLayer(torch.nn.Module):
    def __init__(self, h_in, h_out, act):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(h_in, h_out))
        self.h_in, self.act = h_in, act
        self.cst = normalize2mom_const(act) if act is not None else 1.0

    def forward(self, x):
        x = x @ (self.weight / math.sqrt(self.h_in))
        return self.act(x) * self.cst if self.act is not None else x


class FullyConnectedNet(torch.nn.Sequential):
    """nn.FullyConnectedNet(hs, act): layers `layer0..`, weight [h_in, h_out] ~ N(0,1) used as W / sqrt(h_in); the
    activation (not after the last layer) is rescaled to unit second moment."""

    def __init__(self, hs, act):
        super().__init__()
        for i, (h1, h2) in enumerate(zip(hs, hs[1:])):
            self.add_module("layer%d" % i, _FCLayer(h1, h2, act if i < len(hs) - 2 else None))
