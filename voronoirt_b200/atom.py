"""Host-side, one-time atom and per-site physics set-up: everything Λ_voronoi computes
BEFORE its loop (reference src/lambda_iteration.jl:216-247) and hands to the device engine as arrays.

Names and formulas follow the reference: HydrogenicLine (src/line.jl:14-72), test_atom (:232-247),
sample_λ_line (:259-305), sample_λ_boundfree (:316-345), const_unsold / const_quadratic_stark /
c4_traving (src/broadening.jl:7-61), LTE_populations (src/populations.jl:112-138), B_λ
(src/radiation.jl:17-19), destruction (src/line.jl:367-376).  The Transparency.jl helpers they call
(un-vendored, unpinned) are restated from their published closed forms.  Nothing here is on the hot path.
"""
import numpy as np

from ._abi import vrt_line

# CODATA 2018 (PhysicalConstants.CODATA2018; reference src/atmosphere.jl:1-8)
h = 6.62607015e-34
k_B = 1.380649e-23
c_0 = 299792458.0
e = 1.602176634e-19
m_e = 9.1093837015e-31
m_u = 1.66053906660e-27
m_p = 1.67262192369e-27
eps_0 = 8.8541878128e-12
a_0 = 5.29177210903e-11
R_inf = 10973731.568160

E_inf = R_inf * c_0 * h
hc = h * c_0
Ry = R_inf * c_0 * h
alpha_p = 4.5 * 4 * np.pi * eps_0 * a_0 ** 3
inv_4pieps0 = 1.0 / (4 * np.pi * eps_0)
mass_H = 1.008 * m_u
mass_He = 4.003 * m_u
abund_He = 10 ** 10.99 / 10 ** 12
Ryh = R_inf * c_0 * h / (1 + m_e / m_p)   # Transparency.jl's hydrogen Rydberg energy used by n_eff

BOOST = 2.0e9  # src/rates.jl:3


def test_atom(nλ_bb, nλ_bf):
    """src/line.jl:232-247 — (χu, χl, χ∞ [cm^-1], nλ_bb, nλ_bf, gu, gl, f_value, atom_weight [kg], Z)."""
    return 82258.211, 0.0, 109677.617, nλ_bb, nλ_bf, 8, 2, 4.162e-1, mass_H, 1


def transition_λ(χ1, χ2):
    """src/line.jl:354-356, energies in J -> nm."""
    return (h * c_0) / (χ2 - χ1) * 1e9


def sample_λ_line(nλ, λ0, qwing, qcore):
    """src/line.jl:259-305 (λ0 in nm)."""
    if nλ > 0 and nλ % 2 == 0:
        nλ += 1
    if 1 < nλ < 5:
        nλ = 5
    λ = np.empty(nλ)
    if nλ == 1:
        λ[0] = λ0
    elif nλ >= 5:
        vmicro_char = 2.5e3
        n = nλ / 2  # "Questionable" in the reference: a float (Q9)
        β = qwing / (2 * qcore)
        y = β + np.sqrt(β * β + (β - 1.0) * n + 2.0 - 3.0 * β)
        b = 2.0 * np.log(y) / (n - 1)
        a = qwing / (n - 2.0 + y * y)
        center = nλ // 2
        λ[center] = λ0
        q_to_λ = λ[center] * vmicro_char / c_0
        for w in range(1, nλ // 2 + 1):
            Δλ = a * (w + (np.exp(b * w) - 1.0)) * q_to_λ
            λ[center - w] = λ[center] - Δλ
            λ[center + w] = λ[center] + Δλ
    return λ


def sample_λ_boundfree(nλ, λ_min, χl, χ_inf):
    """src/line.jl:316-345."""
    λ_max = transition_λ(χl, χ_inf)
    λ = np.empty(nλ)
    if nλ == 1:
        λ[0] = λ_max
    elif nλ > 1:
        Δλ = (λ_max - λ_min) / (nλ - 1)
        λ[0] = λ_min
        for w in range(1, nλ):
            λ[w] = λ[w - 1] + Δλ
    return λ


def n_eff(energy_upper, energy_lower, Z):
    return Z * np.sqrt(Ryh / (energy_upper - energy_lower))


class HydrogenicLine:
    """src/line.jl:14-72.  Energies J, λ nm, ΔD nm (per site), Einstein B in m^3 J^-1."""

    def __init__(self, χu, χl, χ_inf, nλ_bb, nλ_bf, gu, gl, f_value, atom_weight, Z, temperature):
        # wavenumber_to_energy: cm^-1 -> J
        χu, χl, χ_inf = (h * c_0 * x * 100.0 for x in (χu, χl, χ_inf))
        assert χ_inf > χu > χl and gu > 0 and gl > 0 and f_value > 0 and atom_weight > 0 and Z >= 1
        λ0 = (h * c_0) / (χu - χl) * 1e9
        λbb = sample_λ_line(nλ_bb, λ0, 600.0, 15.0)
        nbb = len(λbb)
        λ1_min = transition_λ(χl, χ_inf) * (1 / 2.0) ** 2 + 0.001
        λ2_min = transition_λ(χl, χ_inf) * (2 / 2.0) ** 2 + 0.001
        λbf_l = sample_λ_boundfree(nλ_bf, λ1_min, χl, χ_inf)
        λbf_u = sample_λ_boundfree(nλ_bf, λ2_min, χu, χ_inf)
        self.λ = np.concatenate([λbb, λbf_l, λbf_u])
        self.λidx = [0, nbb, nbb + nλ_bf, nbb + 2 * nλ_bf]
        λ0_m = λ0 * 1e-9
        # calc_Aji / calc_Bji (Transparency.jl)
        self.Aji = 2 * np.pi * e ** 2 / (eps_0 * m_e * c_0) * (gl / gu) * f_value / λ0_m ** 2
        self.Bji = λ0_m ** 5 * self.Aji / (2 * h * c_0 ** 2)
        self.Bij = gu / gl * self.Bji
        self.λ0 = λ0
        self.χi, self.χj, self.χ_inf = χl, χu, χ_inf
        self.gi, self.gj = gl, gu
        self.atom_weight = atom_weight
        self.Z = Z
        temperature = np.asarray(temperature, dtype=np.float64)
        self.ΔD = λ0 / c_0 * np.sqrt(2 * k_B * temperature / atom_weight)  # doppler_width, nm
        assert np.all(np.isfinite(self.ΔD)) and np.all(self.ΔD >= 0)

    # ---- broadening constants (src/broadening.jl)
    def c4_traving(self):
        nu = n_eff(self.χ_inf, self.χj, self.Z)
        nl = n_eff(self.χ_inf, self.χi, self.Z)
        return (e ** 2 * inv_4pieps0 * a_0 ** 3 * 2 * np.pi / (h * 18 * self.Z ** 4)
                * ((nu * (5 * nu ** 2 + 1)) ** 2 - (nl * (5 * nl ** 2 + 1)) ** 2))

    def const_unsold(self, H_scaling=1, He_scaling=1):
        Δr = Ry ** 2 * (1 / (self.χ_inf - self.χj) ** 2 - 1 / (self.χ_inf - self.χi) ** 2)
        C6 = 2.5 * e ** 2 * alpha_p * inv_4pieps0 ** 2 * 2 * np.pi * (self.Z * a_0) ** 2 / h * Δr
        v_rel_const = 8 * k_B / (np.pi * self.atom_weight)
        v_rel_H = v_rel_const * (1 + self.atom_weight / mass_H)
        v_rel_He = v_rel_const * (1 + self.atom_weight / mass_He)
        return 8.08 * (H_scaling * v_rel_H ** 0.3 + He_scaling * abund_He * v_rel_He ** 0.3) * C6 ** 0.4

    def const_quadratic_stark(self, mean_atomic_weight=28 * m_u, scaling=1):
        C = 8 * k_B / (np.pi * self.atom_weight)
        Cm = (1 + self.atom_weight / m_e) ** (1 / 6) + (1 + self.atom_weight / mean_atomic_weight) ** (1 / 6)
        cStark23 = 11.37 * (scaling * self.c4_traving()) ** (2 / 3)
        return C ** (1 / 6) * cStark23 * Cm

    @staticmethod
    def const_linear_stark(n_upper=2, n_lower=1):
        """γ_linear_stark(n_e, u, l) = a1*0.6*(u²-l²)*n_e[cm^-3]^(2/3) (Transparency.jl / RH broad.c) as c*n_e[m^-3]^(2/3)."""
        a1 = 0.642 if n_upper - n_lower == 1 else 1.0
        return a1 * 0.6 * (n_upper ** 2 - n_lower ** 2) * 1e-4

    def as_struct(self):
        s = vrt_line()
        s.nlam = len(self.λ)
        for i in range(4):
            s.lidx[i] = self.λidx[i]
        s.lambda0 = self.λ0
        s.Aji, s.Bji, s.Bij = self.Aji, self.Bji, self.Bij
        s.chi_i, s.chi_j, s.chi_inf = self.χi, self.χj, self.χ_inf
        s.gi, s.gj, s.Z = self.gi, self.gj, self.Z
        s.atom_weight = self.atom_weight
        s.c_unsold = self.const_unsold()
        s.gamma_natural = 4.702e8           # src/broadening.jl:76
        s.c_linear_stark = self.const_linear_stark(2, 1)
        s.c_quadratic_stark = self.const_quadratic_stark()
        return s


def B_λ(λ_nm, T):
    """src/radiation.jl:17-19 in kW m^-2 nm^-1."""
    lam = np.asarray(λ_nm, dtype=np.float64) * 1e-9
    return 2 * h * c_0 ** 2 / lam ** 5 / (np.exp(h * c_0 / (lam * k_B * np.asarray(T, dtype=np.float64))) - 1) * 1e-12


def LTE_populations(line, temperature, electron_density, hydrogen_density):
    """src/populations.jl:112-138 -> (n, 3) Fortran-ordered array."""
    T = np.asarray(temperature, dtype=np.float64)
    ne = np.asarray(electron_density, dtype=np.float64)
    NH = np.asarray(hydrogen_density, dtype=np.float64)
    χ = [line.χi, line.χj, line.χ_inf]
    g = [line.gi, line.gj, 1]
    n_rel = np.ones((len(T), 3))
    saha_const = (k_B / h) * (2 * np.pi * m_e) / h
    saha_factor = 2 * ((saha_const * T) ** 1.5 / ne)
    for i in (1, 2):
        n_rel[:, i] = g[i] / g[0] * np.exp(-(χ[i] - χ[0]) / (k_B * T))
    n_rel[:, 2] *= saha_factor
    n_rel[:, 0] = 1 / n_rel.sum(axis=1)
    n_rel[:, 1:] *= n_rel[:, [0]]
    return np.asfortranarray(n_rel * NH[:, None])


def destruction(LTE_pops, C21, temperature, line):
    """src/line.jl:367-376 with the (2,1) collisional rate handed in (it is Transparency.jl's Johnson rate x BOOST)."""
    B_λ0 = B_λ(line.λ0, temperature) * 1e12  # SI
    return C21 / (C21 + line.Aji + line.Bji * B_λ0)
