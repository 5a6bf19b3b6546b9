"""Synthetic inputs for the hot path (BASELINE.json configs 1-5, concretised in SURVEY.md 8d).

The reference builds these through its out-of-scope setup modules (nbody test.F90 / test4.F90, mat_ini_test.F90:210-262
uniform mean-density box, sourceprops_test.F90 source decks, grid.F90:37-100 cell sizes, cosmoparms.f90 WMAP3+ values).
Here they are generated directly: what the hot path sees is only the arrays/scalars returned by `make_problem`.

Array layout: Fortran `A(i,j,k,c)` == numpy C-order `A[c,k,j,i]` (i fastest, component slowest), 1-based srcpos.
"""
import numpy as np

# cosmoparms.f90 (WMAP3+), cgsconstants.f90, cgsastroconstants.f90, abundances.f90 -- double values of the
# reference's default-real literals where that matters for the inputs' plausibility only.
H_LITTLE = float(np.float32(0.7))
OMEGA0 = float(np.float32(0.27))
OMEGA_B = float(np.float32(0.044))
MPC = 1e6 * float(np.float32(3.086e18))
G_GRAV = 6.6732e-8
M_P = 1.672661e-24
PI_F = float(np.float32(3.141592654))
ABU_HE = float(np.float32(0.074))
MU = (1.0 - ABU_HE) + 4.0 * ABU_HE
H0_CGS = H_LITTLE * 100.0 * 1e5 / MPC
RHO_CRIT_0 = 3.0 * H0_CGS * H0_CGS / (8.0 * PI_F * G_GRAV)
YEAR = float(np.float32(3.15576e7))
EPSILON = 1.0e-20
EV2FR = float(np.float32(0.241838e15))
ION_FREQ_HEII = EV2FR * float(np.float32(54.416))


def mean_density(z):
    """mat_ini_test.F90:241  avg_dens=rho_crit_0*Omega_B/(mu*m_p)*(1+z)^3"""
    return RHO_CRIT_0 * OMEGA_B / (MU * M_P) * (1.0 + z) ** 3


def cell_size(boxsize_mpc_h, n, z):
    """grid.F90 (comoving Mpc/h -> cm) then cosmology.f90 cosmo_evol (proper)"""
    return boxsize_mpc_h * MPC / H_LITTLE / n / (1.0 + z)


def lognormal_field(n, sigma, seed, smooth_cells=2.0):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((n, n, n))
    k = np.fft.fftfreq(n) * 2 * np.pi
    kz, ky, kx = np.meshgrid(k, k, np.fft.rfftfreq(n) * 2 * np.pi, indexing="ij")
    filt = np.exp(-0.5 * (kx * kx + ky * ky + kz * kz) * smooth_cells ** 2)
    g = np.fft.irfftn(np.fft.rfftn(g) * filt, s=(n, n, n), axes=(0, 1, 2))
    g *= sigma / g.std()
    rho = np.exp(g - 0.5 * sigma * sigma)
    return rho / rho.mean()


def density_peaks(rho, count, min_sep=2):
    """Indices (k,j,i) of the `count` highest local maxima with periodic min separation."""
    from scipy.ndimage import maximum_filter
    mx = maximum_filter(rho, size=2 * min_sep + 1, mode="wrap")
    flat = np.flatnonzero((rho == mx).ravel())
    order = flat[np.argsort(-rho.ravel()[flat], kind="stable")][:count]
    if len(order) < count:  # fall back to plain top-N cells
        rest = np.argsort(-rho.ravel(), kind="stable")
        rest = rest[~np.isin(rest, order)][:count - len(order)]
        order = np.concatenate([order, rest])
    return np.stack(np.unravel_index(order, rho.shape), axis=1)


def neutral_state(n, T0, isothermal):
    """mat_ini_test.F90:198-202: xh=(1-eps, eps), xhe=(1-2eps, eps, eps)"""
    xh = np.empty((2, n, n, n))
    xhe = np.empty((3, n, n, n))
    xh[0] = 1.0 - EPSILON
    xh[1] = EPSILON
    xhe[0] = 1.0 - 2.0 * EPSILON
    xhe[1] = EPSILON
    xhe[2] = EPSILON
    T = np.full((3, n, n, n), T0, dtype=np.float32)
    return xh, xhe, T


def make_problem(config, n=None, num_src=None, isothermal=False, seed=None):
    """Return a dict describing one synthetic time step of BASELINE config 1..4 (optionally at reduced mesh n)."""
    if config == 1:
        n = n or 128
        z, box, T_eff, S_star, dt = 9.0, 10.0, 5.0e4, 1e48, 5e6 * YEAR
        subbox = 10
        ndens = np.full((n, n, n), mean_density(z))
        c = n // 2
        srcpos = np.array([[c, c, c]], dtype=np.int32)
        nflux = np.array([1e55 / S_star])
        nflux_q = None
        qpl = None
    elif config == 2:
        n = n or 128
        z, box, T_eff, S_star, dt = 8.85, 0.5, 1.0e5, 1e52, 0.05e6 * YEAR
        subbox = n  # c2ray_parameters_TEST4.f90:46 subboxsize=mesh(1)
        rho = lognormal_field(n, 1.0, 4 if seed is None else seed)
        ndens = rho * mean_density(z)
        ns = num_src or 16
        pk = density_peaks(rho, ns, min_sep=2)
        srcpos = (pk[:, ::-1] + 1).astype(np.int32)  # (i,j,k) 1-based
        mass = rho[pk[:, 0], pk[:, 1], pk[:, 2]]
        nflux = 1e53 * mass / mass.sum() / S_star
        nflux_q = None
        qpl = None
    elif config in (3, 4):
        n = n or (256 if config == 3 else 512)
        z, T_eff, S_star, dt = 9.0, 5.0e4, 1e48, 5e6 * YEAR
        box = 37.0 * n / 256.0
        subbox = 10
        rho = lognormal_field(n, 1.2, n if seed is None else seed)
        ndens = rho * mean_density(z)
        ns = num_src or (1000 if config == 3 else 10000)
        pk = density_peaks(rho, ns, min_sep=2)
        srcpos = (pk[:, ::-1] + 1).astype(np.int32)
        mass = rho[pk[:, 0], pk[:, 1], pk[:, 2]]
        total = 1e56 * (ns / 1000.0)
        nflux = total * mass / mass.sum() / S_star
        if config == 3:
            nflux_q = np.zeros(ns)
            nb = max(1, ns // 20)
            nflux_q[:nb] = 0.1 * nflux[:nb] * S_star / 1e48  # 0.1 x the BB photon rate, qpl_S_star = 1e48
            qpl = dict(index=1.8, minfreq=float(np.float32(0.3) * np.float32(1e3)) * EV2FR, maxfreq=ION_FREQ_HEII * 100.0,
                       S_star=1e48)
        else:
            nflux_q = None
            qpl = None
    else:
        raise ValueError(config)
    T0 = 1.0e4
    xh, xhe, T = neutral_state(n, T0, isothermal)
    dr = cell_size(box, n, z)
    return dict(config=config, mesh=np.array([n, n, n], dtype=np.int32), dr=np.array([dr, dr, dr]), vol=dr ** 3,
                ndens=np.ascontiguousarray(ndens), xh=xh, xhe=xhe, temperature_grid=T, srcpos=srcpos,
                NormFlux=np.ascontiguousarray(nflux), NormFluxQPL=nflux_q, qpl=qpl, T_eff=T_eff, S_star=S_star, dt=dt,
                zred=z, H0=H0_CGS, Omega0=OMEGA0, subboxsize=subbox, max_subbox=1150, isothermal=isothermal,
                temper_val=T0, clumping=1.0, cosmological=True)


def make_chemistry_problem(ncells, seed=5, isothermal=False):
    """BASELINE config 5 inputs for `ncells` independent cells: lognormal ndens, Gamma_HI log-uniform in
    [1e-18,1e-10] s^-1 with half the cells at 0, Gamma_HeI=0.3 Gamma_HI, Gamma_HeII=0.01 Gamma_HI,
    heat = Gamma_HI * x_HI * n * (1-abu_he) * 8e-12 erg, neutral gas at 1e4 K, dt = 1e6 yr."""
    rng = np.random.default_rng(seed)
    z = 9.0
    ndens = mean_density(z) * np.exp(rng.standard_normal(ncells) * 1.2 - 0.72)
    g = 10.0 ** rng.uniform(-18.0, -10.0, ncells)
    g[rng.random(ncells) < 0.5] = 0.0
    phih = g
    phihe = np.stack([0.3 * g, 0.01 * g])
    phiheat = g * 1.0 * ndens * (1.0 - ABU_HE) * 8.0e-12
    xh = np.empty((2, ncells)); xhe = np.empty((3, ncells))
    xh[0] = 1.0 - EPSILON; xh[1] = EPSILON
    xhe[0] = 1.0 - 2 * EPSILON; xhe[1] = EPSILON; xhe[2] = EPSILON
    T = np.full((3, ncells), 1.0e4, dtype=np.float32)
    return dict(ncells=ncells, ndens=ndens, phih=phih, phihe=np.ascontiguousarray(phihe), phiheat=phiheat, xh=xh, xhe=xhe,
                temperature_grid=T, dt=1e6 * YEAR, zred=z, H0=H0_CGS, Omega0=OMEGA0, isothermal=isothermal,
                temper_val=1.0e4, clumping=1.0, cosmological=True)
