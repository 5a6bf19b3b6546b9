"""A second, independent restatement (plain numpy/Python, written from the Fortran, sharing no code with oracle/) of the
routines the path spends its time in, checked against the C++ oracle.  The reference ships no expected outputs
(SURVEY F3), so what pins the oracle is agreement between independent transcriptions plus the invariants in
test_oracle_cpu.py.

  photoion_rates   code/radiation_photoionrates.f90:108-277 with set_tau_table_positions :282, read_table :310,
                   photo_lookuptable :331, heat_lookuptable :470, scale_int2/3 :787/:808
  cinterp          code/files_for_3D/column_density.f90:28-345, weightf :351-376
  do_chemistry     code/files_for_3D/evolve_point.F90:444-646 (global branch) with thermal (thermal.f90:22-174), coolin
                   (cooling_h.f90:40-71), cosmo_cool (cosmology.f90:207-234), prepare_doric_factors / coldens
                   (doric.f90:317-372); doric itself through the oracle's single-call hook

Only data is shared: the radiation tables (built by the oracle's rad_ini, itself checked in test_oracle_cpu.py) and the
band constants parsed from oracle/band_data.h (literal arrays extracted from radiation_sizes.f90)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from c2ray_b200 import synth
from common import O, frac_err, oracle_setup

F = lambda x: float(np.float32(x))  # a default-real literal of the reference
NB1, NB2, NB3, NUMTAU = 1, 26, 20, 2000
NFB = NB1 + NB2 + NB3


def band_constants():
    txt = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "band_data.h")).read()
    arr = {m.group(1): np.array([float(v) for v in re.findall(r"[-+0-9.eE]+", m.group(2))])
           for m in re.finditer(r"static const double (BD_\w+)\[\d+\] = \{(.*?)\};", txt, flags=re.S)}
    z1, z26 = np.zeros(1), np.zeros(26)
    bd = {"sHI": np.concatenate([[F(6.346e-18)], arr["BD_SIGMA_HI_B2"], arr["BD_SIGMA_HI_B3"]]),
          "sHeI": np.concatenate([z1, arr["BD_SIGMA_HEI_B2"], arr["BD_SIGMA_HEI_B3"]]),
          "sHeII": np.concatenate([z1, z26, arr["BD_SIGMA_HEII_B3"]])}
    for f in ("F1ION", "F2ION", "F1HEAT", "F2HEAT"):
        for sp in ("HI", "HEI", "HEII"):
            bd[f"{f}_{sp}"] = np.concatenate([z1, arr[f"BD_{f}_{sp}_B2"], arr[f"BD_{f}_{sp}_B3"]])
    assert all(v.shape == (NFB,) for v in bd.values())
    return bd


def table_position(tau):  # :282-306
    t = np.log10(np.maximum(1.0e-20, tau))
    odpos = np.minimum(float(NUMTAU), np.maximum(0.0, 1.0 + (t - (-20.0)) / ((4.0 - (-20.0)) / 2000.0)))
    ipos = odpos.astype(np.int64)
    return ipos, odpos - ipos, np.minimum(NUMTAU, ipos + 1)


def read_table(tab, pos, b, col):  # :310-327 ; b 1-based band (position), col 1-based table column
    ipos, res, ip1 = pos
    return tab[col - 1, ipos[b - 1]] + (tab[col - 1, ip1[b - 1]] - tab[col - 1, ipos[b - 1]]) * res[b - 1]


def photoion_rates_np(col6, vol, seds, i_state, bd, isothermal):
    """seds: list of (NFlux, lo, hi, photo_thick, photo_thin, heat_thick, heat_thin) with tables [column, 0:NumTau]."""
    in_HI, out_HI, in_HeI, out_HeI, in_HeII, out_HeII = col6
    cHI, cHeI, cHeII = out_HI - in_HI, out_HeI - in_HeI, out_HeII - in_HeII                      # :167-169
    tau_in = in_HI * bd["sHI"] + in_HeI * bd["sHeI"] + in_HeII * bd["sHeII"]                      # :172-176
    tau_out = out_HI * bd["sHI"] + out_HeI * bd["sHeI"] + out_HeII * bd["sHeII"]
    pin, pout = table_position(tau_in), table_position(tau_out)
    scHI, scHeI, scHeII = np.zeros(NFB + 1), np.zeros(NFB + 1), np.zeros(NFB + 1)                # 1-based
    for b in range(NB1 + 1, NFB + 1):
        s1, s2, s3 = bd["sHI"][b - 1], bd["sHeI"][b - 1], bd["sHeII"][b - 1]
        if b <= NB1 + NB2:                                                                         # scale_int2 :787
            f = 1.0 / (s1 * cHI + s2 * cHeI)
            scHI[b], scHeI[b] = s1 * cHI * f, s2 * cHeI * f
        else:                                                                                      # scale_int3 :808
            f = 1.0 / (s1 * cHI + s2 * cHeI + s3 * cHeII)
            scHI[b], scHeI[b], scHeII[b] = cHI * s1 * f, cHeI * s2 * f, cHeII * s3 * f
    phi = dict(HI=0.0, HeI=0.0, HeII=0.0, heat=0.0, pin=0.0, pout=0.0)
    for NFlux, lo, hi, thick, thin, _, _ in seds:                                                  # photo_lookuptable :331
        if not NFlux > 0.0:
            continue
        for b in range(lo, hi + 1):
            p_in = NFlux * read_table(thick, pin, b, b)
            phi["pin"] += p_in
            if abs(tau_out[b - 1] - tau_in[b - 1]) > F(1.0e-7):
                p_out = NFlux * read_table(thick, pout, b, b)
                p_all = p_in - p_out
            else:
                p_all = NFlux * (tau_out[b - 1] - tau_in[b - 1]) * read_table(thin, pin, b, b)
                p_out = p_in - p_all
            phi["pout"] += p_out
            if b <= NB1:
                phi["HI"] += p_all / vol
            else:
                phi["HI"] += scHI[b] * p_all / vol
                phi["HeI"] += scHeI[b] * p_all / vol
                if b > NB1 + NB2:
                    phi["HeII"] += scHeII[b] * p_all / vol
    if isothermal:
        return phi
    tcHI, tcHeI, tcHeII = cHI * bd["sHI"], cHeI * bd["sHeI"], cHeII * bd["sHeII"]                # :236-240
    CR1, bR1, dR1 = (0.3908, 0.0554, 1.0), (0.4092, 0.4614, 0.2663), (1.7592, 1.6660, 1.3163)     # :49-55
    CR2, aR2, bR2 = (0.6941, 0.0984, 3.9811), (0.2, 0.2, 0.4), (0.38, 0.38, 0.34)
    y1R = [CR1[i] * (1.0 - i_state ** bR1[i]) ** dR1[i] for i in range(3)]                         # :557-565
    y2R = [CR2[i] * i_state ** aR2[i] * (1.0 - i_state ** bR2[i]) ** 2 for i in range(3)]
    ion_freq_HI = F(0.241838e15) * F(13.598)                                                       # cgsphotoconstants.f90:31
    ion_freq_HeI = F(0.241838e15) * F(24.587)
    hplanck = 6.6260755e-27
    for NFlux, lo, hi, _, _, thick, thin in seds:                                                  # heat_lookuptable :470
        if not NFlux > 0.0:
            continue
        f_heat = f_ion_HI = f_ion_HeI = 0.0
        df_ion_HI = df_ion_HeI = 0.0                          # not reset per band in the reference (:553-554)
        for b in range(lo, hi + 1):
            thick_cell = abs(tau_out[b - 1] - tau_in[b - 1]) > F(1.0e-4)
            if b <= NB1:
                cols, tcs, scs = [b], [tcHI[b - 1]], [1.0]
            elif b <= NB1 + NB2:
                cols = [2 * b - NB1 - 1, 2 * b - NB1]
                tcs, scs = [tcHI[b - 1], tcHeI[b - 1]], [scHI[b], scHeI[b]]
            else:
                base = 3 * b - NB2 - NB1 * 2
                cols = [base - 2, base - 1, base]
                tcs, scs = [tcHI[b - 1], tcHeI[b - 1], tcHeII[b - 1]], [scHI[b], scHeI[b], scHeII[b]]
            ph = []
            for col, tc, sc in zip(cols, tcs, scs):
                if thick_cell:
                    h_in = NFlux * read_table(thick, pin, b, col)
                    h_out = NFlux * read_table(thick, pout, b, col)
                    ph.append(sc * (h_in - h_out) / vol)
                else:
                    ph.append(NFlux * tc * read_table(thin, pin, b, col) / vol)
            df_heat = sum(ph)
            if b > NB1:
                names = ("HI", "HEI", "HEII")[:len(ph)]
                fs = [sum(bd[f"{f}_{n}"][b - 1] * x for n, x in zip(names, ph)) for f in ("F1ION", "F2ION", "F1HEAT", "F2HEAT")]
                df_ion_HeI = y1R[1] * fs[0] - y2R[1] * fs[1]
                df_ion_HI = y1R[0] * fs[0] - y2R[0] * fs[1]
                df_heat = df_heat - y1R[2] * fs[2] + y2R[2] * fs[3]
            f_heat += df_heat; f_ion_HI += df_ion_HI; f_ion_HeI += df_ion_HeI
        phi["heat"] += f_heat
        phi["HI"] += f_ion_HI / (ion_freq_HI * hplanck)
        phi["HeI"] += f_ion_HeI / (ion_freq_HeI * hplanck)
    return phi


@pytest.mark.parametrize("iso,with_qpl", [(False, False), (False, True), (True, True)])
def test_photoion_rates_against_numpy_restatement(iso, with_qpl):
    p = synth.make_problem(3 if with_qpl else 1, n=8, isothermal=iso)
    oracle_setup(p)
    info = O.sed_info()
    bd = band_constants()
    tabs = lambda s: [O.table(s, k) if (k < 2 or not iso) else None for k in range(4)]
    rng = np.random.default_rng(21)
    n = 150
    lin = 10.0 ** rng.uniform(12, 22, (n, 3)); d = 10.0 ** rng.uniform(8, 21, (n, 3))
    d[:30] = 10.0 ** rng.uniform(5, 12, (30, 3))          # optically thin cells (below both tau limits)
    col6 = np.empty((n, 6)); col6[:, 0::2] = lin; col6[:, 1::2] = lin + d
    vol = 10.0 ** rng.uniform(62, 68, n); i_state = 10.0 ** rng.uniform(-8, 0, n) * 0.999
    nflux = [2.0e5, 0.0, 7.0e3 if with_qpl else 0.0]
    ref = O.photoion_rates_batch(col6, vol, nflux, i_state)
    seds = [(nflux[0], 1, info["bb"][1]) + tuple(tabs(0))]
    if with_qpl:
        seds.append((nflux[2], info["qpl"][0], info["qpl"][1]) + tuple(tabs(2)))
    for c in range(n):
        phi = photoion_rates_np(col6[c], vol[c], seds, i_state[c], bd, iso)
        got = np.array([phi["HI"], phi["HeI"], phi["HeII"], phi["heat"], phi["pin"], phi["pout"]])
        scale = np.abs(ref[c]).max()
        # same formulas in the same order up to libm pow/log10 differences and summation order of numpy's sum()
        assert np.all(np.abs(got - ref[c]) <= 1e-11 * np.abs(ref[c]) + 1e-13 * np.array([1, 1, 1, 1, scale, scale]) * np.abs(ref[c]).max()), (c, got, ref[c])


def cinterp_np(pos, src, mesh, cdh, cdhe0, cdhe1):
    """column_density.f90:28-345.  Arrays are [k,j,i]; positions 1-based and unwrapped."""
    sqrt3, sqrt2 = float(np.sqrt(np.float32(3.0))), float(np.sqrt(np.float32(2.0)))
    sig = (F(6.346e-18), F(7.430e-18), F(1.589e-18))                      # cgsphotoconstants.f90:25-29
    weightf = lambda cd, i: 1.0 / max(0.6, cd * sig[i])                   # :351-376
    i, j, k = (int(x) for x in pos); i0, j0, k0 = (int(x) for x in src)
    idel, jdel, kdel = i - i0, j - j0, k - k0
    sgn = lambda v: 1 if v >= 0 else -1                                   # sign(1,0) = +1
    sgni, sgnj, sgnk = sgn(idel), sgn(jdel), sgn(kdel)
    im, jm, km = i - sgni, j - sgnj, k - sgnk
    di, dj, dk = float(idel), float(jdel), float(kdel)
    w = lambda v, n: (v - 1) % n                                          # modulo(v-1,mesh)+1, 0-based here
    get = lambda a, ii, jj, kk: a[w(kk, mesh[2]), w(jj, mesh[1]), w(ii, mesh[0])]
    ia, ja, ka = abs(idel), abs(jdel), abs(kdel)
    if ka >= ja and ka >= ia:
        alam = (float(km - k0) + sgnk * 0.5) / dk
        xc, yc = alam * di + float(i0), alam * dj + float(j0)
        dx = 2.0 * abs(xc - (float(im) + 0.5 * sgni)); dy = 2.0 * abs(yc - (float(jm) + 0.5 * sgnj))
        s = [(1. - dx) * (1. - dy), (1. - dy) * dx, (1. - dx) * dy, dx * dy]
        corners = [(im, jm, km), (i, jm, km), (im, j, km), (i, j, km)]
        special = ka == 1 and (ia == 1 or ja == 1); both = ia == 1 and ja == 1
        path = np.sqrt((di * di + dj * dj) / (dk * dk) + 1.0)
    elif ja >= ia and ja >= ka:
        alam = (float(jm - j0) + sgnj * 0.5) / dj
        zc, xc = alam * dk + float(k0), alam * di + float(i0)
        dz = 2.0 * abs(zc - (float(km) + 0.5 * sgnk)); dx = 2.0 * abs(xc - (float(im) + 0.5 * sgni))
        s = [(1. - dx) * (1. - dz), (1. - dz) * dx, (1. - dx) * dz, dx * dz]
        corners = [(im, jm, km), (i, jm, km), (im, jm, k), (i, jm, k)]
        special = ja == 1 and (ia == 1 or ka == 1); both = ia == 1 and ka == 1
        path = np.sqrt((di * di + dk * dk) / (dj * dj) + 1.0)
    else:
        alam = (float(im - i0) + sgni * 0.5) / di
        zc, yc = alam * dk + float(k0), alam * dj + float(j0)
        dz = 2.0 * abs(zc - (float(km) + 0.5 * sgnk)); dy = 2.0 * abs(yc - (float(jm) + 0.5 * sgnj))
        s = [(1. - dz) * (1. - dy), (1. - dz) * dy, (1. - dy) * dz, dy * dz]
        corners = [(im, jm, km), (im, j, km), (im, jm, k), (im, j, k)]
        special = ia == 1 and (ja == 1 or ka == 1); both = ja == 1 and ka == 1
        path = np.sqrt((dj * dj + dk * dk) / (di * di) + 1.0)
    out = []
    for sp, a in enumerate((cdh, cdhe0, cdhe1)):
        c = [get(a, *q) for q in corners]
        ws = [s[q] * weightf(c[q], sp) for q in range(4)]
        v = (c[0] * ws[0] + c[1] * ws[1] + c[2] * ws[2] + c[3] * ws[3]) / (ws[0] + ws[1] + ws[2] + ws[3])
        if special:
            v = (sqrt3 if both else sqrt2) * v
        out.append(v)
    return out + [path]


def test_cinterp_against_numpy_restatement():
    rng = np.random.default_rng(8)
    n = 11
    mesh = np.array([n, n, n], dtype=np.int32)
    cdh = 10.0 ** rng.uniform(14, 20, (n, n, n))
    cdhe = 10.0 ** rng.uniform(13, 19, (2, n, n, n))
    src = np.array([3, 9, 6], dtype=np.int32)
    pos = [(i, j, k) for k in range(src[2] - 5, src[2] + 6) for j in range(src[1] - 5, src[1] + 6)
           for i in range(src[0] - 5, src[0] + 6) if (i, j, k) != tuple(src)]
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    worst = 0.0
    for q in pos:
        out = np.zeros(4)
        O.lib().orc_cinterp(P(mesh), P(cdh), P(cdhe[0]), P(cdhe[1]), P(np.array(q, dtype=np.int32)), P(src), P(out))
        ref = cinterp_np(q, src, mesh, cdh, cdhe[0], cdhe[1])
        worst = max(worst, float(np.max(np.abs(out - np.array(ref)) / np.abs(ref))))
    assert worst < 1e-14, worst   # the same IEEE operations in the same order: differences only from float() vs real()


# ------------------------------------------------------------------------------------------------------------------
# thermal (code/thermal.f90:22-174), coolin (cooling_h.f90:40-71), cosmo_cool (cosmology.f90:207-234) and the
# do_chemistry iteration (files_for_3D/evolve_point.F90:444-646, global branch) in plain Python; doric itself is taken
# from the oracle through its single-call hook (it has its own check against a 60-digit matrix exponential in
# test_oracle_cpu.py), everything around it -- the order of the calls, the partial averaging, the convergence test, the
# explicit thermal sub-cycling -- is restated here from the Fortran.
# ------------------------------------------------------------------------------------------------------------------
ABU_HE, ABU_C = F(0.074), F(7.1e-7)
K_B, GAMMA1 = 1.381e-16, 5.0 / 3.0 - 1.0
MINITEMP, REL_DENERGY = F(1.0), F(0.1)
MFC, MFA = F(1.0e-2), F(1.0e-8)


def electrondens_py(n, xh, xhe):  # tped.f90:75-84
    return n * (xh[1] * (1.0 - ABU_HE) + ABU_C + ABU_HE * (xhe[1] + 2.0 * xhe[2]))


def make_coolin():
    logT, *cols = O.read_cooling_table()
    lin = [10.0 ** np.asarray(c) for c in cols]                          # cooling_h.f90:163-169
    mintemp, dtemp = logT[0], logT[1] - logT[0]

    def coolin(n, ne, xh, xhe, T):
        tpos = (np.log10(T) - mintemp) / dtemp + 1.0
        it = min(801 - 1, max(1, int(tpos)))
        d = tpos - float(it)
        it1 = min(801, it + 1)
        L = [c[it - 1] + (c[it1 - 1] - c[it - 1]) * d for c in lin]
        return n * ne * ((xh[0] * L[0] + xh[1] * L[1]) * (1.0 - ABU_HE) + (xhe[0] * L[2] + xhe[1] * L[3] + xhe[2] * L[4]) * ABU_HE)
    return coolin


def thermal_py(dt, T, ne, n, ion, heat, coolin, cosmo):
    """ion: dict h, he, h_av, he_av, h_old, he_old.  Returns (end_temper, avg_temper or None when untouched)."""
    e_int = (n + electrondens_py(n, ion["h_old"], ion["he_old"])) * K_B * T / GAMMA1
    cosmo_rate = 0.0
    if cosmo is not None:
        zred, H0, Om = cosmo
        dzdt = H0 * (F(1.0) + zred) * np.sqrt(Om * (F(1.0) + zred) ** 3 + F(1.0) - Om)
        cosmo_rate = e_int * F(2.0) / (F(1.0) + zred) * dzdt
    if not T > MINITEMP:
        return T, None
    t_cum, avg, it, T0 = 0.0, 0.0, 0, T
    ne_av = electrondens_py(n, ion["h_av"], ion["he_av"])
    while True:
        it += 1
        cooling = coolin(n, ne, ion["h_av"], ion["he_av"], T) + cosmo_rate
        rate = max(1e-50, abs(cooling - heat))
        dt_ode = min(REL_DENERGY * (e_int / abs(rate)), dt - t_cum)
        e_int = e_int + dt_ode * (heat - cooling)
        avg = avg + F(0.5) * T * dt_ode
        T = (e_int * GAMMA1) / (K_B * (n + ne_av))
        avg = avg + F(0.5) * T * dt_ode
        if T < MINITEMP:
            e_int = (n + ne_av) * K_B * MINITEMP                                # :141, no /gamma1
            T = MINITEMP
        t_cum += dt_ode
        if t_cum >= dt or abs(t_cum - dt) < F(1e-6) * dt or it > 10000:
            break
    avg = avg / dt if dt > 0.0 else T0
    T_end = (e_int * GAMMA1) / (K_B * (n + electrondens_py(n, ion["h"], ion["he"])))
    return T_end, avg


def do_chemistry_py(dt, n, ion15, phi4, T_avg_in, T_old, coolin, cosmo, isothermal, doric=None):
    """doric: callable(dt, de, n, ion15, phi3, fr4, T) -> ion15; default: the oracle's single-call hook."""
    doric = doric or O.doric
    ion = np.array(ion15, dtype=np.float64)
    sig = dict(H_heth=1.238e-18, H_heLya=9.907e-22, He_heLya=1.301e-20, He_he2=1.690780687052975e-18,
               H_he2=1.230695924714239e-19, HeI=F(7.430e-18), HeII=F(1.589e-18))

    def fracs():  # prepare_doric_factors(coldens(path=1,...)) doric.f90:317-372
        NH = ion[0] * n * 1.0 * (1.0 - ABU_HE); NHe0 = ion[2] * n * 1.0 * ABU_HE; NHe1 = ion[3] * n * 1.0 * ABU_HE
        a, b = NH * sig["H_heth"], NHe0 * sig["HeI"]
        c, d = NH * sig["H_heLya"], NHe0 * sig["He_heLya"]
        e, f, g = NH * sig["H_he2"], NHe0 * sig["He_he2"], NHe1 * sig["HeII"]
        return [a / (a + b), c / (c + d), g / (g + f + e), f / (g + f + e)]

    avg_temper, temper1 = T_avg_in, T_old
    temper0 = temper1
    nit = 0
    while True:
        nit += 1
        temper2 = temper1
        yh0, yhe0, yhe2 = ion[5], ion[7], ion[9]
        de = electrondens_py(n, ion[5:7], ion[7:10])
        ion = doric(dt, de, n, ion, phi4[:3], fracs(), avg_temper)
        de = electrondens_py(n, ion[5:7], ion[7:10])
        fr = fracs()
        old = ion.copy()
        ion = doric(dt, de, n, ion, phi4[:3], fr, avg_temper)
        for q in (0, 1, 2, 3, 4, 5, 7, 8):                       # h(0:1) he(0:2) h_av(0) he_av(0:1); h_av(1), he_av(2) keep pass 2
            ion[q] = (ion[q] + old[q]) / 2.0
        de = electrondens_py(n, ion[5:7], ion[7:10])
        temper1 = temper0
        if not isothermal:
            d = dict(h=ion[0:2], he=ion[2:5], h_av=ion[5:7], he_av=ion[7:10], h_old=ion[10:12], he_old=ion[12:15])
            temper1, av = thermal_py(dt, temper1, de, n, d, phi4[3], coolin, cosmo)
            if av is not None:
                avg_temper = av
        conv = ((abs((ion[5] - yh0) / ion[5]) < MFC or ion[5] < MFA) and (abs((ion[7] - yhe0) / ion[7]) < MFC or ion[7] < MFA) and
                (abs((ion[9] - yhe2) / ion[9]) < MFC or ion[9] < MFA) and abs((temper1 - temper2) / temper1) < MFC)
        if conv or nit > 400:
            break
    return ion, temper1, avg_temper, nit


@pytest.mark.parametrize("iso", [False, True])
def test_do_chemistry_and_thermal_against_python_restatement(iso):
    p = synth.make_problem(2, n=8, isothermal=iso)
    oracle_setup(p)
    coolin = make_coolin()
    cosmo = (p["zred"], p["H0"], p["Omega0"]) if p["cosmological"] else None
    q = synth.make_chemistry_problem(96, seed=9, isothermal=iso)
    n = 96
    rng = np.random.default_rng(4)
    x1 = 10.0 ** rng.uniform(-5, -0.01, n); a = 10.0 ** rng.uniform(-5, -0.3, n); b = a * 10.0 ** rng.uniform(-3, -0.2, n)
    h = np.stack([1 - x1, x1], 1); he = np.stack([1 - a - b, a, b], 1)
    ion15 = np.concatenate([h, he, h, he, h, he], axis=1)
    ndens = np.ravel(q["ndens"])[:n]
    phi4 = np.stack([np.ravel(q["phih"])[:n], np.ravel(q["phihe"][0])[:n], np.ravel(q["phihe"][1])[:n],
                     np.ravel(q["phiheat"])[:n] if not iso else np.zeros(n)], 1)
    T = np.float32(10.0 ** rng.uniform(3.5, 4.5, n)).astype(np.float64)
    T3 = np.stack([T, T, T], 1)
    ion_o, T_o, nit_o = O.chemistry_batch(q["dt"], ndens, ion15, phi4, T3)
    assert nit_o.max() > 1
    for c in range(n):
        ion, t1, tav, nit = do_chemistry_py(q["dt"], ndens[c], ion15[c], phi4[c], p["temper_val"] if iso else T3[c, 1],
                                            p["temper_val"] if iso else T3[c, 2], coolin, cosmo, iso)
        assert nit == nit_o[c], (c, nit, nit_o[c])
        # 1e-16 differences in the inputs handed to doric (summation order of the electron density) come back as
        # ~1e-11 absolute from its cancellations: the parity tolerance of tests/common.py applies here too
        assert frac_err(ion[:10], ion_o[c, :10]) < 1, (c, frac_err(ion[:10], ion_o[c, :10]))
        if not iso:
            assert abs(t1 / T_o[c, 0] - 1) < 1e-9 and abs(tav / T_o[c, 1] - 1) < 1e-9, (c, t1, tav, T_o[c])
