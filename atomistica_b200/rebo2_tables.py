"""Host-side construction of the REBO2 parameter block handed to atx_rebo2_create.

Mirrors what the reference does on the host at bind_to time:
  defaults            rebo2_type.f90:48-394 (Brenner et al. 2002, Tables 2, 3, 6, 7)
  default tables      rebo2_default_tables.f90:32-336 (Tables 4, 8, 9 of the paper)
  derived constants   rebo2_db.f90:81-303
  g(cos theta) spline rebo2_db.f90:405-524          -> atx_host_rebo2_g_spline
  P/F/T tables        table2d.f90:84-226, table3d.f90:85-284 -> atx_host_table{2,3}d_init
The linear solves run in the library's C++ host code (atx_host_gaussn).
"""
import ctypes as C
from math import exp

import numpy as np

from . import _lib as L


def _f32(x):
    # default-real (single precision) literals of the Fortran source, promoted to double
    return float(np.float32(x))


DEFAULTS = dict(
    cc_B1=12388.79197798, cc_B2=17.56740646509, cc_B3=30.71493208065,
    cc_beta1=4.7204523127, cc_beta2=1.4332132499, cc_beta3=1.3826912506,
    cc_Q=0.3134602960833, cc_A=10953.544162170, cc_alpha=4.7465390606595,
    ch_B1=32.3551866587, ch_beta1=1.43445805925, ch_Q=0.340775728, ch_A=149.94098723, ch_alpha=4.10254983,
    hh_B1=29.632593, hh_beta1=1.71589217, hh_Q=0.370471487045, hh_A=32.817355747, hh_alpha=3.536298648,
    hhh_lambda=4.0, cc_re=1.4, ch_re=1.09, hh_re=0.7415886997,
    cc_in_r1=1.70, cc_in_r2=2.00, ch_r1=1.30, ch_r2=1.80, hh_r1=1.10, hh_r2=1.70,
    dihedral=False,
)

G_THETA = np.array([-1.0, -1.0 / 2, -1.0 / 3, 0.0, 1.0 / 2, 1.0])
G_G1 = np.array([_f32(x) for x in (-0.01, 0.05280, 0.09733, 0.37545, 2.0014, 8.0)])
G_DG1 = np.array([_f32(x) for x in (0.10400, 0.17000, 0.40000, 0.0, 0.0, 0.0)])
G_D2G1 = np.array([_f32(x) for x in (0.00000, 0.37000, 1.98000, 0.0, 0.0, 0.0)])
G_G2 = np.array([_f32(x) for x in (0.0, 0.0, 0.09733, 0.271856, 0.416335, 1.0)])

SPGH = [270.467795364007301, 1549.701314596994564, 3781.927258631323866, 4582.337619544424228,
        2721.538161662818368, 630.658598136730774,
        16.956325544514659, -21.059084522755980, -102.394184748124742, -210.527926707779059,
        -229.759473570467513, -94.968528666251945,
        19.065031149937783, 2.017732531534021, -2.566444502991983, 3.291353893907436,
        -2.653536801884563, 0.837650930130006]
IGH = [3] * 18 + [2] * 4 + [1] * 3


def default_tables():
    Fcc = np.zeros((5, 5, 10)); dFdi = np.zeros_like(Fcc); dFdj = np.zeros_like(Fcc); dFdk = np.zeros_like(Fcc)
    Fcc[1, 1, 0] = 0.105000; Fcc[1, 1, 1] = -0.0041775; Fcc[1, 1, 2:9] = -0.0160856
    for k, v in enumerate((0.09444957, 0.02200000, 0.03970587, 0.03308822, 0.02647058, 0.01985293, 0.01323529,
                           0.00661764, 0.0)):
        Fcc[2, 2, k] = v
    Fcc[0, 1, 0] = 0.04338699; Fcc[0, 1, 1:9] = 0.0099172158
    Fcc[0, 2, 0] = 0.0493976637; Fcc[0, 2, 1] = -0.011942669; Fcc[0, 2, 2:9] = Fcc[0, 1, 1]
    Fcc[0, 3, 0:9] = -0.119798935; Fcc[0, 3, 2:9] = Fcc[0, 1, 1]
    Fcc[1, 2, 0] = 0.0096495698; Fcc[1, 2, 1] = 0.030; Fcc[1, 2, 2] = -0.0200
    Fcc[1, 2, 3] = -0.0233778774; Fcc[1, 2, 4] = -0.0267557548; Fcc[1, 2, 5:9] = -0.030133632
    Fcc[1, 3, 1:9] = -0.124836752
    Fcc[2, 3, 0:9] = -0.044709383
    for i in range(3, 8):
        Fcc[2, 2, i] = Fcc[2, 2, 2] + (i - 2) * (Fcc[2, 2, 8] - Fcc[2, 2, 2]) / 6
    for i in range(3, 5):
        Fcc[1, 2, i] = Fcc[1, 2, 2] + (i - 2) * (Fcc[1, 2, 5] - Fcc[1, 2, 2]) / 3
    dFdi[2, 1, 0] = -0.052500; dFdi[2, 1, 4:9] = -0.054376
    dFdi[2, 3, 1:9] = 0.062418
    dFdk[2, 2, 3:8] = -0.006618
    dFdk[1, 1, 1] = -0.060543; dFdk[1, 2, 3] = -0.020044; dFdk[1, 2, 4] = -0.020044
    for k in range(10):     # symmetrisation, rebo2_default_tables.f90:128-160
        for i in range(4):
            for j in range(i + 1, 4):
                x = Fcc[i, j, k] + Fcc[j, i, k]; Fcc[i, j, k] = Fcc[j, i, k] = x
                x = dFdi[i, j, k] + dFdj[j, i, k]; dFdi[i, j, k] = dFdj[j, i, k] = x
                x = dFdi[j, i, k] + dFdj[i, j, k]; dFdi[j, i, k] = dFdj[i, j, k] = x
                x = dFdk[i, j, k] + dFdk[j, i, k]; dFdk[i, j, k] = dFdk[j, i, k] = x
    Fch = np.zeros((5, 5, 10))
    Fch[0, 2, 4:9] = -0.0090477875161288110
    Fch[1, 3, 0:9] = -0.213; Fch[1, 2, 0:9] = -0.25; Fch[1, 1, 0:9] = -0.5
    for k in range(10):
        for i in range(3):
            for j in range(i + 1, 4):
                x = Fch[i, j, k] + Fch[j, i, k]; Fch[i, j, k] = Fch[j, i, k] = x
    Fhh = np.zeros((5, 5, 10)); Fhh[1, 1, 0] = 0.249831916
    Pcc = np.zeros((6, 6))
    Pcc[1, 1] = 0.003026697473481; Pcc[2, 0] = 0.007860700254745; Pcc[3, 0] = 0.016125364564267
    Pcc[1, 2] = 0.003179530830731; Pcc[2, 1] = 0.006326248241119
    Pch = np.zeros((6, 6))
    Pch[1, 0] = 0.2093367328250380; Pch[2, 0] = -0.064449615432525; Pch[3, 0] = -0.303927546346162
    Pch[0, 1] = 0.01; Pch[0, 2] = -0.1220421462782555; Pch[1, 1] = -0.1251234006287090
    Pch[2, 1] = -0.298905245783; Pch[0, 3] = -0.307584705066; Pch[1, 2] = -0.3005291724067579
    Tcc = np.zeros((5, 5, 10)); Tcc[2, 2, 0] = -0.070280085; Tcc[2, 2, 1:9] = -0.00809675
    return dict(Fcc=Fcc, dFdi=dFdi, dFdj=dFdj, dFdk=dFdk, Fch=Fch, Fhh=Fhh, Pcc=Pcc, Pch=Pch, Tcc=Tcc)


def _fort(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order='F'))


def _table3d(v, dx=None, dy=None, dz=None):
    coeff = np.zeros(144 * 64)
    keep = [_fort(v)] + [None if d is None else _fort(d) for d in (dx, dy, dz)]
    args = [None if k is None else L.dptr(k) for k in keep]
    L.check(L.lib().atx_host_table3d_init(4, 4, 9, *args, L.dptr(coeff)))
    return coeff


def _table2d(v):
    coeff = np.zeros(25 * 16)
    k = _fort(v)
    L.check(L.lib().atx_host_table2d_init(5, 5, L.dptr(k), None, None, L.dptr(coeff)))
    return coeff


def build_params(kwargs):
    d = dict(DEFAULTS)
    tabs = default_tables()
    for k, v in kwargs.items():
        if k in tabs:
            tabs[k] = np.asarray(v, dtype=np.float64)
        elif k in d:
            d[k] = v
        elif k == 'with_dihedral':
            d['dihedral'] = v
        else:
            raise RuntimeError("Unknown Rebo2 property '%s'." % k)
    p = L.AtxRebo2Params()
    for name in ('cc_B1', 'cc_B2', 'cc_B3', 'cc_beta1', 'cc_beta2', 'cc_beta3', 'cc_Q', 'cc_A', 'cc_alpha',
                 'ch_B1', 'ch_beta1', 'ch_Q', 'ch_A', 'ch_alpha', 'hh_B1', 'hh_beta1', 'hh_Q', 'hh_A', 'hh_alpha'):
        setattr(p, name, d[name])
    g1c = np.zeros(18); g2c = np.zeros(18)
    L.check(L.lib().atx_host_rebo2_g_spline(L.dptr(G_THETA), L.dptr(G_G1), L.dptr(G_DG1), L.dptr(G_D2G1),
                                            L.dptr(G_G2), L.dptr(g1c), L.dptr(g2c)))
    for i in range(6):
        p.cc_g_theta[i] = G_THETA[i]
    for i in range(18):
        p.cc_g1_coeff[i] = g1c[i]; p.cc_g2_coeff[i] = g2c[i]; p.spgh[i] = SPGH[i]
    for i in range(25):
        p.igh[i] = IGH[i]
    for t in (0, 2):     # rebo2_db.f90:147-152
        p.conpe[t] = -0.5
        p.conan[t] = 0.5 * p.conpe[t]
        p.conpf[t] = p.conpe[t] - 1.0
    al = d['hhh_lambda']
    p.conalp = al
    CC, CH, HH = 0, 2, 5
    ce = np.zeros((6, 6))     # rebo2_db.f90:158-168
    ce[CC, CC] = 1.0
    ce[CC, CH] = exp(al * (d['ch_re'] - d['cc_re']))
    ce[CC, HH] = exp(al * (d['hh_re'] - d['cc_re']))
    ce[CH, CC] = 1.0 / ce[CC, CH]
    ce[CH, CH] = 1.0
    ce[CH, HH] = exp(al * (d['hh_re'] - d['ch_re']))
    ce[HH, CC] = 1.0 / ce[CC, HH]
    ce[HH, CH] = 1.0 / ce[CH, HH]
    ce[HH, HH] = 1.0
    cef = ce.ravel(order='F')
    for i in range(36):
        p.conear[i] = cef[i]
    for idx, lo, hi in ((CC, 'cc_in_r1', 'cc_in_r2'), (CH, 'ch_r1', 'ch_r2'), (HH, 'hh_r1', 'hh_r2')):
        p.cut_in_l[idx] = d[lo]; p.cut_in_h[idx] = d[hi]; p.cut_in_h2[idx] = d[hi] ** 2
    p.with_dihedral = int(bool(d['dihedral']))
    keep = dict(
        Fcc=_table3d(tabs['Fcc'], tabs['dFdi'], tabs['dFdj'], tabs['dFdk']), Fch=_table3d(tabs['Fch']),
        Fhh=_table3d(tabs['Fhh']), Tcc=_table3d(tabs['Tcc']), Pcc=_table2d(tabs['Pcc']), Pch=_table2d(tabs['Pch']))
    for k, v in keep.items():
        setattr(p, k, L.dptr(v))
    return d, tabs, p, keep


# rebo2_type.f90:65-69, :204-213 (SCREENING branch): cutoffs of the C-C bonds of Rebo2Scr
SCR_DEFAULTS = dict(cc_in_r1=1.95, cc_in_r2=2.25, cc_ar_r1=2.179347, cc_ar_r2=2.819732,
                    cc_bo_r1=1.866344, cc_bo_r2=2.758372, cc_nc_r1=1.217335, cc_nc_r2=4.000000,
                    Cmin=1.00, Cmax=2.00)
SCR_KEYS = ('cc_ar_r1', 'cc_ar_r2', 'cc_bo_r1', 'cc_bo_r2', 'cc_nc_r1', 'cc_nc_r2', 'Cmin', 'Cmax')


def build_params_scr(kwargs):
    """parameter blocks of Rebo2Scr: (d, tabs, atx_rebo2_params, keep, atx_rebo2_screening, sd)"""
    sd = dict(SCR_DEFAULTS)
    base = {}
    for k, v in kwargs.items():
        if k in SCR_KEYS:
            sd[k] = v
        else:
            base[k] = v
    base.setdefault('cc_in_r1', sd['cc_in_r1'])
    base.setdefault('cc_in_r2', sd['cc_in_r2'])
    d, tabs, p, keep = build_params(base)
    sd['cc_in_r1'], sd['cc_in_r2'] = d['cc_in_r1'], d['cc_in_r2']
    q = L.AtxRebo2Screening()
    for k in SCR_KEYS:
        setattr(q, k, float(sd[k]))
    return d, tabs, p, keep, q, sd
