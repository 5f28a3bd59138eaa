"""Parameter sets for the bond-order potentials.

Mirrors the public names and list layout of the reference's
src/python/atomistica/parameters.py (pair lists in PAIR_INDEX order,
src/macros.inc:123) and the built-in Fortran defaults
(tersoff_params.f90:85-131, kumagai_params.f90:96-118, brenner_params.f90:78-222).
The numbers are the published parameterisations cited in each "__ref__".
"""
import copy
from math import log, sqrt


def pair_index(i, j, maxval):
    """0-based PAIR_INDEX (src/macros.inc:123)"""
    return min(i + j * maxval, j + i * maxval) - min(i * (i + 1) // 2, j * (j + 1) // 2)


def mix(p, key, rule):
    nel = len(p['el'])
    for i in range(nel):
        for j in range(i + 1, nel):
            p[key][pair_index(i, j, nel)] = rule(p[key][pair_index(i, i, nel)], p[key][pair_index(j, j, nel)])


def mix_arithmetic(p, key):
    mix(p, key, lambda x, y: (x + y) / 2)


def mix_geometric(p, key):
    mix(p, key, lambda x, y: sqrt(x * y))


# --- Tersoff -----------------------------------------------------------------

Tersoff_PRB_39_5566_Si_C = {
    "__ref__": "Tersoff J., Phys. Rev. B 39, 5566 (1989)",
    "el": ["C", "Si"],
    "A": [1.3936e3, sqrt(1.3936e3 * 1.8308e3), 1.8308e3],
    "B": [3.4674e2, sqrt(3.4674e2 * 4.7118e2), 4.7118e2],
    "xi": [1.0, 0.9776e0, 1.0],
    "lambda": [3.4879e0, (3.4879e0 + 2.4799e0) / 2, 2.4799e0],
    "mu": [2.2119e0, (2.2119e0 + 1.7322e0) / 2, 1.7322e0],
    "omega": [1.0, 1.0, 1.0],
    "mubo": [0.0, 0.0, 0.0],
    "m": [1, 1, 1],
    "beta": [1.5724e-7, 1.1000e-6],
    "n": [7.2751e-1, 7.8734e-1],
    "c": [3.8049e4, 1.0039e5],
    "d": [4.3484e0, 1.6217e1],
    "h": [-5.7058e-1, -5.9825e-1],
    "r1": [1.80, sqrt(1.80 * 2.70), 2.70],
    "r2": [2.10, sqrt(2.10 * 3.00), 3.00],
}

Goumri_Said_ChemPhys_302_135_Al_N = {
    "__ref__": "Goumri-Said S., Kanoun M.B., Merad A.E., Merad G., Aourag H., Chem. Phys. 302, 135 (2004)",
    "el": ["Al", "N"],
    "r1": [3.20, 2.185, 1.60],
    "r2": [3.60, 2.485, 2.00],
    "A": [746.698, 3000.214, 636.814],
    "B": [40.451, 298.81, 511.76],
    "xi": [1.0, 1.0, 1.0],
    "lambda": [2.4647, 3.53051, 5.43673],
    "mu": [0.9683, 1.99995, 2.7],
    "beta": [1.094932, 5.2938e-3],
    "n": [6.085605, 1.33041],
    "c": [0.074836, 2.0312e4],
    "d": [19.569127, 20.312],
    "h": [-0.659266, -0.56239],
}

Matsunaga_Fisher_Matsubara_Jpn_J_Appl_Phys_39_48_B_C_N = {
    "__ref__": "Matsunaga K., Fisher C., Matsubara H., Jpn. J. Appl. Phys. 39, 48 (2000)",
    "el": ["C", "N", "B"],
    "A": [1.3936e3, -1.0, -1.0, 1.1e4, -1.0, 2.7702e2],
    "B": [3.4674e2, -1.0, -1.0, 2.1945e2, -1.0, 1.8349e2],
    "xi": [1.0, 0.9685, 1.0025, 1.0, 1.1593, 1.0],
    "lambda": [3.4879, -1.0, -1.0, 5.7708, -1.0, 1.9922],
    "mu": [2.2119, -1.0, -1.0, 2.5115, -1.0, 1.5856],
    "omega": [1.0, 0.6381, 1.0, 1.0, 1.0, 1.0],
    "mubo": [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
    "m": [1, 1, 1, 1, 1, 1],
    "r1": [1.80, -1.0, -1.0, 2.0, -1.0, 1.8],
    "r2": [2.10, -1.0, -1.0, 2.3, -1.0, 2.1],
    "beta": [1.5724e-7, 1.0562e-1, 1.6e-6],
    "n": [7.2751e-1, 12.4498, 3.9929],
    "c": [3.8049e4, 7.9934e4, 5.2629e-1],
    "d": [4.3484e0, 1.3432e2, 1.5870e-3],
    "h": [-5.7058e-1, -0.9973, 0.5],
}
for _k in ('A', 'B', 'r1', 'r2'):
    mix_geometric(Matsunaga_Fisher_Matsubara_Jpn_J_Appl_Phys_39_48_B_C_N, _k)
for _k in ('lambda', 'mu'):
    mix_arithmetic(Matsunaga_Fisher_Matsubara_Jpn_J_Appl_Phys_39_48_B_C_N, _k)

# --- Brenner / Erhart-Albe ---------------------------------------------------

Erhart_PRB_71_035211_SiC = {
    "__ref__": "Erhart P., Albe K., Phys. Rev. B 71, 035211 (2005)",
    "el": ["C", "Si"],
    "D0": [6.00, 4.36, 3.24],
    "r0": [1.4276, 1.79, 2.232],
    "S": [2.167, 1.847, 1.842],
    "beta": [2.0099, 1.6991, 1.4761],
    "gamma": [0.11233, 0.011877, 0.114354],
    "c": [181.910, 273987.0, 2.00494],
    "d": [6.28433, 180.314, 0.81472],
    "h": [0.5556, 0.68, 0.259],
    "mu": [0.0, 0.0, 0.0],
    "n": [1.0, 1.0, 1.0],
    "m": [1, 1, 1],
    "r1": [1.85, 2.20, 2.68],
    "r2": [2.15, 2.60, 2.96],
}

Albe_PRB_65_195124_PtC = {
    "__ref__": "Albe K., Nordlund K., Averback R. S., Phys. Rev. B 65, 195124 (2002)",
    "el": ["Pt", "C"],
    "D0": [3.683, 5.3, 6.0],
    "r0": [2.384, 1.84, 1.39],
    "S": [2.24297, 1.1965, 1.22],
    "beta": [1.64249, 1.836, 2.1],
    "gamma": [8.542e-4, 9.7e-3, 2.0813e-4],
    "c": [34.0, 1.23, 330.0],
    "d": [1.1, 0.36, 3.5],
    "h": [1.0, 1.0, 1.0],
    "mu": [1.335, 0.0, 0.0],
    "n": [1.0, 1.0, 1.0],
    "m": [1, 1, 1],
    "r1": [2.9, 2.5, 1.7],
    "r2": [3.3, 2.8, 2.0],
}

Henriksson_PRB_79_114107_FeC = dict(
    __ref__="Henriksson K.O.E., Nordlund K., Phys. Rev. B 79, 144107 (2009)",
    el=["Fe", "C"],
    D0=[1.5, 4.82645134, 6.0],
    r0=[2.29, 1.47736510, 1.39],
    S=[2.0693109, 1.43134755, 1.22],
    beta=[1.4, 1.63208170, 2.1],
    gamma=[0.0115751, 0.00205862, 0.00020813],
    c=[1.2898716, 8.95583221, 330.0],
    d=[0.3413219, 0.72062047, 3.5],
    h=[-0.26, 0.87099874, 1.0],
    mu=[0.0, 0.0, 0.0],
    n=[1.0, 1.0, 1.0],
    m=[1, 1, 1],
    r1=[2.95, 2.3, 1.70],
    r2=[3.35, 2.7, 2.00],
)

Kioseoglou_PSSb_245_1118_AlN = {
    "__ref__": "Kioseoglou J., Komninou Ph., Karakostas Th., Phys. Stat. Sol. (b) 245, 1118 (2008)",
    "el": ["N", "Al"],
    "D0": [9.9100, 3.3407, 1.5000],
    "r0": [1.1100, 1.8616, 2.4660],
    "S": [1.4922, 1.7269, 2.7876],
    "beta": [2.05945, 1.7219, 1.0949],
    "gamma": [0.76612, 1.1e-6, 0.3168],
    "c": [0.178493, 100390, 0.0748],
    "d": [0.20172, 16.2170, 19.5691],
    "h": [0.045238, 0.5980, 0.6593],
    "mu": [0.0, 0.0, 0.0],
    "n": [1.0, 0.7200, 6.0865],
    "m": [1, 1, 1],
    "r1": [2.00, 2.19, 3.40],
    "r2": [2.40, 2.49, 3.60],
}

Brenner_PRB_42_9458_C_I = {
    "__ref__": "Brenner D., Phys. Rev. B 42, 9458 (1990) [potential I]",
    "el": ["C"],
    "D0": [6.325], "r0": [1.315], "S": [1.29], "beta": [1.5], "gamma": [0.011304],
    "c": [19.0], "d": [2.5], "h": [1.0], "mu": [0.0], "n": [1.0 / (2 * 0.80469)], "m": [1],
    "r1": [1.70], "r2": [2.00],
}

Brenner_PRB_42_9458_C_II = {
    "__ref__": "Brenner D., Phys. Rev. B 42, 9458 (1990) [potential II]",
    "el": ["C"],
    "D0": [6.0], "r0": [1.39], "S": [1.22], "beta": [2.1], "gamma": [0.00020813],
    "c": [330.0], "d": [3.5], "h": [1.0], "mu": [0.0], "n": [1.0 / (2 * 0.5)], "m": [1],
    "r1": [1.70], "r2": [2.00],
}

# --- Kumagai -----------------------------------------------------------------

Kumagai_CompMaterSci_39_457_Si = {
    "__ref__": "Kumagai T., Izumi S., Hara S., Sakai S., Comp. Mater. Sci. 39, 457 (2007)",
    "el": ["Si"],
    "A": [3281.5905], "B": [121.00047], "lambda1": [3.2300135], "lambda2": [1.3457970],
    "eta": [1.0000000], "delta": [0.53298909], "alpha": [2.3890327], "beta": [1],
    "c1": [0.20173476], "c2": [730418.72], "c3": [1000000.0], "c4": [1.0000000], "c5": [26.000000],
    "h": [-0.36500000], "r1": [2.70], "r2": [3.30],
}

# --- screened variants (TersoffScr, BrennerScr, KumagaiScr): parameters.py:79-107, 153-170, 196-218,
#     409-420 of the reference.  r1/r2 become the inner cutoff, or1/or2 the outer (attractive /
#     repulsive) cutoff, bor1/bor2 the bond-order cutoff, Cmin/Cmax the screening bounds.

def _scr(base, **upd):
    out = copy.deepcopy(base)
    out.update(copy.deepcopy(upd))
    return out


Tersoff_PRB_39_5566_Si_C__Scr = _scr(
    Tersoff_PRB_39_5566_Si_C,
    m=[3, 3, 3],
    r1=[2.00, sqrt(2.00 * 2.50), 2.50], r2=[2.00 * 1.2, sqrt(2.00 * 2.50) * 1.2, 2.50 * 1.2],
    or1=[2.00, sqrt(2.00 * 3.00), 3.00], or2=[2.00 * 2.0, sqrt(2.00 * 3.00) * 2.0, 3.00 * 2.0],
    bor1=[2.00, sqrt(2.00 * 3.00), 3.00], bor2=[2.00 * 2.0, sqrt(2.00 * 3.00) * 2.0, 3.00 * 2.0],
    Cmin=[1.00, 1.00, 1.00], Cmax=[3.00, 3.00, 3.00])
# mubo is 1/dimer length
_p = Tersoff_PRB_39_5566_Si_C__Scr
_p['mubo'] = [(_p['lambda'][_i] - _p['mu'][_i]) / log((_p['lambda'][_i] * _p['A'][_i]) / (_p['mu'][_i] * _p['B'][_i]))
              for _i in range(3)]

Matsunaga_Fisher_Matsubara_Jpn_J_Appl_Phys_39_48_B_C_N__Scr = _scr(
    Matsunaga_Fisher_Matsubara_Jpn_J_Appl_Phys_39_48_B_C_N,
    m=[3, 3, 3, 3, 3, 3],
    r1=[2.00, -1.0, -1.0, 2.00, -1.0, 1.8], r2=[2.00 * 1.2, -1.0, -1.0, 2.00 * 1.2, -1.0, 1.8 * 1.2],
    or1=[2.00, -1.0, -1.0, 3.00, -1.0, 1.8], or2=[2.00 * 2.0, -1.0, -1.0, 3.00 * 2.0, -1.0, 1.8 * 2],
    bor1=[2.00, -1.0, -1.0, 3.00, -1.0, 1.8], bor2=[2.00 * 2.0, -1.0, -1.0, 3.00 * 2.0, -1.0, 1.8 * 2],
    Cmin=[1.00] * 6, Cmax=[3.00] * 6)
for _k in ('r1', 'r2', 'or1', 'or2', 'bor1', 'bor2'):
    mix_geometric(Matsunaga_Fisher_Matsubara_Jpn_J_Appl_Phys_39_48_B_C_N__Scr, _k)

Erhart_PRB_71_035211_SiC__Scr = _scr(
    Erhart_PRB_71_035211_SiC,
    mu=[1.0 / 1.4276, 1.0 / 1.79, 1.0 / 1.842], m=[3, 3, 3],
    r1=[2.00, 2.40, 2.50], r2=[2.00 * 1.2, 2.40 * 1.2, 2.50 * 1.2],
    or1=[2.00, 2.40, 3.00], or2=[2.00 * 2.0, 2.40 * 2.0, 3.00 * 2.0],
    bor1=[2.00, 2.40, 3.00], bor2=[2.00 * 2.0, 2.40 * 2.0, 3.00 * 2.0],
    Cmin=[1.00, 1.00, 1.00], Cmax=[3.00, 3.00, 3.00])

Kumagai_CompMaterSci_39_457_Si__Scr = _scr(
    Kumagai_CompMaterSci_39_457_Si,
    r1=[2.50], r2=[2.50 * 1.2], or1=[3.00], or2=[3.00 * 2.0], bor1=[3.00], bor2=[3.00 * 2.0],
    Cmin=[1.00], Cmax=[3.00])

SCR_KEYS = ('or1', 'or2', 'bor1', 'bor2', 'Cmin', 'Cmax')
SCR_DEFAULTS = dict(Tersoff=Tersoff_PRB_39_5566_Si_C__Scr, Kumagai=Kumagai_CompMaterSci_39_457_Si__Scr,
                    Brenner=Erhart_PRB_71_035211_SiC__Scr)


def complete_scr(kind, db):
    """like complete() for the screened classes: the object starts as the screened default
    database (tersoff_params.f90:105-131 under SCREENING) and supplied keys overwrite it"""
    base = copy.deepcopy(SCR_DEFAULTS[kind])
    out = copy.deepcopy(db) if db is not None else base
    nel = len(out['el'])
    npairs = nel * (nel + 1) // 2
    for k, v in base.items():
        if k not in out:
            if kind == 'Tersoff' and k in TERSOFF_FIELD_DEFAULTS:
                n = nel if k in ('beta', 'n', 'c', 'd', 'h') else npairs
                out[k] = [TERSOFF_FIELD_DEFAULTS[k]] * n
            else:
                out[k] = list(v)
    return out


def scr_cutoff(db):
    """interaction range the screened potentials request (default_bind_to_func.f90:44-66, 106-130):
    sqrt(C_dr_cut) * largest of the inner / outer / bond-order cutoffs, maximum over the pairs"""
    rmax = max(max(db['r2']), max(db['or2']), max(db['bor2']))
    return max(sqrt(c * c / (4 * (c - 1)) if c > 2.0 else 1.0) for c in db['Cmax']) * rmax

# --- Juslin (W-C-H) / Kuopanportti (Fe-C-H): Brenner form, non-symmetric pair index (nel**2
#     entries, PAIR_INDEX_NS), triplet-indexed alpha/omega/m (TRIPLET_INDEX_NS); parameters.py:277-330
#     of the reference, defaults of juslin_params.f90:76-101

_J_OMEGA = [1.0] * 16 + [2.94586, 4.54415] + [1.0] * 4 + [0.33946, 0.22006] + [1.0] * 3
_J_ALPHA_CH = [0.0] * 5 + [4.0, 0.0, 4.0, 4.0, 0.0, 0.0, 0.0, 0.0, 4.0, 4.0, 0.0, 4.0, 4.0]

Juslin_JAP_98_123520_WCH = {
    '__ref__': 'Juslin N. et al., J. Appl. Phys. 98, 123520 (2005)',
    'el': ['W', 'C', 'H'],
    'D0': [5.41861, 6.64, 2.748, 0.0, 6.0, 3.6422, 0.0, 3.642, 4.7509],
    'r0': [2.34095, 1.90547, 1.727, -1.0, 1.39, 1.1199, -1.0, 1.1199, 0.74144],
    'S': [1.92708, 2.96149, 1.2489, 0.0, 1.22, 1.69077, 0.0, 1.69077, 2.3432],
    'beta': [1.38528, 1.80370, 1.52328, 0.0, 2.1, 1.9583, 0.0, 1.9583, 1.9436],
    'gamma': [0.00188227, 0.072855, 0.0054, 0.0, 0.00020813, 0.00020813, 0.0, 12.33, 12.33],
    'c': [2.14969, 1.10304, 1.788, 0.0, 330.0, 330.0, 0.0, 0.0, 0.0],
    'd': [0.17126, 0.33018, 0.8255, 0.0, 3.5, 3.5, 0.0, 1.0, 1.0],
    'h': [-0.27780, 0.75107, 0.38912, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0],
    'n': [1.0, 1.0, 1.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0],
    'alpha': [0.45876, 0.0, 0.0, 0.45876, 0.0, 0.0, 0.45876, 0.0, 0.0] + _J_ALPHA_CH,
    'omega': list(_J_OMEGA),
    'm': [1] * 27,
    'r1': [3.20, 2.60, 2.68, 0.0, 1.70, 1.30, 0.0, 1.30, 1.10],
    'r2': [3.80, 3.00, 2.96, 0.0, 2.00, 1.80, 0.0, 1.80, 1.70],
}

Kuopanportti_CMS_111_525_FeCH = {
    '__ref__': 'Kuopanportti P. et al., Comp. Mat. Sci. 111, 525 (2016)',
    'el': ['Fe', 'C', 'H'],
    'D0': [1.5, 4.82645134, 1.630, 0.0, 6.0, 3.6422, 0.0, 3.642, 4.7509],
    'r0': [2.29, 1.47736510, 1.589, -1.0, 1.39, 1.1199, -1.0, 1.1199, 0.74144],
    'S': [2.0693, 1.43134755, 4.000, 0.0, 1.22, 1.69077, 0.0, 1.69077, 2.3432],
    'beta': [1.4, 1.63208170, 1.875, 0.0, 2.1, 1.9583, 0.0, 1.9583, 1.9436],
    'gamma': [0.01158, 0.00205862, 0.01332, 0.0, 0.00020813, 0.00020813, 0.0, 12.33, 12.33],
    'c': [1.2899, 8.95583221, 424.5, 0.0, 330.0, 330.0, 0.0, 0.0, 0.0],
    'd': [0.3413, 0.72062047, 7.282, 0.0, 3.5, 3.5, 0.0, 1.0, 1.0],
    'h': [-0.26, 0.87099874, -0.1091, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0],
    'n': [1.0, 1.0, 1.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0],
    'alpha': [0.0] * 9 + _J_ALPHA_CH,
    'omega': list(_J_OMEGA),
    'm': [1] * 27,
    'r1': [2.95, 2.30, 2.2974, 0.0, 1.70, 1.30, 0.0, 1.30, 1.10],
    'r2': [3.35, 2.70, 2.6966, 0.0, 2.00, 1.80, 0.0, 1.80, 1.70],
}

JUSLIN_PAIR_KEYS = ('D0', 'r0', 'S', 'beta', 'gamma', 'c', 'd', 'h', 'n', 'r1', 'r2')


def complete_juslin(db):
    """default database + the mirroring of BIND_TO_FUNC (juslin_module.f90:283-312): entries with
    r0 < 0 take the parameters of the transposed pair"""
    base = copy.deepcopy(Juslin_JAP_98_123520_WCH)
    out = copy.deepcopy(db) if db is not None else base
    for k, v in base.items():
        if k not in out:
            out[k] = list(v)
    nel = len(out['el'])
    for i in range(nel):
        for j in range(nel):
            a, b = j + i * nel, i + j * nel
            if out['r0'][a] < 0.0:
                for key in JUSLIN_PAIR_KEYS:
                    out[key][a] = out[key][b]
    return out

# JuslinScr.  The Fortran default database (juslin_params.f90:95-103, SCREENING branch) has the outer
# and bond-order cutoffs equal to the inner one and Cmin = 1, Cmax = 3; the Python module's
# Juslin_JAP_98_123520_WCH__Scr (parameters.py:297-307) repeats the r1 / r2 rows in Cmin / Cmax.
_J_R1 = [3.20, 2.60, 2.68, 0.0, 1.70, 1.30, 0.0, 1.30, 1.10]
_J_R2 = [3.80, 3.00, 2.96, 0.0, 2.00, 1.80, 0.0, 1.80, 1.70]
Juslin_WCH__Scr_fortran_default = copy.deepcopy(Juslin_JAP_98_123520_WCH)
Juslin_WCH__Scr_fortran_default.update(
    r1=list(_J_R1), r2=list(_J_R2), or1=list(_J_R1), or2=list(_J_R2), bor1=list(_J_R1), bor2=list(_J_R2),
    Cmin=[1.0, 1.0, 1.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0], Cmax=[3.0, 3.0, 3.0, 0.0, 3.0, 3.0, 0.0, 3.0, 3.0])
Juslin_JAP_98_123520_WCH__Scr = copy.deepcopy(Juslin_JAP_98_123520_WCH)
Juslin_JAP_98_123520_WCH__Scr.update(
    r1=list(_J_R1), r2=list(_J_R2), or1=list(_J_R1), or2=list(_J_R2), bor1=list(_J_R1), bor2=list(_J_R2),
    Cmin=list(_J_R1), Cmax=list(_J_R2))

# parameters.py:331-342 of the reference: Fe-C-H with the screening rows of JuslinScr
_K_R1 = [2.95, 2.30, 2.2974, 0.0, 1.70, 1.30, 0.0, 1.30, 1.10]
_K_R2 = [3.35, 2.70, 2.6966, 0.0, 2.00, 1.80, 0.0, 1.80, 1.70]
Kuopanportti_CMS_111_525_FeCH__Scr = copy.deepcopy(Kuopanportti_CMS_111_525_FeCH)
Kuopanportti_CMS_111_525_FeCH__Scr.update(
    r1=list(_K_R1), r2=list(_K_R2), or1=list(_K_R1), or2=list(_K_R2), bor1=list(_K_R1), bor2=list(_K_R2),
    Cmin=[1.0, 1.0, 1.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0], Cmax=[3.0, 3.0, 3.0, 0.0, 3.0, 3.0, 0.0, 3.0, 3.0],
    m=[1] * 27)


def complete_juslin_scr(db):
    """JuslinScr: default database + mirroring, which in the SCREENING build also covers
    or1 ... Cmax (juslin_module.f90:285-298)"""
    base = copy.deepcopy(Juslin_WCH__Scr_fortran_default)
    out = copy.deepcopy(db) if db is not None else base
    for k, v in base.items():
        if k not in out:
            out[k] = list(v)
    nel = len(out['el'])
    for i in range(nel):
        for j in range(nel):
            a, b = j + i * nel, i + j * nel
            if out['r0'][a] < 0.0:
                for key in JUSLIN_PAIR_KEYS + SCR_KEYS:
                    out[key][a] = out[key][b]
    return out


def juslin_scr_cutoff(db):
    """list cutoff requested by JuslinScr: sqrt(C_dr_cut) of the pair times the largest cutoff of the
    database (juslin_module.f90:379-395), maximum over the pairs in use"""
    m = max(max(db['r2'][k], db['or2'][k], db['bor2'][k]) for k in range(len(db['r2'])))
    x = 0.0
    for k in range(len(db['r2'])):
        cmax = db['Cmax'][k]
        if cmax > 1.0:
            x = max(x, (cmax * cmax / (4 * (cmax - 1))) ** 0.5)
    return x * m


# Fortran built-in defaults used when a key is not supplied
# (type initialisers in tersoff_params.f90:62-77, brenner/kumagai: the default db itself)
TERSOFF_FIELD_DEFAULTS = dict(A=1.0, B=1.0, xi=1.0, mu=1.0, omega=1.0, mubo=0.0, m=1, beta=1.0, n=1.0,
                              c=1.0, d=1.0, h=1.0, r1=1.0, r2=2.0)
TERSOFF_FIELD_DEFAULTS['lambda'] = 1.0

DEFAULTS = dict(Tersoff=Tersoff_PRB_39_5566_Si_C, Kumagai=Kumagai_CompMaterSci_39_457_Si,
                Brenner=Erhart_PRB_71_035211_SiC)


def complete(kind, db):
    """Fill keys missing from a user dict the way the Fortran side does: the object starts
    as the default database and only supplied properties are overwritten."""
    base = copy.deepcopy(DEFAULTS[kind])
    out = copy.deepcopy(db) if db is not None else base
    nel = len(out['el'])
    npairs = nel * (nel + 1) // 2
    for k, v in base.items():
        if k not in out:
            if kind == 'Tersoff' and k in TERSOFF_FIELD_DEFAULTS:
                n = nel if k in ('beta', 'n', 'c', 'd', 'h') else npairs
                out[k] = [TERSOFF_FIELD_DEFAULTS[k]] * n
            else:
                out[k] = list(v)
    return out
