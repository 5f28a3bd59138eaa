"""atomistica_b200 -- B200-native hot path of Atomistica behind the reference's plugin API.

    from atomistica_b200 import Tersoff, Kumagai, Brenner, Rebo2, TabulatedAlloyEAM

are drop-ins for `atomistica.Tersoff()` etc. (src/python/atomistica/aseinterface.py); the
low-level mirror of the `_atomistica` extension lives in `atomistica_b200.native`.
Importing this package does not touch the GPU; the shared library is loaded on first use and
there is no CPU fallback.
"""
from .aseinterface import (Atomistica, BornMayer, r6, Brenner, BrennerScr, DoubleHarmonic, Harmonic, Juslin, JuslinScr, Kumagai, LJCut, KumagaiScr, Rebo2, Rebo2Scr, TabulatedAlloyEAM, TabulatedEAM,  # noqa: F401
                           Tersoff, TersoffScr)
from .parameters import *  # noqa: F401,F403
