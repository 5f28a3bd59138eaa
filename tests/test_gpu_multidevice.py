"""The C ABI must work on a device other than the thread's current one (a DD rank whose host code,
or another library in the process, changed the current device between calls)."""
import numpy as np
import pytest

import oracle
from atomistica_b200 import native, structures as S

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def test_calculator_on_second_device_with_other_device_current(cu_setfl):
    torch = pytest.importorskip('torch')
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    a = S.fcc('Cu', 3.615, (6, 6, 6))
    a.rattle(0.05, seed=5)
    p = native.from_atoms(a, device=1)
    nl = native.Neighbors(200, device=1)
    pot = native.TabulatedAlloyEAM(setfl=cu_setfl, device=1)
    pot.bind_to(p, nl)
    torch.cuda.set_device(0)
    torch.zeros(1, device='cuda:0')          # make device 0 current for this thread
    e, f = pot.energy_and_forces(p, nl)[:2]
    eam = oracle.EAM(cu_setfl)
    onl = oracle.neighbor_list(a.positions, a.cell, a.pbc, eam.cutoff, 200)
    o = eam.energy_and_forces(a.positions, a.cell, onl, eam.eldb(a.symbols))
    assert abs(e - o['epot']) <= RTOL * abs(o['epot'])
    assert np.abs(f - o['f']).max() <= RTOL * max(np.abs(o['f']).max(), 1.0)
