"""GPU: the batched small-system path (abi_b200_getghc_batch_: concurrent lanes + CUDA graphs) against the oracle and against
the plain per-call path, for a set of k-points x 2 spins of a Fe-2-like PAW shape (BASELINE configs[2])."""
import numpy as np
import pytest
import abinit_b200 as ab
from abinit_b200 import api
from problems import make_problem, rel_err_per_band
from oracle import getghc as ogh, nonlop as onl

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

KPTS = [(0.125, 0.25, 0.375), (0.0, 0.0, 0.0), (-0.25, 0.5, 0.0), (0.5, 0.5, 0.5), (0.375, -0.125, 0.25)]


def _problems(ndat, usepaw):
    out = []
    for ik, k in enumerate(KPTS):
        for spin in range(2):
            # one potential per spin (load_spin), one sphere / projector set per k (load_k): 10 independent Hamiltonians
            p = make_problem(9.0, 5.42, k, 1, ndat=ndat, seed=50 + 7 * ik + spin, natom_per_type=(2,), lmax_per_type=(2,),
                             usepaw=usepaw, ngfft=(24, 24, 24))
            out.append(p)
    return out


@pytest.mark.parametrize("usepaw,use_graphs", [(1, True), (1, False), (0, True)])
def test_getghc_batch_matches_oracle_and_plain_calls(lib, usepaw, use_graphs):
    ndat = 12
    probs = _problems(ndat, usepaw)
    dev = torch.device("cuda", 0)
    hams, cws, ghcs, gscs, refs = [], [], [], [], []
    for p in probs:
        h = ab.Hamiltonian(p.ngfft, p.natom, p.ntypat, p.lmnmax, p.indlmn, p.nattyp, p.atindx1, usepaw, p.ucvol)
        h.load_spin(p.vlocal, 1); h.load_enl(p.enl, p.sij if usepaw else None); h.load_k(1, p.kgF, p.kinpw, p.ffnl, p.ph3d)
        hams.append(h)
        cws.append(torch.from_numpy(p.cwavef).to(dev)); ghcs.append(torch.zeros((ndat, p.npw), dtype=torch.complex128, device=dev))
        gscs.append(torch.zeros((ndat, p.npw), dtype=torch.complex128, device=dev))
        P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
        refs.append(ogh.getghc(p.cwavef, p.vlocal, p.kg, p.ngfft, p.kinpw, P, p.enl, p.sij if usepaw else None, p.indlmn, p.nattyp,
                               p.atindx1 - 1, usepaw=usepaw, sij_opt=1 if usepaw else 0))
    torch.cuda.synchronize()
    for rep in range(4):                      # eager, capture + launch, replay, replay
        for g in ghcs + gscs:
            g.zero_()
        torch.cuda.synchronize()
        api.getghc_batch(hams, cws, ghcs, gscs if usepaw else None, ndat=ndat, sij_opt=1 if usepaw else 0, use_graphs=use_graphs)
        torch.cuda.synchronize()
        for p, g, s, r in zip(probs, ghcs, gscs, refs):
            assert rel_err_per_band(g.cpu().numpy(), r[0]) < 1e-11, rep
            if usepaw:
                assert rel_err_per_band(s.cpu().numpy(), r[1]) < 1e-11, rep
    # identical to the plain call, bit for bit (same kernels, same launch geometry)
    plain = torch.zeros_like(ghcs[3])
    ab.getghc(-1, cws[3], None, plain, None if not usepaw else torch.zeros_like(plain), hams[3], None, None, None, ndat,
              sij_opt=1 if usepaw else 0)
    torch.cuda.synchronize()
    assert torch.equal(plain, ghcs[3])
    # a reload drops the graph of that handle: new potential -> new result, still equal to the oracle
    p = probs[0]
    v2 = np.ascontiguousarray(p.vlocal * 0.5)
    hams[0].load_spin(v2, 1)
    P = onl.prep_projectors(p.ffnl, p.ph3d, p.indlmn, p.nattyp, p.ucvol)
    r2 = ogh.getghc(p.cwavef, v2, p.kg, p.ngfft, p.kinpw, P, p.enl, p.sij if usepaw else None, p.indlmn, p.nattyp, p.atindx1 - 1,
                    usepaw=usepaw, sij_opt=1 if usepaw else 0)
    for rep in range(3):
        api.getghc_batch(hams, cws, ghcs, gscs if usepaw else None, ndat=ndat, sij_opt=1 if usepaw else 0, use_graphs=use_graphs)
        torch.cuda.synchronize()
        assert rel_err_per_band(ghcs[0].cpu().numpy(), r2[0]) < 1e-11
    api.graphs_clear()
    for h in hams:
        h.destroy()
