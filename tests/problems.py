"""Seeded synthetic inputs for the parity tests; test infrastructure.
Uses the oracle's G-sphere/box restatement (pinned on the reference's own test outputs) to build shapes."""
from __future__ import annotations
import numpy as np
from oracle import gsphere as g


class Problem:
    pass


def make_problem(ecut, L, kpt=(0.0, 0.0, 0.0), istwf_k=1, ndat=4, seed=1234, ngfft=None, cplex=1,
                 natom_per_type=(2,), lmax_per_type=(1,), nproj_per_l=2, usepaw=0, filter_shell=True):
    """Cubic/orthorhombic cell; lmax -> (l, n) channels with 2l+1 m's each (useylm=1 ordering l, n, m)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rng_op = np.random.Generator(np.random.PCG64(seed + 1000))   # operator data: independent of npw/ndat
    p = Problem()
    rprimd = np.diag(L) if np.ndim(L) else np.eye(3) * float(L)
    p.gprimd, p.gmet, p.ucvol = g.metric(rprimd)
    p.kpt = np.asarray(kpt, dtype=np.float64); p.istwf_k = istwf_k; p.ndat = ndat; p.ecut = ecut
    p.ngfft = tuple(ngfft) if ngfft is not None else g.getng(2.0, ecut, p.gmet, p.kpt)
    p.kg = g.kpgsph(ecut, p.gmet, p.kpt, istwf_k)             # (3, npw) oracle convention
    p.kgF = np.ascontiguousarray(p.kg.T)                       # (npw, 3) == Fortran kg(3,npw) memory
    p.npw = p.kg.shape[1]
    p.me_g0 = 1
    n1, n2, n3 = p.ngfft
    kin = g.mkkin(ecut, 0.0, 1.0, p.gmet, p.kg, p.kpt)
    if filter_shell and p.npw > 20:
        # outermost ~0.5 % shell carries the huge*1e-10 sentinel (m_kg.F90:422-429) to exercise the filter
        thr = np.quantile(kin, 0.995)
        kin = np.where(kin >= thr, g.KIN_SENTINEL, kin)
    p.kinpw = np.ascontiguousarray(kin)
    kin_ok = np.where(kin < g.KIN_FILTER, kin, 0.0)
    c = (rng.standard_normal((ndat, p.npw)) + 1j * rng.standard_normal((ndat, p.npw))) / (1.0 + kin_ok)[None, :]
    if istwf_k == 2:
        c[:, 0] = c[:, 0].real
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    p.cwavef = np.ascontiguousarray(c)
    rng = rng_op
    # smooth local potential: 8 small-G cosines, amplitude 0.5 Ha, mean -0.3 Ha
    i3, i2, i1 = np.meshgrid(np.arange(n3), np.arange(n2), np.arange(n1), indexing="ij")
    v = np.full((n3, n2, n1), -0.3)
    for _ in range(8):
        gv = rng.integers(-2, 3, size=3); ph = rng.uniform(0, 2 * np.pi)
        v += 0.5 / 8 * np.cos(2 * np.pi * (gv[0] * i1 / n1 + gv[1] * i2 / n2 + gv[2] * i3 / n3) + ph)
    if cplex == 2:
        v = v + 1j * 0.1 * np.sin(2 * np.pi * (i1 / n1 - i3 / n3))
    p.cplex = cplex
    p.vlocal = np.ascontiguousarray(v)
    # --- non-local operator ---
    ntypat = len(natom_per_type)
    p.ntypat = ntypat; p.nattyp = np.array(natom_per_type, dtype=np.int32); p.natom = int(sum(natom_per_type))
    chans = []
    for t in range(ntypat):
        lst = []
        iln = 0
        for l in range(lmax_per_type[t] + 1):
            for n in range(nproj_per_l):
                iln += 1
                for m in range(-l, l + 1):
                    lst.append((l, m, n + 1, l * l + l + m + 1, iln, 1))
        chans.append(lst)
    p.lmnmax = max(len(x) for x in chans)
    indlmn = np.zeros((ntypat, p.lmnmax, 6), dtype=np.int32)
    for t in range(ntypat):
        for i, ch in enumerate(chans[t]):
            indlmn[t, i] = ch
    p.indlmn = indlmn
    p.lnmax = max(ch[4] for lst in chans for ch in lst)
    perm = rng.permutation(p.natom)
    p.atindx1 = (perm + 1).astype(np.int32)                    # sorted position -> original atom (1-based)
    p.xred = rng.uniform(0, 1, size=(3, p.natom))              # already type-sorted order
    kpg = p.kg.astype(float) + p.kpt[:, None]
    kpgnorm = np.sqrt(np.einsum("ip,ij,jp->p", kpg, p.gmet, kpg)) * 2 * np.pi
    ffnl = np.zeros((ntypat, p.lmnmax, 1, p.npw))
    for t in range(ntypat):
        for i, ch in enumerate(chans[t]):
            l, m, n = ch[0], ch[1], ch[2]
            ang = rng.standard_normal(3)
            # parity (-1)^l in k+G, so that P(-G) = conj(P(G)) at time-reversal-invariant k
            ylm = 1.0 if l == 0 else ((kpg.T @ ang) / np.maximum(np.linalg.norm(kpg, axis=0), 1e-30)) ** l * (1.0 + 0.1 * m)
            ffnl[t, i, 0] = (kpgnorm ** l) * np.exp(-(0.35 + 0.1 * n + 0.05 * t) * kpgnorm ** 2) * ylm
    p.ffnl = np.ascontiguousarray(ffnl)
    p.ph3d = np.ascontiguousarray(g.ph3d(p.kg, p.kpt, p.xred))  # (natom, npw) complex
    p.usepaw = usepaw
    if usepaw == 0:
        p.enl = np.ascontiguousarray(rng.standard_normal((ntypat, p.lnmax)))           # ekb(lnmax, ntypat)
        p.sij = None
    else:
        lmn2 = p.lmnmax * (p.lmnmax + 1) // 2
        p.enl = np.ascontiguousarray(0.5 * rng.standard_normal((p.natom, lmn2)))         # dij(lmn2, natom)
        s = np.zeros((ntypat, lmn2))
        for t in range(ntypat):
            a = 0.1 * rng.standard_normal((p.lmnmax, p.lmnmax)); a = a @ a.T
            for j in range(p.lmnmax):
                for i in range(j + 1):
                    s[t, j * (j + 1) // 2 + i] = a[i, j]
        p.sij = np.ascontiguousarray(s)
    return p


def rel_err_per_band(a, b):
    a = np.asarray(a).reshape(b.shape[0], -1); b = np.asarray(b).reshape(b.shape[0], -1)
    num = np.linalg.norm(a - b, axis=1)
    den = np.maximum(np.linalg.norm(b, axis=1), 1e-300)
    return float(np.max(num / den))
