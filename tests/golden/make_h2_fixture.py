"""Generates tests/golden/h2_tbase1.npz: the pseudopotential-derived tables of the reference's tutorial test tbase1_1
(H2 molecule in a 10 Bohr box, ecut 10 Ha, Gamma point only => istwfk 2; tests/tutorial/Input/tbase1_1.abi) so that the
Gamma-point SCF pin tests need neither the reference tree nor the psp8 file at run time.  Run in the build container only:
    python tests/golden/make_h2_fixture.py
Reads /root/reference/tests/Pspdir/Psdj_nc_sr_04_pw_std_psp8/H.psp8 (data file, not copied into the repo)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import psp8, scf, gsphere as g

PSP = "/root/reference/tests/Pspdir/Psdj_nc_sr_04_pw_std_psp8/H.psp8"
rprimd = 10.0 * np.eye(3)                                            # acell 10 10 10 (tbase1_1.abi)
xred = np.array([[-0.07, 0.0, 0.0], [0.07, 0.0, 0.0]]).T             # xcart -0.7 / +0.7 Bohr along x
ecut = 10.0
gprimd, gmet, ucvol = g.metric(rprimd)
ngfft = g.getng(2.0, ecut, gmet, (0.0, 0.0, 0.0))
assert tuple(ngfft) == (30, 30, 30), ngfft                           # tbase1_1.abo:63
gsqcut, boxcut = scf.getcut(ecut, gmet, ngfft)
p = psp8.read_psp8(PSP)
assert p.fchrg == 0.0                                                # no model core charge for H
qg = psp8.qgrid(gsqcut)
epsatm, vlspl, q2vq = psp8.psp8lo(p, qg)
ffs = psp8.psp8nl(p, qg)
vpsp = scf.vpsp_r(ngfft, gmet, ucvol, [xred], [vlspl], gsqcut)
xccc3d = np.zeros_like(vpsp)
tabs = np.array([f.cs(qg) for f in ffs])
yps = np.array([[float(f.cs(qg[0], 1)), float(f.cs(qg[-1], 1))] for f in ffs])
out = os.path.join(ROOT, "tests", "golden", "h2_tbase1.npz")
np.savez_compressed(out, rprimd=rprimd, xred=xred, ecut=ecut, ngfft=np.array(ngfft), zion=p.zion, epsatm=epsatm, ekb=p.ekb,
                    indlmn=p.indlmn, qgrid=qg, ffspl_tab=tabs, ffspl_yp=yps, vpsp=vpsp, xccc3d=xccc3d, boxcut=boxcut)
print("wrote", out, os.path.getsize(out), "bytes; epsatm", epsatm, "boxcut", boxcut, "ucvol", ucvol)
