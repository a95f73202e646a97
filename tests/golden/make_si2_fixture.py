"""Generates tests/golden/si2_tbase3.npz: the pseudopotential-derived tables of the reference's tutorial test tbase3_1
(Si-2, ecut 12 Ha; tests/tutorial/Input/tbase3_1.abi) so that the SCF pin tests need neither the reference tree nor
the psp8 file at run time.  Run in the build container only:  python tests/golden/make_si2_fixture.py
Reads /root/reference/tests/Pspdir/Psdj_nc_sr_04_pw_std_psp8/Si.psp8 (data file, not copied into the repo)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import psp8, scf, gsphere as g

PSP = "/root/reference/tests/Pspdir/Psdj_nc_sr_04_pw_std_psp8/Si.psp8"
acell = 10.18
rprimd = acell * np.array([[0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]]).T      # columns = primitive vectors
xred = np.array([[0.0, 0.0, 0.0], [0.25, 0.25, 0.25]]).T
ecut = 12.0
gprimd, gmet, ucvol = g.metric(rprimd)
ngfft = g.getng(2.0, ecut, gmet, (0.0, 0.0, 0.0))
assert tuple(ngfft) == (24, 24, 24), ngfft
gsqcut, boxcut = scf.getcut(ecut, gmet, ngfft)
p = psp8.read_psp8(PSP)
qg = psp8.qgrid(gsqcut)
epsatm, vlspl, q2vq = psp8.psp8lo(p, qg)
ffs = psp8.psp8nl(p, qg)
xccc1d = psp8.psp8cc(p)
vpsp = scf.vpsp_r(ngfft, gmet, ucvol, [xred], [vlspl], gsqcut)
xccc3d = scf.mkcore(ngfft, rprimd, xred, xccc1d, p.rchrg)
tabs = np.array([f.cs(qg) for f in ffs])
yps = np.array([[float(f.cs(qg[0], 1)), float(f.cs(qg[-1], 1))] for f in ffs])
out = os.path.join(ROOT, "tests", "golden", "si2_tbase3.npz")
np.savez_compressed(out, rprimd=rprimd, xred=xred, ecut=ecut, ngfft=np.array(ngfft), zion=p.zion, epsatm=epsatm, ekb=p.ekb,
                    indlmn=p.indlmn, qgrid=qg, ffspl_tab=tabs, ffspl_yp=yps, vpsp=vpsp, xccc3d=xccc3d, boxcut=boxcut)
print("wrote", out, os.path.getsize(out), "bytes; epsatm", epsatm, "boxcut", boxcut, "ucvol", ucvol,
      "core electrons on grid", xccc3d.sum() * ucvol / xccc3d.size)
