"""Generates tests/golden/si2_tw90.npz: the pseudopotential-derived tables of dataset 1 of the reference's test
tests/tutoplugs/Input/tw90_1.abi (Si-2, acell 10.263, ecut 8 Ha, Gamma-centred 2x2x2 mesh: the three irreducible k-points
Gamma, (1/2,0,0), (1/2,1/2,0) are all time-reversal invariant => istwfk 2, 3, 7 storage is possible), so that the
time-reversal SCF pin tests need neither the reference tree nor the psp8 file at run time.  Run in the build container only:
    python tests/golden/make_si2_w90_fixture.py
Reads /root/reference/tests/Pspdir/Psdj_nc_sr_04_pw_std_psp8/Si.psp8 (data file, not copied into the repo)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import psp8, scf, gsphere as g

PSP = "/root/reference/tests/Pspdir/Psdj_nc_sr_04_pw_std_psp8/Si.psp8"
acell = 10.263
rprimd = acell * np.array([[0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]]).T      # columns = primitive vectors
xred = np.array([[0.0, 0.0, 0.0], [0.25, 0.25, 0.25]]).T
ecut = 8.0
gprimd, gmet, ucvol = g.metric(rprimd)
ngfft = g.getng(2.0, ecut, gmet, (0.0, 0.0, 0.0))
assert tuple(ngfft) == (20, 20, 20), ngfft                          # tw90_1.abo:107
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
out = os.path.join(ROOT, "tests", "golden", "si2_tw90.npz")
np.savez_compressed(out, rprimd=rprimd, xred=xred, ecut=ecut, ngfft=np.array(ngfft), zion=p.zion, epsatm=epsatm, ekb=p.ekb,
                    indlmn=p.indlmn, qgrid=qg, ffspl_tab=tabs, ffspl_yp=yps, vpsp=vpsp, xccc3d=xccc3d, boxcut=boxcut)
print("wrote", out, os.path.getsize(out), "bytes; epsatm", epsatm, "boxcut", boxcut, "ucvol", ucvol)
