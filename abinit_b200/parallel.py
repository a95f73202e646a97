"""Multi-GPU plumbing for the getghc path: one process per GPU, torch.distributed (NCCL over NVLink; gloo in the
CPU tests).  getghc shards over independent units and needs no data-path collective:

  * k-points/spins -> ranks round-robin, the reference's proc_distrb skip (src/79_seqpar_mpi/m_vtorho.F90:855-862);
  * band blocks of one k -> ranks, contiguous blocks like npband/bandpp (src/66_wfs/m_getghc.F90:2489-2520 slicing).

The only collective is the Gram-matrix sum of the block Rayleigh-Ritz, xgBlock_mpi_sum -> xmpi_sum
(src/45_xgTools/m_xg.F90:3636-3663, reached from xgBlock_gemm(..., comm=) :1969-1974): an in-place SUM allreduce of
subdim^2 FP64 numbers, issued here as one NCCL allreduce on the stream that produced the partial Gram matrices.
"""
from __future__ import annotations
from typing import List, Tuple


def band_block(nband: int, nranks: int, rank: int) -> Tuple[int, int]:
    """Contiguous [first, last) band range of `rank` (block sizes differ by at most one, larger blocks first)."""
    if nranks < 1 or not 0 <= rank < nranks:
        raise ValueError("bad rank/nranks")
    base, rem = divmod(nband, nranks)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def kpoint_owner(ikpt: int, isppol: int, nkpt: int, nranks: int) -> int:
    """Round-robin owner of (k-point, spin), the proc_distrb_cycle rule for npband = npfft = 1."""
    return (ikpt + isppol * nkpt) % nranks


def my_kpoints(nkpt: int, nsppol: int, nranks: int, rank: int) -> List[Tuple[int, int]]:
    return [(ik, isp) for isp in range(nsppol) for ik in range(nkpt) if kpoint_owner(ik, isp, nkpt, nranks) == rank]


def row_shard(npw: int, nranks: int, rank: int) -> Tuple[int, int]:
    """Plane-wave row range of `rank` in the LINALG (row-sharded, all bands) layout of xgTransposer
    (src/45_xgTools/m_xgTransposer.F90:640-900)."""
    return band_block(npw, nranks, rank)


def gram_allreduce(*mats, group=None):
    """xgBlock_mpi_sum: SUM-allreduce the partial Gram matrices (torch tensors, any device) in place.
    Several matrices (X^H A X and X^H B X) are flattened into one message to pay the launch latency once."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mats
    if len(mats) == 1:
        dist.all_reduce(mats[0], op=dist.ReduceOp.SUM, group=group)
        return mats
    flat = torch.cat([m.reshape(-1) for m in mats])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for m in mats:
        n = m.numel()
        m.copy_(flat[off:off + n].view_as(m)); off += n
    return mats


# ---------------------------------------------------------------------------------------------------------------------
# xgTransposer (src/45_xgTools/m_xgTransposer.F90:640-900, TRANS_ALL2ALL): between the band-sharded layout the filter works
# in (STATE_COLSROWS: my columns, all plane waves) and the row-sharded layout of the Rayleigh-Ritz (STATE_LINALG: all
# columns, my plane-wave rows).  One all-to-all per block over NCCL (NVLink / NVSwitch); blocks are float64 tensors
# (ncols, rows, 2) == the memory of cg(2, rows*ncols).
# ---------------------------------------------------------------------------------------------------------------------
def transpose_cols_to_rows(x_cols, nband: int, npw: int, group=None):
    """(my_ncols, npw, 2) on every rank -> (nband, my_nrows, 2) on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return x_cols
    f, l = band_block(nband, world, rank)
    assert x_cols.shape[0] == l - f and x_cols.shape[1] == npw
    rows = [row_shard(npw, world, q) for q in range(world)]
    send = torch.cat([x_cols[:, lo:hi].reshape(-1) for lo, hi in rows])
    in_splits = [(l - f) * (hi - lo) * 2 for lo, hi in rows]
    my_rows = rows[rank][1] - rows[rank][0]
    out_splits = [(band_block(nband, world, q)[1] - band_block(nband, world, q)[0]) * my_rows * 2 for q in range(world)]
    recv = torch.empty(sum(out_splits), dtype=x_cols.dtype, device=x_cols.device)
    dist.all_to_all_single(recv, send, out_splits, in_splits, group=group)
    return recv.view(nband, my_rows, 2)


def transpose_rows_to_cols(x_rows, nband: int, npw: int, group=None):
    """(nband, my_nrows, 2) on every rank -> (my_ncols, npw, 2) on every rank (inverse of transpose_cols_to_rows)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return x_rows
    f, l = band_block(nband, world, rank)
    rows = [row_shard(npw, world, q) for q in range(world)]
    my_rows = rows[rank][1] - rows[rank][0]
    assert x_rows.shape[0] == nband and x_rows.shape[1] == my_rows
    in_splits = [(band_block(nband, world, q)[1] - band_block(nband, world, q)[0]) * my_rows * 2 for q in range(world)]
    out_splits = [(l - f) * (hi - lo) * 2 for lo, hi in rows]
    recv = torch.empty(sum(out_splits), dtype=x_rows.dtype, device=x_rows.device)
    dist.all_to_all_single(recv, x_rows.contiguous().view(-1), out_splits, in_splits, group=group)
    out = torch.empty((l - f, npw, 2), dtype=x_rows.dtype, device=x_rows.device)
    off = 0
    for (lo, hi), n in zip(rows, out_splits):
        out[:, lo:hi] = recv[off:off + n].view(l - f, hi - lo, 2); off += n
    return out


_LIB_COMM = {"world": 0}


def init_library_comm(group=None):
    """Create the NCCL communicator owned by libabinit_b200.so for the ranks of `group` (ncclCommInitRank): rank 0 draws the
    ncclUniqueId, torch.distributed only carries its 128 bytes.  After this call the band-parallel drivers run entirely inside
    the library (abi_b200_chebfiwf2_paral_): NCCL on the library stream, no torch tensors on the data path."""
    import torch
    import torch.distributed as dist
    from . import xg
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if _LIB_COMM["world"] == world:
        return
    if world == 1:
        xg.comm_init_rank(bytes(128), 1, 0)
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t = torch.frombuffer(bytearray(xg.comm_get_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        xg.comm_init_rank(bytes(t.cpu().numpy().tobytes()), world, rank)
    _LIB_COMM["world"] = world


def chebfi_band_parallel_native(gs_hamk, cg_cols, nband: int, ecut: float, nline: int, bandpp: int = 128, group=None, occ=None,
                                tolwfr_diago: float = 1e-16, nbdbuf: int = 0, chebfi_oracle: int = 0, oracle_factor: float = 1e-2,
                                oracle_min_occ: float = 1e-8, enl_out=None):
    """chebfi_band_parallel through the library's own driver (abi_b200_chebfiwf2_paral_, include/abinit_b200.h): same arguments
    and results; the all-to-all re-layouts, the Gram allreduce (overlapped with the second Gram product) and the MAX reduction
    run on the library stream without host synchronisation in between."""
    import numpy as np
    import torch
    from . import xg
    init_library_comm(group)
    ncols, npw = int(cg_cols.shape[0]), int(cg_cols.shape[1])
    eig = np.zeros(nband); res = np.zeros(max(ncols, 1))
    torch.cuda.current_stream(cg_cols.device).synchronize()
    xg.chebfiwf2_paral(cg_cols, eig, occ, enl_out, res, gs_hamk, nband, ncols, npw, 1, float(tolwfr_diago), float(ecut), int(nline),
                       nbdbuf=nbdbuf, chebfi_oracle=chebfi_oracle, oracle_factor=oracle_factor, oracle_min_occ=oracle_min_occ, bandpp=bandpp)
    return eig, res[:ncols]


def lobpcg_band_parallel_native(gs_hamk, cg_cols, nband: int, nline: int, tolwfr_diago: float = 1e-30, bandpp: int = 128, group=None):
    """lobpcg_band_parallel through the library's own driver (abi_b200_lobpcgwf2_paral_): returns (eig[nband], resid[nband])."""
    import numpy as np
    import torch
    from . import xg
    init_library_comm(group)
    ncols, npw = int(cg_cols.shape[0]), int(cg_cols.shape[1])
    eig = np.zeros(nband); res = np.zeros(nband)
    torch.cuda.current_stream(cg_cols.device).synchronize()
    xg.lobpcgwf2_paral(cg_cols, eig, res, gs_hamk, nband, ncols, npw, 1, float(tolwfr_diago), int(nline), bandpp=bandpp)
    return eig, res


def chebfi_band_parallel(gs_hamk, cg_cols, nband: int, ecut: float, nline: int, bandpp: int = 128, group=None):
    """chebfi_run with paral_kgb=1, npband = world size (src/48_diago/m_chebfi2.F90:466-735): every rank filters its own
    band block (no communication), the Rayleigh quotient extrema are reduced over ranks (:606-611), the blocks are
    transposed to the row-sharded layout (:687-689), the Gram matrices are summed with ONE NCCL allreduce, every rank solves
    the replicated nband x nband eigenproblem and rotates its rows, and the blocks are transposed back.
    cg_cols: CUDA float64 tensor (my_ncols, npw, 2), updated in place.  Returns (eig[nband], resid[my_ncols]) host arrays."""
    import torch
    import torch.distributed as dist
    from . import xg, api
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    # the library runs on its own stream and returns synchronised (async off); torch ops issued here (clone, pack/unpack
    # of the all-to-all) run on torch's current stream: `sync()` orders torch -> library, the library's own
    # synchronisation orders library -> torch.
    api.set_async(False)
    sync = torch.cuda.current_stream(cg_cols.device).synchronize
    ncols, npw = int(cg_cols.shape[0]), int(cg_cols.shape[1])
    f, l = band_block(nband, world, rank)
    if ncols != l - f:
        raise ValueError("cg_cols does not hold this rank's band block")
    istwf_k = gs_hamk.istwf_k
    # PAW: B = S.  The filter needs S X for the Rayleigh quotients and S^-1 (apply_invovl) per band -- both local to the rank's
    # band block -- and the Rayleigh-Ritz step needs X^H S X as its B matrix, so the BX block is transposed, rotated and
    # transposed back with X and AX (m_chebfi2.F90:687-705 with paw = .true.).
    paw = bool(getattr(gs_hamk, "usepaw", 0))
    space = xg.SPACE_CR if istwf_k > 1 else xg.SPACE_C
    x = cg_cols.clone(); ax = torch.empty_like(x); xn = torch.empty_like(x); xp = torch.empty_like(x)
    bx = torch.empty_like(x) if paw else None
    sync()
    div, mx, mn = xg.chebfi_rq(gs_hamk, ncols, bandpp, x, ax, bx)
    if world > 1:
        t = torch.tensor([mx, -mn], dtype=torch.float64, device=x.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        mx, mn = float(t[0]), -float(t[1])
    lambda_minus, lambda_plus = mx, float(ecut)
    ndeg = min(xg.cheb_oracle1(mn, lambda_minus, lambda_plus, 1e-16, 40), int(nline))
    x, xn, xp = xg.chebfi_core(gs_hamk, ncols, bandpp, x, ax, bx, xn, xp, lambda_minus, lambda_plus, ndeg, div)
    del xn, xp
    # ---- Rayleigh-Ritz in the row-sharded layout
    xr = transpose_cols_to_rows(x, nband, npw, group); axr = transpose_cols_to_rows(ax, nband, npw, group)
    bxr = transpose_cols_to_rows(bx, nband, npw, group) if paw else None
    nrows = int(xr.shape[1])
    sync()
    me_g0 = (1 if (istwf_k == 2 and rank == 0 and gs_hamk.me_g0 == 1) else 0) if space == xg.SPACE_CR else -1
    xg.xg_colwise("zero_im_g0", space, nrows, nband, xr, nrows, me_g0=me_g0)
    xg.xg_colwise("zero_im_g0", space, nrows, nband, axr, nrows, me_g0=me_g0)
    if paw:
        xg.xg_colwise("zero_im_g0", space, nrows, nband, bxr, nrows, me_g0=me_g0)
    ldw = (nband + 1) & ~1
    sdt = torch.complex128 if space == xg.SPACE_C else torch.float64
    sub = torch.zeros((2, nband, ldw), dtype=sdt, device=x.device)
    sync()
    xg.xg_gram(space, nrows, nband, nband, xr, nrows, axr, nrows, sub[0], ldw, me_g0)
    xg.xg_gram(space, nrows, nband, nband, xr, nrows, bxr if paw else xr, nrows, sub[1], ldw, me_g0)
    if world > 1:
        dist.all_reduce(torch.view_as_real(sub) if sub.is_complex() else sub, op=dist.ReduceOp.SUM, group=group)   # xgBlock_mpi_sum
    w = torch.empty(nband, dtype=torch.float64, device=x.device)
    sync()
    info = xg.xg_hegvd(xg.SPACE_C if space == xg.SPACE_C else xg.SPACE_R, nband, sub[0], ldw, sub[1], ldw, w)
    if info != 0:
        raise RuntimeError(f"chebfi: hegvd failed with info={info}")
    xg.xg_rotate(space, nrows, nband, nband, xr, nrows, sub[0], ldw)
    xg.xg_rotate(space, nrows, nband, nband, axr, nrows, sub[0], ldw)
    if paw:
        xg.xg_rotate(space, nrows, nband, nband, bxr, nrows, sub[0], ldw)
    x = transpose_rows_to_cols(xr, nband, npw, group); ax = transpose_rows_to_cols(axr, nband, npw, group)
    if paw:
        bx = transpose_rows_to_cols(bxr, nband, npw, group)
    del xr, axr, bxr
    # ---- residuals of my bands: |AX - eig BX|^2 (m_chebfi2.F90:709-716)
    me_g0_cols = (1 if (istwf_k == 2 and gs_hamk.me_g0 == 1) else 0) if space == xg.SPACE_CR else -1
    w_loc = w[f:l].contiguous()
    res = torch.empty(ncols, dtype=torch.float64, device=x.device)
    sync()
    xg.xg_colwise("cymax", space, npw, ncols, ax, npw, bx if paw else x, npw, ax, npw, da=w_loc)
    xg.xg_colwise("norm2", space, npw, ncols, ax, npw, out=res, me_g0=me_g0_cols)
    cg_cols.copy_(x)
    return w.cpu().numpy(), res.cpu().numpy()


def lobpcg_band_parallel(gs_hamk, cg_cols, nband: int, nline: int, kinpw, tolwfr_diago: float = 1e-30, bandpp: int = 128, group=None):
    """lobpcg_run with paral_kgb=1, npband = world size, one block of all bands (src/48_diago/m_lobpcg2.F90:340-765),
    norm-conserving or PAW (B = S: the [BX | BW | BP] blocks travel with the A blocks): getAX_BX on the rank's own band block
    (band-sharded layout), everything else -- B-orthonormalisation,
    X / XW / XWP Rayleigh-Ritz, residuals, preconditioner -- on the rank's plane-wave rows (row-sharded layout) with the Gram
    matrices summed over ranks by NCCL allreduce and the small dense problems solved redundantly on every rank; per LOBPCG
    iteration one all-to-all out (W) and one back (AW), as the reference's xgTransposer does.
    cg_cols: CUDA float64 tensor (my_ncols, npw, 2), updated in place.  Returns (eig[nband], resid[nband]) host arrays."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from . import xg, api
    paw = bool(gs_hamk.usepaw)      # PAW: B = S, the [BX | BW | BP] blocks are carried (and transposed) like the A blocks
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    api.set_async(False)
    dev = cg_cols.device
    sync = torch.cuda.current_stream(dev).synchronize
    ncols, npw = int(cg_cols.shape[0]), int(cg_cols.shape[1])
    f, l = band_block(nband, world, rank)
    if ncols != l - f:
        raise ValueError("cg_cols does not hold this rank's band block")
    n = nband
    istwf_k = gs_hamk.istwf_k
    space = xg.SPACE_CR if istwf_k > 1 else xg.SPACE_C
    sub_space = xg.SPACE_C if space == xg.SPACE_C else xg.SPACE_R
    sdt = torch.complex128 if space == xg.SPACE_C else torch.float64
    lo, hi = row_shard(npw, world, rank)
    nr = hi - lo
    me_g0 = (1 if (istwf_k == 2 and rank == 0 and gs_hamk.me_g0 == 1) else 0) if space == xg.SPACE_CR else -1
    # build_pcon (m_lobpcgwf.F90:316-334) on my rows
    kin = np.asarray(kinpw, dtype=np.float64)[lo:hi]
    big = kin > np.finfo(np.float64).max * 1e-11
    kk = np.where(big, 0.0, kin)
    num = 27 + kk * (18 + kk * (12 + 8 * kk))
    pcon = torch.from_numpy(np.where(big, 0.0, num / (num + 16 * kk ** 4))).to(dev)
    # [X | W | P] and [AX | AW | AP] on my rows; B blocks alias the X blocks (norm-conserving)
    xwp = torch.zeros((3 * n, nr, 2), dtype=torch.float64, device=dev)
    axwp = torch.zeros_like(xwp)
    bxwp = torch.zeros_like(xwp) if paw else xwp
    blocks = (xwp, axwp, bxwp) if paw else (xwp, axwp)
    X, W, AX, AW, BX, BW = xwp[:n], xwp[n:2 * n], axwp[:n], axwp[n:2 * n], bxwp[:n], bxwp[n:2 * n]

    def allsum(t):
        if world > 1:
            dist.all_reduce(torch.view_as_real(t) if t.is_complex() else t, op=dist.ReduceOp.SUM, group=group)

    def apply_h(src_rows, dst_rows, dst_b_rows=None):
        cols = transpose_rows_to_cols(src_rows.contiguous(), n, npw, group).contiguous()
        out = torch.empty_like(cols)
        outs = torch.empty_like(cols) if paw else None
        sync()
        api.getghc(-1, cols, None, out, outs, gs_hamk, None, None, None, ncols, sij_opt=1 if paw else 0)
        if space == xg.SPACE_CR:
            g0 = 1 if (istwf_k == 2 and gs_hamk.me_g0 == 1) else 0
            xg.xg_colwise("zero_im_g0", space, npw, ncols, out, npw, me_g0=g0)
            if paw:
                xg.xg_colwise("zero_im_g0", space, npw, ncols, outs, npw, me_g0=g0)
        dst_rows.copy_(transpose_cols_to_rows(out, n, npw, group))
        if paw:
            dst_b_rows.copy_(transpose_cols_to_rows(outs, n, npw, group))
        return cols

    def gram(a, b, na, nb, w):                       # w: column block view of a (.., ldw) tensor, written in place
        sync()
        xg.xg_gram(space, nr, na, nb, a, nr, b, nr, w, int(w.shape[1]), me_g0)

    def b_orthonormalize(m):
        ldw = (m + 1) & ~1
        buf = torch.zeros((m, ldw), dtype=sdt, device=dev)
        for blk in blocks:
            xg.xg_colwise("zero_im_g0", space, nr, m, blk, nr, me_g0=me_g0)
        gram(xwp, bxwp, m, m, buf)
        allsum(buf)
        sync()
        info = xg.xg_chol_inverse(sub_space, m, buf, ldw)
        if info != 0:
            return info
        for blk in blocks:
            xg.xg_gemm_nn(space, nr, m, m, blk, nr, buf, ldw, blk, nr, upper=True)
        return 0

    def rayleigh_ritz(nvar):
        sub = nvar * n
        ldw = (sub + 1) & ~1
        ab = torch.zeros((2, sub, ldw), dtype=sdt, device=dev)
        for blk in blocks:
            xg.xg_colwise("zero_im_g0", space, nr, sub, blk, nr, me_g0=me_g0)
        for v in range(nvar):
            gram(xwp, axwp[v * n:], (v + 1) * n, n, ab[0, v * n:])
            if nvar > 1:
                gram(xwp, bxwp[v * n:], (v + 1) * n, n, ab[1, v * n:])
        allsum(ab)
        w = torch.empty(sub, dtype=torch.float64, device=dev)
        sync()
        info = xg.xg_hegvd(sub_space, sub, ab[0], ldw, ab[1] if nvar > 1 else None, ldw, w)
        if info != 0:
            raise RuntimeError(f"lobpcg: sub-space eigenproblem failed (info={info})")
        vec = ab[0]                                   # eigenvectors: tensor row j = column j, first `sub` entries
        if nvar > 1:
            ldc1 = (sub - n + 1) & ~1
            c1 = torch.zeros((n, ldc1), dtype=sdt, device=dev)
            c1[:, :sub - n] = vec[:n, n:sub]
        sync()
        for blk in blocks:
            xg.xg_rotate(space, nr, n, n, blk, nr, vec, ldw)
            if nvar > 1:
                xg.xg_gemm_nn(space, nr, sub - n, n, blk[n:], nr, c1, ldc1, blk[2 * n:], nr)
                xg.xg_colwise("add", space, nr, n, blk, nr, blk[2 * n:], nr)
        return w

    X.copy_(transpose_cols_to_rows(cg_cols, n, npw, group))
    apply_h(X, AX, BX)
    b_orthonormalize(n)
    eig = rayleigh_ritz(1)[:n].clone()
    res = torch.zeros(n, dtype=torch.float64, device=dev)

    def residuals():
        sync()
        xg.xg_colwise("cymax", space, nr, n, W, nr, BX, nr, AX, nr, da=eig)      # W = AX - eig BX (BX = X when norm-conserving)
        xg.xg_colwise("norm2", space, nr, n, W, nr, out=res, me_g0=me_g0)
        allsum(res)
        xg.xg_colwise("apply_diag", space, nr, n, W, nr, da=pcon)
        r = res.cpu().numpy()
        return r, float(r.min()), float(r.max())
    compute_residu = True
    r = None
    for iline in range(1, nline + 1):
        r, min_res, max_res = residuals()
        if max_res < tolwfr_diago:
            compute_residu = False
            break
        apply_h(W, AW, BW)
        use_xw = iline == 1 or min_res < 1e-27
        if not use_xw and b_orthonormalize(3 * n) != 0:
            use_xw = True
        if use_xw:
            b_orthonormalize(2 * n)
            for blk in blocks:
                blk[2 * n:].zero_()
        eig = rayleigh_ritz(2 if use_xw else 3)[:n].clone()
    if compute_residu:
        r, _, _ = residuals()
    cg_cols.copy_(transpose_rows_to_cols(X.contiguous(), n, npw, group))
    return eig.cpu().numpy(), r
