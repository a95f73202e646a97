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
    space = xg.SPACE_CR if istwf_k > 1 else xg.SPACE_C
    x = cg_cols.clone(); ax = torch.empty_like(x); xn = torch.empty_like(x); xp = torch.empty_like(x)
    sync()
    div, mx, mn = xg.chebfi_rq(gs_hamk, ncols, bandpp, x, ax)
    if world > 1:
        t = torch.tensor([mx, -mn], dtype=torch.float64, device=x.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        mx, mn = float(t[0]), -float(t[1])
    lambda_minus, lambda_plus = mx, float(ecut)
    ndeg = min(xg.cheb_oracle1(mn, lambda_minus, lambda_plus, 1e-16, 40), int(nline))
    x, xn, xp = xg.chebfi_core(gs_hamk, ncols, bandpp, x, ax, None, xn, xp, lambda_minus, lambda_plus, ndeg, div)
    del xn, xp
    # ---- Rayleigh-Ritz in the row-sharded layout
    xr = transpose_cols_to_rows(x, nband, npw, group); axr = transpose_cols_to_rows(ax, nband, npw, group)
    nrows = int(xr.shape[1])
    sync()
    me_g0 = (1 if (istwf_k == 2 and rank == 0 and gs_hamk.me_g0 == 1) else 0) if space == xg.SPACE_CR else -1
    xg.xg_colwise("zero_im_g0", space, nrows, nband, xr, nrows, me_g0=me_g0)
    xg.xg_colwise("zero_im_g0", space, nrows, nband, axr, nrows, me_g0=me_g0)
    ldw = (nband + 1) & ~1
    sdt = torch.complex128 if space == xg.SPACE_C else torch.float64
    sub = torch.zeros((2, nband, ldw), dtype=sdt, device=x.device)
    sync()
    xg.xg_gram(space, nrows, nband, nband, xr, nrows, axr, nrows, sub[0], ldw, me_g0)
    xg.xg_gram(space, nrows, nband, nband, xr, nrows, xr, nrows, sub[1], ldw, me_g0)
    if world > 1:
        dist.all_reduce(torch.view_as_real(sub) if sub.is_complex() else sub, op=dist.ReduceOp.SUM, group=group)   # xgBlock_mpi_sum
    w = torch.empty(nband, dtype=torch.float64, device=x.device)
    sync()
    info = xg.xg_hegvd(xg.SPACE_C if space == xg.SPACE_C else xg.SPACE_R, nband, sub[0], ldw, sub[1], ldw, w)
    if info != 0:
        raise RuntimeError(f"chebfi: hegvd failed with info={info}")
    xg.xg_rotate(space, nrows, nband, nband, xr, nrows, sub[0], ldw)
    xg.xg_rotate(space, nrows, nband, nband, axr, nrows, sub[0], ldw)
    x = transpose_rows_to_cols(xr, nband, npw, group); ax = transpose_rows_to_cols(axr, nband, npw, group)
    del xr, axr
    # ---- residuals of my bands: |AX - eig X|^2 (m_chebfi2.F90:709-716)
    me_g0_cols = (1 if (istwf_k == 2 and gs_hamk.me_g0 == 1) else 0) if space == xg.SPACE_CR else -1
    w_loc = w[f:l].contiguous()
    res = torch.empty(ncols, dtype=torch.float64, device=x.device)
    sync()
    xg.xg_colwise("cymax", space, npw, ncols, ax, npw, x, npw, ax, npw, da=w_loc)
    xg.xg_colwise("norm2", space, npw, ncols, ax, npw, out=res, me_g0=me_g0_cols)
    cg_cols.copy_(x)
    return w.cpu().numpy(), res.cpu().numpy()
