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
