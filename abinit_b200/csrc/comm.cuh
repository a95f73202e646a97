// NCCL communicator of the library (band-parallel eigensolver drivers).  NCCL is dlopen'ed on first use (libnccl.so.2: the copy the
// host application already loaded, e.g. the one bundled with PyTorch, or the system one) -- not a link dependency.  The only
// collectives of the path are the ones the reference has (src/45_xgTools/m_xgTransposer.F90:640-900 all-to-all re-layout,
// xgBlock_mpi_sum src/45_xgTools/m_xg.F90:3636-3663 Gram allreduce, the MAX/MIN of the Rayleigh quotients m_chebfi2.F90:606-611);
// all of them are issued on the library stream, so no host synchronisation separates them from the kernels around them.
#pragma once
#include "common.cuh"
#include <cstddef>

namespace abi {

struct CommState { void* comm = nullptr; int nranks = 1, rank = 0; bool owned = false; };
CommState& comm_state();
void comm_get_unique_id(char* id128);                                 // ncclGetUniqueId (rank 0, then broadcast by the caller)
void comm_init(const char* id128, int nranks, int rank);              // ncclCommInitRank
void comm_adopt(void* nccl_comm, int nranks, int rank);               // use a communicator the caller created (ncclComm_t)
void comm_destroy();
void comm_allreduce(double* buf, size_t n, bool max_op, cudaStream_t st);      // in place, SUM or MAX; no-op on one rank
// personalised all-to-all in units of doubles: segment q of sbuf (soff[q], scnt[q]) goes to rank q, segment q of rbuf comes from it
void comm_alltoallv(const double* sbuf, const size_t* soff, const size_t* scnt, double* rbuf, const size_t* roff,
                    const size_t* rcnt, cudaStream_t st);


// ---- xgTransposer (src/45_xgTools/m_xgTransposer.F90:640-900, TRANS_ALL2ALL) on the library communicator ----
// Contiguous block distribution used for both the bands (STATE_COLSROWS) and the rows (STATE_LINALG): sizes differ by at most
// one, the larger blocks first (the rule of abinit_b200/parallel.py:band_block).
inline void block_range(long long n, int nranks, int rank, long long* first, long long* last) {
  const long long base = n / nranks, rem = n % nranks;
  *first = rank * base + (rank < rem ? rank : rem);
  *last = *first + base + (rank < rem ? 1 : 0);
}
// cols (my_ncols columns of `rows` complex rows, ld = rows)  ->  lin (nband columns of my_nrows rows, ld = my_nrows); `pack` is a
// scratch block of my_ncols * rows complex numbers.  One pack kernel + one grouped send/recv exchange, all on `st`.
void transpose_cols_to_rows(const double* cols, double* lin, double* pack, long long rows, int nband, cudaStream_t st);
void transpose_rows_to_cols(const double* lin, double* cols, double* pack, long long rows, int nband, cudaStream_t st);

}  // namespace abi
