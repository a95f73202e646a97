#include "comm.cuh"
#include "context.cuh"
#include "fourwf.cuh"
#ifndef ABI_EMU
#include <nccl.h>
#include <dlfcn.h>
#include <cstring>
#endif
#include <algorithm>
#include <vector>

namespace abi {

CommState& comm_state() { static CommState c; return c; }

#ifndef ABI_EMU
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& api() {
  static NcclApi a;
  if (a.lib) return a;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) { a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (a.lib) break; }
  ABI_CHECK(a.lib != nullptr, "band-parallel drivers: libnccl.so.2 not found (NCCL is loaded at run time)");
  auto sym = [&](const char* s) { void* p = dlsym(a.lib, s); ABI_CHECK(p != nullptr, "NCCL symbol missing"); return p; };
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
  a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
  a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
  a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
  a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  return a;
}
void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) { fprintf(stderr, "NCCL error in %s: %s\n", what, api().GetErrorString(r)); ABI_ERROR("NCCL call failed"); }
}
static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes in the NCCL ABI");
}  // namespace

void comm_get_unique_id(char* id128) {
  ncclUniqueId id;
  nccl_check(api().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(id128, &id, 128);
}
void comm_init(const char* id128, int nranks, int rank) {
  ensure_init();
  comm_destroy();
  CommState& c = comm_state();
  c.nranks = nranks; c.rank = rank;
  if (nranks <= 1) return;
  ncclUniqueId id; memcpy(&id, id128, 128);
  ncclComm_t comm;
  nccl_check(api().CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
  c.comm = comm; c.owned = true;
}
void comm_adopt(void* nccl_comm, int nranks, int rank) {
  comm_destroy();
  CommState& c = comm_state();
  c.comm = nccl_comm; c.nranks = nranks; c.rank = rank; c.owned = false;
  if (nranks > 1) { ABI_CHECK(nccl_comm != nullptr, "comm_adopt: null communicator"); api(); }
}
void comm_destroy() {
  CommState& c = comm_state();
  if (c.comm && c.owned) api().CommDestroy(static_cast<ncclComm_t>(c.comm));
  c = CommState();
}
void comm_allreduce(double* buf, size_t n, bool max_op, cudaStream_t st) {
  CommState& c = comm_state();
  if (c.nranks <= 1 || n == 0) return;
  nccl_check(api().AllReduce(buf, buf, n, ncclFloat64, max_op ? ncclMax : ncclSum, static_cast<ncclComm_t>(c.comm), st), "ncclAllReduce");
}
void comm_alltoallv(const double* sbuf, const size_t* soff, const size_t* scnt, double* rbuf, const size_t* roff,
                    const size_t* rcnt, cudaStream_t st) {
  CommState& c = comm_state();
  if (c.nranks <= 1) {
    if (scnt[0]) CUDA_CHECK(cudaMemcpyAsync(rbuf + roff[0], sbuf + soff[0], sizeof(double) * scnt[0], cudaMemcpyDeviceToDevice, st));
    return;
  }
  ncclComm_t comm = static_cast<ncclComm_t>(c.comm);
  nccl_check(api().GroupStart(), "ncclGroupStart");
  for (int q = 0; q < c.nranks; q++) {
    if (scnt[q]) nccl_check(api().Send(sbuf + soff[q], scnt[q], ncclFloat64, q, comm, st), "ncclSend");
    if (rcnt[q]) nccl_check(api().Recv(rbuf + roff[q], rcnt[q], ncclFloat64, q, comm, st), "ncclRecv");
  }
  nccl_check(api().GroupEnd(), "ncclGroupEnd");
}

// segment q of the packed buffer = rows [lo_q, hi_q) of all my columns, stored [column][row - lo_q]
__global__ void k_transposer_pack(double2* __restrict__ cols, double2* __restrict__ pack, long long rows, int ncols,
                                  long long base, long long rem, int unpack) {
  const long long total = rows * ncols;
  const long long big = (base + 1) * rem;                   // rows owned by the `rem` larger shards
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const long long c = o / rows, i = o - c * rows;
    long long q, lo, nr;
    if (i < big) { q = i / (base + 1); lo = q * (base + 1); nr = base + 1; }
    else { q = rem + (i - big) / base; lo = big + (q - rem) * base; nr = base; }
    const long long p = (long long)ncols * lo + c * nr + (i - lo);
    if (unpack) cols[o] = pack[p]; else pack[p] = cols[o];
  }
}

static void transposer_counts(long long rows, int nband, std::vector<size_t>& coff, std::vector<size_t>& ccnt, std::vector<size_t>& loff,
                              std::vector<size_t>& lcnt, long long* my_ncols, long long* my_nrows) {
  const CommState& c = comm_state();
  const int R = c.nranks;
  long long f, l, lo, hi;
  block_range(nband, R, c.rank, &f, &l);
  block_range(rows, R, c.rank, &lo, &hi);
  *my_ncols = l - f; *my_nrows = hi - lo;
  coff.resize(R); ccnt.resize(R); loff.resize(R); lcnt.resize(R);
  for (int q = 0; q < R; q++) {
    long long fq, lq, loq, hiq;
    block_range(nband, R, q, &fq, &lq);
    block_range(rows, R, q, &loq, &hiq);
    coff[q] = 2 * (size_t)(*my_ncols) * loq; ccnt[q] = 2 * (size_t)(*my_ncols) * (hiq - loq);     // packed (column-sharded) side
    loff[q] = 2 * (size_t)fq * (*my_nrows); lcnt[q] = 2 * (size_t)(lq - fq) * (*my_nrows);         // row-sharded side: columns of rank q
  }
}

void transpose_cols_to_rows(const double* cols, double* lin, double* pack, long long rows, int nband, cudaStream_t st) {
  std::vector<size_t> coff, ccnt, loff, lcnt;
  long long nc, nr;
  transposer_counts(rows, nband, coff, ccnt, loff, lcnt, &nc, &nr);
  const int R = comm_state().nranks;
  if (nc > 0) {
    const int blocks = (int)std::min<long long>(kNumSM * 8, ceil_div<long long>(rows * nc, 256));
    k_transposer_pack<<<blocks, 256, 0, st>>>(reinterpret_cast<double2*>(const_cast<double*>(cols)), reinterpret_cast<double2*>(pack), rows,
                                              (int)nc, rows / R, rows % R, 0);
    CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
  }
  comm_alltoallv(pack, coff.data(), ccnt.data(), lin, loff.data(), lcnt.data(), st);
}

void transpose_rows_to_cols(const double* lin, double* cols, double* pack, long long rows, int nband, cudaStream_t st) {
  std::vector<size_t> coff, ccnt, loff, lcnt;
  long long nc, nr;
  transposer_counts(rows, nband, coff, ccnt, loff, lcnt, &nc, &nr);
  const int R = comm_state().nranks;
  comm_alltoallv(lin, loff.data(), lcnt.data(), pack, coff.data(), ccnt.data(), st);
  if (nc > 0) {
    const int blocks = (int)std::min<long long>(kNumSM * 8, ceil_div<long long>(rows * nc, 256));
    k_transposer_pack<<<blocks, 256, 0, st>>>(reinterpret_cast<double2*>(cols), reinterpret_cast<double2*>(pack), rows, (int)nc,
                                              rows / R, rows % R, 1);
    CUDA_CHECK(cudaGetLastError());
    g_kernel_launches++;
  }
}
#else
void transpose_cols_to_rows(const double*, double*, double*, long long, int, cudaStream_t) {}
void transpose_rows_to_cols(const double*, double*, double*, long long, int, cudaStream_t) {}
void comm_get_unique_id(char*) {}
void comm_init(const char*, int, int) {}
void comm_adopt(void*, int, int) {}
void comm_destroy() {}
void comm_allreduce(double*, size_t, bool, cudaStream_t) {}
void comm_alltoallv(const double*, const size_t*, const size_t*, double*, const size_t*, const size_t*, cudaStream_t) {}
#endif

}  // namespace abi
