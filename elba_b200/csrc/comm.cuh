// Multi-GPU plumbing: one context per GPU / process, NCCL over NVLink 5 / NVSwitch.
//
// The reference's MPI call sites on this path and what replaces them here:
//   KmerOps.cpp:117,151   Alltoall(counts) + Alltoallv(k-mers)       -> exchange_partitions(): grouped ncclSend/ncclRecv of the
//                                                                        level-1 partition slabs (8 B per instance, one exchange;
//                                                                        the reference's second, 20 B per instance exchange
//                                                                        KmerOps.cpp:244,274 does not exist: sweep 2 is local)
//   KmerOps.cpp:371-374   Allreduce/Exscan of k-mer counts            -> allgather of the owners' reliable lists; every rank ranks them
//   SpParMat ctor + Transpose + Mult_AnXBn_DoubleBuff broadcasts      -> allgather of A's row blocks; block (i,j) of B is computed
//                                                                        on one GPU with the inner dimension unsplit
// NCCL is dlopen'ed at elba_fe_comm_init: a single-GPU user never needs it, and inside a process that already
// carries a libnccl.so.2 (torch) that copy is the one used.
#pragma once
#include "common.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <string>
#include <vector>

namespace elba {

struct NcclApi
{
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

inline NcclApi *nccl_api(std::string &err)
{
    static NcclApi api; static bool tried = false;
    if (api.lib) return &api;
    if (tried) { err = "libnccl.so.2 could not be loaded"; return nullptr; }
    tried = true;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return nullptr; }
#define ELBA_NCCL_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(lib, name)); if (!api.field) { err = std::string("libnccl lacks ") + name; return nullptr; }
    ELBA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") ELBA_NCCL_SYM(CommInitRank, "ncclCommInitRank") ELBA_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    ELBA_NCCL_SYM(GroupStart, "ncclGroupStart") ELBA_NCCL_SYM(GroupEnd, "ncclGroupEnd") ELBA_NCCL_SYM(Send, "ncclSend") ELBA_NCCL_SYM(Recv, "ncclRecv")
    ELBA_NCCL_SYM(AllReduce, "ncclAllReduce") ELBA_NCCL_SYM(AllGather, "ncclAllGather") ELBA_NCCL_SYM(Broadcast, "ncclBroadcast")
    ELBA_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef ELBA_NCCL_SYM
    api.lib = lib;
    return &api;
}

struct Comm
{
    NcclApi *api = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int grid_rows = 1, grid_cols = 1;       // pr x pc, rank = i * pc + j
};

// CombBLAS block distribution (src/DistributedFastaData.cpp:21-29): n / parts each, the remainder to the last block
__host__ __device__ inline void block_extent(int64_t n, int parts, int idx, int64_t &off, int64_t &len)
{
    int64_t per = n / parts;
    off = per * idx;
    len = (idx == parts - 1) ? n - off : per;
}

// default grids of the north star: 1x1, 1x2, 2x2, 2x4; otherwise the most square factorisation with rows <= cols
inline void default_grid(int nranks, int &rows, int &cols)
{
    rows = 1;
    for (int r = 1; r * r <= nranks; ++r) if (nranks % r == 0) rows = r;
    cols = nranks / rows;
}

// ---- kernels of the exchange steps ---------------------------------------------------------------------
// local A (CSR, local read ids) -> packed triples with GLOBAL read ids, in (read, column) order
__global__ void k_pack_rows(const int64_t *__restrict__ rowptr, const u32 *__restrict__ col, u32 nrows, u64 read_offset, u64 *__restrict__ key)
{
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    int64_t b = rowptr[warp], e = rowptr[warp + 1];
    for (int64_t p = b + lane; p < e; p += 32) key[p] = ((read_offset + warp) << 32) | col[p];
}

// gathered triples (key = read << 32 | col, ascending): ptr[r - r0] = first index with read >= r, r in [r0, r0 + n]
__global__ void k_read_ptr(const u64 *__restrict__ key, u64 nnz, u64 r0, u64 n, int64_t *__restrict__ ptr)
{
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    u64 target = r0 + r, lo = 0, hi = nnz;
    while (lo < hi) { u64 mid = (lo + hi) >> 1; if ((key[mid] >> 32) < target) lo = mid + 1; else hi = mid; }
    ptr[r] = (int64_t)lo;
}

// slice [b, e) of the gathered triples -> left operand columns / right operand sort keys (col << rbits | local read)
__global__ void k_slice_left(const u64 *__restrict__ key, u64 b, u64 n, u32 *__restrict__ col)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) col[i] = (u32)key[b + i];
}
__global__ void k_slice_right(const u64 *__restrict__ key, u64 b, u64 n, u64 r0, int rbits, u64 *__restrict__ out)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { u64 kk = key[b + i]; out[i] = ((kk & 0xFFFFFFFFull) << rbits) | ((kk >> 32) - r0); }
}
__global__ void k_sub_base(int64_t *__restrict__ ptr, u64 n, int64_t base)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) ptr[i] -= base;
}

} // namespace elba
