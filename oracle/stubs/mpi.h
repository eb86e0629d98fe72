/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Nothing under elba_b200/ may include,
 * link or execute this.
 *
 * A thread-backed stand-in for <mpi.h>: "ranks" are threads of one process,
 * collectives are pointer exchanges across a pthread barrier.  It exists so
 * that the reference's own translation units on the hot path
 * (/root/reference/src/{KmerOps,SharedSeeds,Logger,HyperLogLog,...}.cpp)
 * compile UNMODIFIED from where they lie and run as `mpirun -np P` would on
 * one node, without an MPI installation (none exists in this image).
 *
 * Only the MPI-3 subset those files name is provided (SURVEY.md §2.2 lists
 * the call sites): Comm_rank/size, Barrier, Allreduce, Reduce, Exscan,
 * Alltoall, Alltoallv, Gather, Gatherv; and for src/FastaIndex.cpp (the step
 * before the hot path): Bcast, Scatterv, contiguous derived
 * datatypes, MPI_File_open / get_size / read_at_all / close over pread(2).
 */
#ifndef ELBA_ORACLE_FAKE_MPI_H
#define ELBA_ORACLE_FAKE_MPI_H

#include <pthread.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cstdlib>
#include <vector>

#define MPI_VERSION 3

typedef int MPI_Comm;
typedef long long MPI_Count;
typedef long MPI_Aint;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;     /* DistributedFastaData.hpp:54 only declares vectors of them */

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_IN_PLACE ((void*)(intptr_t)-1)

enum { MPI_CHAR = 1, MPI_BYTE, MPI_UINT8_T, MPI_INT, MPI_UNSIGNED_LONG, MPI_INT64_T, MPI_DOUBLE, MPI_COUNT };
enum { MPI_MAX = 1, MPI_SUM, MPI_LAND, MPI_MIN };

namespace fake_mpi {

struct World
{
    int size = 1;
    pthread_barrier_t bar;
    std::vector<const void*> p0, p1, p2;   /* per-rank published pointers */
    explicit World(int n) : size(n), p0(n), p1(n), p2(n) { pthread_barrier_init(&bar, nullptr, n); }
    ~World() { pthread_barrier_destroy(&bar); }
};

inline World *g_world = nullptr;
inline thread_local int t_rank = 0;

inline int nranks() { return g_world ? g_world->size : 1; }
inline void barrier() { if (g_world && g_world->size > 1) pthread_barrier_wait(&g_world->bar); }

/* derived datatypes (MPI_Type_contiguous): ids from 1000 up, sizes in a small table guarded by the callers' own barriers */
inline size_t g_derived[64];
inline int g_nderived = 0;
inline pthread_mutex_t g_derived_mu = PTHREAD_MUTEX_INITIALIZER;

inline size_t dtsize(MPI_Datatype t)
{
    if (t >= 1000) return g_derived[(t - 1000) & 63];
    switch (t)
    {
        case MPI_CHAR: case MPI_BYTE: case MPI_UINT8_T: return 1;
        case MPI_INT: return 4;
        default: return 8;
    }
}

template <class T> inline void red(T *acc, const T *in, int n, MPI_Op op)
{
    for (int i = 0; i < n; ++i)
    {
        switch (op)
        {
            case MPI_MAX:  acc[i] = in[i] > acc[i] ? in[i] : acc[i]; break;
            case MPI_MIN:  acc[i] = in[i] < acc[i] ? in[i] : acc[i]; break;
            case MPI_SUM:  acc[i] = acc[i] + in[i]; break;
            case MPI_LAND: acc[i] = (acc[i] && in[i]); break;
        }
    }
}

inline void red_any(void *acc, const void *in, int n, MPI_Datatype t, MPI_Op op)
{
    switch (t)
    {
        case MPI_CHAR: case MPI_BYTE: case MPI_UINT8_T: red((uint8_t*)acc, (const uint8_t*)in, n, op); break;
        case MPI_INT:           red((int*)acc, (const int*)in, n, op); break;
        case MPI_UNSIGNED_LONG: red((unsigned long*)acc, (const unsigned long*)in, n, op); break;
        case MPI_INT64_T:       red((int64_t*)acc, (const int64_t*)in, n, op); break;
        case MPI_COUNT:         red((long long*)acc, (const long long*)in, n, op); break;
        case MPI_DOUBLE:        red((double*)acc, (const double*)in, n, op); break;
    }
}

/* reduce over ranks [0, upto) of everyone's published input into tmp */
inline void fold_ranks(std::vector<uint8_t>& tmp, int upto, int n, MPI_Datatype t, MPI_Op op)
{
    size_t bytes = (size_t)n * dtsize(t);
    tmp.resize(bytes);
    if (upto <= 0) return;
    std::memcpy(tmp.data(), g_world->p0[0], bytes);
    for (int r = 1; r < upto; ++r) red_any(tmp.data(), g_world->p0[r], n, t, op);
}

} // namespace fake_mpi

static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = fake_mpi::t_rank; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = fake_mpi::nranks(); return 0; }
static inline int MPI_Barrier(MPI_Comm) { fake_mpi::barrier(); return 0; }
#define MPI_COMM_TYPE_SHARED 1
#define MPI_INFO_NULL 0
/* every thread-rank lives on the one node: the node communicator is the world */
static inline int MPI_Comm_split_type(MPI_Comm c, int, int, int, MPI_Comm *out) { *out = c; return 0; }
static inline int MPI_Comm_free(MPI_Comm *) { return 0; }
static inline int MPI_Abort(MPI_Comm, int code) { std::fprintf(stderr, "MPI_Abort(%d)\n", code); std::abort(); return 0; }
/* root's buffer to every rank-thread */
static inline int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm)
{
    using namespace fake_mpi;
    if (nranks() == 1) return 0;
    g_world->p0[t_rank] = buf; barrier();
    if (t_rank != root) std::memcpy(buf, g_world->p0[root], (size_t)n * dtsize(t));
    barrier();
    return 0;
}
static inline double MPI_Wtime()
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static inline int MPI_Allreduce(const void *send, void *recv, int n, MPI_Datatype t, MPI_Op op, MPI_Comm)
{
    using namespace fake_mpi;
    size_t bytes = (size_t)n * dtsize(t);
    if (nranks() == 1) { if (send != MPI_IN_PLACE) std::memcpy(recv, send, bytes); return 0; }
    g_world->p0[t_rank] = (send == MPI_IN_PLACE) ? recv : send;
    barrier();
    std::vector<uint8_t> tmp; fold_ranks(tmp, nranks(), n, t, op);
    barrier();
    std::memcpy(recv, tmp.data(), bytes);
    return 0;
}

static inline int MPI_Reduce(const void *send, void *recv, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm)
{
    using namespace fake_mpi;
    size_t bytes = (size_t)n * dtsize(t);
    if (nranks() == 1) { if (send != MPI_IN_PLACE) std::memcpy(recv, send, bytes); return 0; }
    g_world->p0[t_rank] = (send == MPI_IN_PLACE) ? recv : send;
    barrier();
    std::vector<uint8_t> tmp;
    if (t_rank == root) fold_ranks(tmp, nranks(), n, t, op);
    barrier();
    if (t_rank == root) std::memcpy(recv, tmp.data(), bytes);
    return 0;
}

static inline int MPI_Exscan(const void *send, void *recv, int n, MPI_Datatype t, MPI_Op op, MPI_Comm)
{
    using namespace fake_mpi;
    if (nranks() == 1) return 0; /* recvbuf on rank 0 is undefined by the standard */
    size_t bytes = (size_t)n * dtsize(t);
    g_world->p0[t_rank] = (send == MPI_IN_PLACE) ? recv : send;
    barrier();
    std::vector<uint8_t> tmp; fold_ranks(tmp, t_rank, n, t, op);
    barrier();
    if (t_rank > 0) std::memcpy(recv, tmp.data(), bytes);
    return 0;
}

static inline int MPI_Alltoall(const void *send, int scount, MPI_Datatype st, void *recv, int rcount, MPI_Datatype rt, MPI_Comm)
{
    using namespace fake_mpi;
    size_t sb = (size_t)scount * dtsize(st), rb = (size_t)rcount * dtsize(rt);
    if (nranks() == 1) { std::memcpy(recv, send, sb); return 0; }
    g_world->p0[t_rank] = send;
    barrier();
    for (int p = 0; p < nranks(); ++p)
        std::memcpy((uint8_t*)recv + p * rb, (const uint8_t*)g_world->p0[p] + t_rank * sb, rb);
    barrier();
    return 0;
}

static inline int MPI_Alltoallv(const void *send, const int *scnt, const int *sdis, MPI_Datatype st,
                                void *recv, const int *rcnt, const int *rdis, MPI_Datatype rt, MPI_Comm)
{
    using namespace fake_mpi;
    size_t sz = dtsize(st); (void)rt;
    if (nranks() == 1) { std::memcpy((uint8_t*)recv + rdis[0] * sz, (const uint8_t*)send + sdis[0] * sz, (size_t)scnt[0] * sz); return 0; }
    g_world->p0[t_rank] = send; g_world->p1[t_rank] = sdis; g_world->p2[t_rank] = scnt;
    barrier();
    for (int p = 0; p < nranks(); ++p)
    {
        const int *pdis = (const int*)g_world->p1[p];
        std::memcpy((uint8_t*)recv + (size_t)rdis[p] * sz, (const uint8_t*)g_world->p0[p] + (size_t)pdis[t_rank] * sz, (size_t)rcnt[p] * sz);
    }
    barrier();
    return 0;
}

/* root's blocks send[dis[p] .. dis[p] + cnt[p]) to rank p */
static inline int MPI_Scatterv(const void *send, const int *cnt, const int *dis, MPI_Datatype st, void *recv, int rcount, MPI_Datatype rt, int root, MPI_Comm)
{
    using namespace fake_mpi;
    size_t sz = dtsize(st); (void)rt;
    if (nranks() == 1) { std::memcpy(recv, (const uint8_t*)send + (size_t)dis[0] * sz, (size_t)rcount * sz); return 0; }
    g_world->p0[t_rank] = send; g_world->p1[t_rank] = dis;
    barrier();
    std::memcpy(recv, (const uint8_t*)g_world->p0[root] + (size_t)((const int*)g_world->p1[root])[t_rank] * sz, (size_t)rcount * sz);
    barrier();
    return 0;
}

static inline int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype *out)
{
    using namespace fake_mpi;
    pthread_mutex_lock(&g_derived_mu);
    const int id = g_nderived++ & 63;
    g_derived[id] = (size_t)n * dtsize(old);
    pthread_mutex_unlock(&g_derived_mu);
    *out = 1000 + id;
    return 0;
}
static inline int MPI_Type_commit(MPI_Datatype *) { return 0; }
static inline int MPI_Type_free(MPI_Datatype *) { return 0; }

/* MPI-IO, read side only: every rank-thread opens the file itself */
typedef long long MPI_Offset;
typedef int MPI_File;
typedef int MPI_Status;
#define MPI_MODE_RDONLY 2
#define MPI_STATUS_IGNORE ((MPI_Status*)nullptr)
static inline int MPI_File_open(MPI_Comm, const char *name, int, int, MPI_File *fh) { *fh = ::open(name, O_RDONLY); return *fh < 0; }
static inline int MPI_File_get_size(MPI_File fh, MPI_Offset *size) { struct stat sb; if (fstat(fh, &sb)) return 1; *size = (MPI_Offset)sb.st_size; return 0; }
static inline int MPI_File_read_at_all(MPI_File fh, MPI_Offset off, void *buf, int n, MPI_Datatype t, MPI_Status *)
{
    size_t want = (size_t)n * fake_mpi::dtsize(t), got = 0;
    while (got < want) { ssize_t r = ::pread(fh, (char*)buf + got, want - got, (off_t)(off + (MPI_Offset)got)); if (r <= 0) break; got += (size_t)r; }
    return got != want;
}
static inline int MPI_File_close(MPI_File *fh) { ::close(*fh); *fh = -1; return 0; }

static inline int MPI_Gather(const void *send, int scount, MPI_Datatype st, void *recv, int rcount, MPI_Datatype rt, int root, MPI_Comm)
{
    using namespace fake_mpi;
    size_t sb = (size_t)scount * dtsize(st); (void)rcount; (void)rt;
    if (nranks() == 1) { std::memcpy(recv, send, sb); return 0; }
    g_world->p0[t_rank] = send;
    barrier();
    if (t_rank == root) for (int p = 0; p < nranks(); ++p) std::memcpy((uint8_t*)recv + p * sb, g_world->p0[p], sb);
    barrier();
    return 0;
}

static inline int MPI_Allgather(const void *send, int scount, MPI_Datatype st, void *recv, int rcount, MPI_Datatype rt, MPI_Comm)
{
    using namespace fake_mpi;
    size_t sb = (size_t)scount * dtsize(st); (void)rcount; (void)rt;
    if (nranks() == 1) { std::memcpy(recv, send, sb); return 0; }
    g_world->p0[t_rank] = send;
    barrier();
    for (int p = 0; p < nranks(); ++p) std::memcpy((uint8_t*)recv + p * sb, g_world->p0[p], sb);
    barrier();
    return 0;
}

static inline int MPI_Allgatherv(const void *send, int scount, MPI_Datatype st, void *recv, const int *rcnt, const int *rdis, MPI_Datatype rt, MPI_Comm)
{
    using namespace fake_mpi;
    size_t sz = dtsize(st); (void)rt;
    if (nranks() == 1) { std::memcpy((uint8_t*)recv + (size_t)rdis[0] * sz, send, (size_t)scount * sz); return 0; }
    g_world->p0[t_rank] = send;
    barrier();
    for (int p = 0; p < nranks(); ++p) std::memcpy((uint8_t*)recv + (size_t)rdis[p] * sz, g_world->p0[p], (size_t)rcnt[p] * sz);
    barrier();
    return 0;
}

static inline int MPI_Gatherv(const void *send, int scount, MPI_Datatype st, void *recv, const int *rcnt, const int *rdis, MPI_Datatype rt, int root, MPI_Comm)
{
    using namespace fake_mpi;
    size_t sz = dtsize(st); (void)rt;
    if (nranks() == 1) { std::memcpy((uint8_t*)recv + rdis[0] * sz, send, (size_t)scount * sz); return 0; }
    g_world->p0[t_rank] = send;
    barrier();
    if (t_rank == root) for (int p = 0; p < nranks(); ++p) std::memcpy((uint8_t*)recv + (size_t)rdis[p] * sz, g_world->p0[p], (size_t)rcnt[p] * sz);
    barrier();
    return 0;
}

#endif
