// TEST INFRASTRUCTURE (tests/test_transitive_host.py::test_shim_host_logic_with_a_mock_abi): the shim's TransitiveReduction
// (elba_b200/host/elba_fe_shim.cpp, -DELBA_FE_SHIM_TR) linked against a MOCK of the C ABI whose elba_fe_transitive_reduction
// is the CPU oracle: checks, without a GPU, everything the shim does on the host around the device call (the walk over
// seqptr()->GetDCSC(), global ids, Overlap::Transpose of the payload of mirror images, the triple constructor).
// The product never links this; the GPU run of the same path is tests/test_gpu_shim.py.
#include <mpi.h>
#include "common.h"
#include "KmerOps.hpp"
#include "SharedSeeds.hpp"
#include "Overlap.hpp"
#include "FastaIndex.hpp"
#include "TransitiveReduction.hpp"
#include SHIM_CPP
#include <cstdio>
#include <dlfcn.h>
// mock of the C ABI for the TR path: the oracle's restatement does the work (CPU-only check of the shim's host logic)
typedef uint64_t (*eo_tr_t)(int64_t, uint64_t, const int64_t*, const int64_t*, const int32_t*, int32_t, int64_t*, int64_t*, int32_t*, uint64_t*, uint8_t*);
static uint64_t g_n = 0; static std::vector<int64_t> g_r, g_c; static std::vector<int32_t> g_f; static std::vector<uint64_t> g_s; static std::vector<uint8_t> g_t;
extern "C" {
void elba_fe_default_config(elba_fe_config *c) { memset(c, 0, sizeof *c); }
int elba_fe_create(const elba_fe_config*, elba_fe_ctx **out) { *out = (elba_fe_ctx*)0x1; return 0; }
int elba_fe_destroy(elba_fe_ctx*) { return 0; }
const char *elba_fe_last_error(const elba_fe_ctx*) { return "mock"; }
int elba_fe_device_count(void) { return 1; }
void elba_fe_block_extent(int64_t n, int parts, int idx, int64_t *o, int64_t *l) { int64_t per = n / parts; *o = per * idx; *l = idx == parts - 1 ? n - per * idx : per; }
int elba_fe_transitive_reduction(elba_fe_ctx*, const int64_t *row, const int64_t *col, const int32_t *f, uint64_t nnz, int64_t n, int32_t fuzz, uint64_t *out)
{
    void *h = dlopen(ORACLE_SO, RTLD_NOW); eo_tr_t fn = (eo_tr_t)dlsym(h, "eo_transitive_reduction");
    g_r.assign(2 * nnz + 1, 0); g_c.assign(2 * nnz + 1, 0); g_f.assign(8 * nnz + 4, 0); g_s.assign(2 * nnz + 1, 0); g_t.assign(2 * nnz + 1, 0);
    *out = g_n = fn(n, nnz, row, col, f, fuzz, g_r.data(), g_c.data(), g_f.data(), g_s.data(), g_t.data());
    return 0;
}
int elba_fe_get_string_graph(elba_fe_ctx*, int64_t *row, int64_t *col, int32_t *f, uint64_t *src, uint8_t *tr)
{
    memcpy(row, g_r.data(), 8 * g_n); memcpy(col, g_c.data(), 8 * g_n); memcpy(f, g_f.data(), 16 * g_n); memcpy(src, g_s.data(), 8 * g_n); memcpy(tr, g_t.data(), g_n);
    return 0;
}
}

int main(int argc, char **argv)
{
    // input: n nnz then nnz lines "row col dir dirT suf sufT"
    FILE *fi = fopen(argv[1], "r"); long n, nnz; fscanf(fi, "%ld %ld", &n, &nnz);
    std::vector<int64_t> r(nnz), c(nnz); std::vector<Overlap> v(nnz);
    for (long e = 0; e < nnz; ++e) { long a, b; int d, dt, s, st; fscanf(fi, "%ld %ld %d %d %d %d", &a, &b, &d, &dt, &s, &st); r[e] = a; c[e] = b; v[e].direction = d; v[e].directionT = dt; v[e].suffix = s; v[e].suffixT = st;
        std::get<0>(v[e].len) = 1000 + e; std::get<1>(v[e].len) = 2000 + e; v[e].containedQ = true; }
    auto grid = std::make_shared<CommGrid>(MPI_COMM_WORLD, 0, 0);
    CT<int64_t>::PDistVec dr(r, grid), dc(c, grid); CT<Overlap>::PDistVec dv(v, grid);
    CT<Overlap>::PSpParMat R(n, n, dr, dc, dv, false);
    auto S = TransitiveReduction(R);
    const auto &st = *S->st;
    for (int64_t i = 0; i < st.m; ++i) for (int64_t p = st.rowptr[i]; p < st.rowptr[i+1]; ++p)
    { const Overlap &o = st.val[p]; printf("%ld %ld %d %d %d %d %u %u %d %d\n", (long)i, (long)st.col[p], o.direction, o.directionT, o.suffix, o.suffixT, std::get<0>(o.len), std::get<1>(o.len), (int)o.containedQ, (int)o.containedT); }
}

