/*
 * tiled_exec.cu -- execution of a TiledPlan: one kernel launch per pass on the
 * handle's stream.  Pass descriptors travel as kernel parameters, so the timed
 * region contains nothing but the pass kernels and, when the state is sharded
 * over several GPUs, the global<->local qubit exchanges.
 *
 * Exchange (the reference has no multi-GPU path): the state is
 * sharded on the top g = log2(P) physical index bits, one process per GPU.  An
 * exchange swaps those rank bits with the top g LOCAL bits: the local buffer is
 * P contiguous chunks, chunk j goes to rank j and the chunk received from rank
 * j lands at position j of the second buffer -- an all-to-all of contiguous
 * chunks over NVLink, issued as one NCCL group of send/recv pairs.  NCCL is
 * loaded at run time (dlopen) so the library has no link-time dependency on it.
 */
#include <dlfcn.h>
#include <string.h>

#include "sim.h"
#include "tiled.h"

int tiled_launch_pass(qsb_sim *s, const TiledPlan *p, size_t k, const void *src, void *dst);

int tiled_plan_build(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                     const std::vector<COp> &cops, const double gphase[2], bool /*with_device*/,
                     TiledPlan **out, qsb_run_stats_t *stats)
{
    TiledPlan *p = new TiledPlan();
    int rc = tiled_schedule(n, prec, g, nloc, rank, opt, start, cops, gphase, p);
    if (rc) { delete p; return rc; }
    uint64_t n_ops = 0, n_rounds = 0, sweeps = 0, swaps = 0;
    for (auto &hp : p->passes) {
        if (hp.is_swap) { swaps++; continue; }
        sweeps++; n_ops += hp.ops.size(); n_rounds += hp.rounds.size();
    }
    const uint64_t local_bytes = ((uint64_t)1 << nloc) * amp_bytes(prec);
    stats->device_ops = n_ops;
    stats->passes = (uint32_t)sweeps;
    stats->rounds = (uint32_t)n_rounds;
    stats->kernel_launches = (uint32_t)sweeps;
    stats->bytes_moved = sweeps * 2ULL * local_bytes;
    stats->swaps = (uint32_t)swaps;
    stats->bytes_exchanged = swaps * (local_bytes - (local_bytes >> g));
    *out = p;
    return QSB_OK;
}

void tiled_plan_free(TiledPlan *p) { delete p; }
double tiled_last_exchange_ms(const TiledPlan *p) { return p ? p->last_exchange_ms : 0.0; }

/* ------------------------------------------------------------------ NCCL (dlopen) */
typedef struct { char internal[128]; } qsb_nccl_id_t;
typedef void *qsb_nccl_comm_t;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(qsb_nccl_id_t *) = nullptr;
    int (*CommInitRank)(qsb_nccl_comm_t *, int, qsb_nccl_id_t, int) = nullptr;
    int (*CommDestroy)(qsb_nccl_comm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, qsb_nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, qsb_nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.lib) return QSB_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    void *h = nullptr;
    for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) { qsb_set_error("cannot load NCCL (libnccl.so.2): %s", dlerror()); return QSB_ERR_COMM; }
#define SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { qsb_set_error("NCCL symbol %s missing", name); return QSB_ERR_COMM; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = h;
    return QSB_OK;
}
#define QSB_NCCL(call) do { int r_ = (call); if (r_ != 0) { qsb_set_error("NCCL: %s in %s at line %d", g_nccl.GetErrorString(r_), __FILE__, __LINE__); return QSB_ERR_COMM; } } while (0)

extern "C" int qsb_comm_unique_id(void *id128)
{
    if (!id128) { qsb_set_error("qsb_comm_unique_id: null argument"); return QSB_ERR_ARG; }
    int rc = nccl_load();
    if (rc) return rc;
    qsb_nccl_id_t id;
    QSB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return QSB_OK;
}

extern "C" int qsb_comm_init(qsb_t *s, const void *id128)
{
    if (!s || !id128) { qsb_set_error("qsb_comm_init: null argument"); return QSB_ERR_ARG; }
    if (s->world == 1) return QSB_OK;
    int rc = nccl_load();
    if (rc) return rc;
    QSB_CUDA(cudaSetDevice(s->device));
    qsb_nccl_id_t id; memcpy(&id, id128, 128);
    qsb_nccl_comm_t c = nullptr;
    QSB_NCCL(g_nccl.CommInitRank(&c, s->world, id, s->rank));
    s->comm = c;
    if (!s->state2) {
        cudaError_t e = cudaMalloc(&s->state2, s->state_bytes);
        if (e != cudaSuccess) { qsb_set_error("Malloc error: exchange buffer of %zu bytes (%s)", s->state_bytes, cudaGetErrorString(e)); (void)cudaGetLastError(); return QSB_ERR_NOMEM; }
    }
    return QSB_OK;
}

void tiled_comm_destroy(qsb_sim *s)
{
    if (s->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy((qsb_nccl_comm_t)s->comm); s->comm = nullptr; }
}

static int exchange(qsb_sim *s)
{
    if (!s->comm) { qsb_set_error("plan needs a qubit exchange but qsb_comm_init was not called"); return QSB_ERR_COMM; }
    const int P = s->world;
    const size_t chunk = s->state_bytes / P;
    char *src = (char *)s->state, *dst = (char *)s->state2;
    QSB_NCCL(g_nccl.GroupStart());
    for (int j = 0; j < P; j++) {
        if (j == s->rank) continue;
        QSB_NCCL(g_nccl.Send(src + (size_t)j * chunk, chunk, 1 /* ncclUint8 */, j, (qsb_nccl_comm_t)s->comm, s->stream));
        QSB_NCCL(g_nccl.Recv(dst + (size_t)j * chunk, chunk, 1, j, (qsb_nccl_comm_t)s->comm, s->stream));
    }
    QSB_NCCL(g_nccl.GroupEnd());
    QSB_CUDA(cudaMemcpyAsync(dst + (size_t)s->rank * chunk, src + (size_t)s->rank * chunk, chunk, cudaMemcpyDeviceToDevice, s->stream));
    void *t = s->state; s->state = s->state2; s->state2 = t;
    return QSB_OK;
}

int tiled_execute(qsb_sim *s, TiledPlan *p)
{
    std::vector<cudaEvent_t> ev;
    for (size_t k = 0; k < p->passes.size(); k++) {
        if (p->passes[k].is_swap) {
            cudaEvent_t a, b;
            QSB_CUDA(cudaEventCreate(&a)); QSB_CUDA(cudaEventCreate(&b));
            QSB_CUDA(cudaEventRecord(a, s->stream));
            int rc = exchange(s);
            if (rc) return rc;
            QSB_CUDA(cudaEventRecord(b, s->stream));
            ev.push_back(a); ev.push_back(b);
            continue;
        }
        int rc = tiled_launch_pass(s, p, k, s->state, s->state);
        if (rc) return rc;
    }
    s->perm = p->end_perm;
    p->last_exchange_ms = 0.0;
    if (!ev.empty()) {
        QSB_CUDA(cudaStreamSynchronize(s->stream));
        for (size_t i = 0; i < ev.size(); i += 2) {
            float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            p->last_exchange_ms += ms;
            cudaEventDestroy(ev[i]); cudaEventDestroy(ev[i + 1]);
        }
    }
    return QSB_OK;
}
