/*
 * tiled_exec.cu -- device image of a TiledPlan and its execution.
 * One kernel launch per pass on the handle's stream; the whole circuit's
 * descriptors are uploaded once at plan time, so the timed region contains
 * nothing but the pass kernels (and exchanges, multi-GPU).
 */
#include <string.h>

#include "sim.h"
#include "tiled.h"

int tiled_launch_pass(qsb_sim *s, const TiledPlan *p, size_t k, void *const *src_ptrs, void *dst);

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename R>
static void pack_op(const HostOp &h, DevOp<R> &d)
{
    d.kind = h.kind; d.vmask = h.vmask; d.tmask = h.tmask;
    if (sizeof(R) == 4) { for (int k = 0; k < 32; k++) d.c[k] = (R)h.c[k]; }
    else { for (int set = 0; set < 2; set++) for (int c = 0; c < 8; c++) d.c[set * 8 + c] = (R)h.c[set * 16 + c * 2]; }
}

int tiled_plan_build(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                     const std::vector<COp> &cops, const double gphase[2], bool with_device,
                     TiledPlan **out, qsb_run_stats_t *stats)
{
    TiledPlan *p = new TiledPlan();
    int rc = tiled_schedule(n, prec, g, nloc, rank, opt, start, cops, gphase, p);
    if (rc) { delete p; return rc; }
    const size_t opsz = sizeof(DevOp<float>); /* same for both precisions */
    static_assert(sizeof(DevOp<float>) == sizeof(DevOp<double>), "op size");
    size_t total = 0;
    uint64_t n_ops = 0, n_rounds = 0;
    for (auto &hp : p->passes) {
        total = align_up(total, 256);
        p->pass_off.push_back(total); total += sizeof(DevPass);
        total = align_up(total, 16);
        p->round_off.push_back(total); total += hp.rounds.size() * sizeof(DevRound);
        total = align_up(total, 16);
        p->op_off.push_back(total); total += hp.ops.size() * opsz;
        n_ops += hp.ops.size(); n_rounds += hp.rounds.size();
    }
    stats->device_ops = n_ops;
    stats->passes = (uint32_t)p->passes.size();
    stats->rounds = (uint32_t)n_rounds;
    stats->kernel_launches = (uint32_t)p->passes.size();
    stats->bytes_moved = (uint64_t)p->passes.size() * 2ULL * ((uint64_t)1 << nloc) * amp_bytes(prec);
    stats->swaps = 0; stats->bytes_exchanged = 0;
    p->blob_bytes = total;
    if (with_device && total) {
        std::vector<uint8_t> host(total, 0);
        for (size_t k = 0; k < p->passes.size(); k++) {
            const HostPass &hp = p->passes[k];
            memcpy(host.data() + p->pass_off[k], &hp.hdr, sizeof(DevPass));
            memcpy(host.data() + p->round_off[k], hp.rounds.data(), hp.rounds.size() * sizeof(DevRound));
            for (size_t i = 0; i < hp.ops.size(); i++) {
                if (prec == QSB_F32) pack_op<float>(hp.ops[i], *(DevOp<float> *)(host.data() + p->op_off[k] + i * opsz));
                else pack_op<double>(hp.ops[i], *(DevOp<double> *)(host.data() + p->op_off[k] + i * opsz));
            }
        }
        cudaError_t e = cudaMalloc(&p->d_blob, total);
        if (e != cudaSuccess) { qsb_set_error("Malloc error: plan blob (%s)", cudaGetErrorString(e)); delete p; return QSB_ERR_NOMEM; }
        e = cudaMemcpy(p->d_blob, host.data(), total, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { qsb_set_error("%s in %s at line %d", cudaGetErrorString(e), __FILE__, __LINE__); cudaFree(p->d_blob); delete p; return QSB_ERR_CUDA; }
    }
    *out = p;
    return QSB_OK;
}

void tiled_plan_free(TiledPlan *p)
{
    if (!p) return;
    if (p->d_blob) cudaFree(p->d_blob);
    delete p;
}

double tiled_last_exchange_ms(const TiledPlan *p) { return p ? p->last_exchange_ms : 0.0; }
void tiled_comm_destroy(qsb_sim *) {}

int tiled_execute(qsb_sim *s, TiledPlan *p)
{
    void *src[8];
    for (int i = 0; i < 8; i++) src[i] = s->state;
    for (size_t k = 0; k < p->passes.size(); k++) {
        int rc = tiled_launch_pass(s, p, k, src, s->state);
        if (rc) return rc;
    }
    s->perm = p->end_perm;
    return QSB_OK;
}

/* ---- multi-GPU entry points (NCCL exchange): see tiled_comm section ---- */
extern "C" int qsb_comm_unique_id(void *) { qsb_set_error("multi-GPU exchange is not available in this build"); return QSB_ERR_COMM; }
extern "C" int qsb_comm_init(qsb_t *, const void *) { qsb_set_error("multi-GPU exchange is not available in this build"); return QSB_ERR_COMM; }
