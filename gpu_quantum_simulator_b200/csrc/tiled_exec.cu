/*
 * tiled_exec.cu -- execution of a TiledPlan: one kernel launch per pass on the
 * handle's stream.  Pass descriptors travel as kernel parameters, so the timed
 * region contains nothing but the pass kernels (and exchanges, multi-GPU).
 */
#include <string.h>

#include "sim.h"
#include "tiled.h"

int tiled_launch_pass(qsb_sim *s, const TiledPlan *p, size_t k, void *const *src_ptrs, void *dst, bool peer);

int tiled_plan_build(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                     const std::vector<COp> &cops, const double gphase[2], bool /*with_device*/,
                     TiledPlan **out, qsb_run_stats_t *stats)
{
    TiledPlan *p = new TiledPlan();
    int rc = tiled_schedule(n, prec, g, nloc, rank, opt, start, cops, gphase, p);
    if (rc) { delete p; return rc; }
    uint64_t n_ops = 0, n_rounds = 0;
    for (auto &hp : p->passes) { n_ops += hp.ops.size(); n_rounds += hp.rounds.size(); }
    stats->device_ops = n_ops;
    stats->passes = (uint32_t)p->passes.size();
    stats->rounds = (uint32_t)n_rounds;
    stats->kernel_launches = (uint32_t)p->passes.size();
    stats->bytes_moved = (uint64_t)p->passes.size() * 2ULL * ((uint64_t)1 << nloc) * amp_bytes(prec);
    stats->swaps = 0; stats->bytes_exchanged = 0;
    *out = p;
    return QSB_OK;
}

void tiled_plan_free(TiledPlan *p) { delete p; }

double tiled_last_exchange_ms(const TiledPlan *p) { return p ? p->last_exchange_ms : 0.0; }
void tiled_comm_destroy(qsb_sim *) {}

int tiled_execute(qsb_sim *s, TiledPlan *p)
{
    void *src[8];
    for (int i = 0; i < 8; i++) src[i] = s->state;
    for (size_t k = 0; k < p->passes.size(); k++) {
        int rc = tiled_launch_pass(s, p, k, src, s->state, false);
        if (rc) return rc;
    }
    s->perm = p->end_perm;
    return QSB_OK;
}

/* ---- multi-GPU entry points (NCCL exchange): see tiled_comm section ---- */
extern "C" int qsb_comm_unique_id(void *) { qsb_set_error("multi-GPU exchange is not available in this build"); return QSB_ERR_COMM; }
extern "C" int qsb_comm_init(qsb_t *, const void *) { qsb_set_error("multi-GPU exchange is not available in this build"); return QSB_ERR_COMM; }
