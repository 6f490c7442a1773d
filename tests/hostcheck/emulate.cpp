/*
 * emulate.cpp -- TEST DOUBLE, never shipped: a host interpreter of the tiled
 * schedule (DevPass / DevRound / HostOp tables), thread by thread and slot by
 * slot, exactly as tiled_kernel.cu consumes them.  It lets the CPU test-suite
 * verify the planner's tables (index maps, slot maps, op encodings) against the
 * oracle without a GPU, and it checks what the kernel relies on:
 *   - every smem slot is written exactly once per exchange,
 *   - every 128-bit shared access phase (16-byte slots, both precisions) is bank-conflict free,
 *   - first-round loads / last-round stores of a warp are contiguous.
 * It is linked only into tests/hostcheck/libqsb_hostcheck.so.
 */
#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstring>
#include <set>
#include <vector>

#include "sim.h"
#include "tiled.h"

typedef std::complex<double> cd;

struct Report { int max_conflict; int bad_slots; int noncontig; int passes; int rounds; };

/* outs: fused-exchange passes scatter into the shard of the rank named by the index bits above nloc
 * (outs[rank] = that rank's NEW shard); null for ordinary in-place passes */
int blob_run_pass(const HostPass &hp, bool f32, int nloc, std::vector<cd> &st, cd *const *outs);   /* blob_emulate.cpp */
static int g_use_blob = 0;   /* 1: interpret the device encoding (the kernel-parameter blob) instead of the logical tables */
static int g_low_a = 0;      /* contiguous low index bits of the plan being run (0: precision default), for the coalescing check */

/* Coalescing of the global accesses of an edge round: for a fixed vector, each group of 2^k consecutive lanes
 * (k = lane bits inside the low segment, at most 3) must touch 2^k consecutive 16-byte units of ONE aligned run.
 * idx[tid]: amplitude index that thread tid accesses for that vector (pack bit clear).  Returns the violations. */
static int lanes_not_contiguous(const std::vector<uint64_t> &idx, bool f32, int a)
{
    const int unit_shift = f32 ? 1 : 0;                    /* amplitudes per 16-byte unit: 2 (f32) or 1 (f64) */
    const int k = std::min(3, a - unit_shift);             /* lane bits inside the contiguous low segment */
    if (k <= 0) return 0;
    int bad = 0;
    for (size_t t0 = 0; t0 < idx.size(); t0 += (size_t)1 << k) {
        const uint64_t base = idx[t0] >> unit_shift;
        if (base & ((1ULL << k) - 1)) { bad++; continue; }
        for (int ln = 0; ln < (1 << k); ln++) if ((idx[t0 + ln] >> unit_shift) != base + (uint64_t)ln) { bad++; break; }
    }
    return bad;
}
extern "C" void qsb_hostcheck_use_blob(int on) { g_use_blob = on; }
/* -1 (default): the doubles run the schedule tiled_plan_search picks, i.e. what a GPU run executes; 0..7: one fixed order of
 * the tile hill climbing (tiled_schedule's climb_variant), no search */
static int g_climb = -1;
extern "C" void qsb_hostcheck_set_climb(int v) { g_climb = v; }
static int make_plan(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                     const std::vector<COp> &cops, const double gph[2], TiledPlan *plan)
{
    if (g_climb >= 0) return tiled_schedule(n, prec, g, nloc, rank, opt, start, cops, gph, plan, g_climb);
    TiledPlan *p = nullptr;
    int rc = tiled_plan_search(n, prec, g, nloc, rank, opt, start, cops, gph, &p);
    if (rc) return rc;
    *plan = std::move(*p);
    delete p;
    return 0;
}

static void run_pass(const HostPass &hp, bool f32, int nloc, std::vector<cd> &st, Report &rep, cd *const *outs = nullptr)
{
    if (g_use_blob) { rep.bad_slots += blob_run_pass(hp, f32, nloc, st, outs); return; }
    const int L = f32 ? 2 : 1;                  /* pack lanes */
    const int nb = 3;
    const int phase_lanes = 8;                  /* lanes per 128-bit shared-memory phase (16-byte slots) */
    const uint64_t loc_mask = (1ULL << nloc) - 1;
    std::vector<cd> regs((size_t)QSB_THREADS * QSB_NV * L);
    std::vector<cd> smem((size_t)QSB_SLOTS * L);
    std::vector<int> written(QSB_SLOTS);
    const int nr = (int)hp.hdr.n_rounds;
    for (uint64_t tile = 0; tile < hp.hdr.n_tiles; tile++) {
        uint64_t t = tile, outer = 0;
        for (uint32_t r = 0; r < hp.hdr.n_runs; r++) {
            int len = hp.hdr.run_len[r];
            outer |= (t & ((1ULL << len) - 1)) << hp.hdr.run_start[r];
            t >>= len;
        }
        const uint64_t src_outer = outer | hp.hdr.src_fixed;
        for (int rd = 0; rd < nr; rd++) {
            const DevRound &RD = hp.rounds[rd];
            std::vector<uint64_t> gthr(QSB_THREADS);
            /* ---- load ---- */
            for (int tid = 0; tid < QSB_THREADS; tid++) {
                uint64_t g = src_outer;
                for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) g |= RD.thr[j].gidx;
                gthr[tid] = g;
                uint32_t sb = 0;
                for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) sb ^= RD.thr[j].ld;
                for (int v = 0; v < QSB_NV; v++) {
                    if (rd == 0) {
                        uint64_t gi = g;
                        for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) gi |= RD.vec[b].gidx;
                        for (int l = 0; l < L; l++) regs[((size_t)tid * QSB_NV + v) * L + l] = st[(gi & loc_mask) | (uint64_t)l];
                        if (f32 && (gi & 1)) rep.noncontig++;
                    } else {
                        uint32_t slot = sb;
                        for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) slot ^= RD.vec[b].ld;
                        for (int l = 0; l < L; l++) regs[((size_t)tid * QSB_NV + v) * L + l] = smem[(size_t)slot * L + l];
                    }
                }
            }
            /* contiguity of the first load (the vector bits only add high index bits, so vector 0 stands for all) */
            if (rd == 0) {
                std::vector<uint64_t> idx(QSB_THREADS);
                for (int tid = 0; tid < QSB_THREADS; tid++) idx[tid] = gthr[tid] & loc_mask;
                rep.noncontig += lanes_not_contiguous(idx, f32, g_low_a > 0 ? g_low_a : (f32 ? 4 : 3));
            }
            /* bank check, load side */
            if (rd > 0) {
                for (int w = 0; w < QSB_THREADS / phase_lanes; w++) for (int v = 0; v < QSB_NV; v++) {
                    std::set<uint32_t> banks;
                    for (int ln = 0; ln < phase_lanes; ln++) {
                        int tid = w * phase_lanes + ln;
                        uint32_t slot = 0;
                        for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) slot ^= RD.thr[j].ld;
                        for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) slot ^= RD.vec[b].ld;
                        banks.insert(slot & ((1u << nb) - 1));
                    }
                    int conflict = phase_lanes / (int)banks.size();
                    if (conflict > rep.max_conflict) rep.max_conflict = conflict;
                }
            }
            /* ---- ops ---- */
            for (int tid = 0; tid < QSB_THREADS; tid++) {
                cd *R = &regs[(size_t)tid * QSB_NV * L];
                cd pend(1.0, 0.0);
                for (uint32_t i = 0; i < hp.round_op_count[rd]; i++) {
                    const HostOp &op = hp.ops[hp.round_op_begin[rd] + i];
                    const uint32_t code = op.kind & 0xff, vb = (op.kind >> 8) & 0xf, vmask = OPK_VMASK(op.kind);
                    const bool mux = (op.kind >> 16) & 1;
                    const bool pred = (gthr[tid] & op.tmask) == op.tmask;
                    if (!pred && !mux) continue;
                    const int set = (mux && pred) ? 1 : 0;
                    auto C = [&](int k, int l) { return op.c[set][k][f32 ? l : 1]; };
                    if (code == OP_TPHASE) { pend *= cd(op.tph[0], op.tph[1]); continue; }
                    if (code == OP_XDEF) {   /* the kernel defers this swap into its next store address: same result */
                        for (int v = 0; v < QSB_NV; v++) if (!((v >> vb) & 1))
                            for (int l = 0; l < L; l++) std::swap(R[v * L + l], R[(v | (1 << vb)) * L + l]);
                        continue;
                    }
                    for (int v = 0; v < QSB_NV; v++) {
                        if (code == OP_MAT_U || code == OP_MAT_UI) {
                            if ((v >> vb) & 1) continue;
                            int w = v | (1 << vb);
                            for (int l = 0; l < L; l++) {
                                cd x0 = R[v * L + l], x1 = R[w * L + l];
                                if (code == OP_MAT_U) { double p_ = C(0, l), q = C(1, l), k = C(2, l); x0 += p_ * x1; x1 = k * x1 + q * x0; }
                                else { double p_ = C(0, l), q = C(2, l), k = C(4, l);
                                    if (C(1, l) != -p_ || C(3, l) != -q) rep.bad_slots++;
                                    x0 += cd(0, p_) * x1; x1 = k * x1 + cd(0, q) * x0; }
                                R[v * L + l] = x0; R[w * L + l] = x1;
                            }
                            if (v == 0) pend *= (code == OP_MAT_U ? C(3, 0) : C(5, 0));
                        } else if (code == OP_MAT_R || code == OP_MAT_I || code == OP_MAT_G) {
                            if ((v >> vb) & 1) continue;
                            int w = v | (1 << vb);
                            for (int l = 0; l < L; l++) {
                                cd x0 = R[v * L + l], x1 = R[w * L + l], m00, m01, m10, m11;
                                if (code == OP_MAT_R) { m01 = C(0, l); m10 = C(1, l); m00 = C(2, l); m11 = C(3, l); }
                                else if (code == OP_MAT_I) { m01 = cd(0, C(1, l)); m10 = cd(0, C(3, l)); m00 = C(4, l); m11 = C(5, l);
                                    if (C(0, l) != -C(1, l) || C(2, l) != -C(3, l)) rep.bad_slots++; }
                                else { m00 = cd(C(6, l), C(0, l)); m01 = cd(C(1, l), C(2, l)); m10 = cd(C(3, l), C(4, l)); m11 = cd(C(7, l), C(5, l)); }
                                R[v * L + l] = m00 * x0 + m01 * x1;
                                R[w * L + l] = m10 * x0 + m11 * x1;
                            }
                        } else if (code == OP_MATP_R || code == OP_MATP_G) {
                            cd x0 = R[v * L], x1 = R[v * L + 1];
                            cd A0, A1, B0, B1;
                            if (code == OP_MATP_R) { A0 = C(0, 0); A1 = C(0, 1); B0 = C(1, 0); B1 = C(1, 1); }
                            else { A0 = cd(C(0, 0), C(1, 0)); A1 = cd(C(0, 1), C(1, 1)); B0 = cd(C(2, 0), C(3, 0)); B1 = cd(C(2, 1), C(3, 1)); }
                            R[v * L] = A0 * x0 + B0 * x1;
                            R[v * L + 1] = A1 * x1 + B1 * x0;
                        } else if (code == OP_DIAG_V || code == OP_DIAG_ALL || code == OP_DIAG_GEN) {
                            if (code == OP_DIAG_V && !((v >> vb) & 1)) continue;
                            if (code == OP_DIAG_GEN && (v & vmask) != vmask) continue;
                            for (int l = 0; l < L; l++) R[v * L + l] *= cd(C(0, l), C(1, l));
                        } else rep.bad_slots++;
                    }
                }
                if (hp.rounds[rd].flags & 1) for (int k = 0; k < QSB_NV * L; k++) R[k] *= pend;
                else if (pend != cd(1.0, 0.0)) rep.bad_slots++;
            }
            /* ---- store ---- */
            if (rd == nr - 1) {
                std::vector<uint64_t> didx(QSB_THREADS);
                for (int tid = 0; tid < QSB_THREADS; tid++) {
                    uint64_t d = outer | hp.hdr.dst_fixed;
                    for (uint32_t k = 0; k < hp.hdr.n_xo; k++) {   /* victims outside the tile: the outer bit names a bit of the destination rank */
                        const uint64_t bit = (outer >> hp.hdr.xo_pos[k]) & 1ULL;
                        d = (d & ~(1ULL << hp.hdr.xo_pos[k])) | hp.hdr.dst_fixed & (1ULL << hp.hdr.xo_pos[k]);
                        d |= bit << (nloc + hp.hdr.xo_rank[k]);
                    }
                    for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) d |= hp.hdr.dst_thr[j];
                    didx[tid] = d & loc_mask;
                    for (int v = 0; v < QSB_NV; v++) {
                        uint64_t gi = d;
                        for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) gi |= hp.hdr.dst_vec[b];
                        for (int l = 0; l < L; l++) {
                            const cd val = regs[((size_t)tid * QSB_NV + v) * L + l];
                            if (hp.fused_swap) { if (!outs) { rep.bad_slots++; continue; } outs[gi >> nloc][(gi & loc_mask) | (uint64_t)l] = val; }
                            else st[(gi & loc_mask) | (uint64_t)l] = val;
                        }
                    }
                }
                /* the last store must be as coalesced as the first load (the lanes carry the low DESTINATION bits) */
                rep.noncontig += lanes_not_contiguous(didx, f32, g_low_a > 0 ? g_low_a : (f32 ? 4 : 3));
                for (int b = 0; b < QSB_NVB; b++) if (f32 && (hp.hdr.dst_vec[b] & 1)) rep.noncontig++;
            } else {
                std::fill(written.begin(), written.end(), 0);
                for (int tid = 0; tid < QSB_THREADS; tid++) {
                    uint32_t sb = 0;
                    for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) sb ^= RD.thr[j].st;
                    for (int v = 0; v < QSB_NV; v++) {
                        uint32_t slot = sb;
                        for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) slot ^= RD.vec[b].st;
                        if (slot >= QSB_SLOTS) { rep.bad_slots++; continue; }
                        written[slot]++;
                        for (int l = 0; l < L; l++) smem[(size_t)slot * L + l] = regs[((size_t)tid * QSB_NV + v) * L + l];
                    }
                }
                for (int s = 0; s < QSB_SLOTS; s++) if (written[s] != 1) rep.bad_slots++;
                for (int w = 0; w < QSB_THREADS / phase_lanes; w++) for (int v = 0; v < QSB_NV; v++) {
                    std::set<uint32_t> banks;
                    for (int ln = 0; ln < phase_lanes; ln++) {
                        int tid = w * phase_lanes + ln;
                        uint32_t slot = 0;
                        for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) slot ^= RD.thr[j].st;
                        for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) slot ^= RD.vec[b].st;
                        banks.insert(slot & ((1u << nb) - 1));
                    }
                    int conflict = phase_lanes / (int)banks.size();
                    if (conflict > rep.max_conflict) rep.max_conflict = conflict;
                }
            }
        }
    }
}

/* state: 2^max(num_q, T) complex doubles (interleaved), physical == logical order on entry;
 * on exit amplitudes are in PHYSICAL order and perm_out[q] gives logical q -> physical bit. */
extern "C" int qsb_hostcheck_run(int num_q, int prec, int low_bits, const qsb_gate_t *gates, size_t n,
                                 double *state, int *report5, int8_t *perm_out)
{
    qsb_options_t opt; memset(&opt, 0, sizeof opt);
    opt.precision = prec; opt.low_bits = low_bits; opt.world_size = 1;
    if (getenv("QSB_HC_RES4")) opt.reserved[4] = atoi(getenv("QSB_HC_RES4"));   /* planner A/B knobs for the tests */
    if (getenv("QSB_HC_RES6")) opt.reserved[6] = atoi(getenv("QSB_HC_RES6"));
    g_low_a = low_bits;
    const int T = tiled_min_local_bits(prec, &opt);
    const int nloc = std::max(num_q, T);
    std::vector<COp> cops; double gph[2];
    int rc = qsb_canonicalise(gates, n, num_q, cops, gph);
    if (rc) return rc;
    BitPerm id; for (int q = 0; q < 64; q++) id.pos[q] = (int8_t)q;
    TiledPlan plan;
    rc = make_plan(num_q, prec, 0, nloc, 0, &opt, id, cops, gph, &plan);
    if (rc) return rc;
    std::vector<cd> st((size_t)1 << nloc);
    memcpy((void *)st.data(), state, sizeof(cd) * st.size());
    Report rep; memset(&rep, 0, sizeof rep);
    rep.max_conflict = 1;
    for (const HostPass &hp : plan.passes) { run_pass(hp, prec == QSB_F32, nloc, st, rep); rep.passes++; rep.rounds += (int)hp.hdr.n_rounds; }
    memcpy(state, st.data(), sizeof(cd) * st.size());
    report5[0] = rep.max_conflict; report5[1] = rep.bad_slots; report5[2] = rep.noncontig; report5[3] = rep.passes; report5[4] = rep.rounds;
    for (int q = 0; q < 64; q++) perm_out[q] = plan.end_perm.pos[q];
    return 0;
}


/* ---- step-wise interface for sharded schedules (multi-rank tests) ------------------------------
 * The caller owns one local shard per rank (2^nloc complex doubles) and performs the exchanges
 * itself (numpy in one process, or torch.distributed/gloo across processes). */
struct HcPlan { TiledPlan plan; int prec; int nloc; int low_bits; Report rep; };

extern "C" void *qsb_hostcheck_plan(int num_q, int prec, int low_bits, int world, int rank, int swap_min_ops,
                                    const qsb_gate_t *gates, size_t n)
{
    qsb_options_t opt; memset(&opt, 0, sizeof opt);
    opt.precision = prec; opt.low_bits = low_bits; opt.world_size = world; opt.rank = rank; opt.reserved[0] = swap_min_ops;
    if (getenv("QSB_HC_RES4")) opt.reserved[4] = atoi(getenv("QSB_HC_RES4"));   /* planner A/B knobs for the tests */
    if (getenv("QSB_HC_RES6")) opt.reserved[6] = atoi(getenv("QSB_HC_RES6"));
    int g = 0; while ((1 << g) < world) g++;
    const int T = tiled_min_local_bits(prec, &opt);
    const int nloc = std::max(num_q - g, T);
    std::vector<COp> cops; double gph[2];
    if (qsb_canonicalise(gates, n, num_q, cops, gph)) return nullptr;
    BitPerm id; for (int q = 0; q < 64; q++) id.pos[q] = (int8_t)q;
    HcPlan *h = new HcPlan();
    h->prec = prec; h->nloc = nloc; h->low_bits = low_bits; memset(&h->rep, 0, sizeof h->rep); h->rep.max_conflict = 1;
    if (make_plan(num_q, prec, g, nloc, rank, &opt, id, cops, gph, &h->plan)) { delete h; return nullptr; }
    return h;
}
/* like qsb_hostcheck_plan, with the exchanges fused into the preceding pass (peer scatter) */
extern "C" void *qsb_hostcheck_plan_fused(int num_q, int prec, int low_bits, int world, int rank, int swap_min_ops,
                                          const qsb_gate_t *gates, size_t n)
{
    qsb_options_t opt; memset(&opt, 0, sizeof opt);
    opt.precision = prec; opt.low_bits = low_bits; opt.world_size = world; opt.rank = rank; opt.reserved[0] = swap_min_ops;
    opt.reserved[5] = 1;
    int g = 0; while ((1 << g) < world) g++;
    const int T = tiled_min_local_bits(prec, &opt);
    const int nloc = std::max(num_q - g, T);
    std::vector<COp> cops; double gph[2];
    if (qsb_canonicalise(gates, n, num_q, cops, gph)) return nullptr;
    BitPerm id; for (int q = 0; q < 64; q++) id.pos[q] = (int8_t)q;
    HcPlan *h = new HcPlan();
    h->prec = prec; h->nloc = nloc; h->low_bits = low_bits; memset(&h->rep, 0, sizeof h->rep); h->rep.max_conflict = 1;
    if (make_plan(num_q, prec, g, nloc, rank, &opt, id, cops, gph, &h->plan)) { delete h; return nullptr; }
    return h;
}
/* 0 ordinary pass, 1 exchange marker (all-to-all of chunks), 2 fused-exchange pass */
extern "C" int qsb_hostcheck_step_kind(void *h, int i)
{
    const HostPass &hp = ((HcPlan *)h)->plan.passes[i];
    return hp.is_swap ? 1 : hp.fused_swap ? 2 : 0;
}
/* fused-exchange pass of one rank: reads `state` (this rank's shard), writes into outs[r] (rank r's NEW shard) */
extern "C" int qsb_hostcheck_run_step_fused(void *hv, int i, const double *state, double *const *outs)
{
    HcPlan *h = (HcPlan *)hv;
    const HostPass &hp = h->plan.passes[i];
    if (!hp.fused_swap) return 1;
    std::vector<cd> st((size_t)1 << h->nloc);
    memcpy((void *)st.data(), state, sizeof(cd) * st.size());
    g_low_a = h->low_bits;
    run_pass(hp, h->prec == QSB_F32, h->nloc, st, h->rep, (cd *const *)outs);
    return 0;
}
/* properties of pass i for the race test: out[0] rounds, out[1] out_of_place, out[2] 1 if the pass relocates a qubit
 * inside its tile (tile_dst != tile_src), out[3] the QSB_PASS_SYNC_SCATTER flag of the serialised descriptor */
extern "C" int qsb_hostcheck_step_props(void *hv, int i, int *out4)
{
    const HostPass &hp = ((HcPlan *)hv)->plan.passes[i];
    if (hp.is_swap) return 1;
    out4[0] = (int)hp.rounds.size(); out4[1] = (int)hp.hdr.out_of_place;
    int moved = 0; for (int j = 0; j < hp.T; j++) moved |= hp.tile_src[j] != hp.tile_dst[j];
    out4[2] = moved;
    GPass gp; memcpy(&gp, hp.blob.data(), sizeof gp);
    out4[3] = (int)(gp.flags & QSB_PASS_SYNC_SCATTER);
    return 0;
}
extern "C" int qsb_hostcheck_num_steps(void *h) { return (int)((HcPlan *)h)->plan.passes.size(); }
extern "C" int qsb_hostcheck_step_is_swap(void *h, int i) { return ((HcPlan *)h)->plan.passes[i].is_swap ? 1 : 0; }
extern "C" int qsb_hostcheck_nloc(void *h) { return ((HcPlan *)h)->nloc; }
extern "C" int qsb_hostcheck_run_step(void *hv, int i, double *state)
{
    HcPlan *h = (HcPlan *)hv;
    const HostPass &hp = h->plan.passes[i];
    if (hp.is_swap) return 1;
    std::vector<cd> st((size_t)1 << h->nloc);
    memcpy((void *)st.data(), state, sizeof(cd) * st.size());
    g_low_a = h->low_bits;
    run_pass(hp, h->prec == QSB_F32, h->nloc, st, h->rep);
    memcpy(state, st.data(), sizeof(cd) * st.size());
    return 0;
}
extern "C" void qsb_hostcheck_finish(void *hv, int *report5, int8_t *perm_out)
{
    HcPlan *h = (HcPlan *)hv;
    report5[0] = h->rep.max_conflict; report5[1] = h->rep.bad_slots; report5[2] = h->rep.noncontig;
    int sw = 0, fs = 0; for (auto &p : h->plan.passes) { sw += p.is_swap; fs += p.fused_swap; }
    report5[3] = (int)h->plan.passes.size() - sw; report5[4] = sw + fs;
    for (int q = 0; q < 64; q++) perm_out[q] = h->plan.end_perm.pos[q];
    delete h;
}

/* tile bits of this build (the tile geometry is a build-time switch, tiled.h QSB_TB) */
extern "C" int qsb_hostcheck_tile_bits(int prec) { return tiled_min_local_bits(prec, nullptr); }

/* plan statistics for tuning: per pass "rounds ops | histogram of op codes" on stdout */
extern "C" void qsb_hostcheck_describe(void *hv)
{
    HcPlan *h = (HcPlan *)hv;
    static const char *nm[16] = {"?", "MAT_R", "MAT_I", "MAT_G", "MATP_R", "MATP_G", "U", "UI", "?", "DIAG_V", "DIAG_ALL", "DIAG_GEN", "TPHASE", "XDEF", "?", "?"};
    int tot[16] = {0}, totmux = 0, k = 0;
    for (const HostPass &hp : h->plan.passes) {
        if (hp.is_swap) { printf("pass %d: SWAP\n", k++); continue; }
        int cnt[16] = {0}, mux = 0, flagged = 0;
        for (const HostOp &o : hp.ops) { cnt[o.kind & 0xf]++; tot[o.kind & 0xf]++; if ((o.kind >> 16) & 1) { mux++; totmux++; } }
        for (const DevRound &r : hp.rounds) flagged += r.flags & 1;
        printf("pass %d: rounds %d flagged %d src_ops %d ops %zu mux %d |", k++, (int)hp.hdr.n_rounds, flagged, hp.n_source_ops, hp.ops.size(), mux);
        for (int c = 0; c < 16; c++) if (cnt[c]) printf(" %s:%d", nm[c], cnt[c]);
        printf(" | per-round ops:");
        for (size_t r = 0; r < hp.round_op_count.size(); r++) printf(" %u", hp.round_op_count[r]);
        printf("\n");
    }
    printf("total:");
    for (int c = 0; c < 16; c++) if (tot[c]) printf(" %s:%d", nm[c], tot[c]);
    printf(" mux:%d\n", totmux);
}

/* tuning aid: classify the full-form real ops of a plan (bare X / multiplexer / other) */
extern "C" void qsb_hostcheck_describe_full(void *hv)
{
    HcPlan *h = (HcPlan *)hv;
    int nx = 0, nmux = 0, nother = 0, nx_cond = 0;
    for (const HostPass &hp : h->plan.passes) for (const HostOp &o : hp.ops) {
        const int code = o.kind & 0xff; const bool mux = (o.kind >> 16) & 1;
        if (code != OP_MAT_R) continue;
        const bool isx = o.c[0][0][1] == 1 && o.c[0][1][1] == 1 && o.c[0][2][1] == 0 && o.c[0][3][1] == 0;
        if (mux) nmux++; else if (isx) { nx++; if (o.tmask) nx_cond++; } else nother++;
    }
    printf("MAT_R: bare X %d (conditional %d), mux %d, other %d\n", nx, nx_cond, nmux, nother);
}
