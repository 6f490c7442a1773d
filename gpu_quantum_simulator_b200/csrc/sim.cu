/*
 * sim.cu -- handle lifecycle, the reference-style SWEEP mode, readout and the
 * C ABI of libqsim_b200.  The fused TILED mode (the product path) lives in
 * tiled_plan.cpp / tiled_kernel.cu and is driven from qsb_plan_create /
 * qsb_execute below.
 *
 * Reference functions replaced here (all in /root/reference/):
 *   init_state_vector            naive.cu:64-70      -> k_init
 *   kernel_gate / kernel_gate_2  naive.cu:72-95      -> k_sweep_mat
 *   kernel_cnot                  naive.cu:97-122     -> k_sweep_x
 *   (compute_state_cumulative_distribution / measurement, quantum_simulator.c:256-283 -> readout.cu)
 * None of the reference's code is reused: indices are 64-bit throughout (the
 * reference's `int th_id` caps it at 31 qubits), control masks are generic,
 * and the fp32 state uses the packed pair-interleaved layout of common.cuh.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "sim.h"
#include "tiled.h"

#define QSB_VERSION "qsim-b200 0.1 (sm_100a)"

static double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

/* ===================================================================== kernels */

template <typename R>
__global__ void k_init(R *st, uint64_t n_amps, int is_rank0)
{
    /* whole buffer is zero except amplitude 0 on rank 0; works for both layouts
     * because re(0) == 0 in both. */
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = i; k < 2 * n_amps; k += stride) st[k] = (k == 0 && is_rank0) ? R(1) : R(0);
}

/* insert a zero bit at position `pos` */
__device__ __forceinline__ uint64_t insert_zero(uint64_t x, int pos)
{
    uint64_t lo = x & ((1ULL << pos) - 1);
    return ((x >> pos) << (pos + 1)) | lo;
}

struct Mat8 { double m[8]; };

/* one sweep: 2x2 matrix on physical bit `t` where (global index & ctrl) == ctrl */
template <typename R>
__global__ void k_sweep_mat(R *st, uint64_t n_pairs, int t, uint64_t ctrl, uint64_t rank_bits, Mat8 M)
{
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const R m0 = (R)M.m[0], m1 = (R)M.m[1], m2 = (R)M.m[2], m3 = (R)M.m[3];
    const R m4 = (R)M.m[4], m5 = (R)M.m[5], m6 = (R)M.m[6], m7 = (R)M.m[7];
    for (; p < n_pairs; p += stride) {
        uint64_t i = insert_zero(p, t), j = i | (1ULL << t);
        if (((i | rank_bits) & ctrl) != ctrl) continue;
        R ar = st[Lay<R>::re(i)], ai = st[Lay<R>::im(i)];
        R br = st[Lay<R>::re(j)], bi = st[Lay<R>::im(j)];
        st[Lay<R>::re(i)] = (ar * m0 - ai * m1) + (br * m2 - bi * m3);
        st[Lay<R>::im(i)] = (ar * m1 + ai * m0) + (br * m3 + bi * m2);
        st[Lay<R>::re(j)] = (ar * m4 - ai * m5) + (br * m6 - bi * m7);
        st[Lay<R>::im(j)] = (ar * m5 + ai * m4) + (br * m7 + bi * m6);
    }
}

/* one sweep: conditional swap (X / CX / CCX) */
template <typename R>
__global__ void k_sweep_x(R *st, uint64_t n_pairs, int t, uint64_t ctrl, uint64_t rank_bits)
{
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; p < n_pairs; p += stride) {
        uint64_t i = insert_zero(p, t), j = i | (1ULL << t);
        if (((i | rank_bits) & ctrl) != ctrl) continue;
        R ar = st[Lay<R>::re(i)], ai = st[Lay<R>::im(i)];
        st[Lay<R>::re(i)] = st[Lay<R>::re(j)]; st[Lay<R>::im(i)] = st[Lay<R>::im(j)];
        st[Lay<R>::re(j)] = ar; st[Lay<R>::im(j)] = ai;
    }
}

/* one sweep: multiply by a phase where (global index & mask) == mask */
template <typename R>
__global__ void k_sweep_phase(R *st, uint64_t n_amps, uint64_t mask, uint64_t rank_bits, double pr_, double pi_)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const R pr = (R)pr_, pi = (R)pi_;
    for (; i < n_amps; i += stride) {
        if (((i | rank_bits) & mask) != mask) continue;
        R ar = st[Lay<R>::re(i)], ai = st[Lay<R>::im(i)];
        st[Lay<R>::re(i)] = ar * pr - ai * pi;
        st[Lay<R>::im(i)] = ar * pi + ai * pr;
    }
}

/* logical -> physical index through the qubit permutation */
struct PermArg { int8_t pos[64]; int n; };
__device__ __forceinline__ uint64_t to_phys(uint64_t logical, const PermArg &P)
{
    uint64_t r = 0;
    for (int q = 0; q < P.n; q++) r |= ((logical >> q) & 1ULL) << P.pos[q];
    return r;
}

/* Export kernels: a block handles runs of 256 consecutive logical indices.  Only the low byte of the index varies
 * inside a run, so the physical index is  hi(run) | lo(thread):  the per-thread part is mapped once per kernel, the
 * per-run part is uniform (round 1 walked up to 40 permutation entries for EVERY amplitude). */
__device__ __forceinline__ uint64_t to_phys_bits(uint64_t logical, const PermArg &P, int q0, int q1)
{
    uint64_t r = 0;
    for (int q = q0; q < q1 && q < P.n; q++) r |= ((logical >> q) & 1ULL) << P.pos[q];
    return r;
}
/* export amplitudes [first, first+count) in logical order as O = double (fp64 re, im) or the state precision */
template <typename R, typename O>
__global__ void k_export(const R *st, O *out, uint64_t first, uint64_t count, PermArg P, uint64_t loc_mask)
{
    const uint64_t n_runs = (count + (first & 255) + 255) / 256;
    const uint64_t lo = to_phys_bits(threadIdx.x, P, 0, 8);
    for (uint64_t run = blockIdx.x; run < n_runs; run += gridDim.x) {
        const uint64_t base = (first & ~255ULL) + run * 256;       /* logical index of the run's first element */
        const uint64_t idx = base + threadIdx.x;
        if (idx < first || idx >= first + count) continue;
        const uint64_t i = (to_phys_bits(base, P, 8, 64) | lo) & loc_mask;
        const uint64_t k = idx - first;
        out[2 * k] = (O)st[Lay<R>::re(i)];
        out[2 * k + 1] = (O)st[Lay<R>::im(i)];
    }
}
template <typename R>
__global__ void k_import(R *st, const double *in, uint64_t first, uint64_t count, PermArg P, uint64_t loc_mask)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    uint64_t i = to_phys(first + k, P) & loc_mask;
    st[Lay<R>::re(i)] = (R)in[2 * k];
    st[Lay<R>::im(i)] = (R)in[2 * k + 1];
}
template <typename R>
__global__ void k_probs(const R *st, double *out, uint64_t first, uint64_t count, PermArg P, uint64_t loc_mask)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    uint64_t i = to_phys(first + k, P) & loc_mask;
    double r = (double)st[Lay<R>::re(i)], m = (double)st[Lay<R>::im(i)];
    out[k] = r * r + m * m;
}

/* per-block partial: sum |a|^2 (fp64) and arg max over the PHYSICAL local index */
struct RedOut { double sum; double best; uint64_t idx; uint64_t pad; };
template <typename R>
__global__ void k_norm_argmax(const R *st, uint64_t n_amps, RedOut *out)
{
    __shared__ double s_sum[256];
    __shared__ double s_best[256];
    __shared__ uint64_t s_idx[256];
    double sum = 0.0, best = -1.0;
    uint64_t bidx = 0;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
        double r = (double)st[Lay<R>::re(i)], m = (double)st[Lay<R>::im(i)];
        double p = r * r + m * m;
        sum += p;
        if (p > best) { best = p; bidx = i; }
    }
    s_sum[threadIdx.x] = sum; s_best[threadIdx.x] = best; s_idx[threadIdx.x] = bidx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
            if (s_best[threadIdx.x + o] > s_best[threadIdx.x] ||
                (s_best[threadIdx.x + o] == s_best[threadIdx.x] && s_idx[threadIdx.x + o] < s_idx[threadIdx.x])) {
                s_best[threadIdx.x] = s_best[threadIdx.x + o]; s_idx[threadIdx.x] = s_idx[threadIdx.x + o];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[blockIdx.x].sum = s_sum[0]; out[blockIdx.x].best = s_best[0]; out[blockIdx.x].idx = s_idx[0]; }
}

/* ================================================================ host helpers */

static PermArg perm_arg(const qsb_sim *s)
{
    PermArg P; memset(&P, 0, sizeof P);
    P.n = s->n;
    for (int q = 0; q < s->n; q++) P.pos[q] = s->perm.pos[q];
    return P;
}

static int grid_for(uint64_t work, int block) { return (int)std::min<uint64_t>((work + block - 1) / block, 148ULL * 32); }

/* ================================================================== lifecycle */

extern "C" void qsb_options_default(qsb_options_t *o)
{
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->precision = QSB_F32; o->device = -1; o->mode = QSB_MODE_TILED;
    o->rank = 0; o->world_size = 1; o->use_graph = 0;
}

/* the tile geometry is a build-time switch (tiled.h): keep it visible in every bench line and bug report */
extern "C" const char *qsb_version(void)
{
    static char v[128];
    if (!v[0]) snprintf(v, sizeof v, "%s; %d threads x %d CTAs/SM, tile bits f32/f64 = %d/%d", QSB_VERSION, QSB_THREADS, QSB_CTAS_PER_SM, QSB_T_F32, QSB_T_F64);
    return v;
}

static int ilog2(int v) { int r = 0; while ((1 << r) < v) r++; return r; }

extern "C" int qsb_create(qsb_t **out, int num_qubits, const qsb_options_t *opt_in)
{
    if (!out) { qsb_set_error("qsb_create: null out"); return QSB_ERR_ARG; }
    *out = nullptr;
    qsb_options_t opt;
    if (opt_in) opt = *opt_in; else qsb_options_default(&opt);
    if (opt.precision == 0) opt.precision = QSB_F32;
    if (opt.world_size <= 0) opt.world_size = 1;
    if (opt.precision != QSB_F32 && opt.precision != QSB_F64) { qsb_set_error("precision must be 32 or 64"); return QSB_ERR_ARG; }
    if (num_qubits < 1 || num_qubits > 40) { qsb_set_error("num_qubits %d out of range 1..40", num_qubits); return QSB_ERR_ARG; }
    if (opt.world_size & (opt.world_size - 1)) { qsb_set_error("world_size must be a power of two"); return QSB_ERR_ARG; }
    if (opt.rank < 0 || opt.rank >= opt.world_size) { qsb_set_error("rank %d outside world of %d", opt.rank, opt.world_size); return QSB_ERR_ARG; }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        qsb_set_error("no CUDA device (%s): libqsim_b200 has no CPU path", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return QSB_ERR_NOGPU;
    }
    qsb_sim *s = new qsb_sim();
    s->opt = opt;
    s->n = num_qubits; s->prec = opt.precision; s->rank = opt.rank; s->world = opt.world_size;
    s->g = ilog2(opt.world_size);
    if (opt.device >= 0) { s->device = opt.device; } else { cudaGetDevice(&s->device); }
    if (cudaSetDevice(s->device) != cudaSuccess) { qsb_set_error("cannot select CUDA device %d", s->device); (void)cudaGetLastError(); delete s; return QSB_ERR_CUDA; }
    const int min_loc = tiled_min_local_bits(s->prec, &s->opt);
    if (s->g > 0 && num_qubits - s->g < min_loc + s->g) {
        qsb_set_error("%d qubits are too few to shard over %d ranks", num_qubits, s->world); delete s; return QSB_ERR_ARG;
    }
    s->nloc = std::max(num_qubits - s->g, min_loc); /* small registers are zero-padded up to one tile */
    s->nphys = s->nloc + s->g;
    s->state_bytes = ((size_t)1 << s->nloc) * amp_bytes(s->prec);
    e = cudaMalloc(&s->state, s->state_bytes);
    if (e != cudaSuccess) {
        qsb_set_error("Malloc error: cudaMalloc of %zu bytes failed (%s)", s->state_bytes, cudaGetErrorString(e));
        (void)cudaGetLastError(); delete s; return QSB_ERR_NOMEM;
    }
    /* every failure from here on releases the handle and the state buffer */
#define QSB_CREATE_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        qsb_set_error("%s in %s at line %d", cudaGetErrorString(e_), __FILE__, __LINE__); (void)cudaGetLastError(); qsb_destroy(s); return QSB_ERR_CUDA; } } while (0)
    QSB_CREATE_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    QSB_CREATE_CUDA(cudaStreamCreateWithFlags(&s->dl_stream, cudaStreamNonBlocking));
    QSB_CREATE_CUDA(cudaEventCreate(&s->ev0)); QSB_CREATE_CUDA(cudaEventCreate(&s->ev1));
    QSB_CREATE_CUDA(cudaEventCreate(&s->evx0)); QSB_CREATE_CUDA(cudaEventCreate(&s->evx1));
    for (int b = 0; b < 2; b++) {
        QSB_CREATE_CUDA(cudaEventCreateWithFlags(&s->dl_filled[b], cudaEventDisableTiming));
        QSB_CREATE_CUDA(cudaEventCreateWithFlags(&s->dl_copied[b], cudaEventDisableTiming));
    }
    s->staging_bytes = (size_t)64 << 20;
    QSB_CREATE_CUDA(cudaMalloc(&s->staging, s->staging_bytes));
    QSB_CREATE_CUDA(cudaMalloc(&s->d_scratch, 1 << 20));
#undef QSB_CREATE_CUDA
    int rc = qsb_reset(s);
    if (rc) { qsb_destroy(s); return rc; }
    *out = s;
    return QSB_OK;
}

extern "C" void qsb_destroy(qsb_t *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    tiled_comm_destroy(s);
    if (s->state) cudaFree(s->state);
    if (s->state2) cudaFree(s->state2);
    if (s->staging) cudaFree(s->staging);
    if (s->d_scratch) cudaFree(s->d_scratch);
    for (cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->evx0) cudaEventDestroy(s->evx0);
    if (s->evx1) cudaEventDestroy(s->evx1);
    for (int b = 0; b < 2; b++) { if (s->dl_filled[b]) cudaEventDestroy(s->dl_filled[b]); if (s->dl_copied[b]) cudaEventDestroy(s->dl_copied[b]); }
    if (s->dl_stream) cudaStreamDestroy(s->dl_stream);
    for (int i = 0; i < 4; i++) if (s->copy_stream[i]) cudaStreamDestroy(s->copy_stream[i]);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int qsb_reset(qsb_t *s)
{
    if (!s) { qsb_set_error("null handle"); return QSB_ERR_ARG; }
    QSB_CUDA(cudaSetDevice(s->device));
    for (int q = 0; q < 64; q++) s->perm.pos[q] = (int8_t)q;
    /* logical qubits >= nloc (if any) sit on the rank bits: identity covers it since nphys = nloc + g */
    uint64_t n_amps = 1ULL << s->nloc;
    if (s->prec == QSB_F32) k_init<float><<<grid_for(2 * n_amps, 256), 256, 0, s->stream>>>((float *)s->state, n_amps, s->rank == 0);
    else k_init<double><<<grid_for(2 * n_amps, 256), 256, 0, s->stream>>>((double *)s->state, n_amps, s->rank == 0);
    QSB_CUDA(cudaGetLastError());
    QSB_CUDA(cudaStreamSynchronize(s->stream));
    return QSB_OK;
}

extern "C" int qsb_num_qubits(const qsb_t *s) { return s ? s->n : QSB_ERR_ARG; }
extern "C" int qsb_precision(const qsb_t *s) { return s ? s->prec : QSB_ERR_ARG; }

/* =================================================================== hot path */

static int build_plan(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                      const qsb_gate_t *gates, size_t ngates, bool with_device, qsb_plan **out)
{
    qsb_plan *p = new qsb_plan();
    p->mode = opt->mode; p->n = n; p->prec = prec; p->world = 1 << g;
    double t0 = now_ms();
    int rc = qsb_canonicalise(gates, ngates, n, p->cops, p->gphase);
    if (rc) { delete p; return rc; }
    p->stats.source_gates = ngates;
    const uint64_t pass_bytes = 2ULL * ((uint64_t)1 << nloc) * amp_bytes(prec);
    if (p->mode == QSB_MODE_SWEEP) {
        if (g > 0) { qsb_set_error("SWEEP mode is single-GPU only"); delete p; return QSB_ERR_ARG; }
        if (!(p->gphase[0] == 1.0 && p->gphase[1] == 0.0)) {
            COp c; memset(&c, 0, sizeof c); c.kind = C_PHASE; c.ctrl = 0; c.target = -1; c.m[0] = p->gphase[0]; c.m[1] = p->gphase[1];
            p->cops.push_back(c);
        }
        p->stats.device_ops = p->cops.size();
        p->stats.passes = (uint32_t)p->cops.size();
        p->stats.kernel_launches = (uint32_t)p->cops.size();
        uint64_t b = 0;
        for (auto &c : p->cops) b += (c.kind == C_X && c.ctrl) ? pass_bytes / 2 : pass_bytes; /* a bare CX touches half the state */
        p->stats.bytes_moved = b;
    } else if (p->mode == QSB_MODE_DENSE) {
        if (g > 0) { qsb_set_error("DENSE mode is single-GPU only"); delete p; return QSB_ERR_ARG; }
        rc = dense_fuse(p->cops, p->gphase, n, opt->tile_bits > 0 ? opt->tile_bits : 4, p->dense);
        if (rc) { delete p; return rc; }
        p->stats.device_ops = p->dense.size();
        p->stats.passes = p->stats.kernel_launches = (uint32_t)p->dense.size();
        p->stats.bytes_moved = p->dense.size() * pass_bytes;
    } else {
        rc = tiled_plan_build(n, prec, g, nloc, rank, opt, start, p->cops, p->gphase, with_device, &p->tiled, &p->stats);
        if (rc) { delete p; return rc; }
    }
    p->stats.plan_ms = now_ms() - t0;
    *out = p;
    return QSB_OK;
}

extern "C" int qsb_plan_create(qsb_t *s, const qsb_gate_t *gates, size_t n, qsb_plan_t **out)
{
    if (!s || !out || (n && !gates)) { qsb_set_error("qsb_plan_create: null argument"); return QSB_ERR_ARG; }
    QSB_CUDA(cudaSetDevice(s->device));
    return build_plan(s->n, s->prec, s->g, s->nloc, s->rank, &s->opt, s->perm, gates, n, true, out);
}

extern "C" int qsb_plan_dry_run(int num_qubits, const qsb_options_t *opt_in, const qsb_gate_t *gates, size_t n, qsb_run_stats_t *out)
{
    if (!out || (n && !gates)) { qsb_set_error("qsb_plan_dry_run: null argument"); return QSB_ERR_ARG; }
    if (num_qubits < 1 || num_qubits > 40) { qsb_set_error("num_qubits %d out of range 1..40", num_qubits); return QSB_ERR_ARG; }
    qsb_options_t opt;
    if (opt_in) opt = *opt_in; else qsb_options_default(&opt);
    if (opt.precision == 0) opt.precision = QSB_F32;
    if (opt.world_size <= 0) opt.world_size = 1;
    if (opt.world_size & (opt.world_size - 1)) { qsb_set_error("world_size must be a power of two"); return QSB_ERR_ARG; }
    int g = ilog2(opt.world_size);
    /* the exchange flavour qsb_comm_init chooses when the peer shards can be mapped (direct fused scatter) */
    if (g > 0 && opt.reserved[5] == 0) opt.reserved[5] = 1;
    int nloc = std::max(num_qubits - g, tiled_min_local_bits(opt.precision, &opt));
    BitPerm id; for (int q = 0; q < 64; q++) id.pos[q] = (int8_t)q;
    qsb_plan *p = nullptr;
    int rc = build_plan(num_qubits, opt.precision, g, nloc, opt.rank, &opt, id, gates, n, false, &p);
    if (rc) return rc;
    *out = p->stats;
    qsb_plan_destroy(p);
    return QSB_OK;
}

extern "C" void qsb_plan_destroy(qsb_plan_t *p)
{
    if (!p) return;
    if (p->tiled) tiled_plan_free(p->tiled);
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    delete p;
}

extern "C" int qsb_plan_stats(const qsb_plan_t *p, qsb_run_stats_t *out)
{
    if (!p || !out) { qsb_set_error("null argument"); return QSB_ERR_ARG; }
    *out = p->stats;
    return QSB_OK;
}

template <typename R>
static int sweep_execute(qsb_sim *s, const qsb_plan *p)
{
    R *st = (R *)s->state;
    const uint64_t n_amps = 1ULL << s->nloc;
    const uint64_t rank_bits = (uint64_t)s->rank << s->nloc;
    for (const COp &c : p->cops) {
        uint64_t ctrl = 0;
        for (int q = 0; q < s->n; q++) if ((c.ctrl >> q) & 1) ctrl |= 1ULL << s->perm.pos[q];
        if (c.kind == C_PHASE) {
            k_sweep_phase<R><<<grid_for(n_amps, 256), 256, 0, s->stream>>>(st, n_amps, ctrl, rank_bits, c.m[0], c.m[1]);
        } else {
            int t = s->perm.pos[c.target];
            if (t >= s->nloc) { qsb_set_error("SWEEP mode: target on a global qubit"); return QSB_ERR_ARG; }
            if (c.kind == C_X) k_sweep_x<R><<<grid_for(n_amps / 2, 256), 256, 0, s->stream>>>(st, n_amps / 2, t, ctrl, rank_bits);
            else {
                Mat8 M; memcpy(M.m, c.m, sizeof M.m);
                k_sweep_mat<R><<<grid_for(n_amps / 2, 256), 256, 0, s->stream>>>(st, n_amps / 2, t, ctrl, rank_bits, M);
            }
        }
    }
    QSB_CUDA(cudaGetLastError());
    return QSB_OK;
}

int tiled_prepare_capture(qsb_sim *s);

/* use_graph: capture the pass launches of a single-GPU tiled plan into a CUDA graph (once per plan and
 * state buffer).  Pays off when one plan is executed many times on a small register, where a pass is
 * shorter than its launch (the regime of the reference's own 5-22 qubit benchmarks). */
static int capture_plan(qsb_sim *s, qsb_plan *p)
{
    int rc = tiled_prepare_capture(s);
    if (rc) return rc;
    if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
    const BitPerm before = s->perm;
    QSB_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeRelaxed));
    rc = tiled_execute(s, p->tiled);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    s->perm = before;                       /* nothing ran yet */
    if (rc) { if (graph) cudaGraphDestroy(graph); (void)cudaGetLastError(); return rc; }
    if (e != cudaSuccess) { qsb_set_error("%s while capturing the plan", cudaGetErrorString(e)); (void)cudaGetLastError(); return QSB_ERR_CUDA; }
    e = cudaGraphInstantiate(&p->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { p->graph_exec = nullptr; qsb_set_error("%s while instantiating the plan graph", cudaGetErrorString(e)); (void)cudaGetLastError(); return QSB_ERR_CUDA; }
    p->graph_state = s->state;
    return QSB_OK;
}

extern "C" int qsb_execute(qsb_t *s, qsb_plan_t *p)
{
    if (!s || !p) { qsb_set_error("qsb_execute: null argument"); return QSB_ERR_ARG; }
    if (p->n != s->n || p->prec != s->prec || p->world != s->world) { qsb_set_error("plan was built for a different machine"); return QSB_ERR_ARG; }
    /* fused passes leave the qubits permuted, and a tiled plan addresses the layout it was made for */
    if (p->mode == QSB_MODE_TILED && !tiled_plan_starts_at(p->tiled, s->perm)) {
        qsb_set_error("plan was built for another qubit layout: the state has moved since qsb_plan_create "
                      "(a plan runs once per layout: qsb_reset / qsb_load_state to its layout, or plan again)");
        return QSB_ERR_ARG;
    }
    QSB_CUDA(cudaSetDevice(s->device));
    const bool graphed = s->opt.use_graph && p->mode == QSB_MODE_TILED && s->world == 1;
    if (graphed && (!p->graph_exec || p->graph_state != s->state)) {
        int rc = capture_plan(s, p);
        if (rc) return rc;
    }
    QSB_CUDA(cudaEventRecord(s->ev0, s->stream));
    int rc;
    if (graphed) {
        rc = QSB_OK;
        cudaError_t e = cudaGraphLaunch(p->graph_exec, s->stream);
        if (e != cudaSuccess) { qsb_set_error("%s while launching the plan graph", cudaGetErrorString(e)); rc = QSB_ERR_CUDA; }
        else tiled_plan_end_perm(p->tiled, &s->perm);
    }
    else if (p->mode == QSB_MODE_SWEEP) rc = (s->prec == QSB_F32) ? sweep_execute<float>(s, p) : sweep_execute<double>(s, p);
    else if (p->mode == QSB_MODE_DENSE) rc = dense_execute(s, p->dense);
    else rc = tiled_execute(s, p->tiled);
    if (rc) { cudaStreamSynchronize(s->stream); (void)cudaGetLastError(); return rc; }
    QSB_CUDA(cudaEventRecord(s->ev1, s->stream));
    QSB_CUDA(cudaStreamSynchronize(s->stream));
    float ms = 0;
    QSB_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->last = p->stats;
    s->last.device_ms = ms;
    s->last.exchange_ms = (p->mode == QSB_MODE_TILED) ? tiled_last_exchange_ms(p->tiled) : 0.0;
    return QSB_OK;
}

extern "C" int qsb_apply_gates(qsb_t *s, const qsb_gate_t *gates, size_t n)
{
    qsb_plan_t *p = nullptr;
    int rc = qsb_plan_create(s, gates, n, &p);
    if (rc) return rc;
    rc = qsb_execute(s, p);
    qsb_plan_destroy(p);
    return rc;
}

extern "C" int qsb_last_run_stats(const qsb_t *s, qsb_run_stats_t *out)
{
    if (!s || !out) { qsb_set_error("null argument"); return QSB_ERR_ARG; }
    *out = s->last;
    return QSB_OK;
}

/* ===================================================================== readout */

static int check_range(const qsb_sim *s, uint64_t first, uint64_t count, uint64_t *lo_out)
{
    /* logical indices [first, first+count) must be < 2^n and live on this rank */
    const uint64_t total = 1ULL << s->n;
    if (first > total || count > total - first) { qsb_set_error("range [%llu, +%llu) exceeds 2^%d amplitudes", (unsigned long long)first, (unsigned long long)count, s->n); return QSB_ERR_ARG; }
    (void)lo_out;
    return QSB_OK;
}

/* For sharded states the rank owning logical index L is phys(L) >> nloc; callers
 * must ask each rank only for what it owns.  We verify on the host for the two
 * end points and all rank-bit patterns inside the range (cheap: g <= 6). */
static bool range_is_local(const qsb_sim *s, uint64_t first, uint64_t count)
{
    if (s->g == 0 || count == 0) return true;
    /* logical qubits mapped to rank bits */
    uint64_t lmask = 0; uint64_t want = 0;
    for (int q = 0; q < s->n; q++) if (s->perm.pos[q] >= s->nloc) {
        lmask |= 1ULL << q;
        if ((s->rank >> (s->perm.pos[q] - s->nloc)) & 1) want |= 1ULL << q;
    }
    /* every index in range must satisfy (idx & lmask) == want: check by walking blocks of the lowest global logical bit */
    int lowest = __builtin_ctzll(lmask);
    uint64_t blk = 1ULL << lowest;
    for (uint64_t i = first; i < first + count; i = ((i >> lowest) + 1) << lowest) {
        if ((i & lmask) != want) return false;
        if (blk == 0) break;
    }
    return ((first + count - 1) & lmask) == want;
}

/* Readout pipeline: the staging buffer is used as two halves; the export kernel fills one half on the compute
 * stream while the copy engine drains the other to the host on dl_stream.  With a pinned destination the copy is
 * a direct DMA at PCIe speed; pageable memory is staged by the runtime and is correspondingly slower. */
template <typename F>
static int pipelined_d2h(qsb_sim *s, char *dst, uint64_t count, size_t item_bytes, F fill /* (void *staging, uint64_t off, uint64_t c) -> int */)
{
    const size_t half = s->staging_bytes / 2;
    const uint64_t chunk = half / item_bytes;
    bool used[2] = {false, false};
    int b = 0;
    for (uint64_t off = 0; off < count; off += chunk, b ^= 1) {
        const uint64_t c = std::min(chunk, count - off);
        char *stg = (char *)s->staging + (size_t)b * half;
        if (used[b]) QSB_CUDA(cudaStreamWaitEvent(s->stream, s->dl_copied[b], 0));   /* this half has reached the host */
        int rc = fill(stg, off, c);
        if (rc) { cudaStreamSynchronize(s->dl_stream); return rc; }
        QSB_CUDA(cudaEventRecord(s->dl_filled[b], s->stream));
        QSB_CUDA(cudaStreamWaitEvent(s->dl_stream, s->dl_filled[b], 0));
        QSB_CUDA(cudaMemcpyAsync(dst + off * item_bytes, stg, c * item_bytes, cudaMemcpyDeviceToHost, s->dl_stream));
        QSB_CUDA(cudaEventRecord(s->dl_copied[b], s->dl_stream));
        used[b] = true;
    }
    QSB_CUDA(cudaStreamSynchronize(s->dl_stream));
    QSB_CUDA(cudaStreamSynchronize(s->stream));
    return QSB_OK;
}
static inline unsigned export_grid(uint64_t c) { return (unsigned)std::min<uint64_t>((c + 511) / 256, 148 * 16); }

template <typename R, typename O>
static int download_impl(qsb_sim *s, O *dst, uint64_t first, uint64_t count, const PermArg &P, uint64_t loc_mask)
{
    return pipelined_d2h(s, (char *)dst, count, 2 * sizeof(O), [&](void *stg, uint64_t off, uint64_t c) -> int {
        k_export<R, O><<<export_grid(c), 256, 0, s->stream>>>((const R *)s->state, (O *)stg, first + off, c, P, loc_mask);
        QSB_CUDA(cudaGetLastError());
        return QSB_OK;
    });
}

extern "C" int qsb_download(qsb_t *s, double *re_im, uint64_t first, uint64_t count)
{
    if (!s || (!re_im && count)) { qsb_set_error("qsb_download: null argument"); return QSB_ERR_ARG; }
    int rc = check_range(s, first, count, nullptr);
    if (rc) return rc;
    if (!range_is_local(s, first, count)) { qsb_set_error("range is not owned by rank %d", s->rank); return QSB_ERR_ARG; }
    QSB_CUDA(cudaSetDevice(s->device));
    const PermArg P = perm_arg(s);
    const uint64_t loc_mask = (1ULL << s->nloc) - 1;
    return s->prec == QSB_F32 ? download_impl<float, double>(s, re_im, first, count, P, loc_mask)
                              : download_impl<double, double>(s, re_im, first, count, P, loc_mask);
}

extern "C" int qsb_download_native(qsb_t *s, void *dst, uint64_t first, uint64_t count)
{
    if (!s || (!dst && count)) { qsb_set_error("qsb_download_native: null argument"); return QSB_ERR_ARG; }
    int rc = check_range(s, first, count, nullptr);
    if (rc) return rc;
    if (!range_is_local(s, first, count)) { qsb_set_error("range is not owned by rank %d", s->rank); return QSB_ERR_ARG; }
    QSB_CUDA(cudaSetDevice(s->device));
    const PermArg P = perm_arg(s);
    const uint64_t loc_mask = (1ULL << s->nloc) - 1;
    return s->prec == QSB_F32 ? download_impl<float, float>(s, (float *)dst, first, count, P, loc_mask)
                              : download_impl<double, double>(s, (double *)dst, first, count, P, loc_mask);
}

extern "C" int qsb_get_layout(const qsb_t *s, int8_t *perm64, int *nloc)
{
    if (!s || !perm64) { qsb_set_error("qsb_get_layout: null argument"); return QSB_ERR_ARG; }
    for (int q = 0; q < 64; q++) perm64[q] = q < s->n ? s->perm.pos[q] : (int8_t)-1;
    if (nloc) *nloc = s->nloc;
    return QSB_OK;
}

extern "C" int qsb_download_physical(qsb_t *s, double *re_im, uint64_t first, uint64_t count)
{
    if (!s || (!re_im && count)) { qsb_set_error("qsb_download_physical: null argument"); return QSB_ERR_ARG; }
    const uint64_t nl = 1ULL << s->nloc;
    if (first > nl || count > nl - first) { qsb_set_error("range exceeds the local shard"); return QSB_ERR_ARG; }
    QSB_CUDA(cudaSetDevice(s->device));
    PermArg P; memset(&P, 0, sizeof P);
    P.n = s->nloc;
    for (int q = 0; q < s->nloc; q++) P.pos[q] = (int8_t)q;   /* identity: physical order */
    return s->prec == QSB_F32 ? download_impl<float, double>(s, re_im, first, count, P, nl - 1)
                              : download_impl<double, double>(s, re_im, first, count, P, nl - 1);
}

extern "C" int qsb_upload(qsb_t *s, const double *re_im, uint64_t first, uint64_t count)
{
    if (!s || (!re_im && count)) { qsb_set_error("qsb_upload: null argument"); return QSB_ERR_ARG; }
    int rc = check_range(s, first, count, nullptr);
    if (rc) return rc;
    if (!range_is_local(s, first, count)) { qsb_set_error("range is not owned by rank %d", s->rank); return QSB_ERR_ARG; }
    QSB_CUDA(cudaSetDevice(s->device));
    const uint64_t chunk = s->staging_bytes / 16;
    PermArg P = perm_arg(s);
    const uint64_t loc_mask = (1ULL << s->nloc) - 1;
    for (uint64_t off = 0; off < count; off += chunk) {
        uint64_t c = std::min(chunk, count - off);
        QSB_CUDA(cudaMemcpyAsync(s->staging, re_im + 2 * off, c * 16, cudaMemcpyHostToDevice, s->stream));
        if (s->prec == QSB_F32) k_import<float><<<(unsigned)((c + 255) / 256), 256, 0, s->stream>>>((float *)s->state, (const double *)s->staging, first + off, c, P, loc_mask);
        else k_import<double><<<(unsigned)((c + 255) / 256), 256, 0, s->stream>>>((double *)s->state, (const double *)s->staging, first + off, c, P, loc_mask);
        QSB_CUDA(cudaGetLastError());
        QSB_CUDA(cudaStreamSynchronize(s->stream));
    }
    return QSB_OK;
}

/* physical local index -> logical global index */
static uint64_t to_logical(const qsb_sim *s, uint64_t phys_local)
{
    uint64_t phys = phys_local | ((uint64_t)s->rank << s->nloc);
    uint64_t L = 0;
    for (int q = 0; q < s->n; q++) L |= ((phys >> s->perm.pos[q]) & 1ULL) << q;
    return L;
}

extern "C" int qsb_norm_argmax(qsb_t *s, double *norm, uint64_t *argmax_idx, double *argmax_p)
{
    if (!s) { qsb_set_error("null handle"); return QSB_ERR_ARG; }
    QSB_CUDA(cudaSetDevice(s->device));
    const int blocks = 148 * 4;
    const uint64_t n_amps = 1ULL << s->nloc;
    RedOut *d = (RedOut *)s->d_scratch;
    if (s->prec == QSB_F32) k_norm_argmax<float><<<blocks, 256, 0, s->stream>>>((const float *)s->state, n_amps, d);
    else k_norm_argmax<double><<<blocks, 256, 0, s->stream>>>((const double *)s->state, n_amps, d);
    QSB_CUDA(cudaGetLastError());
    std::vector<RedOut> h(blocks);
    QSB_CUDA(cudaMemcpyAsync(h.data(), d, blocks * sizeof(RedOut), cudaMemcpyDeviceToHost, s->stream));
    QSB_CUDA(cudaStreamSynchronize(s->stream));
    double sum = 0, best = -1; uint64_t bidx = 0;
    for (auto &r : h) {
        sum += r.sum;
        uint64_t L = to_logical(s, r.idx);
        if (r.best > best || (r.best == best && L < bidx)) { best = r.best; bidx = L; }
    }
    if (norm) *norm = sum;
    if (argmax_idx) *argmax_idx = bidx;
    if (argmax_p) *argmax_p = best;
    return QSB_OK;
}

extern "C" int qsb_probabilities(qsb_t *s, double *p, uint64_t first, uint64_t count)
{
    if (!s || (!p && count)) { qsb_set_error("qsb_probabilities: null argument"); return QSB_ERR_ARG; }
    int rc = check_range(s, first, count, nullptr);
    if (rc) return rc;
    if (!range_is_local(s, first, count)) { qsb_set_error("range is not owned by rank %d", s->rank); return QSB_ERR_ARG; }
    QSB_CUDA(cudaSetDevice(s->device));
    const uint64_t chunk = s->staging_bytes / 8;
    PermArg P = perm_arg(s);
    const uint64_t loc_mask = (1ULL << s->nloc) - 1;
    for (uint64_t off = 0; off < count; off += chunk) {
        uint64_t c = std::min(chunk, count - off);
        if (s->prec == QSB_F32) k_probs<float><<<(unsigned)((c + 255) / 256), 256, 0, s->stream>>>((const float *)s->state, (double *)s->staging, first + off, c, P, loc_mask);
        else k_probs<double><<<(unsigned)((c + 255) / 256), 256, 0, s->stream>>>((const double *)s->state, (double *)s->staging, first + off, c, P, loc_mask);
        QSB_CUDA(cudaGetLastError());
        QSB_CUDA(cudaMemcpyAsync(p + off, s->staging, c * 8, cudaMemcpyDeviceToHost, s->stream));
        QSB_CUDA(cudaStreamSynchronize(s->stream));
    }
    return QSB_OK;
}

/* qsb_cdf, qsb_sample, qsb_save_state, qsb_load_state: readout.cu */
