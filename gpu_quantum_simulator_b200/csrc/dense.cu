/*
 * dense.cu -- QSB_MODE_DENSE: the EXPERIMENT BASELINE.json configuration 4 asks for ("fusion k = 2..5 sweep,
 * tensor-core check for large fused blocks"), not the product path.
 *
 * Host side: consecutive gates are merged greedily into dense k-qubit unitaries (k = qsb_options_t.tile_bits,
 * 2..5) -- the generalisation of the reference's 4x4 pair accumulator (quantum_simulator_4x4.cu:327-501: acc4 per
 * CX-linked pair, flushed on conflict) from k = 2 to k <= 5, in fp64 and without its 1e-3 identity cut.
 * Device side: one sweep per block, k_dense<R, K>, the generalisation of kernel_gate_4
 * (quantum_simulator_4x4.cu:109-146): a thread owns the 2^K amplitudes that differ in the K target bits, the
 * matrix travels as a __grid_constant__ kernel parameter (constant-bank operands of the FMAs).
 *
 * What it is for: with ncu, sm__throughput vs dram__throughput per k decides by measurement (north_star) whether a
 * dense block of that width is compute-bound, i.e. whether tensor cores would be the next step, and ms per circuit
 * per k is compared with the sparse register-tile schedule (profiles/r2/dense_k_sweep.md).
 * Arithmetic per amplitude and block: 2^K complex multiply-adds = 4 * 2^K real FMAs, against 16 bytes of traffic.
 */
#include <algorithm>
#include <map>
#include <string.h>

#include "sim.h"
#include "dense.h"

/* ------------------------------------------------------------------ host: greedy dense fusion */
namespace {
typedef std::pair<double, double> cplx;   /* (re, im) */
inline cplx cmul(cplx a, cplx b) { return {a.first * b.first - a.second * b.second, a.first * b.second + a.second * b.first}; }
inline cplx cadd(cplx a, cplx b) { return {a.first + b.first, a.second + b.second}; }

struct Open { uint64_t qmask = 0; std::vector<COp> ops; size_t born = 0; };

/* dense matrix of a list of ops on the sorted qubits of qmask: column by column, each op applied like a sweep */
void materialise(const Open &o, DenseBlock &b)
{
    b.k = 0;
    for (int q = 0; q < 64; q++) if ((o.qmask >> q) & 1) b.q[b.k++] = q;
    const int D = 1 << b.k;
    std::vector<cplx> M((size_t)D * D, cplx(0.0, 0.0));
    for (int i = 0; i < D; i++) M[(size_t)i * D + i] = cplx(1.0, 0.0);
    auto local = [&](uint64_t mask) { uint32_t r = 0; for (int j = 0; j < b.k; j++) if ((mask >> b.q[j]) & 1) r |= 1u << j; return r; };
    for (const COp &c : o.ops) {
        const uint32_t cm = local(c.ctrl);
        if (c.kind == C_PHASE) {
            const cplx ph(c.m[0], c.m[1]);
            for (int r = 0; r < D; r++) if (((uint32_t)r & cm) == cm) for (int col = 0; col < D; col++) M[(size_t)r * D + col] = cmul(ph, M[(size_t)r * D + col]);
            continue;
        }
        const uint32_t tb = local(1ULL << c.target);
        cplx g[4];
        if (c.kind == C_X) { g[0] = g[3] = cplx(0, 0); g[1] = g[2] = cplx(1, 0); }
        else for (int e = 0; e < 4; e++) g[e] = cplx(c.m[2 * e], c.m[2 * e + 1]);
        for (int r = 0; r < D; r++) {
            if ((uint32_t)r & tb) continue;
            if (((uint32_t)r & cm) != cm) continue;
            const int r1 = r | (int)tb;
            for (int col = 0; col < D; col++) {
                const cplx a = M[(size_t)r * D + col], bb = M[(size_t)r1 * D + col];
                M[(size_t)r * D + col] = cadd(cmul(g[0], a), cmul(g[1], bb));
                M[(size_t)r1 * D + col] = cadd(cmul(g[2], a), cmul(g[3], bb));
            }
        }
    }
    b.m.resize((size_t)D * D * 2);
    for (size_t i = 0; i < (size_t)D * D; i++) { b.m[2 * i] = M[i].first; b.m[2 * i + 1] = M[i].second; }
}
}

int dense_fuse(const std::vector<COp> &cops, const double gphase[2], int n, int k, std::vector<DenseBlock> &out)
{
    if (k < 1 || k > QSB_DENSE_MAX_K) { qsb_set_error("dense fusion width %d out of range 1..%d", k, QSB_DENSE_MAX_K); return QSB_ERR_ARG; }
    out.clear();
    std::vector<Open> open;
    size_t born = 0;
    auto flush = [&](size_t i) { DenseBlock b; materialise(open[i], b); out.push_back(std::move(b)); open.erase(open.begin() + i); };
    cplx scalar(gphase[0], gphase[1]);
    for (const COp &c : cops) {
        if (c.kind == C_MUX) { qsb_set_error("internal: dense fusion runs on canonical ops before CX absorption"); return QSB_ERR_ARG; }
        uint64_t qm = c.ctrl | (c.target >= 0 ? 1ULL << c.target : 0);
        if (qm == 0) { scalar = cmul(scalar, cplx(c.m[0], c.m[1])); continue; }     /* global phase */
        if (__builtin_popcountll(qm) > k) { qsb_set_error("gate on %d qubits is wider than the dense fusion width %d", __builtin_popcountll(qm), k); return QSB_ERR_ARG; }
        uint64_t un = qm;
        for (const Open &o : open) if (o.qmask & qm) un |= o.qmask;
        if (__builtin_popcountll(un) > k) {
            /* does not fit: close every block the gate touches, oldest first (blocks on disjoint qubits commute) */
            for (;;) {
                size_t best = open.size();
                for (size_t i = 0; i < open.size(); i++) if ((open[i].qmask & qm) && (best == open.size() || open[i].born < open[best].born)) best = i;
                if (best == open.size()) break;
                flush(best);
            }
            un = qm;
        }
        /* merge every block the gate touches into one (their op lists concatenate: disjoint qubits commute) */
        Open merged; merged.qmask = un; merged.born = born++;
        for (size_t i = 0; i < open.size();) {
            if (open[i].qmask & qm) { merged.ops.insert(merged.ops.end(), open[i].ops.begin(), open[i].ops.end()); merged.born = std::min(merged.born, open[i].born); open.erase(open.begin() + i); }
            else i++;
        }
        merged.ops.push_back(c);
        open.push_back(std::move(merged));
    }
    while (!open.empty()) {
        size_t best = 0;
        for (size_t i = 1; i < open.size(); i++) if (open[i].born < open[best].born) best = i;
        flush(best);
    }
    if (!(scalar.first == 1.0 && scalar.second == 0.0)) {
        if (out.empty()) { DenseBlock b; b.k = 1; b.q[0] = 0; b.m = {1, 0, 0, 0, 0, 0, 1, 0}; out.push_back(b); }
        for (size_t i = 0; i < out[0].m.size(); i += 2) { const cplx v = cmul(scalar, cplx(out[0].m[i], out[0].m[i + 1])); out[0].m[i] = v.first; out[0].m[i + 1] = v.second; }
    }
    (void)n;
    return QSB_OK;
}

/* ------------------------------------------------------------------ device: one sweep per dense block */
template <typename R, int K> struct DenseArg { R m[(1 << K) * (1 << K) * 2]; int pos[K]; };   /* row-major (re, im); pos ascending */

template <typename R, int K>
__global__ void __launch_bounds__(128)
k_dense(R *st, uint64_t n_groups, const __grid_constant__ DenseArg<R, K> A)
{
    constexpr int D = 1 << K;
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    uint64_t base = g;                                    /* insert a zero at every target position (ascending) */
#pragma unroll
    for (int j = 0; j < K; j++) { const uint64_t lo = base & ((1ULL << A.pos[j]) - 1); base = ((base >> A.pos[j]) << (A.pos[j] + 1)) | lo; }
    R xr[D], xi[D];
#pragma unroll
    for (int c = 0; c < D; c++) {
        uint64_t i = base;
#pragma unroll
        for (int j = 0; j < K; j++) if ((c >> j) & 1) i |= 1ULL << A.pos[j];
        xr[c] = st[Lay<R>::re(i)]; xi[c] = st[Lay<R>::im(i)];
    }
#pragma unroll 1
    for (int r = 0; r < D; r++) {                         /* rows in a loop (uniform coefficient loads), columns unrolled */
        R ar = 0, ai = 0;
#pragma unroll
        for (int c = 0; c < D; c++) {
            const R mr = A.m[2 * (r * D + c)], mi = A.m[2 * (r * D + c) + 1];
            ar = fma(mr, xr[c], ar); ar = fma(-mi, xi[c], ar);
            ai = fma(mr, xi[c], ai); ai = fma(mi, xr[c], ai);
        }
        uint64_t i = base;
#pragma unroll
        for (int j = 0; j < K; j++) if ((r >> j) & 1) i |= 1ULL << A.pos[j];
        st[Lay<R>::re(i)] = ar; st[Lay<R>::im(i)] = ai;   /* in place: every input of the group is already in registers */
    }
}

template <typename R, int K>
static int launch_dense(qsb_sim *s, const DenseBlock &b, const int *pos)
{
    DenseArg<R, K> A;
    for (int j = 0; j < K; j++) A.pos[j] = pos[j];
    /* the block matrix is indexed by the block's qubits in ascending LOGICAL order; pos[] is ascending PHYSICAL
     * order: permute rows / columns accordingly */
    int order[K];                                         /* physical slot j <- block qubit order[j] */
    for (int j = 0; j < K; j++) for (int i = 0; i < K; i++) if (s->perm.pos[b.q[i]] == pos[j]) order[j] = i;
    const int D = 1 << K;
    for (int r = 0; r < D; r++) for (int c = 0; c < D; c++) {
        int rl = 0, cl = 0;
        for (int j = 0; j < K; j++) { if ((r >> j) & 1) rl |= 1 << order[j]; if ((c >> j) & 1) cl |= 1 << order[j]; }
        A.m[2 * (r * D + c)] = (R)b.m[2 * ((size_t)rl * D + cl)];
        A.m[2 * (r * D + c) + 1] = (R)b.m[2 * ((size_t)rl * D + cl) + 1];
    }
    const uint64_t n_groups = (1ULL << s->nloc) >> K;
    k_dense<R, K><<<(unsigned)((n_groups + 127) / 128), 128, 0, s->stream>>>((R *)s->state, n_groups, A);
    QSB_CUDA(cudaGetLastError());
    return QSB_OK;
}

template <typename R>
static int dense_execute_t(qsb_sim *s, const std::vector<DenseBlock> &blocks)
{
    for (const DenseBlock &b : blocks) {
        int pos[QSB_DENSE_MAX_K];
        for (int j = 0; j < b.k; j++) {
            pos[j] = s->perm.pos[b.q[j]];
            if (pos[j] >= s->nloc) { qsb_set_error("DENSE mode is single-GPU only"); return QSB_ERR_ARG; }
        }
        std::sort(pos, pos + b.k);
        int rc;
        switch (b.k) {
        case 1: rc = launch_dense<R, 1>(s, b, pos); break;
        case 2: rc = launch_dense<R, 2>(s, b, pos); break;
        case 3: rc = launch_dense<R, 3>(s, b, pos); break;
        case 4: rc = launch_dense<R, 4>(s, b, pos); break;
        case 5: rc = launch_dense<R, 5>(s, b, pos); break;
        default: qsb_set_error("internal: dense block of width %d", b.k); return QSB_ERR_ARG;
        }
        if (rc) return rc;
    }
    return QSB_OK;
}

int dense_execute(qsb_sim *s, const std::vector<DenseBlock> &blocks)
{
    return s->prec == QSB_F32 ? dense_execute_t<float>(s, blocks) : dense_execute_t<double>(s, blocks);
}
