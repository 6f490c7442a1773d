/*
 * tiled_plan.cpp -- host-side gate fusion and scheduling (no CUDA in this file).
 *
 * Turns the canonical op list into PASSES (one HBM sweep each) and ROUNDS
 * (register-resident butterfly groups inside a pass) and emits the constant
 * tables the kernel of tiled_kernel.cu interprets.  This is the B200-native
 * replacement of the reference's host preprocessing:
 *   preproces.cu:215-269   per-qubit 2x2 accumulation, flush before each CX
 *   4x4.cu:327-501         pair (4x4) accumulation state machine
 *   4x4_permute.cu:350-434 usage-histogram qubit relabel
 * Instead of multiplying matrices together (dense k-qubit blocks cost
 * 8*2^k flop per amplitude and become compute-bound at k >= 4)
 * gates stay sparse: a pass fuses every gate whose target is resident in the
 * tile, costing the 2..4 packed FMAs per amplitude each gate really needs,
 * while the pass count -- the only thing HBM sees -- drops by the fusion
 * factor.  Specialisations:
 *   - diagonal gates and controls never need residency: they are per-thread
 *     predicates / phases on physical index bits (OP_TPHASE costs ~nothing);
 *   - a CX next to a one-qubit gate on its target is absorbed into that gate
 *     as a thread-level multiplexer (U vs X.U), costing zero extra arithmetic
 *     (the 4x4.cu pair accumulator did this with dense 4x4 products);
 *   - an X / CX that is the last gate on its qubit in a round is DEFERRED: the
 *     kernel flips one bit of the thread's vector-index mask and the swap
 *     happens in the next store address (OP_XDEF).  Matrices with a small m00
 *     are pivoted (M = X.M') so that every matrix runs in the 3-FMA unit form;
 *   - phase gates wait for a round in which all their qubits are thread-level
 *     (one entry of the pending-scalar list) unless a later gate of the round
 *     depends on them; thin tail rounds of SM-bound passes move to the next pass.
 * Multi-GPU: Belady-style victim choice and three exchange flavours (fused
 * peer scatter, pipelined copies, NCCL), see DESIGN.md section 5.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <functional>
#include <thread>

#include "sim.h"
#include "tiled.h"

/* ---------------------------------------------------------------- canonical */
/* one (multi-)controlled 2x2 matrix -> canonical ops appended to `out`; global scalars go to gphase.
 * Returns false for a singular diagonal matrix. */
static bool canon_one(const double *m, uint64_t controls, int target, std::vector<COp> &out, double gphase[2])
{
    const bool offdiag_zero = m[2] == 0 && m[3] == 0 && m[4] == 0 && m[5] == 0;
    const bool diag_zero = m[0] == 0 && m[1] == 0 && m[6] == 0 && m[7] == 0;
    COp c; memset(&c, 0, sizeof c);
    if (offdiag_zero) {
        /* diag(d0, d1) = d0 * diag(1, d1/d0) */
        double d0r = m[0], d0i = m[1], d1r = m[6], d1i = m[7];
        if (!(d0r == 1.0 && d0i == 0.0)) {
            double den = d0r * d0r + d0i * d0i;
            if (den == 0) return false;
            if (controls == 0) { /* a global scalar: fold */
                double r = gphase[0] * d0r - gphase[1] * d0i, i = gphase[0] * d0i + gphase[1] * d0r;
                gphase[0] = r; gphase[1] = i;
            } else {
                c.kind = C_PHASE; c.ctrl = controls; c.target = -1; c.m[0] = d0r; c.m[1] = d0i;
                out.push_back(c);
            }
            double qr = (d1r * d0r + d1i * d0i) / den, qi = (d1i * d0r - d1r * d0i) / den;
            d1r = qr; d1i = qi;
        }
        if (!(d1r == 1.0 && d1i == 0.0)) {
            c.kind = C_PHASE; c.ctrl = controls | (1ULL << target); c.target = -1; c.m[0] = d1r; c.m[1] = d1i;
            out.push_back(c);
        }
    } else if (diag_zero && m[2] == 1 && m[3] == 0 && m[4] == 1 && m[5] == 0) {
        c.kind = C_X; c.target = target; c.ctrl = controls;
        out.push_back(c);
    } else {
        c.kind = C_MAT; c.target = target; c.ctrl = controls; memcpy(c.m, m, sizeof c.m);
        /* e^{i phi} x (a real or an rx-form matrix), e.g. SX = e^{i pi/4} RX(pi/2) (quantum_simulator.c:187): pull the
         * scalar out -- a global factor, or a phase gate on the controls -- so that the gate runs in a cheap form
         * instead of the general complex one */
        const bool real_form = m[1] == 0 && m[3] == 0 && m[5] == 0 && m[7] == 0;
        const bool rx_form = m[1] == 0 && m[7] == 0 && m[2] == 0 && m[4] == 0;
        const bool j_form = m[0] == 0 && m[6] == 0 && m[3] == 0 && m[5] == 0;
        if (!real_form && !rx_form && !j_form) {
            const int piv = (m[0] != 0 || m[1] != 0) ? 0 : 2;           /* first non-zero entry of the first row */
            const double mag = hypot(m[piv], m[piv + 1]);
            const double ur = m[piv] / mag, ui = m[piv + 1] / mag;     /* e^{i phi} */
            double r[8], sc = 0;
            for (int k = 0; k < 4; k++) {                              /* r = conj(u) * m */
                r[2 * k] = m[2 * k] * ur + m[2 * k + 1] * ui;
                r[2 * k + 1] = m[2 * k + 1] * ur - m[2 * k] * ui;
                sc = fmax(sc, fmax(fabs(r[2 * k]), fabs(r[2 * k + 1])));
            }
            for (int k = 0; k < 8; k++) if (fabs(r[k]) <= 4e-16 * sc) r[k] = 0.0;
            const bool r_real = r[1] == 0 && r[3] == 0 && r[5] == 0 && r[7] == 0;
            const bool r_rx = r[1] == 0 && r[7] == 0 && r[2] == 0 && r[4] == 0;
            if (r_real || r_rx) {
                memcpy(c.m, r, sizeof c.m);
                if (controls == 0) {
                    const double gr = gphase[0] * ur - gphase[1] * ui, gi = gphase[0] * ui + gphase[1] * ur;
                    gphase[0] = gr; gphase[1] = gi;
                } else {
                    COp ph; memset(&ph, 0, sizeof ph);
                    ph.kind = C_PHASE; ph.ctrl = controls; ph.target = -1; ph.m[0] = ur; ph.m[1] = ui;
                    out.push_back(ph);
                }
            }
        }
        out.push_back(c);
    }
    return true;
}

int qsb_canonicalise(const qsb_gate_t *gates, size_t n, int num_qubits, std::vector<COp> &out, double gphase[2])
{
    gphase[0] = 1.0; gphase[1] = 0.0;
    out.clear(); out.reserve(n + 1);
    for (size_t k = 0; k < n; k++) {
        const qsb_gate_t &g = gates[k];
        if (g.target < 0 || g.target >= num_qubits) { qsb_set_error("gate %zu: target %d outside the %d-qubit register", k, g.target, num_qubits); return QSB_ERR_ARG; }
        if (num_qubits < 64 && (g.controls >> num_qubits)) { qsb_set_error("gate %zu: control outside the register", k); return QSB_ERR_ARG; }
        if (g.controls & (1ULL << g.target)) { qsb_set_error("gate %zu: target %d is also a control", k, g.target); return QSB_ERR_ARG; }
        if (!canon_one(g.m, g.controls, g.target, out, gphase)) { qsb_set_error("gate %zu: singular diagonal matrix", k); return QSB_ERR_ARG; }
    }
    return QSB_OK;
}

/* QSB_PLAN_TRACE=1: the lowered op mix of every round on stderr.  tiled_plan_build plans several candidates and
 * silences all but the schedule it keeps. */
static thread_local bool t_trace_off = false;
void tiled_plan_trace_suppress(bool off) { t_trace_off = off; }
static bool plan_trace() { return !t_trace_off && getenv("QSB_PLAN_TRACE") != nullptr; }

/* ------------------------------------------------------------------ helpers */
int tiled_min_local_bits(int prec, const qsb_options_t *) { return prec == QSB_F64 ? QSB_T_F64 : QSB_T_F32; }

namespace {

/* keeps a pass descriptor below QSB_BLOB_LARGE: in the worst case (every op on the same vector bit)
 * each op occupies a whole group, 160 bytes in f32 and 288 bytes in f64 */
inline int max_pass_ops(bool f32) { return f32 ? 120 : 72; }
#ifndef QSB_FUSED_OP_DIV
#define QSB_FUSED_OP_DIV 1   /* fused-exchange passes: op budget = max_pass_ops / this (round 1: 3) */
#endif
const int MAX_PASS_ROUNDS = 32;
const int QSB_PLAN_OVERFLOW = -100;   /* internal: serialise() could not fit the pass into QSB_BLOB_LARGE */

struct Machine {
    int n, prec, g, nloc, rank, T, a, nb;
    bool f32, lazy_diag, defer_diag, sink_phases, tile_search, hform, diaga;
    int trim_thin, cost_cap;
    bool fused_exchange, force_top, fused_direct;
    int lane_reloc;   /* 0 off; 1..: policy of the end-of-pass lane relocation (build_rounds) */
};

inline int popc(uint64_t x) { return __builtin_popcountll(x); }

/* Dependency bookkeeping for the greedy scans.  level 1 = a skipped op uses the
 * qubit diagonally (later diagonal use may still pass), 2 = non-diagonally. */
struct Blocker {
    uint8_t lvl[64];
    int full = 0;
    void clear() { memset(lvl, 0, sizeof lvl); full = 0; }
    bool ok(const COp &o) const
    {
        if (o.target >= 0 && lvl[o.target]) return false;
        for (uint64_t m = o.ctrl; m; m &= m - 1) if (lvl[__builtin_ctzll(m)] == 2) return false;
        return true;
    }
    void block(const COp &o)
    {
        if (o.target >= 0 && lvl[o.target] != 2) { lvl[o.target] = 2; full++; }
        for (uint64_t m = o.ctrl; m; m &= m - 1) { int q = __builtin_ctzll(m); if (!lvl[q]) lvl[q] = 1; }
    }
};

/* snap tiny components so that structure tests are exact */
inline double snap(double x, double scale) { return fabs(x) <= 4e-16 * scale ? 0.0 : x; }
void snap_mat(const double *in, double *out)
{
    double sc = 0; for (int k = 0; k < 8; k++) sc = std::max(sc, fabs(in[k]));
    for (int k = 0; k < 8; k++) out[k] = snap(in[k], sc);
}

int classify(const double *m) /* 1 real, 2 real-diag/imag-offdiag, 3 general */
{
    if (m[1] == 0 && m[3] == 0 && m[5] == 0 && m[7] == 0) return 1;
    if (m[1] == 0 && m[7] == 0 && m[2] == 0 && m[4] == 0) return 2;
    return 3;
}

const double IDENT[8] = {1, 0, 0, 0, 0, 0, 1, 0};
/* imaginary diagonal, real off-diagonal: i times an rx-form matrix */
bool is_jform(const double *m) { return m[0] == 0 && m[6] == 0 && m[3] == 0 && m[5] == 0; }

/* CX (any number of controls) next to an uncontrolled one-qubit gate on its
 * target  ->  one multiplexed gate (U when the controls are off, X.U / U.X when on).
 * For the rx form, X.U = i * (another rx-form matrix): emit() moves that i into the
 * per-thread phase, so the multiplexed pair keeps the cheap form. */
void absorb_cx(std::vector<COp> &ops, int n)
{
    const int N = (int)ops.size();
    std::vector<char> alive(N, 1);
    auto mergeable = [&](const COp &o, int t) {
        if (o.kind != C_MAT || o.ctrl != 0 || o.target != t) return false;
        return true;
    };
    /* backward: gate U then CX  ->  mux(U, X.U) at U's position */
    {
        std::vector<int> last_touch(n, -1), last_nondiag(n, -1);
        for (int i = 0; i < N; i++) {
            COp &o = ops[i];
            if (o.kind == C_X && o.ctrl) {
                const int t = o.target, j = last_touch[t];
                bool ok = j >= 0 && alive[j] && mergeable(ops[j], t);
                if (ok) for (uint64_t m = o.ctrl; m; m &= m - 1) if (last_nondiag[__builtin_ctzll(m)] > j) ok = false;
                if (ok) {
                    COp &u = ops[j];
                    u.kind = C_MUX; u.ctrl = o.ctrl;
                    /* X.U: rows swapped */
                    for (int k = 0; k < 4; k++) { u.m2[k] = u.m[4 + k]; u.m2[4 + k] = u.m[k]; }
                    alive[i] = 0;
                }
            }
            const uint64_t touch = o.ctrl | (o.target >= 0 ? 1ULL << o.target : 0);
            for (uint64_t m = touch; m; m &= m - 1) last_touch[__builtin_ctzll(m)] = i;
            if (o.target >= 0) last_nondiag[o.target] = i;
        }
    }
    /* forward: CX then gate U  ->  mux(U, U.X) at U's position */
    {
        std::vector<int> next_touch(n, N), next_nondiag(n, N);
        for (int i = N - 1; i >= 0; i--) {
            COp &o = ops[i];
            if (!alive[i]) continue;
            if (o.kind == C_X && o.ctrl) {
                const int t = o.target, j = next_touch[t];
                bool ok = j < N && alive[j] && mergeable(ops[j], t);
                if (ok) for (uint64_t m = o.ctrl; m; m &= m - 1) if (next_nondiag[__builtin_ctzll(m)] < j) ok = false;
                if (ok) {
                    COp &u = ops[j];
                    u.kind = C_MUX; u.ctrl = o.ctrl;
                    /* U.X: columns swapped */
                    u.m2[0] = u.m[2]; u.m2[1] = u.m[3]; u.m2[2] = u.m[0]; u.m2[3] = u.m[1];
                    u.m2[4] = u.m[6]; u.m2[5] = u.m[7]; u.m2[6] = u.m[4]; u.m2[7] = u.m[5];
                    alive[i] = 0;
                    /* the mux now reads the controls at position j: keep later scans conservative */
                    for (uint64_t m = u.ctrl; m; m &= m - 1) { int q = __builtin_ctzll(m); next_touch[q] = std::min(next_touch[q], j); }
                    continue;
                }
            }
            const uint64_t touch = o.ctrl | (o.target >= 0 ? 1ULL << o.target : 0);
            for (uint64_t m = touch; m; m &= m - 1) next_touch[__builtin_ctzll(m)] = i;
            if (o.target >= 0) next_nondiag[o.target] = i;
        }
    }
    std::vector<COp> out; out.reserve(N);
    for (int i = 0; i < N; i++) if (alive[i]) out.push_back(ops[i]);
    ops.swap(out);
}

/* 2x2 products per qubit -- the reference's "preprocessing" (preproces.cu:128-163, 215-269: one product
 * accumulator per qubit, flushed before a CX), done in fp64 and only where it pays:
 *   - consecutive uncontrolled one-qubit gates on a qubit (matrix or X; nothing else touches the qubit in
 *     between) multiply into one matrix when the product keeps a cheap form (real, rx-form, diagonal,
 *     anti-diagonal).  H.H disappears, rx(a).rx(b) is one gate, X.H stays real.  A product that would need the
 *     general complex form (H.rx) is left as two cheap gates.
 *   - phase gates with the same qubit mask multiply when no non-diagonal gate on one of their qubits lies
 *     between them (diagonal gates commute with each other and with controls).
 * The reference multiplies in fp32 and drops products within 1e-3 of the identity (preproces.cu:151-160);
 * here nothing is dropped that is not the identity to rounding. */
void fuse_same_qubit(std::vector<COp> &ops, int n, double gphase[2])
{
    const int N = (int)ops.size();
    std::vector<COp> out; out.reserve(N);
    std::vector<char> dead;                    /* per output op */
    std::vector<int> acc(n, -1);               /* qubit -> output index of the matrix still open for products */
    struct PhaseSlot { uint64_t mask; int idx; };
    std::vector<PhaseSlot> open_phase;         /* phases that may still absorb an equal-mask phase */
    auto close_phases_on = [&](uint64_t qubits) {
        for (size_t k = 0; k < open_phase.size();) {
            if (open_phase[k].mask & qubits) { open_phase[k] = open_phase.back(); open_phase.pop_back(); } else k++;
        }
    };
    auto as_matrix = [&](const COp &o, double *m) {
        if (o.kind == C_X) { const double X[8] = {0, 0, 1, 0, 1, 0, 0, 0}; memcpy(m, X, sizeof X); }
        else memcpy(m, o.m, sizeof(double) * 8);
    };
    for (int i = 0; i < N; i++) {
        const COp &o = ops[i];
        if (o.kind == C_PHASE) {
            /* touches its qubits diagonally: closes their matrix accumulators, may join an open phase */
            for (uint64_t m = o.ctrl; m; m &= m - 1) acc[__builtin_ctzll(m)] = -1;
            bool joined = false;
            if (o.ctrl) for (PhaseSlot &ps : open_phase) if (ps.mask == o.ctrl) {
                COp &a = out[ps.idx];
                const double r = a.m[0] * o.m[0] - a.m[1] * o.m[1], im = a.m[0] * o.m[1] + a.m[1] * o.m[0];
                a.m[0] = r; a.m[1] = im;
                joined = true; break;
            }
            if (!joined) {
                out.push_back(o); dead.push_back(0);
                if (o.ctrl) open_phase.push_back({o.ctrl, (int)out.size() - 1});
            }
            continue;
        }
        const int t = o.target;
        close_phases_on(1ULL << t);            /* a non-diagonal gate on t: phases involving t stop absorbing */
        const bool plain = o.ctrl == 0 && (o.kind == C_MAT || o.kind == C_X);
        if (plain && acc[t] >= 0) {
            COp &a = out[acc[t]];
            double A[8], B[8], P[8];
            as_matrix(a, A); as_matrix(o, B);
            /* P = B . A (A acts first) */
            for (int r = 0; r < 2; r++) for (int c = 0; c < 2; c++) {
                double pr = 0, pi = 0;
                for (int k = 0; k < 2; k++) {
                    const double br = B[2 * (2 * r + k)], bi = B[2 * (2 * r + k) + 1], ar = A[2 * (2 * k + c)], ai = A[2 * (2 * k + c) + 1];
                    pr += br * ar - bi * ai; pi += br * ai + bi * ar;
                }
                P[2 * (2 * r + c)] = pr; P[2 * (2 * r + c) + 1] = pi;
            }
            double S[8]; snap_mat(P, S);
            const bool diag = S[2] == 0 && S[3] == 0 && S[4] == 0 && S[5] == 0;
            const bool anti = S[0] == 0 && S[1] == 0 && S[6] == 0 && S[7] == 0;
            if (diag || anti || classify(S) <= 2 || is_jform(S)) {
                std::vector<COp> repl;
                if (canon_one(S, 0, t, repl, gphase) && repl.size() <= 1 && (repl.empty() || repl[0].kind != C_PHASE)) {
                    if (repl.empty()) { dead[acc[t]] = 1; acc[t] = -1; }          /* identity (global scalar folded) */
                    else a = repl[0];                                              /* one matrix or X, same place */
                    continue;
                }
                if (!repl.empty() && repl[0].kind == C_PHASE && repl.size() == 1) {
                    /* the product is a phase gate on t: it replaces the accumulator and closes it */
                    a = repl[0]; acc[t] = -1;
                    continue;
                }
                /* anything else (cannot happen for 2x2 inputs): fall through and keep the gate */
            }
        }
        /* ordinary op: controls close the accumulators of their qubits, the target opens / closes its own */
        for (uint64_t m = o.ctrl; m; m &= m - 1) acc[__builtin_ctzll(m)] = -1;
        out.push_back(o); dead.push_back(0);
        acc[t] = plain ? (int)out.size() - 1 : -1;
    }
    std::vector<COp> res; res.reserve(out.size());
    for (size_t i = 0; i < out.size(); i++) if (!dead[i]) res.push_back(out[i]);
    ops.swap(res);
}

/* CX between one-qubit gates on its target: two algebraic rewrites that remove matrix ops before anything is scheduled.
 *  (1) If H.X.H^-1 = D is diagonal (H = the Hadamard gate, S.H, ...), then "CX, then H" is "H, then controlled-D", and
 *      "H, then CX" is "controlled-(H^-1.X.H), then H": the non-diagonal two-qubit gate becomes a controlled PHASE (no
 *      residency, one entry of a thread-phase list) and no longer separates H from the one-qubit gate on the other side
 *      of the CX -- the 2x2 products of fuse_same_qubit then see them as neighbours.  h, cx, h -- the reference's
 *      spelling of CZ, SURVEY.md 8c -- is one phase gate and no matrix at all.  Applied when such a gate sits on both
 *      sides (one_sided = false, the default) or on either side (one_sided: reserved[4] = 6, the build timed in GPU call 33).
 *  (2) Gates that commute with X (rx, x) on both sides of the CX: the later one hops over the CX and the 2x2 products make
 *      one gate of the two, which then absorbs the CX (absorb_cx).
 * The moved gate commutes with everything it passes: nothing between the two ops touches the target, and the CX or the
 * controlled phase stays at the CX's place, so it reads the controls when the CX did.
 * Random layered workload (one of {h, rx, rz} per qubit and layer): 1 of 9 CX sits between two h, 1 of 9 between two rx. */
void cx_through_h(std::vector<COp> &ops, double gphase[2], bool one_sided)
{
    const int N = (int)ops.size();
    std::vector<char> dead(N, 0);
    std::vector<std::vector<COp>> repl(N);
    auto touches = [](const COp &o, int t) { return o.target == t || ((o.ctrl >> t) & 1); };
    /* D = H.X.H^-1 (forward) or H^-1.X.H (backward), H unitary; 1 if D is diagonal to rounding, 2 if D = X, else 0 */
    auto conj_x = [](const double *H, bool forward, double *D) -> int {
        auto mul = [](const double *A, const double *B, double *C) {       /* C = A.B, row-major complex 2x2 */
            for (int r = 0; r < 2; r++) for (int c = 0; c < 2; c++) {
                double pr = 0, pi = 0;
                for (int k = 0; k < 2; k++) {
                    const double ar = A[2 * (2 * r + k)], ai = A[2 * (2 * r + k) + 1], br = B[2 * (2 * k + c)], bi = B[2 * (2 * k + c) + 1];
                    pr += ar * br - ai * bi; pi += ar * bi + ai * br;
                }
                C[2 * (2 * r + c)] = pr; C[2 * (2 * r + c) + 1] = pi;
            }
        };
        double Hd[8];                                                       /* H^dagger */
        for (int r = 0; r < 2; r++) for (int c = 0; c < 2; c++) { Hd[2 * (2 * r + c)] = H[2 * (2 * c + r)]; Hd[2 * (2 * r + c) + 1] = -H[2 * (2 * c + r) + 1]; }
        static const double X[8] = {0, 0, 1, 0, 1, 0, 0, 0};
        double T[8], P[8];
        if (forward) { mul(H, X, T); mul(T, Hd, P); } else { mul(Hd, X, T); mul(T, H, P); }
        snap_mat(P, D);
        if (D[2] != 0 || D[3] != 0 || D[4] != 0 || D[5] != 0) {
            /* H.X.H^-1 = X: the gate commutes with the CX (rx, x, ...) */
            if (D[0] == 0 && D[1] == 0 && D[6] == 0 && D[7] == 0 && fabs(D[2] - 1.0) <= 8e-16 && D[3] == 0 && fabs(D[4] - 1.0) <= 8e-16 && D[5] == 0) return 2;
            return 0;
        }
        /* the diagonal of a conjugated X has modulus one: take the rounding out of exact values (+-1, +-i) */
        for (int k : {0, 1, 6, 7}) if (fabs(D[k] - rint(D[k])) <= 8e-16) D[k] = rint(D[k]);
        const double n0 = hypot(D[0], D[1]), n1 = hypot(D[6], D[7]);
        return (fabs(n0 - 1.0) <= 1e-12 && fabs(n1 - 1.0) <= 1e-12) ? 1 : 0;       /* H was unitary */
    };
    auto plain_mat_on = [](const COp &o, int t) { return o.kind == C_MAT && o.ctrl == 0 && o.target == t; };
    for (int i = 0; i < N; i++) {
        const COp &cx = ops[i];
        if (cx.kind != C_X || !cx.ctrl || dead[i] || !repl[i].empty()) continue;
        const int t = cx.target;
        double D[8];
        int j = i + 1;
        while (j < N && (dead[j] || !touches(ops[j], t))) j++;
        if (!one_sided) {
            /* U1, CX, U2 with U1 and U2 both commuting with X (rx . CX . rx): U2 hops over the CX, the 2x2 products make
             * one gate of U2.U1 and that gate absorbs the CX -- one multiplexed gate instead of a multiplexer and a
             * gate (22 of 318 matrix slots on the 30 q layered circuit) */
            int k3 = i - 1;
            while (k3 >= 0 && (dead[k3] || !touches(ops[k3], t))) k3--;
            double D3[8];
            if (j < N && repl[j].empty() && plain_mat_on(ops[j], t) && conj_x(ops[j].m, true, D3) == 2 &&
                k3 >= 0 && repl[k3].empty() && plain_mat_on(ops[k3], t) && conj_x(ops[k3].m, false, D3) == 2) {
                repl[i].push_back(ops[j]); repl[i].push_back(cx); dead[j] = 1;
                continue;
            }
        }
        if (!one_sided) {
            /* Default: only when a Hadamard-like gate sits on BOTH sides of the CX, where the rewrite removes two matrix
             * ops.  With one H the rewrite trades a multiplexed gate (the CX rides for free in the gate's slot) for a
             * plain gate plus a controlled phase, and that phase costs a slot of its own whenever another gate on the
             * target follows in the same round: on the 30 q layered circuit the one-sided rewrite saves 1-2 passes and
             * 13 rounds but adds 100 diagonal slots (479 instead of 378 arithmetic slots), and the arithmetic slots are
             * what the kernel's time follows (call 33: 25 rounds and 2 passes fewer were worth 2 ms of 118). */
            int k2 = i - 1;
            while (k2 >= 0 && (dead[k2] || !touches(ops[k2], t))) k2--;
            double D2[8];
            const bool fw = j < N && repl[j].empty() && plain_mat_on(ops[j], t) && conj_x(ops[j].m, true, D) == 1;
            const bool bw = k2 >= 0 && repl[k2].empty() && plain_mat_on(ops[k2], t) && conj_x(ops[k2].m, false, D2) == 1;
            if (!(fw && bw)) continue;
        }
        if (j < N && repl[j].empty() && plain_mat_on(ops[j], t) && conj_x(ops[j].m, true, D) == 1) {
            std::vector<COp> seq; seq.push_back(ops[j]);
            if (!canon_one(D, cx.ctrl, t, seq, gphase)) continue;
            repl[i].swap(seq); dead[j] = 1;
            continue;
        }
        int k = i - 1;
        while (k >= 0 && (dead[k] || !touches(ops[k], t))) k--;
        if (k >= 0 && repl[k].empty() && plain_mat_on(ops[k], t) && conj_x(ops[k].m, false, D) == 1) {
            std::vector<COp> seq;
            if (!canon_one(D, cx.ctrl, t, seq, gphase)) continue;
            seq.push_back(ops[k]);
            repl[i].swap(seq); dead[k] = 1;
        }
    }
    std::vector<COp> out; out.reserve(N + 8);
    for (int i = 0; i < N; i++) {
        if (dead[i]) continue;
        if (repl[i].empty()) out.push_back(ops[i]);
        else out.insert(out.end(), repl[i].begin(), repl[i].end());
    }
    ops.swap(out);
}

/* SWAP as a relabelling.  Three CX in a row on the same pair with alternating direction (what a front end
 * makes of `swap a, b`; SURVEY.md 8c lists SWAP = 3 CX) exchange the states of the two qubits: instead of
 * moving amplitudes, the two logical qubits trade wires -- every later op is rewritten onto the wire that
 * holds its qubit and the plan's final qubit map absorbs the permutation.  Ops between the three CX that
 * touch neither qubit commute with them, so only the ops touching a or b have to be consecutive.
 * wire[q] on return: the wire (= logical label used by the rewritten ops) that holds logical qubit q. */
void relabel_swaps(std::vector<COp> &ops, int n, int8_t *wire)
{
    const int N = (int)ops.size();
    for (int q = 0; q < 64; q++) wire[q] = (int8_t)q;
    auto is_cx = [&](const COp &o) { return o.kind == C_X && popc(o.ctrl) == 1; };
    auto touches = [&](const COp &o, uint64_t m) { return ((o.ctrl | (o.target >= 0 ? 1ULL << o.target : 0)) & m) != 0; };
    std::vector<char> drop(N, 0), event(N, 0);
    for (int i = 0; i < N; i++) {
        if (drop[i] || !is_cx(ops[i])) continue;
        const int a = __builtin_ctzll(ops[i].ctrl), b = ops[i].target;
        const uint64_t m = (1ULL << a) | (1ULL << b);
        int idx[2], found = 0;
        for (int j = i + 1; j < N && found < 2; j++) {
            if (!touches(ops[j], m)) continue;
            const int wc = found == 0 ? b : a, wt = found == 0 ? a : b;     /* expected control / target */
            if (drop[j] || !is_cx(ops[j]) || __builtin_ctzll(ops[j].ctrl) != wc || ops[j].target != wt) break;
            idx[found++] = j;
        }
        if (found < 2) continue;
        drop[i] = drop[idx[0]] = drop[idx[1]] = 1;
        event[i] = 1;
    }
    std::vector<COp> out; out.reserve(N);
    for (int i = 0; i < N; i++) {
        if (event[i]) std::swap(wire[__builtin_ctzll(ops[i].ctrl)], wire[ops[i].target]);
        if (drop[i]) continue;
        COp o = ops[i];
        if (o.target >= 0) o.target = wire[o.target];
        uint64_t c = 0;
        for (uint64_t mm = o.ctrl; mm; mm &= mm - 1) c |= 1ULL << wire[__builtin_ctzll(mm)];
        o.ctrl = c;
        out.push_back(o);
    }
    (void)n;
    ops.swap(out);
}

} // namespace

/* ------------------------------------------------------------ pass building */
struct PassBuilder {
    const Machine &M;
    const BitPerm &perm;        /* logical -> physical at the start of the pass */
    HostPass hp;
    int tile_of_qubit[64];      /* logical qubit -> tile bit or -1              */
    std::vector<int> tile_qubit; /* tile bit -> logical qubit or -1 (padding)   */
    /* lane relocation (build_rounds): the caller allows it and tells, for the ops this pass consumed, at which op
     * each logical qubit is next used as a target; reloc_map (position -> position) is what the pass then did */
    bool may_relocate = false;
    std::function<void(const std::vector<char> &, size_t *)> next_use_cb;
    bool relocated = false;
    int8_t reloc_map[64];

    PassBuilder(const Machine &m, const BitPerm &p) : M(m), perm(p) {}

    /* choose the tile from the set of resident logical qubits (+ positions that must be resident);
     * pos_map, if given, relocates physical positions inside the tile (in-tile qubit permutation:
     * free, because the tile is gathered and scattered anyway). */
    void set_tile(uint64_t resident, uint64_t forced_pos = 0, const int8_t *pos_map = nullptr, bool fuse_exchange = false,
                  const int *victim_pos = nullptr)
    {
        /* fuse_exchange: the pass also performs the global <-> local qubit exchange: every amplitude is scattered
         * straight into the shard of the rank named by its victim bits (peer memory over NVLink).
         *   round 1 flavour (victim_pos == null): an in-tile permutation first moves the victims to the top g local
         *     positions, which then trade places with the rank bits; the writer's rank lands on the top positions.
         *   direct flavour (victim_pos[k] = local position whose qubit becomes rank bit k): position victim_pos[k]
         *     trades places with rank bit k, wherever it is.  A victim inside the tile is a tile bit whose destination
         *     is a rank bit; a victim OUTSIDE the tile is an outer bit, i.e. constant per CTA, so the whole tile goes to
         *     one peer and the pass needs no tile slot for the exchange at all. */
        auto exch = [&](int p) {
            if (!fuse_exchange) return p;
            if (victim_pos) { for (int k = 0; k < M.g; k++) if (victim_pos[k] == p) return M.nloc + k; return p; }
            return (p >= M.nloc - M.g && p < M.nloc) ? p + M.g : p;
        };
        uint64_t posmask = forced_pos;
        for (int q = 0; q < M.n; q++) if ((resident >> q) & 1) posmask |= 1ULL << perm.pos[q];
        for (int p = 0; p < M.a; p++) posmask |= 1ULL << p;              /* contiguous low segment */
        for (int p = 0; p < M.nloc && popc(posmask) < M.T; p++) posmask |= 1ULL << p; /* pad from the bottom */
        int inv[64]; for (int p = 0; p < 64; p++) inv[p] = -1;
        for (int q = 0; q < M.n; q++) inv[perm.pos[q]] = q;
        hp.T = M.T;
        tile_qubit.assign(M.T, -1);
        for (int q = 0; q < 64; q++) tile_of_qubit[q] = -1;
        int j = 0;
        for (int p = 0; p < M.nloc + M.g; p++) if ((posmask >> p) & 1) {
            hp.tile_src[j] = (int8_t)p;
            hp.tile_dst[j] = (int8_t)exch(pos_map ? pos_map[p] : p);
            tile_qubit[j] = inv[p];
            if (inv[p] >= 0) tile_of_qubit[inv[p]] = j;
            j++;
        }
        memset(&hp.hdr, 0, sizeof hp.hdr);
        int nr = 0, p = 0, outer_bits = 0;
        while (p < M.nloc) {
            if ((posmask >> p) & 1) { p++; continue; }
            int st = p; while (p < M.nloc && !((posmask >> p) & 1)) p++;
            hp.hdr.run_start[nr] = (uint8_t)st; hp.hdr.run_len[nr] = (uint8_t)(p - st); nr++;
            outer_bits += p - st;
        }
        hp.hdr.n_runs = nr;
        hp.hdr.n_tiles = 1ULL << outer_bits;
        hp.hdr.src_fixed = hp.hdr.dst_fixed = (uint64_t)M.rank << M.nloc;
        hp.hdr.nloc = M.nloc;
        hp.hdr.out_of_place = 0;
        hp.fused_swap = fuse_exchange;
        if (fuse_exchange) {
            hp.hdr.out_of_place = 1;
            if (!victim_pos) hp.hdr.dst_fixed = (uint64_t)M.rank << (M.nloc - M.g);   /* the old rank bits land on the top local positions */
            else {
                hp.hdr.dst_fixed = 0;
                for (int k = 0; k < M.g; k++) {
                    hp.hdr.dst_fixed |= (uint64_t)((M.rank >> k) & 1) << victim_pos[k];   /* the writer's rank bit k moves in */
                    if (!((posmask >> victim_pos[k]) & 1)) {                              /* victim outside the tile */
                        hp.hdr.xo_pos[hp.hdr.n_xo] = (uint8_t)victim_pos[k]; hp.hdr.xo_rank[hp.hdr.n_xo] = (uint8_t)k; hp.hdr.n_xo++;
                    }
                }
            }
        }
    }

    /* tile bits that must be thread (lane) bits in the first / last round */
    uint32_t lane_forbidden() const
    {
        uint32_t f = 0;
        const int lo = M.f32 ? 1 : 0;
        for (int j = lo; j < M.a && j < lo + 5; j++) f |= 1u << j; /* tile bit j == physical bit j for j < a */
        return f;
    }

    /* Build rounds for the ordered op list `ops` (all targets resident).  done[i] tells the
     * caller which ops were consumed (a pass is cut at MAX_PASS_ROUNDS). */
    int build_rounds(const std::vector<COp> &ops, std::vector<char> &done, bool more_passes_follow = false)
    {
        const int P = M.f32 ? 0 : -1;             /* pack tile bit */
        const uint32_t F = lane_forbidden();
        const size_t n = ops.size();
        done.assign(n, 0);
        size_t left = n, first_open = 0;
        const bool lazy_diag = M.lazy_diag;
        const int trim_thin = M.trim_thin;
        double sm_cost = 0.0;
        std::vector<uint32_t> roundR;             /* vector-bit set (tile-bit mask) per round */
        std::vector<std::vector<int>> round_ops;
        while (left || roundR.empty()) {
            if ((int)roundR.size() >= MAX_PASS_ROUNDS - 1) break;
            if (M.cost_cap > 0 && !roundR.empty() && more_passes_follow && sm_cost >= M.cost_cap) break;   /* fusion-depth limit (sweeps) */
            const bool is_first = roundR.empty();
            uint32_t R = 0;
            uint32_t nonpack = 0;
            for (int tb = 0; tb < M.T; tb++) if (tb != P) nonpack |= 1u << tb;
            /* every round needs QSB_NVB vector bits that are neither controls of its gates nor (edge rounds) lane bits */
            auto paddable = [&](uint32_t R_, uint32_t creal) { return popc(nonpack & ~F & ~(creal | R_)) + popc(R_) >= QSB_NVB; };
            std::vector<int> mine;
            /* One greedy scan of the open ops.  `allowed`: tile bits that may become vector bits; `cap`: how many of
             * them; commit = false is a dry run (used to choose the vector bits) that reports, per tile bit, how many
             * gates with arithmetic it would run as a target. */
            auto scan = [&](bool commit, uint32_t allowed, int cap, std::vector<int> &acc, std::vector<int> *per_tb) -> uint32_t {
                uint32_t ctrl_used = 0; int nR_ = 0; uint32_t R_ = 0, creal_all = 0;
                uint64_t closed = 0;   /* qubits that took a bare X / CX (or a matrix to be pivoted) in this round: nothing
                                          may follow on them here, so that the X is deferred into the store address */
                Blocker B; B.clear();
                std::vector<char> taken;
                if (!commit) taken.assign(n, 0);
                for (size_t i = first_open; i < n; i++) {
                    if (done[i]) continue;
                    const COp &o = ops[i];
                    bool can = B.ok(o);
                    if (can && ((o.ctrl & closed) || (o.target >= 0 && ((closed >> o.target) & 1)))) can = false;
                    uint32_t cbits = 0; /* tile bits this op uses as (non-diagonal-gate) controls */
                    if (can && (o.kind != C_PHASE || lazy_diag)) {
                        /* controls must stay thread-level; with lazy diagonals so must the qubits of a phase gate: it
                         * then costs one entry of the per-thread phase list instead of arithmetic on the vectors */
                        for (uint64_t m = o.ctrl; m; m &= m - 1) { int tb = tile_of_qubit[__builtin_ctzll(m)]; if (tb >= 0 && tb != P) cbits |= 1u << tb; }
                        if (cbits & R_) can = false;
                    }
                    const uint32_t creal = (o.kind != C_PHASE) ? cbits : 0u;
                    if (can && o.target >= 0) {
                        int tb = tile_of_qubit[o.target];
                        if (tb == P) { if (!paddable(R_, creal_all | creal)) can = false; }
                        else if (is_first && ((F >> tb) & 1)) can = false;
                        else if ((R_ >> tb) & 1) { if (!paddable(R_, creal_all | creal)) can = false; }
                        else if (nR_ < cap && ((allowed >> tb) & 1) && !((ctrl_used >> tb) & 1) &&
                                 (cap > QSB_NVB || paddable(R_ | (1u << tb), creal_all | creal))) { R_ |= 1u << tb; nR_++; }
                        else can = false;
                    }
                    if (can) {
                        acc.push_back((int)i); ctrl_used |= cbits; creal_all |= creal;
                        if (commit) { done[i] = 1; left--; }
                        if (per_tb && o.target >= 0 && tile_of_qubit[o.target] != P && o.kind != C_X) (*per_tb)[tile_of_qubit[o.target]]++;
                        if (o.kind == C_X && tile_of_qubit[o.target] != P) {
                            /* a deferred X is free but ends its qubit's turn in this round; in a chain of CX on one
                             * target (Toffoli decompositions) that would cost a round per CX: run those as matrices */
                            int chain = 0;
                            const int need = 3;     /* 1 or 2 more CX on the target: the layered workloads, unchanged */
                            for (size_t j = i + 1; j < n && j < i + 64 && chain < need; j++)
                                if (!done[j] && ops[j].kind == C_X && ops[j].target == o.target) chain++;
                            if (chain < need) closed |= 1ULL << o.target;
                        }
                        /* a matrix with a small m00 (in either variant of a multiplexer) is cheapest as a pivoted unit
                         * form + deferred X (see emit): that needs it to be the last gate on its qubit in this round */
                        if ((o.kind == C_MAT || o.kind == C_MUX) && tile_of_qubit[o.target] != P &&
                            (hypot(o.m[0], o.m[1]) < 0.3 || (o.kind == C_MUX && hypot(o.m2[0], o.m2[1]) < 0.3)))
                            closed |= 1ULL << o.target;
                    }
                    else { B.block(o); if (B.full >= M.n) break; }
                }
                return R_;
            };
            /* (choosing the vector bits by a dry run -- the QSB_NVB targets with the most runnable gates -- was tried
             * and does not reduce the number of rounds: first come, first served stays) */
            R = scan(true, nonpack, QSB_NVB, mine, nullptr);
            /* A heavy pass is bound by the SM, not by HBM: a thin tail round (fewer than `trim` gates with
             * arithmetic) costs a full shared-memory exchange for almost no work.  Leave its ops to the
             * next pass, whose tile is chosen around them. */
            /* Phase gates that touch a vector bit cost arithmetic on the vectors; as thread-level phases they are
             * one entry of the pending-scalar list.  A phase gate that no later gate of this round depends on (no
             * later target among its qubits) can wait for a round in which all its qubits are thread-level. */
            if (M.defer_diag) {
                std::vector<char> gone(mine.size(), 0);
                bool any = false;
                for (int k = (int)mine.size() - 1; k >= 0; k--) {
                    const COp &o = ops[mine[k]];
                    if (o.kind != C_PHASE) continue;
                    uint32_t bits = 0;
                    for (uint64_t m = o.ctrl; m; m &= m - 1) { int tb = tile_of_qubit[__builtin_ctzll(m)]; if (tb >= 0 && tb != P) bits |= 1u << tb; }
                    if (!(bits & R)) continue;
                    bool needed = false;
                    for (size_t k2 = k + 1; k2 < mine.size() && !needed; k2++) {
                        if (gone[k2]) continue;
                        const COp &o2 = ops[mine[k2]];
                        if (o2.target >= 0 && ((o.ctrl >> o2.target) & 1)) needed = true;
                    }
                    if (!needed) { gone[k] = 1; any = true; done[mine[k]] = 0; left++; }
                }
                if (any) {
                    std::vector<int> kept;
                    for (size_t k = 0; k < mine.size(); k++) if (!gone[k]) kept.push_back(mine[k]);
                    mine.swap(kept);
                }
            }
            {
                int useful = 0;
                for (int i : mine) if (ops[i].kind != C_PHASE) useful++;
                /* cost in units of one unit-form gate: ~5 for the exchange, the round tables and the pending scalar;
                 * one HBM sweep hides roughly 17 such units (measured) */
                if (trim_thin > 0 && sm_cost >= 34.0 && more_passes_follow && !roundR.empty() && useful < trim_thin) {
                    for (int i : mine) { done[i] = 0; left++; }
                    break;
                }
                sm_cost += 5.0 + useful;
            }
            while (first_open < n && done[first_open]) first_open++;
            roundR.push_back(R); round_ops.push_back(mine);
            if (!left) break;
        }
        /* The last round must keep the low DESTINATION bits on lanes (coalesced 128-byte stores).  If it uses a
         * qubit of the low segment as a vector bit, that qubit trades places -- inside the tile, which the pass
         * gathers and scatters anyway -- with a qubit that is a thread bit in the last round: the pass leaves the
         * qubits permuted (the caller folds reloc_map into the layout) instead of paying an empty round, i.e. one
         * more trip of the tile through shared memory, just to turn the registers.  Flast: tile bits whose
         * destination is a lane position. */
        uint32_t Flast = F;
        relocated = false;
        if (may_relocate && M.lane_reloc && roundR.size() >= 2) {
            const uint32_t Rl = roundR.back();
            uint32_t nonpack = 0, ctrl_last = 0;
            for (int tb = 0; tb < M.T; tb++) if (tb != P) nonpack |= 1u << tb;
            for (int i : round_ops.back()) if (ops[i].kind != C_PHASE)
                for (uint64_t m = ops[i].ctrl; m; m &= m - 1) { int tb = tile_of_qubit[__builtin_ctzll(m)]; if (tb >= 0 && tb != P) ctrl_last |= 1u << tb; }
            size_t next_use[64];
            for (int q = 0; q < 64; q++) next_use[q] = ~(size_t)0;
            if (next_use_cb) next_use_cb(done, next_use);
            auto key = [&](int tb) { return tile_qubit[tb] >= 0 ? next_use[tile_qubit[tb]] : ~(size_t)0; };
            /* always: every lane position is re-assigned; otherwise only the qubits in conflict move out */
            auto try_relocate = [&](bool always) {
                std::vector<int> cands;       /* thread bits of the last round that may live on the lanes afterwards */
                for (int tb = 0; tb < M.T; tb++) if (((nonpack & ~Rl & (always ? ~0u : ~F)) >> tb) & 1) cands.push_back(tb);
                /* Who lives on the lanes from now on: the qubits whose next use as a target is NEAREST.  The low positions
                 * are part of every tile, so their qubits are resident in every pass without taking one of the freely
                 * chosen tile slots (30 q layered: 22 -> 18 passes); the price -- they cannot be vector bits of a first
                 * round -- is one round of delay for their first gate.  (On the planner's counts, parking the qubits
                 * with the FURTHEST next use there costs passes instead: 22 passes / 134 rounds.) */
                std::stable_sort(cands.begin(), cands.end(), [&](int x, int y) {
                    if (key(x) != key(y)) return key(x) < key(y);
                    const int fx = (F >> x) & 1, fy = (F >> y) & 1;     /* ties: whoever is there already stays */
                    if (fx != fy) return fx > fy;
                    return x > y;
                });
                const int need = always ? popc(F) : popc(Rl & F);
                if ((int)cands.size() < need) return false;
                uint32_t chosen = 0;
                for (int k = 0; k < need; k++) chosen |= 1u << cands[k];
                const uint32_t incoming = chosen & ~F, outgoing = always ? (F & ~chosen) : (Rl & F);
                if (!incoming || popc(incoming) != popc(outgoing)) return false;
                const uint32_t Fl = (F & ~outgoing) | incoming;
                /* the last round must still find its padding vector bits outside the new lane set */
                if (popc(nonpack & ~Fl & ~(ctrl_last | Rl)) + popc(Rl) < QSB_NVB) return false;
                for (int q = 0; q < 64; q++) reloc_map[q] = (int8_t)q;
                for (uint32_t in = incoming, out = outgoing; in; in &= in - 1, out &= out - 1) {
                    const int tj = __builtin_ctz(out), tc = __builtin_ctz(in);
                    const int8_t pj = hp.tile_src[tj], pc = hp.tile_src[tc];
                    hp.tile_dst[tj] = pc; hp.tile_dst[tc] = pj;
                    reloc_map[pj] = pc; reloc_map[pc] = pj;
                }
                Flast = Fl;
                return true;
            };
            if (M.lane_reloc >= 4) relocated = try_relocate(true);
            if (!relocated && (Rl & F)) relocated = try_relocate(false);
        }
        if (roundR.back() & Flast) { roundR.push_back(0); round_ops.push_back({}); Flast = F; }
        if (plan_trace()) {
            for (size_t r = 0; r < roundR.size(); r++) {
                int useful = 0, ph = 0; for (int i : round_ops[r]) (ops[i].kind != C_PHASE ? useful : ph)++;
                fprintf(stderr, "qsb-plan: selected round %zu: vector bits %03x, gates %d, phases %d\n", r, roundR[r], useful, ph);
            }
            fprintf(stderr, "qsb-plan: pass of %zu rounds, %zu of %zu ops left to the next pass, lane relocation %d\n", roundR.size(), left, n, (int)relocated);
        }

        const int nrounds = (int)roundR.size();
        /* Sink thread-level phase gates to the latest round that can take them: a phase whose qubits are
         * thread-level in this round AND the next one commutes with everything in between, so it can ride
         * with the next round's pending scalar.  Rounds left without any phase skip the scalar multiply. */
        if (M.sink_phases) for (int r = 0; r + 1 < nrounds; r++) {
            std::vector<int> keep, moved;
            /* tile bits the next round cannot use as padding vector bits: its targets, controls and phase qubits */
            uint32_t busy = roundR[r + 1], nonpack = 0;
            for (int tb = 0; tb < M.T; tb++) if (tb != P) nonpack |= 1u << tb;
            for (int i : round_ops[r + 1])
                for (uint64_t m = ops[i].ctrl; m; m &= m - 1) { int tb = tile_of_qubit[__builtin_ctzll(m)]; if (tb >= 0 && tb != P) busy |= 1u << tb; }
            const uint32_t edgeF = (r + 1 == nrounds - 1) ? Flast : 0u;
            for (int i : round_ops[r]) {
                const COp &o = ops[i];
                bool move = false;
                if (o.kind == C_PHASE) {
                    uint32_t bits = 0; bool pack = false;
                    for (uint64_t m = o.ctrl; m; m &= m - 1) {
                        int tb = tile_of_qubit[__builtin_ctzll(m)];
                        if (tb == P && tb >= 0) pack = true; else if (tb >= 0) bits |= 1u << tb;
                    }
                    move = !pack && !(bits & roundR[r]) && !(bits & roundR[r + 1]);
                    /* the next round must still find its padding vector bits among qubits without phases */
                    if (move && popc(nonpack & ~edgeF & ~(busy | bits)) + popc(roundR[r + 1]) < QSB_NVB) move = false;
                    if (move) busy |= bits;
                }
                (move ? moved : keep).push_back(i);
            }
            if (!moved.empty()) {
                round_ops[r].swap(keep);
                round_ops[r + 1].insert(round_ops[r + 1].begin(), moved.begin(), moved.end());
            }
        }
        hp.rounds.assign(nrounds, DevRound());
        hp.round_thr.assign(nrounds, {}); hp.round_vec.assign(nrounds, {});
        std::vector<uint32_t> ctrl_of_round(nrounds, 0), phase_of_round(nrounds, 0);
        for (int r = 0; r < nrounds; r++)
            for (int i : round_ops[r])
                for (uint64_t m = ops[i].ctrl; m; m &= m - 1) {
                    int tb = tile_of_qubit[__builtin_ctzll(m)];
                    if (tb >= 0 && tb != P) (ops[i].kind != C_PHASE ? ctrl_of_round[r] : phase_of_round[r]) |= 1u << tb;
                }
        for (int r = 0; r < nrounds; r++) {
            uint32_t R = roundR[r];
            /* lane bits of the edge rounds: the low source positions in the first round, the low destinations in the last */
            const uint32_t edgeF = (r == 0 ? F : 0u) | (r == nrounds - 1 ? Flast : 0u);
            /* pad R with the highest free tile bits: never a control of this round; the qubits of its phase
             * gates only if nothing else is left (they then cost vector arithmetic instead of a thread phase) */
            for (int tb = M.T - 1; tb >= 0 && popc(R) < QSB_NVB; tb--) {
                if (tb == P || ((R >> tb) & 1) || (((ctrl_of_round[r] | phase_of_round[r]) >> tb) & 1)) continue;
                if ((edgeF >> tb) & 1) continue;
                R |= 1u << tb;
            }
            while (popc(R) < QSB_NVB) {   /* second choice: the phase-gate qubit with the fewest phase gates on it */
                int best = -1, best_cnt = 1 << 30;
                for (int tb = M.T - 1; tb >= 0; tb--) {
                    if (tb == P || ((R >> tb) & 1) || ((ctrl_of_round[r] >> tb) & 1)) continue;
                    if ((edgeF >> tb) & 1) continue;
                    int cnt = 0;
                    for (int i : round_ops[r]) if (ops[i].kind == C_PHASE && tile_qubit[tb] >= 0 && ((ops[i].ctrl >> tile_qubit[tb]) & 1)) cnt++;
                    if (cnt < best_cnt) { best_cnt = cnt; best = tb; }
                }
                if (best < 0) break;
                R |= 1u << best;
            }
            if (popc(R) < QSB_NVB) { qsb_set_error("internal: round %d cannot be padded to %d vector bits", r, QSB_NVB); return QSB_ERR_ARG; }
            std::vector<int8_t> vec, thr;
            for (int tb = 0; tb < M.T; tb++) if ((R >> tb) & 1) vec.push_back((int8_t)tb);
            /* thread bits: ascending; on edge rounds this puts the low physical bits on the lanes */
            for (int tb = 0; tb < M.T; tb++) if (tb != P && !((R >> tb) & 1)) thr.push_back((int8_t)tb);
            /* after a lane relocation the last round orders them by destination */
            if (relocated && r == nrounds - 1)
                std::stable_sort(thr.begin(), thr.end(), [&](int8_t x, int8_t y) { return hp.tile_dst[x] < hp.tile_dst[y]; });
            hp.round_vec[r] = vec; hp.round_thr[r] = thr;
            DevRound &D = hp.rounds[r];
            memset(&D, 0, sizeof D);
            for (int j = 0; j < QSB_TB; j++) D.thr[j].gidx = 1ULL << hp.tile_src[thr[j]];
            for (int j = 0; j < QSB_NVB; j++) D.vec[j].gidx = 1ULL << hp.tile_src[vec[j]];
        }
        for (int j = 0; j < QSB_TB; j++) hp.hdr.dst_thr[j] = 1ULL << hp.tile_dst[hp.round_thr[nrounds - 1][j]];
        for (int j = 0; j < QSB_NVB; j++) hp.hdr.dst_vec[j] = 1ULL << hp.tile_dst[hp.round_vec[nrounds - 1][j]];
        hp.hdr.n_rounds = nrounds;

        for (int r = 0; r + 1 < nrounds; r++) slot_map(r);

        hp.ops.clear();
        hp.round_op_begin.assign(nrounds, 0); hp.round_op_count.assign(nrounds, 0);
        /* The scale of an un-multiplexed unit-form op is the same for every thread: collect those
         * into one factor per pass instead of forcing a pending-scalar multiply in each round. */
        double pass_scale = 1.0;
        std::vector<std::vector<HostOp>> rops(nrounds);
        int flagged = -1;
        for (int r = 0; r < nrounds; r++) {
            for (size_t k = 0; k < round_ops[r].size(); k++) {
                const COp &o = ops[round_ops[r][k]];
                /* is this the last op of the round that involves its target qubit?  (an X can then be deferred) */
                bool last_on_target = o.target >= 0;
                for (size_t k2 = k + 1; k2 < round_ops[r].size() && last_on_target; k2++) {
                    const COp &o2 = ops[round_ops[r][k2]];
                    if (o2.target == o.target || ((o2.ctrl >> o.target) & 1)) last_on_target = false;
                }
                emit(o, r, last_on_target);
            }
            rops[r].swap(hp.ops);
            bool need = false;
            for (HostOp &h : rops[r]) {
                const uint32_t cd_ = h.kind & 0xff; const bool mux = (h.kind >> 16) & 1;
                if (cd_ == OP_TPHASE) need = true;
                else if (cd_ == OP_MAT_U || cd_ == OP_MAT_UI) {
                    const int ai = cd_ == OP_MAT_U ? 3 : 5;
                    if (mux && h.c[0][ai][0] == h.c[1][ai][0] && h.c[0][ai][1] == h.c[1][ai][1] && h.c[0][ai][0] == h.c[0][ai][1]) {
                        /* both variants of the multiplexer carry the same scale (e.g. H | H.X): thread-independent */
                        pass_scale *= h.c[0][ai][0];
                        h.c[0][ai][0] = h.c[0][ai][1] = h.c[1][ai][0] = h.c[1][ai][1] = 1.0;
                    }
                    else if (mux || h.tmask) need = true;   /* per-thread scale */
                    else { pass_scale *= h.c[0][ai][0]; h.c[0][ai][0] = h.c[0][ai][1] = 1.0; }
                }
            }
            if (need) { hp.rounds[r].flags |= 1u; if (flagged < 0) flagged = r; }
        }
        if (pass_scale != 1.0) {
            if (flagged < 0) { flagged = nrounds - 1; hp.rounds[flagged].flags |= 1u; }
            HostOp t; memset(&t, 0, sizeof t);
            t.kind = OPK(OP_TPHASE, 0, 0, 0); t.tmask = 0; t.tph[0] = pass_scale; t.tph[1] = 0.0;
            rops[flagged].push_back(t);
        }
        for (int r = 0; r < nrounds; r++) {
            hp.round_op_begin[r] = (uint32_t)hp.ops.size();
            for (HostOp &h : rops[r]) hp.ops.push_back(h);
            hp.round_op_count[r] = (uint32_t)rops[r].size();
        }
        hp.n_source_ops = (int)(n - left);
        return QSB_OK;
    }

    /* GF(2)-linear slot map for the exchange between round r (writer) and r+1 (reader).
     * Slot space = non-pack tile bits.  The low nb slot bits select the bank group; the
     * nb lowest lane bits of BOTH sides must map to independent bank vectors. */
    void slot_map(int r)
    {
        const int P = M.f32 ? 0 : -1;
        const int nb = M.nb;
        uint16_t col[16]; memset(col, 0, sizeof col);
        bool has[16]; memset(has, 0, sizeof has);
        const std::vector<int8_t> &wt = hp.round_thr[r], &rt = hp.round_thr[r + 1];
        bool used[8] = {false};
        for (int i = 0; i < nb; i++) { col[rt[i]] = (uint16_t)(1u << i); has[rt[i]] = true; }
        for (int i = 0; i < nb; i++) if (has[wt[i]]) used[__builtin_ctz(col[wt[i]])] = true;
        for (int i = 0; i < nb; i++) if (!has[wt[i]]) {
            int b = 0; while (used[b]) b++;
            used[b] = true; col[wt[i]] = (uint16_t)(1u << b); has[wt[i]] = true;
        }
        bool isD[16] = {false};
        for (int i = 0; i < nb; i++) isD[rt[i]] = true;
        int up = nb;
        for (int tb = 0; tb < M.T; tb++) {
            if (tb == P || isD[tb]) continue;
            col[tb] |= (uint16_t)(1u << up); up++;
        }
        DevRound &W = hp.rounds[r], &Rd = hp.rounds[r + 1];
        for (int j = 0; j < QSB_TB; j++) { W.thr[j].st = col[wt[j]]; Rd.thr[j].ld = col[rt[j]]; }
        for (int j = 0; j < QSB_NVB; j++) { W.vec[j].st = col[hp.round_vec[r][j]]; Rd.vec[j].ld = col[hp.round_vec[r + 1][j]]; }
    }

    static void set_c(HostOp &h, int set, int k, double lo, double hi) { h.c[set][k][0] = lo; h.c[set][k][1] = hi; }

    /* fill coefficient set `set` of a vector-bit matrix op: lo lane uses mlo, hi lane mhi */
    static void fill_mat(HostOp &h, int set, int form, const double *mlo, const double *mhi)
    {
        if (form == 1) {        /* m01 m10 | m00 m11 */
            set_c(h, set, 0, mlo[2], mhi[2]); set_c(h, set, 1, mlo[4], mhi[4]);
            set_c(h, set, 2, mlo[0], mhi[0]); set_c(h, set, 3, mlo[6], mhi[6]);
        } else if (form == 2) { /* [[a, ib],[ic, d]] -> -b b | -c c | a d */
            set_c(h, set, 0, -mlo[3], -mhi[3]); set_c(h, set, 1, mlo[3], mhi[3]);
            set_c(h, set, 2, -mlo[5], -mhi[5]); set_c(h, set, 3, mlo[5], mhi[5]);
            set_c(h, set, 4, mlo[0], mhi[0]); set_c(h, set, 5, mlo[6], mhi[6]);
        } else {                /* m00i m01r | m01i m10r | m10i m11i | m00r m11r */
            set_c(h, set, 0, mlo[1], mhi[1]); set_c(h, set, 1, mlo[2], mhi[2]);
            set_c(h, set, 2, mlo[3], mhi[3]); set_c(h, set, 3, mlo[4], mhi[4]);
            set_c(h, set, 4, mlo[5], mhi[5]); set_c(h, set, 5, mlo[7], mhi[7]);
            set_c(h, set, 6, mlo[0], mhi[0]); set_c(h, set, 7, mlo[6], mhi[6]);
        }
    }
    /* pack-bit matrix op: A = (m00, m11), B = (m01, m10) */
    static void fill_matp(HostOp &h, int set, int form, const double *m)
    {
        if (form == 1) { set_c(h, set, 0, m[0], m[6]); set_c(h, set, 1, m[2], m[4]); }
        else { set_c(h, set, 0, m[0], m[6]); set_c(h, set, 1, m[1], m[7]); set_c(h, set, 2, m[2], m[4]); set_c(h, set, 3, m[3], m[5]); }
    }

    void emit(const COp &o, int r, bool last_on_target)
    {
        const int P = M.f32 ? 0 : -1;
        HostOp h; memset(&h, 0, sizeof h);
        bool pack_ctrl = false;
        uint32_t vmask = 0;
        for (uint64_t m = o.ctrl; m; m &= m - 1) {
            int q = __builtin_ctzll(m);
            int tb = tile_of_qubit[q];
            int vi = -1;
            if (tb >= 0) for (int j = 0; j < QSB_NVB; j++) if (hp.round_vec[r][j] == tb) vi = j;
            if (tb >= 0 && tb == P) pack_ctrl = true;
            else if (vi >= 0) vmask |= 1u << vi;
            else h.tmask |= 1ULL << perm.pos[q];
        }
        if (o.kind == C_PHASE) {
            if (!pack_ctrl && vmask == 0) {
                h.kind = OPK(OP_TPHASE, 0, 0, 0); h.n_coef = 0; h.tph[0] = o.m[0]; h.tph[1] = o.m[1];
            } else {
                if (vmask == 0) h.kind = OPK(OP_DIAG_ALL, 0, 0, 0);
                else if (popc(vmask) == 1) h.kind = OPK(OP_DIAG_V, __builtin_ctz(vmask), 0, 0);
                else h.kind = OPK(OP_DIAG_GEN, 0, 0, vmask);
                h.n_coef = 2;
                set_c(h, 0, 0, pack_ctrl ? 1.0 : o.m[0], o.m[0]);
                set_c(h, 0, 1, pack_ctrl ? 0.0 : o.m[1], o.m[1]);
            }
            hp.ops.push_back(h);
            return;
        }
        const int tb = tile_of_qubit[o.target];
        int vb = -1;
        for (int j = 0; j < QSB_NVB; j++) if (hp.round_vec[r][j] == tb) vb = j;
        /* X / CX whose target is a vector bit and that nothing else in this round touches afterwards:
         * the swap of the two register halves is deferred into the next store address (free). */
        if (o.kind == C_X && tb != P && !pack_ctrl && vmask == 0 && last_on_target) {
            h.kind = OPK(OP_XDEF, vb, 0, 0); h.n_coef = 0;
            hp.ops.push_back(h);
            return;
        }
        double m0[8], m1[8]; /* m0: controls not satisfied, m1: satisfied */
        /* X / CX that could not be absorbed: the swap is issued as the real matrix [[0,1],[1,0]] --
         * 0*x + y is exact, costs fewer instructions than register moves and keeps every update in place */
        static const double XMAT[8] = {0, 0, 1, 0, 1, 0, 0, 0};
        if (o.kind == C_MUX) { snap_mat(o.m, m0); snap_mat(o.m2, m1); }
        else if (o.kind == C_X) { memcpy(m0, IDENT, sizeof m0); memcpy(m1, XMAT, sizeof m1); }
        else { memcpy(m0, IDENT, sizeof m0); snap_mat(o.m, m1); }
        const bool is_mux = (o.kind == C_MUX);
        /* ---- unit forms completed by a deferred X -------------------------------------------------
         * Nothing else in this round touches the target afterwards, so an X on it costs nothing (OP_XDEF).
         * That removes the need for full-form arithmetic: a matrix with a small m00 is X . M' with M'
         * in unit form, and the multiplexer (U | X.U) of an absorbed CX is U for every thread followed
         * by the deferred CX, with one uniform coefficient set and no per-thread scale. */
        if (tb != P && !pack_ctrl && last_on_target && (!is_mux || h.tmask)) {
            auto u_ok = [](const double *m) { return fabs(m[0]) >= 0.3; };
            /* M = (times_i ? i : 1) * out with out real or rx form and a usable m00; returns the class or 0 */
            auto cand = [&](const double *Mx, double *out, bool &times_i) -> int {
                const int c = classify(Mx);
                times_i = false;
                if (c == 1 || c == 2) { memcpy(out, Mx, 8 * sizeof(double)); return u_ok(out) ? c : 0; }
                if (is_jform(Mx)) {
                    for (int k = 0; k < 4; k++) { out[2 * k] = Mx[2 * k + 1]; out[2 * k + 1] = -Mx[2 * k]; }
                    times_i = true;
                    const int c2 = classify(out);
                    return ((c2 == 1 || c2 == 2) && u_ok(out)) ? c2 : 0;
                }
                return 0;
            };
            auto rowswap = [](const double *Mx, double *out) { for (int k = 0; k < 4; k++) { out[k] = Mx[4 + k]; out[4 + k] = Mx[k]; } };
            auto fill_u = [&](HostOp &hh, int set, int cls, const double *m) {
                const double a = m[0];
                if (cls == 1) {
                    const double pp = m[2] / a, q = m[4] / a, rr = m[6] / a;
                    set_c(hh, set, 0, pp, pp); set_c(hh, set, 1, q, q); set_c(hh, set, 2, rr - q * pp, rr - q * pp); set_c(hh, set, 3, a, a);
                } else {
                    const double pp = m[3] / a, q = m[5] / a, rr = m[6] / a;
                    set_c(hh, set, 0, pp, pp); set_c(hh, set, 1, -pp, -pp); set_c(hh, set, 2, q, q); set_c(hh, set, 3, -q, -q);
                    set_c(hh, set, 4, rr + q * pp, rr + q * pp); set_c(hh, set, 5, a, a);
                }
            };
            auto push_tph = [&](uint64_t tmask, double pr, double pi) {
                HostOp t; memset(&t, 0, sizeof t);
                t.kind = OPK(OP_TPHASE, 0, 0, 0); t.tmask = tmask; t.tph[0] = pr; t.tph[1] = pi;
                hp.ops.push_back(t);
            };
            auto push_xdef = [&](uint64_t tmask) {
                HostOp x; memset(&x, 0, sizeof x);
                x.kind = OPK(OP_XDEF, vb, 0, 0); x.tmask = tmask; x.n_coef = 0;
                hp.ops.push_back(x);
            };
            double d0[8], d1[8], s0[8], s1[8], u0[8];
            bool t0 = false, t1 = false;
            if (is_mux) {
                rowswap(m1, s1);
                bool same = true;
                for (int k = 0; k < 8; k++) if (s1[k] != m0[k]) same = false;
                const int c0 = cand(m0, u0, t0);
                if (same && c0) {
                    /* U for everyone, then the CX: uniform coefficients, scale folded into the pass scale */
                    HostOp g; memset(&g, 0, sizeof g);
                    g.kind = OPK(c0 == 1 ? OP_MAT_U : OP_MAT_UI, vb, 0, 0); g.tmask = 0; g.n_coef = c0 == 1 ? 4 : 6;
                    fill_u(g, 0, c0, u0);
                    hp.ops.push_back(g);
                    if (t0) push_tph(0, 0.0, 1.0);
                    push_xdef(h.tmask);
                    return;
                }
                /* both variants in unit form, each possibly through a pivot */
                bool p0 = false, p1 = false, ok = true;
                int k0 = cand(m0, d0, t0), k1 = cand(m1, d1, t1);
                if (!(k0 && k1 && k0 == k1)) {
                    if (!k0) { rowswap(m0, s0); k0 = cand(s0, d0, t0); p0 = true; }
                    if (!k1) { k1 = cand(s1, d1, t1); p1 = true; }
                    if (!(k0 && k1 && k0 == k1)) ok = false;
                    if (ok) {
                        HostOp g; memset(&g, 0, sizeof g);
                        g.kind = OPK(k0 == 1 ? OP_MAT_U : OP_MAT_UI, vb, 1, 0); g.tmask = h.tmask; g.n_coef = k0 == 1 ? 4 : 6;
                        fill_u(g, 0, k0, d0); fill_u(g, 1, k1, d1);
                        hp.ops.push_back(g);
                        /* factors of i: set 0 only -> i everywhere and -i where the predicate holds */
                        if (t0 && t1) push_tph(0, 0.0, 1.0);
                        else if (t1) push_tph(h.tmask, 0.0, 1.0);
                        else if (t0) { push_tph(0, 0.0, 1.0); push_tph(h.tmask, 0.0, -1.0); }
                        if (p0) push_xdef(0);                 /* X for everyone ... */
                        if (p0 != p1) push_xdef(h.tmask);     /* ... toggled back (or on) where the predicate holds */
                        return;
                    }
                }
            } else if (o.kind != C_X) {
                const int k1 = cand(m1, d1, t1);
                if (!k1) {
                    rowswap(m1, s1);
                    const int kp = cand(s1, d1, t1);
                    if (kp) {
                        HostOp g; memset(&g, 0, sizeof g);
                        g.kind = OPK(kp == 1 ? OP_MAT_U : OP_MAT_UI, vb, 0, 0); g.tmask = h.tmask; g.n_coef = kp == 1 ? 4 : 6;
                        fill_u(g, 0, kp, d1);
                        hp.ops.push_back(g);
                        if (t1) push_tph(h.tmask, 0.0, 1.0);
                        push_xdef(h.tmask);
                        return;
                    }
                }
            }
        }
        /* rx-form multiplexer: m1 = i * m1' with m1' in rx form; the i becomes a thread phase */
        bool extra_i = false;
        if (is_mux && classify(m0) == 2 && classify(m1) == 3 && is_jform(m1) && !pack_ctrl && h.tmask && tile_of_qubit[o.target] != P) {
            for (int k = 0; k < 4; k++) { double re_ = m1[2 * k], im_ = m1[2 * k + 1]; m1[2 * k] = im_; m1[2 * k + 1] = -re_; }
            extra_i = true;
        }
        int form = classify(m1);
        if (is_mux || pack_ctrl) form = std::max(form, classify(m0));
        if (form == 2 && is_mux && classify(m0) != classify(m1)) form = 3;
        if (form == 2 && pack_ctrl && !is_mux) form = 2; /* identity fits the rx form */
        if (tb == P) {
            if (form == 2) form = 3;
            h.kind = OPK(form == 1 ? OP_MATP_R : OP_MATP_G, 0, (is_mux && h.tmask) ? 1 : 0, 0);
            h.n_coef = form == 1 ? 2 : 4;
            fill_matp(h, (is_mux && h.tmask) ? 1 : 0, form, m1);
            if (is_mux && h.tmask) fill_matp(h, 0, form, m0);
            else if (is_mux) fill_matp(h, 0, form, m1); /* mux without any control left cannot occur */
        } else {
            /* unit form (see tiled.h): needs a well-conditioned m00 in every variant and no lane dependence */
            const bool two = is_mux && h.tmask;
            auto unit_ok = [&](const double *m) { return fabs(m[0]) >= 0.3; };
            if (!pack_ctrl && (form == 1 || form == 2) && unit_ok(m1) && (!two || unit_ok(m0)) && (is_mux ? (two || true) : true) && !(is_mux && !h.tmask)) {
                auto fill_unit = [&](int set, const double *m) {
                    const double a = m[0];
                    if (form == 1) {
                        const double pp = m[2] / a, q = m[4] / a, r = m[6] / a;
                        set_c(h, set, 0, pp, pp); set_c(h, set, 1, q, q); set_c(h, set, 2, r - q * pp, r - q * pp); set_c(h, set, 3, a, a);
                    } else {
                        const double pp = m[3] / a, q = m[5] / a, r = m[6] / a;
                        set_c(h, set, 0, pp, pp); set_c(h, set, 1, -pp, -pp); set_c(h, set, 2, q, q); set_c(h, set, 3, -q, -q);
                        set_c(h, set, 4, r + q * pp, r + q * pp); set_c(h, set, 5, a, a);
                    }
                };
                h.kind = OPK(form == 1 ? OP_MAT_U : OP_MAT_UI, vb, two ? 1 : 0, 0);
                h.n_coef = form == 1 ? 4 : 6;
                if (two) { fill_unit(0, m0); fill_unit(1, m1); } else fill_unit(0, m1);
                hp.ops.push_back(h);
                if (extra_i) {
                    HostOp t; memset(&t, 0, sizeof t);
                    t.kind = OPK(OP_TPHASE, 0, 0, 0); t.tmask = h.tmask; t.tph[0] = 0.0; t.tph[1] = 1.0;
                    hp.ops.push_back(t);
                }
                return;
            }
            const int opc = form == 1 ? OP_MAT_R : form == 2 ? OP_MAT_I : OP_MAT_G;
            h.n_coef = form == 1 ? 4 : form == 2 ? 6 : 8;
            if (is_mux && h.tmask) {
                /* thread-level select: failing threads use m0; passing threads use m1 (hi lane only if the pack bit is a control too) */
                h.kind = OPK(opc, vb, 1, 0);
                fill_mat(h, 0, form, m0, m0);
                fill_mat(h, 1, form, pack_ctrl ? m0 : m1, m1);
            } else {
                /* no thread-level control (mux on the pack bit alone) or plain controlled gate: one set */
                h.kind = OPK(opc, vb, 0, 0);
                fill_mat(h, 0, form, pack_ctrl ? m0 : m1, m1);
            }
        }
        hp.ops.push_back(h);
        if (extra_i) {
            HostOp t; memset(&t, 0, sizeof t);
            t.kind = OPK(OP_TPHASE, 0, 0, 0); t.tmask = h.tmask; t.tph[0] = 0.0; t.tph[1] = 1.0;
            hp.ops.push_back(t);
        }
    }

    /* Lower the logical tables (hdr / rounds / ops) to the device encoding of tiled.h (GPass, GRound,
     * op stream, GTPhase lists) and serialise it into the kernel-parameter blob. */
    int serialise()
    {
        const bool f32 = M.f32;
        const uint64_t AMP = f32 ? 8 : 16;             /* bytes per amplitude */
        const uint64_t loc_mask = (1ULL << M.nloc) - 1;
        auto al16 = [](size_t x) { return (x + 15) / 16 * 16; };
        const int nrounds = (int)hp.rounds.size();
        auto put_s = [&](std::vector<uint8_t> &o, double v) {
            if (f32) { float x = (float)v; const uint8_t *q = (const uint8_t *)&x; o.insert(o.end(), q, q + 4); }
            else { const uint8_t *q = (const uint8_t *)&v; o.insert(o.end(), q, q + 8); }
        };
        auto put_v = [&](std::vector<uint8_t> &o, const double lane[2]) {   /* (lo, hi) floats or one double */
            if (f32) { float x[2] = {(float)lane[0], (float)lane[1]}; const uint8_t *q = (const uint8_t *)x; o.insert(o.end(), q, q + 8); }
            else { const uint8_t *q = (const uint8_t *)&lane[1]; o.insert(o.end(), q, q + 8); }
        };
        auto pad16 = [&](std::vector<uint8_t> &o) { o.resize(al16(o.size()), 0); };

        GPass gp; memset(&gp, 0, sizeof gp);
        gp.n_rounds = hp.hdr.n_rounds; gp.n_runs = hp.hdr.n_runs;
        memcpy(gp.run_start, hp.hdr.run_start, sizeof gp.run_start);
        memcpy(gp.run_len, hp.hdr.run_len, sizeof gp.run_len);
        gp.src_fixed = hp.hdr.src_fixed; gp.n_tiles = hp.hdr.n_tiles; gp.nloc = hp.hdr.nloc;
        /* destination index bits -> local byte offset | rank contribution << QSB_RANK_SHIFT (fused exchange only) */
        auto enc_dst = [&](uint64_t bits) { return ((bits & loc_mask) * AMP) | ((bits >> M.nloc) << QSB_RANK_SHIFT); };
        for (int j = 0; j < QSB_TB; j++) {
            gp.ld_thr[j] = (hp.rounds[0].thr[j].gidx & loc_mask) * AMP;
            gp.st_thr[j] = hp.fused_swap ? enc_dst(hp.hdr.dst_thr[j]) : (hp.hdr.dst_thr[j] & loc_mask) * AMP;
        }
        for (int v = 0; v < QSB_NV; v++) {
            uint64_t l = 0, t = 0;
            for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) { l |= hp.rounds[0].vec[b].gidx; t |= hp.hdr.dst_vec[b]; }
            gp.ld_vec[v] = (l & loc_mask) * AMP;
            gp.st_vec[v] = hp.fused_swap ? enc_dst(t) : (t & loc_mask) * AMP;
        }
        gp.st_fixed = hp.fused_swap ? (hp.hdr.dst_fixed & loc_mask) * AMP : 0;
        gp.n_xo = hp.hdr.n_xo;
        for (uint32_t k = 0; k < hp.hdr.n_xo; k++) { gp.xo_pos[k] = hp.hdr.xo_pos[k]; gp.xo_rank[k] = hp.hdr.xo_rank[k]; gp.xo_mask |= 1ULL << hp.hdr.xo_pos[k]; }
        if (nrounds == 1 && !hp.hdr.out_of_place) {   /* a thread's store set differs from its load set: barrier before the scatter */
            bool moved = false;
            for (int j = 0; j < QSB_TB; j++) moved |= gp.ld_thr[j] != gp.st_thr[j];
            for (int v = 0; v < QSB_NV; v++) moved |= gp.ld_vec[v] != gp.st_vec[v];
            if (moved) gp.flags |= QSB_PASS_SYNC_SCATTER;
        }

        std::vector<GRound> gr(nrounds);
        std::vector<std::vector<uint8_t>> segstream(nrounds), bodystream(nrounds), tphstream(nrounds), angstream(nrounds);
        /* body offsets inside segstream entries are relative to the round's body; fixed up below */
        struct SegRec { uint32_t n_special, special_rel, n_groups, group_rel; };
        std::vector<std::vector<SegRec>> segrec(nrounds);
        const int SET16 = QSB_SET16(f32), GROUP16 = QSB_GROUP16(f32);
        uint32_t n_cond = 0;
        /* angle entries (round-level thread phases and merged phase runs): fixed-point turn fraction + ONE 32-bit mask over
         * the per-thread predicate word (thread bits, then the pass's outer-condition bits: angle_mask below) */
        auto turn_fraction = [](double pr, double pi) {   /* e^{2 pi i f}, f as a 64-bit fixed-point fraction of a turn */
            long double turns = (long double)atan2(pi, pr) / (2.0L * 3.14159265358979323846264338327950288L);
            /* exact fractions for the phases circuits are made of (-1, +-i, e^{i pi/4}, pi/2^k ladders) */
            if (pi == 0.0) turns = pr > 0 ? 0.0L : 0.5L;
            else if (pr == 0.0) turns = pi > 0 ? 0.25L : 0.75L;
            turns -= floorl(turns);
            long double scaled = roundl(ldexpl(turns, 64));
            if (scaled >= ldexpl(1.0L, 64)) scaled = 0.0L;
            return (uint64_t)scaled;
        };
        auto put_angle = [&](std::vector<uint8_t> &o, uint32_t m32, uint64_t ang64) {
            if (f32) { GTAngle32 e; e.mask = m32; e.ang32 = (uint32_t)((ang64 + 0x80000000ULL) >> 32);   /* top 32 bits, rounded; wraps to 0 at one turn */
                       const uint8_t *q = (const uint8_t *)&e; o.insert(o.end(), q, q + sizeof e); }
            else { GTAngle64 e; e.mask = m32; e.pad = 0; e.ang64 = ang64; const uint8_t *q = (const uint8_t *)&e; o.insert(o.end(), q, q + sizeof e); }
        };
        for (int r = 0; r < nrounds; r++) {
            const DevRound &D = hp.rounds[r];
            GRound &G = gr[r]; memset(&G, 0, sizeof G);
            G.flags = D.flags;
            for (int j = 0; j < QSB_TB; j++) G.thr_x[j] = ((uint32_t)D.thr[j].ld * 16u) | (((uint32_t)D.thr[j].st * 16u) << 16);
            for (int b = 0; b < QSB_NVB; b++) { G.vld_b[b] = (uint32_t)D.vec[b].ld * 16u; G.vst_b[b] = (uint32_t)D.vec[b].st * 16u; }
            /* predicate masks: source index bit -> thread bit of this round, or outer */
            auto split_mask = [&](uint64_t tmask, uint32_t &tm8, uint64_t &om) {
                tm8 = 0; om = 0;
                for (uint64_t m = tmask; m; m &= m - 1) {
                    const uint64_t bit = m & (~m + 1);
                    int j = -1;
                    for (int k = 0; k < QSB_TB; k++) if (D.thr[k].gidx == bit) j = k;
                    if (j >= 0) tm8 |= 1u << j; else om |= bit;
                }
            };
            /* outer condition -> bit of W (pass-wide table); false if the table is full */
            auto cond_bit = [&](uint64_t om, uint32_t &wbits) {
                wbits = 0;
                if (!om) return true;
                for (uint32_t i = 0; i < n_cond; i++) if (gp.cond[i] == om) { wbits = 1u << i; return true; }
                if (n_cond >= QSB_MAX_COND) return false;
                gp.cond[n_cond] = om; wbits = 1u << n_cond; n_cond++;
                return true;
            };
            /* mask of an angle entry over the predicate word: thread bits + one single-bit outer condition per outer control.
             * false if the outer-condition table is full (the caller keeps the gate in its plain form) */
            auto angle_mask = [&](uint32_t tm8, uint64_t om, uint32_t &m32) {
                m32 = tm8;
                for (uint64_t m = om; m; m &= m - 1) {
                    uint32_t wb;
                    if (!cond_bit(m & (~m + 1), wb)) return false;
                    m32 |= wb << QSB_TB;
                }
                return true;
            };
            std::vector<uint8_t> &ts = tphstream[r];
            uint32_t n_tph = 0, n_ang = 0;
            /* unit-modulus thread-level phases travel as fixed-point angles when the round has enough of them */
            auto is_unit_phase = [](const HostOp &h) { return fabs(h.tph[0] * h.tph[0] + h.tph[1] * h.tph[1] - 1.0) <= 1e-15; };
            int n_unit = 0;
            for (uint32_t k = hp.round_op_begin[r]; k < hp.round_op_begin[r] + hp.round_op_count[r]; k++)
                if ((hp.ops[k].kind & 0xff) == OP_TPHASE && is_unit_phase(hp.ops[k])) n_unit++;
            const bool use_angles = n_unit >= (f32 ? QSB_TANGLE_MIN_F32 : QSB_TANGLE_MIN_F64);

            /* G_DIAGA pre-scan: runs of mergeable controlled phases per vector bit.  Everything between two members of a
             * run commutes with them (ops on other vector bits never name this one as a control; phases are diagonal), so
             * the whole run is lowered at its first member.  A run ends at any other op whose target is that vector bit. */
            const uint32_t rb = hp.round_op_begin[r], rn = hp.round_op_count[r];
            std::vector<int> run_of(rn, -1);                    /* op -> run id, or -1 */
            std::vector<std::vector<uint32_t>> runs;
            {
                auto lanes_eq = [](const HostOp &h) { for (int c = 0; c < h.n_coef; c++) if (h.c[0][c][0] != h.c[0][c][1]) return false; return true; };
                auto mergeable = [&](const HostOp &h) {
                    if ((h.kind & 0xff) != OP_DIAG_V || ((h.kind >> 16) & 1) || !lanes_eq(h)) return false;
                    const double pr = h.c[0][0][1], pi = h.c[0][1][1];
                    return fabs(pr * pr + pi * pi - 1.0) <= 1e-15;
                };
                /* run QSB_NVB: phases whose target is the PACK qubit (OP_DIAG_ALL: low lane 1, high lane the phase) under
                 * thread-level controls.  Only a matrix on the pack qubit ends it: everything else either is diagonal or
                 * uses the pack qubit as a control. */
                auto mergeable_pack = [&](const HostOp &h) {
                    if ((h.kind & 0xff) != OP_DIAG_ALL || ((h.kind >> 16) & 1)) return false;
                    if (h.c[0][0][0] != 1.0 || h.c[0][1][0] != 0.0) return false;
                    const double pr = h.c[0][0][1], pi = h.c[0][1][1];
                    return fabs(pr * pr + pi * pi - 1.0) <= 1e-15;
                };
                int open_run[QSB_NVB + 1]; for (int b = 0; b <= QSB_NVB; b++) open_run[b] = -1;
                std::vector<char> pack_run;                         /* run id -> it is a pack-qubit run */
                for (uint32_t k = 0; k < rn; k++) {
                    const HostOp &h = hp.ops[rb + k];
                    const int code = h.kind & 0xff, vb = (h.kind >> 8) & 0xf;
                    if (code == OP_MATP_R || code == OP_MATP_G) { open_run[QSB_NVB] = -1; continue; }
                    uint32_t tmm, amm; uint64_t omm;
                    split_mask(h.tmask, tmm, omm);
                    if (code == OP_DIAG_ALL) {
                        if (mergeable_pack(h) && angle_mask(tmm, omm, amm)) {
                            if (open_run[QSB_NVB] < 0) { open_run[QSB_NVB] = (int)runs.size(); runs.emplace_back(); pack_run.push_back(1); }
                            runs[open_run[QSB_NVB]].push_back(k); run_of[k] = open_run[QSB_NVB];
                        }
                        continue;
                    }
                    if (code == OP_TPHASE || code == OP_DIAG_GEN) continue;
                    if (mergeable(h) && angle_mask(tmm, omm, amm)) {
                        if (open_run[vb] < 0) { open_run[vb] = (int)runs.size(); runs.emplace_back(); pack_run.push_back(0); }
                        runs[open_run[vb]].push_back(k); run_of[k] = open_run[vb];
                    } else open_run[vb] = -1;
                }
                const size_t dga_min = M.diaga ? (f32 ? QSB_DIAGA_MIN_F32 : QSB_DIAGA_MIN_F64) : ((size_t)1 << 30);
                const size_t dga_min_pack = M.diaga ? QSB_DIAGA_MIN_PACK : ((size_t)1 << 30);
                for (size_t i = 0; i < runs.size(); i++)
                    if (runs[i].size() < (pack_run[i] ? dga_min_pack : dga_min)) { for (uint32_t k : runs[i]) run_of[k] = -1; runs[i].clear(); }
            }
            std::vector<char> run_done(runs.size(), 0);
            unsigned tr_slot = 0, tr_xmerge = 0, tr_run = 0, tr_run_members = 0, tr_code[G_NCODES] = {0};   /* QSB_PLAN_TRACE */
            /* segment under construction */
            std::vector<uint8_t> specials; uint32_t n_special = 0;
            std::vector<std::vector<uint8_t>> groups;       /* each QSB_GROUP16 * 16 bytes */
            std::vector<std::array<bool, QSB_NVB>> slot_single;   /* slot holds an unconditional single-set gate */
            int next_group[QSB_NVB]; for (int b = 0; b < QSB_NVB; b++) next_group[b] = 0;
            auto close_segment = [&]() {
                if (!n_special && groups.empty()) return;
                SegRec sr; sr.n_special = n_special; sr.special_rel = (uint32_t)bodystream[r].size();
                bodystream[r].insert(bodystream[r].end(), specials.begin(), specials.end());
                sr.n_groups = (uint32_t)groups.size(); sr.group_rel = (uint32_t)bodystream[r].size();
                for (auto &g : groups) bodystream[r].insert(bodystream[r].end(), g.begin(), g.end());
                segrec[r].push_back(sr);
                specials.clear(); n_special = 0; groups.clear(); slot_single.clear();
                for (int b = 0; b < QSB_NVB; b++) next_group[b] = 0;
            };

            for (uint32_t k = hp.round_op_begin[r]; k < hp.round_op_begin[r] + hp.round_op_count[r]; k++) {
                const HostOp &h = hp.ops[k];
                const int code = h.kind & 0xff, vb = (h.kind >> 8) & 0xf, vmask = OPK_VMASK(h.kind);
                const bool mux = (h.kind >> 16) & 1;
                uint32_t tm8; uint64_t om;
                split_mask(h.tmask, tm8, om);
                uint32_t am32 = 0;
                if (code == OP_TPHASE && use_angles && is_unit_phase(h) && angle_mask(tm8, om, am32)) {
                    put_angle(angstream[r], am32, turn_fraction(h.tph[0], h.tph[1]));
                    n_ang++;
                    continue;
                }
                if (code == OP_TPHASE) {
                    GTPhase e; memset(&e, 0, sizeof e);
                    e.tmask = tm8; e.omask = om;
                    if (f32) { float v[2] = {(float)h.tph[0], (float)h.tph[1]}; memcpy(e.val, v, 8); }
                    else memcpy(e.val, h.tph, 16);
                    const uint8_t *q = (const uint8_t *)&e; ts.insert(ts.end(), q, q + sizeof e);
                    n_tph++;
                    continue;
                }
                if (run_of[k - rb] >= 0) {
                    /* member of a run of controlled phases on vector bit vb: the whole run is ONE special, lowered at its first member */
                    const int run = run_of[k - rb];
                    if (run_done[run]) continue;
                    run_done[run] = 1;
                    tr_run++; tr_run_members += (unsigned)runs[run].size();
                    if (!groups.empty()) close_segment();
                    std::vector<uint8_t> ents;
                    for (uint32_t km : runs[run]) {
                        const HostOp &hm = hp.ops[rb + km];
                        uint32_t tmm, amm; uint64_t omm; split_mask(hm.tmask, tmm, omm);
                        if (!angle_mask(tmm, omm, amm)) { qsb_set_error("internal: angle word overflow in a merged run"); return QSB_ERR_ARG; }   /* the pre-scan reserved the bits */
                        put_angle(ents, amm, turn_fraction(hm.c[0][0][1], hm.c[0][1][1]));
                    }
                    pad16(ents);                                  /* f32: an odd count ends with a zero entry (adds nothing) */
                    std::vector<uint8_t> body(16, 0);
                    const uint32_t n_u = (uint32_t)(ents.size() / 16);
                    memcpy(body.data(), &n_u, 4);
                    body.insert(body.end(), ents.begin(), ents.end());
                    const size_t bytes = 16 + body.size();
                    if (bytes / 16 > 0xffff) { qsb_set_error("internal: merged phase run too long"); return QSB_ERR_ARG; }
                    uint32_t hdr[4] = {GOPK(G_DIAGA + (code == OP_DIAG_ALL ? QSB_NVB : vb), 0, 0, 0, bytes / 16), 0, 0, 0};
                    const uint8_t *q = (const uint8_t *)hdr; specials.insert(specials.end(), q, q + 16);
                    specials.insert(specials.end(), body.begin(), body.end());
                    n_special++;
                    continue;
                }
                const bool cond = (tm8 | om) != 0;
                /* lanes differ (a control on the pack qubit): scalar-coefficient forms cannot express it */
                auto lanes_equal = [&](int set) { for (int c = 0; c < h.n_coef; c++) if (h.c[set][c][0] != h.c[set][c][1]) return false; return true; };
                const bool leq = lanes_equal(0) && (!mux || lanes_equal(1));

                /* ---- slot forms ---- */
                int sform = S_SKIP;
                if (leq) switch (code) {
                    case OP_MAT_U: sform = S_UNIT_R; break;
                    case OP_MAT_UI: sform = S_UNIT_I; break;
                    case OP_DIAG_V: sform = S_DIAG; break;
                    case OP_XDEF: sform = S_XDEF; break;
                    default: break;
                }
                uint32_t wbits = 0;
                if (sform != S_SKIP && cond_bit(om, wbits)) {
                    std::vector<uint8_t> slot;
                    auto slot_set = [&](int set, bool identity) {
                        auto C = [&](int c) { return h.c[set][c][1]; };
                        std::vector<uint8_t> o;
                        switch (sform) {
                        case S_UNIT_R: if (identity) { put_s(o, 0); put_s(o, 0); put_s(o, 1); put_s(o, 1); } else { put_s(o, C(0)); put_s(o, C(1)); put_s(o, C(2)); put_s(o, C(3)); } break;
                        case S_UNIT_I: if (identity) { put_s(o, 0); put_s(o, 0); put_s(o, 1); put_s(o, 1); } else { put_s(o, C(0)); put_s(o, C(2)); put_s(o, C(4)); put_s(o, C(5)); } break;
                        case S_DIAG: if (identity) { put_s(o, 1); put_s(o, 0); put_s(o, 0); put_s(o, 0); } else { put_s(o, C(0)); put_s(o, C(1)); put_s(o, 0); put_s(o, 0); } break;
                        default: put_s(o, 0); put_s(o, 0); put_s(o, 0); put_s(o, 0); break;
                        }
                        slot.insert(slot.end(), o.begin(), o.end());
                    };
                    /* Hadamard-like: unconditional real unit form with q == 1 and the scale already in the pass scale */
                    const bool hlike = sform == S_UNIT_R && !mux && !cond && h.c[0][1][1] == 1.0 && h.c[0][3][1] == 1.0 && M.hform;
                    if (mux) { slot_set(0, false); slot_set(1, false); }
                    else if (hlike) { slot_set(0, false); slot_set(0, false); }   /* both sets usable: an X merged later brings a predicate */
                    else { slot_set(0, true); slot_set(0, false); }
                    if (hlike) sform = S_UNIT_H;
                    if ((int)slot.size() != 2 * SET16 * 16) { qsb_set_error("internal: slot of %zu bytes", slot.size()); return QSB_ERR_ARG; }
                    const uint32_t pm_new = tm8 | (wbits << QSB_TB);
                    if (sform == S_XDEF && next_group[vb] > 0) {
                        /* merge a deferred X into the gate it follows on the same vector bit: same slot, same predicate.
                         * An unconditional single-set gate takes the X's predicate (both coefficient sets identical). */
                        uint8_t *Gp = groups[next_group[vb] - 1].data();
                        const uint8_t pf = Gp[vb];
                        uint32_t ppm; memcpy(&ppm, Gp + QSB_GROUP_MASK_OFF(vb), 4);
                        uint8_t *sets = Gp + 32 + (size_t)vb * 2 * SET16 * 16;
                        const bool prev_uncond = ppm == 0 && slot_single[next_group[vb] - 1][vb];
                        if ((pf & (S_UNIT_R | S_UNIT_I | S_UNIT_H | S_DIAG)) && !(pf & S_XDEF) && (ppm == pm_new || prev_uncond)) {
                            if (ppm != pm_new) { memcpy(sets, sets + (size_t)SET16 * 16, (size_t)SET16 * 16); memcpy(Gp + QSB_GROUP_MASK_OFF(vb), &pm_new, 4); }
                            Gp[vb] = (uint8_t)(pf | S_XDEF);
                            tr_xmerge++;
                            continue;
                        }
                    }
                    tr_slot++;
                    const int g = next_group[vb]++;
                    if (g == (int)groups.size()) { groups.push_back(std::vector<uint8_t>((size_t)GROUP16 * 16, 0)); slot_single.push_back(std::array<bool, QSB_NVB>{}); }
                    slot_single[g][vb] = !mux && !cond;
                    uint8_t *G0 = groups[g].data();
                    G0[vb] = (uint8_t)sform;                                   /* form byte of slot vb */
                    memcpy(G0 + QSB_GROUP_MASK_OFF(vb), &pm_new, 4);           /* predicate mask */
                    memcpy(G0 + 32 + (size_t)vb * 2 * SET16 * 16, slot.data(), slot.size());
                    continue;
                }

                /* ---- specials (generic interpreter) ---- */
                if (!groups.empty()) close_segment();     /* a special runs before the groups of its segment */
                const int two = mux ? 1 : 0, skip = (!mux && cond) ? 1 : 0;
                std::vector<uint8_t> sets[2];
                int gcode = -1;
                auto write_set = [&](std::vector<uint8_t> &o, int set) {
                    const double zero[2] = {0.0, 0.0};
                    auto L = [&](int c) { return h.c[set][c]; };
                    /* complex 2x2 with lane pairs: m00r m00i m01r m01i m10r m10i m11r m11i */
                    auto put_gen = [&](const double (*m)[2]) { for (int c = 0; c < 8; c++) put_v(o, m[c]); };
                    double m[8][2];
                    for (int c = 0; c < 8; c++) m[c][0] = m[c][1] = 0.0;
                    switch (code) {
                    case OP_MAT_U: case OP_MAT_UI: {   /* only when the outer-condition table is full */
                        gcode = G_FULL_G + vb;
                        for (int l = 0; l < 2; l++) {
                            if (code == OP_MAT_U) {
                                const double pp = h.c[set][0][l], q = h.c[set][1][l], kk = h.c[set][2][l], a = h.c[set][3][l];
                                m[0][l] = a; m[2][l] = a * pp; m[4][l] = a * q; m[6][l] = a * (kk + q * pp);
                            } else {
                                const double pp = h.c[set][0][l], q = h.c[set][2][l], kk = h.c[set][4][l], a = h.c[set][5][l];
                                m[0][l] = a; m[3][l] = a * pp; m[5][l] = a * q; m[6][l] = a * (kk - q * pp);
                            }
                        }
                        put_gen(m);
                        break;
                    }
                    case OP_XDEF:
                        gcode = G_FULL_G + vb;
                        m[2][0] = m[2][1] = 1.0; m[4][0] = m[4][1] = 1.0;
                        put_gen(m);
                        break;
                    case OP_MAT_R:
                        gcode = G_FULL_G + vb;
                        put_v(o, L(2)); put_v(o, zero); put_v(o, L(0)); put_v(o, zero); put_v(o, L(1)); put_v(o, zero); put_v(o, L(3)); put_v(o, zero);
                        break;
                    case OP_MAT_I:
                        gcode = G_FULL_G + vb;
                        put_v(o, L(4)); put_v(o, zero); put_v(o, zero); put_v(o, L(1)); put_v(o, zero); put_v(o, L(3)); put_v(o, L(5)); put_v(o, zero);
                        break;
                    case OP_MAT_G:
                        gcode = G_FULL_G + vb;
                        put_v(o, L(6)); put_v(o, L(0)); put_v(o, L(1)); put_v(o, L(2)); put_v(o, L(3)); put_v(o, L(4)); put_v(o, L(7)); put_v(o, L(5));
                        break;
                    case OP_MATP_R:
                        gcode = G_MATP_R; put_v(o, L(0)); put_v(o, L(1));
                        break;
                    case OP_MATP_G:
                        gcode = G_MATP_G; put_v(o, L(0)); put_v(o, L(1)); put_v(o, L(2)); put_v(o, L(3));
                        break;
                    case OP_DIAG_V: case OP_DIAG_ALL: case OP_DIAG_GEN:
                        gcode = code == OP_DIAG_V ? G_DIAG_V + vb : code == OP_DIAG_ALL ? G_DIAG_ALL : G_DIAG_GEN;
                        put_v(o, L(0)); put_v(o, L(1));
                        break;
                    default: break;
                    }
                    pad16(o);
                };
                write_set(sets[0], 0);
                if (mux) write_set(sets[1], 1);
                if (gcode < 0) { qsb_set_error("internal: op code %d cannot be lowered", code); return QSB_ERR_ARG; }
                /* an unconditional unit-form op that fell through keeps its scale in the matrix: nothing else to do.
                 * A conditional XDEF / unit op lowered to FULL_G is a plain controlled gate (skip when the predicate fails). */
                const size_t bytes = 16 + sets[0].size() + sets[1].size();
                uint32_t hdr[4] = {GOPK(gcode, two, skip, vmask, bytes / 16), tm8, (uint32_t)om, (uint32_t)(om >> 32)};
                const uint8_t *q = (const uint8_t *)hdr; specials.insert(specials.end(), q, q + 16);
                specials.insert(specials.end(), sets[0].begin(), sets[0].end());
                specials.insert(specials.end(), sets[1].begin(), sets[1].end());
                n_special++;
                if (gcode < G_NCODES) tr_code[gcode]++;
            }
            close_segment();
            pad16(angstream[r]);
            G.n_tph = n_tph; G.n_ang = (uint32_t)(angstream[r].size() / 16); G.n_seg = (uint32_t)segrec[r].size();   /* n_ang: 16-byte units */
            if (plan_trace()) {   /* host-side op mix of the lowered round (stderr); tools and DESIGN.md quote it */
                unsigned full = 0, dv = 0;
                for (int b = 0; b < QSB_NVB; b++) { full += tr_code[G_FULL_G + b]; dv += tr_code[G_DIAG_V + b]; }
                fprintf(stderr, "qsb-plan: round %d: segments %u, slots %u (+%u merged X), phase runs %u (%u gates), FULL_G %u, DIAG_V %u, "
                                "DIAG_ALL %u, DIAG_GEN %u, MATP %u, thread phases %u, angle phases %u\n", r, G.n_seg, tr_slot, tr_xmerge,
                        tr_run, tr_run_members, full, dv, tr_code[G_DIAG_ALL], tr_code[G_DIAG_GEN], tr_code[G_MATP_R] + tr_code[G_MATP_G], n_tph, n_ang);
            }
        }
        gp.n_cond = n_cond;
        size_t off = al16(sizeof(GPass));
        gp.rounds_off16 = (uint32_t)(off / 16);
        std::vector<size_t> round_at(nrounds);
        for (int r = 0; r < nrounds; r++) {
            round_at[r] = off;
            off += al16(sizeof(GRound));
            gr[r].seg_off16 = (uint32_t)(off / 16);
            const size_t body0 = off + segrec[r].size() * sizeof(GSegment);
            for (const SegRec &sr : segrec[r]) {
                GSegment gs; gs.n_special = sr.n_special; gs.special_off16 = (uint32_t)((body0 + sr.special_rel) / 16);
                gs.n_groups = sr.n_groups; gs.group_off16 = (uint32_t)((body0 + sr.group_rel) / 16);
                const uint8_t *q = (const uint8_t *)&gs; segstream[r].insert(segstream[r].end(), q, q + sizeof gs);
            }
            off = body0 + bodystream[r].size();
            gr[r].tph_off16 = (uint32_t)(off / 16); off += tphstream[r].size() + angstream[r].size();
            gr[r].next16 = (uint32_t)(off / 16);
        }
        const size_t total = off + 64;   /* slack: the group loop prefetches one group header past the last group */
        if (total > QSB_BLOB_LARGE) { qsb_set_error("internal: pass descriptor of %zu bytes exceeds the limit", total); return QSB_PLAN_OVERFLOW; }
        hp.hdr.blob_bytes = (uint32_t)total;
        std::vector<uint8_t> &b = hp.blob;
        b.assign(total <= QSB_BLOB_SMALL ? QSB_BLOB_SMALL : total <= QSB_BLOB_MEDIUM ? QSB_BLOB_MEDIUM : QSB_BLOB_LARGE, 0);
        memcpy(&b[0], &gp, sizeof gp);
        for (int r = 0; r < nrounds; r++) {
            memcpy(&b[round_at[r]], &gr[r], sizeof(GRound));
            size_t at = (size_t)gr[r].seg_off16 * 16;
            if (!segstream[r].empty()) memcpy(&b[at], segstream[r].data(), segstream[r].size());
            at += segstream[r].size();
            if (!bodystream[r].empty()) memcpy(&b[at], bodystream[r].data(), bodystream[r].size());
            if (!tphstream[r].empty()) memcpy(&b[(size_t)gr[r].tph_off16 * 16], tphstream[r].data(), tphstream[r].size());
            if (!angstream[r].empty()) memcpy(&b[(size_t)gr[r].tph_off16 * 16 + tphstream[r].size()], angstream[r].data(), angstream[r].size());
        }
        return QSB_OK;
    }
};

/* ----------------------------------------------------------------- scheduler */
int tiled_schedule(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                   const std::vector<COp> &cops_in, const double gphase[2], TiledPlan *plan, int climb_variant)
{
    Machine M;
    M.n = n; M.prec = prec; M.g = g; M.nloc = nloc; M.rank = rank;
    M.f32 = (prec == QSB_F32);
    M.T = M.f32 ? QSB_T_F32 : QSB_T_F64;
    M.lazy_diag = opt && opt->reserved[1] == 2;      /* reserved[1] = 2: keep the qubits of phase gates thread-level (A/B runs;
                                                        same speed on random circuits, 2.4x more rounds on QFT) */
    M.trim_thin = (opt && opt->reserved[2] > 0) ? opt->reserved[2] - 1 : 2;   /* reserved[2] = k+1: trim tail rounds with < k gates (1 = off) */
    M.diaga = !(opt && opt->reserved[4] == 4);       /* reserved[4] = 4: no merged controlled phases G_DIAGA (A/B runs) */
#ifdef QSB_UNIT_H
    M.hform = !(opt && opt->reserved[4] == 3);
#else
    M.hform = false;                                 /* S_UNIT_H is not compiled into the default kernel */
#endif       /* reserved[4] = 3: no Hadamard-like slot form S_UNIT_H (A/B runs) */
    M.defer_diag = !(opt && opt->reserved[4] == 1);  /* reserved[4] = 1: do not defer vector-bit phase gates (A/B runs) */
    M.force_top = g > 0 && opt && opt->reserved[5] == 3;        /* reserved[5] = 3: NCCL-style plan executed as a pipelined exchange */
    M.fused_exchange = g > 0 && opt && (opt->reserved[5] == 1 || opt->reserved[5] == 4);   /* reserved[5] = 1 / 4: exchanges fused into a pass (peer stores) */
    M.fused_direct = g > 0 && opt && opt->reserved[5] == 1;     /* 1: victims trade places with the rank bits wherever they are (round 2); 4: round 1
                                                                   flavour, victims moved to the top local positions first (A/B) */
    M.tile_search = !(opt && opt->reserved[6] == 2);   /* reserved[6] = 2: first-come tile choice (A/B runs) */
    M.sink_phases = !(opt && opt->reserved[6] == 1);   /* reserved[6] = 1: keep thread-level phases in the round that accepted them (A/B) */
    /* reserved[6] = 3: no lane relocation at the end of a pass (an empty round turns the registers instead; round 1 / A-B);
     * 4: relocate only the qubits in conflict; default: every pass re-picks the qubits that live on the lanes */
    M.lane_reloc = (opt && opt->reserved[6] == 3) ? 0 : (opt && opt->reserved[6] == 4) ? 2 : 4;
    M.cost_cap = opt ? opt->reserved[3] : 0;   /* reserved[3] = k: stop adding rounds to a pass once its estimated SM cost reaches k gate units */
    M.nb = 3;   /* 16-byte shared-memory slots in both precisions: 8 lanes per 128-bit access phase */
    M.a = opt && opt->low_bits > 0 ? opt->low_bits : (M.f32 ? 4 : 3);   /* measured optimum on B200: DESIGN.md §5 */
    if (M.a < (M.f32 ? 3 : 2) || M.a > (M.f32 ? 6 : 5)) { qsb_set_error("low_bits %d unsupported for this precision", M.a); return QSB_ERR_ARG; }
    if (nloc < M.T) { qsb_set_error("internal: local register smaller than a tile"); return QSB_ERR_ARG; }
    if (g > 0 && nloc - g < M.a + g) { qsb_set_error("local register too small to exchange %d qubits", g); return QSB_ERR_ARG; }

    plan->n = n; plan->prec = prec; plan->g = g; plan->nloc = nloc; plan->rank = rank;
    plan->start_perm = start;
    plan->passes.clear();

    std::vector<COp> cops = cops_in;
    int8_t wire[64];
    relabel_swaps(cops, n, wire);
    double gph[2] = {gphase[0], gphase[1]};
    /* reserved[4] = 5: CX stays CX next to an h (A/B); 6: rewritten with an h on one side too (the build of GPU call 33) */
    if (!(opt && (opt->reserved[4] == 2 || opt->reserved[4] == 5))) cx_through_h(cops, gph, opt && opt->reserved[4] == 6);
    if (!(opt && opt->reserved[4] == 2)) fuse_same_qubit(cops, n, gph);      /* reserved[4] = 2: no 2x2 products (A/B, tests) */
    absorb_cx(cops, n);
    if (!(gph[0] == 1.0 && gph[1] == 0.0)) {
        COp c; memset(&c, 0, sizeof c); c.kind = C_PHASE; c.target = -1; c.ctrl = 0; c.m[0] = gph[0]; c.m[1] = gph[1];
        cops.push_back(c);
    }
    BitPerm perm = start;
    const size_t N = cops.size();
    std::vector<char> done(N, 0);
    size_t left = N, first_open = 0;
    const int SWAP_MIN_OPS = opt && opt->reserved[0] > 0 ? opt->reserved[0] : 10;

    /* greedy op collection for one pass.  S0/n0: qubits / slots already resident. */
    bool strict_count = false;
    auto collect = [&](uint64_t S0, int n0, std::vector<COp> &mine, std::vector<size_t> &mine_idx, uint64_t &S_out, int op_limit = 0) {
        uint64_t S = S0; int nS = n0;
        Blocker B; B.clear();
        mine.clear(); mine_idx.clear();
        /* Budget of the pass descriptor.  A matrix op may take a whole group (the worst case max_pass_ops is sized
         * for); a phase gate is usually one 16/32-byte entry of a thread-phase list, so it is charged a quarter
         * (QFT ladders: hundreds of phases per pass).  Should the optimistic count overflow the descriptor after
         * all, the caller retries with `strict` (every op charged in full). */
        const int max_ops = op_limit > 0 ? op_limit : max_pass_ops(M.f32);
        const bool weighted = op_limit == 0 && !strict_count;
        int budget = 4 * max_ops;
        for (size_t i = first_open; i < N && budget > 0; i++) {
            if (done[i]) continue;
            const COp &o = cops[i];
            bool can = B.ok(o);
            if (can && o.target >= 0) {
                if (perm.pos[o.target] >= nloc) can = false;
                else if ((S >> o.target) & 1) {}
                else if (nS < M.T) { S |= 1ULL << o.target; nS++; }
                else can = false;
            }
            if (can) { mine.push_back(o); mine_idx.push_back(i); budget -= (weighted && o.kind == C_PHASE) ? 1 : 4; }
            else { B.block(o); if (B.full >= n) break; }
        }
        S_out = S;
    };
    /* ops a pass would run if exactly the qubits of S were resident (no growth): the score of a tile choice */
    /* (the hill climbing calls this some 500 times per pass: it runs on a compact copy of the op list -- 16 bytes per
     * op instead of the 160-byte COp -- with the blocker's two levels as bit masks) */
    struct LiteOp { uint64_t ctrl; int8_t target; uint8_t phase; };
    std::vector<LiteOp> lite(N);
    for (size_t i = 0; i < N; i++) { lite[i].ctrl = cops[i].ctrl; lite[i].target = (int8_t)cops[i].target; lite[i].phase = cops[i].kind == C_PHASE; }
    auto count_fixed = [&](uint64_t S) -> int {
        uint64_t any = 0, hard = 0;     /* qubits blocked at level >= 1 / at level 2 (Blocker) */
        uint64_t local = 0;
        for (int q = 0; q < n; q++) if (perm.pos[q] < nloc) local |= 1ULL << q;
        const uint64_t resident = S & local;
        int score = 0, budget = 4 * max_pass_ops(M.f32), full = 0;
        const LiteOp *L = lite.data();
        const char *dn = done.data();
        for (size_t i = first_open; i < N && budget > 0; i++) {
            if (dn[i]) continue;
            const LiteOp &o = L[i];
            const uint64_t tbit = o.target >= 0 ? 1ULL << o.target : 0;
            bool can = !(tbit & any) && !(o.ctrl & hard);
            if (can && tbit && !(tbit & resident)) can = false;
            if (can) { score += tbit ? 8 : 1; budget -= o.phase ? 1 : 4; }
            else {
                if (tbit && !(hard & tbit)) { hard |= tbit; full++; }
                any |= tbit | o.ctrl;
                if (full >= n) break;
            }
        }
        return score;
    };
    /* hill climbing over the tile: trade one chosen qubit for one outside while the pass gets more ops */
    /* climb_variant: the order in which the hill climbing tries its swaps (bit 0: qubits leaving the tile from the top,
     * bit 1: qubits entering from the top, bit 2: best instead of first improvement).  The climb stops in a local optimum
     * and the variants end in different ones, a pass more or less on the 30-34 q workloads: tiled_plan_build plans a few
     * and keeps the cheapest schedule. */
    const int xvar = climb_variant;
    auto improve_tile = [&](uint64_t lowS, uint64_t S) -> uint64_t {
        int best = count_fixed(S);
        for (int sweep = 0; sweep < 3; sweep++) {
            bool any = false;
            for (int qo_ = 0; qo_ < n; qo_++) {
                const int qo = (xvar & 1) ? n - 1 - qo_ : qo_;
                if (!((S >> qo) & 1) || ((lowS >> qo) & 1)) continue;
                uint64_t bestS = S; int bestc = best;
                for (int qi_ = 0; qi_ < n; qi_++) {
                    const int qi = (xvar & 2) ? n - 1 - qi_ : qi_;
                    if (((S >> qi) & 1) || perm.pos[qi] >= nloc) continue;
                    const uint64_t S2 = (S & ~(1ULL << qo)) | (1ULL << qi);
                    const int c = count_fixed(S2);
                    if (c > bestc) { bestc = c; bestS = S2; if (!(xvar & 4)) break; }
                }
                if (bestc > best) { best = bestc; S = bestS; any = true; }
            }
            if (!any) break;
        }
        return S;
    };
    auto emit_pass = [&](uint64_t S, uint64_t forced_pos, const int8_t *pos_map, const std::vector<COp> &mine,
                         const std::vector<size_t> &mine_idx, bool allow_empty, bool fuse_exchange = false, const int *victim_pos = nullptr) -> int {
        PassBuilder pb(M, perm);
        pb.set_tile(S, forced_pos, pos_map, fuse_exchange, victim_pos);
        /* plain passes may end with a lane relocation (build_rounds); it wants to know when each qubit is next a target */
        pb.may_relocate = !pos_map && !fuse_exchange && !forced_pos;
        pb.next_use_cb = [&](const std::vector<char> &used_now, size_t *next_use) {
            std::vector<char> mine_used(N, 0);
            for (size_t k = 0; k < mine_idx.size() && k < used_now.size(); k++) if (used_now[k]) mine_used[mine_idx[k]] = 1;
            int found = 0;
            for (size_t i = first_open; i < N && found < n; i++) {
                if (done[i] || mine_used[i] || cops[i].target < 0) continue;
                const int qq = cops[i].target;
                if (next_use[qq] == ~(size_t)0) { next_use[qq] = i; found++; }
            }
        };
        std::vector<char> used;
        int rc = pb.build_rounds(mine, used, !allow_empty && mine_idx.size() < left);
        if (rc) return rc;
        rc = pb.serialise();
        if (rc) return rc;          /* QSB_PLAN_OVERFLOW: nothing is committed yet, the caller collects again */
        size_t consumed = 0;
        for (size_t k = 0; k < mine_idx.size(); k++) if (used[k]) { done[mine_idx[k]] = 1; left--; consumed++; }
        if (!consumed && !allow_empty) { qsb_set_error("scheduler made no progress inside a pass"); return QSB_ERR_ARG; }
        while (first_open < N && done[first_open]) first_open++;
        if (pos_map) for (int q = 0; q < n; q++) if (perm.pos[q] < nloc) perm.pos[q] = pos_map[perm.pos[q]];
        if (pb.relocated) for (int q = 0; q < n; q++) if (perm.pos[q] < nloc) perm.pos[q] = pb.reloc_map[perm.pos[q]];
        plan->passes.push_back(std::move(pb.hp));
        return QSB_OK;
    };

    /* Termination.  An ordinary pass always consumes ops; an exchange need not (its permutation pass runs "whatever
     * fits" with the victims' positions forced into the tile).  With a high exchange threshold on a circuit whose tail
     * never offers that many runnable gates, the scheduler could exchange for ever: a second exchange without an op
     * consumed since the first one is refused while this pass has anything to run, and a pass count that no schedule
     * can reach ends the planning with an error instead of a hang (a candidate of tiled_plan_search is then dropped). */
    size_t left_at_last_exchange = (size_t)-1;
    const size_t max_plan_steps = 4 * N + 1024;
    while (left) {
        if (plan->passes.size() > max_plan_steps) { qsb_set_error("internal: the scheduler does not terminate (%zu passes for %zu ops)", plan->passes.size(), N); return QSB_ERR_ARG; }
        uint64_t lowS = 0;
        for (int q = 0; q < n; q++) if (perm.pos[q] < M.a) lowS |= 1ULL << q;
        std::vector<COp> mine; std::vector<size_t> mine_idx; uint64_t S = 0;
        /* the low positions always occupy `a` tile slots, whether or not a logical qubit lives there */
        collect(lowS, M.a, mine, mine_idx, S);
        if (M.tile_search && !mine.empty()) {
            const uint64_t S2 = improve_tile(lowS, S);
            /* collect again with the tile fixed: every slot is taken, so no qubit joins */
            if (S2 != S) collect(S2, M.T, mine, mine_idx, S);
        }

        /* is the schedule about to starve behind a gate on a global qubit?  (few gates left that run without an exchange) */
        auto starving = [&](const std::vector<COp> &avail) {
            int nd = 0; for (const COp &o : avail) if (o.target >= 0) nd++;
            bool blocked_global = false;
            size_t seen = 0;
            for (size_t i = first_open; i < N && seen < (size_t)8 * n; i++) {
                if (done[i]) continue;
                seen++;
                if (cops[i].target >= 0 && perm.pos[cops[i].target] >= nloc) { blocked_global = true; break; }
            }
            return blocked_global && nd < SWAP_MIN_OPS;
        };
        bool want_swap = g > 0 && starving(mine);
        const bool exchanged_in_vain = left == left_at_last_exchange && !mine.empty();
        if (exchanged_in_vain) want_swap = false;
        if (M.fused_direct) {
            /* Direct fused exchange (round 2): victims trade places with the rank bits wherever they are, so the
             * exchange needs no tile slot and ANY pass can carry it.  Take the last pass that still has a full load of
             * gates: if the schedule would starve once this pass is through, this pass scatters into the peers. */
            bool fuse_now = want_swap;
            for (size_t k : mine_idx) done[k] = 1;            /* look past this pass */
            if (!fuse_now && !mine.empty() && !exchanged_in_vain) {
                std::vector<COp> m2; std::vector<size_t> i2; uint64_t S2 = 0;
                collect(lowS, M.a, m2, i2, S2);
                fuse_now = starving(m2);
            }
            int victim_pos[8]; std::vector<int> cand;
            if (fuse_now) {
                std::vector<size_t> next_use(n, N + 1);
                std::vector<char> seenq(n, 0); int found = 0;
                for (size_t i = first_open; i < N && found < n; i++) {
                    if (done[i] || cops[i].target < 0) continue;
                    const int qq = cops[i].target;
                    if (!seenq[qq]) { seenq[qq] = 1; next_use[qq] = i; found++; }
                }
                for (int qq = 0; qq < n; qq++) if (perm.pos[qq] >= M.a && perm.pos[qq] < nloc) cand.push_back(qq);
                std::stable_sort(cand.begin(), cand.end(), [&](int x, int y) {
                    if (next_use[x] != next_use[y]) return next_use[x] > next_use[y];
                    return perm.pos[x] > perm.pos[y];
                });
            }
            for (size_t k : mine_idx) done[k] = 0;
            if (fuse_now) {
                if ((int)cand.size() < g) { qsb_set_error("not enough local qubits to exchange"); return QSB_ERR_ARG; }
                for (int k = 0; k < g; k++) victim_pos[k] = perm.pos[cand[k]];
                int rc = emit_pass(S, 0, nullptr, mine, mine_idx, true, true, victim_pos);
                if (rc == QSB_PLAN_OVERFLOW) {
                    strict_count = true; collect(lowS, M.a, mine, mine_idx, S); strict_count = false;
                    rc = emit_pass(S, 0, nullptr, mine, mine_idx, true, true, victim_pos);
                }
                if (rc) return rc == QSB_PLAN_OVERFLOW ? QSB_ERR_ARG : rc;
                left_at_last_exchange = left;
                for (int qq = 0; qq < n; qq++) {          /* position victim_pos[k] <-> rank bit k */
                    const int pp = perm.pos[qq];
                    for (int k = 0; k < g; k++) {
                        if (pp == victim_pos[k]) perm.pos[qq] = (int8_t)(nloc + k);
                        else if (pp == nloc + k) perm.pos[qq] = (int8_t)victim_pos[k];
                    }
                }
                continue;
            }
            want_swap = false;
        }
        if (!want_swap) {
            if (mine.empty()) { qsb_set_error("scheduler made no progress (gate on a non-local qubit?)"); return QSB_ERR_ARG; }
            int rc = emit_pass(S, 0, nullptr, mine, mine_idx, false);
            if (rc == QSB_PLAN_OVERFLOW) {      /* the optimistic phase-gate budget did not fit: charge every op in full */
                strict_count = true; collect(lowS, M.a, mine, mine_idx, S); strict_count = false;
                rc = emit_pass(S, 0, nullptr, mine, mine_idx, false);
            }
            if (rc) return rc == QSB_PLAN_OVERFLOW ? QSB_ERR_ARG : rc;
            continue;
        }

        /* ---- global <-> local exchange --------------------------------------------------------
         * victims: the g local qubits whose next non-diagonal use lies furthest ahead (Belady).
         * They are moved to the top g local positions by an in-tile permutation pass (which also
         * runs whatever gates fit), then those positions are exchanged with the rank bits. */
        std::vector<size_t> next_use(n, N + 1);
        {
            std::vector<char> seenq(n, 0); int found = 0;
            for (size_t i = first_open; i < N && found < n; i++) {
                if (done[i] || cops[i].target < 0) continue;
                int q = cops[i].target;
                if (!seenq[q]) { seenq[q] = 1; next_use[q] = i; found++; }
            }
        }
        std::vector<int> cand;
        for (int q = 0; q < n; q++) if (perm.pos[q] >= M.a && perm.pos[q] < nloc) cand.push_back(q);
        std::stable_sort(cand.begin(), cand.end(), [&](int x, int y) {
            if (next_use[x] != next_use[y]) return next_use[x] > next_use[y];
            return perm.pos[x] > perm.pos[y];
        });
        if ((int)cand.size() < g) { qsb_set_error("not enough local qubits to exchange"); return QSB_ERR_ARG; }
        int8_t pos_map[64]; for (int p = 0; p < 64; p++) pos_map[p] = (int8_t)p;
        uint64_t forced = 0, vict_pos = 0, top_pos = 0;
        for (int k = 0; k < g; k++) { vict_pos |= 1ULL << perm.pos[cand[k]]; top_pos |= 1ULL << (nloc - g + k); }
        uint64_t v_only = vict_pos & ~top_pos, t_only = top_pos & ~vict_pos;
        while (v_only) {
            int pv = __builtin_ctzll(v_only), pt = __builtin_ctzll(t_only);
            v_only &= v_only - 1; t_only &= t_only - 1;
            pos_map[pv] = (int8_t)pt; pos_map[pt] = (int8_t)pv;
            forced |= (1ULL << pv) | (1ULL << pt);
        }
        const bool fuse = M.fused_exchange;
        /* the victim bits select the destination rank (fused exchange) or the chunk (pipelined exchange, which
         * slices the pass along its top outer bits): they must be tile bits */
        if (fuse || M.force_top) forced |= top_pos;
        if (forced) {
            /* resident set: low qubits + the qubits living on the forced positions */
            uint64_t S0 = lowS; int n0 = M.a;
            for (int q = 0; q < n; q++) if ((forced >> perm.pos[q]) & 1) { if (!((S0 >> q) & 1)) S0 |= 1ULL << q; }
            n0 += popc(forced);   /* forced positions are >= a: each takes a slot, qubit or padding */
            /* a fused-exchange pass also carries the peer table as a kernel parameter: keep its descriptor in the medium class */
            collect(S0, n0, mine, mine_idx, S, fuse ? max_pass_ops(M.f32) / QSB_FUSED_OP_DIV : 0);
            int rc = emit_pass(S, forced, pos_map, mine, mine_idx, true, fuse);
            if (rc == QSB_PLAN_OVERFLOW) {
                strict_count = true; collect(S0, n0, mine, mine_idx, S, fuse ? max_pass_ops(M.f32) / QSB_FUSED_OP_DIV : 0); strict_count = false;
                rc = emit_pass(S, forced, pos_map, mine, mine_idx, true, fuse);
            }
            if (rc) return rc == QSB_PLAN_OVERFLOW ? QSB_ERR_ARG : rc;
        }
        if (!fuse) {
            /* exchange marker: top g local positions <-> rank bits (NCCL all-to-all of contiguous chunks) */
            HostPass sw; memset(&sw.hdr, 0, sizeof sw.hdr);
            sw.is_swap = true; sw.hdr.nloc = nloc;
            plan->passes.push_back(std::move(sw));
        }
        for (int q = 0; q < n; q++) {
            int p = perm.pos[q];
            if (p >= nloc) perm.pos[q] = (int8_t)(p - g);
            else if (p >= nloc - g) perm.pos[q] = (int8_t)(p + g);
        }
        left_at_last_exchange = left;
    }
    /* logical qubit q ended on wire[q] (relabel_swaps) */
    for (int q = 0; q < 64; q++) plan->end_perm.pos[q] = q < n ? perm.pos[wire[q]] : perm.pos[q];
    return QSB_OK;
}

/* ------------------------------------------------------------------ plan search */
/* What the product runs: tiled_schedule for every setting of the knobs the caller left open, the cheapest schedule kept.
 * Host only (the CPU test doubles call it too, so that they emulate the very plan a GPU run would get). */
int tiled_plan_search(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                      const std::vector<COp> &cops, const double gphase[2], TiledPlan **out)
{
    TiledPlan *p = nullptr;
    const bool search_threshold = g > 0 && opt && opt->reserved[0] == 0;
    const bool search_lanes = g > 0 && opt && opt->reserved[6] == 0;
    const bool search_climb = g == 0 && cops.size() >= 64 && (!opt || (opt->reserved[6] == 0 && opt->reserved[3] == 0));   /* small circuits: one tile, nothing to choose */
    if (!search_threshold && !search_lanes && !search_climb) {
        p = new TiledPlan();
        int rc = tiled_schedule(n, prec, g, nloc, rank, opt, start, cops, gphase, p);
        if (rc) { delete p; return rc; }
    } else {
        /* Planner knobs the caller did not pin: plan a few candidates, each on its own host thread (~5-10 ms), and keep
         * the cheapest schedule.
         *   sharded run -- two knobs trade passes and rounds against qubit exchanges, and the best setting depends on the
         *     circuit and on the number of ranks: the exchange threshold (how few runnable gates make the scheduler
         *     exchange qubits), reserved[0], and the lane relocation policy (every pass re-picks the qubits on the lane
         *     positions: fewer passes, but the schedule runs dry behind a global qubit sooner; or only on conflict),
         *     reserved[6]: 4 x 2 candidates;
         *   single GPU -- the tile hill climbing ends in a local optimum that depends on the order of its swaps: four
         *     orders (tiled_schedule, climb_variant), a pass more or less on the 30-34 q workloads.
         * The cost model is a fit of the 30 q lines of round 2 ((passes, rounds) -> ms: (22, 130)
         * 126.2, (54, 134) 177.7, (18, 117) 117.8, (20, 114) 119.1), in ms at 2^30 local amplitudes: 1.5 per pass + 0.25 per
         * round, and the extra time of an exchange pass (NVLink-bound scatter of (1 - 2^-g) of the shard at ~700 GB/s, then
         * the barrier; from the 34 q scaling lines): 1.25 with one peer, 2.5 with three, 8 with seven.  Every rank plans the
         * same circuit and picks the same schedule (the rank only enters the descriptors, not the structure); ties go to
         * the earlier candidate. */
        static const int thresholds[] = {10, 8, 14, 12};
        static const int lane_policies[] = {0, 4};
        static const int climb_variants[] = {0, 3, 2, 4};
        struct Cand { qsb_options_t o; int climb; TiledPlan *plan; int rc; char err[512]; };
        std::vector<Cand> cand;
        qsb_options_t base; if (opt) base = *opt; else memset(&base, 0, sizeof base);   /* all knobs at their defaults */
        for (int lp : lane_policies) {
            if (!search_lanes && lp != lane_policies[0]) continue;
            for (int th : thresholds) {
                if (!search_threshold && th != thresholds[0]) continue;
                for (int cv : climb_variants) {
                    if (!search_climb && cv != climb_variants[0]) continue;
                    Cand c; c.o = base; c.climb = cv; c.plan = nullptr; c.rc = QSB_OK; c.err[0] = 0;
                    if (search_threshold) c.o.reserved[0] = th;
                    if (search_lanes) c.o.reserved[6] = lp;
                    cand.push_back(c);
                }
            }
        }
        const bool tracing = getenv("QSB_PLAN_TRACE") != nullptr;   /* candidates stay silent; the winner is planned once more, aloud */
        auto run = [&](Cand &c) {
            tiled_plan_trace_suppress(true);
            c.plan = new TiledPlan();
            c.rc = tiled_schedule(n, prec, g, nloc, rank, &c.o, start, cops, gphase, c.plan, c.climb);
            if (c.rc) { snprintf(c.err, sizeof c.err, "%s", qsb_last_error()); delete c.plan; c.plan = nullptr; }   /* the message is thread-local */
            tiled_plan_trace_suppress(false);
        };
        /* small circuits are planned faster than a thread starts */
        const bool threads = cops.size() >= 64;
        std::vector<std::thread> workers;
        for (size_t i = 1; i < cand.size(); i++) {
            bool started = false;
            if (threads) { try { workers.emplace_back(run, std::ref(cand[i])); started = true; } catch (...) {} }   /* no thread to be had: plan it here */
            if (!started) run(cand[i]);
        }
        run(cand[0]);
        for (std::thread &w : workers) w.join();
        const double xcost = g == 1 ? 1.25 : g == 2 ? 2.5 : 8.0;
        auto cost_of = [&](const TiledPlan *q) {
            double c = 0;
            for (auto &hp : q->passes) {
                if (hp.is_swap) { c += 1.5 + xcost; continue; }     /* the all-to-all costs about a pass on top */
                c += 1.5 + 0.25 * (double)hp.rounds.size() + (hp.fused_swap ? xcost : 0.0);
            }
            return c;
        };
        double best = 0;
        const Cand *winner = nullptr;
        for (Cand &c : cand) {
            if (!c.plan) continue;
            const double cc = cost_of(c.plan);
            if (!p || cc < best - 1e-9) { delete p; p = c.plan; best = cc; winner = &c; } else delete c.plan;
            c.plan = nullptr;
        }
        if (!p) { qsb_set_error("%s", cand[0].err); return cand[0].rc ? cand[0].rc : QSB_ERR_ARG; }
        if (tracing) {
            TiledPlan again;
            tiled_schedule(n, prec, g, nloc, rank, &winner->o, start, cops, gphase, &again, winner->climb);
            fprintf(stderr, "qsb-plan: kept candidate %d of %zu (exchange threshold %d, lane policy %d, climb order %d), modelled cost %.1f\n",
                    (int)(winner - cand.data()), cand.size(), winner->o.reserved[0], winner->o.reserved[6], winner->climb, best);
        }
    }
    *out = p;
    return QSB_OK;
}
