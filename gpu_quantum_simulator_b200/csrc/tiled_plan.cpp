/*
 * tiled_plan.cpp -- host-side gate fusion and scheduling (no CUDA in this file).
 *
 * Turns the canonical op list into PASSES (one HBM sweep each) and ROUNDS
 * (register-resident butterfly groups inside a pass) and emits the uniform
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
 * factor.  Diagonal gates and controls never need residency: they become
 * per-thread predicates / phases on physical index bits.
 */
#include <math.h>
#include <string.h>

#include <algorithm>

#include "sim.h"
#include "tiled.h"

/* ---------------------------------------------------------------- canonical */
int qsb_canonicalise(const qsb_gate_t *gates, size_t n, int num_qubits, std::vector<COp> &out, double gphase[2])
{
    gphase[0] = 1.0; gphase[1] = 0.0;
    out.clear(); out.reserve(n + 1);
    for (size_t k = 0; k < n; k++) {
        const qsb_gate_t &g = gates[k];
        if (g.target < 0 || g.target >= num_qubits) { qsb_set_error("gate %zu: target %d outside the %d-qubit register", k, g.target, num_qubits); return QSB_ERR_ARG; }
        if (num_qubits < 64 && (g.controls >> num_qubits)) { qsb_set_error("gate %zu: control outside the register", k); return QSB_ERR_ARG; }
        if (g.controls & (1ULL << g.target)) { qsb_set_error("gate %zu: target %d is also a control", k, g.target); return QSB_ERR_ARG; }
        const double *m = g.m;
        const bool offdiag_zero = m[2] == 0 && m[3] == 0 && m[4] == 0 && m[5] == 0;
        const bool diag_zero = m[0] == 0 && m[1] == 0 && m[6] == 0 && m[7] == 0;
        COp c; memset(&c, 0, sizeof c);
        if (offdiag_zero) {
            /* diag(d0, d1) = d0 * diag(1, d1/d0) */
            double d0r = m[0], d0i = m[1], d1r = m[6], d1i = m[7];
            if (!(d0r == 1.0 && d0i == 0.0)) {
                double den = d0r * d0r + d0i * d0i;
                if (den == 0) { qsb_set_error("gate %zu: singular diagonal matrix", k); return QSB_ERR_ARG; }
                if (g.controls == 0) { /* a global scalar: fold */
                    double r = gphase[0] * d0r - gphase[1] * d0i, i = gphase[0] * d0i + gphase[1] * d0r;
                    gphase[0] = r; gphase[1] = i;
                } else {
                    c.kind = C_PHASE; c.ctrl = g.controls; c.target = -1; c.m[0] = d0r; c.m[1] = d0i;
                    out.push_back(c);
                }
                double qr = (d1r * d0r + d1i * d0i) / den, qi = (d1i * d0r - d1r * d0i) / den;
                d1r = qr; d1i = qi;
            }
            if (!(d1r == 1.0 && d1i == 0.0)) {
                c.kind = C_PHASE; c.ctrl = g.controls | (1ULL << g.target); c.target = -1; c.m[0] = d1r; c.m[1] = d1i;
                out.push_back(c);
            }
        } else if (diag_zero && m[2] == 1 && m[3] == 0 && m[4] == 1 && m[5] == 0) {
            c.kind = C_X; c.target = g.target; c.ctrl = g.controls;
            out.push_back(c);
        } else {
            c.kind = C_MAT; c.target = g.target; c.ctrl = g.controls; memcpy(c.m, m, sizeof c.m);
            out.push_back(c);
        }
    }
    return QSB_OK;
}

/* ------------------------------------------------------------------ helpers */
int tiled_min_local_bits(int prec, const qsb_options_t *) { return prec == QSB_F64 ? QSB_T_F64 : QSB_T_F32; }

namespace {

struct Machine {
    int n, prec, g, nloc, rank, T, a, nb;
    bool f32;
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

int classify(const double *m) /* 1 real, 2 real-diag/imag-offdiag, 3 general */
{
    if (m[1] == 0 && m[3] == 0 && m[5] == 0 && m[7] == 0) return 1;
    if (m[1] == 0 && m[7] == 0 && m[2] == 0 && m[4] == 0) return 2;
    return 3;
}

void set_coef(HostOp &op, int set, int coef, double lo, double hi) { op.c[set * 16 + coef * 2] = lo; op.c[set * 16 + coef * 2 + 1] = hi; }

} // namespace

/* ------------------------------------------------------------ pass building */
struct PassBuilder {
    const Machine &M;
    const BitPerm &perm;        /* logical -> physical at the start of the pass */
    HostPass hp;
    int tile_of_qubit[64];      /* logical qubit -> tile bit or -1              */
    std::vector<int> tile_qubit; /* tile bit -> logical qubit or -1 (padding)   */

    PassBuilder(const Machine &m, const BitPerm &p) : M(m), perm(p) {}

    /* choose the tile from the set of resident logical qubits */
    void set_tile(uint64_t resident)
    {
        /* physical positions in the tile */
        uint64_t posmask = 0;
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
            hp.tile_src[j] = (int8_t)p; hp.tile_dst[j] = (int8_t)p;
            tile_qubit[j] = inv[p];
            if (inv[p] >= 0) tile_of_qubit[inv[p]] = j;
            j++;
        }
        /* outer runs */
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
    }

    /* tile bits that must be thread (lane) bits in the first / last round */
    uint32_t lane_forbidden() const
    {
        uint32_t f = 0;
        const int lo = M.f32 ? 1 : 0;
        for (int j = lo; j < M.a && j < lo + 5; j++) f |= 1u << j; /* tile bit j == physical bit j for j < a */
        return f;
    }

    /* Build rounds for the ordered op list `ops` (all targets resident). */
    int build_rounds(const std::vector<COp> &ops)
    {
        const int P = M.f32 ? 0 : -1;             /* pack tile bit */
        const uint32_t F = lane_forbidden();
        const size_t n = ops.size();
        std::vector<char> done(n, 0);
        size_t left = n, first_open = 0;
        std::vector<uint32_t> roundR;             /* vector-bit set (tile-bit mask) per round */
        std::vector<std::vector<int>> round_ops;
        while (left || roundR.empty()) {
            const bool is_first = roundR.empty();
            uint32_t R = 0; int nR = 0;
            Blocker B; B.clear();
            std::vector<int> mine;
            for (size_t i = first_open; i < n; i++) {
                if (done[i]) continue;
                const COp &o = ops[i];
                bool can = B.ok(o);
                if (can && o.target >= 0) {
                    int tb = tile_of_qubit[o.target];
                    if (tb == P) { /* pack variants */ }
                    else if (is_first && ((F >> tb) & 1)) can = false;
                    else if ((R >> tb) & 1) {}
                    else if (nR < QSB_NVB) { R |= 1u << tb; nR++; }
                    else can = false;
                }
                if (can) { mine.push_back((int)i); done[i] = 1; left--; }
                else { B.block(o); if (B.full >= M.n) break; }
            }
            while (first_open < n && done[first_open]) first_open++;
            roundR.push_back(R); round_ops.push_back(mine);
            if (!left) break;
        }
        /* last round must keep the low destination bits on lanes */
        if (roundR.back() & F) { roundR.push_back(0); round_ops.push_back({}); }

        const int nrounds = (int)roundR.size();
        hp.rounds.assign(nrounds, DevRound());
        hp.round_thr.assign(nrounds, {}); hp.round_vec.assign(nrounds, {});
        for (int r = 0; r < nrounds; r++) {
            uint32_t R = roundR[r];
            const bool edge = (r == 0 || r == nrounds - 1);
            /* pad R with the highest free tile bits */
            for (int tb = M.T - 1; tb >= 0 && popc(R) < QSB_NVB; tb--) {
                if (tb == P || ((R >> tb) & 1)) continue;
                if (edge && ((F >> tb) & 1)) continue;
                R |= 1u << tb;
            }
            std::vector<int8_t> vec, thr;
            for (int tb = 0; tb < M.T; tb++) if ((R >> tb) & 1) vec.push_back((int8_t)tb);
            /* thread bits: ascending; on edge rounds this puts the low physical bits on the lanes */
            for (int tb = 0; tb < M.T; tb++) if (tb != P && !((R >> tb) & 1)) thr.push_back((int8_t)tb);
            hp.round_vec[r] = vec; hp.round_thr[r] = thr;
            DevRound &D = hp.rounds[r];
            memset(&D, 0, sizeof D);
            for (int j = 0; j < QSB_TB; j++) D.thr_gidx[j] = 1ULL << hp.tile_src[thr[j]];
            for (int j = 0; j < QSB_NVB; j++) D.vec_gidx[j] = 1ULL << hp.tile_src[vec[j]];
        }
        /* destination tables (last round) */
        for (int j = 0; j < QSB_TB; j++) hp.hdr.dst_thr[j] = 1ULL << hp.tile_dst[hp.round_thr[nrounds - 1][j]];
        for (int j = 0; j < QSB_NVB; j++) hp.hdr.dst_vec[j] = 1ULL << hp.tile_dst[hp.round_vec[nrounds - 1][j]];
        hp.hdr.n_rounds = nrounds;

        /* shared-memory slot maps between consecutive rounds */
        for (int r = 0; r + 1 < nrounds; r++) slot_map(r);

        /* ops */
        hp.ops.clear();
        for (int r = 0; r < nrounds; r++) {
            DevRound &D = hp.rounds[r];
            D.op_begin = (uint32_t)hp.ops.size();
            for (int i : round_ops[r]) emit(ops[i], r);
            D.n_ops = (uint32_t)hp.ops.size() - D.op_begin;
            for (uint32_t k = D.op_begin; k < D.op_begin + D.n_ops; k++)
                if ((hp.ops[k].kind & 0xff) == OP_TPHASE) D.flags |= 1u;
        }
        hp.n_source_ops = (int)n;
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
        /* upper slot bits: every tile bit that is not one of the reader's bank lanes */
        bool isD[16] = {false};
        for (int i = 0; i < nb; i++) isD[rt[i]] = true;
        int up = nb;
        for (int tb = 0; tb < M.T; tb++) {
            if (tb == P || isD[tb]) continue;
            col[tb] |= (uint16_t)(1u << up); up++;
        }
        DevRound &W = hp.rounds[r], &Rd = hp.rounds[r + 1];
        for (int j = 0; j < QSB_TB; j++) { W.st_thr[j] = col[wt[j]]; Rd.ld_thr[j] = col[rt[j]]; }
        for (int j = 0; j < QSB_NVB; j++) { W.st_vec[j] = col[hp.round_vec[r][j]]; Rd.ld_vec[j] = col[hp.round_vec[r + 1][j]]; }
    }

    void emit(const COp &o, int r)
    {
        const int P = M.f32 ? 0 : -1;
        HostOp h; memset(&h, 0, sizeof h);
        /* split the condition mask */
        bool pack_ctrl = false;
        for (uint64_t m = o.ctrl; m; m &= m - 1) {
            int q = __builtin_ctzll(m);
            int tb = tile_of_qubit[q];
            int vi = -1;
            if (tb >= 0) for (int j = 0; j < QSB_NVB; j++) if (hp.round_vec[r][j] == tb) vi = j;
            if (tb >= 0 && tb == P) pack_ctrl = true;
            else if (vi >= 0) h.vmask |= 1u << vi;
            else h.tmask |= 1ULL << perm.pos[q];
        }
        const double lo_id = pack_ctrl ? 1.0 : 0.0; /* helper: identity entries for the lo lane */
        if (o.kind == C_PHASE) {
            if (!pack_ctrl && h.vmask == 0) {
                h.kind = OPK(OP_TPHASE, 0, 3, 0);
                set_coef(h, 1, 0, o.m[0], o.m[0]); set_coef(h, 1, 1, o.m[1], o.m[1]);
            } else {
                h.kind = OPK(OP_DIAG, 0, 3, 0);
                set_coef(h, 1, 0, pack_ctrl ? 1.0 : o.m[0], o.m[0]);
                set_coef(h, 1, 1, pack_ctrl ? 0.0 : o.m[1], o.m[1]);
            }
            hp.ops.push_back(h);
            return;
        }
        const int tb = tile_of_qubit[o.target];
        int vb = -1;
        for (int j = 0; j < QSB_NVB; j++) if (hp.round_vec[r][j] == tb) vb = j;
        if (o.kind == C_X) {
            if (tb == P) h.kind = OPK(OP_XP, 0, 3, 0);
            else h.kind = OPK(OP_X, vb, pack_ctrl ? 2 : 3, 0);
            hp.ops.push_back(h);
            return;
        }
        /* C_MAT */
        double m[8];
        double sc = 0; for (int k = 0; k < 8; k++) sc = std::max(sc, fabs(o.m[k]));
        for (int k = 0; k < 8; k++) m[k] = snap(o.m[k], sc);
        const int form = classify(m);
        if (tb == P) {
            /* out = A*x + B*swap(x): A = (m00, m11), B = (m01, m10) */
            h.kind = OPK(form == 1 ? OP_MATP_R : OP_MATP_G, 0, 3, 0);
            set_coef(h, 1, 0, m[0], m[6]); set_coef(h, 1, 1, m[1], m[7]);
            set_coef(h, 1, 2, m[2], m[4]); set_coef(h, 1, 3, m[3], m[5]);
        } else if (form == 1) {
            h.kind = OPK(OP_MAT_R, vb, 3, 0);
            set_coef(h, 1, 0, pack_ctrl ? 1.0 : m[0], m[0]); set_coef(h, 1, 2, pack_ctrl ? 0.0 : m[2], m[2]);
            set_coef(h, 1, 4, pack_ctrl ? 0.0 : m[4], m[4]); set_coef(h, 1, 6, pack_ctrl ? 1.0 : m[6], m[6]);
        } else if (form == 2) {
            h.kind = OPK(OP_MAT_I, vb, 3, 0);
            const double a = m[0], b = m[3], c = m[5], d = m[6];
            set_coef(h, 1, 0, pack_ctrl ? 1.0 : a, a);
            set_coef(h, 1, 1, pack_ctrl ? 0.0 : -b, -b); set_coef(h, 1, 2, pack_ctrl ? 0.0 : b, b);
            set_coef(h, 1, 3, pack_ctrl ? 0.0 : -c, -c); set_coef(h, 1, 4, pack_ctrl ? 0.0 : c, c);
            set_coef(h, 1, 6, pack_ctrl ? 1.0 : d, d);
        } else {
            h.kind = OPK(OP_MAT_G, vb, 3, 0);
            for (int k = 0; k < 8; k++) set_coef(h, 1, k, pack_ctrl ? ((k == 0 || k == 6) ? 1.0 : 0.0) : m[k], m[k]);
        }
        (void)lo_id;
        hp.ops.push_back(h);
    }
};

/* ----------------------------------------------------------------- scheduler */
int tiled_schedule(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                   const std::vector<COp> &cops_in, const double gphase[2], TiledPlan *plan)
{
    Machine M;
    M.n = n; M.prec = prec; M.g = g; M.nloc = nloc; M.rank = rank;
    M.f32 = (prec == QSB_F32);
    M.T = M.f32 ? QSB_T_F32 : QSB_T_F64;
    M.nb = M.f32 ? 4 : 3;
    M.a = opt && opt->low_bits > 0 ? opt->low_bits : (M.f32 ? 6 : 5);
    if (M.a < (M.f32 ? 3 : 2) || M.a > (M.f32 ? 6 : 5)) { qsb_set_error("low_bits %d unsupported for this precision", M.a); return QSB_ERR_ARG; }
    if (nloc < M.T) { qsb_set_error("internal: local register smaller than a tile"); return QSB_ERR_ARG; }
    if (g > 0) { qsb_set_error("multi-GPU tiled schedule not available in this build"); return QSB_ERR_ARG; }

    plan->n = n; plan->prec = prec; plan->g = g; plan->nloc = nloc; plan->rank = rank;
    plan->start_perm = start;
    plan->passes.clear();

    std::vector<COp> cops = cops_in;
    if (!(gphase[0] == 1.0 && gphase[1] == 0.0)) {
        COp c; memset(&c, 0, sizeof c); c.kind = C_PHASE; c.target = -1; c.ctrl = 0; c.m[0] = gphase[0]; c.m[1] = gphase[1];
        cops.push_back(c);
    }
    BitPerm perm = start;
    const size_t N = cops.size();
    std::vector<char> done(N, 0);
    size_t left = N, first_open = 0;

    /* logical qubits sitting on the forced low positions */
    while (left) {
        uint64_t S = 0; int nS = 0;
        for (int q = 0; q < n; q++) if (perm.pos[q] < M.a) { S |= 1ULL << q; }
        nS = M.a; /* the low positions always occupy `a` tile slots, whether or not a logical qubit lives there */
        Blocker B; B.clear();
        std::vector<COp> mine;
        std::vector<size_t> mine_idx;
        for (size_t i = first_open; i < N; i++) {
            if (done[i]) continue;
            const COp &o = cops[i];
            bool can = B.ok(o);
            if (can && o.target >= 0) {
                if (perm.pos[o.target] >= nloc) can = false;
                else if ((S >> o.target) & 1) {}
                else if (nS < M.T) { S |= 1ULL << o.target; nS++; }
                else can = false;
            }
            if (can) { mine.push_back(o); mine_idx.push_back(i); }
            else { B.block(o); if (B.full >= n) break; }
        }
        if (mine.empty()) { qsb_set_error("scheduler made no progress (gate on a non-local qubit?)"); return QSB_ERR_ARG; }
        for (size_t i : mine_idx) { done[i] = 1; left--; }
        while (first_open < N && done[first_open]) first_open++;

        PassBuilder pb(M, perm);
        pb.set_tile(S);
        int rc = pb.build_rounds(mine);
        if (rc) return rc;
        plan->passes.push_back(std::move(pb.hp));
    }
    plan->end_perm = perm;
    return QSB_OK;
}
