/*
 * blob_emulate.cpp -- TEST DOUBLE, never shipped: a host interpreter of the DEVICE ENCODING of a pass
 * (the kernel-parameter blob: GPass, GRound, segments, groups of slots, specials, thread-phase lists),
 * statement by statement what k_tile_pass in tiled_kernel.cu does with it.  emulate.cpp checks the
 * planner's logical tables; this file checks their lowering (PassBuilder::serialise), so that the
 * bytes the GPU reads are verified against the oracle on the CPU as well.
 * Arithmetic is done in double complex; f32 blobs carry float coefficients (compare at 1e-5).
 */
#include <complex>
#include <cstdio>
#include <cstring>
#include <vector>

#include "sim.h"
#include "tiled.h"

typedef std::complex<double> cd;

namespace {
struct Rd {
    const uint8_t *b; bool f32;
    double S(const uint8_t *p, int k) const { if (f32) { float x; memcpy(&x, p + 4 * k, 4); return x; } double x; memcpy(&x, p + 8 * k, 8); return x; }
    /* V entry k: (lo, hi) lanes for f32, one double for f64 */
    void V(const uint8_t *p, int k, double out[2]) const
    {
        if (f32) { float x[2]; memcpy(x, p + 8 * k, 8); out[0] = x[0]; out[1] = x[1]; }
        else { double x; memcpy(&x, p + 8 * k, 8); out[0] = out[1] = x; }
    }
};
}

/* returns the number of inconsistencies found (bad slot, out-of-range address, unknown code) */
/* how many special ops of each code the blob double has interpreted (tests ask whether a lowering was actually used) */
static unsigned long g_code_count[256];
extern "C" unsigned long qsb_hostcheck_blob_code_count(int code, int reset)
{
    const unsigned long v = g_code_count[code & 0xff];
    if (reset) for (auto &c : g_code_count) c = 0;
    return v;
}

/* the largest outer-condition table of the passes interpreted so far (QSB_MAX_COND = full: later gates fell back) */
static unsigned g_max_cond;
extern "C" unsigned qsb_hostcheck_blob_max_cond(int reset) { const unsigned v = g_max_cond; if (reset) g_max_cond = 0; return v; }

int blob_run_pass(const HostPass &hp, bool f32, int nloc, std::vector<cd> &st, cd *const *outs)
{
    int bad = 0;
    const int L = f32 ? 2 : 1;
    const uint64_t AMP = f32 ? 8 : 16;
    const uint8_t *B = hp.blob.data();
    Rd rd{B, f32};
    GPass P; memcpy(&P, B, sizeof P);
    if (P.n_cond > g_max_cond) g_max_cond = P.n_cond;
    const int SET16 = QSB_SET16(f32), G16 = QSB_GROUP16(f32);
    const uint64_t loc_bytes = ((uint64_t)1 << nloc) * AMP;
    std::vector<cd> regs((size_t)QSB_THREADS * QSB_NV * L), smem((size_t)QSB_SLOTS * L);
    std::vector<uint32_t> xm(QSB_THREADS);
    std::vector<cd> pend(QSB_THREADS);
    auto amp_at = [&](uint64_t byte_off, int lane) -> uint64_t { return f32 ? (byte_off / 8 + lane) : byte_off / 16; };
    for (uint64_t tile = 0; tile < P.n_tiles; tile++) {
        uint64_t t = tile, outer = 0;
        for (uint32_t r = 0; r < P.n_runs; r++) { int len = P.run_len[r]; outer |= (t & ((1ULL << len) - 1)) << P.run_start[r]; t >>= len; }
        const uint64_t src_outer = outer | P.src_fixed;
        uint32_t W = 0;
        for (uint32_t i = 0; i < P.n_cond; i++) if ((src_outer & P.cond[i]) == P.cond[i]) W |= 1u << i;
        /* angle entries test ONE 32-bit mask against the predicate word (tid, then the outer-condition bits W) */
        auto angle_word = [&](int tid) { return (uint32_t)tid | (W << QSB_TB); };
        /* sum of the entries of n_u 16-byte units that the word satisfies: f32 two {mask, ang32} per unit, f64 one {mask, -, ang64} */
        auto angle_sum = [&](const uint8_t *p, uint32_t n_u, uint32_t aw) {
            uint64_t acc = 0;
            for (uint32_t u = 0; u < n_u; u++, p += 16) {
                if (f32) { GTAngle32 e[2]; memcpy(e, p, 16); for (int k = 0; k < 2; k++) { if ((uint64_t)e[k].mask >> (QSB_TB + P.n_cond)) bad++; if ((aw & e[k].mask) == e[k].mask) acc += e[k].ang32; } }
                else { GTAngle64 e; memcpy(&e, p, 16); if ((uint64_t)e.mask >> (QSB_TB + P.n_cond)) bad++; if ((aw & e.mask) == e.mask) acc += e.ang64; }
            }
            return acc;
        };
        /* gather */
        for (int tid = 0; tid < QSB_THREADS; tid++) {
            uint64_t off = outer * AMP;
            for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) off += P.ld_thr[j];
            for (int v = 0; v < QSB_NV; v++) {
                const uint64_t o = off + P.ld_vec[v];
                if (o + 16 > loc_bytes) { bad++; continue; }
                for (int l = 0; l < L; l++) regs[((size_t)tid * QSB_NV + v) * L + l] = st[amp_at(o, l)];
            }
            xm[tid] = 0;
        }
        const uint8_t *rp = B + (size_t)P.rounds_off16 * 16;
        for (uint32_t r = 0; r < P.n_rounds; r++) {
            GRound RD; memcpy(&RD, rp, sizeof RD);
            if ((size_t)RD.seg_off16 * 16 != (size_t)(rp - B) + sizeof(GRound)) bad++;   /* the kernel finds the segment table right behind the header */
            rp = B + (size_t)RD.next16 * 16;
            std::vector<uint32_t> sb(QSB_THREADS, 0);
            for (int tid = 0; tid < QSB_THREADS; tid++) for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) sb[tid] ^= RD.thr_x[j];
            if (r > 0) for (int tid = 0; tid < QSB_THREADS; tid++) for (int v = 0; v < QSB_NV; v++) {
                uint32_t vx = 0; for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) vx ^= RD.vld_b[b];
                const uint32_t a = (sb[tid] & 0xffffu) ^ vx;
                if (a % 16 || a / 16 >= QSB_SLOTS) { bad++; continue; }
                for (int l = 0; l < L; l++) regs[((size_t)tid * QSB_NV + v) * L + l] = smem[(size_t)(a / 16) * L + l];
            }
            for (int tid = 0; tid < QSB_THREADS; tid++) pend[tid] = cd(1.0, 0.0);
            /* unit-form / diag helpers on thread tid */
            auto unit = [&](int tid, int vb, bool imag, double p, double q, double k) {
                cd *R = &regs[(size_t)tid * QSB_NV * L];
                for (int v = 0; v < QSB_NV; v++) if (!((v >> vb) & 1)) for (int l = 0; l < L; l++) {
                    cd &x0 = R[v * L + l], &x1 = R[(v | (1 << vb)) * L + l];
                    if (!imag) { x0 += p * x1; x1 = k * x1 + q * x0; } else { x0 += cd(0, p) * x1; x1 = k * x1 + cd(0, q) * x0; }
                }
            };
            const GSegment *seg = (const GSegment *)(B + (size_t)RD.seg_off16 * 16);
            for (uint32_t sg = 0; sg < RD.n_seg; sg++) {
                GSegment SG; memcpy(&SG, &seg[sg], sizeof SG);
                const uint8_t *op = B + (size_t)SG.special_off16 * 16;
                for (uint32_t i = 0; i < SG.n_special; i++) {
                    uint32_t h[4]; memcpy(h, op, 16);
                    const uint8_t *c = op + 16;
                    op += (size_t)(h[0] >> 16) * 16;
                    const uint32_t code = h[0] & 0xff, vmask = GOP_VMASK(h[0]); const bool two = (h[0] >> 8) & 1;
                    const uint64_t om = ((uint64_t)h[3] << 32) | h[2];
                    g_code_count[code]++;
                    if (code >= G_DIAGA && code <= G_DIAGA + QSB_NVB) {   /* merged controlled phases: per-thread fixed-point angle sum, one phase */
                        const int vb = code - G_DIAGA;                     /* == QSB_NVB: run on the pack qubit (high lane of every vector) */
                        if (vb == QSB_NVB && !f32) bad++;
                        uint32_t n_u; memcpy(&n_u, c, 4);                 /* 16-byte units of angle entries */
                        if (two || h[1] || om || (size_t)(h[0] >> 16) != 2 + (size_t)n_u) bad++;
                        for (int tid = 0; tid < QSB_THREADS; tid++) {
                            const uint64_t acc = angle_sum(c + 16, n_u, angle_word(tid));
                            const double half_turns = f32 ? (double)(int32_t)(uint32_t)acc / 2147483648.0 : (double)(int64_t)acc / 9223372036854775808.0;
                            const double PI_ = 3.14159265358979323846;
                            const cd ph(cos(PI_ * half_turns), sin(PI_ * half_turns));
                            cd *R = &regs[(size_t)tid * QSB_NV * L];
                            if (vb == QSB_NVB) { for (int v = 0; v < QSB_NV; v++) R[v * L + (L - 1)] *= ph; }
                            else for (int v = 0; v < QSB_NV; v++) if ((v >> vb) & 1) for (int l = 0; l < L; l++) R[v * L + l] *= ph;
                        }
                        continue;
                    }
                    for (int tid = 0; tid < QSB_THREADS; tid++) {
                        const bool pred = ((src_outer & om) == om) && (((uint32_t)tid & h[1]) == h[1]);
                        if (!two && !pred) continue;
                        const bool s1 = two && pred;
                        cd *R = &regs[(size_t)tid * QSB_NV * L];
                        if (code >= G_FULL_G && code < G_FULL_G + QSB_NVB) {
                            const int vb = code - G_FULL_G; const uint8_t *cs = c + (s1 ? 64 : 0);
                            double m[8][2]; for (int k = 0; k < 8; k++) rd.V(cs, k, m[k]);
                            for (int v = 0; v < QSB_NV; v++) if (!((v >> vb) & 1)) for (int l = 0; l < L; l++) {
                                const int ll = f32 ? l : 1;
                                cd m00(m[0][ll], m[1][ll]), m01(m[2][ll], m[3][ll]), m10(m[4][ll], m[5][ll]), m11(m[6][ll], m[7][ll]);
                                cd x0 = R[v * L + l], x1 = R[(v | (1 << vb)) * L + l];
                                R[v * L + l] = m00 * x0 + m01 * x1; R[(v | (1 << vb)) * L + l] = m10 * x0 + m11 * x1;
                            }
                        } else if ((code >= G_DIAG_V && code < G_DIAG_V + QSB_NVB) || code == G_DIAG_ALL || code == G_DIAG_GEN) {
                            double pr[2], pi[2]; const uint8_t *cs = c + (s1 ? 16 : 0);
                            rd.V(cs, 0, pr); rd.V(cs, 1, pi);
                            for (int v = 0; v < QSB_NV; v++) {
                                if (code != G_DIAG_ALL && code != G_DIAG_GEN && !((v >> (code - G_DIAG_V)) & 1)) continue;
                                if (code == G_DIAG_GEN && (v & vmask) != vmask) continue;
                                for (int l = 0; l < L; l++) { const int ll = f32 ? l : 1; R[v * L + l] *= cd(pr[ll], pi[ll]); }
                            }
                        } else if (code == G_MATP_R || code == G_MATP_G) {
                            if (!f32) { bad++; continue; }
                            double A[2][2] = {{0, 0}, {0, 0}}, Bc[2][2] = {{0, 0}, {0, 0}};   /* [re/im][lane] */
                            if (code == G_MATP_R) { const uint8_t *cs = c + (s1 ? 16 : 0); rd.V(cs, 0, A[0]); rd.V(cs, 1, Bc[0]); }
                            else { const uint8_t *cs = c + (s1 ? 32 : 0); rd.V(cs, 0, A[0]); rd.V(cs, 1, A[1]); rd.V(cs, 2, Bc[0]); rd.V(cs, 3, Bc[1]); }
                            for (int v = 0; v < QSB_NV; v++) {
                                cd x0 = R[v * L], x1 = R[v * L + 1];
                                R[v * L] = cd(A[0][0], A[1][0]) * x0 + cd(Bc[0][0], Bc[1][0]) * x1;
                                R[v * L + 1] = cd(A[0][1], A[1][1]) * x1 + cd(Bc[0][1], Bc[1][1]) * x0;
                            }
                        } else bad++;
                    }
                }
                const uint8_t *gp = B + (size_t)SG.group_off16 * 16;
                for (uint32_t g = 0; g < SG.n_groups; g++, gp += (size_t)G16 * 16) {
                    uint32_t pm[QSB_NVB]; for (int j = 0; j < QSB_NVB; j++) memcpy(&pm[j], gp + QSB_GROUP_MASK_OFF(j), 4);
                    for (int vb = 0; vb < QSB_NVB; vb++) {
                        const uint32_t form = gp[vb];
                        if (!form) continue;
                        const uint8_t *c0 = gp + 32 + (size_t)vb * 2 * SET16 * 16, *c1 = c0 + (size_t)SET16 * 16;
                        for (int tid = 0; tid < QSB_THREADS; tid++) {
                            const uint32_t tw = (uint32_t)tid | (W << QSB_TB);
                            const bool pred = (tw & pm[vb]) == pm[vb];
                            const uint8_t *cs = pred ? c1 : c0;
                            const double a0 = rd.S(cs, 0), a1 = rd.S(cs, 1), a2 = rd.S(cs, 2), a3 = rd.S(cs, 3);
                            if (form & S_UNIT_R) { unit(tid, vb, false, a0, a1, a2); pend[tid] *= a3; }
                            else if (form & S_UNIT_I) { unit(tid, vb, true, a0, a1, a2); pend[tid] *= a3; }
                            else if (form & S_UNIT_H) { if (a1 != 1.0 || a3 != 1.0) bad++; unit(tid, vb, false, a0, 1.0, a2); }
                            else if (form & S_DIAG) {
                                cd *R = &regs[(size_t)tid * QSB_NV * L];
                                for (int v = 0; v < QSB_NV; v++) if ((v >> vb) & 1) for (int l = 0; l < L; l++) R[v * L + l] *= cd(a0, a1);
                            } else if (!(form & S_XDEF)) bad++;
                            if ((form & S_XDEF) && pred) xm[tid] ^= 1u << vb;
                        }
                    }
                }
            }
            const uint8_t *e = B + (size_t)RD.tph_off16 * 16;
            for (uint32_t i = 0; i < RD.n_tph; i++, e += 32) {
                GTPhase T; memcpy(&T, e, sizeof T);
                if ((src_outer & T.omask) != T.omask) continue;
                const cd ph(rd.S(T.val, 0), rd.S(T.val, 1));
                for (int tid = 0; tid < QSB_THREADS; tid++) if (((uint32_t)tid & T.tmask) == T.tmask) pend[tid] *= ph;
            }
            /* angle entries follow the GTPhase entries: fixed-point turn fractions, summed per thread (wrapping) */
            if (RD.n_ang) {                                           /* n_ang: 16-byte units */
                std::vector<uint64_t> acc(QSB_THREADS, 0);
                for (int tid = 0; tid < QSB_THREADS; tid++) acc[tid] = angle_sum(e, RD.n_ang, angle_word(tid));
                for (int tid = 0; tid < QSB_THREADS; tid++) {
                    const double half_turns = f32 ? (double)(int32_t)(uint32_t)acc[tid] / 2147483648.0 : (double)(int64_t)acc[tid] / 9223372036854775808.0;
                    const double PI_ = 3.14159265358979323846;
                    pend[tid] *= cd(cos(PI_ * half_turns), sin(PI_ * half_turns));
                }
            }
            for (int tid = 0; tid < QSB_THREADS; tid++) {
                if (RD.flags & 1u) { for (int k = 0; k < QSB_NV * L; k++) regs[(size_t)tid * QSB_NV * L + k] *= pend[tid]; }
                else if (std::abs(pend[tid] - cd(1.0, 0.0)) > 0) bad++;     /* a scalar nobody applies */
            }
            if (r + 1 < P.n_rounds) {
                std::vector<int> written(QSB_SLOTS, 0);
                for (int tid = 0; tid < QSB_THREADS; tid++) {
                    uint32_t ss = sb[tid] >> 16;
                    for (int b = 0; b < QSB_NVB; b++) if ((xm[tid] >> b) & 1) ss ^= RD.vst_b[b];
                    xm[tid] = 0;
                    for (int v = 0; v < QSB_NV; v++) {
                        uint32_t vx = 0; for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) vx ^= RD.vst_b[b];
                        const uint32_t a = ss ^ vx;
                        if (a % 16 || a / 16 >= QSB_SLOTS) { bad++; continue; }
                        written[a / 16]++;
                        for (int l = 0; l < L; l++) smem[(size_t)(a / 16) * L + l] = regs[((size_t)tid * QSB_NV + v) * L + l];
                    }
                }
                for (int k = 0; k < QSB_SLOTS; k++) if (written[k] != 1) bad++;
            }
        }
        /* scatter */
        for (int tid = 0; tid < QSB_THREADS; tid++) {
            uint64_t off = (outer & ~P.xo_mask) * AMP + P.st_fixed, xoff = 0;
            for (uint32_t k = 0; k < P.n_xo; k++) off += ((outer >> P.xo_pos[k]) & 1ULL) << (QSB_RANK_SHIFT + P.xo_rank[k]);
            for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) off += P.st_thr[j];
            for (int b = 0; b < QSB_NVB; b++) if ((xm[tid] >> b) & 1) xoff ^= P.st_vec[1 << b];
            for (int v = 0; v < QSB_NV; v++) {
                const uint64_t tt = off + (P.st_vec[v] ^ xoff);
                const uint64_t rank = tt >> QSB_RANK_SHIFT, loc = tt & ((1ULL << QSB_RANK_SHIFT) - 1);
                if (loc + 16 > loc_bytes || (!hp.fused_swap && rank)) { bad++; continue; }
                for (int l = 0; l < L; l++) {
                    const cd val = regs[((size_t)tid * QSB_NV + v) * L + l];
                    if (hp.fused_swap) { if (!outs) { bad++; continue; } outs[rank][amp_at(loc, l)] = val; }
                    else st[amp_at(loc, l)] = val;
                }
            }
        }
    }
    return bad;
}
