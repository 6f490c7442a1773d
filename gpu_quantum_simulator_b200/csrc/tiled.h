/*
 * tiled.h -- the fused tile-pass schedule: descriptors shared by the host
 * planner (tiled_plan.cpp) and the sm_100a kernel (tiled_kernel.cu).
 *
 * A PASS is one sweep over the local state.  Every CTA owns one TILE: the
 * 2^T amplitudes obtained by fixing all physical index bits except T "tile
 * bits".  The tile bits are the low `a` physical bits (so a tile is made of
 * contiguous 2^a-amplitude segments and every global access is a full
 * 128-bit coalesced transaction) plus T-a freely chosen high bits -- this is
 * the qubit-remapping layer: high-stride qubits are gathered into the tile
 * instead of being swept with strided passes.
 *
 * Inside a pass the tile lives in REGISTERS: each of the 256 threads holds
 * 2^NVB vectors (NVB = 4).  For f32 a vector is a float2 holding the same
 * component (re or im) of the two amplitudes that differ in the PACK bit
 * (tile bit 0 == physical bit 0), so every butterfly is issued as packed
 * FFMA2/FMUL2 on two amplitudes at once; for f64 a vector is one double.
 * A ROUND applies all gates whose target is one of the NVB "vector bits" (or
 * the pack bit) of that round; between rounds the tile is exchanged through
 * shared memory with a GF(2)-linear slot map chosen by the planner so that
 * both the store of round k and the load of round k+1 are bank-conflict free.
 *
 * The whole description of a pass (header, per-round tables, op stream) is a
 * few KiB and travels as a __grid_constant__ kernel parameter: every table and
 * coefficient read is a constant-bank load, no global traffic besides the state.
 *
 * Replaces (reference, /root/reference/): the per-gate launches of
 * naive.cu:163-189, the 2x2/4x4 host fusion of preproces.cu:215-269 and
 * 4x4.cu:327-501, and the static relabel of 4x4_permute.cu:350-434.
 */
#pragma once
#include <stdint.h>
#include <vector>

#include "common.cuh"

#ifndef QSB_NVB
#define QSB_NVB 4              /* vector bits per round (build-time: 4 -> 256 threads x 16 vectors, 3 -> 512 x 8) */
#endif
#define QSB_NV (1 << QSB_NVB)  /* vectors per thread                         */
#define QSB_TB (12 - QSB_NVB)  /* thread bits                                */
#define QSB_THREADS (1 << QSB_TB)
#define QSB_T_F32 13           /* tile bits f32: pack + NVB + TB             */
#define QSB_T_F64 12           /* tile bits f64: NVB + TB                    */
#define QSB_MAX_RUNS 16
#define QSB_BLOB_SMALL 4000    /* pass descriptor sizes (kernel parameter)   */
#define QSB_BLOB_LARGE 32000

/* ---- op codes ---------------------------------------------------------- */
enum {
    OP_MAT_R = 1,    /* target = vector bit; all entries real                  */
    OP_MAT_I = 2,    /* target = vector bit; real diagonal, imaginary off-diag */
    OP_MAT_G = 3,    /* target = vector bit; general complex                   */
    OP_MATP_R = 4,   /* target = pack bit (f32 only); real                     */
    OP_MATP_G = 5,   /* target = pack bit (f32 only); general                  */
    OP_MAT_U = 6,    /* real, unit form a*[[1,p],[q,r]]: x0 += p*x1; x1 = k*x1 + q*x0 (k = r - q*p);
                        the scale a goes to the per-thread pending scalar.  Pure in-place chain. */
    OP_MAT_UI = 7,   /* rx form a*[[1,ip],[iq,r]] in the same unit form (k = r + q*p)       */
    /* X / CX that were not absorbed are issued as OP_MAT_R / OP_MATP_R with [[0,1],[1,0]]    */
    OP_DIAG_V = 9,   /* phase on the vectors whose vector bit `vb` is set      */
    OP_DIAG_ALL = 10,/* phase on all vectors (lane-dependent, e.g. rz on the pack qubit) */
    OP_DIAG_GEN = 11,/* phase on vectors with (v & vmask) == vmask, vmask in kind bits 20..23 */
    OP_TPHASE = 12,  /* thread-level phase: folded into a per-thread scalar    */
};

/* kind = opcode | vb << 8 | mux << 16 | vmask << 20
 *   vb    : target vector bit (OP_MAT_*, OP_X*, OP_DIAG_V)
 *   mux   : thread-level multiplexer: threads whose tmask test fails use
 *           coefficient set 0 instead of skipping (set 1 follows set 0)       */
#define OPK(op, vb, mux, vmask) ((uint32_t)(op) | ((uint32_t)(vb) << 8) | ((uint32_t)(mux) << 16) | ((uint32_t)(vmask) << 20))

/* Op stream element: 16-byte header followed by the coefficient payload.
 * Payload entries are 8 bytes in both precisions: f32 = (lo lane, hi lane)
 * float2, f64 = one double.
 *   OP_MAT_R : m00 m01 m10 m11                       (4 entries per set)
 *   OP_MAT_I : a -b b -c c d  for [[a, ib],[ic, d]]  (6 entries per set)
 *   OP_MAT_G : m00r m00i m01r m01i m10r m10i m11r m11i (8 entries per set)
 *   OP_MAT_U : p q k a   (a: scale, lo lane = value)    OP_MAT_UI: p -p q -q k a
 *   OP_MATP_R: A B          out = A*x + B*swap(x), A=(m00,m11) B=(m01,m10)
 *   OP_MATP_G: Ar Ai Br Bi
 *   OP_DIAG_*: pr pi
 *   OP_TPHASE: 16 bytes holding (pr, pi) as two scalars of the state precision */
struct OpHdr {
    uint32_t kind;
    uint32_t size16; /* header + payload, in 16-byte units */
    uint64_t tmask;  /* condition on the thread's physical index bits */
};

struct BitEntry {      /* one thread bit or vector bit of a round */
    uint64_t gidx;     /* physical SOURCE index bit(s) it stands for (tmask tests; round 0: load address) */
    uint16_t ld, st;   /* smem slot XOR constants, load side / store side */
    uint32_t pad;
};

struct DevRound {
    uint32_t n_ops;
    uint32_t op_off16;          /* op stream start, 16-byte units from the blob start */
    uint32_t flags;             /* bit0: round has OP_TPHASE ops              */
    uint32_t pad;
    BitEntry thr[QSB_TB];
    BitEntry vec[QSB_NVB];
};

struct DevPass {
    uint32_t n_rounds;
    uint32_t n_runs;            /* runs of consecutive outer bits             */
    uint8_t run_start[QSB_MAX_RUNS];
    uint8_t run_len[QSB_MAX_RUNS];
    uint64_t src_fixed;         /* constant physical index bits of every tile (rank bits, swap-pass selectors) */
    uint64_t dst_fixed;
    uint64_t dst_thr[QSB_TB];   /* last round: thread / vector bit -> DESTINATION index bits */
    uint64_t dst_vec[QSB_NVB];
    uint32_t nloc;              /* local index bits: index >> nloc selects the source rank   */
    uint32_t out_of_place;      /* 1: write to the second buffer              */
    uint64_t n_tiles;
    uint32_t rounds_off16;      /* DevRound array, 16-byte units from the blob start */
    uint32_t blob_bytes;
};

template <int BYTES> struct PassBlob { uint4 q[BYTES / 16]; };

/* ---- host-side plan ----------------------------------------------------- */
struct HostOp {              /* precision-independent; lane-expanded fp64 coefficients */
    uint32_t kind;
    uint64_t tmask;
    int n_coef;              /* payload entries per set */
    double c[2][8][2];       /* [set][entry][lane] */
    double tph[2];           /* OP_TPHASE */
};

struct HostPass {
    DevPass hdr;
    std::vector<DevRound> rounds;
    std::vector<HostOp> ops;
    std::vector<uint8_t> blob;           /* serialised kernel parameter        */
    /* bookkeeping for tests / emulation */
    int T = 0;                           /* tile bits                          */
    int8_t tile_src[16];                 /* tile bit j -> source physical bit  */
    int8_t tile_dst[16];                 /* tile bit j -> destination physical bit */
    std::vector<std::vector<int8_t>> round_thr; /* per round: thread bit j -> tile bit */
    std::vector<std::vector<int8_t>> round_vec; /* per round: vector bit j -> tile bit */
    std::vector<uint32_t> round_op_begin, round_op_count; /* indices into ops */
    bool is_swap = false;
    int n_source_ops = 0;
};

struct TiledPlan {
    int n = 0, prec = QSB_F32, g = 0, nloc = 0, rank = 0;
    std::vector<HostPass> passes;
    BitPerm start_perm{}, end_perm{};
    double last_exchange_ms = 0.0;
};

int tiled_min_local_bits(int prec, const qsb_options_t *opt);
int tiled_plan_build(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                     const std::vector<COp> &cops, const double gphase[2], bool with_device,
                     TiledPlan **out, qsb_run_stats_t *stats);
void tiled_plan_free(TiledPlan *p);
struct qsb_sim;
int tiled_execute(qsb_sim *s, TiledPlan *p);
double tiled_last_exchange_ms(const TiledPlan *p);
void tiled_comm_destroy(qsb_sim *s);

/* host-only planner entry (no CUDA): used by tiled_plan_build and by the test emulator */
int tiled_schedule(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                   const std::vector<COp> &cops, const double gphase[2], TiledPlan *plan);
