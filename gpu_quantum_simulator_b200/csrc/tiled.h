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
 * Inside a pass the tile lives in REGISTERS: each of the 128 threads holds
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
 * coefficient read is a (mostly uniform-datapath) constant-bank load, no global
 * traffic besides the state.  The LOGICAL tables below (DevPass, DevRound, HostOp)
 * are what the planner reasons about and what the host test double interprets;
 * serialise() lowers them to the device encoding further down (GPass, GRound,
 * groups of slots, specials, thread-phase lists).
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
#define QSB_NVB 4              /* vector bits per round: 16 vectors per thread (8 per thread measured slower, round 1a; 32 per thread,
                                  -DQSB_NVB=5 -DQSB_TB=6, is the round-2 experiment of profiles/r2/nvb5_experiment.md) */
#endif
#define QSB_NV (1 << QSB_NVB)  /* vectors per thread                         */
/* Thread bits.  Build-time geometry switch, A/B measured on B200 (profiles/r1h_tile_geometry_ab.txt):
 *   QSB_TB = 8: 256 threads, two CTAs per SM, tiles of 2^13 (f32) / 2^12 (f64) amplitudes
 *   QSB_TB = 7: 128 threads, FOUR CTAs per SM, tiles of 2^12 / 2^11 -- 11 % (f32) / 6.5 % (f64) faster on the
 *               30 q depth-20 circuit although it needs 30 instead of 25 passes: a CTA runs gather -> rounds ->
 *               scatter one after the other, and four resident CTAs overlap HBM, shared-memory and FP32 phases
 *               better than two.  Default. */
#ifndef QSB_TB
#define QSB_TB 7
#endif
#define QSB_THREADS (1 << QSB_TB)
#define QSB_T_F64 (QSB_NVB + QSB_TB)   /* tile bits f64: NVB + TB (11)               */
#define QSB_T_F32 (QSB_T_F64 + 1)      /* tile bits f32: pack + NVB + TB (12)        */
#define QSB_SLOTS (1 << QSB_T_F64)     /* 16-byte shared-memory slots of a tile       */
#define QSB_SMEM_BYTES (QSB_SLOTS * 16)
#define QSB_SMEM_TOTAL QSB_SMEM_BYTES   /* dynamic shared memory of a CTA: the exchange buffer, nothing else */
#ifndef QSB_CTAS_PER_SM
#define QSB_CTAS_PER_SM (512 / QSB_THREADS)
#endif
#define QSB_MAX_RUNS 16
#define QSB_BLOB_SMALL 4000    /* pass descriptor sizes (kernel parameter)   */
#define QSB_BLOB_MEDIUM 12000
#ifndef QSB_BLOB_LARGE
#define QSB_BLOB_LARGE 31744   /* + two pointers + the peer table stay below the 32764-byte parameter limit */
#endif

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
    OP_XDEF = 13,    /* X on vector bit vb for the threads that pass: a swap of the two halves, which the
                        kernel defers into its next store address.  Last op on its qubit in the round. */
};

/* kind = opcode | vb << 8 | mux << 16 | vmask << 20
 *   vb    : target vector bit (OP_MAT_*, OP_X*, OP_DIAG_V)
 *   mux   : thread-level multiplexer: threads whose tmask test fails use
 *           coefficient set 0 instead of skipping (set 1 follows set 0)       */
#define OPK(op, vb, mux, vmask) ((uint32_t)(op) | ((uint32_t)(vb) << 8) | ((uint32_t)(mux) << 16) | ((uint32_t)(vmask) << 20))
#define OPK_VMASK(kind) (((kind) >> 20) & 0xffu)

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
    /* Fused exchange with victims OUTSIDE the tile (round 2): outer index bit xo_pos[k] of the source does not keep
     * its place -- it selects bit xo_rank[k] of the DESTINATION rank (the whole tile of a CTA then goes to one peer),
     * and the writer's own rank bit sits at that local position of the destination (part of dst_fixed). */
    uint32_t n_xo;
    uint8_t xo_pos[8], xo_rank[8];
};

template <int BYTES> struct PassBlob { uint4 q[BYTES / 16]; };

/* ---- device encoding -------------------------------------------------------
 * DevPass / DevRound / HostOp above are the LOGICAL tables (the host test double
 * interprets them).  PassBuilder::serialise() lowers them to the compact form the
 * kernel reads, with everything that does not depend on the thread pre-computed:
 * byte offsets per vector for the global gather / scatter, shared-memory XOR
 * constants per vector, and predicates as ONE 32-bit mask over the per-thread
 * word  tw = threadIdx.x | W << 8,  where bit i of W says whether the CTA's outer
 * index bits satisfy the i-th outer condition of the pass (GPass::cond).
 *
 * The ops of a round are laid out as SEGMENTS.  A segment is a (usually empty)
 * list of SPECIAL ops, run by a generic interpreter, followed by GROUPS.  A
 * group holds one SLOT per vector bit at a fixed position: slot j is the next
 * op whose target is vector bit j (ops on different vector bits commute, their
 * predicates never involve vector bits), so the common gates are dispatched
 * with a uniform load at a static offset and a 5-way uniform branch instead of
 * an op-stream walk.  Everything else (general complex or lane-dependent
 * matrices, pack-bit targets, multi-vector-bit phases) is a special. */
enum {              /* slot forms (one-hot bytes, so the kernel dispatches with bit tests); S = scalar of the state precision */
    S_SKIP = 0,
    S_UNIT_R = 1,   /* real unit form  a[[1,p],[q,r]]: x0 += p x1; x1 = k x1 + q x0     S: p q k a */
    S_UNIT_I = 2,   /* rx unit form    a[[1,ip],[iq,r]]                                 S: p q k a */
    S_UNIT_H = 4,   /* (only with -DQSB_UNIT_H; measured neutral on B200 and left out of the default build to keep the
                       interpreter's code small, profiles/r2/README.md) real unit form with q == 1 (Hadamard-like), unconditional, a == 1:
                       x0 += p x1; x1 = k x1 + x0 -- 4 instead of 6 packed operations per vector pair   S: p 1 k 1 */
    S_DIAG = 8,     /* phase on the vectors whose bit is set                           S: pr pi   */
    S_XDEF = 16,    /* X (swap of the two halves) under the predicate, DEFERRED: the thread only
                       flips bit j of its vector-index mask; the swap happens for free in the
                       address of the next shared / global store.  Last op on its bit in a round. */
    /* matrices that cannot take a unit form (rare after pivoting) run as G_FULL_G specials */
};
/* A group (16-byte units): [0] x = form of slot 0 | slot 1 << 8 | slot 2 << 16 | slot 3 << 24
 *                          [1] the four predicate masks: (tw & pmask) == pmask
 *                          then per slot two coefficient sets of 4 scalars: threads whose predicate
 *                          fails use set 0 (the identity for a controlled gate, the control-off matrix
 *                          for a multiplexer), the others set 1. */
#define QSB_SET16(f32) ((f32) ? 1 : 2)             /* one 4-scalar coefficient set, 16-byte units */
/* byte offsets inside a group: form byte of slot j = byte j of unit [0]; predicate mask of slot j = word j of unit [1]
 * for j < 4, word 2 + (j - 4) of unit [0] beyond (QSB_NVB = 5) */
#define QSB_GROUP_MASK_OFF(j) ((j) < 4 ? 16 + 4 * (j) : 8 + 4 * ((j) - 4))
#define QSB_GROUP16(f32) (2 + QSB_NVB * 2 * QSB_SET16(f32))

enum {              /* special op codes (generic interpreter); V = 8 bytes (f32: (lo, hi) lanes, f64: one double) */
    G_FULL_G = 16,   /* +vb: complex 2x2             V: m00r m00i m01r m01i m10r m10i m11r m11i */
    G_DIAG_V = 24,   /* +vb: phase where vector bit vb is set          V: pr pi       */
    G_DIAGA = 32,    /* +vb: MERGED controlled phases on one vector bit (QFT ladders, round 2): e^{2 pi i A} on the vectors
                        whose bit vb is set, A = the sum of the fixed-point angles of the entries this thread satisfies.
                        One sincospi + one packed complex multiply of half the vectors for the whole run, instead of a
                        multiply (+ dispatch) per gate.  Payload: 16 bytes {n_entries, -, -, -}, then n_entries GTAngle
                        entries (thread mask, outer mask, angle; 16 bytes f32 / 32 bytes f64).  No predicate of its own.
                        +QSB_NVB (f32 only): the same for phases on the PACK qubit -- the merged phase multiplies the high
                        lane of every vector (a run of G_DIAG_ALL ops). */
    G_DIAG_ALL = 40, /* phase on every vector (lane dependent)         V: pr pi       */
    G_DIAG_GEN = 41, /* phase where (v & vmask) == vmask               V: pr pi       */
    G_MATP_R = 42,   /* pack-bit target, real (f32 only)               V: A B         */
    G_MATP_G = 43,   /* pack-bit target, complex (f32 only)            V: Ar Ai Br Bi */
    G_NCODES = 44
};
/* Special op header (16 bytes):
 *   x = code | two << 8 | skip << 9 | vmask << 10 | size16 << 16
 *       two : a multiplexer -- two coefficient sets follow; threads whose predicate fails use set 0
 *             (the control-off matrix), the others set 1.  Otherwise one set: threads whose predicate
 *             fails skip the op (a controlled gate)
 *       skip: informational, the op has a predicate and one set
 *   y = 8-bit mask over threadIdx.x          z, w = 64-bit mask over the outer (per-CTA) index bits
 * predicate = (tid & y) == y && (outer & zw) == zw.  Coefficient sets are padded to 16 bytes. */
#define GOPK(code, two, skip, vmask, size16) \
    ((uint32_t)(code) | ((uint32_t)(two) << 8) | ((uint32_t)(skip) << 9) | ((uint32_t)(vmask) << 10) | ((uint32_t)(size16) << 16))
#define GOP_VMASK(x) (((x) >> 10) & 0x3fu)

struct GTPhase {               /* thread-level phase, applied through the pending scalar (32 bytes) */
    uint32_t tmask, pad;
    uint64_t omask;
    uint8_t val[16];           /* (pr, pi) in the state precision */
};

/* Unit-modulus thread-level phases, when a round holds several of them (QFT ladders): the angle as a
 * fixed-point fraction of a turn.  The kernel adds the angles of the entries a thread satisfies with integer
 * adds (exact, wraps modulo one turn) and pays ONE sincospi per round instead of a complex multiply per entry.
 * f32 passes store 16-byte entries (ang32 = the top 32 bits), f64 passes the full 32 bytes. */
struct GTAngle32 { uint32_t mask, ang32; };                /* f32 passes: two per 16-byte unit (lists are padded with a zero entry) */
struct GTAngle64 { uint32_t mask, pad; uint64_t ang64; };  /* f64 passes */
/* mask is over the per-thread PREDICATE WORD (the one slot predicates test): bits 0..QSB_TB-1 = threadIdx.x, bit
 * QSB_TB + i = outer condition GPass::cond[i]; every outer control of an angle entry takes a single-bit condition (shared
 * with the slots that test the same bit).  An entry that finds the table full stays a plain phase op. */
/* Fewer qualifying phases in a round: they stay GTPhase entries.  Break-even measured on B200 (profiles/
 * r1h_angle_ab.txt): an entry costs ~12 instructions as a complex multiply and ~5 as an angle, the sincospi
 * ~40 (f32) / ~150 (f64); with a threshold of 4 the 30 q layered circuit lost 1.8 % (f32) and 4.8 % (f64). */
#ifdef QSB_NO_TANGLE            /* A/B builds: no GTAngle entries, kernel without the angle loop */
#define QSB_TANGLE_MIN_F32 (1 << 30)
#define QSB_TANGLE_MIN_F64 (1 << 30)
#else
#define QSB_TANGLE_MIN_F32 8
#define QSB_TANGLE_MIN_F64 24
#endif
/* G_DIAGA: a run of controlled phases on one vector bit is merged from this length on (a plain phase slot costs ~32
 * packed operations + dispatch per gate; the merged form one multiply + one sincospi (~40 f32 / ~150 f64
 * instructions) + ~5 per entry) */
#define QSB_DIAGA_MIN_F32 3
#define QSB_DIAGA_MIN_F64 6
#ifndef QSB_DIAGA_MIN_PACK     /* A/B builds pass a huge value: no pack-qubit runs */
#define QSB_DIAGA_MIN_PACK 2   /* pack-qubit runs: a G_DIAG_ALL op multiplies all 2^QSB_NVB vectors (64 packed operations) */
#endif

struct GSegment {              /* 16 bytes */
    uint32_t n_special, special_off16;
    uint32_t n_groups, group_off16;
};

#ifndef QSB_MAX_COND           /* tests build the host doubles with a tiny table to exercise every full-table fallback */
#define QSB_MAX_COND 24        /* distinct outer conditions a pass can name through W */
#endif

/* Blob layout (round 2): everything a round reads lies in one contiguous run -- [GRound][its GSegment table][specials and
 * groups][thread-phase and angle lists] -- followed by the next round's run, so that a round touches few, adjacent
 * constant-cache lines and its segment table sits at a fixed distance from its header (no dependent load for its address). */
struct alignas(16) GRound {
    uint32_t n_seg, seg_off16; /* GSegment array, 16-byte units from the blob start (== this header + sizeof(GRound)) */
    uint32_t n_tph, tph_off16; /* GTPhase array                                                      */
    uint32_t flags, n_ang, next16, pad; /* next16: the next round's GRound, 16-byte units from the blob start; flags bit0: apply the pending scalar at the end of the round; n_ang: GTAngle
                                  entries, stored right after the n_tph GTPhase entries                */
    uint32_t thr_x[QSB_TB];    /* smem byte XOR per thread bit: load side | store side << 16         */
    /* smem byte XOR per VECTOR BIT, load side / store side: the slot map is GF(2)-linear, so the offset of vector v is
     * the XOR of the words of the bits set in v.  (Round 1 stored all 2^NVB combinations: 128 bytes more per round of
     * constant-cache traffic at a kernel whose uniform constant loads miss the SM-level cache 14.5 % of the time.) */
    uint32_t vld_b[QSB_NVB];
    uint32_t vst_b[QSB_NVB];
};

struct GPass {
    uint32_t n_rounds, n_runs;
    uint8_t run_start[QSB_MAX_RUNS], run_len[QSB_MAX_RUNS];
    uint64_t src_fixed;        /* constant source index bits (rank bits): predicates see them         */
    uint64_t n_tiles;
    uint32_t nloc, rounds_off16;
    uint32_t n_cond, flags;    /* flags: QSB_PASS_SYNC_SCATTER */
    uint64_t cond[QSB_MAX_COND]; /* outer conditions: W bit i = (outer & cond[i]) == cond[i]          */
    uint64_t ld_thr[QSB_TB];   /* local BYTE offset of thread bit j, round-0 gather                   */
    uint64_t st_thr[QSB_TB];   /*                                    last-round scatter               */
    uint64_t ld_vec[QSB_NV];   /* local BYTE offset of vector v                                       */
    uint64_t st_vec[QSB_NV];
    uint64_t st_fixed;         /* constant part of the scatter offset (fused exchange: the writer's rank
                                  bits on the local positions the victims leave)                      */
    uint64_t xo_mask;          /* outer bits that leave the local index in a fused exchange (DevPass::xo_pos) */
    uint32_t n_xo, xo_pad;
    uint8_t xo_pos[8], xo_rank[8];
    /* Scatter offsets are  local byte offset | destination-rank contribution << QSB_RANK_SHIFT.  The rank
     * field is non-zero only in a fused-exchange pass (peer stores), whose kernel variant adds the fields
     * up and indexes the peer-pointer table with the result. */
};
/* An in-place pass whose only round relocates qubits inside the tile (the permutation pass in front of an NCCL /
 * pipelined exchange) stores to addresses OTHER threads of the CTA gather from; with a single round there is no
 * exchange barrier between the gather and the scatter, so the kernel must place one. */
#define QSB_PASS_SYNC_SCATTER 1u
#define QSB_RANK_SHIFT 48
#define QSB_MAX_PEERS 16
struct PeerTab { char *p[QSB_MAX_PEERS]; uint64_t shard_bytes; uint32_t world, pad; };

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
    bool is_swap = false;                /* exchange marker: NCCL all-to-all of contiguous chunks             */
    bool fused_swap = false;             /* this pass scatters into the peers' shards: it IS the exchange      */
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
void tiled_plan_end_perm(const TiledPlan *p, BitPerm *out);   /* qubit layout the plan leaves behind */
bool tiled_plan_starts_at(const TiledPlan *p, const BitPerm &perm);   /* was the plan made for this layout? */
struct qsb_sim;
int tiled_execute(qsb_sim *s, TiledPlan *p);
double tiled_last_exchange_ms(const TiledPlan *p);
void tiled_comm_destroy(qsb_sim *s);

/* host-only: plans the candidates for the knobs the caller left open (exchange threshold x lane policy when sharded, four
 * hill-climbing orders on one GPU) on host threads and returns the cheapest schedule; *out is new'ed */
int tiled_plan_search(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                      const std::vector<COp> &cops, const double gphase[2], TiledPlan **out);
void tiled_plan_trace_suppress(bool off);   /* QSB_PLAN_TRACE output of the calling thread's tiled_schedule calls on / off */
/* host-only planner entry (no CUDA): used by tiled_plan_build and by the test emulator */
int tiled_schedule(int n, int prec, int g, int nloc, int rank, const qsb_options_t *opt, const BitPerm &start,
                   const std::vector<COp> &cops, const double gphase[2], TiledPlan *plan, int climb_variant = 0);
