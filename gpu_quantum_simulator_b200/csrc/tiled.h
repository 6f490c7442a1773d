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
 * Replaces (reference, /root/reference/): the per-gate launches of
 * naive.cu:163-189, the 2x2/4x4 host fusion of preproces.cu:215-269 and
 * 4x4.cu:327-501, and the static relabel of 4x4_permute.cu:350-434.
 */
#pragma once
#include <stdint.h>
#include <vector>

#include "common.cuh"

#define QSB_NVB 4              /* vector bits per round                      */
#define QSB_NV (1 << QSB_NVB)  /* vectors per thread                         */
#define QSB_THREADS 256
#define QSB_TB 8               /* thread bits: log2(QSB_THREADS)             */
#define QSB_T_F32 13           /* tile bits f32: pack + NVB + TB             */
#define QSB_T_F64 12           /* tile bits f64: NVB + TB                    */
#define QSB_MAX_RUNS 16

/* ---- op codes ---------------------------------------------------------- */
enum {
    OP_END = 0,
    OP_MAT_R = 1,  /* target = vector bit; all entries real                  */
    OP_MAT_I = 2,  /* target = vector bit; real diagonal, imaginary off-diag */
    OP_MAT_G = 3,  /* target = vector bit; general complex                   */
    OP_MATP_R = 4, /* target = pack bit (f32 only); real                     */
    OP_MATP_G = 5, /* target = pack bit (f32 only); general                  */
    OP_X = 6,      /* swap along a vector bit (X / CX / CCX ...)             */
    OP_XP = 7,     /* swap along the pack bit                                */
    OP_DIAG = 8,   /* phase on vectors selected by vmask (per pack lane)     */
    OP_TPHASE = 9, /* thread-level phase: folded into a per-thread scalar    */
};

/* kind = opcode | vb << 8 | lanes << 12 | mux << 16
 *   vb    : target vector bit (OP_MAT_*, OP_X)
 *   lanes : which pack lanes the op acts on (bit0 = lane .x, bit1 = .y); f64: 1
 *   mux   : thread-level multiplexer: threads whose tmask test fails use
 *           coefficient set 0 instead of skipping                           */
#define OPK(op, vb, lanes, mux) ((uint32_t)(op) | ((uint32_t)(vb) << 8) | ((uint32_t)(lanes) << 12) | ((uint32_t)(mux) << 16))

/* One device op.  Same byte size for both precisions (144 B):
 *   f32: c[32] = two sets of 8 coefficient VECTORS (lo, hi pack lane)
 *   f64: c[16] = two sets of 8 coefficients
 * Coefficient order inside a set: m00r m00i m01r m01i m10r m10i m11r m11i
 * (R form uses the four real parts, I form uses m00r m01i m10i m11r).
 * OP_DIAG / OP_TPHASE use the first two entries of set 1 as the phase.       */
template <typename R> struct DevOp {
    uint32_t kind;
    uint32_t vmask;  /* condition on the vector index: (v & vmask) == vmask   */
    uint64_t tmask;  /* condition on the thread's physical index bits         */
    R c[128 / sizeof(R)];
};

/* Uniform per-round tables. */
struct DevRound {
    uint32_t n_ops;
    uint32_t op_begin;          /* index into the pass' op array              */
    uint32_t flags;             /* bit0: round has OP_TPHASE ops              */
    uint32_t pad;
    uint64_t thr_gidx[QSB_TB];  /* thread bit j set -> these physical SOURCE index bits are set (for tmask tests; round 0: load address) */
    uint64_t vec_gidx[QSB_NVB]; /* same for the vector bits (round 0: load address)  */
    uint16_t ld_thr[QSB_TB];    /* smem slot XOR constants, load side (unused in round 0)   */
    uint16_t ld_vec[QSB_NVB];
    uint16_t st_thr[QSB_TB];    /* store side (unused in the last round)      */
    uint16_t st_vec[QSB_NVB];
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
};

/* ---- host-side plan ----------------------------------------------------- */
struct HostOp {              /* precision-independent, fp64 coefficients      */
    uint32_t kind;
    uint32_t vmask;
    uint64_t tmask;
    double c[32];            /* f32 layout (lane-expanded); f64 uses even entries' .lo semantics, see pack_op() */
};

struct HostPass {
    DevPass hdr;
    std::vector<DevRound> rounds;
    std::vector<HostOp> ops;
    /* bookkeeping for tests / emulation */
    int T = 0;                           /* tile bits                          */
    int8_t tile_src[16];                 /* tile bit j -> source physical bit  */
    int8_t tile_dst[16];                 /* tile bit j -> destination physical bit */
    std::vector<std::vector<int8_t>> round_thr; /* per round: thread bit j -> tile bit */
    std::vector<std::vector<int8_t>> round_vec; /* per round: vector bit j -> tile bit */
    bool is_swap = false;
    int n_source_ops = 0;
};

struct TiledPlan {
    int n = 0, prec = QSB_F32, g = 0, nloc = 0, rank = 0;
    std::vector<HostPass> passes;
    BitPerm start_perm{}, end_perm{};
    /* device image */
    void *d_blob = nullptr;
    size_t blob_bytes = 0;
    std::vector<size_t> pass_off, round_off, op_off; /* byte offsets into the blob */
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
