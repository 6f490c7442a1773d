/* common.cuh -- device-side layout helpers and host-side structs shared by the .cu files. */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include <string>

#include "qsb_internal.h"

/*
 * State layout in HBM (per rank, N_loc = 2^nloc amplitudes):
 *   f64 : interleaved (re, im) doubles, 16 B per amplitude -- the memory image
 *         of the reference's `double complex v[]` (quantum_simulator.c:125).
 *   f32 : "pair-interleaved": amplitudes 2k and 2k+1 share one 16-byte unit
 *         { re[2k], re[2k+1], im[2k], im[2k+1] }.  One 128-bit load yields two
 *         register pairs that are directly usable as packed operands of
 *         FFMA2/FMUL2 (Blackwell packed fp32), with physical bit 0 as the pack
 *         dimension.  (The reference's CUDA variants use two separate float
 *         planes, naive.cu:148-149; this layout keeps their planar arithmetic
 *         but restores 16-byte contiguity per amplitude pair.)
 */
template <typename R> struct Lay;
template <> struct Lay<float> {
    static __host__ __device__ __forceinline__ uint64_t re(uint64_t i) { return ((i >> 1) << 2) | (i & 1); }
    static __host__ __device__ __forceinline__ uint64_t im(uint64_t i) { return (((i >> 1) << 2) | (i & 1)) + 2; }
};
template <> struct Lay<double> {
    static __host__ __device__ __forceinline__ uint64_t re(uint64_t i) { return 2 * i; }
    static __host__ __device__ __forceinline__ uint64_t im(uint64_t i) { return 2 * i + 1; }
};

#define QSB_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            qsb_set_error("%s in %s at line %d", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return (e_ == cudaErrorMemoryAllocation) ? QSB_ERR_NOMEM : QSB_ERR_CUDA;        \
        }                                                                                   \
    } while (0)

/* Canonical op: what a source gate becomes before scheduling (logical qubits). */
enum { C_MAT = 0, C_PHASE = 1, C_X = 2, C_MUX = 3 };
struct COp {
    int kind;
    int target;      /* C_MAT / C_X / C_MUX */
    uint64_t ctrl;   /* control mask; for C_PHASE the full mask the phase is conditioned on */
    double m[8];     /* C_MAT: row-major 2x2; C_PHASE: m[0], m[1] = phase (re, im); C_MUX: matrix when the controls are NOT all set */
    double m2[8];    /* C_MUX: matrix when all controls are set (a CX absorbed into a neighbouring gate) */
};

struct BitPerm { int8_t pos[64]; }; /* logical qubit -> physical bit position */
