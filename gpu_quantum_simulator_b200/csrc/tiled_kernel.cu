/*
 * tiled_kernel.cu -- the sm_100a apply kernel: one launch = one PASS (one
 * read + one write of the local state), any number of gates.
 *
 * See tiled.h for the schedule this kernel interprets.  Per CTA (128 threads,
 * four CTAs per SM; QSB_TB in tiled.h):
 *   1. gather: 16 x 128-bit loads per thread pull the 32 KiB tile straight from
 *      HBM into registers; a warp instruction covers 128-byte contiguous segments
 *      (f32: one load = {re,re',im,im'} of an amplitude pair = two packed
 *      operands for FFMA2/FMUL2; f64: one (re,im) double2).
 *   2. per round: an interpreter applies the fused gates whose target is one of
 *      the 4 register-resident vector bits (or the pack bit); controls and
 *      diagonal phases are per-thread predicates.  Between rounds the tile is
 *      transposed through shared memory in 16-byte slots with a planner-chosen
 *      GF(2)-linear, bank-conflict-free slot map.
 *   3. scatter: 16 x 128-bit stores per thread.
 * Everything that does not depend on the thread is pre-computed by the planner
 * (byte offsets per vector, smem XOR constants, predicate masks) and travels
 * with the op stream as ONE __grid_constant__ kernel parameter: all table and
 * coefficient reads are uniform constant-bank loads, the only global traffic is
 * the state itself.  Algorithmic traffic per launch is 2 * N_loc *
 * sizeof(amplitude), independent of how many gates the pass fuses.
 *
 * Reference kernels replaced: kernel_gate / kernel_gate_2 (naive.cu:72-95),
 * kernel_cnot (naive.cu:97-122), kernel_gate_4 (4x4.cu:109-146).
 */
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include "sim.h"
#include "tiled.h"

/* ------------------------------------------------------------ vector algebra
 * All updates are written as IN-PLACE inline PTX (read-write operands) so the
 * tile keeps the same registers through every op body of the interpreter.
 *   f32: V = b64 register holding (lo lane, hi lane) -> mul.f32x2 / fma.rn.f32x2
 *   f64: V = double                                   -> mul.f64   / fma.rn.f64  */
template <typename R> struct VT;
template <> struct VT<float> {
    typedef unsigned long long V;
    typedef float S;
    static __device__ __forceinline__ V mul(V a, V b) { V d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
    /* acc += a * b */
    static __device__ __forceinline__ void acc(V &acc_, V a, V b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc_) : "l"(a), "l"(b)); }
    /* x = a * x + t */
    static __device__ __forceinline__ void upd(V &x, V a, V t) { asm("fma.rn.f32x2 %0, %1, %0, %2;" : "+l"(x) : "l"(a), "l"(t)); }
    /* x = a * x + t, ordered after the producers of d1..d3 (they read the old x): keeps ptxas from
     * hoisting the in-place update above those reads, which would cost a register copy */
    static __device__ __forceinline__ void updd(V &x, V a, V t, V d1) { asm("fma.rn.f32x2 %0, %1, %0, %2; // %3" : "+l"(x) : "l"(a), "l"(t), "l"(d1)); }
    static __device__ __forceinline__ void updd(V &x, V a, V t, V d1, V d2, V d3) { asm("fma.rn.f32x2 %0, %1, %0, %2; // %3 %4 %5" : "+l"(x) : "l"(a), "l"(t), "l"(d1), "l"(d2), "l"(d3)); }
    /* negated products: ptxas folds the neg into the operand modifier of FMUL2 / FFMA2 (no extra register) */
    static __device__ __forceinline__ V nmul(V a, V b) { V d; asm("{ .reg .b64 t; .reg .f32 lo, hi; mov.b64 {lo, hi}, %1; neg.f32 lo, lo; neg.f32 hi, hi; mov.b64 t, {lo, hi}; mul.rn.f32x2 %0, t, %2; }" : "=l"(d) : "l"(a), "l"(b)); return d; }
    static __device__ __forceinline__ void nacc(V &acc_, V a, V b) { asm("{ .reg .b64 t; .reg .f32 lo, hi; mov.b64 {lo, hi}, %1; neg.f32 lo, lo; neg.f32 hi, hi; mov.b64 t, {lo, hi}; fma.rn.f32x2 %0, t, %2, %0; }" : "+l"(acc_) : "l"(a), "l"(b)); }
    static __device__ __forceinline__ V bc(float s) { unsigned u = __float_as_uint(s); return ((V)u << 32) | u; }
    static __device__ __forceinline__ V lanes(float lo, float hi) { return ((V)__float_as_uint(hi) << 32) | __float_as_uint(lo); }
    static __device__ __forceinline__ V swp(V a) { return (a >> 32) | (a << 32); }
    static __device__ __forceinline__ V as_v(uint2 u) { return ((V)u.y << 32) | u.x; }
    /* scalar k of a coefficient set (16-byte units starting at c) */
    static __device__ __forceinline__ void scalars4(const uint4 *c, S &a, S &b, S &cc, S &d)
    {
        const uint4 u = c[0];
        a = __uint_as_float(u.x); b = __uint_as_float(u.y); cc = __uint_as_float(u.z); d = __uint_as_float(u.w);
    }
    static __device__ __forceinline__ void vec2(const uint4 *c, int k, V &x, V &y) { const uint4 u = c[k]; x = ((V)u.y << 32) | u.x; y = ((V)u.w << 32) | u.z; }
    static __device__ __forceinline__ void tph(const uint4 u, S &pr, S &pi) { pr = __uint_as_float(u.x); pi = __uint_as_float(u.y); }
    enum { SET4 = 1 };   /* 16-byte units per 4-scalar coefficient set */
    /* angle entries (GTAngle32): 8 bytes {mask over the predicate word tw, top 32 bits of the turn fraction}, TWO per 16-byte unit */
    typedef uint32_t A;
    /* acc += ang where the word satisfies the mask, as ONE logic op writing a predicate (mask & ~word == 0) + ONE predicated
     * add; from C++ nvcc makes LOP3 + ISETP + SEL + IADD3 of it (ncu, call 29: angle entries were 20 % of a QFT pass) */
    static __device__ __forceinline__ void ang_add(uint32_t mask, uint32_t ang, uint32_t aw, A &acc)
    {
        asm("{ .reg .pred p; .reg .b32 t;\n\t"
            "lop3.b32 t, %1, %2, 0, 0x30;\n\t"          /* mask & ~word */
            "setp.eq.u32 p, t, 0;\n\t"
            "@p add.u32 %0, %0, %3;\n\t"
            "}" : "+r"(acc) : "r"(mask), "r"(aw), "r"(ang));
    }
    static __device__ __forceinline__ void ang_acc(const uint4 u, uint32_t aw, A &acc)
    {
        ang_add(u.x, u.y, aw, acc);
        ang_add(u.z, u.w, aw, acc);
    }
    static __device__ __forceinline__ void turn(A acc, S &c, S &s) { sincospif((float)(int)acc * 4.656612873077393e-10f, &s, &c); }   /* acc / 2^31 half-turns */
};
template <> struct VT<double> {
    typedef double V;
    typedef double S;
    static __device__ __forceinline__ V mul(V a, V b) { V d; asm("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b)); return d; }
    static __device__ __forceinline__ void acc(V &acc_, V a, V b) { asm("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc_) : "d"(a), "d"(b)); }
    static __device__ __forceinline__ void upd(V &x, V a, V t) { asm("fma.rn.f64 %0, %1, %0, %2;" : "+d"(x) : "d"(a), "d"(t)); }
    static __device__ __forceinline__ void updd(V &x, V a, V t, V d1) { asm("fma.rn.f64 %0, %1, %0, %2; // %3" : "+d"(x) : "d"(a), "d"(t), "d"(d1)); }
    static __device__ __forceinline__ void updd(V &x, V a, V t, V d1, V d2, V d3) { asm("fma.rn.f64 %0, %1, %0, %2; // %3 %4 %5" : "+d"(x) : "d"(a), "d"(t), "d"(d1), "d"(d2), "d"(d3)); }
    static __device__ __forceinline__ V nmul(V a, V b) { V d; asm("{ .reg .f64 t; neg.f64 t, %1; mul.rn.f64 %0, t, %2; }" : "=d"(d) : "d"(a), "d"(b)); return d; }
    static __device__ __forceinline__ void nacc(V &acc_, V a, V b) { asm("{ .reg .f64 t; neg.f64 t, %1; fma.rn.f64 %0, t, %2, %0; }" : "+d"(acc_) : "d"(a), "d"(b)); }
    static __device__ __forceinline__ V bc(double s) { return s; }
    static __device__ __forceinline__ V lanes(double, double hi) { return hi; }   /* no pack qubit in f64 tiles: never emitted */
    static __device__ __forceinline__ V swp(V a) { return a; }
    static __device__ __forceinline__ double lohi(unsigned lo, unsigned hi) { return __hiloint2double((int)hi, (int)lo); }
    static __device__ __forceinline__ void scalars4(const uint4 *c, S &a, S &b, S &cc, S &d)
    {
        const uint4 u = c[0], w = c[1];
        a = lohi(u.x, u.y); b = lohi(u.z, u.w); cc = lohi(w.x, w.y); d = lohi(w.z, w.w);
    }
    static __device__ __forceinline__ void vec2(const uint4 *c, int k, V &x, V &y) { const uint4 u = c[k]; x = lohi(u.x, u.y); y = lohi(u.z, u.w); }
    static __device__ __forceinline__ void tph(const uint4 u, S &pr, S &pi) { pr = lohi(u.x, u.y); pi = lohi(u.z, u.w); }
    enum { SET4 = 2 };
    /* angle entries (GTAngle64): 16 bytes {mask, -, 64-bit turn fraction}, one per unit */
    typedef uint64_t A;
    static __device__ __forceinline__ void ang_acc(const uint4 u, uint32_t aw, A &acc)
    {
        asm("{ .reg .pred p; .reg .b32 t; .reg .b64 a;\n\t"
            "lop3.b32 t, %1, %2, 0, 0x30;\n\t"          /* mask & ~word */
            "setp.eq.u32 p, t, 0;\n\t"
            "mov.b64 a, {%3, %4};\n\t"
            "@p add.u64 %0, %0, a;\n\t"
            "}" : "+l"(acc) : "r"(u.x), "r"(aw), "r"(u.z), "r"(u.w));
    }
    static __device__ __forceinline__ void turn(A acc, S &c, S &s) { sincospi((double)(long long)acc * 1.0842021724855044e-19, &s, &c); }   /* acc / 2^63 half-turns */
};

#define NV QSB_NV

/* (xr, xi) *= (pr, pi) */
template <typename R>
__device__ __forceinline__ void cmul_inplace(typename VT<R>::V &xr, typename VT<R>::V &xi, typename VT<R>::V pr, typename VT<R>::V pi)
{
    typedef VT<R> T;
    const typename T::V t0 = T::nmul(pi, xi), t1 = T::mul(pi, xr);
    T::updd(xr, pr, t0, t1);
    T::updd(xi, pr, t1, t0);
}

/* unit form on vector bit VB: strictly in-place dependency chains.
 *   real: x0 += p*x1 ; x1 = k*x1 + q*x0        rx form: x0 += i p x1 ; x1 = k*x1 + i q x0 */
template <typename R, int VB, bool IMAG>
__device__ __forceinline__ void unit_v(typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], typename VT<R>::S ps, typename VT<R>::S qs, typename VT<R>::S ks)
{
    typedef VT<R> T; typedef typename T::V V;
    const V p = T::bc(ps), q = T::bc(qs), k = T::bc(ks);
    if (!IMAG) {
#pragma unroll
        for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
            const int w = v | (1 << VB);
            T::acc(re[v], p, re[w]); T::acc(im[v], p, im[w]);
            const V t = T::mul(q, re[v]), u = T::mul(q, im[v]);
            T::upd(re[w], k, t); T::upd(im[w], k, u);
        }
    } else {
#pragma unroll
        for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
            const int w = v | (1 << VB);
            T::nacc(re[v], p, im[w]); T::acc(im[v], p, re[w]);
            const V t = T::nmul(q, im[v]), u = T::mul(q, re[v]);
            T::upd(re[w], k, t); T::upd(im[w], k, u);
        }
    }
}

/* real unit form with q == 1 (Hadamard-like): x0 += p*x1 ; x1 = k*x1 + x0 -- two packed operations per component */
template <typename R, int VB>
__device__ __forceinline__ void unit_h(typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], typename VT<R>::S ps, typename VT<R>::S ks)
{
    typedef VT<R> T; typedef typename T::V V;
    const V p = T::bc(ps), k = T::bc(ks);
#pragma unroll
    for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
        const int w = v | (1 << VB);
        T::acc(re[v], p, re[w]); T::acc(im[v], p, im[w]);
        T::upd(re[w], k, re[v]); T::upd(im[w], k, im[v]);
    }
}

/* complex 2x2 on vector bit VB; coefficients are V-typed (lane pairs in f32) */
template <typename R, int VB>
__device__ __forceinline__ void gen_v(typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], const typename VT<R>::V (&m)[8])
{
    typedef VT<R> T; typedef typename T::V V;
    const V ar = m[0], ai = m[1], br = m[2], bi = m[3], cr = m[4], ci = m[5], dr = m[6], di = m[7];
#pragma unroll
    for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
        const int w = v | (1 << VB);
        V t0 = T::nmul(ai, im[v]); T::acc(t0, br, re[w]); T::nacc(t0, bi, im[w]);   /* re[v] minus its own-term */
        V t1 = T::mul(ai, re[v]);  T::acc(t1, br, im[w]); T::acc(t1, bi, re[w]);    /* im[v] */
        V t2 = T::mul(cr, re[v]);  T::nacc(t2, ci, im[v]); T::nacc(t2, di, im[w]);  /* re[w] */
        V t3 = T::mul(cr, im[v]);  T::acc(t3, ci, re[v]); T::acc(t3, di, re[w]);    /* im[w] */
        T::updd(re[v], ar, t0, t1, t2, t3); T::updd(im[v], ar, t1, t0, t2, t3);
        T::updd(re[w], dr, t2, t0, t1, t3); T::updd(im[w], dr, t3, t0, t1, t2);
    }
}

/* phase on the vectors whose bit VB is set */
template <typename R, int VB>
__device__ __forceinline__ void diag_v(typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], typename VT<R>::V pr, typename VT<R>::V pi)
{
#pragma unroll
    for (int v = 0; v < NV; v++) if ((v >> VB) & 1) cmul_inplace<R>(re[v], im[v], pr, pi);
}
template <typename R, int MASK>
__device__ __forceinline__ void diag_mask(typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], typename VT<R>::V pr, typename VT<R>::V pi)
{
#pragma unroll
    for (int v = 0; v < NV; v++) if ((v & MASK) == MASK) cmul_inplace<R>(re[v], im[v], pr, pi);
}
/* (pr, pi) of a diagonal op: set 0, or set 1 for the threads that pass (uniform loads + select) */
template <typename R>
__device__ __forceinline__ void load_phase(const uint4 *c, bool two, bool s1, typename VT<R>::V &pr, typename VT<R>::V &pi)
{
    typedef VT<R> T;
    T::vec2(c, 0, pr, pi);
    if (two) { typename T::V pr1, pi1; T::vec2(c, 1, pr1, pi1); if (s1) { pr = pr1; pi = pi1; } }
}

/* phase on the vectors v with (v & VM) == VM (a controlled phase between vector-bit qubits) */
template <typename R, int VM>
__device__ __forceinline__ void diag_gen(typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], typename VT<R>::V pr, typename VT<R>::V pi)
{
#pragma unroll
    for (int v = 0; v < NV; v++) if ((v & VM) == VM) cmul_inplace<R>(re[v], im[v], pr, pi);
}

/* 2x2 on the pack bit (f32 only): out = A * x + B * swap(x), A = (m00, m11), B = (m01, m10) */
__device__ __forceinline__ void matp_r(unsigned long long (&re)[NV], unsigned long long (&im)[NV], unsigned long long A, unsigned long long B)
{
    typedef VT<float> T; typedef T::V V;
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const V t0 = T::mul(B, T::swp(re[v])), t1 = T::mul(B, T::swp(im[v]));
        T::upd(re[v], A, t0); T::upd(im[v], A, t1);
    }
}
__device__ __forceinline__ void matp_g(unsigned long long (&re)[NV], unsigned long long (&im)[NV], unsigned long long Ar, unsigned long long Ai, unsigned long long Br, unsigned long long Bi)
{
    typedef VT<float> T; typedef T::V V;
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const V sr = T::swp(re[v]), si = T::swp(im[v]);
        V t0 = T::nmul(Ai, im[v]); T::acc(t0, Br, sr); T::nacc(t0, Bi, si);
        V t1 = T::mul(Ai, re[v]);  T::acc(t1, Br, si); T::acc(t1, Bi, sr);
        T::updd(re[v], Ar, t0, t1); T::updd(im[v], Ar, t1, t0);
    }
}
__device__ __forceinline__ void matp_r(double (&)[NV], double (&)[NV], double, double) {}
__device__ __forceinline__ void matp_g(double (&)[NV], double (&)[NV], double, double, double, double) {}

/* ---------------------------------------------------------- global / shared IO
 * One 16-byte unit per vector in both precisions: f32 {re0, re1, im0, im1} (an
 * amplitude pair), f64 {re, im}. */
template <typename R> struct IO;
template <> struct IO<float> {
    typedef unsigned long long V;
    static __device__ __forceinline__ void gload(const char *p, V &re, V &im) { const ulonglong2 x = __ldcs(reinterpret_cast<const ulonglong2 *>(p)); re = x.x; im = x.y; }
    static __device__ __forceinline__ void gstore(char *p, V re, V im) { __stcs(reinterpret_cast<ulonglong2 *>(p), make_ulonglong2(re, im)); }
    static __device__ __forceinline__ void sload(const uint8_t *sm, uint32_t off, V &re, V &im) { const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(sm + off); re = x.x; im = x.y; }
    static __device__ __forceinline__ void sstore(uint8_t *sm, uint32_t off, V re, V im) { *reinterpret_cast<ulonglong2 *>(sm + off) = make_ulonglong2(re, im); }
};
template <> struct IO<double> {
    typedef double V;
    static __device__ __forceinline__ void gload(const char *p, V &re, V &im) { const double2 x = __ldcs(reinterpret_cast<const double2 *>(p)); re = x.x; im = x.y; }
    static __device__ __forceinline__ void gstore(char *p, V re, V im) { __stcs(reinterpret_cast<double2 *>(p), make_double2(re, im)); }
    static __device__ __forceinline__ void sload(const uint8_t *sm, uint32_t off, V &re, V &im) { const double2 x = *reinterpret_cast<const double2 *>(sm + off); re = x.x; im = x.y; }
    static __device__ __forceinline__ void sstore(uint8_t *sm, uint32_t off, V re, V im) { *reinterpret_cast<double2 *>(sm + off) = make_double2(re, im); }
};

/* ------------------------------------------------------------------ the kernel */
#if QSB_NVB == 5
#define CASE_VB4(base, ...) case (base) + 4: { enum { VB = 4 }; __VA_ARGS__ } break;
#else
#define CASE_VB4(base, ...)
#endif
#define CASE4(base, ...)                              \
    case (base) + 0: { enum { VB = 0 }; __VA_ARGS__ } break; \
    case (base) + 1: { enum { VB = 1 }; __VA_ARGS__ } break; \
    case (base) + 2: { enum { VB = 2 }; __VA_ARGS__ } break; \
    case (base) + 3: { enum { VB = 3 }; __VA_ARGS__ } break; \
    CASE_VB4(base, __VA_ARGS__)

/* One slot of a group = the op on vector bit VB: a one-hot form byte, a predicate mask and two
 * 4-scalar coefficient sets.  Slots are software-pipelined: the coefficient loads of the next
 * non-empty slot are issued before the current slot's packed FMAs, the header of the next group
 * before the current group.  No per-thread branches: threads whose predicate fails use coefficient
 * set 0 (the identity for a controlled gate, the control-off matrix for a multiplexer). */
template <typename R> struct SlotC { typename VT<R>::S c[4], d[4]; };
template <typename R>
__device__ __forceinline__ void slot_fetch(const uint4 *cp, SlotC<R> &s)
{
    typedef VT<R> T;
    T::scalars4(cp, s.c[0], s.c[1], s.c[2], s.c[3]);
    T::scalars4(cp + T::SET4, s.d[0], s.d[1], s.d[2], s.d[3]);
}
template <typename R, int VB>
__device__ __forceinline__ void slot_exec(uint32_t form, uint32_t pmask, const SlotC<R> &s, typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV],
                                          uint32_t tw, typename VT<R>::S &psr, typename VT<R>::S &psi, uint32_t &xm)
{
    typedef VT<R> T; typedef typename T::S S;
    const bool pred = (tw & pmask) == pmask;
    const S c0 = pred ? s.d[0] : s.c[0], c1 = pred ? s.d[1] : s.c[1], c2 = pred ? s.d[2] : s.c[2], c3 = pred ? s.d[3] : s.c[3];
    if (form & S_UNIT_R) { unit_v<R, VB, false>(re, im, c0, c1, c2); psr *= c3; psi *= c3; }
    else if (form & S_UNIT_I) { unit_v<R, VB, true>(re, im, c0, c1, c2); psr *= c3; psi *= c3; }
#ifdef QSB_UNIT_H
    else if (form & S_UNIT_H) unit_h<R, VB>(re, im, c0, c2);
#endif
    else if (form & S_DIAG) diag_v<R, VB>(re, im, T::bc(c0), T::bc(c1));
    /* S_XDEF alone, or merged into the gate it follows (the planner then gives both the same predicate) */
    if (form & S_XDEF) xm ^= pred ? (1u << VB) : 0u;
}

/* PEER: the scatter of a fused-exchange pass -- every amplitude goes straight into the shard of the rank
 * named by its victim bits (peer memory over NVLink), so the qubit exchange costs no extra sweep. */
template <typename R, int BLOB, bool PEER>
__global__ void __launch_bounds__(QSB_THREADS, QSB_CTAS_PER_SM)
k_tile_pass(const __grid_constant__ PassBlob<BLOB> blob, const char *src, char *dst, const __grid_constant__ PeerTab peers, uint32_t tile_base)
{
    typedef VT<R> T; typedef typename T::V V; typedef typename T::S S;
    static_assert(QSB_NVB == 4 || QSB_NVB == 5, "the interpreter is written for 4 or 5 vector bits");
    extern __shared__ __align__(16) uint8_t smem[];
    const uint4 *B = blob.q;
    const GPass &P = *reinterpret_cast<const GPass *>(B);
    const uint32_t tid = threadIdx.x;
    const uint64_t AMP = sizeof(R) * 2;

    /* tile id -> outer index bits (uniform) */
    uint64_t tile = (uint64_t)blockIdx.x + tile_base, outer = 0;   /* tile_base: a pass may be launched in slices (pipelined exchange) */
    {
        const int nr = (int)P.n_runs;
        for (int r = 0; r < nr; r++) {
            const int len = P.run_len[r];
            outer |= (tile & ((1ULL << len) - 1)) << P.run_start[r];
            tile >>= len;
        }
    }
    const uint64_t src_outer = outer | P.src_fixed;   /* what the outer predicates test */
    /* per-thread predicate word: tid in bits 0..7, the CTA's outer-condition bits above */
    uint32_t tw = tid;
    {
        const int nc = (int)P.n_cond;
        for (int i = 0; i < nc; i++) { const uint64_t m = P.cond[i]; if ((src_outer & m) == m) tw |= (uint32_t)QSB_THREADS << i; }
    }
    const int n_rounds = (int)P.n_rounds;

    V re[NV], im[NV];

    /* ---- gather ---- */
    {
        uint64_t off = outer * AMP;
#pragma unroll
        for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) off += P.ld_thr[j];
        const char *p = src + off;
#pragma unroll
        for (int v = 0; v < NV; v++) IO<R>::gload(p + P.ld_vec[v], re[v], im[v]);
    }

    uint32_t xm = 0;   /* deferred X: this thread's register v holds logical vector v ^ xm */
    const uint4 *rp = B + P.rounds_off16;
    for (int rd = 0; rd < n_rounds; rd++, rp = B + reinterpret_cast<const GRound *>(rp)->next16) {   /* every round's tables are one contiguous run (tiled.h) */
        const GRound &RD = *reinterpret_cast<const GRound *>(rp);
        uint32_t sb = 0; /* this thread's smem byte offset: load side in the low half, store side in the high half */
#pragma unroll
        for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) sb ^= RD.thr_x[j];

        if (rd > 0) {
            const uint32_t sl = sb & 0xffffu;
            /* the slot map is GF(2)-linear: the 2^NVB vector offsets are the XOR combinations of NVB basis words (one
             * uniform load + uniform XORs instead of 2^NVB table loads) */
            uint32_t bl[QSB_NVB];
#pragma unroll
            for (int b = 0; b < QSB_NVB; b++) bl[b] = RD.vld_b[b];
#pragma unroll
            for (int v = 0; v < NV; v++) {
                uint32_t x = 0;
#pragma unroll
                for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) x ^= bl[b];
                IO<R>::sload(smem, sl ^ x, re[v], im[v]);
            }
            __syncthreads(); /* every thread has its registers before anyone overwrites the tile */
        }

        /* ---- the fused gates of this round ---- */
        S psr = S(1), psi = S(0); /* per-thread pending scalar: unit-form scales and thread-level phases */
        const uint32_t n_seg = RD.n_seg;
        const GSegment *seg = reinterpret_cast<const GSegment *>(rp + sizeof(GRound) / 16);   /* right behind the header */
        for (uint32_t sg = 0; sg < n_seg; sg++) {
            /* -- specials: generic interpreter -- */
            const uint32_t n_ops = seg[sg].n_special;
            uint32_t opi = seg[sg].special_off16;      /* an INDEX into the descriptor, not a pointer: keeps every load a constant-bank load */
            for (uint32_t i = 0; i < n_ops; i++) {
                const uint4 h = B[opi];
                const uint32_t ci = opi + 1;
                const uint4 *c = B + ci;
                opi += h.x >> 16;
                const uint32_t code = h.x & 0xffu;
                const uint64_t om = ((uint64_t)h.w << 32) | h.z;
                const bool two = (h.x >> 8) & 1u;
                const bool pred = ((src_outer & om) == om) && ((tid & h.y) == h.y);
                if (!two && !pred) continue;   /* controlled gate: the other threads sit this op out */
                const bool s1 = two && pred;   /* multiplexer: threads that pass use coefficient set 1 */
                if (code >= G_DIAGA && code <= G_DIAGA + QSB_NVB) {
                    /* merged controlled phases (G_DIAGA): integer angle sum over the entries this thread satisfies, ONE
                     * sincospi (one copy of its code for all vector bits), then the phase on the vectors whose bit is set */
                    const uint32_t n_u = B[ci].x;                       /* 16-byte units of angle entries */
                    typename T::A acc = 0;
                    for (uint32_t k = 0; k < n_u; k++) T::ang_acc(B[ci + 1 + k], tw, acc);
                    S apr, api; T::turn(acc, apr, api);
                    if (sizeof(R) == 4 && code == G_DIAGA + QSB_NVB) {  /* run on the pack qubit: high lane of every vector */
                        const V lpr = T::lanes(S(1), apr), lpi = T::lanes(S(0), api);
#pragma unroll
                        for (int v = 0; v < NV; v++) cmul_inplace<R>(re[v], im[v], lpr, lpi);
                        continue;
                    }
                    const V vpr = T::bc(apr), vpi = T::bc(api);
                    switch (code - G_DIAGA) {
                    CASE4(0, { diag_v<R, VB>(re, im, vpr, vpi); })
                    default: break;
                    }
                    continue;
                }
                switch (code) {
                CASE4(G_FULL_G, {
                    const uint4 *cs = c + (s1 ? 4 : 0);     /* rare form: per-thread constant loads are fine */
                    V m[8];
                    T::vec2(cs, 0, m[0], m[1]); T::vec2(cs, 1, m[2], m[3]); T::vec2(cs, 2, m[4], m[5]); T::vec2(cs, 3, m[6], m[7]);
                    gen_v<R, VB>(re, im, m);
                })
                CASE4(G_DIAG_V, {
                    V pr, pi; load_phase<R>(c, two, s1, pr, pi);
                    diag_v<R, VB>(re, im, pr, pi);
                })
                case G_DIAG_ALL: {
                    V pr, pi; load_phase<R>(c, two, s1, pr, pi);
#pragma unroll
                    for (int v = 0; v < NV; v++) cmul_inplace<R>(re[v], im[v], pr, pi);
                    break;
                }
                case G_DIAG_GEN: {
                    const uint32_t vmask = GOP_VMASK(h.x);   /* uniform */
                    V pr, pi; load_phase<R>(c, two, s1, pr, pi);
#if QSB_NVB == 4
                    /* one copy per mask: only the 4 / 2 / 1 vectors the mask selects are touched (a predicated sweep over
                     * all 16 costs 150 instructions per op -- 21 % of a QFT pass, profiles/r2/qft_pass_ncu_summary.txt) */
#define QSB_DG(m) case m: diag_gen<R, m>(re, im, pr, pi); break;
                    switch (vmask) { QSB_DG(3) QSB_DG(5) QSB_DG(6) QSB_DG(9) QSB_DG(10) QSB_DG(12) QSB_DG(7) QSB_DG(11) QSB_DG(13) QSB_DG(14) QSB_DG(15) default: break; }
#undef QSB_DG
#else
#pragma unroll
                    for (int v = 0; v < NV; v++) if ((v & vmask) == vmask) cmul_inplace<R>(re[v], im[v], pr, pi);
#endif
                    break;
                }
                case G_MATP_R: {
                    V A, Bc; T::vec2(c + (s1 ? 1 : 0), 0, A, Bc);
                    matp_r(re, im, A, Bc);
                    break;
                }
                case G_MATP_G: {
                    const uint4 *cs = c + (s1 ? 2 : 0);
                    V Ar, Ai, Br, Bi; T::vec2(cs, 0, Ar, Ai); T::vec2(cs, 1, Br, Bi);
                    matp_g(re, im, Ar, Ai, Br, Bi);
                    break;
                }
                default: break;
                }
            }
            /* -- groups: one slot per vector bit at a fixed position -- */
            const uint32_t n_groups = seg[sg].n_groups;
            const uint4 *gp = B + seg[sg].group_off16;
            if (n_groups) {
                const int G16 = QSB_GROUP16(sizeof(R) == 4), S16 = 2 * T::SET4;
                uint4 nh = gp[0], nm = gp[1];
                for (uint32_t g = 0; g < n_groups; g++, gp += G16) {
                    const uint4 gh = nh, gm = nm;
                    nh = gp[G16]; nm = gp[G16 + 1];            /* next group's header (or slack) */
                    const uint32_t f0 = gh.x & 0xffu, f1 = (gh.x >> 8) & 0xffu, f2 = (gh.x >> 16) & 0xffu, f3 = gh.x >> 24;
#if QSB_NVB == 5
                    const uint32_t f4 = gh.y & 0xffu;      /* slot 4: form byte 4, predicate mask in word 2 of the header unit */
#endif
                    const uint4 *cp = gp + 2;
                    SlotC<R> sa, sc;
                    if (f0) slot_fetch<R>(cp, sa);
                    if (f1) slot_fetch<R>(cp + S16, sc);
                    if (f0) slot_exec<R, 0>(f0, gm.x, sa, re, im, tw, psr, psi, xm);
                    if (f2) slot_fetch<R>(cp + 2 * S16, sa);
                    if (f1) slot_exec<R, 1>(f1, gm.y, sc, re, im, tw, psr, psi, xm);
                    if (f3) slot_fetch<R>(cp + 3 * S16, sc);
                    if (f2) slot_exec<R, 2>(f2, gm.z, sa, re, im, tw, psr, psi, xm);
#if QSB_NVB == 5
                    if (f4) slot_fetch<R>(cp + 4 * S16, sa);
#endif
                    if (f3) slot_exec<R, 3>(f3, gm.w, sc, re, im, tw, psr, psi, xm);
#if QSB_NVB == 5
                    if (f4) slot_exec<R, 4>(f4, gh.z, sa, re, im, tw, psr, psi, xm);
#endif
                }
            }
        }
        /* ---- thread-level phases of this round ---- */
        {
            const uint32_t n_tph = RD.n_tph;
            const uint32_t ti = RD.tph_off16;
            /* One or two entries per round on the circuits measured.  nvcc unrolls the loop 4x with the masks prefetched;
             * same-box A/B (calls 24-26): the f32 kernel is 0.9 % faster with that, the f64 kernel 1.4 % slower (code size).
             * The body is spelled twice on purpose: behind a lambda the f32 kernel allocates registers differently and
             * QFT loses 1.8 %. */
#define QSB_TPH_ENTRY                                                                  \
                const uint4 h = B[ti + 2 * i];                                         \
                const uint64_t om = ((uint64_t)h.w << 32) | h.z;                       \
                if ((src_outer & om) != om) continue;          /* uniform */          \
                S pr, pi; T::tph(B[ti + 2 * i + 1], pr, pi);                           \
                if ((tid & h.x) != h.x) { pr = S(1); pi = S(0); }                      \
                const S nr = psr * pr - psi * pi;                                      \
                psi = psr * pi + psi * pr; psr = nr;
            if (sizeof(R) == 8) {
#pragma unroll 1
                for (uint32_t i = 0; i < n_tph; i++) { QSB_TPH_ENTRY }
            } else {
                for (uint32_t i = 0; i < n_tph; i++) { QSB_TPH_ENTRY }
            }
#undef QSB_TPH_ENTRY
            /* unit-modulus phases as fixed-point angles (GTAngle): integer adds per entry, one sincospi per round */
#ifdef QSB_NO_TANGLE   /* A/B builds only: the planner must then be told not to emit GTAngle entries */
            const uint32_t n_ang = 0;
#else
            const uint32_t n_ang = RD.n_ang;
#endif
            if (n_ang) {                                               /* n_ang: 16-byte units */
                typename T::A acc = 0;
                const uint32_t ai = ti + 2 * n_tph;
                for (uint32_t i = 0; i < n_ang; i++) T::ang_acc(B[ai + i], tw, acc);
                S pr, pi; T::turn(acc, pr, pi);
                const S nr = psr * pr - psi * pi;
                psi = psr * pi + psi * pr; psr = nr;
            }
        }
        if (RD.flags & 1u) {
            if (psi != S(0)) {
                const V pr = T::bc(psr), pi = T::bc(psi);
#pragma unroll
                for (int v = 0; v < NV; v++) cmul_inplace<R>(re[v], im[v], pr, pi);
            } else if (psr != S(1)) {
                const V pr = T::bc(psr);
#pragma unroll
                for (int v = 0; v < NV; v++) { re[v] = T::mul(pr, re[v]); im[v] = T::mul(pr, im[v]); }
            }
        }

        if (rd < n_rounds - 1) {
            uint32_t ss = sb >> 16;
            /* deferred X: register v holds logical vector v ^ xm; the slot map is GF(2)-linear */
#pragma unroll
            for (int b = 0; b < QSB_NVB; b++) if ((xm >> b) & 1) ss ^= RD.vst_b[b];
            xm = 0;
            uint32_t bs[QSB_NVB];
#pragma unroll
            for (int b = 0; b < QSB_NVB; b++) bs[b] = RD.vst_b[b];
#pragma unroll
            for (int v = 0; v < NV; v++) {
                uint32_t x = 0;
#pragma unroll
                for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) x ^= bs[b];
                IO<R>::sstore(smem, ss ^ x, re[v], im[v]);
            }
            __syncthreads();
        }
    }

    /* ---- scatter ---- */
    /* single-round in-place pass that moves qubits inside the tile: other threads gather what this one overwrites */
    if (P.flags & QSB_PASS_SYNC_SCATTER) __syncthreads();
    {
        uint64_t off = outer * AMP + P.st_fixed, xoff = 0;
        if (PEER) {   /* victims outside the tile: their outer bits leave the local index and name bits of the destination rank (uniform) */
            off = (outer & ~P.xo_mask) * AMP + P.st_fixed;
            const int nx = (int)P.n_xo;
            for (int k = 0; k < nx; k++) off += ((outer >> P.xo_pos[k]) & 1ULL) << (QSB_RANK_SHIFT + P.xo_rank[k]);
        }
#pragma unroll
        for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) off += P.st_thr[j];
#pragma unroll
        for (int b = 0; b < QSB_NVB; b++) if ((xm >> b) & 1) xoff ^= P.st_vec[1 << b];
        if (!PEER) {
            char *p = dst + off;
#pragma unroll
            for (int v = 0; v < NV; v++) IO<R>::gstore(p + (P.st_vec[v] ^ xoff), re[v], im[v]);
        } else {
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const uint64_t t = off + (P.st_vec[v] ^ xoff);            /* fields add without carries */
                char *p = peers.p[t >> QSB_RANK_SHIFT] + (t & ((1ULL << QSB_RANK_SHIFT) - 1));
#ifdef QSB_DEBUG_PEER
                if ((t >> QSB_RANK_SHIFT) >= peers.world || (t & ((1ULL << QSB_RANK_SHIFT) - 1)) + 16 > peers.shard_bytes || !peers.p[t >> QSB_RANK_SHIFT]) {
                    if (threadIdx.x == 0 || true) printf("bad peer store: block %u tid %u v %d t %llx off %llx stvec %llx xoff %llx base %p\n", blockIdx.x, tid, v,
                                               (unsigned long long)t, (unsigned long long)off, (unsigned long long)P.st_vec[v], (unsigned long long)xoff, (void *)peers.p[(t >> QSB_RANK_SHIFT) & 15]);
                    continue;
                }
#endif
                IO<R>::gstore(p, re[v], im[v]);
            }
        }
    }
}

/* ------------------------------------------------------------------ launching */
/* the dynamic shared-memory size attribute (needed above 48 KiB, i.e. for the QSB_TB = 8 geometry) is per device and per instantiation */
template <typename R, int BLOB, bool PEER>
static int ensure_smem_optin(int device)
{
    static std::mutex mu;                 /* handles on different host threads may launch their first pass at the same time */
    static bool attr_set[64] = {false};
    const int dev = device & 63;
    std::lock_guard<std::mutex> lock(mu);
    if (!attr_set[dev]) {
        QSB_CUDA(cudaFuncSetAttribute(k_tile_pass<R, BLOB, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, QSB_SMEM_TOTAL));
        /* The shared-memory carve-out is left to the driver for the default geometry (4 CTAs x 33 KiB): forcing the largest
         * carve-out shrinks L1 and costs 8 % (142.8 vs 131.5 ms at 30 q, call 17 of round 2).  Builds that need more than the
         * driver's choice to reach their residency (-DQSB_CTAS_PER_SM=5: the driver settles for 4 resident CTAs) ask for it. */
        if (QSB_CTAS_PER_SM * (QSB_SMEM_TOTAL + 1024) > 132 * 1024)
            QSB_CUDA(cudaFuncSetAttribute(k_tile_pass<R, BLOB, PEER>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        if (getenv("QSB_VERBOSE_OCC")) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tile_pass<R, BLOB, PEER>, QSB_THREADS, QSB_SMEM_TOTAL);
            fprintf(stderr, "qsim_b200: k_tile_pass<%s, %d, %d>: %d resident CTAs per SM (built for %d)\n", sizeof(R) == 4 ? "float" : "double", BLOB, (int)PEER, nb, QSB_CTAS_PER_SM);
        }
        attr_set[dev] = true;
    }
    return QSB_OK;
}

template <typename R, int BLOB, bool PEER>
static int launch_one(qsb_sim *s, const HostPass &hp, const void *src, void *dst, const PeerTab &peers, uint64_t tile0, uint64_t ntile)
{
    int rc = ensure_smem_optin<R, BLOB, PEER>(s->device);
    if (rc) return rc;
    if (hp.hdr.n_tiles > 0x7fffffffULL) { qsb_set_error("too many tiles"); return QSB_ERR_ARG; }
    if (ntile == 0) { tile0 = 0; ntile = hp.hdr.n_tiles; }
    const PassBlob<BLOB> *blob = reinterpret_cast<const PassBlob<BLOB> *>(hp.blob.data());
    k_tile_pass<R, BLOB, PEER><<<(unsigned)ntile, QSB_THREADS, QSB_SMEM_TOTAL, s->stream>>>(*blob, (const char *)src, (char *)dst, peers, (uint32_t)tile0);
    QSB_CUDA(cudaGetLastError());
    return QSB_OK;
}

/* everything a stream capture must not do later (function attributes of the single-GPU instantiations) */
int tiled_prepare_capture(qsb_sim *s)
{
    int rc;
    if (s->prec == QSB_F32) {
        if ((rc = ensure_smem_optin<float, QSB_BLOB_SMALL, false>(s->device))) return rc;
        if ((rc = ensure_smem_optin<float, QSB_BLOB_MEDIUM, false>(s->device))) return rc;
        return ensure_smem_optin<float, QSB_BLOB_LARGE, false>(s->device);
    }
    if ((rc = ensure_smem_optin<double, QSB_BLOB_SMALL, false>(s->device))) return rc;
    if ((rc = ensure_smem_optin<double, QSB_BLOB_MEDIUM, false>(s->device))) return rc;
    return ensure_smem_optin<double, QSB_BLOB_LARGE, false>(s->device);
}

template <typename R>
static int launch_pass(qsb_sim *s, const HostPass &hp, const void *src, void *dst, const PeerTab *peers, uint64_t tile0, uint64_t ntile)
{
    if (peers) {
        /* the descriptor and the peer table together must stay below the 32764-byte kernel-parameter limit */
        if (hp.blob.size() <= QSB_BLOB_SMALL) return launch_one<R, QSB_BLOB_SMALL, true>(s, hp, src, dst, *peers, tile0, ntile);
        if (hp.blob.size() <= QSB_BLOB_MEDIUM) return launch_one<R, QSB_BLOB_MEDIUM, true>(s, hp, src, dst, *peers, tile0, ntile);
        return launch_one<R, QSB_BLOB_LARGE, true>(s, hp, src, dst, *peers, tile0, ntile);
    }
    static const PeerTab none = {};
    if (hp.blob.size() <= QSB_BLOB_SMALL) return launch_one<R, QSB_BLOB_SMALL, false>(s, hp, src, dst, none, tile0, ntile);
    if (hp.blob.size() <= QSB_BLOB_MEDIUM) return launch_one<R, QSB_BLOB_MEDIUM, false>(s, hp, src, dst, none, tile0, ntile);
    return launch_one<R, QSB_BLOB_LARGE, false>(s, hp, src, dst, none, tile0, ntile);
}

/* peers: the destination shards of a fused-exchange pass (indexed by rank), or null.
 * [tile0, tile0 + ntile): the slice of the pass to launch (ntile = 0: the whole pass). */
int tiled_launch_pass(qsb_sim *s, const TiledPlan *p, size_t k, const void *src, void *dst, const PeerTab *peers, uint64_t tile0, uint64_t ntile)
{
    const HostPass &hp = p->passes[k];
    return s->prec == QSB_F32 ? launch_pass<float>(s, hp, src, dst, peers, tile0, ntile) : launch_pass<double>(s, hp, src, dst, peers, tile0, ntile);
}
