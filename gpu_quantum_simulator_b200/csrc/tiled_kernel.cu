/*
 * tiled_kernel.cu -- the sm_100a apply kernel: one launch = one PASS (one
 * read + one write of the local state), any number of gates.
 *
 * See tiled.h for the schedule this kernel interprets.  Per CTA:
 *   1. 16 x 128-bit coalesced loads per thread pull the tile (64 KiB) straight
 *      from HBM into registers (f32: {re,re',im,im'} units -> two float2
 *      vectors that feed FFMA2/FMUL2 directly; f64: one (re,im) double2).
 *   2. per round: butterflies on register-resident vector bits, per-thread
 *      predicates for controls / diagonal phases, then a conflict-free
 *      exchange through shared memory (planner-chosen GF(2)-linear slot map).
 *   3. 16 x 128-bit coalesced stores per thread.
 * The pass descriptor (tables + op stream) is a __grid_constant__ parameter:
 * all table / coefficient reads are constant-bank loads.
 * The kernel is HBM-bound by design: algorithmic traffic per launch is
 * 2 * N_loc * sizeof(amplitude), independent of how many gates the pass fuses.
 *
 * Reference kernels replaced: kernel_gate / kernel_gate_2 (naive.cu:72-95),
 * kernel_cnot (naive.cu:97-122), kernel_gate_4 (4x4.cu:109-146).
 */
#include "sim.h"
#include "tiled.h"

/* ------------------------------------------------------------ vector algebra
 * All updates are written as IN-PLACE inline PTX (read-write operands) so the
 * tile keeps the same registers through every op body: without this the
 * compiler materialises a 64-register shuffle at each switch join.
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
    static __device__ __forceinline__ V neg(V a) { return a ^ 0x8000000080000000ULL; }
    static __device__ __forceinline__ V bc(float s) { unsigned u = __float_as_uint(s); return ((V)u << 32) | u; }
    static __device__ __forceinline__ V swp(V a) { return (a >> 32) | (a << 32); }
    static __device__ __forceinline__ V coef(const uint2 *c, int k) { return reinterpret_cast<const V *>(c)[k]; }
    static __device__ __forceinline__ void coef2(const uint2 *c, int k, V &x, V &y) { const ulonglong2 u = reinterpret_cast<const ulonglong2 *>(c)[k >> 1]; x = u.x; y = u.y; }
    static __device__ __forceinline__ void tphase(const uint2 *c, float &pr, float &pi) { uint2 u = c[0]; pr = __uint_as_float(u.x); pi = __uint_as_float(u.y); }
    static __device__ __forceinline__ float scalar(const uint2 *c, int k) { return __uint_as_float(c[k].x); }
};
template <> struct VT<double> {
    typedef double V;
    typedef double S;
    static __device__ __forceinline__ V mul(V a, V b) { V d; asm("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b)); return d; }
    static __device__ __forceinline__ void acc(V &acc_, V a, V b) { asm("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc_) : "d"(a), "d"(b)); }
    static __device__ __forceinline__ void upd(V &x, V a, V t) { asm("fma.rn.f64 %0, %1, %0, %2;" : "+d"(x) : "d"(a), "d"(t)); }
    static __device__ __forceinline__ void updd(V &x, V a, V t, V d1) { asm("fma.rn.f64 %0, %1, %0, %2; // %3" : "+d"(x) : "d"(a), "d"(t), "d"(d1)); }
    static __device__ __forceinline__ void updd(V &x, V a, V t, V d1, V d2, V d3) { asm("fma.rn.f64 %0, %1, %0, %2; // %3 %4 %5" : "+d"(x) : "d"(a), "d"(t), "d"(d1), "d"(d2), "d"(d3)); }
    static __device__ __forceinline__ V neg(V a) { return -a; }
    static __device__ __forceinline__ V bc(double s) { return s; }
    static __device__ __forceinline__ V swp(V a) { return a; }
    static __device__ __forceinline__ V coef(const uint2 *c, int k) { return reinterpret_cast<const double *>(c)[k]; }
    static __device__ __forceinline__ void coef2(const uint2 *c, int k, V &x, V &y) { const double2 u = reinterpret_cast<const double2 *>(c)[k >> 1]; x = u.x; y = u.y; }
    static __device__ __forceinline__ void tphase(const uint2 *c, double &pr, double &pi) { pr = coef(c, 0); pi = coef(c, 1); }
    static __device__ __forceinline__ double scalar(const uint2 *c, int k) { return coef(c, k); }
};

#define NV QSB_NV

/* (xr, xi) *= (pr, pi);  npi = -pi */
template <typename R>
__device__ __forceinline__ void cmul_inplace(typename VT<R>::V &xr, typename VT<R>::V &xi, typename VT<R>::V pr, typename VT<R>::V pi, typename VT<R>::V npi)
{
    typedef VT<R> T;
    const typename T::V t0 = T::mul(npi, xi), t1 = T::mul(pi, xr);
    T::updd(xr, pr, t0, t1);
    T::updd(xi, pr, t1, t0);
}

/* 2x2 on a vector bit.  FORM: 1 real, 2 real-diag/imag-offdiag, 3 general.
 * Cross terms go to temporaries first, then each amplitude is updated in place. */
template <typename R, int VB, int FORM>
__device__ __forceinline__ void mat_v(typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], const uint2 *c)
{
    typedef VT<R> T; typedef typename T::V V;
    if (FORM == 1) {
        V a, b, cc, d;                       /* payload order: m01 m10 | m00 m11 (cross terms first) */
        T::coef2(c, 0, b, cc); T::coef2(c, 2, a, d);
#pragma unroll
        for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
            const int w = v | (1 << VB);
            const V t0 = T::mul(b, re[w]), t1 = T::mul(b, im[w]), t2 = T::mul(cc, re[v]), t3 = T::mul(cc, im[v]);
            T::updd(re[v], a, t0, t2); T::updd(im[v], a, t1, t3);
            T::updd(re[w], d, t2, t0); T::updd(im[w], d, t3, t1);
        }
    } else if (FORM == 2) {
        /* [[a, i b],[i c, d]] (host pre-negates) */
        V a, nb, b, nc, cc, d;               /* payload order: -b b | -c c | a d */
        T::coef2(c, 0, nb, b); T::coef2(c, 2, nc, cc); T::coef2(c, 4, a, d);
#pragma unroll
        for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
            const int w = v | (1 << VB);
            const V t0 = T::mul(nb, im[w]), t1 = T::mul(b, re[w]), t2 = T::mul(nc, im[v]), t3 = T::mul(cc, re[v]);
            T::updd(re[v], a, t0, t3); T::updd(im[v], a, t1, t2);
            T::updd(re[w], d, t2, t1); T::updd(im[w], d, t3, t0);
        }
    } else {
        V ar, ai, br, bi, cr, ci, dr, di;    /* payload order: m00i m01r | m01i m10r | m10i m11i | m00r m11r */
        T::coef2(c, 0, ai, br); T::coef2(c, 2, bi, cr); T::coef2(c, 4, ci, di); T::coef2(c, 6, ar, dr);
        const V nai = T::neg(ai), nbi = T::neg(bi), nci = T::neg(ci), ndi = T::neg(di);
#pragma unroll
        for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
            const int w = v | (1 << VB);
            V t0 = T::mul(nai, im[v]); T::acc(t0, br, re[w]); T::acc(t0, nbi, im[w]);   /* re[v] minus its own-term */
            V t1 = T::mul(ai, re[v]);  T::acc(t1, br, im[w]); T::acc(t1, bi, re[w]);    /* im[v] */
            V t2 = T::mul(cr, re[v]);  T::acc(t2, nci, im[v]); T::acc(t2, ndi, im[w]);  /* re[w] */
            V t3 = T::mul(cr, im[v]);  T::acc(t3, ci, re[v]); T::acc(t3, di, re[w]);    /* im[w] */
            T::updd(re[v], ar, t0, t1, t2, t3); T::updd(im[v], ar, t1, t0, t2, t3);
            T::updd(re[w], dr, t2, t0, t1, t3); T::updd(im[w], dr, t3, t0, t1, t2);
        }
    }
}

/* unit form on a vector bit (see tiled.h): strictly in-place dependency chains.
 *   real: x0 += p*x1 ; x1 = k*x1 + q*x0        rx form: x0 += i p x1 ; x1 = k*x1 + i q x0 */
template <typename R, int VB, bool IMAG>
__device__ __forceinline__ void unit_v(typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], const uint2 *c)
{
    typedef VT<R> T; typedef typename T::V V;
    if (!IMAG) {
        V p, q, k, a_; T::coef2(c, 0, p, q); T::coef2(c, 2, k, a_);
#pragma unroll
        for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
            const int w = v | (1 << VB);
            T::acc(re[v], p, re[w]); T::acc(im[v], p, im[w]);
            const V t = T::mul(q, re[v]), u = T::mul(q, im[v]);
            T::upd(re[w], k, t); T::upd(im[w], k, u);
        }
    } else {
        V p, np, q, nq, k, a_; T::coef2(c, 0, p, np); T::coef2(c, 2, q, nq); T::coef2(c, 4, k, a_);
#pragma unroll
        for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
            const int w = v | (1 << VB);
            T::acc(re[v], np, im[w]); T::acc(im[v], p, re[w]);
            const V t = T::mul(nq, im[v]), u = T::mul(q, re[v]);
            T::upd(re[w], k, t); T::upd(im[w], k, u);
        }
    }
}
template <typename R, bool IMAG>
__device__ __forceinline__ void unit_dispatch(int vb, typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], const uint2 *c)
{
    switch (vb) {
    case 0: unit_v<R, 0, IMAG>(re, im, c); break;
    case 1: unit_v<R, 1, IMAG>(re, im, c); break;
#if QSB_NVB > 3
    case 2: unit_v<R, 2, IMAG>(re, im, c); break;
    default: unit_v<R, 3, IMAG>(re, im, c); break;
#else
    default: unit_v<R, 2, IMAG>(re, im, c); break;
#endif
    }
}

template <typename R, int FORM>
__device__ __forceinline__ void mat_dispatch(int vb, typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], const uint2 *c)
{
    switch (vb) {
    case 0: mat_v<R, 0, FORM>(re, im, c); break;
    case 1: mat_v<R, 1, FORM>(re, im, c); break;
#if QSB_NVB > 3
    case 2: mat_v<R, 2, FORM>(re, im, c); break;
    default: mat_v<R, 3, FORM>(re, im, c); break;
#else
    default: mat_v<R, 2, FORM>(re, im, c); break;
#endif
    }
}

/* phase on the vectors whose bit VB is set */
template <typename R, int VB>
__device__ __forceinline__ void diag_v(typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], typename VT<R>::V pr, typename VT<R>::V pi, typename VT<R>::V npi)
{
#pragma unroll
    for (int v = 0; v < NV; v++) if ((v >> VB) & 1) cmul_inplace<R>(re[v], im[v], pr, pi, npi);
}

/* 2x2 on the pack bit (f32 only): out = A * x + B * swap(x), A = (m00, m11), B = (m01, m10) */
template <int FORM>
__device__ __forceinline__ void mat_p(unsigned long long (&re)[NV], unsigned long long (&im)[NV], const uint2 *c)
{
    typedef VT<float> T; typedef T::V V;
    if (FORM == 1) {
        const V A = T::coef(c, 0), B = T::coef(c, 1);
#pragma unroll
        for (int v = 0; v < NV; v++) {
            const V t0 = T::mul(B, T::swp(re[v])), t1 = T::mul(B, T::swp(im[v]));
            T::upd(re[v], A, t0); T::upd(im[v], A, t1);
        }
    } else {
        const V Ar = T::coef(c, 0), Ai = T::coef(c, 1), Br = T::coef(c, 2), Bi = T::coef(c, 3);
        const V nAi = T::neg(Ai), nBi = T::neg(Bi);
#pragma unroll
        for (int v = 0; v < NV; v++) {
            const V sr = T::swp(re[v]), si = T::swp(im[v]);
            V t0 = T::mul(nAi, im[v]); T::acc(t0, Br, sr); T::acc(t0, nBi, si);
            V t1 = T::mul(Ai, re[v]);  T::acc(t1, Br, si); T::acc(t1, Bi, sr);
            T::updd(re[v], Ar, t0, t1); T::updd(im[v], Ar, t1, t0);
        }
    }
}
template <int FORM>
__device__ __forceinline__ void mat_p(double (&)[NV], double (&)[NV], const uint2 *) {}

/* ---------------------------------------------------------- global / shared IO */
template <typename R> struct IO;
template <> struct IO<float> {
    typedef unsigned long long V;
    /* amplitude pair unit: {re0, re1, im0, im1} at 16 * (index >> 1) */
    static __device__ __forceinline__ void gload(const void *base, uint64_t idx, V &re, V &im)
    {
        const ulonglong2 x = __ldcs(reinterpret_cast<const ulonglong2 *>(base) + (idx >> 1));
        re = x.x; im = x.y;
    }
    static __device__ __forceinline__ void gstore(void *base, uint64_t idx, V re, V im)
    {
        __stcs(reinterpret_cast<ulonglong2 *>(base) + (idx >> 1), make_ulonglong2(re, im));
    }
    /* two 32 KiB planes of 8-byte slots */
    static __device__ __forceinline__ void sload(const uint8_t *sm, uint32_t slot, V &re, V &im)
    {
        re = *reinterpret_cast<const V *>(sm + slot * 8u);
        im = *reinterpret_cast<const V *>(sm + 32768u + slot * 8u);
    }
    static __device__ __forceinline__ void sstore(uint8_t *sm, uint32_t slot, V re, V im)
    {
        *reinterpret_cast<V *>(sm + slot * 8u) = re;
        *reinterpret_cast<V *>(sm + 32768u + slot * 8u) = im;
    }
};
template <> struct IO<double> {
    typedef double V;
    static __device__ __forceinline__ void gload(const void *base, uint64_t idx, V &re, V &im)
    {
        const double2 x = __ldcs(reinterpret_cast<const double2 *>(base) + idx);
        re = x.x; im = x.y;
    }
    static __device__ __forceinline__ void gstore(void *base, uint64_t idx, V re, V im)
    {
        __stcs(reinterpret_cast<double2 *>(base) + idx, make_double2(re, im));
    }
    static __device__ __forceinline__ void sload(const uint8_t *sm, uint32_t slot, V &re, V &im)
    {
        const double2 x = *reinterpret_cast<const double2 *>(sm + slot * 16u);
        re = x.x; im = x.y;
    }
    static __device__ __forceinline__ void sstore(uint8_t *sm, uint32_t slot, V re, V im)
    {
        *reinterpret_cast<double2 *>(sm + slot * 16u) = make_double2(re, im);
    }
};

struct PtrTab { void *p[8]; };

/* OR / XOR of the per-vector-bit constants selected by the bits of v (v is a compile-time index) */
template <typename X> __device__ __forceinline__ X vcomb_or(int v, const X (&g)[QSB_NVB])
{
    X r = 0;
#pragma unroll
    for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) r |= g[b];
    return r;
}
template <typename X> __device__ __forceinline__ X vcomb_xor(int v, const X (&g)[QSB_NVB])
{
    X r = 0;
#pragma unroll
    for (int b = 0; b < QSB_NVB; b++) if ((v >> b) & 1) r ^= g[b];
    return r;
}

/* ------------------------------------------------------------------ the kernel
 * PEER: source amplitudes may live on other ranks (exchange passes): the index
 * bits above nloc select the peer buffer. */
template <typename R, int BLOB, bool PEER>
__global__ void __launch_bounds__(QSB_THREADS, 2)
k_tile_pass(const __grid_constant__ PassBlob<BLOB> blob, const __grid_constant__ PtrTab src, void *dst)
{
    typedef VT<R> T; typedef typename T::V V;
    extern __shared__ __align__(16) uint8_t smem[];
    const uint4 *B = blob.q;
    const DevPass &P = *reinterpret_cast<const DevPass *>(B);
    const DevRound *rounds = reinterpret_cast<const DevRound *>(B + P.rounds_off16);
    const uint32_t tid = threadIdx.x;

    /* tile id -> outer index bits */
    uint64_t tile = blockIdx.x, outer = 0;
    {
        const int nr = (int)P.n_runs;
        for (int r = 0; r < nr; r++) {
            const int len = P.run_len[r];
            outer |= (tile & ((1ULL << len) - 1)) << P.run_start[r];
            tile >>= len;
        }
    }
    const uint64_t src_outer = outer | P.src_fixed;
    const uint32_t nloc = P.nloc;
    const uint64_t loc_mask = (1ULL << nloc) - 1;
    const int n_rounds = (int)P.n_rounds;

    V re[NV], im[NV];

    for (int rd = 0; rd < n_rounds; rd++) {
        const DevRound &RD = rounds[rd];
        /* this thread's physical index bits (vector bits zero) and smem slot bases */
        uint64_t gthr = src_outer;
        uint32_t sb = 0; /* ld in the low half, st in the high half */
#pragma unroll
        for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) {
            gthr |= RD.thr[j].gidx;
            sb ^= (uint32_t)RD.thr[j].ld | ((uint32_t)RD.thr[j].st << 16);
        }

        if (rd == 0) {
            uint64_t gv[QSB_NVB];
#pragma unroll
            for (int b = 0; b < QSB_NVB; b++) gv[b] = RD.vec[b].gidx;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const uint64_t gi = gthr | vcomb_or(v, gv);
                if (PEER) IO<R>::gload(src.p[gi >> nloc], gi & loc_mask, re[v], im[v]);
                else IO<R>::gload(src.p[0], gi & loc_mask, re[v], im[v]);
            }
        } else {
            const uint32_t sl = sb & 0xffffu;
            uint32_t sv[QSB_NVB];
#pragma unroll
            for (int b = 0; b < QSB_NVB; b++) sv[b] = RD.vec[b].ld;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const uint32_t slot = sl ^ vcomb_xor(v, sv);
                IO<R>::sload(smem, slot, re[v], im[v]);
            }
            __syncthreads(); /* every thread has its registers before anyone overwrites the tile */
        }

        /* ---- the fused gates of this round ---- */
        R psr = R(1), psi = R(0); /* per-thread pending phase (OP_TPHASE) */
        const uint32_t n_ops = RD.n_ops;
        const uint4 *op = B + RD.op_off16;
        for (uint32_t i = 0; i < n_ops; i++) {
            const uint4 h = *op;
            const uint32_t kind = h.x;
            const uint64_t tmask = ((uint64_t)h.w << 32) | h.z;
            const uint2 *c = reinterpret_cast<const uint2 *>(op + 1);
            op += h.y;
            const bool pred = (gthr & tmask) == tmask;
            const bool mux = (kind >> 16) & 1u;
            if (!pred && !mux) continue;
            const uint32_t code = kind & 0xffu;
            const int vb = (kind >> 8) & 0xf;
            switch (code) {
            case OP_MAT_U: { const uint2 *cs = c + ((mux && pred) ? 4 : 0); unit_dispatch<R, false>(vb, re, im, cs); const R a = T::scalar(cs, 3); psr *= a; psi *= a; break; }
            case OP_MAT_UI: { const uint2 *cs = c + ((mux && pred) ? 6 : 0); unit_dispatch<R, true>(vb, re, im, cs); const R a = T::scalar(cs, 5); psr *= a; psi *= a; break; }
            case OP_MAT_R: mat_dispatch<R, 1>(vb, re, im, c + ((mux && pred) ? 4 : 0)); break;
            case OP_MAT_I: mat_dispatch<R, 2>(vb, re, im, c + ((mux && pred) ? 6 : 0)); break;
            case OP_MAT_G: mat_dispatch<R, 3>(vb, re, im, c + ((mux && pred) ? 8 : 0)); break;
            case OP_MATP_R: mat_p<1>(re, im, c + ((mux && pred) ? 2 : 0)); break;
            case OP_MATP_G: mat_p<3>(re, im, c + ((mux && pred) ? 4 : 0)); break;
            case OP_DIAG_V: {
                const V pr = T::coef(c, 0), pi = T::coef(c, 1), npi = T::neg(pi);
                switch (vb) {
                case 0: diag_v<R, 0>(re, im, pr, pi, npi); break;
                case 1: diag_v<R, 1>(re, im, pr, pi, npi); break;
#if QSB_NVB > 3
                case 2: diag_v<R, 2>(re, im, pr, pi, npi); break;
                default: diag_v<R, 3>(re, im, pr, pi, npi); break;
#else
                default: diag_v<R, 2>(re, im, pr, pi, npi); break;
#endif
                }
                break;
            }
            case OP_DIAG_ALL: {
                const V pr = T::coef(c, 0), pi = T::coef(c, 1), npi = T::neg(pi);
#pragma unroll
                for (int v = 0; v < NV; v++) cmul_inplace<R>(re[v], im[v], pr, pi, npi);
                break;
            }
            case OP_DIAG_GEN: {
                const uint32_t vmask = (kind >> 20) & 0xfu;
                const V pr = T::coef(c, 0), pi = T::coef(c, 1), npi = T::neg(pi);
#pragma unroll
                for (int v = 0; v < NV; v++) if ((v & vmask) == vmask) cmul_inplace<R>(re[v], im[v], pr, pi, npi);
                break;
            }
            case OP_TPHASE: {
                R pr, pi; T::tphase(c, pr, pi);
                const R nr = psr * pr - psi * pi;
                psi = psr * pi + psi * pr; psr = nr;
                break;
            }
            default: break;
            }
        }
        if (RD.flags & 1u) {
            if (!(psr == R(1) && psi == R(0))) {
                const V pr = T::bc(psr), pi = T::bc(psi), npi = T::bc(-psi);
#pragma unroll
                for (int v = 0; v < NV; v++) cmul_inplace<R>(re[v], im[v], pr, pi, npi);
            }
        }

        if (rd == n_rounds - 1) {
            uint64_t dthr = outer | P.dst_fixed;
#pragma unroll
            for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) dthr |= P.dst_thr[j];
            uint64_t gv[QSB_NVB];
#pragma unroll
            for (int b = 0; b < QSB_NVB; b++) gv[b] = P.dst_vec[b];
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const uint64_t gi = dthr | vcomb_or(v, gv);
                IO<R>::gstore(dst, gi & loc_mask, re[v], im[v]);
            }
        } else {
            const uint32_t ss = sb >> 16;
            uint32_t sv[QSB_NVB];
#pragma unroll
            for (int b = 0; b < QSB_NVB; b++) sv[b] = RD.vec[b].st;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const uint32_t slot = ss ^ vcomb_xor(v, sv);
                IO<R>::sstore(smem, slot, re[v], im[v]);
            }
            __syncthreads();
        }
    }
}

/* ------------------------------------------------------------------ launching */
template <typename R, int BLOB, bool PEER>
static int launch_one(qsb_sim *s, const HostPass &hp, const PtrTab &src, void *dst)
{
    static bool attr_set = false;
    if (!attr_set) {
        QSB_CUDA(cudaFuncSetAttribute(k_tile_pass<R, BLOB, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        attr_set = true;
    }
    if (hp.hdr.n_tiles > 0x7fffffffULL) { qsb_set_error("too many tiles"); return QSB_ERR_ARG; }
    const PassBlob<BLOB> *blob = reinterpret_cast<const PassBlob<BLOB> *>(hp.blob.data());
    k_tile_pass<R, BLOB, PEER><<<(unsigned)hp.hdr.n_tiles, QSB_THREADS, 65536, s->stream>>>(*blob, src, dst);
    QSB_CUDA(cudaGetLastError());
    return QSB_OK;
}

template <typename R>
static int launch_pass(qsb_sim *s, const HostPass &hp, const PtrTab &src, void *dst, bool peer)
{
    const bool small = hp.blob.size() <= QSB_BLOB_SMALL;
    if (peer) return small ? launch_one<R, QSB_BLOB_SMALL, true>(s, hp, src, dst) : launch_one<R, QSB_BLOB_LARGE, true>(s, hp, src, dst);
    return small ? launch_one<R, QSB_BLOB_SMALL, false>(s, hp, src, dst) : launch_one<R, QSB_BLOB_LARGE, false>(s, hp, src, dst);
}

int tiled_launch_pass(qsb_sim *s, const TiledPlan *p, size_t k, void *const *src_ptrs, void *dst, bool peer)
{
    PtrTab t;
    for (int i = 0; i < 8; i++) t.p[i] = src_ptrs[i];
    const HostPass &hp = p->passes[k];
    return s->prec == QSB_F32 ? launch_pass<float>(s, hp, t, dst, peer) : launch_pass<double>(s, hp, t, dst, peer);
}
