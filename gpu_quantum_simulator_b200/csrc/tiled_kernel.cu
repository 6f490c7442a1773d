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
 * The kernel is HBM-bound by design: algorithmic traffic per launch is
 * 2 * N_loc * sizeof(amplitude), independent of how many gates the pass fuses.
 *
 * Reference kernels replaced: kernel_gate / kernel_gate_2 (naive.cu:72-95),
 * kernel_cnot (naive.cu:97-122), kernel_gate_4 (4x4.cu:109-146).
 */
#include "sim.h"
#include "tiled.h"

/* ------------------------------------------------------------ vector algebra */
template <typename R> struct VT;
template <> struct VT<float> {
    typedef float2 V;
    static __device__ __forceinline__ V mul(V a, V b) { return __fmul2_rn(a, b); }
    static __device__ __forceinline__ V fma(V a, V b, V c) { return __ffma2_rn(a, b, c); }
    static __device__ __forceinline__ V neg(V a) { return make_float2(-a.x, -a.y); }
    static __device__ __forceinline__ V bc(float s) { return make_float2(s, s); }
    static __device__ __forceinline__ V swp(V a) { return make_float2(a.y, a.x); }
    /* coefficient vector k of a set: (lo lane, hi lane) */
    static __device__ __forceinline__ V coef(const float *c, int k) { return __ldg(reinterpret_cast<const float2 *>(c) + k); }
    static __device__ __forceinline__ void swap_lanes(V &a, V &b, uint32_t lanes)
    {
        if (lanes & 1) { float t = a.x; a.x = b.x; b.x = t; }
        if (lanes & 2) { float t = a.y; a.y = b.y; b.y = t; }
    }
};
template <> struct VT<double> {
    typedef double V;
    static __device__ __forceinline__ V mul(V a, V b) { return a * b; }
    static __device__ __forceinline__ V fma(V a, V b, V c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ V neg(V a) { return -a; }
    static __device__ __forceinline__ V bc(double s) { return s; }
    static __device__ __forceinline__ V swp(V a) { return a; }
    static __device__ __forceinline__ V coef(const double *c, int k) { return __ldg(c + k); }
    static __device__ __forceinline__ void swap_lanes(V &a, V &b, uint32_t) { V t = a; a = b; b = t; }
};

#define NV QSB_NV

/* 2x2 on a vector bit.  FORM: 1 real, 2 real-diag/imag-offdiag, 3 general. */
template <typename R, int VB, int FORM>
__device__ __forceinline__ void mat_v(typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], const R *c, uint32_t vmask)
{
    typedef VT<R> T; typedef typename T::V V;
    if (FORM == 1) {
        const V a = T::coef(c, 0), b = T::coef(c, 2), cc = T::coef(c, 4), d = T::coef(c, 6);
#pragma unroll
        for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
            const int w = v | (1 << VB);
            if ((v & vmask) == vmask) {
                V x0r = re[v], x0i = im[v], x1r = re[w], x1i = im[w];
                re[v] = T::fma(a, x0r, T::mul(b, x1r));
                im[v] = T::fma(a, x0i, T::mul(b, x1i));
                re[w] = T::fma(cc, x0r, T::mul(d, x1r));
                im[w] = T::fma(cc, x0i, T::mul(d, x1i));
            }
        }
    } else if (FORM == 2) {
        /* [[a, i b],[i c, d]]: slots 0:a 1:-b 2:b 3:-c 4:c 6:d (host pre-negates) */
        const V a = T::coef(c, 0), nb = T::coef(c, 1), b = T::coef(c, 2), nc = T::coef(c, 3), cc = T::coef(c, 4), d = T::coef(c, 6);
#pragma unroll
        for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
            const int w = v | (1 << VB);
            if ((v & vmask) == vmask) {
                V x0r = re[v], x0i = im[v], x1r = re[w], x1i = im[w];
                re[v] = T::fma(a, x0r, T::mul(nb, x1i));
                im[v] = T::fma(a, x0i, T::mul(b, x1r));
                re[w] = T::fma(nc, x0i, T::mul(d, x1r));
                im[w] = T::fma(cc, x0r, T::mul(d, x1i));
            }
        }
    } else {
        const V ar = T::coef(c, 0), ai = T::coef(c, 1), br = T::coef(c, 2), bi = T::coef(c, 3);
        const V cr = T::coef(c, 4), ci = T::coef(c, 5), dr = T::coef(c, 6), di = T::coef(c, 7);
        const V nai = T::neg(ai), nbi = T::neg(bi), nci = T::neg(ci), ndi = T::neg(di);
#pragma unroll
        for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
            const int w = v | (1 << VB);
            if ((v & vmask) == vmask) {
                V x0r = re[v], x0i = im[v], x1r = re[w], x1i = im[w];
                re[v] = T::fma(ar, x0r, T::fma(nai, x0i, T::fma(br, x1r, T::mul(nbi, x1i))));
                im[v] = T::fma(ar, x0i, T::fma(ai, x0r, T::fma(br, x1i, T::mul(bi, x1r))));
                re[w] = T::fma(cr, x0r, T::fma(nci, x0i, T::fma(dr, x1r, T::mul(ndi, x1i))));
                im[w] = T::fma(cr, x0i, T::fma(ci, x0r, T::fma(dr, x1i, T::mul(di, x1r))));
            }
        }
    }
}

template <typename R, int VB>
__device__ __forceinline__ void x_v(typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], uint32_t vmask, uint32_t lanes)
{
#pragma unroll
    for (int v = 0; v < NV; v++) if (!((v >> VB) & 1)) {
        const int w = v | (1 << VB);
        if ((v & vmask) == vmask) {
            VT<R>::swap_lanes(re[v], re[w], lanes);
            VT<R>::swap_lanes(im[v], im[w], lanes);
        }
    }
}

template <typename R, int FORM>
__device__ __forceinline__ void mat_dispatch(int vb, typename VT<R>::V (&re)[NV], typename VT<R>::V (&im)[NV], const R *c, uint32_t vmask)
{
    switch (vb) {
    case 0: mat_v<R, 0, FORM>(re, im, c, vmask); break;
    case 1: mat_v<R, 1, FORM>(re, im, c, vmask); break;
    case 2: mat_v<R, 2, FORM>(re, im, c, vmask); break;
    default: mat_v<R, 3, FORM>(re, im, c, vmask); break;
    }
}

/* 2x2 on the pack bit (f32 only): out = A * x + B * swap(x), A = (m00, m11), B = (m01, m10) */
template <int FORM>
__device__ __forceinline__ void mat_p(float2 (&re)[NV], float2 (&im)[NV], const float *c, uint32_t vmask)
{
    typedef VT<float> T;
    if (FORM == 1) {
        const float2 A = T::coef(c, 0), B = T::coef(c, 2);
#pragma unroll
        for (int v = 0; v < NV; v++) if ((v & vmask) == vmask) {
            re[v] = T::fma(A, re[v], T::mul(B, T::swp(re[v])));
            im[v] = T::fma(A, im[v], T::mul(B, T::swp(im[v])));
        }
    } else {
        const float2 Ar = T::coef(c, 0), Ai = T::coef(c, 1), Br = T::coef(c, 2), Bi = T::coef(c, 3);
        const float2 nAi = T::neg(Ai), nBi = T::neg(Bi);
#pragma unroll
        for (int v = 0; v < NV; v++) if ((v & vmask) == vmask) {
            float2 xr = re[v], xi = im[v], sr = T::swp(xr), si = T::swp(xi);
            re[v] = T::fma(Ar, xr, T::fma(nAi, xi, T::fma(Br, sr, T::mul(nBi, si))));
            im[v] = T::fma(Ar, xi, T::fma(Ai, xr, T::fma(Br, si, T::mul(Bi, sr))));
        }
    }
}
template <int FORM>
__device__ __forceinline__ void mat_p(double (&)[NV], double (&)[NV], const double *, uint32_t) {}

__device__ __forceinline__ void xp(float2 (&re)[NV], float2 (&im)[NV], uint32_t vmask)
{
#pragma unroll
    for (int v = 0; v < NV; v++) if ((v & vmask) == vmask) { re[v] = VT<float>::swp(re[v]); im[v] = VT<float>::swp(im[v]); }
}
__device__ __forceinline__ void xp(double (&)[NV], double (&)[NV], uint32_t) {}

/* ---------------------------------------------------------- global / shared IO */
struct PtrTab { void *p[8]; };

template <typename R> struct IO;
template <> struct IO<float> {
    typedef float2 V;
    /* amplitude pair unit: {re0, re1, im0, im1} at 16 * (index >> 1) */
    static __device__ __forceinline__ void gload(const void *base, uint64_t idx, V &re, V &im)
    {
        const float4 x = __ldcs(reinterpret_cast<const float4 *>(base) + (idx >> 1));
        re = make_float2(x.x, x.y); im = make_float2(x.z, x.w);
    }
    static __device__ __forceinline__ void gstore(void *base, uint64_t idx, V re, V im)
    {
        __stcs(reinterpret_cast<float4 *>(base) + (idx >> 1), make_float4(re.x, re.y, im.x, im.y));
    }
    /* two 32 KiB planes of 8-byte slots */
    static __device__ __forceinline__ void sload(const uint8_t *sm, uint32_t slot, V &re, V &im)
    {
        re = *reinterpret_cast<const float2 *>(sm + slot * 8u);
        im = *reinterpret_cast<const float2 *>(sm + 32768u + slot * 8u);
    }
    static __device__ __forceinline__ void sstore(uint8_t *sm, uint32_t slot, V re, V im)
    {
        *reinterpret_cast<float2 *>(sm + slot * 8u) = re;
        *reinterpret_cast<float2 *>(sm + 32768u + slot * 8u) = im;
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

/* ------------------------------------------------------------------ the kernel */
template <typename R>
__global__ void __launch_bounds__(QSB_THREADS, 2)
k_tile_pass(const DevPass *__restrict__ pass_p, const DevRound *__restrict__ rounds,
            const DevOp<R> *__restrict__ ops, PtrTab src, void *dst)
{
    typedef VT<R> T; typedef typename T::V V;
    extern __shared__ __align__(16) uint8_t smem[];
    const DevPass &P = *pass_p;
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
        /* this thread's physical index bits (vector bits zero) */
        uint64_t gthr = src_outer;
#pragma unroll
        for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) gthr |= RD.thr_gidx[j];

        if (rd == 0) {
            const uint64_t g0 = RD.vec_gidx[0], g1 = RD.vec_gidx[1], g2 = RD.vec_gidx[2], g3 = RD.vec_gidx[3];
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const uint64_t gi = gthr | ((v & 1) ? g0 : 0) | ((v & 2) ? g1 : 0) | ((v & 4) ? g2 : 0) | ((v & 8) ? g3 : 0);
                IO<R>::gload(src.p[gi >> nloc], gi & loc_mask, re[v], im[v]);
            }
        } else {
            uint32_t sb = 0;
#pragma unroll
            for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) sb ^= RD.ld_thr[j];
            const uint32_t s0 = RD.ld_vec[0], s1 = RD.ld_vec[1], s2 = RD.ld_vec[2], s3 = RD.ld_vec[3];
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const uint32_t slot = sb ^ ((v & 1) ? s0 : 0) ^ ((v & 2) ? s1 : 0) ^ ((v & 4) ? s2 : 0) ^ ((v & 8) ? s3 : 0);
                IO<R>::sload(smem, slot, re[v], im[v]);
            }
            __syncthreads(); /* every thread has its registers before anyone overwrites the tile */
        }

        /* ---- the fused gates of this round ---- */
        R psr = R(1), psi = R(0); /* per-thread pending phase (OP_TPHASE) */
        const uint32_t n_ops = RD.n_ops;
        const DevOp<R> *op = ops + RD.op_begin;
        for (uint32_t i = 0; i < n_ops; i++, op++) {
            const uint32_t kind = op->kind;
            const uint32_t code = kind & 0xffu;
            const uint64_t tmask = op->tmask;
            const bool pred = (gthr & tmask) == tmask;
            const bool mux = (kind >> 16) & 1u;
            if (!pred && !mux) continue;
            const uint32_t vmask = op->vmask;
            const R *c = op->c + (pred ? (64 / sizeof(R)) : 0); /* coefficient set 1 = condition holds */
            const int vb = (kind >> 8) & 0xf;
            switch (code) {
            case OP_MAT_R: mat_dispatch<R, 1>(vb, re, im, c, vmask); break;
            case OP_MAT_I: mat_dispatch<R, 2>(vb, re, im, c, vmask); break;
            case OP_MAT_G: mat_dispatch<R, 3>(vb, re, im, c, vmask); break;
            case OP_MATP_R: mat_p<1>(re, im, c, vmask); break;
            case OP_MATP_G: mat_p<3>(re, im, c, vmask); break;
            case OP_X: {
                const uint32_t lanes = (kind >> 12) & 3u;
                switch (vb) {
                case 0: x_v<R, 0>(re, im, vmask, lanes); break;
                case 1: x_v<R, 1>(re, im, vmask, lanes); break;
                case 2: x_v<R, 2>(re, im, vmask, lanes); break;
                default: x_v<R, 3>(re, im, vmask, lanes); break;
                }
                break;
            }
            case OP_XP: xp(re, im, vmask); break;
            case OP_DIAG: {
                const V pr = T::coef(c, 0), pi = T::coef(c, 1), npi = T::neg(pi);
#pragma unroll
                for (int v = 0; v < NV; v++) if ((v & vmask) == vmask) {
                    const V xr = re[v], xi = im[v];
                    re[v] = T::fma(pr, xr, T::mul(npi, xi));
                    im[v] = T::fma(pr, xi, T::mul(pi, xr));
                }
                break;
            }
            case OP_TPHASE: {
                const R pr = __ldg(c), pi = __ldg(c + (sizeof(R) == 4 ? 2 : 1)); /* coef 0 / coef 1, lo lane */
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
                for (int v = 0; v < NV; v++) {
                    const V xr = re[v], xi = im[v];
                    re[v] = T::fma(pr, xr, T::mul(npi, xi));
                    im[v] = T::fma(pr, xi, T::mul(pi, xr));
                }
            }
        }

        if (rd == n_rounds - 1) {
            uint64_t dthr = outer | P.dst_fixed;
#pragma unroll
            for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) dthr |= P.dst_thr[j];
            const uint64_t g0 = P.dst_vec[0], g1 = P.dst_vec[1], g2 = P.dst_vec[2], g3 = P.dst_vec[3];
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const uint64_t gi = dthr | ((v & 1) ? g0 : 0) | ((v & 2) ? g1 : 0) | ((v & 4) ? g2 : 0) | ((v & 8) ? g3 : 0);
                IO<R>::gstore(dst, gi & loc_mask, re[v], im[v]);
            }
        } else {
            uint32_t sb = 0;
#pragma unroll
            for (int j = 0; j < QSB_TB; j++) if ((tid >> j) & 1) sb ^= RD.st_thr[j];
            const uint32_t s0 = RD.st_vec[0], s1 = RD.st_vec[1], s2 = RD.st_vec[2], s3 = RD.st_vec[3];
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const uint32_t slot = sb ^ ((v & 1) ? s0 : 0) ^ ((v & 2) ? s1 : 0) ^ ((v & 4) ? s2 : 0) ^ ((v & 8) ? s3 : 0);
                IO<R>::sstore(smem, slot, re[v], im[v]);
            }
            __syncthreads();
        }
    }
}

/* ------------------------------------------------------------------ launching */
static bool g_attr_set[2] = {false, false};

template <typename R>
static int launch_pass(qsb_sim *s, const TiledPlan *p, size_t k, const PtrTab &src, void *dst)
{
    const uint8_t *blob = (const uint8_t *)p->d_blob;
    const HostPass &hp = p->passes[k];
    const int which = sizeof(R) == 4 ? 0 : 1;
    if (!g_attr_set[which]) {
        QSB_CUDA(cudaFuncSetAttribute(k_tile_pass<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        g_attr_set[which] = true;
    }
    if (hp.hdr.n_tiles > 0x7fffffffULL) { qsb_set_error("too many tiles"); return QSB_ERR_ARG; }
    k_tile_pass<R><<<(unsigned)hp.hdr.n_tiles, QSB_THREADS, 65536, s->stream>>>(
        (const DevPass *)(blob + p->pass_off[k]), (const DevRound *)(blob + p->round_off[k]),
        (const DevOp<R> *)(blob + p->op_off[k]), src, dst);
    QSB_CUDA(cudaGetLastError());
    return QSB_OK;
}

int tiled_launch_pass(qsb_sim *s, const TiledPlan *p, size_t k, void *const *src_ptrs, void *dst)
{
    PtrTab t;
    for (int i = 0; i < 8; i++) t.p[i] = src_ptrs[i];
    return s->prec == QSB_F32 ? launch_pass<float>(s, p, k, t, dst) : launch_pass<double>(s, p, k, t, dst);
}
