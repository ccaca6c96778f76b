// In-shared-memory 1-D complex FFT engine (fp32), radix 4/8/16 passes in
// registers, used by every transform on the path (voxelgrids.py:388).
//
// The code is written against an explicit (tid, nthreads) pair instead of
// threadIdx so that the very same templates compile for the host
// (tests/host_emul) where the pass/permutation/twiddle logic is unit-tested
// against numpy.fft without a GPU.
//
// Transform of length M = 2^L, in place, decimation in frequency:
//   pass p has radix R_p and stride S_p = M / (R_0 ... R_p); butterfly b works
//   on elements base + S_p*n, multiplies output k by W_{S_p R_p}^{t k} and
//   stores it at base + S_p*k.  Natural-order input, digit-reversed output:
//   X[k] with k = k_0 + R_0 k_1 + R_0 R_1 k_2 ... ends at S_0 k_0 + S_1 k_1 ...
//   (gx_fft_pos).  Shared-memory addresses go through gx_phys() (one pad word
//   every 16 and every 256 elements) which keeps every pass and the permuted
//   read-out free of bank conflicts for the 16^k schedules.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GX_HD __host__ __device__ __forceinline__
#else
#define GX_HD inline
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
#endif

// ---- radix schedule per log2(M) ------------------------------------------
template <int L> struct GxSched;
#define GX_SCHED(L_, NP_, A_, B_, C_, D_)                                        \
    template <> struct GxSched<L_> {                                            \
        enum { NP = NP_, R0 = A_, R1 = B_, R2 = C_, R3 = D_ };                  \
    }
GX_SCHED(4, 1, 16, 1, 1, 1);
GX_SCHED(5, 2, 8, 4, 1, 1);
GX_SCHED(6, 2, 8, 8, 1, 1);
GX_SCHED(7, 2, 16, 8, 1, 1);
GX_SCHED(8, 2, 16, 16, 1, 1);
GX_SCHED(9, 3, 8, 8, 8, 1);
GX_SCHED(10, 3, 16, 16, 4, 1);
GX_SCHED(11, 3, 16, 16, 8, 1);
GX_SCHED(12, 3, 16, 16, 16, 1);
GX_SCHED(13, 4, 8, 16, 16, 4);
GX_SCHED(14, 4, 16, 16, 16, 4);
#undef GX_SCHED

// runtime view of the same table (host plan builder, tests)
static inline int gx_sched_radices(int L, int r[4])
{
    switch (L) {
#define GX_CASE(L_) case L_: r[0] = GxSched<L_>::R0; r[1] = GxSched<L_>::R1; \
                             r[2] = GxSched<L_>::R2; r[3] = GxSched<L_>::R3; return GxSched<L_>::NP;
        GX_CASE(4) GX_CASE(5) GX_CASE(6) GX_CASE(7) GX_CASE(8)
        GX_CASE(9) GX_CASE(10) GX_CASE(11) GX_CASE(12) GX_CASE(13) GX_CASE(14)
#undef GX_CASE
    default: return 0;
    }
}

// ---- plan table layout (float2 units) --------------------------------------
// [ twiddles pass 0 | pass 1 | ... ][ chirp N ][ bhat M ]   (last two: Bluestein)
struct GxFftLayout {
    int N, M, L, bluestein;
    int tw_off[4];
    int chirp_off, bhat_off, total;
};

static inline GxFftLayout gx_fft_layout(int N)
{
    GxFftLayout g;
    g.N = N; g.M = 0; g.L = 0; g.bluestein = 0; g.total = 0;
    g.chirp_off = g.bhat_off = 0;
    for (int i = 0; i < 4; ++i) g.tw_off[i] = 0;
    if (N < 16) return g;
    int pow2 = (N & (N - 1)) == 0;
    int M = N;
    if (!pow2) { M = 1; while (M < 2 * N - 1) M <<= 1; g.bluestein = 1; }
    int L = 0; while ((1 << L) < M) ++L;
    if (L < 4 || L > 14) return g;
    g.M = M; g.L = L;
    int r[4]; int np = gx_sched_radices(L, r);
    int off = 0, S = M;
    for (int p = 0; p < np; ++p) {
        S /= r[p];
        g.tw_off[p] = off;
        if (S > 1) off += (r[p] - 1) * S;
    }
    if (g.bluestein) { g.chirp_off = off; off += N; g.bhat_off = off; off += M; }
    g.total = off;
    return g;
}

// ---- addressing -----------------------------------------------------------
GX_HD int gx_phys(int i) { return i + (i >> 4) + (i >> 8); }
GX_HD int gx_phys_len(int M) { return M + (M >> 4) + (M >> 8) + 1; }

template <int L> GX_HD int gx_fft_pos(int k)
{
    typedef GxSched<L> S;
    constexpr int M = 1 << L;
    int p = (k % S::R0) * (M / S::R0);
    if (S::NP > 1) { k /= S::R0; p += (k % S::R1) * (M / S::R0 / S::R1); }
    if (S::NP > 2) { k /= S::R1; p += (k % S::R2) * (M / S::R0 / S::R1 / S::R2); }
    if (S::NP > 3) { k /= S::R2; p += (k % S::R3) * (M / S::R0 / S::R1 / S::R2 / S::R3); }
    return p;
}

// ---- complex helpers -------------------------------------------------------
// sm_100a has packed fp32x2 arithmetic (add/sub/fma.rn.f32x2 -> FADD2 / FFMA2: one instruction, one
// issue slot, both halves of an aligned register pair).  A complex add is exactly that shape, and the
// slice kernels are bound by instruction issue, so the butterflies' additions are issued packed; nvcc
// does not form these from scalar code on its own.  Results are bit-identical to the scalar forms
// (same round-to-nearest operations).  The host build (tests/host_emul) uses the scalar forms.
#if defined(__CUDA_ARCH__) && !defined(GX_NO_F32X2)
#define GX_F32X2_OP(name, op)                                                                             \
    __device__ __forceinline__ float2 name(float2 a, float2 b)                                            \
    {                                                                                                     \
        float2 r;                                                                                         \
        asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; " op " rc, ra, rb; "       \
            "mov.b64 {%0,%1}, rc; }"                                                                      \
            : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                             \
        return r;                                                                                         \
    }
GX_F32X2_OP(gx_cadd, "add.rn.f32x2")
GX_F32X2_OP(gx_csub, "sub.rn.f32x2")
#undef GX_F32X2_OP
// acc + c * d with a real coefficient c (both components)
__device__ __forceinline__ float2 gx_caxpy(float c, float2 d, float2 acc)
{
    float2 r;
    asm("{ .reg .b64 rc, rd, ra, rr; mov.b64 rc, {%2,%2}; mov.b64 rd, {%3,%4}; mov.b64 ra, {%5,%6}; "
        "fma.rn.f32x2 rr, rc, rd, ra; mov.b64 {%0,%1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(c), "f"(d.x), "f"(d.y), "f"(acc.x), "f"(acc.y));
    return r;
}
#else
GX_HD float2 gx_cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
GX_HD float2 gx_csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
GX_HD float2 gx_caxpy(float c, float2 d, float2 acc) { return make_float2(acc.x + c * d.x, acc.y + c * d.y); }
#endif
GX_HD float2 gx_cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
GX_HD float2 gx_mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
GX_HD float2 gx_conj(float2 a) { return make_float2(a.x, -a.y); }

// forward DFTs (e^{-2 pi i nk/R}), in place, natural order in and out
GX_HD void gx_dft2(float2 &a, float2 &b)
{
    float2 t = a; a = gx_cadd(t, b); b = gx_csub(t, b);
}
GX_HD void gx_dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3)
{
    const float2 a0 = gx_cadd(v0, v2), a1 = gx_csub(v0, v2);
    const float2 a2 = gx_cadd(v1, v3), d = gx_csub(v1, v3);
    v0 = gx_cadd(a0, a2);
    v2 = gx_csub(a0, a2);
    // a1 -+ i d, component by component: the swapped operand would cost moves in a packed add
    v1 = make_float2(a1.x + d.y, a1.y - d.x);
    v3 = make_float2(a1.x - d.y, a1.y + d.x);
}

template <int R> struct GxDft;
template <> struct GxDft<2> { static GX_HD void run(float2 *v) { gx_dft2(v[0], v[1]); } };
template <> struct GxDft<4> { static GX_HD void run(float2 *v) { gx_dft4(v[0], v[1], v[2], v[3]); } };
template <> struct GxDft<8> {
    static GX_HD void run(float2 *v)
    {
        // n = 2a + b : DFT4 over a for b = 0,1 ; twiddle W8^{b k1} ; DFT2 over b
        const float h = 0.70710678118654752440f;
        gx_dft4(v[0], v[2], v[4], v[6]);   // b = 0 -> y0[k1] in v[0],v[2],v[4],v[6]
        gx_dft4(v[1], v[3], v[5], v[7]);   // b = 1 -> y1[k1] in v[1],v[3],v[5],v[7]
        // y1[k1] *= W8^{k1}
        v[3] = make_float2((v[3].x + v[3].y) * h, (v[3].y - v[3].x) * h);     // W8^1 = (h,-h)
        v[5] = gx_mul_mi(v[5]);                                               // W8^2 = -i
        v[7] = make_float2((v[7].y - v[7].x) * h, -(v[7].x + v[7].y) * h);    // W8^3 = (-h,-h)
        // X[k1 + 4 k2] = y0[k1] + (-1)^{k2} y1[k1]
        float2 o[8];
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            o[k1] = gx_cadd(v[2 * k1], v[2 * k1 + 1]);
            o[k1 + 4] = gx_csub(v[2 * k1], v[2 * k1 + 1]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = o[k];
    }
};
template <> struct GxDft<16> {
    static GX_HD void run(float2 *v)
    {
        // n = 4a + b : DFT4 over a ; twiddle W16^{b k1} ; DFT4 over b ; k = k1 + 4 k2
        const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
        const float h = 0.70710678118654752440f;
#pragma unroll
        for (int b = 0; b < 4; ++b) gx_dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
        // after this, y[b][k1] lives in v[4*k1 + b]
        // W16^m = (cos(pi m/8), -sin(pi m/8))
        v[4 * 1 + 1] = gx_cmul(v[4 * 1 + 1], make_float2(c1, -s1));    // m = 1
        v[4 * 1 + 2] = gx_cmul(v[4 * 1 + 2], make_float2(h, -h));      // m = 2
        v[4 * 1 + 3] = gx_cmul(v[4 * 1 + 3], make_float2(s1, -c1));    // m = 3
        v[4 * 2 + 1] = gx_cmul(v[4 * 2 + 1], make_float2(h, -h));      // m = 2
        v[4 * 2 + 2] = gx_mul_mi(v[4 * 2 + 2]);                        // m = 4
        v[4 * 2 + 3] = gx_cmul(v[4 * 2 + 3], make_float2(-h, -h));     // m = 6
        v[4 * 3 + 1] = gx_cmul(v[4 * 3 + 1], make_float2(s1, -c1));    // m = 3
        v[4 * 3 + 2] = gx_cmul(v[4 * 3 + 2], make_float2(-h, -h));     // m = 6
        v[4 * 3 + 3] = gx_cmul(v[4 * 3 + 3], make_float2(-c1, s1));    // m = 9
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1)
            gx_dft4(v[4 * k1 + 0], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
        // now X[k1 + 4 k2] lives in v[4*k1 + k2]: transpose the 4x4 register tile
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a + 1; b < 4; ++b) { float2 t = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = t; }
    }
};

// ---- band-limited last pass ------------------------------------------------
// The last DIF pass (S = 1) of a 16-16-16 schedule decides the MOST significant digit k2 of the
// output index k = k0 + 16 k1 + 256 k2.  A caller that keeps only the coefficients |k| < 512
// around DC (k2 in {0, 1} or {14, 15}; the slice kernels keep ~ +-290 of 4096) needs at most
// X[0], X[1], X[14], X[15] of each 16-point butterfly, and for |k| < 256 only X[0] and X[15].
// With p_n = v[n] + v[16-n], m_n = v[n] - v[16-n] (n = 1..7) and W = e^{-2 pi i/16}:
//   X[0]  = v0 + v8 + sum p_n
//   X[+-1] = A -+ iB,   A = (v0 - v8) + c1 (p1 - p7) + h (p2 - p6) + s1 (p3 - p5)
//                       B = s1 (m1 + m7) + h (m2 + m6) + c1 (m3 + m5) + m4
//   X[+-2] = A2 -+ iB2, A2 = (v0 + v8 - p4) + h ((p1 + p7) - (p3 + p5))
//                       B2 = (m2 - m6) + h ((m1 + m3) - (m5 + m7))
// 38 complex additions + 7 real-by-complex multiply-adds for X[0], X[15] (76 flops instead of the
// ~170 of the full radix-16 butterfly, 2 stores instead of 16); `wide` adds X[1] and X[14].
// Results go to the slots the full pass would use (base + k2); the other slots keep stale data.
GX_HD void gx_dft16_lowband_vals(const float2 *v, bool wide, float2 &x0, float2 &x15, float2 &x1, float2 &x14)
{
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
    const float h = 0.70710678118654752440f;
    float2 p[8], m[8];
#pragma unroll
    for (int n = 1; n < 8; ++n) { p[n] = gx_cadd(v[n], v[16 - n]); m[n] = gx_csub(v[n], v[16 - n]); }
    const float2 e = gx_cadd(v[0], v[8]), o = gx_csub(v[0], v[8]);
    x0 = gx_cadd(gx_cadd(gx_cadd(e, p[4]), gx_cadd(p[1], p[7])),
                 gx_cadd(gx_cadd(p[2], p[6]), gx_cadd(p[3], p[5])));
    const float2 d17 = gx_csub(p[1], p[7]), d26 = gx_csub(p[2], p[6]), d35 = gx_csub(p[3], p[5]);
    const float2 A = gx_caxpy(s1, d35, gx_caxpy(h, d26, gx_caxpy(c1, d17, o)));
    const float2 a17 = gx_cadd(m[1], m[7]), a26 = gx_cadd(m[2], m[6]), a35 = gx_cadd(m[3], m[5]);
    const float2 B = gx_caxpy(c1, a35, gx_caxpy(h, a26, gx_caxpy(s1, a17, m[4])));
    x15 = make_float2(A.x - B.y, A.y + B.x);                        // A + iB
    x1 = x14 = make_float2(0.f, 0.f);
    if (wide) {
        x1 = make_float2(A.x + B.y, A.y - B.x);                     // A - iB
        const float2 s17 = gx_cadd(p[1], p[7]), s35 = gx_cadd(p[3], p[5]);
        const float2 t = gx_csub(s17, s35), u = gx_csub(e, p[4]);
        const float2 A2 = gx_caxpy(h, t, u);
        const float2 w = gx_csub(gx_cadd(m[1], m[3]), gx_cadd(m[5], m[7])), g = gx_csub(m[2], m[6]);
        const float2 B2 = gx_caxpy(h, w, g);
        x14 = make_float2(A2.x - B2.y, A2.y + B2.x);                // A2 + iB2
    }
}

GX_HD void gx_dft16_lowband(const float2 *v, bool wide, float2 *sb)
{
    float2 x0, x15, x1, x14;
    gx_dft16_lowband_vals(v, wide, x0, x15, x1, x14);
    sb[gx_phys(0)] = x0;
    sb[gx_phys(15)] = x15;
    if (wide) {
        sb[gx_phys(1)] = x1;
        sb[gx_phys(14)] = x14;
    }
}

// Last pass of the 16-16-16 transform for a caller that reads only coefficients
// k in [0, khi) and [M + klo, M) with klo >= -512, khi <= 512 (checked by the caller).
// Thread t owns butterfly (k0, k1) = (t & 15, t >> 4): the few butterflies that need X[1]
// (k0 + 16 k1 + 256 < khi) or X[14] (k0 + 16 k1 + 3584 >= M + klo) sit in the first / last warps.
template <int M, int NBUF = 1, int BUFSTRIDE = 0>
GX_HD void gx_fft_lastpass16_lowband(float2 *s, int klo, int khi, int tid, int nthreads)
{
    constexpr int NBFLY = M / 16;
    for (int w = tid; w < NBUF * NBFLY; w += nthreads) {
        const int buf = w / NBFLY, b = w - buf * NBFLY;
        const int k0 = b & 15, k1 = b >> 4;
        const int low = k0 + 16 * k1;
        const int blk = 16 * k0 + k1;                      // butterfly index of the generic pass (base = 16 blk)
        float2 *sb = s + buf * BUFSTRIDE + gx_phys(16 * blk);
        float2 v[16];
#pragma unroll
        for (int n = 0; n < 16; ++n) v[n] = sb[gx_phys(n)];
        const bool wide = (low + 256 < khi) || (low + 3584 >= M + klo);
        gx_dft16_lowband(v, wide, sb);
    }
}

// ---- twiddles of one butterfly ---------------------------------------------
// v[k] *= W^{t k}, k = 1..R-1, table row k-1 at twt[(k-1)*S].
// TWP == 0: every factor is read from the table (fp32-rounded exact values).
// TWP == 1: only W^t, W^2t, W^4t, W^8t are read; the others are products of at
//           most three of them.  4 loads instead of 15 per radix-16 butterfly --
//           the kernels are bound by the L1/shared-memory pipe, not by FMA -- at
//           the price of twiddles accurate to ~2e-7 instead of 6e-8.
#ifndef GX_TWP
#define GX_TWP 0
#endif
template <int R, int S, int TWP>
GX_HD void gx_apply_twiddles(float2 *v, const float2 *twt)
{
    if (TWP == 0 || R <= 4) {
#pragma unroll
        for (int k = 1; k < R; ++k) v[k] = gx_cmul(v[k], twt[(k - 1) * S]);
    } else {
        const float2 w1 = twt[0], w2 = twt[S], w4 = twt[3 * S];
        const float2 w3 = gx_cmul(w1, w2), w5 = gx_cmul(w1, w4), w6 = gx_cmul(w2, w4), w7 = gx_cmul(w3, w4);
        v[1] = gx_cmul(v[1], w1); v[2] = gx_cmul(v[2], w2); v[3] = gx_cmul(v[3], w3); v[4] = gx_cmul(v[4], w4);
        v[5] = gx_cmul(v[5], w5); v[6] = gx_cmul(v[6], w6); v[7] = gx_cmul(v[7], w7);
        if (R > 8) {
            const float2 w8 = twt[7 * S];
            v[8 % R] = gx_cmul(v[8 % R], w8);
            v[9 % R] = gx_cmul(v[9 % R], gx_cmul(w1, w8));
            v[10 % R] = gx_cmul(v[10 % R], gx_cmul(w2, w8));
            v[11 % R] = gx_cmul(v[11 % R], gx_cmul(w3, w8));
            v[12 % R] = gx_cmul(v[12 % R], gx_cmul(w4, w8));
            v[13 % R] = gx_cmul(v[13 % R], gx_cmul(w5, w8));
            v[14 % R] = gx_cmul(v[14 % R], gx_cmul(w6, w8));
            v[15 % R] = gx_cmul(v[15 % R], gx_cmul(w7, w8));
        }
    }
}

// ---- N = 16 boxes x (N / 16): the split transform of the TMA-fed column kernel ------------
// N = 2^L with a 16-16-R2 schedule (L = 10, 11, 12: R2 = 4, 8, 16), RB = N / 16 = 16 R2 rows per box.
// z = 16 u + c: sample z lives in box c = z mod 16 at row u = z / 16 (slot RB c + u of the work buffer).
//   alpha + beta : Y_c[k'] = sum_u x[16 u + c] W_RB^{u k'}    radix-16 (stride R2) and radix-R2 DIF passes
//                  INSIDE one box; Y_c[k_a + 16 k_b] ends at padded slot RB c + R2 k_a + k_b of the column buffer
//   gamma        : X[k' + RB m] = sum_c (W_N^{c k'} Y_c[k']) W_16^{c m}   radix-16 DIT ACROSS the boxes
// The twiddle tables are those of passes 1 (W_RB^{t k}) and 0 (W_N^{t k}) of the N-point schedule.
// The kernel and the host emulation (tests/host_emul) share these functions.
template <int L> struct GxSplit {
    enum { N = 1 << L, RB = N / 16, R2 = GxSched<L>::R2 };
    static_assert(GxSched<L>::R0 == 16 && GxSched<L>::R1 == 16 && GxSched<L>::NP == 3 && RB == 16 * R2,
                  "split transform needs a 16-16-R2 schedule");
    static GX_HD int slot(int z) { return (z & 15) * RB + (z >> 4); }
};

// alpha for butterfly t (0..R2-1) of box c: src[n * src_stride] = row t + R2 n of the dense box,
// sb = column buffer + gx_phys(RB c + t)
template <int L, int TWP>
GX_HD void gx_split_alpha(const float2 *src, int src_stride, float2 *sb, const float2 *tw1, int t)
{
    constexpr int R2 = GxSplit<L>::R2;
    float2 v[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) v[n] = src[n * src_stride];
    GxDft<16>::run(v);
    gx_apply_twiddles<16, R2, TWP>(v, tw1 + t);
#pragma unroll
    for (int k = 0; k < 16; ++k) sb[gx_phys(R2 * k)] = v[k];
}

// beta for block blk (0..15) of box c, in place: sb = column buffer + gx_phys(RB c + R2 blk)
template <int L>
GX_HD void gx_split_beta(float2 *sb)
{
    constexpr int R2 = GxSplit<L>::R2;
    float2 v[R2];
#pragma unroll
    for (int n = 0; n < R2; ++n) v[n] = sb[n];
    GxDft<R2>::run(v);
#pragma unroll
    for (int k = 0; k < R2; ++k) sb[k] = v[k];
}

// inputs of the gamma butterfly of k' (0..RB-1), twiddled: col = column buffer, tw0 = table of pass 0
template <int L, int TWP>
GX_HD void gx_split_gamma_inputs(const float2 *col, const float2 *tw0, int kp, float2 *v)
{
    constexpr int RB = GxSplit<L>::RB, R2 = GxSplit<L>::R2;
    const float2 *sb = col + gx_phys(R2 * (kp & 15) + (kp >> 4));
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = sb[gx_phys(RB * c)];   // gx_phys(RB c + s) = gx_phys(RB c) + gx_phys(s) for s < RB
    gx_apply_twiddles<16, RB, TWP>(v, tw0 + kp);
}

// ---- one pass over NBUF independent buffers of length M --------------------
// s: shared array holding NBUF padded buffers, buffer j at s + j*BUFSTRIDE.
// DIT == false: butterfly then twiddle (decimation in frequency, natural in ->
//               digit-reversed out, passes run first to last);
// DIT == true : twiddle then butterfly (decimation in time, digit-reversed in
//               -> natural out, passes run last to first).  Same geometry and
//               the same twiddle table W_{S R}^{t k}.
template <int R, int S, int M, int NBUF, int BUFSTRIDE, bool DIT, int TWP = GX_TWP>
GX_HD void gx_fft_pass(float2 *s, const float2 *tw, int tid, int nthreads)
{
    constexpr int NBFLY = M / R;
    for (int w = tid; w < NBUF * NBFLY; w += nthreads) {
        const int buf = w / NBFLY;
        const int b = w - buf * NBFLY;
        const int blk = b / S;
        const int t = b - blk * S;
        const int base = blk * (S * R) + t;
        // gx_phys(base + S*n) == gx_phys(base) + gx_phys(S*n) for every schedule in
        // GxSched (no carries across the pad boundaries; checked exhaustively by
        // tests/host_emul), so the R addresses are one register plus immediates.
        float2 *sb = s + buf * BUFSTRIDE + gx_phys(base);
        float2 v[R];
#pragma unroll
        for (int n = 0; n < R; ++n) v[n] = sb[gx_phys(S * n)];
        if (DIT && S > 1) gx_apply_twiddles<R, S, TWP>(v, tw + t);
        GxDft<R>::run(v);
        if (!DIT && S > 1) gx_apply_twiddles<R, S, TWP>(v, tw + t);
#pragma unroll
        for (int k = 0; k < R; ++k) sb[gx_phys(S * k)] = v[k];
    }
}

// First DIF pass of one butterfly whose R0 inputs are already in registers
// (element n of butterfly t is x[t + S0*n]): butterfly, twiddle, store.  Lets
// a producer hand its row to the transform without a shared-memory round trip.
template <int L, int TWP = GX_TWP>
GX_HD void gx_fft_pass0_from_regs(float2 *v, float2 *s, const float2 *tw, int t)
{
    typedef GxSched<L> Sc;
    constexpr int R = Sc::R0, S = (1 << L) / Sc::R0;
    GxDft<R>::run(v);
    if (S > 1) gx_apply_twiddles<R, S, TWP>(v, tw + t);
    float2 *sb = s + gx_phys(t);
#pragma unroll
    for (int k = 0; k < R; ++k) sb[gx_phys(S * k)] = v[k];
}

// exhaustive check of the address identity used above (host tests)
template <int R, int S, int M>
static inline int gx_fft_pass_offsets_ok()
{
    for (int b = 0; b < M / R; ++b) {
        const int blk = b / S, t = b - blk * S, base = blk * (S * R) + t;
        for (int n = 0; n < R; ++n)
            if (gx_phys(base + S * n) != gx_phys(base) + gx_phys(S * n)) return 0;
    }
    return 1;
}
template <int L>
static inline int gx_fft_offsets_ok()
{
    typedef GxSched<L> Sc;
    constexpr int M = 1 << L;
    int ok = gx_fft_pass_offsets_ok<Sc::R0, M / Sc::R0, M>();
    if constexpr (Sc::NP > 1) ok &= gx_fft_pass_offsets_ok<Sc::R1, M / Sc::R0 / Sc::R1, M>();
    if constexpr (Sc::NP > 2) ok &= gx_fft_pass_offsets_ok<Sc::R2, M / Sc::R0 / Sc::R1 / Sc::R2, M>();
    if constexpr (Sc::NP > 3) ok &= gx_fft_pass_offsets_ok<Sc::R3, M / Sc::R0 / Sc::R1 / Sc::R2 / Sc::R3, M>();
    return ok;
}

#if defined(__CUDACC__)
#define GX_BLOCK_SYNC() __syncthreads()
#define GX_DEV __device__ __forceinline__
#else
#define GX_BLOCK_SYNC() ((void)0)
#define GX_DEV inline
#endif

// Forward transform of NBUF buffers, natural order in, coefficient k left at
// slot gx_fft_pos<L>(k).  On the device every thread of the block calls this
// (it contains __syncthreads); the host-emulation build calls it with
// nthreads == 1, which runs the passes sequentially.
// passes 1 .. NP-1 of the DIF transform (pass 0 done by the caller)
template <int L, int NBUF, int BUFSTRIDE>
GX_DEV void gx_fft_dif_tail(float2 *s, const float2 *tw, const int *tw_off, int tid, int nthreads)
{
    typedef GxSched<L> Sc;
    constexpr int M = 1 << L;
    if constexpr (Sc::NP > 1) {
        gx_fft_pass<Sc::R1, M / Sc::R0 / Sc::R1, M, NBUF, BUFSTRIDE, false>(s, tw + tw_off[1], tid, nthreads);
        GX_BLOCK_SYNC();
    }
    if constexpr (Sc::NP > 2) {
        gx_fft_pass<Sc::R2, M / Sc::R0 / Sc::R1 / Sc::R2, M, NBUF, BUFSTRIDE, false>(s, tw + tw_off[2], tid, nthreads);
        GX_BLOCK_SYNC();
    }
    if constexpr (Sc::NP > 3) {
        gx_fft_pass<Sc::R3, M / Sc::R0 / Sc::R1 / Sc::R2 / Sc::R3, M, NBUF, BUFSTRIDE, false>(s, tw + tw_off[3], tid, nthreads);
        GX_BLOCK_SYNC();
    }
}

template <int L, int NBUF, int BUFSTRIDE>
GX_DEV void gx_fft_dif(float2 *s, const float2 *tw, const int *tw_off, int tid, int nthreads)
{
    typedef GxSched<L> Sc;
    constexpr int M = 1 << L;
    gx_fft_pass<Sc::R0, M / Sc::R0, M, NBUF, BUFSTRIDE, false>(s, tw + tw_off[0], tid, nthreads);
    GX_BLOCK_SYNC();
    if constexpr (Sc::NP > 1) {
        gx_fft_pass<Sc::R1, M / Sc::R0 / Sc::R1, M, NBUF, BUFSTRIDE, false>(s, tw + tw_off[1], tid, nthreads);
        GX_BLOCK_SYNC();
    }
    if constexpr (Sc::NP > 2) {
        gx_fft_pass<Sc::R2, M / Sc::R0 / Sc::R1 / Sc::R2, M, NBUF, BUFSTRIDE, false>(s, tw + tw_off[2], tid, nthreads);
        GX_BLOCK_SYNC();
    }
    if constexpr (Sc::NP > 3) {
        gx_fft_pass<Sc::R3, M / Sc::R0 / Sc::R1 / Sc::R2 / Sc::R3, M, NBUF, BUFSTRIDE, false>(s, tw + tw_off[3], tid, nthreads);
        GX_BLOCK_SYNC();
    }
}

// Forward transform taking its input in the slot order gx_fft_dif leaves
// (value for index k at slot gx_fft_pos<L>(k)) and producing natural order.
template <int L, int NBUF, int BUFSTRIDE>
GX_DEV void gx_fft_dit(float2 *s, const float2 *tw, const int *tw_off, int tid, int nthreads)
{
    typedef GxSched<L> Sc;
    constexpr int M = 1 << L;
    if constexpr (Sc::NP > 3) {
        gx_fft_pass<Sc::R3, M / Sc::R0 / Sc::R1 / Sc::R2 / Sc::R3, M, NBUF, BUFSTRIDE, true>(s, tw + tw_off[3], tid, nthreads);
        GX_BLOCK_SYNC();
    }
    if constexpr (Sc::NP > 2) {
        gx_fft_pass<Sc::R2, M / Sc::R0 / Sc::R1 / Sc::R2, M, NBUF, BUFSTRIDE, true>(s, tw + tw_off[2], tid, nthreads);
        GX_BLOCK_SYNC();
    }
    if constexpr (Sc::NP > 1) {
        gx_fft_pass<Sc::R1, M / Sc::R0 / Sc::R1, M, NBUF, BUFSTRIDE, true>(s, tw + tw_off[1], tid, nthreads);
        GX_BLOCK_SYNC();
    }
    gx_fft_pass<Sc::R0, M / Sc::R0, M, NBUF, BUFSTRIDE, true>(s, tw + tw_off[0], tid, nthreads);
    GX_BLOCK_SYNC();
}

// Length-N DFT (N == M for powers of two, Bluestein otherwise) of NBUF buffers.
// Input: natural order in slots [0,N) -- for Bluestein the caller has already
// multiplied element n by chirp[n] and zeroed slots [N,M).
// Output: read coefficient k with gx_dft_result<L>().
// BLUE: -1 decide at run time from g.bluestein, 0 / 1 fixed at compile time (the
// hot kernels are instantiated per flavour so no test sits in their loops).
// PASS0_DONE: the caller already ran the first DIF pass (gx_fft_pass0_from_regs)
// and synchronised.
template <int L, int NBUF, int BUFSTRIDE, int BLUE = -1, bool PASS0_DONE = false>
GX_DEV void gx_dft_block(float2 *s, const GxFftLayout &g, const float2 *plan, int tid, int nthreads)
{
    constexpr int M = 1 << L;
    if (PASS0_DONE) gx_fft_dif_tail<L, NBUF, BUFSTRIDE>(s, plan, g.tw_off, tid, nthreads);
    else gx_fft_dif<L, NBUF, BUFSTRIDE>(s, plan, g.tw_off, tid, nthreads);
    if (BLUE < 0 ? g.bluestein != 0 : BLUE != 0) {
        // circular convolution with the conjugate chirp: multiply by its spectrum
        // (stored in slot order, 1/M folded in), conjugate, transform again
        // (ifft(v) = conj(fft(conj v))).  The DIT flavour consumes slot order
        // directly, so nothing has to be permuted.
        const float2 *bhat = plan + g.bhat_off;
        for (int w = tid; w < NBUF * M; w += nthreads) {
            const int buf = w / M, p = w - buf * M;
            float2 *sb = s + buf * BUFSTRIDE;
            sb[gx_phys(p)] = gx_conj(gx_cmul(sb[gx_phys(p)], bhat[p]));
        }
        GX_BLOCK_SYNC();
        gx_fft_dit<L, NBUF, BUFSTRIDE>(s, plan, g.tw_off, tid, nthreads);
    }
}

template <int L, int BLUE = -1>
GX_HD float2 gx_dft_result(const float2 *sb, const GxFftLayout &g, const float2 *plan, int k)
{
    if (BLUE < 0 ? g.bluestein != 0 : BLUE != 0)
        return gx_cmul(plan[g.chirp_off + k], gx_conj(sb[gx_phys(k)]));
    return sb[gx_phys(gx_fft_pos<L>(k))];
}
