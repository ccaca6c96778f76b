// Host build of the device FFT engine (giwaxsim_b200/csrc/gx_fft_engine.cuh)
// so that its pass / permutation / Bluestein logic can be unit-tested against
// numpy.fft on a machine without a GPU.  Test infrastructure only.
#include <vector>
#include <cstring>
#include "../../giwaxsim_b200/csrc/gx_fft_engine.cuh"

template <int L>
static void run(const GxFftLayout &g, const float2 *plan, const float2 *in, float2 *out)
{
    const int M = 1 << L;
    std::vector<float2> s(gx_phys_len(M));
    for (int n = 0; n < M; ++n) {
        float2 v = make_float2(0.f, 0.f);
        if (n < g.N) {
            v = in[n];
            if (g.bluestein) v = gx_cmul(v, plan[g.chirp_off + n]);
        }
        s[gx_phys(n)] = v;
    }
    gx_dft_block<L, 1, 0>(s.data(), g, plan, 0, 1);
    for (int k = 0; k < g.N; ++k) out[k] = gx_dft_result<L>(s.data(), g, plan, k);
}

extern "C" int emul_dft(int N, const float *plan, const float *in, float *out)
{
    GxFftLayout g = gx_fft_layout(N);
    if (g.M == 0) return -1;
    const float2 *p = reinterpret_cast<const float2 *>(plan);
    const float2 *i = reinterpret_cast<const float2 *>(in);
    float2 *o = reinterpret_cast<float2 *>(out);
    switch (g.L) {
    case 4: run<4>(g, p, i, o); break;
    case 5: run<5>(g, p, i, o); break;
    case 6: run<6>(g, p, i, o); break;
    case 7: run<7>(g, p, i, o); break;
    case 8: run<8>(g, p, i, o); break;
    case 9: run<9>(g, p, i, o); break;
    case 10: run<10>(g, p, i, o); break;
    case 11: run<11>(g, p, i, o); break;
    case 12: run<12>(g, p, i, o); break;
    case 13: run<13>(g, p, i, o); break;
    case 14: run<14>(g, p, i, o); break;
    default: return -2;
    }
    return 0;
}

// 1 iff the "one base register + immediate offsets" addressing used by
// gx_fft_pass is exact for every pass of every schedule.
extern "C" int emul_offsets_ok()
{
    return gx_fft_offsets_ok<4>() & gx_fft_offsets_ok<5>() & gx_fft_offsets_ok<6>() & gx_fft_offsets_ok<7>() &
           gx_fft_offsets_ok<8>() & gx_fft_offsets_ok<9>() & gx_fft_offsets_ok<10>() & gx_fft_offsets_ok<11>() &
           gx_fft_offsets_ok<12>() & gx_fft_offsets_ok<13>() & gx_fft_offsets_ok<14>();
}

// 4096-point transform with the band-limited last pass (gx_fft_lastpass16_lowband): passes 0 and 1
// as usual, then only X[0], X[15] (and X[1], X[14] where the band needs them) of every last-pass
// butterfly.  out[k - klo] for klo <= k < khi (negative k = coefficient M + k).
extern "C" int emul_dft_lowband(const float *plan, const float *in, float *out, int klo, int khi, int nthreads)
{
    constexpr int L = 12, M = 1 << L;
    GxFftLayout g = gx_fft_layout(M);
    if (klo < -512 || khi > 512) return -1;
    const float2 *p = reinterpret_cast<const float2 *>(plan);
    const float2 *x = reinterpret_cast<const float2 *>(in);
    float2 *o = reinterpret_cast<float2 *>(out);
    std::vector<float2> s(gx_phys_len(M));
    for (int n = 0; n < M; ++n) s[gx_phys(n)] = x[n];
    gx_fft_pass<16, M / 16, M, 1, 0, false>(s.data(), p + g.tw_off[0], 0, 1);
    gx_fft_pass<16, M / 256, M, 1, 0, false>(s.data(), p + g.tw_off[1], 0, 1);
    for (int t = 0; t < nthreads; ++t) gx_fft_lastpass16_lowband<M>(s.data(), klo, khi, t, nthreads);
    for (int k = klo; k < khi; ++k) o[k - klo] = gx_dft_result<L, 0>(s.data(), g, p, k < 0 ? k + M : k);
    return 0;
}

// The split N-point transform of the TMA-fed column kernel (16 boxes of N / 16 rows; gx_split_*), N = 1024, 2048,
// 4096: rows are placed at GxSplit<L>::slot(z) as the row kernel does, every box goes through alpha and beta,
// gamma forms all 16 outputs (full != 0) or only the kept band through gx_dft16_lowband_vals.
// out[k - klo] for klo <= k < khi (negative k = coefficient N + k).
template <int L>
static int run_split(const float2 *p, const float2 *x, float2 *o, int klo, int khi, int full)
{
    typedef GxSplit<L> Sp;
    constexpr int M = Sp::N, RB = Sp::RB, R2 = Sp::R2;
    GxFftLayout g = gx_fft_layout(M);
    std::vector<float2> dense(M), col(gx_phys_len(M));
    for (int z = 0; z < M; ++z) dense[Sp::slot(z)] = x[z];
    for (int c = 0; c < 16; ++c) {
        for (int t = 0; t < R2; ++t)
            gx_split_alpha<L, 1>(dense.data() + RB * c + t, R2, col.data() + gx_phys(RB * c + t), p + g.tw_off[1], t);
        for (int blk = 0; blk < 16; ++blk) gx_split_beta<L>(col.data() + gx_phys(RB * c + R2 * blk));
    }
    std::vector<float2> X(M, make_float2(0.f, 0.f));
    for (int kp = 0; kp < RB; ++kp) {
        float2 v[16];
        gx_split_gamma_inputs<L, 1>(col.data(), p + g.tw_off[0], kp, v);
        if (full) {
            GxDft<16>::run(v);
            for (int m = 0; m < 16; ++m) X[kp + RB * m] = v[m];
        } else {
            if (klo < -2 * RB || khi > 2 * RB) return -1;
            const bool w1 = kp + RB < khi, w14 = kp - 2 * RB >= klo;
            float2 x0, x15, x1, x14;
            gx_dft16_lowband_vals(v, w1 || w14, x0, x15, x1, x14);
            X[kp] = x0; X[kp + 15 * RB] = x15;
            if (w1) X[kp + RB] = x1;
            if (w14) X[kp + 14 * RB] = x14;
        }
    }
    for (int k = klo; k < khi; ++k) o[k - klo] = X[k < 0 ? k + M : k];
    return 0;
}

extern "C" int emul_dft_split(int N, const float *plan, const float *in, float *out, int klo, int khi, int full)
{
    const float2 *p = reinterpret_cast<const float2 *>(plan);
    const float2 *x = reinterpret_cast<const float2 *>(in);
    float2 *o = reinterpret_cast<float2 *>(out);
    switch (N) {
    case 1024: return run_split<10>(p, x, o, klo, khi, full);
    case 2048: return run_split<11>(p, x, o, klo, khi, full);
    case 4096: return run_split<12>(p, x, o, klo, khi, full);
    }
    return -2;
}
