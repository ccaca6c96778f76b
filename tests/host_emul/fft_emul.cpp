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
           gx_fft_offsets_ok<12>() & gx_fft_offsets_ok<13>();
}
