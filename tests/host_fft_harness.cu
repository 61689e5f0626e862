// CPU harness for csrc/fft2_core.cuh: runs the exact pass code with the group's
// threads emulated by loops and compares against a float64 O(N^2) DFT.
// Build + run: see tests/test_host_fft_harness.py (nvcc host compile, no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../track_analyser_b200/csrc/fft_core.cuh"
#include "../track_analyser_b200/csrc/fft2_core.cuh"

using namespace ta;

// Packed two-transform core (fft2_core.cuh): transforms A = a + i*b and B = c + i*d share every
// instruction; barriers of the kernel are the boundaries between the thread loops below.
template <int N>
double run2() {
    using C = FftCfg<N>;
    using namespace ta::p2;
    std::vector<float2> tw1(15 * C::M), tw2(16 * C::Q);
    std::vector<float4> ex(Ex<N>::SLOTS);
    const double PI = 3.14159265358979323846;
    for (int k1 = 1; k1 < 16; ++k1)
        for (int r = 0; r < C::M; ++r) {
            double a = -2.0 * PI * double((r * k1) % N) / N;
            tw1[(k1 - 1) * C::M + r] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int k2 = 0; k2 < 16; ++k2)
        for (int n3 = 0; n3 < C::Q; ++n3) {
            double a = -2.0 * PI * double(n3 * k2) / C::M;
            tw2[k2 * C::Q + n3] = make_float2((float)cos(a), (float)sin(a));
        }
    std::vector<float> in[4];
    srand(4321 + N);
    for (auto& v : in) {
        v.resize(N);
        for (int n = 0; n < N; ++n) v[n] = (float)rand() / RAND_MAX - 0.5f;
    }
    std::vector<C2> regs(C::M * 16);
    auto R = [&](int t) -> C2(&)[16] { return *reinterpret_cast<C2(*)[16]>(&regs[t * 16]); };
    for (int r = 0; r < C::M; ++r) {
        for (int n1 = 0; n1 < 16; ++n1) {
            const int n = n1 * C::M + r;
            R(r)[n1].re = make_float2(0.5f * in[0][n], 0.5f * in[2][n]);
            R(r)[n1].im = make_float2(0.5f * in[1][n], 0.5f * in[3][n]);
        }
        pass1<N>(R(r), r, tw1.data(), ex.data());
    }
    for (int t = 0; t < C::M; ++t) pass2<N>(R(t), t, tw2.data(), ex.data());
    for (int t = 0; t < C::M; ++t) pass3<N>(R(t), t, ex.data());
    // reference spectra in double
    std::vector<double> ref[4][2];
    for (int q = 0; q < 4; ++q) {
        ref[q][0].assign(N / 2 + 1, 0.0);
        ref[q][1].assign(N / 2 + 1, 0.0);
        for (int k = 0; k <= N / 2; ++k)
            for (int n = 0; n < N; ++n) {
                double ang = -2.0 * PI * double((long long)n * k % N) / N;
                ref[q][0][k] += in[q][n] * cos(ang);
                ref[q][1][k] += in[q][n] * sin(ang);
            }
    }
    double maxerr = 0, maxref = 0;
    std::vector<int> seen(N / 2 + 1, 0);
    auto check = [&](int k, const C2& zk, const C2& zn) {
        C2 xa, xb;
        split_pair(zk, zn, xa, xb);
        const double got[4][2] = {{xa.re.x, xa.im.x}, {xb.re.x, xb.im.x}, {xa.re.y, xa.im.y}, {xb.re.y, xb.im.y}};
        for (int q = 0; q < 4; ++q)
            for (int c = 0; c < 2; ++c) maxerr = fmax(maxerr, fabs(got[q][c] - ref[q][c][k]));
        maxref = fmax(maxref, sqrt(ref[0][0][k] * ref[0][0][k] + ref[0][1][k] * ref[0][1][k]));
        seen[k]++;
    };
    for (int t = 0; t < C::M; ++t) {
        for (int i = 0; i < 8; ++i) {
            const int k = kept_bin<N>(t, i);
            const C2 zk = R(t)[kept_reg<N>(i)];
            const C2 zn = (k == 0) ? zk : unpack(ex[Ex<N>::slot_of((N - k) & (N - 1))]);
            check(k, zk, zn);
        }
        if (t == 0) {
            const C2 z = R(0)[C::Q / 2];  // bin N/2 lives in thread 0, butterfly 0, k3 = Q/2
            check(N / 2, z, z);
        }
    }
    for (int k = 0; k <= N / 2; ++k)
        if (seen[k] != 1) { printf("bin %d covered %d times\n", k, seen[k]); return 1.0; }
    printf("packed N=%d maxerr=%.3e maxref=%.3e rel=%.3e\n", N, maxerr, maxref, maxerr / maxref);
    return maxerr / maxref;
}

// Paired pass 3: every mirror bin is in the same thread's registers.
template <int N>
double run3() {
    using C = FftCfg<N>;
    using namespace ta::p2;
    using P = Pair3<N>;
    std::vector<float2> tw1(15 * C::M), tw2(16 * C::Q);
    std::vector<float4> ex(Ex<N>::SLOTS);
    const double PI = 3.14159265358979323846;
    for (int k1 = 1; k1 < 16; ++k1)
        for (int r = 0; r < C::M; ++r) {
            double a = -2.0 * PI * double((r * k1) % N) / N;
            tw1[(k1 - 1) * C::M + r] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int k2 = 0; k2 < 16; ++k2)
        for (int n3 = 0; n3 < C::Q; ++n3) {
            double a = -2.0 * PI * double(n3 * k2) / C::M;
            tw2[k2 * C::Q + n3] = make_float2((float)cos(a), (float)sin(a));
        }
    std::vector<float> in[4];
    srand(777 + N);
    for (auto& v : in) {
        v.resize(N);
        for (int n = 0; n < N; ++n) v[n] = (float)rand() / RAND_MAX - 0.5f;
    }
    std::vector<C2> regs(C::M * 16);
    auto R = [&](int t) -> C2(&)[16] { return *reinterpret_cast<C2(*)[16]>(&regs[t * 16]); };
    for (int r = 0; r < C::M; ++r) {
        for (int n1 = 0; n1 < 16; ++n1) {
            const int n = n1 * C::M + r;
            R(r)[n1].re = make_float2(0.5f * in[0][n], 0.5f * in[2][n]);
            R(r)[n1].im = make_float2(0.5f * in[1][n], 0.5f * in[3][n]);
        }
        pass1<N>(R(r), r, tw1.data(), ex.data());
    }
    for (int t = 0; t < C::M; ++t) pass2<N>(R(t), t, tw2.data(), ex.data());
    for (int t = 0; t < C::M; ++t) pass3_paired<N>(R(t), t, ex.data());
    std::vector<double> ref[4][2];
    for (int q = 0; q < 4; ++q) {
        ref[q][0].assign(N / 2 + 1, 0.0);
        ref[q][1].assign(N / 2 + 1, 0.0);
        for (int k = 0; k <= N / 2; ++k)
            for (int n = 0; n < N; ++n) {
                double ang = -2.0 * PI * double((long long)n * k % N) / N;
                ref[q][0][k] += in[q][n] * cos(ang);
                ref[q][1][k] += in[q][n] * sin(ang);
            }
    }
    double maxerr = 0, maxref = 0;
    std::vector<int> seen(N / 2 + 1, 0);
    auto check = [&](int k, const C2& zk, const C2& zn) {
        C2 xa, xb;
        split_pair(zk, zn, xa, xb);
        const double got[4][2] = {{xa.re.x, xa.im.x}, {xb.re.x, xb.im.x}, {xa.re.y, xa.im.y}, {xb.re.y, xb.im.y}};
        for (int q = 0; q < 4; ++q)
            for (int c = 0; c < 2; ++c) maxerr = fmax(maxerr, fabs(got[q][c] - ref[q][c][k]));
        maxref = fmax(maxref, sqrt(ref[0][0][k] * ref[0][0][k] + ref[0][1][k] * ref[0][1][k]));
        seen[k]++;
    };
    constexpr int Q = C::Q, HB = P::HB;
    for (int t = 0; t < C::M; ++t)
        for (int b = 0; b < HB; ++b) {
            C2* va = &R(t)[b * Q];
            C2* vb = &R(t)[(HB + b) * Q];
            if (t == 0 && b == 0) {  // (0,0) and (0,8) mirror themselves
                for (int k3 = 0; k3 <= Q / 2; ++k3) check(256 * k3, va[k3], va[(Q - k3) % Q]);
                for (int k3 = 0; k3 < Q / 2; ++k3) check(128 + 256 * k3, vb[k3], vb[Q - 1 - k3]);
            } else {
                for (int k3 = 0; k3 < Q / 2; ++k3) {
                    check(P::bin(t, b, 0, k3), va[k3], vb[Q - 1 - k3]);
                    check(P::bin(t, b, 1, k3), vb[k3], va[Q - 1 - k3]);
                }
            }
        }
    for (int k = 0; k <= N / 2; ++k)
        if (seen[k] != 1) { printf("paired: bin %d covered %d times\n", k, seen[k]); return 1.0; }
    printf("paired N=%d maxerr=%.3e maxref=%.3e rel=%.3e\n", N, maxerr, maxref, maxerr / maxref);
    return maxerr / maxref;
}

int main() {
    double e = 0;
    e = fmax(e, run2<1024>());
    e = fmax(e, run2<2048>());
    e = fmax(e, run2<4096>());
    e = fmax(e, run3<1024>());
    e = fmax(e, run3<2048>());
    if (e > 2e-6) { printf("FAIL\n"); return 1; }
    printf("OK\n");
    return 0;
}
