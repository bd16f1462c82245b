// dft_reg.cuh -- in-register DFTs of compile-time size (complex, real->Hermitian, Hermitian->real).
//
// Every index is a constant expression (static_for), so the arrays live in registers and every
// twiddle factor is a literal folded by the compiler (constexpr sin/cos below).  These are the
// butterflies of the shared-memory FFT passes in fft_pass.cuh; DIR = -1 is the forward transform
// exp(-2 pi i nk/R) (numpy fftn convention, reference jubik/convolve.py:165-166), DIR = +1 the
// un-normalised inverse (convolve.py:174 ifftn; its 1/P^2 is folded into the PSF spectra).
#pragma once
#include "jb_rt.h"

namespace jbm {
constexpr double PI = 3.14159265358979323846264338327950288;
constexpr double SQRT1_2 = 0.70710678118654752440084436210484904;

constexpr double sin_small(double x) {  // |x| <= pi/4
    double x2 = x * x, r = 1.0;
    for (int k = 11; k >= 1; --k) r = 1.0 - x2 / double((2 * k) * (2 * k + 1)) * r;
    return x * r;
}
constexpr double cos_small(double x) {  // |x| <= pi/4
    double x2 = x * x, r = 1.0;
    for (int k = 11; k >= 1; --k) r = 1.0 - x2 / double((2 * k - 1) * (2 * k)) * r;
    return r;
}
// a/b of a turn, a/b in [0, 1/4]
constexpr double cosq(long long a, long long b) {
    return (8 * a > b) ? sin_small(2.0 * PI * double(b - 4 * a) / double(4 * b))
                       : cos_small(2.0 * PI * double(a) / double(b));
}
constexpr double sinq(long long a, long long b) {
    return (8 * a > b) ? cos_small(2.0 * PI * double(b - 4 * a) / double(4 * b))
                       : sin_small(2.0 * PI * double(a) / double(b));
}
constexpr double cos2pi(long long a, long long b) {  // cos(2 pi a / b)
    a %= b;
    if (a < 0) a += b;
    if (2 * a > b) a = b - a;
    if (4 * a > b) return -cosq(b - 2 * a, 2 * b);
    return cosq(a, b);
}
constexpr double sin2pi(long long a, long long b) {  // sin(2 pi a / b)
    a %= b;
    if (a < 0) a += b;
    double sg = 1.0;
    if (2 * a > b) { a = b - a; sg = -1.0; }
    if (4 * a > b) return sg * sinq(b - 2 * a, 2 * b);
    return sg * sinq(a, b);
}
constexpr int smallest_odd_factor(int r) {
    for (int f = 3; f * f <= r; f += 2)
        if (r % f == 0) return f;
    return r;
}
constexpr bool is_odd_prime(int r) { return (r & 1) && r > 2 && smallest_odd_factor(r) == r; }
// first radix of the in-register Cooley-Tukey split
constexpr int pick_radix(int r) {
    if (r % 4 == 0) return 4;
    if (r % 2 == 0) return 2;
    return smallest_odd_factor(r);
}
}  // namespace jbm

// v * exp(DIR * 2 pi i * NUM / DEN) with the trivial angles special-cased
template <int NUM, int DEN, int DIR>
JB_HD cd twmul(cd v) {
    constexpr int e = ((NUM % DEN) + DEN) % DEN;
    if constexpr (e == 0) {
        return v;
    } else if constexpr (2 * e == DEN) {
        return cd_make(-v.x, -v.y);
    } else if constexpr (4 * e == DEN) {  // * (DIR i)
        return cd_make(-DIR * v.y, DIR * v.x);
    } else if constexpr (4 * e == 3 * DEN) {  // * (-DIR i)
        return cd_make(DIR * v.y, -DIR * v.x);
    } else if constexpr ((8 * e) % DEN == 0) {  // odd multiples of 1/8 turn
        constexpr double h = jbm::SQRT1_2;
        constexpr int o = 8 * e / DEN;  // 1,3,5,7
        constexpr double c = (o == 1 || o == 7) ? h : -h;
        constexpr double s = DIR * ((o == 1 || o == 3) ? h : -h);
        return cd_make(c * v.x - s * v.y, s * v.x + c * v.y);
    } else {
        constexpr double c = jbm::cos2pi(e, DEN);
        constexpr double s = DIR * jbm::sin2pi(e, DEN);
        return cd_make(c * v.x - s * v.y, s * v.x + c * v.y);
    }
}

template <int R, int DIR> JB_HD void dft(cd (&v)[R]);

// odd prime radix: symmetric/antisymmetric split, (R-1)^2/2 real FMAs per part
template <int R, int DIR>
JB_HD void dft_odd(cd (&v)[R]) {
    constexpr int H = (R - 1) / 2;
    cd t[H + 1], u[H + 1];
    static_for<1, H + 1>([&](auto kk) {
        constexpr int k = kk;
        t[k] = cd_add(v[k], v[R - k]);
        u[k] = cd_sub(v[k], v[R - k]);
    });
    cd x0 = v[0];
    cd acc = x0;
    static_for<1, H + 1>([&](auto kk) { constexpr int k = kk; acc = cd_add(acc, t[k]); });
    v[0] = acc;
    static_for<1, H + 1>([&](auto jj) {
        constexpr int j = jj;
        cd re = x0;
        cd im = cd_make(0.0, 0.0);
        static_for<1, H + 1>([&](auto kk) {
            constexpr int k = kk;
            constexpr double c = jbm::cos2pi((long long)j * k, R);
            constexpr double s = DIR * jbm::sin2pi((long long)j * k, R);
            re.x += c * t[k].x;
            re.y += c * t[k].y;
            im.x += s * u[k].x;
            im.y += s * u[k].y;
        });
        // X_j = re + i*im ; X_{R-j} = re - i*im
        v[j] = cd_make(re.x - im.y, re.y + im.x);
        v[R - j] = cd_make(re.x + im.y, re.y - im.x);
    });
}

template <int R, int DIR>
JB_HD void dft(cd (&v)[R]) {
    if constexpr (R == 1) {
    } else if constexpr (R == 2) {
        cd a = v[0], b = v[1];
        v[0] = cd_add(a, b);
        v[1] = cd_sub(a, b);
    } else if constexpr (R == 4) {
        cd t0 = cd_add(v[0], v[2]), t1 = cd_sub(v[0], v[2]);
        cd t2 = cd_add(v[1], v[3]), d = cd_sub(v[1], v[3]);
        cd t3 = cd_make(-DIR * d.y, DIR * d.x);  // (DIR i) * d
        v[0] = cd_add(t0, t2);
        v[2] = cd_sub(t0, t2);
        v[1] = cd_add(t1, t3);
        v[3] = cd_sub(t1, t3);
    } else if constexpr (jbm::is_odd_prime(R)) {
        dft_odd<R, DIR>(v);
    } else {
        constexpr int A = jbm::pick_radix(R), B = R / A;
        cd y[A][B];
        static_for<0, B>([&](auto bb) {
            constexpr int b = bb;
            cd t[A];
            static_for<0, A>([&](auto aa) { constexpr int a = aa; t[a] = v[B * a + b]; });
            dft<A, DIR>(t);
            static_for<0, A>([&](auto kk) { constexpr int k1 = kk; y[k1][b] = twmul<b * k1, R, DIR>(t[k1]); });
        });
        static_for<0, A>([&](auto kk) {
            constexpr int k1 = kk;
            dft<B, DIR>(y[k1]);
            static_for<0, B>([&](auto k2) { constexpr int q = k2; v[k1 + A * q] = y[k1][q]; });
        });
    }
}

// real x[R] -> X[0..R/2] = sum_n x[n] exp(-2 pi i nk/R)          (R even)
template <int R>
JB_HD void rdft_fwd(const double (&x)[R], cd (&X)[R / 2 + 1]) {
    static_assert(R % 2 == 0, "real DFT size must be even");
    constexpr int H = R / 2;
    cd z[H];
    static_for<0, H>([&](auto nn) { constexpr int n = nn; z[n] = cd_make(x[2 * n], x[2 * n + 1]); });
    dft<H, -1>(z);
    static_for<0, H + 1>([&](auto kk) {
        constexpr int k = kk;
        cd zk = z[k % H];
        cd zc = cd_conj(z[(H - k) % H]);
        cd e = cd_make(0.5 * (zk.x + zc.x), 0.5 * (zk.y + zc.y));
        // o = (zk - zc) / (2i) = (-i/2)(zk - zc)
        cd d = cd_sub(zk, zc);
        cd o = cd_make(0.5 * d.y, -0.5 * d.x);
        X[k] = cd_add(e, twmul<k, R, -1>(o));
    });
}

// Hermitian X[0..R/2] -> real x[n] = sum_{k<R} X[k] exp(+2 pi i nk/R)  (un-normalised; R even)
template <int R>
JB_HD void rdft_inv(const cd (&X)[R / 2 + 1], double (&x)[R]) {
    static_assert(R % 2 == 0, "real DFT size must be even");
    constexpr int H = R / 2;
    cd z[H];
    static_for<0, H>([&](auto kk) {
        constexpr int k = kk;
        cd xk = X[k];
        cd xc = cd_conj(X[H - k]);
        cd s = cd_add(xk, xc);
        cd d = twmul<k, R, +1>(cd_sub(xk, xc));
        // Z = s + i d
        z[k] = cd_make(s.x - d.y, s.y + d.x);
    });
    dft<H, +1>(z);
    static_for<0, H>([&](auto nn) { constexpr int n = nn; x[2 * n] = z[n].x; x[2 * n + 1] = z[n].y; });
}
