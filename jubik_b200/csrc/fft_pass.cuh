// fft_pass.cuh -- CTA-level batched 1-D FFTs of compile-time length L = R1*R2 built from the
// in-register DFTs of dft_reg.cuh with ONE shared-memory exchange per transform.
//
//   rows_r2c / rows_c2r : B real sequences <-> Hermitian half spectra (k = 0..L/2).  The first
//       pass is a real-input radix-R1 butterfly, so no two-for-one packing or unpack pass is
//       needed and a real transform costs about half a complex one.  Lanes run along the
//       sequence (coalesced row access), the exchange buffer is padded to an odd stride.
//   cols_cfft           : CB complex sequences that are the COLUMNS of a row-major tile
//       (element p of column b at [p][b]).  Lanes run across columns, which is both the
//       coalesced direction in HBM and conflict-free in shared memory.  Forward transform,
//       pointwise multiply by the resident PSF spectrum and inverse transform are chained in
//       registers: the spectrum never leaves the thread between the two directions.
//
// Index convention (decimation in time):  n = R2*n1 + n2,  k = k1 + R1*k2,
//   X[k] = sum_{n2} W_L^{n2 k1} W_R2^{n2 k2} ( sum_{n1} x[R2 n1 + n2] W_R1^{n1 k1} ).
// tw_nk[k1*R2 + n2] = tw_kn[n2*R1 + k1] = exp(-2 pi i n2 k1 / L)   (two orderings of one table so
// that either lane direction reads it contiguously).
#pragma once
#include "dft_reg.cuh"

template <int L_, int R1_, int R2_>
struct FftPlan {
    static constexpr int L = L_, R1 = R1_, R2 = R2_;
    static_assert(R1_ * R2_ == L_, "L = R1*R2");
    static_assert(R1_ % 2 == 0, "R1 must be even (real first pass)");
    static constexpr int H1 = R1_ / 2 + 1;          // Hermitian-unique k1 values
    static constexpr int R2P = (R2_ == 1) ? 1 : (R2_ | 1);  // odd stride of the exchange buffer
    static constexpr int LS = (R2_ == 1) ? 0 : H1 * R2P;    // cd per real sequence in smem
};

JB_D int jb_tid() { return threadIdx.x; }
JB_D int jb_nthr() { return blockDim.x; }

// ---------------------------------------------------------------------------------- real rows
// ld(b, p) -> double, p in [0,L) ; st(b, k, X) for k in [0, L/2]
template <class FP, class Ld, class St>
JB_D void rows_r2c(cd* sm, const cd* __restrict__ tw_nk, int B, Ld ld, St st) {
    constexpr int L = FP::L, R1 = FP::R1, R2 = FP::R2, H1 = FP::H1, R2P = FP::R2P, LS = FP::LS;
    if constexpr (R2 == 1) {
        for (int b = jb_tid(); b < B; b += jb_nthr()) {
            double x[R1];
            static_for<0, R1>([&](auto nn) { constexpr int n = nn; x[n] = ld(b, n); });
            cd X[H1];
            rdft_fwd<R1>(x, X);
            static_for<0, H1>([&](auto kk) { constexpr int k = kk; st(b, k, X[k]); });
        }
    } else {
        for (int it = jb_tid(); it < B * R2; it += jb_nthr()) {
            int b = it / R2, n2 = it - b * R2;
            double x[R1];
            static_for<0, R1>([&](auto nn) { constexpr int n1 = nn; x[n1] = ld(b, R2 * n1 + n2); });
            cd Y[H1];
            rdft_fwd<R1>(x, Y);
            cd* dst = sm + (size_t)b * LS + n2;
            static_for<0, H1>([&](auto kk) {
                constexpr int k1 = kk;
                cd y = Y[k1];
                if constexpr (k1 > 0) y = cd_mul(y, jb_ldg(tw_nk + k1 * R2 + n2));
                dst[k1 * R2P] = y;
            });
        }
        JB_SYNC();
        for (int it = jb_tid(); it < B * H1; it += jb_nthr()) {
            int b = it / H1, k1 = it - b * H1;
            cd u[R2];
            const cd* src = sm + (size_t)b * LS + k1 * R2P;
            static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; u[n2] = src[n2]; });
            dft<R2, -1>(u);
            bool mirror = (k1 != 0) && (2 * k1 != R1);
            static_for<0, R2>([&](auto kk) {
                constexpr int k2 = kk;
                int k = k1 + R1 * k2;
                if (2 * k <= L) st(b, k, u[k2]);
                else if (mirror) st(b, L - k, cd_conj(u[k2]));
            });
        }
    }
}

// ld(b, k) -> cd for k in [0, L/2] ; st(b, p, double) for p in [0,L)   (un-normalised inverse)
template <class FP, class Ld, class St>
JB_D void rows_c2r(cd* sm, const cd* __restrict__ tw_kn, int B, Ld ld, St st) {
    constexpr int L = FP::L, R1 = FP::R1, R2 = FP::R2, H1 = FP::H1, R2P = FP::R2P, LS = FP::LS;
    if constexpr (R2 == 1) {
        for (int b = jb_tid(); b < B; b += jb_nthr()) {
            cd X[H1];
            static_for<0, H1>([&](auto kk) { constexpr int k = kk; X[k] = ld(b, k); });
            double x[R1];
            rdft_inv<R1>(X, x);
            static_for<0, R1>([&](auto nn) { constexpr int n = nn; st(b, n, x[n]); });
        }
    } else {
        for (int it = jb_tid(); it < B * H1; it += jb_nthr()) {
            int b = it / H1, k1 = it - b * H1;
            cd u[R2];
            static_for<0, R2>([&](auto kk) {
                constexpr int k2 = kk;
                int k = k1 + R1 * k2;
                u[k2] = (2 * k <= L) ? ld(b, k) : cd_conj(ld(b, L - k));
            });
            dft<R2, +1>(u);
            cd* dst = sm + (size_t)b * LS + k1 * R2P;
            static_for<0, R2>([&](auto nn) {
                constexpr int n2 = nn;
                cd y = u[n2];
                if constexpr (n2 > 0) y = cd_mulc(y, jb_ldg(tw_kn + n2 * R1 + k1));
                dst[n2] = y;
            });
        }
        JB_SYNC();
        for (int it = jb_tid(); it < B * R2; it += jb_nthr()) {
            int b = it / R2, n2 = it - b * R2;
            cd U[H1];
            const cd* src = sm + (size_t)b * LS + n2;
            static_for<0, H1>([&](auto kk) { constexpr int k1 = kk; U[k1] = src[k1 * R2P]; });
            double x[R1];
            rdft_inv<R1>(U, x);
            static_for<0, R1>([&](auto nn) { constexpr int n1 = nn; st(b, R2 * n1 + n2, x[n1]); });
        }
    }
}

// ------------------------------------------------------------------------------ complex columns
// ld(p, b) -> cd ; mid(k1, k2, b, X) -> cd (spectrum-domain operation) ; st(p, b, cd).
// INVERSE=false stops after the forward transform and hands the spectrum to `mid` only.
template <class FP, int CB, bool INVERSE, class Ld, class Mid, class St>
JB_D void cols_cfft(cd* sm, const cd* __restrict__ tw_nk, Ld ld, Mid mid, St st) {
    constexpr int R1 = FP::R1, R2 = FP::R2;
    if constexpr (R2 == 1) {
        for (int b = jb_tid(); b < CB; b += jb_nthr()) {
            cd v[R1];
            static_for<0, R1>([&](auto nn) { constexpr int n = nn; v[n] = ld(n, b); });
            dft<R1, -1>(v);
            static_for<0, R1>([&](auto kk) { constexpr int k = kk; v[k] = mid(k, 0, b, v[k]); });
            if constexpr (INVERSE) {
                dft<R1, +1>(v);
                static_for<0, R1>([&](auto nn) { constexpr int n = nn; st(n, b, v[n]); });
            }
        }
    } else {
        for (int it = jb_tid(); it < R2 * CB; it += jb_nthr()) {
            int n2 = it / CB, b = it - n2 * CB;
            cd v[R1];
            static_for<0, R1>([&](auto nn) { constexpr int n1 = nn; v[n1] = ld(R2 * n1 + n2, b); });
            dft<R1, -1>(v);
            static_for<0, R1>([&](auto kk) {
                constexpr int k1 = kk;
                cd y = v[k1];
                if constexpr (k1 > 0) y = cd_mul(y, jb_ldg(tw_nk + k1 * R2 + n2));
                sm[(k1 * R2 + n2) * CB + b] = y;
            });
        }
        JB_SYNC();
        for (int it = jb_tid(); it < R1 * CB; it += jb_nthr()) {
            int k1 = it / CB, b = it - k1 * CB;
            cd u[R2];
            cd* col = sm + (size_t)(k1 * R2) * CB + b;
            static_for<0, R2>([&](auto nn) { constexpr int n2 = nn; u[n2] = col[n2 * CB]; });
            dft<R2, -1>(u);
            static_for<0, R2>([&](auto kk) { constexpr int k2 = kk; u[k2] = mid(k1, k2, b, u[k2]); });
            if constexpr (INVERSE) {
                dft<R2, +1>(u);
                static_for<0, R2>([&](auto nn) {
                    constexpr int n2 = nn;
                    cd y = u[n2];
                    if constexpr (n2 > 0) y = cd_mulc(y, jb_ldg(tw_nk + k1 * R2 + n2));
                    col[n2 * CB] = y;
                });
            }
        }
        if constexpr (INVERSE) {
            JB_SYNC();
            for (int it = jb_tid(); it < R2 * CB; it += jb_nthr()) {
                int n2 = it / CB, b = it - n2 * CB;
                cd v[R1];
                static_for<0, R1>([&](auto kk) { constexpr int k1 = kk; v[k1] = sm[(k1 * R2 + n2) * CB + b]; });
                dft<R1, +1>(v);
                static_for<0, R1>([&](auto nn) { constexpr int n1 = nn; st(R2 * n1 + n2, b, v[n1]); });
            }
        }
    }
}
