// fft_list.h -- transform lengths compiled into the library: X(L, R1, R2, CB)
//   L = R1*R2 (R1 even; R2 == 1 => single in-register pass), CB = spectrum columns per column-pass CTA.
// The build (Makefile / __graft_entry__.build) compiles fft_inst.cu once per line.
#pragma once
#define JB_FFT_LIST(X) \
    X(20, 20, 1, 8)    \
    X(24, 24, 1, 8)    \
    X(28, 28, 1, 8)    \
    X(160, 16, 10, 16) \
    X(336, 24, 14, 16) \
    X(384, 24, 16, 16) \
    X(640, 32, 20, 8)  \
    X(672, 28, 24, 8)
