// fft_list.h -- transform lengths compiled into the library: X(L, R1, R2, CB, PIPE)
//   L = R1*R2 (R1 even; R2 == 1 => single in-register pass), CB = spectrum columns per column-pass work item,
//   PIPE = 1: column pass runs as the persistent TMA pipeline (cols_pipe.cuh; needs 4*L*CB*16 B <= 227 KB and
//   CB*16 B == 128 B rows), 0: the plain one-CTA-per-strip kernel.
// The build (Makefile / __graft_entry__.build) compiles fft_inst.cu once per line.
#pragma once
// X(L, R1, R2, CB, PIPE, Q1, Q2, Q3): rows use L = R1*R2; PIPE = 0 plain column kernel, 1 two-pass TMA pipeline,
// 2 three-stage TMA pipeline with L = Q1*Q2*Q3 (cols3.cuh).
#define JB_FFT_LIST(X) \
    X(20, 20, 1, 8, 0, 1, 1, 1)    \
    X(24, 24, 1, 8, 0, 1, 1, 1)    \
    X(28, 28, 1, 8, 0, 1, 1, 1)    \
    X(160, 16, 10, 4, 2, 8, 5, 4)  \
    X(336, 24, 14, 4, 2, 8, 6, 7)  \
    X(384, 24, 16, 4, 2, 8, 6, 8)  \
    X(640, 32, 20, 4, 2, 8, 8, 10)  \
    X(672, 28, 24, 4, 2, 8, 7, 12)
