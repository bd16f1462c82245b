// fft_list.h -- transform lengths compiled into the library.
//   X(L, R1, R2, CB, PIPE, Q1, Q2, Q3)
//   rows use the two-pass factorisation L = R1*R2 (R1 even; R2 == 1 => single in-register pass),
//   CB = spectrum columns per column-pass work item,
//   PIPE = 2: the column pass runs as the persistent three-stage TMA pipeline with L = Q1*Q2*Q3 (cols3.cuh; needs
//             4*L*CB*16 B of shared memory <= 227 KB) and the c2r rows as the warp-private kernels of rows5.cuh;
//   PIPE = 0: generic kernels only (kernels.cuh): tiny test lengths and lengths beyond the pipeline's shared memory.
// The build (jubik_b200/build.py, __graft_entry__.build) compiles fft_inst.cu once per line.
// Factor order of the column pass: an ODD last radix Q3 makes stage 3 bank-conflict free with 4-column (64-byte) tile rows
// (two neighbouring items sit Q3*64 bytes apart) and a small one lets the module-sharing kernels keep Q3 spectrum values
// per thread in registers.  r2w, C5 chunk (L = 672): 8*7*12 -> 8*12*7 cols_fwd 8.45 -> 7.87 ms, cols_adj 12.15 -> 9.79 ms;
// 12*8*7: 7.66 / 10.38 ms; C3 (L = 336): 8*6*7 -> 6*8*7 cols_adj 1.157 -> 1.089 ms.  Radix-16 orders with Q3 = 3 / 5 for 384 / 640 were measured slower, 8*8*6 / 6*8*8 and 8*10*8 / 10*8*8 equal to the present orders (profiles/r2w_kbench_factors.txt, r2x_kbench_factors2.txt).
// The c2r row kernels (rows5.cuh, 80 registers) want the big radix LAST (it needs no twiddles there): JB_ROWS_F3_LIST names
// their factor order where it differs from the column pass's.
//
// Which P = 2*int(N/n) + 2*margin the reference's configurations produce (margin = ceil(0.062*N), demos/configs/*.yaml):
//   512^2:  4x4 -> 320, 8x8 -> 192 (the shipped Chandra / eROSITA demos);  256^2: 4x4 -> 160, 8x8 -> 96;
//   1024^2: 4x4 -> 640, 8x8 -> 384, 16x16 -> 256;  2048^2: 8x8 -> 766 (-> 768), 16x16 -> 510 (-> 512);  4096^2/16 -> 640.
// Any other P runs on the next larger length: padded (PSF within the margin) or with the period-P alias fold.
#pragma once
#define JB_FFT_LIST(X) \
    X(20, 20, 1, 8, 0, 1, 1, 1)     \
    X(24, 24, 1, 8, 0, 1, 1, 1)     \
    X(28, 28, 1, 8, 0, 1, 1, 1)     \
    X(96, 12, 8, 4, 2, 4, 4, 6)     \
    X(128, 16, 8, 4, 2, 8, 4, 4)    \
    X(160, 16, 10, 4, 2, 8, 5, 4)   \
    X(192, 16, 12, 4, 2, 8, 4, 6)   \
    X(256, 16, 16, 4, 2, 8, 4, 8)   \
    X(320, 20, 16, 4, 2, 8, 5, 8)   \
    X(336, 24, 14, 4, 2, 6, 8, 7)   \
    X(384, 24, 16, 4, 2, 8, 6, 8)   \
    X(448, 28, 16, 4, 2, 8, 7, 8)   \
    X(512, 32, 16, 4, 2, 8, 8, 8)   \
    X(640, 32, 20, 4, 2, 8, 8, 10)  \
    X(672, 28, 24, 4, 2, 8, 12, 7)  \
    X(768, 32, 24, 4, 2, 8, 8, 12)  \
    X(896, 32, 28, 4, 0, 1, 1, 1)   \
    X(1024, 32, 32, 4, 0, 1, 1, 1)
//   Y(L, Q1, Q2, Q3): factor order of the c2r row kernels for length L (default: the column pass's)
#define JB_ROWS_F3_LIST(Y) \
    Y(336, 8, 6, 7)            \
    Y(672, 8, 7, 12)
