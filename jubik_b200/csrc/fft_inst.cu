// fft_inst.cu -- one FFT length per compilation: nvcc -DJB_INST_L=384 -DJB_INST_R1=24 -DJB_INST_R2=16 -DJB_INST_CB=16
#include "ops.cuh"
#define JB_INST_EXPAND(L, R1, R2, CB) JB_DEFINE_FFT(L, R1, R2, CB)
JB_INST_EXPAND(JB_INST_L, JB_INST_R1, JB_INST_R2, JB_INST_CB)
