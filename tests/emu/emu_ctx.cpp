// storage for the emulated CUDA thread context (tests only; see build_emu.py)
#include "jb_rt.h"
thread_local jb_emu_ctx jb_ctx;
